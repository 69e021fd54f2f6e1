"""The shared-memory address swizzle the staged conv epilogue and the wgrad bias sums apply by hand (conv_umma.cu `swz`,
conv_umma_wgrad.cu): `phys = off ^ (((off >> 7) & mask) << 4)` with mask 7 / 3 / 1 must be exactly the TMA 128 / 64 / 32-byte
swizzle of a 1024-byte-aligned box (16-byte chunk index XOR row bits), a bijection on every 1024-byte block, and conflict-free
for the two access patterns the kernels use."""
import itertools

import pytest


def swz(off, mask):
    return off ^ (((off >> 7) & mask) << 4)


@pytest.mark.parametrize("row_bytes,mask", [(128, 7), (64, 3), (32, 1)])
def test_matches_tma_swizzle_and_is_a_bijection(row_bytes, mask):
    span = {7: 128, 3: 64, 1: 32}[mask]
    assert row_bytes == span
    seen = set()
    for r, chunk in itertools.product(range(256), range(row_bytes // 16)):
        off = r * row_bytes + chunk * 16
        phys = swz(off, mask)
        # TMA swizzle<span>: within each 8-row x 128-byte atom the 16-byte chunk index is XORed with address bits [7, 7+log2(span/16))
        line, in_line = divmod(off, 128)
        want = line * 128 + ((in_line // 16) ^ (line & mask)) * 16
        assert phys == want
        assert phys // 1024 == off // 1024          # never leaves its 1024-byte block
        seen.add(phys)
    assert len(seen) == 256 * (row_bytes // 16)     # bijection


@pytest.mark.parametrize("row_bytes,mask", [(128, 7), (64, 3), (32, 1)])
def test_epilogue_store_pattern_is_bank_conflict_free(row_bytes, mask):
    """Thread = accumulator row, 16-byte stores of chunk j: every quarter-warp (8 consecutive rows) must hit 8 distinct
    16-byte bank groups (shared memory has 32 four-byte banks = 8 groups of 16 bytes)."""
    for j in range(row_bytes // 16):
        for r0 in range(0, 128, 8):
            groups = {(swz(r * row_bytes + j * 16, mask) % 128) // 16 for r in range(r0, r0 + 8)}
            assert len(groups) == 8


@pytest.mark.parametrize("nb", [16, 32, 64])
def test_statistics_read_pattern_is_bank_conflict_free(nb):
    """BatchNorm statistics: lane = (row offset, 4-byte word) over 32 / WPR consecutive rows; the 32 lanes of one read must
    fall into 32 distinct banks."""
    row_bytes, mask = nb * 2, {64: 7, 32: 3, 16: 1}[nb]
    wpr = nb // 2
    rpr = 32 // wpr
    for r0 in range(0, 128, rpr):
        banks = {(swz((r0 + lane // wpr) * row_bytes + 4 * (lane % wpr), mask) // 4) % 32 for lane in range(32)}
        assert len(banks) == 32
