"""CPU: frame arithmetic of the inference entry (oracle/frames.py restating infer_video.py:79-83,138-159) and
the round-robin frame shard."""
import numpy as np

from oracle import frames as F


def test_padded_size_matches_reference_formula():
    assert F.padded_size(1080, 1920) == (1280, 2048)
    assert F.padded_size(256, 512) == (512, 768)            # exact multiples still gain a block (infer_video.py:80)
    assert F.padded_size(1, 255) == (256, 256)
    from denoise_gan_b200.infer import padded_size
    for fh, fw in ((1080, 1920), (720, 1280), (256, 256), (17, 999)):
        assert padded_size(fh, fw) == F.padded_size(fh, fw)


def test_crop_or_pad_centre_offsets():
    a = np.arange(5 * 4, dtype=np.float32).reshape(5, 4, 1)
    out = F.resize_with_crop_or_pad(a, 8, 7)                # pad 3 rows: 1 before, 2 after; 3 cols: 1 before, 2 after
    assert out.shape == (8, 7, 1) and out[0].sum() == 0 and out[6:].sum() == 0 and out[:, 0].sum() == 0 and out[:, 5:].sum() == 0
    np.testing.assert_array_equal(out[1:6, 1:5], a)
    out = F.resize_with_crop_or_pad(a, 2, 1)                # crop 3 rows: offset 1; 3 cols: offset 1
    np.testing.assert_array_equal(out, a[1:3, 1:2])
    out = F.resize_with_crop_or_pad(a, 7, 2)                # mixed: pad rows, crop cols
    np.testing.assert_array_equal(out[1:6], a[:, 1:3])


def test_video_pre_post_ranges():
    rng = np.random.default_rng(0)
    f = rng.integers(0, 256, size=(10, 12, 3), dtype=np.uint8)
    x = F.video_pre(f, 16, 16)
    assert x.dtype == np.float32 and x.min() >= -1 and x.max() <= 1 and x[0, 0, 0] == -1.0      # padding maps to -1
    near = lambda a, b: np.abs(a.astype(int) - b.astype(int)).max() <= 1      # astype(uint8) truncates: 254.99998 -> 254
    assert near(F.video_post(x, 10, 12), f[..., ::-1])                          # identity model round-trips to within a level
    assert F.video_post(np.full((4, 4, 3), 3.0, np.float32), 4, 4).max() == 255                   # clip
    assert near(F.unit_post(F.unit_pre(f) * 2 - 1), f[..., ::-1])
    assert near(F.still_post(F.still_pre(f) * 2 - 1), f)


def test_frames_for_rank_partition():
    from denoise_gan_b200.infer import frames_for_rank
    for n, world in ((10, 4), (7, 8), (100, 3), (0, 2)):
        got = sorted(i for r in range(world) for i in frames_for_rank(n, r, world))
        assert got == list(range(n))
        sizes = [len(frames_for_rank(n, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
