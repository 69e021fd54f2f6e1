"""LossLogger (train_loop.py): scalars are delivered `lag` steps late, in order, only on logging iterations, and flush()
drains the ring — the host-side logic of the reference's `train()` scalar summaries (train_srgan.py:140-148)."""
import torch

from denoise_gan_b200.train_loop import SRGAN_TAGS, LossLogger


def test_logger_lag_order_and_log_iter():
    seen = []
    lg = LossLogger(SRGAN_TAGS, log_iter=3, writer=lambda tag, v, step: seen.append((tag, v, step)), lag=2)
    for it in range(1, 11):
        lg.push(it, torch.arange(7, dtype=torch.float32) + 10 * it)
        # nothing newer than it - lag has been delivered
        assert lg.last is None or lg.last[0] == it - 2
    assert [s for _, _, s in seen] == [3] * 7 + [6] * 7          # step 9 is still in the ring
    last = lg.flush()
    assert last[0] == 10 and last[1] == [100.0 + j for j in range(7)]
    steps = sorted({s for _, _, s in seen})
    assert steps == [3, 6, 9]
    by = {(t, s): v for t, v, s in seen}
    assert by[("Generator Losses/gen_loss", 6)] == 60.0 and by[("Discriminator Losses/disc_loss", 9)] == 95.0
    assert by[("Generator Losses/total_variation", 3)] == 36.0
    assert lg.records == seen


class _Writer:
    def __init__(self):
        self.rows = []

    def add_scalar(self, tag, value, step):
        self.rows.append((tag, value, step))


def test_logger_summary_writer_api_and_zero_lag():
    w = _Writer()
    lg = LossLogger(("a", "b"), log_iter=1, writer=w, lag=0)
    lg.push(1, torch.tensor([1.0, 2.0]))
    assert w.rows == [("a", 1.0, 1), ("b", 2.0, 1)]
    lg.push(2, torch.tensor([3.0, 4.0]))
    assert lg.flush() == (2, [3.0, 4.0]) and len(w.rows) == 4
