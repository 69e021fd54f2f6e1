"""The reference's `train(model, dataset, args, writer)` (train_srgan.py:120-176, same shape in train_fsrgan.py:122,
train_autoencoder.py:114, train_pix2pix.py:73): iterate the dataset, run `train_step`, log the loss scalars every
`args.save_iter` iterations.  Here the loop is built from the pieces the benchmark times end to end:

* `DevicePrefetcher` moves batch k+1.. over PCIe on a copy stream while step k runs (the reference's `dataset.prefetch`);
* `GraphedStep` replays the captured step (the reference's `@tf.function`);
* `LossLogger` copies every step's scalars into a small ring of pinned host buffers and hands them to the writer a few
  steps late, so logging never synchronises the device inside the loop (TensorFlow's eager scalars block on `.numpy()`).

The reference's image / Sobel / total-variation image summaries (train_srgan.py:150-172) are produced on request
(`image_summaries=True`, summaries.py): at a logging iteration the generator runs once more in inference mode and sixteen small
uint8 images come back -- a diagnostic outside the hot path, the one place where the loop waits for the
device.  `writer` is anything with `add_scalar(tag, value, step)` (and optionally `add_image`; the torch / TensorBoard
SummaryWriter API), a callable `(tag, value, step)`, or None."""
from __future__ import annotations

import torch

# return order of train_srgan.train_step (train_srgan.py:118) -> the tags of train_srgan.py:142-148
SRGAN_TAGS = ("Generator Losses/gen_loss", "Generator Losses/adv_loss", "Generator Losses/mae_loss", "Generator Losses/mse_loss",
              "Generator Losses/content_loss", "Discriminator Losses/disc_loss", "Generator Losses/total_variation")


class LossLogger:
    """Ring of `lag + 1` host buffers: `push(step, scalars)` enqueues the device->host copy of one step's scalar vector
    and delivers the vector pushed `lag` steps earlier (by then its copy has completed, so the wait is free)."""

    def __init__(self, tags, log_iter: int, writer=None, lag: int = 4):
        self.tags, self.log_iter, self.writer, self.lag = tuple(tags), max(1, int(log_iter)), writer, max(0, int(lag))
        self.bufs = [None] * (self.lag + 1)
        self.events = [None] * (self.lag + 1)
        self.inflight: list[tuple[int, int]] = []      # (slot, step), oldest first
        self.k = 0
        self.last = None                               # (step, [values]) most recently delivered
        self.records: list[tuple[str, float, int]] = []   # everything written, for callers without a writer

    def _emit(self, step, values):
        self.last = (step, values)
        if step % self.log_iter != 0:
            return
        for tag, v in zip(self.tags, values):
            self.records.append((tag, v, step))
            if self.writer is None:
                continue
            if hasattr(self.writer, "add_scalar"):
                self.writer.add_scalar(tag, v, step)
            else:
                self.writer(tag, v, step)

    def _deliver(self, slot, step):
        if self.events[slot] is not None:
            self.events[slot].synchronize()
        self._emit(step, self.bufs[slot].tolist())

    def push(self, step: int, scalars: torch.Tensor):
        s = self.k % (self.lag + 1)
        self.k += 1
        if self.bufs[s] is None or self.bufs[s].shape != scalars.shape:
            self.bufs[s] = torch.empty(scalars.shape, dtype=torch.float32, pin_memory=scalars.is_cuda)
        self.bufs[s].copy_(scalars, non_blocking=True)
        if scalars.is_cuda:
            if self.events[s] is None:
                self.events[s] = torch.cuda.Event()
            self.events[s].record()
        self.inflight.append((s, step))
        if len(self.inflight) > self.lag:
            self._deliver(*self.inflight.pop(0))

    def flush(self):
        while self.inflight:
            self._deliver(*self.inflight.pop(0))
        return self.last


def train(model, dataset, args, writer=None, *, train_step, tags=SRGAN_TAGS, use_graph: bool = True, lag: int = 4,
          image_summaries: bool = False, image_sink: "list | None" = None):
    """Runs `train_step(model, input, target)` over `dataset` (host or device float32 NHWC batches in [-1, 1], static batch
    shape as the reference's `drop_remainder=True`), logging `tags` every `args.save_iter` iterations.  Returns the last
    step's scalars as floats, in `train_step`'s return order (train_srgan.py:176)."""
    from .graph import DevicePrefetcher, GraphedStep

    logger = LossLogger(tags, getattr(args, "save_iter", 1), writer, lag)
    feed = DevicePrefetcher(dataset, model.device)
    step, eager_steps = None, 0
    for x, y in feed:
        if use_graph and eager_steps >= 2:
            if step is None:
                # the first two batches ran eagerly (they allocate every pooled buffer and pack the weights); capture the step
                # on this batch -- capturing executes nothing on the device but runs the Python side of train_step once
                it0 = model.iterations
                step = GraphedStep(model, train_step, x, y, warmup=0)
                model.iterations = it0
            step(x, y)
            packed = step.packed
        else:
            out = train_step(model, x, y)
            eager_steps += 1
            packed = torch.stack([v.detach().float().reshape(()) for v in out])
        logger.push(model.iterations, packed)
        if image_summaries and model.iterations % logger.log_iter == 0:
            from .summaries import training_image_summaries, write_image_summaries
            images = training_image_summaries(model, x, y)       # eager inference forward between two replays of the step graph
            write_image_summaries(writer, images, model.iterations)
            if image_sink is not None:
                image_sink.append((model.iterations, images))
    last = logger.flush()
    return None if last is None else tuple(last[1])
