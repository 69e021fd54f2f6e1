"""GPU: inference forward (`training=False`: BatchNorm on moving statistics) wrapped in the reference's frame
arithmetic (infer_video.py:138-159, infer.py:50-68, unit_test.py:67-86), through the C ABI, against the oracle."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import frames as F  # noqa: E402
from oracle import models as OM  # noqa: E402


def _frame(h, w, seed=0):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)


def _randomise_stats(p, seed=5):
    """Give BN moving statistics, biases and PReLU slopes non-default values so inference-mode BN is exercised."""
    g = torch.Generator().manual_seed(seed)
    for k in p:
        if k.endswith("moving_mean"):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.2
        elif k.endswith("moving_variance"):
            p[k] = torch.rand(p[k].shape, generator=g) * 1.5 + 0.25
        elif k.endswith(("bias", "beta", "alpha")):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.1
    return p


@pytest.mark.parametrize("src,dst", [((10, 12), (16, 16)), ((13, 9), (8, 16)), ((7, 7), (7, 7)), ((21, 40), (10, 33))])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_frame_kernels_bit_exact(src, dst, mode):
    from denoise_gan_b200 import _lib
    lib, ctx = _lib.load(), _lib.ctx()
    f = _frame(*src, seed=mode)
    fd = torch.from_numpy(f).cuda()
    x = torch.empty((1, dst[0], dst[1], 3), dtype=torch.float32, device="cuda")
    scale, offset = (2.0, -1.0) if mode == 0 else (1.0, 0.0)
    _lib.check(lib.dg_frame_to_float(ctx, fd.data_ptr(), src[0], src[1], 1, mode, scale, offset, _lib.tensor(x), _lib.stream_ptr()))
    rgb = f[..., ::-1]
    norm = {0: rgb.astype(np.float32) * np.float32(1.0 / 255.0), 1: (rgb / 255.0).astype(np.float32),
            2: rgb.astype(np.float32) / np.float32(255.0)}[mode]
    ref = F.resize_with_crop_or_pad(norm, *dst) * np.float32(scale) + np.float32(offset)
    np.testing.assert_array_equal(x[0].cpu().numpy(), ref)
    # post: random floats beyond [-1,1] so the clip and the saturation are exercised; crop-or-pad back to the source size
    y = (torch.randn((1, dst[0], dst[1], 3), generator=torch.Generator().manual_seed(3)) * 0.8).cuda()
    out = torch.empty((src[0], src[1], 3), dtype=torch.uint8, device="cuda")
    _lib.check(lib.dg_float_to_frame(ctx, _lib.tensor(y), 0.5, 0.5, 1, 0, out.data_ptr(), src[0], src[1], _lib.stream_ptr()))
    np.testing.assert_array_equal(out.cpu().numpy(), F.video_post(y[0].cpu().numpy(), *src))
    _lib.check(lib.dg_float_to_frame(ctx, _lib.tensor(y), 0.5, 0.5, 0, 1, out.data_ptr(), src[0], src[1], _lib.stream_ptr()))
    yc = F.resize_with_crop_or_pad(y[0].cpu().numpy(), *src) if src != dst else y[0].cpu().numpy()
    if src == dst:                                    # still_post has no crop-or-pad; compare where the geometry is the identity
        np.testing.assert_array_equal(out.cpu().numpy(), F.still_post(yc))


def _levels(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return int(d.max()), float((d > 0).mean())


def test_fsrgan_video_frame_fp32_matches_oracle():
    """infer_video.py:138-159 end to end at a small frame (pads to 256x256, x4 output cropped back to 4*fh x 4*fw)."""
    from denoise_gan_b200 import params as P
    from denoise_gan_b200.fsrgan import FastSRGAN
    from denoise_gan_b200.infer import FrameRunner
    g0 = _randomise_stats(P.init_fsrgan_generator(0))
    model = FastSRGAN(SimpleNamespace(crop_size=384, scale=4, lr=1e-3, fp16=0, vgg=0, seed=0))
    model.gen_params.load(g0)
    f = _frame(40, 56, seed=1)
    ours = FrameRunner(model, upscale=4).video_frame(f).numpy()
    x = torch.from_numpy(F.video_pre(f, *F.padded_size(40, 56)))[None].double()
    y = OM.fsrgan_generator({k: v.double() for k, v in g0.items()}, x, training=False)[0].float().numpy()
    ref = F.video_post(y, 160, 224)
    assert ours.shape == ref.shape == (160, 224, 3)
    mx, frac = _levels(ours, ref)
    assert mx <= 1 and frac < 2e-3, (mx, frac)       # truncation to uint8: an fp32-vs-fp64 ulp can flip a level


@pytest.mark.parametrize("kind", ["srgan", "autoencoder", "pix2pix"])
def test_generators_inference_mode_fp32(kind):
    """unit_test.py:67-86 path (256x256 crop, [0,1] input) for the other three generators."""
    from denoise_gan_b200 import params as P
    from denoise_gan_b200.infer import FrameRunner
    args = SimpleNamespace(crop_size=256, scale=4, lr=1e-3, fp16=0, vgg=0, seed=0, retrain=0)
    if kind == "srgan":
        from denoise_gan_b200.srgan import SRGAN
        g0 = _randomise_stats(P.init_srgan_generator(0, 4)); model = SRGAN(args); ofn = OM.srgan_generator; size = 64
    elif kind == "autoencoder":
        from denoise_gan_b200.autoencoder import Autoencoder
        g0 = _randomise_stats(P.init_autoencoder_generator(0)); model = Autoencoder(args); ofn = OM.autoencoder_generator; size = 256
    else:
        from denoise_gan_b200.pix2pix import Pix2Pix
        g0, _ = P.init_pix2pix(0); g0 = _randomise_stats(g0); model = Pix2Pix(args); ofn = OM.pix2pix_generator; size = 256
    model.gen_params.load(g0)
    f = _frame(size, size, seed=2)
    ours = FrameRunner(model, upscale=4 if kind == "srgan" else 1).unit_image(f).numpy()
    x = torch.from_numpy(F.unit_pre(f))[None].double()
    y = ofn({k: v.double() for k, v in g0.items()}, x, training=False)[0].float().numpy()
    ref = F.unit_post(y)
    assert ours.shape == ref.shape
    mx, frac = _levels(ours, ref)
    assert mx <= 1 and frac < 2e-3, (mx, frac)


def test_fsrgan_1080p_bf16_frame_and_tile_consistency():
    """C5: one 1080p frame through the bf16 tensor-core path (padded to 1280x2048, output 4320x7680), checked by a
    size-independent property: the interior of the output equals the forward of an interior sub-frame (finite
    receptive field, BatchNorm in inference mode is per-pixel), and against the fp64 oracle on that sub-frame."""
    from denoise_gan_b200 import params as P
    from denoise_gan_b200.fsrgan import FastSRGAN
    from denoise_gan_b200.infer import FrameRunner
    g0 = _randomise_stats(P.init_fsrgan_generator(0))
    model = FastSRGAN(SimpleNamespace(crop_size=384, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0))
    model.gen_params.load(g0)
    run = FrameRunner(model, upscale=4)
    f = _frame(1080, 1920, seed=4)
    full = run.video_frame(f).numpy()
    assert full.shape == (4320, 7680, 3)
    y0, x0, s, m = 400, 800, 128, 16                    # sub-frame and the margin (> receptive field radius of ~11 LR px)
    sub = F.video_pre(np.ascontiguousarray(f[y0:y0 + s, x0:x0 + s]), s, s)
    ysub = run.forward(torch.from_numpy(sub)[None].cuda()).float().cpu().numpy()[0]
    ours_sub = F.video_post(ysub, 4 * s, 4 * s)[4 * m:-4 * m, 4 * m:-4 * m]
    ours_full = full[4 * (y0 + m):4 * (y0 + s - m), 4 * (x0 + m):4 * (x0 + s - m)]
    mx, frac = _levels(ours_full, ours_sub)
    assert mx <= 1 and frac < 1e-2, ("tile consistency", mx, frac)
    yref = OM.fsrgan_generator({k: v.double() for k, v in g0.items()}, torch.from_numpy(sub)[None].double(), training=False)[0].float().numpy()
    ref = F.video_post(yref, 4 * s, 4 * s)[4 * m:-4 * m, 4 * m:-4 * m]
    d = np.abs(ours_full.astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 12 and d.mean() < 1.5, ("bf16 vs fp64", int(d.max()), float(d.mean()))   # bf16 storage: ~2^-8 relative per layer


def test_video_shard_pipeline_matches_single_frames():
    from denoise_gan_b200 import params as P
    from denoise_gan_b200.fsrgan import FastSRGAN
    from denoise_gan_b200.infer import FrameRunner
    model = FastSRGAN(SimpleNamespace(crop_size=384, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0))
    model.gen_params.load(_randomise_stats(P.init_fsrgan_generator(0)))
    run = FrameRunner(model, upscale=4)
    frames = [_frame(48, 64, seed=10 + i) for i in range(5)]
    single = [run.video_frame(fr).numpy() for fr in frames]
    seen = {}
    for rank in range(2):
        for i, out in run.video(frames, rank=rank, world=2):
            seen[i] = out.numpy()
    assert sorted(seen) == list(range(5))
    for i in range(5):
        np.testing.assert_array_equal(seen[i], single[i])


def test_video_frame_tight_padding_and_frame_sink_are_exact():
    """FrameRunner runs a Fast-SRGAN video frame on the frame + a receptive-field margin instead of the reference's padding to a
    multiple of 256 (infer_video.py:79-83,141,152): the cropped output must be the same bit for bit.  And the output convolution
    that writes the uint8 frame itself (dg_conv3x3_tapsum_frame) must agree with convolution + dg_float_to_frame to a level."""
    from denoise_gan_b200 import params as P
    from denoise_gan_b200.fsrgan import FastSRGAN
    from denoise_gan_b200.infer import FrameRunner
    model = FastSRGAN(SimpleNamespace(crop_size=384, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0))
    model.gen_params.load(_randomise_stats(P.init_fsrgan_generator(0)))
    run = FrameRunner(model, upscale=4)
    f = _frame(300, 420, seed=7)
    assert run.compute_size(300, 420) == (332, 452)
    tight = run.video_frame(f).numpy()
    run.tight_padding = False
    assert run.compute_size(300, 420) == (512, 512)
    full = run.video_frame(f).numpy()
    assert tight.shape == full.shape == (1200, 1680, 3)
    np.testing.assert_array_equal(tight, full)
    model.engine.tapsum_infer = False                   # implicit-GEMM output convolution + dg_float_to_frame
    old = run.video_frame(f).numpy()
    mx, frac = _levels(full, old)
    assert mx <= 1 and frac < 1e-3, (mx, frac)          # fp32 accumulation order of the last layer only


def test_srgan_video_frame_tight_padding_is_exact():
    """The same property for the SRGAN generator (srgan.py:129-185: 35 stride-1 3x3 convolutions deep, radius 35.5 -> margin 40)."""
    from denoise_gan_b200 import params as P
    from denoise_gan_b200.infer import FrameRunner
    from denoise_gan_b200.srgan import SRGAN
    model = SRGAN(SimpleNamespace(crop_size=384, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0))
    model.gen_params.load(_randomise_stats(P.init_srgan_generator(0, 4)))
    run = FrameRunner(model, upscale=4)
    assert model.generator.receptive_radius == 35.5
    f = _frame(140, 170, seed=8)                     # padded_size: 256 x 256, margins 58 / 43 >= 40
    assert run.compute_size(140, 170) == (220, 250)
    tight = run.video_frame(f).numpy()
    run.tight_padding = False
    assert run.compute_size(140, 170) == (256, 256)
    full = run.video_frame(f).numpy()
    np.testing.assert_array_equal(tight, full)


def test_srgan_bf16_inference_with_fused_skip_epilogue():
    """SRGAN generator in bf16 at inference: the 17 `conv -> BN -> Add` pairs (srgan.py:166-169, :174-176) run as convolutions with the
    folded BatchNorm and the skip-add in their epilogue (dg_umma_conv2d_fwd_res_prelu).  Against the fp64 oracle, and against the same
    model with the skip-adds as separate passes."""
    from denoise_gan_b200 import params as P
    from denoise_gan_b200.infer import FrameRunner
    from denoise_gan_b200.srgan import SRGAN
    g0 = _randomise_stats(P.init_srgan_generator(0, 4))
    model = SRGAN(SimpleNamespace(crop_size=256, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0, retrain=0))
    model.gen_params.load(g0)
    run = FrameRunner(model, upscale=4)
    f = _frame(64, 64, seed=2)
    fused = run.unit_image(f).numpy()
    y = OM.srgan_generator({k: v.double() for k, v in g0.items()}, torch.from_numpy(F.unit_pre(f))[None].double(), training=False)[0].float().numpy()
    ref = F.unit_post(y)
    d = np.abs(fused.astype(np.int32) - ref.astype(np.int32))
    print("bf16 fused vs fp64 oracle: max", int(d.max()), "mean", float(d.mean()))
    assert d.mean() < 4.0 and d.max() <= 40, (int(d.max()), float(d.mean()))      # 35 bf16 layers deep
    model.engine.fuse_res_epilogue = False
    plain = run.unit_image(f).numpy()
    d2 = np.abs(fused.astype(np.int32) - plain.astype(np.int32))
    d3 = np.abs(plain.astype(np.int32) - ref.astype(np.int32))
    print("fused vs separate passes: max", int(d2.max()), "mean", float(d2.mean()), "| separate vs oracle mean", float(d3.mean()))
    assert d2.mean() < 2.0, (int(d2.max()), float(d2.mean()))
    assert d.mean() <= d3.mean() + 0.25            # one rounding per layer instead of two: not worse than the separate passes


def test_video_graph_replay_matches_eager_frames():
    """FrameRunner.video() replays a captured CUDA graph per staging slot from the third frame of that slot on: eight frames through
    one rank (eager, eager, capture, replay per slot) must equal the eager single-frame results, in order."""
    from denoise_gan_b200 import params as P
    from denoise_gan_b200.fsrgan import FastSRGAN
    from denoise_gan_b200.infer import FrameRunner
    model = FastSRGAN(SimpleNamespace(crop_size=384, scale=4, lr=1e-3, fp16=1, vgg=0, seed=0))
    model.gen_params.load(_randomise_stats(P.init_fsrgan_generator(0)))
    run = FrameRunner(model, upscale=4)
    frames = [_frame(48, 64, seed=30 + i) for i in range(8)]
    single = [run.video_frame(fr).numpy() for fr in frames]
    got = {i: out.numpy() for i, out in run.video(frames, rank=0, world=1)}
    assert sorted(got) == list(range(8)) and len(run._graphs) == 2
    for i in range(8):
        np.testing.assert_array_equal(got[i], single[i])
    # new weights invalidate the captured frames (they read BatchNorm-folded kernels that are rebuilt eagerly)
    model.gen_params.load(_randomise_stats(P.init_fsrgan_generator(3), seed=9))
    single2 = [run.video_frame(fr).numpy() for fr in frames]
    got2 = {i: out.numpy() for i, out in run.video(frames, rank=0, world=1)}
    assert any((single2[i] != single[i]).any() for i in range(8))
    for i in range(8):
        np.testing.assert_array_equal(got2[i], single2[i])
