"""CPU: host-side logic added in round 2 that needs no device -- channel-segment bookkeeping of physically padded activations, the
packed loss vector of the captured step, the image-summary table, the libjpeg quality scaling used by the device kernel."""
import torch


def test_channel_segments_merge():
    from denoise_gan_b200.engine import Engine
    n = Engine._norm_segs
    assert n([(64, 64), (3, 16)]) == ((67, 80),)                       # a dense part joins the padded part that follows it
    assert n([(100, 112), (76, 80)]) == ((100, 112), (76, 80))         # two padded parts stay two segments (autoencoder.py:135)
    assert n([(32, 32), (32, 32)]) is None                             # nothing padded: a dense tensor
    assert n([(152, 160)]) == ((152, 160),)
    assert n([(84, 96), (32, 32)]) == ((84, 96), (32, 32))


def test_packed_loss_vector_reuses_the_kernel_output():
    from denoise_gan_b200.graph import GraphedStep
    t = torch.arange(7, dtype=torch.float32)
    views = [t[i] for i in range(7)]
    packed = GraphedStep._pack(views)
    assert packed.data_ptr() == t.data_ptr() and packed.tolist() == t.tolist()      # no copy: the 7-vector of dg_gan_loss_terms itself
    scattered = [torch.tensor(float(i)) for i in range(3)]
    assert GraphedStep._pack(scattered).tolist() == [0.0, 1.0, 2.0]                 # unrelated scalars are stacked
    assert GraphedStep._pack([]) is None


def test_image_summary_table_matches_the_reference_tags():
    from denoise_gan_b200.summaries import SUMMARIES
    tags = [t for t, *_ in SUMMARIES]
    assert len(tags) == 16 and len(set(tags)) == 16                    # train_srgan.py:156-172
    assert tags[:3] == ["Images/Input", "Images/Target", "Images/Generated"]
    assert sum(1 for t in tags if t.startswith("Image Gradients/")) == 9 and sum(1 for t in tags if t.startswith("Error/")) == 4


def test_c_abi_declares_the_round2_entry_points():
    import os
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "dg_b200.h")).read()
    for sym in ("dg_umma_conv2d_wgrad_batch", "dg_umma_conv2d_fwd_narrow", "dg_umma_conv2d_fwd_d2s_prelu", "dg_umma_conv2d_dgrad_relu_mask",
                "dg_umma_pack_weights_seg", "dg_unpad_weight_grad_seg", "dg_maxpool2x2_bwd_relu", "dg_pair_synthesis", "dg_image_summary",
                "dg_gan_loss_terms", "dg_comm_allreduce", "dg_conv3x3_tapsum_fwd", "dg_conv3x3_tapsum_frame", "dg_conv3x3_tapsum_supported", "dg_umma_conv2d_fwd_res_prelu"):
        assert sym + "(" in hdr, sym
    from denoise_gan_b200 import _lib
    lib = _lib.load()                                                  # binds every declared symbol or raises
    assert all(hasattr(lib, s) for s in ("dg_umma_conv2d_wgrad_batch", "dg_pair_synthesis", "dg_image_summary"))


def test_tight_size_keeps_the_reference_geometry():
    """FrameRunner.compute_size: frame + receptive-field margin when the reference's padding (infer_video.py:79-83) is even and at
    least that wide on both axes, the reference's size otherwise."""
    from denoise_gan_b200.infer import padded_size, tight_size
    assert padded_size(1080, 1920) == (1280, 2048)
    assert tight_size(1080, 1920, None) == (1280, 2048)
    assert tight_size(1080, 1920, 9.75) == (1112, 1952)           # margin 16 = the radius rounded up to a multiple of 8
    assert tight_size(1080, 1920, 40.0) == (1160, 2000)
    assert tight_size(1080, 1920, 70.0) == (1280, 2048)           # 72 > the 64 columns of padding
    assert tight_size(1081, 1920, 9.75) == padded_size(1081, 1920)   # odd padding: the reference's crop is off-centre by half a pixel
    assert tight_size(256, 256, 9.75) == (288, 288)               # a multiple of 256 is padded by a full block (infer_video.py:81-82)
    assert tight_size(250, 250, 9.75) == padded_size(250, 250)    # 3 pixels of padding < margin


def test_concat_segments_with_a_padded_tail():
    """Engine.upsample_concat pads a concat of 16 (mod 32) physical channels with 16 zero channels, counted as padding of the last
    segment: the autoencoder's last concat (64 + 3 of 16 channels, autoencoder.py:181) becomes one (67 logical, 96 physical) segment,
    the others keep two segments -- the two forms conv2d's packed kernels understand."""
    from denoise_gan_b200.engine import Engine
    norm = Engine._norm_segs
    assert norm([(64, 64), (3, 16)]) == ((67, 80),)
    assert norm([(64, 64), (3, 32)]) == ((67, 96),)
    assert norm([(44, 48), (32, 48)]) == ((44, 48), (32, 48))
    assert norm([(64, 64), (32, 32)]) is None
