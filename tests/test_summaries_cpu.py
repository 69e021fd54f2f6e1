"""CPU: closed-form checks of the image-summary oracle (oracle/summaries.py, train_srgan.py:27-59)."""
import numpy as np

from oracle import summaries as OS


def test_renorm_and_image_kind():
    x = np.array([[[-1.5, -1.0, 0.0], [0.5, 1.0, 3.0]]], dtype=np.float32).repeat(2, axis=0)      # [2, 2, 3]
    assert OS.summary_u8(OS.IMAGE, x)[0].tolist() == [[0, 0, 127], [191, 255, 255]]


def test_sobel_of_a_ramp_is_constant():
    yy, xx = np.mgrid[0:9, 0:11].astype(np.float32)
    ramp = ((xx / 10) * 2 - 1)[..., None]                         # renorm(ramp) = xx / 10: d/dx = 0.1 per pixel
    g = OS.sobel_variation(ramp)[1:-1, 1:-1, 0]                   # interior: (1 + 2 + 1) * 2 * 0.1 / 4 = 0.2
    assert np.allclose(g, 0.2, atol=1e-6)
    edge = OS.sobel_variation(ramp)[4, 0, 0]                      # REFLECT padding: the column left of 0 mirrors column 1 -> zero gradient
    assert abs(edge) < 1e-6


def test_total_variation_shapes_and_values():
    x = np.zeros((4, 5, 1), dtype=np.float32); x[1, 2, 0] = 1.0
    dx, dy = OS.high_pass_x_y(x)
    assert dx.shape == (3, 4, 1) and dy.shape == (3, 4, 1)
    tv = OS.values(OS.TV, x)
    assert tv[1, 2, 0] == 2.0 and tv[1, 1, 0] == 1.0 and tv[0, 2, 0] == 1.0
    u = OS.summary_u8(OS.TV, x)
    assert u.max() == 255 and u.min() == 0 and u[1, 1, 0] == 127


def test_sobel_magnitude_matches_opencv():
    """tf.image.sobel_edges = the 3x3 Sobel pair on a REFLECT-padded image; OpenCV's Sobel with BORDER_REFLECT_101 is the same operator."""
    import pytest
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    img = (rng.random((12, 15, 3)).astype(np.float32) * 2 - 1)
    r = OS.renorm(img)
    gx = cv2.Sobel(r, cv2.CV_32F, 1, 0, ksize=3, borderType=cv2.BORDER_REFLECT_101)
    gy = cv2.Sobel(r, cv2.CV_32F, 0, 1, ksize=3, borderType=cv2.BORDER_REFLECT_101)
    ref = np.sqrt((gx / 4) ** 2 + (gy / 4) ** 2)
    assert np.abs(ref - OS.sobel_variation(img)).max() < 1e-6
