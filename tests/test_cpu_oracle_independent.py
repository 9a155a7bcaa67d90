"""The oracle's arithmetic against INDEPENDENT numpy / scipy implementations of the same mathematics.

The oracle is DiffDRR's glue restated from the public package (parity unpinned).  What *can* be
pinned here without DiffDRR is that the restated glue computes what its names say, by a second implementation that
shares no code with it: trilinear line integrals via ``scipy.ndimage.map_coordinates`` in float64, Siddon's exact
radiological path by brute-force sub-sampling, NCC as Pearson correlation (``numpy.corrcoef``), patch NCC over
``sliding_window_view``, Sobel via ``scipy.ndimage.correlate``, the geodesic via scipy rotations.  The remaining
freedom is the list of named conventions in ``oracle/knobs.py``.
"""

import numpy as np
import torch
from numpy.lib.stride_tricks import sliding_window_view
from scipy import ndimage
from scipy.spatial.transform import Rotation

import oracle
from oracle import knobs


def _rays(n_rays, shape, seed):
    """One source outside a volume of the given shape and segments from it through random interior points
    (voxel-index coordinates), as a C-arm pose produces them."""
    g = np.random.default_rng(seed)
    centre = (np.array(shape) - 1) / 2.0
    direction = g.normal(size=3)
    source = centre - direction / np.linalg.norm(direction) * 2.5 * max(shape)
    through = centre + g.uniform(-0.35, 0.35, size=(n_rays, 3)) * np.array(shape)
    target = source + 2.0 * (through - source)
    return source[None].repeat(n_rays, 0), target


def _slab(source, target, lo, hi):
    d = target - source
    with np.errstate(divide="ignore"):
        a0, a1 = (lo - source) / d, (hi - source) / d
    amin = np.clip(np.minimum(a0, a1).max(1), 0.0, 1.0)
    amax = np.clip(np.maximum(a0, a1).min(1), 0.0, 1.0)
    return amin, amax


def test_trilinear_render_is_a_riemann_sum_of_trilinear_interpolation():
    shape = (20, 24, 18)
    vol = np.random.default_rng(0).random(shape)
    source, target = _rays(64, shape, 1)
    raylen = np.linalg.norm(target - source, axis=1) * 0.7  # any world scale: mm per voxel unit
    n = 200
    amin, amax = _slab(source, target, 0.0, np.array(shape) - 1.0)
    alphas = amin[:, None] + (amax - amin)[:, None] * np.linspace(0.0, 1.0, n)[None]
    pts = source[:, None] + alphas[..., None] * (target - source)[:, None]
    samples = ndimage.map_coordinates(vol, pts.reshape(-1, 3).T, order=1, mode="grid-constant", cval=0.0)
    ref = samples.reshape(64, n).sum(1) * raylen * (amax - amin) / (n - 1)

    out = oracle.trilinear_render(torch.as_tensor(vol, dtype=torch.float32),
                                  torch.as_tensor(source[:1], dtype=torch.float32)[None],
                                  torch.as_tensor(target, dtype=torch.float32)[None],
                                  torch.as_tensor(raylen, dtype=torch.float32)[None, None], n_points=n,
                                  step="span/(n-1)")
    out = out[0, 0].double().numpy()
    assert (ref > 0).mean() > 0.9
    assert np.abs(out - ref).max() < 2e-4 * ref.max()


def test_siddon_render_is_the_exact_line_integral_of_the_voxel_grid():
    """Nearest-voxel values integrated along the ray: brute force with 40 000 sub-samples per ray."""
    shape = (12, 10, 14)
    vol = np.random.default_rng(2).random(shape)
    source, target = _rays(24, shape, 3)
    raylen = np.linalg.norm(target - source, axis=1)
    shift = knobs.SIDDON_VOXEL_SHIFT_DEFAULT
    amin, amax = _slab(source, target, -shift, np.array(shape) - shift)
    m = 40_000
    u = (np.arange(m) + 0.5) / m
    alphas = amin[:, None] + (amax - amin)[:, None] * u[None]
    pts = source[:, None] + alphas[..., None] * (target - source)[:, None]
    idx = np.floor(pts + shift).astype(int)  # voxel i covers [i - shift, i + 1 - shift)
    inside = ((idx >= 0) & (idx < np.array(shape))).all(-1)
    idx = np.clip(idx, 0, np.array(shape) - 1)
    vals = vol[idx[..., 0], idx[..., 1], idx[..., 2]] * inside
    ref = vals.mean(1) * (amax - amin) * raylen

    out = oracle.siddon_render(torch.as_tensor(vol, dtype=torch.float32),
                               torch.as_tensor(source[:1], dtype=torch.float32)[None],
                               torch.as_tensor(target, dtype=torch.float32)[None],
                               torch.as_tensor(raylen, dtype=torch.float32)[None, None])
    out = out[0, 0].double().numpy()
    assert (ref > 0).mean() > 0.9
    assert np.abs(out - ref).max() < 2e-3 * ref.max()  # the brute force resolves voxel borders to 1/40 000 of the ray


def test_ncc_is_pearson_correlation():
    g = np.random.default_rng(4)
    a = g.random((3, 1, 40, 36)) * 2.0
    b = 0.6 * a + 0.4 * g.random((3, 1, 40, 36))
    out = oracle.ncc(torch.as_tensor(a, dtype=torch.float32), torch.as_tensor(b, dtype=torch.float32)).numpy()
    ref = np.array([np.corrcoef(a[i].ravel(), b[i].ravel())[0, 1] for i in range(3)])
    assert np.abs(out - ref).max() < 2e-4  # eps = 1e-5 on variances of O(0.1)

    p = 9
    out_p = oracle.ncc(torch.as_tensor(a, dtype=torch.float32), torch.as_tensor(b, dtype=torch.float32), p).numpy()
    ref_p = []
    for i in range(3):
        wa = sliding_window_view(a[i, 0], (p, p)).reshape(-1, p * p)
        wb = sliding_window_view(b[i, 0], (p, p)).reshape(-1, p * p)
        wa = wa - wa.mean(1, keepdims=True)
        wb = wb - wb.mean(1, keepdims=True)
        r = (wa * wb).mean(1) / np.sqrt(((wa * wa).mean(1) + knobs.NCC_EPS) * ((wb * wb).mean(1) + knobs.NCC_EPS))
        ref_p.append(r.mean())
    assert np.abs(out_p - np.array(ref_p)).max() < 1e-5


def test_sobel_is_the_3x3_sobel_cross_correlation_with_zero_padding():
    img = np.random.default_rng(5).random((2, 1, 17, 23))
    out = oracle.sobel(torch.as_tensor(img, dtype=torch.float32)).double().numpy()
    kx = np.array([[1.0, 0.0, -1.0], [2.0, 0.0, -2.0], [1.0, 0.0, -1.0]])
    for i in range(2):
        gx = ndimage.correlate(img[i, 0], kx, mode="constant", cval=0.0)
        gy = ndimage.correlate(img[i, 0], kx.T, mode="constant", cval=0.0)
        assert np.abs(out[i, 0] - gx).max() < 1e-5 and np.abs(out[i, 1] - gy).max() < 1e-5
    # and, up to sign, scipy's own Sobel operator (smoothing [1,2,1] across, derivative [-1,0,1] along the axis)
    assert np.abs(np.abs(out[0, 0]) - np.abs(ndimage.sobel(img[0, 0], axis=1, mode="constant"))).max() < 1e-5
    assert np.abs(np.abs(out[0, 1]) - np.abs(ndimage.sobel(img[0, 0], axis=0, mode="constant"))).max() < 1e-5


def test_double_geodesic_is_the_rotation_angle_times_half_the_source_detector_distance():
    ra, rb = Rotation.random(8, random_state=6), Rotation.random(8, random_state=7)
    g = np.random.default_rng(8)
    ta, tb = g.normal(size=(8, 3)) * 50, g.normal(size=(8, 3)) * 50
    A, B = torch.eye(4).repeat(8, 1, 1), torch.eye(4).repeat(8, 1, 1)
    A[:, :3, :3], A[:, :3, 3] = torch.as_tensor(ra.as_matrix(), dtype=torch.float32), torch.as_tensor(ta, dtype=torch.float32)
    B[:, :3, :3], B[:, :3, 3] = torch.as_tensor(rb.as_matrix(), dtype=torch.float32), torch.as_tensor(tb, dtype=torch.float32)
    sdd = 1020.0
    ang, tra, dbl = (x.double().numpy() for x in oracle.double_geodesic(A, B, sdd))
    ref_ang = (ra.inv() * rb).magnitude() * sdd / 2.0
    ref_tra = np.linalg.norm(ta - tb, axis=1)
    assert np.abs(ang - ref_ang).max() < 0.3  # mm on O(1000): acos of a trace in fp32
    assert np.abs(tra - ref_tra).max() < 1e-3
    assert np.abs(dbl - np.sqrt(ref_ang**2 + ref_tra**2)).max() < 0.3


def test_hu_to_density_piecewise_map():
    hu = np.random.default_rng(9).uniform(-1000, 1500, size=(16, 16, 16))
    for m in (1.0, 4.5):
        out = oracle.hu_to_density(torch.as_tensor(hu, dtype=torch.float32), m).double().numpy()
        soft = (hu > knobs.HU_AIR) & (hu <= knobs.HU_BONE)
        ref = np.where(hu > knobs.HU_BONE, hu * m, np.where(soft, hu, hu[soft].min()))
        ref = ref - ref.min()
        ref = ref / ref.max()
        assert np.abs(out - ref).max() < 1e-5
