"""The host mirror of jaxpm/kernels.py (jaxpm_b200/kernels.py: fftk, gradient_kernel orders 0 and 1, invlaplace_kernel
incl. the fd form, longrange_kernel, cic_compensation) against the fixture produced by the reference's own source
(tests/golden/kernels.npz).  The mirror lays the z axis out as the R2C half axis (nz // 2 + 1 entries): compared with
the matching slice of the reference's full axis, Nyquist handling of the order-0 kernel included."""
import os

import numpy as np
import torch

from jaxpm_b200 import kernels as K

HERE = os.path.dirname(os.path.abspath(__file__))


def test_host_kernels_against_reference_fixture():
    g = np.load(os.path.join(HERE, "golden", "kernels.npz"))
    shape = g["invlap"].shape
    nzh = shape[2] // 2 + 1
    kvec = K.fftk(shape)
    half = lambda a: a[..., :nzh] if a.shape[-1] == shape[2] else a
    for d in range(3):
        np.testing.assert_array_equal(kvec[d].numpy(), half(g[f"k{d}"]))
        np.testing.assert_allclose(K.gradient_kernel(kvec, d).numpy(), half(g[f"grad{d}_o1"]), rtol=1e-6, atol=1e-7)
        # order 0: i k with the Nyquist mode zeroed - on the half axis that is the LAST entry, not index len // 2
        np.testing.assert_allclose(K.gradient_kernel(kvec, d, order=0).numpy(), half(g[f"grad{d}_o0"]), rtol=1e-6,
                                   atol=1e-7)
    np.testing.assert_allclose(K.invlaplace_kernel(kvec).numpy(), half(g["invlap"]), rtol=1e-6)
    np.testing.assert_allclose(K.invlaplace_kernel(kvec, fd=True).numpy(), half(g["invlap_fd"]), rtol=3e-6)
    assert K.longrange_kernel(kvec, 0) == float(g["longrange_r0"]) == 1.0
    np.testing.assert_allclose(K.longrange_kernel(kvec, 1.5).numpy(), half(g["longrange_r1p5"]), rtol=1e-5, atol=1e-30)
    np.testing.assert_allclose(K.cic_compensation(kvec).numpy(), half(g["cic_comp"]), rtol=1e-5)
    assert K.laplace_kernel is K.invlaplace_kernel
    # PGD filter: the tabulated radial form the fused pass consumes == the broadcast form
    tab, kmax = K.pgd_filter_table(0.4, 2.5, 1 << 16)
    kk = sum(k**2 for k in kvec).sqrt().numpy()
    ref = K.PGD_kernel(kvec, 0.4, 2.5).numpy()
    got = np.interp(kk, np.linspace(0, kmax, tab.size), tab)
    np.testing.assert_allclose(got, ref, atol=2e-5)


def test_gaussian_smoothing_matches_the_reference_formula():
    """jaxpm/utils.py:208-222 restated in NumPy (scipy.stats.norm.pdf filter normalised at k = 0) on a square image."""
    from scipy.stats import norm

    from jaxpm_b200.utils import gaussian_smoothing
    rng = np.random.default_rng(4)
    im = rng.standard_normal((24, 24)).astype(np.float32)
    for sigma in (0.7, 2.5):
        kvec = np.stack(np.meshgrid(np.fft.fftfreq(im.shape[0]), np.fft.fftfreq(im.shape[1])), axis=-1)
        k = np.linalg.norm(kvec, axis=-1)
        filt = norm.pdf(k, 0, 1.0 / (2.0 * np.pi * sigma))
        filt /= filt[0, 0]
        ref = np.fft.ifft2(np.fft.fft2(im) * filt).real
        got = gaussian_smoothing(torch.as_tensor(im), sigma).numpy()
        assert got.dtype == np.float32
        np.testing.assert_allclose(got, ref, atol=2e-6)
        assert abs(got.mean() - im.mean()) < 1e-6          # the k = 0 mode is untouched
