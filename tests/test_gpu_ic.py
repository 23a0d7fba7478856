"""SURVEY.md section 8f row 1 on the device: `normal_field` (own counter-based generator, pinned by oracle/rng.py) and
`linear_field` (jaxpm/pm.py:129-144) through the fused FFT chain, against the oracle on SHARED white noise."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import pm as OPM
from oracle import rng as ORNG

pytestmark = pytest.mark.gpu


def T(x, dev):
    return torch.as_tensor(np.ascontiguousarray(x)).to(dev)


@pytest.mark.parametrize("shape", [(16, 16, 16), (32, 16, 64), (12, 20, 18), (8, 8, 7)])
def test_normal_field_matches_the_oracle_generator(cuda, shape):
    from jaxpm_b200.distributed import Sharding, normal_field
    for seed in (0, 42, 2**40 + 17):
        ref = ORNG.normal_field(seed, shape)
        got = normal_field(seed, shape, device=cuda).cpu().numpy()
        assert got.shape == shape
        assert np.abs(got - ref).max() < 4e-6        # float32 log / sincospi against float64
    # a sharded call draws the block of the single-device field (decomposition-independent ICs)
    if shape[0] % 4 == 0 and shape[1] % 2 == 0:
        full = normal_field(7, shape, device=cuda).cpu().numpy()
        for pd in ((2, 2), (4, 1), (1, 2)):
            lx, ly = shape[0] // pd[0], shape[1] // pd[1]
            for r in range(pd[0] * pd[1]):
                sh = Sharding(pd, rank=r)
                blk = normal_field(7, shape, sharding=sh, device=cuda).cpu().numpy()
                np.testing.assert_array_equal(blk, full[sh.rx * lx:(sh.rx + 1) * lx, sh.ry * ly:(sh.ry + 1) * ly])


def test_normal_field_moments_at_size(cuda):
    from jaxpm_b200.distributed import normal_field
    z = normal_field(3, (256, 256, 256), device=cuda).double()
    n = z.numel()
    assert abs(float(z.mean())) < 5 / n**0.5 and abs(float(z.var()) - 1) < 5 * (2 / n)**0.5
    assert abs(float((z**4).mean()) - 3) < 0.01 and float(z.abs().max()) < 6.7
    # neighbouring cells / planes are uncorrelated
    for d in range(3):
        c = float((z * torch.roll(z, 1, dims=d)).mean())
        assert abs(c) < 5 / n**0.5, (d, c)


@pytest.mark.parametrize("shape,box", [((32, 32, 32), (100., 100., 100.)), ((64, 32, 128), (128., 64., 512.)),
                                       ((24, 40, 20), (60., 80., 45.))])
def test_linear_field_against_oracle_on_shared_white_noise(cuda, shape, box):
    """Power-of-two meshes run the fused chain (amplitude table inside the x pass), the others cuFFT + the table
    pass; both against oracle.pm.linear_field (pm.py:129-144 restated) on the same N(0,1) array, incl. the k = 0
    rule (the mode is multiplied by sqrt(P(0) Nc / V) = 0 for this spectrum)."""
    from jaxpm_b200.cosmology import Planck15, linear_matter_power
    from jaxpm_b200.pm import linear_field
    c = Planck15()
    pk = lambda k: linear_matter_power(c, k)
    wn = ORNG.normal_field(11, shape, dtype=np.float32)
    wn += np.float32(0.3)                              # a mean, so that the k = 0 rule is visible
    ref = OPM.linear_field(wn.astype(np.float64), box, pk)
    got = linear_field(shape, box, pk, seed=0, white_noise=T(wn, cuda)).cpu().numpy()
    assert rel_err(got, ref) < 1e-5
    assert abs(float(got.astype(np.float64).mean())) < 1e-5 * np.abs(ref).max()
    # seed path: the device generator feeds the same chain
    f = linear_field(shape, box, pk, seed=11, device=cuda).cpu().numpy()
    wn0 = ORNG.normal_field(11, shape, dtype=np.float64)
    assert rel_err(f, OPM.linear_field(wn0, box, pk)) < 2e-5
