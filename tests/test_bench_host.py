"""Host-side logic of the measurement scripts (no GPU): the step schedule of bench.py and the NVLink counter parser of
bench_multi.py."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_schedule_ends_at_a_equal_one_for_any_step_count():
    import bench
    from jaxpm_b200.cosmology import Planck15
    from jaxpm_b200.ode import kick_drift_coefficients
    cosmo = Planck15()
    ref_d, ref_k = kick_drift_coefficients(cosmo, 0.1, 1.0, 40, "symplectic")
    for K, W, n_pre_expected, total in ((10, 3, 30, 40), (5, 3, 35, 40), (40, 3, 3, 43), (60, 5, 5, 65)):
        a = argparse.Namespace(schedule_steps=40, steps=K, warmup=W)
        n_pre, d, k = bench.schedule(cosmo, a, kick_drift_coefficients)
        assert n_pre == n_pre_expected and len(k) == total and n_pre >= W
        if total == 40:      # the timed K steps are the LAST K of the 40-step run of BASELINE.json's configs
            np.testing.assert_allclose(np.asarray(d, dtype=np.float64), np.asarray(ref_d, dtype=np.float64))
            np.testing.assert_allclose(np.asarray(k, dtype=np.float64), np.asarray(ref_k, dtype=np.float64))
        else:                # more steps than the schedule: same interval in a, proportionally smaller steps
            assert abs(float(np.sum(d)) / float(np.sum(ref_d)) - 1) < 0.05


def test_nvlink_counter_parser(monkeypatch):
    import subprocess

    import bench_multi

    class R:
        def __init__(self, out):
            self.stdout = out
    sample = ("GPU 0: NVIDIA B200 (UUID: GPU-x)\\n\\t Link 0: Data Tx: 1000 KiB\\n\\t Link 0: Data Rx: 10 KiB\\n"
              "\\t Link 1: Data Tx: 24 KiB\\n\\t Link 1: Data Rx: 6 KiB\\n")
    monkeypatch.setattr(subprocess, "run", lambda *a, **kw: R(sample))
    assert bench_multi.nvlink_counters(0) == (1024 * 1024, 16 * 1024)
    # this pool: the driver reports N/A -> None, and the bench line says so instead of inventing a number
    monkeypatch.setattr(subprocess, "run", lambda *a, **kw: R("GPU 0: x\\n\\t Link 0: Data Tx: N/A\\n\\t Link 0: Data Rx: N/A\\n"))
    assert bench_multi.nvlink_counters(0) is None
