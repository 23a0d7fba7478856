"""The C-ABI library loads and exports exactly the symbols include/jaxpm_b200.h declares (CPU)."""
import os
import re

from jaxpm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "jaxpm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(jpm_[a-z0-9_]+)\s*\(", src))


def test_header_matches_bindings():
    hdr = _header_symbols()
    assert hdr == set(_lib.SIGNATURES), hdr ^ set(_lib.SIGNATURES)


def test_library_exports_every_symbol():
    lib = _lib.load()
    for name in _header_symbols():
        assert hasattr(lib, name), name
    assert lib.jpm_abi_version() == 1


def test_no_cpu_fallback():
    import numpy as np
    import pytest
    import torch
    from jaxpm_b200.painting import cic_paint
    with pytest.raises(_lib.JpmError):
        cic_paint(torch.zeros(4, 4, 4), torch.zeros(4, 4, 4, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "jaxpm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f


def test_xla_ffi_handlers_type_check():
    """jaxpm_b200/csrc/xla_ffi.cc (the jax.ffi binding of INTEGRATION.md section 3) compiles against the C ABI: every
    handler body is type-checked with the stand-in for jaxlib's header (tools/xla_ffi_stub; JAX is not installable
    here), so a changed entry-point signature breaks this test instead of a maintainer's build."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["make", "-C", os.path.join(root, "jaxpm_b200", "csrc"), "ffi-check"], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    src = open(os.path.join(root, "jaxpm_b200", "csrc", "xla_ffi.cc")).read()
    for name in ("JpmCicPaint", "JpmCicRead", "JpmCicPaintDx", "JpmCicRead3", "JpmRead3KickDrift",
                 "JpmDensityToForceMeshes", "JpmSimForces", "JpmSimStep", "JpmCicReadGrad", "JpmCicPaintGrad",
                 "JpmGreensDiv", "JpmPmForcesVjp", "JpmSlabForces", "JpmNormalField"):
        assert f"XLA_FFI_DEFINE_HANDLER_SYMBOL({name}," in src, name
