"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/zkm_b200.h declares, and -- with no GPU in this container -- refuses to compute instead
of falling back to a CPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "zkm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zkm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from zkmember_b200 import _lib
    L = _lib.load()
    decl = _declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(L, name), "libzkm_b200.so does not export %s" % name
    assert sorted(_lib.SYMBOLS) == decl, "zkmember_b200/_lib.py SYMBOLS out of sync with the header"


def test_library_is_sm100a_cuda_not_a_cpu_stub():
    """The .so must carry sm_100a device code (the product path is CUDA, not a host stub)."""
    import subprocess
    from zkmember_b200 import _lib
    try:
        out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from zkmember_b200 import _lib
    L = _lib.load()
    assert L.zkm_device_count() == 0
    assert L.zkm_init(0) == -2                        # ZKM_ERR_CUDA
    assert b"no CPU fallback" in L.zkm_last_error()
    x = np.zeros((8, 4), dtype=np.uint64)
    assert L.zkm_ntt(0, ctypes.c_void_p(x.ctypes.data), 3, 0, 0) == -3      # ZKM_ERR_NOT_INIT
    out = np.zeros(12, dtype=np.uint64)
    inf = np.zeros(1, dtype=np.uint8)
    rc = L.zkm_msm_g1(0, ctypes.c_void_p(0), ctypes.c_void_p(0), ctypes.c_void_p(0), 0,
                      ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(inf.ctypes.data))
    assert rc == -3
    import zkmember_b200 as zkm
    with pytest.raises(zkm.ZkmError):
        zkm.VariableBaseMSM.multi_scalar_mul(np.zeros((1, 12), dtype=np.uint64), np.zeros((1, 4), dtype=np.uint64))


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under zkmember_b200/ may import or load it."""
    pkg = os.path.join(ROOT, "zkmember_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "libzkm_oracle" not in text, f


def test_domain_new_mirrors_upstream_size_rules():
    from zkmember_b200.domain import Radix2EvaluationDomain, TWO_ADICITY
    # new() returns None above the two-adicity without touching the GPU
    assert Radix2EvaluationDomain.new((1 << 28) + 1, "bn254") is None
    assert Radix2EvaluationDomain.new((1 << 32) + 1, "bls12_381") is None
    assert Radix2EvaluationDomain.new((1 << 46) + 1, "bw6_761") is None
    assert TWO_ADICITY == {0: 32, 1: 28, 2: 46}


def test_window_bits_heuristic_is_monotone():
    from zkmember_b200 import msm_window_bits
    prev = 0
    for k in range(4, 27):
        c = msm_window_bits("bls12_381", 1, 1 << k)
        assert 2 <= c <= 24 and c >= prev
        prev = c


def test_window_model_matches_measured_optima():
    """The automatic window size is a cost model fitted to B200 sweeps; these are the measured
    optima it must keep reproducing: full top windows (c = 16, 20 for 255-bit scalars) once the sort matters."""
    from zkmember_b200 import msm_window_bits
    assert [msm_window_bits("bls12_381", 1, 1 << k) for k in (20, 21, 22, 23)] == [16, 16, 16, 16]
    assert [msm_window_bits("bls12_381", 1, 1 << k) for k in (24, 25, 26)] == [20, 20, 20]
    assert msm_window_bits("bls12_381", 1, 1 << 17) in (11, 12)
    assert msm_window_bits("bn254", 1, 1 << 22) == 17            # 254 = 14 * 17 + 16: a full top window
    assert msm_window_bits("bw6_761", 1, 1 << 20) == 14          # 377 = 26 * 14 + 13
    for curve in ("bls12_381", "bn254", "bw6_761"):
        prev = 0
        for k in range(4, 27):
            c = msm_window_bits(curve, 1, 1 << k)
            assert 2 <= c <= 24 and c >= prev, (curve, k, c)
            prev = c
