"""CPU: the C-ABI library loads, exports every symbol include/ellc_gn.h declares, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    g.build()
    from egomotion_with_local_loop_closures_b200 import capi as m
    return m


def test_every_declared_symbol_is_exported(capi):
    hdr = open(os.path.join(ROOT, "include", "ellc_gn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ellc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    L = capi.lib()
    for s in declared:
        assert getattr(L, s) is not None


def test_struct_sizes_match_header(capi):
    assert C.sizeof(capi.Result) == 256 and C.sizeof(capi.IterTrace) == 256 and C.sizeof(capi.Pair) == 36
    assert capi.RESULT_DTYPE.itemsize == 256 and capi.TRACE_DTYPE.itemsize == 256


def test_default_config_mirrors_extern_variable_h(capi):
    cfg = capi.default_config(480, 270)
    assert list(cfg.max_iter) == [4, 7, 9, 12]                       # src/main.cpp:34
    assert cfg.huber_d == 3.0 and cfg.camera_pixel_noise_2 == 16.0   # src/ExternVariable.h:148-149
    assert list(cfg.weight) == [1e5, 1e5, 1e5, 1e4, 1e4, 1e4]        # :76
    assert cfg.stop_threshold == 1.0 and cfg.cx == 240.0 and cfg.cy == 135.0


def test_no_cpu_fallback(capi):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present: the failure path is only observable on a CPU box")
    with pytest.raises(capi.EllcError) as e:
        capi.Tracker(capi.default_config(640, 480))
    assert "no CPU fallback" in str(e.value)


def test_invalid_config_rejected(capi):
    h = C.c_void_p()
    cfg = capi.default_config(8, 8)
    assert capi.lib().ellc_create(C.byref(cfg), C.byref(h)) == -1
    assert capi.lib().ellc_create(None, C.byref(h)) == -1


def test_host_pose_algebra_matches_oracle(capi, oracle_mod):
    """The product's pose algebra restates the same published algorithms (Pade fp32 exp, exact log) as the oracle does
    independently: on the host the two agree bit for bit."""
    rng = np.random.default_rng(0)
    for i in range(400):
        sc = [0.02, 0.3, 1.5, 4.0][i % 4]
        a = (rng.standard_normal(6) * sc).astype(np.float32)
        b = (rng.standard_normal(6) * sc * 0.5).astype(np.float32)
        assert np.array_equal(capi.concat_relative(a, b), oracle_mod.concat_relative(a, b))
        assert np.array_equal(capi.concat_origin(a, b), oracle_mod.concat_origin(a, b))
        assert np.array_equal(capi.se3_exp(b), oracle_mod.se3_exp(b))
    z = np.zeros(6, np.float32)
    assert np.all(capi.concat_relative(z, z) == 0) and np.abs(capi.concat_origin(b, b)).max() < 1e-6
