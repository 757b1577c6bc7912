"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/mcl_cuda.h declares, refuses to
run without a GPU (no CPU fallback), and its host-side scalar action model matches the reference's golden values."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from botlab_b200 import engine


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mcl_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcl_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(engine.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = engine.lib()
    for s in declared_symbols():
        assert hasattr(L, s), s


def test_struct_sizes_match_reference_layout():
    assert C.sizeof(engine.Pose) == 24           # sizeof(pose_xyt_t), SURVEY B13
    assert engine.PARTICLE_DTYPE.itemsize == 56  # sizeof(particle_t)


def test_default_params_are_the_reference_constants():
    p = engine.default_params()
    assert p.min_range == np.float32(0.15) and p.weight_floor == 0.001 and p.init_std == 0.01
    assert p.legacy_equal_utime == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.MclError, match="no CPU fallback"):
        engine.Engine(1000)


def test_bad_arguments_are_reported_not_fatal():
    L = engine.lib()
    h = C.c_void_p()
    assert L.mcl_create(None, 1, 0, C.addressof(h)) == -1      # particle_filter.cpp:11 asserts numParticles > 1
    assert b"num_particles" in L.mcl_last_error(None)
    assert L.mcl_sync(None) == -1


def test_host_action_update_matches_reference():
    k = load_golden("kat")
    am = engine.ActionModel()
    assert am.update(0, 0, 0) is False
    assert am.update(0.02, 0.01, 0.01) is True
    assert np.array_equal(am.params, k["action_forward"])       # SURVEY B8, bit for bit
    am2 = engine.ActionModel()
    am2.update(0, 0, 0)
    assert am2.update(-0.02, 0, 0) is True
    assert np.array_equal(am2.params, k["action_backward"])     # B10 (backward branch)
    a = load_golden("action")
    am3 = engine.ActionModel()
    am3.update(0.3, -0.2, 0.1)
    am3.update(0.32, -0.19, 0.11)
    assert np.array_equal(am3.params, a["params"])
    # below the motion threshold (action_model.cpp:52)
    am4 = engine.ActionModel()
    am4.update(1.0, 1.0, 0.5)
    assert am4.update(1.0, 1.0, 0.5) is False
