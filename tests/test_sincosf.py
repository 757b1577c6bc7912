"""The glibc-sincosf restatement (botlab_b200/csrc/glibc_sincosf.h) against the live libm, on the CPU (strided sweep;
the exhaustive sweep over every float in [-4, 4] was run once and is recorded in DESIGN.md) and on the GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def test_cpu_strided_sweep(tmp_path):
    exe = str(tmp_path / "sincosf_sweep")
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-o", exe,
                           os.path.join(ROOT, "tests", "csrc", "sincosf_sweep.c"), "-lm"])
    out = subprocess.check_output([exe, "4.0", "61"]).decode()
    assert "mismatches 0" in out, out
    out = subprocess.check_output([exe, "119.9", "257"]).decode()
    assert "mismatches 0" in out, out


@pytest.mark.gpu
def test_gpu_matches_libm_bit_for_bit():
    from botlab_b200 import engine
    libm = C.CDLL("libm.so.6")
    libm.sincosf.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(0)
    n = 2_000_000
    x = np.concatenate([
        rng.uniform(-np.pi, np.pi, n).astype(np.float32),
        rng.uniform(-100, 100, n // 4).astype(np.float32),
        (rng.standard_normal(n // 4) * 1e-3).astype(np.float32),
        np.float32([0.0, -0.0, np.pi, -np.pi, np.pi / 2, np.pi / 4, 0.78539819, 0.7853981, 2.4e-4, 1e-30]),
    ])
    e = engine.Engine(16)
    s, c = e.debug_sincosf(x)
    s_ref = np.zeros_like(x)
    c_ref = np.zeros_like(x)
    sv, cv = C.c_float(), C.c_float()
    # libm call per element through ctypes is slow: check a 200k sample + all specials
    pick = np.concatenate([rng.choice(len(x) - 10, 200_000, replace=False), np.arange(len(x) - 10, len(x))])
    for i in pick:
        libm.sincosf(C.c_float(float(x[i])), C.addressof(sv), C.addressof(cv))
        s_ref[i], c_ref[i] = sv.value, cv.value
    assert np.array_equal(s[pick].view(np.uint32), s_ref[pick].view(np.uint32))
    assert np.array_equal(c[pick].view(np.uint32), c_ref[pick].view(np.uint32))
