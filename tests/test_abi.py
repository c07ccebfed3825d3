"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the headers declare, the ctypes
mirrors agree with the C struct layouts, and the product path fails loudly (never falls back) without a CUDA device."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import mppi_generic_b200 as m

H = m.host
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for hdr in ("include/mppi_b200.h", "include/mppi_b200/host_twins.h"):
        txt = open(os.path.join(ROOT, hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(mppib_[a-z0-9_]+)\s*\(", txt))
    return names


def test_library_exports_every_declared_symbol():
    L = H.lib()
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"declared in include/ but not exported: {missing}"
    assert set(H.ABI_SYMBOLS) == declared, set(H.ABI_SYMBOLS) ^ declared


def test_ctypes_struct_layouts_match_the_c_headers():
    structs = {
        "mppib_control_limits": H.ControlLimits, "mppib_cartpole_dyn_params": H.CartpoleDynParams,
        "mppib_di_dyn_params": H.DIDynParams, "mppib_ar_nn_dyn_params": H.ARNNDynParams,
        "mppib_cartpole_cost_params": H.CartpoleCostParams, "mppib_di_circle_cost_params": H.DICircleCostParams,
        "mppib_ar_standard_cost_params": H.ARStandardCostParams, "mppib_gaussian_params": H.GaussianParams,
        "mppib_desc": H.Desc, "mppib_solve_stats": H.SolveStats, "mppib_timing": H.Timing,
    }
    offs = [("mppib_ar_standard_cost_params", "map_width"), ("mppib_gaussian_params", "offset_decay_rate"),
            ("mppib_cartpole_cost_params", "desired_terminal_state"), ("mppib_desc", "world_size"),
            ("mppib_cartpole_dyn_params", "gravity")]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "mppi_b200.h"', 'int main(void){']
    for n in structs:
        src.append(f'printf("%zu\\n", sizeof({n}));')
    for n, f in offs:
        src.append(f'printf("%zu\\n", offsetof({n}, {f}));')
    src.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write("\n".join(src))
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe])  # header is plain C
        vals = [int(v) for v in subprocess.check_output([exe]).split()]
    for (n, cls), v in zip(structs.items(), vals):
        assert C.sizeof(cls) == v, (n, C.sizeof(cls), v)
    for (n, f), v in zip(offs, vals[len(structs):]):
        assert getattr(structs[n], f).offset == v, (n, f)


def test_strerror_and_version():
    L = H.lib()
    assert L.mppib_version() >= 100
    assert b"no CUDA device" in L.mppib_strerror(-5)
    assert L.mppib_strerror(0) == b"ok"


def test_invalid_arguments_are_rejected_without_touching_the_gpu():
    L = H.lib()
    h = C.c_void_p()
    assert L.mppib_create(C.byref(h), None) == -1
    d = H.Desc(H.DYN_CARTPOLE, H.COST_CARTPOLE_QUADRATIC, 0, 0, 10, 1, 0, 0, None, 0, 1)
    assert L.mppib_create(C.byref(h), C.byref(d)) == -1          # num_rollouts = 0
    d = H.Desc(H.DYN_CARTPOLE, H.COST_DI_CIRCLE, 0, 64, 10, 1, 0, 0, None, 0, 1)
    assert L.mppib_create(C.byref(h), C.byref(d)) == -2          # no such (dynamics, cost) pair
    d = H.Desc(H.DYN_CARTPOLE, H.COST_CARTPOLE_QUADRATIC, 0, 64, 10, 3, 0, 0, None, 0, 1)
    assert L.mppib_create(C.byref(h), C.byref(d)) == -1          # D = 3
    assert L.mppib_solve(None, None, None, 1, 0, None, None) == -1
    assert L.mppib_destroy(None) == 0


def _gpu_present():
    try:
        e = m.Engine(m.CartpoleDynamics(), m.CartpoleQuadraticCost(), m.GaussianDistribution(1), 32, 4)
        e.close()
        return True
    except m.MppibError:
        return False


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU the product path must fail loudly — it never routes through the oracle or any CPU code."""
    if _gpu_present():
        pytest.skip("CUDA device present")
    with pytest.raises(m.MppibError) as ei:
        m.Engine(m.CartpoleDynamics(), m.CartpoleQuadraticCost(), m.GaussianDistribution(1), 64, 50)
    assert ei.value.status == -5
    w = m.workloads.cartpole(64, 50)
    with pytest.raises(m.MppibError):
        m.VanillaMPPIController(w.dyn, w.cost, None, w.sampler, w.dt, 1, w.lambda_, w.alpha, w.T, w.N)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package or include/ may reference it."""
    bad = []
    for base in ("mppi-generic_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".sh")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(import\s+oracle|from\s+oracle|oracle/|libmppi_oracle|mppi_oracle)", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_host_twin_step_matches_plugin_objects():
    # Dynamics::step host method through the plugin object (dynamics.cuh:283-290)
    dyn = m.CartpoleDynamics(1.0, 1.0, 1.0)
    xn, xd, y = dyn.step(np.array([0, 0, 0.1, 0], np.float32), np.array([1.0], np.float32), 0.01)
    assert xd[0] == 0 and xd[2] == 0 and xd[1] != 0
    np.testing.assert_array_equal(y, xn)
    u = np.array([9.0], np.float32)
    dyn.setControlRanges([(-5, 5)])
    dyn.enforceConstraints(None, u)
    assert u[0] == 5.0
