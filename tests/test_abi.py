"""The C-ABI library loads on a machine without a GPU, exports every symbol that
include/psfm_b200.h declares, and fails loudly (no CPU fallback) when asked to compute."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from particlesfm_b200 import _abi, _lib, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "psfm_b200.h")).read()
    declared = set(re.findall(r"\b(psfm_[a-z_]+)\s*\(", hdr))
    L = _lib.lib()
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in psfm_b200.h but not exported"
    assert set(_lib.EXPORTS) == declared
    assert L.psfm_abi_version() == 2


def test_struct_layouts_match_defaults():
    L = _lib.lib()
    o = _abi.BAOptions()
    L.psfm_ba_global_options(C.byref(o))
    # controllers/global_mapper.cc:41-71
    assert (o.function_tolerance, o.gradient_tolerance, o.parameter_tolerance) == (1e-6, 1.0, 1e-8)
    assert (o.max_num_iterations, o.max_linear_solver_iterations) == (50, 100)
    assert o.loss_function_type == _abi.LOSS_SOFT_L1 and o.refine_rotation == 0 and o.refine_focal_length == 0
    assert o.max_num_consecutive_invalid_steps == 10 and o.jacobi_scaling == 1 and o.eta == 0.1
    t = _abi.TrajOptions()
    L.psfm_traj_default_options(C.byref(t))
    # trajectory_optimize.cpp:74-79 + Ceres defaults
    assert t.max_num_iterations == 200 and t.function_tolerance == 1e-6 and t.gradient_tolerance == 1e-10
    assert t.parameter_tolerance == 1e-8 and t.initial_trust_region_radius == 1e4


def test_oracle_and_product_defaults_agree():
    import oracle
    a, b = _abi.BAOptions(), _abi.BAOptions()
    _lib.lib().psfm_ba_global_options(C.byref(a))
    oracle.lib().psfm_oracle_ba_global_options(C.byref(b))
    b.exact_r_tolerance = a.exact_r_tolerance
    assert bytes(a) == bytes(b)
    ta, tb = _abi.TrajOptions(), _abi.TrajOptions()
    _lib.lib().psfm_traj_default_options(C.byref(ta))
    oracle.lib().psfm_oracle_traj_default_options(C.byref(tb))
    assert bytes(ta) == bytes(tb)


@pytest.mark.skipif(_lib.lib().psfm_device_count() > 0, reason="needs a machine WITHOUT a GPU")
def test_no_cpu_fallback():
    from particlesfm_b200 import ba, traj
    uv12, r1, r2, sc, f12 = syn.make_traj_inputs(10, 32, 32, seed=0)
    with pytest.raises(_lib.PsfmError, match="no CUDA device"):
        traj.optimize_location(uv12, r1, r2, sc, f12, 10, 32, 32)
    prob, _ = syn.make_ba_problem(3, 10, 2, seed=0)
    o = _abi.BAOptions()
    _lib.lib().psfm_ba_global_options(C.byref(o))
    with pytest.raises(_lib.PsfmError, match="no CUDA device"):
        ba.solve_problem(prob, o)


@pytest.mark.skipif(_lib.lib().psfm_device_count() > 0, reason="needs a machine WITHOUT a GPU")
def test_no_cpu_fallback_of_the_initialisation_ops():
    """SURVEY.md 8(f) f-4 ops: host-side argument checks come first, then the library refuses without a device."""
    import numpy as np
    from particlesfm_b200 import init_geometry as ig
    p = np.zeros((5, 2))
    q = np.array([1.0, 0.0, 0.0, 0.0])
    with pytest.raises(ValueError):
        ig.batch_optimize_relative_position_with_known_rotation([(p, p[:-1], q, q)])
    with pytest.raises(_lib.PsfmError, match="no CUDA device"):
        ig.optimize_relative_position_with_known_rotation(p, p, q, q)
    with pytest.raises(_lib.PsfmError, match="no CUDA device"):
        ig.triangulate_multi_view_points([(np.zeros((2, 3, 4)), np.zeros((2, 2)))])
    assert ig.batch_optimize_relative_position_with_known_rotation([]).shape == (0, 3)      # nothing to do: no device needed
    assert ig.triangulate_multi_view_points([]).shape == (0, 3)
