"""CPU-only: libfissgpu.so builds for sm_100a, loads, and exports every symbol include/fiss_abi.h
declares; host-only helpers behave (no kernel is launched here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib():
    from fiss_plus_planner_b200 import _shim, build
    build.build()
    return _shim.load()


def test_header_symbols_are_exported(lib):
    from fiss_plus_planner_b200 import _shim
    header = open(os.path.join(ROOT, "include", "fiss_abi.h")).read()
    declared = set(re.findall(r"\b(fiss_[a-z_0-9]+)\s*\(", header))
    declared -= {"fiss_last_error"} - set(re.findall(r"\bfiss_last_error\b", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in fiss_abi.h but not exported"
    assert declared == set(_shim.EXPORTS)
    assert lib.fiss_abi_version() == 2


def test_params_struct_layout():
    from fiss_plus_planner_b200._shim import FissParams
    assert ctypes.sizeof(FissParams) == 12 * 8 + 4 * 4


def test_arange_len_matches_numpy(lib):
    rng = np.random.default_rng(3)
    ts = np.concatenate((np.linspace(4, 5, 5), np.linspace(8, 10, 5), np.linspace(8, 10, 9), rng.uniform(0.05, 12, 2000)))
    for t in ts:
        assert lib.fiss_arange_len(float(t), 0.1) == len(np.arange(0.0, t, 0.1)), t
    assert lib.fiss_arange_len(0.0, 0.1) == 0
    assert lib.fiss_arange_len(-1.0, 0.1) == 0


def test_create_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fiss_plus_planner_b200.engine import FissEngine
    from fiss_plus_planner_b200._shim import FissError
    with pytest.raises(FissError, match="no CUDA device"):
        FissEngine(0)


def test_lattice_tables_match_reference_enumeration(lib):
    """fop_lattice order = d outer, T middle, v inner; fiss_lattice = [d][v][t] with +0.3 width."""
    from fiss_plus_planner_b200.engine import fiss_lattice, fop_lattice
    from fiss_plus_planner_b200.planners.frenet_optimal_planner import FrenetOptimalPlannerSettings
    st = FrenetOptimalPlannerSettings(3, 4, 5)
    tab = fop_lattice(st, 1.844)
    sw = 3.5 - 1.844
    want = [(d, v, t) for d in np.linspace(-sw / 2, sw / 2, 3) for t in np.linspace(8, 10, 5)
            for v in np.linspace(0, 13.4112, 4)]
    np.testing.assert_array_equal(tab[:, :3], np.array(want))
    np.testing.assert_array_equal(tab[:, 3], [len(np.arange(0.0, t, 0.1)) for _, _, t in want])
    ftab, ds, vs, ts, res = fiss_lattice(st, 1.844)
    sw = 3.5 - 1.844 + 0.3
    want = [(d, v, t) for d in np.linspace(-sw / 2, sw / 2, 3) for v in np.linspace(0, 13.4112, 4)
            for t in np.linspace(8, 10, 5)]
    np.testing.assert_array_equal(ftab[:, :3], np.array(want))
