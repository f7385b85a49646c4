# -*- coding: utf-8 -*-
"""The C-ABI library loads and exports every symbol include/bgp.h declares (no compute calls: no GPU needed), and the
ctypes mirrors of the POD structs have the C layout."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "bgp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bgp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from george_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libbgp_b200.so missing: run __graft_entry__.build()"
    lib = C.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 50
    for name in names:
        assert hasattr(lib, name), "symbol {0} declared in include/bgp.h is not exported".format(name)
    assert sorted(_lib.SIGNATURES) == names, set(_lib.SIGNATURES) ^ set(names)


def test_struct_layouts():
    from george_b200._spec import HodlrNodeInfo, HodlrOpts, KernelNode, KernelSpec
    assert C.sizeof(KernelNode) == 8 * 4 + 8 * 4 + 4 * 8 + 36 * 8 + 8 * 8 + 8 * 8
    assert C.sizeof(KernelSpec) == 8 + 32 * C.sizeof(KernelNode)
    assert C.sizeof(HodlrOpts) == 40
    assert C.sizeof(HodlrNodeInfo) == 40


def test_host_only_entry_points_work_without_gpu():
    from george_b200 import _lib, kernels
    from george_b200._spec import flatten
    lib = _lib.load()
    assert lib.bgp_version() >= 1000
    assert lib.bgp_device_count() >= 0
    k = 1.0 * kernels.ExpSquaredKernel(1.0) + 0.5 * kernels.ExpSine2Kernel(gamma=1.0, log_period=0.3)
    spec = flatten(k)
    assert lib.bgp_spec_validate(C.byref(spec)) == 0
    n = C.c_int()
    assert lib.bgp_spec_num_params(C.byref(spec), C.byref(n)) == 0 and n.value == 5 == len(k)
    spec.nodes[0].kernel_type = 99
    assert lib.bgp_spec_validate(C.byref(spec)) == _lib.BGP_ERR_INVALID
    assert "kernel" in _lib.last_error()


def test_compute_without_device_fails_loudly():
    """No CPU fallback: on a box without a B200 every compute entry point must raise, never return numbers."""
    import numpy as np
    from george_b200 import _lib, kernels
    if _lib.load().bgp_device_count() > 0:
        pytest.skip("a GPU is present")
    k = kernels.ExpSquaredKernel(1.0)
    with pytest.raises(_lib.BGPError):
        k.get_value(np.zeros((3, 1)))
    import george_b200 as george
    gp = george.GP(k, solver=george.HODLRSolver)
    with pytest.raises(_lib.BGPError):
        gp.compute(np.linspace(0, 1, 10), 0.1)


def test_native_handles_are_parked_and_reused_but_look_fresh():
    """The reference builds a new solver on every GP.compute (gp.py:327); the native handle behind it (device buffers,
    rank capacities, the instantiated ACA graph) is parked by a dying Python solver and picked up by the next one.  A
    recycled handle must look like a brand-new solver: not computed, and every query raises until compute() ran."""
    import numpy as np
    import pytest
    from george_b200.solvers._hodlr import HODLRSolver, resolve_rng_mode
    HODLRSolver.release_parked()
    a = HODLRSolver()
    pa = a._ptr.value
    del a
    assert len(HODLRSolver._parked) == 1
    b = HODLRSolver()
    assert b._ptr.value == pa and len(HODLRSolver._parked) == 0
    assert b.computed == 0
    with pytest.raises(RuntimeError):
        b.log_determinant
    with pytest.raises(RuntimeError):
        b.dot_solve(np.zeros(3))
    with pytest.raises(RuntimeError):
        b.apply_inverse(np.zeros(3))
    c, d, e = HODLRSolver(), HODLRSolver(), HODLRSolver()
    del b, c, d, e                       # at most _max_parked handles are kept, the rest is destroyed
    assert len(HODLRSolver._parked) == HODLRSolver._max_parked
    HODLRSolver.release_parked()
    assert HODLRSolver._parked == []
    # default RNG order: the reference's own stream whenever the result depends on the pivots
    assert resolve_rng_mode(None, 0.1) == "reference" and resolve_rng_mode(None, 1e-6) == "pernode"
    assert resolve_rng_mode("reference", 1e-12) == "reference"
