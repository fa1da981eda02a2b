"""Binding of the CPU oracle (oracle/liborc.so) for tests: the checker, never the product."""
import ctypes as C
import os

import numpy as np

from percept_b200.capi import Api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = C.CDLL(os.path.join(ROOT, "oracle", "liborc.so"))
_dp = C.POINTER(C.c_double)
_lib.orc_fixture_cube.argtypes = [C.c_uint] * 3 + [_dp] * 3
_lib.orc_fixture_cube.restype = None
_lib.orc_perturb_cube_rand.argtypes = [C.c_uint, C.c_double, C.c_uint, _dp]
_lib.orc_perturb_cube_rand.restype = None
_lib.orc_total_metric_ld.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_longdouble), C.POINTER(C.c_size_t)]
_lib.orc_nodal_field_dot_ld.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_longdouble)]

oracle = Api(_lib, "orc_", product=False)


def fixture_cube(nn, width=(1.0, 1.0, 1.0), offset=(0.0, 0.0, 0.0)):
    """BlockStructuredGrid::fixture_1 coordinates, SoA (3, nx*ny*nz), i fastest."""
    nn = tuple(nn) if hasattr(nn, "__len__") else (nn, nn, nn)
    x = np.zeros((3, nn[0] * nn[1] * nn[2]))
    _lib.orc_fixture_cube(nn[0], nn[1], nn[2], (C.c_double * 3)(*width), (C.c_double * 3)(*offset),
                          x.ctypes.data_as(_dp))
    return x


def perturb_cube_rand(x, nele, epsilon, seed=1):
    """build_perturbed_cube (HeavyTestMeshSmoother.cpp:124-159): glibc srand(seed)/rand()."""
    x = np.ascontiguousarray(x.copy())
    _lib.orc_perturb_cube_rand(nele, epsilon, seed, x.ctypes.data_as(_dp))
    return x


def total_metric_ld(ctx, stage, alpha=0.0):
    m, n = C.c_longdouble(), C.c_size_t()
    rc = _lib.orc_total_metric_ld(ctx.p, stage, alpha, C.byref(m), C.byref(n))
    assert rc == 0
    return m.value, n.value
