"""Synthetic meshes of the reference's smoother tests and of BASELINE.json's configs (host side, numpy).

Coordinates are SoA arrays of shape (3, nnodes), i fastest (PMS_SOA).
"""
import ctypes as C

import numpy as np

from . import lib as _lib

_dp = C.POINTER(C.c_double)
_lib.pms_fixture_cube.argtypes = [C.c_uint] * 3 + [_dp] * 3
_lib.pms_fixture_cube.restype = None
_lib.pms_fixture_perturb_rand.argtypes = [C.c_uint] * 3 + [C.c_double, C.c_uint, _dp]
_lib.pms_fixture_perturb_rand.restype = None


def _n3(nn):
    return tuple(int(v) for v in nn) if hasattr(nn, "__len__") else (int(nn),) * 3


def cube(nn, width=(1.0, 1.0, 1.0), offset=(0.0, 0.0, 0.0)):
    """BlockStructuredGrid::fixture_1 (structured/StructuredBlock.cpp:101-168): nn nodes per direction."""
    nn = _n3(nn)
    x = np.zeros((3, nn[0] * nn[1] * nn[2]))
    _lib.pms_fixture_cube(nn[0], nn[1], nn[2], (C.c_double * 3)(*width), (C.c_double * 3)(*offset),
                          x.ctypes.data_as(_dp))
    return x


def perturb_rand(x, nn, epsilon, seed=1):
    """build_perturbed_cube (HeavyTestMeshSmoother.cpp:124-159), glibc srand(seed)/rand()."""
    nn = _n3(nn)
    y = np.ascontiguousarray(x.copy())
    _lib.pms_fixture_perturb_rand(nn[0], nn[1], nn[2], float(epsilon), int(seed), y.ctypes.data_as(_dp))
    return y


def perturb_sinusoid(x, h, amp=0.2, freqs=(3, 5, 7)):
    """BASELINE config 2: x_d += amp*h*sin(2pi*3x)sin(2pi*5y)sin(2pi*7z); vanishes on the cube faces."""
    b = amp * h * np.sin(2 * np.pi * freqs[0] * x[0]) * np.sin(2 * np.pi * freqs[1] * x[1]) * \
        np.sin(2 * np.pi * freqs[2] * x[2])
    return x + b[None, :]


def cube_boundary_mask(nn):
    """Nodes on the 6 faces (the 6 fixture_1 boundary conditions / sideset:xXyYzZ)."""
    nn = _n3(nn)
    i = np.arange(nn[0])[None, None, :]
    j = np.arange(nn[1])[None, :, None]
    k = np.arange(nn[2])[:, None, None]
    m = (i == 0) | (i == nn[0] - 1) | (j == 0) | (j == nn[1] - 1) | (k == 0) | (k == nn[2] - 1)
    return np.ascontiguousarray(np.broadcast_to(m, (nn[2], nn[1], nn[0])).reshape(-1).astype(np.uint8))


def hex8_connectivity(ncell):
    """Exodus/Shards hex8 connectivity of an ncell^3 (or (nx,ny,nz)) generated mesh: node id = i + (nx+1)(j + (ny+1)k),
    element id = ci + nx*(cj + ny*ck), local order (000,100,110,010,001,101,111,011) — cf. build-cmake/hex8.e
    connect1[0] = [1 2 6 5 17 18 22 21]."""
    nx, ny, nz = _n3(ncell)
    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ci, cj, ck = (a.transpose(2, 1, 0).reshape(-1) for a in (ci, cj, ck))  # ci fastest
    nid = lambda i, j, k: i + (nx + 1) * (j + (ny + 1) * k)
    offs = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    conn = np.stack([nid(ci + a, cj + b, ck + c) for a, b, c in offs], axis=1)
    return np.ascontiguousarray(conn.astype(np.int32))


def tet4_from_hex8(conn_hex):
    """6 Kuhn tetrahedra per hex around the diagonal (local 0 -> 6), positively oriented for
    jacobian_matrix_tet_3D (JacobianUtil.hpp:175-205).  Tet t of hex e is element 6e+t."""
    paths = [(1, 2), (2, 3), (3, 7), (7, 4), (4, 5), (5, 1)]  # consecutive neighbours around the 0-6 diagonal
    tets = [conn_hex[:, [0, a, b, 6]] for a, b in paths]
    return np.ascontiguousarray(np.stack(tets, axis=1).reshape(-1, 4).astype(np.int32))
