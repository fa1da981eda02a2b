"""Minimal Exodus II (classic netCDF-3) reader/writer for the smoother's inputs and outputs, plus the boundary-node
selection the production caller uses.

Reference side: `PerceptMesh::open` + Ioss (third party) read the mesh; `Refiner::snapAndSmooth`
(src/adapt/Refiner.cpp:1536-1565) builds the boundary selector from the surface parts / `fix_all_block_boundaries`.
Here: hex8 / tet4 element blocks and coordinates are read with scipy's netCDF-3 reader; the fixed mask is the skin of
the mesh (faces that belong to exactly one element) when no side/node sets are requested.
"""
import numpy as np

HEX8_FACES = np.array([[0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [0, 4, 7, 3], [0, 3, 2, 1], [4, 5, 6, 7]])  # Shards Hexahedron<8>
TET4_FACES = np.array([[0, 1, 3], [1, 2, 3], [0, 3, 2], [0, 2, 1]])  # Shards Tetrahedron<4>


def read_exodus(path):
    """Returns dict(coords (3,N) float64, blocks=[(topology, conn (nelem,npe) int32 0-based)], node_num_map, elem_num_map)."""
    from scipy.io import netcdf_file
    nc = netcdf_file(path, "r", mmap=False)
    v = nc.variables
    if "coordx" in v:
        coords = np.stack([v["coordx"][:], v["coordy"][:], v["coordz"][:]]).astype(np.float64)
    else:
        coords = np.ascontiguousarray(v["coord"][:].astype(np.float64))
    blocks = []
    b = 1
    while f"connect{b}" in v:
        c = v[f"connect{b}"]
        et = getattr(c, "elem_type", b"").decode() if isinstance(getattr(c, "elem_type", b""), bytes) else str(getattr(c, "elem_type", ""))
        conn = c[:].astype(np.int64) - 1
        npe = conn.shape[1]
        topo = 8 if npe == 8 else (4 if npe == 4 else None)
        if topo is None:
            raise ValueError(f"element block {b} ({et}, {npe} nodes/element) is outside the hex8/tet4 smoothing path")
        blocks.append((topo, np.ascontiguousarray(conn.astype(np.int32))))
        b += 1
    out = {"coords": np.ascontiguousarray(coords), "blocks": blocks,
           "node_num_map": v["node_num_map"][:].astype(np.int64) if "node_num_map" in v else np.arange(1, coords.shape[1] + 1),
           "elem_num_map": v["elem_num_map"][:].astype(np.int64) if "elem_num_map" in v else None}
    nc.close()
    return out


def skin_nodes(conn, topology):
    """uint8 mask of the nodes on faces that belong to exactly one element (the mesh boundary)."""
    faces_tab = HEX8_FACES if topology == 8 else TET4_FACES
    f = conn[:, faces_tab].reshape(-1, faces_tab.shape[1])
    key = np.sort(f, axis=1)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    boundary_faces = f[cnt[inv.reshape(-1)] == 1]
    mask = np.zeros(int(conn.max()) + 1, dtype=np.uint8)
    mask[boundary_faces.reshape(-1)] = 1
    return mask


def write_coordinates_npz(path, coords, **extra):
    np.savez_compressed(path, coords=coords, **extra)
