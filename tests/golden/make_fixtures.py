"""Extracts the reference's own mesh fixtures (classic netCDF-3 Exodus / Nemesis files) into small .npz files.

Run in the build container where /root/reference exists:   python tests/golden/make_fixtures.py
Sources (read-only): build-cmake/hex8-ref.e            12^3 hex8 = hex8.e refined twice by mesh_adapt (scrambled ids)
                     build-cmake/hex8-ref.e.3.{0,1,2}  the same mesh, 3-rank Nemesis decomposition with n_comm maps
                     test/unit_tests/exodus_files/tet_from_hex_fixture_0.e   1536 tet4 / 429 nodes
The GPU box has no /root/reference; tests read only the committed .npz files.
"""
import os
import sys

import numpy as np
from scipy.io import netcdf_file

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_exo(path):
    nc = netcdf_file(path, "r", mmap=False)
    v = nc.variables
    d = {"coords": np.stack([v["coordx"][:], v["coordy"][:], v["coordz"][:]]).astype(np.float64),
         "conn": (v["connect1"][:].astype(np.int64) - 1),
         "node_num_map": v["node_num_map"][:].astype(np.int64), "elem_num_map": v["elem_num_map"][:].astype(np.int64)}
    if "n_comm_nids" in v:
        d["n_comm_nids"] = v["n_comm_nids"][:].astype(np.int64) - 1
        d["n_comm_proc"] = v["n_comm_proc"][:].astype(np.int64)
    nc.close()
    return d


def main():
    if not os.path.isdir(REF):
        sys.exit("no /root/reference here; fixtures are already committed")
    h = read_exo(f"{REF}/build-cmake/hex8-ref.e")
    np.savez_compressed(f"{OUT}/hex8_ref.npz", **h)
    parts = {}
    for r in range(3):
        p = read_exo(f"{REF}/build-cmake/hex8-ref.e.3.{r}")
        for k, val in p.items():
            parts[f"r{r}_{k}"] = val
    np.savez_compressed(f"{OUT}/hex8_ref_3ranks.npz", **parts)
    t = read_exo(f"{REF}/test/unit_tests/exodus_files/tet_from_hex_fixture_0.e")
    np.savez_compressed(f"{OUT}/tet_from_hex_fixture_0.npz", **t)
    for f in ("hex8_ref.npz", "hex8_ref_3ranks.npz", "tet_from_hex_fixture_0.npz"):
        print(f, os.path.getsize(f"{OUT}/{f}"), "bytes")


if __name__ == "__main__":
    main()
