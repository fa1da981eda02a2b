"""Analytic geometry for the `smooth_surfaces` hooks: node classification against a list of analytic surfaces.

The reference classifies nodes against CAD entities (`MeshGeometry::classify_node`, mesh/geometry/kernel/MeshGeometry.hpp:
dof 0 vertex, 1 curve, 2 surface, 3 volume) on the host, once, before smoothing; the evaluators themselves (normal_at /
snap_points_to_geometry) run on the GPU here (`pms_add_analytic_surface`).  This module is the host-side classifier for
analytic surfaces: a boundary (skin) node lying on one surface is a surface node, on two a curve node (their
intersection), on three or more a vertex; skin nodes on no listed surface are held fixed (treated as vertices), like
the reference treats boundary nodes without a geometry association.  Set-up code, not on the hot path.
"""
import json

import numpy as np

PLANE, SPHERE, CYLINDER = 0, 1, 2
_KINDS = {"plane": PLANE, "sphere": SPHERE, "cylinder": CYLINDER}


def surface_params(s):
    """dict -> (kind, 7 params) as pms_add_analytic_surface takes them."""
    kind = _KINDS[s["kind"]] if isinstance(s["kind"], str) else int(s["kind"])
    if kind == PLANE:
        p = list(s["point"]) + list(s["normal"]) + [0.0]
    elif kind == SPHERE:
        p = list(s["centre"] if "centre" in s else s["center"]) + [float(s["radius"]), 0.0, 0.0, 0.0]
    else:
        p = list(s["point"]) + list(s["axis"]) + [float(s["radius"])]
    return kind, np.asarray(p, dtype=np.float64)


def distance(kind, p, x):
    """unsigned distance of the points x (3, N) from the surface"""
    d = x - p[:3, None]
    if kind == SPHERE:
        return np.abs(np.linalg.norm(d, axis=0) - p[3])
    u = p[3:6] / np.linalg.norm(p[3:6])
    t = u @ d
    if kind == PLANE:
        return np.abs(t)
    return np.abs(np.linalg.norm(d - u[:, None] * t[None, :], axis=0) - p[6])


def classify_nodes(coords, skin_mask, surfaces, tol):
    """coords (3, N) on-geometry coordinates (the reference mesh), skin_mask (N,) boundary nodes, surfaces: list of dicts.
    Returns (dof int8 (N,), surface_id int32 (N,))."""
    n = coords.shape[1]
    hits = np.zeros(n, dtype=np.int32)
    first = np.full(n, -1, dtype=np.int32)
    for sid, s in enumerate(surfaces):
        kind, p = surface_params(s)
        on = (distance(kind, p, coords) <= tol) & (skin_mask != 0)
        first = np.where(on & (first < 0), sid, first)
        hits += on
    dof = np.full(n, 3, dtype=np.int8)
    skin = skin_mask != 0
    dof[skin & (hits == 0)] = 0   # boundary node without a geometry association: held fixed
    dof[skin & (hits == 1)] = 2
    dof[skin & (hits == 2)] = 1
    dof[skin & (hits >= 3)] = 0
    surf = np.where(dof == 2, first, -1).astype(np.int32)
    return dof, surf


def load_surfaces(path):
    """JSON list, e.g. [{"kind": "plane", "point": [0,0,0], "normal": [0,0,-1]}, {"kind": "cylinder", "point": [...],
    "axis": [...], "radius": 3.0}, {"kind": "sphere", "centre": [...], "radius": 1.0}]"""
    with open(path) as f:
        return json.load(f)
