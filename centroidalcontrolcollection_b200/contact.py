"""Contact geometry flattening: surface contact -> per-input (vertex, ridge) tables.

Restates what the hot path reads from ForceColl::SurfaceContact::vertexWithRidgeList_
(reference src/DdpCentroidal.cpp:49-60, tests/src/ContactManager.h:10-21; ForceColl itself is
an external dependency): for every vertex, in the order given, the vertex
and the `ridge_num` unit ridges of a linearised friction pyramid with coefficient mu,
rho_i = normalize(mu cos(2 pi i / n), mu sin(2 pi i / n), 1), rotated into the world frame.
"""
import math

import numpy as np


def friction_pyramid(mu, ridge_num=4):
    out = np.zeros((ridge_num, 3))
    for i in range(ridge_num):
        theta = 2.0 * math.pi * (float(i) / ridge_num)
        v = np.array([mu * math.cos(theta), mu * math.sin(theta), 1.0])
        out[i] = v / np.linalg.norm(v)
    return out


def surface_contact(vertices, mu=0.5, rot=None, ridge_num=4):
    """-> (vertex[m,3], ridge[m,3]) with m = len(vertices) * ridge_num."""
    local = friction_pyramid(mu, ridge_num)
    rot = np.eye(3) if rot is None else np.asarray(rot, dtype=np.float64)
    ridges = local @ rot  # R^T * rho for each row rho
    vtx, rdg = [], []
    for v in vertices:
        for r in ridges:
            vtx.append(np.asarray(v, dtype=np.float64))
            rdg.append(r)
    return np.array(vtx).reshape(-1, 3), np.array(rdg).reshape(-1, 3)


def contact_from_rect(rect_min, rect_max, mu=0.5):
    """makeContactFromRect (reference tests/src/ContactManager.h:10-21): z = 0, identity pose."""
    (x0, y0), (x1, y1) = rect_min, rect_max
    verts = [(x0, y0, 0.0), (x0, y1, 0.0), (x1, y1, 0.0), (x1, y0, 0.0)]
    return surface_contact(verts, mu)


def total_wrench(vertex, ridge, scales, origin):
    """ForceColl::calcTotalWrench: force = sum l_i rho_i, moment about `origin`."""
    f = (scales[:, None] * ridge).sum(axis=0)
    n = (scales[:, None] * np.cross(vertex - np.asarray(origin)[None, :], ridge)).sum(axis=0)
    return f, n
