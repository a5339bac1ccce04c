"""Independent brute-force anchor for the raycast oracle (TEST INFRASTRUCTURE).

The reference has no raycaster (parity unpinned), so the DDA
oracle in ``c/auvpf_oracle.c`` is anchored here against a completely different
formulation: explicit triangle soup + Moller-Trumbore ray/triangle intersection
against EVERY triangle of the DEM, nearest hit wins.  Only usable on small
DEMs; ``tests/test_oracle_raycast.py`` runs it.
"""
import numpy as np


def dem_triangles(dem):
    """(T, 3, 3) float64 vertices; two triangles per cell, diagonal (i,j)->(i+1,j+1)."""
    z = np.asarray(dem["z"], np.float64)
    H, W = z.shape
    xs = dem["x0"] + np.arange(W) * dem["res"]
    ys = dem["y0"] + np.arange(H) * dem["res"]
    X, Y = np.meshgrid(xs, ys)
    V = np.stack([X, Y, z], -1)
    v00, v01, v10, v11 = V[:-1, :-1], V[:-1, 1:], V[1:, :-1], V[1:, 1:]
    lower = np.stack([v00, v01, v11], -2).reshape(-1, 3, 3)     # u >= v
    upper = np.stack([v00, v11, v10], -2).reshape(-1, 3, 3)     # u <  v
    return np.concatenate([lower, upper], 0)


def raycast(dem, origins, dirs, range_max, eps=1e-12):
    """Nearest Moller-Trumbore hit with t in [0, range_max]; misses -> range_max."""
    tris = dem_triangles(dem)
    v0, e1, e2 = tris[:, 0], tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
    out = np.full(len(origins), float(range_max))
    for r, (o, d) in enumerate(zip(np.asarray(origins, float), np.asarray(dirs, float))):
        h = np.cross(d, e2)
        a = np.einsum("ij,ij->i", e1, h)
        ok = np.abs(a) > eps
        f = np.where(ok, 1.0 / np.where(ok, a, 1.0), 0.0)
        s = o - v0
        u = f * np.einsum("ij,ij->i", s, h)
        q = np.cross(s, e1)
        v = f * np.einsum("ij,j->i", q, d)
        t = f * np.einsum("ij,ij->i", e2, q)
        tol = 1e-9
        hit = ok & (u >= -tol) & (v >= -tol) & (u + v <= 1 + tol) & (t >= 0.0) & (t <= range_max)
        if hit.any():
            out[r] = t[hit].min()
    return out
