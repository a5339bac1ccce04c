"""Independent brute-force anchor for the raycast oracle (TEST INFRASTRUCTURE).

The reference has no raycaster (parity unpinned), so the DDA
oracle in ``c/auvpf_oracle.c`` is anchored here against a completely different
formulation: explicit triangle soup + Moller-Trumbore ray/triangle intersection
against EVERY triangle of the DEM, nearest hit wins; no grid walk, no cell
lookup, no line-family crossings.  Only usable on small DEMs;
``tests/test_oracle_raycast.py`` runs it.

``raycast``       nearest triangle hit with t in [0, range_max] -- the plain geometric question.
``raycast_spec``  the same under the RAYCAST SPEC's clip rule (``c/auvpf_oracle.c`` header): the ray
                  is first clipped to the map footprint, and a ray whose clipped start already
                  lies at or below the surface returns that start.  The surface height at the
                  start is found by testing the point against every triangle's plan-view
                  barycentric coordinates (again: no cell lookup).
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


def _nearest_hit(v0, e1, e2, o, d, t_lo, t_hi, eps=1e-12, tol=1e-9):
    h = np.cross(d, e2)
    a = np.einsum("ij,ij->i", e1, h)
    ok = np.abs(a) > eps
    f = np.where(ok, 1.0 / np.where(ok, a, 1.0), 0.0)
    s = o - v0
    u = f * np.einsum("ij,ij->i", s, h)
    q = np.cross(s, e1)
    v = f * np.einsum("ij,j->i", q, d)
    t = f * np.einsum("ij,ij->i", e2, q)
    hit = ok & (u >= -tol) & (v >= -tol) & (u + v <= 1 + tol) & (t >= t_lo) & (t <= t_hi)
    return t[hit].min() if hit.any() else None


def raycast(dem, origins, dirs, range_max, eps=1e-12):
    """Nearest Moller-Trumbore hit with t in [0, range_max]; misses -> range_max."""
    tris = dem_triangles(dem)
    v0, e1, e2 = tris[:, 0], tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
    out = np.full(len(origins), float(range_max))
    for r, (o, d) in enumerate(zip(np.asarray(origins, float), np.asarray(dirs, float))):
        t = _nearest_hit(v0, e1, e2, o, d, 0.0, range_max, eps)
        if t is not None:
            out[r] = t
    return out


def surface_height(tris, x, y, tol=1e-9):
    """Height of the triangulated surface at plan-view point (x, y): every triangle is tested for
    containment (barycentric, edges included); the highest containing facet's value is returned
    (facets that share an edge agree there).  None outside the footprint."""
    ax, ay = tris[:, 0, 0], tris[:, 0, 1]
    bx, by = tris[:, 1, 0], tris[:, 1, 1]
    cx, cy = tris[:, 2, 0], tris[:, 2, 1]
    det = (by - cy) * (ax - cx) + (cx - bx) * (ay - cy)
    l0 = ((by - cy) * (x - cx) + (cx - bx) * (y - cy)) / det
    l1 = ((cy - ay) * (x - cx) + (ax - cx) * (y - cy)) / det
    l2 = 1.0 - l0 - l1
    inside = (l0 >= -tol) & (l1 >= -tol) & (l2 >= -tol)
    if not inside.any():
        return None
    zz = l0 * tris[:, 0, 2] + l1 * tris[:, 1, 2] + l2 * tris[:, 2, 2]
    return float(zz[inside].max())


def raycast_spec(dem, origins, dirs, range_max):
    """The RAYCAST SPEC of ``c/auvpf_oracle.c`` by brute force: clip the ray to the footprint
    [x0, x0 + (W-1) res] x [y0, y0 + (H-1) res] and to t in [0, range_max] -> [ts, te]; an empty clip
    misses; f(ts) <= 0 returns ts; otherwise the nearest triangle hit in (ts, te]; none -> range_max."""
    tris = dem_triangles(dem)
    v0, e1, e2 = tris[:, 0], tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
    H, W = np.asarray(dem["z"]).shape
    lo = np.array([dem["x0"], dem["y0"]], float)
    hi = lo + np.array([W - 1, H - 1], float) * dem["res"]
    out = np.full(len(origins), float(range_max))
    for r, (o, d) in enumerate(zip(np.asarray(origins, float), np.asarray(dirs, float))):
        ts, te, empty = 0.0, float(range_max), False
        for a in range(2):
            if d[a] == 0.0:
                if o[a] < lo[a] or o[a] > hi[a]:
                    empty = True
            else:
                t1, t2 = (lo[a] - o[a]) / d[a], (hi[a] - o[a]) / d[a]
                ts, te = max(ts, min(t1, t2)), min(te, max(t1, t2))
        if empty or not (ts <= te):
            continue
        p = o + ts * d
        zs = surface_height(tris, min(max(p[0], lo[0]), hi[0]), min(max(p[1], lo[1]), hi[1]))
        if zs is not None and p[2] - zs <= 0.0:
            out[r] = ts
            continue
        t = _nearest_hit(v0, e1, e2, o, d, ts, te)
        if t is not None:
            out[r] = t
    return out
