"""Anchor of the raycast specification (CPU only).

The reference has no raycaster, so the DDA oracle that every GPU range is
compared with -- ``oracle/c/auvpf_oracle.c`` -- is itself checked here against an independent
formulation: Moller-Trumbore against EVERY triangle of the DEM (``oracle/raycast_bruteforce.py``),
no grid walk and no cell lookup.  Tolerance 1e-9 m on more than 1e3 rays per map, with misses,
border exits, rays that start outside the footprint, rays that start below the surface, axis- and
diagonal-aligned rays (degenerate line families of the DDA) and occluding ridges."""
import numpy as np
import pytest

from oracle import c_oracle, raycast_bruteforce as bf
from smarc_navigation_b200 import world

TOL = 1e-9


def _smooth():
    return world.make_dem(60, 2.0, seed=5)


def _ridged():
    d = world.make_dem(48, 1.0, seed=8)
    z = d["z"].astype(np.float64)
    ii, jj = np.meshgrid(np.arange(48), np.arange(48), indexing="ij")
    z += 6.0 * np.abs(np.sin(0.45 * ii + 0.2 * jj)) + 4.0 * ((ii // 5 + jj // 7) % 2)     # ridges and terraces
    d["z"] = z.astype(np.float32)
    return d


def _unit(v):
    v = np.asarray(v, float)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _rays(dem, rng, n):
    """A labelled mix of rays; every category the spec distinguishes is in it."""
    H, W = dem["z"].shape
    ext = np.array([(W - 1) * dem["res"], (H - 1) * dem["res"]])
    zmax, zmin = float(dem["z"].max()), float(dem["z"].min())
    cat = {}
    # 1. sensor-like: above the terrain, pointing down into a +-75 degree cone
    o = np.column_stack([rng.uniform(0.1, 0.9, n) * ext[0], rng.uniform(0.1, 0.9, n) * ext[1], zmax + rng.uniform(2, 25, n)])
    th, ph = np.radians(rng.uniform(0, 75, n)), rng.uniform(0, 2 * np.pi, n)
    cat["down"] = (o, np.column_stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), -np.cos(th)]))
    # 2. grazing / horizontal rays at terrain height: occlusions, near misses, border exits
    m = n // 2
    o = np.column_stack([rng.uniform(0.05, 0.95, m) * ext[0], rng.uniform(0.05, 0.95, m) * ext[1], rng.uniform(zmin, zmax + 3, m)])
    ph = rng.uniform(0, 2 * np.pi, m)
    el = np.radians(rng.uniform(-8, 8, m))
    cat["grazing"] = (o, np.column_stack([np.cos(el) * np.cos(ph), np.cos(el) * np.sin(ph), np.sin(el)]))
    # 3. upward rays from above the surface: misses
    m = n // 6
    o = np.column_stack([rng.uniform(0.2, 0.8, m) * ext[0], rng.uniform(0.2, 0.8, m) * ext[1], np.full(m, zmax + 1.0)])
    cat["up"] = (o, _unit(np.column_stack([rng.normal(size=m), rng.normal(size=m), rng.uniform(0.2, 1.0, m)])))
    # 4. starts below the surface (inside the footprint): range 0 whatever the direction
    m = n // 6
    o = np.column_stack([rng.uniform(0.1, 0.9, m) * ext[0], rng.uniform(0.1, 0.9, m) * ext[1], np.full(m, zmin - 3.0)])
    cat["below"] = (o, _unit(rng.normal(size=(m, 3))))
    # 5. starts outside the footprint, aimed at it (enters through a border; above or below the surface there)
    m = n // 4
    side = rng.integers(0, 4, m)
    o = np.column_stack([np.where(side == 0, -rng.uniform(1, 30, m), np.where(side == 1, ext[0] + rng.uniform(1, 30, m), rng.uniform(0, 1, m) * ext[0])),
                         np.where(side == 2, -rng.uniform(1, 30, m), np.where(side == 3, ext[1] + rng.uniform(1, 30, m), rng.uniform(0, 1, m) * ext[1])),
                         rng.uniform(zmin - 2, zmax + 15, m)])
    tgt = np.column_stack([rng.uniform(0.2, 0.8, m) * ext[0], rng.uniform(0.2, 0.8, m) * ext[1], rng.uniform(zmin - 5, zmax, m)])
    cat["outside_in"] = (o, _unit(tgt - o))
    # 6. outside and pointing away / parallel to the border: empty clip
    m = n // 10
    o = np.column_stack([-rng.uniform(1, 20, m), rng.uniform(0, 1, m) * ext[1], np.full(m, zmax + 5.0)])
    cat["outside_away"] = (o, _unit(np.column_stack([-rng.uniform(0, 1, m), rng.normal(size=m), -rng.uniform(0, 1, m)])))
    # 7. degenerate line families: dx = 0, dy = 0, dx = dy, dx = -dy, and straight down
    m = n // 10
    base = np.column_stack([rng.uniform(0.2, 0.8, 5 * m) * ext[0], rng.uniform(0.2, 0.8, 5 * m) * ext[1], zmax + rng.uniform(1, 10, 5 * m)])
    dz = -rng.uniform(0.05, 1.0, 5 * m)
    s = rng.choice([-1.0, 1.0], 5 * m)
    kinds = np.repeat(np.arange(5), m)
    dxy = np.select([kinds[:, None] == 0, kinds[:, None] == 1, kinds[:, None] == 2, kinds[:, None] == 3],
                    [np.column_stack([np.zeros(5 * m), s]), np.column_stack([s, np.zeros(5 * m)]),
                     np.column_stack([s, s]), np.column_stack([s, -s])], np.zeros((5 * m, 2)))
    cat["aligned"] = (base, _unit(np.column_stack([dxy, dz])))
    # 8. origins exactly on grid lines / vertices (tie-breaking of the cell lookup)
    m = n // 10
    o = np.column_stack([rng.integers(2, W - 2, m) * dem["res"], rng.integers(2, H - 2, m) * dem["res"], zmax + rng.uniform(1, 10, m)])
    cat["on_vertex"] = (o, _unit(np.column_stack([rng.normal(size=m), rng.normal(size=m), -rng.uniform(0.2, 1, m)])))
    return cat


@pytest.mark.parametrize("make,range_max", [(_smooth, 120.0), (_smooth, 25.0), (_ridged, 60.0)])
def test_c_oracle_matches_bruteforce(make, range_max):
    dem = make()
    rng = np.random.default_rng(12)
    cats = _rays(dem, rng, 600)
    n_total, n_miss, n_zero = 0, 0, 0
    for name, (o, d) in cats.items():
        got = c_oracle.raycast(dem, o, d, range_max)
        want = bf.raycast_spec(dem, o, d, range_max)
        err = np.abs(got - want)
        assert err.max() <= TOL, (name, int(err.argmax()), float(err.max()), o[err.argmax()], d[err.argmax()])
        n_total += len(o)
        n_miss += int((want >= range_max).sum())
        n_zero += int((want == 0.0).sum())
        if name in ("up", "outside_away"):
            assert (want >= range_max).all()
        if name == "below":
            assert (want == 0.0).all()
    assert n_total >= 1000 and n_miss >= 100 and n_zero >= 50


def test_plain_geometry_agrees_where_the_clip_rule_is_idle():
    """For rays that start inside the footprint above the surface, the spec IS the nearest triangle
    hit: the plain Moller-Trumbore answer (no clip logic at all) must agree too, except for rays that
    leave the footprint before range_max (those the spec ends at the border by definition)."""
    dem = _ridged()
    rng = np.random.default_rng(3)
    o, d = _rays(dem, rng, 900)["down"]
    got = c_oracle.raycast(dem, o, d, 60.0)
    plain = bf.raycast(dem, o, d, 60.0)
    np.testing.assert_allclose(got, plain, rtol=0, atol=TOL)
    assert (got < 60.0).mean() > 0.5          # the rest leave the 47 m map: ended at the border by both


def test_update_mbes_ranges_are_the_same_rays():
    """orc_update_mbes (what the GPU parity tests call) casts exactly the rays get_p_pose defines."""
    from oracle import pf_oracle as po
    dem = _smooth()
    rng = np.random.default_rng(4)
    poses = np.column_stack([rng.uniform(30, 90, 8), rng.uniform(30, 90, 8), np.full(8, -5.0),
                             rng.normal(0, 0.1, 8), rng.normal(0, 0.1, 8), rng.uniform(-np.pi, np.pi, 8)])
    beams = world.make_beams(32)
    out = c_oracle.update_mbes(dem, poses, np.eye(4), beams, None, 1.0, 0.5, 120.0, want_ranges=True)["ranges"]
    p, R = po.get_p_pose(poses, np.eye(4))
    o = np.repeat(p, 32, axis=0)
    d = np.einsum("nij,bj->nbi", R, beams).reshape(-1, 3)
    want = bf.raycast_spec(dem, o, d, 120.0).reshape(8, 32)
    np.testing.assert_allclose(out, want, rtol=0, atol=TOL)
