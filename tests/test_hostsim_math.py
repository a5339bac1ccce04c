"""Kernel ARITHMETIC vs the oracle, on the CPU: the __host__ __device__ functions of
smarc_navigation_b200/csrc/auvpf_math.cuh compiled for the host (tests/hostsim).
This is how kernel logic is checked in a container without a GPU; the GPU parity
tests (test_gpu_*.py) repeat the same comparisons through libauvpf.so."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import pf_oracle as o
from oracle import tf_stub
from smarc_navigation_b200 import world
from tests import fragile
from tests.hostsim import binding as hs

RANGE_TOL = 1e-4      # north_star: beam ranges within 1e-4 m
FRAGILE = 2e-3        # rays whose first-hit decision hinges on < 2 mm are excluded (and counted)


def _cloud(rng, n, centre, spread, roll_pitch=0.03, z=-5.0):
    poses = np.zeros((n, 6))
    poses[:, 0] = centre[0] + rng.normal(0, spread, n)
    poses[:, 1] = centre[1] + rng.normal(0, spread, n)
    poses[:, 2] = z
    poses[:, 3] = rng.normal(0, roll_pitch, n)
    poses[:, 4] = rng.normal(0, roll_pitch, n)
    poses[:, 5] = rng.uniform(-np.pi, np.pi, n)
    poses[:, 2:6] = poses[:, 2:6].astype(np.float32)     # the device keeps z, roll, pitch, yaw in fp32
    return poses


def _compare(dem, shift, poses, beams, m2o, range_max, max_fragile_frac=2e-3, shift_hi=0, fan=1):
    """Sturdy rays within RANGE_TOL of the oracle; fragile rays (decision hinges on < 2 mm) inside the
    bracket of the oracle's answers for a sensor moved 5 mm down / up (tests/fragile.py)."""
    beams = np.asarray(beams, np.float32).astype(np.float64)      # identical (fp32) directions for both arms
    r = hs.raycast_poses(dem, shift, range_max, m2o, poses, beams.astype(np.float32), shift_hi=shift_hi, fan=fan)
    out = c_oracle.update_mbes(dem, poses, m2o, beams, None, 1.0, 0.5, range_max, want_ranges=True, want_diag=True)
    ro, stats = fragile.check_ranges(r, dem, poses, m2o, beams, range_max, max_fragile_frac, label="hostsim", out=out)
    assert np.all((r >= 0) & (r <= np.float32(range_max)))
    return ro, np.abs(r - ro), fragile.fragile_mask(out["diag"])


@pytest.mark.parametrize("n,res,shift,hi,B", [(500, 2.0, 1, 0, 256), (500, 2.0, 2, 0, 256), (500, 2.0, 4, 0, 64),
                                              (1200, 0.5, 3, 0, 512), (1200, 0.5, 5, 0, 128),
                                              (500, 2.0, 1, 3, 256), (1200, 0.5, 2, 5, 512), (1200, 0.5, 1, 2, 128)])
def test_raycast_fan_over_synthetic_bathymetry(n, res, shift, hi, B):
    dem = world.make_dem(n, res)
    rng = np.random.default_rng(4)
    poses = _cloud(rng, 96, (0.45 * n * res,) * 2, 3.0)
    ro, err, fragile = _compare(dem, shift, poses, world.make_beams(B), np.eye(4), world.RANGE_MAX, shift_hi=hi)
    assert (ro < world.RANGE_MAX).all()          # every beam hits the seafloor here


def test_raycast_far_from_origin_keeps_tolerance():
    """World coordinates ~1.9 km: fp32 world coordinates would already be off by 1.2e-4 m."""
    n, res = 4000, 0.5
    dem = world.make_dem(n, res)
    rng = np.random.default_rng(5)
    poses = _cloud(rng, 48, (1900.0, 1850.0), 2.0)
    _compare(dem, 3, poses, world.make_beams(512), np.eye(4), world.RANGE_MAX)


def test_raycast_with_map_to_odom_transform():
    dem = world.make_dem(500, 2.0)
    rng = np.random.default_rng(6)
    q = tf_stub.quaternion_from_euler(0.01, -0.02, 0.7)
    m2o = tf_stub.translation_matrix([400.0, 300.0, 1.5]).dot(tf_stub.quaternion_matrix(q))
    poses = _cloud(rng, 64, (50.0, 80.0), 5.0)
    _compare(dem, 1, poses, world.make_beams(256), m2o, world.RANGE_MAX)


@pytest.mark.parametrize("lo,hi", [(2, 0), (1, 4), (2, 3)])
def test_raycast_edges_misses_and_steep_rays(lo, hi):
    """Particles at / outside the map border, short range_max (misses), large roll
    (beams up to ~horizontal), upward-looking beams."""
    dem = world.make_dem(300, 2.0, seed=11)
    rng = np.random.default_rng(7)
    beams = world.make_beams(128, half_angle_deg=75.0)
    kw = dict(shift_hi=hi)
    edge = _cloud(rng, 64, (8.0, 590.0), 6.0)                       # straddles two borders
    ro, err, fr = _compare(dem, lo, edge, beams, np.eye(4), world.RANGE_MAX, max_fragile_frac=2e-2, **kw)
    assert (ro >= world.RANGE_MAX).any() and (ro < world.RANGE_MAX).any()
    outside = _cloud(rng, 32, (-40.0, 300.0), 3.0)
    _compare(dem, lo, outside, beams, np.eye(4), world.RANGE_MAX, max_fragile_frac=2e-2, **kw)
    tilted = _cloud(rng, 64, (300.0, 300.0), 3.0, roll_pitch=0.5)
    _compare(dem, lo, tilted, beams, np.eye(4), world.RANGE_MAX, max_fragile_frac=2e-2, **kw)
    short = _cloud(rng, 32, (300.0, 300.0), 3.0)
    ro, err, fr = _compare(dem, lo, short, beams, np.eye(4), 30.0, max_fragile_frac=2e-2, **kw)
    assert (ro == 30.0).any()
    up = _cloud(rng, 16, (300.0, 300.0), 3.0)
    up[:, 3] = np.float32(np.pi)                                    # sonar looking up: all miss
    ro, err, fr = _compare(dem, lo, up, beams, np.eye(4), world.RANGE_MAX, **kw)
    assert (ro == world.RANGE_MAX).all()
    below = _cloud(rng, 16, (300.0, 300.0), 3.0, z=-60.0)           # vehicle under the seafloor: range 0
    ro, err, fr = _compare(dem, lo, below, beams, np.eye(4), world.RANGE_MAX, **kw)
    assert (ro == 0.0).all()


def test_raycast_rough_terrain_random_directions():
    dem = world.make_dem(200, 1.0, seed=3)
    dem["z"] = (dem["z"] + np.random.default_rng(1).normal(0, 0.6, dem["z"].shape)).astype(np.float32)
    rng = np.random.default_rng(8)
    d = rng.normal(size=(96, 3))
    d[:, 2] = -np.abs(d[:, 2])
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    poses = _cloud(rng, 48, (100.0, 100.0), 20.0, roll_pitch=0.2, z=-12.0)
    _compare(dem, 2, poses, d, np.eye(4), 80.0, max_fragile_frac=3e-2)
    _compare(dem, 1, poses, d, np.eye(4), 80.0, max_fragile_frac=3e-2, shift_hi=3)


def test_mip_is_conservative():
    dem = world.make_dem(130, 2.0, seed=2)
    z = dem["z"]
    for shift in (1, 2, 3, 5):
        _, mip = hs.raycast_poses(dem, shift, 50.0, np.eye(4), np.zeros((1, 6)), world.make_beams(4).astype(np.float32),
                                  want_mip=True)
        S = 1 << shift
        for bi in range(mip.shape[0]):
            for bj in range(mip.shape[1]):
                blk = z[bi * S:min((bi + 1) * S, z.shape[0] - 1) + 1, bj * S:min((bj + 1) * S, z.shape[1] - 1) + 1]
                assert mip[bi, bj] == blk.max()


def test_propagate_matches_reference_golden(golden):
    """G1: fp32/fp64 SoA propagate vs Particle.motion_pred golden, injected normals."""
    g = golden("particle")
    for case in ("a", "b"):
        poses = g[case + "_poses_start"].copy()
        cov = g[case + "_process_cov"]
        for odom, noise, want in zip(g[case + "_odoms"], g[case + "_pred_noise"], g[case + "_pred_poses"]):
            roll, pitch = o.odom_roll_pitch(odom[4:8])
            dt = odom[9]
            vx, vy, vz = odom[0:3] * dt
            cr, sr, cp, sp = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch)
            params = (cp * vx + sp * sr * vy + sp * cr * vz, cr * vy - sr * vz, odom[3] * dt, roll, pitch, odom[8],
                      np.sqrt(cov[0]), np.sqrt(cov[1]), np.sqrt(cov[5]))
            poses = hs.propagate(poses, params, noise)
            np.testing.assert_allclose(poses[:, 0:2], want[:, 0:2], atol=1e-4)
            np.testing.assert_allclose(poses[:, 2:5], want[:, 2:5], atol=1e-6)
            dyaw = (poses[:, 5] - want[:, 5] + np.pi) % (2 * np.pi) - np.pi
            assert np.abs(dyaw).max() < 1e-5


def test_propagate_200_steps_drift():
    """200 predicts at map coordinates ~1 km stay within 1e-4 m / 1e-5 rad of float64."""
    rng = np.random.default_rng(12)
    N = 256
    poses = np.zeros((N, 6))
    poses[:, 0] = 1000 + rng.normal(0, 2, N)
    poses[:, 1] = 900 + rng.normal(0, 2, N)
    poses[:, 5] = rng.uniform(-3, 3, N)
    ref = poses.copy()
    sim = poses.copy()
    sim[:, 5] = sim[:, 5].astype(np.float32)
    ref[:, 5] = sim[:, 5]
    cov = np.array([1e-4, 1e-4, 0, 0, 0, 1e-6])
    traj = world.Trajectory(2000.0)
    for s in range(200):
        od = traj.step()
        noise = rng.normal(size=(N, 6)).astype(np.float32)
        ref = o.motion_pred(ref, od.v, od.wz, od.quat, od.z, od.dt, cov, noise.astype(np.float64))
        roll, pitch = o.odom_roll_pitch(od.quat)
        vx, vy, vz = od.v * od.dt
        cr, sr, cp, sp = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch)
        params = (cp * vx + sp * sr * vy + sp * cr * vz, cr * vy - sr * vz, od.wz * od.dt, roll, pitch, od.z,
                  np.sqrt(cov[0]), np.sqrt(cov[1]), np.sqrt(cov[5]))
        sim = hs.propagate(sim, params, noise)
    assert np.abs(sim[:, 0:2] - ref[:, 0:2]).max() < 1e-4
    dyaw = (sim[:, 5] - ref[:, 5] + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(dyaw).max() < 1e-5


def test_wrap_pi():
    for a in (0.0, 3.0, -3.0, 3.2, -3.2, 6.0, -6.0, 9.5, -9.5, 40.0, -40.0, np.pi, -np.pi):
        w = hs.wrap_pi(a)
        ref = (np.float64(np.float32(a)) + np.pi) % (2 * np.pi) - np.pi
        d = (w - ref + np.pi) % (2 * np.pi) - np.pi
        assert abs(d) < 4e-6 and -np.pi - 1e-6 <= w <= np.pi + 1e-6


def test_philox_and_normals():
    rng = np.random.default_rng(3)
    ctr = rng.integers(0, 2 ** 32, (64, 4), dtype=np.uint64).astype(np.uint32)
    key = rng.integers(0, 2 ** 32, (64, 2), dtype=np.uint64).astype(np.uint32)
    np.testing.assert_array_equal(hs.philox(ctr, key), o.philox4x32_10(ctr, key))
    ids = np.array([0, 1, 2, 12345, 2 ** 33 + 7], np.int64)
    got = hs.normals6(0xDEADBEEF12345678, ids, 9, 3)
    want = o.philox_pose_normals(0xDEADBEEF12345678, 3, 9, ids)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)
    big = hs.normals6(42, np.arange(200000, dtype=np.int64), 0, 1)
    assert abs(big.mean()) < 5e-3 and abs(big.std() - 1) < 5e-3
    assert abs((big ** 3).mean()) < 2e-2 and abs((big ** 4).mean() - 3) < 5e-2


@pytest.mark.parametrize("N,scale", [(1, 1.0), (2, 1.0), (100, 3.0), (4097, 10.0), (65536, 25.0)])
def test_fixed_point_resample_is_bit_exact(N, scale):
    """G3: device normalise + select arithmetic == oracle quantiser == reference systematic."""
    rng = np.random.default_rng(N)
    logw = (rng.normal(0, scale, N) - 470.0).astype(np.float32)
    if N > 10:
        logw[3] = -np.inf
        logw[5] = np.nan
    u = float(rng.uniform())
    qq, anc = hs.resample(logw, u)
    q_ref, m, A = c_oracle.quantise_logw(logw)
    np.testing.assert_array_equal(qq, q_ref)
    if N <= 5000:
        np.testing.assert_array_equal(qq, o.quantise_logw(logw)[0])
    w = o.weights_from_q(qq)
    np.testing.assert_array_equal(anc, o.systematic_resample(w, u))
    np.testing.assert_array_equal(anc, c_oracle.systematic(w, u))


def test_cdf_exact_is_the_exact_floor():
    """floor(C 2^52 / A) for prefix sums up to 2^90: float64 estimate + WRAPPING 128-bit multiply-back (the
    products exceed 128 bits; only their difference has to fit) == Python big-integer floor division, and
    the estimate is within its documented error of it."""
    rng = np.random.default_rng(11)
    cases = [(1, 1), (5, 5), (1, (1 << 84) + 12345), ((1 << 60), (1 << 60)), ((1 << 60), (1 << 84)),
             ((1 << 60) - 1, (1 << 60) + 1), (3, 7), ((1 << 59) + 1, 3 * (1 << 59) + 2), (0, 17),
             ((1 << 90) - 1, (1 << 90) - 1), ((1 << 90) - 2, (1 << 90) - 1), ((1 << 89) + 1, (1 << 90) - 1),
             ((1 << 64) - 1, (1 << 64)), ((1 << 64), (1 << 64) + 1), ((1 << 64) + 1, (1 << 65))]
    for _ in range(6000):
        abits = int(rng.integers(1, 91))
        A = int(rng.integers(1, 1 << min(abits, 62))) << max(0, abits - 62) | int(rng.integers(0, 1 << 20))
        A = max(A, 1)
        kind = rng.integers(0, 4)
        if kind == 0:
            Cv = int(rng.integers(0, 1 << 62)) % (A + 1)
        elif kind == 1:
            Cv = A - int(rng.integers(0, min(A, 1000)))
        elif kind == 2:                                  # right at a quotient boundary: C = ceil(k A / 2^52) +- 1
            k = int(rng.integers(0, 1 << 52))
            Cv = min(A, max(0, -((-k * A) >> 52) + int(rng.integers(-1, 2))))
        else:
            Cv = (A * int(rng.integers(0, 1 << 30))) >> 30
        cases.append((Cv, A))
    for Cv, A in cases:
        want = (Cv << 52) // A
        assert hs.cdf_exact(Cv, A) == want, (Cv, A)
        assert abs(hs.cdf_estimate(Cv, A) - want) <= 3.0 + 1.0, (Cv, A)


def test_hinted_searchsorted_is_numpys():
    """np.searchsorted on NON-monotone arrays depends on its key-to-key hint; the device walk
    must take the same steps (resampling.py:74 calls it on such an array)."""
    rng = np.random.default_rng(12)
    for n in (1, 2, 17, 1000):
        for _ in range(20):
            arr = np.cumsum(rng.normal(0.0, 1.0, n))          # not sorted
            arr[-1] = 1.0
            keys = rng.uniform(-0.5, 1.0, 200)
            np.testing.assert_array_equal(hs.searchsorted(arr, keys), np.searchsorted(arr, keys))
            ks = np.sort(keys)
            np.testing.assert_array_equal(hs.searchsorted(arr, ks), np.searchsorted(arr, ks))


def test_residual_resample_strict_and_fixed(golden):
    g = golden("residual_quantised")
    for n in ("n100", "n1000", "n4096", "n20000"):
        qq, U = g[n + "_qq"], g[n + "_U"]
        np.testing.assert_array_equal(hs.residual(qq, U, strict=True), g[n + "_idx"])     # the unmodified reference
        np.testing.assert_array_equal(hs.residual(qq, U, strict=False), o.residual_resample_fixed(qq, U))


def test_count_below_matches_positions():
    """sys_count_below (closed form, fixed up with the reference's expression only when the estimate is
    within 1e-5 of an integer) == counting positions[i] < cdf, including CDF values that sit exactly on,
    one unit below and one unit above a position (the estimate is then an integer up to rounding)."""
    rng = np.random.default_rng(1)
    for N in (1, 7, 1000, 1 << 20, (1 << 24) + 3):
        for u in (float(rng.uniform()), 0.0, 1.0 - 2.0 ** -53, 2.0 ** -30):
            pos = (u + np.arange(N)) / float(N)
            cs = [0, 1, 1 << 51, (1 << 52) - 1, 1 << 52] + [int(x) for x in rng.integers(0, 1 << 52, 20)]
            for i in [0, N - 1] + [int(x) for x in rng.integers(0, N, 12)]:
                k = int(np.floor(np.ldexp(pos[i], 52)))
                cs += [max(k - 1, 0), k, min(k + 1, 1 << 52)]
            for c in cs:
                want = int(np.searchsorted(pos, np.ldexp(float(c), -52), side="left"))     # pos is non-decreasing
                assert hs.count_below(u, N, c) == want, (N, u, c)


def test_count_below_beyond_array_sizes():
    """N near 2^31 (no array of positions): the count must satisfy its definition at the boundary."""
    rng = np.random.default_rng(2)
    N = (1 << 31) - 5
    for u in (float(rng.uniform()), 0.0, 0.999999):
        position = lambda i: (u + float(i)) / float(N)
        for t in range(40):
            i = int(rng.integers(0, N))
            k = int(np.floor(np.ldexp(position(i), 52)))
            for c in (k - 1, k, k + 1, int(rng.integers(0, 1 << 52))):
                c = min(max(c, 0), 1 << 52)
                cv = np.ldexp(float(c), -52)
                g = hs.count_below(u, N, c)
                assert 0 <= g <= N
                assert g == 0 or position(g - 1) < cv
                assert g == N or not (position(g) < cv)


def _ridged_dem(n=400, res=0.5, seed=5):
    """Steep short-wavelength ridges: outer beams are occluded (first-hit jumps)."""
    rng = np.random.default_rng(seed)
    x = np.arange(n) * res
    X, Y = np.meshgrid(x, x)
    z = -30.0 + 2.5 * np.sin(2 * np.pi * X / 9.0) * np.cos(2 * np.pi * Y / 13.0) + 1.0 * np.sin(2 * np.pi * (X + Y) / 4.0)
    z += rng.normal(0, 0.05, z.shape)
    return {"z": z.astype(np.float32), "x0": 0.0, "y0": 0.0, "res": res}


@pytest.mark.parametrize("lo,hi", [(2, 0), (2, 5)])
def test_fan_sweep_exact_under_occlusion(lo, hi):
    """The fan sweep (certified start from the inner neighbour's clear extent) must return
    the same first hits as independent rays, also where ridges occlude the outer beams."""
    dem = _ridged_dem()
    rng = np.random.default_rng(9)
    beams = world.make_beams(512, half_angle_deg=70.0)
    poses = _cloud(rng, 48, (100.0, 100.0), 8.0, roll_pitch=0.04, z=-21.0)   # steep scene: sweep valid for tilt < ~14 deg
    rmax = 60.0
    r_fan, nsw = hs.raycast_poses(dem, lo, rmax, np.eye(4), poses, beams.astype(np.float32), shift_hi=hi, want_swept=True)
    r_ind = hs.raycast_poses(dem, lo, rmax, np.eye(4), poses, beams.astype(np.float32), shift_hi=hi, fan=False)
    assert nsw > 0.8 * r_fan.size                                     # the sweep is really in use
    # occlusion jumps are present in this scene
    assert (np.abs(np.diff(r_ind, axis=1)) > 2.0).sum() > 50
    np.testing.assert_allclose(r_fan, r_ind, rtol=0, atol=3e-5)         # same first hits (fp32 rounding only)
    ro, err, fragile = _compare(dem, lo, poses, beams, np.eye(4), rmax, max_fragile_frac=5e-2, shift_hi=hi)


def test_fan_plan():
    """Host-side fan analysis: planarity / ordering detection and per-lane sweep order."""
    import ctypes as C
    for B in (512, 256, 100, 33):
        beams = world.make_beams(B).astype(np.float32)
        r, nsw = hs.raycast_poses(world.make_dem(100, 2.0), 1, 60.0, np.eye(4), np.array([[100.0, 100.0, -5, 0, 0, 0.3]]),
                                  beams, want_swept=True)
        chunk = (B + 31) // 32
        assert nsw >= B - 2 * 32 - 4 if chunk > 1 else nsw == 0
    # non-planar or unsorted beam sets are never swept
    rng = np.random.default_rng(2)
    d = rng.normal(size=(128, 3))
    d[:, 2] = -np.abs(d[:, 2]) - 0.5
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r, nsw = hs.raycast_poses(world.make_dem(100, 2.0), 1, 60.0, np.eye(4), np.array([[100.0, 100.0, -5, 0, 0, 0.3]]),
                              d.astype(np.float32), want_swept=True)
    assert nsw == 0
    shuffled = world.make_beams(128)[rng.permutation(128)].astype(np.float32)
    r, nsw = hs.raycast_poses(world.make_dem(100, 2.0), 1, 60.0, np.eye(4), np.array([[100.0, 100.0, -5, 0, 0, 0.3]]),
                              shuffled, want_swept=True)
    assert nsw == 0


@pytest.mark.parametrize("name", ["C2", "C3", "tilted", "ridged", "ridged_tilted", "border", "short_range"])
def test_profile_merge_vs_oracle(name):
    """Profile merge (one march per fan side, O(cells + beams)) returns the oracle's first hits."""
    rng = np.random.default_rng(4)
    rmax, lo, hi, frag = 120.0, 2, 5, 5e-3
    if name == "C2":
        dem, poses, beams, lo, hi = world.make_dem(500, 2.0), _cloud(rng, 96, (450.0, 450.0), 3.0), world.make_beams(256), 1, 0
    elif name == "C3":
        dem, poses, beams = world.make_dem(1200, 0.5), _cloud(rng, 96, (300.0, 300.0), 3.0), world.make_beams(512)
    elif name == "tilted":
        dem, poses, beams = world.make_dem(1200, 0.5), _cloud(rng, 96, (300.0, 300.0), 3.0, roll_pitch=0.3), world.make_beams(512, 70.0)
    elif name == "ridged":
        dem, poses, beams, rmax, frag = _ridged_dem(), _cloud(rng, 48, (100.0, 100.0), 8.0, roll_pitch=0.04, z=-21.0), \
            world.make_beams(512, 70.0), 60.0, 5e-2
    elif name == "ridged_tilted":
        dem, poses, beams, rmax, frag = _ridged_dem(), _cloud(rng, 48, (100.0, 100.0), 8.0, roll_pitch=0.2, z=-21.0), \
            world.make_beams(512, 70.0), 60.0, 5e-2
    elif name == "border":          # reach leaves the map: the march finishes the exiting side ray by ray; sensors outside fall back
        dem, poses, beams, frag = world.make_dem(300, 2.0, seed=11), _cloud(rng, 32, (8.0, 590.0), 6.0), world.make_beams(128, 75.0), 2e-2
    else:                           # short range_max: outer beams miss
        dem, poses, beams, rmax = world.make_dem(500, 2.0), _cloud(rng, 48, (450.0, 450.0), 3.0), world.make_beams(256), 30.0
    r, n_prof = hs.raycast_poses(dem, lo, rmax, np.eye(4), poses, beams.astype(np.float32), shift_hi=hi, fan=2, want_swept=True)
    ro, err, fragile = _compare(dem, lo, poses, beams, np.eye(4), rmax, max_fragile_frac=frag, shift_hi=hi, fan=2)
    if name == "border":
        inside = (poses[:, 0] > 0.0).sum()
        assert n_prof == inside * r.shape[1]          # every sensor that stands over the map is certified by the march
    elif name == "tilted":
        assert n_prof >= 0.75 * r.size          # steep triangle x large tilt: some particles fall back
    elif name == "ridged_tilted":
        # 60-degree ridge faces: only sensors tilted less than ~14 degrees see a single-valued profile in their fan
        # plane (the certificate of profile_particle); the others are recast ray by ray
        assert 0.25 * r.size <= n_prof < 0.9 * r.size
    else:
        assert n_prof == r.size                 # every particle certified by the profile march
    if name == "short_range":
        assert (ro == rmax).any() and (ro < rmax).any()


@pytest.mark.parametrize("terrain", ["smooth", "ridged"])
def test_profile_merge_near_every_border_and_corner(terrain):
    """Sensors closer to the map border than their reach: the profile curve runs off the map, the march
    finishes that side ray by ray (finish_side_by_rays) with the rays clipped to the footprint -- hits
    right up to the border, misses beyond it, on all four borders and in the corners, any heading."""
    rng = np.random.default_rng(31)
    if terrain == "smooth":
        dem, rmax, z, frag = world.make_dem(400, 1.0, seed=3), 120.0, -5.0, 2e-2
    else:
        dem, rmax, z, frag = _ridged_dem(n=400, res=0.5), 60.0, -21.0, 5e-2
    ext = (dem["z"].shape[0] - 1) * dem["res"]
    spots = [(0.04, 0.5), (0.96, 0.5), (0.5, 0.03), (0.5, 0.97), (0.05, 0.06), (0.95, 0.04), (0.04, 0.95), (0.97, 0.96),
             (0.2, 0.5), (0.5, 0.8)]
    poses = np.concatenate([_cloud(rng, 20, (a * ext, b * ext), 0.01 * ext, roll_pitch=0.05, z=z) for a, b in spots])
    poses = poses[(poses[:, 0] > 0) & (poses[:, 0] < ext) & (poses[:, 1] > 0) & (poses[:, 1] < ext)]
    beams = world.make_beams(256, 70.0)
    r, n_prof = hs.raycast_poses(dem, 2, rmax, np.eye(4), poses, beams.astype(np.float32), shift_hi=5, fan=2, want_swept=True)
    ro, err, fragile_ = _compare(dem, 2, poses, beams, np.eye(4), rmax, max_fragile_frac=frag, shift_hi=5, fan=2)
    assert n_prof >= (0.95 if terrain == "smooth" else 0.7) * r.size      # the march, not the per-ray fallback
    assert (ro >= rmax).mean() > 0.1 and (ro < rmax).mean() > 0.3       # misses beyond the border and hits before it


def test_profile_merge_is_only_certified_where_the_profile_is_single_valued():
    """Regression: a sensor pitched 23 degrees over 60-degree ridge faces sees crests as overhangs in its fan plane
    -- the plane meets the surface in more than one piece, and a beam's first hit (29.35 m) lay on a piece the
    marched curve does not contain (it passes there at 35 m).  With range_max = 35 m the march stopped before it
    met the facet whose gamma would have failed it, and certified two wrong misses.  The certificate is now the
    largest facet gradient within reach against the tilt; such sensors are recast ray by ray."""
    rng = np.random.default_rng(123)
    dem = _ridged_dem(n=700, res=0.5)
    ext = (dem["z"].shape[0] - 1) * dem["res"]
    beams = world.make_beams(256, 65.0)
    poses = _cloud(rng, 600, (0.5 * ext, 0.5 * ext), 6.0, roll_pitch=0.20)
    poses[0] = [169.579322, 174.228105, -5.0, 0.0849569291, 0.406004101, -2.71862006]      # the sensor it was found with
    poses[1:40] = poses[0] + rng.normal(0, 1.0, (39, 6)) * np.array([0.3, 0.3, 0.0, 0.01, 0.01, 0.02])
    for rmax in (35.0, 60.0):
        out = c_oracle.update_mbes(dem, poses, np.eye(4), beams.astype(np.float32).astype(np.float64), None, 1.0, 0.5, rmax,
                                   want_ranges=True, want_diag=True)
        r, n_prof = hs.raycast_poses(dem, 2, rmax, np.eye(4), poses, beams.astype(np.float32), shift_hi=5, fan=2, want_swept=True)
        ro, dg = out["ranges"], out["diag"]
        sturdy = ~((dg[..., 0] < FRAGILE) | (dg[..., 1] < FRAGILE) | (dg[..., 2] < FRAGILE))
        assert sturdy.mean() > 0.95
        assert np.abs(r - ro)[sturdy].max() <= RANGE_TOL
        assert 0.2 * r.size < n_prof < 0.95 * r.size          # the march still takes the sensors it can certify
        assert (np.abs(np.diff(ro, axis=1)) > 2.0).sum() > 100   # occlusion jumps are what this scene is made of


@pytest.mark.parametrize("seed", [2, 3, 7])
def test_randomised_scenes_stay_within_the_range_tolerance(seed):
    """Adversarial sweep over what the update can meet: smooth / ridged / terraced / spiked terrain at 0.5-2 m cells,
    sensors 8-25 m above it tilted by up to ~40 degrees, 64-512 beams over 45-70 degrees, range_max 20-120 m.
    Whatever path takes a sensor (march, border variant, ray by ray), every ray that is not fragile keeps the
    1e-4 m tolerance.  Round 2 found two defects with it: the missing single-valued-profile certificate of the
    march (ranges metres off) and float32 solves of GRAZING per-ray hits (up to 4e-4 m: now re-solved in float64 on
    the facet's plane with the ray's float64 residuals, precise_ray_hit)."""
    rng = np.random.default_rng(seed)

    def terrain(kind, n, res, tseed):
        if kind == "smooth":
            return world.make_dem(n, res, seed=tseed)
        if kind == "ridged":
            return _ridged_dem(n=n, res=res, seed=tseed)
        d = world.make_dem(n, res, seed=tseed)
        z = d["z"].astype(np.float64)
        x = np.arange(n) * res
        X, Y = np.meshgrid(x, x)
        if kind == "steps":
            z = z + np.floor((X + 0.3 * Y) / 11.0) % 3 * 1.5
        else:
            z = z + (rng.uniform(size=z.shape) < 0.002) * rng.uniform(1, 4, z.shape)
        d["z"] = z.astype(np.float32)
        return d

    worst, rays = 0.0, 0
    for trial in range(60):
        kind = rng.choice(["smooth", "ridged", "steps", "spikes"])
        res, n = float(rng.choice([0.5, 1.0, 2.0])), int(rng.choice([300, 500]))
        dem = terrain(kind, n, res, int(rng.integers(1, 100)))
        ext = (n - 1) * res
        beams = world.make_beams(int(rng.choice([64, 256, 512])), float(rng.choice([45, 60, 70]))).astype(np.float32)
        centre = (rng.uniform(0.1, 0.9) * ext, rng.uniform(0.1, 0.9) * ext)
        poses = _cloud(rng, 48, centre, float(rng.choice([0.5, 5.0, 30.0])), roll_pitch=float(rng.choice([0.02, 0.1, 0.25])),
                       z=float(rng.choice([-5.0, -15.0, -22.0])))
        poses = poses[(poses[:, 0] > 0) & (poses[:, 0] < ext) & (poses[:, 1] > 0) & (poses[:, 1] < ext)]
        if not len(poses):
            continue
        rmax = float(rng.choice([20.0, 35.0, 60.0, 120.0]))
        out = c_oracle.update_mbes(dem, poses, np.eye(4), beams.astype(np.float64), None, 1.0, 0.5, rmax, want_ranges=True,
                                   want_diag=True)
        r = hs.raycast_poses(dem, 1, rmax, np.eye(4), poses, beams, shift_hi=4, fan=2)
        ro, dg = out["ranges"], out["diag"]
        sturdy = ~((dg[..., 0] < FRAGILE) | (dg[..., 1] < FRAGILE) | (dg[..., 2] < FRAGILE))
        err = np.abs(r - ro) * sturdy
        assert err.max() <= RANGE_TOL, (trial, kind, res, rmax, float(err.max()))
        worst, rays = max(worst, float(err.max())), rays + int(sturdy.sum())
    assert rays > 500000


def test_profile_merge_random_fans():
    """Random planar fans -- plane rotated about the vertical, asymmetric and one-sided angle ranges,
    1 to 200 beams, either beam order, short and long range_max, small and large roll/pitch -- keep
    the oracle's first hits within the range tolerance (per-ray fallback where the march declines)."""
    rng = np.random.default_rng(123)
    dem = world.make_dem(600, 1.0)
    for trial in range(40):
        B = int(rng.choice([1, 2, 3, 5, 17, 64, 200]))
        a0 = rng.uniform(-70, 10)
        a1 = a0 + rng.uniform(5, 60) if B > 1 else a0
        th = np.radians(np.linspace(a0, a1, B))
        psi = rng.uniform(0, 2 * np.pi)
        lat = np.array([np.cos(psi), np.sin(psi), 0.0])
        beams = np.sin(th)[:, None] * lat[None, :] + np.cos(th)[:, None] * np.array([0, 0, -1.0])[None, :]
        if rng.uniform() < 0.3:
            beams = beams[::-1].copy()
        b32 = beams.astype(np.float32)
        poses = _cloud(rng, 24, (300.0, 300.0), 5.0, roll_pitch=float(rng.choice([0.02, 0.15])))
        rmax = float(rng.choice([40.0, 120.0]))
        out = c_oracle.update_mbes(dem, poses, np.eye(4), b32.astype(np.float64), None, 1.0, 0.5, rmax,
                                   want_ranges=True, want_diag=True)
        r = hs.raycast_poses(dem, 1, rmax, np.eye(4), poses, b32, shift_hi=4, fan=2)
        ro, dg = out["ranges"], out["diag"]
        sturdy = ~((dg[..., 0] < FRAGILE) | (dg[..., 1] < FRAGILE) | (dg[..., 2] < FRAGILE))
        assert sturdy.mean() > 0.97
        assert np.abs(r - ro)[sturdy].max() <= RANGE_TOL, (trial, B, a0, a1, psi, rmax)


def test_neff_gate_arithmetic_is_the_oracles():
    """auvpf_math.cuh neff_s / neff_fixed_a (the device gate of auv_pf.py:174-181) against the oracle's
    restatement, and against the textbook 1 / sum w^2 it approximates; the sums are integers, so any
    sharding of the particles adds up to the same decision."""
    rng = np.random.default_rng(11)
    cases = [(rng.standard_normal(4096) * s).astype(np.float32) for s in (0.0, 0.3, 3.0, 30.0)]
    cases.append(np.r_[np.float32(0.0), np.full(999, -np.inf, np.float32)])        # one particle holds everything
    cases.append(np.zeros(1, np.float32))
    for lw in cases:
        qq, a, A, m = o.quantise_logw(lw)
        got = hs.neff_fixed(a)
        assert got == o.neff_fixed_a(a)
        ref = o.neff(o.weights_from_q(qq))
        assert abs(got - ref) <= 1e-2 * ref
        for ratio in (0.0, 0.1, 0.5, 1.0):
            assert o.resample_gate(a, ratio) == (not (ratio > 0.0 and got >= ratio * len(lw)))
        if len(a) >= 4:
            parts = [o.neff_sums(p) for p in np.array_split(a, 4)]
            assert o.neff_fixed_from_sums(*[sum(x) for x in zip(*parts)]) == got


def _quantised(logw):
    qq, a, A, m = o.quantise_logw(np.asarray(logw, np.float32))
    return qq, a


def _cases(N, rng):
    cases = [rng.normal(0, s, N) for s in (0.0, 0.1, 3.0, 12.0)]
    onehot = np.full(N, -np.inf)
    onehot[N // 2] = 0.0
    cases.append(onehot)
    sparse = np.full(N, -np.inf)
    sparse[rng.choice(N, max(1, N // 300), replace=False)] = rng.normal(0, 1, max(1, N // 300))
    cases.append(sparse)
    ends = rng.normal(0, 0.5, N) - 40.0
    ends[0] = ends[-1] = 0.0                             # almost all the mass on the first and the last particle
    cases.append(ends)
    return cases


@pytest.mark.parametrize("N", [1, 5, 255, 256, 257, 1023, 1024, 1025, 2047, 2048, 2049, 5000, 70001])
def test_source_side_select_is_the_references_systematic_resample(N):
    """The ancestor selection of k_select, run on the host from the same header functions (cdf_estimate /
    rel_end_fast / cdf_exact / rel_end_of): chunk and tile tables, conservative 32-ary tile location, prefix
    sums rebuilt chunk by chunk in float64 with the exact fix-up, run heads, max-scan.  Must equal the
    reference's systematic_resample (resampling.py:135-168) on the same fixed-point weights for diffuse,
    concentrated, one-hot and mostly-zero weights (zero-weight stretches longer than a tile), at the edges
    of the chunk / tile / output-block sizes."""
    rng = np.random.default_rng(N)
    for logw in _cases(N, rng):
        qq, a = _quantised(logw)
        w = o.weights_from_q(qq)
        for u in (float(rng.uniform()), 0.0, 1.0 - 2.0 ** -20):
            np.testing.assert_array_equal(hs.select_source_side(a, u), o.systematic_resample(w, u))
        # u so close to 1 that positions[-1] rounds to 1.0 (u >= 1 - N 2^-54): the reference walks off the end
        # (IndexError, :162-167); every select path hands such an output to the LAST particle
        u = 1.0 - 2.0 ** -53
        pos = (u + np.arange(N)) / float(N)
        idx = np.searchsorted(np.cumsum(w), pos, side="right")
        got = hs.select_source_side(a, u)
        inside = idx < N
        np.testing.assert_array_equal(got[inside], idx[inside])
        assert (~inside).sum() == (1 if N > 1 else 0) and np.all(got[~inside] == N - 1)


@pytest.mark.parametrize("G,M", [(2, 1000), (3, 2048), (4, 5000), (8, 2500), (2, 70001)])
def test_source_side_select_over_shards(G, M):
    """The same selection over G shards of M sources (ragged last tile per shard, shard offsets of the
    prefix sums, ancestors that cross shards): the particle-sharded filter's kernel walk."""
    rng = np.random.default_rng(G * 1000 + M)
    N = G * M
    for logw in _cases(N, rng)[1:]:
        lw = np.asarray(logw, np.float64)
        lw[: N // 2 + 3] -= 30.0 * rng.integers(0, 2)                  # sometimes almost all mass on the upper shards
        qq, a = _quantised(lw)
        w = o.weights_from_q(qq)
        for u in (float(rng.uniform()), 0.0):
            np.testing.assert_array_equal(hs.select_source_side(a, u, G=G), o.systematic_resample(w, u))
        got = hs.select_source_side(a, 1.0 - 2.0 ** -53, G=G)
        assert got[-1] == N - 1


@pytest.mark.parametrize("M,n_out", [(5000, 700), (5000, 1), (700, 5000), (4097, 4096), (2048, 100000), (70001, 1025)])
def test_source_side_select_with_a_different_number_of_positions(M, n_out):
    """The resample that changes the particle count (KLD adaptation): n_out positions over the CDF of M
    sources -- fewer and more outputs than sources, across the tile / chunk / output-block sizes -- against the
    oracle's generalised systematic resample."""
    rng = np.random.default_rng(M + n_out)
    for logw in _cases(M, rng):
        qq, a = _quantised(logw)
        w = o.weights_from_q(qq)
        for u in (float(rng.uniform()), 0.0, 1.0 - 2.0 ** -20):
            np.testing.assert_array_equal(hs.select_source_side(a, u, n_out=n_out), o.systematic_resample_to(w, u, n_out))
        got = hs.select_source_side(a, 1.0 - 2.0 ** -53, n_out=n_out)         # the last position rounds to 1.0: last SOURCE
        assert got[-1] <= M - 1 and np.all(np.diff(got) >= 0)


def test_source_side_select_exact_fixups_are_exercised():
    """Weights built so that many run ends land EXACTLY on an output position (cdf N 2^-52 - u integral):
    the estimate cannot decide those, the exact path must, and the result is still the reference's."""
    N = 4096
    a = np.full(N, 1 << 40, np.uint64)                   # uniform: cdf_j = (j+1) 2^52 / N exactly
    qq = o.quantise_a(a)[0]
    w = o.weights_from_q(qq)
    for u in (0.0, 0.5, 2.0 ** -30):
        got, (n_fast, n_exact) = hs.select_source_side(a, u, want_stats=True)
        np.testing.assert_array_equal(got, o.systematic_resample(w, u))
        if u == 0.0:
            assert n_exact > N // 2                      # positions[i] = i / N sit on the cdf grid
    rng = np.random.default_rng(3)
    a = (rng.integers(1, 1 << 20, N).astype(np.uint64)) << np.uint64(30)
    got, (n_fast, n_exact) = hs.select_source_side(a, float(rng.uniform()), want_stats=True)
    assert n_exact < 0.01 * n_fast + 10                  # generic weights: the exact path is rare


def test_source_side_select_one_million():
    """The same at 1 M + 3 particles with 300 k dead ones in a row (~146 tiles without any mass),
    against the C oracle's systematic resampler."""
    N = (1 << 20) + 3
    rng = np.random.default_rng(7)
    logw = rng.normal(0, 4.0, N).astype(np.float32)
    logw[100000:400000] = -np.inf
    qq, m, A = c_oracle.quantise_logw(logw)
    a = c_oracle.integer_weights(logw, m)
    u = float(rng.uniform())
    np.testing.assert_array_equal(hs.select_source_side(a, u), c_oracle.systematic(o.weights_from_q(qq), u))
    np.testing.assert_array_equal(hs.select_source_side(a, u, G=8 if N % 8 == 0 else 1), c_oracle.systematic(o.weights_from_q(qq), u))
