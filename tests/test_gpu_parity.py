"""GPU parity tests proper: every call goes through the C ABI (libauvpf.so) via the
Python facade and is compared with the oracle / the golden vectors of the
unmodified reference.  Tolerances are the ones BASELINE.json's north_star
states: ancestors bit-exact, beam ranges 1e-4 m, log-weights 1e-5 relative."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import pf_oracle as o
from oracle import tf_stub
from smarc_navigation_b200 import world
from tests import fragile

pytestmark = pytest.mark.gpu

RANGE_TOL = 1e-4
LOGW_RTOL = 1e-5


_DEMS = {}


def _dem(n, res, seed=1):
    key = (n, res, seed)
    if key not in _DEMS:
        _DEMS[key] = world.make_dem(n, res, seed=seed)
    return _DEMS[key]


def _pf(**kw):
    from smarc_navigation_b200 import ParticleFilter
    return ParticleFilter(**kw)


def _dyaw(a, b):
    return (a - b + np.pi) % (2 * np.pi) - np.pi


def _cloud(rng, n, centre, spread, roll_pitch=0.03, z=-5.0):
    poses = np.zeros((n, 6))
    poses[:, 0] = centre[0] + rng.normal(0, spread, n)
    poses[:, 1] = centre[1] + rng.normal(0, spread, n)
    poses[:, 2] = z
    poses[:, 3] = rng.normal(0, roll_pitch, n)
    poses[:, 4] = rng.normal(0, roll_pitch, n)
    poses[:, 5] = rng.uniform(-np.pi, np.pi, n)
    poses[:, 2:6] = poses[:, 2:6].astype(np.float32)
    return poses


# --------------------------------------------------------------------------- a2 / a9 / a3(GPS)
@pytest.mark.parametrize("case", ["a", "b"])
def test_particle_golden_init_predict_gps_roughen(golden, case):
    g = golden("particle")
    k = lambda s: g[case + "_" + s]
    N = k("init_noise").shape[0]
    pf = _pf(particle_count=N, init_covariance=k("init_cov"), motion_covariance=k("process_cov"),
             resampling_noise_covariance=k("res_cov"), measurement_std=float(k("meas_std")), m2o_matrix=k("m2o"),
             init_noise=k("init_noise"))
    np.testing.assert_allclose(pf.poses(), k("poses0"), atol=2e-6)          # fp32 noise * sigma
    pf.set_poses(k("poses_start"))
    for odom, noise, want in zip(k("odoms"), k("pred_noise"), k("pred_poses")):
        pf.predict((odom[0:3], odom[3], odom[4:8], odom[8]), dt=odom[9], noise=noise)
        got = pf.poses()
        np.testing.assert_allclose(got[:, 0:2], want[:, 0:2], atol=1e-4)     # G1
        np.testing.assert_allclose(got[:, 2:5], want[:, 2:5], atol=1e-6)
        assert np.abs(_dyaw(got[:, 5], want[:, 5])).max() < 1e-5
    # GPS weight on the GPU state vs scipy's multivariate_normal.pdf on the reference state:
    # feed the reference's final poses so only the weight arithmetic is compared
    pf.set_poses(k("pred_poses")[-1])
    w = pf.update(tuple(k("gps")))
    want_logw = np.log(k("w"))
    got_logw = pf.log_weights().astype(np.float64)
    tol = LOGW_RTOL * np.abs(want_logw) + 5e-5 * (1.0 / float(k("meas_std")) ** 2)   # fp32 yaw/z state quantisation
    assert np.all(np.abs(got_logw - want_logw) <= tol)
    np.testing.assert_allclose(w, np.exp(got_logw) + 1e-200, rtol=1e-12)
    pf.roughen(k("res_cov"), noise=k("res_noise"))
    got = pf.poses()
    np.testing.assert_allclose(got[:, 0:2], k("poses_rough")[:, 0:2], atol=1e-5)
    np.testing.assert_allclose(got[:, 2:5], k("poses_rough")[:, 2:5], atol=2e-6)
    assert np.abs(_dyaw(got[:, 5], k("poses_rough")[:, 5])).max() < 1e-5
    pf.close()


def test_predict_many_equals_sequential_predicts():
    """One fused pass over the state == the same predicts one by one (state and RNG counters)."""
    N, F = 5000, 3
    rng = np.random.default_rng(9)
    odoms = [[world.Odom(rng.normal(1.0, 0.1, 3), 0.05 * rng.normal(), world.quat_from_rpy(0.02, -0.01, 0.3), -5.0, 0.02)
              for _ in range(F)] for _ in range(5)]
    a = _pf(particle_count=N, n_filters=F, init_covariance=[1.0, 1.0, 0, 0, 0, 0.1],
            motion_covariance=[1e-3, 1e-3, 0, 0, 0, 1e-5], seed=77)
    b = _pf(particle_count=N, n_filters=F, init_covariance=[1.0, 1.0, 0, 0, 0, 0.1],
            motion_covariance=[1e-3, 1e-3, 0, 0, 0, 1e-5], seed=77)
    for od in odoms:
        a.predict(od)
    b.predict_many(odoms)
    np.testing.assert_array_equal(a.poses(), b.poses())
    a.predict(odoms[0])                      # the step counters advanced identically
    b.predict_many(odoms[:1])
    np.testing.assert_array_equal(a.poses(), b.poses())
    nz = rng.normal(size=(2, F * N, 6)).astype(np.float32)
    a.predict(odoms[1], noise=nz[0])
    a.predict(odoms[2], noise=nz[1])
    b.predict_many(odoms[1:3], noise=nz)
    np.testing.assert_array_equal(a.poses(), b.poses())
    a.close()
    b.close()


def test_philox_streams_match_spec():
    """Device-generated normals follow the documented Philox contract."""
    N, seed = 4096, 0x1234567890ABCDEF
    var = [4.0, 1.0, 0.25, 0.01, 0.04, 0.09]
    pf = _pf(particle_count=N, init_covariance=var, seed=seed, global_offset=1000, global_count=100000)
    ids = 1000 + np.arange(N, dtype=np.int64)
    n = o.philox_pose_normals(seed, 1, 0, ids)                           # stream 1 = init, step 0
    np.testing.assert_allclose(pf.poses(), np.sqrt(var)[None, :] * n, rtol=1e-4, atol=2e-4)
    p0 = pf.poses()
    pf.roughen([1.0] * 6)                                                      # stream 3, step 0
    n = o.philox_pose_normals(seed, 3, 0, ids)
    np.testing.assert_allclose(pf.poses() - p0, n, rtol=1e-4, atol=1e-4)
    pf.close()
    # predict: stream 2, normals 0,1 -> x,y ; 2 -> yaw
    pf = _pf(particle_count=N, motion_covariance=[1.0, 1.0, 0, 0, 0, 1e-2], seed=seed)
    pf.predict(((0.0, 0.0, 0.0), 0.0, (0, 0, 0, 1.0), -3.0), dt=0.1)
    n = o.philox_normals(seed, 2, 0, np.arange(N, dtype=np.int64), 1)
    got = pf.poses()
    np.testing.assert_allclose(got[:, 0], n[:, 0], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(got[:, 1], n[:, 1], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(got[:, 5], 0.1 * n[:, 2], rtol=1e-4, atol=1e-5)
    assert np.all(got[:, 2] == -3.0)
    pf.close()


# --------------------------------------------------------------------------- a4 raycast, a3 MBES
def _check_ranges(pf, dem, poses, beams, m2o, range_max, max_fragile=2e-3, label=None):
    """Sturdy rays within RANGE_TOL of the oracle; fragile rays (the oracle's own first-hit decision hinges
    on < 2 mm) are NOT dropped: each must lie inside the bracket of the oracle's answers for a sensor
    moved 5 mm down / up (tests/fragile.py); fraction and worst errors go to gpurun_out/fragile_report.jsonl."""
    r = pf.raycast(poses)
    ro, stats = fragile.check_ranges(r, dem, poses, m2o, beams, range_max, max_fragile, label=label)
    return ro, r


@pytest.mark.parametrize("cfg,coarse", [("C2", -1), ("C2", 1), ("C2", (1, 4)), ("C2", 4), ("C3", -1), ("C3", 5),
                                        ("C3", (1, 3)), ("C3", (3, 6))])
def test_raycast_ranges_vs_oracle(cfg, coarse):
    """G4: >= 1e5 rays per config within 1e-4 m of the float64 triangulated-DEM oracle."""
    c = world.CONFIGS[cfg]
    dem = _dem(c["dem_n"], c["res"])
    beams = world.make_beams(c["beams"])
    rng = np.random.default_rng(21)
    ext = c["dem_n"] * c["res"]
    n = 400 if cfg == "C2" else 200
    poses = np.concatenate([_cloud(rng, n // 2, (0.3 * ext, 0.3 * ext), 3.0),
                            _cloud(rng, n // 2, (0.93 * ext, 0.91 * ext), 2.0)])
    pf = _pf(particle_count=16, range_max=world.RANGE_MAX, coarse_level=coarse)
    pf.set_dem(dem)
    pf.set_beams(beams)
    ro, r = _check_ranges(pf, dem, poses, beams, np.eye(4), world.RANGE_MAX, label="%s coarse=%s" % (cfg, coarse))
    assert ro.size >= 100000 and (ro < world.RANGE_MAX).all()
    pf.close()


def _ridged_dem(n, res, seed=5):
    """Steep short-wavelength ridges on top of the synthetic bathymetry: outer beams are occluded
    (first-hit jumps of metres), near misses over crests are common -- the terrain on which the
    fragile-ray accounting matters."""
    d = world.make_dem(n, res, seed=seed)
    x = np.arange(n, dtype=np.float32) * np.float32(res)
    X, Y = np.meshgrid(x, x)
    d["z"] = (d["z"] + 2.5 * np.sin(2 * np.pi * X / 9.0) * np.cos(2 * np.pi * Y / 13.0)
              + 1.0 * np.sin(2 * np.pi * (X + Y) / 4.0)).astype(np.float32)
    return d


def test_raycast_and_logw_on_ridged_terrain_c3_size():
    """C3-size map (4000^2 at 0.5 m) with occluding ridges, C3 beam count: ranges of > 1e5 rays and the
    log-weights of their particles against the oracle, fragile rays bounded by their bracket and counted."""
    c = world.CONFIGS["C3"]
    dem = _ridged_dem(c["dem_n"], c["res"])
    B = c["beams"]
    beams = world.make_beams(B).astype(np.float32).astype(np.float64)
    rng = np.random.default_rng(17)
    poses = np.concatenate([_cloud(rng, 128, (600.0, 600.0), 4.0, roll_pitch=0.04),
                            _cloud(rng, 128, (1500.0, 400.0), 25.0, roll_pitch=0.04)])
    pf = _pf(particle_count=len(poses), measurement_std=1.0, range_min=world.RANGE_MIN, range_max=world.RANGE_MAX)
    pf.set_dem(dem)
    pf.set_beams(beams)
    ro, r = _check_ranges(pf, dem, poses, beams, np.eye(4), world.RANGE_MAX, max_fragile=5e-2, label="ridged C3-size")
    assert ro.size >= 100000
    assert (np.abs(np.diff(ro, axis=1)) > 2.0).sum() > 100          # occlusion jumps are present
    truth = np.array([[600.0, 600.0, -5.0, 0.0, 0.0, 0.3]])
    meas = (c_oracle.update_mbes(dem, truth, np.eye(4), beams, None, 1.0, 0.5, world.RANGE_MAX, want_ranges=True)["ranges"][0]
            + rng.normal(0, 0.05, B)).astype(np.float32)
    pf.set_poses(poses)
    pf.update_mbes(meas, return_weights=False)
    got = pf.log_weights().astype(np.float64)
    out = c_oracle.update_mbes(dem, poses, np.eye(4), beams, meas.astype(np.float64), 1.0, world.RANGE_MIN,
                               world.RANGE_MAX, want_ranges=True, want_diag=True)
    extra, fr = fragile.logw_bound(dem, poses, np.eye(4), beams, meas, 1.0, world.RANGE_MIN, world.RANGE_MAX, out)
    lognorm = -B * np.log(np.sqrt(2 * np.pi))
    tol = LOGW_RTOL * (np.abs(out["logw"] - lognorm) + abs(lognorm)) + extra
    d = np.abs(got - out["logw"])
    assert np.all(d <= tol), np.max(d / tol)
    fragile.report({"label": "ridged C3-size log-weights", "particles": len(poses), "fragile_frac": float(fr.mean()),
                    "particles_with_fragile_beam": float((extra > 0).mean()),
                    "max_dlogw_fragile_owners": float(d[extra > 0].max()) if (extra > 0).any() else 0.0,
                    "max_dlogw_sturdy": float(d[extra == 0].max()) if (extra == 0).any() else 0.0})
    pf.close()


def test_raycast_edge_cases():
    dem = world.make_dem(300, 2.0, seed=11)
    beams = world.make_beams(128, half_angle_deg=75.0)
    rng = np.random.default_rng(7)
    q = tf_stub.quaternion_from_euler(0.01, -0.02, 0.7)
    m2o = tf_stub.translation_matrix([300.0, 250.0, 1.0]).dot(tf_stub.quaternion_matrix(q))
    for m, rmax in ((np.eye(4), world.RANGE_MAX), (m2o, world.RANGE_MAX), (np.eye(4), 30.0)):
        pf = _pf(particle_count=16, range_max=rmax, m2o_matrix=m)
        pf.set_dem(dem)
        pf.set_beams(beams)
        c = (0.0, 0.0) if m is m2o else (300.0, 300.0)
        _check_ranges(pf, dem, _cloud(rng, 64, c, 3.0), beams, m, rmax, 2e-2)
        if m is not m2o:
            ro, r = _check_ranges(pf, dem, _cloud(rng, 64, (8.0, 590.0), 6.0), beams, m, rmax, 2e-2)       # border
            assert (ro >= rmax).any()
            _check_ranges(pf, dem, _cloud(rng, 32, (-40.0, 300.0), 3.0), beams, m, rmax, 2e-2)           # outside
            _check_ranges(pf, dem, _cloud(rng, 64, (300.0, 300.0), 3.0, roll_pitch=0.5), beams, m, rmax, 2e-2)
            up = _cloud(rng, 16, (300.0, 300.0), 3.0)
            up[:, 3] = np.float32(np.pi)
            ro, r = _check_ranges(pf, dem, up, beams, m, rmax)
            assert (r == np.float32(rmax)).all()                                                         # all miss
            below = _cloud(rng, 16, (300.0, 300.0), 3.0, z=-60.0)
            ro, r = _check_ranges(pf, dem, below, beams, m, rmax)
            assert (r == 0.0).all()
        pf.close()


@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_update_mbes_logw_vs_oracle(cfg):
    """G2 (MBES form): fused raycast + likelihood log-weights within 1e-5 relative."""
    c = world.CONFIGS[cfg]
    dem = _dem(c["dem_n"], c["res"])
    B = c["beams"]
    beams = world.make_beams(B)
    rng = np.random.default_rng(31)
    ext = c["dem_n"] * c["res"]
    N = 512
    poses = _cloud(rng, N, (0.3 * ext, 0.3 * ext), 2.0)
    truth = poses[:1].copy()
    truth[0, 0:2] = 0.3 * ext
    r_true = c_oracle.update_mbes(dem, truth, np.eye(4), beams, None, 1.0, 0.5, world.RANGE_MAX, want_ranges=True)["ranges"][0]
    meas = (r_true + rng.normal(0, 0.05, B)).astype(np.float32)
    meas[5] = np.nan                # invalid beams are skipped
    meas[17] = 0.1                  # below range_min
    meas[40] = 500.0                # above range_max
    for std in (1.0, 0.2):
        pf = _pf(particle_count=N, measurement_std=std, range_min=world.RANGE_MIN, range_max=world.RANGE_MAX)
        pf.set_dem(dem)
        pf.set_beams(beams)
        pf.set_poses(poses)
        w = pf.update(meas)
        got = pf.log_weights().astype(np.float64)
        want = c_oracle.update_mbes(dem, poses, np.eye(4), beams, meas.astype(np.float64), std, world.RANGE_MIN,
                                    world.RANGE_MAX)["logw"]
        lognorm = -(B - 3) * np.log(std * np.sqrt(2 * np.pi))
        tol = LOGW_RTOL * (np.abs(want - lognorm) + abs(lognorm))      # relative to the terms, not their difference
        assert np.all(np.abs(got - want) <= tol), np.max(np.abs(got - want) / tol)
        assert w.shape == (N,) and np.all(w >= 1e-200)
        pf.close()


def test_update_mbes_mixed_profile_and_per_ray_particles():
    """Particles the profile march cannot certify (sensor reach leaves the map, large tilt on steep
    facets) are listed by the profile kernel and recast ray by ray: the log-weights of a cloud that
    mixes both kinds -- in two filters at once -- still match the oracle, and resampling on them
    (max published by both kernels) is exact."""
    dem = _dem(500, 2.0)
    B, N, F = 128, 3000, 2
    beams = world.make_beams(B, 70.0).astype(np.float32).astype(np.float64)     # identical directions for both arms
    rng = np.random.default_rng(77)
    near_border = _cloud(rng, N // 3, (30.0, 500.0), 25.0)            # reach (120 m) crosses x = 0 for most
    tilted = _cloud(rng, N // 3, (500.0, 500.0), 3.0, roll_pitch=0.35)
    plain = _cloud(rng, N - 2 * (N // 3), (500.0, 500.0), 3.0)
    one = np.concatenate([near_border, tilted, plain])[rng.permutation(N)]
    poses = np.concatenate([one, one[::-1]])                            # second filter: the same cloud reversed
    truth = np.array([[500.0, 500.0, -5.0, 0.0, 0.0, 0.3]])
    r_true = c_oracle.update_mbes(dem, truth, np.eye(4), beams, None, 1.0, 0.5, world.RANGE_MAX, want_ranges=True)["ranges"][0]
    meas = np.stack([r_true + rng.normal(0, 0.05, B), r_true + rng.normal(0, 0.05, B)]).astype(np.float32)
    pf = _pf(particle_count=N, n_filters=F, measurement_std=1.0, range_min=world.RANGE_MIN, range_max=world.RANGE_MAX)
    pf.set_dem(dem)
    pf.set_beams(beams)
    pf.set_poses(poses)
    pf.update_mbes(meas, return_weights=False)
    got = pf.log_weights().astype(np.float64).reshape(F, N)
    lognorm = -B * np.log(np.sqrt(2 * np.pi))
    for f in range(F):
        pz = poses[f * N:(f + 1) * N]
        out = c_oracle.update_mbes(dem, pz, np.eye(4), beams, meas[f].astype(np.float64), 1.0,
                                   world.RANGE_MIN, world.RANGE_MAX, want_ranges=True, want_diag=True)
        want = out["logw"]
        # no particle is excluded: one that owns fragile beams gets exactly the slack those beams justify
        extra, fr = fragile.logw_bound(dem, pz, np.eye(4), beams, meas[f], 1.0, world.RANGE_MIN, world.RANGE_MAX, out)
        tol = LOGW_RTOL * (np.abs(want - lognorm) + abs(lognorm)) + extra
        d = np.abs(got[f] - want)
        assert np.all(d <= tol), np.max(d / tol)
        fragile.report({"label": "mixed profile/per-ray filter %d" % f, "particles": N, "fragile_frac": float(fr.mean()),
                        "particles_with_fragile_beam": float((extra > 0).mean()),
                        "max_dlogw_fragile_owners": float(d[extra > 0].max()) if (extra > 0).any() else 0.0,
                        "max_dlogw_bound": float(extra.max())})
    anc = pf.resample(u=[0.25, 0.75], return_ancestors=True, roughen=False)
    qq = pf.weights_q().reshape(F, N)
    lw32 = pf.log_weights().reshape(F, N)
    for f in range(F):
        np.testing.assert_array_equal(qq[f], o.quantise_logw(lw32[f])[0])       # the max came from BOTH kernels
        np.testing.assert_array_equal(anc.reshape(F, N)[f], o.systematic_resample(o.weights_from_q(qq[f]), [0.25, 0.75][f]))
    pf.close()


@pytest.mark.parametrize("rmax", [world.RANGE_MAX, 35.0])
def test_update_mbes_unusable_beams_on_every_path(rmax):
    """A ping whose beams are partly unusable (NaN, inf, below range_min, beyond range_max -- the nadir beam and the
    outermost beam of either side among them) on ridged terrain: the march's no-graze and grazing beam loops, the
    map-border variant with its ray-by-ray finish, and the per-ray fallback must all leave those beams out of the
    sum exactly as the oracle does (the march carries no validity test per beam: an unusable beam is a table entry
    whose range and error come out as zero).  With the short range_max the outer beams MISS as well."""
    dem = _ridged_dem(700, 0.5)
    ext = (dem["z"].shape[0] - 1) * dem["res"]
    B = 256
    beams = world.make_beams(B, 65.0).astype(np.float32).astype(np.float64)
    rng = np.random.default_rng(123)
    poses = np.concatenate([_cloud(rng, 400, (0.5 * ext, 0.5 * ext), 6.0, roll_pitch=0.03),          # interior
                            _cloud(rng, 400, (0.5 * ext, 0.5 * ext), 6.0, roll_pitch=0.20),          # grazing beams / fallback
                            _cloud(rng, 300, (0.04 * ext, 0.5 * ext), 5.0, roll_pitch=0.03),         # reach leaves the map
                            _cloud(rng, 300, (0.95 * ext, 0.93 * ext), 5.0, roll_pitch=0.05)])       # ... in a corner
    poses = poses[(poses[:, 0] > 0) & (poses[:, 0] < ext) & (poses[:, 1] > 0) & (poses[:, 1] < ext)]
    truth = np.array([[0.5 * ext, 0.5 * ext, -5.0, 0.0, 0.0, 0.3]])
    meas = (c_oracle.update_mbes(dem, truth, np.eye(4), beams, None, 1.0, 0.5, rmax, want_ranges=True)["ranges"][0]
            + rng.normal(0, 0.05, B)).astype(np.float32)
    bad = rng.choice(B, B // 4, replace=False)
    meas[bad[0::4]] = np.nan
    meas[bad[1::4]] = np.inf
    meas[bad[2::4]] = 0.5 * world.RANGE_MIN
    meas[bad[3::4]] = 2.0 * rmax
    meas[[0, B - 1, B // 2]] = np.nan               # outermost beam of either side, and the nadir
    pf = _pf(particle_count=len(poses), measurement_std=1.0, range_min=world.RANGE_MIN, range_max=rmax)
    pf.set_dem(dem)
    pf.set_beams(beams)
    pf.set_poses(poses)
    pf.update_mbes(meas, return_weights=False)
    got = pf.log_weights().astype(np.float64)
    flagged = pf.mbes_flagged()
    out = c_oracle.update_mbes(dem, poses, np.eye(4), beams, meas.astype(np.float64), 1.0, world.RANGE_MIN, rmax,
                               want_ranges=True, want_diag=True)
    extra, fr = fragile.logw_bound(dem, poses, np.eye(4), beams, meas, 1.0, world.RANGE_MIN, rmax, out)
    n_valid = int(np.sum(np.isfinite(meas) & (meas >= world.RANGE_MIN) & (meas <= rmax)))
    assert 0.4 * B < n_valid < 0.85 * B
    lognorm = -n_valid * np.log(np.sqrt(2 * np.pi))
    tol = LOGW_RTOL * (np.abs(out["logw"] - lognorm) + abs(lognorm)) + extra
    d = np.abs(got - out["logw"])
    assert np.all(d <= tol), np.max(d / tol)
    assert int(flagged[:, 0].sum()) > 100 and int(flagged[:, 1].sum()) > 0          # every path was really taken
    if rmax < world.RANGE_MAX:
        assert (out["ranges"] >= rmax).mean() > 0.05                       # misses are present
    fragile.report({"label": "unusable beams, rmax %g" % rmax, "particles": len(poses), "fragile_frac": float(fr.mean()),
                    "border_pass": int(flagged[:, 0].sum()), "ray_by_ray": int(flagged[:, 1].sum()),
                    "max_dlogw_over_tol": float(np.max(d / tol))})
    pf.close()


@pytest.mark.parametrize("seed", [2, 5, 11])
def test_randomised_scenes_on_the_gpu(seed):
    """The adversarial sweep of tests/test_hostsim_math.py, on the device: smooth / ridged / terraced / spiked terrain,
    sensors 8-25 m above it tilted by up to ~40 degrees, 64-512 beams, range_max 20-120 m, pings with a few unusable
    beams.  The update kernels (march, border variant, ray by ray: whichever takes a sensor) against the oracle's
    log-weights with the fragile-ray slack, and the probe kernel's ranges with the fragile-ray bracket."""
    rng = np.random.default_rng(seed)

    def terrain(kind, n, res, tseed):
        if kind == "smooth":
            return world.make_dem(n, res, seed=tseed)
        if kind == "ridged":
            return _ridged_dem(n, res, seed=tseed)
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

    worst = 0.0
    for trial in range(40):
        kind = rng.choice(["smooth", "ridged", "steps", "spikes"])
        res, n = float(rng.choice([0.5, 1.0, 2.0])), int(rng.choice([300, 500]))
        dem = terrain(kind, n, res, int(rng.integers(1, 100)))
        ext = (n - 1) * res
        B = int(rng.choice([64, 256, 512]))
        beams = world.make_beams(B, float(rng.choice([45, 60, 70]))).astype(np.float32).astype(np.float64)
        centre = (rng.uniform(0.15, 0.85) * ext, rng.uniform(0.15, 0.85) * ext)
        alt = float(rng.choice([-5.0, -15.0, -22.0]))
        poses = _cloud(rng, 160, centre, float(rng.choice([0.5, 5.0, 30.0])), roll_pitch=float(rng.choice([0.02, 0.1, 0.25])), z=alt)
        poses = poses[(poses[:, 0] > 0) & (poses[:, 0] < ext) & (poses[:, 1] > 0) & (poses[:, 1] < ext)]
        if len(poses) < 8:
            continue
        rmax = float(rng.choice([20.0, 35.0, 60.0, 120.0]))
        truth = np.array([[centre[0], centre[1], alt, 0.0, 0.0, 0.3]])
        meas = (c_oracle.update_mbes(dem, truth, np.eye(4), beams, None, 1.0, 0.5, rmax, want_ranges=True)["ranges"][0]
                + rng.normal(0, 0.05, B)).astype(np.float32)
        meas[rng.choice(B, 5, replace=False)] = np.nan
        pf = _pf(particle_count=len(poses), measurement_std=1.0, range_min=world.RANGE_MIN, range_max=rmax)
        pf.set_dem(dem)
        pf.set_beams(beams)
        pf.set_poses(poses)
        pf.update_mbes(meas, return_weights=False)
        got = pf.log_weights().astype(np.float64)
        out = c_oracle.update_mbes(dem, poses, np.eye(4), beams, meas.astype(np.float64), 1.0, world.RANGE_MIN, rmax,
                                   want_ranges=True, want_diag=True)
        extra, fr = fragile.logw_bound(dem, poses, np.eye(4), beams, meas, 1.0, world.RANGE_MIN, rmax, out)
        n_valid = int(np.sum(np.isfinite(meas) & (meas >= world.RANGE_MIN) & (meas <= rmax)))
        lognorm = -n_valid * np.log(np.sqrt(2 * np.pi))
        tol = LOGW_RTOL * (np.abs(out["logw"] - lognorm) + abs(lognorm)) + extra
        d = np.abs(got - out["logw"])
        assert np.all(d <= tol), (trial, kind, res, rmax, float(np.max(d / tol)))
        worst = max(worst, float(np.max(d[tol > 0] / tol[tol > 0])) if (tol > 0).any() else 0.0)
        _check_ranges(pf, dem, poses[:48], beams, np.eye(4), rmax, max_fragile=0.2, label="sweep %d/%d %s" % (seed, trial, kind))
        pf.close()
    assert worst <= 1.0


def test_quantised_dem_file_parity(tmp_path):
    """Map ingestion: a DEM file with uint16 heights.  The library and the oracle both see the
    dequantised float32 map, so ranges and log-weights keep the usual tolerances ON THAT MAP."""
    from smarc_navigation_b200 import dem_io
    dem = _dem(500, 2.0)
    path = str(tmp_path / "map_u16.npz")
    dem_io.save_dem(path, dem, quantised=True)
    beams = world.make_beams(128)
    rng = np.random.default_rng(3)
    poses = _cloud(rng, 200, (500.0, 500.0), 3.0)
    pf = _pf(particle_count=200, range_min=world.RANGE_MIN, range_max=world.RANGE_MAX)
    loaded = pf.load_dem(path)
    assert loaded["z"].dtype == np.float32 and np.abs(loaded["z"] - dem["z"]).max() < 1e-3 and not np.array_equal(loaded["z"], dem["z"])
    pf.set_beams(beams)
    _check_ranges(pf, loaded, poses, beams, np.eye(4), world.RANGE_MAX, label="uint16 DEM file")
    pf.close()


def test_mesh_rasteriser_round_trip_and_overlap():
    """Map ingestion: the DEM's own triangulation rasterises back to the DEM bit for bit; a coarser
    mesh interpolates linearly; where two sheets overlap the higher one wins; uncovered vertices
    keep the fill value."""
    dem = _dem(120, 2.0)
    pf = _pf(particle_count=4)
    v, f = world.dem_to_mesh(dem)
    z = pf.rasterise_mesh(v, f, dem["z"].shape, dem["x0"], dem["y0"], dem["res"])
    np.testing.assert_array_equal(z, dem["z"])
    # one big tilted triangle pair over part of the grid + a higher small sheet, finer grid than the mesh
    verts = np.array([[10, 10, -30], [110, 10, -20], [110, 90, -25], [10, 90, -35],
                      [40, 40, -5], [60, 40, -5], [60, 60, -5], [40, 60, -5]], np.float32)
    faces = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.int32)
    zz = pf.rasterise_mesh(verts, faces, (101, 121), 0.0, 0.0, 1.0, fill=-999.0)
    assert zz[5, 5] == -999.0 and zz[95, 115] == -999.0                       # outside every triangle
    assert zz[50, 50] == -5.0 and zz[40, 60] == -5.0                          # the higher sheet, edges included
    np.testing.assert_allclose(zz[10, 10], -30.0, atol=1e-5)
    np.testing.assert_allclose(zz[10, 60], -25.0, atol=1e-5)                  # halfway along the first edge
    # the lower triangle (0,1,2) is the plane z = -30 + 0.1 (x - 10) - 0.0625 (y - 10)
    np.testing.assert_allclose(zz[30, 100], -30.0 + 0.1 * 90 - 0.0625 * 20, atol=1e-4)
    pf.set_dem(np.where(zz == -999.0, -40.0, zz).astype(np.float32), 0.0, 0.0, 1.0)
    pf.set_scan_geometry(-0.5, 0.1, 11)
    r = pf.raycast(np.array([[50.0, 50.0, 0.0, 0.0, 0.0, 0.0]]))
    np.testing.assert_allclose(r[0, 5], 5.0, atol=1e-4)                        # nadir beam lands on the upper sheet
    pf.close()


# --------------------------------------------------------------------------- a5 / a7 / a8(sorted) / a9
@pytest.mark.parametrize("N,scale", [(1, 1.0), (100, 3.0), (4097, 10.0), (65536, 25.0), (1048576, 6.0),
                                     (16777216, 4.0)])
def test_systematic_resample_bit_exact(N, scale):
    """G3: ancestors == reference systematic_resample on the GPU's own quantised weights,
    and the quantised weights == the oracle's fixed-point pipeline, bit for bit."""
    rng = np.random.default_rng(N)
    logw = (rng.normal(0, scale, N) - 470.0).astype(np.float32)
    if N > 10:
        logw[3] = -np.inf
        logw[5] = np.nan
    u = float(rng.uniform())
    pf = _pf(particle_count=N, resampling_noise_covariance=[0.0] * 6)
    poses = np.zeros((N, 6))
    poses[:, 0] = np.arange(N)
    poses[:, 1] = -2.0 * np.arange(N)
    poses[:, 5] = (np.arange(N) % 7).astype(np.float32)
    pf.set_poses(poses)
    pf.set_log_weights(logw)
    anc = pf.resample(u=u, return_ancestors=True, roughen=False)
    qq = pf.weights_q()
    q_ref, m, A = c_oracle.quantise_logw(logw)
    np.testing.assert_array_equal(qq, q_ref)
    w = o.weights_from_q(qq)
    np.testing.assert_array_equal(anc, c_oracle.systematic(w, u))
    if N <= 65536:
        np.testing.assert_array_equal(anc, o.systematic_resample(w, u))
        if N >= 100:
            ref = o.normalise_logw(np.where(np.isnan(logw), -np.inf, logw))
            np.testing.assert_allclose(w, ref, atol=2.0 ** -50)
    assert np.all(np.diff(anc) >= 0)
    got = pf.poses()
    np.testing.assert_array_equal(got[:, 0], poses[anc, 0])
    np.testing.assert_array_equal(got[:, 1], poses[anc, 1])
    np.testing.assert_array_equal(got[:, 5], poses[anc, 5])
    neff = pf.neff()[0]
    np.testing.assert_allclose(neff, 1.0 / np.sum(w * w), rtol=1e-9)
    pf.close()


def test_resample_reference_golden_weights(golden):
    """The reference's resample(weights) call: linear float64 weights in, systematic out."""
    g = golden("resampling")
    names = sorted({k[:-2] for k in g.files if k.endswith("_w")})
    for n in names:
        w, u = g[n + "_w"], float(g[n + "_u"])
        N = len(w)
        pf = _pf(particle_count=N)
        anc = pf.resample(weights=w, u=u, return_ancestors=True, roughen=False)
        qq = pf.weights_q()
        np.testing.assert_array_equal(qq, o.quantise_linear(w)[0])
        wq = o.weights_from_q(qq)
        np.testing.assert_allclose(wq, w / w.sum(), atol=2.0 ** -50)
        np.testing.assert_array_equal(anc, o.systematic_resample(wq, u))
        want = g[n + "_systematic"]
        if want[0] >= 0:
            # vs the reference run on the UNquantised weights: identical except where a position
            # falls within 2^-52-grid distance of a CDF step
            assert (anc != want).mean() <= 0.01
        pf.close()


def test_uniform_weights_identity_and_roughening():
    N = 1 << 20
    pf = _pf(particle_count=N, resampling_noise_covariance=[0.25, 0.25, 0, 0, 0, 1e-4], seed=5)
    pf.set_log_weights(np.full(N, -123.0, np.float32))
    p0 = np.zeros((N, 6))
    p0[:, 0] = np.arange(N)
    pf.set_poses(p0)
    anc = pf.resample(u=0.37, return_ancestors=True)                 # fused Philox roughening
    np.testing.assert_array_equal(anc, np.arange(N, dtype=np.int32))
    got = pf.poses()
    n = o.philox_pose_normals(5, 3, 0, np.arange(N, dtype=np.int64))
    # fp32 Box-Muller: u0 -> 1 loses relative precision in r = sqrt(-2 ln u0); absolute error stays < 1e-4
    np.testing.assert_allclose(got[:, 0] - p0[:, 0], 0.5 * n[:, 0], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(got[:, 1], 0.5 * n[:, 1], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(got[:, 5], 0.01 * n[:, 5], rtol=1e-4, atol=1e-6)
    assert np.all(got[:, 2:5] == 0.0)
    pf.close()


def test_degenerate_weights_raise_and_leave_state():
    from smarc_navigation_b200 import DegenerateWeights
    N = 1000
    pf = _pf(particle_count=N, init_covariance=[1.0] * 6)
    p0 = pf.poses()
    pf.set_log_weights(np.full(N, -np.inf, np.float32))
    with pytest.raises(DegenerateWeights):
        pf.resample(u=0.5)
    pf.set_log_weights(np.full(N, np.nan, np.float32))
    with pytest.raises(DegenerateWeights):
        pf.resample(u=0.5)
    with pytest.raises(DegenerateWeights):
        pf.resample(weights=np.zeros(N), u=0.5)
    np.testing.assert_array_equal(pf.poses(), p0)
    # enqueue-only mode: no exception, the state is carried over, the flag can be queried
    pf.resample(u=0.5, wait=False)
    assert pf.resample_status()[0]
    np.testing.assert_array_equal(pf.poses(), p0)
    pf.set_log_weights(np.zeros(N, np.float32))
    pf.resample(u=0.5, wait=False, roughen=False)
    assert not pf.resample_status()[0]
    np.testing.assert_array_equal(pf.poses(), p0)          # uniform weights: identity ancestors
    pf.close()


# --------------------------------------------------------------------------- a8 reference slot map
def test_neff_gated_resample_matches_the_oracle_gate():
    """The gate the reference carries commented out (auv_pf.py:174-181: resample only when
    N_eff < pc/2), decided on the device.  Per filter of one handle: the decision equals the
    oracle's, a held-back filter keeps poses bit for bit (no roughening, identity ancestors), a
    resampled one is exactly what the ungated filter produces."""
    F, M = 4, 6000
    rng = np.random.default_rng(21)
    scales = [0.05, 0.7, 1.6, 8.0]                      # N_eff / M from ~1 down to a handful of particles
    logw = np.concatenate([(rng.normal(0, s, M) - 30.0) for s in scales]).astype(np.float32)
    poses = np.zeros((F * M, 6))
    poses[:, 0] = np.arange(F * M)
    poses[:, 1] = 0.5 * np.arange(F * M)
    u = rng.uniform(size=F)
    res = [0.01, 0.01, 0.0, 0.0, 0.0, 1e-5]
    out = {}
    for ratio in (0.0, 0.5):
        pf = _pf(particle_count=M, n_filters=F, resampling_noise_covariance=res, seed=3, resample_neff_ratio=ratio)
        pf.set_poses(poses)
        pf.set_log_weights(logw)
        anc = pf.resample(u=u, return_ancestors=True)
        out[ratio] = (anc.reshape(F, M), pf.poses().reshape(F, M, 6), pf.resample_skipped(), pf.neff_fixed(), pf.neff())
        pf.close()
    anc0, p0, sk0, nf0, ne0 = out[0.0]
    anc1, p1, sk1, nf1, ne1 = out[0.5]
    assert not sk0.any()
    want_skip = []
    for f in range(F):
        qq, a, A, m = o.quantise_logw(logw[f * M:(f + 1) * M])
        assert nf1[f] == o.neff_fixed_a(a) == nf0[f]
        np.testing.assert_allclose(nf1[f], ne1[f], rtol=1e-2)
        want_skip.append(not o.resample_gate(a, 0.5))
    assert want_skip == [True, True, False, False]       # the cases straddle the gate
    np.testing.assert_array_equal(sk1, want_skip)
    for f in range(F):
        if want_skip[f]:
            np.testing.assert_array_equal(anc1[f], np.arange(M))
            np.testing.assert_array_equal(p1[f], poses.reshape(F, M, 6)[f])
        else:
            np.testing.assert_array_equal(anc1[f], anc0[f])
            np.testing.assert_array_equal(p1[f], p0[f])
    # the synchronising paths (residual draw, reference slot map) take the same decision on the host
    for kw in (dict(resampler="residual"), dict(inplace_reference=True), dict(resampler="residual", inplace_reference=True)):
        for f in (1, 2):
            pf = _pf(particle_count=M, resampling_noise_covariance=res, seed=3, resample_neff_ratio=0.5, **kw)
            pf.set_poses(poses[f * M:(f + 1) * M])
            pf.set_log_weights(logw[f * M:(f + 1) * M])
            a = pf.resample(return_ancestors=True)
            assert bool(pf.resample_skipped()[0]) == want_skip[f]
            if want_skip[f]:
                np.testing.assert_array_equal(a, np.arange(M))
                np.testing.assert_array_equal(pf.poses(), poses[f * M:(f + 1) * M])
            else:
                assert not np.array_equal(pf.poses(), poses[f * M:(f + 1) * M])
            pf.close()
    with pytest.raises(Exception):
        _pf(particle_count=64, n_filters=2, resampler="residual", resample_neff_ratio=0.5).resample()


def test_inplace_reference_map(golden):
    """auv_pf.resample's keep/lost/dupes reassignment (auv_pf.py:183-198) on the golden
    residual_resample indices, then roughening with the reference's own draws."""
    g = golden("auv_pf_cycle")
    N = g["init_noise"].shape[0]
    for c in range(g["gps"].shape[0]):
        pf = _pf(particle_count=N, resampling_noise_covariance=g["res_cov"])
        pf.set_poses(g["poses_pred"][c])
        pf.apply_indices(g["indices"][c])
        got = pf.poses()
        A = o.ancestor_map(g["indices"][c], N)
        np.testing.assert_array_equal(got[:, 0:2], g["poses_pred"][c][A, 0:2])
        pf.roughen(noise=g["res_noise"][c])
        got = pf.poses()
        np.testing.assert_allclose(got[:, 0:2], g["poses_res"][c][:, 0:2], atol=1e-5)
        assert np.abs(_dyaw(got[:, 5], g["poses_res"][c][:, 5])).max() < 1e-5
        pf.close()
    rng = np.random.default_rng(3)
    for N in (1, 2, 1000, 300000):
        idx = rng.integers(0, N, N).astype(np.int32)
        pf = _pf(particle_count=N)
        p = np.zeros((N, 6))
        p[:, 0] = np.arange(N)
        pf.set_poses(p)
        pf.apply_indices(idx)
        np.testing.assert_array_equal(pf.poses()[:, 0], o.ancestor_map(idx, N).astype(np.float64))
        pf.close()


def test_inplace_mode_of_resample():
    N = 5000
    rng = np.random.default_rng(8)
    logw = rng.normal(0, 2, N).astype(np.float32)
    p = np.zeros((N, 6))
    p[:, 0] = np.arange(N)
    pf = _pf(particle_count=N, inplace_reference=True)
    pf.set_poses(p)
    pf.set_log_weights(logw)
    anc = pf.resample(u=0.123, return_ancestors=True, roughen=False)
    np.testing.assert_array_equal(pf.poses()[:, 0], o.ancestor_map(anc, N).astype(np.float64))
    pf.close()


# --------------------------------------------------------------------------- a6 residual resample
def test_residual_resample_strict_is_the_reference(golden):
    """AUVPF_RESIDUAL + AUVPF_RS_STRICT_REF == the UNMODIFIED reference residual_resample
    (resampling.py:27-76, the resampler auv_pf.resample calls) on the same quantised weights and
    the same random(N - k) draw, bit for bit; with the in-place map it is auv_pf.resample()."""
    g = golden("residual_quantised")
    for n in ("n100", "n1000", "n4096", "n20000"):
        logw, qq, U, idx = g[n + "_logw"], g[n + "_qq"], g[n + "_U"], g[n + "_idx"]
        N = len(logw)
        p = np.zeros((N, 6))
        p[:, 0] = np.arange(N)
        p[:, 1] = 0.5 * np.arange(N)
        for inplace in (False, True):
            pf = _pf(particle_count=N, resampler="residual", strict_reference=True, inplace_reference=inplace)
            pf.set_poses(p)
            pf.set_log_weights(logw)
            anc = pf.resample(u=U, return_ancestors=True, roughen=False)
            np.testing.assert_array_equal(pf.weights_q(), qq)
            np.testing.assert_array_equal(anc, idx)
            src = o.ancestor_map(idx, N) if inplace else idx
            np.testing.assert_array_equal(pf.poses()[:, 0:2], p[src, 0:2])
            pf.close()


@pytest.mark.parametrize("N,scale", [(1, 1.0), (100, 2.0), (4097, 6.0), (65536, 1.0), (1 << 20, 3.0)])
def test_residual_resample_default(N, scale):
    """Default residual mode: deterministic floor(N w) copies == the reference's (:61-66), the
    rest == the fixed-point multinomial of the oracle (residual_resample_fixed), bit for bit."""
    rng = np.random.default_rng(1000 + N)
    logw = (rng.normal(0, scale, N) - 30.0).astype(np.float32)
    U = rng.uniform(size=N)
    p = np.zeros((N, 6))
    p[:, 0] = np.arange(N)
    pf = _pf(particle_count=N, resampler="residual")
    pf.set_poses(p)
    pf.set_log_weights(logw)
    anc = pf.resample(u=U, return_ancestors=True, roughen=False)
    qq = pf.weights_q()
    np.testing.assert_array_equal(qq, c_oracle.quantise_logw(logw)[0])
    head, copies = o.residual_copies(o.weights_from_q(qq))
    np.testing.assert_array_equal(anc[:len(head)], head)
    np.testing.assert_array_equal(anc, o.residual_resample_fixed(qq, U))
    np.testing.assert_array_equal(pf.poses()[:, 0], p[anc, 0])
    # library-drawn uniforms (Philox): same head, tail only on particles with a fractional part
    pf.set_poses(p)
    pf.set_log_weights(logw)
    anc2 = pf.resample(return_ancestors=True, roughen=False)
    np.testing.assert_array_equal(anc2[:len(head)], head)
    frac = N * o.weights_from_q(qq) - copies
    assert np.all(frac[anc2[len(head):]] > 0)
    pf.close()


def test_residual_resample_batched_filters(golden):
    """Residual resampling of several filters in one handle: every filter equals its own
    single-filter result -- fixed-point default mode vs the oracle, strict mode vs the unmodified
    reference's golden indices."""
    g = golden("residual_quantised")
    names = ("n1000", "n1000", "n1000")
    N, F = 1000, 3
    rng = np.random.default_rng(55)
    logw = np.stack([g["n1000_logw"], g["n1000_logw"][::-1], (rng.normal(0, 1.5, N) - 20).astype(np.float32)])
    U = np.stack([g["n1000_U"], rng.uniform(size=N), rng.uniform(size=N)])
    p = np.zeros((F * N, 6))
    p[:, 0] = np.arange(F * N)
    for strict in (False, True):
        pf = _pf(particle_count=N, n_filters=F, resampler="residual", strict_reference=strict)
        pf.set_poses(p)
        pf.set_log_weights(logw.reshape(-1))
        anc = pf.resample(u=U, return_ancestors=True, roughen=False).reshape(F, N)
        qq = pf.weights_q().reshape(F, N)
        for f in range(F):
            np.testing.assert_array_equal(qq[f], o.quantise_logw(logw[f])[0])
            want = o.residual_resample(o.weights_from_q(qq[f]), U[f]) if strict else o.residual_resample_fixed(qq[f], U[f])
            np.testing.assert_array_equal(anc[f], want)
            np.testing.assert_array_equal(pf.poses()[f * N:(f + 1) * N, 0], p[f * N + anc[f], 0])
        if strict:
            np.testing.assert_array_equal(anc[0], g["n1000_idx"])            # the unmodified reference
        pf.close()


# --------------------------------------------------------------------------- a10 estimate
def test_estimate_vs_reference_golden(golden):
    g = golden("auv_pf_cycle")
    N = g["init_noise"].shape[0]
    pf = _pf(particle_count=N)
    for c in range(g["gps"].shape[0]):
        pf.set_poses(g["poses_res"][c])
        mean, cov = pf.mean_cov()
        ave, cov_ref, quat = o.estimate(pf.poses())
        np.testing.assert_allclose(mean[0], ave, atol=1e-9)
        np.testing.assert_allclose(cov[0], cov_ref, atol=1e-9)
        e = g["estimate"][c]
        np.testing.assert_allclose(mean[0, 0:3], e[0:3], atol=5e-6)
        np.testing.assert_allclose(o.covariance36(cov[0]), e[7:43], atol=1e-5)
        assert cov[0][2, 0] == 0.0 and cov[0][2, 1] == 0.0 and cov[0][1, 0] == cov[0][0, 1]
    pf.close()


# --------------------------------------------------------------------------- batched filters
def test_batched_filters_match_single_filters():
    F, M, B = 4, 3000, 64
    dem = world.make_dem(200, 2.0, seed=9)
    beams = world.make_beams(B)
    rng = np.random.default_rng(17)
    poses = np.concatenate([_cloud(rng, M, (100.0 + 30 * f, 150.0 + 20 * f), 2.0) for f in range(F)])
    meas = np.stack([rng.uniform(15, 40, B) for _ in range(F)]).astype(np.float32)
    us = rng.uniform(size=F)
    pf = _pf(particle_count=M, n_filters=F, range_max=world.RANGE_MAX)
    pf.set_dem(dem)
    pf.set_beams(beams)
    pf.set_poses(poses)
    pf.update(meas, return_weights=False)
    logw = pf.log_weights()
    anc = pf.resample(u=us, return_ancestors=True, roughen=False)
    after = pf.poses()
    mean, cov = pf.mean_cov()
    for f in range(F):
        sl = slice(f * M, (f + 1) * M)
        want = c_oracle.update_mbes(dem, poses[sl], np.eye(4), beams, meas[f].astype(np.float64), 0.01, 0.0,
                                    world.RANGE_MAX)["logw"]
        np.testing.assert_allclose(logw[sl], want, rtol=LOGW_RTOL)
        qq, m, A = c_oracle.quantise_logw(logw[sl])
        a = c_oracle.systematic(o.weights_from_q(qq), float(us[f]))
        np.testing.assert_array_equal(anc[sl], a)
        np.testing.assert_array_equal(after[sl, 0:2], poses[sl][a, 0:2])
        ave, cref, _ = o.estimate(after[sl])
        np.testing.assert_allclose(mean[f], ave, atol=1e-9)
    # per-filter odometry
    od = [world.Odom(np.array([1.0 + f, 0, 0]), 0.1 * f, world.quat_from_rpy(0.01 * f, 0.0, 0.0), -5.0 - f, 0.1)
          for f in range(F)]
    before = pf.poses()
    pf.predict(od)
    got = pf.poses()
    for f in range(F):
        sl = slice(f * M, (f + 1) * M)
        want = o.motion_pred(before[sl], od[f].v, od[f].wz, od[f].quat, od[f].z, od[f].dt, [0.0] * 6, np.zeros((M, 6)))
        np.testing.assert_allclose(got[sl, 0:2], want[:, 0:2], atol=1e-5)
        assert np.all(got[sl, 2] == np.float32(-5.0 - f))
    pf.close()


# --------------------------------------------------------------------------- G5 cycle
def test_c1_cycle_lockstep_and_convergence():
    """C1 (100 particles x 256 beams, 200 cycles).  Lock-step: every cycle the oracle starts
    from the GPU's state and must reproduce propagate, log-weights, ancestors (bit-exact given
    the GPU log-weights) and the roughened state.  Free-running: the filter tracks the truth."""
    c = world.CONFIGS["C1"]
    dem = _dem(c["dem_n"], c["res"])
    B, N = c["beams"], c["particles"]
    beams = world.make_beams(B)
    P = world.MBES_PARAMS
    ext = c["dem_n"] * c["res"]
    traj = world.Trajectory(ext)
    rng = np.random.default_rng(4)
    init_noise = rng.normal(size=(N, 6)).astype(np.float32)
    pf = _pf(particle_count=N, init_covariance=P["init_covariance"], motion_covariance=P["motion_covariance"],
             resampling_noise_covariance=P["resampling_noise_covariance"], measurement_std=P["measurement_std"],
             range_min=world.RANGE_MIN, range_max=world.RANGE_MAX, init_mean=traj.truth, init_noise=init_noise)
    pf.set_dem(dem)
    pf.set_beams(beams)
    rng5, rng6 = np.random.default_rng(5), np.random.default_rng(6)
    errs = []
    n_fragile, max_dlogw_fragile = 0, 0.0
    for cyc in range(200):
        odoms, truth = traj.cycle()
        state = pf.poses()
        for od in odoms:
            nz = rng5.normal(size=(N, 6)).astype(np.float32)
            pf.predict(od, noise=nz)
            state = o.motion_pred(state, od.v, od.wz, od.quat, od.z, od.dt, P["motion_covariance"], nz.astype(np.float64))
        got = pf.poses()
        assert np.abs(got[:, 0:2] - state[:, 0:2]).max() < 1e-4
        assert np.abs(_dyaw(got[:, 5], state[:, 5])).max() < 1e-5
        r_true = c_oracle.update_mbes(dem, truth[None, :], np.eye(4), beams, None, 1.0, 0.5, world.RANGE_MAX,
                                      want_ranges=True)["ranges"][0]
        meas = (r_true + world.ping_noise(B, cyc)).astype(np.float32)
        pf.update(meas, return_weights=False)
        logw = pf.log_weights()
        ref = c_oracle.update_mbes(dem, got, np.eye(4), beams, meas.astype(np.float64), P["measurement_std"],
                                   world.RANGE_MIN, world.RANGE_MAX, want_ranges=True, want_diag=True)
        want = ref["logw"]
        # no particle is excluded: one with a ray whose first-hit decision hinges on < 2 mm gets exactly the
        # slack the oracle's own +-5 mm bracket of that ray justifies (tests/fragile.py), and is counted
        extra, fr = fragile.logw_bound(dem, got, np.eye(4), beams, meas, P["measurement_std"], world.RANGE_MIN,
                                       world.RANGE_MAX, ref)
        n_fragile += int((extra > 0).sum())
        d = np.abs(logw.astype(np.float64) - want)
        assert np.all(d <= LOGW_RTOL * np.abs(want) + extra), (cyc, float(np.max(d - extra)))
        if (extra > 0).any():
            max_dlogw_fragile = max(max_dlogw_fragile, float(d[extra > 0].max()))
        u = float(rng6.uniform())
        rn = rng5.normal(size=(N, 6)).astype(np.float32)
        anc = pf.resample(u=u, noise=rn, return_ancestors=True)
        qq = o.quantise_logw(logw)[0]
        np.testing.assert_array_equal(pf.weights_q(), qq)
        a_ref = o.systematic_resample(o.weights_from_q(qq), u)
        np.testing.assert_array_equal(anc, a_ref)
        want_state = o.add_noise(got[a_ref], P["resampling_noise_covariance"], rn.astype(np.float64))
        after = pf.poses()
        np.testing.assert_allclose(after[:, 0:2], want_state[:, 0:2], atol=1e-5)
        assert np.abs(_dyaw(after[:, 5], want_state[:, 5])).max() < 1e-5
        mean, cov = pf.mean_cov()
        errs.append(np.hypot(mean[0, 0] - truth[0], mean[0, 1] - truth[1]))
    assert np.mean(errs[100:]) < 1.0, np.mean(errs[100:])
    assert n_fragile <= 0.02 * 200 * N, n_fragile
    fragile.report({"label": "C1 200 cycles", "particle_updates": 200 * N, "with_fragile_beam": n_fragile,
                    "max_dlogw_fragile_owners": max_dlogw_fragile})
    pf.close()


def test_estimate_options_and_pose_array():
    """weighted / circular-yaw / full-covariance estimate against the oracle, the reference form
    unchanged by default, and the PoseArray of loc_loop (auv_pf.py:264-285) for a strided subsample."""
    M, F = 20000, 2
    rng = np.random.default_rng(8)
    poses = np.column_stack([1500.0 + rng.normal(0, 3, F * M), -800.0 + rng.normal(0, 2, F * M), -5 + rng.normal(0, 0.1, F * M),
                             rng.normal(0, 0.02, F * M), rng.normal(0, 0.02, F * M), rng.normal(np.pi, 0.2, F * M)])   # yaw straddles +-pi
    poses[:, 2:6] = poses[:, 2:6].astype(np.float32)
    logw = (rng.normal(0, 2.0, F * M) - 300).astype(np.float32)
    logw[5] = -np.inf
    pf = _pf(particle_count=M, n_filters=F)
    pf.set_poses(poses)
    pf.set_log_weights(logw)
    m0, c0 = pf.mean_cov()
    for f in range(F):
        sl = slice(f * M, (f + 1) * M)
        ave, cref, _ = o.estimate(poses[sl])
        np.testing.assert_allclose(m0[f], ave, atol=1e-9)
        np.testing.assert_allclose(c0[f], cref, atol=1e-9)
        for weighted, circ, full in ((True, False, False), (False, True, False), (True, True, True), (False, False, True)):
            m, c = pf.estimate(weighted=weighted, circular_yaw=circ, full_cov=full)
            wm, wc = o.estimate_options(poses[sl], logw[sl] if weighted else None, circ, full)
            np.testing.assert_allclose(m[f, 0:5], wm[0:5], rtol=0, atol=1e-8)
            assert abs(_dyaw(m[f, 5], wm[5])) < 1e-9
            np.testing.assert_allclose(c[f], wc, rtol=1e-8, atol=1e-9)
        # the arithmetic mean of wrapped yaws is meaningless for this cloud, the circular one is not
        mc, _ = pf.estimate(circular_yaw=True)
        assert abs(_dyaw(mc[f, 5], np.pi)) < 0.01 and abs(_dyaw(m0[f, 5], np.pi)) > 1.0
    # PoseArray: strided subsample, quaternion_from_euler on the device
    for max_poses, stride in ((500, 0), (100000, 0), (64, 7)):
        pa = pf.pose_array(max_poses=max_poses, stride=stride)
        st = stride if stride > 0 else -(-M // min(max_poses, M))
        n = min(-(-M // st), max_poses)
        assert pa.shape == (F, n, 7)
        for f in range(F):
            src = poses[f * M:(f + 1) * M][::st][:n]
            np.testing.assert_allclose(pa[f, :, 0:3], src[:, 0:3], atol=1e-9)
            q = np.array([tf_stub.quaternion_from_euler(r, p, y) for r, p, y in src[:50, 3:6]])
            np.testing.assert_allclose(pa[f, :50, 3:7], q, atol=1e-12)
    pf.close()


@pytest.mark.parametrize("M,F,k", [(20000, 2, 500), (1 << 20, 1, 2000), (70000, 3, 70000), (4097, 1, 1)])
def test_pose_array_topk_is_the_exact_selection(M, F, k):
    """SURVEY 8f-1, the top-k form of the PoseArray subsample: the k particles of largest weight per filter, chosen
    on the device, against numpy's stable ordering (weight descending, index ascending among equals, NaN as -inf)
    -- with heavy ties, -inf, NaN and +0 / -0 in the weights."""
    rng = np.random.default_rng(M + k)
    poses = np.column_stack([rng.normal(100, 30, F * M), rng.normal(-50, 30, F * M), rng.normal(-5, 1, F * M),
                             rng.normal(0, 0.05, F * M), rng.normal(0, 0.05, F * M), rng.uniform(-3.1, 3.1, F * M)])
    poses[:, 2:6] = poses[:, 2:6].astype(np.float32)
    logw = np.round(rng.normal(0, 3.0, F * M), 1).astype(np.float32) - 100       # one decimal: thousands of exact ties
    logw[rng.choice(F * M, F * M // 50, replace=False)] = -np.inf
    logw[rng.choice(F * M, max(1, F * M // 200), replace=False)] = np.nan
    logw[0:4] = [0.0, -0.0, 0.0, -0.0]                                          # the largest, equal under IEEE compare? (no: see below)
    pf = _pf(particle_count=M, n_filters=F)
    pf.set_poses(poses)
    pf.set_log_weights(logw)
    got_p, got_w, got_i = pf.pose_array_topk(k)
    assert got_p.shape == (F, k, 7) and got_w.shape == (F, k) and got_i.shape == (F, k)
    for f in range(F):
        lw = logw[f * M:(f + 1) * M]
        key = np.where(np.isnan(lw), -np.inf, lw).astype(np.float64)
        # the device orders the weights by their bit pattern: +0 above -0 (a total order; IEEE compares them equal)
        key = np.where((lw == 0) & np.signbit(lw), -1e-300, key)
        want = np.lexsort((np.arange(M), -key))[:k]
        np.testing.assert_array_equal(np.sort(got_i[f]), np.sort(want))                  # the exact set
        kk = np.where(np.isnan(got_w[f]), -np.inf, got_w[f])
        assert np.all(kk[:-1] >= kk[1:])                                                 # the facade's order
        src = poses[f * M:(f + 1) * M][got_i[f]]
        np.testing.assert_allclose(got_p[f, :, 0:3], src[:, 0:3], atol=1e-9)
        np.testing.assert_array_equal(got_w[f][~np.isnan(got_w[f])], lw[got_i[f]][~np.isnan(got_w[f])])
        q = np.array([tf_stub.quaternion_from_euler(r, p_, y) for r, p_, y in src[:40, 3:6]])
        np.testing.assert_allclose(got_p[f, :40, 3:7], q, atol=1e-12)
    pf.close()


def test_kld_bins_and_a_resample_that_changes_the_particle_count():
    """SURVEY 8f-4, KLD-style adaptation: the occupied-bin count of the cloud (device hash set) against numpy's unique;
    Fox's bound from the library against the oracle's; and the resample that applies it -- n_new positions over the CDF
    of the current particles -- bit-exact against the oracle, down AND up again, with the cycle continuing at the new
    count.  A held-back or degenerate resample leaves the count alone; batched filters refuse."""
    M = 200000
    rng = np.random.default_rng(31)
    poses = np.column_stack([rng.normal(420.0, 1.5, M), rng.normal(-77.0, 0.8, M), np.full(M, -5.0), np.zeros(M), np.zeros(M),
                             rng.normal(3.1, 0.1, M)])                  # yaw straddles +pi: bins on both sides of the shift
    poses[:, 2:6] = poses[:, 2:6].astype(np.float32)
    pf = _pf(particle_count=M, resampling_noise_covariance=[0.0] * 6, measurement_std=1.0, range_min=world.RANGE_MIN,
             range_max=world.RANGE_MAX)
    pf.set_poses(poses)
    for bxy, byaw in ((0.5, np.radians(5.0)), (0.05, np.radians(1.0)), (7.0, 1.0)):
        k = int(pf.kld_bins(bxy, byaw)[0])
        assert k == o.kld_bins(poses, bxy, byaw)
        for eps, z in ((0.05, 2.3263478740408408), (0.01, 1.6448536269514722)):
            assert int(pf._lib.auvpf_kld_target(k, eps, z, 100, 10 ** 9)) == o.kld_target(k, eps, z, 100, 10 ** 9)
    n1 = int(pf.kld_particle_count(0.5, np.radians(5.0), n_min=1000)[0])
    assert 1000 <= n1 < M and pf.capacity == M
    # ---- down to n1: ancestors, state, count
    logw = (rng.normal(0, 3.0, M) - 200).astype(np.float32)
    pf.set_log_weights(logw)
    anc = pf.resample(u=0.37, return_ancestors=True, roughen=False, new_count=n1)
    qq = o.quantise_logw(logw)[0]
    want = o.systematic_resample_to(o.weights_from_q(qq), 0.37, n1)
    assert pf.pc == n1 and len(anc) == n1
    np.testing.assert_array_equal(anc, want)
    p1 = pf.poses()
    assert p1.shape == (n1, 6)
    np.testing.assert_array_equal(p1, poses[want])
    # ---- the filter goes on at n1: an MBES update on the smaller cloud ...
    dem = _dem(500, 2.0)
    beams = world.make_beams(64).astype(np.float32).astype(np.float64)
    pf.set_dem(dem)
    pf.set_beams(beams)
    truth = np.array([[420.0, -77.0 + 500.0, -5.0, 0.0, 0.0, 0.3]])
    shifted = p1.copy()
    shifted[:, 1] += 500.0                                   # (the cloud was placed off this map: move it onto it)
    pf.set_poses(shifted)
    meas = c_oracle.update_mbes(dem, truth, np.eye(4), beams, None, 1.0, 0.5, world.RANGE_MAX, want_ranges=True)["ranges"][0].astype(np.float32)
    pf.update_mbes(meas, return_weights=False)
    got = pf.log_weights()
    ref = c_oracle.update_mbes(dem, shifted[:4000], np.eye(4), beams, meas.astype(np.float64), 1.0, world.RANGE_MIN,
                               world.RANGE_MAX)["logw"]
    assert got.shape == (n1,)
    lognorm = -len(beams) * np.log(np.sqrt(2 * np.pi))
    tol = LOGW_RTOL * (np.abs(ref - lognorm) + abs(lognorm))
    assert np.all(np.abs(got[:4000].astype(np.float64) - ref) <= tol)
    pf.set_poses(p1)
    # ---- ... weights, an ordinary resample, the estimate
    logw1 = (rng.normal(0, 2.0, n1) - 50).astype(np.float32)
    pf.set_log_weights(logw1)
    anc1 = pf.resample(u=0.81, return_ancestors=True, roughen=False)
    np.testing.assert_array_equal(anc1, o.systematic_resample(o.weights_from_q(o.quantise_logw(logw1)[0]), 0.81))
    p2 = pf.poses()
    em, ec, _ = o.estimate(p2)
    m, c = pf.mean_cov()
    np.testing.assert_allclose(m[0], em, atol=1e-9)
    # ---- up again to the capacity, with the gate holding one attempt back first
    logw2 = np.zeros(n1, np.float32)                       # uniform weights: N_eff = n1, the gate holds the resample back
    pf.set_log_weights(logw2)
    pf.set_neff_gate(0.5)
    pf.resample(u=0.5, roughen=False, new_count=M)
    assert pf.pc == n1 and pf.resample_skipped()[0]
    np.testing.assert_array_equal(pf.poses(), p2)
    pf.set_neff_gate(0.0)
    logw2 = (rng.normal(0, 1.0, n1)).astype(np.float32)
    pf.set_log_weights(logw2)
    anc2 = pf.resample(u=0.12, return_ancestors=True, roughen=False, new_count=M)
    assert pf.pc == M
    want2 = o.systematic_resample_to(o.weights_from_q(o.quantise_logw(logw2)[0]), 0.12, M)
    np.testing.assert_array_equal(anc2, want2)
    np.testing.assert_array_equal(pf.poses(), p2[want2])
    # ---- degenerate weights: error, nothing changes; more than the capacity: refused
    pf.set_log_weights(np.full(M, -np.inf, np.float32))
    with pytest.raises(Exception):
        pf.resample(u=0.3, roughen=False, new_count=5000)
    assert pf.pc == M
    with pytest.raises(Exception):
        pf.resample(u=0.3, roughen=False, new_count=M + 1)
    pf.close()
    pf2 = _pf(particle_count=1000, n_filters=2)
    with pytest.raises(Exception):
        pf2.resample(new_count=500)
    pf2.close()


# --------------------------------------------------------------------------- full-size properties
def test_c3_full_size_properties():
    """At BASELINE C3 size (1M x 512 on the 4000^2 DEM) the oracle is too slow for a full
    comparison; check size-independent properties and an oracle-checked subsample."""
    c = world.CONFIGS["C3"]
    dem = _dem(c["dem_n"], c["res"])
    B, N = c["beams"], c["particles"]
    beams = world.make_beams(B)
    P = world.MBES_PARAMS
    truth = np.array([600.0, 600.0, -5.0, 0.01, -0.02, 0.3])
    pf = _pf(particle_count=N, init_covariance=P["init_covariance"], measurement_std=P["measurement_std"],
             resampling_noise_covariance=[0.0] * 6, range_min=world.RANGE_MIN, range_max=world.RANGE_MAX,
             init_mean=truth, seed=11)
    pf.set_dem(dem)
    pf.set_beams(beams)
    r_true = c_oracle.update_mbes(dem, truth[None, :], np.eye(4), beams, None, 1.0, 0.5, world.RANGE_MAX,
                                  want_ranges=True)["ranges"][0]
    meas = r_true.astype(np.float32)
    poses = pf.poses()
    pf.update(meas, return_weights=False)
    logw = pf.log_weights()
    sub = np.random.default_rng(0).choice(N, 2000, replace=False)
    want = c_oracle.update_mbes(dem, poses[sub], np.eye(4), beams, meas.astype(np.float64), P["measurement_std"],
                                world.RANGE_MIN, world.RANGE_MAX)["logw"]
    np.testing.assert_allclose(logw[sub], want, rtol=LOGW_RTOL)
    # the best particles are the ones nearest the truth
    best = np.argsort(logw)[-100:]
    assert np.hypot(poses[best, 0] - truth[0], poses[best, 1] - truth[1]).mean() < 1.5
    anc = pf.resample(u=0.5, return_ancestors=True, roughen=False)
    qq = pf.weights_q()
    assert int(qq.sum(dtype=np.uint64)) == 1 << 52
    assert np.all(np.diff(anc) >= 0) and anc.min() >= 0 and anc.max() < N
    counts = np.bincount(anc, minlength=N)
    expect = o.weights_from_q(qq) * N
    assert np.abs(counts - expect).max() <= 1.0 + 1e-9           # systematic: |offspring - N w| < 1
    after = pf.poses()
    np.testing.assert_array_equal(after[:, 0], poses[anc, 0])
    # idempotence: resampling uniform weights with the same u leaves the set unchanged
    pf.set_log_weights(np.zeros(N, np.float32))
    anc2 = pf.resample(u=0.5, return_ancestors=True, roughen=False)
    np.testing.assert_array_equal(anc2, np.arange(N, dtype=np.int32))
    np.testing.assert_array_equal(pf.poses(), after)
    pf.close()


def test_replay_equals_cycle_by_cycle():
    """auvpf_replay (a recorded odometry / ping sequence run without the host in the loop) is the
    same filter as predict_many + update + resample + mean_cov called cycle by cycle: poses and
    estimates, two filters with their own odometry rows, the N_eff gate active."""
    from smarc_navigation_b200.filter import _odom_row
    dem = _dem(300, 2.0)
    B, M, F, K, P = 96, 3000, 2, 7, 4
    beams = world.make_beams(B)
    MP = world.MBES_PARAMS
    trajs = [world.Trajectory(600.0, seed=2 + f, predicts_per_cycle=P, offset=(3.0 * f, -2.0 * f)) for f in range(F)]
    odoms = np.zeros((K, P, F, 11))
    ranges = np.zeros((K, F, B), np.float32)
    for c in range(K):
        for f, tr in enumerate(trajs):
            od, truth = tr.cycle()
            odoms[c, :, f] = [_odom_row(x, None) for x in od]
            r = c_oracle.update_mbes(dem, truth[None, :], np.eye(4), beams, None, 1.0, 0.5, world.RANGE_MAX,
                                     want_ranges=True)["ranges"][0]
            ranges[c, f] = r + world.ping_noise(B, c + 100 * f)
    ranges[3, 1, ::2] = np.nan                  # half of one ping unusable
    kw = dict(particle_count=M, n_filters=F, init_covariance=MP["init_covariance"],
              motion_covariance=MP["motion_covariance"], resampling_noise_covariance=MP["resampling_noise_covariance"],
              measurement_std=MP["measurement_std"], range_min=world.RANGE_MIN, range_max=world.RANGE_MAX,
              init_mean=np.array([0.3 * 600.0, 0.3 * 600.0, -5.0, 0, 0, 0]), seed=9,
              resample_neff_ratio=0.5)
    a, b = _pf(**kw), _pf(**kw)
    for pf in (a, b):
        pf.set_dem(dem)
        pf.set_beams(beams)
    np.testing.assert_array_equal(a.poses(), b.poses())
    means, covs, skipped = [], [], []
    for c in range(K):
        a.predict_many(odoms[c])
        a.update(ranges[c], return_weights=False)
        a.resample(wait=False)
        skipped.append(a.resample_skipped())
        m, cv = a.mean_cov()
        means.append(m)
        covs.append(cv)
    got_m, got_c = b.replay(odoms, ranges)
    np.testing.assert_array_equal(b.poses(), a.poses())
    # (the estimate's block partials meet in float64 atomics: equal up to the order of the additions)
    np.testing.assert_allclose(got_m, np.stack(means), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got_c, np.stack(covs), rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(b.resample_skipped(), skipped[-1])
    # enqueue-only form, caller's uniforms, no roughening: still the same as the calls one by one
    u = np.random.default_rng(1).uniform(size=(K, F))
    for c in range(K):
        a.predict_many(odoms[c])
        a.update(ranges[c], return_weights=False)
        a.resample(u=u[c], roughen=False, wait=False)
    assert b.replay(odoms, ranges, u=u, roughen=False, estimates=False) is None
    np.testing.assert_array_equal(b.poses(), a.poses())
    a.close()
    b.close()


def test_replay_cycle_vs_oracle():
    """The oracle leg of auvpf_replay: single-cycle replays in lock-step with the CPU oracle.  Each
    cycle starts from the GPU's state; the oracle runs the Philox-noised motion model, the MBES
    log-weights, the exact normalisation + systematic ancestors (given the GPU's float32 log-weights),
    the roughening and update_loc_pose's estimate, and every stage the replay exposes is compared."""
    from smarc_navigation_b200.filter import _odom_row
    dem = _dem(300, 2.0)
    B, M, K, P, seed = 96, 4000, 4, 5, 21
    beams = world.make_beams(B)
    MP = world.MBES_PARAMS
    tr = world.Trajectory(600.0, seed=2, predicts_per_cycle=P)
    pf = _pf(particle_count=M, init_covariance=MP["init_covariance"], motion_covariance=MP["motion_covariance"],
             resampling_noise_covariance=MP["resampling_noise_covariance"], measurement_std=MP["measurement_std"],
             range_min=world.RANGE_MIN, range_max=world.RANGE_MAX, init_mean=tr.truth, seed=seed)
    pf.set_dem(dem)
    pf.set_beams(beams)
    ids = np.arange(M, dtype=np.int64)
    rng = np.random.default_rng(8)
    for c in range(K):
        od, truth = tr.cycle()
        rows = np.array([_odom_row(x, None) for x in od]).reshape(1, P, 1, 11)
        r = c_oracle.update_mbes(dem, truth[None, :], np.eye(4), beams, None, 1.0, 0.5, world.RANGE_MAX,
                                 want_ranges=True)["ranges"][0]
        meas = (r + world.ping_noise(B, c)).astype(np.float32)
        u = float(rng.uniform())
        s0 = pf.poses()
        got_m, got_c = pf.replay(rows, meas.reshape(1, 1, B), u=[[u]])
        # oracle: motion model with the device RNG contract (stream 2 = predict, step = running predict count)
        s1 = s0.copy()
        for t, x in enumerate(od):
            n3 = o.philox_normals(seed, 2, c * P + t, ids, 1)
            nz = np.zeros((M, 6))
            nz[:, 0], nz[:, 1], nz[:, 5] = n3[:, 0], n3[:, 1], n3[:, 2]
            s1 = o.motion_pred(s1, x.v, x.wz, x.quat, x.z, x.dt, MP["motion_covariance"], nz)
        ref = c_oracle.update_mbes(dem, s1, np.eye(4), beams, meas.astype(np.float64), MP["measurement_std"],
                                   world.RANGE_MIN, world.RANGE_MAX, want_ranges=True, want_diag=True)
        logw = pf.log_weights()
        extra, _ = fragile.logw_bound(dem, s1, np.eye(4), beams, meas, MP["measurement_std"], world.RANGE_MIN,
                                      world.RANGE_MAX, ref)
        # the oracle's poses differ from the device's by float32 rounding of the step (<= 1e-5 m): its effect on
        # sum e^2 / 2 is bounded by sum |e| * 2e-5
        pose_slack = 2e-5 * np.sqrt(2.0 * np.maximum(-(ref["logw"] + B * np.log(np.sqrt(2 * np.pi))), 0.0) * B) + 1e-6
        assert np.all(np.abs(logw - ref["logw"]) <= LOGW_RTOL * np.abs(ref["logw"]) + extra + pose_slack)
        qq = o.quantise_logw(logw)[0]
        np.testing.assert_array_equal(pf.weights_q(), qq)
        anc = o.systematic_resample(o.weights_from_q(qq), u)
        want = o.add_noise(s1[anc], MP["resampling_noise_covariance"], o.philox_pose_normals(seed, 3, c, ids))
        after = pf.poses()
        np.testing.assert_allclose(after[:, 0:2], want[:, 0:2], atol=1e-4)
        assert np.abs(_dyaw(after[:, 5], want[:, 5])).max() < 1e-5
        em, ec, _ = o.estimate(after)
        np.testing.assert_allclose(got_m[0, 0], em, atol=1e-9)
        np.testing.assert_allclose(got_c[0, 0], ec, atol=1e-9)
    pf.close()


def test_device_pointer_inputs():
    """Torch tensors pass zero-copy through data_ptr()."""
    import ctypes as C
    import torch
    from smarc_navigation_b200 import _lib
    N = 10000
    pf = _pf(particle_count=N)
    lw = torch.randn(N, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    L = _lib.load()
    _lib.check(L.auvpf_set_logw(pf._h, C.c_void_p(lw.data_ptr())))
    anc_t = torch.empty(N, device="cuda", dtype=torch.int32)
    u = np.array([0.25])
    _lib.check(L.auvpf_resample(pf._h, 0, 0, C.c_void_p(u.ctypes.data), C.c_void_p(anc_t.data_ptr()), None, None))
    pf.sync()
    qq = o.quantise_logw(lw.cpu().numpy())[0]
    np.testing.assert_array_equal(anc_t.cpu().numpy(), o.systematic_resample(o.weights_from_q(qq), 0.25))
    pf.close()
