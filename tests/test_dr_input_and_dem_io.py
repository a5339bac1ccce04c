"""CPU: the dead-reckoning input model (dr_node.py:165-201, sam_mm.py:30-125) and the DEM file format."""
import os

import numpy as np
import pytest

from smarc_navigation_b200 import dem_io, dr_input, world


def test_sam_motion_matches_the_unmodified_reference(golden):
    g = golden("sam_motion")
    np.testing.assert_allclose(dr_input.sam_motion(g["control"]), g["nudot"], rtol=1e-13, atol=1e-15)
    ref = "/root/reference/sam_dead_reckoning/scripts/sam_mm.py"
    if os.path.isfile(ref):                         # live re-check where the checkout exists
        import importlib.util
        spec = importlib.util.spec_from_file_location("sam_mm_ref", ref)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sam = mod.SAM()
        for c in g["control"][:8]:
            np.testing.assert_allclose(dr_input.sam_motion(c), sam.motion(list(c)), rtol=1e-13, atol=1e-15)


def test_dvl_gate_is_dr_nodes():
    """dr_node.py:179-182: fresh, |v_y| < 0.2, |v_x| < 1.5, v_x > -0.1 -- each bound on its own."""
    ok = np.array([1.0, 0.0, 0.05])
    cases = [(0.1, ok, True), (0.21, ok, False), (0.25, ok, False),                     # stale
             (0.1, [1.0, 0.2, 0.0], False), (0.1, [1.0, -0.2, 0.0], False), (0.1, [1.0, 0.199, 0.0], True),
             (0.1, [1.5, 0.0, 0.0], False), (0.1, [1.499, 0.0, 0.0], True), (0.1, [-0.1, 0.0, 0.0], False),
             (0.1, [-0.099, 0.0, 0.0], True), (0.1, [0.5, 0.0, 9.0], True)]                # v_z is not gated
    for age, v, want in cases:
        assert bool(dr_input.dvl_gate(10.0, 10.0 - age, np.asarray(v, float))) is want, (age, v)
    lv, used = dr_input.linear_velocity(10.0, 9.95, ok, [800.0, 0.05])
    np.testing.assert_array_equal(lv, ok)
    assert bool(used)
    lv, used = dr_input.linear_velocity(10.0, 9.0, ok, [800.0, 0.05])                     # stale DVL: motion model
    acc = dr_input.sam_motion([800.0, 0.05])
    np.testing.assert_allclose(lv, [acc[0] * 0.02, -acc[1] * 0.02, 0.0])
    assert not bool(used)


def test_batched_rows_equal_message_by_message():
    rng = np.random.default_rng(1)
    K, P, F = 6, 5, 3
    t = 100.0 + 0.02 * np.arange(K * P).reshape(K, P, 1) + np.zeros((1, 1, F))
    t_dvl = t - rng.uniform(0.0, 0.4, (K, P, F))
    v = rng.normal([0.9, 0.0, 0.0], [0.5, 0.15, 0.02], (K, P, F, 3))
    u = np.stack([rng.uniform(0, 1500, (K, P, F)), rng.uniform(-0.1, 0.1, (K, P, F))], -1)
    quat = world.quat_from_rpy(0.01, -0.02, 0.4)
    rows, used = dr_input.odom_rows(t, t_dvl, v, u, 0.05, quat, -5.0)
    assert rows.shape == (K, P, F, 11) and 0 < used.mean() < 1
    for idx in [(0, 0, 0), (3, 2, 1), (5, 4, 2)]:
        d = dr_input.DeadReckoningInput()
        d.dvl_cb(t_dvl[idx], v[idx])
        d.u = u[idx]
        np.testing.assert_array_equal(rows[idx][0:3], d.velocity(t[idx]))
        np.testing.assert_array_equal(rows[idx][3:11], [0.05, *quat, -5.0, 0.02, 0.0])


def test_live_producer_feeds_predict_with_gated_twists():
    """Fake DVL / thruster messages through the message-by-message form: a plausible fresh DVL sample is used,
    an implausible one (v_x = 2 m/s: a bottom-lock glitch) and a stale one fall back to the motion model -- and the
    sample has the shape ``ParticleFilter.predict`` / ``AuvPfNode.predict`` consume."""
    from smarc_navigation_b200.filter import _odom_row
    d = dr_input.DeadReckoningInput(dvl_period=0.2, dr_period=0.02)
    d.thrust_cb(900.0, 0.02)
    quat = world.quat_from_rpy(0.0, 0.0, 0.1)
    d.dvl_cb(50.00, [1.0, 0.02, 0.0])
    od = d.odometry(50.02, quat, 0.01, -4.0)
    np.testing.assert_array_equal(od.v, [1.0, 0.02, 0.0])
    d.dvl_cb(50.04, [2.0, 0.0, 0.0])                                  # glitch: gated out
    od2 = d.odometry(50.06, quat, 0.01, -4.0)
    acc = dr_input.sam_motion(d.u)
    np.testing.assert_allclose(od2.v, [acc[0] * 0.02, -acc[1] * 0.02, 0.0])
    d.dvl_cb(50.08, [1.0, 0.0, 0.0])
    od3 = d.odometry(50.40, quat, 0.01, -4.0)                         # 0.32 s old: stale
    np.testing.assert_allclose(od3.v, od2.v)
    assert (d.n_dvl, d.n_model) == (1, 2)
    row = _odom_row(od, None)
    assert row[0:4] == [1.0, 0.02, 0.0, 0.01] and row[8:10] == [-4.0, 0.02]


def test_dem_file_round_trips(tmp_path):
    dem = world.make_dem(120, 2.0, seed=3)
    p = str(tmp_path / "a.npz")
    dem_io.save_dem(p, dem)
    got = dem_io.load_dem(p)
    np.testing.assert_array_equal(got["z"], dem["z"])
    assert (got["x0"], got["y0"], got["res"], got["frame"]) == (0.0, 0.0, 2.0, "map")
    # uint16 heights: half the bytes; the dequantised float32 map is within half a step of the original, and it
    # is THE map from then on (saving it again is lossless)
    q = str(tmp_path / "q.npz")
    dem_io.save_dem(q, dem, quantised=True)
    gq = dem_io.load_dem(q)
    zq, z_min, step = dem_io.quantise(dem["z"])
    assert gq["z"].dtype == np.float32 and np.abs(gq["z"].astype(np.float64) - dem["z"]).max() <= 0.5 * step + 1e-5
    assert step < 1e-3                                   # ~0.4 mm for this relief
    q2 = str(tmp_path / "q2.npz")
    dem_io.save_dem(q2, gq, quantised=True, z_step=step)
    np.testing.assert_array_equal(dem_io.load_dem(q2)["z"], gq["z"])
    with pytest.raises(ValueError):
        dem_io.quantise(dem["z"], z_step=1e-6)
    bad = dict(dem, z=dem["z"].copy())
    bad["z"][3, 3] = np.nan
    dem_io.save_dem(p, bad)
    with pytest.raises(ValueError):
        dem_io.load_dem(p)


def test_oracle_sees_the_quantised_map(tmp_path):
    """Parity on a uint16 map: the oracle casts into the SAME dequantised float32 heights the library gets."""
    from oracle import c_oracle
    dem = world.make_dem(100, 2.0, seed=4)
    q = str(tmp_path / "q.npz")
    dem_io.save_dem(q, dem, quantised=True, z_step=0.01)          # 1 cm steps: visibly different from the original
    gq = dem_io.load_dem(q)
    pose = np.array([[100.0, 100.0, -5.0, 0.0, 0.0, 0.3]])
    beams = world.make_beams(16)
    r0 = c_oracle.update_mbes(dem, pose, np.eye(4), beams, None, 1.0, 0.5, 120.0, want_ranges=True)["ranges"]
    r1 = c_oracle.update_mbes(gq, pose, np.eye(4), beams, None, 1.0, 0.5, 120.0, want_ranges=True)["ranges"]
    assert 1e-5 < np.abs(r0 - r1).max() < 0.02
