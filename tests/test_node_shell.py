"""The ROS-free node shell (smarc_navigation_b200/node.py) driven by fake messages: same gating,
parameter names and published numbers as auv_pf.py's callbacks.  The CPU tests plug a recording
fake in place of the GPU filter (host logic only); the GPU test runs the real one."""
import types

import numpy as np
import pytest

from smarc_navigation_b200 import node as node_mod


def _ns(**kw):
    return types.SimpleNamespace(**kw)


def odom_msg(t, v=(1.0, 0.0, 0.0), wz=0.0, quat=(0.0, 0.0, 0.0, 1.0), z=-5.0):
    return _ns(header=_ns(stamp=_ns(to_sec=lambda: t)),
               twist=_ns(twist=_ns(linear=_ns(x=v[0], y=v[1], z=v[2]), angular=_ns(x=0.0, y=0.0, z=wz))),
               pose=_ns(pose=_ns(orientation=_ns(x=quat[0], y=quat[1], z=quat[2], w=quat[3]),
                                 position=_ns(x=0.0, y=0.0, z=z))))


def gps_msg(x, y):
    return _ns(pose=_ns(pose=_ns(position=_ns(x=x, y=y, z=0.0))))


PARAMS = {"particle_count": 64, "init_covariance": "[0.1, 0.1, 0.0, 0.0, 0.0, 0.0]",
          "motion_covariance": "[0.0, 0.0, 0.0, 0.0, 0.0, 1e-12]",
          "resampling_noise_covariance": "[1.0, 1.0, 0.0, 0.0, 0.0, 1e-4]", "measurement_std": 1.0}
_REQUIRED = object()


def get_param(name, default=_REQUIRED):
    if name in PARAMS:
        return PARAMS[name]
    if default is _REQUIRED:
        raise KeyError(name)                # rospy.get_param without default raises for a missing parameter
    return default


class FakeFilter(object):
    def __init__(self, **kw):
        self.kw, self.calls = kw, []

    def predict(self, odom, dt=None):
        self.calls.append(("predict", round(dt, 6)))

    def update_gps(self, xy):
        self.calls.append(("update_gps", tuple(xy)))
        return np.ones(self.kw["particle_count"])

    def resample(self, weights=None, new_count=None):
        self.calls.append(("resample", None if weights is None else len(weights)) if new_count is None else ("resample_to", new_count))

    def kld_particle_count(self, **kw):
        self.calls.append(("kld", kw["bin_xy"], kw["n_min"]))
        return np.array([4321])

    def poses(self):
        return np.zeros((self.kw["particle_count"], 6))

    def pose_array(self, max_poses=2000, stride=0):
        self.calls.append(("pose_array", max_poses))
        return np.zeros((1, min(max_poses, self.kw["particle_count"]), 7))

    def estimate(self, circular_yaw=False, **kw):
        self.calls.append(("estimate", circular_yaw))
        return self.mean_cov()

    def mean_cov(self):
        return np.array([[1.0, 2.0, -5.0, 0.01, 0.02, 0.5]]), np.arange(9.0).reshape(1, 3, 3)

    def set_scan_geometry(self, a0, da, n):
        self.calls.append(("geometry", n))

    def update_mbes(self, ranges, return_weights=True):
        self.calls.append(("update_mbes", len(ranges)))

    def close(self):
        pass


def make(now=lambda: 0.0, **kw):
    return node_mod.AuvPfNode(get_param, np.eye(4), filter_factory=FakeFilter, now=now, **kw)


def test_parameters_follow_the_reference_names_and_defaults():
    n = make()
    kw = n.pf.kw
    assert kw["particle_count"] == 64 and kw["measurement_std"] == 1.0
    assert kw["init_covariance"] == PARAMS["init_covariance"]          # covariance STRINGS pass through (auv_pf.py:40-56)
    assert kw["resampler"] == "residual" and kw["inplace_reference"] is True    # the reference's choices (:182-198)
    assert kw["resample_neff_ratio"] == 0.0     # the gate of :174-181 is commented out in the reference: off by default
    assert n.odom_frame == "sam/odom" and n.base_frame == "base_link" and n.diving is True
    PARAMS.pop("motion_covariance")
    try:
        with pytest.raises(KeyError):                                  # no default in the reference either (:40)
            make()
    finally:
        PARAMS["motion_covariance"] = "[0.0, 0.0, 0.0, 0.0, 0.0, 1e-12]"


def test_callback_gating_matches_auv_pf():
    n = make()
    n.gps_odom_cb(gps_msg(1.0, 2.0))                    # old_time == 0: nothing (auv_pf.py:126)
    assert n.pf.calls == []
    n.odom_callback(odom_msg(10.0))                     # first odom: old_time is 0 -> falsy -> no predict (:205)
    assert n.pf.calls == [] and n.old_time == 10.0
    n.odom_callback(odom_msg(10.02))
    assert n.pf.calls == [("predict", 0.02)]
    n.odom_callback(odom_msg(10.02))                    # stamp did not advance: no predict
    assert len(n.pf.calls) == 1
    n.gps_odom_cb(gps_msg(1.0, 2.0))                    # time == old_time after the odom callback (:211): gated out
    assert len(n.pf.calls) == 1
    # a fix that arrives while a newer odom stamp is being processed: time > old_time
    n.time = 10.04
    n.gps_odom_cb(gps_msg(1.0, 2.0))                    # still diving (:128)
    assert len(n.pf.calls) == 1
    n.dive_cb(_ns(data=False))
    n.gps_odom_cb(gps_msg(1.0, 2.0))
    assert n.pf.calls[1:] == [("update_gps", (1.0, 2.0)), ("resample", 64)]


def test_utm_to_map_is_applied_once_per_fix():
    seen = []

    def tf(x, y):
        seen.append((x, y))
        return x - 100.0, y + 5.0
    n = make(utm_to_map=tf)
    n.old_time, n.time, n.diving = 1.0, 2.0, False
    n.gps_odom_cb(gps_msg(150.0, 20.0))
    assert seen == [(150.0, 20.0)] and n.pf.calls[0] == ("update_gps", (50.0, 25.0))


def test_loc_loop_publishes_the_reference_layout():
    out = {}
    n = make(publish_estimate=lambda e: out.setdefault("est", e), publish_particles=lambda p: out.setdefault("p", p))
    est = n.loc_loop()
    assert out["p"].shape == (64, 7) and out["est"] is est         # PoseArray entries [x y z qx qy qz qw] (auv_pf.py:267-275)
    assert ("pose_array", 2000) in n.pf.calls                      # a bounded, device-built subsample, not pf.poses()
    np.testing.assert_array_equal(est["position"], [1.0, 2.0, -5.0])
    cov = est["covariance"]
    assert len(cov) == 36 and cov[:9] == list(np.arange(9.0)) and cov[9:] == [0.0] * 27   # stride-3 fill (auv_pf.py:249-252)
    assert est["tf"]["translation"] == [1.0, 2.0, 0.0] and est["tf"]["parent"] == "sam/odom"
    q = est["orientation"]
    from oracle import tf_stub
    np.testing.assert_allclose(q, tf_stub.quaternion_from_euler(0.01, 0.02, 0.5), atol=1e-15)


def test_mbes_callback_sets_geometry_once():
    n = make()
    scan = _ns(ranges=[10.0] * 32, angle_min=-1.0, angle_increment=2.0 / 31)
    n.mbes_cb(scan)
    n.mbes_cb(scan)
    assert [c[0] for c in n.pf.calls] == ["geometry", "update_mbes", "resample", "update_mbes", "resample"]


def test_kld_adaptive_node_resamples_to_the_kld_bound():
    params = dict(PARAMS, kld_adaptive=True, kld_bin_xy=0.25, kld_min_particles=500)
    n = node_mod.AuvPfNode(lambda name, default=_REQUIRED: params[name] if name in params else get_param(name, default),
                           np.eye(4), filter_factory=FakeFilter)
    n.mbes_cb(_ns(ranges=[10.0] * 32, angle_min=-1.0, angle_increment=2.0 / 31))
    assert n.pf.calls == [("geometry", 32), ("update_mbes", 32), ("kld", 0.25, 500), ("resample_to", 4321)]


@pytest.mark.gpu
def test_node_on_the_gpu_tracks_a_gps_fix_and_a_ping():
    from smarc_navigation_b200 import world
    clock = [100.0]
    n = node_mod.AuvPfNode(get_param, np.eye(4), now=lambda: clock[0])
    n.dive_cb(_ns(data=False))
    t = 100.0
    for k in range(20):                                  # 50 Hz odometry, a fix every 5th message
        t += 0.02
        n.time = t                                       # the stamp a concurrent odom callback would have set
        if k % 5 == 4:
            n.gps_odom_cb(gps_msg(3.0, -2.0))
        n.odom_callback(odom_msg(t, v=(0.0, 0.0, 0.0)))
    est = n.loc_loop()
    assert n.updates == 4
    assert np.hypot(est["position"][0] - 3.0, est["position"][1] + 2.0) < 1.0
    # an MBES ping over a synthetic map from the estimated pose itself: weights stay finite, estimate stays put
    dem = world.make_dem(200, 2.0)
    n.pf.set_poses(np.tile([200.0, 200.0, -5.0, 0.0, 0.0, 0.3], (64, 1)) + np.random.default_rng(0).normal(0, 0.2, (64, 6)) * [1, 1, 0, 0, 0, 0])
    n.set_map(dem["z"], dem["x0"], dem["y0"], dem["res"])
    th = np.linspace(-1.0, 1.0, 64)
    n.pf.set_scan_geometry(th[0], th[1] - th[0], 64)
    r = n.pf.raycast(np.array([[200.0, 200.0, -5.0, 0.0, 0.0, 0.3]]))[0]
    n.mbes_cb(_ns(ranges=r.tolist(), angle_min=th[0], angle_increment=th[1] - th[0]))
    est = n.loc_loop()
    assert np.hypot(est["position"][0] - 200.0, est["position"][1] - 200.0) < 2.0
    n.close()
