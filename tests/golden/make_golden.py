#!/usr/bin/env python
"""Generate the golden fixtures in this directory by EXECUTING THE UNMODIFIED
REFERENCE (``/root/reference/auv_particle_filter/scripts``) on seeded inputs.

The reference has no tests, fixtures or known-answer vectors of its own
(SURVEY.md §4), and being Python it cannot travel to the GPU box, so its
outputs are frozen here as small ``.npz`` files and committed together with
this script.  Run from the repo root (only works where /root/reference exists):

    python tests/golden/make_golden.py

How the reference is executed:
* ``auv_particle.py`` and ``resampling.py`` are imported as they lie, with an
  empty ``rospy`` module and ``oracle/tf_stub.py`` standing in for
  ``tf.transformations`` (``oracle/ref_shim.py``).
* ``auv_pf.py`` cannot be imported (ROS master, tf2_ros, message packages, and it
  spins in its constructor).  Its class body is lifted with ``ast`` from the
  source file AS IS and compiled in a namespace that supplies the few ROS names
  its methods touch; the methods ``predict``, ``update``, ``resample``,
  ``reassign_poses`` and ``update_loc_pose`` are then run on an instance created
  with ``object.__new__`` (no ``__init__``).  No reference source is copied.
* NumPy's global legacy generator is seeded and the very same draws are
  re-derived from ``np.random.RandomState(seed)`` in the reference's consumption
  order, so the fixtures carry the noise explicitly.
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim, tf_stub   # noqa: E402

REF = ref_shim.DEFAULT_REF
auv_particle, resampling = ref_shim.load(REF)
Particle = auv_particle.Particle


# ----------------------------------------------------------------------------- lift auv_pf
class _Recorder(object):
    def __init__(self):
        self.calls = []

    def publish(self, msg):
        self.calls.append(("publish", msg))

    def sendTransform(self, *a):
        self.calls.append(("tf", a))


class _Listener(object):
    """utm -> map is the identity in the fixtures (auv_pf.py:151)."""

    def transformPoint(self, frame, pt):
        return pt


def lift_auv_pf():
    src = open(os.path.join(REF, "auv_pf.py")).read()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "auv_pf"][0]
    mod = ast.Module(body=[cls], type_ignores=[])
    rospy = sys.modules["rospy"]

    class _Time(object):
        def __init__(self, *a):
            pass

        @staticmethod
        def now():
            return 0.0
    rospy.Time = _Time
    rospy.logwarn = lambda *a, **k: None
    tfmod = sys.modules["tf"]
    for exc in ("LookupException", "ConnectivityException", "ExtrapolationException"):
        setattr(tfmod, exc, type(exc, (Exception,), {}))

    class PointStamped(object):
        def __init__(self):
            self.header = ref_shim.Msg(frame_id=None, stamp=None)
            self.point = ref_shim.Msg(x=0.0, y=0.0, z=0.0)

    ns = dict(np=np, rospy=rospy, tf=tfmod, PointStamped=PointStamped,
              Quaternion=lambda *q: tuple(q),
              quaternion_from_euler=tf_stub.quaternion_from_euler,
              residual_resample=resampling.residual_resample,
              Particle=Particle)
    exec(compile(mod, os.path.join(REF, "auv_pf.py"), "exec"), ns)
    return ns["auv_pf"]


def make_filter(cls, pc, m2o, init_cov, meas_std, process_cov, res_noise_cov):
    pf = object.__new__(cls)
    pf.pc = pc
    pf.map_frame, pf.utm_frame = "map", "utm"
    pf.listener = _Listener()
    pf.res_noise_cov = res_noise_cov
    pf.particles = np.empty(pc, dtype=object)
    for i in range(pc):                                   # auv_pf.py:90-94
        pf.particles[i] = Particle(pc, i, m2o, init_cov=init_cov, meas_std=meas_std,
                                   process_cov=process_cov)
    pf.loc_pose = ref_shim.Msg(pose=ref_shim.Msg(pose=ref_shim.Msg(position=ref_shim.Msg(), orientation=None),
                                                 covariance=None),
                               header=ref_shim.Msg(stamp=None))
    pf.loc_pub = _Recorder()
    pf.loc_tf = _Recorder()
    pf.base_frame, pf.odom_frame = "base_link", "odom"
    return pf


def poses_of(pf):
    return np.array([np.asarray(p.p_pose, float) for p in pf.particles])


def m2o_matrix(rng, trivial=False):
    if trivial:
        return np.identity(4)
    q = tf_stub.quaternion_from_euler(*(rng.uniform(-0.2, 0.2, 2).tolist() + [rng.uniform(-3, 3)]))
    return tf_stub.translation_matrix(rng.uniform(-50, 50, 3)).dot(tf_stub.quaternion_matrix(q))


def rand_odom(rng):
    v = rng.normal(0, 1, 3) * [1.0, 0.2, 0.1]
    wz = rng.normal(0, 0.2)
    quat = tf_stub.quaternion_from_euler(rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(-3, 3))
    z = rng.uniform(-20, 0)
    return v, wz, quat, z


# ----------------------------------------------------------------------------- fixtures
def golden_particle(path):
    """Particle.__init__/add_noise/motion_pred/get_p_pose/compute_weight."""
    rng = np.random.default_rng(100)
    out = {}
    N = 48
    for case, trivial in (("a", True), ("b", False)):
        m2o = m2o_matrix(rng, trivial)
        init_cov = [4.0, 2.0, 0.5, 0.01, 0.02, 0.3]
        process_cov = [1e-2, 2e-2, 3e-3, 1e-4, 1e-4, 1e-3]
        meas_std = 1.7 if trivial else 0.4
        seed = 1234 + (0 if trivial else 1)
        np.random.seed(seed)
        parts = [Particle(N, i, m2o, init_cov=init_cov, meas_std=meas_std, process_cov=process_cov)
                 for i in range(N)]
        rs = np.random.RandomState(seed)
        init_noise = np.stack([rs.randn(6, 1)[:, 0] for _ in range(N)])
        out[case + "_m2o"] = m2o
        out[case + "_init_cov"] = np.array(init_cov)
        out[case + "_process_cov"] = np.array(process_cov)
        out[case + "_meas_std"] = meas_std
        out[case + "_init_noise"] = init_noise
        out[case + "_poses0"] = np.array([np.asarray(p.p_pose, float) for p in parts])
        # put the particles somewhere interesting (yaw beyond +-pi to exercise the wrap)
        for i, p in enumerate(parts):
            p.p_pose = np.asarray(p.p_pose, float) + np.array([100.0, -40.0, 0, 0, 0, 3.0 * ((i % 3) - 1)])
        out[case + "_poses_start"] = np.array([np.asarray(p.p_pose, float) for p in parts])
        steps = 6
        odoms, noises, traj = [], [], []
        for s in range(steps):
            v, wz, quat, z = rand_odom(rng)
            dt = float(rng.uniform(0.01, 0.2))
            odom = ref_shim.make_odom(v, wz, quat, z)
            noise = []
            for p in parts:                                 # auv_pf.py:215-216
                p.motion_pred(odom, dt)
                noise.append(rs.randn(1, 6)[0])
            odoms.append(np.concatenate([v, [wz], quat, [z, dt]]))
            noises.append(np.array(noise))
            traj.append(np.array([np.asarray(p.p_pose, float) for p in parts]))
        out[case + "_odoms"] = np.array(odoms)              # v(3) wz quat(4) z dt
        out[case + "_pred_noise"] = np.array(noises)
        out[case + "_pred_poses"] = np.array(traj)
        # get_p_pose + compute_weight on the final state
        ps, Rs, ws = [], [], []
        gps = rng.uniform(-3, 3, 2) + traj[-1][:, 0:2].mean(0) @ m2o[:2, :2].T + m2o[:2, 3]
        for p in parts:
            pp, RR = p.get_p_pose()
            ps.append(np.array(pp))
            Rs.append(np.array(RR))
        gps = np.array(ps)[:, 0:2].mean(0) + rng.uniform(-2, 2, 2)
        for p in parts:
            p.compute_weight(ref_shim.make_point(gps[0], gps[1]))
            ws.append(p.w)
        out[case + "_p"] = np.array(ps)
        out[case + "_R"] = np.array(Rs)
        out[case + "_gps"] = gps
        out[case + "_w"] = np.array(ws)
        # add_noise with a resampling covariance
        res_cov = [1.0, 1.0, 0.1, 0.0, 1e-3, 1e-4]
        an = []
        for p in parts:
            p.add_noise(res_cov)
            an.append(rs.randn(6, 1)[:, 0])
        out[case + "_res_cov"] = np.array(res_cov)
        out[case + "_res_noise"] = np.array(an)
        out[case + "_poses_rough"] = np.array([np.asarray(p.p_pose, float) for p in parts])
    # fullRotation including its typo
    ang = rng.uniform(-3, 3, (16, 3))
    out["fullrot_angles"] = ang
    out["fullrot"] = np.array([parts[0].fullRotation(*a) for a in ang])
    np.savez_compressed(path, **out)


def weight_cases(rng):
    cases = {}
    for N in (7, 100, 1000):
        cases["dirichlet_%d" % N] = rng.dirichlet(np.ones(N))
        w = rng.dirichlet(np.ones(N) * 0.05)
        cases["peaked_%d" % N] = w
        cases["uniform_%d" % N] = np.full(N, 1.0 / N)
        w = np.zeros(N)
        w[rng.integers(0, N, 3)] = [0.5, 0.3, 0.2][:3]
        w /= w.sum()
        cases["sparse_%d" % N] = w
        g = np.exp(-0.5 * (np.arange(N) - N / 3.0) ** 2 / (N / 20.0) ** 2) + 1e-200
        cases["gauss_%d" % N] = g / g.sum()
    return cases


def golden_resampling(path):
    """All five resamplers of resampling.py with the module-level ``random``
    replaced by a deterministic source (resampling.py:24)."""
    rng = np.random.default_rng(200)
    out = {}
    orig_random = resampling.random
    try:
        for name, w in weight_cases(rng).items():
            N = len(w)
            out[name + "_w"] = w
            u = float(rng.uniform())
            U = rng.uniform(size=N)
            out[name + "_u"] = u
            out[name + "_U"] = U

            resampling.random = lambda *a: u
            try:
                out[name + "_systematic"] = resampling.systematic_resample(w)
            except IndexError:
                out[name + "_systematic"] = np.array([-1], 'i')     # reference raised
            resampling.random = lambda n=None: U[:n].copy()
            try:
                out[name + "_stratified"] = resampling.stratified_resample(w)
            except IndexError:
                out[name + "_stratified"] = np.array([-1], 'i')
            out[name + "_multinomial"] = resampling.multinomial_resample(w)
            out[name + "_residual"] = resampling.residual_resample(w.copy())
    finally:
        resampling.random = orig_random
    np.savez_compressed(path, **out)


def golden_quantised_systematic(path):
    """systematic_resample on weights that lie on the 2^-52 grid and sum to
    exactly 1 (the form the GPU normaliser emits)."""
    from oracle import pf_oracle
    rng = np.random.default_rng(300)
    out = {}
    orig_random = resampling.random
    try:
        for name, N, scale in (("n100", 100, 3.0), ("n4096", 4096, 8.0), ("n65536", 65536, 20.0)):
            logw = (rng.normal(0, scale, N) - 470.0).astype(np.float32)
            qq, a, A, m = pf_oracle.quantise_logw(logw)
            w = pf_oracle.weights_from_q(qq)
            u = float(rng.uniform())
            resampling.random = lambda *a_: u
            out[name + "_logw"] = logw
            out[name + "_qq"] = qq
            out[name + "_u"] = u
            out[name + "_idx"] = resampling.systematic_resample(w)
    finally:
        resampling.random = orig_random
    np.savez_compressed(path, **out)


def golden_quantised_residual(path):
    """residual_resample -- the resampler auv_pf.resample actually calls (auv_pf.py:182) -- on
    weights that lie on the 2^-52 grid, with the ``random(N - k)`` draw of resampling.py:74
    injected.  The GPU's AUVPF_RS_STRICT_REF mode must reproduce ``idx`` bit for bit."""
    from oracle import pf_oracle
    rng = np.random.default_rng(301)
    out = {}
    orig_random = resampling.random
    try:
        for name, N, scale in (("n100", 100, 2.0), ("n1000", 1000, 1.0), ("n4096", 4096, 4.0),
                               ("n20000", 20000, 0.3)):
            logw = (rng.normal(0, scale, N) - 470.0).astype(np.float32)
            qq, a, A, m = pf_oracle.quantise_logw(logw)
            w = pf_oracle.weights_from_q(qq)
            U = rng.uniform(size=N)
            resampling.random = lambda n_, U=U: U[:n_].copy()
            out[name + "_logw"] = logw
            out[name + "_qq"] = qq
            out[name + "_U"] = U
            out[name + "_idx"] = resampling.residual_resample(w.copy())
    finally:
        resampling.random = orig_random
    np.savez_compressed(path, **out)


def golden_auv_pf(path):
    """auv_pf.predict / update / resample / update_loc_pose, lifted class."""
    cls = lift_auv_pf()
    rng = np.random.default_rng(400)
    out = {}
    orig_random = resampling.random
    N = 100
    m2o = m2o_matrix(rng)
    init_cov = [0.1, 0.1, 0.0, 0.0, 0.0, 0.0]                       # auv_pf.launch:17
    process_cov = [1e-3, 1e-3, 0.0, 0.0, 0.0, 1e-5]
    res_cov = [1.0, 1.0, 0.0, 0.0, 0.0, 1e-4]                       # auv_pf.launch:19
    meas_std = 1.0                                                  # auv_pf.launch:20
    seed = 77
    np.random.seed(seed)
    rs = np.random.RandomState(seed)
    pf = make_filter(cls, N, m2o, init_cov, meas_std, process_cov, res_cov)
    init_noise = np.stack([rs.randn(6, 1)[:, 0] for _ in range(N)])
    out.update(m2o=m2o, init_cov=np.array(init_cov), process_cov=np.array(process_cov),
               res_cov=np.array(res_cov), meas_std=meas_std, init_noise=init_noise,
               poses0=poses_of(pf))
    cycles, P = 12, 3
    odoms, pnoise, rnoise, runi, gpss, weights, indices, poses_pred, poses_res, est = \
        [], [], [], [], [], [], [], [], [], []
    truth = np.zeros(2)
    try:
        for c in range(cycles):
            for s in range(P):
                v, wz, quat, z = rand_odom(rng)
                v[0] = 1.0
                dt = 0.02 * (1 + s)
                pf.time, pf.old_time = dt, 0.0                       # predict: dt = time - old_time (:214)
                pf.predict(ref_shim.make_odom(v, wz, quat, z))
                pnoise.append(np.stack([rs.randn(1, 6)[0] for _ in range(N)]))
                odoms.append(np.concatenate([v, [wz], quat, [z, dt]]))
            poses_pred.append(poses_of(pf))
            # GPS fix near the cloud, in the map frame (utm->map identity)
            pm = poses_pred[-1][:, 0:3] @ m2o[:3, :3].T + m2o[:3, 3]
            gps = pm[:, 0:2].mean(0) + rng.normal(0, 0.7, 2)
            gps_odom = ref_shim.Msg(pose=ref_shim.Msg(pose=ref_shim.Msg(position=ref_shim.Msg(x=gps[0], y=gps[1]))))
            w = pf.update(gps_odom)
            weights.append(w.copy())
            # resample draws its uniforms through resampling.random (:24,74)
            U = rng.uniform(size=N)
            drawn = {}

            def fake_random(n=None, U=U, drawn=drawn):
                drawn["n"] = n
                return U[:n].copy()
            resampling.random = fake_random
            # capture the index draw too
            real_rr = resampling.residual_resample
            cap = {}

            def rr(wts, real_rr=real_rr, cap=cap):
                cap["idx"] = real_rr(wts)
                return cap["idx"]
            pf.resample.__func__.__globals__["residual_resample"] = rr
            pf.resample(w)
            pf.resample.__func__.__globals__["residual_resample"] = real_rr
            rnoise.append(np.stack([rs.randn(6, 1)[:, 0] for _ in range(N)]))
            runi.append(U)
            gpss.append(gps)
            indices.append(cap["idx"].copy())
            poses_res.append(poses_of(pf))
            # estimate
            pf.loc_pub.calls.clear()
            pf.loc_tf.calls.clear()
            pf.update_loc_pose([p.p_pose for p in pf.particles])
            lp = pf.loc_pose.pose
            est.append(np.concatenate([[lp.pose.position.x, lp.pose.position.y, lp.pose.position.z],
                                       np.asarray(lp.pose.orientation, float),
                                       np.asarray(lp.covariance, float)]))
    finally:
        resampling.random = orig_random
    out.update(odoms=np.array(odoms), pred_noise=np.array(pnoise), res_noise=np.array(rnoise),
               res_uniforms=np.array(runi), gps=np.array(gpss), weights=np.array(weights),
               indices=np.array(indices), poses_pred=np.array(poses_pred),
               poses_res=np.array(poses_res), estimate=np.array(est), P=P)
    np.savez_compressed(path, **out)


if __name__ == "__main__":
    golden_particle(os.path.join(HERE, "particle.npz"))
    golden_resampling(os.path.join(HERE, "resampling.npz"))
    golden_quantised_systematic(os.path.join(HERE, "systematic_quantised.npz"))
    golden_quantised_residual(os.path.join(HERE, "residual_quantised.npz"))
    golden_auv_pf(os.path.join(HERE, "auv_pf_cycle.npz"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


def golden_sam_motion(path):
    """sam_dead_reckoning/scripts/sam_mm.py SAM.motion (the motion-model fallback of dr_node.py:191-201) on
    random thruster commands; imported unmodified (it needs only NumPy)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(REF)), "sam_dead_reckoning", "scripts"))
    import sam_mm
    sam = sam_mm.SAM()
    rng = np.random.default_rng(0)
    ctrl = np.column_stack([rng.uniform(-1500, 1500, 64), rng.uniform(-0.13, 0.13, 64)])
    np.savez_compressed(path, control=ctrl, nudot=np.array([sam.motion(list(c)) for c in ctrl]))


if __name__ == "__main__":
    golden_sam_motion(os.path.join(HERE, "sam_motion.npz"))
