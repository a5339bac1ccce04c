"""The oracle restatement vs golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import pf_oracle as o
from oracle import c_oracle, ref_shim, tf_stub


@pytest.mark.parametrize("case", ["a", "b"])
def test_particle_init_predict_pose_weight(golden, case):
    g = golden("particle")
    k = lambda s: g[case + "_" + s]
    N = k("init_noise").shape[0]
    poses = o.init_particles(N, k("init_cov"), k("init_noise"))          # auv_particle.py:30
    np.testing.assert_allclose(poses, k("poses0"), rtol=0, atol=1e-14)
    poses = k("poses_start").copy()
    for odom, noise, want in zip(k("odoms"), k("pred_noise"), k("pred_poses")):
        poses = o.motion_pred(poses, odom[0:3], odom[3], odom[4:8], odom[8], odom[9], k("process_cov"), noise)
        np.testing.assert_allclose(poses, want, rtol=0, atol=1e-11)
        assert np.all(poses[:, 5] >= -np.pi) and np.all(poses[:, 5] < np.pi)
    p, R = o.get_p_pose(poses, k("m2o"))                                  # auv_particle.py:72-83
    np.testing.assert_allclose(p, k("p"), rtol=0, atol=1e-10)
    np.testing.assert_allclose(R, k("R"), rtol=0, atol=1e-13)
    w = np.exp(o.gps_logw(poses, k("m2o"), k("gps"), float(k("meas_std"))))
    np.testing.assert_allclose(w, k("w"), rtol=1e-11, atol=0)
    rough = o.add_noise(poses, k("res_cov"), k("res_noise"))              # auv_particle.py:32-36
    np.testing.assert_allclose(rough, k("poses_rough"), rtol=0, atol=1e-11)


def test_c_motion_pred_matches_golden(golden):
    g = golden("particle")
    poses = g["b_poses_start"].copy()
    for odom, noise, want in zip(g["b_odoms"], g["b_pred_noise"], g["b_pred_poses"]):
        roll, pitch = o.odom_roll_pitch(odom[4:8])
        poses = c_oracle.motion_pred(poses, odom[0:3], odom[3], roll, pitch, odom[8], odom[9],
                                     g["b_process_cov"], noise)
        np.testing.assert_allclose(poses, want, rtol=0, atol=1e-11)


def test_full_rotation_with_typo(golden):
    g = golden("particle")
    ang = g["fullrot_angles"]
    np.testing.assert_allclose(o.full_rotation(ang[:, 0], ang[:, 1], ang[:, 2]), g["fullrot"], atol=1e-15)
    # rows 0,1 equal the correct ZYX rotation -- the only rows that reach the state
    np.testing.assert_allclose(g["fullrot"][:, 0:2, :], o.rpy_matrix(ang[:, 0], ang[:, 1], ang[:, 2])[:, 0:2, :],
                               atol=1e-15)


def _names(g):
    return sorted({k[:-2] for k in g.files if k.endswith("_w")})


def test_resamplers(golden):
    g = golden("resampling")
    names = _names(g)
    assert len(names) == 15
    for n in names:
        w, u, U = g[n + "_w"], float(g[n + "_u"]), g[n + "_U"]
        want = g[n + "_systematic"]
        if want[0] < 0:
            with pytest.raises(IndexError):
                o.systematic_resample(w, u)
        else:
            np.testing.assert_array_equal(o.systematic_resample(w, u), want)
            np.testing.assert_array_equal(c_oracle.systematic(w, u), want)
        if g[n + "_stratified"][0] >= 0:
            np.testing.assert_array_equal(o.stratified_resample(w, U), g[n + "_stratified"])
        np.testing.assert_array_equal(o.multinomial_resample(w, U), g[n + "_multinomial"])
        np.testing.assert_array_equal(o.residual_resample(w, U), g[n + "_residual"])
        head, k = o.residual_copies(w)
        np.testing.assert_array_equal(head, g[n + "_residual"][:len(head)])


def test_quantised_systematic(golden):
    """Weights on the 2^-52 grid: sequential cumsum == exact integer scan, so the
    reference's ancestors are reproduced from the integer CDF alone."""
    g = golden("systematic_quantised")
    for n in ("n100", "n4096", "n65536"):
        logw, qq, u, idx = g[n + "_logw"], g[n + "_qq"], float(g[n + "_u"]), g[n + "_idx"]
        q2, a, A, m = o.quantise_logw(logw)
        np.testing.assert_array_equal(q2, qq)
        q3, m3, A3 = c_oracle.quantise_logw(logw)
        np.testing.assert_array_equal(q3, qq)
        assert A3 == A and m3 == m
        assert int(qq.astype(object).sum()) == 1 << 52
        w = o.weights_from_q(qq)
        assert np.cumsum(w)[-1] == 1.0 and np.cumsum(w[::-1])[-1] == 1.0
        # integer CDF -> float64 exactly equals np.cumsum
        cdf_int = np.cumsum(qq)
        np.testing.assert_array_equal(np.ldexp(cdf_int.astype(np.float64), -52), np.cumsum(w))
        np.testing.assert_array_equal(o.systematic_resample(w, u), idx)
        np.testing.assert_array_equal(c_oracle.systematic(w, u), idx)
        # tolerance leg: quantised weights vs plain float64 normalisation
        ref = o.normalise_logw(logw)
        big = ref > 1e-9
        np.testing.assert_allclose(w[big], ref[big], rtol=1e-6)
        np.testing.assert_allclose(w, ref, atol=2.0 ** -50)


def test_auv_pf_cycle(golden):
    """predict x P -> update -> resample (+bookkeeping, roughening) -> estimate,
    against the lifted auv_pf methods (auv_pf.py:135-260)."""
    g = golden("auv_pf_cycle")
    N = g["init_noise"].shape[0]
    P = int(g["P"])
    poses = o.init_particles(N, g["init_cov"], g["init_noise"])
    np.testing.assert_allclose(poses, g["poses0"], atol=1e-14)
    for c in range(g["gps"].shape[0]):
        for s in range(P):
            od = g["odoms"][c * P + s]
            poses = o.motion_pred(poses, od[0:3], od[3], od[4:8], od[8], od[9], g["process_cov"],
                                  g["pred_noise"][c * P + s])
        np.testing.assert_allclose(poses, g["poses_pred"][c], atol=1e-9)
        w = o.gps_weights(poses, g["m2o"], g["gps"][c], float(g["meas_std"]))
        np.testing.assert_allclose(w, g["weights"][c], rtol=1e-9)
        wn = o.normalise(g["weights"][c])
        idx = o.residual_resample(wn, g["res_uniforms"][c])
        np.testing.assert_array_equal(idx, g["indices"][c])
        poses = o.resample_apply(poses, idx, g["res_cov"], g["res_noise"][c])
        np.testing.assert_allclose(poses, g["poses_res"][c], atol=1e-9)
        ave, cov, quat = o.estimate(poses)
        e = g["estimate"][c]
        np.testing.assert_allclose(ave[0:3], e[0:3], atol=1e-10)
        np.testing.assert_allclose(quat, e[3:7], atol=1e-12)
        np.testing.assert_allclose(o.covariance36(cov), e[7:43], atol=1e-10)


def test_bookkeeping_bruteforce():
    """resample_bookkeeping vs the literal list operations of auv_pf.py:183-187."""
    rng = np.random.default_rng(5)
    for _ in range(50):
        pc = int(rng.integers(1, 60))
        indices = rng.integers(0, pc, pc).astype('i')
        keep = list(set(indices))
        lost = [i for i in range(pc) if i not in keep]
        dupes = indices[:].tolist()
        for i in keep:
            dupes.remove(i)
        l2, d2 = o.resample_bookkeeping(indices, pc)
        assert lost == l2.tolist() and dupes == d2.tolist()


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_live_reference_still_matches_golden(golden):
    """Where the checkout exists, re-run one reference call live."""
    _, resampling = ref_shim.load()
    g = golden("resampling")
    w, u = g["dirichlet_100_w"], float(g["dirichlet_100_u"])
    old = resampling.random
    resampling.random = lambda *a: u
    try:
        np.testing.assert_array_equal(resampling.systematic_resample(w), g["dirichlet_100_systematic"])
    finally:
        resampling.random = old


def test_tf_stub_against_scipy():
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(9)
    for _ in range(100):
        r, p, y = rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5), rng.uniform(-3.1, 3.1)
        q = tf_stub.quaternion_from_euler(r, p, y)
        R = tf_stub.quaternion_matrix(q)[:3, :3]
        Rs = Rotation.from_euler("xyz", [r, p, y]).as_matrix()        # extrinsic xyz == 'sxyz'
        np.testing.assert_allclose(R, Rs, atol=1e-14)
        np.testing.assert_allclose(R, o.rpy_matrix(r, p, y), atol=1e-14)
        np.testing.assert_allclose(tf_stub.euler_from_quaternion(q), [r, p, y], atol=1e-12)
        qs = Rotation.from_euler("xyz", [r, p, y]).as_quat()
        assert min(np.abs(q - qs).max(), np.abs(q + qs).max()) < 1e-14


def test_philox_known_answers():
    """Random123 philox4x32-10 known-answer vectors."""
    c = np.array([[0, 0, 0, 0], [0xffffffff] * 4, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], np.uint32)
    k = np.array([[0, 0], [0xffffffff] * 2, [0xa4093822, 0x299f31d0]], np.uint32)
    want = np.array([[0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8],
                     [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd],
                     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]], np.uint32)
    np.testing.assert_array_equal(o.philox4x32_10(c, k), want)


def test_residual_quantised_golden(golden):
    """The literal residual_resample restatement against the unmodified reference run on
    2^-52-grid weights with the random(N - k) draw injected; the fixed-point variant keeps the
    reference's deterministic head and draws the tail only from particles with a fractional part."""
    g = golden("residual_quantised")
    for n in ("n100", "n1000", "n4096", "n20000"):
        qq, U, idx = g[n + "_qq"], g[n + "_U"], g[n + "_idx"]
        w = o.weights_from_q(qq)
        np.testing.assert_array_equal(o.quantise_logw(g[n + "_logw"])[0], qq)
        np.testing.assert_array_equal(o.residual_resample(w, U), idx)
        head, copies = o.residual_copies(w)
        fx = o.residual_resample_fixed(qq, U)
        np.testing.assert_array_equal(fx[:len(head)], idx[:len(head)])
        frac = len(w) * w - copies
        assert np.all(frac[fx[len(head):]] > 0)
        # the tail follows the fractional parts: chi-square-free sanity on the mean index weight
        if len(w) - len(head) > 200:
            pr = frac / frac.sum()
            assert abs(pr[fx[len(head):]].mean() - (pr * pr).sum()) < 5.0 * np.sqrt((pr ** 3).sum() / (len(w) - len(head)))


def test_kld_oracle_and_library_bound():
    """KLD adaptation: `systematic_resample_to` with n_out = N is the reference function (and the golden
    weights' ancestors); Fox's bound from the library (a pure host function of the C ABI) equals the oracle's for a
    sweep of bin counts; the bin count is numpy's unique over floor(coordinate / bin)."""
    rng = np.random.default_rng(3)
    w = rng.dirichlet(np.ones(500) * 0.3)
    for u in (0.0, 0.25, 0.999):
        np.testing.assert_array_equal(o.systematic_resample_to(w, u, 500), o.systematic_resample(w, u))
    a = o.systematic_resample_to(w, 0.4, 123)
    assert len(a) == 123 and np.all(np.diff(a) >= 0) and a[-1] <= 499
    up = o.systematic_resample_to(w, 0.4, 5000)
    counts = np.bincount(up, minlength=500)
    assert np.all(np.abs(counts - 5000 * w) <= 1.0 + 1e-9)          # systematic: every count within one of N w
    from smarc_navigation_b200 import _lib
    L = _lib.load()
    for k in (0, 1, 2, 3, 10, 499, 500, 12345, 10 ** 6, 10 ** 9):
        for eps, z in ((0.05, 2.3263478740408408), (0.2, 0.0), (1e-4, 3.0)):
            assert int(L.auvpf_kld_target(k, eps, z, 50, 10 ** 12)) == o.kld_target(k, eps, z, 50, 10 ** 12)
    assert o.kld_target(500, 0.05, 2.3263478740408408, 1, 10 ** 9) == 5755            # Fox 2003, eq. 13 by hand
    poses = np.zeros((6, 6))
    poses[:, 0] = [0.1, 0.4, 0.6, -0.1, 0.1, 0.1]
    poses[:, 5] = [0, 0, 0, 0, 0.2, -3.2]
    assert o.kld_bins(poses, 0.5, 0.1) == 5
