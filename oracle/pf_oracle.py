"""NumPy (float64) restatement of the reference particle-filter arithmetic.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function cites the
reference lines it follows (paths relative to
``/root/reference/auv_particle_filter/scripts/``).  The functions are
vectorised over particles where the reference loops in Python; results agree
with the unmodified reference to float64 round-off, which
``tests/test_oracle_golden.py`` checks against ``tests/golden/*.npz``.

State layout: ``poses`` is ``(N, 6)`` float64 ``[x, y, z, roll, pitch, yaw]``
(``auv_particle.py:24,36``).  Random draws are always INJECTED (``noise`` is an
``(N, 6)`` array of standard normals in the order the reference consumes them)
because the reference uses NumPy's global legacy generator
(``auv_particle.py:35,40``, ``resampling.py:24``).
"""
import math

import numpy as np

from . import tf_stub

WEIGHT_FLOOR = 1.0e-200          # auv_pf.py:165


# --------------------------------------------------------------------------- motion
def full_rotation(roll, pitch, yaw):
    """``Particle.fullRotation`` (auv_particle.py:86-97), INCLUDING its
    ``rot_y`` third row ``[-sin p, cos p, 0]`` (the reference's typo, :90-92).
    Vectorised: returns ``(..., 3, 3)``."""
    roll, pitch, yaw = np.broadcast_arrays(np.asarray(roll, float),
                                           np.asarray(pitch, float),
                                           np.asarray(yaw, float))
    shp = roll.shape
    z = np.zeros(shp)
    o = np.ones(shp)
    cy, sy = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cr, sr = np.cos(roll), np.sin(roll)
    rot_z = np.stack([np.stack([cy, -sy, z], -1),
                      np.stack([sy, cy, z], -1),
                      np.stack([z, z, o], -1)], -2)
    rot_y = np.stack([np.stack([cp, z, sp], -1),
                      np.stack([z, o, z], -1),
                      np.stack([-sp, cp, z], -1)], -2)
    rot_x = np.stack([np.stack([o, z, z], -1),
                      np.stack([z, cr, -sr], -1),
                      np.stack([z, sr, cr], -1)], -2)
    return np.matmul(rot_z, np.matmul(rot_y, rot_x))


def wrap_pi(a):
    """``(a + pi) % (2 pi) - pi`` (auv_particle.py:48, auv_pf.py:229); Python /
    NumPy ``%`` takes the sign of the divisor, so the result is in [-pi, pi)."""
    return (np.asarray(a, float) + np.pi) % (2 * np.pi) - np.pi


def odom_roll_pitch(quat_xyzw):
    """roll, pitch read from the odometry quaternion (auv_particle.py:50-56)."""
    e = tf_stub.euler_from_quaternion(np.asarray(quat_xyzw, float))
    return e[0], e[1]


def motion_pred(poses, v, wz, quat_xyzw, z, dt, process_cov, noise):
    """``Particle.motion_pred`` (auv_particle.py:38-70) for all particles.

    ``noise`` is (N, 6) standard normals = the ``np.random.randn(1, 6)`` of
    line :40, one row per particle in particle order (auv_pf.py:215-216)."""
    poses = np.array(poses, dtype=np.float64, copy=True)
    noise_vec = np.sqrt(np.asarray(process_cov, float))[None, :] * np.asarray(noise, float)   # :40
    vel_rot = np.array([0.0, 0.0, wz])                                                     # :43-45
    rot_t = poses[:, 3:6] + vel_rot * dt + noise_vec[:, 3:6]                               # :47
    yaw_t = (rot_t[:, 2] + np.pi) % (2 * np.pi) - np.pi                                    # :48
    roll_t, pitch_t = odom_roll_pitch(quat_xyzw)                                           # :50-56
    poses[:, 3] = roll_t                                                                   # :57
    poses[:, 4] = pitch_t
    poses[:, 5] = yaw_t
    vel_p = np.asarray(v, float)                                                           # :60-62
    R = full_rotation(roll_t, pitch_t, yaw_t)                                              # :64
    step_t = np.matmul(R, vel_p * dt) + noise_vec[:, 0:3]                                  # :65
    poses[:, 0] += step_t[:, 0]                                                            # :67
    poses[:, 1] += step_t[:, 1]                                                            # :68
    poses[:, 2] = z                                                                        # :70
    return poses


def add_noise(poses, cov, noise):
    """``Particle.add_noise`` (auv_particle.py:32-36): ``pose + sqrt(diag(cov)) @ randn(6,1)``.
    ``noise`` is (N, 6): row i = the 6 draws of particle i."""
    return np.asarray(poses, float) + np.sqrt(np.asarray(cov, float))[None, :] * np.asarray(noise, float)


def init_particles(n, init_cov, noise):
    """``Particle.__init__`` (auv_particle.py:24,30): zero pose + ``add_noise(init_cov)``,
    created in index order (auv_pf.py:90-94)."""
    return add_noise(np.zeros((n, 6)), init_cov, noise)


# --------------------------------------------------------------------------- pose in map
def rpy_matrix(roll, pitch, yaw):
    """Rotation of ``quaternion_matrix(quaternion_from_euler(r, p, y))``
    (auv_particle.py:75-76) in closed form: ``Rz(yaw) Ry(pitch) Rx(roll)``."""
    roll, pitch, yaw = np.broadcast_arrays(np.asarray(roll, float),
                                           np.asarray(pitch, float),
                                           np.asarray(yaw, float))
    cy, sy = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cr, sr = np.cos(roll), np.sin(roll)
    R = np.empty(roll.shape + (3, 3))
    R[..., 0, 0] = cy * cp
    R[..., 0, 1] = cy * sp * sr - sy * cr
    R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp
    R[..., 1, 1] = sy * sp * sr + cy * cr
    R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp
    R[..., 2, 1] = cp * sr
    R[..., 2, 2] = cp * cr
    return R


def get_p_pose(poses, m2o):
    """``Particle.get_p_pose`` (auv_particle.py:72-83): ``m2o @ (T(xyz) @ Rq(rpy))``;
    returns ``p`` (N, 3) and ``R`` (N, 3, 3) in the map frame."""
    poses = np.asarray(poses, float)
    m2o = np.asarray(m2o, float)
    R_local = rpy_matrix(poses[:, 3], poses[:, 4], poses[:, 5])
    p = poses[:, 0:3] @ m2o[:3, :3].T + m2o[:3, 3]
    R = np.matmul(m2o[:3, :3], R_local)
    return p, R


def get_p_pose_literal(pose, m2o):
    """One particle, through the literal tf call chain of auv_particle.py:74-81."""
    t_particle = tf_stub.translation_matrix(pose[0:3])
    q = tf_stub.quaternion_from_euler(pose[3], pose[4], pose[5])
    mat = np.dot(t_particle, tf_stub.quaternion_matrix(q))
    trans_mat = np.asarray(m2o, float).dot(mat)
    return trans_mat[0:3, 3], trans_mat[0:3, 0:3]


# --------------------------------------------------------------------------- weights
def gps_logw(poses, m2o, gps_xy, meas_std):
    """log of ``multivariate_normal.pdf(z, mean=p[0:2], cov=diag(std^2, std^2))``
    (auv_particle.py:28,100-106) in closed form."""
    p, _ = get_p_pose(poses, m2o)
    var = float(meas_std) ** 2
    d2 = (gps_xy[0] - p[:, 0]) ** 2 + (gps_xy[1] - p[:, 1]) ** 2
    return -0.5 * d2 / var - math.log(2.0 * math.pi * var)


def gps_weights(poses, m2o, gps_xy, meas_std):
    """``auv_pf.update`` (auv_pf.py:135-167): linear pdf + 1e-200 (:165)."""
    return np.exp(gps_logw(poses, m2o, gps_xy, meas_std)) + WEIGHT_FLOOR


def beam_logw(r_meas, r_exp, meas_std, range_min=0.0, range_max=np.inf):
    """North-star form of ``compute_weight``: independent Gaussian per beam with
    the reference's diagonal ``meas_cov = std^2 I`` (auv_particle.py:28,105-106):
    ``logw = -1/(2 std^2) sum_b (r_b - rhat_b)^2 - B_valid log(std sqrt(2 pi))``.
    Measured beams that are non-finite or outside [range_min, range_max] are
    skipped (LaserScan validity convention,
    mbes_processors/mbes_toy_processor/src/toy_mbes_manipulator.cpp:69-73).
    ``r_exp``: (N, B) expected ranges; ``r_meas``: (B,)."""
    r_meas = np.asarray(r_meas, float)
    valid = np.isfinite(r_meas) & (r_meas >= range_min) & (r_meas <= range_max)
    d = np.asarray(r_exp, float)[:, valid] - r_meas[valid][None, :]
    nv = int(valid.sum())
    return -0.5 * (d * d).sum(axis=1) / meas_std ** 2 - nv * math.log(meas_std * math.sqrt(2.0 * math.pi))


def normalise(weights):
    """``weights /= weights.sum()`` (auv_pf.py:172)."""
    w = np.array(weights, dtype=np.float64, copy=True)
    w /= w.sum()
    return w


def normalise_logw(logw):
    """Plain float64 log-sum-exp normalisation (tolerance reference for the
    exact fixed-point pipeline below)."""
    lw = np.asarray(logw, np.float64)
    m = lw.max()
    e = np.exp(lw - m)
    return e / e.sum()


# --------------------------------------------------------------------------- exact normalise spec
# The build's own specification: turns fp32 log-weights into
# float64 weights on a 2^-52 grid that sum to EXACTLY 1 in any summation order,
# so np.cumsum (sequential), a blocked GPU scan and a cross-GPU offset scan all
# give bit-identical CDFs.  Uses only IEEE single +,* (no FMA) and integer
# arithmetic, so C, CUDA and NumPy agree bit for bit.
Q_BITS = 52
A_BITS = 60      # a_i = floor(E(logw_i - max) * 2^60); A = sum a_i is kept exactly (Python ints / int128)
_F = np.float32
_EXP_COEF32 = [_F(1.0) / _F(5040.0), _F(1.0) / _F(720.0), _F(1.0) / _F(120.0), _F(1.0) / _F(24.0), _F(1.0) / _F(6.0),
               _F(0.5), _F(1.0), _F(1.0)]


def integer_weights(logw32, m):
    """``a_i = floor(E(logw_i - m) 2^60)`` with E the library's fixed float32 evaluation of exp
    (csrc/auvpf_math.cuh ``a_from_logw``): IEEE single multiplies and adds only -- NumPy float32 arithmetic
    is exactly that -- then the significand is shifted into place with integers.  ``m``: the float32 maximum
    the weights refer to (the GLOBAL one for a shard of a sharded filter)."""
    lw = np.asarray(logw32, np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        x = lw - _F(m)
        live = x >= _F(-41.5)                                     # NaN compares false
        x = np.where(live, x, _F(0.0)).astype(np.float32)
        t = x * _F(1.44269502)
        k = (t + _F(12582912.0)) - _F(12582912.0)                 # rint(t) for |t| < 2^22
        r = (x - k * _F(0.693359375)) - k * _F(-2.12194440e-4)
        p = np.full_like(r, _EXP_COEF32[0])
        for c in _EXP_COEF32[1:]:
            p = p * r + c
    assert p.dtype == np.float32 and r.dtype == np.float32
    bits = p.view(np.uint32).astype(np.int64)
    shift = k.astype(np.int64) + (A_BITS - 23) + ((bits >> 23) - 127)
    mant = ((bits & 0x7FFFFF) | 0x800000).astype(np.uint64)
    a = np.where(shift >= 0, mant << np.clip(shift, 0, 63).astype(np.uint64),
                 np.where(shift > -24, mant >> np.clip(-shift, 0, 63).astype(np.uint64), np.uint64(0)))
    return np.where(live, a, np.uint64(0)).astype(np.uint64)


def quantise_logw(logw32):
    """fp32 log-weights -> integer weights ``qq`` (uint64) with sum == 2^52.
    Returns ``(qq, a, A, m)``; raises FloatingPointError when degenerate."""
    lw = np.asarray(logw32, np.float32)
    finite = lw[~np.isnan(lw)]
    if finite.size == 0 or not np.isfinite(finite.max()):
        raise FloatingPointError("degenerate weights")
    m = np.float32(finite.max())
    return quantise_a(integer_weights(lw, m)) + (m,)


def quantise_linear(w):
    """Linear float64 weights -> same fixed-point pipeline: ``a_i =
    floor(w_i / wmax * 2^A_BITS)`` (IEEE division, deterministic)."""
    w = np.asarray(w, np.float64)
    wmax = w.max()
    if not (np.isfinite(wmax) and wmax > 0.0):
        raise FloatingPointError("degenerate weights")
    ab = A_BITS
    e = np.where(w > 0.0, w / wmax, 0.0)
    a = np.floor(np.ldexp(e, ab)).astype(np.uint64)
    return quantise_a(a) + (wmax,)


def quantise_a(a):
    """Integer stage: the CDF as a quotient of PREFIX SUMS, ``cdf_j = floor(C_j 2^52 / A)`` with
    ``C_j = sum_{i<=j} a_i`` and ``A = C_{N-1}``; ``q_j = cdf_j - cdf_{j-1}``.  ``cdf_{N-1} = 2^52``
    exactly, so the weights ``q 2^-52`` sum to exactly 1 and ``np.cumsum`` of them is ``cdf 2^-52`` bit for
    bit in any summation order.  Returns ``(qq, a, A)``.  Python-int arithmetic: fine for the N <= ~1M the
    tests use; larger N go through ``c/auvpf_oracle.c``."""
    a = np.asarray(a, np.uint64)
    C = np.cumsum(a.astype(object))
    A = int(C[-1])
    if A == 0:
        raise FloatingPointError("degenerate weights")
    cdf = np.array([(int(c) << Q_BITS) // A for c in C], dtype=np.uint64)
    qq = np.diff(cdf, prepend=np.uint64(0))
    assert int(cdf[-1]) == 1 << Q_BITS
    return qq, a, A


def _sum_u64(q):
    """Exact sum of a uint64 array (32-bit limbs, no wrap for < 2^32 terms)."""
    lo = int((q & np.uint64(0xFFFFFFFF)).sum(dtype=np.uint64))
    hi = int((q >> np.uint64(32)).sum(dtype=np.uint64))
    return lo + (hi << 32)


def weights_from_q(qq):
    """``qq * 2^-52`` -- exact in float64."""
    return np.ldexp(qq.astype(np.float64), -Q_BITS)


def neff(weights):
    """``N_eff = 1/np.sum(np.square(weights))`` -- the block the reference carries commented out
    (auv_pf.py:174-178), for normalised weights."""
    w = np.asarray(weights, np.float64)
    return 1.0 / np.sum(np.square(w))


NEFF_SHIFT = 29


def neff_sums(a):
    """The exact integer sums of the library's N_eff gate over the 60-bit integer weights ``a``:
    ``s_i = a_i >> 29`` (<= 2^31), ``S1 = sum s_i``, ``S2 = sum s_i^2`` as two limb sums (low 32 bits, high
    bits).  Sums of integers: independent of the order and of the sharding."""
    s = np.asarray(a, np.uint64) >> np.uint64(NEFF_SHIFT)
    s2 = s * s                                            # <= 2^62, no wrap
    S1 = int(s.sum(dtype=np.uint64))
    lo = int((s2 & np.uint64(0xFFFFFFFF)).sum(dtype=np.uint64))
    hi = int((s2 >> np.uint64(32)).sum(dtype=np.uint64))
    return S1, lo, hi


def neff_fixed_from_sums(S1, S2_lo, S2_hi):
    """``float(S1)^2 / (float(S2_hi) 2^32 + float(S2_lo))`` in float64 (csrc/auvpf_math.cuh ``neff_fixed_a``;
    include/auvpf.h ``auvpf_set_neff_gate``): ``(sum w)^2 / sum w^2`` of the integer weights."""
    S2 = np.float64(S2_hi) * np.float64(4294967296.0) + np.float64(S2_lo)
    s1 = np.float64(S1)
    return float(s1 * s1 / S2) if S2 > 0 else 0.0


def neff_fixed_a(a):
    return neff_fixed_from_sums(*neff_sums(a))


def resample_gate(a, ratio):
    """True = resample (``N_eff < ratio * pc``, the ``if`` of auv_pf.py:181); ``a`` as returned by
    ``quantise_logw`` / ``quantise_linear``.  ratio == 0: always resample (the live reference code)."""
    if not ratio > 0.0:
        return True
    return neff_fixed_a(a) < ratio * float(len(a))


# --------------------------------------------------------------------------- resamplers
def systematic_resample(weights, u):
    """``systematic_resample`` (resampling.py:135-168) with ``random()`` = u.
    ``positions[i] < cumsum[j]`` two-pointer (:161-167) == searchsorted-right.
    Like the reference this is only defined when ``positions[-1] < cumsum[-1]``
    (otherwise the reference walks ``j`` off the end, :162-167)."""
    w = np.asarray(weights, np.float64)
    N = len(w)
    positions = (u + np.arange(N)) / float(N)                      # :157
    cumulative_sum = np.cumsum(w)                                  # :160
    idx = np.searchsorted(cumulative_sum, positions, side='right')
    if idx[-1] >= N:
        raise IndexError("positions[-1] >= cumsum[-1]: reference raises IndexError here")
    return idx.astype(np.int32)


def systematic_resample_to(weights, u, n_out):
    """``systematic_resample`` (resampling.py:135-168) generalised to ``n_out`` positions ``(u + i) / n_out`` over the
    same cumulative sum: the resample that changes the particle count (KLD adaptation).  ``n_out =
    len(weights)`` is the reference function."""
    w = np.asarray(weights, np.float64)
    positions = (u + np.arange(int(n_out))) / float(n_out)
    idx = np.searchsorted(np.cumsum(w), positions, side='right')
    if idx[-1] >= len(w):
        raise IndexError("positions[-1] >= cumsum[-1]")
    return idx.astype(np.int32)


def kld_bins(poses, bin_xy, bin_yaw):
    """Occupied bins of a cloud (n, 6) (KLD-sampling, Fox 2003): bin index = floor(coordinate * (1 / bin)) in x and y,
    floor((yaw + pi) * (1 / bin_yaw)) in yaw, yaw as the float32 the state holds."""
    p = np.asarray(poses, np.float64)
    ix = np.floor(p[:, 0] * (1.0 / bin_xy)).astype(np.int64)
    iy = np.floor(p[:, 1] * (1.0 / bin_xy)).astype(np.int64)
    iw = np.floor((p[:, 5].astype(np.float32).astype(np.float64) + np.pi) * (1.0 / bin_yaw)).astype(np.int64)
    return len(np.unique(np.stack([ix, iy, iw], axis=1), axis=0))


def kld_target(bins, eps, z_quantile, n_min, n_max):
    """n >= (k - 1) / (2 eps) (1 - 2 / (9 (k - 1)) + sqrt(2 / (9 (k - 1))) z)^3, clamped (Fox 2003, eq. 13); the same
    float64 operations as auvpf_kld_target."""
    n = n_min
    if bins > 1 and eps > 0.0:
        k1 = float(bins - 1)
        a = 2.0 / (9.0 * k1)
        t = 1.0 - a + math.sqrt(a) * z_quantile
        v = k1 / (2.0 * eps) * (t * t * t)
        n = n_max if v >= 9.0e18 else int(math.ceil(v))
    return max(n_min, min(n_max, n))


def stratified_resample(weights, uniforms):
    """``stratified_resample`` (resampling.py:80-114) with ``random(N)`` = uniforms."""
    w = np.asarray(weights, np.float64)
    N = len(w)
    positions = (np.asarray(uniforms, float) + np.arange(N)) / float(N)   # :103
    cumulative_sum = np.cumsum(w)
    idx = np.empty(N, np.int32)
    i, j = 0, 0
    while i < N:                                                   # :108-113 (positions not sorted in general? they are: strictly increasing)
        if positions[i] < cumulative_sum[j]:
            idx[i] = j
            i += 1
        else:
            j += 1
    return idx


def multinomial_resample(weights, uniforms):
    """``multinomial_resample`` (resampling.py:171-194)."""
    cumulative_sum = np.cumsum(np.asarray(weights, np.float64))
    cumulative_sum[-1] = 1.0
    return np.searchsorted(cumulative_sum, np.asarray(uniforms, float)).astype(np.int64)


def residual_copies(weights):
    """Deterministic part of ``residual_resample`` (resampling.py:61-66):
    ``floor(N w_i)`` copies of i in index order.  Returns ``(indexes_prefix, k)``."""
    w = np.asarray(weights, np.float64)
    N = len(w)
    num_copies = np.floor(N * w).astype(int)                       # :61
    return np.repeat(np.arange(N, dtype=np.int32), num_copies), num_copies


def residual_resample(weights, uniforms):
    """``residual_resample`` (resampling.py:27-76) literally, including the
    FilterPy-inherited ``residual = weights - num_copies`` (:70; should be
    ``N*w - num_copies``), the builtin sequential ``sum`` (:71) and the
    searchsorted on the resulting non-monotone "CDF" (:72-74).
    ``uniforms`` = the ``random(N-k)`` draw of :74 (at least N-k values)."""
    w = np.asarray(weights, np.float64)
    N = len(w)
    indexes = np.zeros(N, 'i')
    head, num_copies = residual_copies(w)
    k = len(head)
    indexes[:k] = head
    residual = w - num_copies                                      # :70
    residual /= sum(residual)                                      # :71
    cumulative_sum = np.cumsum(residual)                           # :72
    cumulative_sum[-1] = 1.0                                       # :73
    indexes[k:N] = np.searchsorted(cumulative_sum, np.asarray(uniforms, float)[:N - k])   # :74
    return indexes


RES_SHIFT = 14      # fractional parts live on a 2^-38 grid (csrc/auvpf_math.cuh kResShift)


def residual_resample_fixed(qq, uniforms):
    """The library's DEFAULT residual resampler on fixed-point weights ``w = qq 2^-52``
    (DESIGN.md 5.3): the deterministic part is the reference's, bit for bit --
    ``floor(N w_i)`` copies with the float64 product of resampling.py:61, written in index
    order (:62-66) -- and the remaining ``N - k`` slots are a multinomial draw on the
    fractional parts ``N w_i - k_i`` (what :70 means to compute; the reference's literal
    ``w - num_copies`` is reproduced by ``residual_resample`` above / AUVPF_RS_STRICT_REF).
    All-integer: ``r_i = (N q_i - k_i 2^52) >> 14``, inclusive scan ``S``, total ``R``; draw t
    picks the first i with ``floor(floor(u_t 2^53) R / 2^53) < S_i``."""
    qq = np.asarray(qq, np.uint64)
    N = len(qq)
    w = weights_from_q(qq)
    num_copies = np.floor(N * w).astype(np.int64)                                    # :61
    head = np.repeat(np.arange(N, dtype=np.int32), num_copies)                        # :62-66
    k = len(head)
    indexes = np.zeros(N, np.int32)
    indexes[:k] = head[:N]
    r = [max(0, int(q) * N - (int(c) << Q_BITS)) >> RES_SHIFT for q, c in zip(qq.tolist(), num_copies.tolist())]
    S = np.cumsum(np.array(r, dtype=object))
    R = int(S[-1]) if N else 0
    S64 = np.array([int(v) for v in S], dtype=np.uint64)
    u = np.asarray(uniforms, np.float64)[:N - k]
    T = np.array([(int(np.floor(np.ldexp(x, 53))) * R) >> 53 for x in u.tolist()], dtype=np.uint64)
    tail = np.searchsorted(S64, T, side='right')
    indexes[k:N] = np.minimum(tail, N - 1)
    return indexes


# --------------------------------------------------------------------------- resample bookkeeping
def resample_bookkeeping(indices, pc):
    """``auv_pf.resample`` lines auv_pf.py:183-187: ``keep`` = set of drawn ids,
    ``lost`` = ascending slots not in keep, ``dupes`` = the drawn list with the
    first occurrence of each kept id removed (order preserved)."""
    indices = np.asarray(indices)
    present = np.zeros(pc, bool)
    present[indices] = True
    lost = np.nonzero(~present)[0]
    _, first = np.unique(indices, return_index=True)
    is_first = np.zeros(len(indices), bool)
    is_first[first] = True
    dupes = indices[~is_first]
    return lost, dupes


def ancestor_map(indices, pc):
    """Effective in-place map of auv_pf.py:183-198: survivors keep their slot,
    ``pose[lost[r]] <- pose[dupes[r]]`` (:195-198)."""
    lost, dupes = resample_bookkeeping(indices, pc)
    A = np.arange(pc)
    n = min(len(lost), len(dupes))
    A[lost[:n]] = dupes[:n]
    return A


def resample_apply(poses, indices, res_noise_cov, noise):
    """``auv_pf.resample`` after the index draw (auv_pf.py:183-192): in-place
    reassignment, then ``add_noise`` on every particle in slot order."""
    poses = np.asarray(poses, float)
    A = ancestor_map(indices, poses.shape[0])
    return add_noise(poses[A], res_noise_cov, noise)


# --------------------------------------------------------------------------- estimate
def estimate(poses):
    """``auv_pf.update_loc_pose`` (auv_pf.py:218-252): unweighted mean pose,
    arithmetic mean of the wrapped yaws (:229-231), 3x3 position second moments
    with only the upper triangle and [1,0] filled (:238-246); returns
    ``(ave_pose(6) with [5] = that yaw mean, cov(3,3), quat_xyzw)``."""
    P = np.array(poses, dtype=np.float64, copy=True)
    ave = P.mean(axis=0)                                            # :221
    yaw = np.mean((P[:, 5] + np.pi) % (2 * np.pi) - np.pi)          # :229-231
    dx = P[:, 0:3] - ave[0:3]
    cov = np.zeros((3, 3))
    cov[0, 0] = (dx[:, 0] * dx[:, 0]).sum()
    cov[1, 1] = (dx[:, 1] * dx[:, 1]).sum()
    cov[2, 2] = (dx[:, 2] * dx[:, 2]).sum()
    cov[0, 1] = (dx[:, 0] * dx[:, 1]).sum()
    cov[0, 2] = (dx[:, 0] * dx[:, 2]).sum()
    cov[1, 2] = (dx[:, 1] * dx[:, 2]).sum()
    cov /= P.shape[0]                                               # :245
    cov[1, 0] = cov[0, 1]                                           # :246
    quat = tf_stub.quaternion_from_euler(ave[3], ave[4], yaw)       # :233
    out = ave.copy()
    out[5] = yaw
    return out, cov, quat


def estimate_options(poses, logw=None, circular_yaw=False, full_cov=False):
    """``estimate`` with the options of ``auvpf_estimate`` (include/auvpf.h; no reference counterpart beyond
    the unweighted form): weights ``exp(logw - max)`` normalised, circular yaw mean
    ``atan2(sum w sin yaw, sum w cos yaw)``, all nine covariance entries."""
    P = np.array(poses, dtype=np.float64, copy=True)
    if logw is None:
        w = np.full(P.shape[0], 1.0 / P.shape[0])
    else:
        lw = np.asarray(logw, np.float64)
        w = np.where(np.isnan(lw), 0.0, np.exp(lw - np.nanmax(lw)))
        w = w / w.sum()
    ave = (w[:, None] * P).sum(0)
    yw = (P[:, 5] + np.pi) % (2 * np.pi) - np.pi
    ave[5] = np.arctan2((w * np.sin(yw)).sum(), (w * np.cos(yw)).sum()) if circular_yaw else (w * yw).sum()
    dx = P[:, 0:3] - ave[0:3]
    cov = (w[:, None, None] * dx[:, :, None] * dx[:, None, :]).sum(0)
    if not full_cov:
        cov[2, 0] = cov[2, 1] = 0.0
    return ave, cov


def covariance36(cov):
    """The reference writes cov into the 36-vector with stride 3 (auv_pf.py:249-252)."""
    c = [0.0] * 36
    for i in range(3):
        for j in range(3):
            c[i * 3 + j] = cov[i, j]
    return np.array(c)


# --------------------------------------------------------------------------- Philox (spec for the device RNG)
_PH_M0, _PH_M1 = 0xD2511F53, 0xCD9E8D57
_PH_W0, _PH_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11; Random123).  ``counter``: (..., 4)
    uint32, ``key``: (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(counter[..., i], np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], np.uint64)
    k1 = np.asarray(key[..., 1], np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_PH_M0) * c[0]
        p1 = np.uint64(_PH_M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(_PH_W0)) & mask
        k1 = (k1 + np.uint64(_PH_W1)) & mask
    return np.stack(c, -1).astype(np.uint32)


def philox_normals(seed, stream, step, index, n_calls):
    """The device RNG contract: for particle ``index`` (global
    id, int64 array), call c = 0..n_calls-1 uses counter
    ``[index_lo, index_hi, step, stream*16 + c]`` and key ``[seed_lo, seed_hi]``; each call
    yields 4 normals by two Box-Muller pairs with ``u = (x + 0.5) * 2^-32``:
    ``(sqrt(-2 ln u0) cos(2 pi u1), sqrt(-2 ln u0) sin(2 pi u1))`` from words
    (0,1) then (2,3).  Returns (len(index), 4*n_calls) float64."""
    index = np.asarray(index, np.int64)
    out = []
    for c in range(n_calls):
        ctr = np.stack([(index & 0xFFFFFFFF).astype(np.uint32),
                        ((index >> 32) & 0xFFFFFFFF).astype(np.uint32),
                        np.full(index.shape, step & 0xFFFFFFFF, np.uint32),
                        np.full(index.shape, (stream * 16 + c) & 0xFFFFFFFF, np.uint32)], -1)
        key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32),
                              index.shape + (2,))
        x = philox4x32_10(ctr, key).astype(np.float64)
        u = (x + 0.5) * 2.0 ** -32
        for a, b in ((0, 1), (2, 3)):
            r = np.sqrt(-2.0 * np.log(u[..., a]))
            th = 2.0 * np.pi * u[..., b]
            out.append(r * np.cos(th))
            out.append(r * np.sin(th))
    return np.stack(out, -1)


def philox_pose_normals(seed, stream, step, index):
    """The six normals of one ``add_noise`` in POSE order ``[x, y, z, roll, pitch, yaw]``
    (device contract ``pose_normals``, csrc/auvpf_math.cuh): call 0 gives x, y (words 0,1) and
    yaw, z (words 2,3); call 1 gives roll, pitch (words 0,1)."""
    n = philox_normals(seed, stream, step, index, 2)
    return n[:, [0, 1, 3, 4, 5, 2]]

