"""Stand-in for the four ``tf.transformations`` functions the reference uses.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

``tf.transformations`` (ROS ``geometry``/``tf`` package, C. Gohlke's
``transformations.py``) is a third-party dependency that is NOT under
``/root/reference`` (declared only as ``tf`` in
``auv_particle_filter/package.xml:57,71``; CI distros melodic/noetic, i.e.
tf 1.12/1.13 -- version otherwise unpinned).  Its published algorithm is
restated here for the call sites on the hot path:

* ``euler_from_quaternion``   -- ``auv_particle.py:50``
* ``quaternion_from_euler``   -- ``auv_particle.py:75``, ``auv_pf.py:233,272``
* ``translation_matrix``      -- ``auv_particle.py:74,122``
* ``quaternion_matrix``       -- ``auv_particle.py:76,123``

Conventions (ROS flavour of the file): quaternions are ``[x, y, z, w]``, the
default axes string is ``'sxyz'`` (static frame, roll about x, then pitch about
y, then yaw about z), so ``R = Rz(yaw) @ Ry(pitch) @ Rx(roll)``.  Only the
``'sxyz'`` convention is implemented; ``tests/test_oracle_tf.py`` cross-checks
it against ``scipy.spatial.transform.Rotation``.
"""
import math

import numpy as np

_EPS = np.finfo(float).eps * 4.0


def translation_matrix(direction):
    M = np.identity(4)
    M[:3, 3] = direction[:3]
    return M


def translation_from_matrix(matrix):
    return np.array(matrix, copy=False)[:3, 3].copy()


def quaternion_from_euler(ai, aj, ak, axes='sxyz'):
    if axes != 'sxyz':
        raise NotImplementedError("only 'sxyz' is used by the reference")
    ai, aj, ak = ai / 2.0, aj / 2.0, ak / 2.0
    ci, si = math.cos(ai), math.sin(ai)
    cj, sj = math.cos(aj), math.sin(aj)
    ck, sk = math.cos(ak), math.sin(ak)
    cc, cs = ci * ck, ci * sk
    sc, ss = si * ck, si * sk
    q = np.empty((4,), dtype=np.float64)
    q[0] = cj * sc - sj * cs
    q[1] = cj * ss + sj * cc
    q[2] = cj * cs - sj * sc
    q[3] = cj * cc + sj * ss
    return q


def quaternion_matrix(quaternion):
    q = np.array(quaternion[:4], dtype=np.float64, copy=True)
    nq = np.dot(q, q)
    if nq < _EPS:
        return np.identity(4)
    q *= math.sqrt(2.0 / nq)
    q = np.outer(q, q)
    return np.array((
        (1.0 - q[1, 1] - q[2, 2], q[0, 1] - q[2, 3], q[0, 2] + q[1, 3], 0.0),
        (q[0, 1] + q[2, 3], 1.0 - q[0, 0] - q[2, 2], q[1, 2] - q[0, 3], 0.0),
        (q[0, 2] - q[1, 3], q[1, 2] + q[0, 3], 1.0 - q[0, 0] - q[1, 1], 0.0),
        (0.0, 0.0, 0.0, 1.0)), dtype=np.float64)


def euler_from_matrix(matrix, axes='sxyz'):
    if axes != 'sxyz':
        raise NotImplementedError("only 'sxyz' is used by the reference")
    M = np.array(matrix, dtype=np.float64, copy=False)[:3, :3]
    i, j, k = 0, 1, 2
    cy = math.sqrt(M[i, i] * M[i, i] + M[j, i] * M[j, i])
    if cy > _EPS:
        ax = math.atan2(M[k, j], M[k, k])
        ay = math.atan2(-M[k, i], cy)
        az = math.atan2(M[j, i], M[i, i])
    else:
        ax = math.atan2(-M[j, k], M[j, j])
        ay = math.atan2(-M[k, i], cy)
        az = 0.0
    return ax, ay, az


def euler_from_quaternion(quaternion, axes='sxyz'):
    return euler_from_matrix(quaternion_matrix(quaternion), axes)


def quaternion_from_matrix(matrix):
    """Imported (unused) by ``auv_particle.py:10``; trace method, [x,y,z,w]."""
    M = np.array(matrix, dtype=np.float64, copy=False)[:4, :4]
    q = np.empty((4,), dtype=np.float64)
    t = np.trace(M)
    if t > M[3, 3]:
        q[3] = t
        q[2] = M[1, 0] - M[0, 1]
        q[1] = M[0, 2] - M[2, 0]
        q[0] = M[2, 1] - M[1, 2]
    else:
        i, j, k = 0, 1, 2
        if M[1, 1] > M[0, 0]:
            i, j, k = 1, 2, 0
        if M[2, 2] > M[i, i]:
            i, j, k = 2, 0, 1
        t = M[i, i] - (M[j, j] + M[k, k]) + M[3, 3]
        q[i] = t
        q[j] = M[i, j] + M[j, i]
        q[k] = M[k, i] + M[i, k]
        q[3] = M[k, j] - M[j, k]
    q *= 0.5 / math.sqrt(t * M[3, 3])
    return q


def identity_matrix():
    return np.identity(4)
