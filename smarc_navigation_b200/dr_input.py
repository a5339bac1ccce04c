"""The dead-reckoning INPUT model in front of ``predict``.

The odometry twist the particle filter's ``predict`` consumes is produced by
``sam_dead_reckoning/scripts/dr_node.py::dr_timer`` (:165-246).  Its linear velocity is

* the latest DVL velocity, IF the message is fresh and plausible -- the gate of :179-182:
  ``t_now - t_dvl_prev < dvl_period and |v_y| < 0.2 and |v_x| < 1.5 and v_x > -0.1``;
* otherwise the motion-model estimate of :191-201: ``sam.motion(u)[0:3]`` (``sam_mm.py:30-123``: the
  acceleration ``inv(M) tau`` of the thruster command ``u = [rpm, dr]``), y negated, z zeroed, integrated once
  over ``dr_period``.

Host-side by nature (11 scalars per message; the reference is a 50 Hz timer callback); it exists here so that a
node built on ``AuvPfNode`` -- or a replayed log -- feeds ``predict`` / ``replay`` exactly what the reference's
producer would have fed it.  ``batch`` evaluates it for whole recorded sequences (and F vehicles) at once, the
form ``ParticleFilter.replay`` takes.
"""
import numpy as np

# sam_mm.SAM.__init__ (:12-16) and eom (:54-57,86-88)
SAM_MASS, SAM_IZZ, SAM_XG, SAM_YG, SAM_KT, SAM_D_SCALE, SAM_RPM_SCALE = 15.4, 1.6202, 0.4, 0.0, 0.3, -1.0, 1.0


def sam_motion(control):
    """``sam_mm.SAM.motion(control)`` (:30-125): ``inv(M) [F_T cos dr, -F_T sin dr, 0]`` with the mass matrix of
    :86-88; ``control = [rpm, dr]`` (broadcasts over leading axes).  Returns (..., 3)."""
    c = np.asarray(control, np.float64)
    rpm, dr = c[..., 0] * SAM_RPM_SCALE, c[..., 1] * SAM_D_SCALE
    M = np.array([[SAM_MASS, 0.0, -SAM_MASS * SAM_YG],
                  [0.0, SAM_MASS, SAM_MASS * SAM_XG],
                  [-SAM_MASS * SAM_YG, SAM_MASS * SAM_XG, SAM_IZZ]])
    F_T = SAM_KT * rpm
    tau = np.stack([F_T * np.cos(dr), -F_T * np.sin(dr), np.zeros_like(F_T)], -1)
    return tau @ np.linalg.inv(M).T


def dvl_gate(t_now, t_dvl_prev, dvl_velocity, dvl_period=0.2):
    """The gate of dr_node.py:179-182 (vectorised): True where the DVL velocity is integrated."""
    v = np.asarray(dvl_velocity, np.float64)
    return ((np.asarray(t_now) - np.asarray(t_dvl_prev)) < dvl_period) & (np.abs(v[..., 1]) < 0.2) & \
        (np.abs(v[..., 0]) < 1.5) & (v[..., 0] > -0.1)


def linear_velocity(t_now, t_dvl_prev, dvl_velocity, control, dvl_period=0.2, dr_period=0.02):
    """``lin_vel_t`` of dr_timer (:176-201) for one message or a whole batch: DVL velocity where the gate
    passes, else the motion-model acceleration ``[a_x, -a_y, 0] * dr_period`` (:195-200).  Returns
    ``(lin_vel (..., 3), used_dvl (...))``."""
    v = np.asarray(dvl_velocity, np.float64)
    ok = dvl_gate(t_now, t_dvl_prev, v, dvl_period)
    acc = sam_motion(control)
    mm = np.stack([acc[..., 0], -acc[..., 1], np.zeros_like(acc[..., 0])], -1) * dr_period
    return np.where(ok[..., None], v, mm), ok


def odom_rows(t_now, t_dvl_prev, dvl_velocity, control, yaw_rate, quat_xyzw, depth, dvl_period=0.2, dr_period=0.02):
    """ABI odometry rows ``[vx vy vz wz qx qy qz qw z dt 0]`` (include/auvpf.h ``auvpf_predict``) of a recorded
    sequence, any leading shape (e.g. (n_cycles, n_steps, F)): what dr_timer publishes in the Odometry twist /
    pose (:212-226) and ``Particle.motion_pred`` reads (auv_particle.py:43-70)."""
    lv, ok = linear_velocity(t_now, t_dvl_prev, dvl_velocity, control, dvl_period, dr_period)
    q = np.asarray(quat_xyzw, np.float64)
    shape = lv.shape[:-1]
    rows = np.zeros(shape + (11,))
    rows[..., 0:3] = lv
    rows[..., 3] = np.broadcast_to(np.asarray(yaw_rate, np.float64), shape)
    rows[..., 4:8] = np.broadcast_to(q, shape + (4,))
    rows[..., 8] = np.broadcast_to(np.asarray(depth, np.float64), shape)
    rows[..., 9] = dr_period
    return rows, ok


class DeadReckoningInput(object):
    """Message-by-message form for a live node: keeps the latest DVL message and thruster command like
    dr_node (``dvl_cb`` :305-307, ``thrust_cb`` :248-251) and hands ``AuvPfNode.odom_callback`` the twist."""

    def __init__(self, dvl_period=0.2, dr_period=0.02):
        self.dvl_period, self.dr_period = float(dvl_period), float(dr_period)
        self.t_dvl_prev = -np.inf
        self.dvl_velocity = np.zeros(3)
        self.u = np.zeros(2)
        self.n_dvl = self.n_model = 0

    def dvl_cb(self, t, velocity_xyz):
        self.t_dvl_prev, self.dvl_velocity = float(t), np.asarray(velocity_xyz, np.float64)

    def thrust_cb(self, rpm_sum, rudder_rad):
        self.u = np.array([float(rpm_sum), float(np.clip(-rudder_rad, -7 * np.pi / 180, 7 * np.pi / 180))])   # :249-251

    def velocity(self, t_now):
        lv, ok = linear_velocity(t_now, self.t_dvl_prev, self.dvl_velocity, self.u, self.dvl_period, self.dr_period)
        if bool(ok):
            self.n_dvl += 1
        else:
            self.n_model += 1
        return lv

    def odometry(self, t_now, quat_xyzw, yaw_rate, depth):
        """The odometry sample dr_timer publishes for this tick (:212-226) in the plain form ``predict`` accepts
        (``v``, ``wz``, ``quat``, ``z``, ``dt``: what ``Particle.motion_pred`` reads, auv_particle.py:43-70)."""
        from .world import Odom
        return Odom(self.velocity(t_now), float(yaw_rate), np.asarray(quat_xyzw, np.float64), float(depth), self.dr_period)
