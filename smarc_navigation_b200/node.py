"""ROS-free shell of the particle-filter node.

``AuvPfNode`` keeps the callbacks, parameter names, gating and published numbers of
``auv_particle_filter/scripts/auv_pf.py`` (class ``auv_pf``, :23-285) around the GPU
``ParticleFilter``; everything ROS-specific is injected, so the same class runs under
``rospy`` (pass ``rospy.get_param`` and the publishers' ``publish`` methods) and in the test
suite with fake messages (``tests/test_node_shell.py``).  A maintainer's ``auv_pf_gpu.py`` is
then ~20 lines: ``rospy.init_node``, look up the map->odom transform (:76-87), build this
object, subscribe the four callbacks, start the 10 Hz timer (:114) and ``rospy.spin()``.

What differs from the reference, deliberately:

* the callbacks take a lock (the reference's three rospy threads race on ``self.particles``,
  SURVEY.md §5);
* the utm->map ``transformPoint`` of :143-151 is loop-invariant and is done ONCE per fix by the
  injected ``utm_to_map`` callable, not once per particle;
* an MBES subscriber (``mbes_cb``) exists: a ``sensor_msgs/LaserScan``-shaped ping (the only
  MBES wire format in the reference, mbes_mapper/mbes_receptor.hpp:26,82) updates the weights by
  raycasting into the DEM.  The reference snapshot only has the GPS update.
"""
import threading

import numpy as np

from .filter import ParticleFilter

# parameter names and defaults of auv_pf.__init__ (auv_pf.py:27-56,64,71,101,106,110); the
# covariance parameters have NO default in the reference (rospy.get_param raises) -- same here
DEFAULTS = {
    "particle_count": 10,                      # :27
    "map_frame": "map",                        # :28
    "base_frame": "base_link",                 # :29
    "utm_frame": "utm",                        # :30
    "odom_frame": "sam/odom",                  # :31
    "measurement_std": 0.01,                   # :39
    "particle_poses_topic": "/particle_poses",  # :64
    "odom_corrected_topic": "/average_pose",   # :71
    "aux_dive": "/dive",                       # :101
    "gps_odom_topic": "/gps",                  # :106
    "odom_topic": "odom",                      # :110
    # additions of the GPU path
    "resampler": "residual",                   # what auv_pf.resample calls (:182)
    "inplace_reference": True,                 # :183-198
    "mbes_topic": "/mbes/scan",
    "range_min": 0.5,
    "range_max": 120.0,
    "seed": 0,
    "device": 0,
    "resample_neff_ratio": 0.0,                # 0 = resample at every fix like the live code; 0.5 = the gate
                                               # the reference carries commented out (:174-181)
    "publish_max_particles": 2000,             # PoseArray size: a strided subsample built on the device (the reference
                                               # publishes every particle, :264-285; >= particle_count does the same)
    "estimate_circular_yaw": False,            # True: circular yaw mean instead of the mean of wrapped yaws (:229-231)
    "kld_adaptive": False,                     # True: every MBES resample draws Fox's KLD bound of particles (the
                                               # reference's particle_count, :27, is then the ceiling)
    "kld_bin_xy": 0.5,                         # [m]
    "kld_bin_yaw": 0.0873,                     # [rad] (5 degrees)
    "kld_epsilon": 0.05,
    "kld_min_particles": 1000,
}


def quaternion_from_euler(roll, pitch, yaw):
    """tf.transformations.quaternion_from_euler, axes 'sxyz', (x, y, z, w) -- used for the
    published orientation only (auv_pf.py:233)."""
    cr, sr = np.cos(roll * 0.5), np.sin(roll * 0.5)
    cp, sp = np.cos(pitch * 0.5), np.sin(pitch * 0.5)
    cy, sy = np.cos(yaw * 0.5), np.sin(yaw * 0.5)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy])


class AuvPfNode(object):
    def __init__(self, get_param, m2o_matrix, utm_to_map=None, publish_estimate=None, publish_particles=None,
                 now=None, filter_factory=ParticleFilter):
        """``get_param(name, default=<required>)`` is ``rospy.get_param`` with the leading
        ``~`` stripped; ``m2o_matrix`` the looked-up map<-odom transform (:76-81);
        ``utm_to_map(x, y) -> (x, y)`` the TF listener (:143-151)."""
        def param(name):
            if name in DEFAULTS:
                return get_param(name, DEFAULTS[name])
            return get_param(name)
        self.pc = int(param("particle_count"))
        self.map_frame, self.base_frame = param("map_frame"), param("base_frame")
        self.utm_frame, self.odom_frame = param("utm_frame"), param("odom_frame")
        self.res_noise_cov = param("resampling_noise_covariance")              # :52-56
        self.pf = filter_factory(particle_count=self.pc, init_covariance=param("init_covariance"),
                                 motion_covariance=param("motion_covariance"),
                                 resampling_noise_covariance=self.res_noise_cov,
                                 measurement_std=float(param("measurement_std")), m2o_matrix=m2o_matrix,
                                 resampler=param("resampler"), inplace_reference=bool(param("inplace_reference")),
                                 range_min=float(param("range_min")), range_max=float(param("range_max")),
                                 seed=int(param("seed")), device=int(param("device")),
                                 resample_neff_ratio=float(param("resample_neff_ratio")))
        self.utm_to_map = utm_to_map or (lambda x, y: (x, y))
        self.publish_estimate = publish_estimate or (lambda est: None)
        self.publish_particles = publish_particles or (lambda poses: None)
        self.now = now or (lambda: 0.0)
        self.time = self.now()                                                   # :97-98
        self.old_time = self.now()
        self.diving = True                                                       # :103
        self.odom_latest = None
        self.cov = np.zeros((3, 3))
        self._lock = threading.Lock()
        self._scan_geometry = None
        self.updates = 0
        self.publish_max = int(param("publish_max_particles"))
        self.circular_yaw = bool(param("estimate_circular_yaw"))
        self.kld = bool(param("kld_adaptive"))
        self.kld_args = dict(bin_xy=float(param("kld_bin_xy")), bin_yaw=float(param("kld_bin_yaw")),
                             eps=float(param("kld_epsilon")), n_min=int(param("kld_min_particles")))

    # -- map / sensor (GPU additions) ---------------------------------------------------------
    def set_map(self, z, x0=0.0, y0=0.0, res=1.0):
        self.pf.set_dem(z, x0, y0, res)

    # -- callbacks ------------------------------------------------------------------------------
    def dive_cb(self, dive_msg):                                                 # :122-123
        self.diving = bool(dive_msg.data)

    def odom_callback(self, odom_msg):                                           # :201-211
        with self._lock:
            self.time = odom_msg.header.stamp.to_sec()
            self.odom_latest = odom_msg
            if self.old_time and self.time > self.old_time:
                self.predict(odom_msg)
            self.old_time = self.time

    def predict(self, odom_t):                                                   # :213-216
        self.pf.predict(odom_t, dt=self.time - self.old_time)

    def gps_odom_cb(self, gps_odom):                                             # :125-133
        with self._lock:
            if self.old_time and self.time > self.old_time:
                if not self.diving:                                              # GPS only at the surface
                    weights = self.update(gps_odom)
                    self.resample(weights)

    def update(self, gps_odom):                                                  # :135-167
        p = gps_odom.pose.pose.position
        gx, gy = self.utm_to_map(p.x, p.y)
        self.updates += 1
        return self.pf.update_gps((gx, gy))

    def resample(self, weights):                                                 # :169-192
        self.pf.resample(weights)

    def mbes_cb(self, scan):
        """A ping: ``scan.ranges``, ``scan.angle_min``, ``scan.angle_increment``."""
        with self._lock:
            geom = (float(scan.angle_min), float(scan.angle_increment), len(scan.ranges))
            if geom != self._scan_geometry:
                self.pf.set_scan_geometry(*geom)
                self._scan_geometry = geom
            self.pf.update_mbes(np.asarray(scan.ranges, np.float32), return_weights=False)
            if self.kld:       # the cloud's occupied bins decide how many particles the resample draws (<= particle_count)
                self.pf.resample(new_count=int(self.pf.kld_particle_count(**self.kld_args)[0]))
            else:
                self.pf.resample()
            self.updates += 1

    def loc_loop(self, event=None):                                              # :264-285
        """10 Hz timer: publish the particle poses and the mean estimate."""
        with self._lock:
            # (n, 7) [x y z qx qy qz qw]: the Pose entries of the reference's PoseArray (:267-275), for a strided
            # subsample assembled on the device -- not 48 MB of raw state per tick at 1 M particles
            poses = self.pf.pose_array(max_poses=self.publish_max)[0]
            est = self.update_loc_pose()
        self.publish_particles(poses)
        self.publish_estimate(est)
        return est

    def update_loc_pose(self):                                                   # :218-260
        mean, cov = self.pf.estimate(circular_yaw=True) if self.circular_yaw else self.pf.mean_cov()
        ave_pose, self.cov = mean[0], cov[0]
        quat_t = quaternion_from_euler(ave_pose[3], ave_pose[4], ave_pose[5])    # :233 (yaw = mean of wrapped yaws)
        covariance = [0.0] * 36
        for i in range(3):
            for j in range(3):
                covariance[i * 3 + j] = self.cov[i, j]                           # :249-252 (stride 3, as the reference writes it)
        return {"position": ave_pose[0:3].copy(), "orientation": quat_t, "covariance": covariance,
                "frame_id": self.odom_frame, "child_frame_id": self.base_frame,
                "tf": {"translation": [ave_pose[0], ave_pose[1], 0.0], "rotation": quat_t,     # :256-260
                       "child": self.base_frame, "parent": self.odom_frame}}

    def close(self):
        self.pf.close()
