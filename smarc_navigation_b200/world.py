"""Synthetic bathymetry, vehicle trajectory, DVL/IMU odometry and MBES fan.

Shared by ``bench.py``, ``__graft_entry__.smoke()`` and the tests so that every
arm (GPU library, CPU oracle) sees the same seeded inputs.
Pure NumPy; generates INPUTS only -- the ping ranges themselves come from a
raycast of the ground-truth pose, done by whoever consumes the world (the GPU
library in ``bench.py``, the oracle in the CPU tests).

Seeds: map 1, trajectory 2, ping noise 3, particle init 4, process/roughening
noise 5, resample uniforms 6.
"""
import math

import numpy as np

CONFIGS = {
    # name: particles, beams, DEM vertices per side, resolution [m]
    "C1": dict(particles=100, beams=256, dem_n=500, res=2.0),
    "C2": dict(particles=65536, beams=256, dem_n=500, res=2.0),
    "C3": dict(particles=1048576, beams=512, dem_n=4000, res=0.5),
    "C4": dict(particles=16777216, beams=512, dem_n=4000, res=0.5),       # 8 GPUs
    "C5": dict(particles=65536, beams=256, dem_n=500, res=2.0, filters=256),
}

# Filter parameters of the MBES case; names as in auv_pf.launch:15-20.
MBES_PARAMS = dict(
    init_covariance=[4.0, 4.0, 0.0, 0.0, 0.0, 0.01],
    motion_covariance=[1e-4, 1e-4, 0.0, 0.0, 0.0, 1e-6],
    resampling_noise_covariance=[0.01, 0.01, 0.0, 0.0, 0.0, 1e-5],
    measurement_std=1.0,
)
RANGE_MAX = 120.0
RANGE_MIN = 0.5


def make_dem(n, res, seed=1, chunk=512):
    """fp32 (n, n) heights: six-octave sinusoid field + white roughness."""
    rng = np.random.default_rng(seed)
    k = np.arange(6)
    lam = 400.0 * 2.0 ** (-k)
    amp = 4.0 * 2.0 ** (-0.8 * k)
    phi = rng.uniform(0.0, 2.0 * np.pi, 6)
    psi = rng.uniform(0.0, 2.0 * np.pi, 6)
    z = np.empty((n, n), np.float32)
    xs = np.arange(n, dtype=np.float64) * res
    for i0 in range(0, n, chunk):
        ys = (np.arange(i0, min(n, i0 + chunk), dtype=np.float64) * res)[:, None]
        acc = np.full((ys.shape[0], n), 30.0)
        for a, l, p, s in zip(amp, lam, phi, psi):
            acc += a * np.sin(2.0 * np.pi * (xs[None, :] * math.cos(p) + ys * math.sin(p)) / l + s)
        z[i0:i0 + ys.shape[0]] = (-acc).astype(np.float32)
    z += rng.normal(0.0, 0.05, size=(n, n)).astype(np.float32)
    return {"z": z, "x0": 0.0, "y0": 0.0, "res": float(res)}


def dem_to_mesh(dem):
    """The DEM's own triangulation as a mesh (vertices (H*W, 3), faces (2 (H-1)(W-1), 3)): two triangles
    per cell split along (i, j) -> (i+1, j+1), the surface of the raycast spec."""
    z = dem["z"]
    H, W = z.shape
    jj, ii = np.meshgrid(np.arange(W), np.arange(H))
    v = np.stack([dem["x0"] + jj * dem["res"], dem["y0"] + ii * dem["res"], z], -1).reshape(-1, 3).astype(np.float32)
    idx = (ii * W + jj)[:-1, :-1].reshape(-1)
    lower = np.stack([idx, idx + 1, idx + W + 1], -1)
    upper = np.stack([idx, idx + W + 1, idx + W], -1)
    return v, np.concatenate([lower, upper]).astype(np.int32)


def make_beams(B, half_angle_deg=60.0):
    """Fan in the sensor y-z plane: d_b = (0, sin th, -cos th)."""
    th = np.linspace(-math.radians(half_angle_deg), math.radians(half_angle_deg), B)
    return np.stack([np.zeros(B), np.sin(th), -np.cos(th)], -1)


def quat_from_rpy(roll, pitch, yaw):
    """[x, y, z, w] of Rz(yaw) Ry(pitch) Rx(roll) (tf 'sxyz')."""
    cr, sr = math.cos(roll / 2), math.sin(roll / 2)
    cp, sp = math.cos(pitch / 2), math.sin(pitch / 2)
    cy, sy = math.cos(yaw / 2), math.sin(yaw / 2)
    return np.array([sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy,
                     cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy])


def _rot_rows01(roll, pitch, yaw):
    cy, sy = math.cos(yaw), math.sin(yaw)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cr, sr = math.cos(roll), math.sin(roll)
    return (np.array([cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr]),
            np.array([sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr]))


class Odom(object):
    """Plain-array odometry sample: the scalars ``Particle.motion_pred`` reads
    (auv_particle.py:43-70)."""
    __slots__ = ("v", "wz", "quat", "z", "dt")

    def __init__(self, v, wz, quat, z, dt):
        self.v, self.wz, self.quat, self.z, self.dt = v, wz, quat, z, dt


class Trajectory(object):
    """Lawnmower survey: 1 m/s body-x, 180 m legs, +-0.1 rad/s turns, 50 Hz odom."""

    def __init__(self, extent, seed=2, start=None, predicts_per_cycle=5, dt=0.02, offset=(0.0, 0.0)):
        self.rng = np.random.default_rng(seed)
        self.dt = dt
        self.P = predicts_per_cycle
        s = 0.3 * extent if start is None else start
        self.truth = np.array([s + offset[0], s + offset[1], -5.0, 0.0, 0.0, 0.0])   # x y z r p yaw
        self.t = 0.0
        self.leg_s = 0.0
        self.turning = 0.0          # remaining turn angle (signed)
        self.turn_dir = 1.0

    def step(self):
        """Advance ground truth by dt; return the (noisy) Odom for that step."""
        dt = self.dt
        v_true = np.array([1.0, 0.0, 0.0])
        if self.turning != 0.0:
            wz = 0.1 * self.turn_dir
            self.turning -= abs(wz) * dt
            if self.turning <= 0.0:
                self.turning = 0.0
                self.turn_dir = -self.turn_dir
                self.leg_s = 0.0
        else:
            wz = 0.0
            self.leg_s += v_true[0] * dt
            if self.leg_s >= 180.0:
                self.turning = math.pi
        self.t += dt
        roll = 0.02 * math.sin(0.2 * self.t)
        pitch = 0.03 * math.sin(0.13 * self.t)
        yaw = self.truth[5] + wz * dt
        yaw = (yaw + math.pi) % (2 * math.pi) - math.pi
        R0, R1 = _rot_rows01(roll, pitch, yaw)
        self.truth[0] += float(R0 @ v_true) * dt
        self.truth[1] += float(R1 @ v_true) * dt
        self.truth[3:6] = (roll, pitch, yaw)
        v_meas = v_true + self.rng.normal(0.0, 0.02, 3)
        wz_meas = wz + 1e-3
        return Odom(v_meas, wz_meas, quat_from_rpy(roll, pitch, yaw), self.truth[2], dt)

    def cycle(self):
        """P odometry samples, then the ground-truth pose at ping time."""
        odoms = [self.step() for _ in range(self.P)]
        return odoms, self.truth.copy()


def ping_noise(B, cycle, seed=3, std=0.05):
    return np.random.default_rng([seed, cycle]).normal(0.0, std, B)
