"""Time the UNMODIFIED reference filter on the host CPU (CPU-A, CPU-B).

TEST / BENCH INFRASTRUCTURE (see ``oracle/__init__.py``): only ``bench.py``'s ``cpu_baseline`` and
``--impl reference`` legs call this.  It executes the reference as it lies under
``/root/reference/auv_particle_filter/scripts`` -- ``auv_particle.Particle`` and
``resampling.residual_resample`` imported unmodified, the ``auv_pf`` class body lifted with ``ast``
(``tests/golden/make_golden.py::lift_auv_pf``; no source is copied) -- on the synthetic sequence of
``smarc_navigation_b200/world.py``.  The reference checkout exists only in the build container, not on
the GPU box, so ``run`` reports ``reachable: False`` there; the numbers measured in the build container
are committed as ``profiles/r2_reference_python_c1.json`` (with the CPU they were taken on) and
``bench.py`` attaches that record, labelled as such, when the live run is impossible.

CPU-A  what the reference really does: P predicts + GPS update (``auv_pf.update``, scipy's
       ``multivariate_normal.pdf`` per particle) + ``auv_pf.resample`` (residual_resample,
       keep/lost/dupes, ``add_noise``), C1 size (100 particles), 200 cycles, one thread.
CPU-B  BASELINE.json ``configs[0]``: the same reference predict / resample code, but the update is
       the MBES form -- every particle's ``get_p_pose()`` (reference code), 256 beams raycast into the
       2 m DEM by the C oracle on ONE thread, independent-beam Gaussian, weights handed to the
       reference's ``resample``.  The reference has no raycaster, so this row is
       "reference + oracle", not the reference alone.
"""
import json
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RECORD = os.path.join(ROOT, "profiles", "r2_reference_python_c1.json")


def reachable():
    from . import ref_shim
    return ref_shim.available()


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


def _stats(ms):
    ms = np.asarray(ms)
    return {"median_ms": float(np.median(ms)), "p99_ms": float(np.percentile(ms, 99)), "mean_ms": float(ms.mean())}


def run(cycles=200, n=100, beams=256, predicts=5, mbes=True):
    """Returns the record dict (``reachable: False`` and nothing else when the checkout is absent)."""
    if not reachable():
        return {"reachable": False, "why": "/root/reference is not present on this machine"}
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import scipy
    from . import c_oracle, ref_shim
    from smarc_navigation_b200 import world
    from tests.golden import make_golden as mg          # executes ref_shim.load(): the unmodified modules
    cls = mg.lift_auv_pf()
    cfg = world.CONFIGS["C1"]
    dem = world.make_dem(cfg["dem_n"], cfg["res"])
    bd = world.make_beams(beams)
    out = {"reachable": True, "kind": "reference-python", "particles": n, "beams": beams, "cycles": cycles,
           "predicts_per_cycle": predicts, "threads": 1, "cpu_model": _cpu_model(), "os_cpu_count": os.cpu_count(),
           "python": platform.python_version(), "numpy": np.__version__, "scipy": scipy.__version__}
    c_oracle.set_num_threads(1)
    for leg in (("cpu_a_gps",) + (("cpu_b_mbes",) if mbes else ())):
        P = world.MBES_PARAMS if leg == "cpu_b_mbes" else dict(
            init_covariance=[0.1, 0.1, 0.0, 0.0, 0.0, 0.0], motion_covariance=[1e-3, 1e-3, 0.0, 0.0, 0.0, 1e-5],
            resampling_noise_covariance=[1.0, 1.0, 0.0, 0.0, 0.0, 1e-4], measurement_std=1.0)     # auv_pf.launch:17-20
        np.random.seed(5)
        traj = world.Trajectory(cfg["dem_n"] * cfg["res"])
        pf = mg.make_filter(cls, n, np.identity(4), P["init_covariance"], P["measurement_std"],
                            P["motion_covariance"], P["resampling_noise_covariance"])
        for p in pf.particles:                        # start the cloud at the vehicle
            p.p_pose = np.asarray(p.p_pose, float) + traj.truth
        t_pred, t_upd, t_res, t_cyc = [], [], [], []
        for c in range(cycles):
            odoms, truth = traj.cycle()
            t0 = time.perf_counter()
            for od in odoms[:predicts]:
                pf.time, pf.old_time = od.dt, 0.0     # predict: dt = time - old_time (auv_pf.py:214)
                pf.predict(ref_shim.make_odom(od.v, od.wz, od.quat, od.z))
            t1 = time.perf_counter()
            if leg == "cpu_a_gps":
                fix = truth[0:2] + np.random.normal(0, 0.5, 2)
                msg = ref_shim.Msg(pose=ref_shim.Msg(pose=ref_shim.Msg(position=ref_shim.Msg(x=fix[0], y=fix[1]))))
                w = pf.update(msg)                    # auv_pf.py:135-167
            else:
                r_true = c_oracle.update_mbes(dem, truth[None, :], np.eye(4), bd, None, 1.0, world.RANGE_MIN,
                                              world.RANGE_MAX, want_ranges=True)["ranges"][0]
                meas = r_true + world.ping_noise(beams, c)
                t1 = time.perf_counter()              # the synthetic ping is not the filter's work
                poses = np.array([np.asarray(p.p_pose, float) for p in pf.particles])
                for p in pf.particles:
                    p.get_p_pose()                    # the reference's own pose -> (p, R) (auv_particle.py:72-83)
                logw = c_oracle.update_mbes(dem, poses, np.eye(4), bd, meas, P["measurement_std"], world.RANGE_MIN,
                                            world.RANGE_MAX)["logw"]
                w = np.exp(logw - logw.max()) + 1.0e-200          # auv_pf.py:165
            t2 = time.perf_counter()
            pf.resample(w)                            # auv_pf.py:169-192
            t3 = time.perf_counter()
            t_pred.append(1e3 * (t1 - t0))
            t_upd.append(1e3 * (t2 - t1))
            t_res.append(1e3 * (t3 - t2))
            t_cyc.append(1e3 * ((t1 - t0) + (t3 - t1)))
        rec = {"cycle": _stats(t_cyc), "predict": _stats(t_pred), "update": _stats(t_upd), "resample": _stats(t_res)}
        if leg == "cpu_b_mbes":
            rec["evals_per_s"] = n * beams / (np.median(t_upd) * 1e-3)
            rec["evals_per_s_whole_cycle"] = n * beams / (np.median(t_cyc) * 1e-3)
        mean = np.mean([np.asarray(p.p_pose, float)[0:2] for p in pf.particles], axis=0)
        rec["track_error_m"] = float(np.hypot(*(mean - traj.truth[0:2])))
        out[leg] = rec
    return out


def recorded():
    """The build-container measurement committed under profiles/ (None when absent)."""
    try:
        return json.load(open(RECORD))
    except (OSError, ValueError):
        return None


if __name__ == "__main__":
    r = run()
    r["where"] = "build container (no GPU); committed because /root/reference does not travel to the GPU box"
    r["when"] = time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())
    print(json.dumps(r, indent=1))
    if r.get("reachable"):
        os.makedirs(os.path.dirname(RECORD), exist_ok=True)
        json.dump(r, open(RECORD, "w"), indent=1)
