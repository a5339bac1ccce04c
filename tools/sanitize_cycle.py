#!/usr/bin/env python
"""A small cycle through every kernel family, for `compute-sanitizer --tool memcheck|racecheck|synccheck`:
    gpurun -- compute-sanitizer --tool memcheck python tools/sanitize_cycle.py
Sizes are tiny on purpose (the tools slow kernels down 10-100x)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smarc_navigation_b200 import ParticleFilter, world  # noqa: E402


def main():
    dem = world.make_dem(200, 2.0)
    B, N, F = 64, 3000, 2
    P = world.MBES_PARAMS
    traj = world.Trajectory(400.0)
    for resampler, inplace, strict in (("systematic", False, False), ("systematic", True, False),
                                       ("residual", False, False), ("residual", True, True)):
        pf = ParticleFilter(particle_count=N, n_filters=F, init_covariance=P["init_covariance"],
                            motion_covariance=P["motion_covariance"],
                            resampling_noise_covariance=P["resampling_noise_covariance"], measurement_std=1.0,
                            range_min=world.RANGE_MIN, range_max=world.RANGE_MAX, init_mean=traj.truth, seed=3,
                            resampler=resampler, inplace_reference=inplace, strict_reference=strict)
        pf.set_dem(dem)
        pf.set_beams(world.make_beams(B))
        odoms, truth = traj.cycle()
        pf.predict_many(odoms)
        pf.predict(odoms[0])
        r = pf.raycast(np.stack([truth, truth + [50.0, 0, 0, 0.3, 0, 0]]))          # second pose: large roll -> per-ray path
        pf.update_mbes(np.stack([r[0], r[0]]), return_weights=False)
        pf.resample()
        pf.update_gps(np.array([[truth[0], truth[1]]] * F))
        pf.resample(wait=False)
        pf.resample_status()
        pf.mean_cov()
        pf.neff()
        pf.weights_q()
        # a cloud at the map border: the profile kernel lists particles, the per-ray kernel walks the list
        p = pf.poses()
        p[::3, 0] = 5.0
        pf.set_poses(p)
        pf.update_mbes(np.stack([r[0], r[0]]), return_weights=False)
        pf.resample()
        if resampler == "systematic" and not inplace:
            # the N_eff gate (held back and let through) and a replayed sequence of three cycles
            from smarc_navigation_b200.filter import _odom_row
            pf.set_neff_gate(0.5)
            pf.set_log_weights(np.zeros(N * F, np.float32))
            pf.resample(wait=False)
            assert pf.resample_skipped().all()
            rows = np.array([[[_odom_row(o, None)] for o in odoms]] * 3)
            pf.replay(rows, np.stack([np.stack([r[0], r[0]])] * 3))
            pf.neff_fixed()
        pf.close()
    print("sanitize cycle ok")


if __name__ == "__main__":
    main()
