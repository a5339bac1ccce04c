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
        pf.estimate(weighted=True, circular_yaw=True, full_cov=True)
        pf.pose_array(max_poses=100)
        pf.pose_array_topk(64)
        pf.kld_bins()
        pf.close()
    # linear weights (resample(weights)) and a particle-sharded filter: 3 ranks on this device, peer-memory path
    import ctypes as C
    from smarc_navigation_b200 import _lib
    L = _lib.load()
    pf = ParticleFilter(particle_count=N, resampling_noise_covariance=P["resampling_noise_covariance"], seed=3)
    pf.resample(weights=np.random.default_rng(0).uniform(0.0, 1.0, N))
    pf.weights_q()
    pf.close()
    G, M = 3, 2500
    pfs = [ParticleFilter(particle_count=M, global_offset=g * M, global_count=G * M, seed=5,
                          resampling_noise_covariance=P["resampling_noise_covariance"]) for g in range(G)]
    handles = (C.c_void_p * G)(*[p_._h for p_ in pfs])
    rng = np.random.default_rng(1)
    for g, p_ in enumerate(pfs):
        _lib.check(L.auvpf_peer_attach(p_._h, handles, G, g))
        p_.set_log_weights(rng.normal(0, 3, M).astype(np.float32))
    rv = (C.c_double * 6)(*P["resampling_noise_covariance"])
    for c in range(2):
        _lib.check(L.auvpf_group_resample(handles, G, 0.3 + 0.2 * c, rv, None))
        for g, p_ in enumerate(pfs):
            p_.sync()
            p_.set_log_weights(rng.normal(0, 3, M).astype(np.float32))
    for p_ in pfs:
        p_.close()
    print("sanitize cycle ok")


if __name__ == "__main__":
    main()
