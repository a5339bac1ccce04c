#!/usr/bin/env python
"""Resample path alone (k_weights + k_select with the roughening fused)
over particle counts: per-kernel CUDA-event times from the library's own profile counters and the
achieved fraction of the HBM roofline (96 algorithmic bytes per particle).
One JSON line per size.  L2 is flushed before every resample; the log-weights are rewritten by
auvpf_set_logw before it (outside the timed kernels), so each resample starts from HBM.

    python tools/bench_resample.py [--sizes 1048576,2097152,...] [--reps 10]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="65536,1048576,2097152,4194304,8388608,16777216")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--scale", type=float, default=3.0, help="standard deviation of the log-weights")
    a = ap.parse_args()
    import torch
    from smarc_navigation_b200 import ParticleFilter
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for n in [int(x) for x in a.sizes.split(",")]:
        rng = np.random.default_rng(n)
        pf = ParticleFilter(particle_count=n, init_covariance=[4, 4, 0, 0, 0, 0.01],
                            resampling_noise_covariance=[0.01, 0.01, 0, 0, 0, 1e-5], seed=1)
        logw = (rng.normal(0, a.scale, n) - 100.0).astype(np.float32)
        stream = torch.cuda.ExternalStream(pf.stream_ptr)
        for r in range(3 + a.reps):
            if r == 3:
                pf.profile(True)
                pf.profile_read(reset=True)
            pf.set_log_weights(logw)                   # (max pass of set_logw is booked under "normalise": reported apart)
            with torch.cuda.stream(stream):
                flush.zero_()
            pf.resample(wait=False)
        pf.sync()
        prof = pf.profile_read(reset=True)
        ms = {k: v[0] / a.reps for k, v in prof.items() if v[1]}
        launches = {k: v[1] / a.reps for k, v in prof.items() if v[1]}
        rs = ms.get("normalise", 0) + ms.get("scan", 0) + ms.get("select_gather", 0)
        print(json.dumps({"particles": n, "ms": ms, "launches_per_resample": launches, "resample_ms": rs,
                          "achieved_gbs": 96.0 * n / (rs * 1e-3) / 1e9, "peak_gbs": peak,
                          "frac": 96.0 * n / (rs * 1e-3) / 1e9 / peak,
                          "note": "normalise includes the k_weight_max_f32 pass that set_logw makes necessary "
                                  "(the update kernels publish the max themselves in the filter cycle)"}))
        pf.close()


if __name__ == "__main__":
    main()
