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
        def run(reps, timed):
            evs = []
            for r in range(reps):
                pf.set_log_weights(logw)               # (max pass of set_logw is booked under "normalise": reported apart)
                pf._lib.auvpf_shard_max_dev(pf._h)     # ... and made here, outside the timed resample
                with torch.cuda.stream(stream):
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                pf.resample(wait=False)
                e1.record(stream)
                evs.append((e0, e1))
            pf.sync()
            return [x.elapsed_time(y) for x, y in evs]
        run(3, False)
        pair = run(a.reps, True)                       # the two kernels as enqueued (programmatic dependent launch active)
        pf.profile(True)
        pf.profile_read(reset=True)
        run(a.reps, False)
        prof = pf.profile_read(reset=True)
        pf.profile(False)
        ms = {k: v[0] / a.reps for k, v in prof.items() if v[1]}
        launches = {k: v[1] / a.reps for k, v in prof.items() if v[1]}
        rs = float(np.median(pair))
        print(json.dumps({"particles": n, "resample_ms": rs, "achieved_gbs": 96.0 * n / (rs * 1e-3) / 1e9, "peak_gbs": peak,
                          "frac": 96.0 * n / (rs * 1e-3) / 1e9 / peak, "frac_80B": 80.0 * n / (rs * 1e-3) / 1e9 / peak,
                          "per_kernel_ms_bracketed": ms, "launches_per_resample": launches,
                          "note": "resample_ms: CUDA events around k_weights + k_select as enqueued, median; "
                                  "per_kernel_ms_bracketed: each kernel between its own events (no overlap)"}))
        pf.close()


if __name__ == "__main__":
    main()
