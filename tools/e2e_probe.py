"""Where the host-side time of one filter cycle through the public API goes (C3, one GPU): per-call enqueue cost and
the wait for the estimate.  gpurun -- python tools/e2e_probe.py"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import torch
from smarc_navigation_b200 import world


def main():
    ctx = bench.Ctx()
    ctx.rank = ctx.local_rank = 0
    ctx.G = 1
    ctx.dist = None
    ctx.dems = {}
    cfg = bench.workload("C3", 1)
    filt, pf, dem, trajs = bench.make_filter(ctx, cfg, "p2p")
    B = cfg["beams"]
    cycles = []
    for s in range(40):
        per = trajs[0].cycle()
        r = pf.raycast(per[1][None])
        cycles.append(([per[0][p] for p in range(bench.P_PREDICTS)], (r + world.ping_noise(B, s)[None]).astype(np.float32).reshape(B)))
    pin = torch.from_numpy(np.stack([c[1] for c in cycles])).pin_memory()
    for s in range(5):
        filt.predict_many(cycles[s][0]); pf.update_mbes(pin[s].numpy(), return_weights=False); filt.resample(wait=False); filt.mean_cov()
    acc = np.zeros(5)
    n = 0
    for s in range(5, 35):
        t = [time.perf_counter()]
        filt.predict_many(cycles[s][0]); t.append(time.perf_counter())
        pf.update_mbes(pin[s].numpy(), return_weights=False); t.append(time.perf_counter())
        filt.resample(wait=False); t.append(time.perf_counter())
        filt.mean_cov(); t.append(time.perf_counter())
        acc[:4] += np.diff(t); acc[4] += t[-1] - t[0]; n += 1
    print(json.dumps({k: round(1e3 * v / n, 4) for k, v in zip(("predict_many_ms", "update_mbes_ms", "resample_ms", "mean_cov_wait_ms", "cycle_ms"), acc)}))
    # the same with the host never waiting inside the loop (estimate read at the end): the device's pace
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for s in range(5, 35):
        filt.predict_many(cycles[s][0]); pf.update_mbes(pin[s].numpy(), return_weights=False); filt.resample(wait=False)
    filt.mean_cov(); print(json.dumps({"no_wait_cycle_ms": round(1e3 * (time.perf_counter() - t0) / 30, 4)}))
    filt.close()


main()
