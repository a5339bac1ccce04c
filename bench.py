#!/usr/bin/env python
"""Benchmark of the auv_particle_filter measurement cycle (BASELINE.json metric:
particle-beam likelihood evaluations/s and filter-cycle latency).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle port)

A "step" is one filter cycle on one ping: P=5 predicts (50 Hz odometry between
10 Hz pings) + MBES update (raycast + beam likelihood) + normalise + systematic
resample + gather with fused roughening.  Workload at 1 GPU = BASELINE config C3
(1,048,576 particles x 512 beams, 0.5 m DEM over 4 km^2); with N GPUs every rank
holds that many particles of ONE filter (weak scaling; `--workload C4` gives the
16M x 512 target configuration at 8 GPUs).

Prints ONE JSON line (rank 0).  `value` = whole-job evals/s with the pings already
resident in HBM, timed with CUDA events on the library's stream per step (L2 flushed
between steps), max over ranks; `e2e` = the same through the public Python API with
pinned HOST buffers (ping H2D, estimate D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from smarc_navigation_b200 import world  # noqa: E402

METRIC = "particle_beam_evals_per_s"
UNIT = "evals/s"
P_PREDICTS = 5
# dram__bytes_read.sum + dram__bytes_write.sum of k_update_mbes_profile per launch, from the committed ncu --set full
# capture (profiles/r1_ncu_update_mbes_profile.txt); only the workload that was captured has a figure
NCU_DRAM_TRAFFIC = {"C3": 33.64e6 + 0.87e6}
# what bounds that kernel, from the same capture (SURVEY 8d asks for these instead of a %-of-HBM figure)
NCU_DOMINANT = {"C3": {"issue_slots_used": 0.72, "ipc_per_sm": 2.85, "active_lanes_of_32": 23.5, "l1_hit_rate": 0.988,
                       "l2_hit_rate": 0.76, "registers": 96, "warps_per_sm": 18, "top_stall": "wait (dependent issue)",
                       "source": "profiles/r1_ncu_update_mbes_profile.txt"}}
RS_BYTES = 96        # algorithmic HBM bytes per particle of normalise + scan + select + gather


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "C2", "C3", "C4", "C5"])
    ap.add_argument("--cpu-seconds", type=float, default=None,
                    help="target CPU work per reference step / baseline sample (default: 15 s for the baseline "
                         "sample; for --impl reference about 75 s split over the steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(args, n_gpus):
    name = args.workload
    if name == "auto":
        name = "C3"
    c = dict(world.CONFIGS[name])
    if name == "C4":
        per_gpu = c["particles"] // 8
    else:
        per_gpu = c["particles"]
    # C5: 256 independent filters over 8 GPUs = 32 whole filters per GPU, no cross-GPU traffic
    c["filters_per_gpu"] = c.get("filters", 8) // 8 if name == "C5" else 1
    c["name"] = name
    c["per_gpu"] = per_gpu
    c["total"] = per_gpu * n_gpus
    return c


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.idx = gpu_index
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "25"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_cycle_setup(cfg, seed_particles=4):
    from oracle import c_oracle
    dem = world.make_dem(cfg["dem_n"], cfg["res"])
    beams = world.make_beams(cfg["beams"])
    ext = cfg["dem_n"] * cfg["res"]
    traj = world.Trajectory(ext)
    return c_oracle, dem, beams, traj


def cpu_run(cfg, n_sample, steps, warmup):
    """The oracle port of the cycle (C/OpenMP float64 raycast + likelihood + motion model,
    NumPy/C normalise + systematic resample + gather + roughen) on n_sample particles."""
    from oracle import pf_oracle as o
    c_oracle, dem, beams, traj = cpu_cycle_setup(cfg)
    P = world.MBES_PARAMS
    B = cfg["beams"]
    rng = np.random.default_rng(4)
    poses = o.add_noise(np.tile(traj.truth, (n_sample, 1)), P["init_covariance"], rng.normal(size=(n_sample, 6)))
    times = []
    for s in range(warmup + steps):
        odoms, truth = traj.cycle()
        r_true = c_oracle.update_mbes(dem, truth[None, :], np.eye(4), beams, None, 1.0, 0.5, world.RANGE_MAX,
                                      want_ranges=True)["ranges"][0]
        meas = r_true + world.ping_noise(B, s)
        pn = [rng.normal(size=(n_sample, 6)) for _ in odoms]
        rn = rng.normal(size=(n_sample, 6))
        u = float(rng.uniform())
        t0 = time.perf_counter()
        for od, nz in zip(odoms, pn):
            roll, pitch = o.odom_roll_pitch(od.quat)
            poses = c_oracle.motion_pred(poses, od.v, od.wz, roll, pitch, od.z, od.dt, P["motion_covariance"], nz)
        logw = c_oracle.update_mbes(dem, poses, np.eye(4), beams, meas, P["measurement_std"], world.RANGE_MIN,
                                    world.RANGE_MAX)["logw"].astype(np.float32)
        qq, m, A = c_oracle.quantise_logw(logw)
        anc = c_oracle.systematic(o.weights_from_q(qq), u)
        poses = o.add_noise(poses[anc], P["resampling_noise_covariance"], rn)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    return float(np.mean(times)), c_oracle.num_threads()


def cpu_sample_size(cfg, seconds):
    """Size the bounded sample so that one CPU cycle is about `seconds` of work."""
    t, cores = cpu_run(cfg, 512, 1, 1)          # one untimed warm-up cycle (thread pool, page faults)
    per_particle = t / 512.0
    n = int(max(256, min(cfg["per_gpu"], seconds / max(per_particle, 1e-9))))
    return n, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload(args, args.gpus)
    # bounded sample: the whole --steps K --warmup W run should end within a couple of minutes
    per_step = args.cpu_seconds if args.cpu_seconds else min(15.0, max(0.5, 75.0 / (args.steps + min(args.warmup, 1))))
    n, cores = cpu_sample_size(cfg, per_step)
    t, cores = cpu_run(cfg, n, args.steps, min(args.warmup, 1))
    v = n * cfg["beams"] / t
    sample = "%d of %d particles x %d beams per step, C/OpenMP float64 oracle port, %d threads" % (
        n, cfg["total"], cfg["beams"], cores)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t, "ms_per_step_full_workload_extrapolated": 1e3 * t * cfg["total"] / n,
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %d particles x %d beams, %.1f m DEM %dx%d" % (
                cfg["name"], cfg["total"], cfg["beams"], cfg["res"], cfg["dem_n"], cfg["dem_n"]),
                "predicts_per_cycle": P_PREDICTS},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    G = world_size
    cfg = workload(args, G)
    M, B, F = cfg["per_gpu"], cfg["beams"], cfg["filters_per_gpu"]
    K, W = args.steps, max(args.warmup, 3)
    P = world.MBES_PARAMS
    ext = cfg["dem_n"] * cfg["res"]
    dem = world.make_dem(cfg["dem_n"], cfg["res"])
    beams = world.make_beams(B)
    # one trajectory per filter (C5: offset starts, own odometry noise seed)
    trajs = [world.Trajectory(ext, seed=2 + (rank * F + f if F > 1 else 0),
                              offset=(12.0 * ((rank * F + f) % 16), 12.0 * ((rank * F + f) // 16)) if F > 1 else (0.0, 0.0))
             for f in range(F)]
    traj = trajs[0]

    from smarc_navigation_b200 import ParticleFilter
    common = dict(init_covariance=P["init_covariance"], motion_covariance=P["motion_covariance"],
                  resampling_noise_covariance=P["resampling_noise_covariance"],
                  measurement_std=P["measurement_std"], range_min=world.RANGE_MIN, range_max=world.RANGE_MAX,
                  init_mean=traj.truth, seed=4, coarse_level=int(os.environ.get("AUVPF_COARSE_LEVEL", "-1"), 0))
    if G == 1 or F > 1:
        pf = ParticleFilter(particle_count=M, n_filters=F, device=local_rank, **common)
        filt = pf
        if F > 1:      # every filter starts around its own vehicle
            rng0 = np.random.default_rng(4 + rank)
            p0 = np.concatenate([t.truth[None, :] + rng0.normal(size=(M, 6)) * np.sqrt(P["init_covariance"]) for t in trajs])
            pf.set_poses(p0)
    else:
        from smarc_navigation_b200.sharded import ShardedParticleFilter
        filt = ShardedParticleFilter(M, dist, device=local_rank, **common)
        pf = filt.pf
    filt.set_dem(dem)
    filt.set_beams(beams)

    # synthetic ping sequence: ground-truth ranges raycast on the GPU (the product path), + noise
    n_steps = 2 * (K + W) + 2 * K + 2
    cycles = []
    for s in range(n_steps):
        per = [t.cycle() for t in trajs]
        truth = np.stack([c[1] for c in per])                                    # (F, 6)
        r_true = pf.raycast(truth)                                               # (F, B)
        noise = np.stack([world.ping_noise(B, s, seed=3 + 7 * f) for f in range(F)]) if F > 1 else world.ping_noise(B, s)[None]
        if F == 1:
            odoms = [per[0][0][p] for p in range(P_PREDICTS)]
        else:       # batched filters: one (F, 11) array of ABI odometry rows per predict
            odoms = [np.array([[o.v[0], o.v[1], o.v[2], o.wz, o.quat[0], o.quat[1], o.quat[2], o.quat[3], o.z, o.dt, 0.0]
                               for o in (per[f][0][p] for f in range(F))]) for p in range(P_PREDICTS)]
        cycles.append((odoms, (r_true + noise).astype(np.float32).reshape(F * B), truth[0], truth))
    pings_dev = torch.from_numpy(np.stack([c[1] for c in cycles])).cuda()
    pings_pin = torch.from_numpy(np.stack([c[1] for c in cycles])).pin_memory()
    stream = torch.cuda.ExternalStream(pf.stream_ptr)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    import ctypes as C
    from smarc_navigation_b200 import _lib
    L = _lib.load()

    def cycle_device(s):
        filt.predict_many(cycles[s][0])            # the 5 odometry messages between two pings, one pass over the state
        _lib.check(L.auvpf_update_mbes(pf._h, C.c_void_p(pings_dev[s].data_ptr()), None))
        filt.resample(wait=False) if (G == 1 or F > 1) else filt.resample()     # enqueue only: no host wait in the cycle

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.zero_()

    # ---- device-resident timing --------------------------------------------------------------
    for s in range(W):
        cycle_device(s)
    pf.sync()
    pf.profile(True)
    pf.profile_read(reset=True)
    clocks = ClockSampler(local_rank)
    barrier()
    evs = []
    for s in range(W, W + K):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        cycle_device(s)
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(np.sum(step_ms))
    prof = pf.profile_read(reset=True)
    pf.profile(False)
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / K
    value = M * F * G * B / (ms_per_step * 1e-3)
    mean, cov = filt.mean_cov()
    mean = np.asarray(mean).reshape(-1, 6) if F > 1 else np.asarray(mean).reshape(1, -1)
    tr = cycles[W + K - 1][3]
    track_err = float(np.max(np.hypot(mean[:, 0] - tr[:len(mean), 0], mean[:, 1] - tr[:len(mean), 1])))

    # ---- end to end through the public API, host buffers ----------------------------------------
    base = W + K
    for s in range(base, base + W):
        filt.predict_many(cycles[s][0])
        filt.update_mbes(pings_pin[s].numpy()) if (G > 1 and F == 1) else pf.update_mbes(pings_pin[s].numpy(), return_weights=False)
        filt.resample(wait=False) if (G == 1 or F > 1) else filt.resample()
        filt.mean_cov()
    barrier()
    t0 = time.perf_counter()
    e2e_lat = []
    for s in range(base + W, base + W + K):
        t1 = time.perf_counter()
        filt.predict_many(cycles[s][0])
        if G > 1 and F == 1:
            filt.update_mbes(pings_pin[s].numpy())
        else:
            pf.update_mbes(pings_pin[s].numpy(), return_weights=False)
        filt.resample(wait=False) if (G == 1 or F > 1) else filt.resample()
        est = filt.mean_cov()                   # reads the estimate back: the cycle's results are on the host here
        e2e_lat.append(1e3 * (time.perf_counter() - t1))
    barrier()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop()                         # sampled over both timed regions (device-resident and end-to-end)
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = M * F * G * B * K / e2e_s

    # ---- the same K cycles as ONE replay call (auvpf_replay): the sequence goes up once from host
    # memory, K cycles are enqueued back to back, the K estimates come back once ------------------------
    replay = None
    if G == 1 or F > 1:
        from smarc_navigation_b200.filter import _odom_row
        def rows(s0, s1):
            od = np.array([[([_odom_row(o, None)] if F == 1 else o) for o in cycles[s][0]] for s in range(s0, s1)], np.float64)
            return od.reshape(s1 - s0, P_PREDICTS, F, 11), np.stack([cycles[s][1] for s in range(s0, s1)]).reshape(s1 - s0, F, B)
        b2 = base + W + K
        od_w, rg_w = rows(b2, b2 + K)           # warm-up of the same length: the replay buffers are sized once
        pf.replay(od_w, rg_w)
        od_k, rg_k = rows(b2 + K, b2 + 2 * K)
        barrier()
        t0 = time.perf_counter()
        rmean, rcov = pf.replay(od_k, rg_k)
        barrier()
        rp_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([rp_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rp_s = float(t.item())
        replay = {"value": M * F * G * B * K / rp_s, "unit": UNIT, "ms_per_step": 1e3 * rp_s / K,
                  "h2d_bytes_per_step": F * (B * 4 + 8 + P_PREDICTS * 40), "d2h_bytes_per_step": F * 16 * 8,
                  "note": "ParticleFilter.replay: one upload, K cycles enqueued back to back, one read-back"}
    h2d = F * (B * 4 + 8) + (P_PREDICTS * F * 40 if F > 1 else 0)   # ping ranges + resample uniforms (+ per-filter odometry rows)
    d2h = F * (16 * 8 + 4)                      # pose sums/moments + the degenerate-weights flag

    # ---- roofline of the dominant kernel + the HBM-bound resample path ----------------------------
    peak, peak_src = peaks()
    dem_bytes = dem["z"].size * 4
    mbes_ms, mbes_n = prof["update_mbes"]
    mbes_n = K                                               # per update (profile-merge kernel + per-ray kernel for flagged particles)
    alg_mbes = M * F * (32 + 4) + dem_bytes + 16 * B        # state read + logw write, DEM once, beam table
    ach = alg_mbes / (mbes_ms / max(mbes_n, 1) * 1e-3) / 1e9 if mbes_n else 0.0
    rs_ms = prof["normalise"][0] + prof["scan"][0] + prof["select_gather"][0]
    rs_n = K                                                # one resample per cycle
    alg_rs = M * F * RS_BYTES                                   # normalise 4 + scan 12 + select 12 + gather 4+32+32
    ach_rs = alg_rs / (rs_ms / rs_n * 1e-3) / 1e9 if rs_ms > 0 else 0.0
    prop_ms, prop_n = prof["propagate"]
    # one launch runs the P predicts of a cycle in registers: algorithmic traffic = one read (x, y, yaw) + one write
    ach_prop = M * F * 52 / (prop_ms / max(prop_n, 1) * 1e-3) / 1e9 if prop_n else 0.0
    launches = int(sum(v[1] for v in prof.values()))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": G, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %s%d particles/GPU x %d beams (%d total), %.1f m DEM %dx%d, %d predicts per cycle" % (
            cfg["name"], ("%d filters x " % F) if F > 1 else "", M, B, M * F * G, cfg["res"], cfg["dem_n"], cfg["dem_n"],
            P_PREDICTS),
            "parallelism": ("filter-sharded x%d (%d filters/GPU, no communication)" % (G, F)) if F > 1 else
            "particle-sharded x%d%s" % (G, (", exchange=%s" % filt.exchange) if G > 1 else ""), "l2": "flushed between steps (256 MiB memset)",
            "resampler": "systematic", "cycle_hz": 1e3 / ms_per_step},
        "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / K,
                # filter-cycle latency, host call to estimate on the host (BASELINE metric, second half)
                "latency_ms": {"median": float(np.median(e2e_lat)), "p99": float(np.percentile(e2e_lat, 99)),
                               "max": float(np.max(e2e_lat))}},
        "cycle_latency_ms": {"median": float(np.median(step_ms)), "p99": float(np.percentile(step_ms, 99)),
                             "max": float(np.max(step_ms)), "of": "device time per cycle (CUDA events), this rank"},
        "e2e_replay": replay,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "k_update_mbes_profile (+ k_update_mbes for flagged particles): fused raycast + beam likelihood",
                     "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": NCU_DRAM_TRAFFIC.get(cfg["name"]) if G == 1 else None,
                     "ncu": NCU_DOMINANT.get(cfg["name"]),
                     "peak_source": peak_src, "ms_per_launch": mbes_ms / max(mbes_n, 1),
                     "evals_per_s": M * F * B / (mbes_ms / max(mbes_n, 1) * 1e-3) if mbes_n else None,
                     "note": "dominant kernel is instruction-issue-bound (ncu: 72 % of issue slots, 2.85 IPC/SM, L1 hit "
                             "99 %) on an L1/L2-resident DEM footprint, not HBM-bound (SURVEY 8d): its algorithmic HBM "
                             "bytes are 0.2 % of what the measured time could move; the HBM-bound path is roofline_resample"},
        "roofline_resample": {"bound": "hbm", "kernel": "k_normalise + k_quotients + k_select_gather",
                              "achieved": ach_rs, "peak": peak, "unit": "GB/s", "frac": ach_rs / peak,
                              "ms_per_cycle": rs_ms / rs_n, "bytes_per_particle": RS_BYTES},
        "roofline_propagate": {"bound": "hbm", "achieved": ach_prop, "peak": peak, "unit": "GB/s",
                               "frac": ach_prop / peak, "ms_per_launch": prop_ms / max(prop_n, 1)},
        "kernel_ms_per_cycle": {k: v[0] / K for k, v in prof.items() if v[1]},
        "track_error_m": track_err,
    }
    if rank == 0 and G == 1 and not args.no_cpu_baseline:
        n, cores = cpu_sample_size(cfg, args.cpu_seconds or 15.0)
        t, cores = cpu_run(cfg, n, 1, 0)
        line["cpu_baseline"] = {"value": n * B / t, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d of %d particles x %d beams, 1 cycle, C/OpenMP float64 oracle port" % (
                                    n, M, B)}
    if rank == 0:
        print(json.dumps(line))
    filt.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
