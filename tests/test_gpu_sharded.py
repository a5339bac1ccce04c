"""Particle-sharded filter: the cross-GPU normalise + systematic resample + exchange must give the same
ancestors, bit for bit, as one GPU holding every particle and as the reference's systematic_resample on
the same quantised weights.

Three layers, so that a box with ONE GPU still runs the peer-memory kernels:
* `test_peer_memory_resample_on_one_gpu`: G handles on one device attached to each other in-process
  (auvpf_peer_attach) -- the mailbox exchange inside k_weights / k_select<peers> and the peer reads, driven
  from one thread with every rank's launch enqueued before any is waited for;
* `test_shard_stages_on_one_gpu`: the staged (NCCL-variant) entry points driven by hand;
* `test_sharded_resample_on_gpus`: the real thing, one process per GPU over NCCL + CUDA IPC (skipped
  below 2 GPUs; tools/gpu_multi.sh runs it on 2 / 4 / 8)."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene(N, seed, spread=5.0, offset=-200.0):
    rng = np.random.default_rng(seed)
    logw = (rng.normal(0, spread, N) + offset).astype(np.float32)
    if seed % 2:
        logw[: N // 2 + 3] -= 80.0                      # almost all mass on the upper shards
    logw[7 % N] = -np.inf
    poses = np.zeros((N, 6))
    poses[:, 0] = np.arange(N)
    poses[:, 1] = -0.25 * np.arange(N)
    poses[:, 5] = rng.uniform(-3, 3, N).astype(np.float32)
    return logw, poses, float(rng.uniform())


def _worker(rank, world, port, M, seed, tmp, exchange):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from smarc_navigation_b200.sharded import ShardedParticleFilter
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    N = M * world
    logw, poses, u = _scene(N, seed)
    if exchange == "p2p":
        os.environ["AUVPF_ESTIMATE_MAILBOX"] = "1"          # the estimate through the mailbox as well (opt-in, see sharded.mean_cov)
    sp = ShardedParticleFilter(M, dist, device=rank, seed=seed, resampling_noise_covariance=[0.0] * 6,
                               exchange=exchange)
    assert sp.exchange == exchange, "peer memory unavailable: %s" % sp.exchange
    sp.pf.set_poses(poses[rank * M:(rank + 1) * M])
    sp.pf.set_log_weights(logw[rank * M:(rank + 1) * M])
    anc_t = torch.full((M,), -1, dtype=torch.int64, device="cuda") if exchange == "p2p" else None
    sp.resample(u=u, roughen=False, check=True, ancestors=anc_t)
    got = sp.poses()
    qq = sp.pf.weights_q()
    mean, cov = sp.mean_cov()
    # a second cycle on the new state: the double buffers and the peers' views flip together
    logw2 = (np.random.default_rng(seed + 100).normal(0, 2.0, N) - 50).astype(np.float32)
    sp.pf.set_log_weights(logw2[rank * M:(rank + 1) * M])
    u2 = 0.618
    sp.resample(u=u2, roughen=False)
    got2 = sp.poses()
    # several more cycles back to back without any host synchronisation in between (mailbox epochs, buffer flips)
    for c in range(4):
        lw = (np.random.default_rng(seed + 200 + c).normal(0, 3.0, N) - 20).astype(np.float32)
        sp.pf.set_log_weights(lw[rank * M:(rank + 1) * M])
        sp.resample(u=0.1 + 0.2 * c, roughen=False)
    got_chain = sp.poses()
    # the estimate again, three times in a row (mailbox path: both parities of its slots and the reuse of the first)
    est_chain = [sp.mean_cov() for _ in range(3)]
    # (equal up to the order of the floating-point atomics inside k_pose_stats, which differs from launch to launch;
    # the scene's x = particle index makes the covariance sums ~1e18 at 1 M particles, hence RELATIVE)
    assert all(np.allclose(e[0], est_chain[0][0], rtol=1e-9, atol=1e-9) and np.allclose(e[1], est_chain[0][1], rtol=1e-9, atol=1e-6)
               for e in est_chain)
    # degenerate weights on every shard: state carried over, flag raised
    sp.pf.set_log_weights(np.full(M, -np.inf, np.float32))
    deg = False
    try:
        sp.resample(u=0.5, roughen=False, check=True)
    except FloatingPointError:
        deg = True
    got3 = sp.poses()
    # the N_eff gate on the global sums
    gated = None
    if exchange == "p2p":
        sp.pf.set_neff_gate(0.5)
        sp.pf.set_log_weights(np.zeros(M, np.float32))            # uniform weights: N_eff = N, held back
        sp.resample(u=0.3, roughen=False)
        gated = bool(sp.resample_skipped()) and np.array_equal(sp.poses(), got3)
        sp.pf.set_neff_gate(0.0)
    np.savez(os.path.join(tmp, "r%d.npz" % rank), poses=got, poses2=got2, poses3=got3, chain=got_chain, deg=deg, u=u,
             logw=logw, logw2=logw2, u2=u2, all_poses=poses, mean=mean, cov=cov, qq=qq, mean_chain=est_chain[0][0],
             anc=anc_t.cpu().numpy() if anc_t is not None else np.zeros(0), gated=-1 if gated is None else int(gated))
    sp.close()
    dist.barrier()
    dist.destroy_process_group()


def _check_parts(parts, world, M, seed, exchange):
    from oracle import c_oracle, pf_oracle as o
    logw, u, poses = parts[0]["logw"], float(parts[0]["u"]), parts[0]["all_poses"]
    N = M * world
    qq = c_oracle.quantise_logw(logw)[0]
    anc = c_oracle.systematic(o.weights_from_q(qq), u)
    got = np.concatenate([p["poses"] for p in parts])
    np.testing.assert_array_equal(got[:, 0], poses[anc, 0])                     # bit-exact ancestors for any G
    np.testing.assert_array_equal(got[:, 1], poses[anc, 1])
    np.testing.assert_array_equal(got[:, 5], poses[anc, 5])
    np.testing.assert_array_equal(np.concatenate([p["qq"] for p in parts]), qq)  # every shard's slice of the GLOBAL weights
    if exchange == "p2p":
        np.testing.assert_array_equal(np.concatenate([p["anc"] for p in parts]), anc.astype(np.int64))
    em, ec, _ = o.estimate(got)
    np.testing.assert_allclose(parts[0]["mean"], em, atol=1e-6)
    np.testing.assert_allclose(parts[0]["cov"], ec, rtol=1e-6, atol=1e-6)
    for p in parts[1:]:
        np.testing.assert_array_equal(p["mean"], parts[0]["mean"])              # every rank returns the same estimate
        np.testing.assert_array_equal(p["mean_chain"], parts[0]["mean_chain"])
    np.testing.assert_allclose(parts[0]["mean_chain"], o.estimate(np.concatenate([p["chain"] for p in parts]))[0], atol=1e-6)
    # second cycle
    logw2, u2 = parts[0]["logw2"], float(parts[0]["u2"])
    anc2 = c_oracle.systematic(o.weights_from_q(c_oracle.quantise_logw(logw2)[0]), u2)
    got2 = np.concatenate([p["poses2"] for p in parts])
    np.testing.assert_array_equal(got2[:, 0], got[anc2, 0])
    np.testing.assert_array_equal(got2[:, 5], got[anc2, 5])
    # the chain of unsynchronised cycles
    state = got2
    for c in range(4):
        lw = (np.random.default_rng(seed + 200 + c).normal(0, 3.0, N) - 20).astype(np.float32)
        a = c_oracle.systematic(o.weights_from_q(c_oracle.quantise_logw(lw)[0]), 0.1 + 0.2 * c)
        state = state[a]
    np.testing.assert_array_equal(np.concatenate([p["chain"] for p in parts])[:, 0], state[:, 0])
    # degenerate cycle: flagged on every rank, state untouched
    assert all(bool(p["deg"]) for p in parts)
    np.testing.assert_array_equal(np.concatenate([p["poses3"] for p in parts]), np.concatenate([p["chain"] for p in parts]))
    if exchange == "p2p":
        assert all(int(p["gated"]) == 1 for p in parts)
    return anc


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("M,seed", [(4096, 0), (100000, 1), (1 << 20, 2)])
def test_sharded_resample_on_gpus(tmp_path, M, seed, exchange):
    world = int(os.environ.get("AUVPF_TEST_WORLD", "0")) or min(_n_gpus(), 4)
    if world < 2 or _n_gpus() < world:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    from smarc_navigation_b200 import ParticleFilter
    mp.spawn(_worker, args=(world, _free_port(), M, seed, str(tmp_path), exchange), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    anc = _check_parts(parts, world, M, seed, exchange)
    # the same filter on ONE GPU
    logw, u, poses = parts[0]["logw"], float(parts[0]["u"]), parts[0]["all_poses"]
    pf = ParticleFilter(particle_count=M * world, resampling_noise_covariance=[0.0] * 6, seed=seed)
    pf.set_poses(poses)
    pf.set_log_weights(logw)
    np.testing.assert_array_equal(pf.resample(u=u, return_ancestors=True, roughen=False), anc)
    pf.close()


@pytest.mark.gpu
@pytest.mark.parametrize("G,M,seed", [(2, 5000, 0), (3, 70001, 1), (4, 1 << 18, 2), (8, 2500, 3), (8, 1 << 17, 5)])
def test_peer_memory_resample_on_one_gpu(G, M, seed):
    """The peer-memory path -- mailbox exchange of the maximum and of the totals inside k_weights / k_select<peers>,
    prefix sums rebuilt from the OWNERS' log-weights, state gathered from the owners' buffers -- with G handles
    on one device attached in-process and driven by auvpf_group_resample (stage-interleaved enqueue: the kernels of
    the ranks wait on each other's mailbox posts, in whatever order the device runs the streams).  Ancestors, the
    weights, three chained cycles, the roughening keyed by the GLOBAL slot, the gate and the degenerate case
    against the oracle and against one handle holding every particle."""
    import torch
    from oracle import c_oracle, pf_oracle as o
    from smarc_navigation_b200 import ParticleFilter, _lib
    L = _lib.load()
    N = G * M
    logw, poses, u = _scene(N, seed, spread=4.0, offset=-150.0)
    res = [0.01, 0.01, 0.0, 0.0, 0.0, 1e-5]
    pfs = [ParticleFilter(particle_count=M, global_offset=g * M, global_count=N, seed=seed,
                          resampling_noise_covariance=res) for g in range(G)]
    handles = (C.c_void_p * G)(*[pf._h for pf in pfs])
    for g, pf in enumerate(pfs):
        _lib.check(L.auvpf_peer_attach(pf._h, handles, G, g))
        pf.set_poses(poses[g * M:(g + 1) * M])
        pf.set_log_weights(logw[g * M:(g + 1) * M])
    ancs = [torch.full((M,), -1, dtype=torch.int64, device="cuda") for _ in range(G)]

    def everyone(uu, rough=None, anc=None):
        # all ranks live in this process and on one device: auvpf_group_resample enqueues the stages rank by rank
        # (a lone auvpf_shard_resample would wait inside its kernel for posts that may be queued behind it)
        rv = None if rough is None else (C.c_double * 6)(*rough)
        ap = None if anc is None else (C.c_void_p * G)(*[a.data_ptr() for a in anc])
        _lib.check(L.auvpf_group_resample(handles, G, uu, rv, ap))
        for pf in pfs:
            pf.sync()

    everyone(u, anc=ancs)
    qq = c_oracle.quantise_logw(logw)[0]
    anc = c_oracle.systematic(o.weights_from_q(qq), u)
    np.testing.assert_array_equal(np.concatenate([a.cpu().numpy() for a in ancs]), anc.astype(np.int64))
    got = np.concatenate([pf.poses() for pf in pfs])
    np.testing.assert_array_equal(got[:, 0], poses[anc, 0])
    np.testing.assert_array_equal(got[:, 5], poses[anc, 5])
    np.testing.assert_array_equal(np.concatenate([pf.weights_q() for pf in pfs]), qq)
    one = ParticleFilter(particle_count=N, resampling_noise_covariance=res, seed=seed)
    one.set_poses(poses)
    one.set_log_weights(logw)
    np.testing.assert_array_equal(one.resample(u=u, return_ancestors=True, roughen=False), anc)
    # chained cycles with roughening: bit-identical to the single handle (noise keyed by the global slot)
    for c in range(3):
        lw = (np.random.default_rng(seed + 50 + c).normal(0, 2.5, N) - 30).astype(np.float32)
        for g, pf in enumerate(pfs):
            pf.set_log_weights(lw[g * M:(g + 1) * M])
        one.set_log_weights(lw)
        uu = 0.05 + 0.3 * c
        everyone(uu, rough=res)
        one.resample(u=uu)
        np.testing.assert_array_equal(np.concatenate([pf.poses() for pf in pfs]), one.poses())
        # the estimate through the mailbox (auvpf_group_pose_sums: every rank's 16 pose sums to every rank): the same
        # table on all ranks, and combined it is update_loc_pose on the whole cloud; three calls in a row exercise
        # both parities of its slots and the reuse of the first
        from smarc_navigation_b200.sharded import combine_pose_sums
        for rep in range(3 if c == 0 else 1):
            tab = np.empty((G, G, 16), np.float64)
            _lib.check(L.auvpf_group_pose_sums(handles, G, tab.ctypes.data_as(C.c_void_p)))
            for g in range(1, G):
                np.testing.assert_array_equal(tab[g], tab[0])
            mean, cov = combine_pose_sums(tab[0], M)
            em, ec, _ = o.estimate(one.poses())
            np.testing.assert_allclose(mean, em, atol=1e-9)
            np.testing.assert_allclose(cov, ec, atol=1e-9)
    # the gate on the global sums: uniform weights are held back at ratio 0.5, concentrated ones are not
    state = one.poses()
    for pf in pfs + [one]:
        pf.set_neff_gate(0.5)
    for g, pf in enumerate(pfs):
        pf.set_log_weights(np.zeros(M, np.float32))
    one.set_log_weights(np.zeros(N, np.float32))
    everyone(0.4, rough=res)
    one.resample(u=0.4)             # (a held-back resample still advances the noise counters, on every path alike)
    assert all(bool(pf.resample_skipped()[0]) for pf in pfs) and bool(one.resample_skipped()[0])
    np.testing.assert_array_equal(np.concatenate([pf.poses() for pf in pfs]), state)
    np.testing.assert_array_equal(one.poses(), state)
    lw = (np.random.default_rng(seed + 99).normal(0, 6.0, N) - 30).astype(np.float32)
    for g, pf in enumerate(pfs):
        pf.set_log_weights(lw[g * M:(g + 1) * M])
    one.set_log_weights(lw)
    everyone(0.4, rough=res)
    one.resample(u=0.4)
    assert not any(bool(pf.resample_skipped()[0]) for pf in pfs) and not bool(one.resample_skipped()[0])
    np.testing.assert_array_equal(np.concatenate([pf.poses() for pf in pfs]), one.poses())
    # no finite weight anywhere: carried over, flagged
    state = one.poses()
    for pf in pfs:
        pf.set_log_weights(np.full(M, -np.inf, np.float32))
    everyone(0.5)
    flag = C.c_int32(0)
    for pf in pfs:
        _lib.check(L.auvpf_shard_degenerate(pf._h, C.byref(flag)))
        assert flag.value == 1
    np.testing.assert_array_equal(np.concatenate([pf.poses() for pf in pfs]), state)
    for pf in pfs + [one]:
        pf.close()


@pytest.mark.gpu
def test_peer_memory_last_particle_edge():
    """u >= 1 - N 2^-54: positions[-1] rounds to 1.0 = cumsum[-1] and the reference raises IndexError
    (resampling.py:162-167); the single-GPU and the sharded select agree on the LAST particle for that
    output, also when the last particles carry no weight."""
    import torch
    from oracle import pf_oracle as o
    from smarc_navigation_b200 import ParticleFilter, _lib
    L = _lib.load()
    u = float(np.nextafter(1.0, 0.0))
    for G, M in ((1, 1000), (1, 1024), (1, 4097), (2, 1000), (4, 2048)):
        N = G * M
        rng = np.random.default_rng(N)
        logw = rng.normal(0, 1.0, N).astype(np.float32)
        logw[-5:] = -np.inf                                   # the last particle WITH offspring is not the last particle
        poses = np.zeros((N, 6))
        poses[:, 0] = np.arange(N)
        qq = o.quantise_logw(logw)[0]
        pos = (u + np.arange(N)) / float(N)
        want = np.minimum(np.searchsorted(np.cumsum(o.weights_from_q(qq)), pos, side="right"), N - 1)
        assert want[-1] == N - 1 and want[-2] < N - 5
        if G == 1:
            pf = ParticleFilter(particle_count=N)
            pf.set_poses(poses)
            pf.set_log_weights(logw)
            np.testing.assert_array_equal(pf.resample(u=u, return_ancestors=True, roughen=False), want)
            pf.close()
            continue
        pfs = [ParticleFilter(particle_count=M, global_offset=g * M, global_count=N) for g in range(G)]
        handles = (C.c_void_p * G)(*[pf._h for pf in pfs])
        ancs = [torch.empty(M, dtype=torch.int64, device="cuda") for _ in range(G)]
        for g, pf in enumerate(pfs):
            _lib.check(L.auvpf_peer_attach(pf._h, handles, G, g))
            pf.set_poses(poses[g * M:(g + 1) * M])
            pf.set_log_weights(logw[g * M:(g + 1) * M])
        ap = (C.c_void_p * G)(*[a.data_ptr() for a in ancs])
        _lib.check(L.auvpf_group_resample(handles, G, u, None, ap))
        for pf in pfs:
            pf.sync()
        np.testing.assert_array_equal(np.concatenate([a.cpu().numpy() for a in ancs]), want)
        for pf in pfs:
            pf.close()


@pytest.mark.gpu
@pytest.mark.parametrize("G,M,seed", [(2, 5000, 0), (3, 70001, 1), (4, 1 << 18, 2)])
def test_shard_stages_on_one_gpu(G, M, seed):
    """The staged entry points (max / integer totals / emit<records> / load: what the NCCL-exchange variant
    runs, with non-trivial shard offsets and output ranges), host-staged AND device-resident, driven by hand
    for G handles that share ONE GPU: ancestors bit-exact against the oracle."""
    import torch
    from oracle import c_oracle, pf_oracle as o
    from smarc_navigation_b200 import ParticleFilter, _lib
    from smarc_navigation_b200.sharded import GpuShardOps, plan_exchange, splits_for, totals_to_ints, TOTAL_WORDS
    L = _lib.load()
    N = G * M
    logw, poses, u = _scene(N, seed, spread=4.0, offset=-150.0)
    anc = c_oracle.systematic(o.weights_from_q(c_oracle.quantise_logw(logw)[0]), u)
    for variant in ("host", "device"):
        pfs = [ParticleFilter(particle_count=M, global_offset=g * M, global_count=N, seed=seed,
                              resampling_noise_covariance=[0.0] * 6) for g in range(G)]
        ops = [GpuShardOps(pf) for pf in pfs]
        for g, pf in enumerate(pfs):
            pf.set_poses(poses[g * M:(g + 1) * M])
            pf.set_log_weights(logw[g * M:(g + 1) * M])
        if variant == "host":
            m = max(op.shard_max() for op in ops)
            table = [op.shard_sum(m) for op in ops]
        else:
            for pf in pfs:
                _lib.check(L.auvpf_shard_max_dev(pf._h))
                pf.sync()
            mx = torch.stack([op.t_wmax for op in ops]).max(0).values      # all-reduce MAX on the order-preserving int32
            for op in ops:
                op.t_wmax.copy_(mx)
            torch.cuda.synchronize()
            for pf in pfs:
                _lib.check(L.auvpf_shard_weights_dev(pf._h))
                pf.sync()
            table = [op.t_totals.tolist() for op in ops]                   # all-gather of the six integers
        rows = [totals_to_ints(r) for r in table]
        a_all = c_oracle.integer_weights(logw, np.float32(np.nanmax(logw)))
        for g in range(G):                                                  # every shard's integers are the oracle's
            ag = a_all[g * M:(g + 1) * M]
            assert rows[g][0] == int(ag.astype(object).sum())
            assert rows[g][1:4] == o.neff_sums(ag)
        c_off, bounds = plan_exchange([r[0] for r in rows], M, u)
        sends = [ops[g].shard_emit(u, table, bounds[g], bounds[g + 1], g) for g in range(G)]
        everything = torch.cat(sends)                        # offspring in global output order
        assert everything.shape[0] == N
        for g in range(G):
            send_counts, recv_counts = splits_for(bounds, M, g)
            assert sum(recv_counts) == M and sum(send_counts) == sends[g].shape[0]
            ops[g].shard_load(everything[g * M:(g + 1) * M].contiguous())
        got = np.concatenate([pf.poses() for pf in pfs])
        np.testing.assert_array_equal(got[:, 0], poses[anc, 0])
        np.testing.assert_array_equal(got[:, 1], poses[anc, 1])
        np.testing.assert_array_equal(got[:, 5], poses[anc, 5])
        for pf in pfs:
            pf.close()
