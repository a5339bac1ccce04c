"""Particle-sharded filter on real GPUs (NCCL): the cross-GPU normalise + systematic resample +
exchange must give the same ancestors, bit for bit, as one GPU holding every particle and as the
reference's systematic_resample on the same quantised weights."""
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


def _worker(rank, world, port, M, seed, tmp, exchange):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from smarc_navigation_b200.sharded import ShardedParticleFilter
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(seed)
    N = M * world
    logw = (rng.normal(0, 5.0, N) - 200).astype(np.float32)
    if seed % 2:
        logw[: N // 2 + 3] -= 80.0                      # almost all mass on the upper shards
    poses = np.zeros((N, 6))
    poses[:, 0] = np.arange(N)
    poses[:, 1] = -0.25 * np.arange(N)
    poses[:, 5] = rng.uniform(-3, 3, N).astype(np.float32)
    u = float(rng.uniform())
    sp = ShardedParticleFilter(M, dist, device=rank, seed=seed, resampling_noise_covariance=[0.0] * 6,
                               exchange=exchange)
    sp.pf.set_poses(poses[rank * M:(rank + 1) * M])
    sp.pf.set_log_weights(logw[rank * M:(rank + 1) * M])
    sp.resample(u=u, roughen=False, check=True)
    got = sp.poses()
    mean, cov = sp.mean_cov()
    # a second cycle on the new state: the double buffers and the peers' views flip together
    logw2 = (np.random.default_rng(seed + 100).normal(0, 2.0, N) - 50).astype(np.float32)
    sp.pf.set_log_weights(logw2[rank * M:(rank + 1) * M])
    u2 = 0.618
    sp.resample(u=u2, roughen=False)
    got2 = sp.poses()
    # degenerate weights on every shard: state carried over, flag raised
    sp.pf.set_log_weights(np.full(M, -np.inf, np.float32))
    deg = False
    try:
        sp.resample(u=0.5, roughen=False, check=True)
    except FloatingPointError:
        deg = True
    got3 = sp.poses()
    np.savez(os.path.join(tmp, "r%d.npz" % rank), poses=got, poses2=got2, poses3=got3, deg=deg, u=u, logw=logw,
             logw2=logw2, u2=u2, all_poses=poses, mean=mean, cov=cov)
    sp.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("M,seed", [(4096, 0), (100000, 1), (1 << 20, 2)])
def test_sharded_resample_on_gpus(tmp_path, M, seed, exchange):
    world = min(_n_gpus(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import c_oracle, pf_oracle as o
    from smarc_navigation_b200 import ParticleFilter
    mp.spawn(_worker, args=(world, _free_port(), M, seed, str(tmp_path), exchange), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    logw, u, poses = parts[0]["logw"], float(parts[0]["u"]), parts[0]["all_poses"]
    N = M * world
    qq = c_oracle.quantise_logw(logw)[0]
    anc = c_oracle.systematic(o.weights_from_q(qq), u)
    got = np.concatenate([p["poses"] for p in parts])
    np.testing.assert_array_equal(got[:, 0], poses[anc, 0])                     # bit-exact ancestors for any G
    np.testing.assert_array_equal(got[:, 1], poses[anc, 1])
    np.testing.assert_array_equal(got[:, 5], poses[anc, 5])
    # the same filter on ONE GPU
    pf = ParticleFilter(particle_count=N, resampling_noise_covariance=[0.0] * 6, seed=seed)
    pf.set_poses(poses)
    pf.set_log_weights(logw)
    anc1 = pf.resample(u=u, return_ancestors=True, roughen=False)
    np.testing.assert_array_equal(anc1, anc)
    m1, c1 = pf.mean_cov()
    np.testing.assert_allclose(parts[0]["mean"][:3], m1[0, :3], rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(parts[0]["cov"], c1[0], rtol=1e-6, atol=1e-6)
    # second cycle
    logw2, u2 = parts[0]["logw2"], float(parts[0]["u2"])
    anc2 = c_oracle.systematic(o.weights_from_q(c_oracle.quantise_logw(logw2)[0]), u2)
    got2 = np.concatenate([p["poses2"] for p in parts])
    np.testing.assert_array_equal(got2[:, 0], got[anc2, 0])
    np.testing.assert_array_equal(got2[:, 5], got[anc2, 5])
    # degenerate cycle: flagged on every rank, state untouched
    assert all(bool(p["deg"]) for p in parts)
    np.testing.assert_array_equal(np.concatenate([p["poses3"] for p in parts]), got2)
    pf.close()


@pytest.mark.gpu
@pytest.mark.parametrize("G,M,seed", [(2, 5000, 0), (3, 70001, 1), (4, 1 << 18, 2)])
def test_shard_stages_on_one_gpu(G, M, seed):
    """The shard stages of the C ABI (max / exact sum / scan / emit / load: the NCCL-exchange variant's
    kernels, k_select_gather<records> with non-trivial q_offset, j_global0 and output ranges) driven
    by hand for G handles that share ONE GPU, so the single-GPU test run covers them too: ancestors
    bit-exact against the oracle and against one handle holding every particle."""
    import torch
    from oracle import c_oracle, pf_oracle as o
    from smarc_navigation_b200 import ParticleFilter
    from smarc_navigation_b200.sharded import GpuShardOps, plan_exchange, splits_for
    rng = np.random.default_rng(seed)
    N = G * M
    logw = (rng.normal(0, 4.0, N) - 150).astype(np.float32)
    if seed % 2:
        logw[: N // 2 + 3] -= 60.0                      # almost all mass on the upper shards
    logw[7] = -np.inf
    poses = np.zeros((N, 6))
    poses[:, 0] = np.arange(N)
    poses[:, 1] = -0.25 * np.arange(N)
    poses[:, 5] = rng.uniform(-3, 3, N).astype(np.float32)
    u = float(rng.uniform())
    pfs = [ParticleFilter(particle_count=M, global_offset=g * M, global_count=N, seed=seed,
                          resampling_noise_covariance=[0.0] * 6) for g in range(G)]
    ops = [GpuShardOps(pf) for pf in pfs]
    for g, pf in enumerate(pfs):
        pf.set_poses(poses[g * M:(g + 1) * M])
        pf.set_log_weights(logw[g * M:(g + 1) * M])
    m = max(op.shard_max() for op in ops)
    limbs = np.sum([op.shard_sum(m) for op in ops], axis=0, dtype=object)
    A = int(limbs[0]) + (int(limbs[1]) << 32)
    Q_all = [op.shard_scan(m, A) for op in ops]
    D, q_offsets, bounds = plan_exchange(Q_all, M, u)
    sends = [ops[g].shard_emit(u, q_offsets[g], D, bounds[g], bounds[g + 1]) for g in range(G)]
    everything = torch.cat(sends)                        # offspring in global output order
    assert everything.shape[0] == N
    for g in range(G):
        send_counts, recv_counts = splits_for(bounds, M, g)
        assert sum(recv_counts) == M and sum(send_counts) == sends[g].shape[0]
        ops[g].shard_load(everything[g * M:(g + 1) * M].contiguous())
    got = np.concatenate([pf.poses() for pf in pfs])
    anc = c_oracle.systematic(o.weights_from_q(c_oracle.quantise_logw(logw)[0]), u)
    np.testing.assert_array_equal(got[:, 0], poses[anc, 0])
    np.testing.assert_array_equal(got[:, 1], poses[anc, 1])
    np.testing.assert_array_equal(got[:, 5], poses[anc, 5])
    one = ParticleFilter(particle_count=N, resampling_noise_covariance=[0.0] * 6, seed=seed)
    one.set_poses(poses)
    one.set_log_weights(logw)
    np.testing.assert_array_equal(one.resample(u=u, return_ancestors=True, roughen=False), anc)
    for pf in pfs + [one]:
        pf.close()


@pytest.mark.gpu
@pytest.mark.parametrize("G,M,seed", [(2, 6000, 3), (3, 1 << 17, 4)])
def test_device_resident_shard_stages_on_one_gpu(G, M, seed):
    """The DEVICE-resident stages ShardedParticleFilter uses (auvpf_shard_max_dev / sum_dev / scan_dev:
    scalars reduced in place on the library's buffers, then auvpf_shard_emit), with the collectives
    done by hand for G handles on ONE GPU.  scan_dev must leave a finished inclusive CDF and a tile
    table -- what a shard's peers and auvpf_shard_emit read -- not the single-GPU kernel's layout."""
    import torch
    from oracle import c_oracle, pf_oracle as o
    from smarc_navigation_b200 import ParticleFilter, _lib
    from smarc_navigation_b200.sharded import GpuShardOps, plan_exchange
    rng = np.random.default_rng(seed)
    N = G * M
    logw = (rng.normal(0, 3.0, N) - 90).astype(np.float32)
    if seed % 2:
        logw[: N // 2 + 3] -= 50.0
    poses = np.zeros((N, 6))
    poses[:, 0] = np.arange(N)
    poses[:, 1] = 0.5 * np.arange(N)
    u = float(rng.uniform())
    pfs = [ParticleFilter(particle_count=M, global_offset=g * M, global_count=N, seed=seed,
                          resampling_noise_covariance=[0.0] * 6) for g in range(G)]
    ops = [GpuShardOps(pf) for pf in pfs]
    L = _lib.load()
    for g, pf in enumerate(pfs):
        pf.set_poses(poses[g * M:(g + 1) * M])
        pf.set_log_weights(logw[g * M:(g + 1) * M])

    def everyone(call):
        for pf in pfs:
            _lib.check(call(pf._h))
        for pf in pfs:
            pf.sync()

    everyone(L.auvpf_shard_max_dev)
    m = torch.stack([op.t_wmax for op in ops]).max(0).values          # all-reduce MAX on the order-preserving int32
    for op in ops:
        op.t_wmax.copy_(m)
    torch.cuda.synchronize()
    everyone(L.auvpf_shard_sum_dev)
    limbs = torch.stack([op.t_limbs for op in ops]).sum(0)            # all-reduce SUM of the exact integer limbs
    for op in ops:
        op.t_limbs.copy_(limbs)
    torch.cuda.synchronize()
    everyone(L.auvpf_shard_scan_dev)
    Q_all = [int(op.t_q.item()) for op in ops]                        # all-gather of the shards' CDF totals
    D, q_offsets, bounds = plan_exchange(Q_all, M, u)
    sends = [ops[g].shard_emit(u, q_offsets[g], D, bounds[g], bounds[g + 1]) for g in range(G)]
    everything = torch.cat(sends)
    assert everything.shape[0] == N
    for g in range(G):
        ops[g].shard_load(everything[g * M:(g + 1) * M].contiguous())
    got = np.concatenate([pf.poses() for pf in pfs])
    anc = c_oracle.systematic(o.weights_from_q(c_oracle.quantise_logw(logw)[0]), u)
    np.testing.assert_array_equal(got[:, 0], poses[anc, 0])
    np.testing.assert_array_equal(got[:, 1], poses[anc, 1])
    for pf in pfs:
        pf.close()
