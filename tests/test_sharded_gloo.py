"""N>1 path on CPU: the sharded-resample orchestration (collectives, CDF offsets,
contiguous exchange splits) of smarc_navigation_b200/sharded.py under
world_size-2/3 ``gloo``, with the ORACLE standing in for the per-shard CUDA stages.
The result must equal the single-process reference systematic resample."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShardOps(object):
    """numpy/oracle restatement of the auvpf_shard_* stages (include/auvpf.h)."""

    def __init__(self, logw, poses, rank, M):
        from oracle import pf_oracle as o
        self.o = o
        self.logw = logw.astype(np.float32)
        self.poses = poses.copy()
        self.rank, self.M = rank, M
        self.device = torch.device("cpu")

    def shard_max(self):
        f = self.logw[~np.isnan(self.logw)]
        return float(f.max()) if f.size else -np.inf

    def _a(self, m):
        return self.o.integer_weights(self.logw, np.float32(m))

    def shard_sum(self, m):
        """This shard's six words: total of the integer weights (low, high), the gate sums, buffer id."""
        a = self._a(m)
        self.a = a
        A = int(a.astype(object).sum())
        S1, S2lo, S2hi = self.o.neff_sums(a)
        return [A & 0xFFFFFFFFFFFFFFFF, A >> 64, S1, S2lo, S2hi, 0]

    def shard_emit(self, u, totals_all, lo, hi, rank):
        """Offspring [lo, hi) of this shard's sources in the GLOBAL cdf = floor((c_off + local prefix) 2^52 / A)."""
        from smarc_navigation_b200.sharded import sys_position, totals_to_ints
        assert rank == self.rank
        As = [totals_to_ints(r)[0] for r in totals_all]
        c_off, A = sum(As[:rank]), sum(As)
        C = c_off + np.cumsum(self.a.astype(object))
        full = np.array([(int(c) << 52) // A for c in C], dtype=np.uint64).astype(np.float64) * 2.0 ** -52
        rec = np.empty((hi - lo, 4))
        for k, i in enumerate(range(lo, hi)):
            pos = sys_position(u, i, self.N_global)
            j = min(int(np.searchsorted(full, pos, side="right")), self.M - 1)
            rec[k, 0:2] = self.poses[j, 0:2]
            rec[k, 2] = self.rank * self.M + j          # smuggle the global ancestor id for the check
            rec[k, 3] = self.poses[j, 5]
        return torch.from_numpy(rec)

    def shard_load(self, rec):
        self.received = rec.numpy().copy()


def _worker(rank, world, port, M, seed, tmp):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from smarc_navigation_b200.sharded import resample_sharded
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    N = M * world
    logw = (rng.normal(0, 6.0, N) - 470).astype(np.float32)
    if seed % 2:
        logw[: N // 2 + 3] -= 60.0                      # almost all mass on the upper shards
    poses = np.zeros((N, 6))
    poses[:, 0] = np.arange(N)
    poses[:, 5] = rng.uniform(-3, 3, N)
    u = float(rng.uniform())
    ops = OracleShardOps(logw[rank * M:(rank + 1) * M], poses[rank * M:(rank + 1) * M], rank, M)
    ops.N_global = N
    plan = resample_sharded(ops, dist, u)
    # the N_eff gate on the GLOBAL integer sums: every rank takes the same decision, the oracle's
    gated_lo = resample_sharded(ops, dist, u, neff_ratio=1e-9)["skipped"]       # N_eff is never below 1e-9 N: held back
    gated_hi = resample_sharded(ops, dist, u, neff_ratio=1.0)["skipped"]        # ... and (almost) always below N
    np.savez(os.path.join(tmp, "r%d.npz" % rank), rec=ops.received, bounds=np.array(plan["bounds"]), u=u, logw=logw,
             gated_lo=gated_lo, gated_hi=gated_hi)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,M,seed", [(2, 500, 0), (2, 1024, 1), (3, 333, 2), (3, 64, 3)])
def test_sharded_resample_equals_single_process(tmp_path, world, M, seed):
    from oracle import pf_oracle as o
    mp.spawn(_worker, args=(world, _free_port(), M, seed, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    logw, u = parts[0]["logw"], float(parts[0]["u"])
    qq = o.quantise_logw(logw)[0]
    anc = o.systematic_resample(o.weights_from_q(qq), u)
    got = np.concatenate([p["rec"][:, 2] for p in parts]).astype(np.int64)
    np.testing.assert_array_equal(got, anc)                                  # bit-exact for any G
    np.testing.assert_array_equal(np.concatenate([p["rec"][:, 0] for p in parts]), anc.astype(np.float64))
    b = parts[0]["bounds"]
    assert b[0] == 0 and b[-1] == world * M and np.all(np.diff(b) >= 0)
    for g in range(world):                                                   # bounds == ancestors' shard ownership
        assert np.all(anc[b[g]:b[g + 1]] // M == g)
    a = o.quantise_logw(logw)[1]
    assert all(bool(p["gated_lo"]) for p in parts)
    assert all(bool(p["gated_hi"]) == (not o.resample_gate(a, 1.0)) for p in parts)


def test_plan_exchange_properties():
    from smarc_navigation_b200.sharded import plan_exchange, splits_for, count_below
    rng = np.random.default_rng(0)
    for _ in range(50):
        G, M = int(rng.integers(1, 9)), int(rng.integers(1, 5000))
        A_all = [int(v) for v in (rng.dirichlet(np.ones(G) * 0.3) * float(1 << 70)).astype(object)]
        A_all = [int(a) * int(rng.integers(0, 2) if G > 1 else 1) + (1 if g == 0 else 0) for g, a in enumerate(A_all)]
        u = float(rng.uniform())
        c_off, bounds = plan_exchange(A_all, M, u)
        assert c_off[0] == 0 and c_off[-1] == sum(A_all) and bounds[0] == 0 and bounds[-1] == G * M
        tot_send = np.zeros((G, G), int)
        for r in range(G):
            s, rc = splits_for(bounds, M, r)
            assert sum(rc) == M and sum(s) == bounds[r + 1] - bounds[r]
            tot_send[r] = s
        for r in range(G):
            s, rc = splits_for(bounds, M, r)
            assert list(tot_send[:, r]) == rc
    assert count_below(0.5, 10, 1 << 52) == 10 and count_below(0.5, 10, 0) == 0


def test_combine_pose_sums_is_the_global_estimate():
    """The sharded estimate: per-shard one-pass sums about the shard's own first particle, combined about
    shard 0's -- equals update_loc_pose's numbers on the concatenated particle set."""
    from oracle import pf_oracle as o
    from smarc_navigation_b200.sharded import combine_pose_sums
    rng = np.random.default_rng(5)
    G, M = 4, 777
    poses = np.column_stack([1500.0 + rng.normal(0, 3, G * M), -800.0 + rng.normal(0, 2, G * M), -5 + rng.normal(0, 0.1, G * M),
                             rng.normal(0, 0.02, G * M), rng.normal(0, 0.02, G * M), rng.uniform(-3.5, 3.5, G * M)])
    sums = np.zeros((G, 16))
    for g in range(G):
        p = poses[g * M:(g + 1) * M]
        c = p[0, 0:3]
        d = p[:, 0:3] - c
        sums[g, 0:3] = d.sum(0)
        sums[g, 3:6] = p[:, 3:6].sum(0)
        sums[g, 6] = ((p[:, 5] + np.pi) % (2 * np.pi) - np.pi).sum()
        sums[g, 7:13] = [(d[:, i] * d[:, j]).sum() for i, j in ((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))]
        sums[g, 13:16] = c
    mean, cov = combine_pose_sums(sums, M)
    em, ec, _ = o.estimate(poses)
    np.testing.assert_allclose(mean, em, atol=1e-9)
    np.testing.assert_allclose(cov, ec, atol=1e-9)
