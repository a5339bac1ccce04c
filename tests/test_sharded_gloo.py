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
        o = self.o
        x = self.logw.astype(np.float64) - np.float64(np.float32(m))
        x = np.where(np.isnan(x), -np.inf, x)
        live = x >= -745.0
        p, k = o.exp_det(np.where(live, x, 0.0))
        s = k + o.A_BITS
        live &= s >= -1
        a = np.zeros(self.M, np.uint64)
        a[live] = np.floor(np.ldexp(p[live], s[live].astype(np.int64))).astype(np.uint64)
        return a

    def shard_sum(self, m):
        a = self._a(m)
        return [int((a & np.uint64(0xFFFFFFFF)).sum(dtype=np.uint64)), int((a >> np.uint64(32)).sum(dtype=np.uint64))]

    def shard_scan(self, m, A):
        a = self._a(m)
        self.q = np.array([(int(v) << 52) // A for v in a.astype(object)], dtype=np.uint64)
        self.cdf = np.cumsum(self.q)
        return int(self.cdf[-1])

    def shard_emit(self, u, q_offset, D, lo, hi):
        from smarc_navigation_b200.sharded import sys_position
        G_N = self.N_global
        jg1 = self.rank * self.M + np.arange(self.M) + 1
        full = (q_offset + self.cdf.astype(object) + np.minimum(jg1, D)).astype(np.float64) * 2.0 ** -52
        rec = np.empty((hi - lo, 4))
        for k, i in enumerate(range(lo, hi)):
            pos = sys_position(u, i, G_N)
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
    np.savez(os.path.join(tmp, "r%d.npz" % rank), rec=ops.received, bounds=np.array(plan["bounds"]), u=u, logw=logw,
             D=plan["D"])
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


def test_plan_exchange_properties():
    from smarc_navigation_b200.sharded import plan_exchange, splits_for, count_below
    rng = np.random.default_rng(0)
    for _ in range(50):
        G, M = int(rng.integers(1, 9)), int(rng.integers(1, 5000))
        w = rng.dirichlet(np.ones(G) * 0.3)
        Q = np.floor(w * (1 << 52)).astype(np.int64)
        Q[-1] -= int(rng.integers(0, G * M))                                  # leave a deficit < N
        if Q[-1] < 0:
            continue
        u = float(rng.uniform())
        D, off, bounds = plan_exchange(list(Q), M, u)
        assert D == (1 << 52) - int(Q.sum()) and bounds[0] == 0 and bounds[-1] == G * M
        tot_send = np.zeros((G, G), int)
        for r in range(G):
            s, rc = splits_for(bounds, M, r)
            assert sum(rc) == M and sum(s) == bounds[r + 1] - bounds[r]
            tot_send[r] = s
        for r in range(G):
            s, rc = splits_for(bounds, M, r)
            assert list(tot_send[:, r]) == rc
    assert count_below(0.5, 10, 1 << 52) == 10 and count_below(0.5, 10, 0) == 0
