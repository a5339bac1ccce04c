"""Particle-sharded filter: one process per GPU, particles split in contiguous
global ranges, ``torch.distributed`` (NCCL over NVLink) for the three things that
cross GPUs:

1. the global log-sum-exp of the weights -- ``all_reduce(MAX)`` of the local
   maxima, then ``all_reduce(SUM)`` of the local fixed-point sums (two integer
   limbs, so the total is EXACT and independent of the number of GPUs);
2. the resampling CDF offsets -- ``all_gather`` of the per-shard integer CDF
   totals;
3. the particle exchange for cross-shard ancestors -- systematic ancestors are
   globally sorted, so the offspring of shard g's particles are ONE contiguous
   range of global output slots; shard g gathers them locally into 32-byte
   records and an ``all_to_all_single`` with contiguous splits delivers each
   output slot to its owner.

Because every cross-GPU quantity is an integer, the ancestors are bit-identical
to the single-GPU (and to the reference's ``systematic_resample``) result for
any number of GPUs.

The per-shard compute is behind a small ``ops`` interface so that the
orchestration can be exercised with ``gloo`` on CPU (tests/test_sharded_gloo.py
plugs the oracle in); the product ``GpuShardOps`` calls the C ABI.
"""
import ctypes as C

import numpy as np

Q_ONE = 1 << 52


def sys_position(u, i, n):
    """positions[i] of resampling.py:157 exactly as NumPy / the kernels round it."""
    return (u + float(i)) / float(n)


def count_below(u, n, c):
    """#{i in [0, n): positions[i] < c * 2^-52} -- same fix-up walk as the device
    function ``sys_count_below`` (auvpf_math.cuh)."""
    cv = float(c) * 2.0 ** -52
    i = int(np.floor(cv * float(n) - u))
    i = min(max(i, 0), n)
    while i > 0 and not (sys_position(u, i - 1, n) < cv):
        i -= 1
    while i < n and sys_position(u, i, n) < cv:
        i += 1
    return i


def plan_exchange(Q_all, M, u):
    """From the gathered per-shard CDF totals compute, identically on every rank:
    deficit D, per-shard q_offset, and ``bounds[g] .. bounds[g+1]`` = the global
    output slots produced by shard g's particles."""
    G = len(Q_all)
    N = G * M
    D = Q_ONE - sum(int(q) for q in Q_all)
    if not (0 <= D < max(N, 1) + 1):
        raise ArithmeticError("inconsistent CDF totals (deficit %d)" % D)
    q_offsets, bounds, acc = [], [0], 0
    for g in range(G):
        q_offsets.append(acc)
        acc += int(Q_all[g])
        c_end = acc + min((g + 1) * M, D)
        bounds.append(count_below(u, N, c_end) if g < G - 1 else N)
    for g in range(G):
        bounds[g + 1] = max(bounds[g + 1], bounds[g])
    return D, q_offsets, bounds


def splits_for(bounds, M, rank):
    """(send_counts, recv_counts) in records for ``all_to_all_single``."""
    G = len(bounds) - 1

    def overlap(a0, a1, b0, b1):
        return max(0, min(a1, b1) - max(a0, b0))
    send = [overlap(bounds[rank], bounds[rank + 1], r * M, (r + 1) * M) for r in range(G)]
    recv = [overlap(bounds[s], bounds[s + 1], rank * M, (rank + 1) * M) for s in range(G)]
    return send, recv


def resample_sharded(ops, dist, u, group=None):
    """Run one sharded normalise + systematic resample.  ``ops`` provides the
    per-shard stages (see ``GpuShardOps``); ``dist`` is ``torch.distributed``.
    Returns the plan (for tests / diagnostics)."""
    import torch
    rank, G = dist.get_rank(group), dist.get_world_size(group)
    dev = ops.device
    M = ops.M
    # 1. global max
    m = torch.tensor([ops.shard_max()], dtype=torch.float32, device=dev)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    m = float(m.item())
    if not np.isfinite(m):
        raise FloatingPointError("degenerate weights on every shard")
    # 2. exact global fixed-point sum
    limbs = torch.tensor(ops.shard_sum(m), dtype=torch.int64, device=dev)
    dist.all_reduce(limbs, op=dist.ReduceOp.SUM, group=group)
    lo, hi = (int(v) for v in limbs.tolist())
    A = lo + (hi << 32)
    # 3. local scan, gather the totals
    Q = torch.tensor([ops.shard_scan(m, A)], dtype=torch.int64, device=dev)
    Q_all = [torch.zeros_like(Q) for _ in range(G)]
    dist.all_gather(Q_all, Q, group=group)
    Q_all = [int(q.item()) for q in Q_all]
    D, q_offsets, bounds = plan_exchange(Q_all, M, u)
    # 4. emit this shard's offspring, 5. exchange, 6. load
    send_counts, recv_counts = splits_for(bounds, M, rank)
    send = ops.shard_emit(u, q_offsets[rank], D, bounds[rank], bounds[rank + 1])     # (n_out, 4) float64 records
    recv = torch.empty((M, 4), dtype=torch.float64, device=dev)
    assert sum(recv_counts) == M and sum(send_counts) == send.shape[0]
    dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group)
    ops.shard_load(recv)
    return dict(m=m, A=A, Q_all=Q_all, D=D, bounds=bounds, send=send_counts, recv=recv_counts)


class _DeviceArray(object):
    """Zero-copy torch view of library-owned device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 2}


def resample_sharded_device(ops, dist, u, group=None, rough_var=None):
    """The same resample with the three cross-GPU scalars kept ON THE DEVICE: the collectives run
    in place on the library's own buffers, ordered on the library's stream, and the only host
    read-back is the list of per-shard CDF totals that sizes the exchange."""
    import torch
    rank, G = dist.get_rank(group), dist.get_world_size(group)
    M = ops.M
    L, h = ops.L, ops.pf._h
    with torch.cuda.stream(ops.stream):
        ops._lib.check(L.auvpf_shard_max_dev(h))
        dist.all_reduce(ops.t_wmax, op=dist.ReduceOp.MAX, group=group)          # order-preserving int32 == float max
        ops._lib.check(L.auvpf_shard_sum_dev(h))
        dist.all_reduce(ops.t_limbs, op=dist.ReduceOp.SUM, group=group)         # exact integer limbs of A
        ops._lib.check(L.auvpf_shard_scan_dev(h))
        dist.all_gather_into_tensor(ops.t_qall, ops.t_q, group=group)
        Q_all = [int(q) for q in ops.t_qall.tolist()]                           # the one host synchronisation
        if sum(Q_all) == 0:
            raise FloatingPointError("degenerate weights on every shard")
        D, q_offsets, bounds = plan_exchange(Q_all, M, u)
        send_counts, recv_counts = splits_for(bounds, M, rank)
        n_out = bounds[rank + 1] - bounds[rank]
        send = ops.send_records(n_out)
        ops._lib.check(L.auvpf_shard_emit(h, u, q_offsets[rank], D, bounds[rank], bounds[rank + 1],
                                          C.c_void_p(send.data_ptr()) if n_out > 0 else None, None))
        assert sum(recv_counts) == M and sum(send_counts) == n_out
        dist.all_to_all_single(ops.recv_buf, send, output_split_sizes=recv_counts, input_split_sizes=send_counts,
                               group=group)
        rv = None if rough_var is None else (C.c_double * 6)(*rough_var)
        ops._lib.check(L.auvpf_shard_load_roughen(h, C.c_void_p(ops.recv_buf.data_ptr()), M, rv, None))
    return dict(Q_all=Q_all, D=D, bounds=bounds, send=send_counts, recv=recv_counts)


def resample_sharded_p2p(ops, dist, u, group=None, rough_var=None):
    """Sharded resample with NO host synchronisation: the three scalar collectives run in place on
    the library's device buffers, and ONE kernel then selects this rank's ancestors in the global
    CDF and gathers their state straight from the owners' memory over NVLink (CUDA IPC peer
    mappings set up by ``GpuShardOps.connect``).  Returns nothing the host would have to wait for."""
    import torch
    L, h = ops.L, ops.pf._h
    with torch.cuda.stream(ops.stream):
        ops._lib.check(L.auvpf_shard_max_dev(h))
        dist.all_reduce(ops.t_wmax, op=dist.ReduceOp.MAX, group=group)          # order-preserving int32 == float max
        ops._lib.check(L.auvpf_shard_sum_dev(h))
        dist.all_reduce(ops.t_limbs, op=dist.ReduceOp.SUM, group=group)         # exact integer limbs of A
        ops._lib.check(L.auvpf_shard_scan_dev(h))
        # this all-gather is also the cross-GPU barrier that makes every shard's CDF and tile table
        # complete before any rank reads them
        dist.all_gather_into_tensor(ops.t_qall, ops.t_q, group=group)
        rv = None if rough_var is None else (C.c_double * 6)(*rough_var)
        ops._lib.check(L.auvpf_shard_gather_p2p(h, u, C.c_void_p(ops.t_qall.data_ptr()), rv, None, None))
    return {}


class GpuShardOps(object):
    """Per-shard stages on the GPU through the C ABI (auvpf_shard_*)."""

    def __init__(self, pf):
        import torch
        from . import _lib
        self.pf = pf
        self._lib = _lib
        self.L = _lib.load()
        self.M = pf.pc
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.torch = torch
        # device-resident flow: torch views of the library's scalars + persistent exchange buffers
        self.stream = torch.cuda.ExternalStream(pf.stream_ptr, device=self.device)
        pw, pl, pq = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(self.L.auvpf_shard_buffers(pf._h, C.byref(pw), C.byref(pl), C.byref(pq)))
        self.t_wmax = torch.as_tensor(_DeviceArray(pw.value, (1,), "<i4"), device=self.device)
        self.t_limbs = torch.as_tensor(_DeviceArray(pl.value, (2,), "<i8"), device=self.device)
        self.t_q = torch.as_tensor(_DeviceArray(pq.value, (1,), "<i8"), device=self.device)
        self.t_qall = None
        self.send_buf = self.recv_buf = None
        self.p2p = False

    def prepare(self, world_size):
        if self.t_qall is None:
            t = self.torch
            self.t_qall = t.zeros(world_size, dtype=t.int64, device=self.device)
            self.send_buf = t.empty((2 * self.M, 4), dtype=t.float64, device=self.device)
            self.recv_buf = t.empty((self.M, 4), dtype=t.float64, device=self.device)

    def connect(self, dist, group=None):
        """Exchange CUDA IPC handles of the state / CDF slabs and map the peers' memory.  Collective:
        every rank takes part in the all-gather even when its own export failed, and no rank imports
        unless every export succeeded.  Returns an error (or None); the caller agrees on the outcome."""
        t = self.torch
        G, rank = dist.get_world_size(group), dist.get_rank(group)
        self.prepare(G)
        nb = self._lib.IPC_BLOB_BYTES
        blob = (C.c_ubyte * nb)()
        err = None
        try:
            self._lib.check(self.L.auvpf_ipc_export(self.pf._h, blob))
        except Exception as e:              # noqa: BLE001
            err = e
        mine = t.tensor(list(bytes(blob)) + [0 if err else 1], dtype=t.uint8, device=self.device)
        everyone = t.empty(G * (nb + 1), dtype=t.uint8, device=self.device)
        dist.all_gather_into_tensor(everyone, mine, group=group)
        rows = everyone.cpu().numpy().reshape(G, nb + 1)
        if err is None and not rows[:, nb].all():
            err = RuntimeError("a peer could not export its memory")
        if err is None:
            raw = rows[:, :nb].tobytes()
            buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
            try:
                self._lib.check(self.L.auvpf_ipc_import(self.pf._h, buf, G, rank))
                self.p2p = True
            except Exception as e:          # noqa: BLE001
                err = e
        return err

    def degenerate(self):
        flag = C.c_int32(0)
        self._lib.check(self.L.auvpf_shard_degenerate(self.pf._h, C.byref(flag)))
        return bool(flag.value)

    def send_records(self, n_out):
        """(n_out, 4) float64 = n_out 32-byte records; grows when this shard's particles own more than
        the buffered share of the offspring (worst case world_size * M: all mass on this shard)."""
        if n_out > self.send_buf.shape[0]:
            self.stream.synchronize()
            self.send_buf = self.torch.empty((n_out + n_out // 4, 4), dtype=self.torch.float64, device=self.device)
        return self.send_buf[:max(n_out, 0)]

    def shard_max(self):
        out = (C.c_float * 1)()
        self._lib.check(self.L.auvpf_shard_max(self.pf._h, out))
        return float(out[0])

    def shard_sum(self, m):
        mm = (C.c_float * 1)(m)
        out = (C.c_uint64 * 2)()
        self._lib.check(self.L.auvpf_shard_sum(self.pf._h, mm, out))
        return [int(out[0]), int(out[1])]

    def shard_scan(self, m, A):
        mm = (C.c_float * 1)(m)
        AA = (C.c_uint64 * 2)(A & 0xFFFFFFFFFFFFFFFF, A >> 64)
        out = (C.c_uint64 * 1)()
        self._lib.check(self.L.auvpf_shard_scan(self.pf._h, mm, AA, out))
        return int(out[0])

    def shard_emit(self, u, q_offset, D, lo, hi):
        rec = self.torch.empty((max(hi - lo, 0), 4), dtype=self.torch.float64, device=self.device)
        self._lib.check(self.L.auvpf_shard_emit(self.pf._h, u, q_offset, D, lo, hi,
                                               C.c_void_p(rec.data_ptr()) if hi > lo else None, None))
        self.pf.sync()                       # records complete before NCCL reads them on its own stream
        return rec

    def shard_load(self, rec):
        self.torch.cuda.current_stream().synchronize()   # exchange complete before our stream reads it
        self._lib.check(self.L.auvpf_shard_load(self.pf._h, C.c_void_p(rec.data_ptr()), rec.shape[0]))
        self.pf.sync()


class ShardedParticleFilter(object):
    """``ParticleFilter`` over G ranks: same predict/update/resample methods; every rank
    calls them collectively with the same arguments."""

    def __init__(self, particles_per_rank, dist, group=None, exchange="p2p", **kw):
        """``exchange``: "p2p" (default; one kernel over CUDA-IPC peer memory, up to 8 ranks of one
        node) or "nccl" (records + ``all_to_all_single``; any topology NCCL supports)."""
        import torch
        from .filter import ParticleFilter
        self.dist, self.group = dist, group
        self.rank, self.G = dist.get_rank(group), dist.get_world_size(group)
        self.M = int(particles_per_rank)
        self.N = self.M * self.G
        kw.setdefault("device", torch.cuda.current_device())
        self.pf = ParticleFilter(particle_count=self.M, global_offset=self.rank * self.M, global_count=self.N, **kw)
        self.ops = GpuShardOps(self.pf)
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.exchange = exchange
        if exchange == "p2p":
            # every rank must end up on the same path: if mapping the peers' memory fails anywhere
            # (no peer access between two devices, IPC not permitted in this container) all ranks
            # use the NCCL record exchange instead -- still the GPU path, only with a host-sized plan
            err = self.ops.connect(dist, group)
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.ops.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self.ops.p2p = False
                self.exchange = "nccl"
                if self.rank == 0:
                    import warnings
                    warnings.warn("peer-memory exchange unavailable (%s); using the NCCL record exchange" % (err,))
        self._step = 0
        self._seed = int(kw.get("seed", 0))

    @property
    def resample_kernels(self):
        """What a resample launches on this path (bench.py labels its roofline line with it)."""
        if self.ops.p2p:
            return ("k_normalise + all_reduce(MAX, SUM) + k_scan_q (look-back) + all_gather(Q) + k_shard_meta + "
                    "k_select_gather_p2p (select + gather + exchange over peer memory)")
        return ("k_normalise + all_reduce(MAX, SUM) + k_scan_q (look-back) + all_gather(Q) + k_select_gather<records> + "
                "all_to_all_single + k_load_records")

    def set_dem(self, *a, **k):
        self.pf.set_dem(*a, **k)

    def set_beams(self, *a, **k):
        self.pf.set_beams(*a, **k)

    def predict(self, odom_t, dt=None, noise=None):
        self.pf.predict(odom_t, dt=dt, noise=noise)

    def predict_many(self, odoms, dt=None, noise=None):
        self.pf.predict_many(odoms, dt=dt, noise=noise)

    def update_mbes(self, ranges):
        self.pf.update_mbes(ranges, return_weights=False)

    def resample(self, u=None, roughen=True, check=False):
        if u is None:      # same Philox-derived uniform on every rank, no communication
            from .rng import resample_uniform
            u = resample_uniform(self._seed, 0, self._step)
        self._step += 1
        self.ops.prepare(self.G)
        rv = self.pf.res_noise_cov if roughen else None
        if self.ops.p2p:
            plan = resample_sharded_p2p(self.ops, self.dist, float(u), self.group, rough_var=rv)
            if check and self.ops.degenerate():
                raise FloatingPointError("degenerate weights on every shard")
            return plan
        return resample_sharded_device(self.ops, self.dist, float(u), self.group, rough_var=rv)

    def mean_cov(self):
        import torch
        mean, cov = self.pf.mean_cov()
        t = torch.tensor(np.concatenate([mean[0], cov[0].reshape(9)]), dtype=torch.float64, device=self.ops.device)
        allt = [torch.zeros_like(t) for _ in range(self.G)]
        self.dist.all_gather(allt, t, group=self.group)
        arr = np.stack([a.cpu().numpy() for a in allt])
        means, covs = arr[:, :6], arr[:, 6:].reshape(-1, 3, 3)
        gm = means.mean(0)
        d = means[:, :3] - gm[:3]
        gc = (covs + d[:, :, None] * d[:, None, :]).mean(0)
        gc[2, 0] = gc[2, 1] = 0.0          # the reference's fill pattern (auv_pf.py:238-246)
        return gm, gc

    def poses(self):
        return self.pf.poses()

    def close(self):
        self.pf.close()
