"""Particle-sharded filter: one process per GPU, particles split in contiguous
global ranges.

The normalised CDF is a quotient of prefix sums, ``cdf_j = floor(C_j 2^52 / A)`` with
``C_j`` the sum of the integer weights ``a_i`` up to ``j`` over ALL shards and ``A`` the global
total, so a shard's CDF is its local prefix shifted by the integer mass of the shards below it.
What crosses GPUs is therefore

1. the global maximum of the log-weights (the ``a_i`` refer to it),
2. six integers per rank: its total (two words), the three integer sums of the N_eff gate, and
   which log-weight buffer it used,
3. the particles of cross-shard ancestors.

``exchange="p2p"`` (default): all three through peer memory (CUDA IPC mappings of the ranks' slabs):
the weights kernel posts (1) and (2) into every peer's mailbox, the select kernel waits for them and
reads the owners' log-weights, chunk tables and state over NVLink -- ``auvpf_shard_resample``, two
kernels, no NCCL call and no host round trip in the cycle.  The estimate is reduced the same way:
every rank's 16 pose sums are gathered with ONE small ``all_gather`` on the device and read back once.

``exchange="nccl"``: the portable variant.  (1) ``all_reduce(MAX)`` in place on the library's
order-preserving int32 maximum, (2) ``all_gather`` of the six integers, then the host plans contiguous
splits (systematic ancestors are globally sorted, so the offspring of shard g's particles are ONE
contiguous range of global output slots), each shard emits its offspring as 32-byte records and one
``all_to_all_single`` delivers them.

Because every cross-GPU quantity is an integer, the ancestors are bit-identical to the single-GPU
(and to the reference's ``systematic_resample``) result for any number of GPUs.

The per-shard compute of the staged variant is behind a small ``ops`` interface so that the
orchestration can be exercised with ``gloo`` on CPU (tests/test_sharded_gloo.py plugs the oracle in);
the product ``GpuShardOps`` calls the C ABI.
"""
import ctypes as C
import os

import numpy as np

Q_BITS = 52
Q_ONE = 1 << Q_BITS
TOTAL_WORDS = 6         # A low, A high, S1, S2 low, S2 high, log-weight buffer id


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


def totals_to_ints(row):
    """One rank's six words -> (A, S1, S2_lo, S2_hi, buffer)."""
    r = [int(v) & 0xFFFFFFFFFFFFFFFF for v in row]
    return r[0] | (r[1] << 64), r[2], r[3], r[4], r[5]


def neff_fixed(S1, S2_lo, S2_hi):
    """``auvpf_neff_fixed`` from the summed integers (csrc/auvpf_math.cuh ``neff_fixed_a``)."""
    S2 = np.float64(S2_hi) * np.float64(4294967296.0) + np.float64(S2_lo)
    s1 = np.float64(S1)
    return float(s1 * s1 / S2) if S2 > 0 else 0.0


def plan_exchange(A_all, M, u):
    """From the gathered per-shard integer totals compute, identically on every rank: the mass
    ``c_off[g]`` below every shard (``c_off[G]`` = the global total) and ``bounds[g] .. bounds[g+1]`` =
    the global output slots produced by shard g's particles (``count_below`` of the cdf value at the
    shard's end, ``floor(c_off[g+1] 2^52 / A)``)."""
    G = len(A_all)
    N = G * M
    c_off, acc = [], 0
    for g in range(G):
        c_off.append(acc)
        acc += int(A_all[g])
    c_off.append(acc)
    A = acc
    if A <= 0:
        raise FloatingPointError("degenerate weights on every shard")
    bounds = [0]
    for g in range(G):
        cdf_end = (c_off[g + 1] << Q_BITS) // A
        bounds.append(count_below(u, N, cdf_end) if g < G - 1 else N)
    for g in range(G):
        bounds[g + 1] = max(bounds[g + 1], bounds[g])
    return c_off, bounds


def splits_for(bounds, M, rank):
    """(send_counts, recv_counts) in records for ``all_to_all_single``."""
    G = len(bounds) - 1

    def overlap(a0, a1, b0, b1):
        return max(0, min(a1, b1) - max(a0, b0))
    send = [overlap(bounds[rank], bounds[rank + 1], r * M, (r + 1) * M) for r in range(G)]
    recv = [overlap(bounds[s], bounds[s + 1], rank * M, (rank + 1) * M) for s in range(G)]
    return send, recv


def _as_i64(v):
    v = int(v) & 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >= (1 << 63) else v


def resample_sharded(ops, dist, u, group=None, neff_ratio=0.0):
    """One sharded normalise + systematic resample with the scalars staged through the host.
    ``ops`` provides the per-shard stages (see ``GpuShardOps``); ``dist`` is ``torch.distributed``.
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
    # 2. integer weights against the global max: this shard's totals; gather everybody's
    mine = torch.tensor([_as_i64(v) for v in ops.shard_sum(m)], dtype=torch.int64, device=dev)
    everyone = [torch.zeros_like(mine) for _ in range(G)]
    dist.all_gather(everyone, mine, group=group)
    table = [t.tolist() for t in everyone]
    rows = [totals_to_ints(r) for r in table]
    c_off, bounds = plan_exchange([r[0] for r in rows], M, u)
    if neff_ratio > 0.0 and neff_fixed(sum(r[1] for r in rows), sum(r[2] for r in rows), sum(r[3] for r in rows)) >= neff_ratio * G * M:
        return dict(m=m, c_off=c_off, bounds=bounds, skipped=True)
    # 3. emit this shard's offspring, 4. exchange, 5. load
    send_counts, recv_counts = splits_for(bounds, M, rank)
    send = ops.shard_emit(u, table, bounds[rank], bounds[rank + 1], rank)     # (n_out, 4) float64 records
    recv = torch.empty((M, 4), dtype=torch.float64, device=dev)
    assert sum(recv_counts) == M and sum(send_counts) == send.shape[0]
    dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group)
    ops.shard_load(recv)
    return dict(m=m, c_off=c_off, bounds=bounds, send=send_counts, recv=recv_counts, skipped=False)


class _DeviceArray(object):
    """Zero-copy torch view of library-owned device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 2}


def resample_sharded_device(ops, dist, u, group=None, rough_var=None, neff_ratio=0.0):
    """The staged resample with the scalars kept ON THE DEVICE: the collectives run in place on the
    library's own buffers, ordered on the library's stream, and the only host read-back is the G x 6
    table of integers that sizes the exchange."""
    import torch
    rank, G = dist.get_rank(group), dist.get_world_size(group)
    M = ops.M
    L, h = ops.L, ops.pf._h
    with torch.cuda.stream(ops.stream):
        ops._lib.check(L.auvpf_shard_max_dev(h))
        dist.all_reduce(ops.t_wmax, op=dist.ReduceOp.MAX, group=group)          # order-preserving int32 == float max
        ops._lib.check(L.auvpf_shard_weights_dev(h))
        dist.all_gather_into_tensor(ops.t_totals_all, ops.t_totals, group=group)
        rows = [totals_to_ints(r) for r in ops.t_totals_all.view(G, TOTAL_WORDS).tolist()]     # the one host synchronisation
        if sum(r[0] for r in rows) == 0:
            raise FloatingPointError("degenerate weights on every shard")
        c_off, bounds = plan_exchange([r[0] for r in rows], M, u)
        if neff_ratio > 0.0 and neff_fixed(sum(r[1] for r in rows), sum(r[2] for r in rows), sum(r[3] for r in rows)) >= neff_ratio * G * M:
            return dict(c_off=c_off, bounds=bounds, skipped=True)
        send_counts, recv_counts = splits_for(bounds, M, rank)
        n_out = bounds[rank + 1] - bounds[rank]
        send = ops.send_records(n_out)
        ops._lib.check(L.auvpf_shard_emit(h, u, C.c_void_p(ops.t_totals_all.data_ptr()), G, rank, bounds[rank], bounds[rank + 1],
                                          C.c_void_p(send.data_ptr()) if n_out > 0 else None, None))
        assert sum(recv_counts) == M and sum(send_counts) == n_out
        dist.all_to_all_single(ops.recv_buf, send, output_split_sizes=recv_counts, input_split_sizes=send_counts,
                               group=group)
        rv = None if rough_var is None else (C.c_double * 6)(*rough_var)
        ops._lib.check(L.auvpf_shard_load_roughen(h, C.c_void_p(ops.recv_buf.data_ptr()), M, rv, None))
    return dict(c_off=c_off, bounds=bounds, send=send_counts, recv=recv_counts, skipped=False)


def resample_sharded_p2p(ops, u, rough_var=None, ancestors=None):
    """Sharded resample over peer memory: ONE library call, two kernels, nothing the host or NCCL has to
    do (``auvpf_shard_resample``; the peers were mapped by ``GpuShardOps.connect``)."""
    rv = None if rough_var is None else (C.c_double * 6)(*rough_var)
    ops._lib.check(ops.L.auvpf_shard_resample(ops.pf._h, u, rv, None,
                                             None if ancestors is None else C.c_void_p(ancestors.data_ptr())))
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
        pw, pt = C.c_void_p(), C.c_void_p()
        _lib.check(self.L.auvpf_shard_buffers(pf._h, C.byref(pw), C.byref(pt)))
        self.t_wmax = torch.as_tensor(_DeviceArray(pw.value, (1,), "<i4"), device=self.device)
        self.t_totals = torch.as_tensor(_DeviceArray(pt.value, (TOTAL_WORDS,), "<i8"), device=self.device)
        self.t_totals_all = None
        self.send_buf = self.recv_buf = None
        self.p2p = False

    def prepare(self, world_size):
        t = self.torch
        if self.t_totals_all is None:
            self.t_totals_all = t.zeros(world_size * TOTAL_WORDS, dtype=t.int64, device=self.device)
        if self.send_buf is None and not self.p2p:
            self.send_buf = t.empty((2 * self.M, 4), dtype=t.float64, device=self.device)
            self.recv_buf = t.empty((self.M, 4), dtype=t.float64, device=self.device)

    def connect(self, dist, group=None):
        """Exchange CUDA IPC handles of the slabs and map the peers' memory.  Collective:
        every rank takes part in the all-gather even when its own export failed, and no rank imports
        unless every export succeeded.  Returns an error (or None); the caller agrees on the outcome."""
        t = self.torch
        G, rank = dist.get_world_size(group), dist.get_rank(group)
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

    # -- host-staged stages (tests, and the gloo-testable orchestration `resample_sharded`) --------------
    def shard_max(self):
        out = (C.c_float * 1)()
        self._lib.check(self.L.auvpf_shard_max(self.pf._h, out))
        return float(out[0])

    def shard_sum(self, m):
        mm = (C.c_float * 1)(m)
        out = (C.c_uint64 * TOTAL_WORDS)()
        self._lib.check(self.L.auvpf_shard_sum(self.pf._h, mm, out))
        return [int(v) for v in out]

    def shard_emit(self, u, totals_all, lo, hi, rank):
        G = len(totals_all)
        flat = (C.c_uint64 * (G * TOTAL_WORDS))(*[int(v) & 0xFFFFFFFFFFFFFFFF for row in totals_all for v in row])
        rec = self.torch.empty((max(hi - lo, 0), 4), dtype=self.torch.float64, device=self.device)
        self._lib.check(self.L.auvpf_shard_emit(self.pf._h, u, flat, G, rank, lo, hi,
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

    def __init__(self, particles_per_rank, dist, group=None, exchange="p2p", resample_neff_ratio=0.0, **kw):
        """``exchange``: "p2p" (default; mailbox + one select kernel over CUDA-IPC peer memory, up to 8
        ranks of one node) or "nccl" (scalar collectives + records + ``all_to_all_single``; any topology
        NCCL supports).  ``resample_neff_ratio``: the N_eff gate (auv_pf.py:174-181) on the GLOBAL sums."""
        import torch
        from .filter import ParticleFilter
        self.dist, self.group = dist, group
        self.rank, self.G = dist.get_rank(group), dist.get_world_size(group)
        self.M = int(particles_per_rank)
        self.N = self.M * self.G
        kw.setdefault("device", torch.cuda.current_device())
        self.pf = ParticleFilter(particle_count=self.M, global_offset=self.rank * self.M, global_count=self.N,
                                 resample_neff_ratio=resample_neff_ratio, **kw)
        self.neff_ratio = float(resample_neff_ratio)
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
        self._sums = None

    @property
    def resample_kernels(self):
        """What a resample launches on this path (bench.py labels its roofline line with it)."""
        if self.ops.p2p:
            return ("k_weights (maximum and totals exchanged through the peer-memory mailbox) + k_select<peers> "
                    "(select + gather + exchange over NVLink)")
        return ("k_weights + all_reduce(MAX) + all_gather(totals) + k_select<records> + all_to_all_single + "
                "k_load_records")

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

    def resample(self, u=None, roughen=True, check=False, ancestors=None):
        """``ancestors``: optional (M,) int64 CUDA tensor for this rank's GLOBAL ancestor ids (p2p only)."""
        if u is None:      # same Philox-derived uniform on every rank, no communication
            from .rng import resample_uniform
            u = resample_uniform(self._seed, 0, self._step)
        self._step += 1
        rv = self.pf.res_noise_cov if roughen else None
        if self.ops.p2p:
            plan = resample_sharded_p2p(self.ops, float(u), rough_var=rv, ancestors=ancestors)
            if check and self.ops.degenerate():
                raise FloatingPointError("degenerate weights on every shard")
            return plan
        self.ops.prepare(self.G)
        return resample_sharded_device(self.ops, self.dist, float(u), self.group, rough_var=rv, neff_ratio=self.neff_ratio)

    def resample_skipped(self):
        return bool(self.pf.resample_skipped()[0])

    def mean_cov(self):
        """Global estimate (auv_pf.py:218-252): every rank's 16 pose sums stay on the device; with peer memory they
        are gathered by ONE small collective on the library's stream -- or, with ``AUVPF_ESTIMATE_MAILBOX=1`` and peer
        memory, travel through the mailbox like the resample's scalars (``auvpf_shard_pose_sums``); read back ONCE and
        combined in float64 in rank order, so every rank returns the same numbers."""
        torch = self.ops.torch
        L, h = self.ops.L, self.pf._h
        if self.exchange == "p2p" and os.environ.get("AUVPF_ESTIMATE_MAILBOX", "0") == "1":
            # through the mailbox in peer memory, like the resample: no collective library, one read-back.  OPT-IN:
            # verified with 2-8 ranks on one GPU and on two GPUs, but the round's GPU budget ended before an 8-GPU run
            s = np.empty((self.G, 16), np.float64)
            self.ops._lib.check(L.auvpf_shard_pose_sums(h, s.ctypes.data_as(C.c_void_p)))
            return combine_pose_sums(s, self.M)
        if self._sums is None:
            p = C.c_void_p()
            self.ops._lib.check(L.auvpf_pose_sums_buffer(h, C.byref(p)))
            self._sums = torch.as_tensor(_DeviceArray(p.value, (16,), "<f8"), device=self.ops.device)
            self._sums_all = torch.empty(16 * self.G, dtype=torch.float64, device=self.ops.device)
        with torch.cuda.stream(self.ops.stream):
            self.ops._lib.check(L.auvpf_pose_sums_dev(h))
            self.dist.all_gather_into_tensor(self._sums_all, self._sums, group=self.group)
            s = self._sums_all.cpu().numpy().reshape(self.G, 16)               # the one read-back (stream-ordered)
        return combine_pose_sums(s, self.M)

    def poses(self):
        return self.pf.poses()

    def close(self):
        self.pf.close()


def combine_pose_sums(s, M):
    """(G, 16) per-shard sums of ``k_pose_stats`` -- [sum d_xyz, sum roll, pitch, yaw, wrapped yaw, sum d d^T
    (xx yy zz xy xz yz), c_xyz] with d = position - c, c the shard's first particle -- -> the global
    ``(mean(6), cov(3,3))`` of auv_pf.update_loc_pose with the reference's fill pattern."""
    s = np.asarray(s, np.float64)
    G = s.shape[0]
    n = float(G * M)
    c = s[:, 13:16]
    c0 = c[0]
    off = c - c0                                          # every shard's reference point relative to shard 0's
    sd = s[:, 0:3] + M * off                              # sums of (pos - c0)
    mean_d = sd.sum(0) / n
    mean = np.empty(6)
    mean[0:3] = c0 + mean_d
    mean[3] = s[:, 3].sum() / n
    mean[4] = s[:, 4].sum() / n
    mean[5] = s[:, 6].sum() / n                           # mean of the wrapped yaws (auv_pf.py:229-231)
    # second moments about c0: sum (d + off)(d + off)^T = sum d d^T + off sum d^T + sum d off^T + M off off^T
    pairs = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]
    cov = np.zeros((3, 3))
    for k, (i, j) in enumerate(pairs):
        m2 = (s[:, 7 + k] + off[:, i] * s[:, j] + s[:, i] * off[:, j] + M * off[:, i] * off[:, j]).sum() / n
        cov[i, j] = m2 - mean_d[i] * mean_d[j]
    cov[1, 0] = cov[0, 1]                                 # only [1,0] is mirrored (auv_pf.py:246)
    return mean, cov
