// auvpf_resample.cuh -- exact normalisation + systematic resampling in TWO kernels.
//
//   k_weights   one read of the log-weights: a_i = floor(exp(logw_i - max) 2^60), one exact sum per chunk
//               of 256 particles and per tile of 2048, the tile sums' prefix (last block of a filter), the
//               total A and the integer sums of the N_eff gate.  Nothing is written per particle.
//   k_select    a block owns 1024 consecutive output slots: it locates the tiles that hold their ancestors,
//               REBUILDS the prefix sums C_j of those tiles from the log-weights (one warp per chunk, no
//               block barrier), turns C_j into the run end count_below(floor(C_j 2^52 / A)) in closed form,
//               marks run heads, max-scans them and gathers the ancestors' state with the roughening fused.
//
// The CDF is cdf_j = floor(C_j 2^52 / A) (auvpf_math.cuh): a function of the PREFIX of the a_i, so the two
// kernels share nothing per particle -- no a_i / q_i array, no scan pass in between (round 1 ran three
// kernels and moved 16 B per particle more).  A particle-sharded filter runs the same two kernels: the
// shards' totals travel through a mailbox in peer memory (one flag-and-payload exchange for the maximum in
// front of k_weights, one for the totals in front of k_select), the select kernel reads the owners'
// log-weights, chunk tables and state over NVLink.  No NCCL call and no host round trip in the cycle.
#pragma once
#include <cstdio>

#include "auvpf_kernels.cuh"

namespace auvpf {

constexpr int kChunk = 256;                       // sources one warp rebuilds per pass: 32 lanes x 8
constexpr int kChunkItems = 8;
constexpr int kChunksPerTile = kScanTile / kChunk;   // 8: one per warp of a 256-thread block
static_assert(kChunksPerTile == kThreads / 32, "one chunk per warp of the weights kernel");

// ---------------------------------------------------------------------------------------------
// Mailbox in peer memory (particle-sharded filter).  Every rank owns kMboxStages x kMaxPeers slots of
// kMboxWords 64-bit words inside its slab; slot (stage, src) is written by rank `src` only: payload words,
// a system-scope fence, then the flag word = the epoch of this resample.  The reader spins on its OWN
// memory.  A slot is rewritten only after every reader consumed it: a rank posts stage 0 of epoch e+1 after
// its select kernel of epoch e, which waited for everybody's stage 1 of epoch e, which every rank posts
// after all its blocks consumed stage 0 of epoch e (and the same one stage later).
// Stages 0 and 1 carry the resample (maximum, totals); stages 2 .. 7 the ESTIMATE's 16 pose sums of a rank: three
// slots of 7 payload words, twice -- by the parity of the estimate's own epoch, because a rank may post estimate
// e + 1 while a slower peer still reads its slots of estimate e (it cannot post e + 2 before that peer has posted
// e + 1, i.e. is done reading e).
constexpr int kMboxStages = 8, kMboxWords = 8, kMboxFlag = 7, kMboxEstStage = 2, kMboxEstParts = 3;
constexpr size_t kMboxBytes = sizeof(unsigned long long) * kMboxStages * kMaxPeers * kMboxWords;
#ifndef AUVPF_MBOX_TIMEOUT_MS
#define AUVPF_MBOX_TIMEOUT_MS 20000
#endif
constexpr unsigned long long kMboxTimeoutNs = (unsigned long long)AUVPF_MBOX_TIMEOUT_MS * 1000000ull;

struct MailboxView {
    unsigned long long *peer[kMaxPeers];      // every rank's mailbox (own included), mapped into this process
    int G, rank;                              // G == 1: no exchange
    unsigned int epoch;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.global.release.sys.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.global.acquire.sys.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.global.relaxed.sys.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.global.relaxed.sys.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// one thread posts n payload words to every rank's slot (stage, my rank)
__device__ __forceinline__ void mbox_post(const MailboxView &mb, int stage, const unsigned long long *payload, int n)
{
    for (int g = 0; g < mb.G; ++g) {
        unsigned long long *slot = mb.peer[g] + (size_t)(stage * kMaxPeers + mb.rank) * kMboxWords;
        for (int k = 0; k < n; ++k) st_relaxed_sys(slot + k, payload[k]);
    }
    __threadfence_system();
    for (int g = 0; g < mb.G; ++g)
        st_release_sys(mb.peer[g] + (size_t)(stage * kMaxPeers + mb.rank) * kMboxWords + kMboxFlag, (unsigned long long)mb.epoch);
}
// one thread waits for rank src's post of this epoch and returns its slot (in this rank's own memory)
__device__ __forceinline__ const unsigned long long *mbox_wait(const MailboxView &mb, int stage, int src)
{
    const unsigned long long *slot = mb.peer[mb.rank] + (size_t)(stage * kMaxPeers + src) * kMboxWords;
    if (ld_acquire_sys(slot + kMboxFlag) == (unsigned long long)mb.epoch) return slot;
    // A peer that never posts (a rank died, or the ranks' calls got out of step) must not hang the GPU: after
    // kMboxTimeoutNs the kernel reports and traps, which fails the stream with an error the host sees.
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        for (int i = 0; i < 64; ++i) {
            if (ld_acquire_sys(slot + kMboxFlag) == (unsigned long long)mb.epoch) return slot;
            __nanosleep(64);
        }
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > kMboxTimeoutNs) {
            printf("auvpf mailbox: rank %d waited %llu ms for rank %d, stage %d, epoch %u (flag holds %llu)\n", mb.rank,
                   (t1 - t0) / 1000000ull, src, stage, mb.epoch, ld_acquire_sys(slot + kMboxFlag));
            __trap();
        }
    }
}

// ---- the estimate of a particle-sharded filter: every rank's 16 pose sums (k_pose_stats) to every rank -----------
// k_est_post: one thread posts this rank's sums; k_est_wait: thread g collects rank g's into out[g][16].  Two
// kernels, so that one process driving several ranks can enqueue all posts before any wait (auvpf_group_pose_sums).
__global__ void k_est_post(MailboxView mb, int parity, const double *__restrict__ sums)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int part = 0; part < kMboxEstParts; ++part) {
        unsigned long long w[kMboxFlag];
        const int n = min(kMboxFlag, 16 - kMboxFlag * part);
        for (int k = 0; k < n; ++k) w[k] = (unsigned long long)__double_as_longlong(sums[kMboxFlag * part + k]);
        mbox_post(mb, kMboxEstStage + parity * kMboxEstParts + part, w, n);
    }
}

__global__ void k_est_wait(MailboxView mb, int parity, double *__restrict__ out)
{
    const int g = threadIdx.x;
    if (blockIdx.x != 0 || g >= mb.G) return;
    for (int part = 0; part < kMboxEstParts; ++part) {
        const unsigned long long *slot = mbox_wait(mb, kMboxEstStage + parity * kMboxEstParts + part, g);
        const int n = min(kMboxFlag, 16 - kMboxFlag * part);
        for (int k = 0; k < n; ++k) out[g * 16 + kMboxFlag * part + k] = __longlong_as_double((long long)ld_relaxed_sys(slot + k));
    }
}

// ---------------------------------------------------------------------------------------------
// k_weights
struct WeightState {
    int *wmax_f32;                  // [F] ordered-int float32 max published by the update kernels, or nullptr ...
    double *wmax;                   // [F] ... in: the float64 max of a linear source (k_weight_max); out: the max used
    unsigned int *done;             // [F] ticket, zero on entry and on exit
    unsigned long long *tile_neff;  // [F * tiles * 3] scratch: per tile S1, S2 low, S2 high of the N_eff gate
    unsigned long long *neff_pub;   // [3F] out: this shard's sums
    U128 *A;                        // [F] out: this shard's total
    int *degenerate;                // [F] out (single GPU: no finite weight; sharded: decided by the select kernel)
    U128 *tile_prefix;              // [F * tiles] out: inclusive prefix of the tile sums
    U128 *chunk_sum;                // [F * tiles * 8] out: raw sums of 256 particles
    int tiles;
    int logw_parity;                // which of the two log-weight buffers the source is (told to the peers)
    int post_max;                   // sharded: this kernel posts the shard's maximum itself (0: k_post_max did, see auvpf_group_resample)
};

// the 8 raw log-weights of one lane (kind 0): rows [row0, row0 + 8), the first `left` of them exist
__device__ __forceinline__ void load_logw8(const float *logw, int64_t row0, int64_t left, float v[kChunkItems])
{
    if (left >= kChunkItems && (reinterpret_cast<uintptr_t>(logw + row0) & 15) == 0) {
        const float4 p = *reinterpret_cast<const float4 *>(logw + row0);
        const float4 q = *reinterpret_cast<const float4 *>(logw + row0 + 4);
        v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w;
        v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < kChunkItems; ++k) v[k] = k < left ? logw[row0 + k] : -INFINITY;
    }
}

// the 8 consecutive a_i of one lane: rows [row0, row0 + 8), the first `left` of them exist
__device__ __forceinline__ void load_a8(const WeightSrc &w, int64_t row0, int64_t left, double m, uint64_t a[kChunkItems])
{
    if (w.kind == 0) {
        float v[kChunkItems];
        load_logw8(w.logw, row0, left, v);
        const float mf = (float)m;           // (the maximum of float32 log-weights: exact)
#pragma unroll
        for (int k = 0; k < kChunkItems; ++k) a[k] = a_from_logw(v[k], mf);
    } else {
#pragma unroll
        for (int k = 0; k < kChunkItems; ++k) a[k] = k < left ? w.apre[row0 + k] : 0ull;
    }
}

__device__ __forceinline__ U128 shfl_u128(U128 v, int src)
{
    U128 r;
    r.lo = __shfl_sync(0xffffffffu, v.lo, src);
    r.hi = __shfl_sync(0xffffffffu, v.hi, src);
    return r;
}
__device__ __forceinline__ U128 shfl_up_u128(U128 v, int delta)
{
    U128 r;
    r.lo = __shfl_up_sync(0xffffffffu, v.lo, delta);
    r.hi = __shfl_up_sync(0xffffffffu, v.hi, delta);
    return r;
}

// The kernel is ONE wave of blocks (a tile of 2048 particles each) and as long as one block's chain of dependent
// global round trips, so the chain is kept short: the log-weights are requested before the maximum is known, the
// gate's partial sums travel with the tile sum as plain stores (no atomics besides the ticket), and on one GPU
// every thread reads the maximum itself (no thread-0 prologue, no barrier in front of the loads).
__global__ void __launch_bounds__(kThreads, 4) k_weights(WeightSrc w, int64_t M, WeightState ws, MailboxView mb)
{
    __shared__ unsigned long long s_part[kThreads / 32][5];
    __shared__ double s_m;
    __shared__ bool s_last;
    const int f = blockIdx.y, tile = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // programmatic dependent launch: the select kernel behind this one may be scheduled now -- its blocks set up
    // and then wait (cudaGridDependencySynchronize) until this grid has completed and its writes are visible
    cudaTriggerProgrammaticLaunchCompletion();
    const int64_t j0 = (int64_t)tile * kScanTile + (int64_t)threadIdx.x * kChunkItems;
    float v[kChunkItems];
    if (w.kind == 0) load_logw8(w.logw, (int64_t)f * M + j0, M - j0, v);
    // ---- the maximum: local, or exchanged through the mailbox (stage 0) ----
    double m;
    if (mb.G > 1) {
        if (threadIdx.x == 0) {
            if (tile == 0 && ws.post_max) {
                const unsigned long long mine = (unsigned long long)(unsigned int)ws.wmax_f32[0];
                mbox_post(mb, 0, &mine, 1);
            }
            int k = kWmaxReset;
            for (int g = 0; g < mb.G; ++g) k = max(k, (int)(unsigned int)ld_relaxed_sys(mbox_wait(mb, 0, g)));
            s_m = k > kWmaxReset ? (double)ordered_to_float(k) : -INFINITY;
        }
        __syncthreads();
        m = s_m;
    } else if (ws.wmax_f32) {
        const int k = ws.wmax_f32[f];
        m = k > kWmaxReset ? (double)ordered_to_float(k) : -INFINITY;
    } else
        m = ws.wmax[f];
    const bool bad = !isfinite(m) || (w.kind == 1 && !(m > 0.0));
    // ---- a_i of this tile: one chunk of 256 per warp, 8 consecutive particles per lane ----
    uint64_t a[kChunkItems];
    if (bad) {
#pragma unroll
        for (int k = 0; k < kChunkItems; ++k) a[k] = 0ull;
    } else if (w.kind == 0) {
        const float mf = (float)m;
#pragma unroll
        for (int k = 0; k < kChunkItems; ++k) a[k] = a_from_logw(v[k], mf);
    } else
        load_a8(w, (int64_t)f * M + j0, M - j0, m, a);
    unsigned long long lo = 0, hi = 0, s1 = 0, s2lo = 0, s2hi = 0;
#pragma unroll
    for (int k = 0; k < kChunkItems; ++k) {
        lo += a[k] & 0xFFFFFFFFull;
        hi += a[k] >> 32;
        const unsigned long long s = neff_s(a[k]);
        s1 += s;
        s2lo += (s * s) & 0xFFFFFFFFull;
        s2hi += (s * s) >> 32;
    }
    lo = warp_sum_u64(lo);
    hi = warp_sum_u64(hi);
    s1 = warp_sum_u64(s1);
    s2lo = warp_sum_u64(s2lo);
    s2hi = warp_sum_u64(s2hi);
    if (lane == 0) {
        ws.chunk_sum[((size_t)f * ws.tiles + tile) * kChunksPerTile + warp] = u128_from_limbs(lo, hi);
        s_part[warp][0] = lo;
        s_part[warp][1] = hi;
        s_part[warp][2] = s1;
        s_part[warp][3] = s2lo;
        s_part[warp][4] = s2hi;
    }
    __syncthreads();
    U128 *tab = ws.tile_prefix + (size_t)f * ws.tiles;
    unsigned long long *tneff = ws.tile_neff + (size_t)f * ws.tiles * 3;
    if (threadIdx.x == 0) {
        unsigned long long t[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k)
#pragma unroll
            for (int c = 0; c < 5; ++c) t[c] += s_part[k][c];
        const U128 tot = u128_from_limbs(t[0], t[1]);
        // (plain 64-bit stores: the reader below is ordered behind the fence and the ticket)
        reinterpret_cast<volatile unsigned long long *>(tab + tile)[0] = tot.lo;
        reinterpret_cast<volatile unsigned long long *>(tab + tile)[1] = tot.hi;
#pragma unroll
        for (int c = 0; c < 3; ++c) reinterpret_cast<volatile unsigned long long *>(tneff)[3 * tile + c] = t[2 + c];
        __threadfence();
        s_last = atomicAdd(ws.done + f, 1u) == (unsigned int)ws.tiles - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- last block of the filter: tile sums -> inclusive prefixes (a run of consecutive tiles per thread),
    // the gate's sums added up (integers: the order does not matter) ----
    const int per = (ws.tiles + kThreads - 1) / kThreads;
    const int t0 = threadIdx.x * per;
    U128 run = u128(0);
    unsigned long long n1 = 0, n2lo = 0, n2hi = 0;
    for (int k = 0; k < per; ++k)
        if (t0 + k < ws.tiles) {
            const volatile unsigned long long *p = reinterpret_cast<volatile unsigned long long *>(tab + t0 + k);
            run = run + u128(p[0], p[1]);
            const volatile unsigned long long *q = reinterpret_cast<volatile unsigned long long *>(tneff) + 3 * (t0 + k);
            n1 += q[0];
            n2lo += q[1];
            n2hi += q[2];
        }
    U128 incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const U128 n = shfl_up_u128(incl, o);
        if (lane >= o) incl = incl + n;
    }
    n1 = warp_sum_u64(n1);
    n2lo = warp_sum_u64(n2lo);
    n2hi = warp_sum_u64(n2hi);
    __shared__ U128 s_w[kThreads / 32];
    if (lane == 31) s_w[warp] = incl;
    if (lane == 0) {                         // (s_part is free: every warp passed the barrier above)
        s_part[warp][2] = n1;
        s_part[warp][3] = n2lo;
        s_part[warp][4] = n2hi;
    }
    __syncthreads();
    U128 acc = u128(0), total = u128(0);
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) {
        if (k < warp) acc = acc + s_w[k];
        total = total + s_w[k];
    }
    {   // exclusive offset of this thread's run: the warps below + the lanes below
        U128 below = shfl_up_u128(incl, 1);
        if (lane == 0) below = u128(0);
        acc = acc + below;
    }
    for (int k = 0; k < per; ++k)
        if (t0 + k < ws.tiles) {
            const volatile unsigned long long *p = reinterpret_cast<volatile unsigned long long *>(tab + t0 + k);
            acc = acc + u128(p[0], p[1]);
            tab[t0 + k] = acc;
        }
    if (threadIdx.x == 0) {
        ws.A[f] = total;
        ws.wmax[f] = m;
        ws.degenerate[f] = (bad || (total.lo == 0 && total.hi == 0)) ? 1 : 0;
        unsigned long long pay[6];
        pay[0] = total.lo;
        pay[1] = total.hi;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            unsigned long long v3 = 0;
#pragma unroll
            for (int k = 0; k < kThreads / 32; ++k) v3 += s_part[k][2 + c];
            pay[2 + c] = v3;
            ws.neff_pub[3 * f + c] = v3;
        }
        pay[5] = (unsigned long long)ws.logw_parity;
        ws.done[f] = 0u;
        if (mb.G > 1) {
            __threadfence_system();           // the tables above are visible system-wide before the flag
            mbox_post(mb, 1, pay, 6);
        }
    }
}

// stage 0 on its own (auvpf_group_resample: every rank's post is enqueued before any rank's wait)
__global__ void k_post_max(const int *wmax_f32, MailboxView mb)
{
    if (threadIdx.x || blockIdx.x) return;
    const unsigned long long mine = (unsigned long long)(unsigned int)wmax_f32[0];
    mbox_post(mb, 0, &mine, 1);
}

// linear float64 weights -> integer weights (pre-pass of auvpf_resample_weights)
__global__ void __launch_bounds__(kThreads) k_linear_to_a(const double *lin, int64_t M, const double *wmax, unsigned long long *a)
{
    const int f = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= M) return;
    const double m = wmax[f];
    a[(int64_t)f * M + i] = (isfinite(m) && m > 0.0) ? a_from_linear(lin[(int64_t)f * M + i], m) : 0ull;
}

// float64 maximum of a linear weight source (the reference's resample(weights) argument)
__global__ void __launch_bounds__(kThreads) k_weight_max_lin(const double *lin, int64_t M, long long *wmax_ordered)
{
    const int f = blockIdx.y;
    double m = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < M; i += (int64_t)gridDim.x * kThreads) {
        const double v = lin[(int64_t)f * M + i];
        if (v == v) m = fmax(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(wmax_ordered + f, double_to_ordered(m));
}
__global__ void k_finish_max_lin(long long *wmax_ordered, double *wmax, int F)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    wmax[f] = ordered_to_double(wmax_ordered[f]);
    wmax_ordered[f] = double_to_ordered(-INFINITY);
}

// ---------------------------------------------------------------------------------------------
// k_select
struct SelShard {                       // what the select kernel reads of one shard
    const float *logw[2];               // the two log-weight buffers (a shard's update writes the one its peers are not reading)
    const unsigned long long *apre;     // integer weights of a linear source (kind 1; single GPU only)
    const U128 *tile_prefix, *chunk_sum;
    StateView st[2];                    // the double-buffered state
};

struct SelArgs {
    int64_t M;                          // sources per shard and filter
    int64_t n_global;                   // N of positions = (u + i) / N
    int64_t n_src_total;                // sources of the whole filter (= n_global unless the resample changes the particle count)
    int64_t out_lo, n_out;              // this launch emits the outputs [out_lo, out_lo + n_out) of every filter
    int tiles;                          // tiles per shard and filter
    int G, rank, kind;
    int cur;                            // state buffer that is the SOURCE of the gather (same on every rank)
    int logw_parity;                    // local / records mode: which log-weight buffer is current
    SelShard local;                     // this shard (all there is when G == 1)
    const SelShard *peers;              // peer mode: [G] in device memory
    const U128 *A;                      // local mode: [F] totals
    const U128 *c_off;                  // records mode: [G + 1] mass below every shard, [G] = A
    const double *wmax;                 // [F] the maximum the a_i refer to
    const double *u;                    // [F]
    StateView dst;
    Record *records;
    int32_t *anc32;
    int64_t *anc64;
    NoiseParams rough;
    const float *rough_noise;
    const int *degenerate;              // local mode: [F]
    const unsigned long long *neff_pub; // local mode: [3F]
    double gate_thresh;                 // ratio * N; 0 = always resample
    int *skipped;                       // [F] out
    int *degenerate_out;                // peer mode: [1] out (no shard has any weight)
    U128 *coff_out;                     // peer mode: [G + 1] out: mass below every shard, [G] = the total (parity read-back)
    MailboxView mb;
};

enum SelMode : int { kSelLocal = 0, kSelPeers = 1, kSelRecords = 2 };

// the rare exact path (the estimate landed within 1e-5 of an integer): kept out of line, it is 128-bit arithmetic
// and two search loops that would otherwise be inlined ten times into the select kernel (its instruction-cache
// misses were 2.4 stall cycles per issued instruction)
__device__ __noinline__ int rel_end_exact(double u, double n, int64_t n_global, int64_t g0, int nblk, U128 C, U128 A)
{
    const SysPos spos = make_sys_pos(u, n);
    const RelEndCtx rctx = make_rel_end(u, n, g0, nblk);
    return rel_end_of(spos, n_global, rctx, cdf_exact(C, make_cdfdiv(A)));
}

template <int kMode>
__global__ void __launch_bounds__(kSelThreads, AUVPF_SEL_MINB) k_select(SelArgs a)
{
    __shared__ int s_head[kSelTile];              // output slot -> 1 + walk index of the source whose run starts there, else 0
    __shared__ int s_wmax[kSelThreads / 32];
    __shared__ int s_t[2];
    __shared__ U128 s_coff[kMaxPeers + 1];
    __shared__ SelShard s_shard[kMode == kSelPeers ? kMaxPeers : 1];
    __shared__ int s_parity[kMaxPeers];
    __shared__ int s_skip;
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t k0 = (int64_t)blockIdx.x * kSelTile;
    if (k0 >= a.n_out) return;
    const int64_t k1 = min(k0 + (int64_t)kSelTile, a.n_out) - 1;
    const int G = kMode == kSelLocal ? 1 : a.G;
#pragma unroll
    for (int r = 0; r < kSelItems; ++r) s_head[threadIdx.x * kSelItems + r] = 0;
    // launched with programmatic stream serialisation right behind k_weights: everything above overlapped its tail;
    // from here on its results are complete and visible
    cudaGridDependencySynchronize();
    // ---- shard offsets, total, gate ----
    bool skip_local = false;
    if (kMode == kSelLocal) {
        // one GPU: every thread reads the filter's scalars itself (uniform loads) -- no thread-0 prologue, and the
        // block's first barrier is the one behind the tile search
        skip_local = a.degenerate[f] != 0;
        const bool gated = !skip_local && a.gate_thresh > 0.0 &&
                           neff_fixed_a(a.neff_pub[3 * f], a.neff_pub[3 * f + 1], a.neff_pub[3 * f + 2]) >= a.gate_thresh;
        if (blockIdx.x == 0 && threadIdx.x == 0 && a.skipped) a.skipped[f] = gated ? 1 : 0;
        skip_local = skip_local || gated;
    } else {
        if (threadIdx.x == 0) {
            unsigned long long S1 = 0, S2lo = 0, S2hi = 0;
            int skip = 0;
            if (kMode == kSelPeers) {
                U128 acc = u128(0);
                for (int g = 0; g < G; ++g) {
                    const unsigned long long *slot = mbox_wait(a.mb, 1, g);
                    s_coff[g] = acc;
                    acc = acc + u128(ld_relaxed_sys(slot), ld_relaxed_sys(slot + 1));
                    S1 += ld_relaxed_sys(slot + 2);
                    S2lo += ld_relaxed_sys(slot + 3);
                    S2hi += ld_relaxed_sys(slot + 4);
                    s_parity[g] = (int)ld_relaxed_sys(slot + 5);
                }
                s_coff[G] = acc;
                if (acc.lo == 0 && acc.hi == 0) skip = 1;
                if (blockIdx.x == 0) {
                    *a.degenerate_out = skip;
                    if (a.coff_out)
                        for (int g = 0; g <= G; ++g) a.coff_out[g] = s_coff[g];
                }
                if (!skip) {
                    // N_eff gate (auv_pf.py:174-181, commented out in the reference) on the GLOBAL integer sums
                    const bool gated = a.gate_thresh > 0.0 && neff_fixed_a(S1, S2lo, S2hi) >= a.gate_thresh;
                    if (blockIdx.x == 0 && a.skipped) a.skipped[f] = gated ? 1 : 0;
                    if (gated) skip = 1;
                } else if (blockIdx.x == 0 && a.skipped)
                    a.skipped[f] = 0;
            } else {
                for (int g = 0; g <= G; ++g) s_coff[g] = a.c_off[g];
                s_parity[a.rank] = a.logw_parity;
            }
            s_skip = skip;
        }
        if (kMode == kSelPeers) {
            if (threadIdx.x < G * (int)(sizeof(SelShard) / sizeof(unsigned long long))) {
                const int g = threadIdx.x / (int)(sizeof(SelShard) / sizeof(unsigned long long));
                const int wd = threadIdx.x % (int)(sizeof(SelShard) / sizeof(unsigned long long));
                reinterpret_cast<unsigned long long *>(&s_shard[g])[wd] = reinterpret_cast<const unsigned long long *>(a.peers + g)[wd];
            }
        } else if (threadIdx.x == 0)
            s_shard[0] = a.local;
    }
    if (kMode != kSelLocal) __syncthreads();
    // (local mode reads the shard straight from the kernel parameters: the constant bank)
    auto shard = [&](int g) -> const SelShard & { return kMode == kSelLocal ? a.local : s_shard[kMode == kSelPeers ? g : 0]; };
    auto coff = [&](int g) -> U128 { return kMode == kSelLocal ? (g == 0 ? u128(0) : a.A[f]) : s_coff[g]; };
    auto parity = [&](int g) -> int { return kMode == kSelLocal ? a.logw_parity : s_parity[kMode == kSelPeers ? g : a.rank]; };
    const int my_g = kMode == kSelLocal ? 0 : a.rank;
    if (kMode == kSelLocal ? skip_local : (s_skip != 0)) {
        // no finite weight (or the N_eff gate holds the resample back): the state is carried over unchanged
        // (identity ancestors), so the buffers can flip without the host having to look at a flag first
        if (kMode != kSelRecords && (a.dst.x != nullptr || a.anc32 != nullptr || a.anc64 != nullptr) && a.n_out == a.M) {
            const StateView src = shard(my_g).st[a.cur];
            for (int64_t k = k0 + threadIdx.x; k <= k1; k += kSelThreads) {
                const int64_t row = (int64_t)f * a.M + k;
                if (a.dst.x != nullptr) {
                    a.dst.x[row] = src.x[row];
                    a.dst.y[row] = src.y[row];
                    a.dst.z[row] = src.z[row];
                    a.dst.roll[row] = src.roll[row];
                    a.dst.pitch[row] = src.pitch[row];
                    a.dst.yaw[row] = src.yaw[row];
                }
                if (a.anc32) a.anc32[row] = (int32_t)k;
                if (a.anc64) a.anc64[row] = a.out_lo + k;
            }
        }
        return;
    }
    const double u = a.u[f], n = (double)a.n_global, m = a.wmax[f];
    const int nblk = (int)(k1 - k0 + 1);
    const int64_t g0 = a.out_lo + k0;                  // global index of the block's first output
    const RelEndCtx rctx = make_rel_end(u, n, g0, nblk);
    const U128 A_total = coff(G);
    const CdfDiv div = make_cdfdiv(A_total);
    const int tiles = a.tiles, n_tiles = G * tiles;
    // run end of the source whose inclusive prefix is C (exact path only when the estimate cannot decide)
    auto rel_end = [&](U128 C) -> int {
        const int e = rel_end_fast(rctx, cdf_estimate(C, div));
        return e >= 0 ? e : rel_end_exact(u, n, a.n_global, g0, nblk, C, A_total);
    };
    // Level 1: the tiles that can hold ancestors of this block's outputs.  Global tile T = shard g = T / tiles,
    // tile t = T % tiles; its inclusive prefix is c_off[g] + tile_prefix_g[t].  Warp 0 finds the first tile whose
    // end is not certainly before the block, warp 1 the first tile whose end is certainly at or behind the
    // block's end (32-ary search on the ESTIMATE, widened by its error: a tile too many costs a pruned chunk,
    // a tile too few would be wrong).
    // (records mode emits the offspring of THIS shard's sources only: its own tiles)
    const int T_min = kMode == kSelRecords ? a.rank * tiles : 0, T_max = kMode == kSelRecords ? T_min + tiles : n_tiles;
    // (Two warps search, one barrier hands the two tiles over -- and orders the clearing of the run heads before the
    // walk.  Letting every warp run both searches itself saves that hand-over but was measured 6-14 % slower: eight
    // times the table probes.)
    if (warp < 2) {
        int lo = T_min, hi = T_max;
        while (lo < hi) {
            const int step = (hi - lo + 31) >> 5;
            const int T = lo + (lane + 1) * step - 1;
            bool before = false;
            if (T < hi) {
                const int g = kMode == kSelLocal ? 0 : T / tiles, t = T - g * tiles;
                const U128 C = coff(g) + shard(g).tile_prefix[(size_t)f * tiles + t];
                const double x = fma(cdf_estimate(C, div), rctx.n_scaled, -rctx.u) - rctx.g0d;
                before = warp == 0 ? (x <= -1.0e-5) : (x < rctx.nblk_d + 1.0e-5);
            }
            const int nb = __popc(__ballot_sync(0xffffffffu, before));     // monotone: a prefix of the lanes
            lo += nb * step;
            hi = min(lo + step - 1, hi);
        }
        if (lane == 0) s_t[warp] = min(lo, T_max - 1);
    }
    __syncthreads();
    const int T_lo = s_t[0], T_hi = s_t[1];
    // Level 2, from the SOURCE side (resampling.py:161-167 read from the other side): output slot i descends from
    // source j iff count_below(cdf[j-1]) <= i < count_below(cdf[j]).  Warp w rebuilds chunk w of every located
    // tile: 8 consecutive sources per lane, a_i from the log-weights, prefix by a thread-local running sum and a
    // warp scan on top of the chunk's base (tile prefix + the tile's chunk sums below it) -- in float64 for the
    // closed-form run ends, in exact integers only for the few sources whose estimate lands within 1e-5 of an
    // integer.  No barrier inside the walk: the warps do not depend on each other.
    const int64_t J_last = a.n_src_total - 1;          // the filter's last particle (see the note at its use)
    for (int T = T_lo; T <= T_hi; ++T) {
        const int g = kMode == kSelLocal ? 0 : T / tiles, t = T - g * tiles;
        const SelShard &S = shard(g);
        U128 cs = u128(0);
        if (lane < kChunksPerTile) cs = S.chunk_sum[((size_t)f * tiles + t) * kChunksPerTile + lane];
#pragma unroll
        for (int o = 1; o < kChunksPerTile; o <<= 1) {
            const U128 v = shfl_up_u128(cs, o);
            if (lane >= o) cs = cs + v;
        }
        U128 tile_base = coff(g);
        if (t > 0) tile_base = tile_base + S.tile_prefix[(size_t)f * tiles + t - 1];
        U128 C_before = shfl_u128(cs, max(warp - 1, 0));
        if (warp == 0) C_before = u128(0);
        C_before = tile_base + C_before;
        const U128 C_after = tile_base + shfl_u128(cs, warp);
        const int64_t jc = (int64_t)t * kScanTile + (int64_t)warp * kChunk;      // first source of the chunk, local to shard g
        if (jc >= a.M) continue;                                                  // ragged last tile of a shard
        const bool has_last = (int64_t)g * a.M + min(jc + kChunk, a.M) - 1 == J_last;
        if (C_after == C_before && !has_last) continue;                            // no mass: no run starts in this chunk
        const int e_begin = rel_end(C_before);
        if (e_begin == nblk) break;                                               // behind the block, like every later chunk
        const int e_end = has_last ? nblk : rel_end(C_after);
        if (e_end == 0) continue;                                                 // before the block
        // ---- rebuild the chunk ----
        const int64_t j0 = jc + (int64_t)lane * kChunkItems;
        uint64_t av[kChunkItems];
        {
            WeightSrc w;
            w.logw = S.logw[parity(g)];
            w.apre = S.apre;
            w.kind = a.kind;
            load_a8(w, (int64_t)f * a.M + j0, a.M - j0, m, av);
        }
        double cd[kChunkItems];
        uint64_t ci = 0;
        {
            double run = 0.0;
#pragma unroll
            for (int k = 0; k < kChunkItems; ++k) {
                run += (double)av[k];            // a_i < 2^61 is an exact float64
                cd[k] = run;
                ci += av[k];                     // 8 x 2^60 < 2^64
            }
        }
        double before_d = cd[kChunkItems - 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double v = __shfl_up_sync(0xffffffffu, before_d, o);
            if (lane >= o) before_d += v;
        }
        before_d = before_d - cd[kChunkItems - 1] + u128_to_double(C_before);
        // run ends: non-decreasing in j, so a lane whose LAST source ends at 0 lies before the block and one whose
        // FIRST source ends at nblk lies behind it; only lanes across the block evaluate all eight
        int e[kChunkItems];
        unsigned need = 0;                       // bit k: source k's estimate cannot decide
        auto fast = [&](int k) -> int {
            const int r = rel_end_fast(rctx, (before_d + cd[k]) * div.scale);
            if (r < 0) need |= 1u << k;
            return r;
        };
        const int e_lastk = fast(kChunkItems - 1);
        const int e_first = e_lastk == 0 ? 0 : fast(0);
        if ((e_lastk == 0 || e_first == nblk) && need == 0) {
#pragma unroll
            for (int k = 0; k < kChunkItems; ++k) e[k] = e_first;
        } else {
            e[0] = e_first;
#pragma unroll
            for (int k = 1; k < kChunkItems - 1; ++k) e[k] = fast(k);
            e[kChunkItems - 1] = e_lastk;
        }
        if (__any_sync(0xffffffffu, need != 0)) {
            // exact prefix of the lanes' totals (two limbs per lane), then the undecided sources exactly
            unsigned long long xl = ci & 0xFFFFFFFFull, xh = ci >> 32;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long vl = __shfl_up_sync(0xffffffffu, xl, o), vh = __shfl_up_sync(0xffffffffu, xh, o);
                if (lane >= o) {
                    xl += vl;
                    xh += vh;
                }
            }
            xl -= ci & 0xFFFFFFFFull;
            xh -= ci >> 32;
            U128 Ck = C_before + u128_from_limbs(xl, xh);
#pragma unroll
            for (int k = 0; k < kChunkItems; ++k) {
                Ck = Ck + av[k];
                if (need & (1u << k)) e[k] = rel_end_exact(u, n, a.n_global, g0, nblk, Ck, A_total);
            }
        }
        // The filter's LAST particle ends where the outputs end, whatever its CDF value says: for
        // u >= 1 - N 2^-54 the last position rounds to 1.0 = cumsum[-1], the reference walks off the end
        // (IndexError, resampling.py:162-167); every path here gives that output to the last particle.
        if (has_last) {
#pragma unroll
            for (int k = 0; k < kChunkItems; ++k)
                if ((int64_t)g * a.M + j0 + k == J_last) e[k] = nblk;
        }
        // where this lane's first run starts = where the previous source's run ends
        int start = __shfl_up_sync(0xffffffffu, e[kChunkItems - 1], 1);
        if (lane == 0) start = e_begin;
        const int wbase = (T - T_lo) * kScanTile + warp * kChunk + lane * kChunkItems;
#pragma unroll
        for (int k = 0; k < kChunkItems; ++k) {
            if (e[k] > start) s_head[start] = wbase + k + 1;       // starts are distinct: one writer per slot
            start = e[k];
        }
    }
    __syncthreads();
    // inclusive max-scan of the run heads (walk indices increase with the slot): thread-local over its
    // kSelItems adjacent outputs, warp shuffle, then the warps' maxima
    int head[kSelItems];
    {
        int mx = 0;
#pragma unroll
        for (int r = 0; r < kSelItems; ++r) {
            mx = max(mx, s_head[threadIdx.x * kSelItems + r]);
            head[r] = mx;
        }
        int incl = mx;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl = max(incl, t);
        }
        int pre = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) pre = 0;
        if (lane == 31) s_wmax[warp] = incl;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSelThreads / 32; ++k)
            if (k < warp) pre = max(pre, s_wmax[k]);
#pragma unroll
        for (int r = 0; r < kSelItems; ++r) head[r] = max(head[r], pre);
    }
    const int64_t kb = k0 + (int64_t)threadIdx.x * kSelItems;
    if (kb > k1) return;
    const int64_t orow0 = (int64_t)f * a.n_out + kb;
    const bool full4 = kb + kSelItems - 1 <= k1;
    double x[kSelItems], y[kSelItems];
    float z[kSelItems], ro[kSelItems], pi[kSelItems], ya[kSelItems];
    int64_t JJ[kSelItems], jl[kSelItems];
#pragma unroll
    for (int r = 0; r < kSelItems; ++r) {
        const int wi = max(head[r] - 1, 0);
        const int T = T_lo + wi / kScanTile;
        const int g = kMode == kSelLocal ? 0 : T / tiles;
        int64_t j = (int64_t)(T - g * tiles) * kScanTile + (wi % kScanTile);
        if (j > a.M - 1) j = a.M - 1;
        const StateView src = shard(g).st[a.cur];         // the owner's memory: local, or over NVLink
        const int64_t srow = (int64_t)f * a.M + j;
        x[r] = src.x[srow];
        y[r] = src.y[srow];
        z[r] = src.z[srow];
        ro[r] = src.roll[srow];
        pi[r] = src.pitch[srow];
        ya[r] = src.yaw[srow];
        jl[r] = j;
        JJ[r] = (int64_t)(kMode == kSelRecords ? a.rank : g) * a.M + j;
    }
    if (a.anc32) {
        if (full4 && (orow0 & 3) == 0)
            *reinterpret_cast<int4 *>(a.anc32 + orow0) = make_int4((int)jl[0], (int)jl[1], (int)jl[2], (int)jl[3]);
        else
            for (int r = 0; r < kSelItems && kb + r <= k1; ++r) a.anc32[orow0 + r] = (int32_t)jl[r];
    }
    if (a.anc64)
        for (int r = 0; r < kSelItems && kb + r <= k1; ++r) a.anc64[orow0 + r] = JJ[r];
    if (kMode == kSelRecords) {
        for (int r = 0; r < kSelItems && kb + r <= k1; ++r) {
            Record rec = {x[r], y[r], z[r], ro[r], pi[r], ya[r]};
            a.records[orow0 + r] = rec;
        }
        return;
    }
    if (a.dst.x == nullptr) return;
    if (a.rough.mode == 1) {       // fused add_noise (auv_pf.py:191-192)
#pragma unroll
        for (int r = 0; r < kSelItems; ++r) {
            if (kb + r > k1) break;
            float nn[6];
            six_normals(a.rough, (int64_t)f * a.rough.global_stride + a.out_lo + kb + r, a.rough_noise, orow0 + r, nn);
            x[r] += (double)(a.rough.sigma[0] * nn[0]);
            y[r] += (double)(a.rough.sigma[1] * nn[1]);
            z[r] += a.rough.sigma[2] * nn[2];
            ro[r] += a.rough.sigma[3] * nn[3];
            pi[r] += a.rough.sigma[4] * nn[4];
            ya[r] += a.rough.sigma[5] * nn[5];
        }
    }
    if (full4 && (orow0 & 3) == 0) {
        *reinterpret_cast<double2 *>(a.dst.x + orow0) = make_double2(x[0], x[1]);
        *reinterpret_cast<double2 *>(a.dst.x + orow0 + 2) = make_double2(x[2], x[3]);
        *reinterpret_cast<double2 *>(a.dst.y + orow0) = make_double2(y[0], y[1]);
        *reinterpret_cast<double2 *>(a.dst.y + orow0 + 2) = make_double2(y[2], y[3]);
        *reinterpret_cast<float4 *>(a.dst.z + orow0) = make_float4(z[0], z[1], z[2], z[3]);
        *reinterpret_cast<float4 *>(a.dst.roll + orow0) = make_float4(ro[0], ro[1], ro[2], ro[3]);
        *reinterpret_cast<float4 *>(a.dst.pitch + orow0) = make_float4(pi[0], pi[1], pi[2], pi[3]);
        *reinterpret_cast<float4 *>(a.dst.yaw + orow0) = make_float4(ya[0], ya[1], ya[2], ya[3]);
    } else {
        for (int r = 0; r < kSelItems && kb + r <= k1; ++r) {
            a.dst.x[orow0 + r] = x[r];
            a.dst.y[orow0 + r] = y[r];
            a.dst.z[orow0 + r] = z[r];
            a.dst.roll[orow0 + r] = ro[r];
            a.dst.pitch[orow0 + r] = pi[r];
            a.dst.yaw[orow0 + r] = ya[r];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The fixed-point weights themselves, q_j = cdf_j - cdf_{j-1} (parity read-back, residual resampler,
// auvpf_neff): one block per tile rebuilds the exact prefix and takes the exact quotient of every particle.
// c_off: mass below this shard ([0] = 0 on one GPU), total: A.  sumsq (optional): sum of (q 2^-52)^2.
__global__ void __launch_bounds__(kThreads)
k_exact_q(WeightSrc w, int64_t M, const double *wmax, const U128 *tile_prefix, const U128 *A, const U128 *c_off, int tiles,
          const int *degenerate, unsigned long long *q_out, double *sumsq)
{
    __shared__ unsigned long long s_lo[kThreads / 32], s_hi[kThreads / 32];
    const int f = blockIdx.y, tile = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t j0 = (int64_t)tile * kScanTile + (int64_t)threadIdx.x * kChunkItems;
    const bool dead = degenerate && degenerate[f];
    uint64_t a[kChunkItems];
    load_a8(w, (int64_t)f * M + j0, dead ? 0 : M - j0, wmax[f], a);
    unsigned long long lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < kChunkItems; ++k) {
        lo += a[k] & 0xFFFFFFFFull;
        hi += a[k] >> 32;
    }
    unsigned long long xl = lo, xh = hi;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long vl = __shfl_up_sync(0xffffffffu, xl, o), vh = __shfl_up_sync(0xffffffffu, xh, o);
        if (lane >= o) {
            xl += vl;
            xh += vh;
        }
    }
    if (lane == 31) {
        s_lo[warp] = xl;
        s_hi[warp] = xh;
    }
    __syncthreads();
    unsigned long long bl = xl - lo, bh = xh - hi;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k)
        if (k < warp) {
            bl += s_lo[k];
            bh += s_hi[k];
        }
    U128 C = (c_off ? c_off[f] : u128(0)) + u128_from_limbs(bl, bh);
    if (tile > 0) C = C + tile_prefix[(size_t)f * tiles + tile - 1];
    const CdfDiv div = make_cdfdiv(A[f]);
    uint64_t prev = dead ? 0 : cdf_exact(C, div);
    double sq = 0.0;
#pragma unroll
    for (int k = 0; k < kChunkItems; ++k) {
        C = C + a[k];
        const uint64_t c = dead ? 0 : cdf_exact(C, div);
        if (j0 + k < M) {
            if (q_out) q_out[(int64_t)f * M + j0 + k] = c - prev;
            const double v = (double)(c - prev) * 2.220446049250313e-16;
            sq += v * v;
        }
        prev = c;
    }
    if (sumsq) {
        sq = warp_sum_f64(sq);
        if (lane == 0 && sq != 0.0) atomicAdd(sumsq + f, sq);
    }
}

}   // namespace auvpf
