// auvpf_api.cu -- C ABI (include/auvpf.h) over the sm_100a kernels.
//
// One handle = one CUDA device, one stream, F filters x M particles.  The handle owns every
// device buffer; callers own what they pass.  No CPU fallback exists: every entry point either
// enqueues kernels on the handle's stream or fails with an error code.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>      // header-only; a no-op unless a profiler (nsys, ncu --nvtx) is attached

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/auvpf.h"
#include "auvpf_kernels.cuh"
#include "auvpf_resample.cuh"

using namespace auvpf;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? AUVPF_ERR_NOMEM : AUVPF_ERR_CUDA, "%s: %s (%s:%d)", \
                        #call, cudaGetErrorString(e_), __FILE__, __LINE__);                              \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

enum Slot { S_IN0 = 0, S_IN1, S_IN2, S_IN3, S_OUT0, S_OUT1, S_COUNT };

// parts of the slab: st[0].{x,y,z,roll,pitch,yaw}, st[1].{...}, logw[0], logw[1], cdf, status, tile_prefix, chunk_sum, mailbox
enum { P_ST0 = 0, P_ST1 = 6, P_LOGW0 = 12, P_LOGW1 = 13, P_CDF = 14, P_STATUS = 15, P_TILE = 16, P_CHUNK = 17, P_MBOX = 18 };
constexpr int kSlabParts = 19;

struct ProfEvent {
    cudaEvent_t a, b;
    int klass;
};

}   // namespace

struct auvpf_handle {
    auvpf_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t M = 0, N = 0;   // per filter, total rows
    int64_t cap_M = 0, cap_N = 0;      // what the buffers were sized for at creation (M can shrink and grow below it)
    int64_t next_count = 0;            // auvpf_set_next_particle_count: the next systematic resample draws this many
    int F = 1;
    int64_t global_offset = 0, global_count = 0;

    // state (double-buffered SoA), the two log-weight buffers, the resampler's chunk / tile tables and the
    // mailbox live in ONE allocation (the slab) so that a single CUDA IPC handle exposes them to the other
    // ranks of a sharded filter
    void *slab = nullptr;
    size_t slab_bytes = 0;
    int64_t slab_off[kSlabParts] = {0};    // see alloc_slab
    void *peer_slab[8] = {nullptr};        // peers' slabs: opened through IPC, or in-process handles' ([rank] == own slab)
    bool peer_ipc[8] = {false};            // opened with cudaIpcOpenMemHandle (to be closed)
    int n_peers = 0, my_peer = 0;
    SelShard *peer_shards = nullptr;            // [n_peers] device: what the select kernel reads of every shard
    MailboxView mbox{};                         // peers' mailboxes (G = 1 until peers are attached)
    unsigned long long *mailbox = nullptr;      // this rank's mailbox (inside the slab)
    unsigned int epoch = 0;                     // resamples of the sharded filter so far
    unsigned int est_epoch = 0;                 // ... and estimates exchanged through the mailbox (auvpf_shard_pose_sums)
    double *est_all = nullptr;                  // [kMaxPeers][16] device: every rank's pose sums
    int *shard_degenerate = nullptr;            // [1] device: no shard had any finite weight (written by k_select)
    unsigned long long *shard_totals = nullptr; // [6] device: A lo, A hi, S1, S2 lo, S2 hi, log-weight buffer (staged variant)
    U128 *c_off = nullptr;                      // [kMaxPeers + 1] device: mass below every shard, [G] = the global total
    bool have_global = false;                   // c_off describes the last sharded resample
    int n_ranks_global = 1, my_rank_global = 0;
    StateView st[2];
    int cur = 0;
    float *logw_buf[2] = {nullptr, nullptr};    // a shard's update writes the buffer its peers are not reading
    int lw = 0;                                 // the current log-weight buffer
    float *logw = nullptr;                      // == logw_buf[lw]
    unsigned long long *cdf = nullptr;          // [N] scratch: the weights q_j (residual resampler), flag scans (slot map)
    U128 *tile_prefix = nullptr, *chunk_sum = nullptr;

    // per-filter scalars
    long long *wmax_ordered = nullptr;
    int *wmax_f32 = nullptr;      // [F] max log-weight published by the update kernels (ordered int)
    bool wmax_valid = false;      // wmax_f32 describes the current logw
    unsigned int *done = nullptr; // [F] ticket counters of k_weights
    double *wmax = nullptr;
    U128 *A = nullptr;
    unsigned long long *deficit = nullptr;     // [F] zero: kept for the residual kernels' q_i + (i < D) form
    double *u = nullptr, *sumsq = nullptr, *sums = nullptr, *gps = nullptr;
    PredictParams *pparams = nullptr;
    int *degenerate = nullptr;
    unsigned long long *neff_acc = nullptr;    // [3 F tiles] per-tile partial sums of the N_eff gate (scratch of k_weights)
    unsigned long long *neff_pub = nullptr;    // [3F] integer sums of the N_eff gate (S1, S2 lo, S2 hi)
    int *skipped = nullptr;                  // [F] 1 = the N_eff gate held the last resample back
    double neff_ratio = 0.0;                 // resample only when N_eff < neff_ratio * N; 0 = always
    bool sumsq_lazy = false;                 // sumsq is computed on request
    DevBuf replay_params, replay_ranges, replay_u, replay_stats;   // auvpf_replay: a whole recorded sequence

    // scan machinery
    unsigned long long *status = nullptr;
    unsigned int *tile_counter = nullptr;
    int tiles_per_filter = 0;

    // resampling extras
    int32_t *anc = nullptr;
    int *first = nullptr;
    unsigned long long *scan2 = nullptr, *totals2 = nullptr;
    unsigned long long *scan3 = nullptr;   // residual resampling: scan of the fractional parts
    int32_t *lost = nullptr, *dupes = nullptr;
    unsigned long long *a_pre = nullptr;   // [N] integer weights of the last LINEAR weight source (auvpf_resample_weights)
    int last_src_kind = 0;
    bool have_norm = false;

    // map / sensor
    float *dem_z = nullptr, *mip = nullptr, *mip_hi = nullptr, *beams = nullptr, *meas = nullptr, *slope_blk = nullptr;
    float *beam_aux = nullptr;   // [3B]: |sin th|, 1/|sin th|, 1/cos th
    double *beam_tan = nullptr;  // [2B]: tan |th|, 1 / cos th (profile merge)
    int *flagged = nullptr, *flagged2 = nullptr;   // [N] each, per filter: particles for the map-border pass / for the per-ray kernel
    DevBuf probe_flags;
    int *beam_order = nullptr;   // [32 * chunk]
    FanParams fan{};
    DemView dem{};
    bool have_dem = false;
    bool use_profile = true;   // AUVPF_NO_PROFILE=1 in the environment forces the per-ray kernel
    bool use_pdl = true;       // AUVPF_NO_PDL=1: launch the select kernel without programmatic stream serialisation
    int B = 0;
    FrameParams frame{};

    // step counters for the Philox streams
    uint32_t step_predict = 0, step_rough = 0, step_resample = 0;

    DevBuf slots[S_COUNT];
    void *pinned = nullptr;   // small pinned staging ring
    size_t pinned_cap = 1 << 16, pinned_off = 0;

    // profiling
    bool prof = false;
    std::vector<ProfEvent> prof_events;
    std::vector<ProfEvent> prof_pool;
    double prof_ms[AUVPF_K_COUNT] = {0};
    int64_t prof_launches[AUVPF_K_COUNT] = {0};
};

namespace {

struct Guard {   // make the handle's device current for the duration of a call
    int prev = -1;
    explicit Guard(auvpf_handle *h)
    {
        cudaGetDevice(&prev);
        if (prev != h->device) cudaSetDevice(h->device);
    }
    ~Guard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

const char *const kClassName[AUVPF_K_COUNT] = {"auvpf:propagate", "auvpf:update_mbes", "auvpf:update_gps", "auvpf:weights",
                                               "auvpf:scan",      "auvpf:select_gather", "auvpf:roughen",  "auvpf:estimate",
                                               "auvpf:other"};

struct Prof {   // NVTX range (always) + CUDA-event bracket (profiling mode) around the kernels of one class
    auvpf_handle *h;
    ProfEvent ev{};
    bool on;
    Prof(auvpf_handle *h_, int klass) : h(h_), on(h_->prof)
    {
        nvtxRangePushA(kClassName[klass]);
        if (!on) return;
        if (!h->prof_pool.empty()) {
            ev = h->prof_pool.back();
            h->prof_pool.pop_back();
        } else {
            cudaEventCreate(&ev.a);
            cudaEventCreate(&ev.b);
        }
        ev.klass = klass;
        cudaEventRecord(ev.a, h->stream);
    }
    ~Prof()
    {
        if (on) {
            cudaEventRecord(ev.b, h->stream);
            h->prof_events.push_back(ev);
        }
        nvtxRangePop();
    }
};

#define LAUNCHED(h, klass, n) ((h)->prof_launches[klass] += (n))

bool is_device_ptr(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int slot_reserve(auvpf_handle *h, int slot, size_t bytes, void **out)
{
    DevBuf &b = h->slots[slot];
    if (b.cap < bytes) {
        if (b.p) {
            CU(cudaStreamSynchronize(h->stream));
            CU(cudaFree(b.p));
            b.p = nullptr;
            b.cap = 0;
        }
        size_t cap = bytes + bytes / 4 + 256;
        CU(cudaMalloc(&b.p, cap));
        b.cap = cap;
    }
    *out = b.p;
    return AUVPF_OK;
}

// device view of an input that may live on the host
template <class T>
int stage_in(auvpf_handle *h, int slot, const T *src, size_t count, const T **dev)
{
    if (is_device_ptr(src)) {
        *dev = src;
        return AUVPF_OK;
    }
    void *d = nullptr;
    int rc = slot_reserve(h, slot, count * sizeof(T), &d);
    if (rc) return rc;
    CU(cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    *dev = static_cast<const T *>(d);
    return AUVPF_OK;
}

// device buffer a kernel writes an output into; finish_out copies back when dst is host memory
template <class T>
int stage_out(auvpf_handle *h, int slot, T *dst, size_t count, T **dev)
{
    if (is_device_ptr(dst)) {
        *dev = dst;
        return AUVPF_OK;
    }
    void *d = nullptr;
    int rc = slot_reserve(h, slot, count * sizeof(T), &d);
    if (rc) return rc;
    *dev = static_cast<T *>(d);
    return AUVPF_OK;
}

template <class T>
int finish_out(auvpf_handle *h, T *dst, const T *dev, size_t count)
{
    if (static_cast<const void *>(dst) == static_cast<const void *>(dev)) return AUVPF_OK;
    CU(cudaMemcpyAsync(dst, dev, count * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return AUVPF_OK;
}

// tiny host->device upload through the pinned ring (fully asynchronous)
int tiny_upload(auvpf_handle *h, void *dst_dev, const void *src_host, size_t bytes)
{
    if (bytes > h->pinned_cap / 4) {
        CU(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, h->stream));
        return AUVPF_OK;
    }
    size_t off = (h->pinned_off + 15) & ~size_t(15);
    if (off + bytes > h->pinned_cap) {
        CU(cudaStreamSynchronize(h->stream));   // wrap: make sure earlier copies have drained
        off = 0;
    }
    memcpy(static_cast<char *>(h->pinned) + off, src_host, bytes);
    CU(cudaMemcpyAsync(dst_dev, static_cast<char *>(h->pinned) + off, bytes, cudaMemcpyHostToDevice, h->stream));
    h->pinned_off = off + bytes;
    return AUVPF_OK;
}

inline dim3 grid1(int64_t n, int per_block, int F) { return dim3((unsigned)((n + per_block - 1) / per_block), (unsigned)F); }

// tf.transformations.euler_from_quaternion(q)[0:2] with 'sxyz' (auv_particle.py:50-56): host, float64.
void roll_pitch_from_quat(const double q_in[4], double &roll, double &pitch)
{
    double q[4] = {q_in[0], q_in[1], q_in[2], q_in[3]};
    double nq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const double eps = 2.220446049250313e-16 * 4.0;
    double Mm[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (nq >= eps) {
        double s = std::sqrt(2.0 / nq);
        for (double &c : q) c *= s;
        double o[4][4];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) o[i][j] = q[i] * q[j];
        Mm[0][0] = 1.0 - o[1][1] - o[2][2];
        Mm[1][0] = o[0][1] + o[2][3];
        Mm[2][0] = o[0][2] - o[1][3];
        Mm[1][1] = 1.0 - o[0][0] - o[2][2];
        Mm[1][2] = o[1][2] - o[0][3];
        Mm[2][1] = o[1][2] + o[0][3];
        Mm[2][2] = 1.0 - o[0][0] - o[1][1];
    }
    double cy = std::sqrt(Mm[0][0] * Mm[0][0] + Mm[1][0] * Mm[1][0]);
    if (cy > eps) {
        roll = std::atan2(Mm[2][1], Mm[2][2]);
        pitch = std::atan2(-Mm[2][0], cy);
    } else {
        roll = std::atan2(-Mm[1][2], Mm[1][1]);
        pitch = std::atan2(-Mm[2][0], cy);
    }
}

PredictParams make_predict(const double *od, const double motion_cov[6])
{
    // od = [vx vy vz wz qx qy qz qw z dt]
    double roll, pitch;
    roll_pitch_from_quat(od + 4, roll, pitch);
    const double dt = od[9];
    const double vx = od[0] * dt, vy = od[1] * dt, vz = od[2] * dt;
    const double cr = std::cos(roll), sr = std::sin(roll), cp = std::cos(pitch), sp = std::sin(pitch);
    PredictParams p;
    p.a = (float)(cp * vx + sp * sr * vy + sp * cr * vz);
    p.b = (float)(cr * vy - sr * vz);
    p.wzdt = (float)(od[3] * dt);
    p.roll = (float)roll;
    p.pitch = (float)pitch;
    p.z = (float)od[8];
    p.sx = (float)std::sqrt(motion_cov[0]);
    p.sy = (float)std::sqrt(motion_cov[1]);
    p.syaw = (float)std::sqrt(motion_cov[5]);
    p.pad = 0.f;
    return p;
}

NoiseParams make_noise(auvpf_handle *h, const double var[6], const double *mean, int mode, uint32_t step,
                       uint32_t stream)
{
    NoiseParams n{};
    for (int c = 0; c < 6; ++c) {
        n.sigma[c] = var ? (float)std::sqrt(var[c]) : 0.f;
        n.mean[c] = mean ? mean[c] : 0.0;
    }
    n.mode = mode;
    n.need = 0;
    for (int c = 0; c < 6; ++c) n.need |= n.sigma[c] != 0.f ? (1u << c) : 0u;
    n.seed = h->cfg.seed;
    n.step = step;
    n.stream = stream;
    n.global_offset = h->global_offset;
    n.global_stride = h->global_count;
    return n;
}

StateView state_at(void *slab, const int64_t *off)
{
    char *b = static_cast<char *>(slab);
    StateView s;
    s.x = reinterpret_cast<double *>(b + off[0]);
    s.y = reinterpret_cast<double *>(b + off[1]);
    s.z = reinterpret_cast<float *>(b + off[2]);
    s.roll = reinterpret_cast<float *>(b + off[3]);
    s.pitch = reinterpret_cast<float *>(b + off[4]);
    s.yaw = reinterpret_cast<float *>(b + off[5]);
    return s;
}

int alloc_slab(auvpf_handle *h)
{
    const size_t n = (size_t)h->N, tiles = (size_t)h->tiles_per_filter * h->F;
    size_t sizes[kSlabParts];
    for (int b = 0; b < 2; ++b) {
        sizes[P_ST0 + 6 * b + 0] = sizes[P_ST0 + 6 * b + 1] = 8 * n;
        for (int c = 2; c < 6; ++c) sizes[P_ST0 + 6 * b + c] = 4 * n;
    }
    sizes[P_LOGW0] = sizes[P_LOGW1] = 4 * n;
    sizes[P_CDF] = 8 * n;
    sizes[P_STATUS] = 8 * tiles;
    sizes[P_TILE] = sizeof(U128) * tiles;
    sizes[P_CHUNK] = sizeof(U128) * tiles * kChunksPerTile;
    sizes[P_MBOX] = kMboxBytes;
    size_t off = 0;
    for (int k = 0; k < kSlabParts; ++k) {
        h->slab_off[k] = (int64_t)off;
        off += (sizes[k] + 255) & ~size_t(255);
    }
    h->slab_bytes = off;
    CU(cudaMalloc(&h->slab, off));
    char *base = static_cast<char *>(h->slab);
    h->st[0] = state_at(h->slab, h->slab_off + P_ST0);
    h->st[1] = state_at(h->slab, h->slab_off + P_ST1);
    h->logw_buf[0] = reinterpret_cast<float *>(base + h->slab_off[P_LOGW0]);
    h->logw_buf[1] = reinterpret_cast<float *>(base + h->slab_off[P_LOGW1]);
    h->logw = h->logw_buf[h->lw];
    h->cdf = reinterpret_cast<unsigned long long *>(base + h->slab_off[P_CDF]);
    h->status = reinterpret_cast<unsigned long long *>(base + h->slab_off[P_STATUS]);
    h->tile_prefix = reinterpret_cast<U128 *>(base + h->slab_off[P_TILE]);
    h->chunk_sum = reinterpret_cast<U128 *>(base + h->slab_off[P_CHUNK]);
    h->mailbox = reinterpret_cast<unsigned long long *>(base + h->slab_off[P_MBOX]);
    CU(cudaMemsetAsync(h->mailbox, 0, kMboxBytes, h->stream));
    h->mbox.G = 1;
    h->mbox.rank = 0;
    h->mbox.epoch = 0;
    h->mbox.peer[0] = h->mailbox;
    return AUVPF_OK;
}

// what the select kernel reads of a shard whose slab is mapped at `base` with the part offsets `off`
SelShard shard_at(void *base, const int64_t *off)
{
    char *c = static_cast<char *>(base);
    SelShard s{};
    s.logw[0] = reinterpret_cast<const float *>(c + off[P_LOGW0]);
    s.logw[1] = reinterpret_cast<const float *>(c + off[P_LOGW1]);
    s.apre = nullptr;
    s.tile_prefix = reinterpret_cast<const U128 *>(c + off[P_TILE]);
    s.chunk_sum = reinterpret_cast<const U128 *>(c + off[P_CHUNK]);
    s.st[0] = state_at(base, off + P_ST0);
    s.st[1] = state_at(base, off + P_ST1);
    return s;
}

// A producer of log-weights (update, set_logw) of a SHARDED filter writes the buffer its peers are not
// reading: a fast rank may already be in its next update while a slow one still rebuilds prefix sums from
// this rank's previous log-weights (k_select reads them over NVLink).  The buffer alternates per producer
// call; the one a resample used is named in the mailbox.  One GPU: a single buffer.
void flip_logw(auvpf_handle *h)
{
    if (h->n_peers > 1 || h->global_count != h->M) {
        h->lw ^= 1;
        h->logw = h->logw_buf[h->lw];
    }
}

// ---- normalisation on the handle's stream -----------------------------------------------------
// float32 log-weights that did not come from an update kernel (auvpf_set_logw): max pass
int ensure_wmax_f32(auvpf_handle *h)
{
    if (h->wmax_valid) return AUVPF_OK;
    Prof pr(h, AUVPF_K_NORMALISE);
    CU(cudaMemsetAsync(h->wmax_f32, 0x80, sizeof(int) * h->F, h->stream));
    int blocks = (int)std::min<int64_t>((h->M + kThreads - 1) / kThreads, 1184);
    k_weight_max_f32<<<dim3(blocks, h->F), kThreads, 0, h->stream>>>(h->logw, h->M, h->wmax_f32);
    LAUNCHED(h, AUVPF_K_NORMALISE, 1);
    CU(cudaGetLastError());
    h->wmax_valid = true;
    return AUVPF_OK;
}

// k_weights: a_i, chunk / tile sums, tile prefix, total, gate sums.  use_mailbox: a sharded filter whose
// peers are attached exchanges the maximum (and afterwards its totals) through the mailbox inside the
// kernel; otherwise the maximum in wmax_f32 is taken as it is (local, or already reduced by the caller).
int run_weights(auvpf_handle *h, const WeightSrc &src, bool use_mailbox, bool post_max = true)
{
    int rc;
    WeightState ws{};
    if (src.kind == 0) {
        if ((rc = ensure_wmax_f32(h))) return rc;
        ws.wmax_f32 = h->wmax_f32;
    } else
        ws.wmax_f32 = nullptr;      // linear weights: run_linear_prepass left their float64 maximum in h->wmax
    ws.wmax = h->wmax;
    ws.done = h->done;
    ws.tile_neff = h->neff_acc;
    ws.neff_pub = h->neff_pub;
    ws.A = h->A;
    ws.degenerate = h->degenerate;
    ws.tile_prefix = h->tile_prefix;
    ws.chunk_sum = h->chunk_sum;
    ws.tiles = h->tiles_per_filter;
    ws.logw_parity = h->lw;
    ws.post_max = post_max ? 1 : 0;
    MailboxView mb = h->mbox;
    if (!use_mailbox) mb.G = 1;
    mb.epoch = h->epoch;
    Prof pr(h, AUVPF_K_NORMALISE);
    k_weights<<<dim3((unsigned)h->tiles_per_filter, (unsigned)h->F), kThreads, 0, h->stream>>>(src, h->M, ws, mb);
    LAUNCHED(h, AUVPF_K_NORMALISE, 1);
    CU(cudaGetLastError());
    h->have_norm = true;
    h->last_src_kind = src.kind;
    h->sumsq_lazy = true;
    return AUVPF_OK;
}

// the weights q_j = cdf_j - cdf_{j-1} of the last k_weights into `q_out` (and / or their sum of squares)
int run_exact_q(auvpf_handle *h, unsigned long long *q_out, double *sumsq, const U128 *A, const U128 *c_off)
{
    WeightSrc src{h->logw, h->a_pre, h->last_src_kind};
    if (sumsq) CU(cudaMemsetAsync(sumsq, 0, sizeof(double) * h->F, h->stream));
    k_exact_q<<<dim3((unsigned)h->tiles_per_filter, (unsigned)h->F), kThreads, 0, h->stream>>>(
        src, h->M, h->wmax, h->tile_prefix, A, c_off, h->tiles_per_filter, h->degenerate, q_out, sumsq);
    LAUNCHED(h, AUVPF_K_OTHER, 1);
    CU(cudaGetLastError());
    return AUVPF_OK;
}

// launch right behind k_weights with programmatic stream serialisation (the prologue of the select kernel
// overlaps the tail of the weights kernel; cudaGridDependencySynchronize inside orders the rest)
template <int kMode>
cudaError_t launch_select(auvpf_handle *h, dim3 grid, const SelArgs &sa)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kSelThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = h->use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k_select<kMode>, sa);
}

SelArgs select_args(auvpf_handle *h, const WeightSrc &src)
{
    SelArgs a{};
    a.M = h->M;
    a.n_global = h->global_count;
    a.n_src_total = h->global_count;
    a.out_lo = h->global_offset;
    a.n_out = h->M;
    a.tiles = h->tiles_per_filter;
    a.G = 1;
    a.rank = 0;
    a.kind = src.kind;
    a.cur = h->cur;
    a.logw_parity = h->lw;
    a.local = shard_at(h->slab, h->slab_off);
    a.local.apre = src.apre;
    a.peers = nullptr;
    a.A = h->A;
    a.c_off = nullptr;
    a.wmax = h->wmax;
    a.u = h->u;
    a.degenerate = h->degenerate;
    a.neff_pub = h->neff_pub;
    a.gate_thresh = h->neff_ratio * (double)h->global_count;
    a.skipped = h->skipped;
    a.degenerate_out = h->shard_degenerate;
    a.mb = h->mbox;
    a.mb.epoch = h->epoch;
    return a;
}

int check_degenerate(auvpf_handle *h)
{
    std::vector<int> deg(h->F);
    CU(cudaMemcpyAsync(deg.data(), h->degenerate, sizeof(int) * h->F, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int f = 0; f < h->F; ++f)
        if (deg[f]) return fail(AUVPF_ERR_DEGENERATE, "filter %d: no finite positive weight; resample skipped", f);
    return AUVPF_OK;
}

int ensure_inplace_buffers(auvpf_handle *h)
{
    if (h->first) return AUVPF_OK;
    CU(cudaMalloc(&h->first, sizeof(int) * h->cap_N));
    CU(cudaMalloc(&h->scan2, sizeof(unsigned long long) * h->cap_N));
    CU(cudaMalloc(&h->totals2, sizeof(unsigned long long) * 2 * h->F));
    CU(cudaMalloc(&h->lost, sizeof(int32_t) * h->cap_N));
    CU(cudaMalloc(&h->dupes, sizeof(int32_t) * h->cap_N));
    return AUVPF_OK;
}

// reference slot map on device-resident indices (auv_pf.py:183-198)
int apply_indices_dev(auvpf_handle *h, const int32_t *idx)
{
    int rc = ensure_inplace_buffers(h);
    if (rc) return rc;
    Prof pr(h, AUVPF_K_SELECT_GATHER);
    const dim3 g = grid1(h->M, kThreads, h->F);
    CU(cudaMemsetAsync(h->first, 0x7f, sizeof(int) * h->N, h->stream));
    k_first_occurrence<<<g, kThreads, 0, h->stream>>>(idx, h->M, h->first);
    const size_t st_bytes = sizeof(unsigned long long) * (size_t)h->tiles_per_filter * h->F;
    CU(cudaMemsetAsync(h->status, 0, st_bytes, h->stream));
    CU(cudaMemsetAsync(h->tile_counter, 0, sizeof(unsigned int), h->stream));
    k_scan_lost<<<h->tiles_per_filter * h->F, kThreads, 0, h->stream>>>(h->first, h->M, h->cdf, h->status,
                                                                        h->tile_counter, h->tiles_per_filter,
                                                                        h->totals2);
    CU(cudaMemsetAsync(h->status, 0, st_bytes, h->stream));
    CU(cudaMemsetAsync(h->tile_counter, 0, sizeof(unsigned int), h->stream));
    k_scan_dupes<<<h->tiles_per_filter * h->F, kThreads, 0, h->stream>>>(h->first, idx, h->M, h->scan2, h->status,
                                                                         h->tile_counter, h->tiles_per_filter,
                                                                         h->totals2 + h->F);
    k_compact_lost_dupes<<<g, kThreads, 0, h->stream>>>(h->first, idx, h->M, h->cdf, h->scan2, h->lost, h->dupes);
    k_reassign<<<g, kThreads, 0, h->stream>>>(h->st[h->cur], h->M, h->lost, h->dupes, h->totals2, h->totals2 + h->F);
    LAUNCHED(h, AUVPF_K_SELECT_GATHER, 6);
    CU(cudaGetLastError());
    return AUVPF_OK;
}

int reset_scan_state(auvpf_handle *h)
{
    CU(cudaMemsetAsync(h->status, 0, sizeof(unsigned long long) * (size_t)h->tiles_per_filter * h->F, h->stream));
    CU(cudaMemsetAsync(h->tile_counter, 0, sizeof(unsigned int), h->stream));
    return AUVPF_OK;
}

// N_eff gate for the paths that synchronise anyway (residual resampling, reference slot map; one
// filter per handle): the same integers and the same expression as the select kernel's gate.
// Call after check_degenerate (the stream is idle).  *skip = the resample is held back.
int host_neff_gate(auvpf_handle *h, bool *skip)
{
    *skip = false;
    if (!(h->neff_ratio > 0.0)) {
        CU(cudaMemsetAsync(h->skipped, 0, sizeof(int) * h->F, h->stream));
        return AUVPF_OK;
    }
    unsigned long long S[3];
    CU(cudaMemcpyAsync(S, h->neff_pub, sizeof(S), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *skip = neff_fixed_a(S[0], S[1], S[2]) >= h->neff_ratio * (double)h->M;
    const int flag = *skip ? 1 : 0;
    return tiny_upload(h, h->skipped, &flag, sizeof(int));
}

int emit_identity_ancestors(auvpf_handle *h, int32_t *ancestors_out)
{
    if (!ancestors_out) return AUVPF_OK;
    if (!h->anc) CU(cudaMalloc(&h->anc, sizeof(int32_t) * h->cap_N));
    int32_t *dev = is_device_ptr(ancestors_out) ? ancestors_out : h->anc;
    k_identity_indices<<<grid1(h->M, kThreads, h->F), kThreads, 0, h->stream>>>(dev, h->M);
    CU(cudaGetLastError());
    if (dev != ancestors_out) return finish_out(h, ancestors_out, (const int32_t *)dev, (size_t)h->N);
    return AUVPF_OK;
}

// residual_resample (resampling.py:27-76) + the reference's application of its indices.
// u: (F, M) float64 uniforms -- the random(N - k) of :74, first M - k of each row consumed -- or
// nullptr for Philox.
int do_resample_residual(auvpf_handle *h, const WeightSrc &src, int32_t flags, const double *u,
                         int32_t *ancestors_out, const double *rough_var, const float *rough_noise)
{
    if (h->global_count != h->M)
        return fail(AUVPF_ERR_BADARG, "residual resampling is not available for a particle-sharded filter");
    const bool strict = (flags & AUVPF_RS_STRICT_REF) != 0;
    const bool inplace = (flags & AUVPF_RS_INPLACE_REF) != 0;
    if (strict && h->M > 65536)
        return fail(AUVPF_ERR_BADARG, "AUVPF_RS_STRICT_REF emulates the reference's sequential draw: particle_count <= 65536");
    if (h->neff_ratio > 0.0 && h->F > 1)
        return fail(AUVPF_ERR_BADARG, "the N_eff gate with residual resampling needs one filter per handle");
    int rc;
    if ((rc = run_weights(h, src, /*use_mailbox=*/false))) return rc;
    if ((rc = check_degenerate(h))) return rc;
    {
        bool skip;
        if ((rc = host_neff_gate(h, &skip))) return rc;
        if (skip) {     // state and weights untouched; the draws count as consumed, like on the systematic path
            h->step_resample++;
            if (rough_var) h->step_rough++;
            return emit_identity_ancestors(h, ancestors_out);
        }
    }
    // the weights themselves, q_j = cdf_j - cdf_{j-1} (sum exactly 2^52: no deficit units to hand out)
    {
        Prof pr(h, AUVPF_K_SCAN);
        if ((rc = run_exact_q(h, h->cdf, nullptr, h->A, nullptr))) return rc;
    }
    if ((rc = ensure_inplace_buffers(h))) return rc;
    if (!h->scan3) CU(cudaMalloc(&h->scan3, sizeof(unsigned long long) * h->cap_N));
    if (!h->anc) CU(cudaMalloc(&h->anc, sizeof(int32_t) * h->cap_N));
    const double *ud = nullptr;
    if (u && (rc = stage_in(h, S_IN3, u, (size_t)h->N, &ud))) return rc;
    const float *rn = nullptr;
    if (rough_noise && (rc = stage_in(h, S_IN1, rough_noise, (size_t)h->N * 6, &rn))) return rc;
    const uint32_t step = h->step_resample++;
    unsigned long long *K = h->totals2, *R = h->totals2 + h->F;
    const dim3 g = grid1(h->M, kThreads, h->F);
    {
        Prof pr(h, AUVPF_K_SCAN);
        if ((rc = reset_scan_state(h))) return rc;
        k_scan_copies<<<h->tiles_per_filter * h->F, kThreads, 0, h->stream>>>(
            h->cdf, h->M, h->deficit, h->global_offset, h->scan2, h->status, h->tile_counter, h->tiles_per_filter, K);
        LAUNCHED(h, AUVPF_K_SCAN, 1);
        if (!strict) {
            if ((rc = reset_scan_state(h))) return rc;
            k_scan_fracs<<<h->tiles_per_filter * h->F, kThreads, 0, h->stream>>>(
                h->cdf, h->M, h->deficit, h->global_offset, h->scan3, h->status, h->tile_counter, h->tiles_per_filter,
                R);
            LAUNCHED(h, AUVPF_K_SCAN, 1);
        }
        CU(cudaGetLastError());
    }
    {
        Prof pr(h, AUVPF_K_SELECT_GATHER);
        k_residual_fill<<<g, kThreads, 0, h->stream>>>(h->scan2, h->scan3, h->M, K, R, ud, h->cfg.seed, step,
                                                      strict ? 0 : 1, h->anc);
        LAUNCHED(h, AUVPF_K_SELECT_GATHER, 1);
        if (strict) {
            k_residual_strict<<<h->F, kThreads, 0, h->stream>>>(h->cdf, h->M, h->deficit, h->global_offset, K,
                                                                reinterpret_cast<double *>(h->scan3), ud, h->cfg.seed,
                                                                step, h->anc);
            LAUNCHED(h, AUVPF_K_SELECT_GATHER, 1);
        }
        CU(cudaGetLastError());
    }
    if (ancestors_out) {
        if (is_device_ptr(ancestors_out))
            CU(cudaMemcpyAsync(ancestors_out, h->anc, sizeof(int32_t) * h->N, cudaMemcpyDeviceToDevice, h->stream));
        else if ((rc = finish_out(h, ancestors_out, (const int32_t *)h->anc, (size_t)h->N)))
            return rc;
    }
    if (inplace) {                       // the reference: keep / lost / dupes, then add_noise on every particle
        if ((rc = apply_indices_dev(h, h->anc))) return rc;
        if (rough_var && (rc = auvpf_roughen(h, rough_var, rough_noise))) return rc;
    } else {
        NoiseParams rough = make_noise(h, rough_var, nullptr, rough_var ? 1 : 0, h->step_rough, kStreamRoughen);
        Prof pr(h, AUVPF_K_SELECT_GATHER);
        k_gather_indices<<<g, kThreads, 0, h->stream>>>(h->anc, h->M, h->st[h->cur], h->st[1 - h->cur], rough, rn);
        LAUNCHED(h, AUVPF_K_SELECT_GATHER, 1);
        CU(cudaGetLastError());
        h->cur = 1 - h->cur;
        if (rough_var) h->step_rough++;
    }
    return AUVPF_OK;
}

// everything that is derived from the particle count (one filter per handle, not sharded: auvpf_set_next_particle_count)
void set_count(auvpf_handle *h, int64_t n)
{
    h->M = h->N = h->global_count = n;
    h->tiles_per_filter = (int)((n + kScanTile - 1) / kScanTile);
    h->have_norm = false;
    h->wmax_valid = false;
}

// The systematic resample that CHANGES the particle count (KLD adaptation): n_new positions (u + i) / n_new
// over the CDF of the M current particles.  Decided on the host, because a held-back or degenerate resample must leave
// the count alone: one read-back, acceptable for a call that is made when the cloud has changed size, not every cycle.
int do_resample_to_count(auvpf_handle *h, const WeightSrc &src, int64_t n_new, int32_t *ancestors_out, const double *rough_var)
{
    int rc;
    if ((rc = run_weights(h, src, /*use_mailbox=*/false))) return rc;
    if ((rc = check_degenerate(h))) return rc;                   // (error: state, weights and count untouched)
    bool skip;
    if ((rc = host_neff_gate(h, &skip))) return rc;
    if (skip) {                                                  // held back: the draws count as consumed, the count stays
        if (rough_var) h->step_rough++;
        return emit_identity_ancestors(h, ancestors_out);
    }
    if (!h->anc) CU(cudaMalloc(&h->anc, sizeof(int32_t) * h->cap_N));
    int32_t *anc_dev = ancestors_out ? ((is_device_ptr(ancestors_out)) ? ancestors_out : h->anc) : nullptr;
    SelArgs sa = select_args(h, src);
    sa.n_global = n_new;                    // positions
    sa.n_out = n_new;                       // ... all of them emitted here
    sa.gate_thresh = 0.0;                   // (decided above)
    sa.rough = make_noise(h, rough_var, nullptr, rough_var ? 1 : 0, h->step_rough, kStreamRoughen);
    sa.anc32 = anc_dev;
    sa.dst = h->st[1 - h->cur];
    {
        Prof pr(h, AUVPF_K_SELECT_GATHER);
        CU(launch_select<kSelLocal>(h, grid1(n_new, kSelTile, 1), sa));
        LAUNCHED(h, AUVPF_K_SELECT_GATHER, 1);
        CU(cudaGetLastError());
    }
    h->cur = 1 - h->cur;
    if (rough_var) h->step_rough++;
    if (ancestors_out && anc_dev != ancestors_out && (rc = finish_out(h, ancestors_out, (const int32_t *)anc_dev, (size_t)n_new))) return rc;
    CU(cudaStreamSynchronize(h->stream));
    set_count(h, n_new);
    h->next_count = 0;
    return AUVPF_OK;
}

int do_resample(auvpf_handle *h, const WeightSrc &src, int32_t mode, int32_t flags, const double *u,
                int32_t *ancestors_out, const double *rough_var, const float *rough_noise)
{
    if (mode == AUVPF_RESIDUAL) return do_resample_residual(h, src, flags, u, ancestors_out, rough_var, rough_noise);
    if (mode != AUVPF_SYSTEMATIC) return fail(AUVPF_ERR_BADARG, "unknown resample mode %d", mode);
    if (h->neff_ratio > 0.0 && h->F > 1 && (flags & AUVPF_RS_INPLACE_REF))
        return fail(AUVPF_ERR_BADARG, "the N_eff gate with AUVPF_RS_INPLACE_REF needs one filter per handle");
    if (h->global_count != h->M)
        return fail(AUVPF_ERR_BADARG, "a particle-sharded filter resamples through auvpf_shard_resample / the staged auvpf_shard_* calls");
    int rc;
    // uniforms: resampling.py:157 random() (uploaded first, so that the two kernels are adjacent in the stream)
    std::vector<double> uh(h->F);
    if (u) {
        if (is_device_ptr(u)) CU(cudaMemcpyAsync(h->u, u, sizeof(double) * h->F, cudaMemcpyDeviceToDevice, h->stream));
        else if ((rc = tiny_upload(h, h->u, u, sizeof(double) * h->F))) return rc;
    } else {
        for (int f = 0; f < h->F; ++f)
            uh[f] = uniform53(philox_for(h->cfg.seed, f, h->step_resample, kStreamResample, 0));
        if ((rc = tiny_upload(h, h->u, uh.data(), sizeof(double) * h->F))) return rc;
    }
    h->step_resample++;

    const bool inplace = (flags & AUVPF_RS_INPLACE_REF) != 0;
    if (h->next_count > 0 && h->next_count != h->M) {
        if (inplace || rough_noise)
            return fail(AUVPF_ERR_BADARG, "a resample that changes the particle count is systematic, gathers into the other buffer and draws its own roughening noise");
        return do_resample_to_count(h, src, h->next_count, ancestors_out, rough_var);
    }
    h->next_count = 0;
    const float *rn = nullptr;
    if (rough_noise && (rc = stage_in(h, S_IN1, rough_noise, (size_t)h->N * 6, &rn))) return rc;
    if ((rc = run_weights(h, src, /*use_mailbox=*/false))) return rc;
    int32_t *anc_dev = nullptr;
    if (ancestors_out || inplace) {
        if (!h->anc) CU(cudaMalloc(&h->anc, sizeof(int32_t) * h->cap_N));
        anc_dev = h->anc;
        if (ancestors_out && is_device_ptr(ancestors_out) && !inplace) anc_dev = ancestors_out;
    }
    SelArgs sa = select_args(h, src);
    sa.rough = make_noise(h, rough_var, nullptr, (rough_var && !inplace) ? 1 : 0, h->step_rough, kStreamRoughen);
    sa.rough_noise = rn;
    sa.anc32 = anc_dev;
    sa.dst = h->st[1 - h->cur];
    if (inplace) sa.dst.x = nullptr;   // ancestors only
    {
        Prof pr(h, AUVPF_K_SELECT_GATHER);
        CU(launch_select<kSelLocal>(h, grid1(h->M, kSelTile, h->F), sa));
        LAUNCHED(h, AUVPF_K_SELECT_GATHER, 1);
        CU(cudaGetLastError());
    }
    // A filter without any finite weight made the kernel above carry its state over unchanged, so the
    // buffers flip either way; the flag is read back (one 4-byte copy at the END of the enqueued cycle,
    // the only host synchronisation of a resample) unless the caller asked not to wait.
    if (!inplace) {
        h->cur = 1 - h->cur;
        if (rough_var) h->step_rough++;
    }
    if (!(flags & AUVPF_RS_ASYNC) || inplace || ancestors_out) {
        if ((rc = check_degenerate(h))) return rc;
    }
    if (ancestors_out && anc_dev != ancestors_out) {
        if (is_device_ptr(ancestors_out))
            CU(cudaMemcpyAsync(ancestors_out, anc_dev, sizeof(int32_t) * h->N, cudaMemcpyDeviceToDevice, h->stream));
        else if ((rc = finish_out(h, ancestors_out, (const int32_t *)anc_dev, (size_t)h->N)))
            return rc;
    }
    if (inplace) {
        if (h->neff_ratio > 0.0) {      // the kernel's decision (identity ancestors were written): nothing to apply
            int skipped = 0;
            CU(cudaMemcpyAsync(&skipped, h->skipped, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            if (skipped) {
                if (rough_var) h->step_rough++;      // (auvpf_roughen below would have advanced it)
                return AUVPF_OK;
            }
        }
        if ((rc = apply_indices_dev(h, anc_dev))) return rc;
        if (rough_var && (rc = auvpf_roughen(h, rough_var, rough_noise))) return rc;
    }
    return AUVPF_OK;
}

}   // namespace

// ================================================================================================
extern "C" {

int auvpf_version(void) { return AUVPF_VERSION; }

const char *auvpf_last_error(void) { return g_err.c_str(); }

int auvpf_create(const auvpf_config *cfg, auvpf_handle **out)
{
    if (!cfg || !out) return fail(AUVPF_ERR_BADARG, "null argument");
    if (cfg->struct_size != (int32_t)sizeof(auvpf_config))
        return fail(AUVPF_ERR_BADARG, "auvpf_config.struct_size %d != %zu", cfg->struct_size, sizeof(auvpf_config));
    if (cfg->particle_count <= 0 || cfg->particle_count > (1ll << 30))
        return fail(AUVPF_ERR_BADARG, "particle_count %lld out of range", (long long)cfg->particle_count);
    if (cfg->global_count < 0 || cfg->global_count > (1ll << 30))
        return fail(AUVPF_ERR_BADARG, "global_count %lld out of range (a filter holds at most 2^30 particles over all shards)", (long long)cfg->global_count);
    const int F = cfg->n_filters <= 0 ? 1 : cfg->n_filters;
    if (!(cfg->meas_std > 0.0)) return fail(AUVPF_ERR_BADARG, "measurement_std must be > 0");
    if (!(cfg->range_max > 0.0)) return fail(AUVPF_ERR_BADARG, "range_max must be > 0");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(AUVPF_ERR_BADARG, "device %d of %d", cfg->device, ndev);

    auvpf_handle *h = new (std::nothrow) auvpf_handle();
    if (!h) return fail(AUVPF_ERR_NOMEM, "host allocation failed");
    h->cfg = *cfg;
    h->device = cfg->device;
    h->M = cfg->particle_count;
    h->F = F;
    h->N = h->M * F;
    h->cap_M = h->M;
    h->cap_N = h->N;
    h->global_offset = cfg->global_offset;
    h->global_count = cfg->global_count > 0 ? cfg->global_count : h->M;
    Guard g(h);
    int rc = [&]() -> int {
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        int rc2;
        h->tiles_per_filter = (int)((h->M + kScanTile - 1) / kScanTile);
        if ((rc2 = alloc_slab(h))) return rc2;
        CU(cudaMalloc(&h->wmax_ordered, sizeof(long long) * F));
        {
            std::vector<long long> init((size_t)F, (long long)0xFFF0000000000000ull ^ 0x7FFFFFFFFFFFFFFFll);   // ordered(-inf)
            CU(cudaMemcpy(h->wmax_ordered, init.data(), sizeof(long long) * F, cudaMemcpyHostToDevice));
        }
        CU(cudaMalloc(&h->wmax, sizeof(double) * F));
        CU(cudaMalloc(&h->wmax_f32, sizeof(int) * 3 * F));   // [F] max log-weight | [F] + [F] counters of the two particle lists
        CU(cudaMalloc(&h->done, sizeof(unsigned int) * F));
        CU(cudaMemsetAsync(h->done, 0, sizeof(unsigned int) * F, h->stream));
        CU(cudaMalloc(&h->A, sizeof(U128) * F));
        CU(cudaMemsetAsync(h->A, 0, sizeof(U128) * F, h->stream));
        CU(cudaMalloc(&h->deficit, sizeof(unsigned long long) * F));
        CU(cudaMemsetAsync(h->deficit, 0, sizeof(unsigned long long) * F, h->stream));
        CU(cudaMalloc(&h->u, sizeof(double) * F));
        CU(cudaMalloc(&h->sumsq, sizeof(double) * F));
        CU(cudaMalloc(&h->sums, sizeof(double) * 16 * F));
        CU(cudaMalloc(&h->gps, sizeof(double) * 2 * F));
        CU(cudaMalloc(&h->pparams, sizeof(PredictParams) * F * AUVPF_MAX_PREDICT_STEPS));
        CU(cudaMalloc(&h->degenerate, sizeof(int) * F));
        CU(cudaMemsetAsync(h->degenerate, 0, sizeof(int) * F, h->stream));
        CU(cudaMalloc(&h->neff_acc, sizeof(unsigned long long) * 3 * F * (size_t)h->tiles_per_filter));
        CU(cudaMalloc(&h->neff_pub, sizeof(unsigned long long) * 3 * F));
        CU(cudaMemsetAsync(h->neff_pub, 0, sizeof(unsigned long long) * 3 * F, h->stream));
        CU(cudaMalloc(&h->shard_degenerate, sizeof(int)));
        CU(cudaMemsetAsync(h->shard_degenerate, 0, sizeof(int), h->stream));
        CU(cudaMalloc(&h->shard_totals, sizeof(unsigned long long) * 6));
        CU(cudaMalloc(&h->c_off, sizeof(U128) * (kMaxPeers + 1)));
        CU(cudaMalloc(&h->skipped, sizeof(int) * F));
        CU(cudaMemsetAsync(h->skipped, 0, sizeof(int) * F, h->stream));
        CU(cudaMalloc(&h->tile_counter, sizeof(unsigned int)));
        CU(cudaMallocHost(&h->pinned, h->pinned_cap));
        for (int b = 0; b < 2; ++b) {
            CU(cudaMemsetAsync(h->st[b].x, 0, sizeof(double) * h->N, h->stream));
            CU(cudaMemsetAsync(h->st[b].y, 0, sizeof(double) * h->N, h->stream));
            CU(cudaMemsetAsync(h->st[b].z, 0, sizeof(float) * h->N, h->stream));
            CU(cudaMemsetAsync(h->st[b].roll, 0, sizeof(float) * h->N, h->stream));
            CU(cudaMemsetAsync(h->st[b].pitch, 0, sizeof(float) * h->N, h->stream));
            CU(cudaMemsetAsync(h->st[b].yaw, 0, sizeof(float) * h->N, h->stream));
        }
        CU(cudaMemsetAsync(h->logw_buf[0], 0, sizeof(float) * h->N, h->stream));
        CU(cudaMemsetAsync(h->logw_buf[1], 0, sizeof(float) * h->N, h->stream));
        return AUVPF_OK;
    }();
    if (rc) {
        auvpf_destroy(h);
        return rc;
    }
    // map <- odom transform (auv_pf.py:79-81)
    bool ident = true;
    for (int a = 0; a < 3; ++a) {
        h->frame.m2o_t[a] = cfg->m2o[4 * a + 3];
        for (int b = 0; b < 3; ++b) {
            double v = cfg->m2o[4 * a + b];
            h->frame.m2o_r[3 * a + b] = v;
            if (v != (a == b ? 1.0 : 0.0)) ident = false;
        }
    }
    h->frame.m2o_identity_r = ident ? 1 : 0;
    {
        const char *np = getenv("AUVPF_NO_PROFILE");
        h->use_profile = !(np && np[0] == '1');
        const char *nd = getenv("AUVPF_NO_PDL");
        h->use_pdl = !(nd && nd[0] == '1');
    }
    *out = h;
    return AUVPF_OK;
}

int auvpf_destroy(auvpf_handle *h)
{
    if (!h) return AUVPF_OK;
    Guard g(h);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int r = 0; r < h->n_peers; ++r)
        if (r != h->my_peer && h->peer_slab[r] && h->peer_ipc[r]) cudaIpcCloseMemHandle(h->peer_slab[r]);
    cudaFree(h->slab);
    cudaFree(h->peer_shards);
    void *bufs[] = {h->wmax_ordered, h->wmax,    h->A,      h->neff_acc, h->neff_pub, h->shard_degenerate, h->shard_totals, h->c_off,
                    h->a_pre, h->deficit, h->u,     h->sumsq,        h->sums,    h->gps,    h->pparams, h->degenerate,
                    h->tile_counter, h->anc,   h->first,   h->scan2,  h->totals2, h->lost,
                    h->wmax_f32, h->done, h->scan3, h->skipped,
                    h->dupes,   h->dem_z, h->mip,          h->beams,   h->meas,   h->mip_hi, h->slope_blk, h->est_all, h->beam_aux, h->beam_order, h->beam_tan, h->flagged, h->flagged2, h->probe_flags.p,
                    h->replay_params.p, h->replay_ranges.p, h->replay_u.p, h->replay_stats.p};
    for (void *b : bufs) cudaFree(b);
    for (auto &s : h->slots) cudaFree(s.p);
    if (h->pinned) cudaFreeHost(h->pinned);
    for (auto &e : h->prof_events) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    for (auto &e : h->prof_pool) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
    return AUVPF_OK;
}

int auvpf_sync(auvpf_handle *h)
{
    if (!h) return fail(AUVPF_ERR_BADARG, "null handle");
    Guard g(h);
    CU(cudaStreamSynchronize(h->stream));
    return AUVPF_OK;
}

void *auvpf_stream(auvpf_handle *h) { return h ? (void *)h->stream : nullptr; }

int64_t auvpf_particle_count(auvpf_handle *h) { return h ? h->N : 0; }

// ------------------------------------------------------------------------------------------------
int auvpf_set_dem(auvpf_handle *h, const float *z, int32_t H, int32_t W, double x0, double y0, double res)
{
    if (!h || !z) return fail(AUVPF_ERR_BADARG, "null argument");
    if (H < 2 || W < 2 || !(res > 0.0)) return fail(AUVPF_ERR_BADARG, "DEM must be at least 2x2 with res > 0");
    Guard g(h);
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(h->dem_z);
    cudaFree(h->mip);
    cudaFree(h->mip_hi);
    h->dem_z = h->mip = h->mip_hi = nullptr;
    h->have_dem = false;
    const size_t n = (size_t)H * W;
    const int pitch = W + 2;
    const size_t n_pad = (size_t)(H + 2) * pitch;
    const float *src = nullptr;
    {
        int rc0 = stage_in(h, S_IN0, z, n, &src);
        if (rc0) return rc0;
    }
    CU(cudaMalloc(&h->dem_z, n_pad * sizeof(float)));
    {
        dim3 tb(32, 8), gr((W + 2 + tb.x - 1) / tb.x, (H + 2 + tb.y - 1) / tb.y);
        k_pad_dem<<<gr, tb, 0, h->stream>>>(src, H, W, h->dem_z);
        CU(cudaGetLastError());
    }
    const float *z00 = h->dem_z + pitch + 1;   // vertex (0, 0) inside the padding
    // max-mip levels used for skipping: a finer one (blocks of about 2 m) and a coarser one
    // (about 16 m) unless the caller chose: coarse_level = lo | (hi << 8), hi == 0 -> one level
    int lo, hi;
    if (h->cfg.coarse_level < 0) {
        lo = (int)std::lround(std::log2(2.0 / res));
        if (lo < 1) lo = 1;
        hi = lo + 3;
    } else {
        lo = h->cfg.coarse_level & 0xff;
        hi = (h->cfg.coarse_level >> 8) & 0xff;
    }
    if (lo < 1) lo = 1;
    if (lo > 10) lo = 10;
    if (hi > 12) hi = 12;
    if (hi <= lo) hi = 0;
    const int top = hi ? hi : lo;
    float *prev = nullptr, *lvl_lo = nullptr, *lvl_hi = nullptr;
    int pH = 0, pW = 0, loH = 0, loW = 0, hiH = 0, hiW = 0;
    dim3 tb(32, 8);
    for (int L = 1; L <= top; ++L) {
        int S = 1 << L;
        int oH = (H - 1 + S - 1) / S, oW = (W - 1 + S - 1) / S;
        if (oH < 1) oH = 1;
        if (oW < 1) oW = 1;
        float *lvl;
        CU(cudaMalloc(&lvl, (size_t)oH * oW * sizeof(float)));
        dim3 gr((oW + tb.x - 1) / tb.x, (oH + tb.y - 1) / tb.y);
        if (L == 1) k_mip_level1<<<gr, tb, 0, h->stream>>>(z00, pitch, H, W, lvl, oH, oW);
        else k_mip_down<<<gr, tb, 0, h->stream>>>(prev, pH, pW, lvl, oH, oW);
        CU(cudaGetLastError());
        if (prev && prev != lvl_lo) {
            CU(cudaStreamSynchronize(h->stream));
            cudaFree(prev);
        }
        prev = lvl;
        pH = oH;
        pW = oW;
        if (L == lo) { lvl_lo = lvl; loH = oH; loW = oW; }
        if (L == hi) { lvl_hi = lvl; hiH = oH; hiW = oW; }
    }
    h->mip = lvl_lo;
    h->mip_hi = lvl_hi;
    // global maximum height
    int *d_max;
    void *tmp;
    int rc = slot_reserve(h, S_OUT0, sizeof(int), &tmp);
    if (rc) return rc;
    d_max = static_cast<int *>(tmp);
    int init = 0x80000000;   // ordered encoding below every float
    CU(cudaMemcpyAsync(d_max, &init, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    k_max_f32<<<296, 256, 0, h->stream>>>(h->dem_z, (int64_t)n_pad, d_max);
    int mx;
    CU(cudaMemcpyAsync(&mx, d_max, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    mx = mx >= 0 ? mx : mx ^ 0x7FFFFFFF;
    float zmax;
    memcpy(&zmax, &mx, sizeof zmax);
    // Lipschitz constant of the surface (validity of the fan sweep)
    init = 0;
    CU(cudaMemcpyAsync(d_max, &init, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    k_max_slope<<<296, 256, 0, h->stream>>>(z00, pitch, H, W, (float)(1.0 / res), d_max);
    CU(cudaMemcpyAsync(&mx, d_max, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float max_slope;
    memcpy(&max_slope, &mx, sizeof max_slope);
    // ... and per block of 128 x 128 cells: what certifies a profile march (local_max_slope)
    const int sbW = std::max(1, (W - 1 + kSlopeBlock - 1) / kSlopeBlock), sbH = std::max(1, (H - 1 + kSlopeBlock - 1) / kSlopeBlock);
    if (h->slope_blk) cudaFree(h->slope_blk);
    h->slope_blk = nullptr;
    CU(cudaMalloc(&h->slope_blk, sizeof(float) * (size_t)sbW * sbH));
    k_slope_blocks<<<dim3((unsigned)sbW, (unsigned)sbH), 256, 0, h->stream>>>(z00, pitch, H, W, (float)(1.0 / res), h->slope_blk, sbW);
    CU(cudaGetLastError());

    DemView &d = h->dem;
    d.z = z00;
    d.W = W;
    d.last_cx = W - 2;
    d.last_cy = H - 2;
    d.H = H;
    d.pitch = pitch;
    d.mip[0] = lvl_lo;
    d.mipW[0] = loW;
    d.mipH[0] = loH;
    d.shift[0] = lo;
    d.mip[1] = lvl_hi ? lvl_hi : lvl_lo;
    d.mipW[1] = lvl_hi ? hiW : loW;
    d.mipH[1] = lvl_hi ? hiH : loH;
    d.shift[1] = lvl_hi ? hi : lo;
    d.n_levels = lvl_hi ? 2 : 1;
    d.range_max = (float)h->cfg.range_max;
    d.zmax = zmax;
    d.max_slope = max_slope * 1.0001f + 1e-6f;
    d.slope_blk = h->slope_blk;
    d.sbW = sbW;
    d.sbH = sbH;
    d.res = (float)res;
    d.x0 = x0;
    d.y0 = y0;
    d.inv_res_d = 1.0 / res;
    h->have_dem = true;

    // Keep the DEM resident in L2 across kernels where the part allows it (best effort).
    int persist_max = 0, win_max = 0;
    cudaDeviceGetAttribute(&persist_max, cudaDevAttrMaxPersistingL2CacheSize, h->device);
    cudaDeviceGetAttribute(&win_max, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
    if (persist_max > 0 && win_max > 0) {
        size_t bytes = std::min(n_pad * sizeof(float), (size_t)win_max);
        size_t carve = std::min(bytes, (size_t)persist_max);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
        cudaStreamAttrValue at{};
        at.accessPolicyWindow.base_ptr = h->dem_z;
        at.accessPolicyWindow.num_bytes = bytes;
        at.accessPolicyWindow.hitRatio = bytes ? (float)std::min(1.0, (double)carve / (double)bytes) : 0.f;
        at.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        at.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &at);
        cudaGetLastError();
    }
    return AUVPF_OK;
}

int auvpf_rasterise_mesh(auvpf_handle *h, const float *vertices, int64_t n_vertices, const int32_t *faces,
                         int64_t n_faces, int32_t H, int32_t W, double x0, double y0, double res, float fill,
                         float *z_out)
{
    if (!h || !vertices || !faces || !z_out) return fail(AUVPF_ERR_BADARG, "null argument");
    if (H < 2 || W < 2 || !(res > 0.0) || n_vertices < 3 || n_faces < 1)
        return fail(AUVPF_ERR_BADARG, "mesh needs >= 3 vertices and >= 1 face, grid at least 2x2 with res > 0");
    Guard g(h);
    int rc;
    const float *vd;
    const int32_t *fd;
    float *zd;
    if ((rc = stage_in(h, S_IN0, vertices, (size_t)n_vertices * 3, &vd))) return rc;
    if ((rc = stage_in(h, S_IN1, faces, (size_t)n_faces * 3, &fd))) return rc;
    const int64_t n = (int64_t)H * W;
    if ((rc = stage_out(h, S_OUT0, z_out, (size_t)n, &zd))) return rc;
    int *zi = reinterpret_cast<int *>(zd);
    k_fill_ordered<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(zi, n);
    k_rasterise_mesh<<<(unsigned)((n_faces + 127) / 128), 128, 0, h->stream>>>(vd, fd, n_faces, H, W, x0, y0, 1.0 / res, zi);
    k_unorder<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(zi, n, fill);
    LAUNCHED(h, AUVPF_K_OTHER, 3);
    CU(cudaGetLastError());
    return finish_out(h, z_out, (const float *)zd, (size_t)n);
}

int auvpf_set_beams(auvpf_handle *h, const float *dirs_xyz, int32_t B)
{
    if (!h || !dirs_xyz || B <= 0 || B > 4096) return fail(AUVPF_ERR_BADARG, "bad beam table (1..4096 beams)");
    Guard g(h);
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(h->beams);
    cudaFree(h->meas);
    cudaFree(h->beam_aux);
    cudaFree(h->beam_order);
    cudaFree(h->beam_tan);
    h->beams = h->meas = h->beam_aux = nullptr;
    h->beam_order = nullptr;
    h->beam_tan = nullptr;
    h->B = 0;
    std::vector<float> host(3 * (size_t)B);
    CU(cudaMemcpy(host.data(), dirs_xyz, sizeof(float) * 3 * B,
                  is_device_ptr(dirs_xyz) ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost));
    // host-side analysis: planar sorted fan? per-lane sweep order (auvpf_fan.h)
    FanPlan plan = plan_fan(host.data(), B);
    std::vector<float> aux(3 * (size_t)B);
    for (int b = 0; b < B; ++b) {
        aux[b] = plan.s[b];
        aux[B + b] = plan.inv_s[b];
        aux[2 * B + b] = (float)plan.sec_t[b];
    }
    CU(cudaMalloc(&h->beams, sizeof(float) * 3 * B));
    CU(cudaMalloc(&h->meas, sizeof(float) * (size_t)B * h->F));
    CU(cudaMalloc(&h->beam_aux, sizeof(float) * 3 * B));
    CU(cudaMalloc(&h->beam_tan, sizeof(double) * 2 * B));
    CU(cudaMemcpy(h->beam_tan, plan.tan_t.data(), sizeof(double) * B, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->beam_tan + B, plan.sec_t.data(), sizeof(double) * B, cudaMemcpyHostToDevice));
    if (!h->flagged) CU(cudaMalloc(&h->flagged, sizeof(int) * (size_t)h->cap_N));
    if (!h->flagged2) CU(cudaMalloc(&h->flagged2, sizeof(int) * (size_t)h->cap_N));
    CU(cudaMalloc(&h->beam_order, sizeof(int) * plan.order.size()));
    CU(cudaMemcpy(h->beams, host.data(), sizeof(float) * 3 * B, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->beam_aux, aux.data(), sizeof(float) * 3 * B, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->beam_order, plan.order.data(), sizeof(int) * plan.order.size(), cudaMemcpyHostToDevice));
    h->fan.planar = plan.planar ? 1 : 0;
    h->fan.chunk = plan.chunk;
    h->fan.lanes_first = plan.lanes_first;
    h->fan.n1 = plan.n1;
    h->fan.sgn1 = plan.sgn1;
    for (int a = 0; a < 3; ++a) h->fan.a0[a] = plan.a0[a];
    for (int a = 0; a < 3; ++a) h->fan.c0[a] = plan.c0[a];
    h->B = B;
    return AUVPF_OK;
}

// ------------------------------------------------------------------------------------------------
int auvpf_init_particles(auvpf_handle *h, const double mean[6], const double var[6], const float *noise)
{
    if (!h || !mean || !var) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    const float *nd = nullptr;
    int rc;
    if (noise && (rc = stage_in(h, S_IN0, noise, (size_t)h->N * 6, &nd))) return rc;
    NoiseParams p = make_noise(h, var, mean, 2, 0, kStreamInit);
    Prof pr(h, AUVPF_K_OTHER);
    k_add_noise<<<grid1(h->M, kThreads, h->F), kThreads, 0, h->stream>>>(h->st[h->cur], h->M, p, nd);
    LAUNCHED(h, AUVPF_K_OTHER, 1);
    CU(cudaGetLastError());
    return AUVPF_OK;
}

int auvpf_roughen(auvpf_handle *h, const double var[6], const float *noise)
{
    if (!h || !var) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    const float *nd = nullptr;
    int rc;
    if (noise && (rc = stage_in(h, S_IN0, noise, (size_t)h->N * 6, &nd))) return rc;
    NoiseParams p = make_noise(h, var, nullptr, 1, h->step_rough++, kStreamRoughen);
    Prof pr(h, AUVPF_K_ROUGHEN);
    k_add_noise<<<grid1(h->M, kThreads, h->F), kThreads, 0, h->stream>>>(h->st[h->cur], h->M, p, nd);
    LAUNCHED(h, AUVPF_K_ROUGHEN, 1);
    CU(cudaGetLastError());
    return AUVPF_OK;
}

int auvpf_set_poses(auvpf_handle *h, const double *poses)
{
    if (!h || !poses) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    const double *pd;
    int rc = stage_in(h, S_IN0, poses, (size_t)h->N * 6, &pd);
    if (rc) return rc;
    k_set_poses<<<grid1(h->N, kThreads, 1), kThreads, 0, h->stream>>>(pd, h->N, h->st[h->cur]);
    LAUNCHED(h, AUVPF_K_OTHER, 1);
    CU(cudaGetLastError());
    return AUVPF_OK;
}

int auvpf_get_poses(auvpf_handle *h, double *poses)
{
    if (!h || !poses) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    double *pd;
    int rc = stage_out(h, S_OUT0, poses, (size_t)h->N * 6, &pd);
    if (rc) return rc;
    k_get_poses<<<grid1(h->N, kThreads, 1), kThreads, 0, h->stream>>>(h->st[h->cur], h->N, pd);
    LAUNCHED(h, AUVPF_K_OTHER, 1);
    CU(cudaGetLastError());
    return finish_out(h, poses, (const double *)pd, (size_t)h->N * 6);
}

int auvpf_get_logw(auvpf_handle *h, float *logw)
{
    if (!h || !logw) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    CU(cudaMemcpyAsync(logw, h->logw, sizeof(float) * h->N,
                       is_device_ptr(logw) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return AUVPF_OK;
}

int auvpf_set_logw(auvpf_handle *h, const float *logw)
{
    if (!h || !logw) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    flip_logw(h);
    CU(cudaMemcpyAsync(h->logw, logw, sizeof(float) * h->N,
                       is_device_ptr(logw) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    h->wmax_valid = false;
    return AUVPF_OK;
}

// ------------------------------------------------------------------------------------------------
int auvpf_predict_many(auvpf_handle *h, const double *odom, int32_t n_steps, int32_t n_odom, const float *noise)
{
    if (!h || !odom) return fail(AUVPF_ERR_BADARG, "null argument");
    if (n_steps < 1 || n_steps > AUVPF_MAX_PREDICT_STEPS)
        return fail(AUVPF_ERR_BADARG, "n_steps must be 1..%d", AUVPF_MAX_PREDICT_STEPS);
    if (n_odom != 1 && n_odom != h->F) return fail(AUVPF_ERR_BADARG, "n_odom must be 1 or n_filters");
    if (is_device_ptr(odom)) return fail(AUVPF_ERR_BADARG, "odom must be host memory (11 scalars per filter and step)");
    Guard g(h);
    int rc;
    const float *nd = nullptr;
    if (noise && (rc = stage_in(h, S_IN0, noise, (size_t)h->N * 6 * n_steps, &nd))) return rc;
    PredictParams p0 = make_predict(odom, h->cfg.motion_cov);
    const PredictParams *pp = nullptr;
    if (n_odom > 1 || n_steps > 1) {
        std::vector<PredictParams> all((size_t)h->F * n_steps);
        for (int t = 0; t < n_steps; ++t)
            for (int f = 0; f < h->F; ++f)
                all[(size_t)t * h->F + f] =
                    make_predict(odom + 11 * ((size_t)t * n_odom + (n_odom > 1 ? f : 0)), h->cfg.motion_cov);
        if ((rc = tiny_upload(h, h->pparams, all.data(), sizeof(PredictParams) * all.size()))) return rc;
        pp = h->pparams;
    }
    Prof pr(h, AUVPF_K_PROPAGATE);
    k_propagate<<<grid1(h->M, kThreads * 4, h->F), kThreads, 0, h->stream>>>(
        h->st[h->cur], h->M, p0, pp, n_steps, h->cfg.seed, h->step_predict, h->global_offset, h->global_count, nd);
    h->step_predict += (uint32_t)n_steps;
    LAUNCHED(h, AUVPF_K_PROPAGATE, 1);
    CU(cudaGetLastError());
    return AUVPF_OK;
}

int auvpf_predict(auvpf_handle *h, const double *odom, int32_t n_odom, const float *noise)
{
    return auvpf_predict_many(h, odom, 1, n_odom, noise);
}

int auvpf_update_gps(auvpf_handle *h, const double *gps_xy, float *logw_out)
{
    if (!h || !gps_xy) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    int rc;
    if (is_device_ptr(gps_xy)) CU(cudaMemcpyAsync(h->gps, gps_xy, sizeof(double) * 2 * h->F, cudaMemcpyDeviceToDevice, h->stream));
    else if ((rc = tiny_upload(h, h->gps, gps_xy, sizeof(double) * 2 * h->F))) return rc;
    const double var = h->cfg.meas_std * h->cfg.meas_std;
    flip_logw(h);
    CU(cudaMemsetAsync(h->wmax_f32, 0x80, sizeof(int) * h->F, h->stream));
    h->wmax_valid = true;
    {
        Prof pr(h, AUVPF_K_UPDATE_GPS);
        k_update_gps<<<grid1(h->M, kThreads, h->F), kThreads, 0, h->stream>>>(
            h->st[h->cur], h->M, h->frame, h->gps, (float)(0.5 / var), (float)(-std::log(2.0 * kPi * var)), h->logw,
            h->wmax_f32);
        LAUNCHED(h, AUVPF_K_UPDATE_GPS, 1);
        CU(cudaGetLastError());
    }
    if (logw_out) return auvpf_get_logw(h, logw_out);
    return AUVPF_OK;
}

int auvpf_update_mbes(auvpf_handle *h, const float *ranges, float *logw_out)
{
    if (!h || !ranges) return fail(AUVPF_ERR_BADARG, "null argument");
    if (!h->have_dem || h->B <= 0) return fail(AUVPF_ERR_STATE, "update_mbes needs set_dem and set_beams first");
    Guard g(h);
    CU(cudaMemcpyAsync(h->meas, ranges, sizeof(float) * (size_t)h->B * h->F,
                       is_device_ptr(ranges) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    const double var = h->cfg.meas_std * h->cfg.meas_std;
    flip_logw(h);
    // one fill resets the maxima AND the two list counters (all count from kWmaxReset)
    CU(cudaMemsetAsync(h->wmax_f32, 0x80, sizeof(int) * 3 * h->F, h->stream));
    h->wmax_valid = true;
    {
        Prof pr(h, AUVPF_K_UPDATE_MBES);
        const float nhiv = (float)(-0.5 / var), lnb = (float)std::log(h->cfg.meas_std * std::sqrt(2.0 * kPi));
        const int *list = nullptr;
        unsigned int *count1 = reinterpret_cast<unsigned int *>(h->wmax_f32 + h->F);        // sensors whose reach leaves the map
        unsigned int *count2 = reinterpret_cast<unsigned int *>(h->wmax_f32 + 2 * h->F);    // not certified at all: ray by ray
        const unsigned int *count = count2;
        int64_t ray_blocks = (h->M + kMbesWarps - 1) / kMbesWarps;
        if (h->fan.planar && h->use_profile) {
            // profile merge, one thread per particle.  Pass 1: every particle, the lean interior variant; a sensor
            // closer to the border than its reach goes on list 1 ...
            const size_t smp = (2 * sizeof(double) + sizeof(BeamAux)) * h->B + 2 * sizeof(BeamAux);
            if (h->N <= kProfSmallLimit)
                k_update_mbes_profile<kProfThreadsSmall, kProfMinBlocksSmall, false>
                    <<<grid1(h->M, kProfThreadsSmall, h->F), kProfThreadsSmall, smp, h->stream>>>(
                        h->st[h->cur], h->M, h->frame, h->dem, h->fan, h->beam_tan, h->beam_aux, h->meas, h->B,
                        (float)h->cfg.range_min, nhiv, lnb, h->logw, nullptr, nullptr, h->flagged, count1, h->wmax_f32);
            else
                k_update_mbes_profile<kProfThreads, kProfMinBlocks, false>
                    <<<grid1(h->M, kProfThreads, h->F), kProfThreads, smp, h->stream>>>(
                        h->st[h->cur], h->M, h->frame, h->dem, h->fan, h->beam_tan, h->beam_aux, h->meas, h->B,
                        (float)h->cfg.range_min, nhiv, lnb, h->logw, nullptr, nullptr, h->flagged, count1, h->wmax_f32);
            // ... pass 2: list 1 with the map-border variant (a side whose curve runs off the map is finished ray by
            // ray); a SMALL grid strides over the list (usually empty: its blocks leave at once).  What even that
            // cannot certify (sensor outside the map, centre beam missed, a facet steeper than the fan tilt allows)
            // goes on list 2 ...
            const int blocks2 = (int)std::min<int64_t>((h->M + kProfThreadsBorder - 1) / kProfThreadsBorder,
                                                       std::max<int64_t>(1, (148 * kProfMinBlocksBorder) / h->F));
            k_update_mbes_profile<kProfThreadsBorder, kProfMinBlocksBorder, true>
                <<<dim3((unsigned)blocks2, (unsigned)h->F), kProfThreadsBorder, smp, h->stream>>>(
                    h->st[h->cur], h->M, h->frame, h->dem, h->fan, h->beam_tan, h->beam_aux, h->meas, h->B,
                    (float)h->cfg.range_min, nhiv, lnb, h->logw, h->flagged, count1, h->flagged2, count2, h->wmax_f32);
            LAUNCHED(h, AUVPF_K_UPDATE_MBES, 2);
            list = h->flagged2;
            // ... that another small grid walks ray by ray (a full-size grid of early-exit blocks cost 0.35 ms at C3)
            ray_blocks = std::min<int64_t>(ray_blocks, std::max<int64_t>(1, (148 * 8) / h->F));
        }
        // per-ray kernel: the listed particles (everything, when the beam set is not a planar fan)
        const size_t smem = sizeof(float) * 6 * h->B + sizeof(int) * 32 * h->fan.chunk;
        k_update_mbes<<<dim3((unsigned)ray_blocks, (unsigned)h->F), kMbesWarps * 32, smem, h->stream>>>(
            h->st[h->cur], h->M, h->frame, h->dem, h->fan, h->beams, h->beam_aux, h->beam_order, h->meas, h->B,
            (float)h->cfg.range_min, nhiv, lnb, h->logw, list, count, h->wmax_f32);
        LAUNCHED(h, AUVPF_K_UPDATE_MBES, 1);
        CU(cudaGetLastError());
    }
    if (logw_out) return auvpf_get_logw(h, logw_out);
    return AUVPF_OK;
}

int auvpf_mbes_flagged(auvpf_handle *h, int64_t *count)
{
    if (!h || !count) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    std::vector<unsigned int> c(2 * (size_t)h->F);
    CU(cudaMemcpyAsync(c.data(), h->wmax_f32 + h->F, sizeof(unsigned int) * 2 * h->F, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    const bool prof = h->fan.planar && h->use_profile;
    for (int f = 0; f < h->F; ++f) {
        count[2 * f] = prof ? (int64_t)(c[f] - (unsigned int)kWmaxReset) : 0;                    // second pass (map-border variant)
        count[2 * f + 1] = prof ? (int64_t)(c[h->F + f] - (unsigned int)kWmaxReset) : h->M;      // recast ray by ray
    }
    return AUVPF_OK;
}

int auvpf_raycast(auvpf_handle *h, const double *poses, int64_t n, float *ranges_out)
{
    if (!h || !poses || !ranges_out || n <= 0) return fail(AUVPF_ERR_BADARG, "bad argument");
    if (!h->have_dem || h->B <= 0) return fail(AUVPF_ERR_STATE, "raycast needs set_dem and set_beams first");
    Guard g(h);
    const double *pd;
    float *od;
    int rc;
    if ((rc = stage_in(h, S_IN0, poses, (size_t)n * 6, &pd))) return rc;
    if ((rc = stage_out(h, S_OUT0, ranges_out, (size_t)n * h->B, &od))) return rc;
    {
        Prof pr(h, AUVPF_K_OTHER);
        const unsigned char *only = nullptr;
        if (h->fan.planar && h->use_profile) {
            void *fl;
            if ((rc = slot_reserve(h, S_OUT1, (size_t)n, &fl))) return rc;
            const size_t smp = (2 * sizeof(double) + sizeof(BeamAux)) * h->B + 2 * sizeof(BeamAux);
            k_raycast_poses_profile<<<grid1(n, kProfThreads, 1), kProfThreads, smp, h->stream>>>(
                pd, n, h->frame, h->dem, h->fan, h->beam_tan, h->beam_aux, h->B, od, static_cast<unsigned char *>(fl));
            only = static_cast<unsigned char *>(fl);
        }
        const size_t smem = sizeof(float) * 6 * h->B + sizeof(int) * 32 * h->fan.chunk;
        k_raycast_poses<<<grid1(n, kMbesWarps, 1), kMbesWarps * 32, smem, h->stream>>>(
            pd, n, h->frame, h->dem, h->fan, h->beams, h->beam_aux, h->beam_order, h->B, od, only);
        LAUNCHED(h, AUVPF_K_OTHER, 1);
        CU(cudaGetLastError());
    }
    return finish_out(h, ranges_out, (const float *)od, (size_t)n * h->B);
}

// ------------------------------------------------------------------------------------------------
int auvpf_resample(auvpf_handle *h, int32_t mode, int32_t flags, const double *u, int32_t *ancestors_out,
                   const double *rough_var, const float *rough_noise)
{
    if (!h) return fail(AUVPF_ERR_BADARG, "null handle");
    Guard g(h);
    WeightSrc src{h->logw, nullptr, 0};
    return do_resample(h, src, mode, flags, u, ancestors_out, rough_var, rough_noise);
}

int auvpf_resample_weights(auvpf_handle *h, const double *weights, int32_t mode, int32_t flags, const double *u,
                           int32_t *ancestors_out, const double *rough_var, const float *rough_noise)
{
    if (!h || !weights) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    const double *wd;
    int rc = stage_in(h, S_IN2, weights, (size_t)h->N, &wd);
    if (rc) return rc;
    // float64 maximum, then a_i = floor(w_i / max * 2^60) into a scratch array: the resample kernels read integers
    if (!h->a_pre) CU(cudaMalloc(&h->a_pre, sizeof(unsigned long long) * h->cap_N));
    {
        Prof pr(h, AUVPF_K_NORMALISE);
        int blocks = (int)std::min<int64_t>((h->M + kThreads - 1) / kThreads, 1184);
        k_weight_max_lin<<<dim3(blocks, h->F), kThreads, 0, h->stream>>>(wd, h->M, h->wmax_ordered);
        k_finish_max_lin<<<(h->F + 127) / 128, 128, 0, h->stream>>>(h->wmax_ordered, h->wmax, h->F);
        k_linear_to_a<<<grid1(h->M, kThreads, h->F), kThreads, 0, h->stream>>>(wd, h->M, h->wmax, h->a_pre);
        LAUNCHED(h, AUVPF_K_NORMALISE, 3);
        CU(cudaGetLastError());
    }
    WeightSrc src{nullptr, h->a_pre, 1};
    return do_resample(h, src, mode, flags, u, ancestors_out, rough_var, rough_noise);
}

int auvpf_get_weights_q(auvpf_handle *h, uint64_t *q)
{
    if (!h || !q) return fail(AUVPF_ERR_BADARG, "null argument");
    if (!h->have_norm) return fail(AUVPF_ERR_STATE, "no normalisation available: call auvpf_resample first");
    Guard g(h);
    unsigned long long *qd;
    int rc = stage_out(h, S_OUT0, reinterpret_cast<unsigned long long *>(q), (size_t)h->N, &qd);
    if (rc) return rc;
    if (h->global_count != h->M && !h->have_global)
        return fail(AUVPF_ERR_STATE, "sharded filter: the global total is known after auvpf_shard_resample / auvpf_shard_emit");
    {
        int rc2 = run_exact_q(h, qd, nullptr, h->global_count != h->M ? h->c_off + h->n_ranks_global : h->A,
                              h->global_count != h->M ? h->c_off + h->my_rank_global : nullptr);
        if (rc2) return rc2;
    }
    return finish_out(h, reinterpret_cast<unsigned long long *>(q), (const unsigned long long *)qd, (size_t)h->N);
}

int auvpf_apply_indices(auvpf_handle *h, const int32_t *indices, int32_t flags)
{
    if (!h || !indices) return fail(AUVPF_ERR_BADARG, "null argument");
    (void)flags;
    Guard g(h);
    const int32_t *id = nullptr;
    int rc = stage_in(h, S_IN0, indices, (size_t)h->N, &id);
    if (rc) return rc;
    return apply_indices_dev(h, id);
}

// ------------------------------------------------------------------------------------------------
namespace {
// update_loc_pose's numbers from the 16 sums k_pose_stats leaves per filter (auv_pf.py:218-252)
void finish_mean_cov(const auvpf_handle *h, const double *sums, double *mean, double *cov)
{
    const double inv = 1.0 / (double)h->M;
    for (int f = 0; f < h->F; ++f) {
        const double *s = sums + 16 * f;
        const double md[3] = {s[0] * inv, s[1] * inv, s[2] * inv};      // mean relative to the first particle
        for (int c = 0; c < 3; ++c) mean[6 * f + c] = s[13 + c] + md[c];
        mean[6 * f + 3] = s[3] * inv;
        mean[6 * f + 4] = s[4] * inv;
        mean[6 * f + 5] = s[6] * inv;   // mean of the wrapped yaws (auv_pf.py:229-231)
        double *C = cov + 9 * f;
        C[0] = s[7] * inv - md[0] * md[0];
        C[4] = s[8] * inv - md[1] * md[1];
        C[8] = s[9] * inv - md[2] * md[2];
        C[1] = s[10] * inv - md[0] * md[1];
        C[2] = s[11] * inv - md[0] * md[2];
        C[5] = s[12] * inv - md[1] * md[2];
        C[3] = C[1];          // only [1,0] is mirrored (auv_pf.py:246)
        C[6] = 0.0;
        C[7] = 0.0;
    }
}

int reserve(DevBuf &b, size_t bytes)
{
    if (b.cap >= bytes) return AUVPF_OK;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    CU(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return AUVPF_OK;
}
}   // namespace

// auv_pf.update_loc_pose with options: weighted by the current log-weights, circular yaw mean,
// full symmetric covariance.  flags == 0 is auvpf_mean_cov (the reference's numbers).
int auvpf_estimate(auvpf_handle *h, int32_t flags, double *mean, double *cov)
{
    if (!h || !mean || !cov) return fail(AUVPF_ERR_BADARG, "null argument");
    if (flags == AUVPF_EST_REFERENCE) return auvpf_mean_cov(h, mean, cov);
    if (flags & ~(AUVPF_EST_WEIGHTED | AUVPF_EST_CIRCULAR_YAW | AUVPF_EST_FULL_COV)) return fail(AUVPF_ERR_BADARG, "unknown estimate flags %d", flags);
    Guard g(h);
    int rc;
    const bool weighted = (flags & AUVPF_EST_WEIGHTED) != 0;
    if (weighted && (rc = ensure_wmax_f32(h))) return rc;
    void *buf;
    if ((rc = slot_reserve(h, S_OUT1, sizeof(double) * 20 * h->F, &buf))) return rc;
    double *stats = static_cast<double *>(buf);
    std::vector<double> host(20 * (size_t)h->F);
    {
        Prof pr(h, AUVPF_K_ESTIMATE);
        CU(cudaMemsetAsync(stats, 0, sizeof(double) * 20 * h->F, h->stream));
        int blocks = (int)std::min<int64_t>((h->M + kThreads - 1) / kThreads, std::max(1, 592 / h->F));
        k_pose_stats_w<<<dim3(blocks, h->F), kThreads, 0, h->stream>>>(h->st[h->cur], h->M, h->logw, h->wmax_f32, weighted ? 1 : 0, stats);
        LAUNCHED(h, AUVPF_K_ESTIMATE, 1);
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(host.data(), stats, sizeof(double) * 20 * h->F, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int f = 0; f < h->F; ++f) {
        const double *s = host.data() + 20 * f;
        if (!(s[0] > 0.0)) return fail(AUVPF_ERR_DEGENERATE, "filter %d: no finite positive weight; estimate undefined", f);
        const double inv = 1.0 / s[0];
        const double md[3] = {s[1] * inv, s[2] * inv, s[3] * inv};
        for (int c = 0; c < 3; ++c) mean[6 * f + c] = s[15 + c] + md[c];
        mean[6 * f + 3] = s[4] * inv;
        mean[6 * f + 4] = s[5] * inv;
        mean[6 * f + 5] = (flags & AUVPF_EST_CIRCULAR_YAW) ? std::atan2(s[7], s[8]) : s[6] * inv;
        double *C = cov + 9 * f;
        C[0] = s[9] * inv - md[0] * md[0];
        C[4] = s[10] * inv - md[1] * md[1];
        C[8] = s[11] * inv - md[2] * md[2];
        C[1] = C[3] = s[12] * inv - md[0] * md[1];
        C[2] = s[13] * inv - md[0] * md[2];
        C[5] = s[14] * inv - md[1] * md[2];
        if (flags & AUVPF_EST_FULL_COV) {
            C[6] = C[2];
            C[7] = C[5];
        } else {
            C[6] = 0.0;          // the reference's fill pattern (auv_pf.py:238-246)
            C[7] = 0.0;
        }
    }
    return AUVPF_OK;
}

// PoseArray of auv_pf.loc_loop (auv_pf.py:264-285) for a strided subsample: (F, n, 7) float64 rows
// [x y z qx qy qz qw], n = min(ceil(M / stride), max_poses) per filter (stride <= 0: chosen so that n <= max_poses).
int auvpf_pose_array(auvpf_handle *h, int64_t stride, int64_t max_poses, double *poses7_out, int64_t *n_out)
{
    if (!h || !poses7_out || !n_out) return fail(AUVPF_ERR_BADARG, "null argument");
    if (max_poses < 1) return fail(AUVPF_ERR_BADARG, "max_poses must be positive");
    if (stride <= 0) stride = (h->M + max_poses - 1) / max_poses;
    const int64_t n = std::min<int64_t>((h->M + stride - 1) / stride, max_poses);
    Guard g(h);
    double *od;
    int rc = stage_out(h, S_OUT0, poses7_out, (size_t)n * 7 * h->F, &od);
    if (rc) return rc;
    k_pose_array<<<grid1(n, kThreads, h->F), kThreads, 0, h->stream>>>(h->st[h->cur], h->M, stride, n, od);
    LAUNCHED(h, AUVPF_K_OTHER, 1);
    CU(cudaGetLastError());
    *n_out = n;
    return finish_out(h, poses7_out, (const double *)od, (size_t)n * 7 * h->F);
}

int auvpf_pose_array_topk(auvpf_handle *h, int64_t k, double *poses7_out, float *logw_out, int32_t *index_out)
{
    if (!h || !poses7_out) return fail(AUVPF_ERR_BADARG, "null argument");
    if (k < 1 || k > h->M) return fail(AUVPF_ERR_BADARG, "k must be in [1, particle_count]");
    if (h->M > 0xFFFFFFFFll) return fail(AUVPF_ERR_BADARG, "top-k: more than 2^32 particles per filter");
    Guard g(h);
    const int F = h->F;
    // scratch: [F] states | [F][65536] histogram | [F][k] indices | [F][k] log-weights
    const size_t off_hist = (sizeof(TopkState) * F + 255) / 256 * 256, off_idx = off_hist + sizeof(unsigned int) * 65536 * (size_t)F;
    const size_t off_lw = off_idx + sizeof(int) * (size_t)k * F, total = off_lw + sizeof(float) * (size_t)k * F;
    void *scratch;
    int rc = slot_reserve(h, S_IN3, total, &scratch);
    if (rc) return rc;
    char *base = static_cast<char *>(scratch);
    TopkState *st = reinterpret_cast<TopkState *>(base);
    unsigned int *hist = reinterpret_cast<unsigned int *>(base + off_hist);
    int *idx = reinterpret_cast<int *>(base + off_idx);
    float *lw = reinterpret_cast<float *>(base + off_lw);
    std::vector<TopkState> init((size_t)F);
    for (auto &s : init) s = TopkState{0ull, (long long)k, 0u, 0u};
    CU(cudaMemcpyAsync(st, init.data(), sizeof(TopkState) * F, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));                      // (init is a stack vector)
    CU(cudaMemsetAsync(hist, 0, sizeof(unsigned int) * 65536 * (size_t)F, h->stream));
    const int blocks = (int)std::min<int64_t>((h->M + kThreads - 1) / kThreads, std::max(1, 1184 / F));
    for (int round = 0; round < 4; ++round) {
        k_topk_hist<<<dim3(blocks, F), kThreads, 0, h->stream>>>(h->logw, h->M, st, round, hist);
        k_topk_pick<<<F, 1024, 0, h->stream>>>(st, hist);
    }
    k_topk_emit<<<dim3(blocks, F), kThreads, 0, h->stream>>>(h->logw, h->M, st, k, idx);
    LAUNCHED(h, AUVPF_K_OTHER, 9);
    double *od;
    if ((rc = stage_out(h, S_OUT0, poses7_out, (size_t)k * 7 * F, &od))) return rc;
    k_pose_array_idx<<<grid1(k, kThreads, F), kThreads, 0, h->stream>>>(h->st[h->cur], h->logw, h->M, idx, k, od, lw);
    LAUNCHED(h, AUVPF_K_OTHER, 1);
    CU(cudaGetLastError());
    if ((rc = finish_out(h, poses7_out, (const double *)od, (size_t)k * 7 * F))) return rc;
    if (logw_out) CU(cudaMemcpyAsync(logw_out, lw, sizeof(float) * (size_t)k * F, cudaMemcpyDeviceToHost, h->stream));
    if (index_out) CU(cudaMemcpyAsync(index_out, idx, sizeof(int) * (size_t)k * F, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return AUVPF_OK;
}

// ---- KLD-style adaptation of the particle count ---------------------------------------------------
int auvpf_kld_bins(auvpf_handle *h, double bin_xy, double bin_yaw, int64_t *bins_out)
{
    if (!h || !bins_out) return fail(AUVPF_ERR_BADARG, "null argument");
    if (!(bin_xy > 0.0) || !(bin_yaw > 0.0)) return fail(AUVPF_ERR_BADARG, "bin sizes must be positive");
    if (h->F > 256) return fail(AUVPF_ERR_BADARG, "kld_bins: at most 256 filters per handle");
    Guard g(h);
    size_t slots = 1024;
    while (slots < 2 * (size_t)h->N) slots <<= 1;
    const size_t off_tab = 256 * ((sizeof(unsigned long long) * (h->F + 1) + 255) / 256);
    void *scratch;
    int rc = slot_reserve(h, S_IN3, off_tab + sizeof(unsigned long long) * slots, &scratch);
    if (rc) return rc;
    unsigned long long *out = static_cast<unsigned long long *>(scratch);
    unsigned long long *table = reinterpret_cast<unsigned long long *>(static_cast<char *>(scratch) + off_tab);
    CU(cudaMemsetAsync(out, 0, sizeof(unsigned long long) * (h->F + 1), h->stream));
    CU(cudaMemsetAsync(table, 0xFF, sizeof(unsigned long long) * slots, h->stream));
    const int blocks = (int)std::min<int64_t>((h->M + kThreads - 1) / kThreads, std::max(1, 1184 / h->F));
    k_kld_bins<<<dim3(blocks, h->F), kThreads, 0, h->stream>>>(h->st[h->cur], h->M, 1.0 / bin_xy, 1.0 / bin_yaw, table, slots - 1, out, h->F);
    LAUNCHED(h, AUVPF_K_OTHER, 1);
    CU(cudaGetLastError());
    std::vector<unsigned long long> host((size_t)h->F + 1);
    CU(cudaMemcpyAsync(host.data(), out, sizeof(unsigned long long) * (h->F + 1), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (host[h->F]) return fail(AUVPF_ERR_BADARG, "kld_bins: a particle lies outside +-2^21 bins in x or y, or +-2^11 in yaw (or is NaN)");
    for (int f = 0; f < h->F; ++f) bins_out[f] = (int64_t)host[f];
    return AUVPF_OK;
}

// Fox's bound: with probability 1 - delta the KL divergence between the sample-based estimate and the true posterior
// stays below eps when n >= (k - 1) / (2 eps) (1 - 2 / (9 (k - 1)) + sqrt(2 / (9 (k - 1))) z)^3, z the 1 - delta
// quantile of the standard normal, k the occupied bins.  Plain float64, no library power: the oracle gets the same bits.
int64_t auvpf_kld_target(int64_t bins, double eps, double z_quantile, int64_t n_min, int64_t n_max)
{
    int64_t n = n_min;
    if (bins > 1 && eps > 0.0) {
        const double k1 = (double)(bins - 1), a = 2.0 / (9.0 * k1);
        const double t = 1.0 - a + sqrt(a) * z_quantile;
        const double v = k1 / (2.0 * eps) * (t * t * t);
        n = v >= 9.0e18 ? n_max : (int64_t)ceil(v);
    }
    if (n < n_min) n = n_min;
    if (n > n_max) n = n_max;
    return n;
}

int auvpf_set_next_particle_count(auvpf_handle *h, int64_t n)
{
    if (!h) return fail(AUVPF_ERR_BADARG, "null handle");
    Guard g(h);
    if (n == 0) {
        h->next_count = 0;
        return AUVPF_OK;
    }
    if (h->F != 1 || h->global_count != h->M || h->n_peers > 1)
        return fail(AUVPF_ERR_BADARG, "the particle count can change for one filter per handle on one GPU only");
    if (n < 1 || n > h->cap_M) return fail(AUVPF_ERR_BADARG, "particle count %lld outside [1, %lld] (the count the handle was created with)", (long long)n, (long long)h->cap_M);
    h->next_count = n;
    return AUVPF_OK;
}

int64_t auvpf_particle_capacity(auvpf_handle *h) { return h ? h->cap_N : 0; }

int auvpf_pose_sums_buffer(auvpf_handle *h, void **sums_f64)
{
    if (!h || !sums_f64) return fail(AUVPF_ERR_BADARG, "null argument");
    *sums_f64 = h->sums;
    return AUVPF_OK;
}

int auvpf_pose_sums_dev(auvpf_handle *h)
{
    if (!h) return fail(AUVPF_ERR_BADARG, "null handle");
    Guard g(h);
    Prof pr(h, AUVPF_K_ESTIMATE);
    CU(cudaMemsetAsync(h->sums, 0, sizeof(double) * 16 * h->F, h->stream));
    int blocks = (int)std::min<int64_t>((h->M + kThreads - 1) / kThreads, std::max(1, 592 / h->F));
    k_pose_stats<<<dim3(blocks, h->F), kThreads, 0, h->stream>>>(h->st[h->cur], h->M, h->sums);
    LAUNCHED(h, AUVPF_K_ESTIMATE, 1);
    CU(cudaGetLastError());
    return AUVPF_OK;
}

int auvpf_mean_cov(auvpf_handle *h, double *mean, double *cov)
{
    if (!h || !mean || !cov) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    std::vector<double> host(16 * h->F);
    {
        Prof pr(h, AUVPF_K_ESTIMATE);
        CU(cudaMemsetAsync(h->sums, 0, sizeof(double) * 16 * h->F, h->stream));
        int blocks = (int)std::min<int64_t>((h->M + kThreads - 1) / kThreads, std::max(1, 592 / h->F));
        k_pose_stats<<<dim3(blocks, h->F), kThreads, 0, h->stream>>>(h->st[h->cur], h->M, h->sums);
        LAUNCHED(h, AUVPF_K_ESTIMATE, 1);
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(host.data(), h->sums, sizeof(double) * 16 * h->F, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    finish_mean_cov(h, host.data(), mean, cov);
    return AUVPF_OK;
}

// ------------------------------------------------------------------------------------------------
// Replay of a recorded sequence: n_cycles x [n_steps predicts, MBES update, resample,
// estimate] enqueued back to back from device-resident inputs.  The host uploads the sequence once,
// reads the estimates back once, and does not wait in between.
int auvpf_replay(auvpf_handle *h, int32_t n_cycles, int32_t n_steps, int32_t n_odom, const double *odom,
                 const float *ranges, int32_t mode, int32_t flags, const double *u, const double *rough_var,
                 double *mean_out, double *cov_out)
{
    if (!h || !odom || !ranges) return fail(AUVPF_ERR_BADARG, "null argument");
    if (n_cycles < 1) return fail(AUVPF_ERR_BADARG, "n_cycles must be positive");
    if (n_steps < 1 || n_steps > AUVPF_MAX_PREDICT_STEPS)
        return fail(AUVPF_ERR_BADARG, "n_steps must be 1..%d", AUVPF_MAX_PREDICT_STEPS);
    if (n_odom != 1 && n_odom != h->F) return fail(AUVPF_ERR_BADARG, "n_odom must be 1 or n_filters");
    if (is_device_ptr(odom)) return fail(AUVPF_ERR_BADARG, "odom must be host memory (11 scalars per filter and step)");
    if (mode != AUVPF_SYSTEMATIC || (flags & (AUVPF_RS_INPLACE_REF | AUVPF_RS_STRICT_REF)))
        return fail(AUVPF_ERR_BADARG, "replay runs the enqueue-only cycle: AUVPF_SYSTEMATIC, AUVPF_RS_SORTED");
    if (h->global_count != h->M) return fail(AUVPF_ERR_BADARG, "replay is not available for a particle-sharded filter");
    if (h->next_count > 0) return fail(AUVPF_ERR_STATE, "a change of the particle count is pending (auvpf_set_next_particle_count): replay keeps the count");
    if (!h->have_dem || h->B <= 0) return fail(AUVPF_ERR_STATE, "replay needs set_dem and set_beams first");
    if ((mean_out == nullptr) != (cov_out == nullptr)) return fail(AUVPF_ERR_BADARG, "mean_out and cov_out go together");
    if (mean_out && (is_device_ptr(mean_out) || is_device_ptr(cov_out)))
        return fail(AUVPF_ERR_BADARG, "mean_out / cov_out must be host memory");
    Guard g(h);
    int rc;
    const int F = h->F;
    // every step's motion parameters (make_predict: quaternion -> roll, pitch on the host), one upload
    std::vector<PredictParams> all((size_t)n_cycles * n_steps * F);
    for (int64_t t = 0; t < (int64_t)n_cycles * n_steps; ++t)
        for (int f = 0; f < F; ++f)
            all[(size_t)t * F + f] = make_predict(odom + 11 * ((size_t)t * n_odom + (n_odom > 1 ? f : 0)), h->cfg.motion_cov);
    if ((rc = reserve(h->replay_params, sizeof(PredictParams) * all.size()))) return rc;
    CU(cudaMemcpyAsync(h->replay_params.p, all.data(), sizeof(PredictParams) * all.size(), cudaMemcpyHostToDevice, h->stream));
    const PredictParams *pp_all = static_cast<const PredictParams *>(h->replay_params.p);
    // pings
    const float *rg = ranges;
    const size_t ping = (size_t)F * h->B;
    if (!is_device_ptr(ranges)) {
        if ((rc = reserve(h->replay_ranges, sizeof(float) * ping * n_cycles))) return rc;
        CU(cudaMemcpyAsync(h->replay_ranges.p, ranges, sizeof(float) * ping * n_cycles, cudaMemcpyHostToDevice, h->stream));
        rg = static_cast<const float *>(h->replay_ranges.p);
    }
    // the resamplers' uniforms: the caller's, or the same Philox draws the cycle-by-cycle calls make
    const double *ud = u;
    std::vector<double> uh;
    if (!u || !is_device_ptr(u)) {
        uh.resize((size_t)n_cycles * F);
        for (int c = 0; c < n_cycles; ++c)
            for (int f = 0; f < F; ++f)
                uh[(size_t)c * F + f] =
                    u ? u[(size_t)c * F + f] : uniform53(philox_for(h->cfg.seed, f, h->step_resample + (uint32_t)c, kStreamResample, 0));
        if ((rc = reserve(h->replay_u, sizeof(double) * uh.size()))) return rc;
        CU(cudaMemcpyAsync(h->replay_u.p, uh.data(), sizeof(double) * uh.size(), cudaMemcpyHostToDevice, h->stream));
        ud = static_cast<const double *>(h->replay_u.p);
    }
    double *stats = nullptr;
    if (mean_out) {
        if ((rc = reserve(h->replay_stats, sizeof(double) * 16 * F * n_cycles))) return rc;
        stats = static_cast<double *>(h->replay_stats.p);
        CU(cudaMemsetAsync(stats, 0, sizeof(double) * 16 * F * n_cycles, h->stream));
    }
    // the staging vectors above must outlive their (pageable-memory) copies: those are synchronous with
    // respect to the host buffer, so they are safe to drop when this function returns
    const WeightSrc src{h->logw, nullptr, 0};
    for (int c = 0; c < n_cycles; ++c) {
        {
            Prof pr(h, AUVPF_K_PROPAGATE);
            const PredictParams *pp = pp_all + (size_t)c * n_steps * F;
            k_propagate<<<grid1(h->M, kThreads * 4, F), kThreads, 0, h->stream>>>(
                h->st[h->cur], h->M, all[(size_t)c * n_steps * F], (n_odom > 1 || n_steps > 1) ? pp : nullptr, n_steps,
                h->cfg.seed, h->step_predict, h->global_offset, h->global_count, nullptr);
            h->step_predict += (uint32_t)n_steps;
            LAUNCHED(h, AUVPF_K_PROPAGATE, 1);
            CU(cudaGetLastError());
        }
        if ((rc = auvpf_update_mbes(h, rg + (size_t)c * ping, nullptr))) return rc;
        if ((rc = do_resample(h, src, mode, flags | AUVPF_RS_ASYNC, ud + (size_t)c * F, nullptr, rough_var, nullptr))) return rc;
        if (stats) {
            Prof pr(h, AUVPF_K_ESTIMATE);
            int blocks = (int)std::min<int64_t>((h->M + kThreads - 1) / kThreads, std::max(1, 592 / F));
            k_pose_stats<<<dim3(blocks, F), kThreads, 0, h->stream>>>(h->st[h->cur], h->M, stats + (size_t)c * 16 * F);
            LAUNCHED(h, AUVPF_K_ESTIMATE, 1);
            CU(cudaGetLastError());
        }
    }
    if (stats) {
        std::vector<double> host((size_t)16 * F * n_cycles);
        CU(cudaMemcpyAsync(host.data(), stats, sizeof(double) * host.size(), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (int c = 0; c < n_cycles; ++c)
            finish_mean_cov(h, host.data() + (size_t)c * 16 * F, mean_out + (size_t)c * 6 * F, cov_out + (size_t)c * 9 * F);
    }
    return AUVPF_OK;
}

int auvpf_resample_status(auvpf_handle *h, int32_t *degenerate)
{
    if (!h || !degenerate) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    CU(cudaMemcpyAsync(degenerate, h->degenerate, sizeof(int) * h->F, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return AUVPF_OK;
}

int auvpf_set_neff_gate(auvpf_handle *h, double ratio)
{
    if (!h) return fail(AUVPF_ERR_BADARG, "null handle");
    if (!(ratio >= 0.0) || ratio > 1.0) return fail(AUVPF_ERR_BADARG, "N_eff gate ratio must lie in [0, 1]");
    h->neff_ratio = ratio;
    return AUVPF_OK;
}

int auvpf_resample_skipped(auvpf_handle *h, int32_t *skipped)
{
    if (!h || !skipped) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    CU(cudaMemcpyAsync(skipped, h->skipped, sizeof(int) * h->F, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return AUVPF_OK;
}

int auvpf_neff_fixed(auvpf_handle *h, double *neff)
{
    if (!h || !neff) return fail(AUVPF_ERR_BADARG, "null argument");
    if (!h->have_norm) return fail(AUVPF_ERR_STATE, "no normalisation available: call auvpf_resample first");
    if (h->global_count != h->M) return fail(AUVPF_ERR_STATE, "not available for a particle-sharded filter");
    Guard g(h);
    std::vector<unsigned long long> S(3 * (size_t)h->F);
    CU(cudaMemcpyAsync(S.data(), h->neff_pub, sizeof(unsigned long long) * 3 * h->F, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int f = 0; f < h->F; ++f) neff[f] = neff_fixed_a(S[3 * f], S[3 * f + 1], S[3 * f + 2]);
    return AUVPF_OK;
}

int auvpf_neff(auvpf_handle *h, double *neff)
{
    if (!h || !neff) return fail(AUVPF_ERR_BADARG, "null argument");
    if (!h->have_norm) return fail(AUVPF_ERR_STATE, "no normalisation available: call auvpf_resample first");
    Guard g(h);
    std::vector<double> s(h->F);
    if (h->global_count != h->M) return fail(AUVPF_ERR_STATE, "not available for a particle-sharded filter");
    if (h->sumsq_lazy) {        // valid while the weights of the last resample are still in place (like auvpf_get_weights_q)
        int rc = run_exact_q(h, nullptr, h->sumsq, h->A, nullptr);
        if (rc) return rc;
        h->sumsq_lazy = false;
    }
    CU(cudaMemcpyAsync(s.data(), h->sumsq, sizeof(double) * h->F, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int f = 0; f < h->F; ++f) neff[f] = s[f] > 0.0 ? 1.0 / s[f] : 0.0;
    return AUVPF_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU (particle-sharded single filter)
namespace {
// per-shard totals (G x 6 words: A lo, A hi, S1, S2 lo, S2 hi, log-weight buffer) -> mass below every shard
__global__ void k_build_coff(const unsigned long long *totals_all, int G, U128 *c_off)
{
    if (threadIdx.x || blockIdx.x) return;
    U128 acc = u128(0);
    for (int g = 0; g < G; ++g) {
        c_off[g] = acc;
        acc = acc + u128(totals_all[6 * g], totals_all[6 * g + 1]);
    }
    c_off[G] = acc;
}

__global__ void k_pack_totals(const U128 *A, const unsigned long long *neff_pub, int parity, unsigned long long *out)
{
    if (threadIdx.x || blockIdx.x) return;
    out[0] = A[0].lo;
    out[1] = A[0].hi;
    out[2] = neff_pub[0];
    out[3] = neff_pub[1];
    out[4] = neff_pub[2];
    out[5] = (unsigned long long)parity;
}

int float_to_ordered_host(float f)
{
    int b;
    memcpy(&b, &f, sizeof b);
    return b >= 0 ? b : b ^ 0x7FFFFFFF;
}

int need_single_filter(auvpf_handle *h)
{
    if (h->F != 1) return fail(AUVPF_ERR_BADARG, "particle sharding is for a single filter (n_filters == 1)");
    return AUVPF_OK;
}

// peers' slabs are known (h->peer_slab, offsets in `offs`): build what the kernels read
int finish_attach(auvpf_handle *h, const int64_t (*offs)[kSlabParts], int n_ranks, int my_rank)
{
    std::vector<SelShard> sh((size_t)n_ranks);
    h->mbox.G = n_ranks;
    h->mbox.rank = my_rank;
    for (int r = 0; r < n_ranks; ++r) {
        sh[r] = shard_at(h->peer_slab[r], offs[r]);
        h->mbox.peer[r] = reinterpret_cast<unsigned long long *>(static_cast<char *>(h->peer_slab[r]) + offs[r][P_MBOX]);
    }
    h->n_peers = n_ranks;
    h->my_peer = my_rank;
    CU(cudaMalloc(&h->peer_shards, sizeof(SelShard) * n_ranks));
    CU(cudaMemcpy(h->peer_shards, sh.data(), sizeof(SelShard) * n_ranks, cudaMemcpyHostToDevice));
    return AUVPF_OK;
}
}   // namespace

// ---- staged variant: the scalars travel through the caller (host, or device buffers it reduces in place) --------
int auvpf_shard_max(auvpf_handle *h, float *m_out)
{
    if (!h || !m_out) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    int rc = need_single_filter(h);
    if (rc) return rc;
    if ((rc = ensure_wmax_f32(h))) return rc;
    int k;
    CU(cudaMemcpyAsync(&k, h->wmax_f32, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (k > kWmaxReset) {
        const int b = k >= 0 ? k : k ^ 0x7FFFFFFF;
        memcpy(m_out, &b, sizeof(float));
    } else
        *m_out = -INFINITY;
    return AUVPF_OK;
}

int auvpf_shard_sum(auvpf_handle *h, const float *m_global, uint64_t *totals_out)
{
    if (!h || !m_global || !totals_out) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    int rc = need_single_filter(h);
    if (rc) return rc;
    const int k = std::isfinite(*m_global) ? float_to_ordered_host(*m_global) : kWmaxReset;
    if ((rc = tiny_upload(h, h->wmax_f32, &k, sizeof(int)))) return rc;
    h->wmax_valid = true;       // the buffer now holds the GLOBAL maximum
    WeightSrc src{h->logw, nullptr, 0};
    if ((rc = run_weights(h, src, /*use_mailbox=*/false))) return rc;
    h->wmax_valid = false;      // ... which is not this shard's: recompute next time unless an update refreshes it
    k_pack_totals<<<1, 32, 0, h->stream>>>(h->A, h->neff_pub, h->lw, h->shard_totals);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(totals_out, h->shard_totals, sizeof(unsigned long long) * 6, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return AUVPF_OK;
}

int auvpf_shard_buffers(auvpf_handle *h, void **wmax_i32, void **totals_u64)
{
    if (!h) return fail(AUVPF_ERR_BADARG, "null handle");
    if (wmax_i32) *wmax_i32 = h->wmax_f32;
    if (totals_u64) *totals_u64 = h->shard_totals;
    return AUVPF_OK;
}

int auvpf_shard_max_dev(auvpf_handle *h)
{
    if (!h) return fail(AUVPF_ERR_BADARG, "null handle");
    Guard g(h);
    return ensure_wmax_f32(h);
}

int auvpf_shard_weights_dev(auvpf_handle *h)
{
    if (!h) return fail(AUVPF_ERR_BADARG, "null handle");
    Guard g(h);
    int rc = need_single_filter(h);
    if (rc) return rc;
    WeightSrc src{h->logw, nullptr, 0};
    h->wmax_valid = true;       // the buffer now holds the GLOBAL maximum
    if ((rc = run_weights(h, src, /*use_mailbox=*/false))) return rc;
    h->wmax_valid = false;
    k_pack_totals<<<1, 32, 0, h->stream>>>(h->A, h->neff_pub, h->lw, h->shard_totals);
    LAUNCHED(h, AUVPF_K_NORMALISE, 1);
    CU(cudaGetLastError());
    return AUVPF_OK;
}

int auvpf_shard_emit(auvpf_handle *h, double u, const uint64_t *totals_all, int32_t n_ranks, int32_t my_rank, int64_t out_lo,
                     int64_t out_hi, void *records, int64_t *ancestors_out)
{
    if (!h || !totals_all) return fail(AUVPF_ERR_BADARG, "null argument");
    if (n_ranks < 1 || n_ranks > kMaxPeers || my_rank < 0 || my_rank >= n_ranks)
        return fail(AUVPF_ERR_BADARG, "ranks: 1..%d, got %d (my rank %d)", kMaxPeers, n_ranks, my_rank);
    const int64_t n_out = out_hi - out_lo;
    if (n_out < 0) return fail(AUVPF_ERR_BADARG, "out_hi < out_lo");
    Guard g(h);
    int rc = need_single_filter(h);
    if (rc) return rc;
    const unsigned long long *td = nullptr;
    if ((rc = stage_in(h, S_IN2, reinterpret_cast<const unsigned long long *>(totals_all), (size_t)n_ranks * 6, &td))) return rc;
    k_build_coff<<<1, 32, 0, h->stream>>>(td, n_ranks, h->c_off);
    CU(cudaGetLastError());
    h->have_global = true;
    h->n_ranks_global = n_ranks;
    h->my_rank_global = my_rank;
    if ((rc = tiny_upload(h, h->u, &u, sizeof u))) return rc;
    if (n_out == 0) return AUVPF_OK;
    if (!records) return fail(AUVPF_ERR_BADARG, "null records");
    Record *rd = nullptr;
    int64_t *ad = nullptr;
    if ((rc = stage_out(h, S_OUT0, static_cast<Record *>(records), (size_t)n_out, &rd))) return rc;
    if (ancestors_out && (rc = stage_out(h, S_OUT1, ancestors_out, (size_t)n_out, &ad))) return rc;
    WeightSrc src{h->logw, nullptr, 0};
    SelArgs sa = select_args(h, src);
    sa.G = n_ranks;
    sa.rank = my_rank;
    sa.out_lo = out_lo;
    sa.n_out = n_out;
    sa.c_off = h->c_off;
    sa.records = rd;
    sa.anc64 = ad;
    sa.rough = make_noise(h, nullptr, nullptr, 0, 0, kStreamRoughen);
    sa.dst = StateView{};
    {
        Prof pr(h, AUVPF_K_SELECT_GATHER);
        k_select<kSelRecords><<<grid1(n_out, kSelTile, 1), kSelThreads, 0, h->stream>>>(sa);
        LAUNCHED(h, AUVPF_K_SELECT_GATHER, 2);
        CU(cudaGetLastError());
    }
    if ((rc = finish_out(h, static_cast<Record *>(records), (const Record *)rd, (size_t)n_out))) return rc;
    if (ancestors_out && (rc = finish_out(h, ancestors_out, (const int64_t *)ad, (size_t)n_out))) return rc;
    return AUVPF_OK;
}

// ---- peer memory: the slabs of all ranks mapped into every rank ----------------------------------------
namespace {
struct IpcBlob {            // what a rank publishes: AUVPF_IPC_BLOB_BYTES
    cudaIpcMemHandle_t handle;
    int64_t M, slab_bytes;
    int64_t off[kSlabParts];
    int32_t tiles, device;
};
static_assert(sizeof(IpcBlob) <= AUVPF_IPC_BLOB_BYTES, "IpcBlob must fit the ABI's blob size");
}   // namespace

int auvpf_ipc_export(auvpf_handle *h, void *blob)
{
    if (!h || !blob) return fail(AUVPF_ERR_BADARG, "null argument");
    Guard g(h);
    int rc = need_single_filter(h);
    if (rc) return rc;
    IpcBlob b{};
    CU(cudaIpcGetMemHandle(&b.handle, h->slab));
    b.M = h->M;
    b.slab_bytes = (int64_t)h->slab_bytes;
    for (int k = 0; k < kSlabParts; ++k) b.off[k] = h->slab_off[k];
    b.tiles = h->tiles_per_filter;
    b.device = h->device;
    memset(blob, 0, AUVPF_IPC_BLOB_BYTES);
    memcpy(blob, &b, sizeof b);
    return AUVPF_OK;
}

int auvpf_ipc_import(auvpf_handle *h, const void *blobs, int32_t n_ranks, int32_t my_rank)
{
    if (!h || !blobs) return fail(AUVPF_ERR_BADARG, "null argument");
    if (n_ranks < 1 || n_ranks > kMaxPeers || my_rank < 0 || my_rank >= n_ranks)
        return fail(AUVPF_ERR_BADARG, "ranks: 1..%d, got %d (my rank %d)", kMaxPeers, n_ranks, my_rank);
    if (h->n_peers) return fail(AUVPF_ERR_STATE, "peers already attached");
    Guard g(h);
    int64_t offs[kMaxPeers][kSlabParts];
    for (int r = 0; r < n_ranks; ++r) {
        IpcBlob b;
        memcpy(&b, static_cast<const char *>(blobs) + (size_t)r * AUVPF_IPC_BLOB_BYTES, sizeof b);
        if (b.M != h->M || b.tiles != h->tiles_per_filter)
            return fail(AUVPF_ERR_BADARG, "rank %d holds %lld particles, this rank %lld", r, (long long)b.M, (long long)h->M);
        void *base = h->slab;
        if (r != my_rank) {
            if (b.device != h->device) {
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, h->device, b.device));
                if (!can) return fail(AUVPF_ERR_CUDA, "device %d cannot access device %d's memory", h->device, b.device);
            }
            CU(cudaIpcOpenMemHandle(&base, b.handle, cudaIpcMemLazyEnablePeerAccess));
            h->peer_ipc[r] = true;
        }
        h->peer_slab[r] = base;
        for (int k = 0; k < kSlabParts; ++k) offs[r][k] = b.off[k];
    }
    return finish_attach(h, offs, n_ranks, my_rank);
}

int auvpf_peer_attach(auvpf_handle *h, auvpf_handle *const *handles, int32_t n_ranks, int32_t my_rank)
{
    if (!h || !handles) return fail(AUVPF_ERR_BADARG, "null argument");
    if (n_ranks < 1 || n_ranks > kMaxPeers || my_rank < 0 || my_rank >= n_ranks || handles[my_rank] != h)
        return fail(AUVPF_ERR_BADARG, "ranks: 1..%d, got %d (my rank %d), handles[my_rank] must be this handle", kMaxPeers,
                    n_ranks, my_rank);
    if (h->n_peers) return fail(AUVPF_ERR_STATE, "peers already attached");
    Guard g(h);
    int rc = need_single_filter(h);
    if (rc) return rc;
    int64_t offs[kMaxPeers][kSlabParts];
    for (int r = 0; r < n_ranks; ++r) {
        auvpf_handle *p = handles[r];
        if (!p || p->M != h->M || p->F != 1)
            return fail(AUVPF_ERR_BADARG, "rank %d must hold one filter of %lld particles", r, (long long)h->M);
        if (p->device != h->device) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, h->device, p->device));
            if (!can) return fail(AUVPF_ERR_CUDA, "device %d cannot access device %d's memory", h->device, p->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(AUVPF_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", p->device, cudaGetErrorString(e));
            cudaGetLastError();
        }
        h->peer_slab[r] = p->slab;
        for (int k = 0; k < kSlabParts; ++k) offs[r][k] = p->slab_off[k];
    }
    return finish_attach(h, offs, n_ranks, my_rank);
}

namespace {
int shard_check(auvpf_handle *h)
{
    if (!h) return fail(AUVPF_ERR_BADARG, "null handle");
    if (h->n_peers < 1) return fail(AUVPF_ERR_STATE, "attach the peers first (auvpf_ipc_import / auvpf_peer_attach)");
    if (h->global_count != h->M * h->n_peers || h->global_offset != h->M * h->my_peer)
        return fail(AUVPF_ERR_STATE, "global_offset / global_count of the handle do not match its rank among the peers");
    return AUVPF_OK;
}

// the three stages of a peer-memory resample on one rank
int shard_stage_max(auvpf_handle *h, bool separate_post, double u)
{
    Guard g(h);
    int rc;
    h->epoch++;                 // every rank resamples the same number of times: the mailbox's epoch
    if ((rc = tiny_upload(h, h->u, &u, sizeof u))) return rc;      // (ahead of the kernels: they stay adjacent in the stream)
    if ((rc = ensure_wmax_f32(h))) return rc;
    if (separate_post) {
        MailboxView mb = h->mbox;
        mb.epoch = h->epoch;
        k_post_max<<<1, 32, 0, h->stream>>>(h->wmax_f32, mb);
        LAUNCHED(h, AUVPF_K_NORMALISE, 1);
        CU(cudaGetLastError());
    }
    return AUVPF_OK;
}

int shard_stage_weights(auvpf_handle *h, bool post_inside)
{
    Guard g(h);
    WeightSrc src{h->logw, nullptr, 0};
    int rc = run_weights(h, src, /*use_mailbox=*/true, post_inside);
    h->wmax_valid = false;      // (k_weights used the global maximum; the local one is refreshed by the next update)
    return rc;
}

int shard_stage_select(auvpf_handle *h, const double *rough_var, const float *rough_noise, int64_t *ancestors_out_dev)
{
    Guard g(h);
    int rc;
    const float *rn = nullptr;
    if (rough_noise && (rc = stage_in(h, S_IN1, rough_noise, (size_t)h->N * 6, &rn))) return rc;
    WeightSrc src{h->logw, nullptr, 0};
    SelArgs sa = select_args(h, src);
    sa.G = h->n_peers;
    sa.rank = h->my_peer;
    sa.peers = h->peer_shards;
    sa.coff_out = h->c_off;
    sa.rough = make_noise(h, rough_var, nullptr, rough_var ? 1 : 0, h->step_rough, kStreamRoughen);
    sa.rough_noise = rn;
    sa.anc64 = ancestors_out_dev;
    sa.dst = h->st[1 - h->cur];
    {
        Prof pr(h, AUVPF_K_SELECT_GATHER);
        CU(launch_select<kSelPeers>(h, grid1(h->M, kSelTile, 1), sa));
        LAUNCHED(h, AUVPF_K_SELECT_GATHER, 1);
        CU(cudaGetLastError());
    }
    h->have_global = true;
    h->n_ranks_global = h->n_peers;
    h->my_rank_global = h->my_peer;
    h->cur = 1 - h->cur;
    if (rough_var) h->step_rough++;
    return AUVPF_OK;
}
}   // namespace

int auvpf_shard_resample(auvpf_handle *h, double u, const double *rough_var, const float *rough_noise,
                         int64_t *ancestors_out_dev)
{
    int rc = shard_check(h);
    if (rc) return rc;
    if ((rc = shard_stage_max(h, /*separate_post=*/false, u))) return rc;
    if ((rc = shard_stage_weights(h, /*post_inside=*/true))) return rc;
    return shard_stage_select(h, rough_var, rough_noise, ancestors_out_dev);
}

// All ranks of ONE process (several GPUs driven by one process, or several handles on one GPU): the same
// resample, enqueued stage by stage over the ranks -- every rank's post of a stage is in its stream before any
// rank's kernel that waits for that stage is enqueued.  Kernels of different streams are not guaranteed to run
// concurrently (streams can share a hardware queue), and with this order a waiting kernel never sits in front
// of the post it waits for, whatever the mapping.
int auvpf_group_resample(auvpf_handle *const *handles, int32_t n_ranks, double u, const double *rough_var,
                         int64_t *const *ancestors_out_dev)
{
    if (!handles || n_ranks < 1 || n_ranks > kMaxPeers) return fail(AUVPF_ERR_BADARG, "bad handle list");
    int rc;
    for (int r = 0; r < n_ranks; ++r) {
        if ((rc = shard_check(handles[r]))) return rc;
        if (handles[r]->n_peers != n_ranks || handles[r]->my_peer != r)
            return fail(AUVPF_ERR_BADARG, "handles[%d] is not rank %d of %d", r, r, n_ranks);
    }
    for (int r = 0; r < n_ranks; ++r)
        if ((rc = shard_stage_max(handles[r], /*separate_post=*/true, u))) return rc;
    for (int r = 0; r < n_ranks; ++r)
        if ((rc = shard_stage_weights(handles[r], /*post_inside=*/false))) return rc;
    for (int r = 0; r < n_ranks; ++r)
        if ((rc = shard_stage_select(handles[r], rough_var, nullptr, ancestors_out_dev ? ancestors_out_dev[r] : nullptr))) return rc;
    return AUVPF_OK;
}

// ---- the estimate across ranks through the mailbox: no NCCL call, one read-back ------------------------------------
namespace {
int est_stage_post(auvpf_handle *h)
{
    Guard g(h);
    if (!h->est_all) CU(cudaMalloc(&h->est_all, sizeof(double) * 16 * kMaxPeers));
    h->est_epoch++;
    Prof pr(h, AUVPF_K_ESTIMATE);
    CU(cudaMemsetAsync(h->sums, 0, sizeof(double) * 16, h->stream));
    const int blocks = (int)std::min<int64_t>((h->M + kThreads - 1) / kThreads, 592);
    k_pose_stats<<<dim3(blocks, 1), kThreads, 0, h->stream>>>(h->st[h->cur], h->M, h->sums);
    MailboxView mb = h->mbox;
    mb.epoch = h->est_epoch;
    k_est_post<<<1, 32, 0, h->stream>>>(mb, (int)(h->est_epoch & 1u), h->sums);
    LAUNCHED(h, AUVPF_K_ESTIMATE, 2);
    CU(cudaGetLastError());
    return AUVPF_OK;
}
int est_stage_wait(auvpf_handle *h)
{
    Guard g(h);
    MailboxView mb = h->mbox;
    mb.epoch = h->est_epoch;
    k_est_wait<<<1, 32, 0, h->stream>>>(mb, (int)(h->est_epoch & 1u), h->est_all);
    LAUNCHED(h, AUVPF_K_ESTIMATE, 1);
    CU(cudaGetLastError());
    return AUVPF_OK;
}
int est_read(auvpf_handle *h, double *sums_all_out)
{
    Guard g(h);
    CU(cudaMemcpyAsync(sums_all_out, h->est_all, sizeof(double) * 16 * h->n_peers, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return AUVPF_OK;
}
}  // namespace

int auvpf_shard_pose_sums(auvpf_handle *h, double *sums_all_out)
{
    if (!h || !sums_all_out) return fail(AUVPF_ERR_BADARG, "null argument");
    int rc;
    if ((rc = shard_check(h))) return rc;
    if (h->F != 1) return fail(AUVPF_ERR_BADARG, "one filter per shard");
    if ((rc = est_stage_post(h))) return rc;
    if ((rc = est_stage_wait(h))) return rc;
    return est_read(h, sums_all_out);
}

int auvpf_group_pose_sums(auvpf_handle *const *handles, int32_t n_ranks, double *sums_all_out)
{
    if (!handles || !sums_all_out || n_ranks < 1 || n_ranks > kMaxPeers) return fail(AUVPF_ERR_BADARG, "bad handle list");
    int rc;
    for (int r = 0; r < n_ranks; ++r) {
        if ((rc = shard_check(handles[r]))) return rc;
        if (handles[r]->n_peers != n_ranks || handles[r]->my_peer != r)
            return fail(AUVPF_ERR_BADARG, "handles[%d] is not rank %d of %d", r, r, n_ranks);
    }
    for (int r = 0; r < n_ranks; ++r)
        if ((rc = est_stage_post(handles[r]))) return rc;
    for (int r = 0; r < n_ranks; ++r)
        if ((rc = est_stage_wait(handles[r]))) return rc;
    for (int r = 0; r < n_ranks; ++r)                      // every rank holds the same table: [r][G][16] out
        if ((rc = est_read(handles[r], sums_all_out + (size_t)r * 16 * n_ranks))) return rc;
    return AUVPF_OK;
}

int auvpf_shard_degenerate(auvpf_handle *h, int32_t *flag)
{
    if (!h || !flag) return fail(AUVPF_ERR_BADARG, "null argument");
    if (!h->n_peers) return fail(AUVPF_ERR_STATE, "attach the peers first");
    Guard g(h);
    int v = 0;
    CU(cudaMemcpyAsync(&v, h->shard_degenerate, sizeof v, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *flag = v ? 1 : 0;
    return AUVPF_OK;
}

int auvpf_shard_load_roughen(auvpf_handle *h, const void *records, int64_t n, const double *rough_var,
                             const float *rough_noise)
{
    if (!h || !records) return fail(AUVPF_ERR_BADARG, "null argument");
    if (n != h->N) return fail(AUVPF_ERR_BADARG, "expected %lld records, got %lld", (long long)h->N, (long long)n);
    Guard g(h);
    const Record *rd = nullptr;
    int rc = stage_in(h, S_IN0, static_cast<const Record *>(records), (size_t)n, &rd);
    if (rc) return rc;
    const float *rn = nullptr;
    if (rough_noise && (rc = stage_in(h, S_IN1, rough_noise, (size_t)h->N * 6, &rn))) return rc;
    NoiseParams rough = make_noise(h, rough_var, nullptr, rough_var ? 1 : 0, h->step_rough, kStreamRoughen);
    {
        Prof pr(h, AUVPF_K_SELECT_GATHER);
        k_load_records<<<grid1(n, kThreads, 1), kThreads, 0, h->stream>>>(rd, n, h->st[1 - h->cur], rough, rn);
        LAUNCHED(h, AUVPF_K_SELECT_GATHER, 1);
        CU(cudaGetLastError());
    }
    h->cur = 1 - h->cur;
    if (rough_var) h->step_rough++;
    return AUVPF_OK;
}

int auvpf_shard_load(auvpf_handle *h, const void *records, int64_t n)
{
    return auvpf_shard_load_roughen(h, records, n, nullptr, nullptr);
}

// ------------------------------------------------------------------------------------------------
int auvpf_profile_enable(auvpf_handle *h, int32_t on)
{
    if (!h) return fail(AUVPF_ERR_BADARG, "null handle");
    h->prof = on != 0;
    return AUVPF_OK;
}

int auvpf_profile_read(auvpf_handle *h, double ms[AUVPF_K_COUNT], int64_t launches[AUVPF_K_COUNT], int32_t reset)
{
    if (!h) return fail(AUVPF_ERR_BADARG, "null handle");
    Guard g(h);
    CU(cudaStreamSynchronize(h->stream));
    for (auto &e : h->prof_events) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, e.a, e.b) == cudaSuccess) h->prof_ms[e.klass] += (double)t;
        h->prof_pool.push_back(e);
    }
    h->prof_events.clear();
    cudaGetLastError();
    for (int k = 0; k < AUVPF_K_COUNT; ++k) {
        if (ms) ms[k] = h->prof_ms[k];
        if (launches) launches[k] = h->prof_launches[k];
        if (reset) {
            h->prof_ms[k] = 0.0;
            h->prof_launches[k] = 0;
        }
    }
    return AUVPF_OK;
}

}   // extern "C"
