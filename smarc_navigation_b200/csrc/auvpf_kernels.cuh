// auvpf_kernels.cuh -- sm_100a kernels of the particle-filter cycle.
//
// Layout in HBM: structure-of-arrays particle state,
// x,y float64 (so a 2 km map keeps 1e-4 m), z,roll,pitch,yaw float32 = 32 B per
// particle, double-buffered for the resampling gather; float32 log-weights;
// uint64 integer CDF.  F filters x M particles, row f*M + i; grids are
// (tiles over M, F) so no block straddles two filters.
#pragma once
#include <cuda_runtime.h>

#include "auvpf_fan.h"
#include "auvpf_math.cuh"

namespace auvpf {

struct StateView {
    double *x, *y;
    float *z, *roll, *pitch, *yaw;
};

struct Record {   // 32-byte exchange record of the multi-GPU ancestor exchange
    double x, y;
    float z, roll, pitch, yaw;
};

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int float_to_ordered(float f)
{
    int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }
__device__ __forceinline__ long long double_to_ordered(double d)
{
    long long b = __double_as_longlong(d);
    return b >= 0 ? b : b ^ 0x7FFFFFFFFFFFFFFFll;
}
__device__ __forceinline__ double ordered_to_double(long long k)
{
    return __longlong_as_double(k >= 0 ? k : k ^ 0x7FFFFFFFFFFFFFFFll);
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per-filter maximum of the log-weights, published by the kernel that produces them so that
// normalisation needs no separate max pass.  float32 in the order-preserving int encoding;
// reset value kWmaxReset (a 0x80 byte fill) decodes to -3.39e38: "no finite weight seen".
constexpr int kWmaxReset = (int)0x80808080;
constexpr float kWmaxNone = -3.0e38f;

__device__ __forceinline__ void warp_publish_max(float v, bool valid, int *wmax_f32)
{
    int k = (valid && v == v) ? float_to_ordered(v) : kWmaxReset;
    k = __reduce_max_sync(0xffffffffu, k);
    if ((threadIdx.x & 31) == 0 && k > kWmaxReset) atomicMax(wmax_f32, k);
}

// ---------------------------------------------------------------------------------------------
// add_noise (auv_particle.py:32-36): pose = base + sigma * n.  `set_mean` != nullptr makes it the
// particle initialiser (pose = mean + sigma n); otherwise roughening of the current state.
struct NoiseParams {
    float sigma[6];
    double mean[6];
    int mode;                // 0 = none, 1 = pose += sigma n, 2 = pose = mean + sigma n
    unsigned need;           // bit c: sigma[c] != 0 (normals multiplied by zero are not generated)
    uint64_t seed;
    uint32_t step;
    uint32_t stream;
    int64_t global_offset;   // global id of local particle 0
    int64_t global_stride;   // global particles per filter (ids: f * global_stride + global_offset + i)
};

__device__ __forceinline__ void six_normals(const NoiseParams &p, int64_t gid, const float *noise, int64_t row,
                                            float n[6])
{
    if (noise) {
#pragma unroll
        for (int c = 0; c < 6; ++c) n[c] = __ldg(noise + row * 6 + c);
    } else {
        pose_normals(p.seed, gid, p.step, p.stream, p.need, n);
    }
}

__global__ void __launch_bounds__(kThreads) k_add_noise(StateView s, int64_t M, NoiseParams p, const float *noise)
{
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= M) return;
    int f = blockIdx.y;
    int64_t row = (int64_t)f * M + i;
    int64_t gid = (int64_t)f * p.global_stride + p.global_offset + i;
    float n[6];
    six_normals(p, gid, noise, row, n);
    if (p.mode == 2) {
        s.x[row] = p.mean[0] + (double)(p.sigma[0] * n[0]);
        s.y[row] = p.mean[1] + (double)(p.sigma[1] * n[1]);
        s.z[row] = (float)p.mean[2] + p.sigma[2] * n[2];
        s.roll[row] = (float)p.mean[3] + p.sigma[3] * n[3];
        s.pitch[row] = (float)p.mean[4] + p.sigma[4] * n[4];
        s.yaw[row] = (float)p.mean[5] + p.sigma[5] * n[5];
    } else {
        s.x[row] += (double)(p.sigma[0] * n[0]);
        s.y[row] += (double)(p.sigma[1] * n[1]);
        s.z[row] += p.sigma[2] * n[2];
        s.roll[row] += p.sigma[3] * n[3];
        s.pitch[row] += p.sigma[4] * n[4];
        s.yaw[row] += p.sigma[5] * n[5];
    }
}

// ---------------------------------------------------------------------------------------------
// Propagate (a2).  4 particles per thread, 128-bit loads/stores; HBM-bound:
// reads x,y,yaw (20 B), writes x,y,z,roll,pitch,yaw (32 B) per particle.
__device__ __forceinline__ void predict_normals(uint64_t seed, int64_t gid, uint32_t step, const float *noise,
                                                int64_t row, float &n0, float &n1, float &n5)
{
    if (noise) {
        n0 = __ldg(noise + row * 6 + 0);
        n1 = __ldg(noise + row * 6 + 1);
        n5 = __ldg(noise + row * 6 + 5);
    } else {
        U4 r = philox_for(seed, gid, step, kStreamPredict, 0);
        float unused;
        box_muller(r.x, r.y, n0, n1);
        box_muller(r.z, r.w, n5, unused);
    }
}

// n_steps > 1 runs that many consecutive predicts in ONE pass over the state (the 50 Hz odometry
// messages between two pings): same arithmetic and RNG counters as n_steps separate launches, the
// state is read and written once.  params: (n_steps, F) rows, or nullptr for one broadcast odom (p0).
__global__ void __launch_bounds__(kThreads)
k_propagate(StateView s, int64_t M, PredictParams p0, const PredictParams *params, int n_steps, uint64_t seed,
            uint32_t step, int64_t global_offset, int64_t global_stride, const float *noise)
{
    const int f = blockIdx.y, F = gridDim.y;
    const int64_t base = (int64_t)f * M;
    const int64_t i0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
    if (i0 >= M) return;
    const int64_t gid0 = (int64_t)f * global_stride + global_offset + i0;
    const int64_t noise_stride = (int64_t)F * M * 6;
    if (i0 + 3 < M && ((base + i0) & 3) == 0) {
        double2 xa = *reinterpret_cast<const double2 *>(s.x + base + i0);
        double2 xb = *reinterpret_cast<const double2 *>(s.x + base + i0 + 2);
        double2 ya = *reinterpret_cast<const double2 *>(s.y + base + i0);
        double2 yb = *reinterpret_cast<const double2 *>(s.y + base + i0 + 2);
        float4 yw = *reinterpret_cast<const float4 *>(s.yaw + base + i0);
        double x[4] = {xa.x, xa.y, xb.x, xb.y}, y[4] = {ya.x, ya.y, yb.x, yb.y};
        float w[4] = {yw.x, yw.y, yw.z, yw.w};
        float z = 0.f, r = 0.f, pt = 0.f;
        for (int t = 0; t < n_steps; ++t) {
            const PredictParams p = params ? params[(int64_t)t * F + f] : p0;
            const float *nz = noise ? noise + t * noise_stride : nullptr;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float n0, n1, n5;
                predict_normals(seed, gid0 + k, step + t, nz, base + i0 + k, n0, n1, n5);
                propagate_one(p, x[k], y[k], z, r, pt, w[k], n0, n1, n5);
            }
        }
        *reinterpret_cast<double2 *>(s.x + base + i0) = make_double2(x[0], x[1]);
        *reinterpret_cast<double2 *>(s.x + base + i0 + 2) = make_double2(x[2], x[3]);
        *reinterpret_cast<double2 *>(s.y + base + i0) = make_double2(y[0], y[1]);
        *reinterpret_cast<double2 *>(s.y + base + i0 + 2) = make_double2(y[2], y[3]);
        *reinterpret_cast<float4 *>(s.yaw + base + i0) = make_float4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<float4 *>(s.z + base + i0) = make_float4(z, z, z, z);
        *reinterpret_cast<float4 *>(s.roll + base + i0) = make_float4(r, r, r, r);
        *reinterpret_cast<float4 *>(s.pitch + base + i0) = make_float4(pt, pt, pt, pt);
    } else {
        for (int k = 0; k < 4 && i0 + k < M; ++k) {
            int64_t row = base + i0 + k;
            double x = s.x[row], y = s.y[row];
            float z = 0.f, r = 0.f, pt = 0.f, w = s.yaw[row];
            for (int t = 0; t < n_steps; ++t) {
                const PredictParams p = params ? params[(int64_t)t * F + f] : p0;
                const float *nz = noise ? noise + t * noise_stride : nullptr;
                float n0, n1, n5;
                predict_normals(seed, gid0 + k, step + t, nz, row, n0, n1, n5);
                propagate_one(p, x, y, z, r, pt, w, n0, n1, n5);
            }
            s.x[row] = x;
            s.y[row] = y;
            s.z[row] = z;
            s.roll[row] = r;
            s.pitch[row] = pt;
            s.yaw[row] = w;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// GPS update (a3, reference form): logw = -d^2 / (2 std^2) - log(2 pi std^2)
__global__ void __launch_bounds__(kThreads)
k_update_gps(StateView s, int64_t M, FrameParams fp, const double *gps_xy, float inv_2var, float lognorm,
             float *logw, int *wmax_f32)
{
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    int f = blockIdx.y;
    const bool active = i < M;
    int64_t row = (int64_t)f * M + (active ? i : 0);
    double x = s.x[row], y = s.y[row], z = (double)s.z[row];
    double px, py;
    if (fp.m2o_identity_r) {
        px = x + fp.m2o_t[0];
        py = y + fp.m2o_t[1];
    } else {
        px = fp.m2o_r[0] * x + fp.m2o_r[1] * y + fp.m2o_r[2] * z + fp.m2o_t[0];
        py = fp.m2o_r[3] * x + fp.m2o_r[4] * y + fp.m2o_r[5] * z + fp.m2o_t[1];
    }
    double dx = gps_xy[2 * f] - px, dy = gps_xy[2 * f + 1] - py;
    const float lw = (float)(-(dx * dx + dy * dy) * (double)inv_2var + (double)lognorm);
    if (active) logw[row] = lw;
    warp_publish_max(lw, active, wmax_f32 + f);
}

// ---------------------------------------------------------------------------------------------
// Fused raycast + beam-range log-likelihood (a3 + a4).  One warp per particle; each lane owns a
// chunk of ADJACENT beams and sweeps it outward from the fan centre (auvpf_math.cuh "Fan
// sweep"): the first beam of a chunk is a full hierarchical traversal, every later beam starts
// at the lateral distance its inner neighbour was found clear to and is resolved by a walk
// of one or two cells.  Log-likelihood by warp-shuffle reduction over the lanes' partial sums.
// The DEM is read through the read-only path; at the benchmark configs the footprint of a
// particle cloud is L1/L2 resident.
constexpr int kMbesWarps = 8;
constexpr int kExactThreshold = 12;   // run the postponed exact-cell tests once this many lanes wait

struct FanParams {
    float c0[3];      // fan centre direction, sensor frame
    int planar;       // beams form a sorted planar fan
    int chunk;        // beams per lane
    int lanes_first;  // lanes holding the first side of the fan (auvpf_fan.h)
    float a0[3];      // lateral unit vector, sensor frame (th > 0 leans towards +a0)
    int n1, sgn1;     // beams [0, n1) on the side of sign sgn1, [n1, B) on the other
};

// Warp-cooperative traversal driver: MUST be called by all 32 lanes of a converged warp.
// Every iteration all unresolved lanes do one uniform node step; lanes whose ray reached a cell
// that cannot be skipped park (pending) until enough lanes are parked -- or nobody else can
// make progress -- and then run the exact two-triangle test together.
__device__ __forceinline__ float raycast_warp(const DemView &m, const Ray &r, bool valid, bool interior)
{
    Trav s;
    float range = m.range_max;
    bool active = valid && trav_init(s, m, r, range, 0.f, interior);
    bool pending = false;
    while (__any_sync(0xffffffffu, active)) {
        if (active && !pending) {
            const int res = trav_step(s, m, r);
            if (res == kMiss) {
                range = m.range_max;
                active = false;
            } else if (res == kPending)
                pending = true;
        }
        const unsigned pm = __ballot_sync(0xffffffffu, pending);
        if (pm) {
            const unsigned wm = __ballot_sync(0xffffffffu, active && !pending);
            if (wm == 0u || __popc(pm) >= kExactThreshold) {
                if (pending) {
                    if (trav_exact(s, m, r, range)) active = false;
                    pending = false;
                }
            }
        }
    }
    return range;
}

// Beam tables staged in shared memory by both kernels:
//   sm[0..B) dx, [B..2B) dy, [2B..3B) dz (sensor frame), [3B..4B) measured range (NaN = not
//   needed), [4B..5B) |sin th|, [5B..6B) 1/|sin th|; then the per-lane order table (ints).
// Casts every beam of one particle (warp-cooperative, call with all 32 lanes) and hands
// (beam, range) of the needed beams to `sink`.
template <class Sink>
__device__ __forceinline__ void fan_cast(const DemView &dem, const FanParams &fan, const float *sm, int B,
                                         const int *s_ord, const double p[3], const double R[9], int lane, Sink sink)
{
    int ix0, iy0;
    float lx0, ly0;
    origin_cell(dem, p, ix0, iy0, lx0, ly0);
    const int C = fan.chunk;
    const bool sweep_ok = fan.planar && fan_sweep_valid(dem, p, R, fan.c0);     // warp-uniform
    if (!sweep_ok) {
        // no certified sweep for this particle: every beam is traversed from the sensor
        for (int k = 0; k < C; ++k) {
            const int o = s_ord[lane * C + k];
            const int b = o >= 0 ? (o & (kOrdFull - 1)) : 0;
            const bool need = o >= 0 && sm[3 * B + b] == sm[3 * B + b];
            Ray ray = make_ray(dem, p, ix0, iy0, lx0, ly0, R, sm[b], sm[B + b], sm[2 * B + b]);
            const float range = raycast_warp(dem, ray, need, false);
            if (need) sink(b, range);
        }
        return;
    }
    // ---- chunk starts: every lane traverses the innermost beam of its chunk from the sensor,
    // all 32 together through the cooperative driver.  (Chaining the starts outward through the
    // neighbouring lane was measured: 2 active lanes x a long sequential walk costs more warp
    // instructions than this one cooperative traversal.)
    float p_clear = 0.f;
    {
        const int o = s_ord[lane * C];
        const bool has = o >= 0;
        const int b = has ? (o & (kOrdFull - 1)) : 0;
        Ray ray = make_ray(dem, p, ix0, iy0, lx0, ly0, R, sm[b], sm[B + b], sm[2 * B + b]);
        const float range = raycast_warp(dem, ray, has, true);
        if (has) p_clear = range * sm[4 * B + b];
        if (has && sm[3 * B + b] == sm[3 * B + b]) sink(b, range);
    }
    // ---- sweep every lane's chunk outward
    for (int k = 1; k < C; ++k) {
        const int o = s_ord[lane * C + k];
        const int b = o >= 0 ? (o & (kOrdFull - 1)) : 0;
        if (o >= 0 && sm[3 * B + b] == sm[3 * B + b]) {
            Ray ray = make_ray(dem, p, ix0, iy0, lx0, ly0, R, sm[b], sm[B + b], sm[2 * B + b]);
            const float range = raycast_from(dem, ray, sweep_bound(p_clear, sm[5 * B + b]), true);
            p_clear = range * sm[4 * B + b];        // a miss (range_max) leaves the whole ray clear
            sink(b, range);
        }
    }
}

__device__ __forceinline__ void stage_fan_tables(float *sm, int *s_ord, int B, const FanParams &fan,
                                                 const float *beam_dirs, const float *beam_aux, const int *order)
{
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        sm[b] = beam_dirs[3 * b];
        sm[B + b] = beam_dirs[3 * b + 1];
        sm[2 * B + b] = beam_dirs[3 * b + 2];
        sm[4 * B + b] = beam_aux[b];
        sm[5 * B + b] = beam_aux[B + b];
    }
    for (int k = threadIdx.x; k < 32 * fan.chunk; k += blockDim.x) s_ord[k] = order[k];
}

__global__ void __launch_bounds__(kMbesWarps * 32)
k_update_mbes(StateView s, int64_t M, FrameParams fp, DemView dem, FanParams fan, const float *__restrict__ beam_dirs,
              const float *__restrict__ beam_aux, const int *__restrict__ order, const float *__restrict__ meas, int B,
              float range_min, float neg_half_inv_var, float log_norm_per_beam, float *__restrict__ logw,
              const int *__restrict__ flagged_list, const unsigned int *__restrict__ flagged_count,
              int *__restrict__ wmax_f32)
{
    extern __shared__ float sm[];
    int *s_ord = reinterpret_cast<int *>(sm + 6 * B);
    __shared__ int s_valid;
    const int f = blockIdx.y;
    // With a list (particles the profile kernel could not certify, usually none) the grid is small
    // and its warps walk the list; without one every particle of the filter is cast.
    const int64_t n_work = flagged_list ? (int64_t)(flagged_count[f] - (unsigned int)kWmaxReset) : M;
    if ((int64_t)blockIdx.x * kMbesWarps >= n_work) return;
    if (threadIdx.x == 0) s_valid = 0;
    stage_fan_tables(sm, s_ord, B, fan, beam_dirs, beam_aux, order);
    __syncthreads();
    int nv = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float r = meas[(size_t)f * B + b];
        bool ok = isfinite(r) && r >= range_min && r <= dem.range_max;
        sm[3 * B + b] = ok ? r : __int_as_float(0x7fc00000);     // invalid measured beam: no raycast needed
        nv += ok ? 1 : 0;
    }
    if (nv) atomicAdd(&s_valid, nv);
    __syncthreads();
    const float lognorm = -(float)s_valid * log_norm_per_beam;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t w = (int64_t)blockIdx.x * kMbesWarps + warp; w < n_work; w += (int64_t)gridDim.x * kMbesWarps) {
        const int64_t i = flagged_list ? (int64_t)flagged_list[(int64_t)f * M + w] : w;
        const int64_t row = (int64_t)f * M + i;
        double p[3];
        double R[9];
        particle_frame(fp, s.x[row], s.y[row], s.z[row], s.roll[row], s.pitch[row], s.yaw[row], p, R);
        float acc = 0.f;
        fan_cast(dem, fan, sm, B, s_ord, p, R, lane, [&](int b, float range) {
            const float e = sm[3 * B + b] - range;
            acc = fmaf(e, e, acc);
        });
        acc = warp_sum(acc);
        if (lane == 0) {
            const float lw = fmaf(neg_half_inv_var, acc, lognorm);
            logw[row] = lw;
            if (lw == lw) {
                const int k = float_to_ordered(lw);
                if (k > kWmaxReset) atomicMax(wmax_f32 + f, k);
            }
        }
        __syncwarp();
    }
}

// Parity probe: expected ranges for explicit poses, (n, B) float32; same traversal as above.
__global__ void __launch_bounds__(kMbesWarps * 32)
k_raycast_poses(const double *__restrict__ poses, int64_t n, FrameParams fp, DemView dem, FanParams fan,
                const float *__restrict__ beam_dirs, const float *__restrict__ beam_aux,
                const int *__restrict__ order, int B, float *__restrict__ out,
                const unsigned char *__restrict__ only_flagged)
{
    extern __shared__ float sm[];
    int *s_ord = reinterpret_cast<int *>(sm + 6 * B);
    stage_fan_tables(sm, s_ord, B, fan, beam_dirs, beam_aux, order);
    for (int b = threadIdx.x; b < B; b += blockDim.x) sm[3 * B + b] = 0.f;      // every beam is needed
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * kMbesWarps + warp;
    if (i >= n) return;
    if (only_flagged && !only_flagged[i]) return;
    const double *ps = poses + 6 * i;
    double p[3];
    double R[9];
    particle_frame(fp, ps[0], ps[1], (float)ps[2], (float)ps[3], (float)ps[4], (float)ps[5], p, R);
    fan_cast(dem, fan, sm, B, s_ord, p, R, lane, [&](int b, float range) { out[(size_t)i * B + b] = range; });
}

// ---------------------------------------------------------------------------------------------
// Profile-merge kernels (auvpf_math.cuh "Profile merge"): ONE THREAD per particle marches the
// terrain profile of each side of the fan once and resolves all its beams on the way --
// O(cells + beams) per particle instead of O(beams x cells per ray).  Lanes of a warp are
// neighbouring particles doing the same march, so the warp stays converged and the DEM
// footprint is shared in L1.  Particles that cannot be certified (near the map border, centre
// beam missed, a triangle steeper than the fan tilt allows) are flagged and recast by the
// per-ray kernel above.
// Block shapes of the profile kernel at C3 (ms per update, converged cloud).  Round 1, 245 instructions per march
// step: (128 thr, 5 blocks/SM, 96 regs) 4.60; (128, 6, 80 regs) 4.23; (192, 3, 96 regs) 4.24; (256, 2, 109 regs)
// 4.44.  Round 2, after the march step went to ~195 instructions: (192, 3, 96 regs) 3.62; (192, 4, 80 regs)
// 3.29; (256, 3, 80) 3.29; (128, 6, 80) 3.30; (128, 7, 72 regs, 24 B of spills) 3.23; (128, 8, 64 regs) 3.31;
// (256, 4, 64) 3.30 -- the leaner loop is a shorter dependent chain per warp, so warps per scheduler (7 instead of
// 4.5) now buy more than registers do.  A cloud that does not fill the machine (fewer than ~4 warps per scheduler)
// still runs fastest with all the registers it wants (C2: 0.118 ms at (256, 2) against 0.142 at (128, 6)).
#ifndef AUVPF_PROF_T          // (kernel-tuning builds: make EXTRA="-DAUVPF_PROF_T=128 -DAUVPF_PROF_MINB=6" OUT=../lib_x.so)
#define AUVPF_PROF_T 128
#define AUVPF_PROF_MINB 7
#endif
constexpr int kProfThreads = AUVPF_PROF_T, kProfMinBlocks = AUVPF_PROF_MINB;
constexpr int kProfThreadsSmall = 256, kProfMinBlocksSmall = 2;
constexpr int kProfThreadsBorder = 192, kProfMinBlocksBorder = 3;   // the map-border variant wants its 96 registers
constexpr int64_t kProfSmallLimit = 148 * 4 * 4 * 32;      // particles below which the small-cloud shape is used

// shared memory: tan (double B) | sec (double B) | aux (BeamAux B + 2: sec, reach, measured range, tan; a sentinel at either end)
// beam_tan = [tan | sec] (2B doubles), beam_aux[0..B) = |lateral component|
__device__ __forceinline__ FanTables make_tables(const FanParams &fan, int B, const double *s_tan, const double *s_sec,
                                                 const BeamAux *s_aux)
{
    FanTables T;
    T.tan_t = s_tan;
    T.sec_t = s_sec;
    T.aux = s_aux;
    T.B = B;
    T.n1 = fan.n1;
    T.sgn1 = fan.sgn1;
    // outermost beam of a side: beam 0 of [0, n1), beam B - 1 of [n1, B)
    T.s_lim[0] = fan.n1 > 0 ? (1.0f - kDenPrecise) / fmaxf(s_aux[0].tn, 1.0e-6f) : 0.f;
    T.s_lim[1] = B > fan.n1 ? (1.0f - kDenPrecise) / fmaxf(s_aux[B - 1].tn, 1.0e-6f) : 0.f;
    for (int a = 0; a < 3; ++a) {
        T.c0[a] = fan.c0[a];
        T.a0[a] = fan.a0[a];
    }
    return T;
}

// kBorder = false: every particle of the filter (grid over M); a sensor whose reach leaves the map -- or that
// cannot be certified for another reason -- is put on `flagged_list`.  kBorder = true: the second pass over
// such a list (`in_list`, `in_count`; the blocks stride over it), with the map-border handling compiled in;
// what even that cannot certify goes on the next list, for the per-ray kernel.
template <int kT, int kMinB, bool kBorder>
__global__ void __launch_bounds__(kT, kMinB)
k_update_mbes_profile(StateView s, int64_t M, FrameParams fp, DemView dem, FanParams fan,
                      const double *__restrict__ beam_tan, const float *__restrict__ beam_aux,
                      const float *__restrict__ meas, int B, float range_min, float neg_half_inv_var,
                      float log_norm_per_beam, float *__restrict__ logw, const int *__restrict__ in_list,
                      const unsigned int *__restrict__ in_count, int *__restrict__ flagged_list,
                      unsigned int *__restrict__ flagged_count, int *__restrict__ wmax_f32)
{
    extern __shared__ double smd[];
    double *s_tan = smd, *s_sec = smd + B;
    BeamAux *s_aux = reinterpret_cast<BeamAux *>(smd + 2 * B) + 1;        // entries -1 and B: the sentinels (beam_sentinel)
    __shared__ int s_valid;
    const int f = blockIdx.y;
    const int64_t n_work = kBorder ? (int64_t)(in_count[f] - (unsigned int)kWmaxReset) : M;
    if ((int64_t)blockIdx.x * kT >= n_work) return;
    if (threadIdx.x == 0) s_valid = 0;
    __syncthreads();
    int nv = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        s_tan[b] = beam_tan[b];
        s_sec[b] = beam_tan[B + b];
        const float r = meas[(size_t)f * B + b];
        const bool ok = isfinite(r) && r >= range_min && r <= dem.range_max;
        BeamAux a;
        // a beam without a valid measurement: sec = 0 and meas = 0 -- its range comes out as 0 * (q at the hit) = 0,
        // its error as 0 - 0, and the sum takes it like any other beam (no test per beam in the march's hot loop)
        a.sec = ok ? (float)beam_tan[B + b] : 0.f;
        a.reach = beam_aux[b] * dem.range_max;
        a.meas = ok ? r : 0.f;
        a.tn = (float)beam_tan[b];
        s_aux[b] = a;
        nv += ok ? 1 : 0;
    }
    if (threadIdx.x == 0) s_aux[-1] = s_aux[B] = beam_sentinel(dem.range_max);
    if (nv) atomicAdd(&s_valid, nv);
    __syncthreads();
    const FanTables T = make_tables(fan, B, s_tan, s_sec, s_aux);
    const float lognorm = -(float)s_valid * log_norm_per_beam;
    // (whole warps enter every trip: the march is a warp-uniform loop, lanes past the end walk through inactive)
    const int64_t n_round = (n_work + 31) & ~(int64_t)31;
    for (int64_t w = (int64_t)blockIdx.x * kT + threadIdx.x; w < n_round; w += kBorder ? (int64_t)gridDim.x * kT : n_round) {
        const bool active = w < n_work;
        const int64_t i = kBorder ? (active ? (int64_t)in_list[(int64_t)f * M + w] : 0) : w;
        const int64_t row = (int64_t)f * M + (active ? i : 0);
        double p[3];
        double R[9];
        particle_frame(fp, s.x[row], s.y[row], s.z[row], s.roll[row], s.pitch[row], s.yaw[row], p, R);
        float acc = 0.f, acc_exit = 0.f;    // acc_exit: beams finished ray by ray at the map border (lives in local memory)
        const bool ok = profile_particle<kBorder>(
            dem, T, p, R, active,
            [&acc](int, float range, float rm) {
                const float e = rm - range;
                acc = fmaf(e, e, acc);
            },
            [&acc_exit](int, float range, float rm) {
                const float e = rm - range;
                acc_exit = fmaf(e, e, acc_exit);
            });
        if (kBorder) acc += acc_exit;
        const float lw = fmaf(neg_half_inv_var, acc, lognorm);
        if (active) {
            if (ok) logw[row] = lw;
            else flagged_list[(int64_t)f * M + (atomicAdd(flagged_count + f, 1u) - (unsigned int)kWmaxReset)] = (int)i;
        }
        warp_publish_max(lw, active && ok, wmax_f32 + f);
    }
}

__global__ void __launch_bounds__(kProfThreads)
k_raycast_poses_profile(const double *__restrict__ poses, int64_t n, FrameParams fp, DemView dem, FanParams fan,
                        const double *__restrict__ beam_tan, const float *__restrict__ beam_aux, int B,
                        float *__restrict__ out, unsigned char *__restrict__ flags)
{
    extern __shared__ double smd[];
    double *s_tan = smd, *s_sec = smd + B;
    BeamAux *s_aux = reinterpret_cast<BeamAux *>(smd + 2 * B) + 1;        // entries -1 and B: the sentinels (beam_sentinel)
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        s_tan[b] = beam_tan[b];
        s_sec[b] = beam_tan[B + b];
        BeamAux a;
        a.sec = (float)beam_tan[B + b];
        a.reach = beam_aux[b] * dem.range_max;
        a.meas = 0.f;
        a.tn = (float)beam_tan[b];
        s_aux[b] = a;
    }
    if (threadIdx.x == 0) s_aux[-1] = s_aux[B] = beam_sentinel(dem.range_max);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kProfThreads + threadIdx.x;
    const bool active = i < n;
    const double *ps = poses + 6 * (active ? i : 0);
    double p[3];
    double R[9];
    particle_frame(fp, ps[0], ps[1], (float)ps[2], (float)ps[3], (float)ps[4], (float)ps[5], p, R);
    const FanTables T = make_tables(fan, B, s_tan, s_sec, s_aux);
    float *row = out + (size_t)(active ? i : 0) * B;
    const bool ok = profile_particle<true>(dem, T, p, R, active, [row](int b, float range, float) { row[b] = range; },
                                           [row](int b, float range, float) { row[b] = range; });
    if (active) flags[i] = ok ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------
// Max-mip pyramid.  Level 1 block (bi,bj) = max over vertices [2bi, 2bi+2] x [2bj, 2bj+2] (closed
// footprint of a 2x2-cell block); level L+1 = max of 2x2 children.
// Copy an H x W DEM into the padded layout: (H+2) x (W+2) with a replicated one-vertex border.
__global__ void k_pad_dem(const float *src, int H, int W, float *dst)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x - 1, i = blockIdx.y * blockDim.y + threadIdx.y - 1;
    if (j > W || i > H) return;
    int si = min(max(i, 0), H - 1), sj = min(max(j, 0), W - 1);
    dst[(size_t)(i + 1) * (W + 2) + (j + 1)] = src[(size_t)si * W + sj];
}

__global__ void k_mip_level1(const float *z, int pitch, int H, int W, float *out, int oH, int oW)
{
    int bj = blockIdx.x * blockDim.x + threadIdx.x, bi = blockIdx.y * blockDim.y + threadIdx.y;
    if (bj >= oW || bi >= oH) return;
    float m = -INFINITY;
    for (int di = 0; di <= 2; ++di)
        for (int dj = 0; dj <= 2; ++dj) {
            int i = min(2 * bi + di, H - 1), j = min(2 * bj + dj, W - 1);
            m = fmaxf(m, z[(size_t)i * pitch + j]);
        }
    out[(size_t)bi * oW + bj] = m;
}

__global__ void k_mip_down(const float *in, int iH, int iW, float *out, int oH, int oW)
{
    int bj = blockIdx.x * blockDim.x + threadIdx.x, bi = blockIdx.y * blockDim.y + threadIdx.y;
    if (bj >= oW || bi >= oH) return;
    float m = -INFINITY;
    for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj) {
            int i = 2 * bi + di, j = 2 * bj + dj;
            if (i < iH && j < iW) m = fmaxf(m, in[(size_t)i * iW + j]);
        }
    out[(size_t)bi * oW + bj] = m;
}

// Map ingestion: rasterise a triangle mesh onto the DEM's vertex grid.  One thread per
// triangle walks the grid vertices of its bounding box; a vertex inside the triangle (edges
// included) takes the interpolated height, the highest surface wins (atomicMax on the
// order-preserving int encoding).  Vertices no triangle covers keep `fill`.
constexpr int kUncovered = (int)0x80000000;     // below the encoding of every float, -inf included

__global__ void k_fill_ordered(int *z, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = kUncovered;
}

__global__ void k_rasterise_mesh(const float *__restrict__ verts, const int *__restrict__ faces, int64_t n_faces,
                                 int H, int W, double x0, double y0, double inv_res, int *z_ordered)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_faces) return;
    double gx[3], gy[3], gz[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float *v = verts + 3 * (int64_t)faces[3 * t + k];
        gx[k] = ((double)v[0] - x0) * inv_res;
        gy[k] = ((double)v[1] - y0) * inv_res;
        gz[k] = (double)v[2];
    }
    const double det = (gy[1] - gy[2]) * (gx[0] - gx[2]) + (gx[2] - gx[1]) * (gy[0] - gy[2]);
    if (det == 0.0) return;                                   // degenerate in plan view
    const int j0 = max(0, (int)ceil(fmin(gx[0], fmin(gx[1], gx[2])) - 1e-9));
    const int j1 = min(W - 1, (int)floor(fmax(gx[0], fmax(gx[1], gx[2])) + 1e-9));
    const int i0 = max(0, (int)ceil(fmin(gy[0], fmin(gy[1], gy[2])) - 1e-9));
    const int i1 = min(H - 1, (int)floor(fmax(gy[0], fmax(gy[1], gy[2])) + 1e-9));
    const double eps = 1e-9;
    for (int i = i0; i <= i1; ++i)
        for (int j = j0; j <= j1; ++j) {
            const double l0 = ((gy[1] - gy[2]) * ((double)j - gx[2]) + (gx[2] - gx[1]) * ((double)i - gy[2])) / det;
            const double l1 = ((gy[2] - gy[0]) * ((double)j - gx[2]) + (gx[0] - gx[2]) * ((double)i - gy[2])) / det;
            const double l2 = 1.0 - l0 - l1;
            if (l0 < -eps || l1 < -eps || l2 < -eps) continue;
            const float zz = (float)(l0 * gz[0] + l1 * gz[1] + l2 * gz[2]);
            atomicMax(z_ordered + (int64_t)i * W + j, float_to_ordered(zz));
        }
}

__global__ void k_unorder(int *z, int64_t n, float fill)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = __float_as_int(z[i] == kUncovered ? fill : ordered_to_float(z[i]));
}

// largest triangle gradient norm of the DEM (two triangles per cell, diagonal (i,j)->(i+1,j+1))
__global__ void k_max_slope(const float *z, int pitch, int H, int W, float inv_res, int *out_ordered)
{
    float m = 0.f;
    const int64_t n = (int64_t)(H - 1) * (W - 1);
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        int i = (int)(c / (W - 1)), j = (int)(c % (W - 1));
        const float *p = z + (size_t)i * pitch + j;
        float z00 = p[0], z01 = p[1], z10 = p[pitch], z11 = p[pitch + 1];
        float g1 = hypotf(z01 - z00, z11 - z01), g2 = hypotf(z11 - z10, z10 - z00);
        m = fmaxf(m, fmaxf(g1, g2) * inv_res);
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(out_ordered, float_to_ordered(m));
}

// the same per block of kSlopeBlock x kSlopeBlock cells (DemView::slope_blk): one thread block per slope block
__global__ void k_slope_blocks(const float *z, int pitch, int H, int W, float inv_res, float *out, int sbW)
{
    __shared__ float s_m[8];
    const int i0 = blockIdx.y * kSlopeBlock, j0 = blockIdx.x * kSlopeBlock;
    float m = 0.f;
    for (int c = threadIdx.x; c < kSlopeBlock * kSlopeBlock; c += blockDim.x) {
        const int i = i0 + c / kSlopeBlock, j = j0 + c % kSlopeBlock;
        if (i >= H - 1 || j >= W - 1) continue;
        const float *p = z + (size_t)i * pitch + j;
        const float z00 = p[0], z01 = p[1], z10 = p[pitch], z11 = p[pitch + 1];
        m = fmaxf(m, fmaxf(hypotf(z01 - z00, z11 - z01), hypotf(z11 - z10, z10 - z00)) * inv_res);
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) m = fmaxf(m, s_m[k]);
        out[(size_t)blockIdx.y * sbW + blockIdx.x] = m * 1.0001f + 1e-6f;       // (the margin max_slope carries)
    }
}

__global__ void k_max_f32(const float *in, int64_t n, int *out_ordered)
{
    float m = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, in[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(out_ordered, float_to_ordered(m));
}

// ---------------------------------------------------------------------------------------------
// Weight source of a resample:
//   kind 0: float32 log-weights (the update kernels' output)
//   kind 1: integer weights a_i = floor(w_i / max w * 2^60) already in memory -- float64 LINEAR weights (the
//           reference's resample(weights) argument) are converted by a pre-pass (k_linear_to_a), so that their
//           float64 divisions stay out of the two resample kernels
struct WeightSrc {
    const float *logw;
    const unsigned long long *apre;
    int kind;
};

// float32 log-weights that did not come from an update kernel (auvpf_set_logw): max pass
__global__ void __launch_bounds__(kThreads) k_weight_max_f32(const float *logw, int64_t M, int *wmax_f32)
{
    const int f = blockIdx.y;
    float m = -INFINITY;
    bool any = false;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < M; i += (int64_t)gridDim.x * kThreads) {
        const float v = logw[(int64_t)f * M + i];
        if (v == v) {
            m = fmaxf(m, v);
            any = true;
        }
    }
    warp_publish_max(m, any, wmax_f32 + f);
}

// ---------------------------------------------------------------------------------------------
// Single-pass decoupled-lookback inclusive scan (Merrill & Garland) of uint64 values.
// One 64-bit status word per tile: [63:62] flag (1 = tile aggregate, 2 = inclusive prefix),
// [61:0] value, so value and flag travel in one store and no fence is needed.
constexpr int kScanItems = 8;
constexpr int kScanTile = kThreads * kScanItems;
constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagPrefix = 2ull << 62, kValMask = (1ull << 62) - 1;

template <class ValueFn>
__device__ __forceinline__ void scan_tile_body(ValueFn value, int64_t M, unsigned long long *__restrict__ out,
                                               volatile unsigned long long *status, unsigned int *tile_counter,
                                               int tiles_per_filter, unsigned long long *totals)
{
    __shared__ unsigned int s_tile;
    __shared__ unsigned long long s_warp[kThreads / 32];
    __shared__ unsigned long long s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const unsigned int gtile = s_tile;
    const int f = gtile / tiles_per_filter, tile = gtile % tiles_per_filter;
    const int64_t base = (int64_t)f * M;
    const int64_t i0 = (int64_t)tile * kScanTile + (int64_t)threadIdx.x * kScanItems;

    unsigned long long v[kScanItems];
    unsigned long long tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (i0 + k < M) ? value(base + i0 + k, f, k) : 0ull;
        tsum += v[k];
    }
    // block exclusive scan of per-thread sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long warp_off = 0, block_total = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) {
        unsigned long long wv = s_warp[k];
        if (k < warp) warp_off += wv;
        block_total += wv;
    }
    unsigned long long excl = warp_off + incl - tsum;

    // publish the aggregate, then look back with the WHOLE block: thread k polls tile idx - k, so
    // one hop covers kThreads predecessors.  (With warp-0-only look-back and all tiles running
    // in lock-step the chain was tiles/32 dependent hops long; ncu showed the block waiting at the
    // barrier for most of the kernel.)
    {
        volatile unsigned long long *st = status + (size_t)f * tiles_per_filter;
        __shared__ int s_first[kThreads / 32];
        __shared__ unsigned long long s_part[kThreads / 32];
        if (threadIdx.x == 0) st[tile] = (tile == 0 ? kFlagPrefix : kFlagAgg) | block_total;
        unsigned long long run = 0;
        int idx = tile - 1;
        while (idx >= 0) {                                   // block-uniform
            const int j = idx - (int)threadIdx.x;
            unsigned long long wd = kFlagPrefix;             // before the filter's first tile: prefix 0
            if (j >= 0) {
                do {
                    wd = st[j];
                } while ((wd >> 62) == 0);
            }
            // nearest predecessor that already carries an inclusive prefix
            const unsigned pm = __ballot_sync(0xffffffffu, (wd >> 62) == 2);
            if (lane == 0) s_first[warp] = pm ? warp * 32 + (__ffs(pm) - 1) : kThreads;
            __syncthreads();
            int L = kThreads;
#pragma unroll
            for (int k = 0; k < kThreads / 32; ++k) L = min(L, s_first[k]);
            unsigned long long c = ((int)threadIdx.x <= L) ? (wd & kValMask) : 0ull;
            c = warp_sum_u64(c);
            if (lane == 0) s_part[warp] = c;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kThreads / 32; ++k) run += s_part[k];
            if (L < kThreads) break;
            idx -= kThreads;
        }
        if (threadIdx.x == 0) {
            if (tile != 0) st[tile] = kFlagPrefix | (run + block_total);
            s_prefix = run;
        }
    }
    __syncthreads();
    unsigned long long acc = s_prefix + excl;
    if (totals && threadIdx.x == 0 && tile == tiles_per_filter - 1) totals[f] = s_prefix + block_total;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        acc += v[k];
        if (i0 + k < M) out[base + i0 + k] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// Shapes shared with the select kernel (auvpf_resample.cuh): a block owns kSelTile consecutive output slots,
// a thread kSelItems adjacent ones.
constexpr int kSelThreads = 256;  // (128 threads x 2 groups, one wave of 7 blocks per SM, measured 30 % slower)
constexpr int kSelItems = 4;      // adjacent outputs a thread resolves together
constexpr int kSelTile = kSelThreads * kSelItems;
constexpr int kMaxPeers = 8;      // ranks of one sharded filter (one node)
#ifndef AUVPF_SEL_MINB
#define AUVPF_SEL_MINB 4          // resident blocks per SM the select kernel is compiled for (tuning builds override)
#endif

// ---------------------------------------------------------------------------------------------
// residual_resample (a6, resampling.py:27-76) on the fixed-point weights w_i = q_i 2^-52.
// q_i as k_exact_q left it in `cdf` (the deficit D of the old formulation is always 0 now: sum q == 2^52).
__device__ __forceinline__ unsigned long long q_from_cdf(const unsigned long long *cdf, int64_t base, int64_t i,
                                                        unsigned long long D, int64_t global_offset)
{
    return cdf[base + i] + ((unsigned long long)(global_offset + i) < D ? 1ull : 0ull);
}

// inclusive scan of the deterministic copy counts floor(N w_i) (:61); totals = k per filter
__global__ void __launch_bounds__(kThreads)
k_scan_copies(const unsigned long long *cdf, int64_t M, const unsigned long long *deficit, int64_t global_offset,
              unsigned long long *out, unsigned long long *status, unsigned int *tile_counter, int tiles_per_filter,
              unsigned long long *totals)
{
    auto value = [&](int64_t row, int f, int) -> unsigned long long {
        const int64_t base = (int64_t)f * M;
        return residual_copies(q_from_cdf(cdf, base, row - base, deficit[f], global_offset), M);
    };
    scan_tile_body(value, M, out, status, tile_counter, tiles_per_filter, totals);
}

// inclusive scan of the fractional parts N w_i - k_i on the 2^-38 grid; totals = R per filter
__global__ void __launch_bounds__(kThreads)
k_scan_fracs(const unsigned long long *cdf, int64_t M, const unsigned long long *deficit, int64_t global_offset,
             unsigned long long *out, unsigned long long *status, unsigned int *tile_counter, int tiles_per_filter,
             unsigned long long *totals)
{
    auto value = [&](int64_t row, int f, int) -> unsigned long long {
        const int64_t base = (int64_t)f * M;
        const unsigned long long q = q_from_cdf(cdf, base, row - base, deficit[f], global_offset);
        return residual_frac(q, residual_copies(q, M), M);
    };
    scan_tile_body(value, M, out, status, tile_counter, tiles_per_filter, totals);
}

// first j in [0, M) with key < c[j]; M if none
__device__ __forceinline__ int64_t first_above(const unsigned long long *c, int64_t M, unsigned long long key)
{
    int64_t lo = 0, hi = M;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (key < c[mid]) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// indexes[0:k] = the copies in index order (:62-66); indexes[k:N] = multinomial draw on the
// fractional parts (what :70-74 mean to do; the literal, numerically odd form is
// k_residual_strict).  uniforms: (F, M) float64, the first M - k of a row are consumed (:74),
// or nullptr for Philox (stream kStreamResample, call 1, id = f * M + t).
__global__ void __launch_bounds__(kThreads)
k_residual_fill(const unsigned long long *copies, const unsigned long long *fracs, int64_t M,
                const unsigned long long *K, const unsigned long long *R, const double *uniforms, uint64_t seed,
                uint32_t step, int fill_tail, int32_t *idx)
{
    const int f = blockIdx.y;
    const int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (o >= M) return;
    const int64_t base = (int64_t)f * M;
    const int64_t k = (int64_t)K[f];
    int64_t j;
    if (o < k) {
        j = first_above(copies + base, M, (unsigned long long)o);
    } else {
        if (!fill_tail) return;
        const int64_t t = o - k;
        const double uu = uniforms ? uniforms[base + t] : uniform53(philox_for(seed, base + t, step, kStreamResample, 1));
        j = first_above(fracs + base, M, residual_target(uu, R[f]));
    }
    idx[base + o] = (int32_t)min(j, M - 1);
}

// The LITERAL residual draw of the reference (:70-74), for lock-step parity at the reference's own
// particle counts: residual = w - num_copies (sic), residual /= sum(residual) with Python's
// sequential sum, sequential cumsum, cumsum[-1] = 1, then NumPy's left searchsorted with its
// key-to-key bounds hint on that non-monotone array.  One block per filter; the element-wise steps
// are parallel, the three order-dependent ones run on thread 0.  c: (F, M) float64 scratch.
__global__ void __launch_bounds__(kThreads)
k_residual_strict(const unsigned long long *cdf, int64_t M, const unsigned long long *deficit, int64_t global_offset,
                  const unsigned long long *K, double *c_all, const double *uniforms, uint64_t seed, uint32_t step,
                  int32_t *idx)
{
    const int f = blockIdx.x;
    const int64_t base = (int64_t)f * M;
    double *c = c_all + base;
    __shared__ double s_sum;
    for (int64_t i = threadIdx.x; i < M; i += kThreads) {
        const unsigned long long q = q_from_cdf(cdf, base, i, deficit[f], global_offset);
        const double w = cdf_value(q);
        c[i] = add_rn(w, -(double)residual_copies(q, M));                 // :70
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int64_t i = 0; i < M; ++i) s = add_rn(s, c[i]);              // builtin sum(), :71
        s_sum = s;
    }
    __syncthreads();
    const double s = s_sum;
    for (int64_t i = threadIdx.x; i < M; i += kThreads) c[i] = div_rn(c[i], s);   // :71
    __syncthreads();
    if (threadIdx.x == 0) {
        double acc = c[0];
        for (int64_t i = 1; i < M; ++i) {                                 // np.cumsum, :72
            acc = add_rn(acc, c[i]);
            c[i] = acc;
        }
        c[M - 1] = 1.0;                                                   // :73
        const int64_t k = (int64_t)K[f];
        SearchHint h;
        h.lo = 0;
        h.hi = M;
        h.last = 0.0;
        for (int64_t t = 0; t < M - k; ++t) {                             // :74
            const double key = uniforms ? uniforms[base + t] : uniform53(philox_for(seed, base + t, step, kStreamResample, 1));
            if (t == 0) h.last = key;
            idx[base + k + t] = (int32_t)searchsorted_left_hinted(c, M, key, h);
        }
    }
}

// new[o] = old[idx[o]] (+ fused add_noise), for index arrays that are not sorted
__global__ void __launch_bounds__(kThreads)
k_gather_indices(const int32_t *__restrict__ idx, int64_t M, StateView src, StateView dst, NoiseParams rough,
                 const float *rough_noise)
{
    const int f = blockIdx.y;
    const int64_t o = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (o >= M) return;
    const int64_t base = (int64_t)f * M, orow = base + o, srow = base + idx[orow];
    double x = src.x[srow], y = src.y[srow];
    float z = src.z[srow], ro = src.roll[srow], pi = src.pitch[srow], ya = src.yaw[srow];
    if (rough.mode == 1) {
        float nn[6];
        six_normals(rough, (int64_t)f * rough.global_stride + rough.global_offset + o, rough_noise, orow, nn);
        x += (double)(rough.sigma[0] * nn[0]);
        y += (double)(rough.sigma[1] * nn[1]);
        z += rough.sigma[2] * nn[2];
        ro += rough.sigma[3] * nn[3];
        pi += rough.sigma[4] * nn[4];
        ya += rough.sigma[5] * nn[5];
    }
    dst.x[orow] = x;
    dst.y[orow] = y;
    dst.z[orow] = z;
    dst.roll[orow] = ro;
    dst.pitch[orow] = pi;
    dst.yaw[orow] = ya;
}

__global__ void __launch_bounds__(kThreads)
k_load_records(const Record *__restrict__ rec, int64_t n, StateView dst, NoiseParams rough, const float *rough_noise)
{
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    Record r = rec[i];
    if (rough.mode == 1) {       // fused add_noise (auv_pf.py:191-192); single filter: global id = offset + i
        float nn[6];
        six_normals(rough, rough.global_offset + i, rough_noise, i, nn);
        r.x += (double)(rough.sigma[0] * nn[0]);
        r.y += (double)(rough.sigma[1] * nn[1]);
        r.z += rough.sigma[2] * nn[2];
        r.roll += rough.sigma[3] * nn[3];
        r.pitch += rough.sigma[4] * nn[4];
        r.yaw += rough.sigma[5] * nn[5];
    }
    dst.x[i] = r.x;
    dst.y[i] = r.y;
    dst.z[i] = r.z;
    dst.roll[i] = r.roll;
    dst.pitch[i] = r.pitch;
    dst.yaw[i] = r.yaw;
}

// ---------------------------------------------------------------------------------------------
// Reference in-place slot map (a8): auv_pf.py:183-189 keep/lost/dupes + reassign_poses :195-198.
constexpr int kNoFirst = 0x7f7f7f7f;

// ancestors of a resample the N_eff gate held back: every particle is its own
__global__ void __launch_bounds__(kThreads) k_identity_indices(int32_t *idx, int64_t M)
{
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i < M) idx[(int64_t)blockIdx.y * M + i] = (int32_t)i;
}

__global__ void __launch_bounds__(kThreads) k_first_occurrence(const int32_t *idx, int64_t M, int *first)
{
    int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (k >= M) return;
    int64_t base = (int64_t)blockIdx.y * M;
    int j = idx[base + k];
    if (j >= 0 && j < M) atomicMin(first + base + j, (int)k);
}

// scans of the "lost" flag over slots j and of the "dupe" flag over draws k
__global__ void __launch_bounds__(kThreads)
k_scan_lost(const int *first, int64_t M, unsigned long long *out, unsigned long long *status,
            unsigned int *tile_counter, int tiles_per_filter, unsigned long long *totals)
{
    auto value = [&](int64_t row, int, int) -> unsigned long long { return first[row] == kNoFirst ? 1ull : 0ull; };
    scan_tile_body(value, M, out, status, tile_counter, tiles_per_filter, totals);
}

__global__ void __launch_bounds__(kThreads)
k_scan_dupes(const int *first, const int32_t *idx, int64_t M, unsigned long long *out, unsigned long long *status,
             unsigned int *tile_counter, int tiles_per_filter, unsigned long long *totals)
{
    auto value = [&](int64_t row, int f, int) -> unsigned long long {
        int64_t base = (int64_t)f * M;
        int j = idx[row];
        return first[base + j] != (int)(row - base) ? 1ull : 0ull;
    };
    scan_tile_body(value, M, out, status, tile_counter, tiles_per_filter, totals);
}

__global__ void __launch_bounds__(kThreads)
k_compact_lost_dupes(const int *first, const int32_t *idx, int64_t M, const unsigned long long *scan_lost,
                     const unsigned long long *scan_dupes, int32_t *lost, int32_t *dupes)
{
    int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (k >= M) return;
    int64_t base = (int64_t)blockIdx.y * M, row = base + k;
    if (first[row] == kNoFirst) lost[base + scan_lost[row] - 1] = (int32_t)k;
    int j = idx[row];
    if (first[base + j] != (int)k) dupes[base + scan_dupes[row] - 1] = j;
}

__global__ void __launch_bounds__(kThreads)
k_reassign(StateView s, int64_t M, const int32_t *lost, const int32_t *dupes, const unsigned long long *n_lost,
           const unsigned long long *n_dupes)
{
    int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    int f = blockIdx.y;
    unsigned long long n = n_lost[f] < n_dupes[f] ? n_lost[f] : n_dupes[f];
    if (r >= (int64_t)n) return;
    int64_t base = (int64_t)f * M;
    int64_t d = base + lost[base + r], sidx = base + dupes[base + r];
    s.x[d] = s.x[sidx];
    s.y[d] = s.y[sidx];
    s.z[d] = s.z[sidx];
    s.roll[d] = s.roll[sidx];
    s.pitch[d] = s.pitch[sidx];
    s.yaw[d] = s.yaw[sidx];
}

// ---------------------------------------------------------------------------------------------
// Estimate (a10): auv_pf.update_loc_pose (auv_pf.py:218-252).  Pass 1: sums of the six pose
// components with yaw wrapped (:229-231); pass 2: centred second moments of x,y,z.
__device__ __forceinline__ double wrap_pi_d(double a)
{
    double t = fmod(a + kPi, 2.0 * kPi);
    if (t < 0.0) t += 2.0 * kPi;
    return t - kPi;
}

// ONE pass: positions are accumulated relative to the filter's first particle c (first and second
// moments of d = pos - c), so  mean = c + sum d / N  and  cov = sum d d^T / N - mean_d mean_d^T  keep
// full precision on a 2 km map.  stats: F x 16 = [sum d_xyz, sum roll, pitch, yaw, wrapped yaw,
// sum d d^T (xx yy zz xy xz yz), c_xyz].
__global__ void __launch_bounds__(kThreads) k_pose_stats(StateView s, int64_t M, double *stats)
{
    const int f = blockIdx.y;
    const int64_t base = (int64_t)f * M;
    const double cx = s.x[base], cy = s.y[base], cz = (double)s.z[base];
    double a[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < M; i += (int64_t)gridDim.x * kThreads) {
        const int64_t row = base + i;
        const double dx = s.x[row] - cx, dy = s.y[row] - cy, dz = (double)s.z[row] - cz;
        const double yw = (double)s.yaw[row];
        a[0] += dx;
        a[1] += dy;
        a[2] += dz;
        a[3] += (double)s.roll[row];
        a[4] += (double)s.pitch[row];
        a[5] += yw;
        a[6] += wrap_pi_d(yw);
        a[7] += dx * dx;
        a[8] += dy * dy;
        a[9] += dz * dz;
        a[10] += dx * dy;
        a[11] += dx * dz;
        a[12] += dy * dz;
    }
    // warp -> block -> 13 atomics per block (one per warp on the same 13 addresses cost 130 us at 1 M particles)
    __shared__ double s_part[kThreads / 32][13];
#pragma unroll
    for (int c = 0; c < 13; ++c) {
        const double v = warp_sum_f64(a[c]);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < 13) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) v += s_part[w][threadIdx.x];
        atomicAdd(stats + 16 * f + threadIdx.x, v);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        stats[16 * f + 13] = cx;
        stats[16 * f + 14] = cy;
        stats[16 * f + 15] = cz;
    }
}

// Estimate options: WEIGHTED by the current log-weights (w_i = exp(logw_i - max), float64) and / or
// with the CIRCULAR yaw mean atan2(sum w sin yaw, sum w cos yaw) instead of the reference's arithmetic mean of
// wrapped yaws (which is wrong for a cloud that straddles +-pi).  stats: F x 20 = [sum w, sum w d_xyz, sum w roll,
// sum w pitch, sum w wrapped yaw, sum w sin yaw, sum w cos yaw, sum w d d^T (xx yy zz xy xz yz), c_xyz, unused].
__global__ void __launch_bounds__(kThreads)
k_pose_stats_w(StateView s, int64_t M, const float *__restrict__ logw, const int *__restrict__ wmax_f32, int weighted,
               double *stats)
{
    const int f = blockIdx.y;
    const int64_t base = (int64_t)f * M;
    const double cx = s.x[base], cy = s.y[base], cz = (double)s.z[base];
    double m = 0.0;
    if (weighted) {
        const int k = wmax_f32[f];
        m = k > kWmaxReset ? (double)ordered_to_float(k) : 0.0;
    }
    double a[15] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < M; i += (int64_t)gridDim.x * kThreads) {
        const int64_t row = base + i;
        double w = 1.0;
        if (weighted) {
            const double x = (double)logw[row] - m;
            w = x == x ? exp(x) : 0.0;
        }
        const double dx = s.x[row] - cx, dy = s.y[row] - cy, dz = (double)s.z[row] - cz;
        const double yw = wrap_pi_d((double)s.yaw[row]);
        double sy, cyw;
        sincos(yw, &sy, &cyw);
        a[0] += w;
        a[1] += w * dx;
        a[2] += w * dy;
        a[3] += w * dz;
        a[4] += w * (double)s.roll[row];
        a[5] += w * (double)s.pitch[row];
        a[6] += w * yw;
        a[7] += w * sy;
        a[8] += w * cyw;
        a[9] += w * dx * dx;
        a[10] += w * dy * dy;
        a[11] += w * dz * dz;
        a[12] += w * dx * dy;
        a[13] += w * dx * dz;
        a[14] += w * dy * dz;
    }
    __shared__ double s_part[kThreads / 32][15];
#pragma unroll
    for (int c = 0; c < 15; ++c) {
        const double v = warp_sum_f64(a[c]);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < 15) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) v += s_part[w][threadIdx.x];
        atomicAdd(stats + 20 * f + threadIdx.x, v);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        stats[20 * f + 15] = cx;
        stats[20 * f + 16] = cy;
        stats[20 * f + 17] = cz;
    }
}

// The PoseArray of auv_pf.loc_loop (auv_pf.py:264-285) for a STRIDED subsample, built on the device: per
// published particle [x, y, z, qx, qy, qz, qw] with quaternion_from_euler(roll, pitch, yaw) ('sxyz', :272-275).
// Publishing every particle costs 48 MB per tick at 1 M particles; RViz shows a few thousand arrows.
__global__ void __launch_bounds__(kThreads)
k_pose_array(StateView s, int64_t M, int64_t stride, int64_t n, double *__restrict__ out)
{
    const int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (k >= n) return;
    const int f = blockIdx.y;
    const int64_t row = (int64_t)f * M + min(k * stride, M - 1);
    double sr, cr, sp, cp, sy, cy;
    sincos(0.5 * (double)s.roll[row], &sr, &cr);
    sincos(0.5 * (double)s.pitch[row], &sp, &cp);
    sincos(0.5 * (double)s.yaw[row], &sy, &cy);
    double *o = out + ((int64_t)f * n + k) * 7;
    o[0] = s.x[row];
    o[1] = s.y[row];
    o[2] = (double)s.z[row];
    o[3] = sr * cp * cy - cr * sp * sy;
    o[4] = cr * sp * cy + sr * cp * sy;
    o[5] = cr * cp * sy - sr * sp * cy;
    o[6] = cr * cp * cy + sr * sp * sy;
}

// ---- the k particles of largest weight ("strided / top-k particle subsample") ----------------------
// Radix select on a 64-bit key that is unique per particle: (order-preserving code of the log-weight) << 32 |
// (0xFFFFFFFF - index), so "larger key" = larger weight, lower index among equals; a NaN log-weight counts as
// -inf.  Four rounds of one 16-bit digit each: histogram of the digit among the keys that match the digits
// fixed so far (k_topk_hist), then one block picks the digit that holds the k-th largest (k_topk_pick).  What
// is left in `prefix` after the fourth round IS the k-th largest key; k_topk_emit collects the k keys >= it.
struct TopkState {
    unsigned long long prefix;       // the digits fixed so far (right-aligned)
    long long k_rem;                 // how many of the keys that match the prefix are still wanted
    unsigned int count;              // k_topk_emit: slots handed out
    unsigned int pad;
};

__device__ __forceinline__ unsigned long long topk_key(float lw, int64_t i)
{
    const float v = lw == lw ? lw : -INFINITY;
    const unsigned int code = (unsigned int)float_to_ordered(v) ^ 0x80000000u;      // signed order -> unsigned order
    return ((unsigned long long)code << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)i);
}

__global__ void __launch_bounds__(kThreads)
k_topk_hist(const float *__restrict__ logw, int64_t M, const TopkState *__restrict__ st, int round, unsigned int *__restrict__ hist)
{
    const int f = blockIdx.y;
    const unsigned long long prefix = st[f].prefix;
    const int shift = 48 - 16 * round;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < M; i += (int64_t)gridDim.x * kThreads) {
        const unsigned long long key = topk_key(logw[(int64_t)f * M + i], i);
        if (round == 0 || (key >> (shift + 16)) == prefix) atomicAdd(hist + (size_t)f * 65536 + ((key >> shift) & 0xFFFFull), 1u);
    }
}

// one block of 1024 threads per filter: the digit d with  #(digits > d) < k_rem <= #(digits >= d)
__global__ void __launch_bounds__(1024) k_topk_pick(TopkState *st, unsigned int *hist)
{
    __shared__ unsigned long long s_sum[1024];
    __shared__ int s_digit;
    __shared__ unsigned long long s_above;
    const int f = blockIdx.x, t = threadIdx.x;
    unsigned int *h = hist + (size_t)f * 65536;
    const long long k_rem = st[f].k_rem;
    // thread t owns the 64 digits [65536 - 64 (t + 1), 65536 - 64 t): thread 0 the largest
    unsigned long long mine = 0;
    const int base = 65536 - 64 * (t + 1);
    for (int j = 0; j < 64; ++j) mine += h[base + j];
    s_sum[t] = mine;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                  // inclusive scan over the threads (largest digits first)
        const unsigned long long v = t >= o ? s_sum[t - o] : 0;
        __syncthreads();
        s_sum[t] += v;
        __syncthreads();
    }
    const unsigned long long incl = s_sum[t], excl = incl - mine;
    if ((long long)excl < k_rem && k_rem <= (long long)incl) {   // exactly one thread
        unsigned long long above = excl;
        int d = base + 63;
        for (; d >= base; --d) {
            if ((long long)(above + h[d]) >= k_rem) break;
            above += h[d];
        }
        s_digit = d;
        s_above = above;
    }
    __syncthreads();
    for (int j = t; j < 65536; j += 1024) h[j] = 0;       // ready for the next round
    if (t == 0) {
        st[f].prefix = (st[f].prefix << 16) | (unsigned long long)s_digit;
        st[f].k_rem = k_rem - (long long)s_above;
    }
}

__global__ void __launch_bounds__(kThreads)
k_topk_emit(const float *__restrict__ logw, int64_t M, TopkState *st, int64_t k, int *__restrict__ idx_out)
{
    const int f = blockIdx.y;
    const unsigned long long kth = st[f].prefix;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < M; i += (int64_t)gridDim.x * kThreads)
        if (topk_key(logw[(int64_t)f * M + i], i) >= kth) {
            const unsigned int slot = atomicAdd(&st[f].count, 1u);
            if ((int64_t)slot < k) idx_out[(int64_t)f * k + slot] = (int)i;
        }
}

// Pose entries [x y z qx qy qz qw] and log-weights of listed particles
__global__ void __launch_bounds__(kThreads)
k_pose_array_idx(StateView s, const float *__restrict__ logw, int64_t M, const int *__restrict__ idx, int64_t n,
                 double *__restrict__ out, float *__restrict__ lw_out)
{
    const int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (k >= n) return;
    const int f = blockIdx.y;
    const int64_t row = (int64_t)f * M + idx[(int64_t)f * n + k];
    double sr, cr, sp, cp, sy, cy;
    sincos(0.5 * (double)s.roll[row], &sr, &cr);
    sincos(0.5 * (double)s.pitch[row], &sp, &cp);
    sincos(0.5 * (double)s.yaw[row], &sy, &cy);
    double *o = out + ((int64_t)f * n + k) * 7;
    o[0] = s.x[row];
    o[1] = s.y[row];
    o[2] = (double)s.z[row];
    o[3] = sr * cp * cy - cr * sp * sy;
    o[4] = cr * sp * cy + sr * cp * sy;
    o[5] = cr * cp * sy - sr * sp * cy;
    o[6] = cr * cp * cy + sr * sp * sy;
    lw_out[(int64_t)f * n + k] = logw[row];
}

// ---- KLD-sampling (Fox 2003): how many bins of size (bin_xy, bin_xy, bin_yaw) the cloud occupies ------
// One open-addressing hash set per call; a key is the bin triple and the filter packed into 64 bits (22 + 22 + 12 + 8).
// The COUNT of distinct keys does not depend on the insertion order.  out[0..F) += newly occupied bins; out[F] != 0
// when a bin index did not fit its field.
__global__ void __launch_bounds__(kThreads)
k_kld_bins(StateView s, int64_t M, double inv_bin_xy, double inv_bin_yaw, unsigned long long *table, unsigned long long mask,
           unsigned long long *out, int F)
{
    const int f = blockIdx.y;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < M; i += (int64_t)gridDim.x * kThreads) {
        const int64_t row = (int64_t)f * M + i;
        const double bx = floor(s.x[row] * inv_bin_xy), by = floor(s.y[row] * inv_bin_xy);
        const double bw = floor(((double)s.yaw[row] + kPi) * inv_bin_yaw);
        if (!(fabs(bx) < 2097152.0 && fabs(by) < 2097152.0 && fabs(bw) < 2048.0)) {      // (NaN poses land here too)
            out[F] = 1ull;
            continue;
        }
        const unsigned long long key = ((unsigned long long)((long long)bx + 2097152ll) << 42) |
                                       ((unsigned long long)((long long)by + 2097152ll) << 20) |
                                       ((unsigned long long)((long long)bw + 2048ll) << 8) | (unsigned long long)f;
        unsigned long long hsh = key * 0x9E3779B97F4A7C15ull;
        hsh ^= hsh >> 29;
        for (unsigned long long slot = hsh & mask;; slot = (slot + 1) & mask) {
            const unsigned long long prev = atomicCAS(table + slot, ~0ull, key);
            if (prev == ~0ull) {
                atomicAdd(out + f, 1ull);
                break;
            }
            if (prev == key) break;
        }
    }
}

// pack / unpack between the ABI's (n,6) float64 rows and the SoA state
__global__ void __launch_bounds__(kThreads) k_set_poses(const double *poses, int64_t n, StateView s)
{
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const double *p = poses + 6 * i;
    s.x[i] = p[0];
    s.y[i] = p[1];
    s.z[i] = (float)p[2];
    s.roll[i] = (float)p[3];
    s.pitch[i] = (float)p[4];
    s.yaw[i] = (float)p[5];
}

__global__ void __launch_bounds__(kThreads) k_get_poses(StateView s, int64_t n, double *poses)
{
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    double *p = poses + 6 * i;
    p[0] = s.x[i];
    p[1] = s.y[i];
    p[2] = (double)s.z[i];
    p[3] = (double)s.roll[i];
    p[4] = (double)s.pitch[i];
    p[5] = (double)s.yaw[i];
}

}   // namespace auvpf
