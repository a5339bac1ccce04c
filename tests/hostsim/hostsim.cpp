// hostsim.cpp -- TEST-ONLY host compile of the device math (auvpf_math.cuh).
//
// The build container has no GPU.  This file compiles the very same
// __host__ __device__ functions the CUDA kernels call and drives them with
// plain loops, so the CPU test-suite can check the kernel ARITHMETIC (raycast
// traversal, propagate, fixed-point normalisation, ancestor search, Philox)
// against the oracle before any GPU time is spent.  It is NOT part of the
// product: nothing under smarc_navigation_b200/ loads it, and it is never used
// as a fallback.  The GPU parity tests (-m gpu) go through libauvpf.so.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#define AUVPF_DEBUG_REASONS 1
#include "../../smarc_navigation_b200/csrc/auvpf_fan.h"
#include "../../smarc_navigation_b200/csrc/auvpf_math.cuh"

using namespace auvpf;

namespace auvpf {
long g_prof_reasons[8] = {0};
}

namespace {
struct HostDem {
    std::vector<float> mip;
    DemView v;
};

void build_mip(const float *z, int H, int W, int shift, std::vector<float> &out, int &oH, int &oW)
{
    std::vector<float> prev;
    int pH = 0, pW = 0;
    for (int L = 1; L <= shift; ++L) {
        int S = 1 << L;
        int h = std::max(1, (H - 1 + S - 1) / S), w = std::max(1, (W - 1 + S - 1) / S);
        std::vector<float> lvl((size_t)h * w);
        for (int bi = 0; bi < h; ++bi)
            for (int bj = 0; bj < w; ++bj) {
                float m = -INFINITY;
                if (L == 1) {
                    for (int di = 0; di <= 2; ++di)
                        for (int dj = 0; dj <= 2; ++dj)
                            m = std::max(m, z[(size_t)std::min(2 * bi + di, H - 1) * W + std::min(2 * bj + dj, W - 1)]);
                } else {
                    for (int di = 0; di < 2; ++di)
                        for (int dj = 0; dj < 2; ++dj) {
                            int i = 2 * bi + di, j = 2 * bj + dj;
                            if (i < pH && j < pW) m = std::max(m, prev[(size_t)i * pW + j]);
                        }
                }
                lvl[(size_t)bi * w + bj] = m;
            }
        prev.swap(lvl);
        pH = h;
        pW = w;
    }
    out.swap(prev);
    oH = pH;
    oW = pW;
}
}   // namespace

extern "C" {

// expected ranges (n, B) for explicit poses: mirrors k_raycast_poses
void hs_raycast_poses(const float *z, int H, int W, double x0, double y0, double res, int shift, int shift_hi,
                      double range_max,
                      const double *m2o, const double *poses, long n, const float *beams, int B, float *out,
                      float *mip_out /* optional, for checking the pyramid */, int use_fan, long *n_swept)
{
    FanPlan plan = plan_fan(beams, B);
    *n_swept = 0;
    HostDem d;
    int mH, mW, mH2 = 0, mW2 = 0;
    std::vector<float> mip_hi;
    build_mip(z, H, W, shift, d.mip, mH, mW);
    if (shift_hi > shift) build_mip(z, H, W, shift_hi, mip_hi, mH2, mW2);
    // padded copy: replicated one-vertex border, pitch W + 2 (as auvpf_set_dem stores it)
    const int pitch = W + 2;
    std::vector<float> zp((size_t)(H + 2) * pitch);
    for (int i = -1; i <= H; ++i)
        for (int j = -1; j <= W; ++j)
            zp[(size_t)(i + 1) * pitch + (j + 1)] = z[(size_t)std::min(std::max(i, 0), H - 1) * W + std::min(std::max(j, 0), W - 1)];
    if (mip_out) memcpy(mip_out, d.mip.data(), sizeof(float) * d.mip.size());
    DemView &v = d.v;
    v.z = zp.data() + pitch + 1;
    v.W = W;
    v.H = H;
    v.pitch = pitch;
    v.mip[0] = d.mip.data();
    v.mipW[0] = mW;
    v.mipH[0] = mH;
    v.shift[0] = shift;
    v.mip[1] = mip_hi.empty() ? d.mip.data() : mip_hi.data();
    v.mipW[1] = mip_hi.empty() ? mW : mW2;
    v.mipH[1] = mip_hi.empty() ? mH : mH2;
    v.shift[1] = mip_hi.empty() ? shift : shift_hi;
    v.n_levels = mip_hi.empty() ? 1 : 2;
    v.range_max = (float)range_max;
    v.zmax = *std::max_element(z, z + (size_t)H * W);
    v.res = (float)res;
    {
        float ms = 0.f;
        for (int i = 0; i + 1 < H; ++i)
            for (int j = 0; j + 1 < W; ++j) {
                float z00 = z[(size_t)i * W + j], z01 = z[(size_t)i * W + j + 1], z10 = z[(size_t)(i + 1) * W + j],
                      z11 = z[(size_t)(i + 1) * W + j + 1];
                ms = std::max(ms, std::max(std::hypot(z01 - z00, z11 - z01), std::hypot(z11 - z10, z10 - z00)) * (float)(1.0 / res));
            }
        v.max_slope = ms * 1.0001f + 1e-6f;
    }
    v.x0 = x0;
    v.y0 = y0;
    v.inv_res_d = 1.0 / res;
    FrameParams fp;
    bool ident = true;
    for (int a = 0; a < 3; ++a) {
        fp.m2o_t[a] = m2o[4 * a + 3];
        for (int b = 0; b < 3; ++b) {
            fp.m2o_r[3 * a + b] = m2o[4 * a + b];
            if (m2o[4 * a + b] != (a == b ? 1.0 : 0.0)) ident = false;
        }
    }
    fp.m2o_identity_r = ident;
    for (long i = 0; i < n; ++i) {
        const double *ps = poses + 6 * i;
        double p[3];
        double R[9];
        particle_frame(fp, ps[0], ps[1], (float)ps[2], (float)ps[3], (float)ps[4], (float)ps[5], p, R);
        int ix0, iy0;
        float lx0, ly0;
        origin_cell(v, p, ix0, iy0, lx0, ly0);
        if (use_fan == 2) {
            // profile merge (k_update_mbes_profile): one march per side, generic fallback
            FanTables T;
            std::vector<BeamAux> aux(B);
            for (int b = 0; b < B; ++b) {
                aux[b].sec = (float)plan.sec_t[b];
                aux[b].reach = plan.s[b] * v.range_max;
                aux[b].meas = 0.f;
                aux[b].rmax = v.range_max;
            }
            T.tan_t = plan.tan_t.data();
            T.sec_t = plan.sec_t.data();
            T.aux = aux.data();
            T.B = B;
            T.n1 = plan.n1;
            T.sgn1 = plan.sgn1;
            for (int a = 0; a < 3; ++a) {
                T.c0[a] = plan.c0[a];
                T.a0[a] = plan.a0[a];
            }
            float *row = out + (size_t)i * B;
            bool ok = plan.planar && profile_particle(v, T, p, R, true, [&](int b, float range, float) { row[b] = range; });
            if (ok) *n_swept += B;
            else
                for (int b = 0; b < B; ++b) {
                    Ray ray = make_ray(v, p, ix0, iy0, lx0, ly0, R, beams[3 * b], beams[3 * b + 1], beams[3 * b + 2]);
                    row[b] = raycast_one(v, ray);
                }
            continue;
        }
        if (!use_fan) {
            for (int b = 0; b < B; ++b) {
                Ray ray = make_ray(v, p, ix0, iy0, lx0, ly0, R, beams[3 * b], beams[3 * b + 1], beams[3 * b + 2]);
                out[(size_t)i * B + b] = raycast_one(v, ray);
            }
            continue;
        }
        // the kernels' fan_cast, lane by lane (auvpf_kernels.cuh)
        const bool sweep_ok = plan.planar && fan_sweep_valid(v, p, R, plan.c0);
        const int C = plan.chunk;
        if (!sweep_ok) {
            for (int b = 0; b < B; ++b) {
                Ray ray = make_ray(v, p, ix0, iy0, lx0, ly0, R, beams[3 * b], beams[3 * b + 1], beams[3 * b + 2]);
                out[(size_t)i * B + b] = raycast_one(v, ray);
            }
            continue;
        }
        float p_clear[kFanLanes];
        const int L1 = plan.lanes_first;
        const int max_chain = std::max(L1, kFanLanes - L1) - 1;
        for (int c = 0; c <= max_chain; ++c)
            for (int lane = 0; lane < kFanLanes; ++lane) {
                const bool first_side = lane < L1;
                const int chain = first_side ? (L1 - 1 - lane) : (lane - L1);
                if (chain != c) continue;
                int o = plan.order[(size_t)lane * C];
                p_clear[lane] = 0.f;
                if (o < 0) continue;
                int b = o & (kOrdFull - 1);
                Ray ray = make_ray(v, p, ix0, iy0, lx0, ly0, R, beams[3 * b], beams[3 * b + 1], beams[3 * b + 2]);
                float t_lo = 0.f;
                (void)first_side;   // chunk starts are traversed from the sensor (see fan_cast)
                float range = raycast_from(v, ray, t_lo, true);
                if (t_lo > 0.f) ++*n_swept;
                p_clear[lane] = range * plan.s[b];
                out[(size_t)i * B + b] = range;
            }
        for (int lane = 0; lane < kFanLanes; ++lane)
            for (int k = 1; k < C; ++k) {
                int o = plan.order[(size_t)lane * C + k];
                if (o < 0) continue;
                int b = o & (kOrdFull - 1);
                Ray ray = make_ray(v, p, ix0, iy0, lx0, ly0, R, beams[3 * b], beams[3 * b + 1], beams[3 * b + 2]);
                float t_lo = sweep_bound(p_clear[lane], plan.inv_s[b]);
                float range = raycast_from(v, ray, t_lo, true);
                if (t_lo > 0.f) ++*n_swept;
                p_clear[lane] = range * plan.s[b];
                out[(size_t)i * B + b] = range;
            }
    }
}

int hs_mip_dims(int H, int W, int shift, int *oH, int *oW)
{
    int S = 1 << shift;
    *oH = std::max(1, (H - 1 + S - 1) / S);
    *oW = std::max(1, (W - 1 + S - 1) / S);
    return 0;
}

// propagate: mirrors k_propagate with injected noise (n,6); params computed by the caller
void hs_propagate(double *x, double *y, float *z, float *roll, float *pitch, float *yaw, long n, float a, float b,
                  float wzdt, float proll, float ppitch, float pz, float sx, float sy, float syaw, const float *noise)
{
    PredictParams p{a, b, wzdt, proll, ppitch, pz, sx, sy, syaw, 0.f};
    for (long i = 0; i < n; ++i)
        propagate_one(p, x[i], y[i], z[i], roll[i], pitch[i], yaw[i], noise[6 * i], noise[6 * i + 1], noise[6 * i + 5]);
}

void hs_philox(const uint32_t *ctr, const uint32_t *key, long n, uint32_t *out)
{
    for (long i = 0; i < n; ++i) {
        U4 c{ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3]};
        U4 r = philox4x32_10(c, key[2 * i], key[2 * i + 1]);
        out[4 * i] = r.x;
        out[4 * i + 1] = r.y;
        out[4 * i + 2] = r.z;
        out[4 * i + 3] = r.w;
    }
}

// the 6 normals of six_normals() for particle ids, pose order
void hs_normals6(uint64_t seed, const int64_t *ids, long n, uint32_t step, uint32_t stream, float *out)
{
    for (long i = 0; i < n; ++i) pose_normals(seed, ids[i], step, stream, 0x3fu, out + 6 * i);
}

// fixed-point normalisation + systematic select, sequentially, with the device functions
int hs_resample(const float *logw, long M, double u, uint64_t *qq_out, int32_t *anc_out)
{
    float m = -INFINITY;
    for (long i = 0; i < M; ++i)
        if (logw[i] == logw[i] && logw[i] > m) m = logw[i];
    if (!std::isfinite(m)) return -1;
    uint64_t lo = 0, hi = 0;
    for (long i = 0; i < M; ++i) {
        uint64_t a = a_from_logw(logw[i], m);
        lo += a & 0xFFFFFFFFull;
        hi += a >> 32;
    }
    unsigned __int128 A = (unsigned __int128)lo + ((unsigned __int128)hi << 32);
    uint64_t A_lo = (uint64_t)A, A_hi = (uint64_t)(A >> 64);
    std::vector<uint64_t> cdf(M);
    uint64_t acc = 0;
    for (long i = 0; i < M; ++i) {
        acc += q_from_a(a_from_logw(logw[i], m), A_lo, A_hi);
        cdf[i] = acc;
    }
    uint64_t D = (1ull << kQBits) - acc;
    for (long i = 0; i < M; ++i) {
        uint64_t q = q_from_a(a_from_logw(logw[i], m), A_lo, A_hi);
        if (qq_out) qq_out[i] = q + ((uint64_t)i < D ? 1 : 0);
    }
    double n = (double)M;
    for (long i = 0; i < M; ++i) {
        double pos = sys_position(u, i, n);
        long l = 0, h = M;
        while (l < h) {
            long mid = (l + h) >> 1;
            uint64_t jg1 = (uint64_t)(mid + 1);
            double c = cdf_value(cdf[mid] + (jg1 < D ? jg1 : D));
            if (pos < c) h = mid;
            else l = mid + 1;
        }
        anc_out[i] = (int32_t)std::min(l, M - 1);
    }
    return 0;
}

// residual_resample with the device functions, sequentially (mirrors k_scan_copies / k_scan_fracs /
// k_residual_fill and, with strict != 0, k_residual_strict)
int hs_residual(const uint64_t *qq, long M, const double *U, int strict, int32_t *idx)
{
    std::vector<uint64_t> copies(M), fracs(M);
    uint64_t k = 0, R = 0;
    for (long i = 0; i < M; ++i) {
        const uint64_t c = residual_copies(qq[i], M);
        k += c;
        copies[i] = k;
        R += residual_frac(qq[i], c, M);
        fracs[i] = R;
    }
    auto first_above = [&](const std::vector<uint64_t> &c, uint64_t key) {
        long lo = 0, hi = M;
        while (lo < hi) {
            long mid = (lo + hi) >> 1;
            if (key < c[mid]) hi = mid;
            else lo = mid + 1;
        }
        return lo;
    };
    for (long o = 0; o < (long)k && o < M; ++o) idx[o] = (int32_t)std::min<long>(first_above(copies, (uint64_t)o), M - 1);
    if (!strict) {
        for (long o = (long)k; o < M; ++o)
            idx[o] = (int32_t)std::min<long>(first_above(fracs, residual_target(U[o - k], R)), M - 1);
        return 0;
    }
    std::vector<double> c(M);
    double s = 0.0;
    for (long i = 0; i < M; ++i) {
        c[i] = add_rn(cdf_value(qq[i]), -(double)residual_copies(qq[i], M));
        s = add_rn(s, c[i]);
    }
    for (long i = 0; i < M; ++i) c[i] = div_rn(c[i], s);
    double acc = c[0];
    for (long i = 1; i < M; ++i) {
        acc = add_rn(acc, c[i]);
        c[i] = acc;
    }
    c[M - 1] = 1.0;
    SearchHint h;
    h.lo = 0;
    h.hi = M;
    h.last = 0.0;
    for (long t = 0; t < M - (long)k; ++t) {
        if (t == 0) h.last = U[t];
        idx[k + t] = (int32_t)searchsorted_left_hinted(c.data(), M, U[t], h);
    }
    return 0;
}

void hs_searchsorted(const double *arr, long n, const double *keys, long nk, long *out)
{
    SearchHint h;
    h.lo = 0;
    h.hi = n;
    h.last = nk ? keys[0] : 0.0;
    for (long t = 0; t < nk; ++t) out[t] = (long)searchsorted_left_hinted(arr, n, keys[t], h);
}

uint64_t hs_q_from_a(uint64_t a, uint64_t A_lo, uint64_t A_hi) { return q_from_a(a, A_lo, A_hi); }

void hs_prof_reasons(long *out, int reset)
{
    for (int k = 0; k < 8; ++k) {
        out[k] = g_prof_reasons[k];
        if (reset) g_prof_reasons[k] = 0;
    }
}

long hs_count_below(double u, long N, uint64_t c) { return (long)sys_count_below(u, N, c); }

float hs_wrap_pi(float a) { return wrap_pi(a); }

// Host twin of k_select_gather's ancestor selection (single-GPU form): tile table from the raw quotients,
// 32-ary tile location, source-side walk with rel_end_of (the SAME function the kernel calls), pruning
// by thread, run heads, inclusive max-scan -- block by block, "thread" by "thread", in the kernel's order.
void hs_select_source_side(const uint64_t *q_raw, long M, double u, uint64_t D, int32_t *anc)
{
    const int kTile = 2048, kItems = 8, kThreads = 256, kOut = 1024;
    const long tiles = (M + kTile - 1) / kTile;
    std::vector<uint64_t> tab(tiles);
    {
        uint64_t acc = 0;
        for (long t = 0; t < tiles; ++t) {
            for (long j = t * kTile; j < std::min<long>((t + 1) * kTile, M); ++j) acc += q_raw[j];
            tab[t] = acc;
        }
    }
    const double n = (double)M;
    const SysPos spos = make_sys_pos(u, n);
    auto full = [&](uint64_t raw_cdf, long j) { const uint64_t jg1 = (uint64_t)(j + 1); return raw_cdf + (jg1 < D ? jg1 : D); };
    for (long k0 = 0; k0 < M; k0 += kOut) {
        const long k1 = std::min<long>(k0 + kOut, M) - 1;
        const int nblk = (int)(k1 - k0 + 1);
        // level 1: 32-ary search, one "warp" per key
        long tl[2];
        for (int w = 0; w < 2; ++w) {
            const uint64_t key = (uint64_t)floor(sys_position(u, w == 0 ? k0 : k1, n) * 4503599627370496.0);
            long lo = 0, hi = tiles;
            while (lo < hi) {
                const long step = (hi - lo + 31) >> 5;
                int nb = 0;
                for (int lane = 0; lane < 32; ++lane) {
                    const long t = lo + (lane + 1) * step - 1;
                    bool below = false;
                    if (t < hi) {
                        const long je = std::min<long>((t + 1) * kTile, M) - 1;
                        below = !(key < full(tab[t], je));
                    }
                    nb += below ? 1 : 0;
                }
                lo += nb * step;
                hi = std::min(lo + step - 1, hi);
            }
            tl[w] = std::min(lo, tiles - 1);
        }
        const long j_lo = tl[0] * kTile, j_hi = std::min<long>((tl[1] + 1) * kTile, M) - 1;
        const RelEndCtx rctx = make_rel_end(u, n, k0, nblk);
        auto rel_end = [&](uint64_t raw_cdf, long j) { return rel_end_of(spos, M, rctx, full(raw_cdf, j)); };
        std::vector<int> head(kOut, 0);
        for (long jb = j_lo; jb <= j_hi; jb += (long)kThreads * kItems) {
            const long tile = jb / kTile;
            const uint64_t tile_excl = tile > 0 ? tab[tile - 1] : 0;
            if (jb > j_lo && rel_end(tile_excl, jb - 1) >= nblk) break;
            uint64_t run_before = tile_excl;
            int prev_thread_last = 0;
            for (int tid = 0; tid < kThreads; ++tid) {
                const long j0 = jb + (long)tid * kItems;
                uint64_t cv[kItems];
                uint64_t acc = run_before;
                for (int k = 0; k < kItems; ++k) {
                    acc += (j0 + k <= j_hi) ? q_raw[j0 + k] : 0;
                    cv[k] = acc;
                }
                const uint64_t before = run_before;
                run_before = acc;
                int e[kItems];
                const bool whole = j0 + kItems - 1 <= j_hi;
                const int e_last = whole ? rel_end(cv[kItems - 1], j0 + kItems - 1) : nblk;
                const int e_first = (e_last == 0) ? 0 : (j0 <= j_hi ? rel_end(cv[0], j0) : nblk);
                if (e_last == 0 || e_first == nblk) {
                    for (int k = 0; k < kItems; ++k) e[k] = e_first;
                } else {
                    e[0] = e_first;
                    for (int k = 1; k < kItems - 1; ++k) e[k] = j0 + k <= j_hi ? rel_end(cv[k], j0 + k) : nblk;
                    e[kItems - 1] = e_last;
                }
                for (int k = 0; k < kItems; ++k)
                    if (j0 + k == M - 1) e[k] = nblk;       // the last particle ends where the outputs end (kernel: j_last)
                int start;
                if ((tid & 31) == 0) {
                    const bool has_prev = j0 > j_lo && j0 - 1 <= j_hi;
                    start = has_prev ? rel_end(before, j0 - 1) : (j0 > j_lo ? nblk : 0);
                } else
                    start = prev_thread_last;
                for (int k = 0; k < kItems; ++k) {
                    if (e[k] > start) head[start] = (int)(j0 + k - j_lo) + 1;
                    start = e[k];
                }
                prev_thread_last = e[kItems - 1];
            }
        }
        int m = 0;
        for (int k = 0; k < nblk; ++k) {
            m = std::max(m, head[k]);
            anc[k0 + k] = (int32_t)std::min<long>(j_lo + std::max(m - 1, 0), M - 1);
        }
    }
}

// the N_eff gate's integer sum of squares, accumulated like the scan kernel does (two limbs)
double hs_neff_fixed(const uint64_t *q_raw, long n)
{
    uint64_t lo = 0, hi = 0;
    for (long i = 0; i < n; ++i) {
        const uint64_t s2 = neff_square(q_raw[i]);
        lo += s2 & 0xFFFFFFFFull;
        hi += s2 >> 32;
    }
    return neff_fixed(lo, hi);
}
}
