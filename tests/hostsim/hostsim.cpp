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
    std::vector<float> mip, slope_blk;
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
    v.last_cx = W - 2;
    v.last_cy = H - 2;
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
        // per block of kSlopeBlock x kSlopeBlock cells (k_slope_blocks)
        v.sbW = std::max(1, (W - 1 + kSlopeBlock - 1) / kSlopeBlock);
        v.sbH = std::max(1, (H - 1 + kSlopeBlock - 1) / kSlopeBlock);
        d.slope_blk.assign((size_t)v.sbW * v.sbH, 0.f);
        for (int i = 0; i + 1 < H; ++i)
            for (int j = 0; j + 1 < W; ++j) {
                float z00 = z[(size_t)i * W + j], z01 = z[(size_t)i * W + j + 1], z10 = z[(size_t)(i + 1) * W + j],
                      z11 = z[(size_t)(i + 1) * W + j + 1];
                float &b = d.slope_blk[(size_t)(i >> kSlopeShift) * v.sbW + (j >> kSlopeShift)];
                b = std::max(b, std::max(std::hypot(z01 - z00, z11 - z01), std::hypot(z11 - z10, z10 - z00)) * (float)(1.0 / res));
            }
        for (float &b : d.slope_blk) b = b * 1.0001f + 1e-6f;
        v.slope_blk = d.slope_blk.data();
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
            std::vector<BeamAux> aux_store(B + 2, beam_sentinel(v.range_max));      // entries -1 and B: the sentinels
            BeamAux *aux = aux_store.data() + 1;
            for (int b = 0; b < B; ++b) {
                aux[b].sec = (float)plan.sec_t[b];
                aux[b].reach = plan.s[b] * v.range_max;
                aux[b].meas = 0.f;
                aux[b].tn = (float)plan.tan_t[b];
            }
            T.tan_t = plan.tan_t.data();
            T.sec_t = plan.sec_t.data();
            T.aux = aux;
            T.B = B;
            T.n1 = plan.n1;
            T.sgn1 = plan.sgn1;
            T.s_lim[0] = plan.n1 > 0 ? (1.0f - kDenPrecise) / fmaxf(aux[0].tn, 1.0e-6f) : 0.f;
            T.s_lim[1] = B > plan.n1 ? (1.0f - kDenPrecise) / fmaxf(aux[B - 1].tn, 1.0e-6f) : 0.f;
            for (int a = 0; a < 3; ++a) {
                T.c0[a] = plan.c0[a];
                T.a0[a] = plan.a0[a];
            }
            float *row = out + (size_t)i * B;
            bool ok = plan.planar && profile_particle<true>(v, T, p, R, true, [&](int b, float range, float) { row[b] = range; },
                                                      [row](int b, float range, float) { row[b] = range; });
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
    unsigned __int128 A = 0;
    for (long i = 0; i < M; ++i) A += a_from_logw(logw[i], m);
    if (A == 0) return -1;
    const CdfDiv div = make_cdfdiv(u128_from(A));
    std::vector<uint64_t> cdf(M);
    unsigned __int128 C = 0;
    uint64_t prev = 0;
    for (long i = 0; i < M; ++i) {
        C += a_from_logw(logw[i], m);
        cdf[i] = cdf_exact(u128_from(C), div);
        if (qq_out) qq_out[i] = cdf[i] - prev;
        prev = cdf[i];
    }
    double n = (double)M;
    for (long i = 0; i < M; ++i) {
        double pos = sys_position(u, i, n);
        long l = 0, h = M;
        while (l < h) {
            long mid = (l + h) >> 1;
            if (pos < cdf_value(cdf[mid])) h = mid;
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

uint64_t hs_cdf_exact(uint64_t C_lo, uint64_t C_hi, uint64_t A_lo, uint64_t A_hi)
{
    return cdf_exact(u128(C_lo, C_hi), make_cdfdiv(u128(A_lo, A_hi)));
}
double hs_cdf_estimate(uint64_t C_lo, uint64_t C_hi, uint64_t A_lo, uint64_t A_hi)
{
    return cdf_estimate(u128(C_lo, C_hi), make_cdfdiv(u128(A_lo, A_hi)));
}

void hs_prof_reasons(long *out, int reset)
{
    for (int k = 0; k < 8; ++k) {
        out[k] = g_prof_reasons[k];
        if (reset) g_prof_reasons[k] = 0;
    }
}

long hs_count_below(double u, long N, uint64_t c) { return (long)sys_count_below(u, N, c); }

float hs_wrap_pi(float a) { return wrap_pi(a); }

// Host twin of k_select's ancestor selection for G shards of M sources each (G = 1: one GPU): chunk / tile
// tables as k_weights leaves them, the conservative 32-ary tile location on the cdf ESTIMATE, then block by
// block, "warp" by "warp", "lane" by "lane" in the kernel's order: chunk bases from the tables, float64 running
// sums and the warp scan's order of additions, rel_end_fast, the exact fix-up (cdf_exact + rel_end_of) only where
// the estimate cannot decide, run heads, inclusive max-scan.  a: the integer weights of all G*M sources.
// stats (optional, 2 longs): sources evaluated by the fast path / by the exact path.
// n_out: positions (u + i) / n_out, i < n_out, over the CDF of the G*M sources (the resample that changes the particle
// count; n_out = G*M: the ordinary one) -- SelArgs::n_global vs n_src_total.
void hs_select_source_side_n(const uint64_t *a, long M, int G, double u, long n_out, int32_t *anc, long *stats)
{
    const int kTile = 2048, kItems = 8, kOut = 1024, kCh = 256, kCpt = 8;
    const long tiles = (M + kTile - 1) / kTile, n_tiles = G * tiles, Ns = (long)G * M, N = n_out;
    std::vector<U128> chunk((size_t)n_tiles * kCpt), tprefix((size_t)n_tiles), coff(G + 1);
    {
        unsigned __int128 acc_all = 0;
        for (int g = 0; g < G; ++g) {
            coff[g] = u128_from(acc_all);
            unsigned __int128 acc = 0;
            for (long t = 0; t < tiles; ++t) {
                for (int c = 0; c < kCpt; ++c) {
                    unsigned __int128 cs = 0;
                    for (long j = t * kTile + c * kCh; j < std::min<long>(t * kTile + (c + 1) * kCh, M); ++j) cs += a[(long)g * M + j];
                    chunk[((size_t)g * tiles + t) * kCpt + c] = u128_from(cs);
                    acc += cs;
                }
                tprefix[(size_t)g * tiles + t] = u128_from(acc);
            }
            acc_all += acc;
        }
        coff[G] = u128_from(acc_all);
    }
    const CdfDiv div = make_cdfdiv(coff[G]);
    const double n = (double)N;
    const SysPos spos = make_sys_pos(u, n);
    long n_fast = 0, n_exact = 0;
    for (long k0 = 0; k0 < N; k0 += kOut) {
        const long k1 = std::min<long>(k0 + kOut, N) - 1;
        const int nblk = (int)(k1 - k0 + 1);
        const RelEndCtx rctx = make_rel_end(u, n, k0, nblk);
        auto rel_end = [&](U128 C) -> int {
            const int e = rel_end_fast(rctx, cdf_estimate(C, div));
            return e >= 0 ? e : rel_end_of(spos, N, rctx, cdf_exact(C, div));
        };
        long tl[2];
        for (int w = 0; w < 2; ++w) {
            long lo = 0, hi = n_tiles;
            while (lo < hi) {
                const long step = (hi - lo + 31) >> 5;
                int nb = 0;
                for (int lane = 0; lane < 32; ++lane) {
                    const long T = lo + (lane + 1) * step - 1;
                    bool before = false;
                    if (T < hi) {
                        const int g = (int)(T / tiles);
                        const U128 C = coff[g] + tprefix[T];
                        const double x = fma(cdf_estimate(C, div), rctx.n_scaled, -rctx.u) - rctx.g0d;
                        before = w == 0 ? (x <= -1.0e-5) : (x < rctx.nblk_d + 1.0e-5);
                    }
                    nb += before ? 1 : 0;
                }
                lo += nb * step;
                hi = std::min(lo + step - 1, hi);
            }
            tl[w] = std::min(lo, n_tiles - 1);
        }
        const long T_lo = tl[0], T_hi = tl[1];
        std::vector<int> head(kOut, 0);
        for (int warp = 0; warp < 8; ++warp) {
            for (long T = T_lo; T <= T_hi; ++T) {
                const int g = (int)(T / tiles);
                const long t = T - (long)g * tiles;
                U128 incl[kCpt];
                U128 run = u128(0);
                for (int c = 0; c < kCpt; ++c) {
                    run = run + chunk[(size_t)T * kCpt + c];
                    incl[c] = run;
                }
                U128 tile_base = coff[g];
                if (t > 0) tile_base = tile_base + tprefix[T - 1];
                const U128 C_before = tile_base + (warp > 0 ? incl[warp - 1] : u128(0)), C_after = tile_base + incl[warp];
                const long jc = t * kTile + (long)warp * kCh;
                if (jc >= M) continue;
                const bool has_last = (long)g * M + std::min<long>(jc + kCh, M) - 1 == Ns - 1;
                if (C_after == C_before && !has_last) continue;
                const int e_begin = rel_end(C_before);
                if (e_begin == nblk) break;
                const int e_end = has_last ? nblk : rel_end(C_after);
                if (e_end == 0) continue;
                // the 32 lanes: thread-local float64 running sums, then the warp scan in the kernel's order
                double cd[32][kItems], tot[32];
                uint64_t av[32][kItems];
                for (int lane = 0; lane < 32; ++lane) {
                    double r = 0.0;
                    for (int k = 0; k < kItems; ++k) {
                        const long j = jc + lane * kItems + k;
                        av[lane][k] = j < M ? a[(long)g * M + j] : 0;
                        r += (double)av[lane][k];
                        cd[lane][k] = r;
                    }
                    tot[lane] = r;
                }
                double sc[32];
                for (int lane = 0; lane < 32; ++lane) sc[lane] = tot[lane];
                for (int o = 1; o < 32; o <<= 1) {
                    double nx[32];
                    for (int lane = 0; lane < 32; ++lane) nx[lane] = lane >= o ? sc[lane] + sc[lane - o] : sc[lane];
                    for (int lane = 0; lane < 32; ++lane) sc[lane] = nx[lane];
                }
                int prev_last = e_begin;
                unsigned __int128 Cint = u128_int(C_before);
                for (int lane = 0; lane < 32; ++lane) {
                    const double before_d = sc[lane] - tot[lane] + u128_to_double(C_before);
                    int e[kItems];
                    unsigned need = 0;
                    auto fast = [&](int k) -> int {
                        const int r = rel_end_fast(rctx, (before_d + cd[lane][k]) * div.scale);
                        if (r < 0) need |= 1u << k;
                        return r;
                    };
                    const int e_lastk = fast(kItems - 1);
                    const int e_first = e_lastk == 0 ? 0 : fast(0);
                    if ((e_lastk == 0 || e_first == nblk) && need == 0) {
                        for (int k = 0; k < kItems; ++k) e[k] = e_first;
                        n_fast += 2;
                    } else {
                        e[0] = e_first;
                        for (int k = 1; k < kItems - 1; ++k) e[k] = fast(k);
                        e[kItems - 1] = e_lastk;
                        n_fast += kItems;
                    }
                    unsigned __int128 Ck = Cint;
                    for (int k = 0; k < kItems; ++k) {
                        Ck += av[lane][k];
                        if (need & (1u << k)) {
                            e[k] = rel_end_of(spos, N, rctx, cdf_exact(u128_from(Ck), div));
                            ++n_exact;
                        }
                    }
                    Cint = Ck;
                    for (int k = 0; k < kItems; ++k)
                        if ((long)g * M + jc + lane * kItems + k == Ns - 1) e[k] = nblk;
                    int start = prev_last;
                    const int wbase = (int)((T - T_lo) * kTile + warp * kCh + lane * kItems);
                    for (int k = 0; k < kItems; ++k) {
                        if (e[k] > start) head[start] = wbase + k + 1;
                        start = e[k];
                    }
                    prev_last = e[kItems - 1];
                }
            }
        }
        int mx = 0;
        for (int k = 0; k < nblk; ++k) {
            mx = std::max(mx, head[k]);
            const int wi = std::max(mx - 1, 0);
            const long T = T_lo + wi / kTile;
            const int g = (int)(T / tiles);
            long j = (T - (long)g * tiles) * kTile + (wi % kTile);
            if (j > M - 1) j = M - 1;
            anc[k0 + k] = (int32_t)((long)g * M + j);
        }
    }
    if (stats) {
        stats[0] = n_fast;
        stats[1] = n_exact;
    }
}

void hs_select_source_side(const uint64_t *a, long M, int G, double u, int32_t *anc, long *stats)
{
    hs_select_source_side_n(a, M, G, u, (long)G * M, anc, stats);
}

// the N_eff gate's integer sums over the integer weights, accumulated like k_weights does
double hs_neff_fixed(const uint64_t *a, long n)
{
    uint64_t s1 = 0, lo = 0, hi = 0;
    for (long i = 0; i < n; ++i) {
        const uint64_t s = neff_s(a[i]);
        s1 += s;
        lo += (s * s) & 0xFFFFFFFFull;
        hi += (s * s) >> 32;
    }
    return neff_fixed_a(s1, lo, hi);
}
}
