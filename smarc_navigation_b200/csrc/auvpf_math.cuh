// auvpf_math.cuh -- per-element math of the particle-filter cycle.
//
// Everything here is __host__ __device__ so that the SAME source is compiled
// into the CUDA kernels (auvpf_kernels.cuh) and into the test-only host
// simulator (tests/hostsim), which lets the CPU test-suite check the kernel
// arithmetic against the oracle on a box without a GPU.  The product library
// only ever runs the __device__ instantiations.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <type_traits>

#if defined(__CUDACC__)
#define AUVPF_HD __host__ __device__ __forceinline__
#else
#define AUVPF_HD inline
#endif

namespace auvpf {

constexpr float kPiF = 3.14159265358979323846f;
constexpr double kPi = 3.14159265358979323846;
constexpr int kQBits = 52;   // weights live on a 2^-52 grid
constexpr int kABits = 60;   // a_i = floor(exp(logw_i - max) * 2^60)

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  Counter layout of the device RNG contract
// (DESIGN.md section 4): [id_lo, id_hi, step, stream*16 + call], key [seed_lo, seed_hi].
struct U4 {
    uint32_t x, y, z, w;
};

AUVPF_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
{
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

AUVPF_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(0xD2511F53u, c.x, hi0, lo0);
        mulhilo32(0xCD9E8D57u, c.z, hi1, lo1);
        U4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

enum RngStream : uint32_t { kStreamInit = 1, kStreamPredict = 2, kStreamRoughen = 3, kStreamResample = 4 };

AUVPF_HD U4 philox_for(uint64_t seed, int64_t id, uint32_t step, uint32_t stream, uint32_t call)
{
    U4 c;
    c.x = (uint32_t)((uint64_t)id & 0xFFFFFFFFu);
    c.y = (uint32_t)((uint64_t)id >> 32);
    c.z = step;
    c.w = stream * 16u + call;
    return philox4x32_10(c, (uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32));
}

// Box-Muller on two 32-bit words: u = (x + 0.5) 2^-32,
//   n0 = sqrt(-2 ln u0) cos(2 pi u1),  n1 = sqrt(-2 ln u0) sin(2 pi u1).
// Device: special-function-unit versions (the library logf / sincospif were 25 % of the gather
// kernel's instructions).  -ln u0 is MUFU.LG2 except next to u0 = 1, where its ABSOLUTE error
// would dominate the small result: there 1 - u0 is formed exactly from the complemented integer
// and -ln(1 - t) is summed as a series.  The angle is shifted into [-pi, pi), where
// __sinf / __cosf keep 2^-21 absolute accuracy, and the sign folded back.
AUVPF_HD void box_muller(uint32_t a, uint32_t b, float &n0, float &n1)
{
#if defined(__CUDA_ARCH__)
    const float u0 = ((float)a + 0.5f) * 2.3283064365386963e-10f;
    const float t = ((float)(~a) + 0.5f) * 2.3283064365386963e-10f;          // 1 - u0
    const float series = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 0.2f, 0.25f), 0.33333334f), 0.5f), 1.0f);
    const float L = a >= 0xF8000000u ? series : -__logf(u0);                 // -ln u0 > 0
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(2.0f * L));              // one MUFU; 2^-23 relative
    const float ang = ((float)b + 0.5f) * 1.4629180792671596e-9f - kPiF;     // 2 pi u1 - pi
    float s, c;
    __sincosf(ang, &s, &c);
    n0 = -r * c;
    n1 = -r * s;
#else
    float u0 = ((float)a + 0.5f) * 2.3283064365386963e-10f;
    float u1 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
    float r = sqrtf(-2.0f * logf(u0));
    n0 = r * cosf(2.0f * kPiF * u1);
    n1 = r * sinf(2.0f * kPiF * u1);
#endif
}

// The six standard normals of one add_noise (auv_particle.py:32-36) for particle `id`, in POSE
// order [x, y, z, roll, pitch, yaw].  Contract: call 0 -> words (0,1): x, y; words (2,3): yaw, z;
// call 1 -> words (0,1): roll, pitch.  x, y and yaw -- the only components the reference's
// default covariances excite (auv_pf.launch:17-19) -- come from ONE Philox call; `need` (bit c =
// component c has non-zero sigma) lets the caller skip what is multiplied by zero anyway.
AUVPF_HD void pose_normals(uint64_t seed, int64_t id, uint32_t step, uint32_t stream, unsigned need, float n[6])
{
#pragma unroll
    for (int c = 0; c < 6; ++c) n[c] = 0.f;
    if (need & 0x27u) {                      // x, y, z or yaw
        const U4 r0 = philox_for(seed, id, step, stream, 0);
        if (need & 0x03u) box_muller(r0.x, r0.y, n[0], n[1]);
        if (need & 0x24u) box_muller(r0.z, r0.w, n[5], n[2]);
    }
    if (need & 0x18u) {                      // roll, pitch
        const U4 r1 = philox_for(seed, id, step, stream, 1);
        box_muller(r1.x, r1.y, n[3], n[4]);
    }
}

AUVPF_HD double uniform53(U4 r)
{
    uint64_t v = ((uint64_t)r.x << 32) | r.y;
    return (double)(v >> 11) * 1.1102230246251565e-16;   // 2^-53
}

// ---------------------------------------------------------------------------------------------
// Angles
AUVPF_HD float wrap_pi(float a)
{
    // (a + pi) % (2 pi) - pi of auv_particle.py:48 without the add/subtract round trip
    if (a >= kPiF) {
        a -= 2.0f * kPiF;
        if (a >= kPiF) a = (float)(fmod((double)a + kPi, 2.0 * kPi) - kPi);
    } else if (a < -kPiF) {
        a += 2.0f * kPiF;
        if (a < -kPiF) {
            double t = fmod((double)a + kPi, 2.0 * kPi);
            if (t < 0.0) t += 2.0 * kPi;
            a = (float)(t - kPi);
        }
    }
    return a;
}

AUVPF_HD void sincos_f(float a, float &s, float &c)
{
#if defined(__CUDA_ARCH__)
    sincosf(a, &s, &c);
#else
    s = sinf(a);
    c = cosf(a);
#endif
}

// Rz(yaw) Ry(pitch) Rx(roll): quaternion_matrix(quaternion_from_euler(r,p,y)), auv_particle.py:75-76
AUVPF_HD void rpy_matrix(float roll, float pitch, float yaw, float R[9])
{
    float sr, cr, sp, cp, sy, cy;
    sincos_f(roll, sr, cr);
    sincos_f(pitch, sp, cp);
    sincos_f(yaw, sy, cy);
    R[0] = cy * cp;
    R[1] = cy * sp * sr - sy * cr;
    R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp;
    R[4] = sy * sp * sr + cy * cr;
    R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;
    R[7] = cp * sr;
    R[8] = cp * cr;
}

// Same in float64: one call per particle per update, amortised over its B beams.  Keeps the
// beam directions good to ~6e-8 so that 100 m grazing rays stay inside the 1e-4 m tolerance.
AUVPF_HD void rpy_matrix_d(double roll, double pitch, double yaw, double R[9])
{
    double sr, cr, sp, cp, sy, cy;
#if defined(__CUDA_ARCH__)
    sincos(roll, &sr, &cr);
    sincos(pitch, &sp, &cp);
    sincos(yaw, &sy, &cy);
#else
    sr = sin(roll); cr = cos(roll);
    sp = sin(pitch); cp = cos(pitch);
    sy = sin(yaw); cy = cos(yaw);
#endif
    R[0] = cy * cp;
    R[1] = cy * sp * sr - sy * cr;
    R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp;
    R[4] = sy * sp * sr + cy * cr;
    R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;
    R[7] = cp * sr;
    R[8] = cp * cr;
}

// ---------------------------------------------------------------------------------------------
// Propagate: Particle.motion_pred (auv_particle.py:38-70).  Host precomputes, per
// filter, everything that is loop-invariant in the reference's per-particle loop.
struct PredictParams {
    float a, b;        // R(roll,pitch) (v dt) folded: step_x = cy a - sy b, step_y = sy a + cy b
    float wzdt;        // omega_z * dt                                   (:43-47)
    float roll, pitch; // euler_from_quaternion(odom orientation)[0:2]   (:50-56)
    float z;           // odom depth                                     (:70)
    float sx, sy, syaw;// sqrt(process_cov[0,1,5])                       (:40)
    float pad;
};

AUVPF_HD void propagate_one(const PredictParams &p, double &x, double &y, float &z, float &roll, float &pitch,
                            float &yaw, float n0, float n1, float n5)
{
    float yw = wrap_pi(yaw + p.wzdt + p.syaw * n5);          // :47-48
    float s, c;
#if defined(__CUDA_ARCH__)
    __sincosf(yw, &s, &c);       // yw is in [-pi, pi): 2^-21 absolute accuracy, ~1e-8 m on a 2 cm step
#else
    sincos_f(yw, s, c);
#endif
    float stepx = c * p.a - s * p.b + p.sx * n0;             // rows 0,1 of fullRotation (:64-65)
    float stepy = s * p.a + c * p.b + p.sy * n1;
    x += (double)stepx;                                      // :67
    y += (double)stepy;                                      // :68
    z = p.z;                                                 // :70
    roll = p.roll;                                           // :57
    pitch = p.pitch;
    yaw = yw;
}

// ---------------------------------------------------------------------------------------------
// Pose -> map frame: Particle.get_p_pose (auv_particle.py:72-83)
struct FrameParams {
    double m2o_t[3];
    double m2o_r[9];
    int m2o_identity_r;  // skip the 3x3 product when the map<-odom rotation is I
};

AUVPF_HD void particle_frame(const FrameParams &f, double x, double y, float z, float roll, float pitch, float yaw,
                             double p[3], double R[9])
{
    double Rl[9];
    rpy_matrix_d((double)roll, (double)pitch, (double)yaw, Rl);
    if (f.m2o_identity_r) {
        p[0] = x + f.m2o_t[0];
        p[1] = y + f.m2o_t[1];
        p[2] = (double)z + f.m2o_t[2];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = Rl[i];
    } else {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            p[a] = f.m2o_r[3 * a] * x + f.m2o_r[3 * a + 1] * y + f.m2o_r[3 * a + 2] * (double)z + f.m2o_t[a];
#pragma unroll
            for (int b = 0; b < 3; ++b)
                R[3 * a + b] = f.m2o_r[3 * a] * Rl[b] + f.m2o_r[3 * a + 1] * Rl[3 + b] + f.m2o_r[3 * a + 2] * Rl[6 + b];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Heightfield raycast (spec: oracle/c/auvpf_oracle.c RAYCAST SPEC).
//
// The DEM is stored with a one-vertex replicated border (pitch W + 2) so that a cell index
// one outside the map can be read without going out of bounds; `z` points at vertex (0, 0)
// inside the padding.
//
// Traversal = ONE uniform loop of "node steps".  A node is either a max-mip block (levels
// lv[0] < lv[1] of the pyramid, block = 2^shift cells, value = max over the block's closed
// vertex footprint) or, below lv[0], a single cell (max of its 4 corners, computed on the
// fly).  A step computes the node's exit time; if the ray stays above the node's maximum
// over [t, t_exit] the node is skipped (and the level raised again when a coarser boundary is
// crossed), otherwise the level is lowered; a cell that cannot be skipped becomes PENDING for
// the exact two-triangle test.  Because the level is data, not control flow, all lanes of a
// warp execute the same instructions whatever level each ray is at.
constexpr int kMaxLevels = 2;

struct DemView {
    const float *z;      // vertex (0,0); row pitch `pitch` floats; rows/cols -1 .. H / W are readable
    int W, H, pitch;
    int last_cx, last_cy;   // W - 2, H - 2: the last cell column / row (a kernel parameter the march compares against directly)
    const float *mip[kMaxLevels];   // mip[0] = finer block level, mip[1] = coarser (or unused)
    int mipW[kMaxLevels], mipH[kMaxLevels], shift[kMaxLevels];
    int n_levels;        // 1 or 2
    float range_max;
    float zmax;          // global maximum height
    float max_slope;     // largest triangle gradient norm |grad h| (Lipschitz constant of the surface)
    const float *slope_blk;   // the same per block of kSlopeBlock x kSlopeBlock cells, [sbH][sbW] (local_max_slope)
    int sbW, sbH;
    float res;           // cell size [m]
    double x0, y0, inv_res_d;
};

struct Ray {
    int ix0, iy0;        // origin cell (absolute column / row of the local frame's origin)
    float lx0, ly0;      // origin in local grid units, in [0,1) when inside the map
    float oz;
    float dgx, dgy, dz;  // direction in grid units / metres
    float oz_lo, dgx_lo, dgy_lo, dz_lo;   // what float32 dropped of oz and of the direction: read only by the float64
                                          // re-solve of a grazing hit (precise_ray_hit), where 1e-6 m of sensor height
                                          // is 2e-4 m of range
};

AUVPF_HD float ldg_f(const float *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

AUVPF_HD float rcp_rn(float x)
{
#if defined(__CUDA_ARCH__)
    return __frcp_rn(x);      // correctly rounded, same value as the host's 1.0f / x
#else
    return 1.0f / x;
#endif
}

// a / b where 2^-23 relative accuracy is enough (profile slopes and segment lengths; the results
// that must be exact are re-derived in float64).  Device: ONE MUFU.RCP and a multiply --
// __fdividef's range fix-ups were 11 % of the profile kernel's instructions.
AUVPF_HD float fast_div(float a, float b)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return a * r;
#else
    return a / b;
#endif
}

AUVPF_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

constexpr int kSlopeShift = 7, kSlopeBlock = 1 << kSlopeShift;     // 128 x 128 cells: a 120 m reach on a 0.5 m map touches 5 x 5 blocks

// Largest triangle gradient norm among the cells within `radius` cells of the grid position (gx, gy).
AUVPF_HD float local_max_slope(const DemView &m, double gx, double gy, double radius)
{
    const int j0 = clampi((int)floor(gx - radius) >> kSlopeShift, 0, m.sbW - 1), j1 = clampi((int)floor(gx + radius) >> kSlopeShift, 0, m.sbW - 1);
    const int i0 = clampi((int)floor(gy - radius) >> kSlopeShift, 0, m.sbH - 1), i1 = clampi((int)floor(gy + radius) >> kSlopeShift, 0, m.sbH - 1);
    float s = 0.f;
    for (int i = i0; i <= i1; ++i)
        for (int j = j0; j <= j1; ++j) s = fmaxf(s, ldg_f(m.slope_blk + (size_t)i * m.sbW + j));
    return s;
}


constexpr float kBigT = 3.0e38f;
constexpr float kSkipMargin = 2.0e-3f;   // [m] slack that keeps max skipping conservative in fp32

struct Trav {
    // per-ray constants
    float ix, iy;        // 1 / dgx, 1 / dgy (0 when the ray never crosses that family)
    float te;            // end of the clipped ray
    int bx, by;          // 1 when moving in +x / +y
    // state
    float t;             // entry time into the current node
    int cx, cy;          // absolute level-0 cell entered at t
    int li;              // level index: n_levels-1 .. 0 = mip levels, -1 = cells
    // pending exact test (valid when the step returned kPending)
    float t_out;
    float z00, z01, z10, z11;
};

enum StepResult : int { kContinue = 0, kPending = 1, kMiss = 2 };

// Clip the ray and place the traversal at its start.  Returns false (with `range` set) when
// the ray is resolved without traversal.  `t_lo` is a CERTIFIED clear prefix: the caller
// guarantees f > 0 on [0, t_lo) (fan sweep, see sweep_bound); 0 = none.
// `interior` = the map covers the sensor's whole reach (fan_sweep_valid): no clipping needed.
AUVPF_HD bool trav_init(Trav &s, const DemView &m, const Ray &r, float &range, float t_lo = 0.f,
                        bool interior = false)
{
    s.ix = r.dgx != 0.f ? rcp_rn(r.dgx) : 0.f;
    s.iy = r.dgy != 0.f ? rcp_rn(r.dgy) : 0.f;
    s.bx = r.dgx > 0.f ? 1 : 0;
    s.by = r.dgy > 0.f ? 1 : 0;
    float ts = 0.f, te = m.range_max;
    range = m.range_max;
    if (!interior) {
        float lo = -(float)r.ix0, hi = (float)(m.W - 1 - r.ix0);
        if (r.dgx != 0.f) {
            float t1 = (lo - r.lx0) * s.ix, t2 = (hi - r.lx0) * s.ix;
            ts = fmaxf(ts, fminf(t1, t2));
            te = fminf(te, fmaxf(t1, t2));
        } else if (r.lx0 < lo || r.lx0 > hi)
            return false;
        lo = -(float)r.iy0;
        hi = (float)(m.H - 1 - r.iy0);
        if (r.dgy != 0.f) {
            float t1 = (lo - r.ly0) * s.iy, t2 = (hi - r.ly0) * s.iy;
            ts = fmaxf(ts, fminf(t1, t2));
            te = fminf(te, fmaxf(t1, t2));
        } else if (r.ly0 < lo || r.ly0 > hi)
            return false;
    }
    ts = fmaxf(ts, t_lo);
    if (!(ts <= te)) return false;
    // nothing can be hit above the global maximum height
    if (fmaf(ts, r.dz, r.oz) > m.zmax + kSkipMargin) {
        if (r.dz >= 0.f) return false;
        float tz = (m.zmax + kSkipMargin - r.oz) * rcp_rn(r.dz);
        if (tz >= te) return false;
        ts = fmaxf(ts, tz);
    }
    s.te = te;
    s.t = ts;
    s.cx = clampi(r.ix0 + (int)floorf(fmaf(ts, r.dgx, r.lx0)), 0, m.W - 2);
    s.cy = clampi(r.iy0 + (int)floorf(fmaf(ts, r.dgy, r.ly0)), 0, m.H - 2);
    s.li = t_lo > 0.f ? -1 : m.n_levels - 1;   // a swept ray starts next to its hit: look at cells first
    return true;
}

// One node step (see the header comment).  kPending leaves the cell corners and its exit
// time in `s` for trav_exact; kMiss means the ray ended without a hit.
AUVPF_HD int trav_step(Trav &s, const DemView &m, const Ray &r)
{
    const int li = s.li;
    const int sh = li < 0 ? 0 : (li == 0 ? m.shift[0] : m.shift[1]);
    const int nx = s.cx >> sh, ny = s.cy >> sh;
    const int X = (nx + s.bx) << sh, Y = (ny + s.by) << sh;      // node's far grid lines (absolute)
    const float tX = s.ix != 0.f ? ((float)(X - r.ix0) - r.lx0) * s.ix : kBigT;
    const float tY = s.iy != 0.f ? ((float)(Y - r.iy0) - r.ly0) * s.iy : kBigT;
    const float t_out = fminf(fminf(tX, tY), s.te);
    float nodemax;
    if (li >= 0) {
        const float *mp = li == 0 ? m.mip[0] : m.mip[1];
        const int mw = li == 0 ? m.mipW[0] : m.mipW[1], mh = li == 0 ? m.mipH[0] : m.mipH[1];
        nodemax = ldg_f(mp + (size_t)clampi(ny, 0, mh - 1) * mw + clampi(nx, 0, mw - 1));
    } else {
        const float *p = m.z + (ptrdiff_t)s.cy * m.pitch + s.cx;
        s.z00 = ldg_f(p);
        s.z01 = ldg_f(p + 1);
        s.z10 = ldg_f(p + m.pitch);
        s.z11 = ldg_f(p + m.pitch + 1);
        nodemax = fmaxf(fmaxf(s.z00, s.z01), fmaxf(s.z10, s.z11));
    }
    const float zlow = fmaf(r.dz < 0.f ? t_out : s.t, r.dz, r.oz);
    if (zlow > nodemax + kSkipMargin) {
        // skip the node
        if (t_out >= s.te) return kMiss;
        int ncx, ncy;
        if (tX <= tY) {
            ncx = s.bx ? X : X - 1;
            ncy = s.cy;
            if (sh > 0) {
                int c = r.iy0 + (int)floorf(fmaf(t_out, r.dgy, r.ly0));
                ncy = clampi(c, ny << sh, ((ny + 1) << sh) - 1);    // exit through x: still inside the node's rows
            }
        } else {
            ncy = s.by ? Y : Y - 1;
            ncx = s.cx;
            if (sh > 0) {
                int c = r.ix0 + (int)floorf(fmaf(t_out, r.dgx, r.lx0));
                ncx = clampi(c, nx << sh, ((nx + 1) << sh) - 1);
            }
        }
        // raise the level again when a coarser block boundary was crossed
        int nli = li;
        if (li + 1 < m.n_levels) {
            const int ps = li + 1 == 0 ? m.shift[0] : m.shift[1];
            if (((ncx >> ps) != (s.cx >> ps)) || ((ncy >> ps) != (s.cy >> ps))) {
                nli = li + 1;
                if (nli + 1 < m.n_levels) {
                    const int ps2 = m.shift[1];
                    if (((ncx >> ps2) != (s.cx >> ps2)) || ((ncy >> ps2) != (s.cy >> ps2))) nli = nli + 1;
                }
            }
        }
        s.li = nli;
        s.cx = clampi(ncx, -1, m.W - 1);
        s.cy = clampi(ncy, -1, m.H - 1);
        s.t = t_out;
        return kContinue;
    }
    if (li >= 0) {
        s.li = li - 1;     // look closer
        return kContinue;
    }
    s.t_out = t_out;
    return kPending;
}

AUVPF_HD float tri_height4(float z00, float z01, float z10, float z11, float u, float v)
{
    return (u >= v) ? z00 + u * (z01 - z00) + v * (z11 - z01) : z00 + u * (z11 - z10) + v * (z10 - z00);
}

// A ray that meets a facet at a grazing angle: the float32 evaluation of f at the two ends of the sub-segment
// (~5e-6 m each: differences of heights of tens of metres, end points rounded to 1e-6 cells) is divided by the
// closing rate, and 0.02 m/m turns that into 2.5e-4 m of range.  Such hits are re-solved on the facet's plane,
//   o_z + t d_z = z00 + u(t) Du + v(t) Dv,  u(t) = lx0 - cxl + t dgx,  v(t) = ly0 - cyl + t dgy,
// in float64 from the (exact) float32 ray and vertex heights.  (cxl, cyl) = the cell relative to the ray's origin cell.
constexpr float kRayGraze = 0.25f;       // closing rate [m of height per m of range] below which the hit is re-solved

#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
float precise_ray_hit(float z00, float z01, float z10, float z11, bool lower, float cxl, float cyl, const Ray *r, float t_lo,
                      float t_hi)
{
    const double Du = lower ? (double)z01 - (double)z00 : (double)z11 - (double)z10;
    const double Dv = lower ? (double)z11 - (double)z01 : (double)z10 - (double)z00;
    const double num = (double)z00 + ((double)r->lx0 - (double)cxl) * Du + ((double)r->ly0 - (double)cyl) * Dv -
                       ((double)r->oz + (double)r->oz_lo);
    const double den = ((double)r->dz + (double)r->dz_lo) - ((double)r->dgx + (double)r->dgx_lo) * Du -
                       ((double)r->dgy + (double)r->dgy_lo) * Dv;
    const double t = num / den;
    return (float)(t < (double)t_lo ? (double)t_lo : (t > (double)t_hi ? (double)t_hi : t));   // (NaN: den = 0 -> t_hi... see caller)
}

// Exact two-triangle test of one cell over the ray segment [ta, tb]; (ua,va), (ub,vb) are the
// segment end points in cell coordinates.  Returns true and the range when f has its first
// root in the segment (or f(ta) <= 0, spec: range = ta).
AUVPF_HD bool exact_segment(float z00, float z01, float z10, float z11, float ua, float va, float ub, float vb,
                            float ta, float tb, const Ray &r, float &range, float cxl, float cyl)
{
    const float fa = fmaf(ta, r.dz, r.oz) - tri_height4(z00, z01, z10, z11, ua, va);
    if (fa <= 0.f) {
        range = ta;
        return true;
    }
    const float fb = fmaf(tb, r.dz, r.oz) - tri_height4(z00, z01, z10, z11, ub, vb);
    const float ga = ua - va, gb = ub - vb;
    float t0 = ta, f0 = fa, t1 = tb, f1 = fb;
    bool hit;
    if (ga * gb < 0.f) {             // the segment crosses the cell diagonal
        const float sd = fast_div(ga, ga - gb);
        const float td = fmaf(sd, tb - ta, ta);
        const float ud = fmaf(sd, ub - ua, ua);
        const float fd = fmaf(td, r.dz, r.oz) - fmaf(ud, z11 - z00, z00);
        if (fd <= 0.f) {
            t1 = td;
            f1 = fd;
            hit = true;
        } else {
            t0 = td;
            f0 = fd;
            hit = fb <= 0.f;
        }
    } else
        hit = fb <= 0.f;
    if (hit) {
        range = t1 > t0 ? fmaf(t1 - t0, fast_div(f0, f0 - f1), t0) : t0;
        if (t1 > t0 && f0 - f1 < kRayGraze * (t1 - t0)) {
            // which half the sub-segment [t0, t1] lies in: its middle decides (the diagonal, if crossed, is one of its ends)
            const float wa = fmaf(0.5f * (t0 + t1) - ta, fast_div(ga - gb, ta - tb), ga);     // (u - v) at the middle
            const float tp = precise_ray_hit(z00, z01, z10, z11, wa >= 0.f, cxl, cyl, &r, t0, t1);
            if (tp == tp) range = tp;
        }
    }
    return hit;
}

// Exact test of the pending cell over [s.t, s.t_out].  Returns true with the range when the
// ray is resolved (hit, or miss at the end of the ray); otherwise advances into the next cell.
AUVPF_HD bool trav_exact(Trav &s, const DemView &m, const Ray &r, float &range)
{
    const float ta = s.t, tb = s.t_out;
    const float cxl = (float)(s.cx - r.ix0), cyl = (float)(s.cy - r.iy0);
    const float ua = fmaf(ta, r.dgx, r.lx0) - cxl, va = fmaf(ta, r.dgy, r.ly0) - cyl;
    const float ub = fmaf(tb, r.dgx, r.lx0) - cxl, vb = fmaf(tb, r.dgy, r.ly0) - cyl;
    if (exact_segment(s.z00, s.z01, s.z10, s.z11, ua, va, ub, vb, ta, tb, r, range, cxl, cyl)) return true;
    if (tb >= s.te) {                // end of the ray: miss
        range = m.range_max;
        return true;
    }
    // advance into the neighbouring cell (integer step, exact topology)
    const int X = s.cx + s.bx, Y = s.cy + s.by;
    const float tX = s.ix != 0.f ? ((float)(X - r.ix0) - r.lx0) * s.ix : kBigT;
    const float tY = s.iy != 0.f ? ((float)(Y - r.iy0) - r.ly0) * s.iy : kBigT;
    int ncx = s.cx, ncy = s.cy;
    if (tX <= tY) ncx = s.bx ? X : X - 1;
    else ncy = s.by ? Y : Y - 1;
    int nli = -1;
    const int ps = m.shift[0];
    if (((ncx >> ps) != (s.cx >> ps)) || ((ncy >> ps) != (s.cy >> ps))) {
        nli = 0;
        if (m.n_levels > 1) {
            const int ps2 = m.shift[1];
            if (((ncx >> ps2) != (s.cx >> ps2)) || ((ncy >> ps2) != (s.cy >> ps2))) nli = 1;
        }
    }
    s.li = nli;
    s.cx = clampi(ncx, -1, m.W - 1);
    s.cy = clampi(ncy, -1, m.H - 1);
    s.t = tb;
    return false;
}

// Tight cell-by-cell walk for a ray that starts next to its hit (fan sweep): incremental
// DDA over cells, 4-corner max pre-check, exact test.  Gives up after kNearCells cells and
// hands the state back to the hierarchical traversal.
constexpr int kNearCells = 12;

AUVPF_HD bool near_walk(Trav &s, const DemView &m, const Ray &r, float &range)
{
    float t = s.t;
    int cx = s.cx, cy = s.cy;
    float cxl = (float)(cx - r.ix0), cyl = (float)(cy - r.iy0);
    const float sxf = s.bx ? 1.f : -1.f, syf = s.by ? 1.f : -1.f;
    const float bxf = s.bx ? 1.f : 0.f, byf = s.by ? 1.f : 0.f;
    const int sxi = s.bx ? 1 : -1, syi = s.by ? 1 : -1;
    float tx = s.ix != 0.f ? ((cxl + bxf) - r.lx0) * s.ix : kBigT;
    float ty = s.iy != 0.f ? ((cyl + byf) - r.ly0) * s.iy : kBigT;
    const float *p = m.z + (ptrdiff_t)cy * m.pitch + cx;
    for (int n = 0; n < kNearCells; ++n) {
        const float t_out = fminf(fminf(tx, ty), s.te);
        const float z00 = ldg_f(p), z01 = ldg_f(p + 1), z10 = ldg_f(p + m.pitch), z11 = ldg_f(p + m.pitch + 1);
        const float zmax4 = fmaxf(fmaxf(z00, z01), fmaxf(z10, z11));
        const float zlow = fmaf(r.dz < 0.f ? t_out : t, r.dz, r.oz);
        if (!(zlow > zmax4 + kSkipMargin)) {
            const float ua = fmaf(t, r.dgx, r.lx0) - cxl, va = fmaf(t, r.dgy, r.ly0) - cyl;
            const float ub = fmaf(t_out, r.dgx, r.lx0) - cxl, vb = fmaf(t_out, r.dgy, r.ly0) - cyl;
            if (exact_segment(z00, z01, z10, z11, ua, va, ub, vb, t, t_out, r, range, cxl, cyl)) return true;
        }
        if (t_out >= s.te) {
            range = m.range_max;
            return true;
        }
        if (tx <= ty) {
            cxl += sxf;
            cx += sxi;
            p += sxi;
            tx = ((cxl + bxf) - r.lx0) * s.ix;
        } else {
            cyl += syf;
            cy += syi;
            p += syi * m.pitch;
            ty = ((cyl + byf) - r.ly0) * s.iy;
        }
        t = t_out;
        if (cx < 0 || cy < 0 || cx > m.W - 2 || cy > m.H - 2) break;   // left the map by rounding: let the generic path finish
    }
    s.t = t;
    s.cx = clampi(cx, -1, m.W - 1);
    s.cy = clampi(cy, -1, m.H - 1);
    s.li = -1;
    return false;
}

// Scalar drivers (host simulator; swept rays in the kernels).  The kernels run FULL traversals
// through the warp-cooperative driver in auvpf_kernels.cuh around the same functions.
AUVPF_HD float raycast_from(const DemView &m, const Ray &r, float t_lo, bool interior = false)
{
    Trav s;
    float range;
    if (!trav_init(s, m, r, range, t_lo, interior)) return range;
    if (t_lo > 0.f && near_walk(s, m, r, range)) return range;
    for (;;) {
        int res = trav_step(s, m, r);
        if (res == kMiss) return m.range_max;
        if (res == kPending && trav_exact(s, m, r, range)) return range;
    }
}

AUVPF_HD float raycast_one(const DemView &m, const Ray &r) { return raycast_from(m, r, 0.f); }

// ---------------------------------------------------------------------------------------------
// Fan sweep.  When the beams form a planar fan d_b = cos(th_b) c0 +
// sin(th_b) a0 and the height above the terrain is strictly decreasing along the fan's
// centre direction c = R c0 (true when c_z + max_slope |c_xy| < 0), the terrain profile in the
// fan plane is a single-valued function of the lateral coordinate p = t sin(th).  Then a ray
// that is clear of the terrain up to lateral distance p_clear certifies every ray further out
// on the same side (|th'| > |th|) to be clear up to the same lateral distance: at equal p the
// outer ray is higher.  So beams are swept outward and each starts at
// t_lo = (p_clear - margin) / |sin th'| instead of at the sensor.
// The terrain must also be DEFINED wherever the fan can reach (outside the map there is no
// surface, and "clear because outside" certifies nothing for a ray that is inside), so the
// sweep is only used when the whole reach of the sensor lies inside the map footprint.
AUVPF_HD bool fan_sweep_valid(const DemView &m, const double p[3], const double R[9], const float c0[3])
{
    const double reach = (double)m.range_max;
    const double xmax = m.x0 + (double)(m.W - 1) * (double)m.res, ymax = m.y0 + (double)(m.H - 1) * (double)m.res;
    if (!(p[0] - reach > m.x0 && p[0] + reach < xmax && p[1] - reach > m.y0 && p[1] + reach < ymax)) return false;
    const float max_slope = m.max_slope;
    const double cx = R[0] * c0[0] + R[1] * c0[1] + R[2] * c0[2];
    const double cy = R[3] * c0[0] + R[4] * c0[1] + R[5] * c0[2];
    const double cz = R[6] * c0[0] + R[7] * c0[1] + R[8] * c0[2];
    return cz + (double)max_slope * sqrt(cx * cx + cy * cy) < -0.05;
}

AUVPF_HD float sweep_bound(float p_clear, float inv_s)
{
    // small back-off for fp32 rounding of p_clear and near-planarity of the beam table; should
    // rounding still leave f(t_lo) <= 0, the hit lies within ~1e-5 m of t_lo (f(t_lo) >= 0 exactly)
    const float p = p_clear * (1.0f - 2.0e-5f) - 2.0e-4f;
    return p > 0.f ? p * inv_s : 0.f;
}

// Build the ray of beam direction d (sensor frame) for a particle frame (p, R).
AUVPF_HD Ray make_ray(const DemView &m, const double p[3], int ix0, int iy0, float lx0, float ly0, const double R[9],
                      float bx, float by, float bz)
{
    Ray r;
    r.ix0 = ix0;
    r.iy0 = iy0;
    r.lx0 = lx0;
    r.ly0 = ly0;
    r.oz = (float)p[2];
    double dx = R[0] * (double)bx + R[1] * (double)by + R[2] * (double)bz;
    double dy = R[3] * (double)bx + R[4] * (double)by + R[5] * (double)bz;
    double dz = R[6] * (double)bx + R[7] * (double)by + R[8] * (double)bz;
    r.dgx = (float)(dx * m.inv_res_d);
    r.dgy = (float)(dy * m.inv_res_d);
    r.dz = (float)dz;
    r.oz_lo = (float)(p[2] - (double)r.oz);
    r.dgx_lo = (float)(dx * m.inv_res_d - (double)r.dgx);
    r.dgy_lo = (float)(dy * m.inv_res_d - (double)r.dgy);
    r.dz_lo = (float)(dz - (double)r.dz);
    return r;
}

AUVPF_HD void origin_cell(const DemView &m, const double p[3], int &ix0, int &iy0, float &lx0, float &ly0)
{
    double gx = (p[0] - m.x0) * m.inv_res_d, gy = (p[1] - m.y0) * m.inv_res_d;
    double fx = floor(gx), fy = floor(gy);
    // keep the integer part bounded so far-away particles cannot overflow int
    fx = fx < -1.0e6 ? -1.0e6 : (fx > 1.0e6 ? 1.0e6 : fx);
    fy = fy < -1.0e6 ? -1.0e6 : (fy > 1.0e6 ? 1.0e6 : fy);
    ix0 = (int)fx;
    iy0 = (int)fy;
    lx0 = (float)(gx - fx);
    ly0 = (float)(gy - fy);
}

// ---------------------------------------------------------------------------------------------
// Profile merge: all beams of one side of a planar fan in
// O(cells + beams).  In fan-plane coordinates (p lateral, q along the fan centre c) the
// terrain is a single-valued profile q = T(p) that is LINEAR inside every DEM triangle, with
// slope  s = dT/dp = beta / gamma,  beta = a_g . grad h - a_z,  gamma = c_z - c_g . grad h
// (gamma < 0 when the height above the terrain decreases along c).
// Beam th is the line p = q tan th; it hits where it first meets the profile, and because a
// more outward beam is higher at equal p, the hits of th_0 < th_1 < ... occur at increasing p.
// So the profile is marched ONCE outward, triangle by triangle, and the beams are resolved in
// order on the way (a merge of two monotone sequences) -- the same first hits as casting each
// ray separately, occlusions included.
//
// The march is a DDA over the TRIANGULATION in cell-local coordinates: the state is the
// triangle (cell + which half), the curve point (u, v) inside the cell, and the fan
// coordinates (p_cur, q_cur) of that point in float64.  Inside a triangle the curve is the
// straight line (u, v) + dp (a_g + s c_g); it leaves through the nearest of the triangle's
// other two edges, and the neighbour across that edge is known exactly (no point location, no
// rounding-dependent topology).  Everything that is differenced stays small (cell-local), so
// float32 suffices except for the two accumulators and the beam's gap to the curve point.
// Needs: planar sorted fan, the map covering the sensor's reach, gamma < 0 on every triangle
// met (checked on the fly; otherwise the caller falls back to per-ray traversal).
struct ProfFrame {
    float agx, agy, az;    // lateral unit vector of this side: xy in grid units per metre, z
    float cgx, cgy, cz;    // fan centre direction, same units
};

struct ProfFrameD {        // the same frame in float64, read only by the rare re-anchored hit (precise_hit)
    double gox, goy, oz;   // sensor in absolute grid units, height
    double agx, agy, az;   // lateral unit vector of the FIRST side (the other side negates it)
    double cgx, cgy, cz;
};

struct ProfStart {         // the curve point under the sensor (p = 0), shared by both sides
    int cx, cy;            // its cell (absolute)
    float u, v;            // position inside the cell, [0, 1]
    double q0;             // centre-beam range, re-solved in float64 on that triangle's plane
};

#if defined(AUVPF_DEBUG_REASONS) && !defined(__CUDA_ARCH__)
extern long g_prof_reasons[8];      // test-only: why profile marches gave up (tests/hostsim)
#define AUVPF_REASON(k) (++g_prof_reasons[k])
#else
#define AUVPF_REASON(k) ((void)0)
#endif

constexpr float kGammaMin = 0.05f;       // -gamma must exceed this (|cos| of the worst slope/tilt combination)
constexpr float kDenPrecise = 0.25f;    // a beam closing on the profile slower than this is re-solved in float64

// A beam that meets a facet at a grazing angle amplifies the ~1e-6 m drift of the incremental
// (p_cur, q_cur) by 1 / den.  Such hits are re-solved from scratch on the facet's plane:
// q = alpha / (gamma - beta tan th) with everything in float64 relative to the sensor.
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
double precise_hit(const ProfFrameD *D, float sgf, int cx, int cy, float z00, float z01, float z10, float z11, bool lower,
                   double tn)
{
    // float32 arguments on purpose: with float64 ones the compiler hoisted six conversions and two
    // float64 subtractions out of the beam loop into EVERY step (ncu: the slowest instructions there)
    const double Du = lower ? (double)z01 - (double)z00 : (double)z11 - (double)z10;
    const double Dv = lower ? (double)z11 - (double)z01 : (double)z10 - (double)z00;
    const double sg = (double)sgf;
    const double alpha = ((double)z00 - D->oz) + (D->gox - (double)cx) * Du + (D->goy - (double)cy) * Dv;
    const double beta = sg * (D->agx * Du + D->agy * Dv - D->az);
    const double gamma = D->cz - D->cgx * Du - D->cgy * Dv;
    return alpha / (gamma - beta * tn);
}

// Warp convergence.  With independent thread scheduling, lanes that leave a data-dependent
// loop at different iterations are only reconverged where the compiler places a barrier; for the
// first version of the march it placed none inside the loop, so every lane ran its whole march ALONE
// (ncu: 1.03 threads per instruction, 17x slower).  The march is therefore written as a warp-uniform
// loop: every lane of the warp enters (lanes without work carry a finished state), the loop runs
// while ANY lane is still marching, and the ballot at its head reconverges the warp once per pair of steps
// (inside a step the compiler's own barriers around the beam loop do).
// profile_particle must be called by all 32 lanes of a warp.
#if defined(__CUDA_ARCH__)
#define AUVPF_WARP_BALLOT(p) __ballot_sync(0xffffffffu, (p))
#define AUVPF_WARP_SYNC() __syncwarp()
#else
#define AUVPF_WARP_BALLOT(p) ((p) ? 1u : 0u)
#define AUVPF_WARP_SYNC() ((void)0)
#endif

// Beam tables of a planar fan as the profile merge needs them (device: shared memory).
struct BeamAux {                             // one 16-byte load per beam: everything the beam loop reads
    float sec;                               // 1 / along-centre component (float32 copy)
    float reach;                             // |lateral component| * range_max: lateral distance the beam can cover
    float meas;                              // measured range of the ping; 0 for probes.  A beam of the ping without a
                                             // valid measurement has sec = 0 AND meas = 0: range 0, error 0.
    float tn;                                // |lateral| / along-centre component (float32 copy of tan_t)
};

// Per-beam table reads.  The march counts beams by ADDRESS: ba = base + 16 * beam, stepped by +-16, compared with
// the address one past the side's last beam -- the counter is the load address.  On the device the table is in
// shared memory and `base` is its 32-bit shared address (through the generic pointer the compiler re-derived
// the shared window base, S2R SR_CgaCtaId + LEA, in every trip of the beam loop: 3 % of the stall samples);
// on the host base = 0.
struct BeamTabs {
#if defined(__CUDA_ARCH__)
    uint32_t base;
#else
    const BeamAux *aux;
    static constexpr uint32_t base = 0;
#endif
};

// Table entries -1 and B: nothing compares true against a NaN, so the beam loop reads the entry past a side's last
// beam like any other, finds neither a hit nor a miss and stops -- it needs no end-of-table test of its own.
// (Its unused first word carries range_max: read back from shared memory the value is a plain register to the
// compiler, which otherwise re-reads the kernel parameter from the constant bank in every trip of the beam loop.)
AUVPF_HD BeamAux beam_sentinel(float range_max)
{
    BeamAux a;
    a.sec = range_max;
    a.reach = a.meas = a.tn = nanf("");
    return a;
}

AUVPF_HD uint32_t beam_addr(const BeamTabs &t, int b) { return t.base + 16u * (uint32_t)b; }
AUVPF_HD int beam_index(const BeamTabs &t, uint32_t ba) { return (int)(ba - t.base) >> 4; }

AUVPF_HD BeamAux beam_aux(const BeamTabs &t, uint32_t ba)
{
#if defined(__CUDA_ARCH__)
    BeamAux a;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.sec), "=f"(a.reach), "=f"(a.meas), "=f"(a.tn) : "r"(ba));
    return a;
#else
    return t.aux[(int32_t)ba >> 4];      // (entry -1 is the lower sentinel)
#endif
}

struct FanTables {
    const double *tan_t;                     // |lateral| / along-centre component of the beam direction
    const double *sec_t;                     // 1 / along-centre component (float64: grazing hits only)
    const BeamAux *aux;
    int B, n1, sgn1;                         // beams [0,n1) on the side sign sgn1, [n1,B) on the other
    float s_lim[2];                          // per side: profile slope up to which NO beam of the side is grazing
                                             // (den = 1 - slope tan >= kDenPrecise for the side's outermost tan)
    float c0[3], a0[3];                      // fan centre / lateral unit vectors, sensor frame
};

// The profile curve of one side left the map footprint while beams of that side are still unresolved
// (sensor closer to the border than its reach).  Those beams are clear of the terrain up to the lateral
// distance p_clear the curve was marched to -- at equal p they are above it -- so each is finished on its
// own: a ray clipped to the footprint (spec: no surface outside, a clipped ray that ends clear MISSES),
// started at the certified clear prefix and walked the few cells that remain to the border.  The ray is
// rebuilt from the float64 fan frame D (sensor in grid units, centre / lateral unit vectors):
// dir = cos th c + sgn sin th a.  (The map and the tables come BY VALUE: a reference would force the caller's copies --
// kernel parameters in the constant bank -- into local memory for the whole hot loop.)
template <class Sink>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
void finish_side_by_rays(const DemView m, const FanTables T, const ProfFrameD *D, float sgf, int j, int n_side, int bb,
                         int bstep, float p_clear, Sink &sink)
{
    const double fx = floor(D->gox), fy = floor(D->goy);
    Ray r;
    r.ix0 = (int)fx;
    r.iy0 = (int)fy;
    r.lx0 = (float)(D->gox - fx);
    r.ly0 = (float)(D->goy - fy);
    r.oz = (float)D->oz;
    r.oz_lo = (float)(D->oz - (double)r.oz);
    const double sg = (double)sgf;
    for (; j < n_side; ++j, bb += bstep) {
        const double tn = T.tan_t[bb], ct = 1.0 / T.sec_t[bb], st = tn * ct;
        const double dgx = ct * D->cgx + sg * st * D->agx, dgy = ct * D->cgy + sg * st * D->agy, dz = ct * D->cz + sg * st * D->az;
        r.dgx = (float)dgx;
        r.dgy = (float)dgy;
        r.dz = (float)dz;
        r.dgx_lo = (float)(dgx - (double)r.dgx);
        r.dgy_lo = (float)(dgy - (double)r.dgy);
        r.dz_lo = (float)(dz - (double)r.dz);
        const float t_lo = st > 0.0 ? sweep_bound(p_clear, (float)(1.0 / st)) : 0.f;
        const float range = raycast_from(m, r, t_lo, false);
        if (T.aux[bb].sec == 0.f) continue;          // no valid measurement on this beam: nothing to add (see BeamAux)
        sink(bb, range, T.aux[bb].meas);
    }
}

// Marches both sides of the fan, one after the other, inside ONE warp-uniform step loop (a lane
// that finishes its first side starts its second while its neighbours are still on their first,
// so the warp waits for the longest SUM of the two sides, not the sum of the longest).
// Calls sink(beam, range) for every beam in outward order per side.  Returns false if the
// particle cannot be certified (caller must recast its beams ray by ray).
// kBorder: the variant for sensors whose reach leaves the map (the curve of a side may run off the map and the
// side is then finished ray by ray).  It is a COMPILE-TIME variant because its bookkeeping costs the hot loop 15 %
// (30 more instructions per step and two more live values): interior particles -- normally all of them -- run the
// lean instantiation, the few near the border are listed and run the other one.
template <bool kBorder, class Sink, class ExitSink>
AUVPF_HD bool profile_march(const DemView &m, const FanTables &T, const ProfFrame &F0, const ProfFrameD *D,
                            const ProfStart &S, bool active, Sink sink, ExitSink &exit_sink)
{
    // per-side state
    int side = T.n1 > 0 ? 0 : 1;
    BeamTabs tabs;
#if defined(__CUDA_ARCH__)
    tabs.base = (uint32_t)__cvta_generic_to_shared(T.aux);
#else
    tabs.aux = T.aux;
#endif
    // current beam (as its table address, see BeamTabs), the step to the next one outward, and the address one past
    // the side's last beam: the side is done when ba == ba_end (one counter: the loop below is the hottest code of
    // the library)
    uint32_t ba = beam_addr(tabs, side == 0 ? T.n1 - 1 : T.n1), ba_end = beam_addr(tabs, side == 0 ? -1 : T.B);
    uint32_t bstep = side == 0 ? (uint32_t)-16 : 16u;
    float s_lim = T.s_lim[side];
    float sgf = (float)(side == 0 ? T.sgn1 : -T.sgn1);
    float lgx = sgf * F0.agx, lgy = sgf * F0.agy, lz = sgf * F0.az;    // lateral unit vector of the side being marched
    BeamAux ax;                            // the current beam's float32 constants; reloaded only when the beam changes
    ax.sec = 1.f;
    ax.reach = ax.meas = ax.tn = 0.f;
#if defined(__CUDA_ARCH__)
    float rmax;                            // range_max from the sentinel entry (see beam_sentinel)
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(rmax) : "r"((uint32_t)__cvta_generic_to_shared(T.aux) - 16u));
#else
    const float rmax = T.aux[-1].sec;
#endif
    int cx = S.cx, cy = S.cy;
    // position inside the triangle in CANONICAL coordinates a >= b: (a, b) = (u, v) in the lower half of the cell
    // (u >= v), (v, u) in the upper half.  Edges: 0: a = 1, 1: b = 0, 2: a = b (the diagonal).
    const bool lower0 = S.u >= S.v;        // the half the curve starts in (both sides); EVERY edge crossing changes it
    float a = lower0 ? S.u : S.v, b = lower0 ? S.v : S.u;
    int prev_be = -1;                      // edge the previous triangle was left through: this one was entered through
                                           // edge 1 after an exit through 0, edge 0 after 1, the diagonal after the diagonal
    double p_cur = 0.0, q_cur = S.q0;
    float p_f = 0.f, q_f = (float)S.q0;    // float32 copies of the two accumulators
    // one word of state: marching in the lower / the upper half of a cell, then kDone or kFail (not certified).
    // (Three bools were kept byte-packed in one register by the compiler and cost half a dozen byte-insert and
    // mask instructions per step.)
    enum : int { kLower = 0, kUpper = 1, kDone = 2, kFail = 3 };
    int state = active && T.B > 0 ? (lower0 ? kLower : kUpper) : kDone;
    int ex_j[2] = {-1, -1}, ex_bb[2] = {0, 0};     // a side whose curve ran off the map: first unresolved beam (count, index) ...
    float ex_p[2] = {0.f, 0.f};                    // ... and the lateral distance the curve was marched to
    const int max_steps = 2 * (8 * (int)(m.range_max / m.res) + 256);
    float z00 = 0.f, z01 = 0.f, z10 = 0.f, z11 = 0.f;
    if (state < kDone) {
        const float *zp = m.z + (cy * m.pitch + cx);       // the cell's vertex (0, 0) in the padded DEM
        z00 = ldg_f(zp);
        z01 = ldg_f(zp + 1);
        z10 = ldg_f(zp + m.pitch);
        z11 = ldg_f(zp + m.pitch + 1);
        ax = beam_aux(tabs, ba);
    }
    // One march step = one triangle.  Because every crossing changes the half, the loop below runs the steps in
    // pairs -- a lower-half instantiation, then an upper-half one -- and `lower` is a compile-time constant in both
    // (gradient, direction and neighbour selects fold away).  A lane whose triangle is of the other half sits a
    // half-trip out; that happens once per side (at its start), after that the lane alternates with the loop.
    auto march_step = [&](auto lower_c) {
        constexpr bool lower = decltype(lower_c)::value;
        {
            bool fail = false;
            const bool fresh = !kBorder || ba != ba_end;     // false only in the step after a side ran off the map: nothing left to march
            const float Du = lower ? z01 - z00 : z11 - z10;      // dh/du, dh/dv of this triangle (per cell)
            const float Dv = lower ? z11 - z01 : z10 - z00;
            const float gamma = F0.cz - F0.cgx * Du - F0.cgy * Dv;
            if (!(gamma < -kGammaMin) && fresh) { AUVPF_REASON(0); fail = true; }
            const float beta = lgx * Du + lgy * Dv - lz;
            const float s = fast_div(beta, gamma);                // profile slope dT/dp
            const float X1 = fmaf(s, F0.cgx, lgx), Y1 = fmaf(s, F0.cgy, lgy);   // d(u,v)/dp
            // the exit: edge 0 is reached when A1 > 0, edge 1 when B1 < 0, edge 2 when B1 > A1; the nearest wins
            // (fractions compared by cross-multiplication)
            const float A1 = lower ? X1 : Y1, B1 = lower ? Y1 : X1;
            float bn = kBigT, bd = 1.f;                            // nearest exit = bn / bd
            int be = -1;
            if (prev_be != 1 && A1 > 0.f) {
                bn = 1.f - a;
                bd = A1;
                be = 0;
            }
            if (prev_be != 0 && B1 < 0.f && b * bd < bn * -B1) {
                bn = b;
                bd = -B1;
                be = 1;
            }
            {
                const float wn = a - b, wd = B1 - A1;
                const bool c2 = prev_be != 2 && wd > 0.f && wn * bd < bn * wd;
                bn = c2 ? wn : bn;
                bd = c2 ? wd : bd;
                be = c2 ? 2 : be;
            }
            // (Tried: picking the exit edge by the signs of (A1, B1) x (vertex - point) -- three cross products, no
            // cross-multiplied comparisons.  Bit-identical on 10 M rays of the host twin, but with the turn-back test
            // it needs it came to 33 instructions against these 29, and its nested selects compiled to branches:
            // 2.69 against 2.33 ms.)
            if (be < 0 && !fail && fresh) { AUVPF_REASON(1); fail = true; }   // curve turns back through its entry edge
            const float dp = fast_div(bn, bd);
            const float p_out_f = p_f + dp;
            // beams whose line meets the profile inside this segment.  gap = lateral distance of the
            // beam line ahead of the curve point at q_cur; it closes at the rate den per unit p.
            // A beam is resolved here when it hits inside the segment, or when the segment's end is
            // beyond what the beam can reach within range_max (miss).
            // gap = q tan - p cancels (both ~ the lateral distance, tens of metres; the difference is a fraction of a
            // cell), but ONE float32 FMA forms the product exactly and rounds the small difference once: what is left
            // is the rounding of its inputs (the float32 copies of the two accumulators and of tan, < 5e-6 m in gap),
            // which the range sees scaled by slope / den -- grazing hits (den < kDenPrecise) go the float64 way.
            // Two instantiations of the loop: on a segment whose slope cannot make any beam of this side graze
            // (s <= s_lim, all of a smooth seabed) the float64 re-solve and its test are not compiled in.
            auto resolve_beams = [&](auto may_graze) {
                for (;;) {                                 // ends at the side's sentinel entry at the latest
                    const float den = fmaf(-s, ax.tn, 1.0f);
                    const float gap = fmaf(q_f, ax.tn, -p_f);
                    const bool hit = gap <= dp * den;
                    if (!(hit || p_out_f > ax.reach)) break;
                    // range = (q at the hit) / cos th.  float32 from the float32 copy of q_cur: three
                    // roundings at <= 120 m, < 2e-5 m (tolerance 1e-4 m).  (gap < 0: the hit lies a rounding before
                    // the segment, on the same plane.  den <= 0 with a hit is a grazing hit: whatever the division
                    // makes of it is replaced.)
                    float t = fmaf(s, fast_div(gap, den), q_f) * ax.sec;
                    if (decltype(may_graze)::value) {
                        if (hit && den < kDenPrecise) {
                            const int bb = beam_index(tabs, ba);
                            t = (float)(precise_hit(D, sgf, cx, cy, z00, z01, z10, z11, lower, T.tan_t[bb]) * T.sec_t[bb]);
                            if (ax.sec == 0.f) t = 0.f;
                        }
                    }
                    // not hit: out of reach, a miss; a hit beyond range_max is one as well
                    // On the no-graze path den >= kDenPrecise > 0 for every beam, so (a) the hit lies on a segment of
                    // the curve, in front of the sensor: no clamp at zero; (b) a beam resolved as a MISS (segment end
                    // beyond its reach, no hit) meets the extended facet at a lateral distance > reach, i.e. its t
                    // is >= range_max by itself: the clamp at range_max is all it needs.
                    float range;
                    // (kBigT * sec: range_max for a miss, but 0 for a beam without a valid measurement, sec = 0)
                    if (decltype(may_graze)::value) range = fminf(hit ? fmaxf(t, 0.f) : kBigT * ax.sec, rmax);
                    else range = fminf(t, rmax);
                    sink(beam_index(tabs, ba), range, ax.meas);
                    ba += bstep;
                    ax = beam_aux(tabs, ba);               // the next beam's constants (or the sentinel's)
                }
            };
            if (!fail) {
                if (s <= s_lim) resolve_beams(std::false_type{});
                else resolve_beams(std::true_type{});
            }
            // (Lanes leave the beam loop after different trip counts.  Round 1 needed an explicit __syncwarp(mask)
            // here -- without it each group of lanes ran the code below on its own, 2.3 executions per step at 13
            // lanes; with the loop in its present form the compiler brackets it with its own reconvergence barrier
            // (BSSY / BSYNC), and the explicit one, with the ballot that made its mask, measured 4.3 % slower.)
            if (fail) {
                state = kFail;
            } else if (ba == ba_end) {
                // this side is done: start the other one from the anchor, or finish
                if (side == 0 && T.B - T.n1 > 0) {
                    side = 1;
                    ba = beam_addr(tabs, T.n1);
                    bstep = 16u;
                    ba_end = beam_addr(tabs, T.B);
                    s_lim = T.s_lim[1];
                    sgf = -sgf;
                    lgx = -lgx;
                    lgy = -lgy;
                    lz = -lz;
                    ax = beam_aux(tabs, ba);
                    cx = S.cx;
                    cy = S.cy;
                    state = lower0 ? kLower : kUpper;
                    a = lower0 ? S.u : S.v;
                    b = lower0 ? S.v : S.u;
                    prev_be = -1;
                    p_cur = 0.0;
                    q_cur = S.q0;
                    p_f = 0.f;
                    q_f = (float)S.q0;
                    const float *zp = m.z + (cy * m.pitch + cx);
                    z00 = ldg_f(zp);
                    z01 = ldg_f(zp + 1);
                    z10 = ldg_f(zp + m.pitch);
                    z11 = ldg_f(zp + m.pitch + 1);
                } else
                    state = kDone;
            } else {
                // advance to the exit point and step into the neighbour across that edge (selects, not
                // branches: lanes leaving through different edges stay converged)
                state = lower ? kUpper : kLower;
                q_cur += (double)(s * dp);
                p_cur += (double)dp;
                q_f = (float)q_cur;
                p_f = (float)p_cur;
                // In canonical coordinates the step across an edge does not depend on the half the curve is in:
                // edge 0 (a = 1): (a, b) -> (b, 0); edge 1 (b = 0): (a, b) -> (1, a); the diagonal: both -> their mean.
                // Which cell that is does: edge 0 leads to x + 1 from the lower half, y + 1 from the upper one,
                // edge 1 to y - 1 / x - 1; the diagonal stays in the cell.  Every crossing changes the half.
                const float a_adv = fminf(fmaxf(fmaf(dp, A1, a), 0.f), 1.f);
                const float b_adv = fminf(fmaxf(fmaf(dp, B1, b), 0.f), 1.f);
                const float mid = 0.5f * (a_adv + b_adv);
                const bool diag = be == 2;
                a = be == 0 ? b_adv : (diag ? mid : 1.f);
                b = be == 0 ? 0.f : (diag ? mid : a_adv);
                if (lower) {                                       // (a compile-time constant)
                    cx += be == 0 ? 1 : 0;
                    cy -= be == 1 ? 1 : 0;
                } else {
                    cy += be == 0 ? 1 : 0;
                    cx -= be == 1 ? 1 : 0;
                }
                if ((unsigned)cx > (unsigned)m.last_cx || (unsigned)cy > (unsigned)m.last_cy) {   // left the map
                    AUVPF_REASON(2);
                    if (kBorder) {
                        // the rest of this side is finished ray by ray AFTER the march loop (finish_side_by_rays);
                        // remember where it stopped
                        ex_bb[side] = beam_index(tabs, ba);
                        ex_j[side] = side == 0 ? T.n1 - 1 - ex_bb[side] : ex_bb[side] - T.n1;
                        ex_p[side] = p_f;
                        ba = ba_end;                               // the side switch happens in the next step
                        ax = beam_aux(tabs, ba);                   // (the sentinel: the beam loop finds nothing)
                        cx = S.cx;                                 // (any cell inside the map: the step's loads stay in bounds)
                        cy = S.cy;
                    } else {
                        state = kFail;
                    }
                } else if (!diag) {
                    const float *zp = m.z + (cy * m.pitch + cx);
                    z00 = ldg_f(zp);
                    z01 = ldg_f(zp + 1);
                    z10 = ldg_f(zp + m.pitch);
                    z11 = ldg_f(zp + m.pitch + 1);
                }
                prev_be = be;
            }
        }
    };
    for (int step = 0; step < max_steps; step += 2) {
        if (!AUVPF_WARP_BALLOT(state < kDone)) break;        // (also the point where the warp reconverges once per pair)
        if (state == kLower) march_step(std::true_type{});
        if (state == kUpper) march_step(std::false_type{});
    }
    if (state < kDone) {                     // ran out of steps: not certified
        AUVPF_REASON(3);
        state = kFail;
    }
    const bool ok = state != kFail;
    if (kBorder && ok && (ex_j[0] >= 0 || ex_j[1] >= 0)) {
        for (int sd = 0; sd < 2; ++sd)
            if (ex_j[sd] >= 0)
                finish_side_by_rays(m, T, D, (float)(sd == 0 ? T.sgn1 : -T.sgn1), ex_j[sd], sd == 0 ? T.n1 : T.B - T.n1,
                                    ex_bb[sd], sd == 0 ? -1 : 1, ex_p[sd], exit_sink);
    }
    return ok;
}

// All beams of one particle by profile merge; MUST be called by all 32 lanes of a warp
// (`active` = this lane has a particle).  Returns false when the particle cannot be certified
// (not inside the map's interior, centre beam missed, a triangle with gamma >= 0): the caller
// must then cast its beams ray by ray.  sink(beam, range) may have been called for some beams
// before a failure; callers discard those.
// `exit_sink` receives the beams of a side that had to be finished ray by ray (finish_side_by_rays, a
// __noinline__ call that takes it by reference): it must NOT share captured state with `sink` -- a variable
// whose address escapes into that call lives in local memory, and the hot loop's accumulator must stay in a
// register (sharing them cost the kernel 15 %).
template <bool kBorder, class Sink, class ExitSink>
AUVPF_HD bool profile_particle(const DemView &m, const FanTables &T, const double p[3], const double R[9], bool active,
                               Sink sink, ExitSink exit_sink)
{
    bool ok = active;
    const double reach = (double)m.range_max;
    const double xmax = m.x0 + (double)(m.W - 1) * (double)m.res, ymax = m.y0 + (double)(m.H - 1) * (double)m.res;
    // the sensor must stand over the map; its REACH may leave it (the march finishes a side whose curve
    // runs off the map ray by ray, finish_side_by_rays), the rays are then clipped to the footprint
    const bool interior = p[0] - reach > m.x0 && p[0] + reach < xmax && p[1] - reach > m.y0 && p[1] + reach < ymax;
    if (kBorder ? !(p[0] > m.x0 && p[0] < xmax && p[1] > m.y0 && p[1] < ymax) : !interior) ok = false;
    // centre beam from the sensor: q0 = T(0)
    float q0f = 0.f;
    if (ok) {
        int ix0, iy0;
        float lx0, ly0;
        origin_cell(m, p, ix0, iy0, lx0, ly0);
        Ray rc = make_ray(m, p, ix0, iy0, lx0, ly0, R, T.c0[0], T.c0[1], T.c0[2]);
        q0f = raycast_from(m, rc, 0.f, kBorder ? interior : true);
    }
    AUVPF_WARP_SYNC();
    bool all_zero = false;
    ProfStart S;
    S.cx = S.cy = 0;
    S.u = S.v = 0.f;
    S.q0 = 0.0;
    ProfFrame F;
    F.agx = F.agy = F.az = F.cgx = F.cgy = F.cz = 0.f;
    ProfFrameD D;
    D.gox = D.goy = D.oz = D.agx = D.agy = D.az = D.cgx = D.cgy = D.cz = 0.0;
    if (ok) do {
        if (!(q0f < m.range_max)) { AUVPF_REASON(6); ok = false; break; }
        if (q0f <= 0.f) {                     // sensor at / under the surface: every range is 0 (spec)
            all_zero = true;
            break;
        }
        double aw[3], cw[3];
        for (int a = 0; a < 3; ++a) {
            aw[a] = R[3 * a] * T.a0[0] + R[3 * a + 1] * T.a0[1] + R[3 * a + 2] * T.a0[2];
            cw[a] = R[3 * a] * T.c0[0] + R[3 * a + 1] * T.c0[1] + R[3 * a + 2] * T.c0[2];
        }
        if (!(aw[0] * aw[0] + aw[1] * aw[1] > 1.0e-6)) { ok = false; break; }   // fan plane (nearly) horizontal
        // The march follows ONE curve -- the connected piece of (fan plane) x (surface) through the centre hit --
        // and that is every beam's first hit only if the intersection is a single-valued profile q = T(p): if
        // gamma < 0 on EVERY facet the fan plane crosses within reach, not just on those the curve meets.  (A
        // sensor pitched 23 deg over 60 deg ridge faces sees crests as overhangs: a beam hit one at 29.35 m that
        // the curve passes at 35 m; the march, stopped early by a short range_max, never met the facet whose
        // gamma would have failed it.)  gamma = c_z - c_xy . grad h <= c_z + |c_xy| max|grad h|: certified from
        // the largest facet gradient in the blocks the reach touches; otherwise the particle goes ray by ray.
        {
            const double cxy = sqrt(cw[0] * cw[0] + cw[1] * cw[1]);
            const double radius = reach * (1.0 + cxy) * m.inv_res_d + 2.0;          // lateral reach + drift of the curve along c
            const float slope = local_max_slope(m, (p[0] - m.x0) * m.inv_res_d, (p[1] - m.y0) * m.inv_res_d, radius);
            if (!(cw[2] + cxy * (double)slope < -(double)kGammaMin)) { AUVPF_REASON(7); ok = false; break; }
        }
        // anchor: the triangle under the centre hit, and the hit re-solved on its plane in float64
        const double gox = (p[0] - m.x0) * m.inv_res_d, goy = (p[1] - m.y0) * m.inv_res_d;   // sensor, grid units
        const double cgx = cw[0] * m.inv_res_d, cgy = cw[1] * m.inv_res_d;
        const double hx = gox + (double)q0f * cgx, hy = goy + (double)q0f * cgy;
        S.cx = clampi((int)floor(hx), 0, m.W - 2);
        S.cy = clampi((int)floor(hy), 0, m.H - 2);
        const bool lower = (hx - (double)S.cx) >= (hy - (double)S.cy);
        const float *zp = m.z + (ptrdiff_t)S.cy * m.pitch + S.cx;
        const double z00 = ldg_f(zp), z01 = ldg_f(zp + 1), z10 = ldg_f(zp + m.pitch), z11 = ldg_f(zp + m.pitch + 1);
        const double Du = lower ? z01 - z00 : z11 - z10, Dv = lower ? z11 - z01 : z10 - z00;
        const double alpha = (z00 - p[2]) + (gox - (double)S.cx) * Du + (goy - (double)S.cy) * Dv;
        const double gamma = cw[2] - cgx * Du - cgy * Dv;
        if (!(gamma < -(double)kGammaMin)) { AUVPF_REASON(4); ok = false; break; }
        S.q0 = alpha / gamma;
        if (!(fabs(S.q0 - (double)q0f) < 1.0e-3)) { AUVPF_REASON(5); ok = false; break; }   // centre hit not on this triangle's plane
        const double ux = gox + S.q0 * cgx - (double)S.cx, vy = goy + S.q0 * cgy - (double)S.cy;
        S.u = (float)(ux < 0.0 ? 0.0 : (ux > 1.0 ? 1.0 : ux));
        S.v = (float)(vy < 0.0 ? 0.0 : (vy > 1.0 ? 1.0 : vy));
        if (lower != (S.u >= S.v)) S.u = S.v = 0.5f * (S.u + S.v);      // keep the half the plane was taken from
        F.cgx = (float)cgx;
        F.cgy = (float)cgy;
        F.cz = (float)cw[2];
        F.agx = (float)(aw[0] * m.inv_res_d);     // lateral unit vector towards +a0; the march applies the side's sign
        F.agy = (float)(aw[1] * m.inv_res_d);
        F.az = (float)aw[2];
        D.gox = gox;
        D.goy = goy;
        D.oz = p[2];
        D.agx = aw[0] * m.inv_res_d;
        D.agy = aw[1] * m.inv_res_d;
        D.az = aw[2];
        D.cgx = cgx;
        D.cgy = cgy;
        D.cz = cw[2];
    } while (0);
    AUVPF_WARP_SYNC();
    if (all_zero)
        for (int b = 0; b < T.B; ++b) sink(b, 0.f, T.aux[b].meas);
    const bool marched = profile_march<kBorder>(m, T, F, &D, S, ok && !all_zero, sink, exit_sink);
    return ok && (all_zero || marched);
}

// ---------------------------------------------------------------------------------------------
// Exact fixed-point normalisation (spec: oracle/pf_oracle.py quantise_logw).
// Only IEEE mul/add with explicit no-FMA intrinsics and integers.
AUVPF_HD double mul_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
AUVPF_HD double add_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

AUVPF_HD float fmul_rn(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
AUVPF_HD float fadd_rn(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}

// The integer weight a = floor(E(lw - m) 2^60) with E a FIXED float32 evaluation of exp: IEEE single
// multiplies and adds only (no FMA contraction), so NumPy float32 arithmetic (the oracle), C and CUDA agree
// bit for bit.  x = lw - m <= 0 (float32 subtract), k = rint(x log2 e) by the 1.5 2^23 trick, r = x - k ln2 in
// two pieces (k LN2_HI is exact), degree-7 Horner, then the float's 24-bit significand is shifted into place
// with integer operations -- no float64, no ldexp / floor calls.  a carries 24 significant bits: the
// float32 log-weight it comes from carries fewer (one ulp of logw ~ -600 moves the weight by 6e-5 relative).
// Round 1 evaluated exp in float64 (26 dependent double operations per particle); the two resample kernels
// were bound by exactly that arithmetic.
AUVPF_HD uint64_t a_from_logw(float lw, float m)
{
    const float x = fadd_rn(lw, -m);
    if (!(x >= -41.5f)) return 0;                 // NaN, -inf and everything below 2^-59.9 of the maximum
    const float t = fmul_rn(x, 1.44269502f);
    const float k = fadd_rn(fadd_rn(t, 12582912.0f), -12582912.0f);
    const float r = fadd_rn(fadd_rn(x, -fmul_rn(k, 0.693359375f)), -fmul_rn(k, -2.12194440e-4f));
    float p = 1.0f / 5040.0f;
    p = fadd_rn(fmul_rn(p, r), 1.0f / 720.0f);
    p = fadd_rn(fmul_rn(p, r), 1.0f / 120.0f);
    p = fadd_rn(fmul_rn(p, r), 1.0f / 24.0f);
    p = fadd_rn(fmul_rn(p, r), 1.0f / 6.0f);
    p = fadd_rn(fmul_rn(p, r), 0.5f);
    p = fadd_rn(fmul_rn(p, r), 1.0f);
    p = fadd_rn(fmul_rn(p, r), 1.0f);
    uint32_t bits;
#if defined(__CUDA_ARCH__)
    bits = __float_as_uint(p);
#else
    memcpy(&bits, &p, sizeof bits);
#endif
    const int shift = (int)k + kABits - 23 + ((int)(bits >> 23) - 127);       // p in [0.70, 1.42]: always a normal float
    const uint64_t mant = (uint64_t)((bits & 0x7FFFFFu) | 0x800000u);
    return shift >= 0 ? mant << shift : (shift > -24 ? mant >> -shift : 0ull);
}

AUVPF_HD uint64_t a_from_linear(double w, double wmax)
{
    if (!(w > 0.0)) return 0;
    return (uint64_t)floor(ldexp(w / wmax, kABits));
}

// ---------------------------------------------------------------------------------------------
// The normalised CDF as a quotient of PREFIX SUMS:
//   C_j = sum_{i <= j} a_i (exact integer, < 2^90),  A = C_{N-1},  cdf_j = floor(C_j 2^52 / A),
//   q_j = cdf_j - cdf_{j-1},  w_j = q_j 2^-52.
// cdf_{N-1} = 2^52 exactly, every cdf_j is an integer <= 2^52 (exact in float64), so np.cumsum(w) is
// cdf 2^-52 bit for bit in any summation order.  Because the CDF is a function of the prefix of the a_i
// alone, nothing per particle has to be stored between the pass that sums the a_i and the pass that
// selects ancestors: the select kernel rebuilds C_j from the log-weights and a table of per-chunk sums.
struct U128 {
    uint64_t lo, hi;
};
AUVPF_HD U128 u128(uint64_t lo, uint64_t hi = 0)
{
    U128 r;
    r.lo = lo;
    r.hi = hi;
    return r;
}
AUVPF_HD U128 operator+(U128 a, U128 b)
{
    U128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
    return r;
}
AUVPF_HD U128 operator+(U128 a, uint64_t b)
{
    U128 r;
    r.lo = a.lo + b;
    r.hi = a.hi + (r.lo < b ? 1ull : 0ull);
    return r;
}
AUVPF_HD bool operator==(U128 a, U128 b) { return a.lo == b.lo && a.hi == b.hi; }
AUVPF_HD unsigned __int128 u128_int(U128 a) { return ((unsigned __int128)a.hi << 64) | a.lo; }
AUVPF_HD U128 u128_from(unsigned __int128 v) { return u128((uint64_t)v, (uint64_t)(v >> 64)); }
// two-limb form used by the warp / block reductions: lo = sum of the low 32 bits, hi = sum of the high bits
AUVPF_HD U128 u128_from_limbs(uint64_t lo32_sum, uint64_t hi_sum)
{
    return u128_from((unsigned __int128)lo32_sum + ((unsigned __int128)hi_sum << 32));
}
AUVPF_HD double u128_to_double(U128 c) { return (double)c.hi * 18446744073709551616.0 + (double)c.lo; }

struct CdfDiv {
    U128 A;
    double scale;        // 2^52 / A, rounded
};
AUVPF_HD CdfDiv make_cdfdiv(U128 A)
{
    CdfDiv d;
    d.A = A;
    const double Ad = u128_to_double(A);
    d.scale = Ad > 0.0 ? 4503599627370496.0 / Ad : 0.0;
    return d;
}
// cdf estimate: within kCdfEstErr units of floor(C 2^52 / A) (roundings of the conversion, of `scale` and of
// the product: < 2^-50.5 relative, i.e. < 3 units at the top of the range)
constexpr double kCdfEstErr = 3.0;
AUVPF_HD double cdf_estimate(U128 C, const CdfDiv &d) { return u128_to_double(C) * d.scale; }
// the exact floor: the estimate is corrected with a multiply-back.  t = C 2^52 and q A both exceed 128
// bits, but their difference r = t - q A lies in (-4 A, 5 A), |r| < 2^93, so it is exact in WRAPPING
// 128-bit arithmetic read as a signed number.  tests/hostsim checks it against Python integers.
AUVPF_HD uint64_t cdf_exact(U128 C, const CdfDiv &d)
{
    const unsigned __int128 A = u128_int(d.A), Ci = u128_int(C);
    if (A == 0) return 0;
    double qe = cdf_estimate(C, d);
    if (!(qe < 4503599627370496.0)) qe = 4503599627370496.0;      // C <= A, so the quotient is <= 2^52
    if (!(qe > 0.0)) qe = 0.0;
    uint64_t q = (uint64_t)qe;
    __int128 r = (__int128)((Ci << kQBits) - (unsigned __int128)q * A);
    while (r < 0) {
        --q;
        r += (__int128)A;
    }
    while (r >= (__int128)A) {
        ++q;
        r -= (__int128)A;
    }
    return q;
}

// Effective sample size for the resampling gate (the block commented out at auv_pf.py:174-181), on exact
// integers so that every block, every GPU and the CPU oracle take the SAME decision: s_i = a_i >> 28
// (a_i <= 2^60, so s_i <= 2^31 and s_i^2 fits 64 bits), S1 = sum s_i, S2 = sum s_i^2 as two limbs (low 32 bits, high bits);
//   neff_fixed = double(S1)^2 / (double(S2_hi) 2^32 + double(S2_lo))     (= (sum w)^2 / sum w^2)
// Sums of integers do not depend on the order or on the sharding.
constexpr int kNeffShift = 29;
AUVPF_HD uint64_t neff_s(uint64_t a) { return a >> kNeffShift; }
AUVPF_HD double neff_fixed_a(uint64_t S1, uint64_t S2_lo, uint64_t S2_hi)
{
    const double S2 = add_rn(mul_rn((double)S2_hi, 4294967296.0), (double)S2_lo);
    const double s1 = (double)S1;
    return S2 > 0.0 ? mul_rn(s1, s1) / S2 : 0.0;
}

// positions[i] = (u + i) / N of resampling.py:157, as the reference rounds it
AUVPF_HD double sys_position(double u, int64_t i, double n)
{
    return add_rn(u, (double)i) / n;
}

// The same quotient for many i and one n without a division per call: with rn = RN(1 / n),
//   q0 = RN(x rn),  r = x - n q0 (exact in one FMA),  q = RN(q0 + r rn)
// is the correctly rounded x / n (Markstein's theorem: holds for round-to-nearest whenever the
// significand of n is not all ones and nothing over/underflows; n is a particle count < 2^31 and
// x = u + i lies in [0, n)), i.e. bit for bit the value the division above returns.  The exact
// tests at N = 1 ... 16 777 216 compare the resulting ancestors with the reference's.
struct SysPos {
    double u, n, rn;
};
AUVPF_HD SysPos make_sys_pos(double u, double n)
{
    SysPos p;
    p.u = u;
    p.n = n;
    p.rn = 1.0 / n;
    return p;
}
AUVPF_HD double sys_position(const SysPos &p, int64_t i)
{
#if defined(__CUDA_ARCH__)
    const double x = __dadd_rn(p.u, (double)i);
    const double q0 = __dmul_rn(x, p.rn);
    const double r = __fma_rn(-p.n, q0, x);
    return __fma_rn(r, p.rn, q0);
#else
    return add_rn(p.u, (double)i) / p.n;
#endif
}

// exact float64 value of an integer CDF entry (<= 2^52)
AUVPF_HD double cdf_value(uint64_t c) { return (double)c * 2.220446049250313e-16; }

// number of outputs i in [0, N) with positions[i] < cdf(c), positions[i] = RN(RN(u + i) / N) as the
// reference rounds them (resampling.py:157).
// Fast path: est = fma(cv, N, -u) differs from the real threshold T_i of "positions[i] < cv <=> i < T_i"
// by the roundings of u + i, of the quotient and of est itself: at most 3 N 2^-53 (< 7.2e-7 for
// N < 2^31).  So when est is further than 1e-5 from every integer, i < T_i <=> i < est for all i and
// the count is ceil(est) -- no position has to be evaluated.  Otherwise (2e-5 of the cases) the count
// is fixed up with the reference's own expression, which is monotone in i.
AUVPF_HD int64_t sys_count_below(const SysPos &sp, int64_t N, uint64_t c)
{
    const double cv = cdf_value(c);
    const double est = fma(cv, sp.n, -sp.u);
    if (fabs(est - rint(est)) > 1.0e-5) {
        const int64_t i = (int64_t)ceil(est);
        return i < 0 ? 0 : (i > N ? N : i);
    }
    int64_t i = (int64_t)floor(est);
    if (i < 0) i = 0;
    if (i > N) i = N;
    while (i > 0 && !(sys_position(sp, i - 1) < cv)) --i;
    while (i < N && sys_position(sp, i) < cv) ++i;
    return i;
}

AUVPF_HD int64_t sys_count_below(double u, int64_t N, uint64_t c)
{
    return sys_count_below(make_sys_pos(u, (double)N), N, c);
}

// Block-relative end of a particle's run of offspring in k_select_gather: with c the particle's inclusive
// CDF value and a block that owns the outputs [g0, g0 + nblk),
//   rel_end = clamp(sys_count_below(c) - g0, 0, nblk).
// Only values inside (0, nblk) need the count itself; x = estimate - g0 decides the rest, with the same
// 1e-5 margin as sys_count_below, which is called when x is that close to an integer.
struct RelEndCtx {
    double n_scaled;     // N 2^-52 (exact)
    double u;
    double g0d, nblk_d;  // (double)g0, (double)nblk
    int64_t g0;
    int nblk;
};
AUVPF_HD RelEndCtx make_rel_end(double u, double n, int64_t g0, int nblk)
{
    RelEndCtx r;
    r.n_scaled = n * 2.220446049250313e-16;
    r.u = u;
    r.g0d = (double)g0;
    r.nblk_d = (double)nblk;
    r.g0 = g0;
    r.nblk = nblk;
    return r;
}
AUVPF_HD int rel_end_of(const SysPos &sp, int64_t n_global, const RelEndCtx &r, unsigned long long c)
{
    // (double)c for c <= 2^52 without a 64-bit conversion: 2^52 + c is exact in the mantissa
#if defined(__CUDA_ARCH__)
    const double cd = __longlong_as_double((long long)(0x4330000000000000ull + c)) - 4503599627370496.0;
#else
    const unsigned long long bits = 0x4330000000000000ull + c;
    double cd0;
    memcpy(&cd0, &bits, sizeof(cd0));
    const double cd = cd0 - 4503599627370496.0;
#endif
    const double x = fma(cd, r.n_scaled, -r.u) - r.g0d;
    if (x < -1.0e-5) return 0;
    if (x > r.nblk_d + 1.0e-5) return r.nblk;
#if defined(__CUDA_ARCH__)
    const int ie = __double2int_ru(x);
#else
    const int ie = (int)ceil(x);
#endif
    const double d = (double)ie - x;                       // in [0, 1)
#if defined(__CUDA_ARCH__)
    if (d > 1.0e-5 && d < 1.0 - 1.0e-5) return min(max(ie, 0), r.nblk);
#else
    if (d > 1.0e-5 && d < 1.0 - 1.0e-5) return ie < 0 ? 0 : (ie > r.nblk ? r.nblk : ie);
#endif
    const int64_t g = sys_count_below(sp, n_global, c) - r.g0;
    return (int)(g < 0 ? 0 : (g > r.nblk ? r.nblk : g));
}

// The same from an ESTIMATE of the cdf value (cdf_estimate plus float64 running sums: within a few units of
// the exact integer, i.e. within ~3e-6 of the true x for N <= 2^30): decided whenever x is further than 1e-5
// from an integer; -1 = too close to call, the caller evaluates the exact cdf and takes rel_end_of.
AUVPF_HD int rel_end_fast(const RelEndCtx &r, double est)
{
    const double x = fma(est, r.n_scaled, -r.u) - r.g0d;
    if (x < -1.0e-5) return 0;
    if (x > r.nblk_d + 1.0e-5) return r.nblk;
#if defined(__CUDA_ARCH__)
    const int ie = __double2int_ru(x);
#else
    const int ie = (int)ceil(x);
#endif
    const double d = (double)ie - x;                       // in [0, 1)
    if (d > 1.0e-5 && d < 1.0 - 1.0e-5) return ie < 0 ? 0 : (ie > r.nblk ? r.nblk : ie);
    return -1;
}


// ---------------------------------------------------------------------------------------------
// residual_resample (resampling.py:27-76) on fixed-point weights w = q 2^-52.
// Deterministic copies: floor(N * w) with the reference's float64 product (:61).
AUVPF_HD uint64_t residual_copies(uint64_t q, int64_t N)
{
    return (uint64_t)floor(mul_rn((double)N, cdf_value(q)));
}

// Fractional part N w - k (what :70 means to compute) on a 2^-38 grid: (N q - k 2^52) >> 14.
// k can exceed the true floor by one when the float64 product rounds up to an integer: 0 then.
constexpr int kResShift = 14;
AUVPF_HD uint64_t residual_frac(uint64_t q, uint64_t k, int64_t N)
{
    const unsigned __int128 nq = (unsigned __int128)q * (uint64_t)N;
    const unsigned __int128 kk = (unsigned __int128)k << kQBits;
    if (nq <= kk) return 0;
    return (uint64_t)((nq - kk) >> kResShift);
}

// A uniform u in [0,1) on the integer residual CDF of total R: floor(floor(u 2^53) R / 2^53)
AUVPF_HD uint64_t residual_target(double u, uint64_t R)
{
    const uint64_t m = (uint64_t)floor(ldexp(u, 53));
    return (uint64_t)(((unsigned __int128)m * R) >> 53);
}

AUVPF_HD double div_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b;
    return r;
#endif
}

// np.searchsorted(arr, keys) (side='left') as NumPy's binsearch runs it, including the
// key-to-key bounds hint that makes the result on a NON-monotone array depend on the key order
// (numpy/_core/src/npysort/binsearch.cpp).  The reference calls it on such an array
// (resampling.py:70-74), so strict parity needs the same walk.  State = (lo, hi, last key).
struct SearchHint {
    int64_t lo, hi;
    double last;
};
AUVPF_HD int64_t searchsorted_left_hinted(const double *arr, int64_t n, double key, SearchHint &h)
{
    if (h.last < key) {
        h.hi = n;
    } else {
        h.lo = 0;
        h.hi = h.hi < n ? h.hi + 1 : n;
    }
    h.last = key;
    while (h.lo < h.hi) {
        const int64_t mid = h.lo + ((h.hi - h.lo) >> 1);
        if (arr[mid] < key) h.lo = mid + 1;
        else h.hi = mid;
    }
    return h.lo;
}

}   // namespace auvpf
