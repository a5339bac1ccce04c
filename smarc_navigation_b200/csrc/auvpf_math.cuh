// auvpf_math.cuh -- per-element math of the particle-filter cycle.
//
// Everything here is __host__ __device__ so that the SAME source is compiled
// into the CUDA kernels (auvpf_kernels.cuh) and into the test-only host
// simulator (tests/hostsim), which lets the CPU test-suite check the kernel
// arithmetic against the oracle on a box without a GPU.  The product library
// only ever runs the __device__ instantiations.
#pragma once
#include <cmath>
#include <cstdint>

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

// Box-Muller on two 32-bit words: u = (x + 0.5) 2^-32.
AUVPF_HD void box_muller(uint32_t a, uint32_t b, float &n0, float &n1)
{
    float u0 = ((float)a + 0.5f) * 2.3283064365386963e-10f;
    float u1 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
    float r = sqrtf(-2.0f * logf(u0));
    float s, c;
#if defined(__CUDA_ARCH__)
    sincospif(2.0f * u1, &s, &c);
#else
    s = sinf(2.0f * kPiF * u1);
    c = cosf(2.0f * kPiF * u1);
#endif
    n0 = r * c;
    n1 = r * s;
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
    sincos_f(yw, s, c);
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
struct DemView {
    const float *z;      // H x W vertex heights
    const float *mip;    // coarse max-mip level: mipH x mipW block maxima (closed vertex footprint)
    int W, H;
    int mipW, mipH;
    int shift;           // block = 2^shift cells
    float inv_res;
    float range_max;
    float zmax;          // global maximum height
    double x0, y0, inv_res_d;
};

struct Ray {
    int ix0, iy0;        // origin cell (absolute column / row of the local frame's origin)
    float lx0, ly0;      // origin in local grid units, in [0,1) when inside the map
    float oz;
    float dgx, dgy, dz;  // direction in grid units / metres
};

AUVPF_HD float dem_at(const DemView &m, int row, int col)
{
#if defined(__CUDA_ARCH__)
    return __ldg(m.z + (size_t)row * m.W + col);
#else
    return m.z[(size_t)row * m.W + col];
#endif
}

AUVPF_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// General surface evaluation at local grid point (px, py).
AUVPF_HD float surf_local(const DemView &m, const Ray &r, float px, float py)
{
    int cx = clampi(r.ix0 + (int)floorf(px), 0, m.W - 2);
    int cy = clampi(r.iy0 + (int)floorf(py), 0, m.H - 2);
    float u = px - (float)(cx - r.ix0), v = py - (float)(cy - r.iy0);
    float z00 = dem_at(m, cy, cx), z01 = dem_at(m, cy, cx + 1);
    float z10 = dem_at(m, cy + 1, cx), z11 = dem_at(m, cy + 1, cx + 1);
    return (u >= v) ? z00 + u * (z01 - z00) + v * (z11 - z01) : z00 + u * (z11 - z10) + v * (z10 - z00);
}

// Crossing walker over the three line families x = k, y = k, x - y = k.
struct Walker {
    float tx, ty, td;    // next crossing time of each family
    float kx, ky, kd;    // its line index (local grid units), kept as float
    float sx, sy, sd;    // +-1 step (0 = family never crossed)
    float ix, iy, id;    // 1 / derivative of the family coordinate
    float c0d;           // lx0 - ly0
};

constexpr float kBigT = 3.0e38f;

AUVPF_HD void walker_setup(Walker &w, const Ray &r)
{
    float dd = r.dgx - r.dgy;
    w.sx = r.dgx > 0.f ? 1.f : (r.dgx < 0.f ? -1.f : 0.f);
    w.sy = r.dgy > 0.f ? 1.f : (r.dgy < 0.f ? -1.f : 0.f);
    w.sd = dd > 0.f ? 1.f : (dd < 0.f ? -1.f : 0.f);
    w.ix = w.sx != 0.f ? 1.0f / r.dgx : 0.f;
    w.iy = w.sy != 0.f ? 1.0f / r.dgy : 0.f;
    w.id = w.sd != 0.f ? 1.0f / dd : 0.f;
    w.c0d = r.lx0 - r.ly0;
}

AUVPF_HD void family_init(float coord, float c0, float s, float inv, float t_from, float &k, float &t)
{
    if (s == 0.f) {
        k = 0.f;
        t = kBigT;
        return;
    }
    k = s > 0.f ? floorf(coord) + 1.f : ceilf(coord) - 1.f;
    t = (k - c0) * inv;
    // a crossing that rounds to <= t_from is the one we are standing on: move on
    for (int g = 0; g < 3 && t <= t_from; ++g) {
        k += s;
        t = (k - c0) * inv;
    }
}

// (Re)aim the walker at the first crossings strictly after t_from.
AUVPF_HD void walker_seek(Walker &w, const Ray &r, float t_from)
{
    float px = r.lx0 + t_from * r.dgx, py = r.ly0 + t_from * r.dgy;
    family_init(px, r.lx0, w.sx, w.ix, t_from, w.kx, w.tx);
    family_init(py, r.ly0, w.sy, w.iy, t_from, w.ky, w.ty);
    family_init(px - py, w.c0d, w.sd, w.id, t_from, w.kd, w.td);
}

// f = ray height - surface height at the next crossing (<= t_stop); advances the walker.
// Returns false when the next crossing lies beyond t_stop.
AUVPF_HD bool walker_next(Walker &w, const DemView &m, const Ray &r, float t_stop, float &t_out, float &f_out)
{
    float t = fminf(w.tx, fminf(w.ty, w.td));
    if (!(t <= t_stop)) return false;
    float px = r.lx0 + t * r.dgx, py = r.ly0 + t * r.dgy;
    int row, col, stride;
    float frac;
    if (t == w.tx) {                    // vertical grid line x = kx: edge (iy,kx)-(iy+1,kx)
        int k = (int)w.kx;
        float fy = floorf(py);
        col = clampi(r.ix0 + k, 0, m.W - 1);
        row = clampi(r.iy0 + (int)fy, 0, m.H - 2);
        frac = py - (float)(row - r.iy0);
        stride = m.W;
        w.kx += w.sx;
        w.tx = (w.kx - r.lx0) * w.ix;
    } else if (t == w.ty) {             // horizontal grid line y = ky: edge (ky,ix)-(ky,ix+1)
        int k = (int)w.ky;
        float fx = floorf(px);
        row = clampi(r.iy0 + k, 0, m.H - 1);
        col = clampi(r.ix0 + (int)fx, 0, m.W - 2);
        frac = px - (float)(col - r.ix0);
        stride = 1;
        w.ky += w.sy;
        w.ty = (w.ky - r.ly0) * w.iy;
    } else {                            // diagonal x - y = kd: edge (iy,ix)-(iy+1,ix+1)
        int k = (int)w.kd;
        float fx = floorf(px);
        col = clampi(r.ix0 + (int)fx, 0, m.W - 2);
        row = clampi(r.iy0 + ((int)fx - k), 0, m.H - 2);
        frac = px - (float)(col - r.ix0);
        stride = m.W + 1;
        w.kd += w.sd;
        w.td = (w.kd - w.c0d) * w.id;
    }
    const float *p = m.z + (size_t)row * m.W + col;
#if defined(__CUDA_ARCH__)
    float za = __ldg(p), zb = __ldg(p + stride);
#else
    float za = p[0], zb = p[stride];
#endif
    float h = za + frac * (zb - za);
    t_out = t;
    f_out = (r.oz + t * r.dz) - h;
    return true;
}

AUVPF_HD float mip_at(const DemView &m, int by, int bx)
{
#if defined(__CUDA_ARCH__)
    return __ldg(m.mip + (size_t)by * m.mipW + bx);
#else
    return m.mip[(size_t)by * m.mipW + bx];
#endif
}

constexpr float kSkipMargin = 2.0e-3f;   // [m] slack that keeps max-mip skipping conservative in fp32

// First-hit range of one ray; misses return range_max.
AUVPF_HD float raycast_one(const DemView &m, const Ray &r)
{
    float ts = 0.f, te = m.range_max;
    // clip to the map footprint (local coordinates)
    {
        float lo = -(float)r.ix0, hi = (float)(m.W - 1 - r.ix0);
        if (r.dgx != 0.f) {
            float inv = 1.0f / r.dgx;
            float t1 = (lo - r.lx0) * inv, t2 = (hi - r.lx0) * inv;
            ts = fmaxf(ts, fminf(t1, t2));
            te = fminf(te, fmaxf(t1, t2));
        } else if (r.lx0 < lo || r.lx0 > hi)
            return m.range_max;
        lo = -(float)r.iy0;
        hi = (float)(m.H - 1 - r.iy0);
        if (r.dgy != 0.f) {
            float inv = 1.0f / r.dgy;
            float t1 = (lo - r.ly0) * inv, t2 = (hi - r.ly0) * inv;
            ts = fmaxf(ts, fminf(t1, t2));
            te = fminf(te, fmaxf(t1, t2));
        } else if (r.ly0 < lo || r.ly0 > hi)
            return m.range_max;
    }
    if (!(ts <= te)) return m.range_max;

    // Sample at the clip start: spec says f(ts) <= 0 -> range = ts.
    float t_prev = ts;
    float f_prev = (r.oz + ts * r.dz) - surf_local(m, r, r.lx0 + ts * r.dgx, r.ly0 + ts * r.dgy);
    if (f_prev <= 0.f) return ts;

    // Nothing can be hit above the global maximum height.
    float t_cur = ts;
    if (r.oz + ts * r.dz > m.zmax + kSkipMargin) {
        if (r.dz >= 0.f) return m.range_max;
        float tz = (m.zmax + kSkipMargin - r.oz) / r.dz;
        if (tz >= te) return m.range_max;
        if (tz > t_cur) t_cur = tz;
    }
    bool synced = (t_cur == ts);      // (t_prev, f_prev) is a sample at t_cur?

    Walker w;
    walker_setup(w, r);

    // coarse DDA over max-mip blocks (absolute block indices)
    const int S = 1 << m.shift;
    float px = r.lx0 + t_cur * r.dgx, py = r.ly0 + t_cur * r.dgy;
    int bx = clampi((r.ix0 + (int)floorf(px)) >> m.shift, 0, m.mipW - 1);
    int by = clampi((r.iy0 + (int)floorf(py)) >> m.shift, 0, m.mipH - 1);
    const int sbx = r.dgx > 0.f ? 1 : -1, sby = r.dgy > 0.f ? 1 : -1;
    // local coordinate of the next coarse line ahead of the block, per axis
    float lxn = (float)(((bx + (sbx > 0 ? 1 : 0)) << m.shift) - r.ix0);
    float lyn = (float)(((by + (sby > 0 ? 1 : 0)) << m.shift) - r.iy0);
    const float fS = (float)S;
    float tcx = w.sx != 0.f ? (lxn - r.lx0) * w.ix : kBigT;
    float tcy = w.sy != 0.f ? (lyn - r.ly0) * w.iy : kBigT;
    // standing on (or a hair past) a coarse line: we are already in the next block
    for (int g = 0; g < 2 && tcx <= t_cur; ++g) { bx += sbx; lxn += sbx * fS; tcx = (lxn - r.lx0) * w.ix; }
    for (int g = 0; g < 2 && tcy <= t_cur; ++g) { by += sby; lyn += sby * fS; tcy = (lyn - r.ly0) * w.iy; }
    bool walking = false;             // walker aimed at t_prev?

    while (t_cur < te) {
        float t_next = fminf(fminf(tcx, tcy), te);
        int cbx = clampi(bx, 0, m.mipW - 1), cby = clampi(by, 0, m.mipH - 1);
        float zlow = r.oz + (r.dz < 0.f ? t_next : t_cur) * r.dz;
        if (zlow > mip_at(m, cby, cbx) + kSkipMargin) {
            synced = false;           // skipped: no sample carried across
            walking = false;
        } else {
            if (!synced) {
                t_prev = t_cur;
                f_prev = (r.oz + t_cur * r.dz) - surf_local(m, r, r.lx0 + t_cur * r.dgx, r.ly0 + t_cur * r.dgy);
                if (f_prev <= 0.f) return t_cur;
                synced = true;
                walking = false;
            }
            if (!walking) {
                walker_seek(w, r, t_prev);
                walking = true;
            }
            float t, f;
            while (walker_next(w, m, r, t_next, t, f)) {
                if (f <= 0.f) return t_prev + (t - t_prev) * (f_prev / (f_prev - f));
                t_prev = t;
                f_prev = f;
            }
        }
        if (t_next >= te) {
            if (synced) {             // close the last partial segment at te
                float f = (r.oz + te * r.dz) - surf_local(m, r, r.lx0 + te * r.dgx, r.ly0 + te * r.dgy);
                if (f <= 0.f && te > t_prev) return t_prev + (te - t_prev) * (f_prev / (f_prev - f));
            }
            break;
        }
        // step the coarse DDA
        t_cur = t_next;
        if (tcx <= tcy) { bx += sbx; lxn += sbx * fS; tcx = (lxn - r.lx0) * w.ix; }
        else { by += sby; lyn += sby * fS; tcy = (lyn - r.ly0) * w.iy; }
    }
    return m.range_max;
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
// Exact fixed-point normalisation (spec: oracle/pf_oracle.py quantise_logw).
// Only IEEE mul/add with explicit no-FMA intrinsics, rint, floor and integers.
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

AUVPF_HD uint64_t a_from_x(double x)   // x = logw - max <= 0 (or NaN / -inf)
{
    if (!(x >= -745.0)) return 0;
    const double LOG2E = 0x1.71547652b82fep+0, LN2_HI = 0x1.62e42fee00000p-1, LN2_LO = 0x1.a39ef35793c76p-33;
    const double C[14] = {1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
                          1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0};
    double k = rint(mul_rn(x, LOG2E));
    double r = add_rn(add_rn(x, -mul_rn(k, LN2_HI)), -mul_rn(k, LN2_LO));
    double p = C[13];
#pragma unroll
    for (int n = 12; n >= 0; --n) p = add_rn(mul_rn(p, r), C[n]);
    int s = (int)k + kABits;
    if (s < -1) return 0;
    return (uint64_t)floor(ldexp(p, s));
}

AUVPF_HD uint64_t a_from_logw(float lw, float m) { return a_from_x((double)lw - (double)m); }

AUVPF_HD uint64_t a_from_linear(double w, double wmax)
{
    if (!(w > 0.0)) return 0;
    return (uint64_t)floor(ldexp(w / wmax, kABits));
}

// q = floor(a * 2^52 / A), A given as two 64-bit halves
AUVPF_HD uint64_t q_from_a(uint64_t a, uint64_t A_lo, uint64_t A_hi)
{
    unsigned __int128 A = ((unsigned __int128)A_hi << 64) | A_lo;
    unsigned __int128 t = (unsigned __int128)a << kQBits;
    return (uint64_t)(t / A);
}

// positions[i] = (u + i) / N of resampling.py:157, as the reference rounds it
AUVPF_HD double sys_position(double u, int64_t i, double n)
{
    return add_rn(u, (double)i) / n;
}

// exact float64 value of an integer CDF entry (<= 2^52)
AUVPF_HD double cdf_value(uint64_t c) { return (double)c * 2.220446049250313e-16; }

// number of outputs i in [0, N) with positions[i] < cdf(c): the first i with positions[i] >= c 2^-52
AUVPF_HD int64_t sys_count_below(double u, int64_t N, uint64_t c)
{
    double n = (double)N;
    double cv = cdf_value(c);
    double est = cv * n - u;
    int64_t i = (int64_t)floor(est);
    if (i < 0) i = 0;
    if (i > N) i = N;
    // fix up with the reference's own expression (monotone in i)
    while (i > 0 && !(sys_position(u, i - 1, n) < cv)) --i;
    while (i < N && sys_position(u, i, n) < cv) ++i;
    return i;
}

}   // namespace auvpf
