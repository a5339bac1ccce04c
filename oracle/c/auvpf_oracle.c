/* auvpf_oracle.c -- CPU (float64) oracle for the MBES measurement cycle.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): linked/executed only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  The product library (smarc_navigation_b200/csrc) never links this.
 *
 * PARITY UNPINNED for the raycast and the beam-range likelihood: the reference
 * checkout contains no raycaster, mesh or MBES likelihood.
 * What the reference does fix, and what this file follows:
 *   - particle pose -> map-frame (p, R):  Particle.get_p_pose,
 *       auv_particle_filter/scripts/auv_particle.py:72-83
 *       (m2o @ (T(xyz) @ Rq(roll,pitch,yaw)), R = Rz Ry Rx, tf 'sxyz');
 *   - independent Gaussian with meas_cov = std^2 I:  auv_particle.py:28,105-106;
 *   - dead reckoning:  Particle.motion_pred, auv_particle.py:38-70;
 *   - systematic resampling:  resampling.py:135-168 (sequential cumsum +
 *       two-pointer walk, strict '<').
 * Everything else about the raycast is the build's own specification:
 *
 * RAYCAST SPEC
 *   DEM: H x W float32 vertex heights, row i at y = y0 + i*res, column j at
 *   x = x0 + j*res.  Each cell is split by the diagonal (i,j)->(i+1,j+1) into
 *   the triangle u >= v {(i,j),(i,j+1),(i+1,j+1)} and the triangle u < v
 *   {(i,j),(i+1,j+1),(i+1,j)} (u, v = fractional x, y inside the cell).  The
 *   surface is the union of those triangles (a continuous piecewise-linear
 *   height field over [x0, x0+(W-1)res] x [y0, y0+(H-1)res]); there are no
 *   side walls.  For a ray o + t d, |d| = 1, the ray is clipped to the map
 *   footprint and to t in [0, range_max] giving [ts, te].  With
 *   f(t) = z(t) - surface(x(t), y(t)):  if f(ts) <= 0 the range is ts;
 *   otherwise the range is the smallest root of f in (ts, te]; if there is
 *   none (or the clip interval is empty) the beam MISSES and the range is
 *   range_max.  f is linear between consecutive crossings of the three line
 *   families x = k, y = k, x - y = k (grid units), so the root is found exactly
 *   by linear interpolation inside the first sign-changing interval.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    const float *z;
    int H, W;
    double x0, y0, res;
} orc_dem;

static double surf(const orc_dem *m, double gx, double gy)
{
    int j = (int)floor(gx), i = (int)floor(gy);
    if (j < 0) j = 0;
    if (j > m->W - 2) j = m->W - 2;
    if (i < 0) i = 0;
    if (i > m->H - 2) i = m->H - 2;
    double u = gx - j, v = gy - i;
    const float *r0 = m->z + (size_t)i * m->W + j, *r1 = r0 + m->W;
    double z00 = r0[0], z01 = r0[1], z10 = r1[0], z11 = r1[1];
    if (u >= v) return z00 + u * (z01 - z00) + v * (z11 - z01);
    return z00 + u * (z11 - z10) + v * (z10 - z00);
}

/* One ray.  diag[0] = smallest local minimum of f before the hit that stayed
 * positive ("near miss"; +inf if none), diag[1] = depth of the dip below the
 * surface right after the hit (+inf if the ray ends below the surface),
 * diag[2] = te - range (distance of the hit from the clip end). */
static double raycast_one(const orc_dem *m, const double o[3], const double d[3],
                          double range_max, double diag[3])
{
    const double BIG = 1e300;
    double near_miss = INFINITY, dip = INFINITY;
    if (diag) { diag[0] = INFINITY; diag[1] = INFINITY; diag[2] = INFINITY; }
    double gx0 = (o[0] - m->x0) / m->res, gy0 = (o[1] - m->y0) / m->res;
    double dgx = d[0] / m->res, dgy = d[1] / m->res;
    double ts = 0.0, te = range_max;
    /* slab clip to [0, W-1] x [0, H-1] */
    double lo[2] = {0.0, 0.0}, hi[2] = {(double)(m->W - 1), (double)(m->H - 1)};
    double g0[2] = {gx0, gy0}, dg[2] = {dgx, dgy};
    for (int a = 0; a < 2; ++a) {
        if (dg[a] == 0.0) {
            if (g0[a] < lo[a] || g0[a] > hi[a]) return range_max;
        } else {
            double t1 = (lo[a] - g0[a]) / dg[a], t2 = (hi[a] - g0[a]) / dg[a];
            if (t1 > t2) { double s = t1; t1 = t2; t2 = s; }
            if (t1 > ts) ts = t1;
            if (t2 < te) te = t2;
        }
    }
    if (!(ts <= te)) return range_max;

    double t_prev = ts;
    double f_prev = (o[2] + ts * d[2]) - surf(m, gx0 + ts * dgx, gy0 + ts * dgy);
    if (f_prev <= 0.0) { if (diag) diag[2] = te - ts; return ts; }

    /* next crossing index of each family, strictly after ts */
    double px = gx0 + ts * dgx, py = gy0 + ts * dgy;
    double c0[3] = {gx0, gy0, gx0 - gy0};          /* line coordinate at t = 0   */
    double cd[3] = {dgx, dgy, dgx - dgy};          /* its derivative             */
    double cp[3] = {px, py, px - py};              /* line coordinate at ts      */
    double kk[3], step[3], tn[3];
    for (int a = 0; a < 3; ++a) {
        if (cd[a] > 0.0) { kk[a] = floor(cp[a]) + 1.0; step[a] = 1.0; }
        else if (cd[a] < 0.0) { kk[a] = ceil(cp[a]) - 1.0; step[a] = -1.0; }
        else { kk[a] = 0.0; step[a] = 0.0; }
        tn[a] = (step[a] != 0.0) ? (kk[a] - c0[a]) / cd[a] : BIG;
        /* guard against a crossing that rounds to <= ts */
        while (step[a] != 0.0 && tn[a] <= ts) { kk[a] += step[a]; tn[a] = (kk[a] - c0[a]) / cd[a]; }
    }

    double range = range_max;
    int hit = 0, done = 0;
    double f_prev2 = INFINITY;  /* f two points back, for local-minimum detection */
    while (!done) {
        int a = 0;
        if (tn[1] < tn[a]) a = 1;
        if (tn[2] < tn[a]) a = 2;
        double t = tn[a];
        if (t >= te) { t = te; done = 1; }
        else { kk[a] += step[a]; tn[a] = (kk[a] - c0[a]) / cd[a]; }
        double f = (o[2] + t * d[2]) - surf(m, gx0 + t * dgx, gy0 + t * dgy);
        if (!hit) {
            if (f <= 0.0) {
                range = t_prev + (t - t_prev) * f_prev / (f_prev - f);
                hit = 1;
                dip = -f;
                if (!diag) break;
                if (done) dip = INFINITY;
            } else {
                if (f > f_prev && f_prev < f_prev2 && f_prev < near_miss) near_miss = f_prev;
            }
        } else {
            if (f > 0.0) break;          /* came back above the surface */
            if (-f > dip) dip = -f;
            if (done) dip = INFINITY;    /* ended below the surface: robust */
        }
        f_prev2 = f_prev;
        t_prev = t;
        f_prev = f;
    }
    if (!hit) {
        /* a miss is fragile when the ray came close to the surface anywhere */
        if (f_prev < near_miss && f_prev < f_prev2) near_miss = f_prev;
    }
    if (diag) { diag[0] = near_miss; diag[1] = hit ? dip : INFINITY; diag[2] = hit ? te - range : INFINITY; }
    return range;
}

/* ------------------------------------------------------------------ exported */
void orc_raycast(const float *z, int H, int W, double x0, double y0, double res,
                 const double *origins, const double *dirs, long n, double range_max,
                 double *ranges, double *diag /* n*3 or NULL */)
{
    orc_dem m = {z, H, W, x0, y0, res};
#pragma omp parallel for schedule(dynamic, 256)
    for (long r = 0; r < n; ++r)
        ranges[r] = raycast_one(&m, origins + 3 * r, dirs + 3 * r, range_max, diag ? diag + 3 * r : NULL);
}

static void rpy_matrix(double roll, double pitch, double yaw, double R[9])
{
    double cy = cos(yaw), sy = sin(yaw), cp = cos(pitch), sp = sin(pitch), cr = cos(roll), sr = sin(roll);
    R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

/* get_p_pose (auv_particle.py:72-83): p = m2o_R xyz + m2o_t, R = m2o_R Rq(rpy) */
static void particle_frame(const double pose[6], const double m2o[16], double p[3], double R[9])
{
    double Rl[9];
    rpy_matrix(pose[3], pose[4], pose[5], Rl);
    for (int a = 0; a < 3; ++a) {
        p[a] = m2o[4 * a + 0] * pose[0] + m2o[4 * a + 1] * pose[1] + m2o[4 * a + 2] * pose[2] + m2o[4 * a + 3];
        for (int b = 0; b < 3; ++b)
            R[3 * a + b] = m2o[4 * a + 0] * Rl[b] + m2o[4 * a + 1] * Rl[3 + b] + m2o[4 * a + 2] * Rl[6 + b];
    }
}

/* Expected ranges for every (particle, beam): ranges_out is N*B (may be NULL),
 * logw_out is N (may be NULL).  logw follows pf_oracle.beam_logw. */
void orc_update_mbes(const float *z, int H, int W, double x0, double y0, double res,
                     const double *poses, long N, const double *m2o,
                     const double *beam_dirs, const double *r_meas, int B,
                     double meas_std, double range_min, double range_max,
                     double *ranges_out, double *diag_out, double *logw_out)
{
    orc_dem m = {z, H, W, x0, y0, res};
    int nvalid = 0;
    for (int b = 0; b < B; ++b)
        if (r_meas && isfinite(r_meas[b]) && r_meas[b] >= range_min && r_meas[b] <= range_max) ++nvalid;
    double lognorm = -(double)nvalid * log(meas_std * sqrt(2.0 * M_PI));
#pragma omp parallel for schedule(dynamic, 4)
    for (long i = 0; i < N; ++i) {
        double p[3], R[9];
        particle_frame(poses + 6 * i, m2o, p, R);
        double acc = 0.0;
        for (int b = 0; b < B; ++b) {
            const double *db = beam_dirs + 3 * b;
            double d[3];
            for (int a = 0; a < 3; ++a) d[a] = R[3 * a] * db[0] + R[3 * a + 1] * db[1] + R[3 * a + 2] * db[2];
            double r = raycast_one(&m, p, d, range_max, diag_out ? diag_out + 3 * ((size_t)i * B + b) : NULL);
            if (ranges_out) ranges_out[(size_t)i * B + b] = r;
            if (r_meas && isfinite(r_meas[b]) && r_meas[b] >= range_min && r_meas[b] <= range_max) {
                double e = r_meas[b] - r;
                acc += e * e;
            }
        }
        if (logw_out) logw_out[i] = -0.5 * acc / (meas_std * meas_std) + lognorm;
    }
}

/* Particle.motion_pred (auv_particle.py:38-70), all particles; roll/pitch are the
 * loop-invariant euler_from_quaternion(odom) values (:50-56) computed by the caller.
 * noise: N*6 standard normals (row = np.random.randn(1,6) of :40). */
void orc_motion_pred(double *poses, long N, const double v[3], double wz, double roll, double pitch,
                     double z, double dt, const double cov[6], const double *noise)
{
    double s[6];
    for (int a = 0; a < 6; ++a) s[a] = sqrt(cov[a]);
    const double two_pi = 2.0 * M_PI;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) {
        double *p = poses + 6 * i;
        const double *n = noise + 6 * i;
        double yaw = p[5] + wz * dt + s[5] * n[5];                 /* :47 */
        yaw = fmod(yaw + M_PI, two_pi);                            /* :48, Python % */
        if (yaw < 0.0) yaw += two_pi;
        yaw -= M_PI;
        p[3] = roll; p[4] = pitch; p[5] = yaw;                     /* :57 */
        double cy = cos(yaw), sy = sin(yaw), cp = cos(pitch), sp = sin(pitch), cr = cos(roll), sr = sin(roll);
        /* rows 0,1 of fullRotation (:86-97); its rot_y typo only touches row 2 */
        double R0[3] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr};
        double R1[3] = {sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr};
        double vx = v[0] * dt, vy = v[1] * dt, vz = v[2] * dt;
        p[0] += R0[0] * vx + R0[1] * vy + R0[2] * vz + s[0] * n[0]; /* :65-67 */
        p[1] += R1[0] * vx + R1[1] * vy + R1[2] * vz + s[1] * n[1]; /* :68 */
        p[2] = z;                                                  /* :70 */
    }
}

/* ------------------------------------------------------------------ exact normalise spec
 * Mirrors pf_oracle.integer_weights / quantise_a.  Built with
 * -ffp-contract=off so that no FMA is formed: every float operation below is one IEEE single
 * multiply or add, exactly what NumPy float32 arithmetic and the CUDA __fmul_rn / __fadd_rn do. */
#define A_BITS 60   /* a_i = floor(E(logw_i - max) * 2^60); A = sum a_i kept exactly in 128 bits */

static uint64_t a_from_logw(float lw, float m, int abits)
{
    volatile float x = lw - m;
    if (!(x >= -41.5f)) return 0;                 /* also NaN, -inf */
    volatile float t = x * 1.44269502f;
    volatile float k0 = t + 12582912.0f;          /* rint(t) by the 1.5 * 2^23 trick */
    volatile float k = k0 - 12582912.0f;
    volatile float h = k * 0.693359375f, l = k * -2.12194440e-4f;
    volatile float r0 = x - h;
    volatile float r = r0 - l;
    static const float C[8] = {1.0f / 5040.0f, 1.0f / 720.0f, 1.0f / 120.0f, 1.0f / 24.0f, 1.0f / 6.0f, 0.5f, 1.0f, 1.0f};
    volatile float p = C[0];
    for (int n = 1; n < 8; ++n) { volatile float q = p * r; p = q + C[n]; }
    float pf = p;
    uint32_t bits;
    memcpy(&bits, &pf, sizeof bits);
    int shift = (int)k + abits - 23 + ((int)(bits >> 23) - 127);
    uint64_t mant = (uint64_t)((bits & 0x7FFFFFu) | 0x800000u);
    return shift >= 0 ? mant << shift : (shift > -24 ? mant >> -shift : 0);
}

/* floor(C 2^52 / A) for C <= A < 2^90 in 128-bit arithmetic, by two 26-bit steps of long division:
 * C 2^26 = q1 A + r1,  r1 2^26 = q2 A + r2  =>  floor(C 2^52 / A) = q1 2^26 + q2.
 * (The device takes a float64 estimate and corrects it with a wrapping multiply-back instead,
 * csrc/auvpf_math.cuh cdf_exact; the two formulations check each other.) */
static uint64_t cdf_floor(unsigned __int128 C, unsigned __int128 A)
{
    unsigned __int128 t = C << 26;
    unsigned __int128 q1 = t / A, r1 = t - q1 * A;
    unsigned __int128 q2 = (r1 << 26) / A;
    return (uint64_t)((q1 << 26) + q2);
}

/* returns 0 ok, -1 degenerate.  qq: uint64[N] with sum == 2^52: q_j = cdf_j - cdf_{j-1},
 * cdf_j = floor(C_j 2^52 / A), C_j = sum_{i<=j} a_i (pf_oracle.quantise_a). */
int orc_quantise_logw(const float *logw, long N, uint64_t *qq, float *m_out, uint64_t A_out[2])
{
    float m = -INFINITY;
    for (long i = 0; i < N; ++i) if (logw[i] > m) m = logw[i];   /* NaN compares false */
    if (!isfinite(m)) return -1;
    unsigned __int128 A = 0;
    for (long i = 0; i < N; ++i) { qq[i] = a_from_logw(logw[i], m, A_BITS); A += qq[i]; }
    if (A == 0) return -1;
    unsigned __int128 C = 0;
    uint64_t prev = 0;
    for (long i = 0; i < N; ++i) {
        C += qq[i];
        uint64_t c = cdf_floor(C, A);
        qq[i] = c - prev;
        prev = c;
    }
    if (m_out) *m_out = m;
    if (A_out) { A_out[0] = (uint64_t)A; A_out[1] = (uint64_t)(A >> 64); }
    return 0;
}

/* the integer weights a_i themselves (for the gate sums and for sharded-total checks) */
int orc_integer_weights(const float *logw, long N, float m, uint64_t *a)
{
    for (long i = 0; i < N; ++i) a[i] = a_from_logw(logw[i], m, A_BITS);
    return 0;
}

/* systematic_resample (resampling.py:135-168): positions = (u + i)/N (:157),
 * sequential cumsum (:160), two-pointer with strict '<' (:161-167).
 * Returns 0, or -1 where the reference would raise IndexError (j == N). */
int orc_systematic(const double *w, long N, double u, int32_t *idx)
{
    double *cs = (double *)malloc(sizeof(double) * (size_t)N);
    double acc = 0.0;
    for (long j = 0; j < N; ++j) { acc += w[j]; cs[j] = acc; }
    long i = 0, j = 0;
    int rc = 0;
    while (i < N) {
        if (j >= N) { rc = -1; break; }
        double pos = (u + (double)i) / (double)N;
        if (pos < cs[j]) { idx[i] = (int32_t)j; ++i; }
        else ++j;
    }
    free(cs);
    return rc;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
