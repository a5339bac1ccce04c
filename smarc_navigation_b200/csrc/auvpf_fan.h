// auvpf_fan.h -- host-side analysis of the beam table (set_beams): is it a planar fan that
// the kernels may sweep (auvpf_math.cuh "Fan sweep"), and in which order does each lane
// process its chunk of beams.  Plain C++, shared by the C ABI and the test-only host simulator.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace auvpf {

constexpr int kFanLanes = 32;
constexpr int kOrdFull = 1 << 30;    // flag in the order table: start of a sweep -> full traversal

struct FanPlan {
    bool planar = false;             // beams form a sorted planar fan (sweep allowed)
    float c0[3] = {0, 0, -1};        // fan centre direction (sensor frame)
    int chunk = 1;                   // beams per lane (loop trip count)
    int lanes_first = 0;             // lanes holding the first side (lower beam indices) of the fan
    std::vector<float> s, inv_s;     // |lateral component| of beam b (= |sin th_b| for a unit vector), reciprocal (0 = unusable)
    std::vector<double> tan_t;       // |lateral| / along-centre component (profile merge; float64: it multiplies q ~ 30 m)
    std::vector<double> sec_t;       // 1 / along-centre component: range = q sec (exact for the vector as given)
    float a0[3] = {0, 1, 0};         // lateral unit vector (sensor frame): th > 0 leans towards +a0
    int n1 = 0;                      // beams [0, n1) are one side of the centre, [n1, B) the other
    int sgn1 = -1;                   // sign of th on the first side
    std::vector<int32_t> order;      // [32 * chunk]: beam index | kOrdFull, or -1 (padding)
};

// Lane layout of a planar fan: the beams of each side of the centre are split into contiguous
// chunks, one per lane, and every lane processes its chunk OUTWARD (entry 0 = the beam nearest
// the centre).  No chunk straddles the centre.  Lanes [0, lanes_first) hold the first side
// (their inner neighbour is lane + 1), lanes [lanes_first, 32) the second side (inner
// neighbour lane - 1); the kernels chain the chunk starts outward through those neighbours.
inline FanPlan plan_fan(const float *dirs, int B)
{
    FanPlan f;
    f.s.assign(B, 0.f);
    f.inv_s.assign(B, 0.f);
    f.tan_t.assign(B, 0.0);
    f.sec_t.assign(B, 1.0);
    std::vector<double> th(B, 0.0);
    auto D = [&](int b, int a) { return (double)dirs[3 * b + a]; };
    // plane normal from the most separated pair (first / last), centre = normalised mean
    if (B >= 3) {
        double n[3] = {D(0, 1) * D(B - 1, 2) - D(0, 2) * D(B - 1, 1), D(0, 2) * D(B - 1, 0) - D(0, 0) * D(B - 1, 2),
                       D(0, 0) * D(B - 1, 1) - D(0, 1) * D(B - 1, 0)};
        double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        double c[3] = {0, 0, 0};
        for (int b = 0; b < B; ++b)
            for (int a = 0; a < 3; ++a) c[a] += D(b, a);
        double cn = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        if (nn > 1e-3 && cn > 1e-3 * B) {
            for (int a = 0; a < 3; ++a) {
                n[a] /= nn;
                c[a] /= cn;
            }
            double dot = c[0] * n[0] + c[1] * n[1] + c[2] * n[2];      // project the centre into the plane
            for (int a = 0; a < 3; ++a) c[a] -= dot * n[a];
            cn = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
            for (int a = 0; a < 3; ++a) c[a] /= cn;
            const double a0[3] = {n[1] * c[2] - n[2] * c[1], n[2] * c[0] - n[0] * c[2], n[0] * c[1] - n[1] * c[0]};
            bool ok = true;
            for (int b = 0; b < B && ok; ++b) {
                double off = D(b, 0) * n[0] + D(b, 1) * n[1] + D(b, 2) * n[2];
                double pc = D(b, 0) * c[0] + D(b, 1) * c[1] + D(b, 2) * c[2];
                double pa = D(b, 0) * a0[0] + D(b, 1) * a0[1] + D(b, 2) * a0[2];
                double len = std::sqrt(D(b, 0) * D(b, 0) + D(b, 1) * D(b, 1) + D(b, 2) * D(b, 2));
                th[b] = std::atan2(pa, pc);
                // off-plane by more than float32 rounding would move a 100 m hit by more than the range tolerance
                if (std::fabs(off) > 2e-7 || std::fabs(len - 1.0) > 1e-4 || pc < 0.17) ok = false;   // |th| < ~80 deg
            }
            for (int b = 1; b < B && ok; ++b)
                if (!((th[b] > th[b - 1]) == (th[1] > th[0])) || th[b] == th[b - 1]) ok = false;       // strictly monotone
            if (ok) {
                f.planar = true;
                for (int a = 0; a < 3; ++a) {
                    f.c0[a] = (float)c[a];
                    f.a0[a] = (float)a0[a];
                }
                for (int b = 0; b < B; ++b) {
                    const double pc = D(b, 0) * c[0] + D(b, 1) * c[1] + D(b, 2) * c[2];
                    const double pa = std::fabs(D(b, 0) * a0[0] + D(b, 1) * a0[1] + D(b, 2) * a0[2]);
                    f.s[b] = (float)pa;
                    f.inv_s[b] = pa > 1e-6 ? (float)(1.0 / pa) : 0.f;
                    f.tan_t[b] = pa / pc;
                    f.sec_t[b] = 1.0 / pc;
                }
            }
        }
    }
    if (!f.planar) {
        // no fan: every beam is traversed from the sensor; lane l takes beams l, l+32, ... (adjacent
        // lanes = adjacent beams keeps the warp's rays coherent)
        f.chunk = (B + kFanLanes - 1) / kFanLanes;
        f.lanes_first = 0;
        f.order.assign((size_t)kFanLanes * f.chunk, -1);
        for (int b = 0; b < B; ++b) f.order[(size_t)(b % kFanLanes) * f.chunk + b / kFanLanes] = b | kOrdFull;
        return f;
    }
    // sides of the centre in index order: [0, n1) has one sign of th, [n1, B) the other
    int n1 = 0;
    while (n1 < B && (th[n1] < 0.0) == (th[0] < 0.0)) ++n1;
    const int n2 = B - n1;
    f.n1 = n1;
    f.sgn1 = th[0] < 0.0 ? -1 : 1;
    int l1 = n2 == 0 ? kFanLanes : (int)std::lround((double)kFanLanes * n1 / B);
    if (n1 > 0 && l1 < 1) l1 = 1;
    if (n2 > 0 && l1 > kFanLanes - 1) l1 = kFanLanes - 1;
    l1 = std::min(l1, n1);
    const int l2 = std::min(kFanLanes - l1, n2);
    const int c1 = l1 ? (n1 + l1 - 1) / l1 : 0, c2 = l2 ? (n2 + l2 - 1) / l2 : 0;
    f.chunk = std::max(1, std::max(c1, c2));
    f.lanes_first = l1;
    f.order.assign((size_t)kFanLanes * f.chunk, -1);
    // first side: inner end is index n1-1; lane l1-1 is innermost.  Chunks are cut from the inside.
    for (int j = 0; j < n1; ++j) {                 // j = distance from the centre
        int lane = l1 - 1 - j / c1, k = j % c1, b = n1 - 1 - j;
        f.order[(size_t)lane * f.chunk + k] = b | (k == 0 ? kOrdFull : 0);
    }
    for (int j = 0; j < n2; ++j) {
        int lane = l1 + j / c2, k = j % c2, b = n1 + j;
        f.order[(size_t)lane * f.chunk + k] = b | (k == 0 ? kOrdFull : 0);
    }
    return f;
}

}   // namespace auvpf
