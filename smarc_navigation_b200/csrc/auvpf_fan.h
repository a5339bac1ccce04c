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
    int chunk = 1;                   // beams per lane
    std::vector<float> s, inv_s;     // |sin th_b| and its reciprocal (0 = unusable)
    std::vector<int32_t> order;      // [32 * chunk]: beam index | kOrdFull, or -1 (padding)
};

inline FanPlan plan_fan(const float *dirs, int B)
{
    FanPlan f;
    f.chunk = (B + kFanLanes - 1) / kFanLanes;
    f.s.assign(B, 0.f);
    f.inv_s.assign(B, 0.f);
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
                if (std::fabs(off) > 1e-5 || std::fabs(len - 1.0) > 1e-4 || pc < 0.17) ok = false;   // |th| < ~80 deg
            }
            for (int b = 1; b < B && ok; ++b)
                if (!((th[b] > th[b - 1]) == (th[1] > th[0])) || th[b] == th[b - 1]) ok = false;       // strictly monotone
            if (ok) {
                f.planar = true;
                for (int a = 0; a < 3; ++a) f.c0[a] = (float)c[a];
                for (int b = 0; b < B; ++b) {
                    double sn = std::fabs(std::sin(th[b]));
                    f.s[b] = (float)sn;
                    f.inv_s[b] = sn > 1e-3 ? (float)(1.0 / sn) : 0.f;
                }
            }
        }
    }
    // per-lane processing order: each side of the centre outward; without a fan every beam is a start
    f.order.assign((size_t)kFanLanes * f.chunk, -1);
    for (int l = 0; l < kFanLanes; ++l) {
        int lo = l * f.chunk, hi = std::min(B, lo + f.chunk), k = 0;
        if (lo >= hi) continue;
        std::vector<int> neg, pos;
        for (int b = lo; b < hi; ++b) (f.planar && th[b] < 0.0 ? neg : pos).push_back(b);
        auto outward = [&](int x, int y) { return std::fabs(th[x]) < std::fabs(th[y]); };
        if (f.planar) {
            std::sort(neg.begin(), neg.end(), outward);
            std::sort(pos.begin(), pos.end(), outward);
        }
        for (size_t i = 0; i < neg.size(); ++i) f.order[(size_t)l * f.chunk + k++] = neg[i] | ((i == 0 || !f.planar) ? kOrdFull : 0);
        for (size_t i = 0; i < pos.size(); ++i) f.order[(size_t)l * f.chunk + k++] = pos[i] | ((i == 0 || !f.planar) ? kOrdFull : 0);
    }
    return f;
}

}   // namespace auvpf
