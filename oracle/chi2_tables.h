// chi2_tables.h — TEST INFRASTRUCTURE: the two tables of the reference's chi-square BSDF test.
//
// Methodology of crates/akari_api/src/bin/akari_test.rs:31-112 (pdf_histogram + integrate_pdf):
//   observed[theta_bin][phi_bin] = how many of `n_samples` directions drawn by sample_wi(wo, u) fall into the bin
//   expected[theta_bin][phi_bin] = n_samples * integral over the bin of pdf(wo, wi) sin(theta) dtheta dphi
// with theta = acos(z) and phi = atan2(y, x) (the reference's `xyz_to_spherical` / `spherical_to_xyz`,
// geometry.rs:325-336).  The chi-square statistic itself (pooling, Sidak) is evaluated by
// tests/test_bsdf_chi2.py.  Generic over the closure under test: the oracle instantiates it with the literal closure
// tree, tests/hostsim with the device closures of akari_render_b200/csrc/device/akr_bsdf.cuh.
#pragma once
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

namespace chi2 {

struct Dir {
    float x, y, z;
};

// PCG32 (sampler/mod.rs:77-132 restates the same generator); one stream per worker thread
struct Pcg32 {
    uint64_t state, inc;
    explicit Pcg32(uint64_t seq, uint64_t seed = 0x853c49e6748fea9bULL) {
        state = 0u;
        inc = (seq << 1u) | 1u;
        next();
        state += seed;
        next();
    }
    uint32_t next() {
        uint64_t old = state;
        state = old * 6364136223846793005ULL + inc;
        uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((32u - rot) & 31u));
    }
    float uniform() {  // [0, 1)
        float f = (float)(next() >> 8) * (1.0f / 16777216.0f);
        return f;
    }
};

inline void bin_of(Dir w, uint32_t theta_res, uint32_t phi_res, uint32_t &bt, uint32_t &bp) {
    const double PI = 3.14159265358979323846;
    double len = std::sqrt((double)w.x * w.x + (double)w.y * w.y + (double)w.z * w.z);
    double theta = std::acos(std::fmin(1.0, std::fmax(-1.0, w.z / len)));
    double phi = std::atan2((double)w.y, (double)w.x);
    if (phi < 0.0) phi += 2.0 * PI;
    double t = theta / PI, p = phi / (2.0 * PI);
    bt = (uint32_t)std::fmin((double)theta_res - 1.0, std::floor(t * theta_res));
    bp = (uint32_t)std::fmin((double)phi_res - 1.0, std::floor(p * phi_res));
}
inline Dir dir_of(double theta, double phi) {
    double s = std::sin(theta);
    return Dir{(float)(s * std::cos(phi)), (float)(s * std::sin(phi)), (float)std::cos(theta)};
}

// Sample(u_select, u0, u1, Dir &wi) -> bool valid
template <class Sample> void histogram(Sample sample, uint64_t n_samples, uint64_t seed, uint32_t theta_res, uint32_t phi_res, uint32_t *hist, int n_threads) {
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    const size_t n_bins = (size_t)theta_res * phi_res;
    std::vector<std::vector<uint32_t>> part((size_t)n_threads, std::vector<uint32_t>(n_bins, 0u));
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&, t] {
            Pcg32 rng((uint64_t)t + 1u, seed * 0x9e3779b97f4a7c15ULL + 0x853c49e6748fea9bULL);
            uint64_t begin = n_samples * (uint64_t)t / (uint64_t)n_threads, end = n_samples * (uint64_t)(t + 1) / (uint64_t)n_threads;
            for (uint64_t i = begin; i < end; ++i) {
                float us = rng.uniform(), u0 = rng.uniform(), u1 = rng.uniform();
                Dir wi;
                if (!sample(us, u0, u1, wi)) continue;
                uint32_t bt, bp;
                bin_of(wi, theta_res, phi_res, bt, bp);
                part[(size_t)t][(size_t)bt * phi_res + bp]++;
            }
        });
    for (auto &th : pool) th.join();
    for (size_t b = 0; b < n_bins; ++b) {
        uint32_t s = 0;
        for (int t = 0; t < n_threads; ++t) s += part[(size_t)t][b];
        hist[b] = s;
    }
}

// adaptive Simpson in one variable (util/integration.rs:18-106 does the same recursion on the device)
template <class F> double simpson_rec(F &f, double a, double fa, double b, double fb, double m, double fm, double whole, double eps, int depth) {
    double lm = 0.5 * (a + m), rm = 0.5 * (m + b);
    double flm = f(lm), frm = f(rm);
    double left = (m - a) / 6.0 * (fa + 4.0 * flm + fm), right = (b - m) / 6.0 * (fm + 4.0 * frm + fb);
    double delta = left + right - whole;
    if (depth <= 0 || std::fabs(delta) <= 15.0 * eps) return left + right + delta / 15.0;
    return simpson_rec(f, a, fa, m, fm, lm, flm, left, 0.5 * eps, depth - 1) + simpson_rec(f, m, fm, b, fb, rm, frm, right, 0.5 * eps, depth - 1);
}
template <class F> double simpson(F f, double a, double b, double eps, int depth) {
    double m = 0.5 * (a + b);
    double fa = f(a), fb = f(b), fm = f(m);
    double whole = (b - a) / 6.0 * (fa + 4.0 * fm + fb);
    return simpson_rec(f, a, fa, b, fb, m, fm, whole, eps, depth);
}

// Pdf(Dir wi) -> float.  expected[bin] = n_samples * integral of pdf * sin(theta) over the bin
template <class Pdf> void expected(Pdf pdf, uint64_t n_samples, uint32_t theta_res, uint32_t phi_res, double *out, int n_threads) {
    const double PI = 3.14159265358979323846;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    const double th = PI / theta_res, ph = 2.0 * PI / phi_res;
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&, t] {
            for (uint32_t i = (uint32_t)t; i < theta_res; i += (uint32_t)n_threads)
                for (uint32_t j = 0; j < phi_res; ++j) {
                    const double t0 = th * i, t1 = th * (i + 1), p0 = ph * j, p1 = ph * (j + 1);
                    // a fixed 2 x 2 split in front of the recursion keeps narrow lobes from hiding between the first three nodes
                    double sum = 0.0;
                    for (int a = 0; a < 2; ++a)
                        for (int b = 0; b < 2; ++b) {
                            const double ta = t0 + (t1 - t0) * 0.5 * a, tb = ta + (t1 - t0) * 0.5;
                            const double pa = p0 + (p1 - p0) * 0.5 * b, pb = pa + (p1 - p0) * 0.5;
                            sum += simpson(
                                [&](double theta) {
                                    const double s = std::sin(theta);
                                    return simpson(
                                        [&](double phi) {
                                            float v = pdf(dir_of(theta, phi));
                                            return std::isfinite(v) ? (double)v * s : 0.0;
                                        },
                                        pa, pb, 1e-7, 8);
                                },
                                ta, tb, 1e-7, 8);
                        }
                    out[(size_t)i * phi_res + j] = sum * (double)n_samples;
                }
        });
    for (auto &th2 : pool) th2.join();
}

}  // namespace chi2
