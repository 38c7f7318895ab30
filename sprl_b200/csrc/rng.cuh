// rng.cuh -- counter-based random streams and deterministic transcendental math
// for the self-play kernels.
//
// Replaces the reference's global PCG32 + libstdc++ distributions
// (/root/reference/cpp/src/utils/random.hpp:32-118, utils/random.cpp:29-98) with
// one Philox4x32-10 stream per game, indexed (seed, game id, draw index), so a
// warp that owns a tree can replay the reference's draw order exactly:
// tie-break draws during select, one symmetry draw per queued leaf, Dirichlet
// draws at root expansion, the move sample.  All arithmetic is IEEE + - * / sqrt
// in a fixed order (this library is compiled with -fmad=false), so a CPU that
// evaluates the same expressions gets the same bits.
#pragma once
#include <cstdint>

namespace sprl {

struct Rng {
    uint64_t seed;
    uint64_t game;
    uint64_t ctr;
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// Philox4x32-10; only word 0 of each block is used (one block per draw index).
__host__ __device__ __forceinline__ uint32_t philox_word0(uint64_t seed, uint64_t game, uint64_t ctr) {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32);
    uint32_t c2 = (uint32_t)game, c3 = (uint32_t)(game >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

__host__ __device__ __forceinline__ uint32_t rng_u32(Rng& r) {
    uint32_t v = philox_word0(r.seed, r.game, r.ctr);
    r.ctr += 1;
    return v;
}

// Random::UniformInt (utils/random.cpp:76-79).  a == b draws nothing.
__host__ __device__ __forceinline__ int rng_uniform_int(Rng& r, int a, int b) {
    if (a == b) return a;
    return a + (int)mulhi32(rng_u32(r), (uint32_t)(b - a + 1));
}

__host__ __device__ __forceinline__ float rng_unit_f32(Rng& r) {
    return (float)(rng_u32(r) >> 8) * 5.9604644775390625e-08f;
}
__host__ __device__ __forceinline__ double rng_unit_open(Rng& r) {
    return ((double)(rng_u32(r) >> 8) + 0.5) * 5.9604644775390625e-08;
}

// ---- deterministic log / exp / pow ------------------------------------------
__host__ __device__ __forceinline__ double bits_to_double(uint64_t b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double d; __builtin_memcpy(&d, &b, 8); return d;
#endif
}
__host__ __device__ __forceinline__ uint64_t double_to_bits(double d) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b; __builtin_memcpy(&b, &d, 8); return b;
#endif
}
__host__ __device__ __forceinline__ double pow2i(int k) { return bits_to_double((uint64_t)(k + 1023) << 52); }

__host__ __device__ inline double det_log(double x) {
    int k = 0;
    uint64_t b = double_to_bits(x);
    if ((b >> 52) == 0) { x = x * 18014398509481984.0; b = double_to_bits(x); k = -54; }
    k += (int)((b >> 52) & 0x7ff) - 1023;
    double m = bits_to_double((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > 1.4142135623730951) { m = m * 0.5; k += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 23.0;
    p = p * z + 1.0 / 21.0;
    p = p * z + 1.0 / 19.0;
    p = p * z + 1.0 / 17.0;
    p = p * z + 1.0 / 15.0;
    p = p * z + 1.0 / 13.0;
    p = p * z + 1.0 / 11.0;
    p = p * z + 1.0 / 9.0;
    p = p * z + 1.0 / 7.0;
    p = p * z + 1.0 / 5.0;
    p = p * z + 1.0 / 3.0;
    p = p * z + 1.0;
    double lm = 2.0 * s * p;
    double kd = (double)k;
    return kd * 0.693147180369123816490 + (kd * 1.90821492927058770002e-10 + lm);
}

__host__ __device__ inline double det_exp(double x) {
    if (x < -745.0) return 0.0;
    if (x > 709.0) return 1.7976931348623157e308;
    double t = x * 1.4426950408889634 + 0.5;
    long long ki = (long long)t;
    if ((double)ki > t) ki -= 1;
    int k = (int)ki;
    double kd = (double)k;
    double r = (x - kd * 0.693147180369123816490) - kd * 1.90821492927058770002e-10;
    double p = 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    int k1 = k / 2, k2 = k - k1;
    return (p * pow2i(k1)) * pow2i(k2);
}

// pdf.pow(0.98f | 10.0f) of selfplay/SelfPlay.hpp:115-119
__host__ __device__ inline float det_powf(float x, float e) {
    if (x == 0.0f) return 0.0f;
    if (x == 1.0f) return 1.0f;
    return (float)det_exp((double)e * det_log((double)x));
}
// policy.exp() of networks/GridNetwork.hpp:114
__host__ __device__ inline float det_expf(float x) { return (float)det_exp((double)x); }

__host__ __device__ __forceinline__ double det_sqrt(double x) {
#ifdef __CUDA_ARCH__
    return __dsqrt_rn(x);
#else
    return __builtin_sqrt(x);
#endif
}

// ---- Normal / Gamma ------------------------------------------------------------
__host__ __device__ inline double rng_normal(Rng& r) {
    for (;;) {
        double u1 = 2.0 * rng_unit_open(r) - 1.0;
        double u2 = 2.0 * rng_unit_open(r) - 1.0;
        double s = u1 * u1 + u2 * u2;
        if (s >= 1.0 || s == 0.0) continue;
        double f = det_sqrt(-2.0 * det_log(s) / s);
        return u1 * f;
    }
}

// Gamma(alpha, 1): Marsaglia-Tsang with the alpha < 1 boost.  Stands in for the
// std::gamma_distribution<float> of Random::Dirichlet (utils/random.cpp:62-67).
__host__ __device__ inline float rng_gamma(Rng& r, float alpha_f) {
    double alpha = (double)alpha_f;
    double a = (alpha < 1.0) ? alpha + 1.0 : alpha;
    double d = a - 1.0 / 3.0;
    double c = 1.0 / det_sqrt(9.0 * d);
    double g;
    for (;;) {
        double x, v;
        do {
            x = rng_normal(r);
            v = 1.0 + c * x;
        } while (v <= 0.0);
        v = v * v * v;
        double u = rng_unit_open(r);
        double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) { g = d * v; break; }
        if (det_log(u) < 0.5 * x2 + d * (1.0 - v + det_log(v))) { g = d * v; break; }
    }
    if (alpha < 1.0) {
        double u = rng_unit_open(r);
        g = g * det_exp(det_log(u) / alpha);
    }
    return (float)g;
}

// ---- HashNet: deterministic test evaluator --------------------------------------
__host__ __device__ __forceinline__ uint64_t hash_mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t hashnet_seed(int player) {
    return hash_mix(0x5350524C42323030ULL ^ (uint64_t)player);
}
__host__ __device__ __forceinline__ uint64_t hashnet_salt(uint64_t h, uint64_t salt) {   // second net of a match; 0 = plain
    return salt ? hash_mix(h ^ (salt * 0xD6E8FEB86659FD93ULL)) : h;
}
__host__ __device__ __forceinline__ float hashnet_prior_raw(uint64_t h, int i) {
    return (float)((hash_mix(h ^ ((uint64_t)(i + 1) << 32)) >> 40) + 1);
}
__host__ __device__ __forceinline__ float hashnet_value(uint64_t h) {
    int v = (int)(hash_mix(h ^ 0xABCDEFULL) >> 40);
    return ((float)v - 8388608.0f) * 1.1920928955078125e-07f;
}

}  // namespace sprl
