/*
 * oracle_rng.h -- TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
 *
 * The random-number and transcendental-math CONTRACT shared by the three
 * implementations of the self-play hot path:
 *   (1) the reference sources compiled verbatim (oracle/_ref), whose
 *       utils/random.cpp is replaced at link time by ref_harness/random_shim.cpp
 *       which includes this header;
 *   (2) the plain-C restatement in oracle/ (this directory);
 *   (3) the CUDA product path, which RESTATES this contract independently in
 *       sprl_b200/csrc/rng.cuh (it does not include this file).
 *
 * Why a contract at all: the reference draws from one global PCG32 stream
 * seeded from std::random_device (/root/reference/cpp/src/utils/random.cpp:38-47,
 * constants.hpp:4 SEED=0) through libstdc++ distributions
 * (utils/random.cpp:61-98), which are implementation defined.  Nothing can be
 * bit-compared against that, CPU or GPU.  So the stream is redefined as a
 * counter-based one keyed (seed, game, draw index); the reference's call order
 * inside one game is kept exactly (tie-break draws, symmetry draws, Dirichlet
 * draws, the move sample), which a sequential-per-tree GPU warp can replay.
 *
 * Everything here uses only IEEE-754 + - * / sqrt on float/double, evaluated
 * in a fixed order, so that gcc (-ffp-contract=off) and nvcc (-fmad=false)
 * produce identical bits.  No libm call is made.
 */
#ifndef SPRL_ORACLE_RNG_H
#define SPRL_ORACLE_RNG_H

#include <stdint.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t seed;  /* run seed                                  */
    uint64_t game;  /* game id (stream id)                       */
    uint64_t ctr;   /* index of the next 32-bit draw in the game */
} orng_t;

/* ---- Philox4x32-10 (Salmon et al. 2011), word 0 of the block ------------- */
static inline uint32_t orng_philox_word0(uint64_t seed, uint64_t game, uint64_t ctr) {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32);
    uint32_t c2 = (uint32_t)game, c3 = (uint32_t)(game >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

static inline uint32_t orng_u32(orng_t* r) {
    uint32_t v = orng_philox_word0(r->seed, r->game, r->ctr);
    r->ctr += 1;
    return v;
}

/* Replaces Random::UniformInt (utils/random.cpp:76-79).  a==b consumes NO draw
 * (the reference's single-candidate tie-break, uct/UCTNode.hpp:250, calls the
 * RNG; under the contract that call is free).  Otherwise one draw, multiply-
 * shift range reduction (bias < 2^-25 for ranges <= 128). */
static inline int orng_uniform_int(orng_t* r, int a, int b) {
    if (a == b) return a;
    uint32_t n = (uint32_t)(b - a + 1);
    uint32_t v = orng_u32(r);
    return a + (int)(((uint64_t)v * n) >> 32);
}

/* Replaces Random::UniformUint64 (utils/random.cpp:81-84); only the full range
 * is ever requested (utils/Zobrist.hpp:39-40). */
static inline uint64_t orng_u64(orng_t* r) {
    uint64_t hi = orng_u32(r);
    uint64_t lo = orng_u32(r);
    return (hi << 32) | lo;
}

/* 24-bit uniform in [0,1): exact in float. */
static inline float orng_unit_f32(orng_t* r) {
    return (float)(orng_u32(r) >> 8) * 5.9604644775390625e-08f; /* 2^-24 */
}
/* 24-bit uniform in (0,1), never 0: (k + 0.5) * 2^-24, exact in double. */
static inline double orng_unit_open(orng_t* r) {
    return ((double)(orng_u32(r) >> 8) + 0.5) * 5.9604644775390625e-08;
}

/* ---- deterministic log / exp in double ----------------------------------- */
static inline double odet_from_bits(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static inline uint64_t odet_to_bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }

/* 2^k for -1022 <= k <= 1023 */
static inline double odet_pow2i(int k) { return odet_from_bits((uint64_t)(k + 1023) << 52); }

/* natural log of a positive finite double; ~1 ulp; deterministic. */
static inline double odet_log(double x) {
    int k = 0;
    uint64_t b = odet_to_bits(x);
    if ((b >> 52) == 0) {            /* subnormal: scale up by 2^54 */
        x = x * 18014398509481984.0;
        b = odet_to_bits(x);
        k = -54;
    }
    k += (int)((b >> 52) & 0x7ff) - 1023;
    double m = odet_from_bits((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL); /* [1,2) */
    if (m > 1.4142135623730951) { m = m * 0.5; k += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    /* atanh series: log(m) = 2 s (1 + z/3 + z^2/5 + ... + z^11/23) */
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
    /* ln2 split: hi has 32 significant bits so kd*hi is exact */
    return kd * 0.693147180369123816490 + (kd * 1.90821492927058770002e-10 + lm);
}

/* e^x; deterministic; returns 0 / +huge outside the double range. */
static inline double odet_exp(double x) {
    if (x < -745.0) return 0.0;
    if (x > 709.0) return 1.7976931348623157e308;
    double t = x * 1.4426950408889634 + 0.5;
    long long ki = (long long)t;          /* trunc toward zero */
    if ((double)ki > t) ki -= 1;          /* -> floor */
    int k = (int)ki;
    double kd = (double)k;
    double r = (x - kd * 0.693147180369123816490) - kd * 1.90821492927058770002e-10;
    /* Taylor to degree 13, Horner */
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
    int k1 = k / 2, k2 = k - k1;          /* two-step scaling keeps both factors normal */
    return (p * odet_pow2i(k1)) * odet_pow2i(k2);
}

/* powf replacement for pdf.pow(0.98f | 10.0f) (selfplay/SelfPlay.hpp:115-119,
 * games/GameActionDist.hpp:111-119).  x >= 0. */
static inline float odet_powf(float x, float e) {
    if (x == 0.0f) return 0.0f;
    if (x == 1.0f) return 1.0f;
    return (float)odet_exp((double)e * odet_log((double)x));
}
/* expf replacement for policy.exp() (networks/GridNetwork.hpp:114). */
static inline float odet_expf(float x) { return (float)odet_exp((double)x); }

/* sqrt: IEEE exact everywhere; spelled through the compiler builtin so no libm
 * symbol is needed (gcc emits sqrtsd). */
static inline double odet_sqrt(double x) { return __builtin_sqrt(x); }

/* ---- Normal / Gamma / Dirichlet ------------------------------------------ */
/* Marsaglia polar method; second variate discarded. */
static inline double orng_normal(orng_t* r) {
    for (;;) {
        double u1 = 2.0 * orng_unit_open(r) - 1.0;
        double u2 = 2.0 * orng_unit_open(r) - 1.0;
        double s = u1 * u1 + u2 * u2;
        if (s >= 1.0 || s == 0.0) continue;
        double f = odet_sqrt(-2.0 * odet_log(s) / s);
        return u1 * f;
    }
}

/* Gamma(alpha, 1), Marsaglia & Tsang 2000, with the alpha<1 boost. */
static inline float orng_gamma(orng_t* r, float alpha_f) {
    double alpha = (double)alpha_f;
    double a = (alpha < 1.0) ? alpha + 1.0 : alpha;
    double d = a - 1.0 / 3.0;
    double c = 1.0 / odet_sqrt(9.0 * d);
    double g;
    for (;;) {
        double x, v;
        do {
            x = orng_normal(r);
            v = 1.0 + c * x;
        } while (v <= 0.0);
        v = v * v * v;
        double u = orng_unit_open(r);
        double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) { g = d * v; break; }
        if (odet_log(u) < 0.5 * x2 + d * (1.0 - v + odet_log(v))) { g = d * v; break; }
    }
    if (alpha < 1.0) {
        double u = orng_unit_open(r);
        g = g * odet_exp(odet_log(u) / alpha);
    }
    return (float)g;
}

/* Replaces Random::Dirichlet (utils/random.cpp:61-74): same float arithmetic
 * around the gamma draws (sequential float sum, norm = 1/sum, multiply). */
static inline void orng_dirichlet(orng_t* r, float alpha, float* samples, int n) {
    float sum = 0;
    for (int i = 0; i < n; ++i) {
        samples[i] = orng_gamma(r, alpha);
        sum += samples[i];
    }
    float norm = 1 / sum;
    for (int i = 0; i < n; ++i) samples[i] *= norm;
}

/* Replaces Random::SampleCDF (utils/random.cpp:86-98): reject e==0,
 * x = cdf.back()*e, first index with cdf[i] >= x (std::lower_bound). */
static inline int orng_sample_cdf(orng_t* r, const float* cdf, int n) {
    float e;
    do { e = orng_unit_f32(r); } while (e == 0);
    float x = cdf[n - 1] * e;
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        if (cdf[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* ---- HashNet: deterministic evaluator for search parity (ours, both sides) */
static inline uint64_t ohash_mix(uint64_t z) {   /* splitmix64 finaliser */
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* Hash of a (symmetrised) network input state.  `words` holds, for each of the
 * `size` valid history boards t (0 = current), four 64-bit words:
 *   own cells 0..63, own cells 64..127, opponent cells 0..63, opponent 64..127
 * where "own" = stones of the side to move, bit i = cell i (row-major).
 * player: 0 / 1 = side to move. */
static inline uint64_t ohashnet_state_hash(const uint64_t* words, int size, int player) {
    uint64_t h = ohash_mix(0x5350524C42323030ULL ^ (uint64_t)player);
    for (int i = 0; i < 4 * size; ++i) h = ohash_mix(h ^ words[i]);
    return h;
}
/* A second, independent HashNet for two-network matches (Evaluate.cpp): salt 0 is the plain net. */
static inline uint64_t ohashnet_salt(uint64_t h, uint64_t salt) {
    return salt ? ohash_mix(h ^ (salt * 0xD6E8FEB86659FD93ULL)) : h;
}
/* Un-normalised prior of action index i: an integer in [1, 2^24], exact in fp32.
 * The caller then applies the reference's own mask -> sequential sum ->
 * x * (1/sum) pipeline (networks/GridNetwork.hpp:117-139). */
static inline float ohashnet_prior_raw(uint64_t h, int i) {
    return (float)((ohash_mix(h ^ ((uint64_t)(i + 1) << 32)) >> 40) + 1);
}
/* Value in [-1, 1), a multiple of 2^-23. */
static inline float ohashnet_value(uint64_t h) {
    int v = (int)(ohash_mix(h ^ 0xABCDEFULL) >> 40);
    return ((float)v - 8388608.0f) * 1.1920928955078125e-07f;
}

#ifdef __cplusplus
}
#endif
#endif
