/* oracle/detmath.h — TEST INFRASTRUCTURE (parity oracle), not product code.
 *
 * "detmath": a deterministic fp32 libm subset + the counter-based RNG that the
 * oracle and the CUDA path both implement from the same written spec (DESIGN.md
 * "Determinism contract"). Every function uses only IEEE-754 binary32 +,-,*,/,sqrt,
 * rint/floor, integer ops and comparisons, evaluated in the order written (compile
 * with -ffp-contract=off), so a CPU and a GPU produce identical bits.
 *
 * Why: the reference calls the platform libm through Rust's f32::{sin,cos,acos,atan2,
 * exp,ln,powf} (e.g. /root/reference/src/mc.rs:50, src/bxdf/microfacet/beckmann.rs:35-46,
 * src/bxdf/merl.rs:62-75). glibc and CUDA libdevice differ by ulps, which would make
 * whole-path parity tolerance-only. With detmath the per-sample radiance of the GPU
 * path is bit-identical to this oracle. Building with -DORC_SYSTEM_LIBM swaps these for
 * glibc (what the Rust binary would call) to show the choice is statistically neutral.
 *
 * Polynomials are the classic Cephes single-precision kernels (public domain
 * coefficients); accuracy is ~1-2 ulp on the ranges the path uses.
 */
#ifndef ORC_DETMATH_H
#define ORC_DETMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#define DM_PI 3.14159265358979323846f
#define DM_PIO2 1.57079632679489661923f
#define DM_PIO4 0.78539816339744830962f

static inline float dm_from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t dm_to_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
/* 2^k for k in [-126, 127] */
static inline float dm_pow2i(int k) { return dm_from_bits((uint32_t)(k + 127) << 23); }

/* x = q*(pi/2) + r, |r| <= pi/4, 3-term Cody-Waite; valid for |x| <= 1e5 */
static inline float dm_reduce_pio2(float x, int* q) {
    float kf = rintf(x * 0.636619772367581343f);
    *q = (int)kf;
    float r = x - kf * 1.5703125f;
    r = r - kf * 4.837512969970703125e-4f;
    r = r - kf * 7.54978995489188216e-8f;
    return r;
}
static inline float dm_sin_poly(float r) {
    float z = r * r;
    return r + r * z * (-1.6666654611e-1f + z * (8.3321608736e-3f + z * -1.9515295891e-4f));
}
static inline float dm_cos_poly(float r) {
    float z = r * r;
    return 1.0f - 0.5f * z + z * z * (4.166664568298827e-2f + z * (-1.388731625493765e-3f + z * 2.443315711809948e-5f));
}
static inline float dm_sinf(float x) {
    if (!(fabsf(x) <= 1.0e5f)) return x - x; /* NaN for inf/NaN, 0 beyond the supported range */
    int q;
    float r = dm_reduce_pio2(x, &q);
    switch (q & 3) {
        case 0: return dm_sin_poly(r);
        case 1: return dm_cos_poly(r);
        case 2: return -dm_sin_poly(r);
        default: return -dm_cos_poly(r);
    }
}
static inline float dm_cosf(float x) {
    if (!(fabsf(x) <= 1.0e5f)) return x - x;
    int q;
    float r = dm_reduce_pio2(x, &q);
    switch (q & 3) {
        case 0: return dm_cos_poly(r);
        case 1: return -dm_sin_poly(r);
        case 2: return -dm_cos_poly(r);
        default: return dm_sin_poly(r);
    }
}
/* asin on |z| <= 0.5 */
static inline float dm_asin_core(float z) {
    float z2 = z * z;
    float p = ((((4.2163199048e-2f * z2 + 2.4181311049e-2f) * z2 + 4.5470025998e-2f) * z2 + 7.4953002686e-2f) * z2 +
               1.6666752422e-1f);
    return z + z * z2 * p;
}
/* acos for x in [-1, 1] (callers clamp, as the reference does: linalg/mod.rs:64-66) */
static inline float dm_acosf(float x) {
    if (x != x) return x;
    if (x >= 1.0f) return 0.0f;
    if (x <= -1.0f) return DM_PI;
    if (x > 0.5f) return 2.0f * dm_asin_core(sqrtf(0.5f * (1.0f - x)));
    if (x < -0.5f) return DM_PI - 2.0f * dm_asin_core(sqrtf(0.5f * (1.0f + x)));
    return DM_PIO2 - dm_asin_core(x);
}
/* atan for t >= 0 (including +inf) */
static inline float dm_atan_pos(float t) {
    float y0, u;
    if (t > 2.414213562373095f) { y0 = DM_PIO2; u = -(1.0f / t); }
    else if (t > 0.4142135623730950f) { y0 = DM_PIO4; u = (t - 1.0f) / (t + 1.0f); }
    else { y0 = 0.0f; u = t; }
    float z = u * u;
    float p = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * u + u;
    return y0 + p;
}
/* atan2; signed zeros are treated as +0 (atan2(0,-0) = 0 here, pi in libm) */
static inline float dm_atan2f(float y, float x) {
    if (x != x || y != y) return x + y;
    float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax == 0.0f && ay == 0.0f) a = 0.0f;
    else if (ax == INFINITY && ay == INFINITY) a = DM_PIO4;
    else a = dm_atan_pos(ay / ax);
    if (x < 0.0f) a = DM_PI - a;
    return y < 0.0f ? -a : a;
}
static inline float dm_expf(float x) {
    if (x != x) return x;
    if (x > 88.72283905206835f) return INFINITY;
    if (x < -87.33654475055310f) return 0.0f;
    float kf = floorf(1.44269504088896341f * x + 0.5f);
    float r = x - kf * 0.693359375f;
    r = r - kf * -2.12194440e-4f;
    float z = r * r;
    float p = (((((1.9875691500e-4f * r + 1.3981999507e-3f) * r + 8.3334519073e-3f) * r + 4.1665795894e-2f) * r +
                1.6666665459e-1f) * r + 5.0000001201e-1f) * z + r + 1.0f;
    int k = (int)kf;
    if (k > 127) { p = p * 2.0f; k -= 1; }
    return p * dm_pow2i(k);
}
static inline float dm_logf(float x) {
    if (x != x) return x;
    if (x < 0.0f) return x - x + dm_from_bits(0x7fc00000u);
    if (x == 0.0f) return -INFINITY;
    if (x == INFINITY) return x;
    int e = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
    uint32_t b = dm_to_bits(x);
    e += (int)((b >> 23) & 0xffu) - 126;
    float m = dm_from_bits((b & 0x007fffffu) | 0x3f000000u); /* [0.5, 1) */
    if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.0f; }
    else { m = m - 1.0f; }
    float z = m * m;
    float y = ((((((((7.0376836292e-2f * m - 1.1514610310e-1f) * m + 1.1676998740e-1f) * m - 1.2420140846e-1f) * m +
                   1.4249322787e-1f) * m - 1.6668057665e-1f) * m + 2.0000714765e-1f) * m - 2.4999993993e-1f) * m +
               3.3333331174e-1f) * m * z;
    float fe = (float)e;
    y = y + -2.12194440e-4f * fe;
    y = y + -0.5f * z;
    z = m + y;
    z = z + 0.693359375f * fe;
    return z;
}
/* x^y for x > 0 (only use: sRGB encode, src/film/color.rs:67) */
static inline float dm_powf(float x, float y) { return dm_expf(y * dm_logf(x)); }

/* ---- counter-based RNG (replaces rand::StdRng, multithreaded.rs:79) -------------- */
static inline uint32_t dm_mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
/* One 32-bit draw addressed by (seed, a, b, c). */
static inline uint32_t dm_rng(uint32_t seed, uint32_t a, uint32_t b, uint32_t c) {
    uint32_t h = dm_mix(seed ^ 0x9e3779b9U);
    h = dm_mix(h ^ a);
    h = dm_mix(h ^ b);
    h = dm_mix(h ^ c);
    return h;
}
/* Range<u32>::new(0, u32::MAX).ind_sample (ld.rs:27,55-56,61): uniform on [0, 2^32-1) */
static inline uint32_t dm_scramble(uint32_t h) { return h == 0xffffffffU ? 0xfffffffeU : h; }
/* Rng::next_f32 of rand 0.4: 24 random bits / 2^24, in [0,1) (path.rs:99) */
static inline float dm_next_f32(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

/* Random permutation of [0, l) evaluated per index (Kensler, "Correlated Multi-Jittered
 * Sampling", 2013) — stands in for the serial Fisher-Yates of Rng::shuffle (ld.rs:58,63).
 * Bijective on [0,l) for every key p (checked exhaustively in tests/test_oracle_sampler.py). */
static inline uint32_t dm_permute(uint32_t i, uint32_t l, uint32_t p) {
    uint32_t w = l - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p; i *= 0xe170893dU;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8; i *= 0x0929eb3fU;
        i ^= p >> 23;
        i ^= (i & w) >> 1; i *= 1 | p >> 27;
        i *= 0x6935fa69U;
        i ^= (i & w) >> 11; i *= 0x74dcb303U;
        i ^= (i & w) >> 2; i *= 0x9e501cc3U;
        i ^= (i & w) >> 2; i *= 0xc860a3dfU;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    i += p % l; /* rotation by the key without 32-bit wrap-around (i, p % l < l) */
    return i >= l ? i - l : i;
}

/* RNG stream addressing shared by oracle and GPU (DESIGN.md "RNG"):
 *   per pixel  : dm_rng(seed, pixel, DM_PIXEL_STREAM, dim)   dim in DM_PX_*
 *   per sample : dm_rng(seed, pixel, sample_index, dim)      dim in DM_S_* or DM_S_RR + bounce */
#define DM_PIXEL_STREAM 0xffffffffU
enum { DM_PX_POS0 = 0, DM_PX_POS1 = 1, DM_PX_POS_PERM = 2, DM_PX_TIME = 3, DM_PX_TIME_PERM = 4 };
enum {
    DM_S_L0 = 0, DM_S_L1 = 1, DM_S_L_PERM = 2,
    DM_S_B0 = 3, DM_S_B1 = 4, DM_S_B_PERM = 5,
    DM_S_P0 = 6, DM_S_P1 = 7, DM_S_P_PERM = 8,
    DM_S_LC = 9, DM_S_LC_PERM = 10,
    DM_S_BC = 11, DM_S_BC_PERM = 12,
    DM_S_PC = 13, DM_S_PC_PERM = 14,
    DM_S_RR = 32,
    /* Whitted (integrator/whitted.rs, integrator/mod.rs:41-103): every illumination / specular_reflection / specular_transmission call
     * asks the sampler for fresh 1-element arrays (a new random scramble each, ld.rs:55-63). Node n of the recursion tree (root 1,
     * reflection child 2n, transmission child 2n+1) draws dimension DM_S_WHITTED + 8n + slot:
     *   0,1 light sample_2d   2,3 reflection sample_2d   4 reflection sample_1d   5,6 transmission sample_2d   7 transmission sample_1d */
    DM_S_WHITTED = 4096
};

#endif
