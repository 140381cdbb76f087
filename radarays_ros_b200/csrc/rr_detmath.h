/* rr_detmath.h — deterministic elementary math shared by the sm_100a kernels and the CPU oracle.
 *
 * Why this exists: the reference's wave model (radar_algorithms.h:55-139,168-187; RadarCPU.cpp:459-528)
 * calls libm transcendentals (acos/asin/cos/sin/tan/pow/exp). glibc and CUDA libdevice disagree in the
 * last ulp, which would flip `energy > 0.001` / `angle <= limit` branches and 1-ulp-perturb refracted
 * directions, i.e. break bit-exact face ids / bounce counts / range bins between CPU and GPU.
 * Everything here is built ONLY from IEEE-754 correctly-rounded primitives (+ - * / sqrt fma rint and
 * int<->fp conversion), evaluated in one fixed order, so g++ (-ffp-contract=off) and nvcc (--fmad=false)
 * produce identical bits.  Accuracy of each function is ~1 ulp (tests/test_detmath.py checks vs libm).
 *
 * This header holds PRIMITIVES only (elementary functions, fp32 vector/quaternion algebra as Rmagine
 * defines it, the ray/triangle test that stands in for Embree, Philox4x32-10).  The radar ALGORITHM
 * (Fresnel split, BRDF, pruning, drawing, noise) is written twice, independently: oracle/rr_oracle.cpp
 * (following the reference text) and csrc/rr_frame_kernel.cu (fused GPU form).
 *
 * Coefficient tables are exact Taylor coefficients rounded once to double (tools/gen_detmath_tables.py).
 */
#ifndef RR_DETMATH_H
#define RR_DETMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RR_HD __host__ __device__ __forceinline__
/* heavy transcendental bodies: real calls on the device keep the trace kernel's code inside the instruction cache */
#define RR_HD_CALL static __host__ __device__ __noinline__
#else
#define RR_HD static inline
#define RR_HD_CALL static inline
#endif

/* ------------------------------------------------------------------------------------------------
 * bit casts
 * ---------------------------------------------------------------------------------------------- */
RR_HD uint64_t rr_d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
RR_HD double   rr_u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
RR_HD uint32_t rr_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
RR_HD float    rr_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

#define RR_PIO2_HI 0x1.921fb54442d18p+0   /* nearest double to pi/2            */
#define RR_PIO2_LO 0x1.1a62633145c07p-54  /* pi/2 - RR_PIO2_HI                 */
#define RR_PI_HI   0x1.921fb54442d18p+1
#define RR_PI_LO   0x1.1a62633145c07p-53
#define RR_LN2_HI  0x1.62e42fefa39efp-1
#define RR_LN2_LO  0x1.abc9e3b39803fp-56
#define RR_INV_LN2 0x1.71547652b82fep+0
#define RR_2_OVER_PI 0x1.45f306dc9c883p-1

/* ------------------------------------------------------------------------------------------------
 * sin / cos / tan (double), |x| <~ 1e5 (the wave model only produces |x| < 7)
 * ---------------------------------------------------------------------------------------------- */
RR_HD double rr_ksin(double r)   /* |r| <= pi/4 */
{
    const double z = r * r;
    double p = 0x1.71b8ef6dcf572p-66;
    p = fma(p, z, -0x1.2f49b46814157p-57);
    p = fma(p, z, 0x1.952c77030ad4ap-49);
    p = fma(p, z, -0x1.ae7f3e733b81fp-41);
    p = fma(p, z, 0x1.6124613a86d09p-33);
    p = fma(p, z, -0x1.ae64567f544e4p-26);
    p = fma(p, z, 0x1.71de3a556c734p-19);
    p = fma(p, z, -0x1.a01a01a01a01ap-13);
    p = fma(p, z, 0x1.1111111111111p-7);
    p = fma(p, z, -0x1.5555555555555p-3);
    return fma(r * z, p, r);
}

RR_HD double rr_kcos(double r)   /* |r| <= pi/4 */
{
    const double z = r * r;
    double p = -0x1.0ce396db7f853p-70;
    p = fma(p, z, 0x1.e542ba4020225p-62);
    p = fma(p, z, -0x1.6827863b97d97p-53);
    p = fma(p, z, 0x1.ae7f3e733b81fp-45);
    p = fma(p, z, -0x1.93974a8c07c9dp-37);
    p = fma(p, z, 0x1.1eed8eff8d898p-29);
    p = fma(p, z, -0x1.27e4fb7789f5cp-22);
    p = fma(p, z, 0x1.a01a01a01a01ap-16);
    p = fma(p, z, -0x1.6c16c16c16c17p-10);
    p = fma(p, z, 0x1.5555555555555p-5);
    p = fma(p, z, -0x1.0000000000000p-1);
    return fma(z, p, 1.0);
}

/* x = k*(pi/2) + r, returns r and the quadrant (k & 3) */
RR_HD double rr_rem_pio2(double x, int* quadrant)
{
    const double kd = rint(x * RR_2_OVER_PI);
    double r = fma(-kd, RR_PIO2_HI, x);
    r = fma(-kd, RR_PIO2_LO, r);
    *quadrant = ((int)kd) & 3;
    return r;
}

/* shared front end of sin/cos/tan: kernel values on the reduced argument + quadrant. rr_sin / rr_cos / rr_tan are
 * pure selections/quotients of (s, c), so evaluating the pair once and deriving several functions from it gives
 * exactly the bits of the separate calls (the frame kernel does that for sin and tan of the same angle). */
typedef struct { double s, c; int q; } rr_sincos_t;
RR_HD_CALL rr_sincos_t rr_sincos_parts(double x)
{
    rr_sincos_t r;
    if (!(fabs(x) < 1.0e5)) { r.s = x - x; r.c = x - x; r.q = 0; return r; }   /* NaN for inf/NaN/out-of-domain */
    const double red = rr_rem_pio2(x, &r.q);
    r.s = rr_ksin(red); r.c = rr_kcos(red);
    return r;
}
RR_HD double rr_sin_of(rr_sincos_t p) { return (p.q == 0) ? p.s : (p.q == 1) ? p.c : (p.q == 2) ? -p.s : -p.c; }
RR_HD double rr_cos_of(rr_sincos_t p) { return (p.q == 0) ? p.c : (p.q == 1) ? -p.s : (p.q == 2) ? -p.c : p.s; }
RR_HD double rr_tan_of(rr_sincos_t p) { return (p.q & 1) ? (-p.c / p.s) : (p.s / p.c); }

RR_HD double rr_sin(double x) { return rr_sin_of(rr_sincos_parts(x)); }
RR_HD double rr_cos(double x) { return rr_cos_of(rr_sincos_parts(x)); }
RR_HD double rr_tan(double x) { return rr_tan_of(rr_sincos_parts(x)); }

/* ------------------------------------------------------------------------------------------------
 * asin / acos (double).  asin(x) = x + x*z*P(z), z = x^2 <= 1/4, exact Maclaurin coefficients.
 * ---------------------------------------------------------------------------------------------- */
RR_HD double rr_kasin(double x)  /* |x| <= 0.5 */
{
    const double z = x * x;
    double p = 0x1.cf7dea5b6e830p-10;
    p = fma(p, z, 0x1.e82be60d9127ep-10);
    p = fma(p, z, 0x1.018f963c229bfp-9);
    p = fma(p, z, 0x1.1052bc5fa960ap-9);
    p = fma(p, z, 0x1.208d3570ae5a6p-9);
    p = fma(p, z, 0x1.3275586c5f2f0p-9);
    p = fma(p, z, 0x1.464c0950f7d47p-9);
    p = fma(p, z, 0x1.5c5f56efaaaabp-9);
    p = fma(p, z, 0x1.750de64d7d05fp-9);
    p = fma(p, z, 0x1.90cb77f60c7cep-9);
    p = fma(p, z, 0x1.b026f57b13b14p-9);
    p = fma(p, z, 0x1.d3d2a8e0dd67dp-9);
    p = fma(p, z, 0x1.fcaf8fb6db6dbp-9);
    p = fma(p, z, 0x1.15ee9d45d1746p-8);
    p = fma(p, z, 0x1.31683bdef7bdfp-8);
    p = fma(p, z, 0x1.51ba308d3dcb1p-8);
    p = fma(p, z, 0x1.782dda12f684cp-8);
    p = fma(p, z, 0x1.a6863d70a3d71p-8);
    p = fma(p, z, 0x1.df3bd37a6f4dfp-8);
    p = fma(p, z, 0x1.12ef3cf3cf3cfp-7);
    p = fma(p, z, 0x1.3fde50d79435ep-7);
    p = fma(p, z, 0x1.7a87878787878p-7);
    p = fma(p, z, 0x1.c99999999999ap-7);
    p = fma(p, z, 0x1.1c4ec4ec4ec4fp-6);
    p = fma(p, z, 0x1.6e8ba2e8ba2e9p-6);
    p = fma(p, z, 0x1.f1c71c71c71c7p-6);
    p = fma(p, z, 0x1.6db6db6db6db7p-5);
    p = fma(p, z, 0x1.3333333333333p-4);
    p = fma(p, z, 0x1.5555555555555p-3);
    return fma(x * z, p, x);
}

RR_HD_CALL double rr_asin(double x)
{
    const double ax = fabs(x);
    if (!(ax <= 1.0)) return (x - x) / (x - x); /* NaN outside [-1,1] (and for NaN) */
    if (ax <= 0.5) return rr_kasin(x);
    const double s = sqrt((1.0 - ax) * 0.5);
    const double a = rr_kasin(s);
    const double r = RR_PIO2_HI - (2.0 * a - RR_PIO2_LO);
    return (x < 0.0) ? -r : r;
}

RR_HD_CALL double rr_acos(double x)
{
    const double ax = fabs(x);
    if (!(ax <= 1.0)) return (x - x) / (x - x);
    if (ax <= 0.5) return RR_PIO2_HI - (rr_kasin(x) - RR_PIO2_LO);
    const double s = sqrt((1.0 - ax) * 0.5);
    const double a2 = 2.0 * rr_kasin(s);
    return (x > 0.0) ? a2 : (RR_PI_HI - (a2 - RR_PI_LO));
}

/* float acos as the reference gets it from `acos(float)` (C++ float overload, radar_algorithms.h:69):
 * computed in double and rounded once. NaN for |x| > 1, exactly like libm — the reference then drops
 * the wave (NaN energy fails `> threshold`), a quirk we keep. */
RR_HD float rr_acosf(float x) { return (float)rr_acos((double)x); }
RR_HD float rr_cosf(float x)  { return (float)rr_cos((double)x); }

/* ------------------------------------------------------------------------------------------------
 * exp / log (double) — enough range for results that are consumed as float
 * ---------------------------------------------------------------------------------------------- */
RR_HD_CALL double rr_exp(double x)
{
    if (x != x) return x;
    if (x > 709.0) return rr_u2d(0x7ff0000000000000ull);
    if (x < -700.0) return 0.0;            /* < 1e-304: zero for every float consumer */
    const double kd = rint(x * RR_INV_LN2);
    double r = fma(-kd, RR_LN2_HI, x);
    r = fma(-kd, RR_LN2_LO, r);
    double p = 0x1.ae7f3e733b81fp-41;      /* 1/15! */
    p = fma(p, r, 0x1.93974a8c07c9dp-37);
    p = fma(p, r, 0x1.6124613a86d09p-33);
    p = fma(p, r, 0x1.1eed8eff8d898p-29);
    p = fma(p, r, 0x1.ae64567f544e4p-26);
    p = fma(p, r, 0x1.27e4fb7789f5cp-22);
    p = fma(p, r, 0x1.71de3a556c734p-19);
    p = fma(p, r, 0x1.a01a01a01a01ap-16);
    p = fma(p, r, 0x1.a01a01a01a01ap-13);
    p = fma(p, r, 0x1.6c16c16c16c17p-10);
    p = fma(p, r, 0x1.1111111111111p-7);
    p = fma(p, r, 0x1.5555555555555p-5);
    p = fma(p, r, 0x1.5555555555555p-3);
    p = fma(p, r, 0x1.0000000000000p-1);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int k = (int)kd;                  /* |k| <= 1023 here */
    return p * rr_u2d((uint64_t)(k + 1023) << 52);
}

RR_HD_CALL double rr_log(double x)   /* x > 0, finite, normal */
{
    uint64_t u = rr_d2u(x);
    int e = (int)(u >> 52) - 1023;
    u = (u & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
    double m = rr_u2d(u);                  /* [1,2) */
    if (m > 0x1.6a09e667f3bcdp+0) { m *= 0.5; e += 1; }   /* -> [sqrt2/2, sqrt2) */
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double z = s * s;
    double p = 0x1.2f684bda12f68p-4;       /* 2/27 */
    p = fma(p, z, 0x1.47ae147ae147bp-4);
    p = fma(p, z, 0x1.642c8590b2164p-4);
    p = fma(p, z, 0x1.8618618618618p-4);
    p = fma(p, z, 0x1.af286bca1af28p-4);
    p = fma(p, z, 0x1.e1e1e1e1e1e1ep-4);
    p = fma(p, z, 0x1.1111111111111p-3);
    p = fma(p, z, 0x1.3b13b13b13b14p-3);
    p = fma(p, z, 0x1.745d1745d1746p-3);
    p = fma(p, z, 0x1.c71c71c71c71cp-3);
    p = fma(p, z, 0x1.2492492492492p-2);
    p = fma(p, z, 0x1.999999999999ap-2);
    p = fma(p, z, 0x1.5555555555555p-1);
    const double lm = fma(s * z, p, 2.0 * s);            /* log(m) */
    const double ed = (double)e;
    return fma(ed, RR_LN2_HI, fma(ed, RR_LN2_LO, lm));
}

RR_HD float rr_expf(float x) { return (float)rr_exp((double)x); }

/* powf(x, y) with C99 special cases for the inputs the BRDF produces (radar_algorithms.h:181:
 * pow(cos(angle), specular_exp), both float). */
RR_HD_CALL float rr_powf(float x, float y)
{
    if (y == 0.0f || x == 1.0f) return 1.0f;
    if (x != x || y != y) return x + y;
    const float ax = fabsf(x);
    const float yi = rintf(y);
    const int y_is_int = (yi == y) && (fabsf(y) < 1.0e9f);
    const int y_is_odd = y_is_int && ((((long long)yi) & 1ll) != 0);
    if (ax == 0.0f) {
        if (y > 0.0f) return y_is_odd ? x : 0.0f;
        return y_is_odd ? (1.0f / x) : (1.0f / ax);
    }
    if (x < 0.0f && !y_is_int) return (x - x) / (x - x);
    if (ax > 3.0e38f) { /* inf base */
        const float r = (y > 0.0f) ? ax : 0.0f;
        return (x < 0.0f && y_is_odd) ? -r : r;
    }
    double lx;
    if (ax < 1.1754944e-38f) lx = rr_log((double)ax * 0x1p+100) - 100.0 * RR_LN2_HI; /* fp32 denormal */
    else lx = rr_log((double)ax);
    const float r = (float)rr_exp((double)y * lx);
    return (x < 0.0f && y_is_odd) ? -r : r;
}

/* std::pow(float, 4.0) (RadarCPU.cpp:509): promoted to double; x*x is exact for a float input and
 * (x*x)*(x*x) is rounded once, so this equals the correctly rounded pow(x, 4.0). */
RR_HD double rr_pow4(float x) { const double d = (double)x; const double d2 = d * d; return d2 * d2; }

/* ------------------------------------------------------------------------------------------------
 * fp32 vector / quaternion algebra, restating rmagine::Vector3_<float> / Quaternion_<float>
 * (third-party, not in the reference tree). All products unfused, left-to-right.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float x, y, z; } rr_vec3;
typedef struct { float x, y, z, w; } rr_quat;

RR_HD rr_vec3 rr_v3(float x, float y, float z) { rr_vec3 v; v.x = x; v.y = y; v.z = z; return v; }
RR_HD rr_vec3 rr_add(rr_vec3 a, rr_vec3 b) { return rr_v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RR_HD rr_vec3 rr_sub(rr_vec3 a, rr_vec3 b) { return rr_v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RR_HD rr_vec3 rr_neg(rr_vec3 a) { return rr_v3(-a.x, -a.y, -a.z); }
RR_HD rr_vec3 rr_muls(rr_vec3 a, float s) { return rr_v3(a.x * s, a.y * s, a.z * s); }
RR_HD rr_vec3 rr_divs(rr_vec3 a, float s) { return rr_v3(a.x / s, a.y / s, a.z / s); }
RR_HD float   rr_dot(rr_vec3 a, rr_vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RR_HD rr_vec3 rr_cross(rr_vec3 a, rr_vec3 b)
{
    return rr_v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
RR_HD float   rr_l2norm(rr_vec3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
RR_HD rr_vec3 rr_normalize(rr_vec3 a) { return rr_divs(a, rr_l2norm(a)); }

RR_HD rr_quat rr_qmul(rr_quat a, rr_quat b)
{
    rr_quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return r;
}
RR_HD rr_quat rr_qinv(rr_quat q) { rr_quat r; r.x = -q.x; r.y = -q.y; r.z = -q.z; r.w = q.w; return r; }
/* q * v = (q (x) (v,0) (x) q^-1).xyz */
RR_HD rr_vec3 rr_qrot(rr_quat q, rr_vec3 v)
{
    rr_quat p; p.x = v.x; p.y = v.y; p.z = v.z; p.w = 0.0f;
    const rr_quat t = rr_qmul(rr_qmul(q, p), rr_qinv(q));
    return rr_v3(t.x, t.y, t.z);
}

/* ------------------------------------------------------------------------------------------------
 * Ray/triangle test — OUR definition of the closest-hit primitive that Embree provides to the
 * reference (call site RadarCPU.cpp:236; Embree itself is not reproducible bit-for-bit).
 * fp32 Moeller-Trumbore on (v0, e1 = v1 - v0, e2 = v2 - v0), explicit FMAs in a fixed order,
 * two-sided, det == 0 never hits. Returns 1 and *t_out when 0 <= t <= tmax.
 * Closest-hit rule everywhere: smallest t, ties -> lowest face id.
 * ---------------------------------------------------------------------------------------------- */
RR_HD float rr_fdot(rr_vec3 a, rr_vec3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
RR_HD rr_vec3 rr_fcross(rr_vec3 a, rr_vec3 b)
{
    return rr_v3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
RR_HD int rr_ray_triangle(rr_vec3 o, rr_vec3 d, rr_vec3 v0, rr_vec3 e1, rr_vec3 e2, float tmax, float* t_out)
{
    const rr_vec3 p = rr_fcross(d, e2);
    const float det = rr_fdot(e1, p);
    if (!(det != 0.0f)) return 0;
    const float inv = 1.0f / det;
    const rr_vec3 tv = rr_sub(o, v0);
    const float u = rr_fdot(tv, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return 0;
    const rr_vec3 q = rr_fcross(tv, e1);
    const float v = rr_fdot(d, q) * inv;
    if (!(v >= 0.0f && (u + v) <= 1.0f)) return 0;
    const float t = rr_fdot(e2, q) * inv;
    if (!(t >= 0.0f && t <= tmax)) return 0;
    *t_out = t;
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11), counter-based RNG. The reference seeds std::mt19937 from
 * std::random_device (radar_algorithms.cpp:258-259, RadarCPU.cpp:461-462) and is not reproducible;
 * north_star asks for a counter-based generator keyed by (pose, azimuth, sample).
 * ---------------------------------------------------------------------------------------------- */
RR_HD void rr_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* uniform float in [0,1) with 24 random bits (what std::uniform_real_distribution<float>(0,1) spans) */
RR_HD float rr_u01(uint32_t bits) { return (float)(bits >> 8) * 0x1p-24f; }

/* Noise stream of one azimuth column: draw #0 is `random_begin`, draw #(1+i) is cell i's uniform p
 * (the order RadarCPU.cpp:472,482 consumes its generator in). */
RR_HD float rr_noise_u01(uint64_t seed, uint64_t frame_id, uint32_t azimuth, uint32_t draw)
{
    uint32_t ctr[4], key[2], out[4];
    ctr[0] = draw; ctr[1] = azimuth; ctr[2] = (uint32_t)frame_id; ctr[3] = (uint32_t)(frame_id >> 32);
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    rr_philox4x32_10(ctr, key, out);
    return rr_u01(out[0]);
}

#endif /* RR_DETMATH_H */
