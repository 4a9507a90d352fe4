// ccu_tonemap.cuh - the post-processing filters of the reference's tonemap kernel
// (src/main/opencl/tonemap/include/post_processing_filter.cl:5-51, double.h:17-19, rgba.h:6-16).
//
// Arithmetic contract as for the render path: every fp32 operation the reference writes out is one IEEE operation in
// source order (no contraction); pow() is a fixed kernel ("detmath"): log2 / exp2 series evaluated in fp64 from + - * /
// only and rounded once to fp32, so CPU oracle and GPU agree bit for bit and the result is within 1 ulp of a correctly
// rounded powf (the OpenCL builtin is allowed 16 ulp).
#pragma once
#include "ccu_math.cuh"

namespace ccu {

// x^y for the filter's use (y = 1/2.2, positive and not an integer), C99 / OpenCL special cases for such y:
// pow(+-0, y) = 0, pow(finite x < 0, y) = NaN, pow(+-inf, y) = +inf
__device__ __forceinline__ float dm_powf(float x, float y) {
    if (x != x || y != y) return nanf_();
    if (x == 0.0f) return 0.0f;
    if (x == inff_() || x == -inff_()) return inff_();
    if (x < 0.0f) return nanf_();
    // x = m * 2^e with m in [sqrt(1/2), sqrt(2)), exactly, through the fp64 representation
    double dx = (double)x;
    long long bits = __double_as_longlong(dx);
    int e = (int)((bits >> 52) & 0x7FF) - 1023;
    double m = __longlong_as_double((bits & 0x000FFFFFFFFFFFFFLL) | 0x3FF0000000000000LL);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    // log2(m) = 2/ln2 * atanh(t), t = (m-1)/(m+1), |t| <= 0.1716
    const double t = (m - 1.0) / (m + 1.0);
    const double t2 = t * t;
    double s = 1.0 / 15.0;
    s = s * t2 + 1.0 / 13.0;
    s = s * t2 + 1.0 / 11.0;
    s = s * t2 + 1.0 / 9.0;
    s = s * t2 + 1.0 / 7.0;
    s = s * t2 + 1.0 / 5.0;
    s = s * t2 + 1.0 / 3.0;
    s = s * t2 + 1.0;
    const double l2 = (double)e + (t * s) * 2.8853900817779268;   // 2 / ln 2
    const double p = (double)y * l2;
    if (p > 200.0) return inff_();
    if (p < -200.0) return 0.0f;
    const double k = floor(p + 0.5);
    const double f = (p - k) * 0.6931471805599453;                // ln 2, |f| <= 0.3466
    double r = 1.0 / 479001600.0;                                 // exp(f), Taylor to f^12
    r = r * f + 1.0 / 39916800.0;
    r = r * f + 1.0 / 3628800.0;
    r = r * f + 1.0 / 362880.0;
    r = r * f + 1.0 / 40320.0;
    r = r * f + 1.0 / 5040.0;
    r = r * f + 1.0 / 720.0;
    r = r * f + 1.0 / 120.0;
    r = r * f + 1.0 / 24.0;
    r = r * f + 1.0 / 6.0;
    r = r * f + 0.5;
    r = r * f + 1.0;
    r = r * f + 1.0;
    const double scale = __longlong_as_double((long long)((int)k + 1023) << 52);   // 2^k, |k| <= 200
    return (float)(r * scale);
}

// cvt.rzi.u32.f32: toward zero, saturating, NaN -> 0 (what the GPU the reference runs on does for (uint)float)
__device__ __forceinline__ unsigned f2u(float f) { return __float2uint_rz(f); }

// post_processing_filter.cl:14-50 for one pixel; in[] = the three doubles of the pixel
__device__ __forceinline__ unsigned tonemap_pixel(const double *__restrict__ in, float exposure, int type) {
    float c[3];
    for (int i = 0; i < 3; i++) c[i] = (float)in[i] * exposure;             // double.h:19, :22
    const float inv_gamma = (float)(1.0 / 2.2);                             // the double literal of :27,:38 as the float argument of pow
    for (int i = 0; i < 3; i++) {
        float x = c[i];
        switch (type) {
            case 0:   // GAMMA :25-28
                x = dm_powf(x, inv_gamma);
                break;
            case 1:   // TONEMAP1 :29-33
                x = fmaxf(0.0f, x - 0.004f);
                x = (x * (6.2f * x + 0.5f)) / (x * (6.2f * x + 1.7f) + 0.06f);
                break;
            case 2:   // ACES :34-39
                x = (x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f);
                x = fminf(fmaxf(x, 0.0f), 1.0f);
                x = dm_powf(x, inv_gamma);
                break;
            case 3: { // HABLE :40-45
                x = x * 16.0f;
                const float a = 0.10f * 0.50f, b = 0.20f * 0.02f, d = 0.20f * 0.30f, g = 0.02f / 0.30f;
                x = ((x * (0.15f * x + a) + b) / (x * (0.15f * x + 0.50f) + d)) - g;
                const float w = ((11.2f * (0.15f * 11.2f + a) + b) / (11.2f * (0.15f * 11.2f + 0.50f) + d)) - g;
                x = x / w;
                break;
            }
            default:
                break;
        }
        c[i] = x;
    }
    // rgba.h:6-16 (alpha = 1)
    unsigned r = min(f2u(c[0] * 255.0f + 0.5f), 255u);
    unsigned g = min(f2u(c[1] * 255.0f + 0.5f), 255u);
    unsigned b = min(f2u(c[2] * 255.0f + 0.5f), 255u);
    return (255u << 24) | (r << 16) | (g << 8) | b;
}

}  // namespace ccu
