// fs_math.cuh -- arithmetic shared by every kernel on the BDPT path.
//
// Bit-exactness contract (histograms must equal the CPU oracle's bit for bit): this file uses
// only IEEE-754 binary32 + - * / sqrt, explicit fmaf(), rintf(), float<->int conversions and
// integer bit operations.  The translation units that include it are compiled with
// --fmad=false, so the compiler never fuses or splits a multiply-add on its own; every FFMA in
// the SASS comes from an fmaf() written here.  No libm/libdevice transcendental is called:
// exp/log/pow/sincos are polynomial, built from fmaf.
//
// Reference mapping (SUB.cpp = Plugins/FrequenSee/Source/FrequenSee/Private/
// AudioRayTracingSubsystem.cpp): Philox replaces FMath::FRand/VRand/VRandCone (SUB.cpp:301,
// 308, 313); fs_exp replaces exp() at SUB.cpp:396, fs_pow replaces powf() at SUB.cpp:398;
// fs_intersect_tri replaces UWorld::LineTraceSingleByObjectType (SUB.cpp:252, 340).
#pragma once

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define FS_HD __host__ __device__ __forceinline__
#else
#define FS_HD inline
#endif

#define FS_MAX_BANDS 8

#define FS_PI       3.14159274101257324f
#define FS_HALF_PI  1.57079637050628662f
#define FS_INV_PI   0.318309873342514038f
#define FS_FOUR_PI  12.5663709640502930f
#define FS_INV_4PI  0.0795774683356285095f
#define FS_LOG2E    1.44269502162933350f
#define FS_LN2_HI   0.693145751953125f
#define FS_LN2_LO   1.42860676533018518e-06f
#define FS_LN2      0.693147182464599609f

struct fs_vec3 { float x, y, z; };

FS_HD uint32_t fs_f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
FS_HD float fs_u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

FS_HD uint32_t fs_mulhi(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// Philox4x32-10, counter = (g_lo, g_hi, bounce, side), key = seed
FS_HD void fs_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                            uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = fs_mulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = fs_mulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

FS_HD float fs_u01(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }

FS_HD float fs_exp(float x)
{
    if (!(x >= -87.0f)) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    float k = rintf(x * FS_LOG2E);
    float r = fmaf(k, -FS_LN2_HI, x);
    r = fmaf(k, -FS_LN2_LO, r);
    float p = 1.98412701138295233e-04f;
    p = fmaf(p, r, 1.38888892251998186e-03f);
    p = fmaf(p, r, 8.33333376795053482e-03f);
    p = fmaf(p, r, 4.16666679084300995e-02f);
    p = fmaf(p, r, 1.66666671633720398e-01f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    int ki = (int)k;
    return p * fs_u2f((uint32_t)(ki + 127) << 23);
}

FS_HD float fs_log(float x)
{
    uint32_t ix = fs_f2u(x);
    ix += 0x3f800000u - 0x3f3504f3u;
    int k = (int)(ix >> 23) - 127;
    ix = (ix & 0x007fffffu) + 0x3f3504f3u;
    float m = fs_u2f(ix);
    float f = m - 1.0f;
    float s = f / (2.0f + f);
    float z = s * s;
    float p = 1.11111111938953400e-01f;
    p = fmaf(p, z, 1.42857149243354797e-01f);
    p = fmaf(p, z, 2.00000002980232239e-01f);
    p = fmaf(p, z, 3.33333343267440796e-01f);
    p = fmaf(p, z, 1.0f);
    float lm = (2.0f * s) * p;
    return fmaf((float)k, FS_LN2, lm);
}

FS_HD float fs_pow(float x, float e) { return fs_exp(e * fs_log(x)); }

FS_HD void fs_sincos_2pi(float u, float& c, float& s)
{
    float a = u * 4.0f;
    int q = (int)a;
    float f = a - (float)q;
    float x = f * FS_HALF_PI;
    float z = x * x;
    float ps = -2.50521079437465232e-08f;
    ps = fmaf(ps, z, 2.75573188446287531e-06f);
    ps = fmaf(ps, z, -1.98412701138295233e-04f);
    ps = fmaf(ps, z, 8.33333376795053482e-03f);
    ps = fmaf(ps, z, -1.66666671633720398e-01f);
    ps = fmaf(ps, z, 1.0f);
    float sn = x * ps;
    float pc = 2.08767569864244458e-09f;
    pc = fmaf(pc, z, -2.75573199814971304e-07f);
    pc = fmaf(pc, z, 2.48015876422869042e-05f);
    pc = fmaf(pc, z, -1.38888892251998186e-03f);
    pc = fmaf(pc, z, 4.16666679084300995e-02f);
    pc = fmaf(pc, z, -0.5f);
    float cs = fmaf(pc, z, 1.0f);
    q &= 3;
    c = (q == 0) ? cs : (q == 1) ? -sn : (q == 2) ? -cs : sn;
    s = (q == 0) ? sn : (q == 1) ? cs : (q == 2) ? -sn : -cs;
}

FS_HD fs_vec3 fs_sample_sphere(float u1, float u2)
{
    float c, s;
    fs_sincos_2pi(u2, c, s);
    float z = fmaf(-2.0f, u1, 1.0f);
    float rr = fmaf(-z, z, 1.0f);
    if (!(rr > 0.0f)) rr = 0.0f;
    float r = sqrtf(rr);
    fs_vec3 d;
    d.x = r * c; d.y = r * s; d.z = z;
    return d;
}

// cosine-weighted hemisphere about unit normal n; cos_theta = local z
FS_HD fs_vec3 fs_sample_cos_hemisphere(fs_vec3 n, float u1, float u2, float& cos_theta)
{
    float c, s;
    fs_sincos_2pi(u2, c, s);
    float r = sqrtf(u1);
    float zl = sqrtf(1.0f - u1);
    float lx = r * c, ly = r * s;
    float sign = (n.z >= 0.0f) ? 1.0f : -1.0f;
    float a = -1.0f / (sign + n.z);
    float b = (n.x * n.y) * a;
    float tx = fmaf(sign * (n.x * n.x), a, 1.0f);
    float ty = sign * b;
    float tz = -sign * n.x;
    float bx = b;
    float by = fmaf(n.y * n.y, a, sign);
    float bz = -n.y;
    fs_vec3 d;
    d.x = fmaf(lx, tx, fmaf(ly, bx, zl * n.x));
    d.y = fmaf(lx, ty, fmaf(ly, by, zl * n.y));
    d.z = fmaf(lx, tz, fmaf(ly, bz, zl * n.z));
    cos_theta = zl;
    return d;
}

FS_HD float fs_dot(fs_vec3 a, fs_vec3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
FS_HD fs_vec3 fs_cross(fs_vec3 a, fs_vec3 b)
{
    fs_vec3 o;
    o.x = fmaf(a.y, b.z, -(a.z * b.y));
    o.y = fmaf(a.z, b.x, -(a.x * b.z));
    o.z = fmaf(a.x, b.y, -(a.y * b.x));
    return o;
}
FS_HD fs_vec3 fs_sub(fs_vec3 a, fs_vec3 b) { fs_vec3 o; o.x = a.x - b.x; o.y = a.y - b.y; o.z = a.z - b.z; return o; }
FS_HD fs_vec3 fs_mk(float x, float y, float z) { fs_vec3 o; o.x = x; o.y = y; o.z = z; return o; }

// Moller-Trumbore, two-sided, t > 0.  The (t, triangle id) of every hit is a pure function of
// (ray, triangle), so the closest hit is independent of the acceleration structure.
FS_HD bool fs_intersect_tri(fs_vec3 o, fs_vec3 d, fs_vec3 v0, fs_vec3 e1, fs_vec3 e2, float& t_out)
{
    fs_vec3 pv = fs_cross(d, e2);
    float det = fs_dot(e1, pv);
    float inv = 1.0f / det;
    fs_vec3 tv = fs_sub(o, v0);
    float u = fs_dot(tv, pv) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    fs_vec3 qv = fs_cross(tv, e1);
    float v = fs_dot(d, qv) * inv;
    if (!(v >= 0.0f && (u + v) <= 1.0f)) return false;
    float t = fs_dot(e2, qv) * inv;
    if (!(t > 0.0f)) return false;
    t_out = t;
    return true;
}

// unit geometric normal of a triangle given its edges (same sequence as the oracle)
FS_HD fs_vec3 fs_tri_normal(fs_vec3 e1, fs_vec3 e2)
{
    fs_vec3 c = fs_cross(e1, e2);
    float len = sqrtf(fs_dot(c, c));
    float inv = 1.0f / len;
    return fs_mk(c.x * inv, c.y * inv, c.z * inv);
}

// One segment of EvaluatePath (SUB.cpp:368-399) over all bands.
//   total += d;  if d < min_seg skip;  E_b *= bsdf_b; E_b *= 1/(4 pi d^2); E_b *= exp(-air_b d);
//   E_b /= prob^pdf_exponent
struct fs_eval_params {
    float min_seg, pdf_exponent;
    uint32_t n_bands;
    float air[FS_MAX_BANDS];
};

template <int NB>
FS_HD void fs_eval_segment(const fs_eval_params& ep, const float* refl_over_pi_row /* null: BSDF=1 */,
                           float prob, float d, float& total, float E[NB])
{
    total += d;
    if (d < ep.min_seg) return;
    float G = 1.0f / (FS_FOUR_PI * (d * d));
    float P = fs_pow(prob, ep.pdf_exponent);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (b < (int)ep.n_bands) {
            float bs = refl_over_pi_row ? refl_over_pi_row[b] : 1.0f;
            float e = E[b];
            e *= bs;
            e *= G;
            e *= fs_exp(-ep.air[b] * d);
            e /= P;
            E[b] = e;
        }
    }
}
