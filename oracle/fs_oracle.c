/*
 * fs_oracle.c -- CPU oracle (plain C99).  See fs_oracle.h: TEST INFRASTRUCTURE ONLY; pinned by the reference's own
 * function bodies compiled into oracle/_ref (tests/test_oracle_ref_ue.py) and by analytic KATs.
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off -fno-fast-math [-mfma] -pthread -fPIC -shared
 *
 * Arithmetic contract (what makes the integer histogram bit-exact against the CUDA path):
 * only IEEE-754 binary32 + - * / sqrt, fmaf, rintf, float<->int conversions and integer bit
 * operations are used on the histogram path; no libm transcendental.  Contraction is off,
 * every fused multiply-add is an explicit fmaf().  The CUDA side (compiled --fmad=false)
 * executes the same operation sequence.
 */
#include "fs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------------------------------
 * constants (float literals; the same bit patterns appear in the CUDA math header)
 * ---------------------------------------------------------------------------------------- */
#define FSO_PI        3.14159274101257324f   /* (float)pi, 0x40490fdb */
#define FSO_HALF_PI   1.57079637050628662f   /* 0x3fc90fdb */
#define FSO_INV_PI    0.318309873342514038f  /* 0x3ea2f983 */
#define FSO_FOUR_PI   12.5663709640502930f   /* 0x41490fdb */
#define FSO_INV_4PI   0.0795774683356285095f /* 0x3da2f983 */
#define FSO_LOG2E     1.44269502162933350f   /* 0x3fb8aa3b */
#define FSO_LN2_HI    0.693145751953125f     /* 0x3f317200 */
#define FSO_LN2_LO    1.42860676533018518e-06f /* 0x35bfbe8e */
#define FSO_LN2       0.693147182464599609f  /* 0x3f317218 */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

void fso_default_config(fso_config* c)
{
    memset(c, 0, sizeof(*c));
    c->n_bands = 8;
    c->n_bins = 1000;
    c->bin_ms = 1.0f;
    c->rr_prob = 0.9f;
    c->eps_offset = 1e-3f;
    c->eps_connect = 1e-3f;
    c->min_seg = 1e-2f;
    c->sound_speed = 343.0f;
    c->pdf_exponent = 0.1f;
    c->energy_clamp = 1.0f;
    c->energy_gain = 10.0f;
    /* air absorption per metre, 8 octave bands 63 Hz .. 8 kHz (fixed table; the reference has a
     * single 0.05 per 10 m unit = 0.005 /m, SUB.cpp:395, and 0.0017 /m in COMP.cpp:264) */
    static const float air[FSO_MAX_BANDS] = {0.0001f, 0.0003f, 0.0006f, 0.0010f,
                                             0.0017f, 0.0035f, 0.0050f, 0.0120f};
    for (int b = 0; b < FSO_MAX_BANDS; ++b) c->air_absorption[b] = air[b];
    c->sample_rate = 48000;
    c->n_channels = 2;
    c->ir_threshold = 1e-6f;
    c->ir_lowpass = 0.25f;
    c->conv_block = 1024;
    c->conv_clamp = 1;
    c->conv_wet = 1.0f;
}

/* ------------------------------------------------------------------------------------------
 * RNG: Philox4x32-10 (Salmon et al., SC'11; Random123).  Replaces FMath::FRand / VRand /
 * VRandCone (SUB.cpp:301, 308, 313), which are libc rand() based and unseeded.
 * ---------------------------------------------------------------------------------------- */
void fso_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 24-bit uniform in [0,1): exact */
float fso_u01(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }

/* ------------------------------------------------------------------------------------------
 * exp / log / pow built from fmaf only.  Stand in for exp() (SUB.cpp:396) and powf()
 * (SUB.cpp:398), whose libm results differ between glibc and CUDA in the last ulp.
 * ---------------------------------------------------------------------------------------- */
float fso_expf(float x)
{
    if (!(x >= -87.0f)) return 0.0f;       /* also catches NaN */
    if (x > 88.0f) x = 88.0f;
    float k = rintf(x * FSO_LOG2E);
    float r = fmaf(k, -FSO_LN2_HI, x);
    r = fmaf(k, -FSO_LN2_LO, r);
    /* Taylor degree 7 on |r| <= ln2/2 */
    float p = 1.98412701138295233e-04f;          /* 1/5040 */
    p = fmaf(p, r, 1.38888892251998186e-03f);    /* 1/720 */
    p = fmaf(p, r, 8.33333376795053482e-03f);    /* 1/120 */
    p = fmaf(p, r, 4.16666679084300995e-02f);    /* 1/24 */
    p = fmaf(p, r, 1.66666671633720398e-01f);    /* 1/6 */
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    int ki = (int)k;                             /* in [-126, 127] */
    return p * u2f((uint32_t)(ki + 127) << 23);
}

float fso_logf(float x)
{
    /* x > 0, normal.  x = 2^k * m, m in [sqrt(1/2), sqrt(2)) */
    uint32_t ix = f2u(x);
    ix += 0x3f800000u - 0x3f3504f3u;
    int k = (int)(ix >> 23) - 127;
    ix = (ix & 0x007fffffu) + 0x3f3504f3u;
    float m = u2f(ix);
    float f = m - 1.0f;
    float s = f / (2.0f + f);
    float z = s * s;
    /* log(1+f) = 2s (1 + z/3 + z^2/5 + z^3/7 + z^4/9) */
    float p = 1.11111111938953400e-01f;          /* 1/9 */
    p = fmaf(p, z, 1.42857149243354797e-01f);    /* 1/7 */
    p = fmaf(p, z, 2.00000002980232239e-01f);    /* 1/5 */
    p = fmaf(p, z, 3.33333343267440796e-01f);    /* 1/3 */
    p = fmaf(p, z, 1.0f);
    float lm = (2.0f * s) * p;
    return fmaf((float)k, FSO_LN2, lm);
}

float fso_powf(float x, float e) { return fso_expf(e * fso_logf(x)); }

/* cos/sin of 2*pi*u, u in [0,1): exact quadrant reduction + Taylor on [0, pi/2) */
void fso_sincos_2pi(float u, float* c, float* s)
{
    float a = u * 4.0f;
    int   q = (int)a;
    float f = a - (float)q;
    float x = f * FSO_HALF_PI;
    float z = x * x;
    float ps = -2.50521079437465232e-08f;        /* -1/39916800 */
    ps = fmaf(ps, z, 2.75573188446287531e-06f);  /*  1/362880 */
    ps = fmaf(ps, z, -1.98412701138295233e-04f); /* -1/5040 */
    ps = fmaf(ps, z, 8.33333376795053482e-03f);  /*  1/120 */
    ps = fmaf(ps, z, -1.66666671633720398e-01f); /* -1/6 */
    ps = fmaf(ps, z, 1.0f);
    float sn = x * ps;
    float pc = 2.08767569864244458e-09f;         /*  1/479001600 */
    pc = fmaf(pc, z, -2.75573199814971304e-07f); /* -1/3628800 */
    pc = fmaf(pc, z, 2.48015876422869042e-05f);  /*  1/40320 */
    pc = fmaf(pc, z, -1.38888892251998186e-03f); /* -1/720 */
    pc = fmaf(pc, z, 4.16666679084300995e-02f);  /*  1/24 */
    pc = fmaf(pc, z, -0.5f);
    float cs = fmaf(pc, z, 1.0f);
    switch (q & 3) {
    case 0:  *c = cs;  *s = sn;  break;
    case 1:  *c = -sn; *s = cs;  break;
    case 2:  *c = -cs; *s = -sn; break;
    default: *c = sn;  *s = -cs; break;
    }
}

/* Uniform direction on the sphere (stands in for FMath::VRand at node 0, SUB.cpp:308) */
void fso_sample_sphere(float u1, float u2, float dir[3])
{
    float c, s;
    fso_sincos_2pi(u2, &c, &s);
    float z = fmaf(-2.0f, u1, 1.0f);
    float rr = fmaf(-z, z, 1.0f);
    if (!(rr > 0.0f)) rr = 0.0f;
    float r = sqrtf(rr);
    dir[0] = r * c;
    dir[1] = r * s;
    dir[2] = z;
}

/* Cosine-weighted hemisphere about unit normal n (FIX of VRandCone(n, 90deg), SUB.cpp:313,
 * whose claimed pdf cos/pi at SUB.cpp:315-317 it does not actually sample).  Frame: Duff et
 * al. 2017 branchless ONB. */
void fso_sample_cos_hemisphere(const float n[3], float u1, float u2, float dir[3], float* cos_theta)
{
    float c, s;
    fso_sincos_2pi(u2, &c, &s);
    float r  = sqrtf(u1);
    float zl = sqrtf(1.0f - u1);
    float lx = r * c, ly = r * s;
    float sign = (n[2] >= 0.0f) ? 1.0f : -1.0f;
    float a = -1.0f / (sign + n[2]);
    float b = (n[0] * n[1]) * a;
    float tx = fmaf(sign * (n[0] * n[0]), a, 1.0f);
    float ty = sign * b;
    float tz = -sign * n[0];
    float bx = b;
    float by = fmaf(n[1] * n[1], a, sign);
    float bz = -n[1];
    dir[0] = fmaf(lx, tx, fmaf(ly, bx, zl * n[0]));
    dir[1] = fmaf(lx, ty, fmaf(ly, by, zl * n[1]));
    dir[2] = fmaf(lx, tz, fmaf(ly, bz, zl * n[2]));
    *cos_theta = zl;
}

/* ------------------------------------------------------------------------------------------
 * geometry: replaces UWorld::LineTraceSingleByObjectType (SUB.cpp:252, 340) with closest hit
 * on triangles.  Closest = lexicographic min of (t, triangle id) so the answer does not
 * depend on the acceleration structure.
 * ---------------------------------------------------------------------------------------- */
static inline float dot3(const float a[3], const float b[3])
{
    return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0]));
}
static inline void cross3(const float a[3], const float b[3], float o[3])
{
    o[0] = fmaf(a[1], b[2], -(a[2] * b[1]));
    o[1] = fmaf(a[2], b[0], -(a[0] * b[2]));
    o[2] = fmaf(a[0], b[1], -(a[1] * b[0]));
}

int fso_intersect_tri(const float o[3], const float d[3], const float v0[3], const float e1[3],
                      const float e2[3], float* t_out)
{
    float pv[3], tv[3], qv[3];
    cross3(d, e2, pv);
    float det = dot3(e1, pv);
    float inv = 1.0f / det;
    tv[0] = o[0] - v0[0]; tv[1] = o[1] - v0[1]; tv[2] = o[2] - v0[2];
    float u = dot3(tv, pv) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return 0;
    cross3(tv, e1, qv);
    float v = dot3(d, qv) * inv;
    if (!(v >= 0.0f && (u + v) <= 1.0f)) return 0;
    float t = dot3(e2, qv) * inv;
    if (!(t > 0.0f)) return 0;
    *t_out = t;
    return 1;
}

typedef struct { float lo[3], hi[3]; int32_t left, right; uint32_t first, count; } onode;

struct fso_scene {
    uint64_t n_tris;
    uint32_t n_mats, n_bands;
    float*    v0;     /* [T][3] */
    float*    e1;     /* [T][3] */
    float*    e2;     /* [T][3] */
    float*    nrm;    /* [T][3] unit geometric normal */
    uint32_t* mat;    /* [T] */
    float*    refl_over_pi; /* [M][B]  (1-alpha)/pi */
    float*    bsdf_tab;     /* FSO_FLAG_MATERIAL_MODEL: [3][M][B] diffuse / specular / transmitted factor per event */
    float*    lobes;        /* [M][4] (t1, t2, P_spec, P_diff): event thresholds on u3 and event probabilities */
    float*    absorption;   /* [M][B] copy of the input */
    int       use_bvh;
    onode*    nodes;
    uint32_t  n_nodes;
    uint32_t* order;  /* BVH leaf order -> tri id */
    fso_stats counters;
};

typedef struct { uint64_t node_visits, tri_tests; } ocount;

/* --- oracle BVH: median split on the longest centroid axis, <=4 tris per leaf, boxes padded */
typedef struct { fso_scene* sc; float* cen; float pad; } obuild;

static void tri_bounds(const fso_scene* sc, uint32_t t, float lo[3], float hi[3])
{
    for (int a = 0; a < 3; ++a) {
        float p0 = sc->v0[3 * t + a];
        float p1 = p0 + sc->e1[3 * t + a];
        float p2 = p0 + sc->e2[3 * t + a];
        float mn = p0 < p1 ? p0 : p1; mn = mn < p2 ? mn : p2;
        float mx = p0 > p1 ? p0 : p1; mx = mx > p2 ? mx : p2;
        /* e1/e2 were rounded when formed, so v0+e1 may differ from v1 by an ulp: the padding
         * applied by the caller covers it */
        lo[a] = mn; hi[a] = mx;
    }
}

static void select_nth(uint32_t* idx, const float* cen, int axis, int64_t lo, int64_t hi, int64_t nth)
{
    /* quickselect on idx[lo..hi] by cen[3*id+axis], ties by id (deterministic) */
    while (lo < hi) {
        uint32_t pid = idx[lo + (hi - lo) / 2];
        float pv = cen[3 * (uint64_t)pid + axis];
        int64_t i = lo, j = hi;
        while (i <= j) {
            for (;;) {
                uint32_t a = idx[i]; float av = cen[3 * (uint64_t)a + axis];
                if (av < pv || (av == pv && a < pid)) ++i; else break;
            }
            for (;;) {
                uint32_t a = idx[j]; float av = cen[3 * (uint64_t)a + axis];
                if (av > pv || (av == pv && a > pid)) --j; else break;
            }
            if (i <= j) { uint32_t tmp = idx[i]; idx[i] = idx[j]; idx[j] = tmp; ++i; --j; }
        }
        if (nth <= j) hi = j; else if (nth >= i) lo = i; else return;
    }
}

static int32_t build_rec(obuild* b, int64_t lo, int64_t hi /* inclusive */)
{
    fso_scene* sc = b->sc;
    int32_t me = (int32_t)sc->n_nodes++;
    onode* n = &sc->nodes[me];
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int a = 0; a < 3; ++a) { n->lo[a] = INFINITY; n->hi[a] = -INFINITY; }
    for (int64_t i = lo; i <= hi; ++i) {
        uint32_t t = sc->order[i];
        float tl[3], th[3];
        tri_bounds(sc, t, tl, th);
        for (int a = 0; a < 3; ++a) {
            if (tl[a] < n->lo[a]) n->lo[a] = tl[a];
            if (th[a] > n->hi[a]) n->hi[a] = th[a];
            float c = b->cen[3 * (uint64_t)t + a];
            if (c < clo[a]) clo[a] = c;
            if (c > chi[a]) chi[a] = c;
        }
    }
    for (int a = 0; a < 3; ++a) { n->lo[a] -= b->pad; n->hi[a] += b->pad; }
    int64_t cnt = hi - lo + 1;
    if (cnt <= 4) { n->left = n->right = -1; n->first = (uint32_t)lo; n->count = (uint32_t)cnt; return me; }
    int axis = 0;
    float ext = chi[0] - clo[0];
    if (chi[1] - clo[1] > ext) { axis = 1; ext = chi[1] - clo[1]; }
    if (chi[2] - clo[2] > ext) { axis = 2; }
    int64_t mid = lo + cnt / 2;
    select_nth(sc->order, b->cen, axis, lo, hi, mid);
    n->first = 0; n->count = 0;
    int32_t l = build_rec(b, lo, mid - 1);
    int32_t r = build_rec(b, mid, hi);
    n = &sc->nodes[me];
    n->left = l; n->right = r;
    return me;
}

fso_scene* fso_scene_create(const float* verts, const uint32_t* tri_mat, uint64_t n_tris,
                            const float* absorption, uint32_t n_mats, uint32_t n_bands, int use_bvh)
{
    if (n_bands == 0 || n_bands > FSO_MAX_BANDS) return NULL;
    fso_scene* sc = (fso_scene*)calloc(1, sizeof(*sc));
    sc->n_tris = n_tris; sc->n_mats = n_mats; sc->n_bands = n_bands; sc->use_bvh = use_bvh;
    size_t T = (size_t)(n_tris ? n_tris : 1);
    sc->v0 = (float*)malloc(T * 12); sc->e1 = (float*)malloc(T * 12);
    sc->e2 = (float*)malloc(T * 12); sc->nrm = (float*)malloc(T * 12);
    sc->mat = (uint32_t*)malloc(T * 4);
    float smin[3] = {INFINITY, INFINITY, INFINITY}, smax[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint64_t t = 0; t < n_tris; ++t) {
        const float* p = verts + 9 * t;
        float e1[3], e2[3], c[3];
        for (int a = 0; a < 3; ++a) {
            sc->v0[3 * t + a] = p[a];
            e1[a] = p[3 + a] - p[a];
            e2[a] = p[6 + a] - p[a];
            sc->e1[3 * t + a] = e1[a];
            sc->e2[3 * t + a] = e2[a];
            for (int k = 0; k < 3; ++k) {
                float q = p[3 * k + a];
                if (q < smin[a]) smin[a] = q;
                if (q > smax[a]) smax[a] = q;
            }
        }
        cross3(e1, e2, c);
        float len = sqrtf(dot3(c, c));
        float inv = 1.0f / len;
        sc->nrm[3 * t + 0] = c[0] * inv;
        sc->nrm[3 * t + 1] = c[1] * inv;
        sc->nrm[3 * t + 2] = c[2] * inv;
        sc->mat[t] = tri_mat ? tri_mat[t] : 0;
    }
    sc->refl_over_pi = (float*)malloc((size_t)(n_mats ? n_mats : 1) * n_bands * 4);
    sc->absorption = (float*)malloc((size_t)(n_mats ? n_mats : 1) * n_bands * 4);
    memcpy(sc->absorption, absorption, (size_t)n_mats * n_bands * 4);
    for (uint32_t m = 0; m < n_mats; ++m)
        for (uint32_t b = 0; b < n_bands; ++b)
            /* reference: Absorption[2].Value / PI used as reflectivity (SUB.cpp:381-386);
             * FIX+EXTEND: rho_b = 1 - alpha_b per band */
            sc->refl_over_pi[m * n_bands + b] = (1.0f - absorption[m * n_bands + b]) / FSO_PI;
    if (use_bvh && n_tris > 0) {
        sc->order = (uint32_t*)malloc(T * 4);
        float* cen = (float*)malloc(T * 12);
        for (uint64_t t = 0; t < n_tris; ++t) {
            sc->order[t] = (uint32_t)t;
            float lo[3], hi[3];
            tri_bounds(sc, (uint32_t)t, lo, hi);
            for (int a = 0; a < 3; ++a) cen[3 * t + a] = 0.5f * lo[a] + 0.5f * hi[a];
        }
        float ext = 0.0f;
        for (int a = 0; a < 3; ++a) {
            float m1 = fabsf(smin[a]), m2 = fabsf(smax[a]);
            if (m1 > ext) ext = m1;
            if (m2 > ext) ext = m2;
        }
        obuild b;
        b.sc = sc; b.cen = cen;
        b.pad = ext * (1.0f / 8192.0f);
        if (b.pad < 1e-4f) b.pad = 1e-4f;
        sc->nodes = (onode*)malloc(sizeof(onode) * (2 * T));
        sc->n_nodes = 0;
        build_rec(&b, 0, (int64_t)n_tris - 1);
        free(cen);
    }
    return sc;
}

/* the rest of the UAcousticMaterial asset (MAT.h:26-33).  The model is the authors' own split of the
 * reflected energy (MaterialAcousticProcessor.cpp:50-66): Refl = 1 - alpha, tau limited to 1 - Refl, specular gain
 * Refl (1 - sigma), diffuse gain Refl sigma, transmitted gain tau; the mirror direction is the legacy tracer's
 * GetReflectionVector (COMP.cpp:186), transmission its pass-through (COMP.cpp:271-275).  ThicknessCm scales the
 * transmitted gain as a Beer-Lambert layer, the asset's default 2.5 cm being the reference thickness:
 * tau_eff = tau ^ (ThicknessCm / 2.5).  The lobe a walk takes cannot depend on the band, so the choice uses band means.
 * Every value is float arithmetic in this exact order -- the device tables (fs_api.cu) are built by the same sequence.
 * transmission / scattering / thickness may be NULL (0, 1, 2.5 cm: a purely diffuse surface, which reproduces the
 * reference model bit for bit). */
int fso_scene_set_material_model(fso_scene* sc, const float* transmission, const float* scattering, const float* thickness_cm)
{
    const uint32_t M = sc->n_mats, B = sc->n_bands;
    free(sc->bsdf_tab); free(sc->lobes);
    sc->bsdf_tab = (float*)calloc((size_t)3 * (M ? M : 1) * B, 4);
    sc->lobes = (float*)calloc((size_t)(M ? M : 1) * 4, 4);
    for (uint32_t m = 0; m < M; ++m) {
        float sum_r = 0.0f, sum_t = 0.0f, sum_s = 0.0f;
        const float th = thickness_cm ? thickness_cm[m] : 2.5f;
        const float ex = th / 2.5f;
        for (uint32_t b = 0; b < B; ++b) {
            const float a = sc->absorption[m * B + b];
            const float refl = 1.0f - a;
            float tau = transmission ? transmission[m * B + b] : 0.0f;
            if (refl + tau > 1.0f) tau = 1.0f - refl;                      /* MaterialAcousticProcessor.cpp:59-60 */
            const float sig = scattering ? scattering[m * B + b] : 1.0f;
            float tau_eff = 0.0f;
            if (tau > 0.0f) tau_eff = (ex == 1.0f) ? tau : fso_powf(tau, ex);
            sc->bsdf_tab[((size_t)0 * M + m) * B + b] = (refl * sig) / FSO_PI;        /* diffuse: Refl sigma / pi */
            sc->bsdf_tab[((size_t)1 * M + m) * B + b] = refl * (1.0f - sig);           /* specular: Refl (1 - sigma) */
            sc->bsdf_tab[((size_t)2 * M + m) * B + b] = tau_eff;                       /* transmitted */
            sum_r += refl; sum_t += tau; sum_s += sig;
        }
        const float mr = sum_r / (float)B, mt = sum_t / (float)B, ms = sum_s / (float)B;
        const float tot = mr + mt;
        const float t1 = (tot > 0.0f) ? mt / tot : 0.0f;                   /* P(transmit) */
        const float ps = (1.0f - t1) * (1.0f - ms);                         /* P(specular) */
        const float t2 = t1 + ps;
        const float pd = 1.0f - t2;                                         /* P(diffuse) */
        sc->lobes[4 * m + 0] = t1; sc->lobes[4 * m + 1] = t2; sc->lobes[4 * m + 2] = ps; sc->lobes[4 * m + 3] = pd;
    }
    return 0;
}

void fso_scene_destroy(fso_scene* sc)
{
    if (!sc) return;
    free(sc->v0); free(sc->e1); free(sc->e2); free(sc->nrm); free(sc->mat);
    free(sc->refl_over_pi); free(sc->bsdf_tab); free(sc->lobes); free(sc->absorption); free(sc->nodes); free(sc->order);
    free(sc);
}

/* reciprocal direction with zero components replaced by +-1e-30 so that (box - o) * inv is
 * never 0 * inf = NaN */
static inline void safe_inv(const float d[3], float inv[3])
{
    for (int a = 0; a < 3; ++a) {
        float da = d[a];
        if (!(fabsf(da) > 1e-30f)) da = (da < 0.0f) ? -1e-30f : 1e-30f;
        inv[a] = 1.0f / da;
    }
}

static inline int slab(const onode* n, const float o[3], const float inv[3], float tmax, float* tn)
{
    float t0 = 0.0f, t1 = tmax;
    for (int a = 0; a < 3; ++a) {
        float ta = (n->lo[a] - o[a]) * inv[a];
        float tb = (n->hi[a] - o[a]) * inv[a];
        float mn = ta < tb ? ta : tb;
        float mx = ta > tb ? ta : tb;
        /* NaN (0 * inf) compares false on both branches -> keeps the wider interval */
        if (mn > t0) t0 = mn;
        if (mx < t1) t1 = mx;
    }
    *tn = t0;
    return t0 <= t1;
}

static int closest_hit_cnt(const fso_scene* sc, const float o[3], const float d[3], float* t_out,
                           uint32_t* tri_out, ocount* cnt)
{
    float best = INFINITY;
    uint32_t best_id = 0xffffffffu;
    if (!sc->use_bvh || sc->n_tris == 0) {
        for (uint64_t t = 0; t < sc->n_tris; ++t) {
            float tt;
            if (cnt) cnt->tri_tests++;
            if (fso_intersect_tri(o, d, sc->v0 + 3 * t, sc->e1 + 3 * t, sc->e2 + 3 * t, &tt)) {
                if (tt < best || (tt == best && (uint32_t)t < best_id)) { best = tt; best_id = (uint32_t)t; }
            }
        }
    } else {
        float inv[3];
        safe_inv(d, inv);
        int32_t stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const onode* n = &sc->nodes[stack[--sp]];
            float tn;
            if (cnt) cnt->node_visits++;
            if (!slab(n, o, inv, best, &tn)) continue;
            if (n->left < 0) {
                for (uint32_t i = 0; i < n->count; ++i) {
                    uint32_t t = sc->order[n->first + i];
                    float tt;
                    if (cnt) cnt->tri_tests++;
                    if (fso_intersect_tri(o, d, sc->v0 + 3 * (uint64_t)t, sc->e1 + 3 * (uint64_t)t,
                                          sc->e2 + 3 * (uint64_t)t, &tt)) {
                        if (tt < best || (tt == best && t < best_id)) { best = tt; best_id = t; }
                    }
                }
            } else {
                stack[sp++] = n->left;
                stack[sp++] = n->right;
            }
        }
    }
    if (best_id == 0xffffffffu) return 0;
    *t_out = best; *tri_out = best_id;
    return 1;
}

static int any_hit_cnt(const fso_scene* sc, const float o[3], const float d[3], float tmax, ocount* cnt)
{
    if (!sc->use_bvh || sc->n_tris == 0) {
        for (uint64_t t = 0; t < sc->n_tris; ++t) {
            float tt;
            if (cnt) cnt->tri_tests++;
            if (fso_intersect_tri(o, d, sc->v0 + 3 * t, sc->e1 + 3 * t, sc->e2 + 3 * t, &tt) && tt < tmax)
                return 1;
        }
        return 0;
    }
    float inv[3];
    safe_inv(d, inv);
    int32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const onode* n = &sc->nodes[stack[--sp]];
        float tn;
        if (cnt) cnt->node_visits++;
        if (!slab(n, o, inv, tmax, &tn)) continue;
        if (n->left < 0) {
            for (uint32_t i = 0; i < n->count; ++i) {
                uint32_t t = sc->order[n->first + i];
                float tt;
                if (cnt) cnt->tri_tests++;
                if (fso_intersect_tri(o, d, sc->v0 + 3 * (uint64_t)t, sc->e1 + 3 * (uint64_t)t,
                                      sc->e2 + 3 * (uint64_t)t, &tt) && tt < tmax)
                    return 1;
            }
        } else {
            stack[sp++] = n->left;
            stack[sp++] = n->right;
        }
    }
    return 0;
}

int fso_closest_hit(const fso_scene* sc, const float o[3], const float d[3], float* t, uint32_t* tri)
{
    return closest_hit_cnt(sc, o, d, t, tri, NULL);
}
int fso_any_hit(const fso_scene* sc, const float o[3], const float d[3], float tmax)
{
    return any_hit_cnt(sc, o, d, tmax, NULL);
}
void fso_scene_stats(const fso_scene* sc, fso_stats* st) { *st = sc->counters; }

/* ------------------------------------------------------------------------------------------
 * BDPT
 * ---------------------------------------------------------------------------------------- */
typedef struct { float p[3]; float n[3]; int32_t mat; float prob; uint32_t ev; float ro[3]; } pnode;   /* ev: FSO_EV_* the walk took AT this node; ro: origin of the ray that left it (= p, or the far side of the surface after a pass-through) */
#define FSO_EV_DIFFUSE 0u
#define FSO_EV_SPECULAR 1u
#define FSO_EV_TRANSMIT 2u

/* GeneratePath, SUB.cpp:279-355.  Dispositions: Philox replaces FRand; miss
 * terminates the subpath (FIX; the reference keeps looping on duplicate nodes); max_depth
 * bounds the number of rays (PARAM; reference unbounded); cosine hemisphere (FIX). */
static uint32_t gen_subpath(const fso_scene* sc, const fso_config* cfg, const float start[3],
                            uint64_t g, uint32_t side, uint32_t max_depth, uint64_t seed,
                            pnode* nodes, uint64_t* rays, ocount* cnt)
{
    float pos[3] = {start[0], start[1], start[2]};
    float nrm[3] = {0.0f, 0.0f, 0.0f};           /* CurrentNormal = ZeroVector, SUB.cpp:289 */
    int32_t mat = -1;                             /* CurrentMaterial = nullptr, SUB.cpp:290 */
    float prob = 1.0f;                            /* CurrentProbability = 1, SUB.cpp:291 */
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t nn = 0;
    const int model = (cfg->reserved[1] & FSO_FLAG_MATERIAL_MODEL) && sc->lobes;
    float din[3] = {0.0f, 0.0f, 0.0f};            /* direction the walk arrived with (mirror / pass-through events) */
    for (uint32_t k = 0;; ++k) {
        pnode* nd = &nodes[nn++];                 /* SUB.cpp:297-298: push node first */
        nd->p[0] = pos[0]; nd->p[1] = pos[1]; nd->p[2] = pos[2];
        nd->n[0] = nrm[0]; nd->n[1] = nrm[1]; nd->n[2] = nrm[2];
        nd->mat = mat; nd->prob = prob;
        nd->ev = FSO_EV_DIFFUSE;                  /* a node the walk ends at can only be connected through its diffuse lobe */
        nd->ro[0] = pos[0]; nd->ro[1] = pos[1]; nd->ro[2] = pos[2];
        if (k >= max_depth) break;
        uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), k, side};
        uint32_t r[4];
        fso_philox4x32_10(ctr, key, r);
        float u0 = fso_u01(r[0]), u1 = fso_u01(r[1]), u2 = fso_u01(r[2]);
        if (!(u0 < cfg->rr_prob)) break;          /* SUB.cpp:301-302, 349-353 */
        float dir[3], newprob;
        if (k == 0) {                             /* CurrentNormal.IsNearlyZero(), SUB.cpp:306 */
            fso_sample_sphere(u1, u2, dir);
            newprob = FSO_INV_4PI * cfg->rr_prob; /* SUB.cpp:309-310 */
        } else if (model) {
            /* the lobe is chosen by the fourth word of the bounce's Philox block */
            const float* lb = sc->lobes + 4 * (uint32_t)mat;
            const float u3 = fso_u01(r[3]);
            if (u3 < lb[0]) {                     /* pass through (COMP.cpp:271-275) */
                nd->ev = FSO_EV_TRANSMIT;
                dir[0] = din[0]; dir[1] = din[1]; dir[2] = din[2];
                newprob = cfg->rr_prob * lb[0];
            } else if (u3 < lb[1]) {              /* mirror (GetReflectionVector, COMP.cpp:186) */
                nd->ev = FSO_EV_SPECULAR;
                const float sdn = -2.0f * dot3(din, nrm);
                dir[0] = fmaf(sdn, nrm[0], din[0]); dir[1] = fmaf(sdn, nrm[1], din[1]); dir[2] = fmaf(sdn, nrm[2], din[2]);
                newprob = cfg->rr_prob * lb[2];
            } else {
                float ct;
                fso_sample_cos_hemisphere(nrm, u1, u2, dir, &ct);
                newprob = ((ct * FSO_INV_PI) * cfg->rr_prob) * lb[3];
            }
        } else {
            float ct;
            fso_sample_cos_hemisphere(nrm, u1, u2, dir, &ct);
            newprob = (ct * FSO_INV_PI) * cfg->rr_prob;   /* SUB.cpp:315-317 */
        }
        (*rays)++;
        float o[3] = {pos[0], pos[1], pos[2]};
        if (nd->ev == FSO_EV_TRANSMIT)            /* continue from the far side of the surface: 2 x the offset along -n */
            for (int a = 0; a < 3; ++a) { o[a] = fmaf(-2.0f * cfg->eps_offset, nrm[a], pos[a]); nd->ro[a] = o[a]; }
        float t; uint32_t tri;
        if (!closest_hit_cnt(sc, o, dir, &t, &tri, cnt)) { nd->ev = FSO_EV_DIFFUSE; break; }   /* FIX: miss terminates */
        /* SUB.cpp:345-347: pos = ImpactPoint + 0.1 * ImpactNormal; normal; material */
        const float* tn = sc->nrm + 3 * (uint64_t)tri;
        float fn[3] = {tn[0], tn[1], tn[2]};
        if (dot3(fn, dir) > 0.0f) { fn[0] = -fn[0]; fn[1] = -fn[1]; fn[2] = -fn[2]; }
        for (int a = 0; a < 3; ++a) {
            float hp = fmaf(t, dir[a], o[a]);
            pos[a] = fmaf(cfg->eps_offset, fn[a], hp);
            nrm[a] = fn[a];
            din[a] = dir[a];
        }
        mat = (int32_t)sc->mat[tri];
        prob = newprob;
    }
    return nn;
}

static inline float dist3(const float a[3], const float b[3])
{
    float d[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    return sqrtf(dot3(d, d));
}

/* one segment of EvaluatePath, SUB.cpp:368-399 */
static inline void eval_segment_ev(const fso_scene* sc, const fso_config* cfg, int32_t mat, uint32_t ev, float prob,
                                   float d, float* total, float* E)
{
    const float* tab = ((cfg->reserved[1] & FSO_FLAG_MATERIAL_MODEL) && sc->bsdf_tab)
                       ? sc->bsdf_tab + (size_t)ev * sc->n_mats * sc->n_bands : sc->refl_over_pi;
    *total += d;                                  /* ScaledDistance += NodeDistance, :374 */
    if (d < cfg->min_seg) return;                 /* :375-378 */
    float G = 1.0f / (FSO_FOUR_PI * (d * d));     /* :391 */
    float P = fso_powf(prob, cfg->pdf_exponent);  /* :398 */
    for (uint32_t b = 0; b < cfg->n_bands; ++b) {
        float bs = (mat >= 0) ? tab[(uint32_t)mat * sc->n_bands + b] : 1.0f; /* :382-386 */
        float e = E[b];
        e *= bs;                                  /* :392 */
        e *= G;                                   /* :393 */
        e *= fso_expf(-cfg->air_absorption[b] * d); /* :395-397 */
        e /= P;                                   /* :398 */
        E[b] = e;
    }
}
static inline void eval_segment(const fso_scene* sc, const fso_config* cfg, int32_t mat, float prob, float d, float* total, float* E)
{
    eval_segment_ev(sc, cfg, mat, FSO_EV_DIFFUSE, prob, d, total, E);
}

/* AddEnergyAtDelay, COMP.h:87-91: BinIndex = Clamp(FloorToInt(DelaySeconds * 1000 / BinSizeMs), 0, Num - 1) */
int32_t fso_bin_index(const fso_config* cfg, float delay_s)
{
    float fb = floorf((delay_s * 1000.0f) / cfg->bin_ms);
    if (!(fb >= 0.0f)) return 0;
    if (fb >= (float)(cfg->n_bins - 1)) return (int32_t)cfg->n_bins - 1;      /* late energy clamps into the last bin */
    return (int32_t)fb;
}

/* EvaluatePath (SUB.cpp:360-420) on an explicit node list, exactly the reference's loop over consecutive node pairs
 * i -> i + 1 with node i's material and probability; pos [n][3] metres, mat[i] < 0 = no material.  Returns the delay and
 * the clamped, gained energy per band.  (splat_path below is the same walk over F ++ reverse(B) without building the list.) */
void fso_evaluate_nodes(const fso_scene* sc, const fso_config* cfg, const float* pos, const int32_t* mat, const float* prob,
                        uint32_t n, float* delay_out, float* energy_out)
{
    float E[FSO_MAX_BANDS];
    for (uint32_t b = 0; b < cfg->n_bands; ++b) E[b] = 1.0f;
    float total = 0.0f;
    for (uint32_t i = 0; i + 1 < n; ++i)
        eval_segment(sc, cfg, mat[i], prob[i], dist3(pos + 3 * i, pos + 3 * (i + 1)), &total, E);
    *delay_out = total / cfg->sound_speed;
    for (uint32_t b = 0; b < cfg->n_bands; ++b) {
        float e = E[b];
        e = (e < cfg->energy_clamp) ? e : cfg->energy_clamp;
        energy_out[b] = e * cfg->energy_gain;
    }
}

/* EvaluatePath over F[0..nf) ++ reverse(B[0..nb)) (SUB.cpp:259-267, 360-420), then AddEnergyAtDelay (COMP.h:87-91).
 * `len` = length of the connecting segment F[nf-1] -> B[nb-1]; `weight` multiplies the clamped, gained energy
 * (1 for the reference's endpoint connection). */
static void splat_path(const fso_scene* sc, const fso_config* cfg, const pnode* fn, uint32_t nf, const pnode* bn, uint32_t nb,
                       float len, float weight, uint64_t* hist_src, fso_path_dbg* dbg)
{
    float E[FSO_MAX_BANDS];
    for (uint32_t b = 0; b < cfg->n_bands; ++b) E[b] = 1.0f;
    float total = 0.0f;
    /* the factor of a node is the one of the event the walk took there; where a subpath is cut (its last node, or the
     * prefix end of an all-prefix connection) the connection leaves in an arbitrary direction: the diffuse lobe */
    for (uint32_t i = 0; i + 1 < nf; ++i)
        eval_segment_ev(sc, cfg, fn[i].mat, fn[i].ev, fn[i].prob, dist3(fn[i].ro, fn[i + 1].p), &total, E);
    eval_segment_ev(sc, cfg, fn[nf - 1].mat, FSO_EV_DIFFUSE, fn[nf - 1].prob, len, &total, E);
    for (uint32_t j = nb - 1; j >= 1; --j)
        eval_segment_ev(sc, cfg, bn[j].mat, (j == nb - 1) ? FSO_EV_DIFFUSE : bn[j].ev, bn[j].prob, dist3(bn[j - 1].ro, bn[j].p), &total, E);
    float delay = total / cfg->sound_speed;       /* :419 */
    const int32_t bin = fso_bin_index(cfg, delay);
    for (uint32_t b = 0; b < cfg->n_bands; ++b) {
        float e = E[b];
        e = (e < cfg->energy_clamp) ? e : cfg->energy_clamp;   /* :410, NaN -> clamp */
        e = e * cfg->energy_gain;                               /* :413 */
        if (weight != 1.0f) e = e * weight;
        uint64_t q = (uint64_t)(e * 4294967296.0f);             /* Q32.32, truncation */
        hist_src[(uint64_t)b * cfg->n_bins + (uint32_t)bin] += q;
        if (dbg) dbg->energy[b] = e;
    }
    if (dbg) { dbg->connected = 1; dbg->bin = bin; dbg->delay_s = delay; dbg->total_dist = total; }
}

/* all prefix connections.  Is_NaiveConnections (SUB.cpp:508-535) connects every forward prefix path
 * with every backward prefix path through ConnectSubpaths; its MIS weights are unfinished (hard-coded pdf, SUB.cpp:537-597).
 * Here: every (s, t), 1 <= s <= nf, 1 <= t <= nb, is connected with the visibility rule of SUB.cpp:252-257, evaluated
 * exactly like the endpoint connection, and weighted 1 / (s + t - 1) = one over the number of (s', t') strategies that
 * build a path of the same number of nodes.  (nf, nb) is the reference's connection, (1, 1) the direct path. */
/* SURVEY 8f rank 1, second half -- FSO_FLAG_MIS: all prefix connections combined with the BALANCE HEURISTIC (Veach), which is what
 * the reference's unfinished getPdf / getExpectedWeight / MISEnergy aim at (SUB.cpp:537-597: hard-coded pdf 0.9, integer
 * division).  Multiple importance sampling needs ONE integrand for all strategies, which the reference's per-segment product
 * (no cosines, pdf^0.1) is not; this mode therefore evaluates the physically based contribution in area measure
 *     f(x_0 .. x_{k-1}) = 1/(4 pi) * prod_edges [ cos_a cos_b / d^2 * exp(-air d) ] * prod_interior [ rho / pi ]
 * (cosines at the point source / point listener = 1; for k = 2 this is the reference's direct-path value 1/(4 pi d^2)) and
 * weights the sample of strategy (s, t) by p_{s,t} / sum_{s'} p_{s',k-s'}, i.e. contributes f / sum_{s'} p_{s'}.
 * Densities in area measure: a vertex generated from x_a towards x_b has rr * D_a * cos_b / d^2 with D = 1/(4 pi) at the two
 * end points (uniform sphere, SUB.cpp:308-310) and cos_a / pi on surfaces (cosine lobe, SUB.cpp:312-318 with the FIX of A3).
 * Strategies that a walk of at most max_depth rays cannot produce have density zero.  Edges shorter than min_seg drop the
 * sample (the reference's glitch guard, SUB.cpp:375-378).  Everything is float arithmetic in a fixed order: the CUDA kernel
 * (k_eval_mis) executes the same sequence.  FSO_FLAG_MIS_T1 / _S1 (oracle only, tests): a single strategy family with weight
 * one -- an independent unbiased estimator of the same integral. */
#define FSO_MIS_MAXV 66
static void mis_splat(const fso_scene* sc, const fso_config* cfg, const pnode* fn, uint32_t s, const pnode* bn, uint32_t t,
                      uint32_t max_depth, uint64_t* hist_src)
{
    const uint32_t k = s + t;
    if (k > FSO_MIS_MAXV) return;
    float pf[FSO_MIS_MAXV], pb[FSO_MIS_MAXV], gg[FSO_MIS_MAXV], dlen[FSO_MIS_MAXV];
    float total = 0.0f;
    for (uint32_t i = 0; i + 1 < k; ++i) {
        const pnode* a = (i < s) ? &fn[i] : &bn[k - 1u - i];
        const pnode* b = (i + 1u < s) ? &fn[i + 1u] : &bn[k - 2u - i];
        const float dl[3] = {b->p[0] - a->p[0], b->p[1] - a->p[1], b->p[2] - a->p[2]};
        const float d2 = dot3(dl, dl);
        const float d = sqrtf(d2);
        total += d;
        if (d < cfg->min_seg) return;
        const float inv = 1.0f / d;
        const float dir[3] = {dl[0] * inv, dl[1] * inv, dl[2] * inv};
        const float cp = (i == 0u) ? 1.0f : fabsf(dot3(a->n, dir));
        const float cm = (i + 2u == k) ? 1.0f : fabsf(dot3(b->n, dir));
        const float rd2 = 1.0f / d2;
        gg[i] = (cp * cm) * rd2;
        if (!(gg[i] > 0.0f)) return;
        pf[i] = (cfg->rr_prob * ((i == 0u) ? FSO_INV_4PI : cp * FSO_INV_PI)) * (cm * rd2);
        pb[i] = (cfg->rr_prob * ((i + 2u == k) ? FSO_INV_4PI : cm * FSO_INV_PI)) * (cp * rd2);
        dlen[i] = d;
    }
    /* sum over the strategies s' of p_{s'} / p_s; valid s': s' - 1 <= max_depth and k - s' - 1 <= max_depth */
    float sum = 1.0f;
    const uint32_t flags = cfg->reserved[1];
    if (flags & (FSO_FLAG_MIS_T1 | FSO_FLAG_MIS_S1)) {
        if ((flags & FSO_FLAG_MIS_T1) && t != 1u) return;
        if ((flags & FSO_FLAG_MIS_S1) && s != 1u) return;
    } else {
        float r = 1.0f;
        for (uint32_t sp = s; sp + 1u < k && sp <= max_depth; ++sp) {            /* s' = sp + 1 */
            r = r * (pf[sp - 1u] / pb[sp]);
            sum += r;
        }
        r = 1.0f;
        for (uint32_t sp = s; sp > 1u && k - sp <= max_depth; --sp) {            /* s' = sp - 1: t' - 1 = k - sp */
            r = r * (pb[sp - 1u] / pf[sp - 2u]);
            sum += r;
        }
    }
    const float delay = total / cfg->sound_speed;
    const int32_t bin = fso_bin_index(cfg, delay);
    for (uint32_t b = 0; b < cfg->n_bands; ++b) {
        float val = FSO_INV_4PI;
        for (uint32_t i = 0; i + 1 < k; ++i) {
            val = val * gg[i];
            val = val * fso_expf(-cfg->air_absorption[b] * dlen[i]);
            const uint32_t j = i + 1u;
            if (j + 1u < k) {                                                    /* interior vertex j */
                const pnode* v = (j < s) ? &fn[j] : &bn[k - 1u - j];
                val = val * sc->refl_over_pi[(uint32_t)v->mat * sc->n_bands + b];
                val = val / ((j < s) ? pf[j - 1u] : pb[j]);
            }
        }
        float e = val / sum;
        e = (e < cfg->energy_clamp) ? e : cfg->energy_clamp;
        e = e * cfg->energy_gain;
        hist_src[(uint64_t)b * cfg->n_bins + (uint32_t)bin] += (uint64_t)(e * 4294967296.0f);
    }
}

static void connect_all(const fso_scene* sc, const fso_config* cfg, const pnode* fn, uint32_t nf, const pnode* bn, uint32_t nb,
                        uint32_t max_depth, uint64_t* hist_src, fso_stats* st, ocount* cnt)
{
    const int mis = (cfg->reserved[1] & FSO_FLAG_MIS) != 0;
    for (uint32_t s = 1; s <= nf; ++s)
        for (uint32_t t = 1; t <= nb; ++t) {
            const pnode* F = &fn[s - 1];
            const pnode* Bn = &bn[t - 1];
            float dl[3] = {Bn->p[0] - F->p[0], Bn->p[1] - F->p[1], Bn->p[2] - F->p[2]};
            float len = sqrtf(dot3(dl, dl));
            float tmax = len - cfg->eps_connect;
            if (tmax > 0.0f) {
                float inv = 1.0f / len;
                float dir[3] = {dl[0] * inv, dl[1] * inv, dl[2] * inv};
                st->shadow_rays++;
                if (any_hit_cnt(sc, F->p, dir, tmax, cnt)) continue;
            }
            st->connected++;
            if (mis) mis_splat(sc, cfg, fn, s, bn, t, max_depth, hist_src);
            else splat_path(sc, cfg, fn, s, bn, t, len, 1.0f / (float)(s + t - 1u), hist_src, NULL);
        }
}

static void trace_one(const fso_scene* sc, const fso_config* cfg, const float* src, const float* lis,
                      uint64_t g, uint64_t n_paths, uint32_t max_depth, uint64_t seed, pnode* fn, pnode* bn,
                      uint64_t* hist_src, fso_stats* st, ocount* cnt, fso_path_dbg* dbg)
{
    uint64_t rays = 0;
    /* GenerateFullPaths, SUB.cpp:215-230: forward from the source, backward from the listener */
    uint32_t nf = gen_subpath(sc, cfg, src, g, 0u, max_depth, seed, fn, &rays, cnt);
    /* with FSO_FLAG_SHARE_LISTENER the listener subpath of pair (source, i) depends on i only (its
     * Philox stream is keyed by i = g mod n_paths), so the sources of a multi-emitter update share it.  The reference
     * regenerates it per (source, i) (SUB.cpp:215-230): a separate mode because the results differ. */
    const uint64_t gl = (cfg->reserved[1] & FSO_FLAG_SHARE_LISTENER) ? g % n_paths : g;
    uint32_t nb = gen_subpath(sc, cfg, lis, gl, 1u, max_depth, seed, bn, &rays, cnt);
    st->ext_rays += rays;
    st->paths++;
    if (cfg->reserved[1] & (FSO_FLAG_CONNECT_ALL | FSO_FLAG_MIS)) {           /* fs_config.flags lives in reserved[1] */
        connect_all(sc, cfg, fn, nf, bn, nb, max_depth, hist_src, st, cnt);
        return;
    }
    /* ConnectSubpaths, SUB.cpp:235-277: one visibility test between the two LAST nodes */
    const pnode* F = &fn[nf - 1];
    const pnode* Bn = &bn[nb - 1];
    float dl[3] = {Bn->p[0] - F->p[0], Bn->p[1] - F->p[1], Bn->p[2] - F->p[2]};
    float len = sqrtf(dot3(dl, dl));
    float tmax = len - cfg->eps_connect;          /* :253: end 0.1 short of the listener node */
    int occluded = 0;
    if (tmax > 0.0f) {
        float inv = 1.0f / len;
        float dir[3] = {dl[0] * inv, dl[1] * inv, dl[2] * inv};
        st->shadow_rays++;
        occluded = any_hit_cnt(sc, F->p, dir, tmax, cnt);
    }
    if (dbg) {
        memset(dbg, 0, sizeof(*dbg));
        dbg->n_src_nodes = nf; dbg->n_lis_nodes = nb; dbg->bin = -1;
        for (int a = 0; a < 3; ++a) { dbg->src_end[a] = F->p[a]; dbg->lis_end[a] = Bn->p[a]; }
    }
    if (occluded) return;
    st->connected++;
    splat_path(sc, cfg, fn, nf, bn, nb, len, 1.0f, hist_src, dbg);
}

/* the node lists of ONE path pair (known-answer tests of the walk): per node 8 floats (p.xyz, ro.xyz, prob, bits: mat + 1 | ev << 24) */
int fso_debug_path(const fso_scene* sc, const fso_config* cfg, const float* src, const float* lis, uint64_t g, uint64_t n_paths,
                   uint32_t max_depth, uint64_t seed, float* f_nodes, uint32_t* nf_out, float* b_nodes, uint32_t* nb_out)
{
    pnode* fn = (pnode*)malloc(sizeof(pnode) * (max_depth + 2));
    pnode* bn = (pnode*)malloc(sizeof(pnode) * (max_depth + 2));
    uint64_t rays = 0;
    const uint64_t gl = (cfg->reserved[1] & FSO_FLAG_SHARE_LISTENER) ? g % n_paths : g;
    const uint32_t nf = gen_subpath(sc, cfg, src, g, 0u, max_depth, seed, fn, &rays, NULL);
    const uint32_t nb = gen_subpath(sc, cfg, lis, gl, 1u, max_depth, seed, bn, &rays, NULL);
    for (int side = 0; side < 2; ++side) {
        const pnode* nd = side ? bn : fn;
        float* out = side ? b_nodes : f_nodes;
        const uint32_t n = side ? nb : nf;
        for (uint32_t i = 0; i < n; ++i) {
            for (int a = 0; a < 3; ++a) { out[8 * i + a] = nd[i].p[a]; out[8 * i + 3 + a] = nd[i].ro[a]; }
            out[8 * i + 6] = nd[i].prob;
            out[8 * i + 7] = u2f((uint32_t)(nd[i].mat + 1) | (nd[i].ev << 24));
        }
    }
    *nf_out = nf; *nb_out = nb;
    free(fn); free(bn);
    return 0;
}

typedef struct {
    const fso_scene* sc; const fso_config* cfg; const float* src_pos; const float* lis_pos;
    uint64_t n_paths, g_first, g_count; uint32_t max_depth; uint64_t seed;
    uint64_t* hist; uint64_t hsz; fso_path_dbg* dbg; int private_hist;
    uint64_t* cursor;                 /* shared chunk cursor (atomic) */
    fso_stats st; ocount cnt; uint64_t* h;
} worker;

static void* worker_main(void* arg)
{
    worker* w = (worker*)arg;
    const fso_config* cfg = w->cfg;
    pnode* fn = (pnode*)malloc(sizeof(pnode) * (w->max_depth + 2));
    pnode* bn = (pnode*)malloc(sizeof(pnode) * (w->max_depth + 2));
    w->h = w->private_hist ? (uint64_t*)calloc(w->hsz, 8) : w->hist;
    const uint64_t chunk = 256;
    for (;;) {
        uint64_t lo = __atomic_fetch_add(w->cursor, chunk, __ATOMIC_RELAXED);
        if (lo >= w->g_count) break;
        uint64_t hi = lo + chunk < w->g_count ? lo + chunk : w->g_count;
        for (uint64_t i = lo; i < hi; ++i) {
            uint64_t g = w->g_first + i;
            uint32_t s = (uint32_t)(g / w->n_paths);
            trace_one(w->sc, cfg, w->src_pos + 3 * s, w->lis_pos, g, w->n_paths, w->max_depth, w->seed, fn, bn,
                      w->h + (uint64_t)s * cfg->n_bands * cfg->n_bins, &w->st, &w->cnt,
                      w->dbg ? w->dbg + i : NULL);
        }
    }
    free(fn); free(bn);
    return NULL;
}

int fso_trace(const fso_scene* sc, const fso_config* cfg, const float* src_pos, uint32_t n_src,
              const float lis_pos[3], uint64_t n_paths, uint64_t g_first, uint64_t g_count,
              uint32_t max_depth, uint64_t seed, uint64_t* hist, fso_stats* stats,
              fso_path_dbg* dbg, int n_threads)
{
    if (!sc || !cfg || !hist) return -1;
    if (cfg->n_bands != sc->n_bands || cfg->n_bands > FSO_MAX_BANDS) return -2;
    if (n_paths == 0 || g_first + g_count > (uint64_t)n_src * n_paths) return -3;
    const uint64_t hsz = (uint64_t)n_src * cfg->n_bands * cfg->n_bins;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 512) n_threads = 512;
    uint64_t cursor = 0;
    worker* ws = (worker*)calloc((size_t)n_threads, sizeof(worker));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) {
        worker* w = &ws[t];
        w->sc = sc; w->cfg = cfg; w->src_pos = src_pos; w->lis_pos = lis_pos;
        w->n_paths = n_paths; w->g_first = g_first; w->g_count = g_count;
        w->max_depth = max_depth; w->seed = seed; w->hist = hist; w->hsz = hsz; w->dbg = dbg;
        w->private_hist = (n_threads > 1); w->cursor = &cursor;
    }
    if (n_threads == 1) worker_main(&ws[0]);
    else {
        for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, worker_main, &ws[t]);
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    }
    fso_stats total;
    memset(&total, 0, sizeof(total));
    for (int t = 0; t < n_threads; ++t) {
        worker* w = &ws[t];
        if (w->private_hist) {      /* integer sums: order independent, result stays bit-exact */
            for (uint64_t k = 0; k < hsz; ++k) hist[k] += w->h[k];
            free(w->h);
        }
        total.paths += w->st.paths; total.ext_rays += w->st.ext_rays;
        total.shadow_rays += w->st.shadow_rays; total.connected += w->st.connected;
        total.node_visits += w->cnt.node_visits; total.tri_tests += w->cnt.tri_tests;
    }
    free(ws); free(th);
    if (stats) *stats = total;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * IR: ReconstructImpulseResponse, COMP.cpp:320-380
 * ---------------------------------------------------------------------------------------- */
/* NormalizeImpulseResponse, COMP.cpp:382-406 -- the L2 normalisation the reference computes (and then zeroes, ":404 FIXME",
 * on a copy it throws away, :377-378): optional here (FSO_FLAG_IR_NORMALIZE), applied per channel to the IR that is output;
 * a channel whose norm is below KINDA_SMALL_NUMBER (1e-4) is left alone (:394-397).  Sum of squares in double. */
void fso_ir_normalize(const fso_config* cfg, float* ir)
{
    const uint32_t NS = cfg->sample_rate;
    for (uint32_t c = 0; c < cfg->n_channels; ++c) {
        float* x = ir + (uint64_t)c * NS;
        double ss = 0.0;
        for (uint32_t i = 0; i < NS; ++i) ss += (double)x[i] * (double)x[i];
        const float norm = (float)sqrt(ss);
        if (norm < 1e-4f) continue;
        for (uint32_t i = 0; i < NS; ++i) x[i] = x[i] / norm;
    }
}

int fso_build_ir_from_energy(const fso_config* cfg, const float* energy, float* ir_out)
{
    const uint32_t K = cfg->n_bins;
    const uint32_t NS = cfg->sample_rate;                       /* NumSamples, 1 s (COMP.h:139) */
    /* FIX: 48 samples per bin (reference ceil(0.001f*48000) = 49, COMP.cpp:324) */
    const uint32_t spb = (uint32_t)((double)cfg->bin_ms * 1e-3 * cfg->sample_rate + 0.5);
    const float Pi4 = sqrtf(4.0f * FSO_PI);                     /* COMP.cpp:323 */
    float* raw = (float*)calloc(NS, 4);
    for (uint32_t bin = 0; bin < K; ++bin) {
        if ((uint64_t)bin * spb >= NS) break;
        float e = 0.0f, pe = 0.0f;
        if (fabsf(energy[bin]) >= cfg->ir_threshold)            /* COMP.cpp:343-346 */
            e = energy[bin] / sqrtf(energy[bin] * Pi4);
        if (bin == 0) pe = e;                                    /* :348-351 */
        else if (fabsf(energy[bin - 1]) >= cfg->ir_threshold)   /* :352-355 */
            pe = energy[bin - 1] / sqrtf(energy[bin - 1] * Pi4);
        uint32_t nb = spb;
        if (NS - bin * spb < nb) nb = NS - bin * spb;            /* :342 */
        for (uint32_t j = 0; j < nb; ++j) {                      /* :357-363 */
            float w = (float)j / (float)spb;
            raw[bin * spb + j] = (1.0f - w) * pe + w * e;
        }
    }
    /* one-pole low-pass, COMP.cpp:366-375; output = filtered, un-normalised (:377-378) */
    const float a = cfg->ir_lowpass;
    float* out0 = ir_out;
    out0[0] = raw[0];
    for (uint32_t i = 1; i < NS; ++i) out0[i] = a * raw[i] + (1.0f - a) * out0[i - 1];
    for (uint32_t c = 1; c < cfg->n_channels; ++c)               /* both channels read the same mono histogram, :327-330 */
        memcpy(ir_out + (uint64_t)c * NS, out0, (size_t)NS * 4);
    free(raw);
    if (cfg->reserved[1] & FSO_FLAG_IR_NORMALIZE) fso_ir_normalize(cfg, ir_out);
    return 0;
}

int fso_build_ir(const fso_config* cfg, const uint64_t* hist, uint64_t n_paths, float* ir_out)
{
    /* band-summed integer histogram -> energy with the 1/N normalisation of SUB.cpp:164-168
     * applied after the integer reduction */
    const uint32_t K = cfg->n_bins;
    float* energy = (float*)malloc((size_t)K * 4);
    for (uint32_t k = 0; k < K; ++k) {
        uint64_t s = 0;
        for (uint32_t b = 0; b < cfg->n_bands; ++b) s += hist[(uint64_t)b * K + k];
        energy[k] = (float)(((double)s * (1.0 / 4294967296.0)) / (double)n_paths);
    }
    int rc = fso_build_ir_from_energy(cfg, energy, ir_out);
    free(energy);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Per-band IR synthesis (EXTENDS the reference, which collapses its IR to one
 * low-passed broadband envelope, COMP.cpp:337-378, README.md:118-122 "overly reverberant").
 * Every band keeps its own envelope a_b[k] (the mapping of COMP.cpp:339-363 applied per band) and
 * modulates a band-limited noise carrier:  ir[c][t] = sum_b ramp_b[t] * n_{c,b}[t].
 *   bands      octave bands, centre f_b = 62.5 * 2^b Hz (62.5 Hz ... 8 kHz for B = 8), capped at 0.45 fs
 *   carrier    white noise in [-1,1) from Philox4x32-10 (counter (t/4, 'IRNZ', c, b), key = seed), through
 *              two cascaded RBJ band-pass biquads (constant 0 dB peak gain, Q = sqrt 2), double precision,
 *              Direct Form I, then scaled to unit RMS over the 1 s; stored as float
 * No one-pole low-pass in this mode: the carriers are band-limited already.
 * ---------------------------------------------------------------------------------------- */
void fso_band_carriers(const fso_config* cfg, uint64_t seed, float* out /* [C][B][NS] */)
{
    const uint32_t NS = cfg->sample_rate, B = cfg->n_bands, Cn = cfg->n_channels;
    const double fs = (double)cfg->sample_rate;
    double* y = (double*)malloc((size_t)NS * sizeof(double));
    for (uint32_t c = 0; c < Cn; ++c)
        for (uint32_t b = 0; b < B; ++b) {
            double fc = 62.5 * (double)(1u << b);
            if (fc > 0.45 * fs) fc = 0.45 * fs;
            const double w0 = 2.0 * 3.14159265358979323846 * fc / fs, q = 1.4142135623730951;
            const double alpha = sin(w0) / (2.0 * q), a0 = 1.0 + alpha;
            const double b0 = alpha / a0, b2 = -alpha / a0, a1 = -2.0 * cos(w0) / a0, a2 = (1.0 - alpha) / a0;
            double x1 = 0, x2 = 0, u1 = 0, u2 = 0, v1 = 0, v2 = 0, ss = 0;
            uint32_t r[4] = {0, 0, 0, 0};
            for (uint32_t t = 0; t < NS; ++t) {
                if ((t & 3u) == 0) {
                    const uint32_t ctr[4] = {t >> 2, 0x49524E5Au, c, b};
                    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
                    fso_philox4x32_10(ctr, key, r);
                }
                const double x = (double)r[t & 3u] * (2.0 / 4294967296.0) - 1.0;
                const double u = b0 * x + b2 * x2 - a1 * u1 - a2 * u2;        /* first section */
                const double v = b0 * u + b2 * u2 - a1 * v1 - a2 * v2;        /* second section */
                x2 = x1; x1 = x; u2 = u1; u1 = u; v2 = v1; v1 = v;
                y[t] = v; ss += v * v;
            }
            const double g = ss > 0.0 ? 1.0 / sqrt(ss / (double)NS) : 0.0;
            float* o = out + ((size_t)c * B + b) * NS;
            for (uint32_t t = 0; t < NS; ++t) o[t] = (float)(y[t] * g);
        }
    free(y);
}

int fso_build_ir_bands(const fso_config* cfg, const uint64_t* hist, uint64_t n_paths, uint64_t noise_seed, float* ir_out)
{
    const uint32_t K = cfg->n_bins, NS = cfg->sample_rate, B = cfg->n_bands, Cn = cfg->n_channels;
    const uint32_t spb = (uint32_t)((double)cfg->bin_ms * 1e-3 * cfg->sample_rate + 0.5);
    const float Pi4 = sqrtf(4.0f * FSO_PI);
    float* car = (float*)malloc((size_t)Cn * B * NS * 4);
    float* amp = (float*)malloc((size_t)B * K * 4);
    fso_band_carriers(cfg, noise_seed, car);
    for (uint32_t b = 0; b < B; ++b)
        for (uint32_t k = 0; k < K; ++k) {
            const float e = (float)(((double)hist[(uint64_t)b * K + k] * (1.0 / 4294967296.0)) / (double)n_paths);
            amp[b * K + k] = (fabsf(e) >= cfg->ir_threshold) ? e / sqrtf(e * Pi4) : 0.0f;    /* COMP.cpp:343-346 per band */
        }
    for (uint32_t c = 0; c < Cn; ++c)
        for (uint32_t t = 0; t < NS; ++t) {
            const uint32_t bin = t / spb, j = t - bin * spb;
            float acc = 0.0f;
            if (bin < K) {
                const float w = (float)j / (float)spb;
                for (uint32_t b = 0; b < B; ++b) {
                    const float e = amp[b * K + bin], pe = bin ? amp[b * K + bin - 1] : e;   /* :347-363 */
                    const float raw = (1.0f - w) * pe + w * e;
                    acc = acc + raw * car[((size_t)c * B + b) * NS + t];
                }
            }
            ir_out[(uint64_t)c * NS + t] = acc;
        }
    free(car); free(amp);
    if (cfg->reserved[1] & FSO_FLAG_IR_NORMALIZE) fso_ir_normalize(cfg, ir_out);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Convolution semantics: ProcessSourceAudio + ConvolveFFT, REV.cpp:118-213.
 * y[n] = sum_m h[m] x[n-m] for the newest block, history initially zero (CIRC.cpp:15-21),
 * the IR current at the time of the call applied to the whole retained history.  Direct form
 * in double.  FIX: proper de-interleave (REV.cpp:147-148 copies interleaved data).
 * ---------------------------------------------------------------------------------------- */
struct fso_conv {
    fso_config cfg;
    uint32_t ir_len, hist_len;
    float* ir;        /* [C][ir_len] */
    float* hist;      /* [C][hist_len + block] */
};

fso_conv* fso_conv_create(const fso_config* cfg)
{
    fso_conv* cv = (fso_conv*)calloc(1, sizeof(*cv));
    cv->cfg = *cfg;
    cv->ir_len = cfg->sample_rate;                 /* IRSize = SamplingRate * 1 s, REV.cpp:79 */
    cv->hist_len = cv->ir_len - 1;                 /* tail rings of IRSize-1, REV.cpp:80-81 */
    cv->ir = (float*)calloc((size_t)cfg->n_channels * cv->ir_len, 4);
    cv->hist = (float*)calloc((size_t)cfg->n_channels * (cv->hist_len + cfg->conv_block), 4);
    return cv;
}
void fso_conv_destroy(fso_conv* cv) { if (cv) { free(cv->ir); free(cv->hist); free(cv); } }
void fso_conv_set_ir(fso_conv* cv, const float* ir)
{
    memcpy(cv->ir, ir, (size_t)cv->cfg.n_channels * cv->ir_len * 4);
}
int fso_conv_process(fso_conv* cv, const float* in, float* out, uint32_t frames)
{
    const uint32_t C = cv->cfg.n_channels, Bk = cv->cfg.conv_block;
    if (frames != Bk) return -1;
    const uint32_t L = cv->hist_len + Bk;
    for (uint32_t c = 0; c < C; ++c) {
        float* h = cv->hist + (uint64_t)c * L;
        for (uint32_t i = 0; i < Bk; ++i) h[cv->hist_len + i] = in[i * C + c];
        const float* ir = cv->ir + (uint64_t)c * cv->ir_len;
        for (uint32_t i = 0; i < Bk; ++i) {
            double acc = 0.0;
            const float* x = h + cv->hist_len + i;   /* x[n], x[n-m] = x[-m] */
            for (uint32_t m = 0; m < cv->ir_len; ++m) acc += (double)ir[m] * (double)x[-(int64_t)m];
            float y = (float)acc;
            if (cv->cfg.conv_clamp) { if (y > 1.0f) y = 1.0f; if (y < -1.0f) y = -1.0f; }  /* REV.cpp:162-168 */
            out[i * C + c] = y * cv->cfg.conv_wet + in[i * C + c] * (1.0f - cv->cfg.conv_wet);
        }
        memmove(h, h + Bk, (size_t)cv->hist_len * 4);  /* ring advance, CIRC.cpp:43-75 */
    }
    return 0;
}
