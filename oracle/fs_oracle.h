/*
 * fs_oracle.h -- CPU oracle for the FrequenSee BDPT -> histogram -> IR -> convolution path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under audio-pathtracer_b200/ (the product) may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY STATUS: pinned by the reference's own code.  The reference (henreedev/audio-pathtracer) ships no tests, golden
 * vectors or fixtures for this path and its plugin cannot be built without Unreal Engine 5.4, but the
 * function bodies on the path are plain C++ over a handful of engine types.  oracle/Makefile therefore compiles, from the
 * sources where they lie under /root/reference (nothing is copied into this repo):
 *   oracle/_ref/libref_ue_bodies.so  the UNMODIFIED bodies of UpdateSource, GenerateFullPaths, ConnectSubpaths, GeneratePath,
 *                                    EvaluatePath (SUB.cpp:128-420), FlushEnergyBuffer / AddEnergyAtDelay (COMP.h:76-91),
 *                                    ReconstructImpulseResponse (COMP.cpp:320-380), the reverb plugin's Initialize and
 *                                    ConvolveFFT (REV.cpp:74-102, 172-213), FCircularAudioBuffer (CIRC.cpp) and KissFFT,
 *                                    against the engine stand-in oracle/ue_shim/CoreMinimal.h (oracle/ref_ue_bodies.cpp,
 *                                    oracle/extract_ue_bodies.py)
 *   oracle/_ref/libref_kissfft.so    the reference's KissFFT + a restated driver of its convolution scheme
 * and tests/test_oracle_ref_ue.py runs this file against them with the reference's constants: same rays traced, same
 * per-path node counts / connections / delays / gains, histogram within 2e-5 (float vs double rounding), bin index exact,
 * ReconstructImpulseResponse BIT-EXACT.  What stands in for the engine (the absent third-party dependency): the scene query
 * (UWorld::LineTraceSingleByObjectType) and the random numbers (FMath::FRand / VRand / VRandCone -> Philox) -- that is where the
 * REPLACE / FIX dispositions of SURVEY.md section 8a enter.  The remaining dispositions (metres instead of "/ 1000", B bands,
 * 48 samples per bin, Q32.32 integers, miss terminates, max_depth) are parameters of this file whose reference values the pin
 * tests use; analytic known-answer tests (tests/test_oracle_kat.py) cover the rest.
 *
 * Reference aliases: SUB.cpp = Plugins/FrequenSee/Source/FrequenSee/Private/AudioRayTracingSubsystem.cpp
 *                    COMP.h/.cpp = .../FrequenSeeAudioComponent.{h,cpp}
 *                    REV.cpp = .../Private/FrequenSeeAudioReverbPlugin.cpp
 *                    CIRC.cpp = .../Private/CircularBuffer.cpp
 */
#ifndef FS_ORACLE_H
#define FS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSO_MAX_BANDS 8

/* Same field order/meaning as fs_config in include/frequensee.h (kept layout-compatible on
 * purpose so a test can hand the same bytes to both sides); defined independently here. */
typedef struct fso_config {
    uint32_t n_bands;          /* B, 1..8 (reference: effectively 1, Absorption[2], SUB.cpp:385) */
    uint32_t n_bins;           /* K, reference 1000 (COMP.h:137-138) */
    float    bin_ms;           /* reference BinSizeMs = 1 (COMP.h:72) */
    float    rr_prob;          /* RUSSIAN_ROULETTE_PROB 0.9 (SUB.cpp:282); 1.0 disables */
    float    eps_offset;       /* hit offset along normal: 0.1 cm -> 1e-3 m (SUB.cpp:345) */
    float    eps_connect;      /* connection-ray shortening: 0.1 cm -> 1e-3 m (SUB.cpp:253) */
    float    min_seg;          /* short-segment skip threshold in metres (SUB.cpp:375-378) */
    float    sound_speed;      /* 343 (SUB.cpp:362) */
    float    pdf_exponent;     /* 0.1 (SUB.cpp:398); 1.0 = unbiased estimator */
    float    energy_clamp;     /* 1.0 (SUB.cpp:410) */
    float    energy_gain;      /* 10.0 (SUB.cpp:413) */
    float    air_absorption[FSO_MAX_BANDS]; /* per metre per band (SUB.cpp:395: 0.05 per 10 m unit) */
    uint32_t sample_rate;      /* 48000 (COMP.h:133) */
    uint32_t n_channels;       /* 2 (COMP.h:135) */
    float    ir_threshold;     /* 1e-6 (COMP.cpp:322) */
    float    ir_lowpass;       /* 0.25 (COMP.cpp:366) */
    uint32_t conv_block;       /* 1024 (Config/DefaultEngine.ini:15) */
    uint32_t conv_clamp;       /* 1 = clamp output to +-1 (REV.cpp:162-168) */
    float    conv_wet;         /* MixAlpha = 1 (REV.cpp:161) */
    uint32_t reserved[3];      /* fs_config: max_batch_paths, flags, device; the oracle reads flags & FSO_FLAG_CONNECT_ALL */
} fso_config;
#define FSO_FLAG_SHARE_LISTENER 128u /* listener subpath keyed by the path index only: shared by all sources */
#define FSO_FLAG_MATERIAL_MODEL 256u /* transmission / scattering / thickness of the material asset drive the walk */
#define FSO_FLAG_IR_NORMALIZE 1024u  /* every IR that is built is scaled to unit L2 norm per channel (NormalizeImpulseResponse, COMP.cpp:382-406) */
#define FSO_FLAG_MIS 512u             /* all prefix connections weighted by the balance heuristic over a physically based contribution */
#define FSO_FLAG_MIS_T1 (1u << 20)    /* oracle only (tests): only the strategies t = 1, weight 1 -- an independent estimator of the same integral */
#define FSO_FLAG_MIS_S1 (1u << 21)    /* oracle only (tests): only the strategies s = 1 */
#define FSO_FLAG_CONNECT_ALL 64u   /* all prefix connections, weight 1/(s+t-1) */

typedef struct fso_stats {
    uint64_t paths;            /* path pairs processed */
    uint64_t ext_rays;         /* extension rays traced (closest hit) */
    uint64_t shadow_rays;      /* connection rays traced (any hit) */
    uint64_t connected;        /* pairs whose connection was unoccluded (SUB.cpp:232 log) */
    uint64_t node_visits;      /* oracle BVH nodes popped (0 in brute-force mode) */
    uint64_t tri_tests;        /* triangle tests executed */
} fso_stats;

typedef struct fso_scene fso_scene;

/* per-path debug record (one per path pair), used to localise a parity failure */
typedef struct fso_path_dbg {
    uint32_t n_src_nodes;      /* nodes in the source subpath (>=1) */
    uint32_t n_lis_nodes;      /* nodes in the listener subpath (>=1) */
    uint32_t connected;        /* 1 if the endpoints saw each other */
    int32_t  bin;              /* histogram bin, -1 if not connected */
    float    delay_s;
    float    total_dist;
    float    energy[FSO_MAX_BANDS];   /* after clamp * gain */
    float    src_end[3];
    float    lis_end[3];
} fso_path_dbg;

void fso_default_config(fso_config* cfg);

/* ---- shared-arithmetic primitives (exposed so tests can pin them individually) ---- */
void  fso_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float fso_u01(uint32_t r);
float fso_expf(float x);
float fso_logf(float x);
float fso_powf(float x, float e);
void  fso_sincos_2pi(float u, float* c, float* s);
void  fso_sample_sphere(float u1, float u2, float dir[3]);
void  fso_sample_cos_hemisphere(const float n[3], float u1, float u2, float dir[3], float* cos_theta);
/* Moller-Trumbore, two sided; returns 1 on hit and writes t */
int   fso_intersect_tri(const float o[3], const float d[3], const float v0[3], const float e1[3],
                        const float e2[3], float* t);

/* ---- scene ---- */
/* verts: [T][3][3] metres; tri_mat: [T]; absorption: [M][B].  use_bvh=0 -> brute force */
fso_scene* fso_scene_create(const float* verts, const uint32_t* tri_mat, uint64_t n_tris,
                            const float* absorption, uint32_t n_mats, uint32_t n_bands, int use_bvh);
void fso_scene_destroy(fso_scene* sc);
/* SURVEY 8f rank 3 (FSO_FLAG_MATERIAL_MODEL): transmission [M][B], scattering [M][B], thickness_cm [M]; any may be NULL */
int  fso_scene_set_material_model(fso_scene* sc, const float* transmission, const float* scattering, const float* thickness_cm);
/* closest hit by lexicographic min of (t, tri_id); returns 1 on hit */
int  fso_closest_hit(const fso_scene* sc, const float o[3], const float d[3], float* t, uint32_t* tri);
int  fso_any_hit(const fso_scene* sc, const float o[3], const float d[3], float tmax);
void fso_scene_stats(const fso_scene* sc, fso_stats* st);   /* visit counters since creation */

/* ---- BDPT (SUB.cpp:128-420) ---- */
/* Traces global work indices g in [g_first, g_first+g_count) where g = source*n_paths + i.
 * hist: [S][B][K] u64 Q32.32, ACCUMULATED into (caller zeroes).  dbg may be NULL, else [g_count].
 * n_threads <= 1 -> single thread (reference-faithful: game thread, SUB.cpp:55-85). */
int fso_trace(const fso_scene* sc, const fso_config* cfg, const float* src_pos, uint32_t n_src,
              const float lis_pos[3], uint64_t n_paths, uint64_t g_first, uint64_t g_count,
              uint32_t max_depth, uint64_t seed, uint64_t* hist, fso_stats* stats,
              fso_path_dbg* dbg, int n_threads);

/* node lists of one path pair g: per node 8 floats (p.xyz, ray origin.xyz, prob, bits(mat + 1 | event << 24)); [max_depth + 1] each */
int fso_debug_path(const fso_scene* sc, const fso_config* cfg, const float* src, const float* lis, uint64_t g, uint64_t n_paths,
                   uint32_t max_depth, uint64_t seed, float* f_nodes, uint32_t* nf_out, float* b_nodes, uint32_t* nb_out);
/* EvaluatePath (SUB.cpp:360-420) on an explicit node list; AddEnergyAtDelay's bin index (COMP.h:87-91).  Exposed so that
 * tests can pin them one to one against the reference's own bodies (oracle/_ref/libref_ue_bodies.so). */
void fso_evaluate_nodes(const fso_scene* sc, const fso_config* cfg, const float* pos, const int32_t* mat, const float* prob,
                        uint32_t n, float* delay_out, float* energy_out);
int32_t fso_bin_index(const fso_config* cfg, float delay_s);

/* ---- IR (COMP.cpp:320-380) ---- */
/* hist: [B][K] for one source; ir_out: [C][sample_rate] */
int fso_build_ir(const fso_config* cfg, const uint64_t* hist, uint64_t n_paths, float* ir_out);
/* float-histogram variant: the reference's own signature, EnergyBuffer float[K] -> IR */
void fso_ir_normalize(const fso_config* cfg, float* ir /* [C][sample_rate], in place */);
int fso_build_ir_from_energy(const fso_config* cfg, const float* energy, float* ir_out);
/* per-band synthesis: band envelopes x band-limited noise carriers; carriers [C][B][sample_rate] */
void fso_band_carriers(const fso_config* cfg, uint64_t seed, float* out);
int fso_build_ir_bands(const fso_config* cfg, const uint64_t* hist, uint64_t n_paths, uint64_t noise_seed, float* ir_out);

/* ---- convolution semantics (REV.cpp:118-213, CIRC.cpp:43-75) ---- */
/* Streaming direct-form reference in double precision.  State = per-channel history of
 * (ir_len-1) samples, initially zero (CIRC.cpp:15-21). */
typedef struct fso_conv fso_conv;
fso_conv* fso_conv_create(const fso_config* cfg);
void fso_conv_destroy(fso_conv* cv);
/* ir: [C][sample_rate]; takes effect for the next processed block */
void fso_conv_set_ir(fso_conv* cv, const float* ir);
/* in/out interleaved [frames][C]; frames must equal cfg->conv_block */
int  fso_conv_process(fso_conv* cv, const float* in_interleaved, float* out_interleaved, uint32_t frames);

#ifdef __cplusplus
}
#endif
#endif
