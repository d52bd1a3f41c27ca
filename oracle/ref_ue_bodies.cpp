/*
 * ref_ue_bodies.cpp -- C entry points around the REFERENCE's own BDPT / IR / convolution code.
 *
 * TEST INFRASTRUCTURE ONLY (see fs_oracle.h).  Built by `make -C oracle ref_ue` into oracle/_ref/libref_ue_bodies.so
 * when /root/reference is present.  What is compiled:
 *   - the reference's headers, included where they lie: AudioRayTracingSubsystem.h, FrequenSeeAudioComponent.h,
 *     AcousticGeometryComponent.h, AcousticMaterial.h, CircularBuffer.h (and CircularBuffer.cpp + KissFFT as their
 *     own translation units, in place);
 *   - the UNMODIFIED bodies of UpdateSource, GenerateFullPaths, ConnectSubpaths, GeneratePath, EvaluatePath
 *     (SUB.cpp:128-420), the component's constructor, ReconstructImpulseResponse and NormalizeImpulseResponse
 *     (COMP.cpp:16-40, 320-406), FlushEnergyBuffer / AddEnergyAtDelay (inline in COMP.h:76-91), and the reverb
 *     plugin's Initialize and ConvolveFFT (REV.cpp:74-102, 172-213), cut out verbatim at build time by
 *     oracle/extract_ue_bodies.py into a scratch directory (never committed);
 *   - oracle/ue_shim/CoreMinimal.h for the engine types, and below the three engine services the path calls:
 *     the scene query, the random numbers and the actor location.
 * Nothing in this file restates the reference's algorithm except the eight lines of ProcessSourceAudio glue in
 * ref_ue_conv_process (REV.cpp:138-168 with the de-interleave FIX), marked there.
 */
#define private public          /* the BDPT methods are private members of UAudioRayTracingSubsystem (SUB.h:109-150) */
#define protected public
#include "CoreMinimal.h"
#include "AudioRayTracingSubsystem.h"
#include "FrequenSeeAudioComponent.h"
#include "CircularBuffer.h"
#include "kiss_fftr.h"
#undef private
#undef protected

#include <stdio.h>
#include <stdlib.h>

/* ---- shell of the reverb plugin: the reference's state members + the two extracted methods ------------------------- */
class USoundSubmix;
typedef void* FSoundEffectSubmixPtr;
struct FFrequenSeeAudioReverbSource { bool bApplyReflections = true; float PrevDuration = 0.f; };
struct FAudioPluginInitializationParams { uint32 NumSources = 0, NumOutputChannels = 0, SampleRate = 0, BufferLength = 0; };
class FFrequenSeeAudioReverbPlugin {
public:
    void Initialize(const FAudioPluginInitializationParams InitializationParams);
#include "rev_members.inc"
};

/* ---- the reference's bodies ------------------------------------------------------------------------------------------ */
#include "sub_bodies.inc"
#include "comp_bodies.inc"
#include "rev_bodies.inc"

/* ---- declared in the reference's headers, defined in files that are out of scope (glue, visualisation) ------- */
UAudioRayTracingSubsystem::UAudioRayTracingSubsystem() {}
void UAudioRayTracingSubsystem::Initialize(FSubsystemCollectionBase&) {}
void UAudioRayTracingSubsystem::Deinitialize() {}
void UAudioRayTracingSubsystem::Tick(float) {}
void UAudioRayTracingSubsystem::Visualize(FActiveSource&, int, float) {}          /* DEBUG_RAY_COUNT = 0: draws nothing */
void UFrequenSeeAudioComponent::OnRegister() {}
void UFrequenSeeAudioComponent::OnUnregister() {}
void UFrequenSeeAudioComponent::BeginPlay() {}
void UFrequenSeeAudioComponent::TickComponent(float, ELevelTick, FActorComponentTickFunction*) {}
void UAcousticGeometryComponent::OnRegister() {}
void UAcousticGeometryComponent::OnUnregister() {}

/* =========================================================================================================================
 * engine services
 * ========================================================================================================================= */
const FVector FVector::ZeroVector = FVector(0, 0, 0);

void ue_shim_check_failed(const char* what, const char* file, int line)
{
    fprintf(stderr, "check(%s) failed at %s:%d\n", what, file, line);
    abort();
}

static thread_local ue_shim_rng_state g_rng;
ue_shim_rng_state& ue_shim_rng() { return g_rng; }

static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static double u01(uint32_t r) { return (double)(r >> 8) * (1.0 / 16777216.0); }

/* GeneratePath reads the actor location once, first thing (SUB.cpp:287): a new subpath, hence a new random stream */
FVector AActor::GetActorLocation() const
{
    ue_shim_rng_state& s = g_rng;
    s.side = (this == s.side_actor[1]) ? 1u : 0u;
    s.g = s.index[s.side]++;
    s.bounce = 0;
    return Location;
}
/* Russian roulette draw (SUB.cpp:301): opens the Philox block of this bounce; the direction uses the same block */
float FMath::FRand()
{
    ue_shim_rng_state& s = g_rng;
    philox4x32_10((uint32_t)s.g, (uint32_t)(s.g >> 32), s.bounce, s.side, (uint32_t)s.seed, (uint32_t)(s.seed >> 32), s.r);
    ++s.bounce;
    return (float)u01(s.r[0]);
}
FVector FMath::VRand()                                             /* uniform on the sphere */
{
    const double u1 = u01(g_rng.r[1]), u2 = u01(g_rng.r[2]);
    const double z = 1.0 - 2.0 * u1, r = sqrt(fmax(0.0, 1.0 - z * z)), phi = 2.0 * M_PI * u2;
    return FVector(r * cos(phi), r * sin(phi), z);
}
FVector FMath::VRandCone(const FVector& N, float)                  /* FIX: cosine-weighted hemisphere about N */
{
    const double u1 = u01(g_rng.r[1]), u2 = u01(g_rng.r[2]);
    const double r = sqrt(u1), phi = 2.0 * M_PI * u2, lx = r * cos(phi), ly = r * sin(phi), lz = sqrt(1.0 - u1);
    const double sg = N.Z >= 0.0 ? 1.0 : -1.0, a = -1.0 / (sg + N.Z), b = N.X * N.Y * a;
    const FVector T(1.0 + sg * N.X * N.X * a, sg * b, -sg * N.X), B(b, sg + N.Y * N.Y * a, -N.Y);
    return T * lx + B * ly + N * lz;
}

bool UWorld::LineTraceSingleByObjectType(FHitResult& H, const FVector& Start, const FVector& End,
                                         const FCollisionObjectQueryParams&, const FCollisionQueryParams&)
{
    ++n_traces;
    const FVector seg = End - Start;
    const double len = seg.Size();
    if (!(len > 0.0)) return false;
    const double d[3] = {seg.X / len, seg.Y / len, seg.Z / len}, o[3] = {Start.X, Start.Y, Start.Z};
    double best = INFINITY; size_t bi = (size_t)-1;
    const size_t T = tri_actor.size();
    for (size_t t = 0; t < T; ++t) {                              /* Moller-Trumbore, two sided, t > 0 */
        const double *a = &v0[3 * t], *p = &e1[3 * t], *q = &e2[3 * t];
        const double pv[3] = {d[1] * q[2] - d[2] * q[1], d[2] * q[0] - d[0] * q[2], d[0] * q[1] - d[1] * q[0]};
        const double det = p[0] * pv[0] + p[1] * pv[1] + p[2] * pv[2];
        if (det == 0.0) continue;
        const double inv = 1.0 / det, tv[3] = {o[0] - a[0], o[1] - a[1], o[2] - a[2]};
        const double u = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) * inv;
        if (!(u >= 0.0 && u <= 1.0)) continue;
        const double qv[3] = {tv[1] * p[2] - tv[2] * p[1], tv[2] * p[0] - tv[0] * p[2], tv[0] * p[1] - tv[1] * p[0]};
        const double v = (d[0] * qv[0] + d[1] * qv[1] + d[2] * qv[2]) * inv;
        if (!(v >= 0.0 && u + v <= 1.0)) continue;
        const double tt = (q[0] * qv[0] + q[1] * qv[1] + q[2] * qv[2]) * inv;
        if (tt > 0.0 && tt < best) { best = tt; bi = t; }
    }
    if (bi == (size_t)-1 || !(best < len)) return false;
    H.ImpactPoint = FVector(o[0] + best * d[0], o[1] + best * d[1], o[2] + best * d[2]);
    FVector n(nrm[3 * bi], nrm[3 * bi + 1], nrm[3 * bi + 2]);
    if (n.X * d[0] + n.Y * d[1] + n.Z * d[2] > 0.0) n = n * -1.0;     /* the impact normal faces the incoming ray */
    H.ImpactNormal = n;
    H.Actor = Actors[tri_actor[bi]];
    return true;
}

/* =========================================================================================================================
 * C entry points
 * ========================================================================================================================= */
struct ref_ue_world {
    UWorld world;
    std::vector<AActor*> actors;
    std::vector<UAcousticGeometryComponent*> geo;
    std::vector<UAcousticMaterial*> mats;
    double unit;
    ~ref_ue_world()
    {
        for (auto p : actors) delete p;
        for (auto p : geo) delete p;
        for (auto p : mats) delete p;
    }
};

struct ref_ue_path_rec {                  /* one per path pair, in work-index order */
    uint32_t n_src_nodes, n_lis_nodes, connected, pad;
    float delay_s, gain;
    double src_end[3], lis_end[3];        /* world units */
};

extern "C" {

/* verts [T][3][3] float in METRES, scaled by `unit` world units per metre (1000: EvaluatePath's "/ 1000" then yields metres);
 * value2[m] = UAcousticMaterial::Absorption[2].Value of material m (the one band the reference reads, SUB.cpp:385) */
void* ref_ue_world_create(const float* verts, const uint32_t* tri_mat, uint64_t n_tris, const float* value2, uint32_t n_mats,
                          double unit)
{
    ref_ue_world* w = new ref_ue_world();
    w->unit = unit;
    for (uint32_t m = 0; m < n_mats; ++m) {
        UAcousticMaterial* mat = new UAcousticMaterial();
        mat->Absorption[2].Value = value2[m];
        UAcousticGeometryComponent* g = new UAcousticGeometryComponent();
        g->Material = mat;
        AActor* a = new AActor();
        a->Geometry = g; g->Owner = a;
        w->mats.push_back(mat); w->geo.push_back(g); w->actors.push_back(a);
    }
    w->world.Actors = w->actors;
    for (uint64_t t = 0; t < n_tris; ++t) {
        double p[3][3];
        for (int k = 0; k < 3; ++k) for (int a = 0; a < 3; ++a) p[k][a] = (double)verts[t * 9 + k * 3 + a] * unit;
        double e1[3], e2[3];
        for (int a = 0; a < 3; ++a) { e1[a] = p[1][a] - p[0][a]; e2[a] = p[2][a] - p[0][a]; }
        double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const double l = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (int a = 0; a < 3; ++a) {
            w->world.v0.push_back(p[0][a]); w->world.e1.push_back(e1[a]); w->world.e2.push_back(e2[a]);
            w->world.nrm.push_back(l > 0.0 ? n[a] / l : 0.0);
        }
        w->world.tri_actor.push_back(tri_mat[t]);
    }
    return w;
}
void ref_ue_world_destroy(void* h) { delete (ref_ue_world*)h; }

int ref_ue_used_ray_count(void) { return UAudioRayTracingSubsystem::USED_RAY_COUNT; }

struct scene_objs {
    UAudioRayTracingSubsystem sub;
    UFrequenSeeAudioComponent comp;
    AActor owner;
    APawn pawn;
    FActiveSource src;
    scene_objs(ref_ue_world* w, const float s[3], const float l[3], uint64_t seed, uint64_t g_first)
    {
        sub.World = &w->world;
        owner.Location = FVector(s[0] * w->unit, s[1] * w->unit, s[2] * w->unit);
        pawn.Location = FVector(l[0] * w->unit, l[1] * w->unit, l[2] * w->unit);
        owner.Audio = &comp; comp.Owner = &owner;
        sub.PlayerPawn = &pawn;
        src.AudioComp = &comp;
        ue_shim_rng_state& r = ue_shim_rng();
        r = ue_shim_rng_state();
        r.seed = seed; r.index[0] = r.index[1] = g_first;
        r.side_actor[0] = &owner; r.side_actor[1] = &pawn;
    }
};

/* ONE call of the reference's UpdateSource (SUB.cpp:128-195), unmodified, on a freshly constructed component:
 * USED_RAY_COUNT path pairs with work indices g_first ..; returns EnergyBuffer [NumBins] and ImpulseBuffer [2][NumSamples] */
int ref_ue_update_source(void* h, const float src[3], const float lis[3], uint64_t seed, uint64_t g_first,
                         float* energy_out, float* ir_out, uint64_t* n_traces)
{
    ref_ue_world* w = (ref_ue_world*)h;
    scene_objs o(w, src, lis, seed, g_first);
    w->world.n_traces = 0;
    o.sub.UpdateSource(o.src);
    if (o.comp.EnergyBuffer.Num() != o.comp.NumBins) return -1;
    memcpy(energy_out, o.comp.EnergyBuffer.GetData(), sizeof(float) * o.comp.NumBins);
    for (int c = 0; c < o.comp.NumChannels; ++c)
        memcpy(ir_out + (size_t)c * o.comp.NumSamples, o.comp.ImpulseBuffer[c].GetData(), sizeof(float) * o.comp.NumSamples);
    if (n_traces) *n_traces = w->world.n_traces;
    return 0;
}

/* GenerateFullPaths (SUB.cpp:201-233) for n pairs, then EvaluatePath (SUB.cpp:360-420) on every connected path:
 * per-pair records for localising a disagreement */
int ref_ue_paths(void* h, const float src[3], const float lis[3], uint64_t seed, uint64_t g_first, int n, ref_ue_path_rec* recs)
{
    ref_ue_world* w = (ref_ue_world*)h;
    scene_objs o(w, src, lis, seed, g_first);
    TArray<FSoundPath> F, B, Cn;
    o.sub.GenerateFullPaths(o.src, F, B, Cn, n);
    if (F.Num() != n || B.Num() != n) return -1;
    int k = 0;
    for (int i = 0; i < n; ++i) {
        ref_ue_path_rec& r = recs[i];
        memset(&r, 0, sizeof(r));
        r.n_src_nodes = (uint32_t)F[i].Nodes.Num(); r.n_lis_nodes = (uint32_t)B[i].Nodes.Num();
        const FVector fe = F[i].Nodes.Last().Position, be = B[i].Nodes.Last().Position;
        r.src_end[0] = fe.X; r.src_end[1] = fe.Y; r.src_end[2] = fe.Z;
        r.lis_end[0] = be.X; r.lis_end[1] = be.Y; r.lis_end[2] = be.Z;
        if (k < Cn.Num() && Cn[k].Nodes.Num() == F[i].Nodes.Num() + B[i].Nodes.Num() &&
            FVector::Dist(Cn[k].ForwardConnectionPos, fe) == 0.0 && FVector::Dist(Cn[k].BackwardConnectionPos, be) == 0.0) {
            const FPathEnergyResult e = o.sub.EvaluatePath(Cn[k]);
            r.connected = 1; r.delay_s = e.DelaySeconds; r.gain = e.Gain;
            ++k;
        }
    }
    return k == Cn.Num() ? 0 : -2;
}

/* EvaluatePath (SUB.cpp:360-420) on an explicit node list: pos [n][3] world units, value2[i] < 0 = node without material */
int ref_ue_evaluate_path(const double* pos, const float* value2, const float* prob, int n, float* delay_out, float* gain_out)
{
    UAudioRayTracingSubsystem sub;
    std::vector<UAcousticMaterial> mats((size_t)n);
    std::vector<UAcousticGeometryComponent> geo((size_t)n);
    FSoundPath P;
    for (int i = 0; i < n; ++i) {
        FSoundPathNode nd;
        nd.Position = FVector(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        nd.Probability = prob[i];
        if (value2[i] >= 0.0f) {
            mats[(size_t)i].Absorption[2].Value = value2[i];
            geo[(size_t)i].Material = &mats[(size_t)i];
            nd.Material = &geo[(size_t)i];
        }
        P.Nodes.Add(nd);
    }
    const FPathEnergyResult e = sub.EvaluatePath(P);
    *delay_out = e.DelaySeconds; *gain_out = e.Gain;
    return 0;
}

/* FlushEnergyBuffer + AddEnergyAtDelay (COMP.h:76-91) for m (delay, energy) pairs on a fresh component */
int ref_ue_add_energy(const float* delay, const float* energy, int m, float* buf_out)
{
    UFrequenSeeAudioComponent comp;
    comp.FlushEnergyBuffer();
    for (int i = 0; i < m; ++i) comp.AddEnergyAtDelay(delay[i], energy[i]);
    memcpy(buf_out, comp.EnergyBuffer.GetData(), sizeof(float) * comp.NumBins);
    return comp.NumBins;
}

/* ReconstructImpulseResponse (COMP.cpp:320-380) from a given EnergyBuffer [NumBins]; ir_out [NumChannels][NumSamples] */
int ref_ue_reconstruct_ir(const float* energy, float* ir_out)
{
    UFrequenSeeAudioComponent comp;
    comp.FlushEnergyBuffer();
    memcpy(comp.EnergyBuffer.GetData(), energy, sizeof(float) * comp.NumBins);
    comp.ReconstructImpulseResponse();
    for (int c = 0; c < comp.NumChannels; ++c)
        memcpy(ir_out + (size_t)c * comp.NumSamples, comp.ImpulseBuffer[c].GetData(), sizeof(float) * comp.NumSamples);
    return (int)ceilf(comp.BinDuration * comp.SampleRate);          /* the NumSamplesPerBin the reference computes: 49 */
}

/* ---- reverb: the reference's Initialize + ConvolveFFT + FCircularAudioBuffer + KissFFT -------------------------------- */
struct ref_ue_conv { FFrequenSeeAudioReverbPlugin plugin; TArray<TArray<float>> ir; };

void* ref_ue_conv_create(int sample_rate, int frame)
{
    ref_ue_conv* c = new ref_ue_conv();
    FAudioPluginInitializationParams p;
    p.SampleRate = (uint32)sample_rate; p.BufferLength = (uint32)frame; p.NumSources = 1; p.NumOutputChannels = 2;
    c->plugin.Initialize(p);                                                     /* REV.cpp:74-102 */
    c->ir.SetNum(2);
    c->ir[0].Init(0.0f, sample_rate); c->ir[1].Init(0.0f, sample_rate);
    return c;
}
void ref_ue_conv_destroy(void* h)
{
    ref_ue_conv* c = (ref_ue_conv*)h;
    kiss_fftr_free(c->plugin.ForwardCfg); kiss_fftr_free(c->plugin.InverseCfg);  /* the plugin's destructor, REV.cpp:39-43 */
    delete c;
}
int ref_ue_conv_fft_size(void* h) { return ((ref_ue_conv*)h)->plugin.FFTSize; }
void ref_ue_conv_set_ir(void* h, const float* ir)                                /* [2][sample_rate] */
{
    ref_ue_conv* c = (ref_ue_conv*)h;
    for (int ch = 0; ch < 2; ++ch) memcpy(c->ir[ch].GetData(), ir + (size_t)ch * c->ir[ch].Num(), sizeof(float) * c->ir[ch].Num());
}
/* One audio callback, interleaved stereo in/out.  The lines between the markers restate ProcessSourceAudio
 * (REV.cpp:138-168) -- the only restated glue in this file -- with the FIX of SURVEY A8: the newest block is
 * de-interleaved into the two tails (REV.cpp:147-148 copies the interleaved buffer into both). */
void ref_ue_conv_process(void* h, const float* in, float* out, int clamp)
{
    ref_ue_conv* c = (ref_ue_conv*)h;
    FFrequenSeeAudioReverbPlugin& P = c->plugin;
    /* >>> restated glue */
    const int TailSize = P.AudioTailBufferLeft.GetSize();
    P.AudioTailBufferLeft.GetLastSamples(P.CurrAudioTailLeft, TailSize);         /* REV.cpp:141-142 */
    P.AudioTailBufferRight.GetLastSamples(P.CurrAudioTailRight, TailSize);
    P.AudioTailBufferLeft.AddSamples(in, P.FrameSize, 0, 2);                     /* REV.cpp:144-145 */
    P.AudioTailBufferRight.AddSamples(in, P.FrameSize, 1, 2);
    for (int i = 0; i < P.FrameSize; ++i) {                                      /* REV.cpp:147-148, FIX: de-interleave */
        P.CurrAudioTailLeft[TailSize + i] = in[2 * i];
        P.CurrAudioTailRight[TailSize + i] = in[2 * i + 1];
    }
    P.ConvolveFFT(c->ir[0], P.CurrAudioTailLeft, P.ConvOutputLeft);              /* REV.cpp:151, 154: the reference's body */
    P.ConvolveFFT(c->ir[1], P.CurrAudioTailRight, P.ConvOutputRight);
    for (int i = 0; i < P.FrameSize; ++i) {                                      /* REV.cpp:162-168, MixAlpha = 1 */
        const float l = P.ConvOutputLeft[TailSize + i], r = P.ConvOutputRight[TailSize + i];
        out[2 * i] = clamp ? FMath::Clamp(l, -1.0f, 1.0f) : l;
        out[2 * i + 1] = clamp ? FMath::Clamp(r, -1.0f, 1.0f) : r;
    }
    /* <<< restated glue */
}

}  /* extern "C" */
