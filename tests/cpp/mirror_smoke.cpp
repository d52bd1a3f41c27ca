// Drives the C++ mirror of the reference interface (include/frequensee.hpp) the way the reference's own call
// sites do: RegisterGeometry / RegisterSource / ForceUpdateSources (SUB.cpp:883-886), GetImpulseResponse
// (REV.cpp:136), Initialize / OnInitSource / ProcessSourceAudio (REV.cpp:74-170).
// Prints one line of numbers that tests/test_cpp_mirror.py compares with the oracle.  "nogpu" mode only checks
// that context creation fails loudly without a device.
#include <cstdio>
#include <cstring>
#include <vector>
#include "frequensee.hpp"

using namespace FrequenSee;

int main(int argc, char** argv)
{
    auto ctx = FContext::Create();
    if (argc > 1 && !strcmp(argv[1], "nogpu")) {
        if (ctx) { printf("UNEXPECTED: context created without a GPU\n"); return 1; }
        printf("no-gpu: %s\n", fs_last_error(nullptr));
        return strstr(fs_last_error(nullptr), "no CPU fallback") ? 0 : 2;
    }
    if (!ctx) { printf("fs_create failed: %s\n", fs_last_error(nullptr)); return 3; }
    // shoebox 7 x 5 x 3 m, one material per wall (config 1 of BASELINE.json)
    const float X = 7, Y = 5, Z = 3;
    const float quads[6][4][3] = {
        {{0,0,0},{0,Y,0},{0,Y,Z},{0,0,Z}}, {{X,0,0},{X,Y,0},{X,Y,Z},{X,0,Z}},
        {{0,0,0},{X,0,0},{X,0,Z},{0,0,Z}}, {{0,Y,0},{X,Y,0},{X,Y,Z},{0,Y,Z}},
        {{0,0,0},{X,0,0},{X,Y,0},{0,Y,0}}, {{0,0,Z},{X,0,Z},{X,Y,Z},{0,Y,Z}}};
    const float alpha[6] = {0.02f, 0.05f, 0.12f, 0.07f, 0.37f, 0.75f};
    UAudioRayTracingSubsystem sub(ctx);
    for (int f = 0; f < 6; ++f) {
        float tris[2][3][3];
        const int idx[2][3] = {{0, 1, 2}, {0, 2, 3}};
        for (int t = 0; t < 2; ++t) for (int v = 0; v < 3; ++v) memcpy(tris[t][v], quads[f][idx[t][v]], 12);
        FAcousticMaterial m; m.Absorption.assign(8, alpha[f]);
        sub.RegisterGeometry(&tris[0][0][0], 2, m);
    }
    UFrequenSeeAudioComponent comp(ctx, /*SourceId=*/2, FVector3f{1.5f, 1.2f, 1.0f});
    comp.RaycastsPerTick = 4096; comp.RaycastBounces = 8;
    sub.RegisterSource(&comp);
    sub.SetPlayerPawnLocation(FVector3f{5.0f, 3.5f, 1.6f});
    sub.Seed = 0x5EED;
    FFrequenSeeAudioReverbPlugin rev(ctx);
    rev.Initialize(FAudioPluginInitializationParams{1, 2, 48000, 1024});
    rev.OnInitSource(2, 2);
    sub.ForceUpdateSources();
    if (ctx->LastStatus != FS_OK) { printf("update failed: %s\n", ctx->LastError().c_str()); return 4; }
    auto& ir = comp.GetImpulseResponse();
    double e = 0; int peak = 0;
    for (int i = 0; i < 48000; ++i) { e += (double)ir[0][i] * ir[0][i]; if (ir[0][i] > ir[0][peak]) peak = i; }
    std::vector<float> in(2048, 0.0f), out(2048, 0.0f);
    in[0] = 1.0f; in[1] = 0.5f;                         // unit impulse L, half impulse R at frame 0
    rev.ProcessSourceAudio(comp, in.data(), out.data());
    if (ctx->LastStatus != FS_OK) { printf("process failed: %s\n", ctx->LastError().c_str()); return 5; }
    // impulse in -> first block of the IR out
    double d = 0, n = 0;
    for (int i = 0; i < 1024; ++i) {
        double a = out[2 * i] - ir[0][i], b = out[2 * i + 1] - 0.5 * ir[1][i];
        d += a * a + b * b; n += (double)ir[0][i] * ir[0][i] * 1.25;
    }
    // seam 1 taken literally: AddEnergyAtDelay on the float EnergyBuffer, then ReconstructImpulseResponse
    UFrequenSeeAudioComponent comp2(ctx, 3, FVector3f{0, 0, 0});
    comp2.FlushEnergyBuffer();
    comp2.AddEnergyAtDelay(0.0105f, 0.04f);             // -> bin 10
    comp2.AddEnergyAtDelay(5.0f, 0.01f);                // late -> clamps into bin 999
    comp2.ReconstructImpulseResponse();
    comp.bApplyReverb = false;
    std::vector<float> pass(2048, 0.0f);
    rev.ProcessSourceAudio(comp, in.data(), pass.data());
    printf("connected=%llu ir_energy=%.9g ir_peak=%d conv_rel=%.3g bin10=%.9g bin999=%.9g ir2_sample=%.9g passthrough=%d\n",
           (unsigned long long)sub.LastConnected, e, peak, n > 0 ? std::sqrt(d / n) : -1.0, comp2.EnergyBuffer[10],
           comp2.EnergyBuffer[999], comp2.GetImpulseResponse()[1][10 * 48 + 47], (int)(pass[0] == 1.0f && pass[1] == 0.5f));
    return 0;
}
