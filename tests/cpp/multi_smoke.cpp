// Several GPUs behind the C-ABI, driven from plain C++ (no Python, no torch, no collective library): fs_multi_* shards one
// IR update over the devices and reduces the histograms on device 0 by peer stores.  The result must equal one device's
// fs_trace bit for bit.  usage: multi_smoke [n_devices] [same]   ("same": every context on device 0 -- a one-GPU box)
// Prints one JSON line; exit code 0 only if every comparison is exact.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "frequensee.h"

static void add_box(std::vector<float>& v, std::vector<uint32_t>& m, const float lo[3], const float hi[3], uint32_t mat)
{
    const float c[8][3] = {{lo[0], lo[1], lo[2]}, {hi[0], lo[1], lo[2]}, {hi[0], hi[1], lo[2]}, {lo[0], hi[1], lo[2]},
                           {lo[0], lo[1], hi[2]}, {hi[0], lo[1], hi[2]}, {hi[0], hi[1], hi[2]}, {lo[0], hi[1], hi[2]}};
    const int q[6][4] = {{0, 3, 7, 4}, {1, 2, 6, 5}, {0, 1, 5, 4}, {3, 2, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
    for (int f = 0; f < 6; ++f) {
        const int t[2][3] = {{q[f][0], q[f][1], q[f][2]}, {q[f][0], q[f][2], q[f][3]}};
        for (int k = 0; k < 2; ++k) {
            for (int a = 0; a < 3; ++a) for (int x = 0; x < 3; ++x) v.push_back(c[t[k][a]][x]);
            m.push_back(mat == 0xffffffffu ? (uint32_t)f : mat);
        }
    }
}

int main(int argc, char** argv)
{
    int n_dev = argc > 1 ? atoi(argv[1]) : 2;
    const bool same = argc > 2 && !strcmp(argv[2], "same");
    std::vector<float> verts; std::vector<uint32_t> mats;
    const float rlo[3] = {0, 0, 0}, rhi[3] = {12, 9, 3.6f};
    add_box(verts, mats, rlo, rhi, 0xffffffffu);                       // the room, one material per wall
    unsigned s = 12345u;
    auto rnd = [&] { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.0f; };
    for (int i = 0; i < 60; ++i) {                                     // furniture
        const float w = 0.3f + rnd(), d = 0.3f + rnd(), h = 0.3f + 1.8f * rnd();
        const float x = 0.5f + rnd() * (11.0f - w), y = 0.5f + rnd() * (8.0f - d);
        if (x < 3.0f && y < 3.0f) continue;                            // keep the source clear
        const float lo[3] = {x, y, 0.0f}, hi[3] = {x + w, y + d, h};
        add_box(verts, mats, lo, hi, 6u + (uint32_t)(i & 1));
    }
    std::vector<float> ab(8 * 8);
    for (int m = 0; m < 8; ++m) for (int b = 0; b < 8; ++b) ab[m * 8 + b] = 0.03f + 0.1f * m + 0.01f * b;
    const float src[2][3] = {{2.0f, 1.8f, 1.4f}, {1.2f, 2.4f, 1.1f}}, lis[3] = {9.5f, 6.5f, 1.6f};
    const uint64_t n_paths = 150000; const uint32_t depth = 12; const uint64_t seed = 42;
    fs_config cfg; fs_default_config(&cfg);
    const size_t hn = 2 * (size_t)cfg.n_bands * cfg.n_bins;

    // one device, the single-context API
    fs_ctx* one = nullptr;
    cfg.device = 0;
    if (fs_create(&cfg, &one) != FS_OK) { printf("fs_create: %s\n", fs_last_error(nullptr)); return 3; }
    if (fs_scene_set_triangles(one, verts.data(), mats.data(), mats.size()) || fs_scene_set_materials(one, ab.data(), 8, 8) || fs_scene_commit(one)) {
        printf("scene: %s\n", fs_last_error(one)); return 3;
    }
    std::vector<uint64_t> h1(hn), hm(hn), hm2(hn);
    if (fs_trace(one, &src[0][0], 2, lis, n_paths, depth, seed, h1.data())) { printf("fs_trace: %s\n", fs_last_error(one)); return 3; }
    std::vector<float> ir1(2 * 2 * 48000), irm(2 * 2 * 48000);
    fs_build_ir_all(one, 2, ir1.data());

    // n devices, fs_multi
    std::vector<int> devs(n_dev);
    for (int i = 0; i < n_dev; ++i) devs[i] = same ? 0 : i;
    fs_multi* m = nullptr;
    cfg.device = -1;
    if (fs_multi_create(&cfg, devs.data(), (uint32_t)n_dev, &m) != FS_OK) { printf("fs_multi_create: %s\n", fs_multi_last_error()); return 4; }
    if (fs_multi_scene_set_triangles(m, verts.data(), mats.data(), mats.size()) || fs_multi_scene_set_materials(m, ab.data(), 8, 8) ||
        fs_multi_scene_commit(m)) { printf("multi scene: %s\n", fs_multi_last_error()); return 4; }
    if (fs_multi_trace(m, &src[0][0], 2, lis, n_paths, depth, seed, hm.data())) { printf("fs_multi_trace: %s\n", fs_multi_last_error()); return 4; }
    // asynchronous form: three updates back to back (staging slots reused), IRs built on context 0 without a host sync between
    float total = 0.f, red = 0.f;
    for (int it = 0; it < 3; ++it)
        if (fs_multi_trace(m, &src[0][0], 2, lis, n_paths, depth, seed + 1 + it, nullptr)) { printf("fs_multi_trace: %s\n", fs_multi_last_error()); return 4; }
    if (fs_multi_trace(m, &src[0][0], 2, lis, n_paths, depth, seed, nullptr)) return 4;
    if (fs_build_ir_all(fs_multi_context(m, 0), 2, irm.data())) { printf("build_ir: %s\n", fs_last_error(fs_multi_context(m, 0))); return 4; }
    fs_multi_last_ms(m, &total, &red);
    if (fs_get_histogram(fs_multi_context(m, 0), hm2.data())) return 4;
    // timing: K updates, wall clock around enqueue + synchronise
    const int K = 10;
    auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < K; ++it) fs_multi_trace(m, &src[0][0], 2, lis, n_paths, depth, seed + 100 + it, nullptr);
    fs_multi_synchronize(m);
    const double ms_multi = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / K;
    t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < K; ++it) fs_trace(one, &src[0][0], 2, lis, n_paths, depth, seed + 100 + it, nullptr);
    fs_synchronize(one);
    const double ms_one = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / K;
    const bool eq1 = memcmp(h1.data(), hm.data(), hn * 8) == 0, eq2 = memcmp(h1.data(), hm2.data(), hn * 8) == 0;
    const bool eq3 = memcmp(ir1.data(), irm.data(), ir1.size() * 4) == 0;
    uint64_t sum = 0; for (uint64_t x : h1) sum += x;
    printf("{\"devices\": %d, \"same_device\": %s, \"triangles\": %zu, \"pairs_per_update\": %llu, \"hist_sum\": %llu, "
           "\"multi_equals_single\": %s, \"async_multi_equals_single\": %s, \"ir_equal\": %s, \"ms_per_update_multi\": %.3f, "
           "\"ms_per_update_single\": %.3f, \"device_ms_last_update\": %.3f, \"device_ms_wait_and_reduce\": %.3f}\n",
           n_dev, same ? "true" : "false", mats.size(), (unsigned long long)(2 * n_paths), (unsigned long long)sum,
           eq1 ? "true" : "false", eq2 ? "true" : "false", eq3 ? "true" : "false", ms_multi, ms_one, total, red);
    fs_multi_destroy(m);
    fs_destroy(one);
    return (eq1 && eq2 && eq3 && sum > 0) ? 0 : 1;
}
