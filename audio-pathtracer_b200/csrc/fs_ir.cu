// fs_ir.cu -- energy histogram -> impulse response.
//
// Replaces UFrequenSeeAudioComponent::ReconstructImpulseResponse (COMP.cpp:320-380):
//   k_energy : E_k = (sum_b hist[b][k]) * 2^-32 / n_paths   (1/N of SUB.cpp:164 applied after the
//              integer reduction), a_k = E_k / sqrt(E_k * sqrt(4 pi)) if |E_k| >= threshold else 0
//              (COMP.cpp:339-346)
//   k_ir     : sample j of bin k: (1-w) a_{k-1} + w a_k, w = j/48 (COMP.cpp:347-363; FIX: 48
//              samples per bin, the reference's ceil(0.001f*48000.f) is 49); one-pole low-pass
//              y_i = 0.25 x_i + 0.75 y_{i-1}, y_0 = x_0 (COMP.cpp:366-375) evaluated per output
//              sample over a 128-tap warm-up window (0.75^128 ~ 1e-16, below float resolution);
//              output = filtered, un-normalised (COMP.cpp:377-378); every channel gets the same
//              mono IR (COMP.cpp:327-330).
// Algorithmic bytes per update: read B*K*8, write C*sample_rate*4.
#include "fs_internal.h"

namespace {

__global__ void k_energy(const unsigned long long* __restrict__ hist, uint32_t n_bands, uint32_t n_bins,
                         double inv_scale, const float* __restrict__ energy_in, float threshold,
                         float* __restrict__ amp)
{
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_bins) return;
    float e;
    if (energy_in) e = energy_in[k];
    else {
        unsigned long long s = 0;
        for (uint32_t b = 0; b < n_bands; ++b) s += hist[(size_t)b * n_bins + k];
        e = (float)(((double)s * (1.0 / 4294967296.0)) * inv_scale);
    }
    const float Pi4 = sqrtf(4.0f * FS_PI);
    float a = 0.0f;
    if (fabsf(e) >= threshold) a = e / sqrtf(e * Pi4);
    amp[k] = a;
}

__device__ __forceinline__ float raw_sample(const float* __restrict__ amp, uint32_t i, uint32_t spb, uint32_t n_bins)
{
    uint32_t bin = i / spb, jj = i - bin * spb;
    if (bin >= n_bins) return 0.0f;
    float e = amp[bin];
    float pe = bin ? amp[bin - 1] : e;
    float w = (float)jj / (float)spb;
    return (1.0f - w) * pe + w * e;
}

constexpr int IR_WINDOW = 128;

__global__ void k_ir(const float* __restrict__ amp, uint32_t n_bins, uint32_t spb, uint32_t n_samples,
                     uint32_t n_channels, float a, float* __restrict__ ir)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples) return;
    uint32_t j0 = i >= (uint32_t)IR_WINDOW ? i - IR_WINDOW : 0u;
    float y = raw_sample(amp, j0, spb, n_bins);
    if (j0 > 0) y = a * y;
    for (uint32_t j = j0 + 1; j <= i; ++j) y = a * raw_sample(amp, j, spb, n_bins) + (1.0f - a) * y;
    for (uint32_t c = 0; c < n_channels; ++c) ir[(size_t)c * n_samples + i] = y;
}

}  // namespace

cudaError_t fs_ir_build(fs_ctx* ctx, const unsigned long long* d_hist_src, uint64_t n_paths,
                        const float* d_energy_in, float* d_ir_out)
{
    const fs_config& c = ctx->cfg;
    const uint32_t spb = (uint32_t)((double)c.bin_ms * 1e-3 * c.sample_rate + 0.5);
    const double inv_scale = n_paths ? 1.0 / (double)n_paths : 0.0;
    k_energy<<<(c.n_bins + 255) / 256, 256, 0, ctx->stream>>>(d_hist_src, c.n_bands, c.n_bins, inv_scale,
                                                              d_energy_in, c.ir_threshold, ctx->d_amp);
    k_ir<<<(c.sample_rate + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_amp, c.n_bins, spb, c.sample_rate,
                                                                c.n_channels, c.ir_lowpass, d_ir_out);
    ctx->stats.kernel_launches += 2;
    return cudaGetLastError();
}
