// fs_conv.cu -- uniformly partitioned overlap-save FFT convolution of dry audio with the IR.
//
// Replaces FFrequenSeeAudioReverbPlugin::ProcessSourceAudio + ConvolveFFT (REV.cpp:118-213) and
// FCircularAudioBuffer (CIRC.cpp:43-75).  The reference transforms a 49 023-sample history and
// the whole 48 000-tap IR with three 65 536-point KissFFTs per channel per callback; here the IR
// is cut into P = ceil(ir_len / Bk) partitions of one callback block (Bk = 1024) whose 2Bk-point
// spectra H_p are computed once per IR refresh, the input spectra X_k live in a frequency-domain
// delay line, and one callback costs one 2Bk FFT, a P-term spectral multiply-accumulate and one
// inverse FFT:   y_block = IFFT( sum_p X_{k-p} * H_p )[Bk .. 2Bk)
// which equals "history (*) current IR" on the newest block, exactly the reference's result
// (history initially zero, CIRC.cpp:15-21).  FIX: channels are de-interleaved (REV.cpp:147-148
// copies interleaved samples into both channel tails).
//
// Kernels: k_ir_spectra (grid P x C), k_conv_blocks (one CTA per channel, loops over blocks,
// shared-memory Stockham radix-2 FFT).  Algorithmic bytes per block and channel:
// (P+1)*(Bk+1)*8*2 (H_p and X_{k-p} spectra) + 2*Bk*4*2.
#include "fs_internal.h"

#include <math.h>
#include <vector>

namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-block Stockham autosort radix-2 FFT of n = 2^logn points between two shared buffers.
// tw[m] = exp(-2 pi i m / tw_n), m < tw_n/2.  Returns the buffer holding the result.
__device__ float2* fft_stockham(float2* a, float2* b, uint32_t n, const float2* __restrict__ tw,
                                uint32_t tw_n, bool inverse)
{
    const uint32_t half = n >> 1;
    float2* in = a; float2* out = b;
    for (uint32_t ns = 1; ns < n; ns <<= 1) {
        const uint32_t tstep = tw_n / (2u * ns);
        for (uint32_t j = threadIdx.x; j < half; j += blockDim.x) {
            const uint32_t k = j & (ns - 1u);
            float2 w = tw[k * tstep];
            if (inverse) w.y = -w.y;
            const float2 x0 = in[j];
            const float2 x1 = cmul(in[j + half], w);
            const uint32_t j0 = ((j - k) << 1) + k;
            out[j0] = make_float2(x0.x + x1.x, x0.y + x1.y);
            out[j0 + ns] = make_float2(x0.x - x1.x, x0.y - x1.y);
        }
        __syncthreads();
        float2* t = in; in = out; out = t;
    }
    return in;
}

// H[p][c][f] = FFT_2Bk( ir[c][p*Bk .. (p+1)*Bk) zero-padded )
__global__ void k_ir_spectra(const float* __restrict__ ir, uint32_t ir_len, uint32_t bk, uint32_t n_ch,
                             const float2* __restrict__ tw, uint32_t tw_n, float2* __restrict__ H)
{
    extern __shared__ float2 sm[];
    const uint32_t n = 2u * bk, nf = bk + 1u;
    float2* a = sm; float2* b = sm + n;
    const uint32_t p = blockIdx.x, c = blockIdx.y;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t src = p * bk + i;
        float v = (i < bk && src < ir_len) ? ir[(size_t)c * ir_len + src] : 0.0f;
        a[i] = make_float2(v, 0.0f);
    }
    __syncthreads();
    float2* r = fft_stockham(a, b, n, tw, tw_n, false);
    float2* dst = H + ((size_t)p * n_ch + c) * nf;
    for (uint32_t f = threadIdx.x; f < nf; f += blockDim.x) dst[f] = r[f];
}

// the same for up to FS_PTR_TABLE sources in one launch: tab.p[s] = device IR, tab.q[s] = destination H of source s
__global__ void k_ir_spectra_multi(fs_ptr_table tab, uint32_t ir_len, uint32_t bk, uint32_t n_ch,
                                   const float2* __restrict__ tw, uint32_t tw_n)
{
    extern __shared__ float2 sm[];
    const uint32_t n = 2u * bk, nf = bk + 1u;
    float2* a = sm; float2* b = sm + n;
    const uint32_t p = blockIdx.x, c = blockIdx.y;
    const float* ir = (const float*)tab.p[blockIdx.z];
    float2* H = (float2*)tab.q[blockIdx.z];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t src = p * bk + i;
        float v = (i < bk && src < ir_len) ? ir[(size_t)c * ir_len + src] : 0.0f;
        a[i] = make_float2(v, 0.0f);
    }
    __syncthreads();
    float2* r = fft_stockham(a, b, n, tw, tw_n, false);
    float2* dst = H + ((size_t)p * n_ch + c) * nf;
    for (uint32_t f = threadIdx.x; f < nf; f += blockDim.x) dst[f] = r[f];
}

struct conv_args {
    float2* fdl;            // [P][C][NF]
    const float2* H;        // [P][C][NF]
    float* prev;            // [C][Bk]
    const float* in;        // [n_blocks][Bk][C] interleaved
    float* out;             // [n_blocks][Bk][C]
    const float2* tw;
    uint32_t bk, n_part, n_ch, n_blocks, head, tw_n;
    float wet; int clamp;
};

__global__ void k_conv_blocks(conv_args a)
{
    extern __shared__ float2 sm[];
    const uint32_t bk = a.bk, n = 2u * bk, nf = bk + 1u, C = a.n_ch, P = a.n_part;
    float2* s0 = sm; float2* s1 = sm + n;
    const uint32_t c = blockIdx.x;
    uint32_t head = a.head;
    float* prev = a.prev + (size_t)c * bk;
    for (uint32_t blk = 0; blk < a.n_blocks; ++blk) {
        const float* in = a.in + (size_t)blk * bk * C;
        float* out = a.out + (size_t)blk * bk * C;
        // window = previous block ++ current block (de-interleaved)
        for (uint32_t i = threadIdx.x; i < bk; i += blockDim.x) {
            const float cur = in[(size_t)i * C + c];
            s0[i] = make_float2(prev[i], 0.0f);
            s0[bk + i] = make_float2(cur, 0.0f);
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < bk; i += blockDim.x) prev[i] = s0[bk + i].x;
        float2* X = fft_stockham(s0, s1, n, a.tw, a.tw_n, false);
        float2* Y = (X == s0) ? s1 : s0;
        float2* slot = a.fdl + ((size_t)head * C + c) * nf;
        // spectral multiply-accumulate over the partitions; X_k itself comes from shared memory
        for (uint32_t f = threadIdx.x; f < nf; f += blockDim.x) {
            const float2 x = X[f];
            slot[f] = x;
            float2 acc = cmul(x, a.H[((size_t)0 * C + c) * nf + f]);
            uint32_t s = head;
            for (uint32_t p = 1; p < P; ++p) {
                s = s ? s - 1u : P - 1u;
                const float2 xv = a.fdl[((size_t)s * C + c) * nf + f];
                const float2 hv = a.H[((size_t)p * C + c) * nf + f];
                acc.x += xv.x * hv.x - xv.y * hv.y;
                acc.y += xv.x * hv.y + xv.y * hv.x;
            }
            Y[f] = acc;
            if (f > 0 && f < bk) Y[n - f] = make_float2(acc.x, -acc.y);
        }
        __syncthreads();
        float2* X2 = (Y == s0) ? s1 : s0;
        float2* y = fft_stockham(Y, X2, n, a.tw, a.tw_n, true);
        const float scale = 1.0f / (float)n;
        for (uint32_t i = threadIdx.x; i < bk; i += blockDim.x) {
            float v = y[bk + i].x * scale;
            if (a.clamp) v = fminf(fmaxf(v, -1.0f), 1.0f);               // REV.cpp:162-168
            const float dry = in[(size_t)i * C + c];
            out[(size_t)i * C + c] = v * a.wet + dry * (1.0f - a.wet);   // MixAlpha, REV.cpp:161
        }
        __syncthreads();
        head = (head + 1u == P) ? 0u : head + 1u;
    }
}

__global__ void k_rfft(const float* __restrict__ in, uint32_t n, const float2* __restrict__ tw, uint32_t tw_n,
                       float2* __restrict__ out)
{
    extern __shared__ float2 sm[];
    float2* a = sm; float2* b = sm + n;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) a[i] = make_float2(in[i], 0.0f);
    __syncthreads();
    float2* r = fft_stockham(a, b, n, tw, tw_n, false);
    for (uint32_t f = threadIdx.x; f <= n / 2; f += blockDim.x) out[f] = r[f];
}

size_t spectra_elems(const fs_ctx* ctx) { return (size_t)ctx->n_part * ctx->cfg.n_channels * ctx->n_freq; }

}  // namespace

cudaError_t fs_conv_setup(fs_ctx* ctx)
{
    const fs_config& c = ctx->cfg;
    ctx->fft_n = 2u * c.conv_block;
    ctx->n_freq = c.conv_block + 1u;
    ctx->n_part = (c.sample_rate + c.conv_block - 1u) / c.conv_block;   // IRSize = SampleRate * 1 s (REV.cpp:79)
    // twiddles for a 4096-point table (covers every n <= 4096 by striding), computed in double
    const uint32_t tn = ctx->fft_n > 4096u ? ctx->fft_n : 4096u;
    std::vector<float2> tw(tn / 2);
    for (uint32_t m = 0; m < tn / 2; ++m) {
        double ang = -2.0 * 3.14159265358979323846 * (double)m / (double)tn;
        tw[m] = make_float2((float)cos(ang), (float)sin(ang));
    }
    cudaError_t e;
    if ((e = cudaMalloc(&ctx->d_twiddle, sizeof(float2) * tw.size())) != cudaSuccess) return e;
    if ((e = cudaMemcpy(ctx->d_twiddle, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    const size_t smem = sizeof(float2) * 2 * ctx->fft_n;
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(k_ir_spectra, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_conv_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    cudaFuncSetAttribute(k_rfft, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    return cudaSuccess;
}

void fs_conv_teardown(fs_ctx* ctx)
{
    for (uint32_t s = 0; s < ctx->conv_cap; ++s) fs_conv_source_free(ctx, s);
    delete[] ctx->conv; ctx->conv = nullptr; ctx->conv_cap = 0;
    cudaFree(ctx->d_twiddle); ctx->d_twiddle = nullptr;
    cudaFree(ctx->d_conv_in); cudaFree(ctx->d_conv_out); ctx->d_conv_in = ctx->d_conv_out = nullptr;
    if (ctx->h_pin_in) cudaFreeHost(ctx->h_pin_in);
    if (ctx->h_pin_out) cudaFreeHost(ctx->h_pin_out);
    ctx->h_pin_in = ctx->h_pin_out = nullptr;
}

static cudaError_t ensure_sources(fs_ctx* ctx, uint32_t source)
{
    if (source < ctx->conv_cap) return cudaSuccess;
    uint32_t ncap = ctx->conv_cap ? ctx->conv_cap : 4;
    while (ncap <= source) ncap *= 2;
    fs_conv_source* n = new fs_conv_source[ncap];
    for (uint32_t i = 0; i < ncap; ++i) {
        n[i].active = false; n[i].fdl = nullptr; n[i].H[0] = n[i].H[1] = nullptr; n[i].prev = nullptr;
        n[i].ir = nullptr; n[i].h_published.store(0); n[i].h_valid = 0; n[i].head = 0;
    }
    for (uint32_t i = 0; i < ctx->conv_cap; ++i) {
        n[i].active = ctx->conv[i].active; n[i].fdl = ctx->conv[i].fdl; n[i].H[0] = ctx->conv[i].H[0];
        n[i].H[1] = ctx->conv[i].H[1]; n[i].prev = ctx->conv[i].prev; n[i].ir = ctx->conv[i].ir;
        n[i].h_published.store(ctx->conv[i].h_published.load()); n[i].h_valid = ctx->conv[i].h_valid;
        n[i].head = ctx->conv[i].head;
    }
    delete[] ctx->conv;
    ctx->conv = n; ctx->conv_cap = ncap;
    return cudaSuccess;
}

// (re)initialises the source: zero history (OnInitSource, REV.cpp:104-110; SetSize zeroes the ring, CIRC.cpp:15-21)
cudaError_t fs_conv_source_alloc(fs_ctx* ctx, uint32_t source)
{
    cudaError_t e;
    if ((e = ensure_sources(ctx, source)) != cudaSuccess) return e;
    fs_conv_source& s = ctx->conv[source];
    const fs_config& c = ctx->cfg;
    const size_t sp = spectra_elems(ctx);
    if (!s.fdl) {
        if ((e = cudaMalloc(&s.fdl, sizeof(float2) * sp)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.H[0], sizeof(float2) * sp)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.H[1], sizeof(float2) * sp)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.prev, sizeof(float) * c.n_channels * c.conv_block)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&s.ir, sizeof(float) * c.n_channels * c.sample_rate)) != cudaSuccess) return e;
        if ((e = cudaMemset(s.H[0], 0, sizeof(float2) * sp)) != cudaSuccess) return e;
        if ((e = cudaMemset(s.H[1], 0, sizeof(float2) * sp)) != cudaSuccess) return e;
        if ((e = cudaMemset(s.ir, 0, sizeof(float) * c.n_channels * c.sample_rate)) != cudaSuccess) return e;
    }
    if ((e = cudaMemset(s.fdl, 0, sizeof(float2) * sp)) != cudaSuccess) return e;
    if ((e = cudaMemset(s.prev, 0, sizeof(float) * c.n_channels * c.conv_block)) != cudaSuccess) return e;
    s.head = 0;
    s.active = true;
    return cudaSuccess;
}

void fs_conv_source_free(fs_ctx* ctx, uint32_t source)
{
    if (source >= ctx->conv_cap) return;
    fs_conv_source& s = ctx->conv[source];
    cudaFree(s.fdl); cudaFree(s.H[0]); cudaFree(s.H[1]); cudaFree(s.prev); cudaFree(s.ir);
    s.fdl = nullptr; s.H[0] = s.H[1] = nullptr; s.prev = nullptr; s.ir = nullptr;
    s.active = false; s.h_valid = 0;
}

// partition spectra of the source's current device IR into the inactive buffer, then publish
cudaError_t fs_conv_update_ir(fs_ctx* ctx, uint32_t source, cudaStream_t st)
{
    fs_conv_source& s = ctx->conv[source];
    const fs_config& c = ctx->cfg;
    const int nxt = 1 - s.h_published.load(std::memory_order_acquire);
    const size_t smem = sizeof(float2) * 2 * ctx->fft_n;
    const uint32_t tw_n = ctx->fft_n > 4096u ? ctx->fft_n : 4096u;
    dim3 grid(ctx->n_part, c.n_channels);
    uint32_t threads = c.conv_block < 1024u ? c.conv_block : 1024u;
    // twiddle table is 4096-point (or fft_n if larger); fft_stockham strides by tw_n / (2 ns)
    k_ir_spectra<<<grid, threads, smem, st>>>(s.ir, c.sample_rate, c.conv_block, c.n_channels, ctx->d_twiddle,
                                              tw_n, s.H[nxt]);
    ++ctx->stats.kernel_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // the convolver runs on the same stream unless the caller moved it; publication is ordered by the stream
    s.h_published.store(nxt, std::memory_order_release);
    s.h_valid = 1;
    return cudaSuccess;
}

// sources [s0, s0 + n), n <= FS_PTR_TABLE: all partition spectra in one launch, then publish each
cudaError_t fs_conv_update_ir_multi(fs_ctx* ctx, uint32_t s0, uint32_t n, cudaStream_t st)
{
    const fs_config& c = ctx->cfg;
    fs_ptr_table tab;
    int nxt[FS_PTR_TABLE];
    for (uint32_t i = 0; i < n; ++i) {
        fs_conv_source& s = ctx->conv[s0 + i];
        nxt[i] = 1 - s.h_published.load(std::memory_order_acquire);
        tab.p[i] = s.ir; tab.q[i] = s.H[nxt[i]];
    }
    const size_t smem = sizeof(float2) * 2 * ctx->fft_n;
    const uint32_t tw_n = ctx->fft_n > 4096u ? ctx->fft_n : 4096u;
    uint32_t threads = c.conv_block < 1024u ? c.conv_block : 1024u;
    k_ir_spectra_multi<<<dim3(ctx->n_part, c.n_channels, n), threads, smem, st>>>(tab, c.sample_rate, c.conv_block, c.n_channels,
                                                                                   ctx->d_twiddle, tw_n);
    ++ctx->stats.kernel_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    for (uint32_t i = 0; i < n; ++i) {
        fs_conv_source& s = ctx->conv[s0 + i];
        s.h_published.store(nxt[i], std::memory_order_release);
        s.h_valid = 1;
    }
    return cudaSuccess;
}

cudaError_t fs_conv_run(fs_ctx* ctx, uint32_t source, const float* d_in, float* d_out, uint32_t n_blocks,
                        cudaStream_t st)
{
    fs_conv_source& s = ctx->conv[source];
    const fs_config& c = ctx->cfg;
    conv_args a;
    a.fdl = s.fdl;
    a.H = s.H[s.h_published.load(std::memory_order_acquire)];
    a.prev = s.prev; a.in = d_in; a.out = d_out; a.tw = ctx->d_twiddle;
    a.bk = c.conv_block; a.n_part = ctx->n_part; a.n_ch = c.n_channels; a.n_blocks = n_blocks; a.head = s.head;
    a.tw_n = ctx->fft_n > 4096u ? ctx->fft_n : 4096u;
    a.wet = c.conv_wet; a.clamp = (int)c.conv_clamp;
    const size_t smem = sizeof(float2) * 2 * ctx->fft_n;
    uint32_t threads = c.conv_block < 1024u ? c.conv_block : 1024u;
    k_conv_blocks<<<c.n_channels, threads, smem, st>>>(a);
    ++ctx->stats.kernel_launches;
    s.head = (s.head + n_blocks) % ctx->n_part;
    return cudaGetLastError();
}

cudaError_t fs_conv_rfft(fs_ctx* ctx, const float* d_in, uint32_t n, float2* d_out, cudaStream_t st)
{
    const uint32_t tw_n = ctx->fft_n > 4096u ? ctx->fft_n : 4096u;
    uint32_t threads = n / 2 < 1024u ? n / 2 : 1024u;
    k_rfft<<<1, threads, sizeof(float2) * 2 * n, st>>>(d_in, n, ctx->d_twiddle, tw_n, d_out);
    ++ctx->stats.kernel_launches;
    return cudaGetLastError();
}
