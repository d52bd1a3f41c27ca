/*
 * ref_kissfft_conv.c -- thin driver around the REFERENCE's own vendored KissFFT sources
 * (compiled in place from /root/reference/Plugins/FrequenSee/Source/FrequenSee/Private/
 * FrequenSeeFFTConvolver/KissFFT/{kiss_fft.c,kiss_fftr.c} by oracle/Makefile; the sources are
 * NOT copied into this repo).  TEST INFRASTRUCTURE ONLY (see fs_oracle.h).
 *
 * It restates the reference's convolution scheme -- FFFrequenSeeAudioReverbPlugin::Initialize
 * (REV.cpp:74-102), ProcessSourceAudio (REV.cpp:118-170) and ConvolveFFT (REV.cpp:172-213):
 * per callback and channel, a 65 536-point kiss_fftr of the 49 023-sample history, a
 * kiss_fftr of the zero-padded 48 000-tap IR (re-transformed every callback, REV.cpp:188),
 * complex multiply over 32 769 bins (REV.cpp:191-199), kiss_fftri (REV.cpp:202), 1/FFTSize
 * scaling of [IRSize-1, IRSize-1+FrameSize) (REV.cpp:205-209).  Used (a) to pin the oracle's
 * direct-form convolution against the real reference FFT, and (b) as the "reference"-kind CPU
 * baseline of the convolution stage.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "kiss_fftr.h"

typedef struct ref_conv {
    int sample_rate, frame, channels, ir_size, tail, fft_size;
    kiss_fftr_cfg fwd, inv;
    float* ring;        /* [C][ir_size-1] most recent history (FCircularAudioBuffer, CIRC.cpp) */
    float* cur_tail;    /* [tail] CurrAudioTail, REV.cpp:83-84 */
    float* ir;          /* [C][ir_size] */
    float* in_pad; float* ir_pad; float* td;
    kiss_fft_cpx* fi; kiss_fft_cpx* fh; kiss_fft_cpx* fo;
} ref_conv;

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

ref_conv* ref_conv_create(int sample_rate, int frame, int channels)
{
    ref_conv* r = (ref_conv*)calloc(1, sizeof(*r));
    r->sample_rate = sample_rate; r->frame = frame; r->channels = channels;
    r->ir_size = sample_rate;                      /* IRSize = SamplingRate * 1.0, REV.cpp:79 */
    r->tail = r->ir_size - 1 + frame;              /* CurrTailSize, REV.cpp:82 */
    r->fft_size = next_pow2(r->tail);              /* RoundUpToPowerOfTwo, REV.cpp:90 */
    r->fwd = kiss_fftr_alloc(r->fft_size, 0, NULL, NULL);
    r->inv = kiss_fftr_alloc(r->fft_size, 1, NULL, NULL);
    r->ring = (float*)calloc((size_t)channels * (r->ir_size - 1), 4);
    r->cur_tail = (float*)calloc(r->tail, 4);
    r->ir = (float*)calloc((size_t)channels * r->ir_size, 4);
    r->in_pad = (float*)calloc(r->fft_size, 4);
    r->ir_pad = (float*)calloc(r->fft_size, 4);
    r->td = (float*)calloc(r->fft_size, 4);
    int nb = r->fft_size / 2 + 1;
    r->fi = (kiss_fft_cpx*)calloc(nb, sizeof(kiss_fft_cpx));
    r->fh = (kiss_fft_cpx*)calloc(nb, sizeof(kiss_fft_cpx));
    r->fo = (kiss_fft_cpx*)calloc(nb, sizeof(kiss_fft_cpx));
    return r;
}

void ref_conv_destroy(ref_conv* r)
{
    if (!r) return;
    kiss_fftr_free(r->fwd); kiss_fftr_free(r->inv);
    free(r->ring); free(r->cur_tail); free(r->ir); free(r->in_pad); free(r->ir_pad); free(r->td);
    free(r->fi); free(r->fh); free(r->fo); free(r);
}

int ref_conv_fft_size(const ref_conv* r) { return r->fft_size; }

void ref_conv_set_ir(ref_conv* r, const float* ir)
{
    memcpy(r->ir, ir, (size_t)r->channels * r->ir_size * 4);
}

/* in/out interleaved [frame][channels]; de-interleave FIXed */
void ref_conv_process(ref_conv* r, const float* in, float* out, int clamp)
{
    const int C = r->channels, F = r->frame, H = r->ir_size - 1;
    const int nb = r->fft_size / 2 + 1;
    for (int c = 0; c < C; ++c) {
        float* ring = r->ring + (size_t)c * H;
        memcpy(r->cur_tail, ring, (size_t)H * 4);                 /* GetLastSamples, REV.cpp:141-142 */
        for (int i = 0; i < F; ++i) r->cur_tail[H + i] = in[i * C + c];
        memmove(ring, ring + F, (size_t)(H - F) * 4);             /* AddSamples, REV.cpp:144-145 */
        for (int i = 0; i < F; ++i) ring[H - F + i] = in[i * C + c];
        /* ConvolveFFT, REV.cpp:172-213 */
        memset(r->in_pad, 0, (size_t)r->fft_size * 4);
        memset(r->ir_pad, 0, (size_t)r->fft_size * 4);
        memcpy(r->in_pad, r->cur_tail, (size_t)r->tail * 4);
        memcpy(r->ir_pad, r->ir + (size_t)c * r->ir_size, (size_t)r->ir_size * 4);
        kiss_fftr(r->fwd, r->in_pad, r->fi);
        kiss_fftr(r->fwd, r->ir_pad, r->fh);
        for (int i = 0; i < nb; ++i) {
            r->fo[i].r = r->fi[i].r * r->fh[i].r - r->fi[i].i * r->fh[i].i;
            r->fo[i].i = r->fi[i].r * r->fh[i].i + r->fi[i].i * r->fh[i].r;
        }
        kiss_fftri(r->inv, r->fo, r->td);
        const float scale = 1.0f / (float)r->fft_size;
        for (int i = 0; i < F; ++i) {
            float y = r->td[H + i] * scale;
            if (clamp) { if (y > 1.0f) y = 1.0f; if (y < -1.0f) y = -1.0f; }
            out[i * C + c] = y;
        }
    }
}

/* plain real FFT through the reference's kiss_fftr, for pinning the CUDA FFT kernel */
void ref_kiss_fftr(int n, const float* in, float* out_ri /* [n/2+1][2] */)
{
    kiss_fftr_cfg cfg = kiss_fftr_alloc(n, 0, NULL, NULL);
    kiss_fftr(cfg, in, (kiss_fft_cpx*)out_ri);
    kiss_fftr_free(cfg);
}
void ref_kiss_fftri(int n, const float* in_ri, float* out)
{
    kiss_fftr_cfg cfg = kiss_fftr_alloc(n, 1, NULL, NULL);
    kiss_fftri(cfg, (const kiss_fft_cpx*)in_ri, out);
    kiss_fftr_free(cfg);
}
