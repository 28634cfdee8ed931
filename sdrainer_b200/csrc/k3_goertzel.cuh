// k3_goertzel.cuh -- K3: Goertzel / envelope bank.
//
// (a) audio path -- dsp.Goertzel (dsp/dsp.go:34-136) behind cw.AudioDemodulator (cw/audio.go:169-211):
//     every filter owns a real float32 audio stream cut into blocks of Blocksize() samples.
//       goertzel_audio_mag_kernel   one thread per (filter, block): autoscale/clip (cw/audio.go:184-192)
//                                   and the float64 recurrence q0 = coeff*q1 - q2 + x (dsp/dsp.go:98-106)
//                                   with the reference's operation order and no FMA contraction, so the
//                                   magnitudes are bit-identical to the Go/oracle values.
//       goertzel_audio_norm_kernel  one thread per filter: the running magnitudeLimit recurrence
//                                   (dsp/dsp.go:111-123) is sequential across blocks by construction.
// (b) IQ path -- the north-star's multi-listener bank on complex blocks: a block-length Goertzel at a
//     bin-centre frequency equals that DFT bin, so each (block, listener) evaluates
//     X[k] = sum_n x[n] W_N^(nk) directly with table twiddles (an fp32 Goertzel recurrence is numerically
//     unsafe at long N / low omega) and projects it to dB exactly like K1 (rx/receiver.go:376-378,393).
//     The block is staged once in shared memory by a TMA bulk copy and shared by all listeners; the twiddle is a
//     rotating phasor re-synchronised from the table every 32 samples (see goertzel_iq_kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "k1_spectral.cuh"

namespace sdr {

struct GoertzelFilter {
    double coeff;                // 2*cos(2*pi*k/bs)
    double magnitude_limit_low;  // bs/2
    double magnitude_limit;      // running state, carried across calls
    double magnitude_threshold;  // 0.75
    int blocksize;
    int pad;
};

struct GoertzelAudioArgs {
    const GoertzelFilter *filters;
    const float *const *audio;  // [n_filters] device pointers
    const int *n_blocks;        // [n_filters]
    const float *scale;         // [n_filters] 0 = auto, 1 = none
    double max_scale;
    double *magnitude;          // [n_filters][out_stride] raw magnitude, normalised in place by the 2nd kernel
    uint8_t *state;             // [n_filters][out_stride]
    int out_stride;
    int n_filters;
    int max_blocks;
};

__device__ __forceinline__ float truncate_f32(float v) {  // cw/audio.go:213-221
    if (v > 1.f) return 1.f;
    if (v < -1.f) return -1.f;
    return v;
}

__global__ void __launch_bounds__(128) goertzel_audio_mag_kernel(const GoertzelAudioArgs a) {
    const int f = blockIdx.y;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.n_filters || b >= a.n_blocks[f]) return;
    const GoertzelFilter flt = a.filters[f];
    const int bs = flt.blocksize;
    const float *x = a.audio[f] + (size_t)b * bs;
    float scale = a.scale[f];
    if (scale == 0.f) {  // autoscale: scale = float32(min(1/float64(max|x|), maxScale)), cw/audio.go:185-188
        float mx = 0.f;
        for (int i = 0; i < bs; i++) {
            const float v = fabsf(__ldg(&x[i]));
            if (v > mx) mx = v;
        }
        const double inv = 1.0 / (double)mx;
        scale = (float)(inv < a.max_scale ? inv : a.max_scale);
    }
    const bool do_scale = scale != 1.f;
    double q1 = 0.0, q2 = 0.0;
    const double coeff = flt.coeff;
    for (int i = 0; i < bs; i++) {
        float s = __ldg(&x[i]);
        if (do_scale) s = truncate_f32(__fmul_rn(s, scale));
        // q0 = coeff*q1 - q2 + float64(sample): ((coeff*q1) - q2) + x, three roundings
        const double q0 = __dadd_rn(__dsub_rn(__dmul_rn(coeff, q1), q2), (double)s);
        q2 = q1;
        q1 = q0;
    }
    // sqrt((q1*q1) + (q2*q2) - q1*q2*coeff)
    const double m2 = __dsub_rn(__dadd_rn(__dmul_rn(q1, q1), __dmul_rn(q2, q2)), __dmul_rn(__dmul_rn(q1, q2), coeff));
    a.magnitude[(size_t)f * a.out_stride + b] = __dsqrt_rn(m2);
}

__global__ void goertzel_audio_norm_kernel(const GoertzelAudioArgs a, GoertzelFilter *filters_rw) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.n_filters) return;
    GoertzelFilter flt = filters_rw[f];
    double lim = flt.magnitude_limit;
    const int nb = a.n_blocks[f];
    for (int b = 0; b < nb; b++) {
        const double mag = a.magnitude[(size_t)f * a.out_stride + b];
        if (mag > flt.magnitude_limit_low) lim = __dadd_rn(lim, __ddiv_rn(__dsub_rn(mag, lim), 6.0));
        if (lim < flt.magnitude_limit_low) lim = flt.magnitude_limit_low;
        const double norm = __ddiv_rn(mag, lim);
        a.magnitude[(size_t)f * a.out_stride + b] = norm;
        a.state[(size_t)f * a.out_stride + b] = norm > flt.magnitude_threshold ? 1 : 0;
    }
    filters_rw[f].magnitude_limit = lim;
}

// ---- IQ bank -------------------------------------------------------------------------------
struct GoertzelIqArgs {
    const float *iq;        // [n_blocks][2N] device
    const float2 *twiddle;  // [N] W_N^m
    const int *bins;        // [n_bins] fftshifted bin index (Listener.SignalBin)
    float *out_db;          // [n_blocks][n_bins]
    int n;                  // N
    int n_blocks;
    int n_bins;
    float db_offset;        // 10*log10(20/N^2)
};

constexpr int K3_THREADS = 256;
constexpr int K3_CHUNK = 32;  // samples between two table re-synchronisations of the rotating twiddle

// One CTA per block (persistent over blocks); the block is staged ONCE in shared memory by a TMA bulk copy and
// shared by all listeners.  Work split: lane = listener (LPT listeners per lane, 32*LPT per pass), warp = sample
// chunk.  Per sample the lane reads x[n] as a shared-memory broadcast and, per listener, does acc += x * w and
// w *= W_N^k in packed f32x2 arithmetic (4 instructions); w is re-read from the fp64-rounded table every K3_CHUNK
// samples, which bounds the drift of the recurrence at ~K3_CHUNK * 2^-24 (a free-running fp32 Goertzel / phasor
// recurrence is numerically unsafe at long N).  Dynamic smem = 8N bytes (block) + 8 * 32 * LPT * 8 (warp partials).
// ACT of the lane's LPT listener chains are live: acc[i] += x[n] * W_N^(k_i n) over this warp's sample chunks
template <int LPT, int ACT>
__device__ __forceinline__ void k3_accumulate(const float2 *X, const float2 *__restrict__ twiddle, const int (&k)[LPT],
                                              const float2 (&step)[LPT], float2 (&acc)[LPT], int warp, int n_chunks, int N) {
    for (int c = warp; c < n_chunks; c += K3_THREADS / 32) {
        const int n0 = c * K3_CHUNK;
        float2 w[ACT];
#pragma unroll
        for (int i = 0; i < ACT; i++) w[i] = __ldg(&twiddle[(k[i] * n0) & (N - 1)]);
#pragma unroll
        for (int j = 0; j < K3_CHUNK; j++) {
            const float2 x = X[n0 + j];  // broadcast
#pragma unroll
            for (int i = 0; i < ACT; i++) {
                // acc += x * w ; w *= step
                acc[i] = __ffma2_rn(make_float2(x.x, x.x), w[i], acc[i]);
                acc[i] = __ffma2_rn(make_float2(x.y, x.y), make_float2(-w[i].y, w[i].x), acc[i]);
                w[i] = cmul(w[i], step[i]);
            }
        }
    }
}

template <int LPT>
__global__ void __launch_bounds__(K3_THREADS) goertzel_iq_kernel(const GoertzelIqArgs a) {
    extern __shared__ __align__(128) unsigned char k3_smem[];
    __shared__ __align__(8) uint64_t bar;
    float2 *X = reinterpret_cast<float2 *>(k3_smem);
    const int N = a.n;
    float2 *RED = reinterpret_cast<float2 *>(k3_smem + (size_t)8 * N);  // [warp][32 * LPT]
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    constexpr int NW = K3_THREADS / 32;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    const int n_chunks = N / K3_CHUNK;
    for (int blk = blockIdx.x; blk < a.n_blocks; blk += gridDim.x) {
        if (threadIdx.x == 0) {
            fence_proxy_async();
            mbar_expect_tx(&bar, (uint32_t)(8 * N));
            tma_load_1d(X, a.iq + (size_t)blk * 2 * N, (uint32_t)(8 * N), &bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
        for (int g0 = 0; g0 < a.n_bins; g0 += 32 * LPT) {
            // listeners per lane in this pass (uniform): the last pass runs only as many chains as it has listeners
            const int lpt_here = (a.n_bins - g0 + 31) / 32 < LPT ? (a.n_bins - g0 + 31) / 32 : LPT;
            int k[LPT];
            float2 step[LPT], acc[LPT];
#pragma unroll
            for (int i = 0; i < LPT; i++) {
                const int l = g0 + lane + 32 * i;
                const int kk = l < a.n_bins ? __ldg(&a.bins[l]) : 0;
                k[i] = (kk + N / 2) & (N - 1);  // undo dsp/fft.go:54-57
                step[i] = __ldg(&a.twiddle[k[i]]);
                acc[i] = make_float2(0.f, 0.f);
            }
            // chunk loop specialised on the number of active chains (uniform per pass)
            switch (lpt_here) {
                case 1: k3_accumulate<LPT, 1>(X, a.twiddle, k, step, acc, warp, n_chunks, N); break;
                case 2: k3_accumulate<LPT, (LPT >= 2 ? 2 : LPT)>(X, a.twiddle, k, step, acc, warp, n_chunks, N); break;
                case 3: k3_accumulate<LPT, (LPT >= 3 ? 3 : LPT)>(X, a.twiddle, k, step, acc, warp, n_chunks, N); break;
                default: k3_accumulate<LPT, LPT>(X, a.twiddle, k, step, acc, warp, n_chunks, N); break;
            }
#pragma unroll
            for (int i = 0; i < LPT; i++) RED[warp * 32 * LPT + lane + 32 * i] = acc[i];
            __syncthreads();
            for (int q = threadIdx.x; q < 32 * LPT; q += K3_THREADS) {
                const int l = g0 + q;
                if (l < a.n_bins) {
                    float re = 0.f, im = 0.f;
#pragma unroll
                    for (int wq = 0; wq < NW; wq++) {
                        re += RED[wq * 32 * LPT + q].x;
                        im += RED[wq * 32 * LPT + q].y;
                    }
                    const float psd = fmaf(re, re, im * im);
                    const float t = fmaf(3.01029995663981195f, __log2f(psd), a.db_offset);
                    a.out_db[(size_t)blk * a.n_bins + l] = __fadd_rn(t, 120.0f);
                }
            }
            __syncthreads();  // RED is rewritten by the next listener group; X by the next bulk copy
        }
    }
}

}  // namespace sdr
