// k2_post.cuh -- K2: per-stream threshold scan, key states and peak lists.
//
// Replaces, per stream and batch:
//   rx/receiver.go:383-385   dB conversion of the noise-floor scalars + two float32 rolling means
//                            over 60 blocks (dsp.RollingMean.Put, dsp/dsp.go:257-268), peak and
//                            listener thresholds
//   cw/spectral.go:49        state := value > threshold for every (block, listener)
//   dsp.FindPeaks            dsp/fft.go:254-285 on every flushed cumulation vector, emitted in bin
//                            order by a block-wide prefix sum (no atomics)
//
// The rolling mean is a *running* float32 sum (subtract oldest, add newest) whose value depends
// on the whole history, so the recurrence is executed sequentially by one thread, in block order,
// with the same float32 operations as the reference; everything around it is parallel.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/sdrgpu.h"

namespace sdr {

struct RollingState {  // two dsp.RollingMean[float32] of size 60 advancing in lock step
    float floor_values[SDR_NOISE_WINDOW];
    float dev_values[SDR_NOISE_WINDOW];
    float floor_sum, dev_sum;
    int next;
    int pad;
};

// dsp.BoolDebouncer (dsp/dsp.go:139-182) of one listener position, carried across submits:
// bit 0 effectiveState, bit 1 lastRawState, bits 2.. stateCount
typedef uint32_t DebounceState;

struct PostWork {
    int stream;
    int block_out;          // first block of this work in the per-block arrays
    int n_blocks;
    int flush_out;          // first flush slot of this work
    int n_flushes;
    int first_flush_block;  // index (within the work) of the block that closes the first window
    float peak_threshold;   // rx.Receiver.peakThreshold
    int n_listeners;
    int do_peaks;
    int debounce;           // SpectralDemodulator.SetSignalDebounce (cw/spectral.go:33-35); < 2: pass-through
    int lflags_off;         // offset of this work's listener flags in K2Args::lflags, -1: every listener active, no reset
    int pad;
};

struct K2Args {
    const PostWork *works;
    int n_works;
    int n_blocks;               // blocks of the batch
    RollingState *rolling;      // [max_streams]
    const float *psd_floor;     // [blocks]
    const double *variance;     // [blocks]
    float *thresholds;          // [blocks][4]
    const float *taps;          // [blocks][tap_stride]
    uint8_t *keys;              // [blocks][tap_stride] raw value > threshold, or nullptr (SDR_NO_RAW_KEYS)
    uint32_t *key_bits;         // [blocks][key_words] debounced key state, bit l%32 of word l/32
    int key_words;
    DebounceState *deb;         // [max_streams][tap_stride]
    const uint8_t *lflags;      // per-work listener flags (SDR_LISTENER_*)
    int tap_stride;
    const float *flush_cum;     // [flushes][N]
    int *flush_block;           // [flushes]
    int *flush_n_peaks;         // [flushes]
    sdr_peak *flush_peaks;      // [flushes][max_peaks]
    int max_peaks;
    int n;                      // block size N
};

// dsp.PSDValueIndB (dsp/fft.go:83-85) + dBmShift as a float32 add (rx/receiver.go:383-384)
// inv_n2 = 1 / N^2: N is a power of two, so the product equals the reference's quotient bit for bit (no float64 divide)
__device__ __forceinline__ float psd_value_in_db_shifted(float psd_value, double inv_n2) {
    const float db = (float)(10.0 * log10(20.0 * (double)psd_value * inv_n2));
    return __fadd_rn(db, (float)SDR_DBM_SHIFT);
}

constexpr int K2_THREADS = 128;
constexpr int K2_CHUNK = 1024;  // blocks staged in shared memory per sequential pass

// The smallest float32 x with fl(x / c) > thr (c > 0).  fl(x / c) is non-decreasing in x, so `value > threshold` of
// dsp.FindPeaks (dsp/fft.go:259-260, value := v / T(cumulationSize)) is exactly `cum >= cut` and, for a non-NaN bin,
// `value <= threshold` is exactly `cum < cut`: the scan over the bins needs no division.  NaN when no x qualifies
// (threshold NaN or +Inf): every comparison with it is false, as the reference's are.
__device__ float find_peaks_cut(float thr, float c) {
    if (isnan(thr) || thr == __int_as_float(0x7f800000)) return __int_as_float(0x7fc00000);
    float x = __fmul_rn(thr, c);
    if (isnan(x)) x = 0.f;  // -Inf * c stays -Inf, so only 0 * Inf style cases land here
    for (int it = 0; it < 4096 && __fdiv_rn(x, c) > thr; it++) x = nextafterf(x, -INFINITY);  // now fl(x / c) <= thr (or x = -Inf)
    for (int it = 0; it < 4096 && !(__fdiv_rn(x, c) > thr); it++) x = nextafterf(x, INFINITY);
    return x;
}

// dsp.FindPeaks on one vector `cum` of n bins; all K2_THREADS threads of the CTA participate.
__device__ void find_peaks_block(const float *__restrict__ cum, int n, float cumulation_size, float thr,
                                 sdr_peak *__restrict__ out, int max_peaks, int *__restrict__ n_out, int *s_scan) {
    const int tid = threadIdx.x;
    const int per = (n + K2_THREADS - 1) / K2_THREADS;
    const int lo = tid * per, hi = min(n, lo + per);
    auto val = [&](int i) { return __fdiv_rn(cum[i], cumulation_size); };  // v / T(cumulationSize)
    if (tid == 0) s_scan[0] = __float_as_int(find_peaks_cut(thr, cumulation_size));
    __syncthreads();
    const float cut = __int_as_float(s_scan[0]);
    __syncthreads();
    // The Go loop is a two-state machine: a run opens at value > threshold while closed, and closes at the first
    // value <= threshold (a NaN neither opens nor closes a run).  Each thread replays it over its chunk; the
    // state at the chunk start is found by looking back (past NaNs) at the last decisive bin.
    bool open0 = false;
    for (int j = lo - 1; j >= 0 && lo < hi; j--) {
        const float p = cum[j];
        if (p >= cut) {
            open0 = true;
            break;
        }
        if (p < cut) break;
    }
    int count = 0;
    {
        bool open = open0;
        auto step = [&](float v) {
            if (!open && v >= cut) {
                count++;
                open = true;
            } else if (open && v < cut) {
                open = false;
            }
        };
        if ((per & 3) == 0 && hi - lo == per) {  // whole chunk, 16-byte aligned: four bins per load, loads issued ahead
            const float4 *c4 = reinterpret_cast<const float4 *>(cum + lo);
            for (int i0 = 0; i0 < per / 4; i0 += 4) {
                float4 q[4];
#pragma unroll
                for (int u = 0; u < 4; u++) q[u] = (i0 + u < per / 4) ? c4[i0 + u] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (i0 + u < per / 4) {
                        step(q[u].x);
                        step(q[u].y);
                        step(q[u].z);
                        step(q[u].w);
                    }
            }
        } else {
            for (int i = lo; i < hi; i++) step(cum[i]);
        }
    }
    // block-wide exclusive scan of `count`
    s_scan[tid] = count;
    __syncthreads();
    for (int off = 1; off < K2_THREADS; off <<= 1) {
        int v = (tid >= off) ? s_scan[tid - off] : 0;
        __syncthreads();
        s_scan[tid] += v;
        __syncthreads();
    }
    int slot = s_scan[tid] - count;
    if (tid == K2_THREADS - 1) *n_out = s_scan[tid];
    if (count > 0) {
        bool open = open0;
        for (int i = lo; i < hi; i++) {
            const float c0 = cum[i];
            if (open) {
                if (c0 < cut) open = false;
                continue;
            }
            if (!(c0 >= cut)) continue;
            open = true;
            int bin = i;
            float best = val(i);
            int j = i + 1;
            while (j < n) {  // the run may extend past this thread's chunk
                if (cum[j] < cut) break;
                const float v = val(j);
                if (best < v) {  // strict: the first maximum wins
                    best = v;
                    bin = j;
                }
                j++;
            }
            if (slot < max_peaks) {
                sdr_peak p;
                p.from = i;
                p.to = j - 1;
                p.signal_bin = bin;
                p.signal_value = best;
                const bool inner = bin > 0 && bin < n - 1;
                p.y1 = inner ? cum[bin - 1] : 0.f;
                p.y2 = cum[bin];
                p.y3 = inner ? cum[bin + 1] : 0.f;
                out[slot] = p;
            }
            slot++;
        }
    }
    __syncthreads();
}

// K2 is three launches so that every stage has its own parallelism (a launch with few streams but many blocks -- 64
// receivers, 21 s each -- must not collapse onto a handful of CTAs, and one with thousands of short works -- 7 104
// streams of 100 blocks at N = 512 -- must not pay a CTA per work):
//   k2_thresholds_kernel  one WARP per work: the two rolling means (inherently sequential per stream) and the thresholds
//   k2_keys_kernel        one WARP per (work, 64 blocks): value > threshold for every listener, bit-packed
//   k2_debounce_kernel    (only when a work has a real debouncer) one CTA per such work, sequential over its blocks
//   k2_peaks_kernel       one CTA per flush: dsp.FindPeaks
constexpr int K2_WARPS = K2_THREADS / 32;
constexpr int K2_WCHUNK = 512;  // blocks staged in shared memory per pass of a warp
constexpr int K2_SHORT_WORK = 128;  // launches whose works are all this short convert their noise scalars inside the chain kernel

// Inputs of the two rolling means for every block of the batch (rx/receiver.go:383-384), one thread per block, written to
// thresholds[b].xy.  Used when a launch has long works (64 receivers x 2000 blocks): the float64 log10 of 2000 blocks
// would otherwise sit in front of one warp's sequential chain, 6 us per 128 blocks.
__global__ void __launch_bounds__(K2_THREADS) k2_db_kernel(const K2Args a) {
    const int b = blockIdx.x * K2_THREADS + threadIdx.x;
    if (b >= a.n_blocks) return;
    const double inv_n2 = 1.0 / ((double)a.n * (double)a.n);  // exact: N is a power of two
    const float dev_db = psd_value_in_db_shifted((float)sqrt(a.variance[b]), inv_n2);
    float4 t;
    t.x = psd_value_in_db_shifted(a.psd_floor[b], inv_n2);
    t.y = (float)((double)dev_db * 0.25);  // T(float64(PSDValueIndB(T(math.Sqrt(var)), N) + dBmShift) * 0.25)
    t.z = t.w = 0.f;
    reinterpret_cast<float4 *>(a.thresholds)[b] = t;
}

// dsp.RollingMean.Put (dsp/dsp.go:257-268) is sum = fl(fl(sum - oldest) + new) per block, float32, in block order: only
// that two-operation chain is sequential.  The value that falls out of the 60-deep ring at step i is known up front (the
// ring's content for i < 60, the input of step i - 60 afterwards), so the lanes stage inputs and outgoing values in
// parallel and lanes 0 / 1 walk the two chains (noise floor / deviation) over plain arrays.
// PRE: the inputs were staged by k2_db_kernel in thresholds[b].xy
// GROUP: threads per work -- a warp (many works), or the whole CTA when the launch has few, long works: the parallel
// phases around the chain then run four times as wide (one stream x 4000 blocks: 83 -> 30 us)
template <bool PRE, int GROUP>
__global__ void __launch_bounds__(K2_THREADS) k2_thresholds_kernel(const K2Args a) {
    static_assert(GROUP == 32 || GROUP == K2_THREADS, "a warp or the CTA per work");
    constexpr int NGRP = K2_THREADS / GROUP;
    constexpr int CH = GROUP == 32 ? K2_WCHUNK : 4 * K2_WCHUNK;  // blocks per pass: the CTA form stages 2048 (32 KB)
    __shared__ float s_in_all[NGRP][2][CH];   // inputs of the two means
    __shared__ float s_run_all[NGRP][2][CH];  // outgoing ring values, then the running sums
    __shared__ RollingState s_roll_all[NGRP];
    const int wq = threadIdx.x / GROUP, lane = threadIdx.x % GROUP;  // lane: index within the work's thread group
    const int wi = blockIdx.x * NGRP + wq;
    if (wi >= a.n_works) return;  // uniform in the group; nothing below synchronises across groups
    auto gsync = [&]() {
        if (GROUP == 32) __syncwarp();
        else __syncthreads();
    };
    float (*s_in)[CH] = s_in_all[wq];
    float (*s_run)[CH] = s_run_all[wq];
    RollingState &s_roll = s_roll_all[wq];
    const PostWork w = a.works[wi];
    const double n2 = 1.0 / ((double)a.n * (double)a.n);  // exact: N is a power of two
    constexpr int RW = (int)(sizeof(RollingState) / 4), RPL = (RW + GROUP - 1) / GROUP;

    {  // the lanes copy the stream's rolling state (124 words), all loads in flight at once
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.rolling + w.stream);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_roll);
        uint32_t r[RPL];
#pragma unroll
        for (int k = 0; k < RPL; k++) r[k] = (lane + GROUP * k < RW) ? src[lane + GROUP * k] : 0u;
#pragma unroll
        for (int k = 0; k < RPL; k++)
            if (lane + GROUP * k < RW) dst[lane + GROUP * k] = r[k];
    }
    // flushes of this work (rx/receiver.go:409-425): which block closed each window
    for (int f = lane; f < w.n_flushes; f += GROUP) {
        a.flush_block[w.flush_out + f] = w.block_out + w.first_flush_block + f * SDR_CUMULATION_SIZE;
        if (!w.do_peaks) a.flush_n_peaks[w.flush_out + f] = 0;
    }
    gsync();

    for (int c0 = 0; c0 < w.n_blocks; c0 += CH) {
        const int cn = min(CH, w.n_blocks - c0);
        // phase A (parallel): inputs of the two rolling means (rx/receiver.go:383-384)
        if (PRE) {  // every load of the pass in flight at once
            float4 t[CH / GROUP];
#pragma unroll
            for (int k = 0; k < CH / GROUP; k++) {
                const int i = lane + GROUP * k;
                t[k] = i < cn ? reinterpret_cast<const float4 *>(a.thresholds)[w.block_out + c0 + i] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int k = 0; k < CH / GROUP; k++) {
                const int i = lane + GROUP * k;
                if (i < cn) {
                    s_in[0][i] = t[k].x;
                    s_in[1][i] = t[k].y;
                }
            }
        }
        for (int i0 = 0; !PRE && i0 < cn; i0 += 4 * GROUP) {  // four blocks per thread and pass, the loads first
            float pf[4];
            double vr[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = i0 + lane + GROUP * k;
                pf[k] = i < cn ? a.psd_floor[w.block_out + c0 + i] : 1.f;
                vr[k] = i < cn ? a.variance[w.block_out + c0 + i] : 1.0;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = i0 + lane + GROUP * k;
                if (i < cn) {
                    // T(float64(PSDValueIndB(T(math.Sqrt(var)), N) + dBmShift) * 0.25)
                    const float dev_db = psd_value_in_db_shifted((float)sqrt(vr[k]), n2);
                    s_in[1][i] = (float)((double)dev_db * 0.25);
                    s_in[0][i] = psd_value_in_db_shifted(pf[k], n2);
                }
            }
        }
        gsync();
        // the value each step pushes out of the ring
        const int next0 = s_roll.next;
        for (int i = lane; i < cn; i += GROUP) {
            int p = next0 + i;
            if (p >= SDR_NOISE_WINDOW) p -= SDR_NOISE_WINDOW;
            if (p >= SDR_NOISE_WINDOW) p -= SDR_NOISE_WINDOW;  // only used for i < 60 (next0 < 60: one wrap)
            s_run[0][i] = i < SDR_NOISE_WINDOW ? s_roll.floor_values[p] : s_in[0][i - SDR_NOISE_WINDOW];
            s_run[1][i] = i < SDR_NOISE_WINDOW ? s_roll.dev_values[p] : s_in[1][i - SDR_NOISE_WINDOW];
        }
        gsync();
        // phase B (sequential, float32, block order): lane 0 the noise floor, lane 1 the deviation
        if (lane < 2) {
            float sum = lane == 0 ? s_roll.floor_sum : s_roll.dev_sum;
            const float *in = s_in[lane];
            float *run = s_run[lane];
            // Eight steps per pass through registers, two register sets in ping-pong: the operands of the pass after next
            // are requested before a pass computes, so the chain costs its two dependent float32 operations per step and
            // no shared-memory latency (a plain loop over the arrays: 64 cycles per step in store -> load ordering; one
            // register set with copies: 22).  The arrays hold CH entries; reading past cn (inside them) is harmless.
            float o0[8], x0[8], o1[8], x1[8];
            auto load = [&](float (&o)[8], float (&x)[8], int i0) {
                const int j0 = (i0 + 8 <= CH) ? i0 : 0;  // stay inside the arrays
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    o[u] = run[j0 + u];
                    x[u] = in[j0 + u];
                }
            };
            auto pass = [&](const float (&o)[8], const float (&x)[8], int i0) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (i0 + u < cn) {  // the last pass may be partial
                        sum = __fadd_rn(__fsub_rn(sum, o[u]), x[u]);
                        run[i0 + u] = sum;
                    }
                }
            };
            load(o0, x0, 0);
            for (int i0 = 0; i0 < cn; i0 += 16) {
                load(o1, x1, i0 + 8);
                pass(o0, x0, i0);
                load(o0, x0, i0 + 16);
                pass(o1, x1, i0 + 8);
            }
            if (lane == 0) s_roll.floor_sum = sum;
            else s_roll.dev_sum = sum;
        }
        gsync();
        // phase C (parallel): means, thresholds (rx/receiver.go:384-385,394); the ring takes the chunk's last 60 inputs
        for (int i = lane; i < cn; i += GROUP) {
            const int b = w.block_out + c0 + i;
            const float noise_floor = __fdiv_rn(s_run[0][i], (float)SDR_NOISE_WINDOW);
            const float noise_dev = __fdiv_rn(s_run[1][i], (float)SDR_NOISE_WINDOW);
            float4 th;
            th.x = noise_floor;
            th.y = noise_dev;
            th.z = __fadd_rn(w.peak_threshold, noise_floor);
            th.w = __fadd_rn(noise_floor, noise_dev);  // the listeners' threshold (rx/receiver.go:394)
            reinterpret_cast<float4 *>(a.thresholds)[b] = th;
            if (i >= cn - SDR_NOISE_WINDOW) {
                const int p = (next0 + i) % SDR_NOISE_WINDOW;
                s_roll.floor_values[p] = s_in[0][i];
                s_roll.dev_values[p] = s_in[1][i];
            }
        }
        if (lane == 0) s_roll.next = (next0 + cn) % SDR_NOISE_WINDOW;
        gsync();
    }
    {
        uint32_t *dst = reinterpret_cast<uint32_t *>(a.rolling + w.stream);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&s_roll);
        for (int i = lane; i < RW; i += GROUP) dst[i] = src[i];
    }
}

constexpr int K2_KEY_ROWS = 64;  // blocks per warp of the key kernel

// Key states (cw/spectral.go:48-54): state := value > threshold, the listener's BoolDebouncer (dsp/dsp.go:164-182), 32
// listener positions of a block packed into a word.  Pass-through debouncers (threshold < 2, the reference's default) are
// stateless, every (block, 32 listeners) word is independent: k2_keys_kernel, grid = (ceil(max blocks of a work / 64),
// ceil(works / 4)), one warp per (work, 64 blocks).  A real debouncer is sequential per listener: k2_debounce_kernel.
__global__ void __launch_bounds__(K2_THREADS) k2_keys_kernel(const K2Args a) {
    // The warp's 64 rows are one contiguous run of floats: every lane takes whole float4s (tap_stride is a multiple of
    // four, so a float4 never straddles a row), eight of them in flight at once, compares four listeners per load and
    // writes their raw key bytes as one 32-bit store; the packed words are put together through shared memory.
    __shared__ float s_thr_all[K2_WARPS][K2_KEY_ROWS];
    __shared__ uint8_t s_act_all[K2_WARPS][64];                   // per float4 of a row: which of its four positions are active
    __shared__ uint8_t s_nib_all[K2_WARPS][K2_KEY_ROWS][64 + 4];  // four key bits per float4 of a row
    const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wi = blockIdx.y * K2_WARPS + wq;
    if (wi >= a.n_works) return;  // warp-uniform; nothing below synchronises across warps
    const PostWork w = a.works[wi];
    if (w.debounce >= 2) return;  // k2_debounce_kernel
    const int row0 = blockIdx.x * K2_KEY_ROWS;
    const int nrow = min(K2_KEY_ROWS, w.n_blocks - row0);
    if (nrow <= 0) return;
    float *s_thr = s_thr_all[wq];
    uint8_t *s_act = s_act_all[wq];
    uint8_t (*s_nib)[64 + 4] = s_nib_all[wq];
    const int L = w.n_listeners;
    const float *__restrict__ taps = a.taps + (size_t)w.block_out * a.tap_stride;
    const float *__restrict__ thr = a.thresholds + (size_t)w.block_out * 4 + 3;  // listen threshold of block i at thr[4 i]
    uint8_t *__restrict__ keys = a.keys ? a.keys + (size_t)w.block_out * a.tap_stride : nullptr;
    uint32_t *__restrict__ kbits = a.key_bits + (size_t)w.block_out * a.key_words;
    const int q4 = a.tap_stride >> 2;
    for (int i = lane; i < nrow; i += 32) s_thr[i] = thr[(size_t)(row0 + i) * 4];
    for (int c4 = lane; c4 < q4; c4 += 32) {
        unsigned m = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int l = 4 * c4 + j;
            if (l < L && (w.lflags_off < 0 || (a.lflags[w.lflags_off + l] & SDR_LISTENER_ACTIVE))) m |= 1u << j;
        }
        s_act[c4] = (uint8_t)m;
    }
    __syncwarp();
    const unsigned inv_q4 = ((1u << 20) + q4 - 1) / q4;  // e / q4 == (e * inv_q4) >> 20 for e < 4096, q4 <= 64
    const float4 *t4 = reinterpret_cast<const float4 *>(taps + (size_t)row0 * a.tap_stride);
    uchar4 *k4 = keys ? reinterpret_cast<uchar4 *>(keys + (size_t)row0 * a.tap_stride) : nullptr;
    const int total = nrow * q4;
    for (int e0 = lane; e0 < total; e0 += 8 * 32) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int e = e0 + u * 32;
            v[u] = e < total ? t4[e] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int e = e0 + u * 32;
            if (e < total) {
                const int row = (int)(((unsigned)e * inv_q4) >> 20), c4 = e - row * q4;
                const float t = s_thr[row];
                const unsigned nib = ((v[u].x > t ? 1u : 0u) | (v[u].y > t ? 2u : 0u) | (v[u].z > t ? 4u : 0u) | (v[u].w > t ? 8u : 0u)) & s_act[c4];
                // bits 0..3 -> bytes 0..3 (one raw key byte per position)
                if (k4) reinterpret_cast<uint32_t *>(k4)[e] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
                s_nib[row][c4] = (uint8_t)nib;
            }
        }
    }
    __syncwarp();
    for (int i = lane; i < nrow * a.key_words; i += 32) {
        const int row = i / a.key_words, lg = i - row * a.key_words;
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int c4 = lg * 8 + j;
            if (c4 < q4) word |= (uint32_t)s_nib[row][c4] << (4 * j);
        }
        kbits[(size_t)(row0 + row) * a.key_words + lg] = word;
    }
}

// A real debouncer (SetSignalDebounce >= 2): one CTA per work walks all its blocks in order, a warp per 32 listener
// positions, the state carried per (stream, position) across submits.  Works with a pass-through debouncer return at once.
__global__ void __launch_bounds__(K2_THREADS) k2_debounce_kernel(const K2Args a) {
    const PostWork w = a.works[blockIdx.x];
    if (w.debounce < 2) return;
    const int tid = threadIdx.x, wq = tid >> 5, lane = tid & 31;
    constexpr int NWQ = K2_THREADS / 32;
    const int L = w.n_listeners;
    const float *__restrict__ taps = a.taps + (size_t)w.block_out * a.tap_stride;
    const float *__restrict__ thr = a.thresholds + (size_t)w.block_out * 4 + 3;  // listen threshold of block i at thr[4 i]
    uint8_t *__restrict__ keys = a.keys ? a.keys + (size_t)w.block_out * a.tap_stride : nullptr;
    uint32_t *__restrict__ kbits = a.key_bits + (size_t)w.block_out * a.key_words;
    for (int lg = wq; lg < a.key_words; lg += NWQ) {
        const int l = lg * 32 + lane;
        const uint8_t lf = (l < L) ? (w.lflags_off >= 0 ? a.lflags[w.lflags_off + l] : (uint8_t)SDR_LISTENER_ACTIVE) : (uint8_t)0;
        const bool active = (lf & SDR_LISTENER_ACTIVE) != 0;
        DebounceState *dst = a.deb + (size_t)w.stream * a.tap_stride + l;
        DebounceState st = 0;
        if (l < L && !(lf & SDR_LISTENER_RESET)) st = *dst;
        bool eff = (st & 1u) != 0, last = (st & 2u) != 0;
        int count = (int)(st >> 2);
        for (int i0 = 0; i0 < w.n_blocks; i0 += 8) {
            float v[8], t[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                v[u] = (active && i0 + u < w.n_blocks) ? taps[(size_t)(i0 + u) * a.tap_stride + l] : 0.f;
                t[u] = (i0 + u < w.n_blocks) ? thr[(size_t)(i0 + u) * 4] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u;
                if (i < w.n_blocks) {  // uniform across the warp
                    const bool raw = active && v[u] > t[u];
                    if (keys && l < L) keys[(size_t)i * a.tap_stride + l] = raw ? 1 : 0;
                    bool out = false;
                    if (active) {
                        count = (raw != last) ? 1 : count + 1;
                        last = raw;
                        if (count >= w.debounce) eff = raw;
                        out = eff;
                    }
                    const uint32_t word = __ballot_sync(0xffffffffu, out);
                    if (lane == 0) kbits[(size_t)i * a.key_words + lg] = word;
                }
            }
        }
        if (l < L && (active || (lf & SDR_LISTENER_RESET))) *dst = (eff ? 1u : 0u) | (last ? 2u : 0u) | ((uint32_t)count << 2);
    }
}

// dsp.FindPeaks (dsp/fft.go:254-285) on every flushed cumulation vector: one CTA per flush
__global__ void __launch_bounds__(K2_THREADS) k2_peaks_kernel(const K2Args a) {
    __shared__ int s_scan[K2_THREADS];
    const int slot = blockIdx.x;
    const float thr = a.thresholds[(size_t)a.flush_block[slot] * 4 + 2];
    find_peaks_block(a.flush_cum + (size_t)slot * a.n, a.n, (float)SDR_CUMULATION_SIZE, thr, a.flush_peaks + (size_t)slot * a.max_peaks,
                     a.max_peaks, &a.flush_n_peaks[slot], s_scan);
}

// ---- stand-alone kernels behind the dsp-signature-compatible calls --------------------------
__global__ void __launch_bounds__(K2_THREADS) find_peaks_kernel(const float *cum, int n, float cumulation_size, float thr,
                                                                sdr_peak *out, int max_peaks, int *n_out) {
    __shared__ int s_scan[K2_THREADS];
    find_peaks_block(cum, n, cumulation_size, thr, out, max_peaks, n_out, s_scan);
}

}  // namespace sdr
