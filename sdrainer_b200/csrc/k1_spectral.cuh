// k1_spectral.cuh -- K1: fused spectral front end, one pass over the IQ in HBM.
//
// Replaces, for every block of N complex samples (rx/receiver.go:379-407):
//   dsp.FFT.IQToSpectrumAndPSD   dsp/fft.go:23-37   (FFT, fftshift, |X|^2, dB + 120)
//   dsp.FindNoiseFloor           dsp/fft.go:215-252 (10 window means, min, variance)
//   listener taps                rx/receiver.go:393 (spectrum[l.SignalBin()])
//   cumulation                   rx/receiver.go:404-407 (float32, block order), flushed every 100
//
// Data movement: each IQ block (8N bytes) is read from HBM exactly once, by a TMA bulk copy
// (cp.async.bulk -> UBLKCP) into a shared-memory ring guarded by mbarriers; spectrum and PSD never
// go to HBM (unless DEBUG_STORE).  Per block only 4L+12 bytes are written; 4N bytes per flush.
//
// Work decomposition: a *segment* is a run of <=100 consecutive blocks of one stream inside one
// cumulation window.  A thread *group* of T = N/16 threads owns a segment: each thread keeps 16
// points in registers, does radix-16 / radix-16 / radix-(N/256) passes with two shared-memory
// exchanges, and keeps its 16 cumulation bins in registers for the whole segment (sequential
// float32 adds in block order, exactly like the reference).  Groups are independent; a CTA holds
// G groups so that it always has 128..256 threads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_radix.cuh"

#ifndef SDR_K1_MINB
#define SDR_K1_MINB 4
#endif

namespace sdr {

struct Segment {
    const float *iq;      // first block of the segment (device, 16-byte aligned)
    int n_blocks;         // 1..100
    int stream;           // stream slot (cumulation state row)
    int work;             // index into WorkParams
    int block_out;        // index of the first block in the per-block output arrays
    int flush_idx;        // >=0: this segment closes a window -> write cumulation there; -1: store to state
    int state_in;         // >= 0: start from the partial cumulation saved in this row of cum_state; -1: start at 0
    int state_out;        // row of cum_state that receives the partial cumulation when the segment does not flush.
                          // A stream owns two rows and alternates: the segment that continues a window (reads a row)
                          // and the one that leaves the next window open (writes a row) can share a launch.
};

struct WorkParams {
    int edge_width;
    int n_listeners;
    int listener_off;  // offset into the listener bin array
    int pad;
    // noise-floor window sums: lane group g (TPW consecutive threads) sums window nf_map[g].  The host picks the
    // permutation (per N and edge width) that minimises shared-memory bank conflicts between the windows that
    // share a warp (engine.cu choose_nf_map); the identity is always correct.
    unsigned char nf_map[16];
};

struct K1Args {
    const Segment *segs;
    int n_segs;
    const WorkParams *works;
    const int *listener_bins;
    const float2 *tw1;      // [15][M]  W_N^(j*k1)
    const float2 *tw2;      // [15][R3] W_M^(n3*k2)
    const float *window;    // [N] or nullptr
    float *cum_state;       // [2 * max_streams][N]
    float *psd_floor;       // [blocks]
    double *variance;       // [blocks]
    float *taps;            // [blocks][tap_stride]
    int tap_stride;
    float *flush_cum;       // [flushes][N]
    float *dbg_spectrum;    // [blocks][N] (DEBUG_STORE)
    float *dbg_psd;         // [blocks][N]
};

// ---- PTX helpers: mbarrier + TMA bulk copy -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk async copy global -> shared, completion on mbarrier (TMA engine; SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// kiwi.decodeIQBytes (kiwi/client.go:298-308): one sample = 4 wire bytes I_hi I_lo Q_hi Q_lo (big-endian int16
// pair) -> (float32(int16) / float32(32767)) for I and Q.  The IEEE division is done as a reciprocal multiply plus
// one FMA residual correction (Markstein), which is correctly rounded for these operands -- verified bit-exact
// over all 65536 inputs (tests/test_gpu_kiwi.py) -- so the FFT input equals the reference's float32 values.
__device__ __forceinline__ float2 kiwi_decode_sample(uint32_t raw) {
    const uint32_t sw = __byte_perm(raw, 0, 0x2301);  // swap the bytes of each half-word
    const float2 x = make_float2((float)(short)(sw & 0xffffu), (float)(short)(sw >> 16));
    const float r = 1.0f / 32767.0f;                  // correctly rounded reciprocal (compile-time constant)
    const float2 q = __fmul2_rn(x, make_float2(r, r));
    const float2 rem = __ffma2_rn(make_float2(-q.x, -q.y), make_float2(32767.0f, 32767.0f), x);  // exact residual
    return __ffma2_rn(rem, make_float2(r, r), q);
}

template <int N>
struct K1Geom {
    static constexpr int M = N / 16;       // points per pass-1 column set == threads per group
    static constexpr int R3 = N / 256;     // last radix
    static constexpr int T = N / 16;       // threads per group
    static constexpr int PAIRS = 16 / R3;  // (k1,k2) pairs per thread in pass 3
    static constexpr int S1 = M + R3;      // E1 row stride (complex), bank-conflict free
    static constexpr int S2 = 256 + 16 / R3;  // E2 row stride (complex)
    static constexpr int E1_BYTES = 16 * S1 * 8;
    static constexpr int E2_BYTES = R3 * S2 * 8;
    static constexpr int STAGE_BYTES = ((E2_BYTES > 8 * N ? E2_BYTES : 8 * N) + 127) / 128 * 128;
    static constexpr int NSTAGE = 2;
    static constexpr int LMAX = 256;       // listener bins cached in smem per group
    static constexpr int TW2_BYTES = 15 * R3 * 8;
    static constexpr int TPW = (T / 10) / 3 * 3;  // noise floor: threads per window (multiple of 3: lanes 3w..3w+2 combine)
    static constexpr int NF_MAX = ((N / 10 + TPW - 1) / TPW) | 1;  // longest per-thread share of a noise window (edge 0)
    static constexpr int PART_PITCH = TPW + 1;   // partial sums [window][PART_PITCH]: the windows one warp combines sit 8 banks apart
    static constexpr int PART_SLOTS = T + 16;    // >= 10 * PART_PITCH
    static constexpr int NFB = 16;               // noise floor: blocks whose window sums are batched for selection
    static constexpr int NF_BYTES = NFB * 10 * (8 + 8 + 4);
    static constexpr int MISC_BYTES = NF_BYTES + PART_SLOTS * 8 /*share partials (s1,s2)*/ + LMAX * 4 + NSTAGE * 8 /*mbar*/ + 64;
    static_assert(10 * PART_PITCH + (T - 10 * TPW) <= PART_SLOTS, "partial sums fit");
    static constexpr int GROUP_BYTES = ((NSTAGE * STAGE_BYTES + E1_BYTES + TW2_BYTES + MISC_BYTES) + 127) / 128 * 128;
    static constexpr int G = (T >= 128) ? 1 : 128 / T;  // groups per CTA
    static constexpr int CTA_THREADS = G * T;
    static constexpr int SMEM_BYTES = G * GROUP_BYTES;
    static_assert(R3 == 2 || R3 == 4 || R3 == 8 || R3 == 16, "fused path covers N = 512..4096");
    static_assert(E1_BYTES >= 4 * N, "PSD aliases E1");
};

template <int T, int G>
__device__ __forceinline__ void group_sync(int g) {
    if (T == 32) {
        __syncwarp();
    } else if (G == 1) {
        __syncthreads();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(T) : "memory");
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// MUFU.LG2 without the denormal fix-up sequence: a denormal |X|^2 (< 1.2e-38) is treated as 0 -> -Inf dB
__device__ __forceinline__ float fast_log2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// dB projection of rx/receiver.go:376-378: float32(10*log10(20*psd/N^2)) + float32(120)
template <int N>
__device__ __forceinline__ float psd_to_db(float psd) {
    // 10*log10(20*p/N^2) = A*log2(p) + B
    constexpr float A = 3.01029995663981195f;  // 10*log10(2)
    constexpr double LOG2N = (N == 512) ? 9.0 : (N == 1024) ? 10.0 : (N == 2048) ? 11.0 : (N == 4096) ? 12.0 : 0.0;
    constexpr float B = (float)(13.0102999566398120 - 20.0 * LOG2N * 0.30102999566398120);  // 10*log10(20) - 20*log10(N)
    float t = fmaf(A, fast_log2(psd), B);
    return __fadd_rn(t, 120.0f);
}

// ---- dsp.FindNoiseFloor (dsp/fft.go:215-252), parallel form --------------------------------
// Reference behaviour reproduced here (quirks included; dsp/fft.go:215-252):
//  * window w covers bins [e + w*ws, e + (w+1)*ws), ws = (N-2e)/10; its mean is only evaluated when
//    the loop reaches the NEXT bin, so the 10th window is skipped when 10*ws == N-2e (n_win = 9);
//  * the first window is always taken, later ones only on a strict `<` (earliest minimum wins;
//    a NaN mean never replaces, a NaN first mean is never replaced);
//  * `from` is only ever assigned in the very first iteration (count is 0 at the loop top only
//    there), so the variance runs over [e .. first bin of the window after the winner] INCLUSIVE,
//    i.e. (w+1)*ws + 1 terms around the winner's mean, divided by ws.
// float64 sums of float32 values as in the reference.  The variance is evaluated from per-window
// sums of x and x^2 (both exact products in float64): sum (x-m)^2 = S2 - m*(2*S1 - n*m); rounding
// differs from the reference's sequential loop at the 1e-15 relative level.
__host__ __device__ __forceinline__ int nf_window_size(int n, int e) { return (n - 2 * e) / 10; }
__host__ __device__ __forceinline__ int nf_window_count(int n, int e) {
    const int ws = nf_window_size(n, e);
    return (10 * ws < n - 2 * e) ? 10 : 9;
}

// all T threads of a group (T % 32 == 0); WS1/WS2[0..n_win) valid after the next group barrier
template <int T>
__device__ __forceinline__ void nf_window_sums(const float *PSD, double *WS1, double *WS2, int e, int ws, int n_win, int t) {
    const int warp = t / 32, lane = t % 32;
    constexpr int NW = T / 32;
    for (int w = warp; w < n_win; w += NW) {
        const int from = e + w * ws;
        double a1 = 0.0, a2 = 0.0;
        for (int i = lane; i < ws; i += 32) {
            const double x = (double)PSD[from + i];
            a1 += x;
            a2 = fma(x, x, a2);
        }
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (lane == 0) {
            WS1[w] = a1;
            WS2[w] = a2;
        }
    }
}

// one full warp: lane w owns window w and passes in that window's sums (s1 = sum x, s2 = sum x^2)
// and x_to = psd[e + (w+1)*ws] (the first bin of the next window); lanes >= n_win pass zeros.
__device__ __forceinline__ void nf_select_variance(double s1, double s2, double x_to, int ws, int n_win, int lane,
                                                   float *out_min, double *out_var) {
    const unsigned full = 0xffffffffu;
    const double dws = (double)ws;
    const bool real = lane < n_win;
    const double mean = s1 / dws;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double key = mean;
    if (!real || (lane > 0 && isnan(mean))) key = inf;
    int idx = lane;
    const double mean0 = __shfl_sync(full, mean, 0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ok = __shfl_xor_sync(full, key, o);
        const int oi = __shfl_xor_sync(full, idx, o);
        if (ok < key || (ok == key && oi < idx)) {
            key = ok;
            idx = oi;
        }
    }
    const int wsel = isnan(mean0) ? 0 : idx;
    // inclusive prefix sums over windows 0..lane (only lanes < 16 matter: n_win <= 10)
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const double p1 = __shfl_up_sync(full, s1, o);
        const double p2 = __shfl_up_sync(full, s2, o);
        if (lane >= o) {
            s1 += p1;
            s2 += p2;
        }
    }
    const double m = __shfl_sync(full, mean, wsel);
    double P1 = __shfl_sync(full, s1, wsel);
    double P2 = __shfl_sync(full, s2, wsel);
    const double x = __shfl_sync(full, x_to, wsel);
    if (lane == 0) {
        P1 += x;
        P2 = fma(x, x, P2);
        const double n = (double)((wsel + 1) * ws + 1);
        *out_min = (float)m;
        *out_var = (P2 - m * (2.0 * P1 - n * m)) / dws;
    }
}

// One thread runs the sequential selection of dsp.FindNoiseFloor (dsp/fft.go:217-251) for one block from
// that block's window sums: s1[w*stride] = sum x, s2[w*stride] = sum x^2, xto[w*stride] = psd[e+(w+1)*ws].
__device__ __forceinline__ void nf_select_serial(const double *s1, const double *s2, const float *xto, int stride, int ws,
                                                 int n_win, float *out_min, double *out_var) {
    const double inv_ws = 1.0 / (double)ws;
    double min_value = 0.0, P1 = 0.0, P2 = 0.0, bP1 = 0.0, bP2 = 0.0;
    int best = 0;
    for (int w = 0; w < n_win; w++) {
        const double a1 = s1[w * stride];
        P1 += a1;
        P2 += s2[w * stride];
        const double mean = a1 * inv_ws;
        if (w == 0 || mean < min_value) {  // `mean < minValue || first`
            min_value = mean;
            best = w;
            bP1 = P1;
            bP2 = P2;
        }
    }
    const double x = (double)xto[best * stride];
    bP1 += x;
    bP2 = fma(x, x, bP2);
    const double n = (double)((best + 1) * ws + 1);  // bins e .. e+(best+1)*ws inclusive (the reference's `from` quirk)
    *out_min = (float)min_value;
    *out_var = (bP2 - min_value * (2.0 * bP1 - n * min_value)) * inv_ws;
}

// two bins at once (FFMA2 + FADD2); bit-identical to psd_to_db on each half
template <int N>
__device__ __forceinline__ float2 psd_to_db2(float2 psd) {
    constexpr float A = 3.01029995663981195f;
    constexpr double LOG2N = (N == 512) ? 9.0 : (N == 1024) ? 10.0 : (N == 2048) ? 11.0 : (N == 4096) ? 12.0 : 0.0;
    constexpr float B = (float)(13.0102999566398120 - 20.0 * LOG2N * 0.30102999566398120);
    const float2 lg = make_float2(fast_log2(psd.x), fast_log2(psd.y));
    const float2 t = __ffma2_rn(make_float2(A, A), lg, make_float2(B, B));
    return __fadd2_rn(t, make_float2(120.0f, 120.0f));
}

// TW2R: keep the pass-2 twiddles in registers too (15 fewer shared-memory loads per thread and block; ptxas
// still fits 128 registers = 4 resident CTAs per SM with a 12-byte spill)
// IN_I16: the IQ blocks are KiwiSDR wire bytes (4 per sample); the conversion is fused into the pass-1 load.
template <int N, bool DEBUG_STORE, bool HAS_WINDOW, bool TW2R, bool IN_I16 = false>
__global__ void __launch_bounds__(K1Geom<N>::CTA_THREADS, SDR_K1_MINB) k1_spectral_kernel(const K1Args a) {
    using Gm = K1Geom<N>;
    constexpr int M = Gm::M, R3 = Gm::R3, T = Gm::T, PAIRS = Gm::PAIRS, S1 = Gm::S1, S2 = Gm::S2;
    constexpr int NSTAGE = Gm::NSTAGE, G = Gm::G;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int g = threadIdx.x / T;  // group within CTA
    const int t = threadIdx.x % T;  // thread within group
    unsigned char *base = smem_raw + (size_t)g * Gm::GROUP_BYTES;
    unsigned char *stage_base = base;
    float2 *E1 = reinterpret_cast<float2 *>(base + NSTAGE * Gm::STAGE_BYTES);
    float *PSD = reinterpret_cast<float *>(E1);  // aliases E1 (E1 is dead after pass 2 loads)
    float2 *TW2 = reinterpret_cast<float2 *>(base + NSTAGE * Gm::STAGE_BYTES + Gm::E1_BYTES);
    double *NFS1 = reinterpret_cast<double *>(base + NSTAGE * Gm::STAGE_BYTES + Gm::E1_BYTES + Gm::TW2_BYTES);
    double *NFS2 = NFS1 + Gm::NFB * 10;
    float *NFX = reinterpret_cast<float *>(NFS2 + Gm::NFB * 10);
    float2 *PART = reinterpret_cast<float2 *>(NFX + Gm::NFB * 10);
    int *LB = reinterpret_cast<int *>(PART + Gm::PART_SLOTS);
    uint64_t *FULL = reinterpret_cast<uint64_t *>(LB + Gm::LMAX);

    const int group_id = blockIdx.x * G + g;
    const int group_stride = gridDim.x * G;

    // ---- one-time setup ----
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) mbar_init(&FULL[s], 1);
        fence_mbar_init();
    }
    for (int i = t; i < 15 * R3; i += T) TW2[i] = a.tw2[i];
    float2 tw1[15];
#pragma unroll
    for (int k = 0; k < 15; k++) tw1[k] = __ldg(&a.tw1[k * M + t]);
    float win[16];
    if (HAS_WINDOW) {
#pragma unroll
        for (int m = 0; m < 16; m++) win[m] = __ldg(&a.window[m * M + t]);
    }
    group_sync<T, G>(g);

    // ---- producer iterator (thread 0 of the group runs NSTAGE items ahead) ----
    int pseg = group_id, pblk = 0, pn = 0;
    if (pseg < a.n_segs) pn = __ldg(&a.segs[pseg].n_blocks);
    uint32_t issued = 0;
    auto issue_next = [&]() {
        if (pseg >= a.n_segs) return;
        constexpr uint32_t BLOCK_BYTES = IN_I16 ? 4 * N : 8 * N;
        const unsigned char *src = reinterpret_cast<const unsigned char *>(a.segs[pseg].iq) + (size_t)pblk * BLOCK_BYTES;
        int s = issued % NSTAGE;
        mbar_expect_tx(&FULL[s], BLOCK_BYTES);
        tma_load_1d(stage_base + (size_t)s * Gm::STAGE_BYTES, src, BLOCK_BYTES, &FULL[s]);
        issued++;
        pblk++;
        if (pblk == pn) {
            pseg += group_stride;
            pblk = 0;
            if (pseg < a.n_segs) pn = a.segs[pseg].n_blocks;
        }
    };
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) issue_next();
    }

    // pass-2 role of this thread
    const int k1_p2 = t / R3, n3_p2 = t % R3;
    float2 tw2r[TW2R ? 15 : 1];
    if (TW2R) {
#pragma unroll
        for (int k = 0; k < 15; k++) tw2r[k] = TW2[k * R3 + n3_p2];
    }
    uint32_t item = 0;

    for (int seg = group_id; seg < a.n_segs; seg += group_stride) {
        const Segment sg = a.segs[seg];
        const WorkParams wp = a.works[sg.work];
        const int L = wp.n_listeners;
        for (int i = t; i < L; i += T) LB[i] = a.listener_bins[wp.listener_off + i];

        // noise-floor window geometry (dsp/fft.go:216,224)
        const int e = wp.edge_width;
        const int ws = nf_window_size(N, e);
        const int n_win = nf_window_count(N, e);  // the 10th window only closes if a later bin exists
        // this thread's share of the window sums: TPW threads per window, <= nf_len contiguous bins each
        constexpr int TPW = Gm::TPW;
        // nf_part: slot of this thread's partial sums in PART ([window][PART_PITCH]); threads without a share write zeros behind it
        int nf_lo = 0, nf_len = 0, nf_part = 10 * Gm::PART_PITCH + max(t - 10 * Gm::TPW, 0);
        if (t < 10 * TPW) {
            const int grp = t / TPW, part = t - grp * TPW;
            const int w = wp.nf_map[grp];
            nf_part = w * Gm::PART_PITCH + part;
            // an ODD share length makes the lane stride odd: the strided reads below are bank-conflict free
            // within a window (lanes of neighbouring windows can still collide)
            const int per = ((ws + TPW - 1) / TPW) | 1;
            nf_lo = e + w * ws + part * per;
            const int hi = min(nf_lo + per, e + (w + 1) * ws);
            nf_len = max(hi - nf_lo, 0);
        }

        // cumulation registers in pass-3 register order: cum2[i][q] holds positions p = 2q, 2q+1 of pair i,
        // i.e. bins kk = (c + 256*OutIdx<R3>(p) + N/2) % N with c = t + i*T
        float2 cum2[PAIRS][R3 / 2];
        if (sg.state_in >= 0) {
            const float *cs = a.cum_state + (size_t)sg.state_in * N;
#pragma unroll
            for (int i = 0; i < PAIRS; i++)
#pragma unroll
                for (int q = 0; q < R3 / 2; q++) {
                    cum2[i][q].x = cs[((t + i * T) + 256 * OutIdx<R3>::of(2 * q) + N / 2) % N];
                    cum2[i][q].y = cs[((t + i * T) + 256 * OutIdx<R3>::of(2 * q + 1) + N / 2) % N];
                }
        } else {
#pragma unroll
            for (int i = 0; i < PAIRS; i++)
#pragma unroll
                for (int q = 0; q < R3 / 2; q++) cum2[i][q] = make_float2(0.f, 0.f);
        }
        group_sync<T, G>(g);  // LB visible
        int nf_fill = 0, nf_first = sg.block_out;  // warp 0: batched noise-floor selection
        // noise floor, phase 2, spread over all warps: warp q owns windows q, q+NW, q+2NW, ...; its lanes
        // 3*slot .. 3*slot+2 add the TPW partial sums of window w = q + NW*slot and lane 3*slot keeps the result
        // (and x_to = psd[first bin of the next window], read while PSD is still alive) for a batched selection:
        // every NFB blocks lane l of warp 0 runs dsp.FindNoiseFloor's sequential selection (dsp/fft.go:217-251)
        // for block l of the batch.
        constexpr int NW = T / 32;
        constexpr int NF_SLOTS = (10 + NW - 1) / NW;
        const int nf_slot = (t & 31) / 3, nf_j = (t & 31) - 3 * nf_slot;
        const int nf_w = (t >> 5) + NW * nf_slot;
        const bool nf_owner = nf_slot < NF_SLOTS && nf_w < n_win;
        float x_to = 0.f;
        auto nf_phase2 = [&]() {
            float s1 = 0.f, s2 = 0.f;
            if (nf_owner) {
                const float2 *pp = PART + nf_w * Gm::PART_PITCH + nf_j;
#pragma unroll
                for (int m = 0; m < TPW / 3; m++) {
                    const float2 pr = pp[3 * m];
                    s1 += pr.x;
                    s2 += pr.y;
                }
            }
            s1 += __shfl_down_sync(0xffffffffu, s1, 1) + __shfl_down_sync(0xffffffffu, s1, 2);
            s2 += __shfl_down_sync(0xffffffffu, s2, 1) + __shfl_down_sync(0xffffffffu, s2, 2);
            if (nf_owner && nf_j == 0) {
                NFS1[nf_fill * 10 + nf_w] = (double)s1;
                NFS2[nf_fill * 10 + nf_w] = (double)s2;
                NFX[nf_fill * 10 + nf_w] = x_to;
            }
            nf_fill++;
        };
        auto nf_select = [&]() {  // after a group barrier that follows nf_phase2; counters stay uniform in the group
            if (t < nf_fill) nf_select_serial(NFS1 + t * 10, NFS2 + t * 10, NFX + t * 10, 1, ws, n_win,
                                              &a.psd_floor[nf_first + t], &a.variance[nf_first + t]);
            nf_first += nf_fill;
            nf_fill = 0;
        };

        for (int blk = 0; blk < sg.n_blocks; blk++, item++) {
            const int s = item % NSTAGE;
            const uint32_t parity = (item / NSTAGE) & 1u;
            float2 *IN = reinterpret_cast<float2 *>(stage_base + (size_t)s * Gm::STAGE_BYTES);
            float2 *E2 = IN;  // aliases the stage once pass 1 has consumed it
            const int ob = sg.block_out + blk;

            mbar_wait(&FULL[s], parity);

            // ---------------- pass 1: radix-16 over n1, column j = t ----------------
            float2 v[16];
            // loads are issued in the order the first butterfly layer consumes them (b, b+4, b+8, b+12)
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int m = (q & 3) * 4 + (q >> 2);
                if (IN_I16) v[m] = kiwi_decode_sample(reinterpret_cast<const uint32_t *>(IN)[m * M + t]);
                else v[m] = IN[m * M + t];
                if (HAS_WINDOW) v[m] = __fmul2_rn(v[m], make_float2(win[m], win[m]));
            }
            dft16(v);
#pragma unroll
            for (int p = 0; p < 16; p++) {
                const int k1 = OutIdx<16>::of(p);
                if (k1 > 0) v[p] = cmul(v[p], tw1[k1 - 1]);
            }
            // B4 (of the previous block): its partial sums are complete and nobody reads PSD any more, so E1
            // (which PSD aliases) may be overwritten.  Everything above overlaps the slower warps' tail.
            if (blk > 0) group_sync<T, G>(g);
#pragma unroll
            for (int p = 0; p < 16; p++) E1[OutIdx<16>::of(p) * S1 + t] = v[p];
            group_sync<T, G>(g);  // B1: E1 complete, IN consumed
            if (blk > 0) nf_phase2();  // previous block's window sums; PART is rewritten only after B3

            // ---------------- pass 2: radix-16 over n2 for (k1, n3) ----------------
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int n2 = (q & 3) * 4 + (q >> 2);
                v[n2] = E1[k1_p2 * S1 + n2 * R3 + n3_p2];
            }
            dft16(v);
#pragma unroll
            for (int p = 0; p < 16; p++) {
                const int k2 = OutIdx<16>::of(p);
                float2 x = v[p];
                if (k2 > 0) x = cmul(x, TW2R ? tw2r[k2 - 1] : TW2[(k2 - 1) * R3 + n3_p2]);
                E2[n3_p2 * S2 + k1_p2 + 16 * k2] = x;
            }
            group_sync<T, G>(g);  // B2: E2 complete, E1 dead
            if (nf_fill == Gm::NFB) nf_select();  // all warps' NFS stores of the batch happened before B2

            // ---------------- pass 3: radix-R3 over n3 for pairs c = t + i*T ----------------
#pragma unroll
            for (int i = 0; i < PAIRS; i++) {
                const int c = t + i * T;
                float2 u[R3];
#pragma unroll
                for (int q = 0; q < R3; q++) {  // issue order = consumption order of the first butterfly layer
                    const int n3 = (R3 == 16) ? (q & 3) * 4 + (q >> 2) : (R3 == 8) ? (q & 1) * 4 + (q >> 1) : q;
                    u[n3] = E2[n3 * S2 + c];
                }
                dftR<R3>(u);
#pragma unroll
                for (int q = 0; q < R3 / 2; q++) {
                    // two bins at a time: |X|^2 (dsp/fft.go:71-73), dB + 120 (rx/receiver.go:376-378)
#ifdef SDR_K1_PSD_PACKED
                    const float2 sq0 = __fmul2_rn(u[2 * q], u[2 * q]);
                    const float2 sq1 = __fmul2_rn(u[2 * q + 1], u[2 * q + 1]);
                    const float2 psd = make_float2(sq0.x + sq0.y, sq1.x + sq1.y);
#else
                    // FMUL + FFMA per bin: 2 FP32-pipe cycles instead of 3 (FMUL2 + FADD); one rounding fewer than
                    // the reference's two float64 products, far below the fp32-FFT error of the bin
                    const float2 psd = make_float2(fmaf(u[2 * q].x, u[2 * q].x, u[2 * q].y * u[2 * q].y),
                                                   fmaf(u[2 * q + 1].x, u[2 * q + 1].x, u[2 * q + 1].y * u[2 * q + 1].y));
#endif
                    const int kk0 = (c + 256 * OutIdx<R3>::of(2 * q) + N / 2) % N;  // dsp/fft.go:54-57 fftshift
                    const int kk1 = (c + 256 * OutIdx<R3>::of(2 * q + 1) + N / 2) % N;
                    PSD[kk0] = psd.x;
                    PSD[kk1] = psd.y;
                    const float2 db = psd_to_db2<N>(psd);
                    cum2[i][q] = __fadd2_rn(cum2[i][q], db);  // rx/receiver.go:404-406
                    if (DEBUG_STORE) {
                        a.dbg_spectrum[(size_t)ob * N + kk0] = db.x;
                        a.dbg_spectrum[(size_t)ob * N + kk1] = db.y;
                        a.dbg_psd[(size_t)ob * N + kk0] = psd.x;
                        a.dbg_psd[(size_t)ob * N + kk1] = psd.y;
                    }
                }
            }
            group_sync<T, G>(g);  // B3: PSD complete, stage (E2) dead

            // stage s is free again: prefetch item + NSTAGE into it
            if (t == 0) {
                fence_proxy_async();
                issue_next();
            }

            // ---------------- noise floor, phase 1: per-thread partial sums of x and x^2 over a
            // contiguous share of one window (float32 inside the <=17-bin share, float64 across shares)
            {
                // branch-free: a share never exceeds NF_MAX bins for the supported edge widths (host-checked), all
                // loads are issued up front, bins beyond the share contribute +0; four independent packed chains
                float2 a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f), b1 = a1, b2 = a1;
                const float *pp = PSD + nf_lo;
                float xs[Gm::NF_MAX];
#pragma unroll
                for (int i = 0; i < Gm::NF_MAX; i++) xs[i] = (i < nf_len) ? pp[i] : 0.f;
#pragma unroll
                for (int i = 0; i + 3 < Gm::NF_MAX; i += 4) {
                    const float2 x01 = make_float2(xs[i], xs[i + 1]), x23 = make_float2(xs[i + 2], xs[i + 3]);
                    a1 = __fadd2_rn(a1, x01);
                    a2 = __ffma2_rn(x01, x01, a2);
                    b1 = __fadd2_rn(b1, x23);
                    b2 = __ffma2_rn(x23, x23, b2);
                }
                float s1 = (a1.x + a1.y) + (b1.x + b1.y), s2 = (a2.x + a2.y) + (b2.x + b2.y);
#pragma unroll
                for (int i = Gm::NF_MAX / 4 * 4; i < Gm::NF_MAX; i++) {
                    s1 += xs[i];
                    s2 = fmaf(xs[i], xs[i], s2);
                }
                PART[nf_part] = make_float2(s1, s2);
            }
            // x_to = psd[first bin of the next window] is read here (PSD is gone after the next B4);
            // lane 3w of warp 0 owns window w in phase 2
            if (nf_owner && nf_j == 0) x_to = PSD[e + (nf_w + 1) * ws];
            // listener taps (rx/receiver.go:393): same dB function as the owner thread; served from the last
            // threads of the group so that warp 0 (TMA issue) is not the straggler
            for (int l = T - 1 - t; l < L; l += T) a.taps[(size_t)ob * a.tap_stride + l] = psd_to_db<N>(PSD[LB[l]]);
        }
        group_sync<T, G>(g);  // B4 of the last block
        nf_phase2();
        group_sync<T, G>(g);
        nf_select();

        // ---- end of segment: flush or save the cumulation ----
        float *dst = (sg.flush_idx >= 0) ? a.flush_cum + (size_t)sg.flush_idx * N : a.cum_state + (size_t)sg.state_out * N;
#pragma unroll
        for (int i = 0; i < PAIRS; i++)
#pragma unroll
            for (int q = 0; q < R3 / 2; q++) {
                dst[((t + i * T) + 256 * OutIdx<R3>::of(2 * q) + N / 2) % N] = cum2[i][q].x;
                dst[((t + i * T) + 256 * OutIdx<R3>::of(2 * q + 1) + N / 2) % N] = cum2[i][q].y;
            }
        // after a flush the state row is never read: the next segment of the stream has state_in = -1
        group_sync<T, G>(g);  // LB / PART reuse by the next segment
    }
}

}  // namespace sdr
