// k1_mid.cuh -- fused spectral front end for N = R1 x 256, R1 = 16 | 32 (N = 4096, 8192): one pass over the IQ,
// one launch, the cumulation in registers -- the K1 contract (k1_spectral.cuh) at block sizes whose exchange no
// longer fits the three-pass kernel's register/shared-memory budget.
//
// Reference arithmetic: dsp/fft.go:23-85 (FFT, fftshift, |X|^2, dB + 120), dsp/fft.go:215-252 (FindNoiseFloor),
// rx/receiver.go:393 (listener taps), rx/receiver.go:404-407 (cumulation, float32, block order).
//
// A CTA of 256 threads owns a segment (<= 100 consecutive blocks of one stream) and walks it in order:
//   pass A  thread c = column: x[r*256 + c], r < R1, straight from global memory (2 KB contiguous per r and CTA), the
//           whole R1-point transform in registers (dft16 / dft32), twiddle W_N^(c k1), row-major store E[k1][c];
//   pass B  half-warp f = row k1 (= f, f + 16): the 256-point transform of fft256_halfwarp_regs (k1_large.cuh) on
//           E[k1][.], transposing through the row's own storage; lane hl ends with X[k1 + R1*k2], k2 = hl + 16 q;
//   epilogue in registers: |X|^2, dB, cumulation (N / 256 bins per thread, sequential float32 adds in block order);
//           |X|^2 is parked in the (dead) row storage so that the ten noise windows, x_to and the taps read it.
// Three CTA barriers per block.  The next block's IQ is prefetched into L2 while the current one is transformed.
// Used when a launch has enough segments to fill the GPU (engine.cu); otherwise N = 8192 takes the block-parallel
// two-kernel path of k1_large.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "k1_large.cuh"

#ifndef SDR_K1M_MINB16
#define SDR_K1M_MINB16 2
#endif

namespace sdr {

template <int R1>
struct K1MidGeom {
    static constexpr int N = R1 * 256;
    static constexpr int T = (R1 >= 16) ? 256 : 16 * R1;  // threads per CTA: one half-warp per row (two rows at R1 = 32)
    static constexpr int CPT = 256 / T;                // columns per thread in pass A
    static constexpr int JR = (R1 + 15) / 16;          // rows per half-warp
    static constexpr int LOG_R1 = (R1 == 8) ? 3 : (R1 == 16) ? 4 : 5;
    static constexpr int MINB = (R1 == 16) ? SDR_K1M_MINB16 : 2;
    static constexpr int E_BYTES = R1 * HW_PITCH * 8;  // 34 944 / 69 888
    static constexpr int TPW = (T / 10) / 3 * 3;       // noise floor: threads per window (3 lanes x TPW/3 partials combine)
    static constexpr int NFB = 8;                      // noise floor: blocks per batched selection
    static constexpr int PART_BYTES = T * 8;
    static constexpr int NF_BYTES = NFB * 10 * (8 + 8 + 4);
    static constexpr int SMEM_BYTES = E_BYTES + PART_BYTES + NF_BYTES + 64;
    static_assert(R1 == 16 || R1 == 32, "N = 4096 or 8192");
};

template <int R1>
__device__ __forceinline__ void dft_r1(float2 (&v)[R1]);
template <>
__device__ __forceinline__ void dft_r1<16>(float2 (&v)[16]) { dft16(v); }
template <>
__device__ __forceinline__ void dft_r1<32>(float2 (&v)[32]) { dft32(v); }

template <int R1, bool DEBUG_STORE, bool HAS_WINDOW>
__global__ void __launch_bounds__(K1MidGeom<R1>::T, K1MidGeom<R1>::MINB) k1_mid_kernel(const K1Args a, const float2 *__restrict__ tw_step,
                                                        const float2 *__restrict__ tw256) {
    using Gm = K1MidGeom<R1>;
    constexpr int N = Gm::N, JR = Gm::JR, T = Gm::T, CPT = Gm::CPT;
    extern __shared__ __align__(16) unsigned char mid_smem[];
    float2 *E = reinterpret_cast<float2 *>(mid_smem);  // [R1][HW_PITCH]
    float *Ef = reinterpret_cast<float *>(mid_smem);   // |X|^2 of bin kk lives at Ef[(kk % R1) * 2*HW_PITCH + kk / R1]
    float2 *PART = reinterpret_cast<float2 *>(mid_smem + Gm::E_BYTES);  // per-thread (sum x, sum x^2) of a window share
    double *NFS1 = reinterpret_cast<double *>(mid_smem + Gm::E_BYTES + Gm::PART_BYTES);  // [NFB][10]
    double *NFS2 = NFS1 + Gm::NFB * 10;
    float *NFX = reinterpret_cast<float *>(NFS2 + Gm::NFB * 10);
    constexpr int TPW = Gm::TPW, NFB = Gm::NFB;
    const int tid = threadIdx.x, hl = tid & 15, f = tid >> 4, lane = tid & 31, warp = tid >> 5;
    const float db_offset = (float)(13.0102999566398120 - 20.0 * (double)(8 + Gm::LOG_R1) * 0.30102999566398120);
    auto psd_at = [&](int kk) -> float { return Ef[(kk & (R1 - 1)) * (2 * HW_PITCH) + (kk >> Gm::LOG_R1)]; };
    auto to_db = [&](float psd) -> float { return __fadd_rn(fmaf(3.01029995663981195f, fast_log2(psd), db_offset), 120.0f); };

    HwTwiddle t;
    if (R1 != 32) hw_twiddle_load(t, tw256, hl);

    for (int seg = blockIdx.x; seg < a.n_segs; seg += gridDim.x) {
        const Segment sg = a.segs[seg];
        const WorkParams wp = a.works[sg.work];
        const int L = wp.n_listeners;
        const int *lbins = a.listener_bins + wp.listener_off;
        const int e = wp.edge_width;
        const int ws = nf_window_size(N, e), n_win = nf_window_count(N, e);
        // noise floor, phase 1: thread (w, part) sums a contiguous share of window w in float32 (as K1: <= 35 bins,
        // combined in float64 per window); phase 2: lanes 3w..3w+2 of warp 0 add the 24 partial sums of window w
        const int per = (ws + TPW - 1) / TPW;
        int nf_lo = 0, nf_len = 0;
        if (tid < 10 * TPW) {
            const int w = tid / TPW, part = tid - w * TPW;
            nf_lo = e + w * ws + part * per;
            const int hi = min(nf_lo + per, e + (w + 1) * ws);
            nf_len = max(hi - nf_lo, 0);
        }
        int nf_fill = 0, nf_first = sg.block_out;
        auto nf_select = [&]() {  // warp 0, after its own phase 2 (same warp: __syncwarp is enough)
            __syncwarp();
            if (lane < nf_fill) nf_select_serial(NFS1 + lane * 10, NFS2 + lane * 10, NFX + lane * 10, 1, ws, n_win,
                                                 &a.psd_floor[nf_first + lane], &a.variance[nf_first + lane]);
            __syncwarp();
        };

        // cumulation registers: cum[j][p] is bin kk = k1 + R1*((k2 + 128) & 255), k1 = f + 16 j, k2 = hl + 16*OutIdx<16>(p)
        float cum[JR][16];
#pragma unroll
        for (int j = 0; j < JR; j++)
#pragma unroll
            for (int p = 0; p < 16; p++) {
                const int kk = (f + 16 * j) + R1 * ((hl + 16 * OutIdx<16>::of(p) + 128) & 255);
                cum[j][p] = sg.state_in >= 0 ? a.cum_state[(size_t)sg.state_in * N + kk] : 0.f;
            }

        const float2 *iq = reinterpret_cast<const float2 *>(sg.iq);
        for (int blk = 0; blk < sg.n_blocks; blk++) {
            const int ob = sg.block_out + blk;
            // ---------------- pass A: columns c = tid + T*cc ----------------
            if (blk + 1 < sg.n_blocks) {  // next block -> L2 (N*8/128 lines, 1 or 2 per thread)
#pragma unroll
                for (int i = 0; i < (N * 8 / 128 + T - 1) / T; i++)
                    if (tid + T * i < N * 8 / 128)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(iq + (size_t)(blk + 1) * N) +
                                                                      (size_t)(tid + T * i) * 128));
            }
            float2 v[CPT][R1];
#pragma unroll
            for (int cc = 0; cc < CPT; cc++) {  // all loads of the block in flight before the first butterfly
                const int c = tid + T * cc;
                const float2 *src = iq + (size_t)blk * N + c;
#pragma unroll
                for (int q = 0; q < R1; q++) {
                    // issue order = consumption order of the first butterfly layer
                    const int r = (R1 == 32) ? (q & 3) * 8 + (q >> 2) : (R1 == 16) ? (q & 3) * 4 + (q >> 2) : (q & 1) * 4 + (q >> 1);
                    v[cc][r] = __ldg(&src[(size_t)r * 256]);
                    if (HAS_WINDOW) {
                        const float w = __ldg(&a.window[r * 256 + c]);
                        v[cc][r] = __fmul2_rn(v[cc][r], make_float2(w, w));
                    }
                }
            }
#pragma unroll
            for (int cc = 0; cc < CPT; cc++) {
                const int c = tid + T * cc;
                dft_r1<R1>(v[cc]);
#pragma unroll
                for (int p = 1; p < R1; p++) {
                    const int k1 = OutIdx<R1>::of(p);
                    if (k1 > 0) v[cc][p] = cmul(v[cc][p], __ldg(&tw_step[k1 * 256 + c]));
                }
            }
            if (blk > 0) __syncthreads();  // the previous block's noise-floor / tap reads of E are done
#pragma unroll
            for (int cc = 0; cc < CPT; cc++)
#pragma unroll
                for (int p = 0; p < R1; p++) E[OutIdx<R1>::of(p) * HW_PITCH + tid + T * cc] = v[cc][p];
            __syncthreads();
            // ---------------- pass B: half-warp f = rows f, f + 16 ----------------
            if (R1 == 32) {
                // N = 8192: pass A needs 64 registers for its column; the 15 row twiddles are re-read (L1 hits) per block
                // instead of staying live across pass A (the laundered pointer keeps the loads inside the loop)
                const float2 *twp = tw256;
                asm volatile("" : "+l"(twp));
                hw_twiddle_load(t, twp, hl);
            }
#pragma unroll
            for (int j = 0; j < JR; j++) {
                const int k1 = f + 16 * j;
                float2 *col = E + k1 * HW_PITCH;
                float2 u[16];
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int n1 = (q & 3) * 4 + (q >> 2);
                    u[n1] = col[16 * n1 + hl];
                }
                fft256_halfwarp_regs(u, col, t, hl);
                __syncwarp();  // transpose reads done: the row storage may take the |X|^2 values
                float *prow = reinterpret_cast<float *>(col);
#pragma unroll
                for (int p = 0; p < 16; p++) {
                    const int k2s = (hl + 16 * OutIdx<16>::of(p) + 128) & 255;  // fftshift (dsp/fft.go:54-57)
                    const float psd = fmaf(u[p].x, u[p].x, u[p].y * u[p].y);    // dsp/fft.go:71-73
                    const float db = to_db(psd);                                 // rx/receiver.go:376-378
                    cum[j][p] = __fadd_rn(cum[j][p], db);                        // rx/receiver.go:404-406
                    prow[k2s] = psd;
                    if (DEBUG_STORE) {
                        const int kk = k1 + R1 * k2s;
                        a.dbg_spectrum[(size_t)ob * N + kk] = db;
                        a.dbg_psd[(size_t)ob * N + kk] = psd;
                    }
                }
            }
            __syncthreads();
            // ---------------- dsp.FindNoiseFloor (dsp/fft.go:215-252), phase 1: per-thread share sums ----------------
            {
                float s1 = 0.f, s2 = 0.f;
                for (int i = 0; i < nf_len; i++) {
                    const float x = psd_at(nf_lo + i);
                    s1 += x;
                    s2 = fmaf(x, x, s2);
                }
                PART[tid] = make_float2(s1, s2);
            }
            // listener taps (rx/receiver.go:393)
            for (int l = tid; l < L; l += T) a.taps[(size_t)ob * a.tap_stride + l] = to_db(psd_at(__ldg(&lbins[l])));
            __syncthreads();
            if (warp == 0) {  // phase 2 + batched selection: warp 0 only, everything else moves on to the next block
                const int w = lane / 3, j = lane - 3 * w;
                float s1 = 0.f, s2 = 0.f;
                if (w < 10) {
#pragma unroll
                    for (int m = 0; m < TPW / 3; m++) {
                        const float2 pr = PART[w * TPW + j + 3 * m];
                        s1 += pr.x;
                        s2 += pr.y;
                    }
                }
                s1 += __shfl_down_sync(0xffffffffu, s1, 1) + __shfl_down_sync(0xffffffffu, s1, 2);
                s2 += __shfl_down_sync(0xffffffffu, s2, 1) + __shfl_down_sync(0xffffffffu, s2, 2);
                if (w < n_win && j == 0) {
                    NFS1[nf_fill * 10 + w] = (double)s1;
                    NFS2[nf_fill * 10 + w] = (double)s2;
                    NFX[nf_fill * 10 + w] = psd_at(e + (w + 1) * ws);  // x_to (dsp/fft.go:238-243)
                }
            }
            nf_fill++;
            if (nf_fill == NFB) {
                if (warp == 0) nf_select();
                nf_first += nf_fill;
                nf_fill = 0;
            }
            // the barrier before the next block's E stores (top of the loop) also orders PART / E against warp 0's reads
        }
        if (warp == 0 && nf_fill > 0) nf_select();
        __syncthreads();  // last block's reads of E before the next segment's pass A

        float *dst = (sg.flush_idx >= 0) ? a.flush_cum + (size_t)sg.flush_idx * N : a.cum_state + (size_t)sg.state_out * N;
#pragma unroll
        for (int j = 0; j < JR; j++)
#pragma unroll
            for (int p = 0; p < 16; p++) dst[(f + 16 * j) + R1 * ((hl + 16 * OutIdx<16>::of(p) + 128) & 255)] = cum[j][p];
    }
}

}  // namespace sdr
