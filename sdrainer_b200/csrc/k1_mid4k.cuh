// k1_mid4k.cuh -- fused spectral front end for N = 4096: one pass over the IQ, one launch, TMA-staged blocks, the
// cumulation in registers -- the K1 contract (k1_spectral.cuh) at 32 KB blocks, built like k1_mid8k_kernel.
//
// Reference arithmetic: dsp/fft.go:23-85 (FFT, fftshift, |X|^2, dB + 120), dsp/fft.go:215-252 (FindNoiseFloor),
// rx/receiver.go:393 (listener taps), rx/receiver.go:404-407 (cumulation, float32, block order).
//
// Round 1's kernel for this size loaded global -> registers directly (nothing in flight while it computed) and re-read its
// step twiddles from L1 per block: 43 % of the HBM roofline.  Here a CTA of 256 threads owns a segment (<= 100 consecutive
// blocks of one stream), two CTAs per SM:
//   staging   thread 0 keeps NSTAGE whole blocks (32 KB each) in flight with cp.async.bulk (TMA, SASS UBLKCP) into a
//             shared-memory ring guarded by mbarriers -- block b+1 lands while block b is transformed;
//   pass A    4096 = 16 x 256: thread c takes column c (x[256 m + c], m < 16) from the stage, runs the 16-point
//             transform in registers and multiplies by the sixteen step twiddles W_N^(c k1), which stay in registers
//             for the whole kernel (c never changes);
//   pass B    half-warp f = row k1: the 256-point half-warp transform of k1_large.cuh on E[k1][.]; lane hl ends with
//             X[k1 + 16 k2], k2 = hl + 16 q;
//   epilogue  in registers: |X|^2, dB, cumulation (16 bins per thread, sequential float32 adds in block order); |X|^2
//             is parked in the dead row storage (plane_skew layout) for the noise windows, x_to and the taps;
//   noise     thread (window w, row k1) of warps 0-4 sums the <= 26 bins of window w that live in row k1
//             (nf_row_share), a float64 half-warp reduction over the 16 rows gives the window's (sum x, sum x^2);
//             dsp.FindNoiseFloor's sequential selection runs batched, one lane per block, every 8 blocks.
// Three CTA barriers per block; the next block's pass A (stage reads only) overlaps this block's noise phase.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "k1_large.cuh"

namespace sdr {

template <int NSTAGE>
struct K1Mid4kGeom {
    static constexpr int N = 4096, T = 256;
    static constexpr int STAGE_BYTES = N * 8;
    static constexpr int E_BYTES = 16 * HW_PITCH * 8;  // 34 944
    static constexpr int TW256_BYTES = 256 * 8;
    static constexpr int NFB = 8;                      // blocks per batched noise-floor selection
    static constexpr int NF_BYTES = NFB * 10 * (8 + 8 + 4);
    static constexpr int NF_M = 25;                    // whole positions of one noise window in one row: 409 / 16
    static constexpr int OFF_E = NSTAGE * STAGE_BYTES;
    static constexpr int OFF_TW = OFF_E + E_BYTES;
    static constexpr int OFF_NF = OFF_TW + TW256_BYTES;
    static constexpr int OFF_BAR = OFF_NF + NF_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + NSTAGE * 8 + 64;
};

template <int NSTAGE, bool DEBUG_STORE, bool HAS_WINDOW>
__global__ void __launch_bounds__(256, 2) k1_mid4k_kernel(const K1Args a, const float2 *__restrict__ tw_step,
                                                           const float2 *__restrict__ tw256) {
    using Gm = K1Mid4kGeom<NSTAGE>;
    constexpr int N = Gm::N, NFB = Gm::NFB;
    extern __shared__ __align__(128) unsigned char m4_smem[];
    float2 *E = reinterpret_cast<float2 *>(m4_smem + Gm::OFF_E);      // [16][HW_PITCH]
    float *Ef = reinterpret_cast<float *>(m4_smem + Gm::OFF_E);       // |X|^2 of bin kk at Ef[plane_of(kk & 15) + (kk >> 4)]
    float2 *TW = reinterpret_cast<float2 *>(m4_smem + Gm::OFF_TW);    // [15][16] W_256^(hl k)
    double *NFS1 = reinterpret_cast<double *>(m4_smem + Gm::OFF_NF);  // [NFB][10]
    double *NFS2 = NFS1 + NFB * 10;
    float *NFX = reinterpret_cast<float *>(NFS2 + NFB * 10);
    uint64_t *FULL = reinterpret_cast<uint64_t *>(m4_smem + Gm::OFF_BAR);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = tid;                          // pass A: column c
    const int f = tid >> 4, hl = tid & 15;      // pass B: row k1 = f, lane hl of its half-warp
    const float db_offset = (float)(13.0102999566398120 - 20.0 * 12.0 * 0.30102999566398120);  // 10 log10(20) - 20 log10(N)
    auto to_db = [&](float psd) -> float { return __fadd_rn(fmaf(3.01029995663981195f, fast_log2(psd), db_offset), 120.0f); };
    // the |X|^2 plane of row r starts at word plane_of(r) of E, inside the row's own column (k1_large.cuh: plane_skew)
    auto plane_of = [](int r) { return r * (2 * HW_PITCH) + plane_skew(r * (2 * HW_PITCH), r); };
    auto psd_at = [&](int kk) -> float { return Ef[plane_of(kk & 15) + (kk >> 4)]; };

    // ---- one-time setup ----
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) mbar_init(&FULL[s], 1);
        fence_mbar_init();
    }
    if (tid < 240) TW[tid] = __ldg(&tw256[((tid & 15) * ((tid >> 4) + 1)) & 255]);  // [k - 1][hl] = W_256^(hl k)
    // step twiddles of this thread's sixteen outputs: register p of dft16 holds k1 = OutIdx<16>(p)
    float2 tws[16];
#pragma unroll
    for (int p = 0; p < 16; p++) tws[p] = __ldg(&tw_step[OutIdx<16>::of(p) * 256 + c]);
    __syncthreads();

    // ---- producer iterator (thread 0 runs NSTAGE blocks ahead, across segment boundaries) ----
    int pseg = blockIdx.x, pblk = 0, pn = 0;
    if (pseg < a.n_segs) pn = __ldg(&a.segs[pseg].n_blocks);
    uint32_t issued = 0;
    auto issue_next = [&]() {
        if (pseg >= a.n_segs) return;
        const unsigned char *src = reinterpret_cast<const unsigned char *>(a.segs[pseg].iq) + (size_t)pblk * Gm::STAGE_BYTES;
        const int s = issued % NSTAGE;
        mbar_expect_tx(&FULL[s], Gm::STAGE_BYTES);
        tma_load_1d(m4_smem + (size_t)s * Gm::STAGE_BYTES, src, Gm::STAGE_BYTES, &FULL[s]);
        issued++;
        if (++pblk == pn) {
            pseg += gridDim.x;
            pblk = 0;
            if (pseg < a.n_segs) pn = a.segs[pseg].n_blocks;
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) issue_next();
    }
    uint32_t item = 0;

    for (int seg = blockIdx.x; seg < a.n_segs; seg += gridDim.x) {
        const Segment sg = a.segs[seg];
        const WorkParams wp = a.works[sg.work];
        const int L = wp.n_listeners;
        const int *lbins = a.listener_bins + wp.listener_off;
        const int e = wp.edge_width;
        const int ws = nf_window_size(N, e), n_win = nf_window_count(N, e);
        // noise floor: thread (window w = 2 warp + lane/16, row j = lane % 16) of warps 0-4; the window's bins in row k1 = j
        // are the positions [first(k1), last(k1)) (bin kk = k1 + 16 p).  All sixteen rows of a window are read from the
        // position nf_p of the LAST row, the row's own range is [nf_lo, nf_hi) relative to it (k1_large.cuh: nf_row_share)
        int nf_p = 0, nf_lo = 0, nf_hi = 0, nf_rot = 0;
        const int nf_w = 2 * warp + (lane >> 4);
        if (warp < 5) {
            auto first = [&](int w, int k1) { const int lo = e + w * ws; return lo < k1 ? 0 : (lo - k1 + 15) >> 4; };
            const int k1 = lane & 15, hi = e + (nf_w + 1) * ws;
            nf_p = first(nf_w, 15);
            nf_lo = first(nf_w, k1) - nf_p;
            const int p1 = hi <= k1 ? 0 : min(256, (hi - k1 + 15) >> 4);
            nf_hi = max(p1 - nf_p, 0);
            nf_rot = (lane >> 4) & ~(first(2 * warp + 1, 15) - first(2 * warp, 15)) & 1;
        }
        int nf_fill = 0, nf_first = sg.block_out;

        // cumulation registers: cum[p] is bin kk = f + 16*((hl + 16*OutIdx<16>(p) + 128) & 255)
        float cum[16];
#pragma unroll
        for (int p = 0; p < 16; p++) {
            const int kk = f + 16 * ((hl + 16 * OutIdx<16>::of(p) + 128) & 255);
            cum[p] = sg.state_in >= 0 ? a.cum_state[(size_t)sg.state_in * N + kk] : 0.f;
        }

        for (int blk = 0; blk < sg.n_blocks; blk++, item++) {
            const int s = item % NSTAGE;
            const uint32_t parity = (item / NSTAGE) & 1u;
            const float2 *IN = reinterpret_cast<const float2 *>(m4_smem + (size_t)s * Gm::STAGE_BYTES);
            const int ob = sg.block_out + blk;
            mbar_wait(&FULL[s], parity);

            // ---------------- pass A: column c ----------------
            float2 v[16];
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int m = (q & 3) * 4 + (q >> 2);  // issue order = consumption order of dft16's first layer
                v[m] = IN[m * 256 + c];
                if (HAS_WINDOW) {
                    const float w = __ldg(&a.window[m * 256 + c]);
                    v[m] = __fmul2_rn(v[m], make_float2(w, w));
                }
            }
            dft16(v);
#pragma unroll
            for (int p = 1; p < 16; p++) v[p] = cmul(v[p], tws[p]);  // register 0 is k1 = 0: W^0
            // B0: every thread has consumed stage s (it can be refilled) and, for blk > 0, the previous block's
            // noise-floor / tap reads of E are done (E can be overwritten)
            __syncthreads();
            if (tid == 0) {
                fence_proxy_async();
                issue_next();
            }
            if (nf_fill == NFB) {  // batched dsp.FindNoiseFloor selection (dsp/fft.go:217-251); NFS is rewritten after B2
                if (warp == 7 && lane < NFB)
                    nf_select_serial(NFS1 + lane * 10, NFS2 + lane * 10, NFX + lane * 10, 1, ws, n_win,
                                     &a.psd_floor[nf_first + lane], &a.variance[nf_first + lane]);
                nf_first += NFB;
                nf_fill = 0;
            }
#pragma unroll
            for (int p = 0; p < 16; p++) E[OutIdx<16>::of(p) * HW_PITCH + c] = v[p];
            __syncthreads();  // B1: E complete

            // ---------------- pass B: half-warp f = row k1 ----------------
            {
                float2 *col = E + f * HW_PITCH;
                // four table reads (W256^(hl), ^(2 hl), ^(4 hl), ^(8 hl)) and eleven products instead of fifteen reads: this kernel
                // is bound by the shared-memory pipe and has registers to spare (47.7 -> 50.7 % of the HBM roofline; the
                // same change costs k1_mid8k2, at its 128-register cap, 5 %; keeping the four resident across blocks: 50.4 %)
                const float2 w1 = TW[hl], w2 = TW[16 + hl], w4 = TW[48 + hl], w8 = TW[112 + hl];
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int n1 = (q & 3) * 4 + (q >> 2);
                    v[n1] = col[16 * n1 + hl];
                }
                fft256_halfwarp_regs_p2(v, col, w1, w2, w4, w8, hl);
                __syncwarp();  // transpose reads done: the row storage may take the |X|^2 values
                float *prow = Ef + plane_of(f);
#pragma unroll
                for (int p = 0; p < 16; p++) {
                    const int k2s = (hl + 16 * OutIdx<16>::of(p) + 128) & 255;  // fftshift (dsp/fft.go:54-57)
                    const float psd = fmaf(v[p].x, v[p].x, v[p].y * v[p].y);    // dsp/fft.go:71-73
                    const float db = to_db(psd);                                 // rx/receiver.go:376-378
                    cum[p] = __fadd_rn(cum[p], db);                              // rx/receiver.go:404-406
                    prow[k2s] = psd;
                    if (DEBUG_STORE) {
                        const int kk = f + 16 * k2s;
                        a.dbg_spectrum[(size_t)ob * N + kk] = db;
                        a.dbg_psd[(size_t)ob * N + kk] = psd;
                    }
                }
            }
            __syncthreads();  // B2: |X|^2 complete

            // ---------------- dsp.FindNoiseFloor (dsp/fft.go:215-252): window sums ----------------
            if (warp < 5) {
                float s1, s2;
                nf_row_share<Gm::NF_M>(Ef, plane_of(lane & 15) + nf_p, nf_lo, nf_hi, nf_rot, max(ws >> 4, 2), s1, s2);
                double d1 = (double)s1, d2 = (double)s2;  // float32 inside the share of one row, float64 across the 16 rows
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
                    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
                }
                if ((lane & 15) == 0 && nf_w < n_win) {
                    NFS1[nf_fill * 10 + nf_w] = d1;
                    NFS2[nf_fill * 10 + nf_w] = d2;
                    NFX[nf_fill * 10 + nf_w] = psd_at(e + (nf_w + 1) * ws);  // x_to (dsp/fft.go:238-243)
                }
            } else {
                // listener taps (rx/receiver.go:393): same dB function as the owner thread
                for (int l = tid - 160; l < L; l += 96) a.taps[(size_t)ob * a.tap_stride + l] = to_db(psd_at(__ldg(&lbins[l])));
            }
            nf_fill++;
        }
        __syncthreads();  // the last block's window sums are in NFS; its reads of E are done
        if (warp == 7 && lane < nf_fill)
            nf_select_serial(NFS1 + lane * 10, NFS2 + lane * 10, NFX + lane * 10, 1, ws, n_win, &a.psd_floor[nf_first + lane],
                             &a.variance[nf_first + lane]);

        float *dst = (sg.flush_idx >= 0) ? a.flush_cum + (size_t)sg.flush_idx * N : a.cum_state + (size_t)sg.state_out * N;
#pragma unroll
        for (int p = 0; p < 16; p++) dst[f + 16 * ((hl + 16 * OutIdx<16>::of(p) + 128) & 255)] = cum[p];
        __syncthreads();  // NFS is free for the next segment
    }
}

}  // namespace sdr
