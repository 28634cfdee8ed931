// k1_mid8k.cuh -- fused spectral front end for N = 8192 (BASELINE config 3: 768 kS/s contest band): one pass over the
// IQ, one launch, TMA-staged blocks, the cumulation in registers -- the K1 contract (k1_spectral.cuh) at 64 KB blocks.
//
// Reference arithmetic: dsp/fft.go:23-85 (FFT, fftshift, |X|^2, dB + 120), dsp/fft.go:215-252 (FindNoiseFloor),
// rx/receiver.go:393 (listener taps), rx/receiver.go:404-407 (cumulation, float32, block order).
//
// Round 1's kernel for this size (removed) loaded global -> registers directly (nothing in flight while it computed: 33 %
// of the HBM roofline, long-scoreboard bound, 16 warps per SM in two CTAs) and re-read 31 step twiddles per column from
// L1.  Here ONE CTA of 512 threads per SM owns a segment (<= 100 consecutive blocks of one
// stream):
//   staging   thread 0 keeps NSTAGE whole blocks (64 KB each) in flight with cp.async.bulk (TMA, SASS UBLKCP) into a
//             shared-memory ring guarded by mbarriers -- block b+1 lands while block b is transformed;
//   pass A    8192 = 32 x 256.  The 32-point column transform is split by output parity (decimation in frequency)
//             between TWO threads: thread (c, h) forms u[m] = x[m] + (-1)^h x[m+16] (rows m, m+16 of column c, read
//             from the stage), multiplies by W32^m when h = 1 (h is warp-uniform: warps 0-7 / 8-15), runs a 16-point
//             transform in registers and owns the outputs k1 = 2j + h.  Sixteen points and sixteen step twiddles
//             W_N^(c k1) per thread: the twiddles stay in registers for the whole kernel (c and h never change);
//   pass B    half-warp f = row k1: the 256-point half-warp transform of k1_large.cuh on E[k1][.], transposing
//             through the row's own storage; lane hl ends with X[k1 + 32 k2], k2 = hl + 16 q;
//   epilogue  in registers: |X|^2, dB, cumulation (16 bins per thread, sequential float32 adds in block order);
//             |X|^2 is parked in the dead row storage for the noise windows, x_to and the taps;
//   noise     thread (window w, row k1) sums the <= 26 bins of window w that live in row k1 (nf_row_share_tested, k1_large.cuh),
//             a float64 half-warp reduction over a group's 16 rows gives its share of the window's (sum x, sum x^2);
//             large_nf_finish_kernel adds the two groups' shares and runs dsp.FindNoiseFloor's selection.
// Three group barriers per block; the next block's pass A (stage reads only) overlaps this block's noise phase.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "k1_large.cuh"

namespace sdr {

template <int NSTAGE>
struct K1Mid8kGeom {
    static constexpr int N = 8192, T = 512;
    static constexpr int STAGE_BYTES = N * 8;
    static constexpr int E_BYTES = 32 * HW_PITCH * 8;  // 69 888
    static constexpr int TW256_BYTES = 256 * 8;
    static constexpr int NFB = 8;                      // blocks per batched noise-floor selection
    static constexpr int NF_BYTES = NFB * 10 * (8 + 8 + 4);
    static constexpr int NFMAX = 26;                   // bins of one window in one row: ceil(819 / 32)
    static constexpr int OFF_E = NSTAGE * STAGE_BYTES;
    static constexpr int OFF_TW = OFF_E + E_BYTES;
    static constexpr int OFF_NF = OFF_TW + TW256_BYTES;
    static constexpr int OFF_BAR = OFF_NF + NF_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + NSTAGE * 8 + 64;
};

// ---------------------------------------------------------------------------------------------------------------------
// k1_mid8k2_kernel: the transform above as two DECOUPLED groups of 256 threads (the first build ran the CTA as one
// 512-thread group with three CTA-wide barriers per block and the selection inside the kernel: 42.6 % of the HBM roofline
// against 44.4 % for this one on the same box, removed).  Parity h of the column split owns the
// rows k1 = 2 fl + h in pass B as well, so group h (warps 8h .. 8h+7) never touches the other group's half of E: the
// groups only share the TMA stages (both read every stage; the second one to finish a stage re-arms its copy) and run
// out of phase like two CTAs would -- one group's barrier waits and shared-memory bursts are the other's compute time.
// The window sums of the two groups meet in global memory (nf_part[block][h][window]); large_nf_finish_kernel adds them
// and runs dsp.FindNoiseFloor's selection after the launch.
struct Mid8kNf {
    double2 *nf_part;  // [blocks][2][10] (sum x, sum x^2) of the group's rows per noise window
    float *xto;        // [blocks][10] psd[first bin of the next window]
    int *nf_edge;      // [blocks]
};

template <int NSTAGE, bool DEBUG_STORE, bool HAS_WINDOW>
__global__ void __launch_bounds__(512, 1) k1_mid8k2_kernel(const K1Args a, const float2 *__restrict__ tw_step,
                                                            const float2 *__restrict__ tw256, const Mid8kNf nf) {
    using Gm = K1Mid8kGeom<NSTAGE>;
    constexpr int N = Gm::N;
    extern __shared__ __align__(128) unsigned char m8_smem[];
    float2 *E = reinterpret_cast<float2 *>(m8_smem + Gm::OFF_E);      // [32][HW_PITCH]
    float *Ef = reinterpret_cast<float *>(m8_smem + Gm::OFF_E);       // |X|^2 of bin kk at Ef[plane_of(kk & 31) + (kk >> 5)]
    float2 *TW = reinterpret_cast<float2 *>(m8_smem + Gm::OFF_TW);    // [15][16] W_256^(hl k)
    uint64_t *FULL = reinterpret_cast<uint64_t *>(m8_smem + Gm::OFF_BAR);
    int *DONE = reinterpret_cast<int *>(m8_smem + Gm::OFF_NF);        // [NSTAGE] groups that have consumed the stage

    const int tid = threadIdx.x, h = tid >> 8, tg = tid & 255;        // group h (warp-uniform), thread tg of the group
    const int lane = tg & 31, wg = tg >> 5;
    const int c = tg;                                                  // pass A: column c, output parity h
    const int fl = tg >> 4, hl = tg & 15, k1row = 2 * fl + h;         // pass B: row k1row, lane hl of its half-warp
    const float db_offset = (float)(13.0102999566398120 - 20.0 * 13.0 * 0.30102999566398120);  // 10 log10(20) - 20 log10(N)
    auto to_db = [&](float psd) -> float { return __fadd_rn(fmaf(3.01029995663981195f, fast_log2(psd), db_offset), 120.0f); };
    // the |X|^2 plane of row r = 2 fl + h starts at word plane_of(r) of E, inside the row's own column (k1_large.cuh: plane_skew)
    auto plane_of = [](int r) { return r * (2 * HW_PITCH) + plane_skew(r * (2 * HW_PITCH), r >> 1); };
    auto psd_at = [&](int kk) -> float { return Ef[plane_of(kk & 31) + (kk >> 5)]; };
    auto group_sync = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + h) : "memory"); };
    const float2 sgn = h ? make_float2(-1.f, -1.f) : make_float2(1.f, 1.f);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) {
            mbar_init(&FULL[s], 1);
            DONE[s] = 0;
        }
        fence_mbar_init();
    }
    if (tid < 240) TW[tid] = __ldg(&tw256[((tid & 15) * ((tid >> 4) + 1)) & 255]);
    float2 tws[16];
#pragma unroll
    for (int p = 0; p < 16; p++) tws[p] = __ldg(&tw_step[(2 * OutIdx<16>::of(p) + h) * 256 + c]);
    __syncthreads();

    // producer iterator: BOTH group leaders walk the same item sequence (one step per consumed stage); whichever group
    // finishes a stage second issues the copy its iterator points at
    int pseg = blockIdx.x, pblk = 0, pn = 0;
    if (pseg < a.n_segs) pn = __ldg(&a.segs[pseg].n_blocks);
    auto advance = [&](bool issue, int s) {
        if (pseg >= a.n_segs) return;
        if (issue) {
            const unsigned char *src = reinterpret_cast<const unsigned char *>(a.segs[pseg].iq) + (size_t)pblk * Gm::STAGE_BYTES;
            mbar_expect_tx(&FULL[s], Gm::STAGE_BYTES);
            tma_load_1d(m8_smem + (size_t)s * Gm::STAGE_BYTES, src, Gm::STAGE_BYTES, &FULL[s]);
        }
        if (++pblk == pn) {
            pseg += gridDim.x;
            pblk = 0;
            if (pseg < a.n_segs) pn = a.segs[pseg].n_blocks;
        }
    };
    if (tg == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) advance(h == 0, s);
    }
    uint32_t item = 0;

    for (int seg = blockIdx.x; seg < a.n_segs; seg += gridDim.x) {
        const Segment sg = a.segs[seg];
        const WorkParams wp = a.works[sg.work];
        const int L = wp.n_listeners;
        const int *lbins = a.listener_bins + wp.listener_off;
        const int e = wp.edge_width;
        const int ws = nf_window_size(N, e), n_win = nf_window_count(N, e);
        // noise floor: warps 0-4 of the group; lane -> (window 2 wg + lane/16, row 2 (lane % 16) + h); the window's bins in
        // that row are the positions [first(row), last(row)) (bin kk = k1 + 32 p).  The sixteen rows of a window are read
        // from the position nf_p of the group's LAST row (the smallest first()), which keeps their banks plane_of's; the
        // row's own range is [nf_lo, nf_hi) relative to it, nf_lo in {0, 1} (k1_large.cuh: nf_row_share)
        int nf_p = 0, nf_lo = 0, nf_hi = 0, nf_rot = 0;
        const int nf_w = 2 * wg + (lane >> 4), nf_row = 2 * (lane & 15) + h;
        if (wg < 5) {
            auto first = [&](int w, int k1) { const int lo = e + w * ws; return lo < k1 ? 0 : (lo - k1 + 31) >> 5; };
            const int hi = e + (nf_w + 1) * ws;
            nf_p = first(nf_w, 30 + h);
            nf_lo = first(nf_w, nf_row) - nf_p;
            const int p1 = hi <= nf_row ? 0 : min(256, (hi - nf_row + 31) >> 5);
            nf_hi = max(p1 - nf_p, 0);
            // the window of the upper half-warp walks rotated when its start has the parity of the lower one's
            nf_rot = (lane >> 4) & ~(first(2 * wg + 1, 30 + h) - first(2 * wg, 30 + h)) & 1;
        }
        float cum[16];
#pragma unroll
        for (int p = 0; p < 16; p++) {
            const int kk = k1row + 32 * ((hl + 16 * OutIdx<16>::of(p) + 128) & 255);
            cum[p] = sg.state_in >= 0 ? a.cum_state[(size_t)sg.state_in * N + kk] : 0.f;
        }

        for (int blk = 0; blk < sg.n_blocks; blk++, item++) {
            const int s = item % NSTAGE;
            const uint32_t parity = (item / NSTAGE) & 1u;
            const float2 *IN = reinterpret_cast<const float2 *>(m8_smem + (size_t)s * Gm::STAGE_BYTES);
            const int ob = sg.block_out + blk;
            mbar_wait(&FULL[s], parity);

            // ---------------- pass A: column c, outputs k1 = 2j + h ----------------
            float2 v[16];
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int m = (q & 3) * 4 + (q >> 2);
                float2 x0 = IN[m * 256 + c], x1 = IN[(m + 16) * 256 + c];
                if (HAS_WINDOW) {
                    const float w0 = __ldg(&a.window[m * 256 + c]), w1 = __ldg(&a.window[(m + 16) * 256 + c]);
                    x0 = __fmul2_rn(x0, make_float2(w0, w0));
                    x1 = __fmul2_rn(x1, make_float2(w1, w1));
                }
                v[m] = __ffma2_rn(x1, sgn, x0);
            }
            if (h) {
                v[1] = mul_w64<2>(v[1]);
                v[2] = mul_w64<4>(v[2]);
                v[3] = mul_w64<6>(v[3]);
                v[4] = mul_w64<8>(v[4]);
                v[5] = mul_w64<10>(v[5]);
                v[6] = mul_w64<12>(v[6]);
                v[7] = mul_w64<14>(v[7]);
                v[8] = mul_w64<16>(v[8]);
                v[9] = mul_w64<18>(v[9]);
                v[10] = mul_w64<20>(v[10]);
                v[11] = mul_w64<22>(v[11]);
                v[12] = mul_w64<24>(v[12]);
                v[13] = mul_w64<26>(v[13]);
                v[14] = mul_w64<28>(v[14]);
                v[15] = mul_w64<30>(v[15]);
            }
            dft16(v);
#pragma unroll
            for (int p = 0; p < 16; p++) v[p] = cmul(v[p], tws[p]);
            // G0: the group has consumed stage s and finished the previous block's reads of its half of E
            group_sync();
            if (tg == 0) {
                __threadfence_block();
                const bool second = atomicAdd(&DONE[s], 1) == 1;  // the other group is done with the stage too
                if (second) {
                    atomicExch(&DONE[s], 0);
                    __threadfence_block();
                    fence_proxy_async();
                }
                advance(second, s);
            }
#pragma unroll
            for (int p = 0; p < 16; p++) E[(2 * OutIdx<16>::of(p) + h) * HW_PITCH + c] = v[p];
            group_sync();  // G1: the group's sixteen rows of E are complete

            // ---------------- pass B: half-warp = row k1row ----------------
            {
                HwTwiddle t;
#pragma unroll
                for (int k = 1; k < 16; k++) t.w[k - 1] = TW[(k - 1) * 16 + hl];
                float2 *col = E + k1row * HW_PITCH;
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int n1 = (q & 3) * 4 + (q >> 2);
                    v[n1] = col[16 * n1 + hl];
                }
                fft256_halfwarp_regs(v, col, t, hl);
                __syncwarp();
                float *prow = Ef + plane_of(k1row);  // inside the row's own (dead) column
#pragma unroll
                for (int p = 0; p < 16; p++) {
                    const int k2s = hl + ((16 * OutIdx<16>::of(p) + 128) & 255);  // fftshift (dsp/fft.go:54-57)
                    const float psd = fmaf(v[p].x, v[p].x, v[p].y * v[p].y);      // dsp/fft.go:71-73
                    const float db = to_db(psd);                                   // rx/receiver.go:376-378
                    cum[p] = __fadd_rn(cum[p], db);                                // rx/receiver.go:404-406
                    prow[k2s] = psd;
                    if (DEBUG_STORE) {
                        const int kk = k1row + 32 * k2s;
                        a.dbg_spectrum[(size_t)ob * N + kk] = db;
                        a.dbg_psd[(size_t)ob * N + kk] = psd;
                    }
                }
            }
            group_sync();  // G2: |X|^2 of the group's rows complete

            // ---------------- dsp.FindNoiseFloor (dsp/fft.go:215-252): this group's share of the window sums ----------------
            if (wg < 5) {
                float s1, s2;
                nf_row_share_tested<28>(Ef + plane_of(nf_row) + nf_p, nf_lo, nf_hi, nf_rot, s1, s2);  // <= 26 positions per row, + 1, even
                double d1 = (double)s1, d2 = (double)s2;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
                    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
                }
                if ((lane & 15) == 0) nf.nf_part[((size_t)ob * 2 + h) * 10 + nf_w] = make_double2(d1, d2);
            } else if (wg == 5) {
                if (lane < n_win) {  // x_to = psd[e + (w+1)*ws] (dsp/fft.go:238-243): the group that owns the bin's row writes it
                    const int kk = e + (lane + 1) * ws;
                    if ((kk & 1) == h) nf.xto[(size_t)ob * 10 + lane] = psd_at(kk);
                }
                if (lane == 31 && h == 0) nf.nf_edge[ob] = e;
            } else {
                // listener taps (rx/receiver.go:393) on the bins of this group's rows
                for (int l = tg - 192; l < L; l += 64) {
                    const int kk = __ldg(&lbins[l]);
                    if ((kk & 1) == h) a.taps[(size_t)ob * a.tap_stride + l] = to_db(psd_at(kk));
                }
            }
        }
        float *dst = (sg.flush_idx >= 0) ? a.flush_cum + (size_t)sg.flush_idx * N : a.cum_state + (size_t)sg.state_out * N;
#pragma unroll
        for (int p = 0; p < 16; p++) dst[k1row + 32 * ((hl + 16 * OutIdx<16>::of(p) + 128) & 255)] = cum[p];
    }
}

}  // namespace sdr
