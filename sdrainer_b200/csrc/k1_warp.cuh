// k1_warp.cuh -- K1 for N = 512 (the TCI shape: 48 kS/s, tci/tci.go:157-158; and KiwiSDR's 12 kS/s, kiwi/kiwi.go:13):
// one WARP owns a block, no CTA-wide barrier anywhere.
//
// Same reference arithmetic as k1_spectral.cuh (dsp/fft.go:23-85 FFT / fftshift / |X|^2 / dB+120,
// dsp/fft.go:215-252 FindNoiseFloor, rx/receiver.go:393 listener taps, rx/receiver.go:404-407 cumulation).
//
// The three-pass kernel spends a whole shared-memory exchange on N = 512's last radix-2 layer (16 x 16 x 2) and is
// bound by the shared-memory pipe at ~50 % of the HBM roofline.  Here 512 = 2 x 256:
//   pass A  lane = 8 columns c: Z0[c] = x[c] + x[256+c], Z1[c] = (x[c] - x[256+c]) * W512^c, straight from global
//           memory (every warp load is 256 contiguous bytes; no staging), stored as two rows of 256;
//   pass B  half-warp h = row h: the 256-point half-warp transform of k1_large.cuh (radix-16, per-lane W256 twiddles,
//           16x17 transpose through the row's own storage, radix-16): lane hl ends with X[h + 2*k2], k2 = hl + 16 q;
//   epilogue in registers (|X|^2, dB, 16 cumulation bins per lane); |X|^2 goes to a 512-float plane in natural
//           (fftshifted) bin order for the noise windows, x_to and the taps.
// Noise floor: lane shares -- 17 consecutive bins per lane (odd stride: bank-conflict free for
// every edge width), split at the one window boundary a share can contain, window sums gathered by lanes 0..9 and the
// sequential selection batched NFB blocks at a time -- all inside the warp (__syncwarp only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "k1_large.cuh"

#ifndef SDR_K1W_MINB
#define SDR_K1W_MINB 4
#endif

namespace sdr {

struct K1WarpGeom {
    static constexpr int N = 512;
    static constexpr int WARPS = 4;                       // independent warps per CTA
    static constexpr int ROW_BYTES = HW_PITCH * 8;        // one 256-point row incl. transpose slack
    static constexpr int E_BYTES = 2 * ROW_BYTES;         // 4368; PSD plane (+ over-read) and PART alias it
    static constexpr int NF_SHARE = 17;                   // bins per lane: 32 * 17 = 544 >= 512
    static constexpr int PART_OFF = 3072;                 // float4[32] after the over-read PSD plane (<= 716 floats)
    static constexpr int NFB = 8;
    static constexpr int WS_BYTES = NFB * 10 * 8;         // [blk][w] (s1, s2)
    static constexpr int XTO_BYTES = NFB * 10 * 4;
    static constexpr int WARP_BYTES = (E_BYTES + WS_BYTES + XTO_BYTES + 15) / 16 * 16;
    static constexpr int SMEM_BYTES = WARPS * WARP_BYTES;
    static constexpr int MIN_WS = NF_SHARE;               // a share may straddle ONE window boundary
    static_assert(PART_OFF + 32 * 16 <= E_BYTES, "PART inside the row storage");
};

// selection of dsp.FindNoiseFloor (dsp/fft.go:217-251) for one block from float32 window sums
__device__ __forceinline__ void nf_select_f2(const float2 *wsum, const float *xto, int ws, int n_win, float *out_min, double *out_var) {
    const double inv_ws = 1.0 / (double)ws;
    double min_value = 0.0, P1 = 0.0, P2 = 0.0, bP1 = 0.0, bP2 = 0.0;
    int best = 0;
    for (int w = 0; w < n_win; w++) {
        const float2 p = wsum[w];
        const double a1 = (double)p.x;
        P1 += a1;
        P2 += (double)p.y;
        const double mean = a1 * inv_ws;
        if (w == 0 || mean < min_value) {  // `mean < minValue || first`
            min_value = mean;
            best = w;
            bP1 = P1;
            bP2 = P2;
        }
    }
    const double x = (double)xto[best];
    bP1 += x;
    bP2 = fma(x, x, bP2);
    const double n = (double)((best + 1) * ws + 1);  // bins e .. e+(best+1)*ws inclusive (the reference's `from` quirk)
    *out_min = (float)min_value;
    *out_var = (bP2 - min_value * (2.0 * bP1 - n * min_value)) * inv_ws;
}

template <bool DEBUG_STORE, bool HAS_WINDOW, bool IN_I16>
__global__ void __launch_bounds__(32 * K1WarpGeom::WARPS, SDR_K1W_MINB) k1_warp_kernel(const K1Args a, const float2 *__restrict__ tw512,
                                                                                      const float2 *__restrict__ tw256) {
    using Gm = K1WarpGeom;
    constexpr int N = Gm::N, SH = Gm::NF_SHARE, NFB = Gm::NFB;
    extern __shared__ __align__(16) unsigned char warp_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int h = lane >> 4, hl = lane & 15;
    unsigned char *base = warp_smem + (size_t)wid * Gm::WARP_BYTES;
    float2 *E = reinterpret_cast<float2 *>(base);                      // [2][HW_PITCH]
    float *PSD = reinterpret_cast<float *>(base);                      // [512] natural fftshifted order (aliases E)
    float4 *PART = reinterpret_cast<float4 *>(base + Gm::PART_OFF);
    float2 *WSUM = reinterpret_cast<float2 *>(base + Gm::E_BYTES);     // [NFB][10]
    float *XTO = reinterpret_cast<float *>(base + Gm::E_BYTES + Gm::WS_BYTES);

    // per-lane constants: W512^c for this lane's 8 columns, W256^(hl k1) for the row transform.  Kept in registers
    // (128 registers, 4 CTAs = 16 warps per SM).  SDR_K1W_RELOAD re-reads them from the L1-resident tables in every block
    // instead (96 registers, 5 CTAs): measured slower -- the 23 extra loads per lane and block cost more of the L1/shared
    // pipe (82 % busy) than the fifth CTA brings.
#ifndef SDR_K1W_RELOAD
    float2 twc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) twc[i] = __ldg(&tw512[lane + 32 * i]);
    HwTwiddle t;
    hw_twiddle_load(t, tw256, hl);
#endif

    const int gid = blockIdx.x * Gm::WARPS + wid, gstride = gridDim.x * Gm::WARPS;
    for (int seg = gid; seg < a.n_segs; seg += gstride) {
        const Segment sg = a.segs[seg];
        const WorkParams wp = a.works[sg.work];
        const int L = wp.n_listeners;
        const int *lbins = a.listener_bins + wp.listener_off;
        const int lb0 = lane < L ? __ldg(&lbins[lane]) : -1;

        // ---- noise-floor geometry (dsp/fft.go:216,224) ----
        const int e = wp.edge_width;
        const int ws = nf_window_size(N, e);
        const int n_win = nf_window_count(N, e);
        // phase 1: lane q sums bins [e + 17q, e + 17q + 17); the share starts in window wl and may cross into wl + 1 at
        // relative index bnd (windows are at least as long as a share: host-checked ws >= MIN_WS)
        const int nf_start = e + SH * lane;
        const int wl = (SH * lane) / ws;
        const int bnd = e + (wl + 1) * ws - nf_start;
        // phase 2: lane w < 10 adds the partial sums that belong to window w
        int qa = 0, qb = -1;
        bool first_left = true;
        if (lane < 10) {
            qa = (lane * ws) / SH;
            qb = ((lane + 1) * ws - 1) / SH;
            first_left = (lane * ws == SH * qa);
        }
        const int xbin = e + (lane + 1) * ws;  // x_to = psd[first bin of the next window] (dsp/fft.go:238-243)

        // cumulation registers: cum2[q] = positions p = 2q, 2q+1 of the row transform;
        // position p is bin kk = h + 2*((hl + 16*OutIdx<16>(p) + 128) & 255)   (fftshift, dsp/fft.go:54-57)
        float2 cum2[8];
        if (sg.state_in >= 0) {
            const float *cs = a.cum_state + (size_t)sg.state_in * N + h;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                cum2[q].x = cs[2 * ((hl + 16 * OutIdx<16>::of(2 * q) + 128) & 255)];
                cum2[q].y = cs[2 * ((hl + 16 * OutIdx<16>::of(2 * q + 1) + 128) & 255)];
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; q++) cum2[q] = make_float2(0.f, 0.f);
        }
        int nf_fill = 0, nf_first = sg.block_out;
        auto nf_select = [&]() {
            __syncwarp();
            if (lane < nf_fill) nf_select_f2(WSUM + lane * 10, XTO + lane * 10, ws, n_win, &a.psd_floor[nf_first + lane],
                                             &a.variance[nf_first + lane]);
            __syncwarp();
            nf_first += nf_fill;
            nf_fill = 0;
        };

        for (int blk = 0; blk < sg.n_blocks; blk++) {
            const int ob = sg.block_out + blk;
            // ---------------- pass A: radix-2 layer over the two half blocks, 8 columns per lane ----------------
            {
                float2 x0[8], x1[8];
                if (IN_I16) {
                    const uint32_t *src = reinterpret_cast<const uint32_t *>(sg.iq) + (size_t)blk * N + lane;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        x0[i] = kiwi_decode_sample(__ldg(&src[32 * i]));
                        x1[i] = kiwi_decode_sample(__ldg(&src[256 + 32 * i]));
                    }
                } else {
                    const float2 *src = reinterpret_cast<const float2 *>(sg.iq) + (size_t)blk * N + lane;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        x0[i] = __ldg(&src[32 * i]);
                        x1[i] = __ldg(&src[256 + 32 * i]);
                    }
                }
                if (HAS_WINDOW) {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float w0 = __ldg(&a.window[lane + 32 * i]), w1 = __ldg(&a.window[256 + lane + 32 * i]);
                        x0[i] = __fmul2_rn(x0[i], make_float2(w0, w0));
                        x1[i] = __fmul2_rn(x1[i], make_float2(w1, w1));
                    }
                }
#ifdef SDR_K1W_RELOAD
                const float2 *twp = tw512;
                asm volatile("" : "+l"(twp));  // laundered pointer: keeps the loads inside the block loop
                float2 twc[8];
#pragma unroll
                for (int i = 0; i < 8; i++) twc[i] = __ldg(&twp[lane + 32 * i]);
#endif
                __syncwarp();  // the previous block's reads of the PSD plane / PART (aliases of E) are done
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int c = lane + 32 * i;
                    E[c] = cadd(x0[i], x1[i]);
                    const float2 d = csub(x0[i], x1[i]);
                    E[HW_PITCH + c] = cmul(d, twc[i]);
                }
            }
            __syncwarp();
            // ---------------- pass B: half-warp h = row h ----------------
            float2 u[16];
            float2 *row = E + h * HW_PITCH;
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int n1 = (q & 3) * 4 + (q >> 2);
                u[n1] = row[16 * n1 + hl];
            }
#ifdef SDR_K1W_RELOAD
            HwTwiddle t;
            {
                const float2 *twq = tw256;
                asm volatile("" : "+l"(twq));
                hw_twiddle_load(t, twq, hl);
            }
#endif
            fft256_halfwarp_regs(u, row, t, hl);
            __syncwarp();  // both rows' transposes are done: the PSD plane may overwrite them
            // ---------------- |X|^2 (dsp/fft.go:71-73), dB + 120 (rx/receiver.go:376-378), cumulation ----------------
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float2 psd = make_float2(fmaf(u[2 * q].x, u[2 * q].x, u[2 * q].y * u[2 * q].y),
                                               fmaf(u[2 * q + 1].x, u[2 * q + 1].x, u[2 * q + 1].y * u[2 * q + 1].y));
                const int kk0 = h + 2 * ((hl + 16 * OutIdx<16>::of(2 * q) + 128) & 255);
                const int kk1 = h + 2 * ((hl + 16 * OutIdx<16>::of(2 * q + 1) + 128) & 255);
                PSD[kk0] = psd.x;
                PSD[kk1] = psd.y;
                const float2 db = psd_to_db2<N>(psd);
                cum2[q] = __fadd2_rn(cum2[q], db);  // rx/receiver.go:404-406
                if (DEBUG_STORE) {
                    a.dbg_spectrum[(size_t)ob * N + kk0] = db.x;
                    a.dbg_spectrum[(size_t)ob * N + kk1] = db.y;
                    a.dbg_psd[(size_t)ob * N + kk0] = psd.x;
                    a.dbg_psd[(size_t)ob * N + kk1] = psd.y;
                }
            }
            __syncwarp();  // PSD plane complete
            // ---------------- noise floor, phase 1: this lane's share, split at the window boundary ----------------
            {
                const float *pp = PSD + nf_start;
                float l1 = 0.f, l2 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
                for (int i = 0; i < SH; i++) {
                    const float x = pp[i];
                    if (i < bnd) {
                        l1 += x;
                        l2 = fmaf(x, x, l2);
                    } else {
                        r1 += x;
                        r2 = fmaf(x, x, r2);
                    }
                }
                const float xt = (lane < n_win) ? PSD[xbin] : 0.f;
                // listener taps (rx/receiver.go:393)
                if (lb0 >= 0) a.taps[(size_t)ob * a.tap_stride + lane] = psd_to_db<N>(PSD[lb0]);
                for (int l = lane + 32; l < L; l += 32) a.taps[(size_t)ob * a.tap_stride + l] = psd_to_db<N>(PSD[__ldg(&lbins[l])]);
                __syncwarp();  // every lane has read the plane: PART (inside the row storage, beyond the plane) is free
                PART[lane] = make_float4(l1, l2, r1, r2);
                if (lane < n_win) XTO[nf_fill * 10 + lane] = xt;
            }
            __syncwarp();
            // ---------------- noise floor, phase 2: lane w gathers window w's partial sums ----------------
            if (lane < 10) {
                const float4 f4 = PART[qa];
                float s1 = first_left ? f4.x : f4.z, s2 = first_left ? f4.y : f4.w;
                for (int q = qa + 1; q <= qb; q++) {
                    const float4 g = PART[q];
                    s1 += g.x;
                    s2 += g.y;
                }
                WSUM[nf_fill * 10 + lane] = make_float2(s1, s2);
            }
            nf_fill++;
            if (nf_fill == NFB) nf_select();
        }
        if (nf_fill > 0) nf_select();

        // ---- end of segment: flush or save the cumulation ----
        float *dst = ((sg.flush_idx >= 0) ? a.flush_cum + (size_t)sg.flush_idx * N : a.cum_state + (size_t)sg.state_out * N) + h;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            dst[2 * ((hl + 16 * OutIdx<16>::of(2 * q) + 128) & 255)] = cum2[q].x;
            dst[2 * ((hl + 16 * OutIdx<16>::of(2 * q + 1) + 128) & 255)] = cum2[q].y;
        }
        __syncwarp();
    }
}

}  // namespace sdr
