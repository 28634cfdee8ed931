// k1_pair.cuh -- K1 for N = 2048 (192 kS/s streams, BASELINE configs[1] and [3]): the fused spectral front end
// of k1_spectral.cuh re-cut so that ONE shared-memory exchange per block is enough.
//
// Same reference arithmetic as k1_spectral.cuh (dsp/fft.go:23-85 FFT / fftshift / |X|^2 / dB+120,
// dsp/fft.go:215-252 FindNoiseFloor, rx/receiver.go:393 listener taps, rx/receiver.go:404-407 cumulation).
//
// Why a second cut: the three-pass kernel moves every block through shared memory three and a half times
// (TMA stage, two register<->smem exchanges, PSD staging) and is bound by the shared-memory data pipe
// (profiles/r1_ncu_summary.md: 1030 wavefronts per 16 KB block, LSU pipe 82 %).  Here a block is owned by a
// PAIR of warps and the 2048-point DFT is split by output parity (decimation in frequency):
//
//   n = 32*n1 + l   (l = lane = column, n1 = 0..63),   k = k1 + 64*k2
//   warp h (0/1) computes the rows k1 = 2j + h:
//     c_h[m]   = (x[32m + l] + (-1)^h x[32(m+32) + l]) * W64^(m h)          m = 0..31   (radix-2 layer, in the load)
//     Y[2j+h]  = DFT32_j(c_h)                                                            (registers)
//     Z        = Y * W2048^(l (2j+h))                                                    (registers, per-lane table)
//     -- exchange inside the warp: E_h[j][l], 8.25 KB, __syncwarp only --
//     X[2j+h + 64 k2] = DFT32_k2( Z[2j+h][0..31] )       thread lane = j                 (registers)
//
// so each warp runs a self-contained 1024-point transform (32 x 32) and the two warps never exchange spectrum
// data.  Both read the whole TMA stage (that is the price: the stage is read twice), which still leaves
// ~780 shared-memory wavefronts per block instead of ~1030, no CTA-wide barrier in the block loop, and 32
// independent butterflies per thread for latency hiding.  The warps meet only (i) on the stage's
// "consumed" counter -- the second warp to finish its loads re-arms the TMA copy of the next block --, (ii) every
// NFB blocks to run dsp.FindNoiseFloor's window selection on the batched window sums.
//
// PSD staging is de-interleaved by parity: warp h owns the bins kk = 2i + h and keeps them as a plane
// PSD_h[i] (i = 0..1023, fftshifted), so its stores and the noise-floor reads are bank-conflict free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "k1_spectral.cuh"

#ifndef SDR_K1P_MINB
#define SDR_K1P_MINB 5
#endif

namespace sdr {

// ---- W64 constants (compile-time) ---------------------------------------------------------------------------
__host__ __device__ constexpr double w64_quarter(int k) {  // cos(2 pi k / 64), k = 0..16
    constexpr double q[17] = {1.0,
                              0.99518472667219693,
                              0.98078528040323043,
                              0.95694033573220882,
                              0.92387953251128674,
                              0.88192126434835505,
                              0.83146961230254524,
                              0.77301045336273699,
                              0.70710678118654757,
                              0.63439328416364549,
                              0.55557023301960229,
                              0.47139673682599781,
                              0.38268343236508984,
                              0.29028467725446233,
                              0.19509032201612833,
                              0.09801714032956077,
                              0.0};
    return q[k];
}
__host__ __device__ constexpr double w64_cos(int k) {
    k &= 63;
    return k <= 16 ? w64_quarter(k) : k <= 32 ? -w64_quarter(32 - k) : k <= 48 ? -w64_quarter(k - 32) : w64_quarter(64 - k);
}
__host__ __device__ constexpr double w64_sin(int k) { return w64_cos(k - 16); }  // sin(t) = cos(t - pi/2)

// a * W64^E, E a compile-time exponent; the multiples of 8 cost at most two packed instructions
template <int E>
__device__ __forceinline__ float2 mul_w64(float2 a) {
    constexpr int e = E & 63;
    if constexpr (e == 0) return a;
    else if constexpr (e == 16) return mul_mi(a);
    else if constexpr (e == 32) return make_float2(-a.x, -a.y);
    else if constexpr (e == 48) return make_float2(-a.y, a.x);
    else if constexpr (e == 8) return mul_w8_1(a);
    else if constexpr (e == 24) return mul_w8_3(a);
    else {
        constexpr float c = (float)w64_cos(e), s = (float)(-w64_sin(e));
        return cmul(a, make_float2(c, s));
    }
}

template <>
struct OutIdx<32> {  // position p = 8*ka + q holds X[ka + 4*OutIdx<8>(q)]
    __host__ __device__ static constexpr int of(int p) { return (p >> 3) + 4 * OutIdx<8>::of(p & 7); }
};

template <int KA, int B>
__device__ __forceinline__ void dft32_twiddle_one(float2 (&v)[32]) {
    v[8 * KA + B] = mul_w64<2 * B * KA>(v[8 * KA + B]);  // W32^(b ka) = W64^(2 b ka)
}
template <int KA>
__device__ __forceinline__ void dft32_twiddle_row(float2 (&v)[32]) {
    dft32_twiddle_one<KA, 1>(v);
    dft32_twiddle_one<KA, 2>(v);
    dft32_twiddle_one<KA, 3>(v);
    dft32_twiddle_one<KA, 4>(v);
    dft32_twiddle_one<KA, 5>(v);
    dft32_twiddle_one<KA, 6>(v);
    dft32_twiddle_one<KA, 7>(v);
}

// 32-point forward DFT in registers: n = 8a + b; DFT4 over a, twiddle W32^(b ka), DFT8 over b.
// Natural order in; v[p] = X[OutIdx<32>::of(p)] out.  64 + 40 + 112 = 216 packed instructions.
__device__ __forceinline__ void dft32(float2 (&v)[32]) {
#pragma unroll
    for (int b = 0; b < 8; b++) dft4(v[b], v[b + 8], v[b + 16], v[b + 24]);
    dft32_twiddle_row<1>(v);
    dft32_twiddle_row<2>(v);
    dft32_twiddle_row<3>(v);
#pragma unroll
    for (int ka = 0; ka < 4; ka++) dft8(*reinterpret_cast<float2(*)[8]>(&v[8 * ka]));
}

struct K1PairGeom {
    static constexpr int N = 2048;
    static constexpr int ES = 33;                    // E row stride in complex values (odd: conflict-free transpose)
    static constexpr int E_BYTES = 32 * ES * 8;      // 8448 per warp; also holds the warp's PSD plane + NF partials
    static constexpr int PART_OFF = 4608;            // NF partial sums inside E_h, after the (over-read) PSD plane
    static constexpr int NFB = 8;                    // blocks per noise-floor selection batch
    static constexpr int NF_SHARE = 33;              // plane elements per lane (odd stride: conflict-free)
    static constexpr int WS_BYTES = 2 * NFB * 2 * 10 * 8;  // [buf][blk][h][w] (s1, s2)
    static constexpr int XTO_BYTES = 2 * NFB * 10 * 4;     // [buf][blk][w]
    static constexpr int TAP_BYTES = 2 * 64 * 2;           // [h][listener < 64] plane index of the tap, -1: other warp's
    __host__ __device__ static constexpr int stage_bytes(bool i16) { return i16 ? 4 * N : 8 * N; }
    __host__ __device__ static constexpr int smem_bytes(int nstage) {
        return nstage * 8 * N + 2 * E_BYTES + WS_BYTES + XTO_BYTES + TAP_BYTES + 64;
    }
    // smallest noise window the lane shares can split (a share may straddle ONE window boundary)
    static constexpr int MIN_WS = 2 * NF_SHARE + 1;
};

// selection of dsp.FindNoiseFloor (dsp/fft.go:217-251) for one block from the two warps' window sums
__device__ __forceinline__ void nf_select_pair(const float2 *w0, const float2 *w1, const float *xto, int ws, int n_win,
                                               float *out_min, double *out_var) {
    const double inv_ws = 1.0 / (double)ws;
    double min_value = 0.0, P1 = 0.0, P2 = 0.0, bP1 = 0.0, bP2 = 0.0;
    int best = 0;
    for (int w = 0; w < n_win; w++) {
        const float2 p = w0[w], q = w1[w];
        const double a1 = (double)p.x + (double)q.x;
        P1 += a1;
        P2 += (double)p.y + (double)q.y;
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

template <int M>
struct PairLoad {
    // radix-2 (decimation in frequency) layer fused into the stage read: element m of c_h, h = H
    template <int H, bool IN_I16, bool HAS_WINDOW>
    static __device__ __forceinline__ void run(float2 (&v)[32], const void *in, const float *window, int lane) {
        float2 x0, x1;
        if (IN_I16) {
            x0 = kiwi_decode_sample(reinterpret_cast<const uint32_t *>(in)[32 * M + lane]);
            x1 = kiwi_decode_sample(reinterpret_cast<const uint32_t *>(in)[32 * (M + 32) + lane]);
        } else {
            x0 = reinterpret_cast<const float2 *>(in)[32 * M + lane];
            x1 = reinterpret_cast<const float2 *>(in)[32 * (M + 32) + lane];
        }
        if (HAS_WINDOW) {
            const float w0 = __ldg(&window[32 * M + lane]), w1 = __ldg(&window[32 * (M + 32) + lane]);
            x0 = __fmul2_rn(x0, make_float2(w0, w0));
            x1 = __fmul2_rn(x1, make_float2(w1, w1));
        }
        if (H == 0) v[M] = cadd(x0, x1);
        else v[M] = mul_w64<M>(csub(x0, x1));
    }
};

// all 32 elements, issued in the order dft32's first layer consumes them (b, b+8, b+16, b+24)
template <int H, bool IN_I16, bool HAS_WINDOW>
__device__ __forceinline__ void pair_load_all(float2 (&v)[32], const void *in, const float *window, int lane) {
#define SDR_PL(m) PairLoad<m>::template run<H, IN_I16, HAS_WINDOW>(v, in, window, lane);
    SDR_PL(0) SDR_PL(8) SDR_PL(16) SDR_PL(24) SDR_PL(1) SDR_PL(9) SDR_PL(17) SDR_PL(25)
    SDR_PL(2) SDR_PL(10) SDR_PL(18) SDR_PL(26) SDR_PL(3) SDR_PL(11) SDR_PL(19) SDR_PL(27)
    SDR_PL(4) SDR_PL(12) SDR_PL(20) SDR_PL(28) SDR_PL(5) SDR_PL(13) SDR_PL(21) SDR_PL(29)
    SDR_PL(6) SDR_PL(14) SDR_PL(22) SDR_PL(30) SDR_PL(7) SDR_PL(15) SDR_PL(23) SDR_PL(31)
#undef SDR_PL
}

#ifdef SDR_K1P_MAXREG
#define SDR_K1P_BOUNDS __maxnreg__(SDR_K1P_MAXREG)
#else
#define SDR_K1P_BOUNDS __launch_bounds__(64, SDR_K1P_MINB)
#endif

template <bool DEBUG_STORE, bool HAS_WINDOW, bool IN_I16, int NSTAGE>
__global__ void SDR_K1P_BOUNDS k1_pair_kernel(const K1Args a) {
    using Gm = K1PairGeom;
    constexpr int N = Gm::N, ES = Gm::ES, NFB = Gm::NFB, SH = Gm::NF_SHARE;
    constexpr uint32_t BLOCK_BYTES = IN_I16 ? 4 * N : 8 * N;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int h = threadIdx.x >> 5;  // warp of the pair = output parity
    const int lane = threadIdx.x & 31;
    unsigned char *stage_base = smem_raw;
    float2 *E = reinterpret_cast<float2 *>(smem_raw + NSTAGE * 8 * N + h * Gm::E_BYTES);
    float *PSD = reinterpret_cast<float *>(E);  // plane of this warp's bins, aliases E (dead after the pass-2 loads)
    float4 *PART = reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(E) + Gm::PART_OFF);
    float2 *WSUM = reinterpret_cast<float2 *>(smem_raw + NSTAGE * 8 * N + 2 * Gm::E_BYTES);
    float *XTO = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(WSUM) + Gm::WS_BYTES);
    short *TAPI = reinterpret_cast<short *>(reinterpret_cast<unsigned char *>(XTO) + Gm::XTO_BYTES) + h * 64;
    uint64_t *FULL = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(XTO) + Gm::XTO_BYTES + Gm::TAP_BYTES);
    unsigned int *CONS = reinterpret_cast<unsigned int *>(FULL + NSTAGE);

    // ---- one-time setup ----
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) {
            mbar_init(&FULL[s], 1);
            CONS[s] = 0;
        }
        fence_mbar_init();
    }
    // Z = Y * W2048^(l (2j+h)), stored in dft32 output order: tw[p] belongs to j = OutIdx<32>::of(p)
    float2 tw[32];
#pragma unroll
    for (int p = 0; p < 32; p++) tw[p] = __ldg(&a.twp[(h * 32 + p) * 32 + lane]);
    __syncthreads();

    // ---- producer iterator: the item NSTAGE ahead of the one being consumed (kept by every thread, uniform) ----
    const int stride = gridDim.x;
    int pseg = blockIdx.x, pblk = 0, pn = 0;
    if (pseg < a.n_segs) pn = __ldg(&a.segs[pseg].n_blocks);
    auto advance = [&]() {
        pblk++;
        if (pblk >= pn) {
            pseg += stride;
            pblk = 0;
            pn = (pseg < a.n_segs) ? __ldg(&a.segs[pseg].n_blocks) : 0x7fffffff;
        }
    };
    auto issue = [&](int s) {  // one thread
        if (pseg >= a.n_segs) return;
        const unsigned char *src = reinterpret_cast<const unsigned char *>(a.segs[pseg].iq) + (size_t)pblk * BLOCK_BYTES;
        mbar_expect_tx(&FULL[s], BLOCK_BYTES);
        tma_load_1d(stage_base + (size_t)s * 8 * N, src, BLOCK_BYTES, &FULL[s]);
    };
#pragma unroll
    for (int s = 0; s < NSTAGE; s++) {
        if (threadIdx.x == 0) issue(s);
        advance();
    }

    uint32_t item = 0;
    int nf_buf = 0;

    for (int seg = blockIdx.x; seg < a.n_segs; seg += stride) {
        const Segment sg = a.segs[seg];
        const WorkParams wp = a.works[sg.work];
        const int L = wp.n_listeners;
        const int *lbins = a.listener_bins + wp.listener_off;
        // listener bins (rx/listener.go:119-124): this warp serves the listeners on bins of its parity; the plane
        // indices of the first 64 listeners are cached in shared memory, later ones are re-read per block
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int l = lane + 32 * k;
            const int bin = (l < L) ? __ldg(&lbins[l]) : -1;
            TAPI[l] = (short)((bin >= 0 && (bin & 1) == h) ? (bin >> 1) : -1);
        }
        __syncwarp();

        // ---- noise-floor geometry (dsp/fft.go:216,224) in plane coordinates: bin kk = 2i + h ----
        const int e = wp.edge_width;
        const int ws = nf_window_size(N, e);
        const int n_win = nf_window_count(N, e);
        auto plane_lo = [&](int bin) { return (bin - h + 1) >> 1; };  // first plane index whose bin is >= `bin`
        const int B0 = plane_lo(e);
        // phase 1: lane q sums the plane elements [B0 + 33q, B0 + 33q + 33); the share starts in window wl and
        // may cross into wl + 1 at relative index bnd (windows are longer than a share: host-checked ws >= MIN_WS)
        const int nf_start = B0 + SH * lane;
        const int wl = (2 * nf_start + h - e) / ws;
        const int bnd = plane_lo(e + (wl + 1) * ws) - nf_start;
        // phase 2: lane w < 10 adds the partial sums that belong to window w
        int qa = 0, qb = -1;
        bool first_left = true;
        if (lane < 10) {
            const int bw = plane_lo(e + lane * ws), bw1 = plane_lo(e + (lane + 1) * ws);
            qa = (bw - B0) / SH;
            qb = (bw1 - 1 - B0) / SH;
            first_left = (bw == B0 + SH * qa);
        }
        // x_to = psd[first bin of the next window] (dsp/fft.go:238-243): owned by the warp of that bin's parity
        const int xbin = e + (lane + 1) * ws;
        const bool x_owner = lane < n_win && (xbin & 1) == h;

        // cumulation registers in dft32 output order: cum2[q] = positions p = 2q, 2q+1;
        // position p is bin kk = 2*lane + h + 64*((OutIdx<32>(p) + 16) & 31)   (fftshift, dsp/fft.go:54-57)
        float2 cum2[16];
        if (sg.state_in >= 0) {
            const float *cs = a.cum_state + (size_t)sg.state_in * N + 2 * lane + h;
#pragma unroll
            for (int q = 0; q < 16; q++) {
                cum2[q].x = cs[64 * ((OutIdx<32>::of(2 * q) + 16) & 31)];
                cum2[q].y = cs[64 * ((OutIdx<32>::of(2 * q + 1) + 16) & 31)];
            }
        } else {
#pragma unroll
            for (int q = 0; q < 16; q++) cum2[q] = make_float2(0.f, 0.f);
        }
        int nf_fill = 0, nf_first = sg.block_out;
        auto nf_select = [&]() {  // after a pair barrier; the warps alternate
            if (h == (nf_buf & 1) && lane < nf_fill) {
                const float2 *w0 = WSUM + ((nf_buf * NFB + lane) * 2 + 0) * 10;
                nf_select_pair(w0, w0 + 10, XTO + (nf_buf * NFB + lane) * 10, ws, n_win, &a.psd_floor[nf_first + lane],
                               &a.variance[nf_first + lane]);
            }
            nf_first += nf_fill;
            nf_fill = 0;
            nf_buf ^= 1;
        };

        for (int blk = 0; blk < sg.n_blocks; blk++, item++) {
            const int s = item % NSTAGE;
            const uint32_t parity = (item / NSTAGE) & 1u;
            const void *IN = stage_base + (size_t)s * 8 * N;
            const int ob = sg.block_out + blk;

            mbar_wait(&FULL[s], parity);

            // ---------------- radix-2 layer + pass 1 (DFT32 over m), column l = lane ----------------
            float2 v[32];
if (h == 0) pair_load_all<0, IN_I16, HAS_WINDOW>(v, IN, a.window, lane);
            else pair_load_all<1, IN_I16, HAS_WINDOW>(v, IN, a.window, lane);
            __syncwarp();
            // stage consumed by this warp; the second warp to get here re-arms it with the next block
            if (lane == 0) {
                __threadfence_block();
                const unsigned int old = atomicAdd(&CONS[s], 1u);
                if (old & 1u) {
                    __threadfence_block();
                    fence_proxy_async();
                    issue(s);
                }
            }
            advance();

            dft32(v);
#pragma unroll
            for (int p = 0; p < 32; p++) {
                if (p == 0) {
                    if (h == 1) v[0] = cmul(v[0], tw[0]);  // k1 = h: W^(l h)
                } else {
                    v[p] = cmul(v[p], tw[p]);
                }
            }
#pragma unroll
            for (int p = 0; p < 32; p++) E[OutIdx<32>::of(p) * ES + lane] = v[p];
            __syncwarp();

            // ---------------- pass 2: DFT32 over l for row j = lane (k1 = 2*lane + h) ----------------
#pragma unroll
            for (int q = 0; q < 32; q++) {
                const int n2 = (q & 3) * 8 + (q >> 2);  // issue order = consumption order of the first layer
                v[n2] = E[lane * ES + n2];
            }
            __syncwarp();  // E consumed: the PSD plane may overwrite it
            dft32(v);

            // ---------------- |X|^2 (dsp/fft.go:71-73), dB + 120 (rx/receiver.go:376-378), cumulation ----------------
#pragma unroll
            for (int q = 0; q < 16; q++) {
                // re^2 + im^2 as FMUL + FFMA (one rounding fewer than the reference's two float64 products, far
                // below the fp32-FFT error of the bin; 2 FP32-pipe cycles per bin instead of 3)
                const float2 psd = make_float2(fmaf(v[2 * q].x, v[2 * q].x, v[2 * q].y * v[2 * q].y),
                                               fmaf(v[2 * q + 1].x, v[2 * q + 1].x, v[2 * q + 1].y * v[2 * q + 1].y));
                const int c0 = (OutIdx<32>::of(2 * q) + 16) & 31, c1 = (OutIdx<32>::of(2 * q + 1) + 16) & 31;
                PSD[lane + 32 * c0] = psd.x;
                PSD[lane + 32 * c1] = psd.y;
                const float2 db = psd_to_db2<N>(psd);
                cum2[q] = __fadd2_rn(cum2[q], db);  // rx/receiver.go:404-406
                if (DEBUG_STORE) {
                    const int kk0 = 2 * lane + h + 64 * c0, kk1 = 2 * lane + h + 64 * c1;
                    a.dbg_spectrum[(size_t)ob * N + kk0] = db.x;
                    a.dbg_spectrum[(size_t)ob * N + kk1] = db.y;
                    a.dbg_psd[(size_t)ob * N + kk0] = psd.x;
                    a.dbg_psd[(size_t)ob * N + kk1] = psd.y;
                }
            }
            __syncwarp();  // PSD plane complete

            // ---------------- noise floor, phase 1: sums of x and x^2 over this lane's share, split at the
            // window boundary: .x accumulates the part in window wl, .y the part in window wl + 1 ----------------
            {
                const float *pp = PSD + nf_start;
#ifdef SDR_K1P_NF_PACKED
                float2 a1 = make_float2(0.f, 0.f), a2 = a1, b1 = a1, b2 = a1;
#pragma unroll
                for (int i = 0; i < SH; i++) {
                    const float x = pp[i];
                    const bool left = i < bnd;
                    const float2 xv = make_float2(left ? x : 0.f, left ? 0.f : x);
                    if (i & 1) {
                        b1 = __fadd2_rn(b1, xv);
                        b2 = __ffma2_rn(xv, xv, b2);
                    } else {
                        a1 = __fadd2_rn(a1, xv);
                        a2 = __ffma2_rn(xv, xv, a2);
                    }
                }
                a1 = __fadd2_rn(a1, b1);
                a2 = __fadd2_rn(a2, b2);
#else
                // predicated scalar accumulation: every element costs one FADD + one FFMA on exactly one side
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
                const float2 a1 = make_float2(l1, r1), a2 = make_float2(l2, r2);
#endif
                PART[lane] = make_float4(a1.x, a2.x, a1.y, a2.y);  // (left s1, left s2, right s1, right s2)
            }
            if (x_owner) XTO[(nf_buf * NFB + nf_fill) * 10 + lane] = PSD[xbin >> 1];
            // listener taps (rx/receiver.go:393): this warp serves the listeners on bins of its parity
            {
                const int t0 = TAPI[lane], t1 = TAPI[lane + 32];
                float *tp = a.taps + (size_t)ob * a.tap_stride + lane;
                if (t0 >= 0) tp[0] = psd_to_db<N>(PSD[t0]);
                if (t1 >= 0) tp[32] = psd_to_db<N>(PSD[t1]);
            }
            for (int l = lane + 64; l < L; l += 32) {
                const int bin = __ldg(&lbins[l]);
                if ((bin & 1) == h) a.taps[(size_t)ob * a.tap_stride + l] = psd_to_db<N>(PSD[bin >> 1]);
            }
            __syncwarp();  // PART complete

            // ---------------- noise floor, phase 2: lane w gathers window w's partial sums ----------------
            if (lane < 10) {
                const float4 f = PART[qa];
                float s1 = first_left ? f.x : f.z, s2 = first_left ? f.y : f.w;
                for (int q = qa + 1; q <= qb; q++) {
                    const float4 g = PART[q];
                    s1 += g.x;
                    s2 += g.y;
                }
                WSUM[((nf_buf * NFB + nf_fill) * 2 + h) * 10 + lane] = make_float2(s1, s2);
            }
            nf_fill++;
            __syncwarp();  // PSD / PART reads done before the next block's E stores
            if (nf_fill == NFB) {
                __syncthreads();
                nf_select();
            }
        }
        __syncthreads();
        nf_select();

        // ---- end of segment: flush or save the cumulation ----
        float *dst = ((sg.flush_idx >= 0) ? a.flush_cum + (size_t)sg.flush_idx * N : a.cum_state + (size_t)sg.state_out * N) +
                     2 * lane + h;
#pragma unroll
        for (int q = 0; q < 16; q++) {
            dst[64 * ((OutIdx<32>::of(2 * q) + 16) & 31)] = cum2[q].x;
            dst[64 * ((OutIdx<32>::of(2 * q + 1) + 16) & 31)] = cum2[q].y;
        }
    }
}

}  // namespace sdr
