// k1_large.cuh -- large-block path (N = 8192 .. 65536: BASELINE configs 3 and 5) as TWO launches per round.
//
// One block no longer fits a CTA's registers/shared memory, so the transform is the classic four-step N = N1 x 256
// split (N1 = 32 | 64 | 128 | 256) whose intermediate stays L2-resident where it can (126 MB L2):
//   step 1  fast_cols{32,64,256}_kernel: the 256 column transforms of length N1 (x[r*256 + c], r < N1) in registers,
//           multiplied by W_N^(c*k1), written to tmp[k1*256 + c]
//   step 2  fast_rows256_kernel: N1 row transforms of 256 points (a half-warp each) over tmp[k1*256 + .]; X[k1 + N1*k2]
//           goes through the same |X|^2 / dB projection as K1 (dsp/fft.go:32-36,71-85, rx/receiver.go:376-378), the
//           CTA's share of the ten noise-window sums, x_to and the taps are fused in; the dB spectrum goes to one buffer
//   cum     large_round_cum_kernel: float32 cumulation in block order (rx/receiver.go:404-407), flush / state save
//           (or fast_rows256_seg_kernel: rows + cumulation in one segment-sequential kernel when there are enough segments)
//   finish  large_nf_finish_kernel: adds the per-CTA window sums and runs dsp.FindNoiseFloor's selection
// This is the block-parallel path: launches with few segments (a single wideband stream) and N = 16384 / 32768 use it;
// the single-pass kernels (k1_mid8k.cuh, k1_wide.cuh) take N = 8192 / 65536 when a launch has enough segments.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "k1_spectral.cuh"

namespace sdr {

struct LargeGeom {
    int n, n1, n2;  // N = n1 * n2; (32 | 64 | 128 | 256) x 256 run the register-resident kernels
};
__host__ __device__ inline bool large_fast_geom(int n1, int n2) { return n2 == 256 && (n1 == 32 || n1 == 64 || n1 == 128 || n1 == 256); }
__host__ __device__ inline bool large_geom(int n, LargeGeom *g) {
    g->n = n;
    switch (n) {
        case 8192: g->n1 = 32; g->n2 = 256; return true;
        case 16384: g->n1 = 64; g->n2 = 256; return true;
        case 32768: g->n1 = 128; g->n2 = 256; return true;
        case 65536: g->n1 = 256; g->n2 = 256; return true;
    }
    return false;
}

// ---- register-resident sub-transforms (N = N1 x 256) ---------------------------------------------------------------
// A 256-point transform is run by a HALF-WARP: lane hl keeps 16 points, radix-16 over n1 (x[16 n1 + hl]), twiddle
// W256^(hl k1), one 16x16 transpose through the FFT's own shared-memory column (pitch 17, conflict-free, __syncwarp
// only), radix-16 over n2.  Lane hl ends with X[hl + 16*OutIdx<16>(p)] in register p.  Same packed-f32x2 butterflies
// as K1 (fft_radix.cuh); no CTA barrier inside the transform.
constexpr int HW_PITCH = 273;  // complex slots per column: >= 16*17 (transpose), odd multiple-of-16 remainder 1

// Once a row's transform is done, its (dead) transpose column takes the row's |X|^2 (and dB) as a float plane.  With
// the plain pitch (2*HW_PITCH = 546 words == 2 mod 32) the two half-warps of a warp -- rows j and j + 1 -- store to
// banks that overlap in 14 of 16 positions (two wavefronts per STS) and the sixteen rows a noise-window share reads
// side by side fall on all even banks, so two shares in one warp collide whenever their offsets differ by an even
// number.  The plane of row j therefore starts at bank beta(j) = 2 (j / 2) + 16 (j % 2): neighbouring rows are exactly
// 16 banks apart (one wavefront per warp-wide STS.32) and the sixteen rows still cover sixteen distinct (even) banks.
// row_word0: word offset of the row's column in the scratch; j: the row's index among the sixteen rows that are stored /
// read together.  The skew is < 32 words and the planes take 512: both fit the column's 2*HW_PITCH words.
static_assert(2 * HW_PITCH >= 512 + 32, "planes + skew must fit the row's own column");
__device__ __forceinline__ int plane_skew(int row_word0, int j) { return (2 * (j >> 1) + 16 * (j & 1) - row_word0) & 31; }

// One thread's share of a noise-window sum (dsp/fft.go:226-236): the window's positions in one row, float32 inside the
// share.  base[off] is position P of the row, the same P for the sixteen rows of the half-warp (so their banks stay
// beta(j) + const); the row's positions are the indices [lo, hi) with lo in {0, 1}.  A window has m = ws / stride or
// m + 1 positions in every row (m: the caller's uniform minimum, clamped to >= 2), hence hi <= m + 2 and the indices
// 1 .. m - 1 are inside the window for EVERY lane: they are read without a per-lane test.  rot in {0, 1} makes the lane
// walk that body rotated by one (index i + rot in step i): the upper half-warp uses it when its P has the parity of the
// lower one's, which puts the two halves on disjoint (even / odd) banks.  The at most four remaining indices
// (0, m - 1 or 1, m, m + 1) are tested against [lo, hi).
template <int M_MAX>
__device__ __forceinline__ void nf_row_share(const float *base, int off, int lo, int hi, int rot, int m, float &s1, float &s2) {
    int o = off + rot;
    asm volatile("" : "+r"(o));  // one address register for the whole body (otherwise re-derived under every predicate)
    const float *pr = base + o;
    float2 a1 = make_float2(0.f, 0.f), a2 = a1, b1 = a1, b2 = a1;
#pragma unroll
    for (int i = 1; i <= M_MAX - 2; i += 4) {  // body: steps 1 .. m - 2 (uniform bound), two packed chains
        const float x0 = (i <= m - 2) ? pr[i] : 0.f, x1 = (i + 1 <= m - 2) ? pr[i + 1] : 0.f;
        const float x2 = (i + 2 <= m - 2) ? pr[i + 2] : 0.f, x3 = (i + 3 <= m - 2) ? pr[i + 3] : 0.f;
        const float2 p = make_float2(x0, x1), q = make_float2(x2, x3);
        a1 = __fadd2_rn(a1, p);
        a2 = __ffma2_rn(p, p, a2);
        b1 = __fadd2_rn(b1, q);
        b2 = __ffma2_rn(q, q, b2);
    }
    const float *pp = base + off;
    const int t1 = rot ? 1 : m - 1;
    const float y0 = (lo == 0 && hi > 0) ? pp[0] : 0.f;
    const float y1 = (t1 >= lo && t1 < hi) ? pp[t1] : 0.f;
    const float y2 = (m < hi) ? pp[m] : 0.f;
    const float y3 = (m + 1 < hi) ? pp[m + 1] : 0.f;
    const float2 p = make_float2(y0, y1), q = make_float2(y2, y3);
    a1 = __fadd2_rn(a1, p);
    a2 = __ffma2_rn(p, p, a2);
    b1 = __fadd2_rn(b1, q);
    b2 = __ffma2_rn(q, q, b2);
    s1 = (a1.x + a1.y) + (b1.x + b1.y);
    s2 = (a2.x + a2.y) + (b2.x + b2.y);
}

// The same share with every index tested per lane, scalar chains (k1_mid8k2: measured 2.7 % faster there than the packed
// form above, which costs that kernel eight more bytes of spill at its 128-register cap).  Indices [lo, hi), hi <= NFMAX - 1;
// rot as above (index i + rot, the last one wraps to 0).
template <int NFMAX>
__device__ __forceinline__ void nf_row_share_tested(const float *pp, int lo, int hi, int rot, float &s1, float &s2) {
    static_assert(NFMAX % 2 == 0, "pairs");
    const float *pr = pp + rot;
    const int hi_r = hi - rot;  // i + rot < hi  <=>  i < hi_r
    const bool first_ok = lo == 0 && hi > 0;
    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
    for (int i = 0; i < NFMAX; i += 2) {
        float x0, x1;
        if (i == 0) x0 = (rot ? 0 < hi_r : first_ok) ? pr[0] : 0.f;
        else x0 = (i < hi_r) ? pr[i] : 0.f;
        if (i + 1 == NFMAX - 1) x1 = rot ? (first_ok ? pp[0] : 0.f) : ((i + 1 < hi_r) ? pr[i + 1] : 0.f);
        else x1 = (i + 1 < hi_r) ? pr[i + 1] : 0.f;
        s1a += x0;
        s2a = fmaf(x0, x0, s2a);
        s1b += x1;
        s2b = fmaf(x1, x1, s2b);
    }
    s1 = s1a + s1b;
    s2 = s2a + s2b;
}

struct HwTwiddle {
    float2 w[15];  // W256^(hl * k1), k1 = 1..15
};
__device__ __forceinline__ void hw_twiddle_load(HwTwiddle &t, const float2 *__restrict__ tw256, int hl) {
#pragma unroll
    for (int k = 1; k < 16; k++) t.w[k - 1] = __ldg(&tw256[(hl * k) & 255]);
}
// v[n1] = x[16 n1 + hl] on entry; col: 16*17 complex slots of scratch for the transpose; all 32 lanes of the warp
// call it (two FFTs per warp)
__device__ __forceinline__ void fft256_halfwarp_regs(float2 (&v)[16], float2 *col, const HwTwiddle &t, int hl) {
    dft16(v);
#pragma unroll
    for (int p = 0; p < 16; p++) {
        const int k1 = OutIdx<16>::of(p);
        if (k1 > 0) v[p] = cmul(v[p], t.w[k1 - 1]);
    }
    __syncwarp();  // every lane has read its inputs: the column may be overwritten by the transpose
#pragma unroll
    for (int p = 0; p < 16; p++) col[OutIdx<16>::of(p) * 17 + hl] = v[p];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const int n2 = (q & 3) * 4 + (q >> 2);
        v[n2] = col[hl * 17 + n2];
    }
    dft16(v);
}

// The same transform with the fifteen lane twiddles formed from four table values (w1, w2, w4, w8 = W256^(hl), ^(2 hl),
// ^(4 hl), ^(8 hl)) by eleven products: 4 shared-memory reads per transform instead of 15, 22 more packed instructions.
__device__ __forceinline__ void fft256_halfwarp_regs_p2(float2 (&v)[16], float2 *col, float2 w1, float2 w2, float2 w4, float2 w8, int hl) {
    dft16(v);
#define SDR_PK(k) (4 * ((k) & 3) + ((k) >> 2))  // register that holds frequency index k (OutIdx<16>)
    v[SDR_PK(1)] = cmul(v[SDR_PK(1)], w1);
    v[SDR_PK(2)] = cmul(v[SDR_PK(2)], w2);
    v[SDR_PK(4)] = cmul(v[SDR_PK(4)], w4);
    v[SDR_PK(8)] = cmul(v[SDR_PK(8)], w8);
    {
        const float2 w3 = cmul(w1, w2);
        v[SDR_PK(3)] = cmul(v[SDR_PK(3)], w3);
        v[SDR_PK(11)] = cmul(v[SDR_PK(11)], cmul(w3, w8));
        const float2 w7 = cmul(w3, w4);
        v[SDR_PK(7)] = cmul(v[SDR_PK(7)], w7);
        v[SDR_PK(15)] = cmul(v[SDR_PK(15)], cmul(w7, w8));
    }
    {
        const float2 w5 = cmul(w1, w4);
        v[SDR_PK(5)] = cmul(v[SDR_PK(5)], w5);
        v[SDR_PK(13)] = cmul(v[SDR_PK(13)], cmul(w5, w8));
        const float2 w6 = cmul(w2, w4);
        v[SDR_PK(6)] = cmul(v[SDR_PK(6)], w6);
        v[SDR_PK(14)] = cmul(v[SDR_PK(14)], cmul(w6, w8));
    }
    v[SDR_PK(9)] = cmul(v[SDR_PK(9)], cmul(w1, w8));
    v[SDR_PK(10)] = cmul(v[SDR_PK(10)], cmul(w2, w8));
    v[SDR_PK(12)] = cmul(v[SDR_PK(12)], cmul(w4, w8));
#undef SDR_PK
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 16; p++) col[OutIdx<16>::of(p) * 17 + hl] = v[p];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const int n2 = (q & 3) * 4 + (q >> 2);
        v[n2] = col[hl * 17 + n2];
    }
    dft16(v);
}

// The register-resident path runs in ROUNDS of consecutive blocks [blk0, blk0 + round blocks): the four-step
// intermediate and the dB spectrum of a round are addressed relative to blk0 and sized to stay L2-resident, so HBM
// sees the IQ once; everything a block contributes per block (noise-window partial sums, taps) is produced by the
// row kernel, the cumulation by large_round_cum_kernel over the round's segments.
// col: this FFT's 256 inputs in natural order (pitch HW_PITCH), overwritten by the transpose
__device__ __forceinline__ void fft256_halfwarp(float2 (&v)[16], float2 *col, const HwTwiddle &t, int hl) {
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const int n1 = (q & 3) * 4 + (q >> 2);
        v[n1] = col[16 * n1 + hl];
    }
    __syncwarp();  // every lane has read its inputs: the column may be overwritten by the transpose
    fft256_halfwarp_regs(v, col, t, hl);
}

struct FastStepArgs {
    float2 *tmp;            // [round blocks][N] four-step intermediate
    float *spec_round;      // [round blocks][N] dB spectrum of the round (cumulation input)
    float *spectrum, *psd;  // [blocks][N] or nullptr (parity / scope)
    const float2 *tw256;    // W_256^m
    const float2 *tw_n1;    // W_N1^m (the N1 = 128 column kernel's parity twiddle)
    const float2 *tw_step;  // [k1][c] = W_N^(c k1), k1 < N1, c < N2 (step 1)
    const float *window;
    const Segment *segs;    // all segments of the batch
    const int *block_seg;   // [blocks] segment of each block
    const WorkParams *works;
    const int *listener_bins;
    double2 *nf_part;       // [blocks][N1/16][10] (sum x, sum x^2) over this CTA's bins per noise window
    float *xto;             // [blocks][10] psd[first bin of the next window]
    int *nf_edge;           // [blocks] edge width (for the finish kernel)
    float *taps;
    int tap_stride;
    int blk0;               // first block of the round
    int n, n1, n2;
    float db_offset;
};

// number of tile elements (row k2' < rows, column f < 16; bin kk = r0 + f + n1*k2') whose bin is below `bin`
__device__ __forceinline__ int tile_count_below(int bin, int r0, int n1, int rows) {
    if (bin <= r0) return 0;
    const int d = bin - r0, q = d / n1, rem = d - q * n1;
    if (q >= rows) return rows * 16;
    return q * 16 + (rem < 16 ? rem : 16);
}

// step 1, N1 = 256: grid (N2 / 16, blocks), 256 threads; half-warp f owns column c0 + f
__global__ void __launch_bounds__(256, 3) fast_cols256_kernel(const FastStepArgs a) {
    extern __shared__ __align__(16) unsigned char sub_smem[];
    float2 *cols = reinterpret_cast<float2 *>(sub_smem);  // [16][HW_PITCH]
    const int tid = threadIdx.x, hl = tid & 15, f = tid >> 4;
    const int blk = a.blk0 + blockIdx.y, c0 = blockIdx.x * 16;
    const int N = a.n, N2 = a.n2;
    const Segment sg = a.segs[a.block_seg[blk]];
    const float2 *src = reinterpret_cast<const float2 *>(sg.iq) + (size_t)(blk - sg.block_out) * N + c0;
    HwTwiddle t;
    hw_twiddle_load(t, a.tw256, hl);
    // tile load: half-warp = one 128-byte row segment (16 columns); transposed into per-column buffers
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int r = f + 16 * i;
        float2 x = __ldg(&src[(size_t)r * N2 + hl]);
        if (a.window) {
            const float w = __ldg(&a.window[r * N2 + c0 + hl]);
            x = __fmul2_rn(x, make_float2(w, w));
        }
        cols[hl * HW_PITCH + r] = x;
    }
    __syncthreads();
    float2 v[16];
    float2 *col = cols + f * HW_PITCH;
    fft256_halfwarp(v, col, t, hl);
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 16; p++) col[hl + 16 * OutIdx<16>::of(p)] = v[p];  // natural order k = hl + 16 k2
    __syncthreads();
    // A_c[k] * W_N^(c k) -> tmp[k*N2 + c]: half-warp = 128 contiguous bytes
    float2 *dst = a.tmp + (size_t)blockIdx.y * N + c0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int k = f + 16 * i;
        const float2 w = __ldg(&a.tw_step[(size_t)k * N2 + c0 + hl]);
        dst[(size_t)k * N2 + hl] = cmul(cols[hl * HW_PITCH + k], w);
    }
}

// step 1, N1 = 32: one thread per column, the whole 32-point transform in registers; grid (N2 / 256, blocks)
__global__ void __launch_bounds__(256) fast_cols32_kernel(const FastStepArgs a) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int blk = a.blk0 + blockIdx.y;
    const int N = a.n, N2 = a.n2;
    const Segment sg = a.segs[a.block_seg[blk]];
    const float2 *src = reinterpret_cast<const float2 *>(sg.iq) + (size_t)(blk - sg.block_out) * N + c;
    float2 v[32];
#pragma unroll
    for (int q = 0; q < 32; q++) {
        const int r = (q & 3) * 8 + (q >> 2);
        v[r] = __ldg(&src[(size_t)r * N2]);
        if (a.window) {
            const float w = __ldg(&a.window[r * N2 + c]);
            v[r] = __fmul2_rn(v[r], make_float2(w, w));
        }
    }
    dft32(v);
    float2 *dst = a.tmp + (size_t)blockIdx.y * N + c;
#pragma unroll
    for (int p = 0; p < 32; p++) {
        const int k1 = OutIdx<32>::of(p);
        float2 x = v[p];
        if (k1 > 0) x = cmul(x, __ldg(&a.tw_step[(size_t)k1 * N2 + c]));
        dst[(size_t)k1 * N2] = x;
    }
}

// step 1, N1 = 64 (SPLIT = false: one thread per column, grid (N2 / 256, blocks), 256 threads) and N1 = 128 (SPLIT = true:
// the 128-point column transform split by output parity between two threads, decimation in frequency as in k1_mid8k:
// thread (c, h) forms u[m] = x[m] + (-1)^h x[m + 64], multiplies by W128^m when h = 1 (h is warp-uniform), runs the
// 64-point transform and owns the outputs k1 = 2j + h; grid (N2 / 128, blocks), 256 threads).  The whole transform in
// registers: 64 complex values per thread, one CTA per SM -- these block sizes (16384 / 32768) are no BASELINE shape, the
// kernel exists so that they take the register-resident two-kernel path instead of the generic Stockham one.
template <bool SPLIT>
__global__ void __launch_bounds__(256, 1) fast_cols64_kernel(const FastStepArgs a) {
    const int c = SPLIT ? blockIdx.x * 128 + (threadIdx.x & 127) : blockIdx.x * 256 + threadIdx.x;
    const int h = SPLIT ? threadIdx.x >> 7 : 0;
    const int blk = a.blk0 + blockIdx.y;
    const int N = a.n, N2 = a.n2;
    const Segment sg = a.segs[a.block_seg[blk]];
    const float2 *src = reinterpret_cast<const float2 *>(sg.iq) + (size_t)(blk - sg.block_out) * N + c;
    float2 v[64];
#pragma unroll
    for (int q = 0; q < 64; q++) {
        const int r = (q & 7) * 8 + (q >> 3);  // issue order = consumption order of the first layer
        float2 x0 = __ldg(&src[(size_t)r * N2]);
        if (a.window) {
            const float w = __ldg(&a.window[r * N2 + c]);
            x0 = __fmul2_rn(x0, make_float2(w, w));
        }
        if (SPLIT) {
            float2 x1 = __ldg(&src[(size_t)(r + 64) * N2]);
            if (a.window) {
                const float w = __ldg(&a.window[(r + 64) * N2 + c]);
                x1 = __fmul2_rn(x1, make_float2(w, w));
            }
            const float2 sgn = h ? make_float2(-1.f, -1.f) : make_float2(1.f, 1.f);
            x0 = __ffma2_rn(x1, sgn, x0);  // x0 + (-1)^h x1, exact
            if (h && r > 0) x0 = cmul(x0, __ldg(&a.tw_n1[r]));
        }
        v[r] = x0;
    }
    dft64(v);
    float2 *dst = a.tmp + (size_t)blockIdx.y * N + c;
#pragma unroll
    for (int p = 0; p < 64; p++) {
        const int k1 = SPLIT ? 2 * OutIdx<64>::of(p) + h : OutIdx<64>::of(p);
        float2 x = v[p];
        if (SPLIT || OutIdx<64>::of(p) > 0) x = cmul(x, __ldg(&a.tw_step[(size_t)k1 * N2 + c]));
        dst[(size_t)k1 * N2] = x;
    }
}

// step 2 + epilogue, N2 = 256: grid (N1 / 16, round blocks), 256 threads; half-warp f owns row r0 + f, i.e. the bins
// r0 + f + N1*k2.  Fused in: |X|^2 (dsp/fft.go:71-73), dB + 120 (rx/receiver.go:376-378), this CTA's share of the ten
// noise-window sums of dsp.FindNoiseFloor (dsp/fft.go:215-252), the listener taps on bins it owns (rx/receiver.go:393).
__global__ void __launch_bounds__(256) fast_rows256_kernel(const FastStepArgs a) {
    extern __shared__ __align__(16) unsigned char sub_smem[];
    float2 *cols = reinterpret_cast<float2 *>(sub_smem);  // [16][HW_PITCH]; reused as the (psd, dB) output tile
    __shared__ int cnt[11];
    const int tid = threadIdx.x, hl = tid & 15, f = tid >> 4, lane = tid & 31, warp = tid >> 5;
    const int blk = a.blk0 + blockIdx.y, r0 = blockIdx.x * 16;
    const int N = a.n, N1 = a.n1;
    const Segment sg = a.segs[a.block_seg[blk]];
    const WorkParams wp = a.works[sg.work];
    const int e = wp.edge_width;
    const int ws = nf_window_size(N, e), n_win = nf_window_count(N, e);
    if (tid < 11) cnt[tid] = tile_count_below(e + tid * ws, r0, N1, 256);
    HwTwiddle t;
    hw_twiddle_load(t, a.tw256, hl);
    // row r0 + f straight into the registers of its half-warp: lane hl takes x[16 n1 + hl], 128 contiguous bytes
    // per half-warp and n1 (no shared-memory staging of the input)
    const float2 *src = a.tmp + (size_t)blockIdx.y * N + (size_t)(r0 + f) * 256 + hl;
    float2 v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const int n1 = (q & 3) * 4 + (q >> 2);
        v[n1] = src[16 * n1];
    }
    float2 *col = cols + f * HW_PITCH;
    fft256_halfwarp_regs(v, col, t, hl);
    __syncwarp();
    // X[k1 + N1*k2], k2 = hl + 16*OutIdx<16>(p)
#pragma unroll
    for (int p = 0; p < 16; p++) {
        const float psd = fmaf(v[p].x, v[p].x, v[p].y * v[p].y);
        const float db = __fadd_rn(fmaf(3.01029995663981195f, fast_log2(psd), a.db_offset), 120.0f);
        col[hl + 16 * OutIdx<16>::of(p)] = make_float2(psd, db);
    }
    __syncthreads();
    // fftshifted stores (dsp/fft.go:54-57): half-warp = 16 consecutive bins
    float *spec = a.spec_round + (size_t)blockIdx.y * N;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int k2 = f + 16 * i;
        const float2 o = cols[hl * HW_PITCH + k2];
        const int kk = ((r0 + hl) + N1 * k2 + N / 2) & (N - 1);
        spec[kk] = o.y;
        if (a.psd) {
            a.psd[(size_t)blk * N + kk] = o.x;
            a.spectrum[(size_t)blk * N + kk] = o.y;
        }
    }
    // noise-window sums over this CTA's bins in ascending bin order t = 16*k2' + f (k2' = fftshifted k2):
    // float64 sums of float32 values as in the reference
    for (int w = warp; w < 10; w += 8) {
        double a1 = 0.0, a2 = 0.0;
        for (int i = cnt[w] + lane; i < cnt[w + 1]; i += 32) {
            const double x = (double)cols[(i & 15) * HW_PITCH + (((i >> 4) + 128) & 255)].x;
            a1 += x;
            a2 = fma(x, x, a2);
        }
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (lane == 0) a.nf_part[((size_t)blk * gridDim.x + blockIdx.x) * 10 + w] = make_double2(a1, a2);
    }
    if (tid < n_win) {  // x_to = psd[e + (w+1)*ws] (dsp/fft.go:238-243) if this CTA owns that bin
        const int k = (e + (tid + 1) * ws - N / 2) & (N - 1);
        const int k1 = k & (N1 - 1);
        if (k1 >= r0 && k1 < r0 + 16) a.xto[(size_t)blk * 10 + tid] = cols[(k1 - r0) * HW_PITCH + k / N1].x;
    }
    if (blockIdx.x == 0 && tid == 0) a.nf_edge[blk] = e;
    const int *lbins = a.listener_bins + wp.listener_off;
    for (int l = tid; l < wp.n_listeners; l += 256) {
        const int k = (__ldg(&lbins[l]) - N / 2) & (N - 1);
        const int k1 = k & (N1 - 1);
        if (k1 >= r0 && k1 < r0 + 16) a.taps[(size_t)blk * a.tap_stride + l] = cols[(k1 - r0) * HW_PITCH + k / N1].y;
    }
}

// step 2 + epilogue + cumulation, segment-sequential variant of fast_rows256_kernel: grid (N1 / 16, segments of the
// round).  A CTA walks the blocks of ONE segment in order and keeps the cumulation of its 4096 bins in registers
// (16 per thread, sequential float32 adds in block order, rx/receiver.go:404-407): no dB-spectrum round trip and no
// separate cumulation launch.  Chosen by the engine when segments x N1/16 CTAs fill the GPU several times over;
// with few streams the block-parallel kernel + large_round_cum_kernel is faster.
struct FastRowsSegArgs {
    FastStepArgs s;
    int seg0;  // first segment of the round
    float *cum_state, *flush_cum;
};
__global__ void __launch_bounds__(256) fast_rows256_seg_kernel(const FastRowsSegArgs sa) {
    const FastStepArgs &a = sa.s;
    extern __shared__ __align__(16) unsigned char sub_smem[];
    float2 *cols = reinterpret_cast<float2 *>(sub_smem);  // [16][HW_PITCH]: transpose scratch, then the (psd, dB) tile
    __shared__ int cnt[11];
    const int tid = threadIdx.x, hl = tid & 15, f = tid >> 4, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.x * 16;
    const int N = a.n, N1 = a.n1;
    const Segment sg = a.segs[sa.seg0 + blockIdx.y];
    const WorkParams wp = a.works[sg.work];
    const int e = wp.edge_width;
    const int ws = nf_window_size(N, e), n_win = nf_window_count(N, e);
    if (tid < 11) cnt[tid] = tile_count_below(e + tid * ws, r0, N1, 256);
    HwTwiddle t;
    hw_twiddle_load(t, a.tw256, hl);
    const int *lbins = a.listener_bins + wp.listener_off;
    float cum[16];
#pragma unroll
    for (int p = 0; p < 16; p++) {
        const int kk = ((r0 + f) + N1 * (hl + 16 * OutIdx<16>::of(p)) + N / 2) & (N - 1);
        cum[p] = sg.state_in >= 0 ? sa.cum_state[(size_t)sg.state_in * N + kk] : 0.f;
    }
    float2 *col = cols + f * HW_PITCH;
    for (int b = 0; b < sg.n_blocks; b++) {
        const int blk = sg.block_out + b;
        const float2 *src = a.tmp + (size_t)(blk - a.blk0) * N + (size_t)(r0 + f) * 256 + hl;
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int n1 = (q & 3) * 4 + (q >> 2);
            v[n1] = src[16 * n1];
        }
        fft256_halfwarp_regs(v, col, t, hl);
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 16; p++) {
            const float psd = fmaf(v[p].x, v[p].x, v[p].y * v[p].y);
            const float db = __fadd_rn(fmaf(3.01029995663981195f, fast_log2(psd), a.db_offset), 120.0f);
            cum[p] = __fadd_rn(cum[p], db);
            col[hl + 16 * OutIdx<16>::of(p)] = make_float2(psd, db);
        }
        __syncthreads();
        if (a.psd) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int k2 = f + 16 * i;
                const float2 o = cols[hl * HW_PITCH + k2];
                const int kk = ((r0 + hl) + N1 * k2 + N / 2) & (N - 1);
                a.psd[(size_t)blk * N + kk] = o.x;
                a.spectrum[(size_t)blk * N + kk] = o.y;
            }
        }
        for (int w = warp; w < 10; w += 8) {
            double a1 = 0.0, a2 = 0.0;
            for (int i = cnt[w] + lane; i < cnt[w + 1]; i += 32) {
                const double x = (double)cols[(i & 15) * HW_PITCH + (((i >> 4) + 128) & 255)].x;
                a1 += x;
                a2 = fma(x, x, a2);
            }
            a1 = warp_sum(a1);
            a2 = warp_sum(a2);
            if (lane == 0) a.nf_part[((size_t)blk * gridDim.x + blockIdx.x) * 10 + w] = make_double2(a1, a2);
        }
        if (tid < n_win) {
            const int k = (e + (tid + 1) * ws - N / 2) & (N - 1);
            const int k1 = k & (N1 - 1);
            if (k1 >= r0 && k1 < r0 + 16) a.xto[(size_t)blk * 10 + tid] = cols[(k1 - r0) * HW_PITCH + k / N1].x;
        }
        if (blockIdx.x == 0 && tid == 0) a.nf_edge[blk] = e;
        for (int l = tid; l < wp.n_listeners; l += 256) {
            const int k = (__ldg(&lbins[l]) - N / 2) & (N - 1);
            const int k1 = k & (N1 - 1);
            if (k1 >= r0 && k1 < r0 + 16) a.taps[(size_t)blk * a.tap_stride + l] = cols[(k1 - r0) * HW_PITCH + k / N1].y;
        }
        __syncthreads();  // the tile is read: the next block's transposes may overwrite it
    }
    float *dst = (sg.flush_idx >= 0) ? sa.flush_cum + (size_t)sg.flush_idx * N : sa.cum_state + (size_t)sg.state_out * N;
#pragma unroll
    for (int p = 0; p < 16; p++) dst[((r0 + f) + N1 * (hl + 16 * OutIdx<16>::of(p)) + N / 2) & (N - 1)] = cum[p];
}

struct LargeFinishArgs {
    const double2 *nf_part;
    const float *xto;
    const int *nf_edge;
    float *psd_floor;
    double *variance;
    int n_blocks, n_cta, n;
};

// one warp per block: fixed-order sum of the per-CTA window sums, then dsp.FindNoiseFloor's selection
__global__ void __launch_bounds__(128) large_nf_finish_kernel(const LargeFinishArgs a) {
    const int blk = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (blk >= a.n_blocks) return;
    const int e = a.nf_edge[blk];
    const int ws = nf_window_size(a.n, e), n_win = nf_window_count(a.n, e);
    double s1 = 0.0, s2 = 0.0, xt = 0.0;
    if (lane < n_win) {
        for (int c = 0; c < a.n_cta; c++) {
            const double2 p = a.nf_part[((size_t)blk * a.n_cta + c) * 10 + lane];
            s1 += p.x;
            s2 += p.y;
        }
        xt = (double)a.xto[(size_t)blk * 10 + lane];
    }
    nf_select_variance(s1, s2, xt, ws, n_win, lane, &a.psd_floor[blk], &a.variance[blk]);
}

struct RoundCumArgs {
    const float *spec_round;  // [round blocks][N]
    const Segment *segs;      // segments of this round
    float *cum_state;
    float *flush_cum;
    int blk0, n;
};

// grid (N / 256, segments of the round): thread = bin, blocks added in order (sequential float32, rx/receiver.go:404-407)
__global__ void __launch_bounds__(256) large_round_cum_kernel(const RoundCumArgs a) {
    const Segment sg = a.segs[blockIdx.y];
    const int bin = blockIdx.x * 256 + threadIdx.x;
    float cum = sg.state_in >= 0 ? a.cum_state[(size_t)sg.state_in * a.n + bin] : 0.f;
    const float *sp = a.spec_round + (size_t)(sg.block_out - a.blk0) * a.n + bin;
    for (int b = 0; b < sg.n_blocks; b++) cum = __fadd_rn(cum, sp[(size_t)b * a.n]);
    float *dst = (sg.flush_idx >= 0) ? a.flush_cum + (size_t)sg.flush_idx * a.n : a.cum_state + (size_t)sg.state_out * a.n;
    dst[bin] = cum;
}

}  // namespace sdr
