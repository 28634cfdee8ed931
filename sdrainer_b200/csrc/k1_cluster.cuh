// k1_cluster.cuh -- N = 65536 (BASELINE config 5: 24.576 MS/s wideband peak scan) in ONE pass over HBM: a thread-block
// CLUSTER of 16 CTAs holds a whole 512 KB block in distributed shared memory.
//
// Same reference arithmetic as the rest of K1 (dsp/fft.go:23-85, 215-252; rx/receiver.go:393, 404-407).
//
// The two-kernel four-step path (k1_large.cuh) moves 32 N bytes per block through HBM/L2 (IQ, intermediate out and in,
// dB spectrum out and in) against 8 N algorithmic.  Here the 256 x 256 factorisation stays on chip:
//   phase 1  CTA r owns the 16 columns c = 16 r .. 16 r + 15: tile load (128-byte row segments) into per-column
//            buffers, half-warp 256-point transforms (k1_large.cuh), twiddle W_N^(c k) (the [k][c] table is symmetric
//            for N1 = N2, so it is read [c][k]: coalesced), results stored row-major in the CTA's exchange tile
//            T[k][f];
//   cluster barrier;
//   phase 2  CTA j owns the 16 rows k = 16 j .. 16 j + 15: half-warp f gathers row k = 16 j + f through DSMEM --
//            lane hl reads Z[k][16 n1 + hl] from CTA n1's tile, 128 contiguous bytes per half-warp and peer -- runs
//            the 256-point transform, and finishes like fast_rows256_kernel: |X|^2, dB, this CTA's share of the ten
//            noise-window sums, x_to, its listener taps -- plus the cumulation, which now lives in registers (16 bins
//            per thread, sequential float32 adds in block order) because a cluster walks its segment's blocks in order;
//   cluster barrier before the tiles are rewritten.
// HBM sees the IQ once, 10 partial window sums per CTA and block, and the cumulation once per 100 blocks.
// DSMEM moves 32 KB per CTA and block at ~21 B/clk (B300_MICROARCH.md), about the SM's share of HBM bandwidth.
//
// Status (round 1, measured on B200): results identical to the two-kernel path within fp32 error
// (tests/test_gpu_cluster.py), DRAM traffic down to the algorithmic 8 N -- but 5.1 ms against 2.8 ms per 64 streams x
// 100 blocks: only 13 clusters (208 CTAs, 1.4 per SM) are resident, and inside a CTA the phases run back to back
// (tile load latency, two cluster barriers, the DSMEM gather), so the SMs idle on long-scoreboard / barrier stalls
// (issue slots 21 % busy).  It needs the next block's tile in flight (cp.async into a second buffer) and a
// producer/consumer split before it can pay off; until then it is opt-in (SDR_K1_CLUSTER=1).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "k1_large.cuh"

namespace sdr {

constexpr int K1C_CLUSTER = 16;
constexpr int K1C_TPITCH = 17;  // exchange tile: [256 rows k][16 columns f], pitch 17 complex (conflict-free column writes)
constexpr int K1C_SMEM_BYTES = 16 * HW_PITCH * 8 + 256 * K1C_TPITCH * 8;

struct ClusterArgs {
    K1Args a;
    const float2 *tw256;    // W_256^m
    const float2 *tw_step;  // [k][c] = W_N^(c k) (symmetric)
    double2 *nf_part;       // [blocks][16][10]
    float *xto;             // [blocks][10]
    int *nf_edge;           // [blocks]
    float db_offset;        // 10*log10(20/N^2)
};

__global__ void __launch_bounds__(256, 2) k1_cluster_kernel(const ClusterArgs ca) {
    namespace cg = cooperative_groups;
    constexpr int N = 65536, N1 = 256;
    const K1Args &a = ca.a;
    extern __shared__ __align__(16) unsigned char cl_smem[];
    float2 *cols = reinterpret_cast<float2 *>(cl_smem);                      // [16][HW_PITCH]: column buffers, then row scratch + (psd, dB) tile
    float2 *T = reinterpret_cast<float2 *>(cl_smem + 16 * HW_PITCH * 8);     // [256][K1C_TPITCH] exchange tile
    __shared__ int cnt[11];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / K1C_CLUSTER, ncl = gridDim.x / K1C_CLUSTER;
    const int tid = threadIdx.x, hl = tid & 15, f = tid >> 4, lane = tid & 31, warp = tid >> 5;
    const int c0 = 16 * rank;  // phase 1: my columns; phase 2: my rows r0 = 16 * rank
    const int r0 = 16 * rank;
    HwTwiddle t;
    hw_twiddle_load(t, ca.tw256, hl);
    // peers' exchange tiles (generic pointers into distributed shared memory)
    bool first = true;

    for (int seg = cid; seg < a.n_segs; seg += ncl) {
        const Segment sg = a.segs[seg];
        const WorkParams wp = a.works[sg.work];
        const int e = wp.edge_width;
        const int ws = nf_window_size(N, e), n_win = nf_window_count(N, e);
        __syncthreads();  // cnt of the previous segment is no longer read
        if (tid < 11) cnt[tid] = tile_count_below(e + tid * ws, r0, N1, 256);
        const int *lbins = a.listener_bins + wp.listener_off;
        // cumulation registers: cum[p] is bin kk = (r0 + f) + 256*((hl + 16*OutIdx<16>(p) + 128) & 255)
        float cum[16];
#pragma unroll
        for (int p = 0; p < 16; p++) {
            const int kk = (r0 + f) + N1 * ((hl + 16 * OutIdx<16>::of(p) + 128) & 255);
            cum[p] = sg.state_in >= 0 ? a.cum_state[(size_t)sg.state_in * N + kk] : 0.f;
        }
        const float2 *iq = reinterpret_cast<const float2 *>(sg.iq);

        for (int blk = 0; blk < sg.n_blocks; blk++) {
            const int ob = sg.block_out + blk;
            // ---------------- phase 1: columns c0 .. c0 + 15 ----------------
            const float2 *src = iq + (size_t)blk * N + c0;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int r = f + 16 * i;
                float2 x = __ldg(&src[(size_t)r * 256 + hl]);
                if (a.window) {
                    const float w = __ldg(&a.window[r * 256 + c0 + hl]);
                    x = __fmul2_rn(x, make_float2(w, w));
                }
                cols[hl * HW_PITCH + r] = x;
            }
            __syncthreads();
            float2 v[16];
            fft256_halfwarp(v, cols + f * HW_PITCH, t, hl);
            // every CTA of the cluster has finished gathering the previous block from my tile
            if (!first) cluster.sync();
            first = false;
            {
                const float2 *tw = ca.tw_step + (size_t)(c0 + f) * 256;  // W_N^(c k), c = c0 + f, read along k
#pragma unroll
                for (int p = 0; p < 16; p++) {
                    const int k = hl + 16 * OutIdx<16>::of(p);
                    T[k * K1C_TPITCH + f] = cmul(v[p], __ldg(&tw[k]));
                }
            }
            cluster.sync();  // all sixteen tiles are complete
            // ---------------- phase 2: row k = r0 + f gathered through distributed shared memory ----------------
            {
                const int krow = r0 + f;
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int n1 = (q & 3) * 4 + (q >> 2);  // first-layer order; peer n1 holds columns 16 n1 .. 16 n1 + 15
                    const float2 *peer = cluster.map_shared_rank(T, n1);
                    v[n1] = peer[krow * K1C_TPITCH + hl];
                }
            }
            float2 *col = cols + f * HW_PITCH;  // scratch for the transposes, then the (psd, dB) tile of this row
            fft256_halfwarp_regs(v, col, t, hl);
            __syncwarp();
            // X[k + 256*k2], k2 = hl + 16*OutIdx<16>(p)
#pragma unroll
            for (int p = 0; p < 16; p++) {
                const float psd = fmaf(v[p].x, v[p].x, v[p].y * v[p].y);                                     // dsp/fft.go:71-73
                const float db = __fadd_rn(fmaf(3.01029995663981195f, fast_log2(psd), ca.db_offset), 120.0f);  // rx/receiver.go:376-378
                cum[p] = __fadd_rn(cum[p], db);                                                                // rx/receiver.go:404-406
                col[hl + 16 * OutIdx<16>::of(p)] = make_float2(psd, db);
            }
            __syncthreads();
            if (a.dbg_psd) {  // parity / scope only: fftshifted stores (dsp/fft.go:54-57), half-warp = 16 consecutive bins
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int k2 = f + 16 * i;
                    const float2 o = cols[hl * HW_PITCH + k2];
                    const int kk = ((r0 + hl) + N1 * k2 + N / 2) & (N - 1);
                    a.dbg_psd[(size_t)ob * N + kk] = o.x;
                    a.dbg_spectrum[(size_t)ob * N + kk] = o.y;
                }
            }
            // noise-window sums over this CTA's bins in ascending bin order t = 16*k2' + f (float64 sums of float32 values)
            for (int w = warp; w < 10; w += 8) {
                double a1 = 0.0, a2 = 0.0;
                for (int i = cnt[w] + lane; i < cnt[w + 1]; i += 32) {
                    const double x = (double)cols[(i & 15) * HW_PITCH + (((i >> 4) + 128) & 255)].x;
                    a1 += x;
                    a2 = fma(x, x, a2);
                }
                a1 = warp_sum(a1);
                a2 = warp_sum(a2);
                if (lane == 0) ca.nf_part[((size_t)ob * K1C_CLUSTER + rank) * 10 + w] = make_double2(a1, a2);
            }
            if (tid < n_win) {  // x_to = psd[e + (w+1)*ws] (dsp/fft.go:238-243) if this CTA owns that bin
                const int k = (e + (tid + 1) * ws - N / 2) & (N - 1);
                const int k1 = k & (N1 - 1);
                if (k1 >= r0 && k1 < r0 + 16) ca.xto[(size_t)ob * 10 + tid] = cols[(k1 - r0) * HW_PITCH + k / N1].x;
            }
            if (rank == 0 && tid == 0) ca.nf_edge[ob] = e;
            for (int l = tid; l < wp.n_listeners; l += 256) {  // listener taps (rx/receiver.go:393) on bins this CTA owns
                const int k = (__ldg(&lbins[l]) - N / 2) & (N - 1);
                const int k1 = k & (N1 - 1);
                if (k1 >= r0 && k1 < r0 + 16) a.taps[(size_t)ob * a.tap_stride + l] = cols[(k1 - r0) * HW_PITCH + k / N1].y;
            }
            __syncthreads();  // the (psd, dB) tile is read: the column buffers may be reloaded
        }
        float *dst = (sg.flush_idx >= 0) ? a.flush_cum + (size_t)sg.flush_idx * N : a.cum_state + (size_t)sg.state_out * N;
#pragma unroll
        for (int p = 0; p < 16; p++) dst[(r0 + f) + N1 * ((hl + 16 * OutIdx<16>::of(p) + 128) & 255)] = cum[p];
    }
    cluster.sync();  // no CTA leaves while a peer may still read its tile
}

}  // namespace sdr
