// k1_wide.cuh -- N = 65536 (BASELINE config 5: 24.576 MS/s wideband peak scan) in ONE pass over HBM.
//
// Same reference arithmetic as the rest of K1: dsp/fft.go:23-85 (FFT, fftshift, |X|^2, dB + 120), dsp/fft.go:215-252
// (FindNoiseFloor), rx/receiver.go:393 (listener taps), rx/receiver.go:404-407 (cumulation, float32, block order).
//
// A 512 KB block fits neither a CTA nor (usefully) a cluster, so the 256 x 256 four-step transform keeps its
// intermediate -- but in a small ring that never leaves the 126 MB L2, inside ONE persistent launch:
//   * a TEAM of 16 CTAs owns a segment (<= 100 consecutive blocks of one stream) and walks its blocks in order;
//   * per step every CTA of the team does two things.  PRODUCE step i + D - 1: its column tile of the IQ block (16
//     columns x 256 rows: ONE tensor-map TMA load, cp.async.bulk.tensor.3d -> SASS UTMALDG, 128-byte swizzle,
//     L2 evict-first) -> half-warp 256-point transforms -> twiddle W_N^(c k) -> swizzled tile in shared memory -> ONE
//     tensor-map TMA store (UTMASTG) into ring slot (i + D - 1) mod R.  CONSUME step i: its row tile (16 rows x 256 of
//     the intermediate, 32 KB contiguous, one cp.async.bulk) -> half-warp 256-point transforms -> |X|^2, dB, cumulation
//     in registers (16 bins per thread for the whole segment), this CTA's share of the ten noise-window sums, x_to, its
//     taps; the tile's lines are then discarded from L2 (they are dead until the slot is rewritten);
//   * warps 0-7 compute; warps 8 and 9 (one lane each) are the DMA warps: they own every wait on another CTA, every TMA
//     issue and every publication, the compute warps only wait on shared-memory mbarriers.  Warp 8 stores and publishes
//     what the CTA produces (and refills A), warp 9 requests what it consumes;
//   * a global counter per step (`ready`) is released by each producer's DMA thread after its tile store has completed
//     and acquired by each consumer's DMA thread before it requests the row tile: the only inter-CTA synchronisation.
//     All CTAs of the launch are co-resident (cooperative launch) and every wait is on work of EARLIER steps only:
//       - step j is produced in iteration j - D + 1 and published as soon as its store completes; its row tile is
//         requested in iteration j - 1 (once B was read there), D - 2 >= 1 iterations later: the team only has to stay
//         within D - 2 iterations of each other;
//       - ring safety: a CTA produces step j = i + D - 1 at the start of iteration i, after it consumed step i - 1, i.e.
//         after ready[i - 1] was complete: every team mate had produced step i - 1, which it does once it has consumed
//         (read into registers) every step <= i - D - 1; the row tiles of later steps may be in flight there.  The
//         slot's previous occupant is step j - R: safe iff j - R <= i - D - 1, i.e. R >= 2 D.
//     Waits are bounded all the same and report through `err` (sdr_collect turns it into SDR_ECUDA).
// HBM sees the IQ once (8 N per block), the ten partial window sums per CTA and block, and the cumulation once per 100
// blocks; the intermediate (8 N written + 8 N read per block) is L2 traffic: R x 512 KB per team.
// The two-kernel path of k1_large.cuh (32 N bytes of HBM traffic per block) remains for launches with too few segments
// to fill the GPU with teams.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "k1_large.cuh"

namespace sdr {

constexpr int K1W_TEAM = 16;
constexpr int K1W_TILE_BYTES = 32768;
constexpr int K1W_MAX_SPIN = 1 << 24;  // polls of one wait before the launch gives up (seconds; a step takes microseconds)

struct WideArgs {
    K1Args a;
    const CUtensorMap *seg_maps;  // [n_segs] IQ of each segment as {512 floats, 256 rows, n_blocks}, box {32, 256, 1}, 128B swizzle
    const CUtensorMap *tmp_map;   // the ring as {512 floats, 256 rows, n_teams * R}, same box
    float2 *tmp;                  // [n_teams * R][65536] intermediate Z[k][c] (row-major)
    int *ready;                   // [blocks of the batch] producers that have published the step (16 = complete)
    int *err;                     // set to 1 when a bounded wait ran out
    const float2 *tw256;          // W_256^m
    const float2 *tw_step;        // [k][c] = W_N^(c k) (symmetric: read as [c][k])
    double2 *nf_part;             // [blocks][16][10]
    float *xto;                   // [blocks][10]
    int *nf_edge;                 // [blocks]
    float db_offset;              // 10*log10(20/N^2)
    int lookahead;                // D >= 3: step i + D - 1 is produced in iteration i; its row tile is requested in iteration i + D - 2
    int ring;                     // R >= 2 D slots per team
    int discard;                  // drop consumed row tiles from L2 (discard.global.L2) instead of letting them be written back
#ifdef SDR_K1W_TRACE
    long long *trace;             // [32 CTAs][512 steps][32] clock64 stamps (tools/wide_trace.py), measurement builds only
#endif
};
#ifdef SDR_K1W_TRACE
#define K1W_TR(step, slot)                                                                                  \
    do {                                                                                                    \
        if (wa.trace && blockIdx.x < 32 && (step) < 512) wa.trace[((size_t)blockIdx.x * 512 + (step)) * 32 + (slot)] = clock64(); \
    } while (0)
#else
#define K1W_TR(step, slot) do { } while (0)
#endif

constexpr int K1W_OFF_A = 0;                                   // IQ column tile, later the outgoing tile (1024-byte aligned: 128B swizzle)
constexpr int K1W_OFF_B = K1W_OFF_A + K1W_TILE_BYTES;          // row tile of the intermediate
constexpr int K1W_PITCH_P = 280;                               // column pitch of the PRODUCE transposes (== 8 mod 16, see produce)
constexpr int K1W_OFF_S = K1W_OFF_B + K1W_TILE_BYTES;          // transpose scratch ([16][K1W_PITCH_P] produce, [16][HW_PITCH] consume), then the |X|^2 / dB planes
constexpr int K1W_OFF_TW = K1W_OFF_S + 16 * K1W_PITCH_P * 8;  // [15][16] W_256^(hl k)
constexpr int K1W_OFF_TQ = K1W_OFF_TW + 256 * 8;             // [16][17] W_N^(16 c q)
constexpr int K1W_OFF_MISC = K1W_OFF_TQ + 16 * 17 * 8;
constexpr int K1W_SMEM_BYTES = K1W_OFF_MISC + 128;
constexpr int K1W_NF_M = 25;                                   // whole positions of one noise window in one row: 6553 / 256
constexpr int K1W_THREADS = 320;                               // eight compute warps + the two DMA warps

__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_tile_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_store_tile_3d(const CUtensorMap *map, int c0, int c1, int c2, const void *src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0), "r"(c1),
                 "r"(c2), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int *p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// a tensor map that lives in global memory and was written by the host (cudaMemcpy) is acquired through the tensormap
// proxy before its first use (system scope: the writer is the host)
__device__ __forceinline__ void tensormap_acquire(const CUtensorMap *m) {
    asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(m) : "memory");
}

// the block sequence of one team: segments team, team + n_teams, ... in order, each segment's blocks in order
struct WideIter {
    int seg, blk, nb, step;
    __device__ __forceinline__ void start(const Segment *segs, int n_segs, int team) {
        seg = team;
        blk = 0;
        step = 0;
        nb = seg < n_segs ? segs[seg].n_blocks : 0;
    }
    __device__ __forceinline__ void next(const Segment *segs, int n_segs, int n_teams) {
        step++;
        if (++blk == nb) {
            seg += n_teams;
            blk = 0;
            nb = seg < n_segs ? segs[seg].n_blocks : 0;
        }
    }
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }  // the eight compute warps
__device__ __forceinline__ void discard_l2_line(const void *p) { asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory"); }

// Warps 0-7 compute; warps 8 and 9 (one lane each) are the DMA warps: they own every wait on another CTA, every TMA issue
// and every publication, so that the compute warps only ever wait on shared-memory mbarriers.
//   FULL_A   tx barrier: the IQ column tile has landed in A            (warp 8 -> compute)
//   FULL_B   tx barrier: the row tile of the intermediate has landed in B (warp 9 -> compute)
//   OUT_RDY  256 arrivals: the outgoing tile is complete in A           (compute -> warp 8: store it)
//   B_FREE   256 arrivals: B has been read into registers               (compute -> warp 9: refill it)
// SDR_K1WIDE_MINB / SDR_K1W_MAXNREG: measurement builds (one CTA per SM with more registers).
#ifndef SDR_K1WIDE_MINB
#define SDR_K1WIDE_MINB 2
#endif
#ifdef SDR_K1W_MAXNREG
__global__ void __maxnreg__(SDR_K1W_MAXNREG) k1_wide_kernel(const WideArgs wa) {
#else
__global__ void __launch_bounds__(K1W_THREADS, SDR_K1WIDE_MINB) k1_wide_kernel(const WideArgs wa) {
#endif
    constexpr int N = 65536, N1 = 256;
    const K1Args &a = wa.a;
    extern __shared__ __align__(1024) unsigned char wd_smem[];
    unsigned char *A = wd_smem + K1W_OFF_A;
    const float2 *B = reinterpret_cast<const float2 *>(wd_smem + K1W_OFF_B);
    float2 *S = reinterpret_cast<float2 *>(wd_smem + K1W_OFF_S);
    float2 *TW = reinterpret_cast<float2 *>(wd_smem + K1W_OFF_TW);  // [15][16] W_256^(hl k)
    float2 *TQ = reinterpret_cast<float2 *>(wd_smem + K1W_OFF_TQ);  // [16 columns][17] W_N^(16 c q)
    uint64_t *FULL_A = reinterpret_cast<uint64_t *>(wd_smem + K1W_OFF_MISC);
    uint64_t *FULL_B = FULL_A + 1, *OUT_RDY = FULL_A + 2, *B_FREE = FULL_A + 3;

    const int team = blockIdx.x / K1W_TEAM, rank = blockIdx.x % K1W_TEAM, n_teams = gridDim.x / K1W_TEAM;
    const int tid = threadIdx.x;
    const int c0 = 16 * rank, r0 = 16 * rank;  // my columns when producing, my rows when consuming
    const int D = wa.lookahead, R = wa.ring;

    if (tid < 240) TW[tid] = __ldg(&wa.tw256[((tid & 15) * ((tid >> 4) + 1)) & 255]);
    // W_N^(c k) for k = hl + 16 q is W_N^(c hl) (one register pair per thread) times W_N^(16 c q) (this table):
    // sixteen resident step twiddles per thread would not fit next to the cumulation at 2 CTAs per SM
    if (tid < 256) TQ[(tid >> 4) * 17 + (tid & 15)] = __ldg(&wa.tw_step[(size_t)(c0 + (tid >> 4)) * 256 + 16 * (tid & 15)]);
    if (tid == 0) {
        mbar_init(FULL_A, 1);
        mbar_init(FULL_B, 1);
        mbar_init(OUT_RDY, 256);
        mbar_init(B_FREE, 256);
        fence_mbar_init();
    }
    __syncthreads();

    if (tid >= 256) {
        // ================================ DMA warps ================================
        // Warp 8 (lane 0) moves the tiles this CTA PRODUCES: it stores the outgoing tile, refills A as soon as the store has
        // read it, and publishes the step once the store has completed.  Warp 9 (lane 0) fetches the tiles this CTA
        // CONSUMES: it waits until B has been read, until all sixteen producers have published the next step, and requests
        // the row tile.  They are two threads because each chain is long and serial (measured on B200, SM cycles: tile read
        // out of A 2200, release of the counter 1500; acquire poll of the counter 1900, tile landing 1800): one thread
        // running both requested the row tile 6000 cycles after the produce phase ended, exactly when it was needed.
        if (tid == 256) {
            tensormap_acquire(wa.tmp_map);
            const uint64_t policy = l2_evict_first_policy();
            WideIter jp, ja;
            jp.start(a.segs, a.n_segs, team);
            ja = jp;
            uint32_t ph_out = 0;
            auto issue_a = [&]() {  // IQ column tile of the next step to produce (A is free)
                if (ja.nb == 0) return;
                if (ja.blk == 0) tensormap_acquire(&wa.seg_maps[ja.seg]);  // first use of this segment's map
                mbar_expect_tx(FULL_A, K1W_TILE_BYTES);
                tma_load_tile_3d(A, &wa.seg_maps[ja.seg], 2 * c0, 0, ja.blk, FULL_A, policy);
                ja.next(a.segs, a.n_segs, n_teams);
            };
            issue_a();
            while (jp.nb != 0) {  // the compute warps finished producing step jp
                mbar_wait(OUT_RDY, ph_out);
                ph_out ^= 1u;
                K1W_TR(jp.step, 10);
                fence_proxy_async_all();  // consumers' discards of the slot's old lines (generic proxy) before this async-proxy write
                tma_store_tile_3d(wa.tmp_map, 2 * c0, 0, team * R + jp.step % R, A);
                bulk_commit();
                bulk_wait_read();  // the tile has been read out of A ...
                K1W_TR(jp.step, 11);
                issue_a();         // ... which takes the next column tile at once
                bulk_wait_all();   // the store has completed: publish the step
                K1W_TR(jp.step, 12);
                red_release_gpu_add(&wa.ready[a.segs[jp.seg].block_out + jp.blk], 1);
                K1W_TR(jp.step, 13);
                jp.next(a.segs, a.n_segs, n_teams);
            }
        } else if (tid == 288) {
            WideIter jb;
            jb.start(a.segs, a.n_segs, team);
            uint32_t ph_free = 0;
            bool first = true;
            while (jb.nb != 0) {  // row tile of the next step to consume, once all sixteen column tiles are published
                if (!first) {     // B has been read into registers
                    mbar_wait(B_FREE, ph_free);
                    ph_free ^= 1u;
                    K1W_TR(jb.step - 1, 14);
                }
                first = false;
                const int ob = a.segs[jb.seg].block_out + jb.blk;
                int spins = 0;
                while (ld_acquire_gpu(&wa.ready[ob]) < K1W_TEAM) {
                    ++spins;
                    if (spins > K1W_MAX_SPIN || ((spins & 1023) == 0 && ld_acquire_gpu(wa.err) != 0)) {  // give up; once one wait failed, all do
                        atomicExch(wa.err, 1);
                        break;
                    }
                    __nanosleep(64);
                }
                fence_proxy_async_all();  // the acquire above orders the async-proxy read below after the producers' stores
                K1W_TR(jb.step, 15);
                mbar_expect_tx(FULL_B, K1W_TILE_BYTES);
                tma_load_1d(wd_smem + K1W_OFF_B, wa.tmp + ((size_t)(team * R + jb.step % R) * N + (size_t)r0 * 256), K1W_TILE_BYTES, FULL_B);
                jb.next(a.segs, a.n_segs, n_teams);
            }
        }
        return;
    }

    // ================================ compute warps ================================
    const int hl = tid & 15, f = tid >> 4, lane = tid & 31, warp = tid >> 5;  // consume: half-warp f = row r0 + f
    // PRODUCE runs the same two transforms per warp (columns 2 warp, 2 warp + 1) with another lane assignment: lanes
    // 0-7 / 8-15 of each 16-lane half are elements hp .. hp + 7 of the even / odd column (hp = 0 or 8).  A 64-bit
    // shared-memory access is served per 16 lanes; in the 128B-swizzled tile the elements hl and hl + 8 of one column
    // share their banks (two wavefronts per half), whereas elements hp .. hp + 7 of two neighbouring columns cover the
    // 128-byte line exactly once.  The transposes of these transforms use the column pitch K1W_PITCH_P (== 8 mod 16), which
    // keeps them conflict-free under this assignment (the even / odd column's eight lanes take slots s .. s + 7 / s + 8 ..
    // s + 15 mod 16).
    const int hlp = (lane & 7) | ((lane >> 4) << 3), fp = 2 * warp + ((lane >> 3) & 1);
    // byte offset of element (row hlp + 16 j, column fp) in a 128B-swizzled [256][16] tile, minus j * 2048
    const int swz_off = hlp * 128 + ((((fp >> 1) ^ (hlp & 7))) << 4) + ((fp & 1) << 3);
    const float2 tw_base = __ldg(&wa.tw_step[(size_t)(c0 + fp) * 256 + hlp]);  // W_N^(c hl)
    // the fifteen W256^(hl k) of the half-warp transform are re-read from shared memory per transform
    auto load_hw_twiddle = [&](HwTwiddle &t, int l16) {
#pragma unroll
        for (int k = 1; k < 16; k++) t.w[k - 1] = TW[(k - 1) * 16 + l16];
    };

    WideIter ic, ip;
    ic.start(a.segs, a.n_segs, team);
    ip = ic;
    uint32_t phase_a = 0, phase_b = 0;

    // ---------------- produce: column tile of step ip -> outgoing tile in A ----------------
    auto produce = [&]() {
        if (tid == 0) K1W_TR(ip.step, 0);
        mbar_wait(FULL_A, phase_a);
        phase_a ^= 1u;
        if (tid == 0) K1W_TR(ip.step, 1);
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int n1 = (q & 3) * 4 + (q >> 2);
            v[n1] = *reinterpret_cast<const float2 *>(A + swz_off + n1 * 2048);  // x[(16 n1 + hlp) * 256 + c0 + fp]
        }
        if (a.window) {
#pragma unroll
            for (int n1 = 0; n1 < 16; n1++) {
                const float w = __ldg(&a.window[(16 * n1 + hlp) * 256 + c0 + fp]);
                v[n1] = __fmul2_rn(v[n1], make_float2(w, w));
            }
        }
        compute_sync();  // the tile is in registers: A can take the outgoing tile; S is free (consume's plane reads are done)
        if (tid == 0) K1W_TR(ip.step, 2);
        {
            HwTwiddle t;
            load_hw_twiddle(t, hlp);
            fft256_halfwarp_regs(v, S + fp * K1W_PITCH_P, t, hlp);
        }
#pragma unroll
        for (int p = 0; p < 16; p++) {  // Z[k][c0 + fp], k = hlp + 16 q, swizzled like the TMA box
            const int q = OutIdx<16>::of(p);
            *reinterpret_cast<float2 *>(A + swz_off + q * 2048) = cmul(cmul(v[p], tw_base), TQ[fp * 17 + q]);
        }
        fence_proxy_async();  // generic-proxy writes of the tile before the async-proxy (TMA) read
        mbar_arrive(OUT_RDY);
        if (tid == 0) K1W_TR(ip.step, 3);
        ip.next(a.segs, a.n_segs, n_teams);
    };

    for (int d = 0; d + 1 < D; d++) {  // prologue: D - 1 steps are produced before anything is consumed
        if (ip.nb == 0) break;
        produce();
    }

    float cum[16];
    int e = 0, ws = 1, n_win = 9, L = 0, nf_p = 0, nf_lo = 0, nf_hi = 0, nf_rot = 0, seg_block_out = 0;
    // S belongs to the warps in private slices of 2 * K1W_PITCH_P complex slots (no barrier separates one warp's produce
    // transposes from another warp's consume transposes): produce puts its two columns at pitch K1W_PITCH_P, consume at
    // pitch HW_PITCH inside the same slice.  The |X|^2 / dB planes of row j start at word plane_of(j) of S, inside the
    // row's own consume column (k1_large.cuh: plane_skew)
    auto col_of = [](int j) { return (j >> 1) * (2 * K1W_PITCH_P) + (j & 1) * HW_PITCH; };  // complex slots
    auto plane_of = [&](int j) { return 2 * col_of(j) + plane_skew(2 * col_of(j), j); };    // words
    const int *lbins = nullptr;

    while (ic.nb != 0) {
        // produce step ic + D - 1 first: its store then has the whole consume phase to complete and be published
        if (ip.nb != 0) produce();
        if (tid == 0) K1W_TR(ic.step, 20);
        if (ic.blk == 0) {  // a new segment: window geometry, listeners, cumulation registers
            const Segment sg = a.segs[ic.seg];
            seg_block_out = sg.block_out;
            const WorkParams wp = a.works[sg.work];
            e = wp.edge_width;
            ws = nf_window_size(N, e);
            n_win = nf_window_count(N, e);
            L = wp.n_listeners;
            lbins = a.listener_bins + wp.listener_off;
            // noise floor: thread (window w = 2 warp + lane/16, row j = lane % 16) of warps 0-4; the window's bins in row
            // k1 = r0 + j are the positions [first(k1), last(k1)) (bin kk = k1 + 256 p).  All sixteen rows of a window are
            // read from the position nf_p of the LAST row (the smallest first()), so that their banks stay plane_of's; the
            // row's own range is [nf_lo, nf_hi) relative to it, nf_lo in {0, 1} (k1_large.cuh: nf_row_share)
            nf_p = nf_lo = nf_hi = nf_rot = 0;
            if (warp < 5) {
                auto first = [&](int w, int k1) { const int lo = e + w * ws; return lo < k1 ? 0 : (lo - k1 + 255) >> 8; };
                const int w = 2 * warp + (lane >> 4), k1 = r0 + (lane & 15);
                const int hi = e + (w + 1) * ws;
                nf_p = first(w, r0 + 15);
                nf_lo = first(w, k1) - nf_p;
                const int p1 = hi <= k1 ? 0 : min(256, (hi - k1 + 255) >> 8);
                nf_hi = max(p1 - nf_p, 0);
                // the window of the upper half-warp walks rotated when its start has the parity of the lower one's
                nf_rot = (lane >> 4) & ~(first(2 * warp + 1, r0 + 15) - first(2 * warp, r0 + 15)) & 1;
            }
#pragma unroll
            for (int p = 0; p < 16; p++) {
                const int kk = (r0 + f) + N1 * ((hl + 16 * OutIdx<16>::of(p) + 128) & 255);
                cum[p] = sg.state_in >= 0 ? a.cum_state[(size_t)sg.state_in * N + kk] : 0.f;
            }
        }
        const int ob = seg_block_out + ic.blk;

        // ---------------- consume: row tile of step ic ----------------
        if (tid == 0) K1W_TR(ic.step, 4);
        mbar_wait(FULL_B, phase_b);
        phase_b ^= 1u;
        if (tid == 0) K1W_TR(ic.step, 5);
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int n1 = (q & 3) * 4 + (q >> 2);
            v[n1] = B[f * 256 + 16 * n1 + hl];  // Z[r0 + f][16 n1 + hl]
        }
        // The tile's lines in the ring are dead until a later step overwrites them: drop them from L2 instead of
        // letting their eviction write 8 N bytes per block back to HBM (one 128-byte line per thread)
        if (wa.discard)
            discard_l2_line(reinterpret_cast<const unsigned char *>(wa.tmp + ((size_t)(team * R + ic.step % R) * N + (size_t)r0 * 256)) + tid * 128);
        mbar_arrive(B_FREE);  // B is in registers: warp 9 refills it
        float2 *col = S + col_of(f);
        {
            HwTwiddle t;
            load_hw_twiddle(t, hl);
            fft256_halfwarp_regs(v, col, t, hl);
        }
        __syncwarp();
        // X[(r0 + f) + 256 k2], k2 = hl + 16*OutIdx<16>(p).  The row's (dead) transpose storage takes two float planes in
        // fftshifted order (dsp/fft.go:54-57): |X|^2 at prow[k2s], dB at prow[256 + k2s], k2s = (k2 + 128) & 255, i.e.
        // bin kk = (r0 + f) + 256 k2s
        float *prow = reinterpret_cast<float *>(S) + plane_of(f);  // inside the row's own column
        const bool need_db = L > 0 || a.dbg_psd != nullptr;  // uniform: a peak scan without listeners (config 5) skips the stores
#pragma unroll
        for (int p = 0; p < 16; p++) {
            const float psd = fmaf(v[p].x, v[p].x, v[p].y * v[p].y);                                      // dsp/fft.go:71-73
            const float db = __fadd_rn(fmaf(3.01029995663981195f, fast_log2(psd), wa.db_offset), 120.0f);  // rx/receiver.go:376-378
            cum[p] = __fadd_rn(cum[p], db);                                                                // rx/receiver.go:404-406
            const int k2s = hl + ((16 * OutIdx<16>::of(p) + 128) & 255);
            prow[k2s] = psd;
            if (need_db) prow[256 + k2s] = db;  // only the taps (and the parity store) read the dB plane
        }
        if (tid == 0) K1W_TR(ic.step, 6);
        compute_sync();
        if (tid == 0) K1W_TR(ic.step, 7);
        const float *Sf = reinterpret_cast<const float *>(S);  // row j of the tile: Sf[plane_of(j) + (0 | 256) + k2s]
        if (a.dbg_psd) {  // parity / scope only: half-warp = 16 consecutive bins
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int k2s = f + 16 * i;
                const int kk = (r0 + hl) + N1 * k2s;
                a.dbg_psd[(size_t)ob * N + kk] = Sf[plane_of(hl) + k2s];
                a.dbg_spectrum[(size_t)ob * N + kk] = Sf[plane_of(hl) + 256 + k2s];
            }
        }
        // dsp.FindNoiseFloor (dsp/fft.go:215-252), window sums over this CTA's bins: thread (window w, row j) sums the <= 27
        // positions of window w that live in row j (contiguous floats, float32 inside the share), a float64 half-warp
        // reduction over the sixteen rows gives this CTA's (sum x, sum x^2) of the window
        if (warp < 5) {
            const int w = 2 * warp + (lane >> 4);
            float s1, s2;
            nf_row_share<K1W_NF_M>(Sf, plane_of(lane & 15) + nf_p, nf_lo, nf_hi, nf_rot, max(ws >> 8, 2), s1, s2);
            if (tid == 0) K1W_TR(ic.step, 16);
            double d1 = (double)s1, d2 = (double)s2;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                d1 += __shfl_xor_sync(0xffffffffu, d1, o);
                d2 += __shfl_xor_sync(0xffffffffu, d2, o);
            }
            if (tid == 0) K1W_TR(ic.step, 17);
            if ((lane & 15) == 0) wa.nf_part[((size_t)ob * K1W_TEAM + rank) * 10 + w] = make_double2(d1, d2);
        }
        if (tid == 0) K1W_TR(ic.step, 18);
        if (tid < n_win) {  // x_to = psd[e + (w+1)*ws] (dsp/fft.go:238-243) if this CTA owns that bin
            const int kk = e + (tid + 1) * ws;
            const int k1 = kk & (N1 - 1);
            if (k1 >= r0 && k1 < r0 + 16) wa.xto[(size_t)ob * 10 + tid] = Sf[plane_of(k1 - r0) + (kk >> 8)];
        }
        if (rank == 0 && tid == 0) wa.nf_edge[ob] = e;
        for (int l = tid; l < L; l += 256) {  // listener taps (rx/receiver.go:393) on bins this CTA owns
            const int kk = __ldg(&lbins[l]);
            const int k1 = kk & (N1 - 1);
            if (k1 >= r0 && k1 < r0 + 16) a.taps[(size_t)ob * a.tap_stride + l] = Sf[plane_of(k1 - r0) + 256 + (kk >> 8)];
        }
        if (tid == 0) K1W_TR(ic.step, 19);
        if (ic.blk == ic.nb - 1) {  // end of the segment: flush or save the cumulation
            const Segment sg = a.segs[ic.seg];
            float *dst = (sg.flush_idx >= 0) ? a.flush_cum + (size_t)sg.flush_idx * N : a.cum_state + (size_t)sg.state_out * N;
#pragma unroll
            for (int p = 0; p < 16; p++) dst[(r0 + f) + N1 * ((hl + 16 * OutIdx<16>::of(p) + 128) & 255)] = cum[p];
        }
        // The planes must be read before the next transform's transposes overwrite S.  When a produce follows, its own
        // barrier (after the column-tile reads, before the transform) already orders that, and the warps without
        // window-sum work start reading the next column tile meanwhile; only a consume that follows directly needs one.
        if (ip.nb == 0) compute_sync();
        if (tid == 0) K1W_TR(ic.step, 8);
        if (tid == 255) K1W_TR(ic.step, 9);
        ic.next(a.segs, a.n_segs, n_teams);
    }
}

}  // namespace sdr
