// engine.cu -- libsdrgpu.so: engine, stream state, batching and the C ABI of include/sdrgpu.h.
//
// The engine is the device-side half of rx.Receiver.run (rx/receiver.go:336-464): it owns, per
// stream, what the reference keeps in run()'s locals -- the float32 cumulation vector and its
// counter (:340,:347) and the two 60-block rolling means (:343-344) -- and executes the frame
// iteration (:364-461) for whole batches of queued frames with two kernels:
//   K1 k1_spectral_kernel  FFT + |X|^2 + dB + noise floor + listener taps + cumulation
//   K2 k2_thresholds / k2_keys / k2_peaks   rolling means + thresholds, key states (debounced, packed), FindPeaks at flushes
// Three CUDA streams (H2D, compute, D2H) with events give copy/compute overlap across in-flight
// slots; kernels of successive batches stay ordered on the compute stream, which is what keeps
// the per-stream state sequential.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>
#include <optional>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only: ranges cost nothing unless a profiler injects itself

#include "../../include/sdrgpu.h"
#include "k1_spectral.cuh"
#include "k2_post.cuh"
#include "k1_large.cuh"
#include "k1_mid4k.cuh"
#include "k1_mid8k.cuh"
#include "k1_warp.cuh"
#include "k1_wide.cuh"

using namespace sdr;

namespace {

struct NvtxRange {  // sdr_submit / K1 / K2 / D2H show up as named ranges in Nsight Systems and ncu --nvtx
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

std::string g_create_error;

struct StreamInfo {
    bool open = false;
    int sample_rate = 0;
    int cum_count = 0;  // cumulationCount (rx/receiver.go:347)
    int state_row = 0;  // which of the stream's two cum_state rows holds the open window's partial cumulation
};

struct Slot {
    bool busy = false;
    bool collected = false;
    sdr_ticket ticket = 0;
    int flags = 0;
    // descriptors: one pinned staging block + one device block
    unsigned char *h_desc = nullptr, *d_desc = nullptr;
    size_t desc_bytes = 0;
    // device results
    float *d_psd_floor = nullptr;
    double *d_variance = nullptr;
    float *d_thresholds = nullptr;
    float *d_taps = nullptr;
    uint8_t *d_keys = nullptr;
    uint32_t *d_key_bits = nullptr, *h_key_bits = nullptr;
    int *d_flush_block = nullptr, *d_flush_n_peaks = nullptr;
    sdr_peak *d_flush_peaks = nullptr;
    float *d_flush_cum = nullptr;
    float *d_spectrum = nullptr, *d_psd = nullptr;  // lazy (SDR_WANT_SPECTRUM)
    float *d_iq = nullptr;                          // lazy (host inputs)
    float2 *d_tmp = nullptr;                        // large-block path: four-step intermediate (of one round)
    float *d_spec_round = nullptr;                  // register-resident large path: dB spectrum of one round
    double2 *d_nf_part = nullptr;                   //   per-CTA noise-window sums
    float *d_xto = nullptr;
    int *d_nf_edge = nullptr;
    // k1_wide: intermediate ring, per-step ready counters, the ring's tensor map, error flag
    float2 *d_wide_tmp = nullptr;
    int *d_wide_ready = nullptr, *d_wide_err = nullptr, *h_wide_err = nullptr;
    CUtensorMap *d_wide_map = nullptr;
    // pinned host mirrors
    float *h_psd_floor = nullptr;
    double *h_variance = nullptr;
    float *h_thresholds = nullptr;
    float *h_taps = nullptr;
    uint8_t *h_keys = nullptr;
    int *h_flush_block = nullptr, *h_flush_n_peaks = nullptr;
    sdr_peak *h_flush_peaks = nullptr;
    float *h_flush_cum = nullptr;
    float *h_spectrum = nullptr, *h_psd = nullptr;  // lazy
    std::vector<int> work_block_offset, work_flush_offset;
    int n_works = 0, n_blocks = 0, n_flushes = 0, launches = 0;
    cudaEvent_t ev_desc = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_k0 = nullptr, ev_km = nullptr, ev_k2s = nullptr, ev_k1 = nullptr, ev_done = nullptr;
    cudaEvent_t ev_thr = nullptr, ev_peaks = nullptr;
};

}  // namespace

struct sdr_engine {
    // the C ABI is safe for concurrent callers (one goroutine per rx.Receiver, rx/receiver.go:145,336): every entry
    // point takes this lock; a blocking sdr_collect drops it while it waits on the ticket's event
    std::mutex mu;
    sdr_engine_config cfg{};
    int N = 0;
    int tap_stride = 4;
    int max_segs = 0, max_flushes = 0;
    int sm_count = 148;
    std::string err;
    cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    // descriptor uploads always go through an internal stream (also when the caller supplied the compute stream): a
    // slot's descriptor block is only rewritten after its ticket was released, so the copy of batch i+1 overlaps the
    // kernels of batch i instead of sitting between them on the compute stream
    cudaStream_t s_desc = nullptr;
    // K2 (thresholds, keys, peaks) can run on its own stream behind an event so that batch i's K2 overlaps batch i+1's K1
    // (K1 never reads what K2 writes; K2s stay ordered among themselves, which keeps the per-stream rolling means
    // sequential).  Opt-in with SDR_K2_OVERLAP=1: see the measurement at the stream's creation.
    cudaStream_t s_post = nullptr;
    cudaStream_t s_aux = nullptr;  // k2_peaks runs here, next to k2_keys on the post stream (SDR_K2_AUX=0: back to back)
    bool own_post = false;
    cudaEvent_t ev_post = nullptr;  // last K2 (sdr_engine_fence)
    bool own_streams = true;
    float2 *d_tw1 = nullptr, *d_tw2 = nullptr;
    // large-block path (N >= 8192): four-step split N = n1 * n2
    bool large = false;
    LargeGeom lg{};
    float2 *d_tw_sub1 = nullptr, *d_tw_sub2 = nullptr;  // W_N1^m, W_N2^m
    float2 *d_tw_step = nullptr;  // [k1][c] = W_N^(c k1): step-1 twiddles of the register-resident kernels
    int round_blocks = 0;         // register-resident large path: blocks per L2-resident round (SDR_LARGE_ROUND_MB)
    // N = 65536: single-pass persistent kernel (k1_wide.cuh): teams of 16 CTAs, the four-step intermediate in an
    // L2-resident ring.  SDR_K1_WIDE=0 disables it (two-kernel path), =force takes it for every launch
    int k1_wide = 0;          // 0 off, 1 when the launch has enough segments, 2 always
    int k1w_max_teams = 0;    // co-resident CTAs / 16
    int k1w_lookahead = 3, k1w_ring = 6;
    bool k1w_discard = true;  // SDR_K1_WIDE_DISCARD=0: leave consumed ring tiles to L2's write-back
    PFN_cuTensorMapEncodeTiled_v12000 encode_tiled = nullptr;
    // N = 512: warp-per-block kernel (k1_warp.cuh), SDR_K1_WARP=0 selects the three-pass kernel
    float2 *d_tw512 = nullptr, *d_tw256w = nullptr;
    bool k1_warp = false;
    int k1w_grid_cap = 0;
    // N = 4096 / 8192: step twiddles [k1][c] = W_N^(c k1) and W_256^m of the fused single-pass kernels
    float2 *d_tw_mid = nullptr, *d_tw256m = nullptr;
    bool k1_mid4k = false;    // N = 4096: TMA-staged k1_mid4k_kernel (SDR_K1_MID4K=0: the three-pass kernel)
    int k1m4_grid_cap = 0;
    // N = 8192: TMA-staged 512-thread kernel (k1_mid8k.cuh), one CTA per SM.  SDR_K1_MID8K=0 disables it (the launch then
    // takes the two-kernel path), =force takes it for every launch, whatever the segment count
    int k1_mid8k = 0;  // 0 off, 1 when the launch has enough segments, 2 always
    int k1m8_stages = 2;
    float *d_window = nullptr;
    float *d_cum_state = nullptr;
    RollingState *d_rolling = nullptr;
    DebounceState *d_deb = nullptr;  // [max_streams][tap_stride] BoolDebouncer state of every listener position
    int key_words = 1;
    std::vector<StreamInfo> streams;
    std::vector<Slot> slots;
    sdr_ticket next_ticket = 1;
    int64_t launches = 0;
    const char *last_kernel = "";  // spectral kernel of the most recent submit (sdr_engine_last_kernel)
    int k1_grid_cap = 0;  // resident CTAs of K1 on this device
    bool k1_tw2r = true;  // kernel variant: pass-2 twiddles in registers (SDR_K1_TW2R=0 selects the smem-table variant)
    // cache of choose_nf_map results, indexed by edge width (first byte 0xff = not computed)
    std::vector<unsigned char> nf_map_cache;
    // scratch for the dsp single calls
    float *d_scratch = nullptr;
    size_t scratch_bytes = 0;
};

namespace {

#define CK(e, call)                                                                                  \
    do {                                                                                             \
        cudaError_t _st = (call);                                                                    \
        if (_st != cudaSuccess) {                                                                    \
            (e)->err = std::string(#call) + ": " + cudaGetErrorString(_st);                          \
            return SDR_ECUDA;                                                                        \
        }                                                                                            \
    } while (0)

// dsp.FindNoiseFloor (dsp/fft.go:215-252) closes a window whenever `count == ws` on entry, for every i < N-e: with
// ws = (N-2e)/10 >= 9 the remainder (N-2e) - 10*ws <= 9 never completes an 11th window, which is what the kernels
// assume (9 or 10 evaluated windows).  Narrower windows (ws < 9: the reference would evaluate up to 19 of them, and
// 0/0 for ws == 0) are rejected.
bool nf_edge_supported(int n, int e) { return e >= 0 && nf_window_size(n, e) >= 9; }

bool supported_fused_n(int n) { return n == 512 || n == 1024 || n == 2048 || n == 4096; }

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// one work whose edge width the fused noise-floor code does not cover (fewer than 9 bins per window, or no window at
// all): dsp.FindNoiseFloor is replayed literally, one thread per block, on the stored PSD
struct ExactNf {
    int block_out, n_blocks, edge_width, pad;
};

// descriptor block layout (same on host and device)
struct DescLayout {
    size_t segs, works, post, lbins, lflags, block_seg, exact, segmaps, total;
};
DescLayout desc_layout(const sdr_engine *e) {
    DescLayout l;
    size_t off = 0;
    l.segs = off;
    off = align_up(off + sizeof(Segment) * (size_t)e->max_segs, 256);
    l.works = off;
    off = align_up(off + sizeof(WorkParams) * (size_t)e->cfg.max_streams, 256);
    l.post = off;
    off = align_up(off + sizeof(PostWork) * (size_t)e->cfg.max_streams, 256);
    l.lbins = off;
    off = align_up(off + sizeof(int) * (size_t)e->cfg.max_streams * (size_t)(e->cfg.max_listeners > 0 ? e->cfg.max_listeners : 1), 256);
    l.lflags = off;
    off = align_up(off + (size_t)e->cfg.max_streams * (size_t)(e->cfg.max_listeners > 0 ? e->cfg.max_listeners : 1), 256);
    l.block_seg = off;
    if (e->large) off = align_up(off + sizeof(int) * (size_t)e->cfg.max_blocks_per_batch, 256);
    l.exact = off;
    off = align_up(off + sizeof(ExactNf) * (size_t)e->cfg.max_streams, 256);
    l.segmaps = off;
    if (e->k1_wide) off = align_up(off + sizeof(CUtensorMap) * (size_t)e->max_segs, 256);
    l.total = off;
    return l;
}

template <int N>
const void *k1_fn(bool dbg, bool win, bool tw2r, bool i16 = false) {
    if (i16) return dbg ? (const void *)k1_spectral_kernel<N, true, false, true, true> : (const void *)k1_spectral_kernel<N, false, false, true, true>;
    if (tw2r) {
        if (dbg) return win ? (const void *)k1_spectral_kernel<N, true, true, true> : (const void *)k1_spectral_kernel<N, true, false, true>;
        return win ? (const void *)k1_spectral_kernel<N, false, true, true> : (const void *)k1_spectral_kernel<N, false, false, true>;
    }
    if (dbg) return win ? (const void *)k1_spectral_kernel<N, true, true, false> : (const void *)k1_spectral_kernel<N, true, false, false>;
    return win ? (const void *)k1_spectral_kernel<N, false, true, false> : (const void *)k1_spectral_kernel<N, false, false, false>;
}

template <int N>
int k1_occupancy(bool dbg, bool win, bool tw2r) {
    using Gm = K1Geom<N>;
    int occ = 0;
    const void *fn = k1_fn<N>(dbg, win, tw2r);
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, Gm::SMEM_BYTES);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, Gm::CTA_THREADS, Gm::SMEM_BYTES);
    if (!win) {  // the KiwiSDR int16 variants exist for the unwindowed configuration only
        cudaFuncSetAttribute(k1_fn<N>(dbg, false, true, true), cudaFuncAttributeMaxDynamicSharedMemorySize, Gm::SMEM_BYTES);
    }
    return occ;
}

template <int N>
cudaError_t launch_k1_n(const sdr_engine *e, const K1Args &a, bool dbg, cudaStream_t st, bool i16) {
    using Gm = K1Geom<N>;
    const bool win = a.window != nullptr;
    // every group walks the segment list with stride grid*G: pick the grid so that all groups get the
    // same number of rounds (no straggler CTA running alone on its SM at the end)
    const int need = (a.n_segs + Gm::G - 1) / Gm::G;
    int grid = need;
    if (need > e->k1_grid_cap) {
        const int rounds = (need + e->k1_grid_cap - 1) / e->k1_grid_cap;
        grid = (need + rounds - 1) / rounds;
    }
    if (grid < 1) grid = 1;
    K1Args args = a;
    void *params[] = {&args};
    return cudaLaunchKernel(k1_fn<N>(dbg, win, e->k1_tw2r, i16), dim3(grid), dim3(Gm::CTA_THREADS), params, Gm::SMEM_BYTES, st);
}

// register-resident large path (k1_large.cuh): per round of consecutive blocks step 1, step 2 + epilogue, cumulation;
// once per batch the noise-floor finish
struct LargeRound {
    int blk0, n_blocks, seg0, n_segs;
};
struct LargeFastBufs {
    float2 *tmp;
    float *spec_round;
    double2 *nf_part;
    float *xto;
    int *nf_edge;
};
cudaError_t launch_large_fast(const sdr_engine *e, const K1Args &a, const LargeFastBufs &lb, const int *d_block_seg,
                              const std::vector<LargeRound> &rounds, int n_blocks, bool dbg, cudaStream_t st, int *n_launches) {
    const int N = e->N;
    FastStepArgs fa{};
    fa.tmp = lb.tmp;
    fa.spec_round = lb.spec_round;
    fa.spectrum = dbg ? a.dbg_spectrum : nullptr;
    fa.psd = dbg ? a.dbg_psd : nullptr;
    fa.tw256 = e->d_tw_sub2;
    fa.tw_n1 = e->d_tw_sub1;
    fa.tw_step = e->d_tw_step;
    fa.window = e->d_window;
    fa.segs = a.segs;
    fa.block_seg = d_block_seg;
    fa.works = a.works;
    fa.listener_bins = a.listener_bins;
    fa.nf_part = lb.nf_part;
    fa.xto = lb.xto;
    fa.nf_edge = lb.nf_edge;
    fa.taps = a.taps;
    fa.tap_stride = a.tap_stride;
    fa.n = N;
    fa.n1 = e->lg.n1;
    fa.n2 = e->lg.n2;
    fa.db_offset = (float)(10.0 * log10(20.0 / ((double)N * (double)N)));
    const size_t smem = (size_t)16 * HW_PITCH * sizeof(float2);
    int launches = 0;
    for (const LargeRound &r : rounds) {
        fa.blk0 = r.blk0;
        if (e->lg.n1 == 256) fast_cols256_kernel<<<dim3(e->lg.n2 / 16, r.n_blocks), 256, smem, st>>>(fa);
        else if (e->lg.n1 == 128) fast_cols64_kernel<true><<<dim3(e->lg.n2 / 128, r.n_blocks), 256, 0, st>>>(fa);
        else if (e->lg.n1 == 64) fast_cols64_kernel<false><<<dim3(e->lg.n2 / 256, r.n_blocks), 256, 0, st>>>(fa);
        else fast_cols32_kernel<<<dim3(e->lg.n2 / 256, r.n_blocks), 256, 0, st>>>(fa);
        cudaError_t rc = cudaGetLastError();
        if (rc != cudaSuccess) return rc;
        // enough segments to fill the GPU several times over: rows + cumulation in one segment-sequential kernel
        if ((long long)r.n_segs * (e->lg.n1 / 16) >= 4LL * e->sm_count && !getenv("SDR_LARGE_NO_SEGROWS")) {
            FastRowsSegArgs sa{};
            sa.s = fa;
            sa.seg0 = r.seg0;
            sa.cum_state = a.cum_state;
            sa.flush_cum = a.flush_cum;
            fast_rows256_seg_kernel<<<dim3(e->lg.n1 / 16, r.n_segs), 256, smem, st>>>(sa);
            rc = cudaGetLastError();
            if (rc != cudaSuccess) return rc;
            launches += 2;
            continue;
        }
        fast_rows256_kernel<<<dim3(e->lg.n1 / 16, r.n_blocks), 256, smem, st>>>(fa);
        rc = cudaGetLastError();
        if (rc != cudaSuccess) return rc;
        RoundCumArgs ca{};
        ca.spec_round = lb.spec_round;
        ca.segs = a.segs + r.seg0;
        ca.cum_state = a.cum_state;
        ca.flush_cum = a.flush_cum;
        ca.blk0 = r.blk0;
        ca.n = N;
        large_round_cum_kernel<<<dim3(N / 256, r.n_segs), 256, 0, st>>>(ca);
        rc = cudaGetLastError();
        if (rc != cudaSuccess) return rc;
        launches += 3;
    }
    LargeFinishArgs fin{};
    fin.nf_part = lb.nf_part;
    fin.xto = lb.xto;
    fin.nf_edge = lb.nf_edge;
    fin.psd_floor = a.psd_floor;
    fin.variance = a.variance;
    fin.n_blocks = n_blocks;
    fin.n_cta = e->lg.n1 / 16;
    fin.n = N;
    large_nf_finish_kernel<<<(n_blocks + 3) / 4, 128, 0, st>>>(fin);
    if (n_launches) *n_launches = launches + 1;
    return cudaGetLastError();
}

// tensor map over [n_outer][256 rows][256 complex] float32 pairs with a {16 complex, 256 rows, 1} box and the 128-byte
// swizzle k1_wide.cuh assumes
bool encode_tile_map(const sdr_engine *e, CUtensorMap *m, const void *base, int n_outer) {
    const cuuint64_t dims[3] = {512, 256, (cuuint64_t)n_outer};
    const cuuint64_t strides[2] = {2048, 524288};
    const cuuint32_t box[3] = {32, 256, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return e->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// N = 65536 in one pass: teams of 16 CTAs walk the segments (k1_wide.cuh), then the noise-floor finish
struct WideBufs {
    const CUtensorMap *seg_maps, *tmp_map;
    float2 *tmp;
    int *ready, *err, *h_err;
};
#ifdef SDR_K1W_TRACE
static long long *g_k1w_trace = nullptr;  // measurement builds: clock stamps of the last k1_wide launch
constexpr size_t K1W_TRACE_BYTES = (size_t)32 * 512 * 32 * sizeof(long long);
#endif
cudaError_t launch_k1_wide(const sdr_engine *e, const K1Args &a, const LargeFastBufs &lb, const WideBufs &wb, int n_blocks, bool dbg,
                           cudaStream_t st) {
    WideArgs wa{};
    wa.a = a;
    if (!dbg) wa.a.dbg_psd = wa.a.dbg_spectrum = nullptr;
    wa.seg_maps = wb.seg_maps;
    wa.tmp_map = wb.tmp_map;
    wa.tmp = wb.tmp;
    wa.ready = wb.ready;
    wa.err = wb.err;
    wa.tw256 = e->d_tw_sub2;
    wa.tw_step = e->d_tw_step;
    wa.nf_part = lb.nf_part;
    wa.xto = lb.xto;
    wa.nf_edge = lb.nf_edge;
    wa.db_offset = (float)(10.0 * log10(20.0 / (65536.0 * 65536.0)));
    wa.lookahead = e->k1w_lookahead;
    wa.ring = e->k1w_ring;
    wa.discard = e->k1w_discard ? 1 : 0;
#ifdef SDR_K1W_TRACE
    if (!g_k1w_trace) cudaMalloc((void **)&g_k1w_trace, K1W_TRACE_BYTES);
    cudaMemsetAsync(g_k1w_trace, 0, K1W_TRACE_BYTES, st);
    wa.trace = g_k1w_trace;
#endif
    const int n_teams = a.n_segs < e->k1w_max_teams ? a.n_segs : e->k1w_max_teams;
    cudaError_t rc = cudaMemsetAsync(wb.ready, 0, (size_t)n_blocks * sizeof(int), st);
    if (rc != cudaSuccess) return rc;
    void *params[] = {&wa};
    // cooperative: every CTA of the launch is resident at the same time (the teams wait on one another's tiles)
    rc = cudaLaunchCooperativeKernel((const void *)k1_wide_kernel, dim3(K1W_TEAM * n_teams), dim3(K1W_THREADS), params, K1W_SMEM_BYTES, st);
    if (rc != cudaSuccess) return rc;
    rc = cudaMemcpyAsync(wb.h_err, wb.err, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (rc != cudaSuccess) return rc;
    LargeFinishArgs fin{};
    fin.nf_part = lb.nf_part;
    fin.xto = lb.xto;
    fin.nf_edge = lb.nf_edge;
    fin.psd_floor = a.psd_floor;
    fin.variance = a.variance;
    fin.n_blocks = n_blocks;
    fin.n_cta = K1W_TEAM;
    fin.n = 65536;
    large_nf_finish_kernel<<<(n_blocks + 3) / 4, 128, 0, st>>>(fin);
    return cudaGetLastError();
}

const void *k1w_fn(bool dbg, bool win, bool i16) {
    if (i16) return dbg ? (const void *)k1_warp_kernel<true, false, true> : (const void *)k1_warp_kernel<false, false, true>;
    if (dbg) return win ? (const void *)k1_warp_kernel<true, true, false> : (const void *)k1_warp_kernel<true, false, false>;
    return win ? (const void *)k1_warp_kernel<false, true, false> : (const void *)k1_warp_kernel<false, false, false>;
}
int k1w_grid_cap_for(bool win, int sm_count) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k1w_fn(false, win, false), 32 * K1WarpGeom::WARPS, K1WarpGeom::SMEM_BYTES);
    if (occ < 1) occ = 1;
    return occ * sm_count;
}
cudaError_t launch_k1_warp(const sdr_engine *e, const K1Args &a, bool dbg, cudaStream_t st, bool i16) {
    // one warp per segment, WARPS warps per CTA; whole rounds of the resident grid
    const int need = (a.n_segs + K1WarpGeom::WARPS - 1) / K1WarpGeom::WARPS;
    int grid = need;
    if (need > e->k1w_grid_cap) {
        const int rounds = (need + e->k1w_grid_cap - 1) / e->k1w_grid_cap;
        grid = (need + rounds - 1) / rounds;
    }
    if (grid < 1) grid = 1;
    K1Args args = a;
    const float2 *t512 = e->d_tw512, *t256 = e->d_tw256w;
    void *params[] = {&args, &t512, &t256};
    return cudaLaunchKernel(k1w_fn(dbg, a.window != nullptr, i16), dim3(grid), dim3(32 * K1WarpGeom::WARPS), params,
                            K1WarpGeom::SMEM_BYTES, st);
}

const void *k1m4_fn(bool dbg, bool win) {
    if (dbg) return win ? (const void *)k1_mid4k_kernel<2, true, true> : (const void *)k1_mid4k_kernel<2, true, false>;
    return win ? (const void *)k1_mid4k_kernel<2, false, true> : (const void *)k1_mid4k_kernel<2, false, false>;
}
int k1m4_grid_cap_for(int sm_count) {  // 0: the kernel cannot run here
    int occ = 0;
    for (int v = 0; v < 4; v++) {
        const void *fn = k1m4_fn((v & 1) != 0, (v & 2) != 0);
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, K1Mid4kGeom<2>::SMEM_BYTES) != cudaSuccess) return 0;
        if (v == 0 && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 256, K1Mid4kGeom<2>::SMEM_BYTES) != cudaSuccess) return 0;
    }
    return occ * sm_count;
}

int k1m8_smem(int stages) { return stages == 1 ? K1Mid8kGeom<1>::SMEM_BYTES : K1Mid8kGeom<2>::SMEM_BYTES; }
template <int NSTAGE>
const void *k1m8g_fn(bool dbg, bool win) {
    if (dbg) return win ? (const void *)k1_mid8k2_kernel<NSTAGE, true, true> : (const void *)k1_mid8k2_kernel<NSTAGE, true, false>;
    return win ? (const void *)k1_mid8k2_kernel<NSTAGE, false, true> : (const void *)k1_mid8k2_kernel<NSTAGE, false, false>;
}
const void *k1m8g_fn(int stages, bool dbg, bool win) { return stages == 1 ? k1m8g_fn<1>(dbg, win) : k1m8g_fn<2>(dbg, win); }

// nfb: per-block window-sum buffers of the slot (decoupled-group variant; the finish kernel follows)
cudaError_t launch_k1_mid8k(const sdr_engine *e, const K1Args &a, const Mid8kNf &nfb, int n_blocks, bool dbg, cudaStream_t st, int *n_launches) {
    // one CTA per SM walks the segment list with stride grid: whole rounds, no straggler
    int grid = a.n_segs;
    if (grid > e->sm_count) {
        const int rounds = (grid + e->sm_count - 1) / e->sm_count;
        grid = (a.n_segs + rounds - 1) / rounds;
    }
    if (grid < 1) grid = 1;
    K1Args args = a;
    const float2 *tws = e->d_tw_mid, *tw256 = e->d_tw256m;
    Mid8kNf nfa = nfb;
    void *params[] = {&args, &tws, &tw256, &nfa};
    cudaError_t rc = cudaLaunchKernel(k1m8g_fn(e->k1m8_stages, dbg, a.window != nullptr), dim3(grid), dim3(512), params, k1m8_smem(e->k1m8_stages), st);
    if (rc != cudaSuccess) return rc;
    LargeFinishArgs fin{};
    fin.nf_part = nfb.nf_part;
    fin.xto = nfb.xto;
    fin.nf_edge = nfb.nf_edge;
    fin.psd_floor = a.psd_floor;
    fin.variance = a.variance;
    fin.n_blocks = n_blocks;
    fin.n_cta = 2;
    fin.n = 8192;
    large_nf_finish_kernel<<<(n_blocks + 3) / 4, 128, 0, st>>>(fin);
    *n_launches = 2;
    return cudaGetLastError();
}

cudaError_t launch_k1_mid4k(const sdr_engine *e, const K1Args &a, bool dbg, cudaStream_t st) {
    // two CTAs per SM walk the segment list with stride grid: whole rounds, no straggler
    int grid = a.n_segs;
    if (grid > e->k1m4_grid_cap) {
        const int rounds = (grid + e->k1m4_grid_cap - 1) / e->k1m4_grid_cap;
        grid = (a.n_segs + rounds - 1) / rounds;
    }
    if (grid < 1) grid = 1;
    K1Args args = a;
    const float2 *tws = e->d_tw_mid, *tw256 = e->d_tw256m;
    void *params[] = {&args, &tws, &tw256};
    return cudaLaunchKernel(k1m4_fn(dbg, a.window != nullptr), dim3(grid), dim3(256), params, K1Mid4kGeom<2>::SMEM_BYTES, st);
}

// warp_ok: N = 512 and every work has noise windows of at least K1WarpGeom::MIN_WS bins
cudaError_t launch_k1(const sdr_engine *e, const K1Args &a, bool dbg, cudaStream_t st, bool i16 = false, bool warp_ok = true) {
    if (e->k1_mid4k && e->N == 4096 && !i16) return launch_k1_mid4k(e, a, dbg, st);
    if (e->k1_warp && warp_ok) return launch_k1_warp(e, a, dbg, st, i16);
    switch (e->N) {
        case 512: return launch_k1_n<512>(e, a, dbg, st, i16);
        case 1024: return launch_k1_n<1024>(e, a, dbg, st, i16);
        case 2048: return launch_k1_n<2048>(e, a, dbg, st, i16);
        case 4096: return launch_k1_n<4096>(e, a, dbg, st, i16);
    }
    return cudaErrorInvalidValue;
}

int k1_grid_cap_for(int n, bool win, bool tw2r, int sm_count) {
    int occ = 0;
    switch (n) {
        case 512: occ = k1_occupancy<512>(false, win, tw2r); k1_occupancy<512>(true, win, tw2r); break;
        case 1024: occ = k1_occupancy<1024>(false, win, tw2r); k1_occupancy<1024>(true, win, tw2r); break;
        case 2048: occ = k1_occupancy<2048>(false, win, tw2r); k1_occupancy<2048>(true, win, tw2r); break;
        case 4096: occ = k1_occupancy<4096>(false, win, tw2r); k1_occupancy<4096>(true, win, tw2r); break;
    }
    if (occ < 1) occ = 1;
    return occ * sm_count;
}

// Noise-floor window sums (k1_spectral_kernel, phase 1): thread group g of TPW consecutive threads sums window
// map[g] in shares of `per` contiguous bins.  Windows that share a warp can collide in shared-memory banks; this
// host-side search simulates the warp access pattern and keeps the window permutation with the fewest wavefronts.
template <int N>
void choose_nf_map_n(int e, unsigned char *map) {
    using Gm = K1Geom<N>;
    constexpr int T = Gm::T, TPW = Gm::TPW;
    const int ws = nf_window_size(N, e);
    const int per = ((ws + TPW - 1) / TPW) | 1;
    auto cost = [&](const unsigned char *m) {
        long wf = 0;
        for (int warp = 0; warp < (T + 31) / 32; warp++)
            for (int i = 0; i < per; i++) {
                int cnt[32] = {0}, mx = 0;
                for (int lane = 0; lane < 32; lane++) {
                    const int t = warp * 32 + lane;
                    if (t >= T || t >= 10 * TPW) continue;
                    const int w = m[t / TPW], part = t % TPW;
                    const int lo = e + w * ws + part * per;
                    int hi = lo + per;
                    if (hi > e + (w + 1) * ws) hi = e + (w + 1) * ws;
                    if (lo + i >= hi) continue;
                    if (++cnt[(lo + i) & 31] > mx) mx = cnt[(lo + i) & 31];
                }
                wf += mx;
            }
        return wf;
    };
    unsigned char cur[16], best[16];
    for (int i = 0; i < 16; i++) cur[i] = best[i] = (unsigned char)i;
    long best_cost = cost(best);
    uint64_t rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)(e * 2654435761u);
    for (int trial = 0; trial < 1500 && T >= 64; trial++) {
        for (int i = 9; i > 0; i--) {  // Fisher-Yates with xorshift
            rng ^= rng >> 12; rng ^= rng << 25; rng ^= rng >> 27;
            const int j = (int)((rng * 0x2545F4914F6CDD1Dull >> 33) % (uint64_t)(i + 1));
            const unsigned char tmp = cur[i]; cur[i] = cur[j]; cur[j] = tmp;
        }
        const long c = cost(cur);
        if (c < best_cost) {
            best_cost = c;
            memcpy(best, cur, 16);
        }
    }
    memcpy(map, best, 16);
}

void choose_nf_map(int n, int e, unsigned char *map) {
    switch (n) {
        case 512: choose_nf_map_n<512>(e, map); return;
        case 1024: choose_nf_map_n<1024>(e, map); return;
        case 2048: choose_nf_map_n<2048>(e, map); return;
        case 4096: choose_nf_map_n<4096>(e, map); return;
    }
    for (int i = 0; i < 16; i++) map[i] = (unsigned char)i;
}

void build_twiddles(int n, std::vector<float2> &tw1, std::vector<float2> &tw2) {
    const int M = n / 16, R3 = n / 256;
    tw1.resize((size_t)15 * M);
    tw2.resize((size_t)15 * R3);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 1; k < 16; k++)
        for (int j = 0; j < M; j++) {
            const double ang = -two_pi * (double)((long long)j * k % n) / (double)n;
            tw1[(size_t)(k - 1) * M + j] = make_float2((float)cos(ang), (float)sin(ang));
        }
    for (int k = 1; k < 16; k++)
        for (int n3 = 0; n3 < R3; n3++) {
            const double ang = -two_pi * (double)((n3 * k) % M) / (double)M;
            tw2[(size_t)(k - 1) * R3 + n3] = make_float2((float)cos(ang), (float)sin(ang));
        }
}

void free_slot(Slot &s) {
    cudaFreeHost(s.h_desc);
    cudaFree(s.d_desc);
    cudaFree(s.d_psd_floor);
    cudaFree(s.d_variance);
    cudaFree(s.d_thresholds);
    cudaFree(s.d_taps);
    cudaFree(s.d_keys);
    cudaFree(s.d_key_bits);
    cudaFreeHost(s.h_key_bits);
    cudaFree(s.d_flush_block);
    cudaFree(s.d_flush_n_peaks);
    cudaFree(s.d_flush_peaks);
    cudaFree(s.d_flush_cum);
    cudaFree(s.d_spectrum);
    cudaFree(s.d_spec_round);
    cudaFree(s.d_nf_part);
    cudaFree(s.d_xto);
    cudaFree(s.d_nf_edge);
    cudaFree(s.d_wide_tmp);
    cudaFree(s.d_wide_ready);
    cudaFree(s.d_wide_err);
    cudaFree(s.d_wide_map);
    cudaFreeHost(s.h_wide_err);
    cudaFree(s.d_psd);
    cudaFree(s.d_iq);
    cudaFree(s.d_tmp);
    cudaFreeHost(s.h_psd_floor);
    cudaFreeHost(s.h_variance);
    cudaFreeHost(s.h_thresholds);
    cudaFreeHost(s.h_taps);
    cudaFreeHost(s.h_keys);
    cudaFreeHost(s.h_flush_block);
    cudaFreeHost(s.h_flush_n_peaks);
    cudaFreeHost(s.h_flush_peaks);
    cudaFreeHost(s.h_flush_cum);
    cudaFreeHost(s.h_spectrum);
    cudaFreeHost(s.h_psd);
    if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
    if (s.ev_desc) cudaEventDestroy(s.ev_desc);
    if (s.ev_k0) cudaEventDestroy(s.ev_k0);
    if (s.ev_km) cudaEventDestroy(s.ev_km);
    if (s.ev_k2s) cudaEventDestroy(s.ev_k2s);
    if (s.ev_thr) cudaEventDestroy(s.ev_thr);
    if (s.ev_peaks) cudaEventDestroy(s.ev_peaks);
    if (s.ev_k1) cudaEventDestroy(s.ev_k1);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    s = Slot();
}

int alloc_slot(sdr_engine *e, Slot &s) {
    const size_t MB = (size_t)e->cfg.max_blocks_per_batch, MF = (size_t)e->max_flushes, MP = (size_t)e->cfg.max_peaks_per_flush;
    const size_t N = (size_t)e->N, TS = (size_t)e->tap_stride;
    const DescLayout dl = desc_layout(e);
    s.desc_bytes = dl.total;
    CK(e, cudaMallocHost((void **)&s.h_desc, dl.total));
    CK(e, cudaMalloc((void **)&s.d_desc, dl.total));
    CK(e, cudaMalloc((void **)&s.d_psd_floor, MB * sizeof(float)));
    CK(e, cudaMalloc((void **)&s.d_variance, MB * sizeof(double)));
    CK(e, cudaMalloc((void **)&s.d_thresholds, MB * 4 * sizeof(float)));
    CK(e, cudaMalloc((void **)&s.d_taps, MB * TS * sizeof(float)));
    CK(e, cudaMalloc((void **)&s.d_keys, MB * TS));
    CK(e, cudaMalloc((void **)&s.d_key_bits, MB * (size_t)e->key_words * sizeof(uint32_t)));
    CK(e, cudaMallocHost((void **)&s.h_key_bits, MB * (size_t)e->key_words * sizeof(uint32_t)));
    CK(e, cudaMalloc((void **)&s.d_flush_block, MF * sizeof(int)));
    CK(e, cudaMalloc((void **)&s.d_flush_n_peaks, MF * sizeof(int)));
    CK(e, cudaMalloc((void **)&s.d_flush_peaks, MF * MP * sizeof(sdr_peak)));
    CK(e, cudaMalloc((void **)&s.d_flush_cum, MF * N * sizeof(float)));
    CK(e, cudaMallocHost((void **)&s.h_psd_floor, MB * sizeof(float)));
    CK(e, cudaMallocHost((void **)&s.h_variance, MB * sizeof(double)));
    CK(e, cudaMallocHost((void **)&s.h_thresholds, MB * 4 * sizeof(float)));
    CK(e, cudaMallocHost((void **)&s.h_taps, MB * TS * sizeof(float)));
    CK(e, cudaMallocHost((void **)&s.h_keys, MB * TS));
    CK(e, cudaMallocHost((void **)&s.h_flush_block, MF * sizeof(int)));
    CK(e, cudaMallocHost((void **)&s.h_flush_n_peaks, MF * sizeof(int)));
    CK(e, cudaMallocHost((void **)&s.h_flush_peaks, MF * MP * sizeof(sdr_peak)));
    CK(e, cudaMallocHost((void **)&s.h_flush_cum, MF * N * sizeof(float)));
    if (e->large && e->d_tw_step) {
        // register-resident path: intermediate + dB spectrum of ONE round; spectrum / psd only on request (lazy)
        const size_t RB = (size_t)e->round_blocks < MB ? (size_t)e->round_blocks : MB;
        CK(e, cudaMalloc((void **)&s.d_tmp, RB * N * sizeof(float2)));
        CK(e, cudaMalloc((void **)&s.d_spec_round, RB * N * sizeof(float)));
        CK(e, cudaMalloc((void **)&s.d_nf_part, MB * (size_t)(e->lg.n1 / 16) * 10 * sizeof(double2)));
        CK(e, cudaMalloc((void **)&s.d_xto, MB * 10 * sizeof(float)));
        CK(e, cudaMalloc((void **)&s.d_nf_edge, MB * sizeof(int)));
        if (e->k1_wide) {
            const int n_slots = e->k1w_max_teams * e->k1w_ring;
            CK(e, cudaMalloc((void **)&s.d_wide_tmp, (size_t)n_slots * N * sizeof(float2)));
            CK(e, cudaMalloc((void **)&s.d_wide_ready, MB * sizeof(int)));
            CK(e, cudaMalloc((void **)&s.d_wide_err, sizeof(int)));
            CK(e, cudaMemset(s.d_wide_err, 0, sizeof(int)));
            CK(e, cudaMalloc((void **)&s.d_wide_map, sizeof(CUtensorMap)));
            CK(e, cudaMallocHost((void **)&s.h_wide_err, sizeof(int)));
            *s.h_wide_err = 0;
            CUtensorMap m;
            if (!encode_tile_map(e, &m, s.d_wide_tmp, n_slots)) {
                e->err = "cuTensorMapEncodeTiled failed for the intermediate ring";
                return SDR_ECUDA;
            }
            CK(e, cudaMemcpy(s.d_wide_map, &m, sizeof(m), cudaMemcpyHostToDevice));
        }
    }
    CK(e, cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
    CK(e, cudaEventCreateWithFlags(&s.ev_desc, cudaEventDisableTiming));
    CK(e, cudaEventCreate(&s.ev_k0));
    CK(e, cudaEventCreate(&s.ev_km));
    CK(e, cudaEventCreate(&s.ev_k2s));
    CK(e, cudaEventCreateWithFlags(&s.ev_thr, cudaEventDisableTiming));
    CK(e, cudaEventCreateWithFlags(&s.ev_peaks, cudaEventDisableTiming));
    CK(e, cudaEventCreate(&s.ev_k1));
    CK(e, cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    return SDR_OK;
}

Slot *find_slot(sdr_engine *e, sdr_ticket t) {
    for (auto &s : e->slots)
        if (s.busy && s.ticket == t) return &s;
    return nullptr;
}

__global__ void __launch_bounds__(128) noise_floor_kernel(const float *psd, int n, int e, float *out_min, double *out_var) {
    __shared__ double wsum[32];
    const int ws = nf_window_size(n, e);
    const int n_win = nf_window_count(n, e);
    nf_window_sums<128>(psd, wsum, wsum + 16, e, ws, n_win, threadIdx.x);
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const bool real = lane < n_win;
        nf_select_variance(real ? wsum[lane] : 0.0, real ? wsum[16 + lane] : 0.0,
                           real ? (double)psd[e + (lane + 1) * ws] : 0.0, ws, n_win, lane, out_min, out_var);
    }
}

// dsp.FindNoiseFloor (dsp/fft.go:215-252) replayed statement by statement: the slow, exact path for edge widths whose
// windows are too narrow for the parallel window sums (the reference then closes up to 19 windows, or none).
// float64 sums in index order, no FMA contraction (Go/amd64 does not fuse).
__device__ void nf_exact_block(const float *psd, int n, int edge, float *out_min, double *out_var) {
    const int ws = (n - 2 * edge) / 10;
    double lowest = (double)psd[0], acc = 0.0, win_mean = 0.0;
    int filled = 0, start = 0, win_from = 0, win_to = 0;
    bool none_yet = true;
    for (int i = edge; i < n - edge; i++) {
        if (filled == 0) start = i;
        if (filled == ws) {
            filled = 0;
            const double mean = __ddiv_rn(acc, (double)ws);
            if (mean < lowest || none_yet) {
                lowest = mean;
                none_yet = false;
                win_mean = mean;
                win_from = start;
                win_to = i;
            }
            acc = 0.0;
        }
        acc = __dadd_rn(acc, (double)psd[i]);
        filled++;
    }
    acc = 0.0;
    for (int i = win_from; i <= win_to; i++) {
        const double d = __dsub_rn((double)psd[i], win_mean);
        acc = __dadd_rn(acc, __dmul_rn(d, d));
    }
    *out_var = __ddiv_rn(acc, (double)ws);
    *out_min = (float)lowest;
}

__global__ void __launch_bounds__(128) nf_exact_kernel(const ExactNf *works, const float *psd, int n, float *psd_floor, double *variance) {
    const ExactNf w = works[blockIdx.x];
    for (int b = threadIdx.x; b < w.n_blocks; b += blockDim.x)
        nf_exact_block(psd + (size_t)(w.block_out + b) * n, n, w.edge_width, &psd_floor[w.block_out + b], &variance[w.block_out + b]);
}

__global__ void nf_exact_single_kernel(const float *psd, int n, int edge, float *out_min, double *out_var) {
    nf_exact_block(psd, n, edge, out_min, out_var);
}

__global__ void kiwi_decode_kernel(const uint32_t *raw, int n_samples, float2 *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_samples) out[i] = kiwi_decode_sample(raw[i]);
}

int ensure_scratch(sdr_engine *e, size_t bytes) {
    if (e->scratch_bytes >= bytes) return SDR_OK;
    if (e->d_scratch) cudaFree(e->d_scratch);
    e->d_scratch = nullptr;
    e->scratch_bytes = 0;
    CK(e, cudaMalloc((void **)&e->d_scratch, bytes));
    e->scratch_bytes = bytes;
    return SDR_OK;
}

}  // namespace

extern "C" {

const char *sdr_version(void) { return "sdrgpu 0.1 (sm_100a)"; }

int sdr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *sdr_last_error(const sdr_engine *e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int sdr_engine_create(const sdr_engine_config *cfg, sdr_engine **out) {
    if (!cfg || !out) return SDR_EINVAL;
    *out = nullptr;
    LargeGeom lgeom{};
    const bool is_large = large_geom(cfg->block_size, &lgeom);
    if (!supported_fused_n(cfg->block_size) && !is_large) {
        g_create_error = "block_size must be 512..4096 (fused path) or 8192..65536 (large-block path), a power of two";
        return SDR_EINVAL;
    }
    if (cfg->max_streams < 1 || cfg->max_blocks_per_batch < 1 || cfg->n_slots < 1 || cfg->max_listeners < 0 ||
        cfg->max_listeners > K1Geom<2048>::LMAX || cfg->max_peaks_per_flush < 1) {
        g_create_error = "bad engine configuration (max_listeners <= 256, everything else >= 1)";
        return SDR_EINVAL;
    }
    sdr_engine *e = new (std::nothrow) sdr_engine();
    if (!e) return SDR_ENOMEM;
    e->cfg = *cfg;
    e->N = cfg->block_size;
    e->large = is_large;
    e->lg = lgeom;
    e->tap_stride = ((cfg->max_listeners > 0 ? cfg->max_listeners : 1) + 3) / 4 * 4;
    e->key_words = (e->tap_stride + 31) / 32;
    e->max_segs = cfg->max_blocks_per_batch / SDR_CUMULATION_SIZE + 2 * cfg->max_streams + 2;
    e->max_flushes = cfg->max_blocks_per_batch / SDR_CUMULATION_SIZE + cfg->max_streams + 1;
    if (is_large) e->max_segs += cfg->max_blocks_per_batch / 16 + 2;  // segments are also cut at round boundaries (>= 16 blocks)
    auto fail = [&](int code) {
        g_create_error = e->err;
        sdr_engine_destroy(e);
        return code;
    };
#define CKC(call)                                                                      \
    do {                                                                               \
        cudaError_t _st = (call);                                                      \
        if (_st != cudaSuccess) {                                                      \
            e->err = std::string(#call) + ": " + cudaGetErrorString(_st);              \
            return fail(SDR_ECUDA);                                                    \
        }                                                                              \
    } while (0)
    CKC(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, cfg->device));
    e->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        e->err = "libsdrgpu requires an sm_100a device (B200); found compute capability " + std::to_string(prop.major) + "." +
                 std::to_string(prop.minor);
        return fail(SDR_ECUDA);
    }
    CKC(cudaStreamCreateWithFlags(&e->s_desc, cudaStreamNonBlocking));
    CKC(cudaEventCreateWithFlags(&e->ev_post, cudaEventDisableTiming));
    {
        // Measured on B200 (tools/bench_configs.py, 100 steps): K1 fills every SM's registers, so a K2 on a second stream
        // only runs when a K1 drains -- 0.827 ms per step against 0.825 ms with K2 behind K1 on the compute stream.  The
        // second stream stays available (SDR_K2_OVERLAP=1) for engines whose K1 leaves room.
        const char *v = getenv("SDR_K2_OVERLAP");
        if (v && v[0] == '1') {
            // K1 fills every SM (registers and shared memory), so K2's CTAs run when a K1 drains: with the higher
            // priority they go first and the batch's results reach the D2H stream as early as possible, while the next
            // K1's CTAs fill in behind them (SDR_K2_PRIO=low: the other order; measured within 0.5 % of each other)
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            const char *pv = getenv("SDR_K2_PRIO");
            CKC(cudaStreamCreateWithPriority(&e->s_post, cudaStreamNonBlocking, (pv && pv[0] == 'l') ? lo : hi));
            e->own_post = true;
        }
    }
    {
        const char *av = getenv("SDR_K2_AUX");
        if (!(av && av[0] == '0')) CKC(cudaStreamCreateWithFlags(&e->s_aux, cudaStreamNonBlocking));
    }
    if (cfg->cuda_stream) {
        e->own_streams = false;
        e->s_compute = e->s_h2d = e->s_d2h = (cudaStream_t)cfg->cuda_stream;
    } else {
        CKC(cudaStreamCreateWithFlags(&e->s_compute, cudaStreamNonBlocking));
        CKC(cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking));
        CKC(cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking));
    }
    if (!e->s_post) e->s_post = e->s_compute;
    if (!e->large) {
        std::vector<float2> tw1, tw2;
        build_twiddles(e->N, tw1, tw2);
        CKC(cudaMalloc((void **)&e->d_tw1, tw1.size() * sizeof(float2)));
        CKC(cudaMalloc((void **)&e->d_tw2, tw2.size() * sizeof(float2)));
        CKC(cudaMemcpy(e->d_tw1, tw1.data(), tw1.size() * sizeof(float2), cudaMemcpyHostToDevice));
        CKC(cudaMemcpy(e->d_tw2, tw2.data(), tw2.size() * sizeof(float2), cudaMemcpyHostToDevice));
    } else {
        auto table = [&](int len, float2 **dst) -> cudaError_t {
            std::vector<float2> t(len);
            const double two_pi = 6.283185307179586476925286766559;
            for (int m = 0; m < len; m++) {
                const double ang = -two_pi * (double)m / (double)len;
                t[m] = make_float2((float)cos(ang), (float)sin(ang));
            }
            cudaError_t st = cudaMalloc((void **)dst, (size_t)len * sizeof(float2));
            if (st != cudaSuccess) return st;
            return cudaMemcpy(*dst, t.data(), (size_t)len * sizeof(float2), cudaMemcpyHostToDevice);
        };
        CKC(table(e->lg.n1, &e->d_tw_sub1));
        CKC(table(e->lg.n2, &e->d_tw_sub2));
        if (large_fast_geom(e->lg.n1, e->lg.n2)) {
            const int n1 = e->lg.n1, n2 = e->lg.n2;
            std::vector<float2> t((size_t)n1 * n2);
            const double two_pi = 6.283185307179586476925286766559;
            for (int k = 0; k < n1; k++)
                for (int c = 0; c < n2; c++) {
                    const double ang = -two_pi * (double)(((long long)c * k) % e->N) / (double)e->N;
                    t[(size_t)k * n2 + c] = make_float2((float)cos(ang), (float)sin(ang));
                }
            CKC(cudaMalloc((void **)&e->d_tw_step, t.size() * sizeof(float2)));
            CKC(cudaMemcpy(e->d_tw_step, t.data(), t.size() * sizeof(float2), cudaMemcpyHostToDevice));
            // Rounds: by default the whole batch is one round.  Measured on B200 (tools/bench_configs.py --only large):
            // rounds small enough to keep the intermediate (8N) and the dB spectrum (4N bytes per block) L2-resident
            // (SDR_LARGE_ROUND_MB=32..96) cost more in per-launch ramp and tail than they save in HBM traffic.
            const char *mb = getenv("SDR_LARGE_ROUND_MB");
            if (mb && atoi(mb) > 0) {
                e->round_blocks = (int)(((long long)atoi(mb) << 20) / ((long long)12 * e->N));
                if (e->round_blocks < 16) e->round_blocks = 16;
            } else {
                e->round_blocks = cfg->max_blocks_per_batch > 16 ? cfg->max_blocks_per_batch : 16;
            }
            if (e->N == 65536) {
                const char *wv = getenv("SDR_K1_WIDE");
                e->k1_wide = (wv && wv[0] == '0') ? 0 : (wv && wv[0] == 'f') ? 2 : 1;
                const char *dv = getenv("SDR_K1_WIDE_LOOKAHEAD");
                if (dv && atoi(dv) >= 3 && atoi(dv) <= 8) e->k1w_lookahead = atoi(dv);
                e->k1w_ring = 2 * e->k1w_lookahead;
                const char *rv = getenv("SDR_K1_WIDE_RING");
                if (rv && atoi(rv) >= 2 * e->k1w_lookahead && atoi(rv) <= 16) e->k1w_ring = atoi(rv);  // R >= 2 D (k1_wide.cuh)
                const char *kv = getenv("SDR_K1_WIDE_DISCARD");
                e->k1w_discard = !(kv && kv[0] == '0');
                int occ = 0, coop = 0;
                cudaDriverEntryPointQueryResult qres;
                void *fn = nullptr;
                if (e->k1_wide &&
                    (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg->device) != cudaSuccess || !coop ||
                     cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
                     qres != cudaDriverEntryPointSuccess ||
                     cudaFuncSetAttribute(k1_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K1W_SMEM_BYTES) != cudaSuccess ||
                     cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k1_wide_kernel, K1W_THREADS, K1W_SMEM_BYTES) != cudaSuccess || occ < 1))
                    e->k1_wide = 0;
                cudaGetLastError();
                if (e->k1_wide) {
                    e->encode_tiled = (PFN_cuTensorMapEncodeTiled_v12000)fn;
                    e->k1w_max_teams = occ * e->sm_count / K1W_TEAM;
                    if (const char *v = getenv("SDR_K1_WIDE_TEAMS")) {  // measurement: fewer teams than fit (e.g. one CTA per SM)
                        const int t = atoi(v);
                        if (t >= 1 && t < e->k1w_max_teams) e->k1w_max_teams = t;
                    }
                    if (e->k1w_max_teams < 1) e->k1_wide = 0;
                }
            }
        }
    }
    if (cfg->window) {
        CKC(cudaMalloc((void **)&e->d_window, (size_t)e->N * sizeof(float)));
        CKC(cudaMemcpy(e->d_window, cfg->window, (size_t)e->N * sizeof(float), cudaMemcpyHostToDevice));
        e->cfg.window = nullptr;  // never retain the caller's pointer (cgo rule)
    }
    CKC(cudaMalloc((void **)&e->d_cum_state, (size_t)2 * cfg->max_streams * e->N * sizeof(float)));
    CKC(cudaMemset(e->d_cum_state, 0, (size_t)2 * cfg->max_streams * e->N * sizeof(float)));
    if (e->N == 512) {
        std::vector<float2> t512(256), t256(256);
        const double two_pi = 6.283185307179586476925286766559;
        for (int m = 0; m < 256; m++) {
            t512[m] = make_float2((float)cos(-two_pi * m / 512.0), (float)sin(-two_pi * m / 512.0));
            t256[m] = make_float2((float)cos(-two_pi * m / 256.0), (float)sin(-two_pi * m / 256.0));
        }
        CKC(cudaMalloc((void **)&e->d_tw512, 256 * sizeof(float2)));
        CKC(cudaMemcpy(e->d_tw512, t512.data(), 256 * sizeof(float2), cudaMemcpyHostToDevice));
        CKC(cudaMalloc((void **)&e->d_tw256w, 256 * sizeof(float2)));
        CKC(cudaMemcpy(e->d_tw256w, t256.data(), 256 * sizeof(float2), cudaMemcpyHostToDevice));
        const char *v = getenv("SDR_K1_WARP");
        e->k1_warp = !(v && v[0] == '0');
        if (e->k1_warp) e->k1w_grid_cap = k1w_grid_cap_for(e->d_window != nullptr, e->sm_count);
    }
    if (e->N == 4096 || e->N == 8192) {
        const int r1 = e->N / 256;
        std::vector<float2> t((size_t)r1 * 256), t256(256);
        const double two_pi = 6.283185307179586476925286766559;
        for (int k = 0; k < r1; k++)
            for (int c = 0; c < 256; c++) {
                const double ang = -two_pi * (double)((c * k) % e->N) / (double)e->N;
                t[(size_t)k * 256 + c] = make_float2((float)cos(ang), (float)sin(ang));
            }
        for (int m = 0; m < 256; m++) {
            const double ang = -two_pi * (double)m / 256.0;
            t256[m] = make_float2((float)cos(ang), (float)sin(ang));
        }
        CKC(cudaMalloc((void **)&e->d_tw_mid, t.size() * sizeof(float2)));
        CKC(cudaMemcpy(e->d_tw_mid, t.data(), t.size() * sizeof(float2), cudaMemcpyHostToDevice));
        CKC(cudaMalloc((void **)&e->d_tw256m, t256.size() * sizeof(float2)));
        CKC(cudaMemcpy(e->d_tw256m, t256.data(), t256.size() * sizeof(float2), cudaMemcpyHostToDevice));
        if (e->N == 4096) {
            const char *v4 = getenv("SDR_K1_MID4K");
            if (!(v4 && v4[0] == '0')) {
                e->k1m4_grid_cap = k1m4_grid_cap_for(e->sm_count);
                e->k1_mid4k = e->k1m4_grid_cap > 0;
                cudaGetLastError();
            }
        }
        if (e->N == 8192) {
            const char *v8 = getenv("SDR_K1_MID8K");
            e->k1_mid8k = (v8 && v8[0] == '0') ? 0 : (v8 && v8[0] == 'f') ? 2 : 1;
            const char *st = getenv("SDR_K1_MID8K_STAGES");
            e->k1m8_stages = (st && st[0] == '1') ? 1 : 2;
            for (int dbgv = 0; dbgv < 2 && e->k1_mid8k; dbgv++)
                if (cudaFuncSetAttribute(k1m8g_fn(e->k1m8_stages, dbgv != 0, e->d_window != nullptr), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         k1m8_smem(e->k1m8_stages)) != cudaSuccess)
                    e->k1_mid8k = 0;
            cudaGetLastError();
        }
    }
    CKC(cudaMalloc((void **)&e->d_rolling, (size_t)cfg->max_streams * sizeof(RollingState)));
    CKC(cudaMemset(e->d_rolling, 0, (size_t)cfg->max_streams * sizeof(RollingState)));
    CKC(cudaMalloc((void **)&e->d_deb, (size_t)cfg->max_streams * e->tap_stride * sizeof(DebounceState)));
    CKC(cudaMemset(e->d_deb, 0, (size_t)cfg->max_streams * e->tap_stride * sizeof(DebounceState)));
    e->streams.resize(cfg->max_streams);
    e->nf_map_cache.assign((size_t)(e->N / 2 + 1) * 16, 0xff);
    {
        const char *v = getenv("SDR_K1_TW2R");  // experiment switch; default: pass-2 twiddles in registers
        e->k1_tw2r = !(v && v[0] == '0');
    }
    if (!e->large) e->k1_grid_cap = k1_grid_cap_for(e->N, e->d_window != nullptr, e->k1_tw2r, e->sm_count);
    CKC(cudaGetLastError());
    e->slots.resize(cfg->n_slots);
    for (auto &s : e->slots) {
        int rc = alloc_slot(e, s);
        if (rc != SDR_OK) return fail(rc);
    }
    CKC(cudaDeviceSynchronize());
#undef CKC
    *out = e;
    return SDR_OK;
}

void sdr_engine_destroy(sdr_engine *e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    cudaDeviceSynchronize();
#ifdef SDR_K1W_TRACE
    if (g_k1w_trace && getenv("SDR_K1_WIDE_TRACE")) {
        std::vector<long long> h(K1W_TRACE_BYTES / sizeof(long long));
        cudaMemcpy(h.data(), g_k1w_trace, K1W_TRACE_BYTES, cudaMemcpyDeviceToHost);
        if (FILE *f = fopen(getenv("SDR_K1_WIDE_TRACE"), "wb")) {
            fwrite(h.data(), 1, K1W_TRACE_BYTES, f);
            fclose(f);
        }
    }
#endif
    for (auto &s : e->slots) free_slot(s);
    cudaFree(e->d_tw1);
    cudaFree(e->d_tw2);
    cudaFree(e->d_tw_step);
    cudaFree(e->d_tw_mid);
    cudaFree(e->d_tw512);
    cudaFree(e->d_tw256w);
    cudaFree(e->d_tw256m);
    cudaFree(e->d_tw_sub1);
    cudaFree(e->d_tw_sub2);
    cudaFree(e->d_window);
    cudaFree(e->d_cum_state);
    cudaFree(e->d_rolling);
    cudaFree(e->d_deb);
    cudaFree(e->d_scratch);
    if (e->s_desc) cudaStreamDestroy(e->s_desc);
    if (e->own_post && e->s_post) cudaStreamDestroy(e->s_post);
    if (e->s_aux) cudaStreamDestroy(e->s_aux);
    if (e->ev_post) cudaEventDestroy(e->ev_post);
    if (e->own_streams) {
        if (e->s_compute) cudaStreamDestroy(e->s_compute);
        if (e->s_h2d) cudaStreamDestroy(e->s_h2d);
        if (e->s_d2h) cudaStreamDestroy(e->s_d2h);
    }
    delete e;
}

int sdr_alloc_pinned(sdr_engine *e, size_t bytes, void **out) {
    if (!e || !out) return SDR_EINVAL;
    CK(e, cudaSetDevice(e->cfg.device));
    CK(e, cudaMallocHost(out, bytes));
    return SDR_OK;
}

int sdr_free_pinned(sdr_engine *e, void *p) {
    if (!e) return SDR_EINVAL;
    CK(e, cudaFreeHost(p));
    return SDR_OK;
}

static int stream_reset_locked(sdr_engine *e, int stream) {
    CK(e, cudaSetDevice(e->cfg.device));
    // cumulationCount := 0 (rx/receiver.go:347): with cum_count == 0 the next segment of the stream starts from
    // state_in = -1, so the stream's two cum_state rows (2*stream, 2*stream+1) are never read and need no clearing
    e->streams[stream].cum_count = 0;
    e->streams[stream].state_row = 0;
    CK(e, cudaMemsetAsync(e->d_rolling + stream, 0, sizeof(RollingState), e->s_post));
    CK(e, cudaMemsetAsync(e->d_deb + (size_t)stream * e->tap_stride, 0, (size_t)e->tap_stride * sizeof(DebounceState), e->s_post));
    return SDR_OK;
}

int sdr_stream_open(sdr_engine *e, int sample_rate, int *out_stream) {
    if (!e || !out_stream || sample_rate <= 0) return SDR_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    for (size_t i = 0; i < e->streams.size(); i++) {
        if (!e->streams[i].open) {
            e->streams[i].open = true;
            e->streams[i].sample_rate = sample_rate;
            *out_stream = (int)i;
            return stream_reset_locked(e, (int)i);
        }
    }
    e->err = "no free stream slot (max_streams reached)";
    return SDR_EBUSY;
}

int sdr_stream_close(sdr_engine *e, int stream) {
    if (!e) return SDR_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    if (stream < 0 || stream >= (int)e->streams.size() || !e->streams[stream].open) return SDR_EINVAL;
    e->streams[stream].open = false;
    return SDR_OK;
}

int sdr_stream_reset(sdr_engine *e, int stream) {
    if (!e) return SDR_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    if (stream < 0 || stream >= (int)e->streams.size() || !e->streams[stream].open) return SDR_EINVAL;
    return stream_reset_locked(e, stream);
}

int sdr_stream_cumulation_count(sdr_engine *e, int stream, int *out) {
    if (!e || !out) return SDR_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    if (stream < 0 || stream >= (int)e->streams.size() || !e->streams[stream].open) return SDR_EINVAL;
    *out = e->streams[stream].cum_count;
    return SDR_OK;
}

int64_t sdr_engine_launch_count(const sdr_engine *e) { return e ? e->launches : 0; }

const char *sdr_engine_last_kernel(const sdr_engine *e) { return e ? e->last_kernel : ""; }

int sdr_engine_fence(sdr_engine *e) {
    if (!e) return SDR_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(e, cudaSetDevice(e->cfg.device));
    if (e->s_post != e->s_compute) CK(e, cudaStreamWaitEvent(e->s_compute, e->ev_post, 0));
    return SDR_OK;
}

int sdr_submit(sdr_engine *e, const sdr_work *works, int n_works, int flags, sdr_ticket *out) {
    if (!e || !works || !out || n_works < 1) return SDR_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    NvtxRange nvtx_submit("sdr_submit");
    if (n_works > e->cfg.max_streams) {
        e->err = "more works than max_streams";
        return SDR_EINVAL;
    }
    CK(e, cudaSetDevice(e->cfg.device));
    const int N = e->N;
    // ---- validate (programmer errors: the Go reference panics or logs-and-drops) ----
    long long total_blocks = 0;
    bool any_host = false;
    bool need_exact = false;  // some work's edge width needs the literal FindNoiseFloor replay on the stored PSD
    {
        std::vector<char> seen(e->streams.size(), 0);
        for (int w = 0; w < n_works; w++) {
            const sdr_work &wk = works[w];
            if (wk.stream < 0 || wk.stream >= (int)e->streams.size() || !e->streams[wk.stream].open) {
                e->err = "work " + std::to_string(w) + ": stream not open";
                return SDR_EINVAL;
            }
            if (seen[wk.stream]) {
                e->err = "work " + std::to_string(w) + ": a stream may appear only once per submit";
                return SDR_EINVAL;
            }
            seen[wk.stream] = 1;
            if (wk.n_blocks < 1 || !wk.iq) {
                e->err = "work " + std::to_string(w) + ": empty";
                return SDR_EINVAL;
            }
            if (wk.n_listeners < 0 || wk.n_listeners > e->cfg.max_listeners || (wk.n_listeners > 0 && !wk.listener_bins)) {
                e->err = "work " + std::to_string(w) + ": bad listener count";
                return SDR_EINVAL;
            }
            for (int l = 0; l < wk.n_listeners; l++)
                if (wk.listener_bins[l] < 0 || wk.listener_bins[l] >= N) {
                    e->err = "work " + std::to_string(w) + ": listener bin out of range";
                    return SDR_EINVAL;
                }
            if (wk.edge_width < 0) {  // the reference indexes psd[edgeWidth]: a negative width panics
                e->err = "work " + std::to_string(w) + ": negative edge_width";
                return SDR_EINVAL;
            }
            if (!nf_edge_supported(N, wk.edge_width)) need_exact = true;
            if (wk.mem == SDR_MEM_DEVICE && ((uintptr_t)wk.iq & 15)) {
                e->err = "work " + std::to_string(w) + ": device iq must be 16-byte aligned";
                return SDR_EINVAL;
            }
            if (wk.format != works[0].format || (wk.format != SDR_FMT_F32 && wk.format != SDR_FMT_KIWI_I16BE)) {
                e->err = "work " + std::to_string(w) + ": all works of one submit must share one valid sample format";
                return SDR_EINVAL;
            }
            if (wk.format == SDR_FMT_KIWI_I16BE && (e->large || e->d_window)) {
                e->err = "SDR_FMT_KIWI_I16BE is supported on the fused, unwindowed path (N <= 4096)";
                return SDR_EINVAL;
            }
            if (wk.mem != SDR_MEM_DEVICE) any_host = true;
            total_blocks += wk.n_blocks;
        }
    }
    if (total_blocks > e->cfg.max_blocks_per_batch) {
        e->err = "batch exceeds max_blocks_per_batch";
        return SDR_EINVAL;
    }
    Slot *sp = nullptr;
    for (auto &s : e->slots)
        if (!s.busy) {
            sp = &s;
            break;
        }
    if (!sp) {
        e->err = "all in-flight slots busy";
        return SDR_EBUSY;
    }
    Slot &s = *sp;
    const bool want_spec = (flags & SDR_WANT_SPECTRUM) != 0;
    const bool dbg = want_spec || need_exact;  // kernel variant that stores spectrum[] and psd[]
    const bool i16 = works[0].format == SDR_FMT_KIWI_I16BE;
    const size_t sample_floats = i16 ? 1 : 2;  // a complex sample is 4 bytes on the Kiwi wire, 8 as float32 pairs
    if (dbg) {
        const size_t bytes = (size_t)e->cfg.max_blocks_per_batch * N * sizeof(float);
        if (!s.d_spectrum) {
            CK(e, cudaMalloc((void **)&s.d_spectrum, bytes));
            CK(e, cudaMalloc((void **)&s.d_psd, bytes));
        }
        if (want_spec && !s.h_spectrum) {
            CK(e, cudaMallocHost((void **)&s.h_spectrum, bytes));
            CK(e, cudaMallocHost((void **)&s.h_psd, bytes));
        }
    }
    if (any_host && !s.d_iq) CK(e, cudaMalloc((void **)&s.d_iq, (size_t)e->cfg.max_blocks_per_batch * 2 * N * sizeof(float)));

    // ---- build descriptors ----
    const DescLayout dl = desc_layout(e);
    Segment *segs = reinterpret_cast<Segment *>(s.h_desc + dl.segs);
    WorkParams *wps = reinterpret_cast<WorkParams *>(s.h_desc + dl.works);
    PostWork *pws = reinterpret_cast<PostWork *>(s.h_desc + dl.post);
    int *lbins = reinterpret_cast<int *>(s.h_desc + dl.lbins);
    uint8_t *lflags = reinterpret_cast<uint8_t *>(s.h_desc + dl.lflags);
    ExactNf *exs = reinterpret_cast<ExactNf *>(s.h_desc + dl.exact);
    int n_exact = 0;
    s.work_block_offset.assign(n_works + 1, 0);
    s.work_flush_offset.assign(n_works + 1, 0);
    int n_segs = 0, n_flushes = 0, block_off = 0, lb_off = 0;
    size_t iq_off = 0;  // floats into d_iq
    std::vector<StreamInfo> new_state(n_works);
    int max_work_blocks = 1;
    bool any_debounce = false;  // a work with a real debouncer (SetSignalDebounce >= 2): k2_debounce_kernel runs too
    bool warp_ok = (N == 512);        // ... and for k1_warp_kernel
    const float *pend_src = nullptr;  // pending coalesced H2D copy
    float *pend_dst = nullptr;
    size_t pend_n = 0;
    for (int w = 0; w < n_works; w++) {
        const sdr_work &wk = works[w];
        // bookkeeping advances on a copy and is committed only when every enqueue of this submit succeeded
        StreamInfo si = e->streams[wk.stream];
        const float *dev_iq = wk.iq;
        if (wk.mem != SDR_MEM_DEVICE) {
            // host IQ: stage it in the slot's device buffer.  Works whose host ranges are adjacent (one pinned
            // ring filled stream after stream) are coalesced into ONE copy: fewer, larger H2D transfers.
            dev_iq = s.d_iq + iq_off;
            const size_t nfl = (size_t)wk.n_blocks * sample_floats * N;
            if (pend_n > 0 && pend_src + pend_n == wk.iq && pend_dst + pend_n == s.d_iq + iq_off) {
                pend_n += nfl;
            } else {
                if (pend_n > 0) CK(e, cudaMemcpyAsync(pend_dst, pend_src, pend_n * sizeof(float), cudaMemcpyHostToDevice, e->s_h2d));
                pend_src = wk.iq;
                pend_dst = s.d_iq + iq_off;
                pend_n = nfl;
            }
            iq_off += nfl;
        }
        // the fused kernels get a width they cover; the exact replay overwrites this work's noise scalars afterwards
        const bool exact = !nf_edge_supported(N, wk.edge_width);
        wps[w].edge_width = exact ? 0 : wk.edge_width;
        if (exact) exs[n_exact++] = ExactNf{block_off, wk.n_blocks, wk.edge_width, 0};
        if (N != 512 || nf_window_size(N, wps[w].edge_width) < K1WarpGeom::MIN_WS) warp_ok = false;
        wps[w].n_listeners = wk.n_listeners;
        wps[w].listener_off = lb_off;
        wps[w].pad = 0;
        {
            unsigned char *cm = &e->nf_map_cache[(size_t)wps[w].edge_width * 16];
            if (cm[0] == 0xff) choose_nf_map(N, wps[w].edge_width, cm);
            memcpy(wps[w].nf_map, cm, 16);
        }
        for (int l = 0; l < wk.n_listeners; l++) lbins[lb_off + l] = wk.listener_bins[l];
        if (wk.listener_flags)
            for (int l = 0; l < wk.n_listeners; l++) lflags[lb_off + l] = wk.listener_flags[l];
        PostWork &pw = pws[w];
        pw.stream = wk.stream;
        pw.block_out = block_off;
        pw.n_blocks = wk.n_blocks;
        pw.flush_out = n_flushes;
        pw.n_flushes = 0;
        pw.first_flush_block = SDR_CUMULATION_SIZE - si.cum_count - 1;
        pw.peak_threshold = wk.peak_threshold;
        pw.n_listeners = wk.n_listeners;
        pw.do_peaks = (flags & SDR_NO_PEAKS) ? 0 : 1;
        pw.debounce = wk.signal_debounce;
        if (wk.signal_debounce >= 2) any_debounce = true;
        pw.lflags_off = wk.listener_flags ? lb_off : -1;
        pw.pad = 0;
        lb_off += wk.n_listeners;
        s.work_block_offset[w] = block_off;
        s.work_flush_offset[w] = n_flushes;
        int c = si.cum_count, pos = 0, rem = wk.n_blocks;
        while (rem > 0) {
            int take = rem < SDR_CUMULATION_SIZE - c ? rem : SDR_CUMULATION_SIZE - c;
            if (e->round_blocks > 0) {  // register-resident large path: a segment never straddles a round
                const int room = e->round_blocks - (block_off + pos) % e->round_blocks;
                if (take > room) take = room;
            }
            if (n_segs >= e->max_segs || n_flushes >= e->max_flushes) {
                e->err = "internal: descriptor capacity exceeded";
                return SDR_ESTATE;
            }
            if (e->large) {
                int *bs = reinterpret_cast<int *>(s.h_desc + dl.block_seg);
                for (int b = 0; b < take; b++) bs[block_off + pos + b] = n_segs;
            }
            Segment &sg = segs[n_segs++];
            sg.iq = dev_iq + (size_t)pos * sample_floats * N;
            sg.n_blocks = take;
            sg.stream = wk.stream;
            sg.work = w;
            sg.block_out = block_off + pos;
            // the partial cumulation of a stream ping-pongs between its two state rows: a work that crosses a window
            // boundary has a first segment that READS the saved state and a last one that WRITES the new state,
            // both in this launch
            sg.state_in = c > 0 ? 2 * wk.stream + si.state_row : -1;
            sg.flush_idx = (c + take == SDR_CUMULATION_SIZE) ? n_flushes++ : -1;
            sg.state_out = 2 * wk.stream + (si.state_row ^ 1);
            if (sg.flush_idx < 0) si.state_row ^= 1;  // the next segment of the stream reads what this one saves
            if (sg.flush_idx >= 0) pw.n_flushes++;
            c = (c + take) % SDR_CUMULATION_SIZE;
            pos += take;
            rem -= take;
        }
        si.cum_count = c;
        new_state[w] = si;
        if (wk.n_blocks > max_work_blocks) max_work_blocks = wk.n_blocks;
        block_off += wk.n_blocks;
    }
    s.work_block_offset[n_works] = block_off;
    s.work_flush_offset[n_works] = n_flushes;
    s.n_works = n_works;
    s.n_blocks = block_off;
    s.n_flushes = n_flushes;
    s.flags = flags;
    s.launches = 0;

    // N = 65536 with enough segments for whole teams: the single-pass kernel; its per-segment tensor maps travel with
    // the descriptors.  Below 4 segments the block-parallel two-kernel path fills the GPU better.
    // (a batch cut into rounds -- SDR_LARGE_ROUND_MB -- has segments that hand a partial window to each other inside one
    // launch: only the round-by-round two-kernel path orders those)
    const bool use_wide = e->N == 65536 && block_off <= e->round_blocks && (e->k1_wide == 2 || (e->k1_wide == 1 && n_segs >= 4));
    if (use_wide) {
        CUtensorMap *maps = reinterpret_cast<CUtensorMap *>(s.h_desc + dl.segmaps);
        for (int i = 0; i < n_segs; i++)
            if (!encode_tile_map(e, &maps[i], segs[i].iq, segs[i].n_blocks)) {
                e->err = "cuTensorMapEncodeTiled failed for an IQ segment";
                return SDR_ECUDA;
            }
    }
    if (pend_n > 0) CK(e, cudaMemcpyAsync(pend_dst, pend_src, pend_n * sizeof(float), cudaMemcpyHostToDevice, e->s_h2d));
    // ---- H2D: descriptors (+ IQ queued above) ----
    CK(e, cudaMemcpyAsync(s.d_desc, s.h_desc, dl.total, cudaMemcpyHostToDevice, e->s_desc));
    CK(e, cudaEventRecord(s.ev_desc, e->s_desc));
    CK(e, cudaEventRecord(s.ev_h2d, e->s_h2d));
    CK(e, cudaStreamWaitEvent(e->s_compute, s.ev_h2d, 0));
    CK(e, cudaStreamWaitEvent(e->s_compute, s.ev_desc, 0));

    // ---- kernels ----
    K1Args a1;
    a1.segs = reinterpret_cast<const Segment *>(s.d_desc + dl.segs);
    a1.n_segs = n_segs;
    a1.works = reinterpret_cast<const WorkParams *>(s.d_desc + dl.works);
    a1.listener_bins = reinterpret_cast<const int *>(s.d_desc + dl.lbins);
    a1.tw1 = e->d_tw1;
    a1.tw2 = e->d_tw2;
    a1.window = e->d_window;
    a1.cum_state = e->d_cum_state;
    a1.psd_floor = s.d_psd_floor;
    a1.variance = s.d_variance;
    a1.taps = s.d_taps;
    a1.tap_stride = e->tap_stride;
    a1.flush_cum = s.d_flush_cum;
    a1.dbg_spectrum = s.d_spectrum;
    a1.dbg_psd = s.d_psd;
    CK(e, cudaEventRecord(s.ev_k0, e->s_compute));
    int k1_launches = 1;
    std::optional<NvtxRange> nvtx_k1;
    nvtx_k1.emplace("K1 spectral (FFT, |X|^2, dB, noise floor, taps, cumulation)");
    if (!e->large) {
        CK(e, launch_k1(e, a1, dbg, e->s_compute, i16, warp_ok));
        e->last_kernel = (e->k1_mid4k && e->N == 4096 && !i16) ? "k1_mid4k_kernel"
                         : (e->k1_warp && warp_ok) ? "k1_warp_kernel"
                         : e->N == 512 ? "k1_spectral_kernel<512>" : e->N == 1024 ? "k1_spectral_kernel<1024>"
                         : e->N == 2048 ? "k1_spectral_kernel<2048>" : "k1_spectral_kernel<4096>";
    } else if (e->N == 8192 && block_off <= e->round_blocks && (e->k1_mid8k == 2 || (e->k1_mid8k == 1 && n_segs >= e->sm_count / 3))) {
        // TMA-staged single pass, one 512-thread CTA per SM (k1_mid8k.cuh).  Below ~SMs/3 segments the block-parallel
        // two-kernel path (which does not serialise the blocks of a stream) is faster.
        const Mid8kNf nfb{s.d_nf_part, s.d_xto, s.d_nf_edge};
        CK(e, launch_k1_mid8k(e, a1, nfb, block_off, dbg, e->s_compute, &k1_launches));
        e->last_kernel = "k1_mid8k2_kernel";
    } else if (use_wide) {
        // single pass over HBM: one team of 16 CTAs per segment, the intermediate in an L2-resident ring (k1_wide.cuh)
        const LargeFastBufs lb{s.d_tmp, s.d_spec_round, s.d_nf_part, s.d_xto, s.d_nf_edge};
        const WideBufs wb{reinterpret_cast<const CUtensorMap *>(s.d_desc + dl.segmaps), s.d_wide_map, s.d_wide_tmp, s.d_wide_ready,
                          s.d_wide_err, s.h_wide_err};
        CK(e, launch_k1_wide(e, a1, lb, wb, block_off, dbg, e->s_compute));
        e->last_kernel = "k1_wide_kernel";
        k1_launches = 2;
    } else if (e->round_blocks > 0) {
        // rounds of consecutive blocks; the segment table is in block order and no segment straddles a round
        std::vector<LargeRound> rounds;
        int sgi = 0;
        for (int b0 = 0; b0 < block_off; b0 += e->round_blocks) {
            LargeRound r;
            r.blk0 = b0;
            r.n_blocks = block_off - b0 < e->round_blocks ? block_off - b0 : e->round_blocks;
            r.seg0 = sgi;
            while (sgi < n_segs && segs[sgi].block_out < b0 + r.n_blocks) sgi++;
            r.n_segs = sgi - r.seg0;
            rounds.push_back(r);
        }
        const LargeFastBufs lb{s.d_tmp, s.d_spec_round, s.d_nf_part, s.d_xto, s.d_nf_edge};
        CK(e, launch_large_fast(e, a1, lb, reinterpret_cast<const int *>(s.d_desc + dl.block_seg), rounds, block_off, dbg, e->s_compute,
                                &k1_launches));
        e->last_kernel = e->lg.n1 == 256 ? "fast_cols256_kernel + fast_rows256_kernel"
                         : e->lg.n1 == 32 ? "fast_cols32_kernel + fast_rows256_kernel" : "fast_cols64_kernel + fast_rows256_kernel";
    } else {
        e->err = "internal: no spectral kernel for this block size";
        return SDR_EINVAL;
    }
    if (n_exact > 0) {
        nf_exact_kernel<<<n_exact, 128, 0, e->s_compute>>>(reinterpret_cast<const ExactNf *>(s.d_desc + dl.exact), s.d_psd, N,
                                                            s.d_psd_floor, s.d_variance);
        CK(e, cudaGetLastError());
        k1_launches++;
    }
    CK(e, cudaEventRecord(s.ev_km, e->s_compute));
    nvtx_k1.reset();
    NvtxRange nvtx_k2("K2 post (thresholds, keys, peaks) + D2H");
    K2Args a2;
    a2.works = reinterpret_cast<const PostWork *>(s.d_desc + dl.post);
    a2.n_works = n_works;
    a2.n_blocks = s.n_blocks;
    a2.rolling = e->d_rolling;
    a2.psd_floor = s.d_psd_floor;
    a2.variance = s.d_variance;
    a2.thresholds = s.d_thresholds;
    a2.taps = s.d_taps;
    a2.keys = (flags & SDR_NO_RAW_KEYS) ? nullptr : s.d_keys;
    a2.key_bits = s.d_key_bits;
    a2.key_words = e->key_words;
    a2.deb = e->d_deb;
    a2.lflags = reinterpret_cast<const uint8_t *>(s.d_desc + dl.lflags);
    a2.tap_stride = e->tap_stride;
    a2.flush_cum = s.d_flush_cum;
    a2.flush_block = s.d_flush_block;
    a2.flush_n_peaks = s.d_flush_n_peaks;
    a2.flush_peaks = s.d_flush_peaks;
    a2.max_peaks = e->cfg.max_peaks_per_flush;
    a2.n = N;
    if (e->s_post != e->s_compute) CK(e, cudaStreamWaitEvent(e->s_post, s.ev_km, 0));
    CK(e, cudaEventRecord(s.ev_k2s, e->s_post));
    int k2_launches = 2;
    if (max_work_blocks > K2_SHORT_WORK) {  // long works: the float64 dB conversions run block-parallel in front of the chains
        k2_db_kernel<<<(s.n_blocks + K2_THREADS - 1) / K2_THREADS, K2_THREADS, 0, e->s_post>>>(a2);
        CK(e, cudaGetLastError());
        // few long works: the whole CTA serves one work (the parallel phases around the chain run four times as wide)
        if (n_works <= 4 * e->sm_count) k2_thresholds_kernel<true, K2_THREADS><<<n_works, K2_THREADS, 0, e->s_post>>>(a2);
        else k2_thresholds_kernel<true, 32><<<(n_works + K2_WARPS - 1) / K2_WARPS, K2_THREADS, 0, e->s_post>>>(a2);
        k2_launches++;
    } else {
        k2_thresholds_kernel<false, 32><<<(n_works + K2_WARPS - 1) / K2_WARPS, K2_THREADS, 0, e->s_post>>>(a2);
    }
    CK(e, cudaGetLastError());
    // keys and peaks both depend on the thresholds only: the peak scan goes to a side stream so that the two small,
    // latency-bound grids share the GPU instead of running back to back
    const bool do_peaks = n_flushes > 0 && !(flags & SDR_NO_PEAKS);
    if (do_peaks && e->s_aux) {
        CK(e, cudaEventRecord(s.ev_thr, e->s_post));
        CK(e, cudaStreamWaitEvent(e->s_aux, s.ev_thr, 0));
        k2_peaks_kernel<<<n_flushes, K2_THREADS, 0, e->s_aux>>>(a2);
        CK(e, cudaGetLastError());
        CK(e, cudaEventRecord(s.ev_peaks, e->s_aux));
        k2_launches++;
    }
    k2_keys_kernel<<<dim3((max_work_blocks + K2_KEY_ROWS - 1) / K2_KEY_ROWS, (n_works + K2_WARPS - 1) / K2_WARPS), K2_THREADS, 0,
                     e->s_post>>>(a2);
    CK(e, cudaGetLastError());
    if (any_debounce) {
        k2_debounce_kernel<<<n_works, K2_THREADS, 0, e->s_post>>>(a2);
        CK(e, cudaGetLastError());
        k2_launches++;
    }
    if (do_peaks && e->s_aux) {
        CK(e, cudaStreamWaitEvent(e->s_post, s.ev_peaks, 0));
    } else if (do_peaks) {
        k2_peaks_kernel<<<n_flushes, K2_THREADS, 0, e->s_post>>>(a2);
        CK(e, cudaGetLastError());
        k2_launches++;
    }
    CK(e, cudaEventRecord(s.ev_k1, e->s_post));
    CK(e, cudaEventRecord(e->ev_post, e->s_post));
    s.launches = k1_launches + k2_launches;
    e->launches += k1_launches + k2_launches;

    // ---- D2H ----
    if (!(flags & SDR_NO_D2H)) {
        CK(e, cudaStreamWaitEvent(e->s_d2h, s.ev_k1, 0));
        const size_t nb = (size_t)block_off, TS = (size_t)e->tap_stride;
        CK(e, cudaMemcpyAsync(s.h_psd_floor, s.d_psd_floor, nb * sizeof(float), cudaMemcpyDeviceToHost, e->s_d2h));
        CK(e, cudaMemcpyAsync(s.h_variance, s.d_variance, nb * sizeof(double), cudaMemcpyDeviceToHost, e->s_d2h));
        CK(e, cudaMemcpyAsync(s.h_thresholds, s.d_thresholds, nb * 4 * sizeof(float), cudaMemcpyDeviceToHost, e->s_d2h));
        if (!(flags & SDR_NO_TAPS))
            CK(e, cudaMemcpyAsync(s.h_taps, s.d_taps, nb * TS * sizeof(float), cudaMemcpyDeviceToHost, e->s_d2h));
        if (!(flags & SDR_NO_RAW_KEYS)) CK(e, cudaMemcpyAsync(s.h_keys, s.d_keys, nb * TS, cudaMemcpyDeviceToHost, e->s_d2h));
        CK(e, cudaMemcpyAsync(s.h_key_bits, s.d_key_bits, nb * e->key_words * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->s_d2h));
        if (n_flushes > 0) {
            const size_t nf = (size_t)n_flushes;
            CK(e, cudaMemcpyAsync(s.h_flush_block, s.d_flush_block, nf * sizeof(int), cudaMemcpyDeviceToHost, e->s_d2h));
            CK(e, cudaMemcpyAsync(s.h_flush_n_peaks, s.d_flush_n_peaks, nf * sizeof(int), cudaMemcpyDeviceToHost, e->s_d2h));
            CK(e, cudaMemcpyAsync(s.h_flush_peaks, s.d_flush_peaks, nf * e->cfg.max_peaks_per_flush * sizeof(sdr_peak),
                                  cudaMemcpyDeviceToHost, e->s_d2h));
            if (flags & SDR_WANT_FLUSH_CUM)
                CK(e, cudaMemcpyAsync(s.h_flush_cum, s.d_flush_cum, nf * N * sizeof(float), cudaMemcpyDeviceToHost, e->s_d2h));
        }
        if (want_spec) {
            CK(e, cudaMemcpyAsync(s.h_spectrum, s.d_spectrum, nb * N * sizeof(float), cudaMemcpyDeviceToHost, e->s_d2h));
            CK(e, cudaMemcpyAsync(s.h_psd, s.d_psd, nb * N * sizeof(float), cudaMemcpyDeviceToHost, e->s_d2h));
        }
        CK(e, cudaEventRecord(s.ev_done, e->s_d2h));
    } else {
        CK(e, cudaEventRecord(s.ev_done, e->s_post));
    }
    for (int w = 0; w < n_works; w++) e->streams[works[w].stream] = new_state[w];
    s.busy = true;
    s.collected = false;
    s.ticket = e->next_ticket++;
    *out = s.ticket;
    return SDR_OK;
}

int sdr_collect(sdr_engine *e, sdr_ticket t, int blocking, sdr_result *out) {
    if (!e || !out) return SDR_EINVAL;
    std::unique_lock<std::mutex> lk(e->mu);
    Slot *sp = find_slot(e, t);
    if (!sp) {
        e->err = "unknown ticket";
        return SDR_EINVAL;
    }
    Slot &s = *sp;
    if (blocking) {
        // wait without the engine lock: other receivers keep submitting while this one blocks on its ticket (the
        // slot cannot go away: only this ticket's owner releases it)
        cudaEvent_t ev = s.ev_done;
        lk.unlock();
        const cudaError_t st = cudaEventSynchronize(ev);
        lk.lock();
        CK(e, st);
    } else {
        cudaError_t st = cudaEventQuery(s.ev_done);
        if (st == cudaErrorNotReady) return SDR_ENOTREADY;
        CK(e, st);
    }
    if (s.h_wide_err && *s.h_wide_err) {
        // report it once for this ticket and re-arm the flag: the slot is usable again after sdr_release
        *s.h_wide_err = 0;
        cudaMemsetAsync(s.d_wide_err, 0, sizeof(int), e->s_compute);
        e->err = "k1_wide: a dependency wait ran out (the batch's results are invalid)";
        s.collected = true;
        return SDR_ECUDA;
    }
    s.collected = true;
    memset(out, 0, sizeof(*out));
    out->n_works = s.n_works;
    out->n_blocks = s.n_blocks;
    out->n_flushes = s.n_flushes;
    out->tap_stride = e->tap_stride;
    out->key_words = e->key_words;
    out->block_size = e->N;
    out->max_peaks_per_flush = e->cfg.max_peaks_per_flush;
    out->work_block_offset = s.work_block_offset.data();
    out->work_flush_offset = s.work_flush_offset.data();
    if (!(s.flags & SDR_NO_D2H)) {
        out->psd_noise_floor = s.h_psd_floor;
        out->noise_variance = s.h_variance;
        out->thresholds = s.h_thresholds;
        out->taps = (s.flags & SDR_NO_TAPS) ? nullptr : s.h_taps;
        out->keys = (s.flags & SDR_NO_RAW_KEYS) ? nullptr : s.h_keys;
        out->key_bits = s.h_key_bits;
        out->flush_block = s.h_flush_block;
        out->flush_n_peaks = s.h_flush_n_peaks;
        out->flush_peaks = s.h_flush_peaks;
        out->flush_cum = (s.flags & SDR_WANT_FLUSH_CUM) ? s.h_flush_cum : nullptr;
        out->spectrum = (s.flags & SDR_WANT_SPECTRUM) ? s.h_spectrum : nullptr;
        out->psd = (s.flags & SDR_WANT_SPECTRUM) ? s.h_psd : nullptr;
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1) == cudaSuccess) out->gpu_ms = ms;
    if (cudaEventElapsedTime(&ms, s.ev_k0, s.ev_km) == cudaSuccess) out->k1_ms = ms;
    if (cudaEventElapsedTime(&ms, s.ev_k2s, s.ev_k1) == cudaSuccess) out->k2_ms = ms;
    out->gpu_launches = s.launches;
    return SDR_OK;
}

int sdr_release(sdr_engine *e, sdr_ticket t) {
    if (!e) return SDR_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    Slot *sp = find_slot(e, t);
    if (!sp) {
        e->err = "unknown ticket";
        return SDR_EINVAL;
    }
    if (!sp->collected) CK(e, cudaEventSynchronize(sp->ev_done));
    sp->busy = false;
    return SDR_OK;
}

int sdr_ticket_device_ptrs(sdr_engine *e, sdr_ticket t, void **psd_noise_floor, void **noise_variance, void **thresholds,
                           void **taps, void **keys) {
    if (!e) return SDR_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    Slot *sp = find_slot(e, t);
    if (!sp) {
        e->err = "unknown ticket";
        return SDR_EINVAL;
    }
    if (psd_noise_floor) *psd_noise_floor = sp->d_psd_floor;
    if (noise_variance) *noise_variance = sp->d_variance;
    if (thresholds) *thresholds = sp->d_thresholds;
    if (taps) *taps = sp->d_taps;
    if (keys) *keys = sp->d_keys;
    return SDR_OK;
}

// ---- dsp-signature-compatible single calls ---------------------------------------------------

int sdr_dsp_iq_to_spectrum_and_psd(sdr_engine *e, const float *iq, int n_blocks, float *spectrum, float *psd) {
    if (!e || !iq || !spectrum || !psd || n_blocks < 1) return SDR_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(e, cudaSetDevice(e->cfg.device));
    const int N = e->N;
    // every block is its own segment so the calls stay independent of any stream state
    const size_t iq_bytes = (size_t)n_blocks * 2 * N * sizeof(float);
    const size_t out_bytes = (size_t)n_blocks * N * sizeof(float);
    const size_t seg_bytes = align_up(sizeof(Segment) * (size_t)n_blocks, 256);
    const size_t misc = 4096;
    const size_t nf_part_bytes = e->round_blocks > 0 ? align_up((size_t)n_blocks * (size_t)(e->lg.n1 / 16) * 10 * sizeof(double2), 256) : 0;
    const size_t fast_bytes = e->round_blocks > 0 ? align_up(out_bytes, 256) + nf_part_bytes + align_up((size_t)n_blocks * 10 * sizeof(float), 256) +
                                                        align_up((size_t)n_blocks * sizeof(int), 256)
                                                  : 0;
    const size_t large_bytes = e->large ? align_up(iq_bytes, 256) + align_up((size_t)n_blocks * sizeof(int), 256) + fast_bytes : 0;
    size_t need = align_up(iq_bytes, 256) + 2 * align_up(out_bytes, 256) + seg_bytes + misc +
                  align_up((size_t)n_blocks * 16, 256) * 2 + align_up((size_t)N * sizeof(float), 256) + large_bytes;
    int rc = ensure_scratch(e, need);
    if (rc != SDR_OK) return rc;
    unsigned char *p = reinterpret_cast<unsigned char *>(e->d_scratch);
    float *d_iq = reinterpret_cast<float *>(p);
    p += align_up(iq_bytes, 256);
    float *d_spec = reinterpret_cast<float *>(p);
    p += align_up(out_bytes, 256);
    float *d_psd = reinterpret_cast<float *>(p);
    p += align_up(out_bytes, 256);
    Segment *d_segs = reinterpret_cast<Segment *>(p);
    p += seg_bytes;
    WorkParams *d_work = reinterpret_cast<WorkParams *>(p);
    p += 256;
    float *d_floor = reinterpret_cast<float *>(p);
    p += align_up((size_t)n_blocks * 16, 256);
    double *d_var = reinterpret_cast<double *>(p);
    p += align_up((size_t)n_blocks * 16, 256);
    float *d_taps = reinterpret_cast<float *>(p);
    p += 256;
    float *d_cum = reinterpret_cast<float *>(p);
    p += align_up((size_t)N * sizeof(float), 256);
    float2 *d_tmp = reinterpret_cast<float2 *>(p);  // large-block path only
    p += e->large ? align_up(iq_bytes, 256) : 0;
    int *d_block_seg = reinterpret_cast<int *>(p);
    p += e->large ? align_up((size_t)n_blocks * sizeof(int), 256) : 0;
    float *d_spec_round = reinterpret_cast<float *>(p);
    p += e->round_blocks > 0 ? align_up(out_bytes, 256) : 0;
    double2 *d_nf_part = reinterpret_cast<double2 *>(p);
    p += nf_part_bytes;
    float *d_xto = reinterpret_cast<float *>(p);
    p += e->round_blocks > 0 ? align_up((size_t)n_blocks * 10 * sizeof(float), 256) : 0;
    int *d_nf_edge = reinterpret_cast<int *>(p);
    std::vector<Segment> segs(n_blocks);
    for (int b = 0; b < n_blocks; b++) {
        Segment &sg = segs[b];
        sg.iq = d_iq + (size_t)b * 2 * N;
        sg.n_blocks = 1;
        sg.stream = 0;
        sg.work = 0;
        sg.block_out = b;
        sg.flush_idx = 0;  // cumulation goes to the throw-away row
        sg.state_in = -1;
        sg.state_out = 0;
    }
    WorkParams wp;
    wp.edge_width = 0;
    wp.n_listeners = 0;
    wp.listener_off = 0;
    wp.pad = 0;
    for (int i = 0; i < 16; i++) wp.nf_map[i] = (unsigned char)i;
    CK(e, cudaMemcpyAsync(d_iq, iq, iq_bytes, cudaMemcpyHostToDevice, e->s_compute));
    CK(e, cudaMemcpyAsync(d_segs, segs.data(), sizeof(Segment) * (size_t)n_blocks, cudaMemcpyHostToDevice, e->s_compute));
    CK(e, cudaMemcpyAsync(d_work, &wp, sizeof(wp), cudaMemcpyHostToDevice, e->s_compute));
    K1Args a;
    a.segs = d_segs;
    a.n_segs = n_blocks;
    a.works = d_work;
    a.listener_bins = nullptr;
    a.tw1 = e->d_tw1;
    a.tw2 = e->d_tw2;
    a.window = e->d_window;
    a.cum_state = d_cum;
    a.psd_floor = d_floor;
    a.variance = d_var;
    a.taps = d_taps;
    a.tap_stride = 0;
    a.flush_cum = d_cum;
    a.dbg_spectrum = d_spec;
    a.dbg_psd = d_psd;
    if (!e->large) {
        CK(e, launch_k1(e, a, true, e->s_compute));
        e->launches += 1;
    } else {
        std::vector<int> bs(n_blocks);
        for (int b = 0; b < n_blocks; b++) bs[b] = b;
        CK(e, cudaMemcpyAsync(d_block_seg, bs.data(), sizeof(int) * (size_t)n_blocks, cudaMemcpyHostToDevice, e->s_compute));
        CK(e, cudaStreamSynchronize(e->s_compute));  // bs is a stack-lifetime buffer
        {  // one round: the scratch holds the whole call
            const LargeFastBufs lb{d_tmp, d_spec_round, d_nf_part, d_xto, d_nf_edge};
            const std::vector<LargeRound> rounds{LargeRound{0, n_blocks, 0, n_blocks}};
            int nl = 0;
            CK(e, launch_large_fast(e, a, lb, d_block_seg, rounds, n_blocks, true, e->s_compute, &nl));
            e->launches += nl;
        }
    }
    CK(e, cudaMemcpyAsync(spectrum, d_spec, out_bytes, cudaMemcpyDeviceToHost, e->s_compute));
    CK(e, cudaMemcpyAsync(psd, d_psd, out_bytes, cudaMemcpyDeviceToHost, e->s_compute));
    CK(e, cudaStreamSynchronize(e->s_compute));
    return SDR_OK;
}

int sdr_dsp_find_noise_floor(sdr_engine *e, const float *psd, int edge_width, float *min_value, double *variance) {
    if (!e || !psd || !min_value || !variance) return SDR_EINVAL;
    const int N = e->N;
    if (edge_width < 0) {
        e->err = "negative edge_width";
        return SDR_EINVAL;
    }
    std::lock_guard<std::mutex> lk(e->mu);
    CK(e, cudaSetDevice(e->cfg.device));
    int rc = ensure_scratch(e, (size_t)N * sizeof(float) + 256);
    if (rc != SDR_OK) return rc;
    unsigned char *p = reinterpret_cast<unsigned char *>(e->d_scratch);
    double *d_var = reinterpret_cast<double *>(p);
    float *d_min = reinterpret_cast<float *>(p + 8);
    float *d_psd = reinterpret_cast<float *>(p + 256);
    CK(e, cudaMemcpyAsync(d_psd, psd, (size_t)N * sizeof(float), cudaMemcpyHostToDevice, e->s_compute));
    if (nf_edge_supported(N, edge_width)) noise_floor_kernel<<<1, 128, 0, e->s_compute>>>(d_psd, N, edge_width, d_min, d_var);
    else nf_exact_single_kernel<<<1, 1, 0, e->s_compute>>>(d_psd, N, edge_width, d_min, d_var);
    CK(e, cudaGetLastError());
    e->launches += 1;
    CK(e, cudaMemcpyAsync(min_value, d_min, sizeof(float), cudaMemcpyDeviceToHost, e->s_compute));
    CK(e, cudaMemcpyAsync(variance, d_var, sizeof(double), cudaMemcpyDeviceToHost, e->s_compute));
    CK(e, cudaStreamSynchronize(e->s_compute));
    return SDR_OK;
}

int sdr_dsp_find_peaks(sdr_engine *e, const float *cumulation, int cumulation_size, float threshold, sdr_peak *peaks,
                       int max_peaks, int *n_peaks) {
    if (!e || !cumulation || !peaks || !n_peaks || max_peaks < 1 || cumulation_size < 1) return SDR_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    const int N = e->N;
    CK(e, cudaSetDevice(e->cfg.device));
    int rc = ensure_scratch(e, (size_t)N * sizeof(float) + 256 + (size_t)max_peaks * sizeof(sdr_peak));
    if (rc != SDR_OK) return rc;
    unsigned char *p = reinterpret_cast<unsigned char *>(e->d_scratch);
    int *d_n = reinterpret_cast<int *>(p);
    float *d_cum = reinterpret_cast<float *>(p + 256);
    sdr_peak *d_peaks = reinterpret_cast<sdr_peak *>(p + 256 + (size_t)N * sizeof(float));
    CK(e, cudaMemcpyAsync(d_cum, cumulation, (size_t)N * sizeof(float), cudaMemcpyHostToDevice, e->s_compute));
    find_peaks_kernel<<<1, K2_THREADS, 0, e->s_compute>>>(d_cum, N, (float)cumulation_size, threshold, d_peaks, max_peaks, d_n);
    CK(e, cudaGetLastError());
    e->launches += 1;
    int n = 0;
    CK(e, cudaMemcpyAsync(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost, e->s_compute));
    CK(e, cudaStreamSynchronize(e->s_compute));
    *n_peaks = n;
    const int ncopy = n < max_peaks ? n : max_peaks;
    if (ncopy > 0) CK(e, cudaMemcpy(peaks, d_peaks, (size_t)ncopy * sizeof(sdr_peak), cudaMemcpyDeviceToHost));
    return SDR_OK;
}

int sdr_kiwi_decode_iq_bytes(sdr_engine *e, const unsigned char *bytes, int n_bytes, float *out) {
    if (!e || !bytes || !out || n_bytes < 4 || (n_bytes & 3)) return SDR_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    CK(e, cudaSetDevice(e->cfg.device));
    const int n_samples = n_bytes / 4;
    int rc = ensure_scratch(e, (size_t)n_bytes + 256 + (size_t)n_samples * sizeof(float2));
    if (rc != SDR_OK) return rc;
    unsigned char *p = reinterpret_cast<unsigned char *>(e->d_scratch);
    uint32_t *d_raw = reinterpret_cast<uint32_t *>(p);
    float2 *d_out = reinterpret_cast<float2 *>(p + align_up((size_t)n_bytes, 256));
    CK(e, cudaMemcpyAsync(d_raw, bytes, (size_t)n_bytes, cudaMemcpyHostToDevice, e->s_compute));
    kiwi_decode_kernel<<<(n_samples + 255) / 256, 256, 0, e->s_compute>>>(d_raw, n_samples, d_out);
    CK(e, cudaGetLastError());
    e->launches += 1;
    CK(e, cudaMemcpyAsync(out, d_out, (size_t)n_samples * sizeof(float2), cudaMemcpyDeviceToHost, e->s_compute));
    CK(e, cudaStreamSynchronize(e->s_compute));
    return SDR_OK;
}

}  // extern "C"
