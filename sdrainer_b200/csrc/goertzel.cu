// goertzel.cu -- C ABI of the Goertzel / envelope bank (include/sdrgpu.h, sdr_goertzel_*).
// Host side of K3 (k3_goertzel.cuh): filter construction follows dsp.NewGoertzel (dsp/dsp.go:55-75).
#include <cuda_runtime.h>
#include <math.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/sdrgpu.h"
#include "k3_goertzel.cuh"

using namespace sdr;

struct sdr_goertzel_bank {
    sdr_goertzel_config cfg{};
    std::string err;
    std::vector<GoertzelFilter> filters;
    GoertzelFilter *d_filters = nullptr;
    cudaStream_t stream = nullptr;
    // audio scratch
    const float **d_audio_ptrs = nullptr;
    int *d_n_blocks = nullptr;
    float *d_scale = nullptr;
    double *d_magnitude = nullptr;
    uint8_t *d_state = nullptr;
    float *d_audio = nullptr;
    size_t audio_cap = 0;
    // iq scratch
    float2 *d_twiddle = nullptr;
    int twiddle_n = 0;
    float *d_iq = nullptr;
    size_t iq_cap = 0;
    int *d_bins = nullptr;
    float *d_out = nullptr;
    size_t bins_cap = 0, out_cap = 0;
    int sm_count = 148;
};

namespace {
std::string g_goertzel_create_error;

#define GCK(b, call)                                                            \
    do {                                                                        \
        cudaError_t _st = (call);                                               \
        if (_st != cudaSuccess) {                                               \
            (b)->err = std::string(#call) + ": " + cudaGetErrorString(_st);     \
            return SDR_ECUDA;                                                   \
        }                                                                       \
    } while (0)

// math.Round: half away from zero
double go_round(double x) { return round(x); }

int calculate_blocksize(double pitch, int sample_rate, double ratio) {  // dsp/dsp.go:72-75
    const double min_blocksize = go_round((double)sample_rate / pitch);
    return (int)go_round((ratio * (double)sample_rate) / min_blocksize) * (int)min_blocksize;
}
}  // namespace

extern "C" {

const char *sdr_goertzel_last_error(const sdr_goertzel_bank *b) { return b ? b->err.c_str() : g_goertzel_create_error.c_str(); }

int sdr_goertzel_create(const sdr_goertzel_config *cfg, sdr_goertzel_bank **out) {
    if (!cfg || !out) return SDR_EINVAL;
    *out = nullptr;
    if (cfg->n_filters < 1 || cfg->sample_rate <= 0 || cfg->max_blocks < 1 || (cfg->n_filters > 0 && !cfg->pitch)) {
        g_goertzel_create_error = "bad goertzel configuration";
        return SDR_EINVAL;
    }
    sdr_goertzel_bank *b = new (std::nothrow) sdr_goertzel_bank();
    if (!b) return SDR_ENOMEM;
    b->cfg = *cfg;
    b->cfg.pitch = nullptr;
    const double ratio = cfg->blocksize_ratio > 0 ? cfg->blocksize_ratio : 0.005;
    const double two_pi = 2 * 3.14159265358979323846;
    for (int i = 0; i < cfg->n_filters; i++) {
        const double pitch = cfg->pitch[i];
        if (!(pitch > 0)) {
            g_goertzel_create_error = "pitch must be positive";
            delete b;
            return SDR_EINVAL;
        }
        GoertzelFilter f;
        f.blocksize = calculate_blocksize(pitch, cfg->sample_rate, ratio);
        if (f.blocksize < 1) {
            g_goertzel_create_error = "pitch/sample rate give an empty Goertzel block";
            delete b;
            return SDR_EINVAL;
        }
        const int bin_index = (int)(0.5 + ((double)f.blocksize * pitch / (double)cfg->sample_rate));  // dsp/dsp.go:57
        const double omega = two_pi * (double)bin_index / (double)f.blocksize;
        f.coeff = 2 * cos(omega);
        f.magnitude_limit_low = (double)f.blocksize / 2;
        f.magnitude_limit = 0;
        f.magnitude_threshold = 0.75;  // dsp.DefaultMagnitudeThreshold
        f.pad = 0;
        b->filters.push_back(f);
    }
    auto fail = [&](int code) {
        g_goertzel_create_error = b->err;
        sdr_goertzel_destroy(b);
        return code;
    };
#define GCKC(call)                                                          \
    do {                                                                    \
        cudaError_t _st = (call);                                           \
        if (_st != cudaSuccess) {                                           \
            b->err = std::string(#call) + ": " + cudaGetErrorString(_st);   \
            return fail(SDR_ECUDA);                                         \
        }                                                                   \
    } while (0)
    GCKC(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    GCKC(cudaGetDeviceProperties(&prop, cfg->device));
    b->sm_count = prop.multiProcessorCount;
    GCKC(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    const size_t nf = (size_t)cfg->n_filters, mb = (size_t)cfg->max_blocks;
    GCKC(cudaMalloc((void **)&b->d_filters, nf * sizeof(GoertzelFilter)));
    GCKC(cudaMemcpy(b->d_filters, b->filters.data(), nf * sizeof(GoertzelFilter), cudaMemcpyHostToDevice));
    GCKC(cudaMalloc((void **)&b->d_audio_ptrs, nf * sizeof(float *)));
    GCKC(cudaMalloc((void **)&b->d_n_blocks, nf * sizeof(int)));
    GCKC(cudaMalloc((void **)&b->d_scale, nf * sizeof(float)));
    GCKC(cudaMalloc((void **)&b->d_magnitude, nf * mb * sizeof(double)));
    GCKC(cudaMalloc((void **)&b->d_state, nf * mb));
#undef GCKC
    *out = b;
    return SDR_OK;
}

void sdr_goertzel_destroy(sdr_goertzel_bank *b) {
    if (!b) return;
    cudaSetDevice(b->cfg.device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    cudaFree(b->d_filters);
    cudaFree(b->d_audio_ptrs);
    cudaFree(b->d_n_blocks);
    cudaFree(b->d_scale);
    cudaFree(b->d_magnitude);
    cudaFree(b->d_state);
    cudaFree(b->d_audio);
    cudaFree(b->d_twiddle);
    cudaFree(b->d_iq);
    cudaFree(b->d_bins);
    cudaFree(b->d_out);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

int sdr_goertzel_blocksize(const sdr_goertzel_bank *b, int filter) {
    if (!b || filter < 0 || filter >= (int)b->filters.size()) return SDR_EINVAL;
    return b->filters[filter].blocksize;
}

int sdr_goertzel_process_audio(sdr_goertzel_bank *b, const float *const *audio, const int *n_blocks, const float *scale,
                               double max_scale, double *magnitude, uint8_t *state, int out_stride) {
    if (!b || !audio || !n_blocks || !magnitude || !state) return SDR_EINVAL;
    const int nf = (int)b->filters.size();
    int max_nb = 0;
    size_t total = 0;
    for (int i = 0; i < nf; i++) {
        if (n_blocks[i] < 0 || n_blocks[i] > b->cfg.max_blocks || out_stride < n_blocks[i] || (n_blocks[i] > 0 && !audio[i])) {
            b->err = "bad block count for filter " + std::to_string(i);
            return SDR_EINVAL;
        }
        if (n_blocks[i] > max_nb) max_nb = n_blocks[i];
        total += (size_t)n_blocks[i] * b->filters[i].blocksize;
    }
    if (max_nb == 0) return SDR_OK;
    GCK(b, cudaSetDevice(b->cfg.device));
    if (total > b->audio_cap) {
        cudaFree(b->d_audio);
        b->d_audio = nullptr;
        b->audio_cap = 0;
        GCK(b, cudaMalloc((void **)&b->d_audio, total * sizeof(float)));
        b->audio_cap = total;
    }
    std::vector<const float *> ptrs(nf);
    std::vector<float> sc(nf);
    size_t off = 0;
    for (int i = 0; i < nf; i++) {
        ptrs[i] = b->d_audio + off;
        const size_t cnt = (size_t)n_blocks[i] * b->filters[i].blocksize;
        if (cnt) GCK(b, cudaMemcpyAsync(b->d_audio + off, audio[i], cnt * sizeof(float), cudaMemcpyHostToDevice, b->stream));
        off += cnt;
        sc[i] = scale ? scale[i] : 1.f;
    }
    GCK(b, cudaMemcpyAsync(b->d_audio_ptrs, ptrs.data(), nf * sizeof(float *), cudaMemcpyHostToDevice, b->stream));
    GCK(b, cudaMemcpyAsync(b->d_n_blocks, n_blocks, nf * sizeof(int), cudaMemcpyHostToDevice, b->stream));
    GCK(b, cudaMemcpyAsync(b->d_scale, sc.data(), nf * sizeof(float), cudaMemcpyHostToDevice, b->stream));
    GoertzelAudioArgs a;
    a.filters = b->d_filters;
    a.audio = b->d_audio_ptrs;
    a.n_blocks = b->d_n_blocks;
    a.scale = b->d_scale;
    a.max_scale = max_scale > 0 ? max_scale : 12.0;  // cw/audio.go:18 defaultMaxScale
    a.magnitude = b->d_magnitude;
    a.state = b->d_state;
    a.out_stride = b->cfg.max_blocks;
    a.n_filters = nf;
    a.max_blocks = b->cfg.max_blocks;
    dim3 grid((max_nb + 127) / 128, nf);
    goertzel_audio_mag_kernel<<<grid, 128, 0, b->stream>>>(a);
    GCK(b, cudaGetLastError());
    goertzel_audio_norm_kernel<<<(nf + 63) / 64, 64, 0, b->stream>>>(a, b->d_filters);
    GCK(b, cudaGetLastError());
    for (int i = 0; i < nf; i++) {
        if (!n_blocks[i]) continue;
        GCK(b, cudaMemcpyAsync(magnitude + (size_t)i * out_stride, b->d_magnitude + (size_t)i * b->cfg.max_blocks,
                               (size_t)n_blocks[i] * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
        GCK(b, cudaMemcpyAsync(state + (size_t)i * out_stride, b->d_state + (size_t)i * b->cfg.max_blocks, (size_t)n_blocks[i],
                               cudaMemcpyDeviceToHost, b->stream));
    }
    GCK(b, cudaStreamSynchronize(b->stream));
    return SDR_OK;
}

int sdr_goertzel_process_iq(sdr_goertzel_bank *b, const float *iq, int mem, int block_size, int n_blocks, const int *bins,
                            int n_bins, float *out_db) {
    if (!b || !iq || !bins || !out_db || n_blocks < 1 || n_bins < 1) return SDR_EINVAL;
    const int N = block_size;
    if (N < 32 || (N & (N - 1)) || (size_t)N * 8 > 200 * 1024) {
        b->err = "block_size must be a power of two, 32..16384 (one block is staged in shared memory)";
        return SDR_EINVAL;
    }
    for (int i = 0; i < n_bins; i++)
        if (bins[i] < 0 || bins[i] >= N) {
            b->err = "bin out of range";
            return SDR_EINVAL;
        }
    GCK(b, cudaSetDevice(b->cfg.device));
    if (b->twiddle_n != N) {
        cudaFree(b->d_twiddle);
        b->d_twiddle = nullptr;
        b->twiddle_n = 0;
        std::vector<float2> tw(N);
        const double two_pi = 2 * 3.14159265358979323846;
        for (int m = 0; m < N; m++) {
            const double ang = -two_pi * (double)m / (double)N;
            tw[m] = make_float2((float)cos(ang), (float)sin(ang));
        }
        GCK(b, cudaMalloc((void **)&b->d_twiddle, (size_t)N * sizeof(float2)));
        GCK(b, cudaMemcpy(b->d_twiddle, tw.data(), (size_t)N * sizeof(float2), cudaMemcpyHostToDevice));
        b->twiddle_n = N;
        GCK(b, cudaFuncSetAttribute(goertzel_iq_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
        GCK(b, cudaFuncSetAttribute(goertzel_iq_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
        GCK(b, cudaFuncSetAttribute(goertzel_iq_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    }
    const float *d_iq = iq;
    const size_t iq_floats = (size_t)n_blocks * 2 * N;
    if (mem != SDR_MEM_DEVICE) {
        if (iq_floats > b->iq_cap) {
            cudaFree(b->d_iq);
            b->d_iq = nullptr;
            b->iq_cap = 0;
            GCK(b, cudaMalloc((void **)&b->d_iq, iq_floats * sizeof(float)));
            b->iq_cap = iq_floats;
        }
        GCK(b, cudaMemcpyAsync(b->d_iq, iq, iq_floats * sizeof(float), cudaMemcpyHostToDevice, b->stream));
        d_iq = b->d_iq;
    } else if ((uintptr_t)iq & 15) {
        b->err = "device iq must be 16-byte aligned";
        return SDR_EINVAL;
    }
    if ((size_t)n_bins > b->bins_cap) {
        cudaFree(b->d_bins);
        b->d_bins = nullptr;
        b->bins_cap = 0;
        GCK(b, cudaMalloc((void **)&b->d_bins, (size_t)n_bins * sizeof(int)));
        b->bins_cap = n_bins;
    }
    const size_t out_n = (size_t)n_blocks * n_bins;
    if (out_n > b->out_cap) {
        cudaFree(b->d_out);
        b->d_out = nullptr;
        b->out_cap = 0;
        GCK(b, cudaMalloc((void **)&b->d_out, out_n * sizeof(float)));
        b->out_cap = out_n;
    }
    GCK(b, cudaMemcpyAsync(b->d_bins, bins, (size_t)n_bins * sizeof(int), cudaMemcpyHostToDevice, b->stream));
    GoertzelIqArgs a;
    a.iq = d_iq;
    a.twiddle = b->d_twiddle;
    a.bins = b->d_bins;
    a.out_db = b->d_out;
    a.n = N;
    a.n_blocks = n_blocks;
    a.n_bins = n_bins;
    a.db_offset = (float)(10.0 * log10(20.0 / ((double)N * (double)N)));
    // listeners per lane: 32 * LPT listeners per pass over the staged block
    const int lpt = n_bins <= 32 ? 1 : n_bins <= 64 ? 2 : 4;
    const size_t smem = (size_t)N * 8 + (size_t)(K3_THREADS / 32) * 32 * lpt * sizeof(float2);
    int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    const int grid = n_blocks < per_sm * b->sm_count ? n_blocks : per_sm * b->sm_count;
    if (lpt == 1) goertzel_iq_kernel<1><<<grid, K3_THREADS, smem, b->stream>>>(a);
    else if (lpt == 2) goertzel_iq_kernel<2><<<grid, K3_THREADS, smem, b->stream>>>(a);
    else goertzel_iq_kernel<4><<<grid, K3_THREADS, smem, b->stream>>>(a);
    GCK(b, cudaGetLastError());
    GCK(b, cudaMemcpyAsync(out_db, b->d_out, out_n * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    GCK(b, cudaStreamSynchronize(b->stream));
    return SDR_OK;
}

}  // extern "C"
