/*
 * sdrgpu.h -- C ABI of libsdrgpu.so: the B200-native (sm_100a) DSP hot path of SDRainer.
 *
 * The reference (ftl/sdrainer, pure Go) has no FFI seam: rx.Receiver calls dsp.FFT /
 * dsp.FindNoiseFloor / dsp.FindPeaks inline (rx/receiver.go:379,381,411).  This header is the
 * seam a maintainer binds with cgo; each entry point names the reference
 * code it replaces.  Plain C, no exceptions cross the boundary, every call returns an int status
 * (0 = ok, <0 = error; text via sdr_last_error), the library never aborts the process.
 *
 * Threading: every entry point is safe for concurrent callers on one engine (one goroutine per rx.Receiver,
 * rx/receiver.go:145,336: the engine serialises them internally; a blocking sdr_collect does not hold the lock while
 * it waits).  A stream and a ticket have one owner at a time, like rx.Receiver.run owns its state.  Many receivers
 * sharing one engine fill the GPU best through ONE sdr_submit per tick (the dispatcher in host/sdrhost.hpp and
 * go/sdrgpu does that).
 *
 * Memory: `iq` pointers passed to sdr_submit may be device pointers (SDR_MEM_DEVICE, zero copy)
 * or host pointers (SDR_MEM_HOST; pinned memory from sdr_alloc_pinned makes the copy async).
 * cgo rule: C never retains a Go pointer after the call returns -- host IQ is copied (or DMA'd
 * from C-owned pinned memory) before sdr_submit returns control only when SDR_MEM_HOST data lives
 * in memory obtained from sdr_alloc_pinned; otherwise the call copies synchronously.
 */
#ifndef SDRGPU_H
#define SDRGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDR_OK 0
#define SDR_EINVAL (-1)   /* bad argument (programmer error; the Go reference panics, dsp/fft.go:28-30) */
#define SDR_ECUDA (-2)    /* CUDA runtime error, text in sdr_last_error */
#define SDR_ENOMEM (-3)
#define SDR_EBUSY (-4)    /* no free in-flight slot: caller should drop or collect (rx/receiver.go:328-333) */
#define SDR_ENOTREADY (-5)/* non-blocking collect: ticket still running */
#define SDR_ESTATE (-6)

/* constants of rx/receiver.go:15-27 (compile-time in the reference) */
#define SDR_CUMULATION_SIZE 100
#define SDR_DBM_SHIFT 120
#define SDR_NOISE_WINDOW 60

#define SDR_MEM_HOST 0
#define SDR_MEM_DEVICE 1

/* sdr_work.format: sample encoding of `iq` */
#define SDR_FMT_F32 0        /* interleaved float32 I,Q (tci/tci.go:264; kiwi after decodeIQBytes) */
#define SDR_FMT_KIWI_I16BE 1 /* KiwiSDR wire format: interleaved big-endian int16 I,Q; the division by 32767
                                of kiwi/client.go:298-308 is fused into the kernel's load (4 bytes per sample) */

/* sdr_submit flags */
#define SDR_WANT_FLUSH_CUM 0x1   /* copy each flushed cumulation vector back (scope feed, rx/receiver.go:428-457) */
#define SDR_WANT_SPECTRUM 0x2    /* debug/parity: also store spectrum[] and psd[] per block (dsp/fft.go:23 outputs) */
#define SDR_NO_PEAKS 0x4         /* skip FindPeaks at flushes (decode mode / pool full, rx/receiver.go:410) */
#define SDR_NO_D2H 0x8           /* leave results on the device (throughput measurement of the kernels alone) */
#define SDR_NO_TAPS 0x10         /* do not copy the float32 tap values back (keys, thresholds and peaks still are) */
#define SDR_NO_RAW_KEYS 0x20     /* neither write nor copy back the one-byte raw key states: the packed, debounced
                                    key_bits (1 bit per listener and block) are the result */

/* sdr_work.listener_flags[l] */
#define SDR_LISTENER_ACTIVE 0x1  /* an attached listener sits at this position (rx/listener.go:142-147: Listen ticks only
                                    attached listeners); inactive positions produce key bit 0 and keep their state */
#define SDR_LISTENER_RESET 0x2   /* a new listener was bound to this position (rx/listener.go:205-219 creates a new
                                    SpectralDemodulator): its debouncer starts from zero state with this work */

typedef struct sdr_engine sdr_engine;
typedef int64_t sdr_ticket;

typedef struct {
    int device;              /* CUDA device ordinal */
    int block_size;          /* N: complex samples per block; power of two, 512..4096 fused path,
                                8192 and 65536 via the large-N path */
    int max_streams;         /* independent IQ streams (rx.Receiver instances) this engine serves */
    int max_listeners;       /* listener taps per stream (reference pool: 30, rx/receiver.go:26) */
    int max_blocks_per_batch;/* sum of n_blocks over all works of one submit */
    int max_peaks_per_flush; /* capacity of the peak list per flush (<= N/2+1) */
    int n_slots;             /* in-flight batches (double buffering): >= 1 */
    const float *window;     /* N float32 window or NULL = rectangular (the reference applies none) */
    void *cuda_stream;       /* optional cudaStream_t to launch on (NULL: engine-owned streams) */
} sdr_engine_config;

/* One work = consecutive blocks of ONE stream, processed with one listener set and one set of
 * thresholds; equals the body of rx.Receiver.run's `case frame` (rx/receiver.go:364-461) applied to
 * n_blocks queued frames.  Control values are per-work because the reference applies setter
 * closures between frames (rx/receiver.go:166-172,357-358). */
typedef struct {
    int stream;               /* from sdr_stream_open */
    int n_blocks;             /* >= 1 */
    const float *iq;          /* n_blocks * 2N float32, interleaved I,Q (tci/tci.go:264, kiwi/kiwi.go:94) */
    int mem;                  /* SDR_MEM_HOST or SDR_MEM_DEVICE */
    int edge_width;           /* rx.Receiver.edgeWidth (default 70, rx/receiver.go:25), >= 0.  Widths that leave noise
                                 windows of fewer than 9 bins ((N-2e)/10 < 9) take a slow exact replay of
                                 dsp.FindNoiseFloor for that work (same results as the reference, dsp/fft.go:215-252) */
    float peak_threshold;     /* rx.Receiver.peakThreshold (default 15 dB, :24) */
    int n_listeners;          /* attached listeners */
    const int *listener_bins; /* Listener.SignalBin() of each (rx/listener.go:119-124); host memory */
    int format;               /* SDR_FMT_F32 (default 0) or SDR_FMT_KIWI_I16BE; one format per submit */
    int signal_debounce;      /* SpectralDemodulator.SetSignalDebounce (cw/spectral.go:33-35; default 1, and < 2 means
                                 pass-through, dsp/dsp.go:165-167): the BoolDebouncer of every listener position runs on
                                 the device, its state is carried per (stream, position) across submits */
    const uint8_t *listener_flags; /* [n_listeners] SDR_LISTENER_* or NULL (= every position active, none reset); host memory.
                                 Callers that use the debounced key_bits keep a listener at a stable position (its pool
                                 slot) for as long as it is bound */
} sdr_work;

/* dsp.Peak as found by dsp.FindPeaks (dsp/fft.go:254-285), before the host's frequency mapping.
 * y1,y2,y3 = cumulation[bin-1], [bin], [bin+1] for dsp.PeakCenterCorrection (dsp/fft.go:292-309). */
typedef struct {
    int from, to, signal_bin;
    float signal_value;
    float y1, y2, y3;
} sdr_peak;

/* Result view of one ticket.  All pointers are engine-owned host memory, valid until sdr_release. */
typedef struct {
    int n_works, n_blocks, n_flushes, tap_stride, block_size, max_peaks_per_flush;
    const int *work_block_offset;  /* [n_works+1] first block of each work in the arrays below */
    const int *work_flush_offset;  /* [n_works+1] first flush of each work */
    /* per block (dsp.FindNoiseFloor outputs, dsp/fft.go:215-252) */
    const float *psd_noise_floor;  /* [n_blocks] */
    const double *noise_variance;  /* [n_blocks] */
    /* per block, device mirror of rx/receiver.go:383-385 (float32 rolling means over 60 blocks):
       [n_blocks][4] = noiseFloor, noiseDeviation, peakThreshold, listenThreshold(noiseFloor+noiseDeviation) */
    const float *thresholds;
    /* per block and listener: spectrum[l.SignalBin()] (rx/receiver.go:393) */
    const float *taps;             /* [n_blocks][tap_stride] */
    /* per block and listener: value > threshold (cw/spectral.go:49), before debouncing */
    const uint8_t *keys;           /* [n_blocks][tap_stride], NULL with SDR_NO_RAW_KEYS */
    /* per block and listener, packed: the DEBOUNCED key state (cw/spectral.go:50, dsp/dsp.go:164-182) that
       cw.Decoder.Tick consumes -- bit (l % 32) of word [block][l / 32] */
    const uint32_t *key_bits;      /* [n_blocks][key_words] */
    int key_words;                 /* tap_stride / 32 rounded up */
    /* per flush (every SDR_CUMULATION_SIZE blocks of a stream, rx/receiver.go:409) */
    const int *flush_block;        /* [n_flushes] index of the block that closed the window */
    const int *flush_n_peaks;      /* [n_flushes] peaks found (may exceed capacity: list is truncated) */
    const sdr_peak *flush_peaks;   /* [n_flushes][max_peaks_per_flush], in bin order */
    const float *flush_cum;        /* [n_flushes][N] or NULL (SDR_WANT_FLUSH_CUM) */
    /* debug/parity only (SDR_WANT_SPECTRUM) */
    const float *spectrum;         /* [n_blocks][N] or NULL */
    const float *psd;              /* [n_blocks][N] or NULL */
    float gpu_ms;                  /* device time of this ticket's kernels (CUDA events) */
    float k1_ms, k2_ms;            /* split: fused spectral kernel / post kernel */
    int gpu_launches;              /* kernels launched for this ticket */
} sdr_result;

/* ---- library ------------------------------------------------------------------------------ */
const char *sdr_version(void);
int sdr_device_count(void);

/* ---- engine lifecycle ---------------------------------------------------------------------- */
int sdr_engine_create(const sdr_engine_config *cfg, sdr_engine **out);
void sdr_engine_destroy(sdr_engine *e);
const char *sdr_last_error(const sdr_engine *e); /* e may be NULL: last create error */

/* C-owned pinned host memory for the IQ rings (Go fills it through unsafe.Slice) */
int sdr_alloc_pinned(sdr_engine *e, size_t bytes, void **out);
int sdr_free_pinned(sdr_engine *e, void *p);

/* ---- streams: one per rx.Receiver (Start/Stop, rx/receiver.go:130-164) ---------------------- */
int sdr_stream_open(sdr_engine *e, int sample_rate, int *out_stream);
int sdr_stream_close(sdr_engine *e, int stream);
/* clears cumulation and the two rolling means (new Receiver.run, rx/receiver.go:339-346) */
int sdr_stream_reset(sdr_engine *e, int stream);
/* cumulationCount of the stream (blocks since the last flush) */
int sdr_stream_cumulation_count(sdr_engine *e, int stream, int *out);

/* ---- the hot path --------------------------------------------------------------------------- */
/* Asynchronous: queues H2D (if host memory), the fused kernels and D2H on one CUDA stream. */
int sdr_submit(sdr_engine *e, const sdr_work *works, int n_works, int flags, sdr_ticket *out);
/* blocking != 0 waits; otherwise SDR_ENOTREADY while running.  Results stay valid until release. */
int sdr_collect(sdr_engine *e, sdr_ticket t, int blocking, sdr_result *out);
int sdr_release(sdr_engine *e, sdr_ticket t);
/* device pointers of a ticket's result arrays (SDR_NO_D2H consumers, tests) -- same layout as sdr_result */
int sdr_ticket_device_ptrs(sdr_engine *e, sdr_ticket t, void **psd_noise_floor, void **noise_variance,
                           void **thresholds, void **taps, void **keys);
/* kernels launched by this engine since creation (bench.py's gpu_launches) */
int64_t sdr_engine_launch_count(const sdr_engine *e);
/* The post kernels (thresholds, keys, peaks) of a batch may run on an engine-internal stream (SDR_K2_OVERLAP=1).  A
 * caller that supplied cfg.cuda_stream and wants "everything submitted so far" ordered before later work on that
 * stream (e.g. a timing event) calls this: the stream waits for the last post kernel.  No-op otherwise. */
int sdr_engine_fence(sdr_engine *e);
/* name of the spectral kernel the most recent sdr_submit launched (which one serves a block size / batch shape is the
 * engine's choice: DESIGN.md section 4); for benchmarks and logs */
const char *sdr_engine_last_kernel(const sdr_engine *e);

/* ---- dsp-signature-compatible single calls (drop-in correctness, not throughput) ------------- */
/* dsp.FFT.IQToSpectrumAndPSD with the receiver's shiftedMagnitude projection
 * (dsp/fft.go:23-37, rx/receiver.go:376-379).  Host pointers. */
int sdr_dsp_iq_to_spectrum_and_psd(sdr_engine *e, const float *iq, int n_blocks, float *spectrum, float *psd);
/* dsp.FindNoiseFloor (dsp/fft.go:215-252) on a host psd vector of length N */
int sdr_dsp_find_noise_floor(sdr_engine *e, const float *psd, int edge_width, float *min_value, double *variance);
/* dsp.FindPeaks (dsp/fft.go:254-285) on a host cumulation vector of length N; returns count in *n_peaks */
int sdr_dsp_find_peaks(sdr_engine *e, const float *cumulation, int cumulation_size, float threshold,
                       sdr_peak *peaks, int max_peaks, int *n_peaks);

/* kiwi.decodeIQBytes (kiwi/client.go:298-308) on the GPU: n_bytes/2 big-endian int16 -> float32 / 32767, bit-exact
 * (the same conversion the fused SDR_FMT_KIWI_I16BE load performs).  Host pointers. */
int sdr_kiwi_decode_iq_bytes(sdr_engine *e, const unsigned char *bytes, int n_bytes, float *out);

/* ---- Goertzel / envelope bank (dsp.Goertzel, dsp/dsp.go:34-136; cw/audio.go:184-195) --------- */
typedef struct sdr_goertzel_bank sdr_goertzel_bank;
typedef struct {
    int device;
    int sample_rate;
    int n_filters;          /* listeners / pitches */
    const double *pitch;    /* [n_filters] Hz */
    double blocksize_ratio; /* dsp.DefaultBlocksizeRatio = 0.005 */
    int max_blocks;         /* per filter per call */
} sdr_goertzel_config;
int sdr_goertzel_create(const sdr_goertzel_config *cfg, sdr_goertzel_bank **out);
void sdr_goertzel_destroy(sdr_goertzel_bank *b);
const char *sdr_goertzel_last_error(const sdr_goertzel_bank *b);
/* dsp.Goertzel.Blocksize() of filter i (dsp/dsp.go:72-75) */
int sdr_goertzel_blocksize(const sdr_goertzel_bank *b, int filter);
/* Audio path: every filter i gets its own real float32 audio of n_blocks[i]*blocksize(i) samples
 * (audio[i], host memory).  Applies cw/audio.go:184-192 autoscale/clip (scale[i]; 0 = auto,
 * 1 = none), dsp.Goertzel.NormalizedMagnitude with its running magnitudeLimit carried across
 * calls, and Detect's threshold.  Outputs per filter and block: normalized magnitude and state. */
int sdr_goertzel_process_audio(sdr_goertzel_bank *b, const float *const *audio, const int *n_blocks,
                               const float *scale, double max_scale, double *magnitude, uint8_t *state,
                               int out_stride);
/* IQ path (north-star "multi-listener Goertzel"): one complex block of N samples shared by all
 * filters; filter i evaluates DFT bin `bins[i]` (fftshifted index, like Listener.SignalBin) and
 * returns the same dB value K1's tap returns (rx/receiver.go:393).  iq: device or host. */
int sdr_goertzel_process_iq(sdr_goertzel_bank *b, const float *iq, int mem, int block_size, int n_blocks,
                            const int *bins, int n_bins, float *out_db);

#ifdef __cplusplus
}
#endif
#endif
