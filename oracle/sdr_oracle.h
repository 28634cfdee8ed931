/*
 * sdr_oracle.h -- CPU ORACLE for the SDRainer DSP hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's Go arithmetic (ftl/sdrainer), function by
 * function, in the reference's own types (float32 magnitudes, float64/complex128 internals,
 * int64 frequencies).  Each function cites the reference file:line it follows.
 *
 * It is NOT part of the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (sdrainer_b200/, libsdrgpu.so)
 * never links or calls anything in oracle/.
 *
 * PARITY PINNING STATUS
 *   - cw decoder (cw/decode.go): PINNED by the reference's nine golden key streams
 *     (the nine cw/testdata key streams, cw/decode_test.go:184-192), reproduced 9/9 (tests/test_oracle_golden.py).
 *   - BoolDebouncer, binToSpectrumIndex, FrequencyMapping, Goertzel (deterministic rows of
 *     dsp/dsp_test.go, dsp/fft_test.go): PINNED by the reference's own unit-test tables.
 *   - FFT values, spectrum/PSD, FindNoiseFloor, FindPeaks, thresholds: PARITY UNPINNED.
 *     The FFT itself lives in github.com/mjibson/go-dsp v0.0.0-20180508042940-11479a337f12
 *     (go.mod:22), which is not vendored in /root/reference, and no reference test holds a value
 *     for these functions.  The oracle restates go-dsp's published radix-2 algorithm and is
 *     cross-checked against numpy.fft and analytic tones; there is no Go toolchain in this image,
 *     so the reference itself cannot be run to produce fixtures.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: Go/amd64 never fuses a*b+c).
 */
#ifndef SDR_ORACLE_H
#define SDR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants of rx/receiver.go:15-27 -------------------------------------------------- */
#define ORC_CUMULATION_SIZE 100
#define ORC_DBM_SHIFT 120
#define ORC_PEAK_PADDING 0
#define ORC_NOISE_WINDOW 60
#define ORC_DEFAULT_PEAK_THRESHOLD 15
#define ORC_DEFAULT_EDGE_WIDTH 70
#define ORC_DEFAULT_LISTENER_POOL_SIZE 30

/* ---- Go math restatements --------------------------------------------------------------- */
double orc_go_log(double x);   /* Go math.Log (FreeBSD e_log.c port) */
double orc_go_log2(double x);  /* Go math.Log2 */
double orc_go_log10(double x); /* Go math.Log10 = log2(x) * (Ln2/Ln10) */
int64_t orc_go_int(double x);  /* Go int(float64) on amd64 (CVTTSD2SQ) */

/* ---- dsp/fft.go -------------------------------------------------------------------------- */
/* go-dsp fft.FFT for power-of-two n: in-place on interleaved (re,im) float64.  n >= 1. */
int orc_fft(double *x, int n);
/* dsp/fft.go:54-57 */
int orc_bin_to_spectrum_index(int bin, int block_size);
/* dsp/fft.go:71-73 (T=float32) */
float orc_psd(double re, double im);
/* dsp/fft.go:79-81 (T=float32) */
float orc_magnitude_in_db(double re, double im, int block_size);
/* dsp/fft.go:83-85 (T=float32) */
float orc_psd_value_in_db(float psd_value, int block_size);
/* dsp/fft.go:23-37 with the projection of rx/receiver.go:376-378 (MagnitudeIndB + dBmShift).
 * iq: 2n interleaved float32; window: n float32 or NULL (reference: none); spectrum, psd: n. */
int orc_iq_to_spectrum_and_psd(const float *iq, int n, const float *window, float *spectrum, float *psd);
/* dsp/fft.go:215-252 */
void orc_find_noise_floor(const float *psd, int n, int edge_width, float *min_value, double *variance);

typedef struct {
    int64_t from, to;
    int64_t from_frequency, to_frequency;
    int64_t signal_frequency;
    float signal_value;
    int64_t signal_bin;
} orc_peak;

typedef struct {
    int sample_rate, block_size;
    double bin_size;
    int center_bin;
    int64_t center_frequency, from_frequency;
} orc_freqmap;

/* dsp/fft.go:106-135 */
void orc_freqmap_init(orc_freqmap *m, int sample_rate, int block_size, int64_t center_frequency);
int64_t orc_freqmap_bin_to_frequency(const orc_freqmap *m, int64_t bin, double location);
int64_t orc_freqmap_frequency_to_bin(const orc_freqmap *m, int64_t frequency);
/* dsp/fft.go:292-309 */
double orc_peak_center_correction(int64_t bin, const float *spectrum, int n);
/* dsp/fft.go:254-285.  Returns number of peaks written (<= max_peaks). */
int orc_find_peaks(orc_peak *peaks, int max_peaks, const float *spectrum, int n, int cumulation_size,
                   float threshold, const orc_freqmap *m);

/* ---- dsp/dsp.go -------------------------------------------------------------------------- */
typedef struct {
    float values[256];
    int len;
    float n;
    int next;
    float sum_for_mean, mean;
} orc_rolling_mean; /* dsp/dsp.go:239-282 (T=float32), len <= 256 */
void orc_rolling_mean_init(orc_rolling_mean *m, int n);
float orc_rolling_mean_put(orc_rolling_mean *m, float value);

typedef struct {
    int threshold;
    int effective_state, last_raw_state, state_count;
} orc_debouncer; /* dsp/dsp.go:139-182 */
void orc_debouncer_init(orc_debouncer *d, int threshold);
int orc_debouncer_debounce(orc_debouncer *d, int raw_state);

typedef struct {
    double pitch;
    int sample_rate;
    int blocksize;
    double coeff;
    double magnitude_limit_low, magnitude_limit, magnitude_threshold;
} orc_goertzel; /* dsp/dsp.go:34-136 */
int orc_goertzel_calculate_blocksize(double pitch, int sample_rate, double blocksize_ratio);
void orc_goertzel_init(orc_goertzel *g, double pitch, int sample_rate, double blocksize_ratio);
double orc_goertzel_magnitude(const orc_goertzel *g, const float *block, int n);
double orc_goertzel_normalized_magnitude(orc_goertzel *g, const float *block, int n);
/* returns 0 ok / -1 buffer too short (dsp/dsp.go:127-136) */
int orc_goertzel_detect(orc_goertzel *g, const float *buf, int n, double *magnitude, int *state);
float orc_filter_block_max(const float *block, int n); /* dsp/dsp.go:19-28 */

/* ---- cw/decode.go ------------------------------------------------------------------------ */
typedef struct {
    double preset, upper_bound, low, high, last, threshold;
} orc_adaptive_threshold;

#define ORC_MAX_SYMBOLS 8
#define ORC_TEXT_CAP 16384
typedef struct {
    double tick_seconds, ticks;
    int last_state;
    double on_start, off_start, wpm;
    int decoding;
    int abort_decode_after_dits;
    unsigned char current_char[ORC_MAX_SYMBOLS]; /* 0 none, 1 dit, 2 da */
    int current_char_invalid;
    orc_adaptive_threshold on_threshold, off_threshold;
    /* output: UTF-8 text appended here; n_writes counts writeToOutput calls */
    char text[ORC_TEXT_CAP];
    int text_len;
    int64_t n_writes;
} orc_decoder;
void orc_decoder_init(orc_decoder *d, int sample_rate, int block_size); /* cw/decode.go:131-147 */
void orc_decoder_reset(orc_decoder *d);                                 /* :166-170 */
void orc_decoder_tick(orc_decoder *d, int state);                       /* :202-250 */
void orc_decoder_stop(orc_decoder *d);                                  /* :356-358 */
void orc_decoder_clear_text(orc_decoder *d);
/* morse table lookup used by the decoder; returns unicode code point or -1 */
int orc_morse_lookup(const unsigned char *symbols);
/* encode a text to a 0/1 key stream (ticks) at dit_ticks per dit, default 1:3:1:3:7 timing
 * (restates cw/decode_test.go:255-287 generateStream incl. the 3*wordBreak tail).
 * returns number of ticks written (<= cap) */
int orc_morse_keying(const char *utf8_text, int dit_ticks, unsigned char *out, int cap);

/* ---- cw/spectral.go ---------------------------------------------------------------------- */
typedef struct {
    orc_debouncer debouncer;
    orc_decoder decoder;
} orc_spectral_demod;
void orc_spectral_demod_init(orc_spectral_demod *d, int sample_rate, int block_size);
/* cw/spectral.go:48-54; returns the debounced key state */
int orc_spectral_demod_tick(orc_spectral_demod *d, float value, float threshold);

/* ---- cw/audio.go ------------------------------------------------------------------------- */
typedef struct {
    orc_goertzel filter;
    orc_debouncer debouncer;
    orc_decoder decoder;
    double max_scale;
    float scale;
} orc_audio_demod;
void orc_audio_demod_init(orc_audio_demod *d, double pitch, int sample_rate);
/* one iteration of cw/audio.go:184-203 on a full block (modified in place, like the reference).
 * outputs: normalized magnitude, raw state, debounced state */
int orc_audio_demod_block(orc_audio_demod *d, float *block, int n, double *magnitude, int *state, int *debounced);

/* ---- rx/peaks.go, rx/listener.go, rx/receiver.go ----------------------------------------- */
typedef struct orc_receiver orc_receiver;

typedef struct {
    int sample_rate, block_size;
    int strain_mode;        /* 1 = StrainMode, 0 = DecodeMode */
    float peak_threshold;   /* rx/receiver.go:24 default 15 */
    int edge_width;         /* :25 default 70 */
    int listener_pool_size; /* :26 default 30 (reference has no setter; parameterised here) */
    int64_t center_frequency;
    double silence_timeout_s;    /* rx/listener.go:15 default 20 */
    double attachment_timeout_s; /* :16 default 120 */
    int signal_debounce;         /* cw/spectral.go:14 default 1 */
    uint64_t rng_seed;           /* reference uses unseeded math/rand (rx/peaks.go:185) */
    int deterministic_find_next; /* 1: skip the random probe, take the lowest new peak */
    const float *window;         /* NULL = rectangular (reference) */
} orc_receiver_config;

void orc_receiver_config_default(orc_receiver_config *c, int sample_rate, int block_size);
orc_receiver *orc_receiver_new(const orc_receiver_config *c);
void orc_receiver_free(orc_receiver *r);
/* decode-mode/forced attach of a listener at a bin (rx/receiver.go:280-296 ForcePut+Activate+Attach).
 * returns listener slot or -1 */
int orc_receiver_force_attach(orc_receiver *r, int bin);
/* one iteration of the hot loop rx/receiver.go:364-461 on one block of 2n float32 */
int orc_receiver_process_block(orc_receiver *r, const float *iq);

typedef struct {
    int64_t block_index;
    float psd_noise_floor;
    double noise_variance;
    float noise_floor;     /* rolling mean, dB */
    float noise_deviation; /* rolling mean, dB */
    float peak_threshold;  /* r.peakThreshold + noiseFloor */
    float listen_threshold;/* noiseFloor + noiseDeviation */
    int flushed;           /* 1 if this block closed a cumulation window */
    int n_peaks;           /* peaks found at this flush (strain mode, pool available) */
    int attached_bin;      /* bin of the listener attached at this flush or -1 */
} orc_block_report;
const orc_block_report *orc_receiver_last_report(const orc_receiver *r);
const float *orc_receiver_spectrum(const orc_receiver *r);
const float *orc_receiver_psd(const orc_receiver *r);
const float *orc_receiver_cumulation(const orc_receiver *r);     /* running, pre-clear */
const float *orc_receiver_last_flush(const orc_receiver *r);     /* cumulation at last flush */
const orc_peak *orc_receiver_last_peaks(const orc_receiver *r, int *n);
/* listeners: slots 0..pool-1 in attach order since start (never reused by the oracle driver's log) */
int orc_receiver_listener_count(const orc_receiver *r);          /* total ever attached */
int orc_receiver_listener_bin(const orc_receiver *r, int idx);
int orc_receiver_listener_attached(const orc_receiver *r, int idx);
int64_t orc_receiver_listener_attach_block(const orc_receiver *r, int idx);
int64_t orc_receiver_listener_detach_block(const orc_receiver *r, int idx);
const char *orc_receiver_listener_text(const orc_receiver *r, int idx);
/* per-listener key-state log (debounced), one byte per listened block; n out */
const unsigned char *orc_receiver_listener_keys(const orc_receiver *r, int idx, int64_t *n);

/* ---- kiwi/client.go:298-308 -------------------------------------------------------------- */
void orc_kiwi_decode_iq_bytes(const unsigned char *bytes, int n_bytes, float *out);

/* ---- bulk helpers for tests and the CPU baseline ----------------------------------------- */
/* Runs the DSP part of the hot loop (rx/receiver.go:379-407 + flush :409-459 peak scan) over
 * n_blocks consecutive blocks of one stream with a fixed listener set.  Any output may be NULL.
 *   noise[n_blocks*2]   : psdNoiseFloor (f32 widened) , variance
 *   thresholds[n_blocks*3]: noiseFloor, noiseDeviation (rolling means), peakThreshold  (f32)
 *   taps[n_blocks*n_listeners] : spectrum[bin] per listener
 *   flush_cum[(n_blocks/100)*n] : cumulation at each flush
 *   peaks / n_peaks_per_flush  : FindPeaks output per flush (up to max_peaks_per_flush each)
 */
typedef struct {
    orc_rolling_mean noise_floor_mean, noise_deviation_mean;
    int cumulation_count;
    float *cumulation; /* n floats, owned by caller */
} orc_stream_state;
void orc_stream_state_init(orc_stream_state *s, float *cumulation, int n);
int orc_process_stream(orc_stream_state *st, const float *iq, int n, int64_t n_blocks, const float *window,
                       int edge_width, float peak_threshold, const int *listener_bins, int n_listeners,
                       const orc_freqmap *fm, double *noise, float *thresholds, float *taps, float *flush_cum,
                       orc_peak *peaks, int max_peaks_per_flush, int *n_peaks_per_flush, float *spectrum_out,
                       float *psd_out);

#ifdef __cplusplus
}
#endif
#endif
