/*
 * sdr_oracle.c -- CPU ORACLE (test infrastructure, never shipped, never on the product path).
 * See sdr_oracle.h for scope and parity-pinning status.  All cites are into /root/reference.
 *
 * Arithmetic rules followed throughout:
 *   - Go on amd64 never contracts a*b+c: build with -ffp-contract=off.
 *   - T(x) with T=float32 rounds to nearest-even at every cast site the reference has.
 *   - int(x) truncates toward zero; out-of-range / NaN gives INT64_MIN (CVTTSD2SQ).
 */
#include "sdr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ========================================================================================== */
/* Go math restatements (Go stdlib src/math/log.go, log10.go, frexp.go -- published algorithm;   */
/* the Go toolchain is not present in this image so these are restated, not linked).            */
/* ========================================================================================== */

double orc_go_log(double x) {
    const double Ln2Hi = 6.93147180369123816490e-01; /* 3fe62e42 fee00000 */
    const double Ln2Lo = 1.90821492927058770002e-10; /* 3dea39ef 35793c76 */
    const double L1 = 6.666666666666735130e-01;
    const double L2 = 3.999999999940941908e-01;
    const double L3 = 2.857142874366239149e-01;
    const double L4 = 2.222219843214978396e-01;
    const double L5 = 1.818357216161805012e-01;
    const double L6 = 1.531383769920937332e-01;
    const double L7 = 1.479819860511658591e-01;
    if (isnan(x) || (isinf(x) && x > 0)) return x;
    if (x < 0) return NAN;
    if (x == 0) return -INFINITY;
    int ki;
    double f1 = frexp(x, &ki);
    if (f1 < 0.70710678118654752440 /* Sqrt2/2 */) {
        f1 *= 2;
        ki--;
    }
    double f = f1 - 1;
    double k = (double)ki;
    double s = f / (2 + f);
    double s2 = s * s;
    double s4 = s2 * s2;
    double t1 = s2 * (L1 + s4 * (L3 + s4 * (L5 + s4 * L7)));
    double t2 = s4 * (L2 + s4 * (L4 + s4 * L6));
    double R = t1 + t2;
    double hfsq = 0.5 * f * f;
    return k * Ln2Hi - ((hfsq - (s * (hfsq + R) + k * Ln2Lo)) - f);
}

double orc_go_log2(double x) {
    /* src/math/log10.go: log2 */
    int e;
    double frac = frexp(x, &e);
    if (isnan(x) || isinf(x) || x == 0) {
        /* Frexp returns (x, 0) for these; fall through to Log semantics */
        return orc_go_log(x) * (1.0 / 0.693147180559945309417232121458176568);
    }
    if (frac == 0.5) return (double)(e - 1);
    return orc_go_log(frac) * (1.0 / 0.693147180559945309417232121458176568) + (double)e;
}

double orc_go_log10(double x) {
    /* src/math/log10.go: log10(x) = math.Log2(x) * (Ln2 / Ln10) */
    const double k = 0.693147180559945309417232121458176568 / 2.30258509299404568401799145468436421;
    return orc_go_log2(x) * k;
}

int64_t orc_go_int(double x) {
    if (isnan(x) || x >= 9223372036854775808.0 || x < -9223372036854775808.0) return INT64_MIN;
    return (int64_t)x;
}

static inline int64_t wrap_add(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }

/* ========================================================================================== */
/* go-dsp fft (github.com/mjibson/go-dsp/fft radix2.go, pinned go.mod:22) -- restated.          */
/* radix-2 DIT: bit-reversal reorder, then log2(n) stages t[idx]=r[idx]+w*r[idx2],              */
/* t[idx2]=r[idx]-w*r[idx2] with w = factors[blocks*j]; the factor table for size i is built    */
/* from the table of size i/2 at even slots and Sincos(-2*Pi/i*n) at odd slots, seeded by the   */
/* exact 4-point table {1,-i,-1,i}.                                                            */
/* ========================================================================================== */

#define ORC_MAX_LOG2 20
static double *g_factors[ORC_MAX_LOG2 + 1]; /* g_factors[l] has 2*(1<<l) doubles */

static const double *get_factors(int log2n) {
    if (g_factors[log2n]) return g_factors[log2n];
    if (!g_factors[2]) {
        double *f4 = (double *)malloc(8 * sizeof(double));
        f4[0] = 1; f4[1] = 0;
        f4[2] = 0; f4[3] = -1;
        f4[4] = -1; f4[5] = 0;
        f4[6] = 0; f4[7] = 1;
        g_factors[2] = f4;
    }
    for (int l = 3; l <= log2n; l++) {
        if (g_factors[l]) continue;
        int i = 1 << l;
        double *f = (double *)malloc((size_t)2 * i * sizeof(double));
        const double *p = g_factors[l - 1];
        for (int n = 0, j = 0; n < i; n += 2, j++) {
            f[2 * n] = p[2 * j];
            f[2 * n + 1] = p[2 * j + 1];
        }
        for (int n = 1; n < i; n += 2) {
            /* -2 * math.Pi / float64(i) * float64(n) : (const / i) * n */
            double a = -2 * M_PI / (double)i * (double)n;
            f[2 * n] = cos(a);
            f[2 * n + 1] = sin(a);
        }
        g_factors[l] = f;
    }
    return g_factors[log2n];
}

static unsigned reverse_bits(unsigned v, int s) {
    unsigned r = 0;
    for (int i = 0; i < s; i++) {
        r = (r << 1) | (v & 1u);
        v >>= 1;
    }
    return r;
}

int orc_fft(double *x, int n) {
    if (n <= 1) return 0;
    int log2n = 0;
    while ((1 << log2n) < n) log2n++;
    if ((1 << log2n) != n || log2n > ORC_MAX_LOG2) return -1;
    if (n == 2) {
        double a0 = x[0], a1 = x[1], b0 = x[2], b1 = x[3];
        x[0] = a0 + b0; x[1] = a1 + b1; x[2] = a0 - b0; x[3] = a1 - b1;
        return 0;
    }
    const double *factors = get_factors(log2n);
    double *r = (double *)malloc((size_t)2 * n * sizeof(double));
    double *t = (double *)malloc((size_t)2 * n * sizeof(double));
    for (unsigned i = 0; i < (unsigned)n; i++) {
        unsigned j = reverse_bits(i, log2n);
        r[2 * j] = x[2 * i];
        r[2 * j + 1] = x[2 * i + 1];
    }
    for (int stage = 2; stage <= n; stage <<= 1) {
        int blocks = n / stage;
        int s_2 = stage / 2;
        for (int nb = 0; nb < n; nb += stage) {
            if (stage != 2) {
                for (int j = 0; j < s_2; j++) {
                    int idx = j + nb, idx2 = idx + s_2;
                    double ar = r[2 * idx], ai = r[2 * idx + 1];
                    double br = r[2 * idx2], bi = r[2 * idx2 + 1];
                    double wr = factors[2 * (blocks * j)], wi = factors[2 * (blocks * j) + 1];
                    /* Go complex128 multiply: (br*wr - bi*wi) + (br*wi + bi*wr)i */
                    double pr = br * wr - bi * wi;
                    double pi = br * wi + bi * wr;
                    t[2 * idx] = ar + pr; t[2 * idx + 1] = ai + pi;
                    t[2 * idx2] = ar - pr; t[2 * idx2 + 1] = ai - pi;
                }
            } else {
                int n1 = nb + 1;
                double ar = r[2 * nb], ai = r[2 * nb + 1];
                double br = r[2 * n1], bi = r[2 * n1 + 1];
                t[2 * nb] = ar + br; t[2 * nb + 1] = ai + bi;
                t[2 * n1] = ar - br; t[2 * n1 + 1] = ai - bi;
            }
        }
        double *tmp = r; r = t; t = tmp;
    }
    memcpy(x, r, (size_t)2 * n * sizeof(double));
    free(r);
    free(t);
    return 0;
}

/* ========================================================================================== */
/* dsp/fft.go                                                                                   */
/* ========================================================================================== */

int orc_bin_to_spectrum_index(int bin, int block_size) { /* dsp/fft.go:54-57 */
    int center_bin = block_size / 2;
    return (bin + center_bin) % block_size;
}

float orc_psd(double re, double im) { /* dsp/fft.go:71-73: T(Pow(re,2)+Pow(im,2)); Pow(x,2)==x*x */
    return (float)(re * re + im * im);
}

float orc_magnitude_in_db(double re, double im, int block_size) { /* dsp/fft.go:79-81 */
    double bs = (double)block_size;
    return (float)(10.0 * orc_go_log10(20.0 * (double)orc_psd(re, im) / (bs * bs)));
}

float orc_psd_value_in_db(float psd_value, int block_size) { /* dsp/fft.go:83-85 */
    double bs = (double)block_size;
    return (float)(10.0 * orc_go_log10(20.0 * (double)psd_value / (bs * bs)));
}

int orc_iq_to_spectrum_and_psd(const float *iq, int n, const float *window, float *spectrum, float *psd) {
    /* dsp/fft.go:23-37; setSamplesFromIQ :59-69; projection rx/receiver.go:376-378 */
    double *s = (double *)malloc((size_t)2 * n * sizeof(double));
    if (!s) return -1;
    for (int i = 0; i < n; i++) {
        float fi = iq[2 * i], fq = iq[2 * i + 1];
        if (window) { /* extension (reference has no window): float32 multiply before widening */
            fi = fi * window[i];
            fq = fq * window[i];
        }
        s[2 * i] = (double)fi;
        s[2 * i + 1] = (double)fq;
    }
    if (orc_fft(s, n) != 0) {
        free(s);
        return -1;
    }
    for (int i = 0; i < n; i++) {
        int k = orc_bin_to_spectrum_index(i, n);
        double re = s[2 * i], im = s[2 * i + 1];
        spectrum[k] = orc_magnitude_in_db(re, im, n) + (float)ORC_DBM_SHIFT; /* float32 add */
        psd[k] = orc_psd(re, im);
    }
    free(s);
    return 0;
}

void orc_find_noise_floor(const float *psd, int n, int edge_width, float *min_value_out, double *variance_out) {
    /* dsp/fft.go:215-252, quirks preserved: `to` is the next window's first bin, variance has
     * windowSize+1 terms over windowSize, last window only closes if a later bin exists. */
    int window_size = (n - 2 * edge_width) / 10;
    double min_value = (double)psd[0];
    double sum = 0;
    int count = 0;
    int first = 1;
    int from = 0;
    double result_mean = 0;
    int result_from = 0, result_to = 0;
    for (int i = edge_width; i < n - edge_width; i++) {
        if (count == 0) from = i;
        if (count == window_size) {
            count = 0;
            double mean = sum / (double)window_size;
            if (mean < min_value || first) {
                min_value = mean;
                first = 0;
                result_mean = mean;
                result_from = from;
                result_to = i;
            }
            sum = 0;
        }
        sum += (double)psd[i];
        count++;
    }
    sum = 0;
    for (int i = result_from; i <= result_to; i++) {
        double d = (double)psd[i] - result_mean;
        sum += d * d;
    }
    *variance_out = sum / (double)window_size;
    *min_value_out = (float)min_value;
}

void orc_freqmap_init(orc_freqmap *m, int sample_rate, int block_size, int64_t center_frequency) {
    /* dsp/fft.go:106-124 */
    m->sample_rate = sample_rate;
    m->block_size = block_size;
    m->bin_size = (double)sample_rate / (double)block_size;
    m->center_bin = block_size / 2;
    m->center_frequency = center_frequency;
    m->from_frequency = center_frequency - sample_rate / 2;
}

int64_t orc_freqmap_bin_to_frequency(const orc_freqmap *m, int64_t bin, double location) {
    /* dsp/fft.go:126-130 */
    double location_delta = m->bin_size * location;
    return wrap_add(m->from_frequency, orc_go_int((double)bin * m->bin_size + location_delta));
}

int64_t orc_freqmap_frequency_to_bin(const orc_freqmap *m, int64_t frequency) {
    /* dsp/fft.go:132-135 */
    int64_t bin = orc_go_int(((double)frequency - (double)m->from_frequency) / m->bin_size);
    int64_t hi = m->block_size - 1;
    if (bin > hi) bin = hi;
    if (bin < 0) bin = 0;
    return bin;
}

double orc_peak_center_correction(int64_t bin, const float *spectrum, int n) {
    /* dsp/fft.go:292-309 */
    if (bin <= 0 || bin >= n - 1) return 0;
    double y1 = fabs((double)spectrum[bin - 1]);
    double y2 = fabs((double)spectrum[bin]);
    double y3 = fabs((double)spectrum[bin + 1]);
    return (y3 - y1) / (2 * (2 * y2 - y1 - y3));
}

int orc_find_peaks(orc_peak *peaks, int max_peaks, const float *spectrum, int n, int cumulation_size,
                   float threshold, const orc_freqmap *m) {
    /* dsp/fft.go:254-285 */
    int count = 0;
    int have = 0;
    orc_peak cur;
    memset(&cur, 0, sizeof(cur));
    for (int i = 0; i < n; i++) {
        float value = spectrum[i] / (float)cumulation_size;
        if (!have && value > threshold) {
            memset(&cur, 0, sizeof(cur));
            cur.from = i;
            cur.signal_value = value;
            cur.signal_bin = i;
            have = 1;
        } else if (have && value <= threshold) {
            cur.to = i - 1;
            cur.from_frequency = orc_freqmap_bin_to_frequency(m, cur.from, -0.5);
            cur.to_frequency = orc_freqmap_bin_to_frequency(m, cur.to, 0.5);
            double corr = orc_peak_center_correction(cur.signal_bin, spectrum, n);
            cur.signal_frequency = orc_freqmap_bin_to_frequency(m, cur.signal_bin, corr);
            if (count < max_peaks) peaks[count] = cur;
            count++;
            have = 0;
        } else if (have && cur.signal_value < value) {
            cur.signal_value = value;
            cur.signal_bin = i;
        }
    }
    if (have) {
        cur.to = n - 1;
        cur.from_frequency = orc_freqmap_bin_to_frequency(m, cur.from, -0.5);
        cur.to_frequency = orc_freqmap_bin_to_frequency(m, cur.to, 0.5);
        double corr = orc_peak_center_correction(cur.signal_bin, spectrum, n);
        cur.signal_frequency = orc_freqmap_bin_to_frequency(m, cur.signal_bin, corr);
        if (count < max_peaks) peaks[count] = cur;
        count++;
    }
    return count < max_peaks ? count : max_peaks;
}

/* ========================================================================================== */
/* dsp/dsp.go                                                                                   */
/* ========================================================================================== */

void orc_rolling_mean_init(orc_rolling_mean *m, int n) { /* dsp/dsp.go:249-255 */
    memset(m, 0, sizeof(*m));
    m->len = n;
    m->n = (float)n;
}

float orc_rolling_mean_put(orc_rolling_mean *m, float value) { /* dsp/dsp.go:257-268 */
    m->sum_for_mean -= m->values[m->next];
    m->values[m->next] = value;
    m->sum_for_mean += m->values[m->next];
    m->mean = m->sum_for_mean / m->n;
    m->next = (m->next + 1) % m->len;
    return m->mean;
}

void orc_debouncer_init(orc_debouncer *d, int threshold) {
    memset(d, 0, sizeof(*d));
    d->threshold = threshold;
}

int orc_debouncer_debounce(orc_debouncer *d, int raw_state) { /* dsp/dsp.go:164-182 */
    raw_state = raw_state ? 1 : 0;
    if (d->threshold < 2) return raw_state;
    if (raw_state != d->last_raw_state) d->state_count = 1;
    else d->state_count++;
    d->last_raw_state = raw_state;
    if (d->state_count >= d->threshold) {
        if (raw_state != d->effective_state) d->effective_state = raw_state;
    }
    return d->effective_state;
}

static double go_round(double x) { return round(x); } /* math.Round: half away from zero */

int orc_goertzel_calculate_blocksize(double pitch, int sample_rate, double blocksize_ratio) {
    /* dsp/dsp.go:72-75 */
    double min_blocksize = go_round((double)sample_rate / pitch);
    return (int)orc_go_int(go_round((blocksize_ratio * (double)sample_rate) / min_blocksize)) * (int)orc_go_int(min_blocksize);
}

void orc_goertzel_init(orc_goertzel *g, double pitch, int sample_rate, double blocksize_ratio) {
    /* dsp/dsp.go:55-70 */
    int blocksize = orc_goertzel_calculate_blocksize(pitch, sample_rate, blocksize_ratio);
    int bin_index = (int)orc_go_int(0.5 + ((double)blocksize * pitch / (double)sample_rate));
    double omega = 2 * M_PI * (double)bin_index / (double)blocksize;
    g->pitch = pitch;
    g->sample_rate = sample_rate;
    g->blocksize = blocksize;
    g->coeff = 2 * cos(omega);
    g->magnitude_limit_low = (double)blocksize / 2;
    g->magnitude_limit = 0;
    g->magnitude_threshold = 0.75;
}

double orc_goertzel_magnitude(const orc_goertzel *g, const float *block, int n) {
    /* dsp/dsp.go:98-106 */
    double q0, q1 = 0, q2 = 0;
    for (int i = 0; i < n; i++) {
        q0 = g->coeff * q1 - q2 + (double)block[i];
        q2 = q1;
        q1 = q0;
    }
    return sqrt((q1 * q1) + (q2 * q2) - q1 * q2 * g->coeff);
}

double orc_goertzel_normalized_magnitude(orc_goertzel *g, const float *block, int n) {
    /* dsp/dsp.go:111-123 */
    double magnitude = orc_goertzel_magnitude(g, block, n);
    if (magnitude > g->magnitude_limit_low) {
        g->magnitude_limit = (g->magnitude_limit + ((magnitude - g->magnitude_limit) / 6));
    }
    if (g->magnitude_limit < g->magnitude_limit_low) g->magnitude_limit = g->magnitude_limit_low;
    return magnitude / g->magnitude_limit;
}

int orc_goertzel_detect(orc_goertzel *g, const float *buf, int n, double *magnitude, int *state) {
    /* dsp/dsp.go:127-136 */
    if (n < g->blocksize) {
        *magnitude = 0;
        *state = 0;
        return -1;
    }
    double m = orc_goertzel_normalized_magnitude(g, buf, g->blocksize);
    *magnitude = m;
    *state = m > g->magnitude_threshold;
    return 0;
}

float orc_filter_block_max(const float *block, int n) { /* dsp/dsp.go:19-28 */
    float max = 0;
    for (int i = 0; i < n; i++) {
        float a = (float)fabs((double)block[i]);
        if (a > max) max = a;
    }
    return max;
}

/* ========================================================================================== */
/* cw/decode.go                                                                                 */
/* ========================================================================================== */

/* Morse table.  The reference takes it from github.com/ftl/digimodes/cw (cw.Code,
 * v0.0.0-20231231131023-cffadad68e9e, go.mod:16), which is not vendored.  Pinned by the
 * reference's tests: 'a', '/', U+00A7 (eight dits) (cw/decode_test.go:26-28) and every
 * letter/digit/'ä' used by the nine golden strings (:184-192).  The remaining entries are the
 * published ITU-R M.1677-1 table. */
typedef struct {
    int rune;
    const char *code;
} morse_entry;
static const morse_entry MORSE[] = {
    {'a', ".-"},     {'b', "-..."},   {'c', "-.-."},   {'d', "-.."},    {'e', "."},      {'f', "..-."},
    {'g', "--."},    {'h', "...."},   {'i', ".."},     {'j', ".---"},   {'k', "-.-"},    {'l', ".-.."},
    {'m', "--"},     {'n', "-."},     {'o', "---"},    {'p', ".--."},   {'q', "--.-"},   {'r', ".-."},
    {'s', "..."},    {'t', "-"},      {'u', "..-"},    {'v', "...-"},   {'w', ".--"},    {'x', "-..-"},
    {'y', "-.--"},   {'z', "--.."},   {'0', "-----"},  {'1', ".----"},  {'2', "..---"},  {'3', "...--"},
    {'4', "....-"},  {'5', "....."},  {'6', "-...."},  {'7', "--..."},  {'8', "---.."},  {'9', "----."},
    {'.', ".-.-.-"}, {',', "--..--"}, {'?', "..--.."}, {'/', "-..-."},  {'=', "-...-"},  {'+', ".-.-."},
    {'-', "-....-"}, {'@', ".--.-."}, {':', "---..."}, {';', "-.-.-."}, {'\'', ".----."}, {'"', ".-..-."},
    {'(', "-.--."},  {')', "-.--.-"}, {'!', "-.-.--"}, {'&', ".-..."},  {'_', "..--.-"}, {'$', "...-..-"},
    {0xE4, ".-.-"},  {0xF6, "---."},  {0xFC, "..--"},  {0xA7, "........"},
};
#define N_MORSE ((int)(sizeof(MORSE) / sizeof(MORSE[0])))

int orc_morse_lookup(const unsigned char *symbols) {
    char code[ORC_MAX_SYMBOLS + 1];
    int n = 0;
    while (n < ORC_MAX_SYMBOLS && symbols[n] != 0) {
        code[n] = symbols[n] == 1 ? '.' : '-';
        n++;
    }
    code[n] = 0;
    for (int i = 0; i < N_MORSE; i++)
        if (strcmp(MORSE[i].code, code) == 0) return MORSE[i].rune;
    return -1;
}

static const char *morse_code_for(int rune) {
    for (int i = 0; i < N_MORSE; i++)
        if (MORSE[i].rune == rune) return MORSE[i].code;
    return NULL;
}

static int utf8_next(const char **p) {
    const unsigned char *s = (const unsigned char *)*p;
    int c = s[0];
    if (c < 0x80) { *p += 1; return c; }
    if ((c & 0xE0) == 0xC0 && s[1]) { *p += 2; return ((c & 0x1F) << 6) | (s[1] & 0x3F); }
    if ((c & 0xF0) == 0xE0 && s[1] && s[2]) { *p += 3; return ((c & 0x0F) << 12) | ((s[1] & 0x3F) << 6) | (s[2] & 0x3F); }
    *p += 1;
    return '?';
}

int orc_morse_keying(const char *text, int dit_ticks, unsigned char *out, int cap) {
    /* cw/decode_test.go:255-287 generateStream with defaultTiming {1,3,1,3,7} */
    int n = 0;
    int pending_break = 0; /* 0 none, 3 char break, 7 word break */
    const char *p = text;
#define EMIT(v, cnt)                                      \
    do {                                                  \
        for (int _i = 0; _i < (cnt) * dit_ticks; _i++) {  \
            if (n < cap) out[n] = (v);                    \
            n++;                                          \
        }                                                 \
    } while (0)
    while (*p) {
        int r = utf8_next(&p);
        if (r == ' ') {
            if (pending_break) pending_break = 7;
            continue;
        }
        if (r >= 'A' && r <= 'Z') r += 'a' - 'A';
        const char *code = morse_code_for(r);
        if (!code) continue;
        if (pending_break) EMIT(0, pending_break);
        for (int i = 0; code[i]; i++) {
            if (i > 0) EMIT(0, 1);
            EMIT(1, code[i] == '.' ? 1 : 3);
        }
        pending_break = 3;
    }
    EMIT(0, 3 * 7);
#undef EMIT
    return n < cap ? n : cap;
}

static void at_update(orc_adaptive_threshold *t) { t->threshold = sqrt(t->low * t->high); } /* :413-416 */
static void at_reset(orc_adaptive_threshold *t) { /* cw/decode.go:380-385 */
    t->low = t->preset;
    t->high = 3 * t->low;
    t->last = t->low;
    at_update(t);
}
static void at_init(orc_adaptive_threshold *t, double preset) { /* :371-378 */
    t->preset = preset;
    t->upper_bound = 10;
    at_reset(t);
}
static void at_preset(orc_adaptive_threshold *t, double preset) { /* :387-390 */
    t->preset = preset;
    at_reset(t);
}
static void at_put(orc_adaptive_threshold *t, double duration) { /* :392-411 */
    const double high_factor = 2;
    const double avg_weight = 0.75;
    const double current_weight = 1.0 - 0.75;
    if (duration >= t->low * t->upper_bound) return;
    if (t->last >= duration * high_factor) {
        t->low = avg_weight * t->low + current_weight * duration;
        t->high = avg_weight * t->high + current_weight * t->last;
    } else if (duration >= t->last * high_factor) {
        t->low = avg_weight * t->low + current_weight * t->last;
        t->high = avg_weight * t->high + current_weight * duration;
    }
    t->last = duration;
    at_update(t);
}

static double dec_wpm_to_dit(const orc_decoder *d, double wpm) { /* :191-195 */
    double dit_seconds = 60.0 / (50.0 * wpm);
    return ceil(dit_seconds / d->tick_seconds);
}
static double dec_dit_to_wpm(const orc_decoder *d, double dit_ticks) { /* :197-200 */
    double dit_seconds = dit_ticks * d->tick_seconds;
    return 60.0 / (50.0 * dit_seconds);
}

static void dec_write(orc_decoder *d, int rune) { /* :351-354 */
    d->n_writes++;
    if (d->text_len + 4 >= ORC_TEXT_CAP) return;
    if (rune < 0x80) {
        d->text[d->text_len++] = (char)rune;
    } else if (rune < 0x800) {
        d->text[d->text_len++] = (char)(0xC0 | (rune >> 6));
        d->text[d->text_len++] = (char)(0x80 | (rune & 0x3F));
    } else {
        d->text[d->text_len++] = (char)(0xE0 | (rune >> 12));
        d->text[d->text_len++] = (char)(0x80 | ((rune >> 6) & 0x3F));
        d->text[d->text_len++] = (char)(0x80 | (rune & 0x3F));
    }
    d->text[d->text_len] = 0;
}

static void dec_clear_char(orc_decoder *d) { memset(d->current_char, 0, sizeof(d->current_char)); }

static void dec_decode_current_char(orc_decoder *d) { /* :314-349 */
    if (d->current_char[0] == 0) return;
    if (d->current_char_invalid) {
        d->current_char_invalid = 0;
        dec_clear_char(d);
        dec_write(d, 0xA6);
        return;
    }
    int r = orc_morse_lookup(d->current_char);
    if (r >= 0) dec_write(d, r);
    else dec_write(d, 0xA6);
    dec_clear_char(d);
}

static int dec_append(orc_decoder *d, unsigned char s) { /* cwChar.append :73-81 */
    for (int i = 0; i < ORC_MAX_SYMBOLS; i++) {
        if (d->current_char[i] == 0) {
            d->current_char[i] = s;
            return 1;
        }
    }
    return 0;
}

static void dec_append_symbol(orc_decoder *d, unsigned char s) { /* :306-312 */
    if (!dec_append(d, s)) {
        dec_decode_current_char(d);
        dec_append(d, s);
    }
}

static void dec_preset_wpm(orc_decoder *d, int wpm) { /* :180-185 */
    d->wpm = (double)wpm;
    double dit_time = dec_wpm_to_dit(d, d->wpm);
    at_preset(&d->on_threshold, dit_time);
    at_preset(&d->off_threshold, dit_time);
}

void orc_decoder_init(orc_decoder *d, int sample_rate, int block_size) { /* :131-147 */
    memset(d, 0, sizeof(*d));
    d->tick_seconds = (double)block_size / (double)sample_rate;
    d->wpm = 20;
    d->abort_decode_after_dits = 10;
    double dit_time = dec_wpm_to_dit(d, d->wpm);
    at_init(&d->on_threshold, dit_time);
    at_init(&d->off_threshold, dit_time);
}

static void dec_clear(orc_decoder *d) { /* :172-178 */
    d->decoding = 0;
    dec_clear_char(d);
    d->ticks = 0;
    d->on_start = 0;
    d->off_start = 0;
}

void orc_decoder_reset(orc_decoder *d) { /* :166-170: lastState and currentCharInvalid survive */
    dec_preset_wpm(d, 20);
    dec_clear(d);
    at_reset(&d->on_threshold);
}

static void dec_on_rising_edge(orc_decoder *d, double off_duration) { /* :252-275 */
    if (off_duration < 2.0) return;
    at_put(&d->off_threshold, off_duration);
    double threshold = d->off_threshold.threshold;
    double upper_threshold = 4.5 * d->off_threshold.low;
    if (off_duration >= upper_threshold) {
        dec_decode_current_char(d);
        dec_write(d, ' ');
    } else if (off_duration >= threshold) {
        dec_decode_current_char(d);
    }
}

static void dec_on_falling_edge(orc_decoder *d, double on_duration) { /* :277-297 */
    if (on_duration < 2.0) return;
    at_put(&d->on_threshold, on_duration);
    double threshold = d->on_threshold.threshold;
    double upper_threshold = 2 * d->on_threshold.high;
    if (on_duration >= upper_threshold) {
        d->current_char_invalid = 1;
    } else if (on_duration >= threshold) {
        dec_append_symbol(d, 2);
        d->wpm = (d->wpm + dec_dit_to_wpm(d, d->on_threshold.low)) / 2.0;
    } else {
        dec_append_symbol(d, 1);
    }
}

void orc_decoder_tick(orc_decoder *d, int state) { /* :202-250 (scope frames omitted) */
    state = state ? 1 : 0;
    d->ticks++;
    double now = d->ticks;
    if (state != d->last_state) {
        if (state) {
            d->on_start = now;
            double off_duration = now - d->off_start;
            dec_on_rising_edge(d, off_duration);
        } else {
            d->off_start = now;
            double on_duration = now - d->on_start;
            dec_on_falling_edge(d, on_duration);
        }
        d->decoding = 1;
    }
    d->last_state = state;
    double current_duration = state ? now - d->on_start : now - d->off_start;
    double upper_bound = d->off_threshold.threshold * (double)d->abort_decode_after_dits;
    if (d->decoding && current_duration > upper_bound) {
        d->decoding = 0;
        dec_decode_current_char(d);
    }
}

void orc_decoder_stop(orc_decoder *d) { dec_decode_current_char(d); } /* :356-358 */

void orc_decoder_clear_text(orc_decoder *d) {
    d->text_len = 0;
    d->text[0] = 0;
}

/* ========================================================================================== */
/* cw/spectral.go, cw/audio.go                                                                  */
/* ========================================================================================== */

void orc_spectral_demod_init(orc_spectral_demod *d, int sample_rate, int block_size) { /* cw/spectral.go:26-34 */
    orc_debouncer_init(&d->debouncer, 1);
    orc_decoder_init(&d->decoder, sample_rate, block_size);
}

int orc_spectral_demod_tick(orc_spectral_demod *d, float value, float threshold) { /* cw/spectral.go:48-54 */
    int state = value > threshold;
    int debounced = orc_debouncer_debounce(&d->debouncer, state);
    orc_decoder_tick(&d->decoder, debounced);
    return debounced;
}

void orc_audio_demod_init(orc_audio_demod *d, double pitch, int sample_rate) { /* cw/audio.go:37-58 */
    orc_goertzel_init(&d->filter, pitch, sample_rate, 0.005);
    orc_debouncer_init(&d->debouncer, 3);
    d->max_scale = 12;
    d->scale = 1;
    orc_decoder_init(&d->decoder, sample_rate, d->filter.blocksize);
}

static float truncate_f32(float v) { /* cw/audio.go:213-221 */
    if (v > 1) return 1;
    if (v < -1) return -1;
    return v;
}

int orc_audio_demod_block(orc_audio_demod *d, float *block, int n, double *magnitude, int *state, int *debounced) {
    /* cw/audio.go:184-203 */
    float scale = d->scale;
    if (scale == 0) {
        float max = orc_filter_block_max(block, n);
        double a = 1 / (double)max;
        scale = (float)(a < d->max_scale ? a : d->max_scale); /* math.Min */
        if (isnan(a)) scale = (float)a;
    }
    if (scale != 1) {
        for (int i = 0; i < n; i++) block[i] = truncate_f32(block[i] * scale);
    }
    double m;
    int s;
    if (orc_goertzel_detect(&d->filter, block, n, &m, &s) != 0) return -1;
    int deb = orc_debouncer_debounce(&d->debouncer, s);
    orc_decoder_tick(&d->decoder, deb);
    if (magnitude) *magnitude = m;
    if (state) *state = s;
    if (debounced) *debounced = deb;
    return 0;
}

/* ========================================================================================== */
/* kiwi/client.go:298-308                                                                       */
/* ========================================================================================== */

void orc_kiwi_decode_iq_bytes(const unsigned char *bytes, int n_bytes, float *out) {
    int n = n_bytes / 2;
    for (int i = 0; i < n; i++) {
        uint16_t raw = (uint16_t)((bytes[2 * i] << 8) | bytes[2 * i + 1]);
        out[i] = (float)(int16_t)raw / (float)32767;
    }
}

/* ========================================================================================== */
/* rx/peaks.go, rx/listener.go, rx/receiver.go driver                                           */
/* ========================================================================================== */

enum { PEAK_NONE = 0, PEAK_NEW, PEAK_ACTIVE, PEAK_INACTIVE };

typedef struct ipeak {
    orc_peak peak;
    int state;
    int64_t since_ns;
    struct ipeak *next_alloc;
} ipeak;

typedef struct {
    int id_num; /* numeric suffix of the id, rx/listener.go:151-159 */
    int record; /* index into records */
    ipeak *peak;
    int64_t last_attach_ns;
    int64_t last_write_ns;
    orc_spectral_demod demod;
} listener;

typedef struct {
    int bin;
    int attached;
    int64_t attach_block, detach_block;
    char *text;
    int text_len, text_cap;
    unsigned char *keys;
    int64_t n_keys, keys_cap;
} listener_record;

struct orc_receiver {
    orc_receiver_config cfg;
    orc_freqmap fm;
    int n;
    float *spectrum, *psd, *cumulation, *last_flush;
    orc_peak *peaks_buf;
    int n_last_peaks;
    orc_rolling_mean noise_floor_mean, noise_deviation_mean;
    int cumulation_count;
    int64_t block_index; /* blocks processed so far */
    int64_t now_ns;
    int64_t last_cleanup_s;
    /* peaks table */
    ipeak **bins;
    ipeak *alloc_list;
    /* listener pool */
    listener *listeners; /* active slice, pool_size capacity */
    int n_listeners;
    int *ids; /* id stack */
    int n_ids;
    listener_record *records;
    int n_records, cap_records;
    uint64_t rng;
    orc_block_report report;
};

void orc_receiver_config_default(orc_receiver_config *c, int sample_rate, int block_size) {
    memset(c, 0, sizeof(*c));
    c->sample_rate = sample_rate;
    c->block_size = block_size;
    c->strain_mode = 1;
    c->peak_threshold = ORC_DEFAULT_PEAK_THRESHOLD;
    c->edge_width = ORC_DEFAULT_EDGE_WIDTH;
    c->listener_pool_size = ORC_DEFAULT_LISTENER_POOL_SIZE;
    c->center_frequency = 0;
    c->silence_timeout_s = 20;
    c->attachment_timeout_s = 120;
    c->signal_debounce = 1;
    c->rng_seed = 1;
    c->deterministic_find_next = 1;
    c->window = NULL;
}

orc_receiver *orc_receiver_new(const orc_receiver_config *c) {
    orc_receiver *r = (orc_receiver *)calloc(1, sizeof(*r));
    r->cfg = *c;
    int n = c->block_size;
    r->n = n;
    orc_freqmap_init(&r->fm, c->sample_rate, n, c->center_frequency);
    r->spectrum = (float *)calloc((size_t)n, sizeof(float));
    r->psd = (float *)calloc((size_t)n, sizeof(float));
    r->cumulation = (float *)calloc((size_t)n, sizeof(float));
    r->last_flush = (float *)calloc((size_t)n, sizeof(float));
    r->peaks_buf = (orc_peak *)calloc((size_t)n, sizeof(orc_peak));
    orc_rolling_mean_init(&r->noise_floor_mean, ORC_NOISE_WINDOW);
    orc_rolling_mean_init(&r->noise_deviation_mean, ORC_NOISE_WINDOW);
    r->bins = (ipeak **)calloc((size_t)n, sizeof(ipeak *));
    int pool = c->strain_mode ? c->listener_pool_size : 1; /* rx/receiver.go:113-116 */
    r->cfg.listener_pool_size = pool;
    r->listeners = (listener *)calloc((size_t)pool, sizeof(listener));
    r->ids = (int *)calloc((size_t)pool + 1, sizeof(int));
    for (int i = 0; i < pool; i++) r->ids[i] = pool - i; /* NewIDPool: prefix + (size-i) */
    r->n_ids = pool;
    r->rng = c->rng_seed ? c->rng_seed : 0x9E3779B97F4A7C15ull;
    return r;
}

void orc_receiver_free(orc_receiver *r) {
    if (!r) return;
    for (int i = 0; i < r->n_records; i++) {
        free(r->records[i].text);
        free(r->records[i].keys);
    }
    free(r->records);
    ipeak *p = r->alloc_list;
    while (p) {
        ipeak *nx = p->next_alloc;
        free(p);
        p = nx;
    }
    free(r->bins);
    free(r->listeners);
    free(r->ids);
    free(r->spectrum);
    free(r->psd);
    free(r->cumulation);
    free(r->last_flush);
    free(r->peaks_buf);
    free(r);
}

/* xorshift64* -- shared, documented PRNG standing in for Go's unseeded math/rand */
static uint64_t rng_next(orc_receiver *r) {
    uint64_t x = r->rng;
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    r->rng = x;
    return x * 0x2545F4914F6CDD1Dull;
}

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

static void pt_clear(orc_receiver *r, int64_t from, int64_t to) { /* rx/peaks.go:108-112 */
    for (int64_t i = from < 0 ? 0 : from; i <= (to < r->n - 1 ? to : r->n - 1); i++) r->bins[i] = NULL;
}
static void pt_put_internal(orc_receiver *r, ipeak *p) { /* :102-106 */
    for (int64_t i = p->peak.from < 0 ? 0 : p->peak.from; i <= (p->peak.to < r->n - 1 ? p->peak.to : r->n - 1); i++)
        r->bins[i] = p;
}
static ipeak *pt_new(orc_receiver *r, const orc_peak *p) {
    ipeak *ip = (ipeak *)calloc(1, sizeof(ipeak));
    ip->peak = *p;
    ip->state = PEAK_NEW;
    ip->since_ns = r->now_ns;
    ip->next_alloc = r->alloc_list;
    r->alloc_list = ip;
    return ip;
}
static ipeak *pt_put(orc_receiver *r, const orc_peak *p, int force) { /* Put :73-100, ForcePut :46-71 */
    int64_t clear_from = -1, clear_to = -1;
    for (int64_t i = p->from < 0 ? 0 : p->from; i <= (p->to < r->n - 1 ? p->to : r->n - 1); i++) {
        ipeak *e = r->bins[i];
        if (!e) continue;
        if (!force && (e->state == PEAK_ACTIVE || e->state == PEAK_INACTIVE)) return NULL;
        if (clear_from == -1) clear_from = e->peak.from;
        clear_to = e->peak.to;
    }
    if (clear_from > -1 && clear_to > -1) pt_clear(r, clear_from, clear_to);
    ipeak *ip = pt_new(r, p);
    pt_put_internal(r, ip);
    return ip;
}
static void pt_cleanup(orc_receiver *r) { /* :127-147 */
    int64_t timeout_ns = 120ll * 1000000000ll;
    int i = 0;
    while (i < r->n) {
        ipeak *p = r->bins[i];
        i++;
        if (!p) continue;
        if (p->state == PEAK_ACTIVE) continue;
        if (r->now_ns - p->since_ns < timeout_ns) continue;
        pt_clear(r, p->peak.from, p->peak.to);
        i = (int)p->peak.to + 1;
    }
}
static ipeak *pt_get_internal(orc_receiver *r, const ipeak *p) { /* :161-171 */
    ipeak *ip = r->bins[p->peak.from];
    if (!ip) return NULL;
    if (ip->peak.to != p->peak.to) return NULL;
    return ip;
}
static void pt_activate(orc_receiver *r, ipeak *p) { /* :153-159 */
    ipeak *ip = pt_get_internal(r, p);
    if (!ip) return; /* the reference would nil-deref here */
    if (ip->state != PEAK_NEW && ip->state != PEAK_INACTIVE) return;
    ip->state = PEAK_ACTIVE;
}
static void pt_deactivate(orc_receiver *r, ipeak *p) { /* :173-181 */
    ipeak *ip = pt_get_internal(r, p);
    if (!ip) return;
    if (ip->state != PEAK_ACTIVE) return;
    ip->state = PEAK_INACTIVE;
}
static ipeak *pt_find_next(orc_receiver *r) { /* :183-207 */
    if (!r->cfg.deterministic_find_next) {
        for (int i = 0; i < r->n / 2; i++) {
            int j = (int)(rng_next(r) % (uint64_t)r->n);
            ipeak *p = r->bins[j];
            if (!p) continue;
            if (p->state != PEAK_NEW) continue;
            return p;
        }
    }
    for (int i = 0; i < r->n; i++) {
        ipeak *p = r->bins[i];
        if (!p) continue;
        if (p->state != PEAK_NEW) continue;
        return p;
    }
    return NULL;
}

static orc_peak new_peak_centered_on_bin(orc_receiver *r, int center_bin) { /* rx/receiver.go:491-500 */
    orc_peak p;
    memset(&p, 0, sizeof(p));
    p.from = imax(0, center_bin - ORC_PEAK_PADDING);
    p.to = imin(center_bin + ORC_PEAK_PADDING, r->n - 1);
    p.from_frequency = orc_freqmap_bin_to_frequency(&r->fm, p.from, -0.5);
    p.to_frequency = orc_freqmap_bin_to_frequency(&r->fm, p.to, 0.5);
    p.signal_frequency = p.from_frequency + ((p.to_frequency - p.from_frequency) / 2); /* CenterFrequency */
    return p;
}

static listener *pool_bind_next(orc_receiver *r) { /* rx/listener.go:212-226 + newListener */
    if (r->n_listeners == r->cfg.listener_pool_size) return NULL;
    if (r->n_ids == 0) return NULL;
    int id = r->ids[--r->n_ids];
    listener *l = &r->listeners[r->n_listeners++];
    memset(l, 0, sizeof(*l));
    l->id_num = id;
    l->record = -1;
    orc_spectral_demod_init(&l->demod, r->cfg.sample_rate, r->n);
    orc_debouncer_init(&l->demod.debouncer, r->cfg.signal_debounce);
    l->last_write_ns = r->now_ns; /* NewTextProcessor: lastWrite = clock.Now() */
    return l;
}

static void listener_attach(orc_receiver *r, listener *l, ipeak *p) { /* rx/listener.go:84-93 */
    l->peak = p;
    l->last_attach_ns = r->now_ns;
    orc_decoder_reset(&l->demod.decoder); /* demodulator.Reset */
    l->last_write_ns = r->now_ns;         /* textProcessor.Restart */
    if (r->n_records == r->cap_records) {
        r->cap_records = r->cap_records ? 2 * r->cap_records : 16;
        r->records = (listener_record *)realloc(r->records, (size_t)r->cap_records * sizeof(listener_record));
    }
    listener_record *rec = &r->records[r->n_records];
    memset(rec, 0, sizeof(*rec));
    rec->bin = (int)p->peak.signal_bin;
    rec->attached = 1;
    rec->attach_block = r->block_index;
    rec->detach_block = -1;
    l->record = r->n_records++;
}

static void record_sync_text(orc_receiver *r, listener *l) {
    listener_record *rec = &r->records[l->record];
    orc_decoder *d = &l->demod.decoder;
    if (d->text_len > 0) {
        if (rec->text_len + d->text_len + 1 > rec->text_cap) {
            rec->text_cap = 2 * (rec->text_len + d->text_len + 1);
            rec->text = (char *)realloc(rec->text, (size_t)rec->text_cap);
        }
        memcpy(rec->text + rec->text_len, d->text, (size_t)d->text_len);
        rec->text_len += d->text_len;
        rec->text[rec->text_len] = 0;
        orc_decoder_clear_text(d);
    }
}

int orc_receiver_force_attach(orc_receiver *r, int bin) { /* rx/receiver.go:280-296 */
    listener *l = pool_bind_next(r);
    if (!l) return -1;
    orc_peak p = new_peak_centered_on_bin(r, bin);
    p.signal_bin = bin;
    p.signal_frequency = orc_freqmap_bin_to_frequency(&r->fm, bin, 0);
    p.signal_value = 80;
    ipeak *ip = pt_put(r, &p, 1);
    pt_activate(r, ip);
    listener_attach(r, l, ip);
    return l->record;
}

int orc_receiver_process_block(orc_receiver *r, const float *iq) {
    const int n = r->n;
    const int fs = r->cfg.sample_rate;
    /* manual clock: the block with index b is processed at t = (b+1)*N/fs */
    r->now_ns = (int64_t)(((__int128)(r->block_index + 1) * n * 1000000000ll) / fs);
    int64_t now_s = r->now_ns / 1000000000ll;
    if (now_s > r->last_cleanup_s) { /* cleanupTicker rx/receiver.go:359-363 (1 s) */
        r->last_cleanup_s = now_s;
        pt_cleanup(r);
    }
    orc_block_report *rep = &r->report;
    memset(rep, 0, sizeof(*rep));
    rep->block_index = r->block_index;
    rep->attached_bin = -1;

    /* :379 */
    if (orc_iq_to_spectrum_and_psd(iq, n, r->cfg.window, r->spectrum, r->psd) != 0) return -1;
    /* :381 */
    float psd_noise_floor;
    double noise_variance;
    orc_find_noise_floor(r->psd, n, r->cfg.edge_width, &psd_noise_floor, &noise_variance);
    /* :383 noiseDeviation := noiseDeviationMean.Put(T(float64(PSDValueIndB(T(Sqrt(var)), N)+dBmShift) * 0.25)) */
    float dev_in = (float)((double)(orc_psd_value_in_db((float)sqrt(noise_variance), n) + (float)ORC_DBM_SHIFT) * 0.25);
    float noise_deviation = orc_rolling_mean_put(&r->noise_deviation_mean, dev_in);
    /* :384 */
    float noise_floor = orc_rolling_mean_put(&r->noise_floor_mean, orc_psd_value_in_db(psd_noise_floor, n) + (float)ORC_DBM_SHIFT);
    /* :385 */
    float peak_threshold = r->cfg.peak_threshold + noise_floor;
    float listen_threshold = noise_floor + noise_deviation;
    rep->psd_noise_floor = psd_noise_floor;
    rep->noise_variance = noise_variance;
    rep->noise_floor = noise_floor;
    rep->noise_deviation = noise_deviation;
    rep->peak_threshold = peak_threshold;
    rep->listen_threshold = listen_threshold;

    /* :387-402 */
    int detached[1024];
    int n_detached = 0;
    int64_t silence_ns = (int64_t)(r->cfg.silence_timeout_s * 1e9);
    int64_t attach_ns = (int64_t)(r->cfg.attachment_timeout_s * 1e9);
    for (int i = 0; i < r->n_listeners; i++) {
        listener *l = &r->listeners[i];
        if (!l->peak) continue;
        float signal_value = r->spectrum[l->peak->peak.signal_bin];
        int64_t writes_before = l->demod.decoder.n_writes;
        int key = orc_spectral_demod_tick(&l->demod, signal_value, listen_threshold);
        if (l->demod.decoder.n_writes != writes_before) l->last_write_ns = r->now_ns; /* TextProcessor.Write */
        listener_record *rec = &r->records[l->record];
        if (rec->n_keys == rec->keys_cap) {
            rec->keys_cap = rec->keys_cap ? 2 * rec->keys_cap : 1024;
            rec->keys = (unsigned char *)realloc(rec->keys, (size_t)rec->keys_cap);
        }
        rec->keys[rec->n_keys++] = (unsigned char)key;
        record_sync_text(r, l);
        if (r->cfg.strain_mode) {
            int attachment_exceeded = (r->now_ns - l->last_attach_ns) > attach_ns;
            int silence_exceeded = (r->now_ns - l->last_write_ns) > silence_ns;
            if (attachment_exceeded || silence_exceeded) {
                pt_deactivate(r, l->peak);
                l->peak = NULL; /* Detach */
                rec->attached = 0;
                rec->detach_block = r->block_index;
                if (n_detached < 1024) detached[n_detached++] = l->id_num;
            }
        }
    }
    for (int d = 0; d < n_detached; d++) { /* ListenerPool.release rx/listener.go:234-246 */
        int index = -1;
        for (int i = 0; i < r->n_listeners; i++)
            if (r->listeners[i].id_num == detached[d]) { index = i; break; }
        if (index == -1) continue;
        r->ids[r->n_ids++] = detached[d];
        if (r->n_listeners > 1) r->listeners[index] = r->listeners[r->n_listeners - 1];
        r->n_listeners--;
    }

    /* :404-407 */
    for (int i = 0; i < n; i++) r->cumulation[i] += r->spectrum[i];
    r->cumulation_count++;

    /* :409-460 */
    if (r->cumulation_count == ORC_CUMULATION_SIZE) {
        rep->flushed = 1;
        memcpy(r->last_flush, r->cumulation, (size_t)n * sizeof(float));
        r->n_last_peaks = 0;
        if (r->cfg.strain_mode && r->n_listeners < r->cfg.listener_pool_size) {
            int np = orc_find_peaks(r->peaks_buf, n, r->cumulation, n, ORC_CUMULATION_SIZE, peak_threshold, &r->fm);
            r->n_last_peaks = np;
            rep->n_peaks = np;
            for (int i = 0; i < np; i++) {
                /* newPeakCenteredOnSignal :474-480 */
                orc_peak c = new_peak_centered_on_bin(r, (int)r->peaks_buf[i].signal_bin);
                c.signal_frequency = r->peaks_buf[i].signal_frequency;
                c.signal_value = r->peaks_buf[i].signal_value;
                c.signal_bin = r->peaks_buf[i].signal_bin;
                pt_put(r, &c, 0);
            }
            ipeak *selected = pt_find_next(r);
            if (selected) {
                listener *l = pool_bind_next(r);
                if (l) {
                    pt_activate(r, selected);
                    listener_attach(r, l, selected);
                    rep->attached_bin = (int)selected->peak.signal_bin;
                }
            }
        }
        memset(r->cumulation, 0, (size_t)n * sizeof(float));
        r->cumulation_count = 0;
    }
    r->block_index++;
    return 0;
}

const orc_block_report *orc_receiver_last_report(const orc_receiver *r) { return &r->report; }
const float *orc_receiver_spectrum(const orc_receiver *r) { return r->spectrum; }
const float *orc_receiver_psd(const orc_receiver *r) { return r->psd; }
const float *orc_receiver_cumulation(const orc_receiver *r) { return r->cumulation; }
const float *orc_receiver_last_flush(const orc_receiver *r) { return r->last_flush; }
const orc_peak *orc_receiver_last_peaks(const orc_receiver *r, int *n) {
    *n = r->n_last_peaks;
    return r->peaks_buf;
}
int orc_receiver_listener_count(const orc_receiver *r) { return r->n_records; }
int orc_receiver_listener_bin(const orc_receiver *r, int idx) { return r->records[idx].bin; }
int orc_receiver_listener_attached(const orc_receiver *r, int idx) { return r->records[idx].attached; }
int64_t orc_receiver_listener_attach_block(const orc_receiver *r, int idx) { return r->records[idx].attach_block; }
int64_t orc_receiver_listener_detach_block(const orc_receiver *r, int idx) { return r->records[idx].detach_block; }
const char *orc_receiver_listener_text(const orc_receiver *r, int idx) {
    return r->records[idx].text ? r->records[idx].text : "";
}
const unsigned char *orc_receiver_listener_keys(const orc_receiver *r, int idx, int64_t *n) {
    *n = r->records[idx].n_keys;
    return r->records[idx].keys;
}

/* ========================================================================================== */
/* bulk DSP driver (fixed listener set) -- used by parity tests and as the CPU baseline          */
/* ========================================================================================== */

void orc_stream_state_init(orc_stream_state *s, float *cumulation, int n) {
    orc_rolling_mean_init(&s->noise_floor_mean, ORC_NOISE_WINDOW);
    orc_rolling_mean_init(&s->noise_deviation_mean, ORC_NOISE_WINDOW);
    s->cumulation_count = 0;
    s->cumulation = cumulation;
    memset(cumulation, 0, (size_t)n * sizeof(float));
}

int orc_process_stream(orc_stream_state *st, const float *iq, int n, int64_t n_blocks, const float *window,
                       int edge_width, float peak_threshold_cfg, const int *listener_bins, int n_listeners,
                       const orc_freqmap *fm, double *noise, float *thresholds, float *taps, float *flush_cum,
                       orc_peak *peaks, int max_peaks_per_flush, int *n_peaks_per_flush, float *spectrum_out,
                       float *psd_out) {
    float *spectrum = (float *)malloc((size_t)n * sizeof(float));
    float *psd = (float *)malloc((size_t)n * sizeof(float));
    int64_t flush = 0;
    for (int64_t b = 0; b < n_blocks; b++) {
        const float *blk = iq + (size_t)b * 2 * n;
        if (orc_iq_to_spectrum_and_psd(blk, n, window, spectrum, psd) != 0) {
            free(spectrum);
            free(psd);
            return -1;
        }
        float psd_noise_floor;
        double variance;
        orc_find_noise_floor(psd, n, edge_width, &psd_noise_floor, &variance);
        float dev_in = (float)((double)(orc_psd_value_in_db((float)sqrt(variance), n) + (float)ORC_DBM_SHIFT) * 0.25);
        float noise_deviation = orc_rolling_mean_put(&st->noise_deviation_mean, dev_in);
        float noise_floor = orc_rolling_mean_put(&st->noise_floor_mean, orc_psd_value_in_db(psd_noise_floor, n) + (float)ORC_DBM_SHIFT);
        float peak_threshold = peak_threshold_cfg + noise_floor;
        if (noise) {
            noise[2 * b] = (double)psd_noise_floor;
            noise[2 * b + 1] = variance;
        }
        if (thresholds) {
            thresholds[3 * b] = noise_floor;
            thresholds[3 * b + 1] = noise_deviation;
            thresholds[3 * b + 2] = peak_threshold;
        }
        if (taps)
            for (int l = 0; l < n_listeners; l++) taps[(size_t)b * n_listeners + l] = spectrum[listener_bins[l]];
        if (spectrum_out) memcpy(spectrum_out + (size_t)b * n, spectrum, (size_t)n * sizeof(float));
        if (psd_out) memcpy(psd_out + (size_t)b * n, psd, (size_t)n * sizeof(float));
        for (int i = 0; i < n; i++) st->cumulation[i] += spectrum[i];
        st->cumulation_count++;
        if (st->cumulation_count == ORC_CUMULATION_SIZE) {
            if (flush_cum) memcpy(flush_cum + (size_t)flush * n, st->cumulation, (size_t)n * sizeof(float));
            if (peaks && n_peaks_per_flush && fm) {
                n_peaks_per_flush[flush] = orc_find_peaks(peaks + (size_t)flush * max_peaks_per_flush, max_peaks_per_flush,
                                                          st->cumulation, n, ORC_CUMULATION_SIZE, peak_threshold, fm);
            }
            memset(st->cumulation, 0, (size_t)n * sizeof(float));
            st->cumulation_count = 0;
            flush++;
        }
    }
    free(spectrum);
    free(psd);
    return 0;
}
