"""ctypes binding of oracle/libsdroracle.so -- the CPU ORACLE (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under sdrainer_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsdroracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sdr_oracle.c")
    hdr = os.path.join(_HERE, "sdr_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsdroracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Peak(C.Structure):
    _fields_ = [("from_", C.c_int64), ("to", C.c_int64), ("from_frequency", C.c_int64),
                ("to_frequency", C.c_int64), ("signal_frequency", C.c_int64),
                ("signal_value", C.c_float), ("signal_bin", C.c_int64)]

    def key(self):
        return (self.from_, self.to, self.signal_bin)


class FreqMap(C.Structure):
    _fields_ = [("sample_rate", C.c_int), ("block_size", C.c_int), ("bin_size", C.c_double),
                ("center_bin", C.c_int), ("center_frequency", C.c_int64), ("from_frequency", C.c_int64)]


class RollingMean(C.Structure):
    _fields_ = [("values", C.c_float * 256), ("len", C.c_int), ("n", C.c_float), ("next", C.c_int),
                ("sum_for_mean", C.c_float), ("mean", C.c_float)]


class Debouncer(C.Structure):
    _fields_ = [("threshold", C.c_int), ("effective_state", C.c_int), ("last_raw_state", C.c_int),
                ("state_count", C.c_int)]


class Goertzel(C.Structure):
    _fields_ = [("pitch", C.c_double), ("sample_rate", C.c_int), ("blocksize", C.c_int), ("coeff", C.c_double),
                ("magnitude_limit_low", C.c_double), ("magnitude_limit", C.c_double),
                ("magnitude_threshold", C.c_double)]


class AdaptiveThreshold(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("preset", "upper_bound", "low", "high", "last", "threshold")]


TEXT_CAP = 16384


class Decoder(C.Structure):
    _fields_ = [("tick_seconds", C.c_double), ("ticks", C.c_double), ("last_state", C.c_int),
                ("on_start", C.c_double), ("off_start", C.c_double), ("wpm", C.c_double), ("decoding", C.c_int),
                ("abort_decode_after_dits", C.c_int), ("current_char", C.c_ubyte * 8),
                ("current_char_invalid", C.c_int), ("on_threshold", AdaptiveThreshold),
                ("off_threshold", AdaptiveThreshold), ("text", C.c_char * TEXT_CAP), ("text_len", C.c_int),
                ("n_writes", C.c_int64)]

    def get_text(self) -> str:
        return bytes(self.text[: self.text_len]).decode("utf-8")


class SpectralDemod(C.Structure):
    _fields_ = [("debouncer", Debouncer), ("decoder", Decoder)]


class AudioDemod(C.Structure):
    _fields_ = [("filter", Goertzel), ("debouncer", Debouncer), ("decoder", Decoder), ("max_scale", C.c_double),
                ("scale", C.c_float)]


class ReceiverConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_int), ("block_size", C.c_int), ("strain_mode", C.c_int),
                ("peak_threshold", C.c_float), ("edge_width", C.c_int), ("listener_pool_size", C.c_int),
                ("center_frequency", C.c_int64), ("silence_timeout_s", C.c_double),
                ("attachment_timeout_s", C.c_double), ("signal_debounce", C.c_int), ("rng_seed", C.c_uint64),
                ("deterministic_find_next", C.c_int), ("window", C.POINTER(C.c_float))]


class BlockReport(C.Structure):
    _fields_ = [("block_index", C.c_int64), ("psd_noise_floor", C.c_float), ("noise_variance", C.c_double),
                ("noise_floor", C.c_float), ("noise_deviation", C.c_float), ("peak_threshold", C.c_float),
                ("listen_threshold", C.c_float), ("flushed", C.c_int), ("n_peaks", C.c_int),
                ("attached_bin", C.c_int)]


class StreamState(C.Structure):
    _fields_ = [("noise_floor_mean", RollingMean), ("noise_deviation_mean", RollingMean),
                ("cumulation_count", C.c_int), ("cumulation", C.POINTER(C.c_float))]


_lib = None
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    L.orc_go_log.restype = C.c_double
    L.orc_go_log.argtypes = [C.c_double]
    L.orc_go_log2.restype = C.c_double
    L.orc_go_log2.argtypes = [C.c_double]
    L.orc_go_log10.restype = C.c_double
    L.orc_go_log10.argtypes = [C.c_double]
    L.orc_go_int.restype = C.c_int64
    L.orc_go_int.argtypes = [C.c_double]
    L.orc_fft.argtypes = [_f64p, C.c_int]
    L.orc_bin_to_spectrum_index.argtypes = [C.c_int, C.c_int]
    L.orc_psd.restype = C.c_float
    L.orc_psd.argtypes = [C.c_double, C.c_double]
    L.orc_magnitude_in_db.restype = C.c_float
    L.orc_magnitude_in_db.argtypes = [C.c_double, C.c_double, C.c_int]
    L.orc_psd_value_in_db.restype = C.c_float
    L.orc_psd_value_in_db.argtypes = [C.c_float, C.c_int]
    L.orc_iq_to_spectrum_and_psd.argtypes = [_f32p, C.c_int, _f32p, _f32p, _f32p]
    L.orc_find_noise_floor.argtypes = [_f32p, C.c_int, C.c_int, _f32p, _f64p]
    L.orc_find_noise_floor.restype = None
    L.orc_freqmap_init.argtypes = [C.POINTER(FreqMap), C.c_int, C.c_int, C.c_int64]
    L.orc_freqmap_init.restype = None
    L.orc_freqmap_bin_to_frequency.restype = C.c_int64
    L.orc_freqmap_bin_to_frequency.argtypes = [C.POINTER(FreqMap), C.c_int64, C.c_double]
    L.orc_freqmap_frequency_to_bin.restype = C.c_int64
    L.orc_freqmap_frequency_to_bin.argtypes = [C.POINTER(FreqMap), C.c_int64]
    L.orc_peak_center_correction.restype = C.c_double
    L.orc_peak_center_correction.argtypes = [C.c_int64, _f32p, C.c_int]
    L.orc_find_peaks.argtypes = [C.POINTER(Peak), C.c_int, _f32p, C.c_int, C.c_int, C.c_float, C.POINTER(FreqMap)]
    L.orc_rolling_mean_init.argtypes = [C.POINTER(RollingMean), C.c_int]
    L.orc_rolling_mean_init.restype = None
    L.orc_rolling_mean_put.restype = C.c_float
    L.orc_rolling_mean_put.argtypes = [C.POINTER(RollingMean), C.c_float]
    L.orc_debouncer_init.argtypes = [C.POINTER(Debouncer), C.c_int]
    L.orc_debouncer_init.restype = None
    L.orc_debouncer_debounce.argtypes = [C.POINTER(Debouncer), C.c_int]
    L.orc_goertzel_calculate_blocksize.argtypes = [C.c_double, C.c_int, C.c_double]
    L.orc_goertzel_init.argtypes = [C.POINTER(Goertzel), C.c_double, C.c_int, C.c_double]
    L.orc_goertzel_init.restype = None
    L.orc_goertzel_magnitude.restype = C.c_double
    L.orc_goertzel_magnitude.argtypes = [C.POINTER(Goertzel), _f32p, C.c_int]
    L.orc_goertzel_normalized_magnitude.restype = C.c_double
    L.orc_goertzel_normalized_magnitude.argtypes = [C.POINTER(Goertzel), _f32p, C.c_int]
    L.orc_goertzel_detect.argtypes = [C.POINTER(Goertzel), _f32p, C.c_int, _f64p, _i32p]
    L.orc_filter_block_max.restype = C.c_float
    L.orc_filter_block_max.argtypes = [_f32p, C.c_int]
    L.orc_decoder_init.argtypes = [C.POINTER(Decoder), C.c_int, C.c_int]
    L.orc_decoder_init.restype = None
    for fn in ("orc_decoder_reset", "orc_decoder_stop", "orc_decoder_clear_text"):
        getattr(L, fn).argtypes = [C.POINTER(Decoder)]
        getattr(L, fn).restype = None
    L.orc_decoder_tick.argtypes = [C.POINTER(Decoder), C.c_int]
    L.orc_decoder_tick.restype = None
    L.orc_morse_lookup.argtypes = [C.POINTER(C.c_ubyte)]
    L.orc_morse_keying.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_ubyte), C.c_int]
    L.orc_spectral_demod_init.argtypes = [C.POINTER(SpectralDemod), C.c_int, C.c_int]
    L.orc_spectral_demod_init.restype = None
    L.orc_spectral_demod_tick.argtypes = [C.POINTER(SpectralDemod), C.c_float, C.c_float]
    L.orc_audio_demod_init.argtypes = [C.POINTER(AudioDemod), C.c_double, C.c_int]
    L.orc_audio_demod_init.restype = None
    L.orc_audio_demod_block.argtypes = [C.POINTER(AudioDemod), _f32p, C.c_int, _f64p, _i32p, _i32p]
    L.orc_receiver_config_default.argtypes = [C.POINTER(ReceiverConfig), C.c_int, C.c_int]
    L.orc_receiver_config_default.restype = None
    L.orc_receiver_new.restype = C.c_void_p
    L.orc_receiver_new.argtypes = [C.POINTER(ReceiverConfig)]
    L.orc_receiver_free.argtypes = [C.c_void_p]
    L.orc_receiver_free.restype = None
    L.orc_receiver_force_attach.argtypes = [C.c_void_p, C.c_int]
    L.orc_receiver_process_block.argtypes = [C.c_void_p, _f32p]
    L.orc_receiver_last_report.restype = C.POINTER(BlockReport)
    L.orc_receiver_last_report.argtypes = [C.c_void_p]
    for fn in ("orc_receiver_spectrum", "orc_receiver_psd", "orc_receiver_cumulation", "orc_receiver_last_flush"):
        getattr(L, fn).restype = _f32p
        getattr(L, fn).argtypes = [C.c_void_p]
    L.orc_receiver_last_peaks.restype = C.POINTER(Peak)
    L.orc_receiver_last_peaks.argtypes = [C.c_void_p, _i32p]
    L.orc_receiver_listener_count.argtypes = [C.c_void_p]
    for fn in ("orc_receiver_listener_bin", "orc_receiver_listener_attached"):
        getattr(L, fn).argtypes = [C.c_void_p, C.c_int]
    for fn in ("orc_receiver_listener_attach_block", "orc_receiver_listener_detach_block"):
        getattr(L, fn).argtypes = [C.c_void_p, C.c_int]
        getattr(L, fn).restype = C.c_int64
    L.orc_receiver_listener_text.restype = C.c_char_p
    L.orc_receiver_listener_text.argtypes = [C.c_void_p, C.c_int]
    L.orc_receiver_listener_keys.restype = C.POINTER(C.c_ubyte)
    L.orc_receiver_listener_keys.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
    L.orc_kiwi_decode_iq_bytes.argtypes = [C.c_char_p, C.c_int, _f32p]
    L.orc_kiwi_decode_iq_bytes.restype = None
    L.orc_stream_state_init.argtypes = [C.POINTER(StreamState), _f32p, C.c_int]
    L.orc_stream_state_init.restype = None
    L.orc_process_stream.argtypes = [C.POINTER(StreamState), _f32p, C.c_int, C.c_int64, _f32p, C.c_int, C.c_float,
                                     _i32p, C.c_int, C.POINTER(FreqMap), _f64p, _f32p, _f32p, _f32p,
                                     C.POINTER(Peak), C.c_int, _i32p, _f32p, _f32p]
    _lib = L
    return L


def _fp(a):
    return None if a is None else a.ctypes.data_as(_f32p)


# ---- numpy-friendly wrappers -----------------------------------------------------------------

def fft(x: np.ndarray) -> np.ndarray:
    """go-dsp fft.FFT restatement on a complex128 vector (power-of-two length)."""
    buf = np.ascontiguousarray(x, dtype=np.complex128).copy()
    rc = lib().orc_fft(buf.view(np.float64).ctypes.data_as(_f64p), buf.shape[0])
    if rc != 0:
        raise ValueError("length must be a power of two")
    return buf


def iq_to_spectrum_and_psd(iq: np.ndarray, window: np.ndarray | None = None):
    """dsp/fft.go:23 with the receiver's shiftedMagnitude projection.  iq: 2N float32."""
    iq = np.ascontiguousarray(iq, dtype=np.float32)
    n = iq.shape[0] // 2
    spectrum = np.empty(n, np.float32)
    psd = np.empty(n, np.float32)
    w = None if window is None else np.ascontiguousarray(window, np.float32)
    rc = lib().orc_iq_to_spectrum_and_psd(_fp(iq), n, _fp(w), _fp(spectrum), _fp(psd))
    if rc != 0:
        raise ValueError("bad block size")
    return spectrum, psd


def find_noise_floor(psd: np.ndarray, edge_width: int):
    psd = np.ascontiguousarray(psd, np.float32)
    mn = C.c_float()
    var = C.c_double()
    lib().orc_find_noise_floor(_fp(psd), psd.shape[0], edge_width, C.byref(mn), C.byref(var))
    return np.float32(mn.value), var.value


def freqmap(sample_rate: int, block_size: int, center_frequency: int = 0) -> FreqMap:
    m = FreqMap()
    lib().orc_freqmap_init(C.byref(m), sample_rate, block_size, center_frequency)
    return m


def find_peaks(cum: np.ndarray, threshold: float, fm: FreqMap, cumulation_size: int = 100):
    cum = np.ascontiguousarray(cum, np.float32)
    n = cum.shape[0]
    arr = (Peak * n)()
    cnt = lib().orc_find_peaks(arr, n, _fp(cum), n, cumulation_size, C.c_float(threshold), C.byref(fm))
    return [arr[i] for i in range(cnt)]


def psd_value_in_db(v, n):
    return np.float32(lib().orc_psd_value_in_db(C.c_float(v), n))


class StreamResult:
    pass


def process_stream(iq: np.ndarray, n: int, edge_width: int = 70, peak_threshold: float = 15.0,
                   listener_bins=(), sample_rate: int = 48000, center_frequency: int = 0, window=None,
                   want_spectrum: bool = False, state=None, max_peaks_per_flush: int | None = None):
    """Bulk hot loop over consecutive blocks of one stream (rx/receiver.go:379-460, fixed listeners)."""
    iq = np.ascontiguousarray(iq, np.float32).reshape(-1)
    n_blocks = iq.shape[0] // (2 * n)
    L = lib()
    bins = np.ascontiguousarray(np.asarray(listener_bins, dtype=np.int32))
    nl = int(bins.shape[0])
    fm = freqmap(sample_rate, n, center_frequency)
    if state is None:
        cum = np.zeros(n, np.float32)
        st = StreamState()
        L.orc_stream_state_init(C.byref(st), _fp(cum), n)
        state = (st, cum)
    st, cum = state
    n_flush = (st.cumulation_count + n_blocks) // 100
    mp = max_peaks_per_flush or (n // 2 + 1)
    res = StreamResult()
    res.noise = np.zeros((n_blocks, 2), np.float64)
    res.thresholds = np.zeros((n_blocks, 3), np.float32)
    res.taps = np.zeros((n_blocks, max(nl, 1)), np.float32)
    res.flush_cum = np.zeros((max(n_flush, 1), n), np.float32)
    peaks = (Peak * (max(n_flush, 1) * mp))()
    npk = np.zeros(max(n_flush, 1), np.int32)
    res.spectrum = np.zeros((n_blocks, n), np.float32) if want_spectrum else None
    res.psd = np.zeros((n_blocks, n), np.float32) if want_spectrum else None
    w = None if window is None else np.ascontiguousarray(window, np.float32)
    rc = L.orc_process_stream(C.byref(st), _fp(iq), n, n_blocks, _fp(w), edge_width, C.c_float(peak_threshold),
                              bins.ctypes.data_as(_i32p), nl, C.byref(fm),
                              res.noise.ctypes.data_as(_f64p), _fp(res.thresholds), _fp(res.taps), _fp(res.flush_cum),
                              peaks, mp, npk.ctypes.data_as(_i32p), _fp(res.spectrum), _fp(res.psd))
    if rc != 0:
        raise ValueError("oracle process_stream failed")
    res.taps = res.taps[:, :nl]
    res.flush_cum = res.flush_cum[:n_flush]
    res.peaks = [[peaks[f * mp + i] for i in range(int(npk[f]))] for f in range(n_flush)]
    res.n_flush = n_flush
    res.state = state
    return res


def decode_key_stream(keys, sample_rate=48000, block_size=512, decoder: Decoder | None = None, stop=True) -> str:
    """cw.Decoder fed with a 0/1 key stream (cw/decode_test.go:195-206)."""
    L = lib()
    d = decoder
    if d is None:
        d = Decoder()
        L.orc_decoder_init(C.byref(d), sample_rate, block_size)
    for k in keys:
        L.orc_decoder_tick(C.byref(d), int(k))
    if stop:
        L.orc_decoder_stop(C.byref(d))
    return d.get_text()


def morse_keying(text: str, dit_ticks: int) -> np.ndarray:
    cap = 64 * dit_ticks * (len(text) + 8)
    out = (C.c_ubyte * cap)()
    n = lib().orc_morse_keying(text.encode("utf-8"), dit_ticks, out, cap)
    return np.frombuffer(out, dtype=np.uint8, count=n).copy()
