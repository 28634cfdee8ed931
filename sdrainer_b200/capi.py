"""ctypes binding of libsdrgpu.so (include/sdrgpu.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build

OK, EINVAL, ECUDA, ENOMEM, EBUSY, ENOTREADY, ESTATE = 0, -1, -2, -3, -4, -5, -6
MEM_HOST, MEM_DEVICE = 0, 1
FMT_F32, FMT_KIWI_I16BE = 0, 1
WANT_FLUSH_CUM, WANT_SPECTRUM, NO_PEAKS, NO_D2H, NO_TAPS, NO_RAW_KEYS = 1, 2, 4, 8, 16, 32
LISTENER_ACTIVE, LISTENER_RESET = 1, 2
CUMULATION_SIZE = 100

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_ubyte)


class EngineConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("block_size", C.c_int), ("max_streams", C.c_int), ("max_listeners", C.c_int),
                ("max_blocks_per_batch", C.c_int), ("max_peaks_per_flush", C.c_int), ("n_slots", C.c_int),
                ("window", _f32p), ("cuda_stream", C.c_void_p)]


class Work(C.Structure):
    _fields_ = [("stream", C.c_int), ("n_blocks", C.c_int), ("iq", C.c_void_p), ("mem", C.c_int),
                ("edge_width", C.c_int), ("peak_threshold", C.c_float), ("n_listeners", C.c_int),
                ("listener_bins", _i32p), ("format", C.c_int), ("signal_debounce", C.c_int), ("listener_flags", _u8p)]


class Peak(C.Structure):
    _fields_ = [("from_", C.c_int), ("to", C.c_int), ("signal_bin", C.c_int), ("signal_value", C.c_float),
                ("y1", C.c_float), ("y2", C.c_float), ("y3", C.c_float)]

    def key(self):
        return (self.from_, self.to, self.signal_bin)


PEAK_DTYPE = np.dtype([("from", "<i4"), ("to", "<i4"), ("signal_bin", "<i4"), ("signal_value", "<f4"),
                       ("y1", "<f4"), ("y2", "<f4"), ("y3", "<f4")])


class Result(C.Structure):
    _fields_ = [("n_works", C.c_int), ("n_blocks", C.c_int), ("n_flushes", C.c_int), ("tap_stride", C.c_int),
                ("block_size", C.c_int), ("max_peaks_per_flush", C.c_int),
                ("work_block_offset", _i32p), ("work_flush_offset", _i32p),
                ("psd_noise_floor", _f32p), ("noise_variance", _f64p), ("thresholds", _f32p), ("taps", _f32p),
                ("keys", _u8p), ("key_bits", C.POINTER(C.c_uint32)), ("key_words", C.c_int), ("flush_block", _i32p), ("flush_n_peaks", _i32p), ("flush_peaks", C.POINTER(Peak)),
                ("flush_cum", _f32p), ("spectrum", _f32p), ("psd", _f32p), ("gpu_ms", C.c_float),
                ("k1_ms", C.c_float), ("k2_ms", C.c_float), ("gpu_launches", C.c_int)]


class GoertzelConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("sample_rate", C.c_int), ("n_filters", C.c_int), ("pitch", _f64p),
                ("blocksize_ratio", C.c_double), ("max_blocks", C.c_int)]


# every symbol include/sdrgpu.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "sdr_version", "sdr_device_count", "sdr_engine_create", "sdr_engine_destroy", "sdr_last_error",
    "sdr_alloc_pinned", "sdr_free_pinned", "sdr_stream_open", "sdr_stream_close", "sdr_stream_reset",
    "sdr_stream_cumulation_count", "sdr_submit", "sdr_collect", "sdr_release", "sdr_ticket_device_ptrs",
    "sdr_engine_launch_count", "sdr_engine_fence", "sdr_engine_last_kernel", "sdr_dsp_iq_to_spectrum_and_psd", "sdr_dsp_find_noise_floor",
    "sdr_dsp_find_peaks", "sdr_kiwi_decode_iq_bytes", "sdr_goertzel_create", "sdr_goertzel_destroy", "sdr_goertzel_last_error",
    "sdr_goertzel_blocksize", "sdr_goertzel_process_audio", "sdr_goertzel_process_iq",
]

_lib = None


class SdrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sdrgpu error {code}: {msg}")
        self.code = code


def lib_path() -> str:
    # SDRGPU_LIB selects another build of the same library (kernel-variant experiments); default: the in-tree build
    return os.environ.get("SDRGPU_LIB") or _build.LIB


def lib():
    """Loads libsdrgpu.so.  No fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA hot path has no CPU fallback)")
    L = C.CDLL(path)
    _build.LOADED = True
    L.sdr_version.restype = C.c_char_p
    L.sdr_device_count.restype = C.c_int
    L.sdr_engine_create.argtypes = [C.POINTER(EngineConfig), C.POINTER(C.c_void_p)]
    L.sdr_engine_destroy.argtypes = [C.c_void_p]
    L.sdr_engine_destroy.restype = None
    L.sdr_last_error.argtypes = [C.c_void_p]
    L.sdr_last_error.restype = C.c_char_p
    L.sdr_alloc_pinned.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.sdr_free_pinned.argtypes = [C.c_void_p, C.c_void_p]
    L.sdr_stream_open.argtypes = [C.c_void_p, C.c_int, _i32p]
    L.sdr_stream_close.argtypes = [C.c_void_p, C.c_int]
    L.sdr_stream_reset.argtypes = [C.c_void_p, C.c_int]
    L.sdr_stream_cumulation_count.argtypes = [C.c_void_p, C.c_int, _i32p]
    L.sdr_submit.argtypes = [C.c_void_p, C.POINTER(Work), C.c_int, C.c_int, C.POINTER(C.c_int64)]
    L.sdr_collect.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.POINTER(Result)]
    L.sdr_release.argtypes = [C.c_void_p, C.c_int64]
    L.sdr_ticket_device_ptrs.argtypes = [C.c_void_p, C.c_int64] + [C.POINTER(C.c_void_p)] * 5
    L.sdr_engine_launch_count.argtypes = [C.c_void_p]
    L.sdr_engine_launch_count.restype = C.c_int64
    L.sdr_engine_fence.argtypes = [C.c_void_p]
    L.sdr_engine_last_kernel.argtypes = [C.c_void_p]
    L.sdr_engine_last_kernel.restype = C.c_char_p
    L.sdr_dsp_iq_to_spectrum_and_psd.argtypes = [C.c_void_p, _f32p, C.c_int, _f32p, _f32p]
    L.sdr_dsp_find_noise_floor.argtypes = [C.c_void_p, _f32p, C.c_int, _f32p, _f64p]
    L.sdr_dsp_find_peaks.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.POINTER(Peak), C.c_int, _i32p]
    L.sdr_kiwi_decode_iq_bytes.argtypes = [C.c_void_p, C.c_char_p, C.c_int, _f32p]
    L.sdr_goertzel_create.argtypes = [C.POINTER(GoertzelConfig), C.POINTER(C.c_void_p)]
    L.sdr_goertzel_destroy.argtypes = [C.c_void_p]
    L.sdr_goertzel_destroy.restype = None
    L.sdr_goertzel_last_error.argtypes = [C.c_void_p]
    L.sdr_goertzel_last_error.restype = C.c_char_p
    L.sdr_goertzel_blocksize.argtypes = [C.c_void_p, C.c_int]
    L.sdr_goertzel_process_audio.argtypes = [C.c_void_p, C.POINTER(_f32p), _i32p, _f32p, C.c_double, _f64p, _u8p, C.c_int]
    L.sdr_goertzel_process_iq.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _f32p]
    _lib = L
    return L


def _np_from(ptr, shape, dtype):
    if not ptr:
        return None
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(C.addressof(ptr.contents))
    return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


class BatchResult:
    """Host copy of one ticket's sdr_result."""

    def __init__(self, r: Result):
        nb, nf, ts, n, mp = r.n_blocks, r.n_flushes, r.tap_stride, r.block_size, r.max_peaks_per_flush
        self.n_works, self.n_blocks, self.n_flushes = r.n_works, nb, nf
        self.work_block_offset = _np_from(r.work_block_offset, (r.n_works + 1,), np.int32)
        self.work_flush_offset = _np_from(r.work_flush_offset, (r.n_works + 1,), np.int32)
        self.psd_noise_floor = _np_from(r.psd_noise_floor, (nb,), np.float32)
        self.noise_variance = _np_from(r.noise_variance, (nb,), np.float64)
        self.thresholds = _np_from(r.thresholds, (nb, 4), np.float32)
        self.taps = _np_from(r.taps, (nb, ts), np.float32)
        self.keys = _np_from(r.keys, (nb, ts), np.uint8)
        self.key_bits = _np_from(r.key_bits, (nb, r.key_words), np.uint32)
        self.flush_block = _np_from(r.flush_block, (nf,), np.int32)
        self.flush_n_peaks = _np_from(r.flush_n_peaks, (nf,), np.int32)
        self.flush_peaks = _np_from(r.flush_peaks, (nf, mp), PEAK_DTYPE)
        self.flush_cum = _np_from(r.flush_cum, (nf, n), np.float32)
        self.spectrum = _np_from(r.spectrum, (nb, n), np.float32)
        self.psd = _np_from(r.psd, (nb, n), np.float32)
        self.gpu_ms = float(r.gpu_ms)
        self.k1_ms, self.k2_ms = float(r.k1_ms), float(r.k2_ms)
        self.gpu_launches = int(r.gpu_launches)

    def debounced_keys(self, n_listeners: int) -> np.ndarray:
        """[blocks, n_listeners] uint8 view of the packed, debounced key states"""
        bits = np.unpackbits(self.key_bits.view(np.uint8), axis=1, bitorder="little")
        return bits[:, :n_listeners]

    def peaks(self, flush: int):
        n = min(int(self.flush_n_peaks[flush]), self.flush_peaks.shape[1])
        return self.flush_peaks[flush, :n]


class Engine:
    """Thin RAII wrapper over sdr_engine_* (one caller at a time, like rx.Receiver.run)."""

    def __init__(self, block_size: int, max_streams: int = 1, max_listeners: int = 64, max_blocks_per_batch: int = 4096,
                 max_peaks_per_flush: int = 256, n_slots: int = 2, device: int = 0, window=None, cuda_stream: int = 0):
        self.L = lib()
        cfg = EngineConfig()
        cfg.device, cfg.block_size, cfg.max_streams = device, block_size, max_streams
        cfg.max_listeners, cfg.max_blocks_per_batch = max_listeners, max_blocks_per_batch
        cfg.max_peaks_per_flush, cfg.n_slots = max_peaks_per_flush, n_slots
        self._window = None
        if window is not None:
            self._window = np.ascontiguousarray(window, np.float32)
            cfg.window = self._window.ctypes.data_as(_f32p)
        cfg.cuda_stream = cuda_stream or None
        h = C.c_void_p()
        rc = self.L.sdr_engine_create(C.byref(cfg), C.byref(h))
        if rc != OK:
            raise SdrError(rc, self.L.sdr_last_error(None).decode())
        self.h = h
        self.block_size = block_size
        self._keep = {}

    def close(self):
        if getattr(self, "h", None):
            self.L.sdr_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != OK:
            raise SdrError(rc, self.L.sdr_last_error(self.h).decode())

    def open_stream(self, sample_rate: int) -> int:
        s = C.c_int()
        self._ck(self.L.sdr_stream_open(self.h, sample_rate, C.byref(s)))
        return s.value

    def close_stream(self, s: int):
        self._ck(self.L.sdr_stream_close(self.h, s))

    def reset_stream(self, s: int):
        self._ck(self.L.sdr_stream_reset(self.h, s))

    def cumulation_count(self, s: int) -> int:
        v = C.c_int()
        self._ck(self.L.sdr_stream_cumulation_count(self.h, s, C.byref(v)))
        return v.value

    def alloc_pinned(self, nbytes: int) -> np.ndarray:
        p = C.c_void_p()
        self._ck(self.L.sdr_alloc_pinned(self.h, nbytes, C.byref(p)))
        buf = (C.c_char * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8)
        self._keep[arr.ctypes.data] = p
        return arr

    def free_pinned(self, arr: np.ndarray):
        p = self._keep.pop(arr.ctypes.data)
        self._ck(self.L.sdr_free_pinned(self.h, p))

    def prepare(self, works):
        """Builds the sdr_work array once (for callers that resubmit the same batch shape, e.g. bench.py).
        works: list of dicts(stream, iq (np.float32 array | int device ptr), n_blocks, edge_width,
        peak_threshold, listener_bins)."""
        arr = (Work * len(works))()
        keep = []
        for i, w in enumerate(works):
            iq = w["iq"]
            arr[i].stream = w["stream"]
            fmt = w.get("format", FMT_F32)
            arr[i].format = fmt
            if isinstance(iq, np.ndarray):
                want = np.float32 if fmt == FMT_F32 else np.uint8
                if iq.dtype != want or not iq.flags["C_CONTIGUOUS"]:
                    raise SdrError(EINVAL, "iq must be contiguous float32 (FMT_F32) or uint8 wire bytes (FMT_KIWI_I16BE)")
                arr[i].iq = iq.ctypes.data
                arr[i].mem = MEM_HOST
                per_block = 2 * self.block_size if fmt == FMT_F32 else 4 * self.block_size
                arr[i].n_blocks = w.get("n_blocks", iq.size // per_block)
                keep.append(iq)
            else:
                arr[i].iq = int(iq)
                arr[i].mem = MEM_DEVICE
                arr[i].n_blocks = w["n_blocks"]
            arr[i].edge_width = w.get("edge_width", 70)
            arr[i].peak_threshold = w.get("peak_threshold", 15.0)
            bins = np.ascontiguousarray(np.asarray(w.get("listener_bins", ()), dtype=np.int32))
            keep.append(bins)
            arr[i].n_listeners = bins.size
            arr[i].listener_bins = bins.ctypes.data_as(_i32p)
            arr[i].signal_debounce = w.get("signal_debounce", 1)
            if w.get("listener_flags") is not None:
                lf = np.ascontiguousarray(np.asarray(w["listener_flags"], dtype=np.uint8))
                if lf.size != bins.size:
                    raise SdrError(EINVAL, "listener_flags must have one entry per listener bin")
                keep.append(lf)
                arr[i].listener_flags = lf.ctypes.data_as(_u8p)
        return (arr, len(works), keep)

    def submit_prepared(self, prepared, flags: int = 0) -> int:
        arr, n, keep = prepared
        t = C.c_int64()
        self._ck(self.L.sdr_submit(self.h, arr, n, flags, C.byref(t)))
        self._keep[("t", t.value)] = keep
        return t.value

    def submit(self, works, flags: int = 0) -> int:
        return self.submit_prepared(self.prepare(works), flags)

    def collect_raw(self, ticket: int, blocking: bool = True) -> Result:
        r = Result()
        rc = self.L.sdr_collect(self.h, ticket, 1 if blocking else 0, C.byref(r))
        if rc == ENOTREADY:
            return None
        self._ck(rc)
        return r

    def collect(self, ticket: int, release: bool = True) -> BatchResult:
        r = self.collect_raw(ticket, True)
        out = BatchResult(r)
        if release:
            self.release(ticket)
        return out

    def release(self, ticket: int):
        self._ck(self.L.sdr_release(self.h, ticket))
        self._keep.pop(("t", ticket), None)

    def device_ptrs(self, ticket: int):
        ps = [C.c_void_p() for _ in range(5)]
        self._ck(self.L.sdr_ticket_device_ptrs(self.h, ticket, *[C.byref(p) for p in ps]))
        return [p.value for p in ps]

    def launch_count(self) -> int:
        return int(self.L.sdr_engine_launch_count(self.h))

    def last_kernel(self) -> str:
        return self.L.sdr_engine_last_kernel(self.h).decode()

    def fence(self):
        """orders everything submitted so far before later work on the caller's cuda_stream (sdr_engine_fence)"""
        self._ck(self.L.sdr_engine_fence(self.h))

    # ---- dsp-signature-compatible single calls ----
    def iq_to_spectrum_and_psd(self, iq: np.ndarray):
        iq = np.ascontiguousarray(iq, np.float32).reshape(-1)
        nb = iq.size // (2 * self.block_size)
        spectrum = np.empty((nb, self.block_size), np.float32)
        psd = np.empty((nb, self.block_size), np.float32)
        self._ck(self.L.sdr_dsp_iq_to_spectrum_and_psd(self.h, iq.ctypes.data_as(_f32p), nb,
                                                        spectrum.ctypes.data_as(_f32p), psd.ctypes.data_as(_f32p)))
        return spectrum, psd

    def find_noise_floor(self, psd: np.ndarray, edge_width: int):
        psd = np.ascontiguousarray(psd, np.float32)
        if psd.size != self.block_size:
            raise SdrError(EINVAL, "psd length must equal the block size")
        mn, var = C.c_float(), C.c_double()
        self._ck(self.L.sdr_dsp_find_noise_floor(self.h, psd.ctypes.data_as(_f32p), edge_width, C.byref(mn), C.byref(var)))
        return np.float32(mn.value), var.value

    def find_peaks(self, cum: np.ndarray, threshold: float, cumulation_size: int = 100, max_peaks: int = 4096):
        cum = np.ascontiguousarray(cum, np.float32)
        if cum.size != self.block_size:
            raise SdrError(EINVAL, "cumulation length must equal the block size")
        arr = (Peak * max_peaks)()
        n = C.c_int()
        self._ck(self.L.sdr_dsp_find_peaks(self.h, cum.ctypes.data_as(_f32p), cumulation_size, C.c_float(threshold), arr,
                                           max_peaks, C.byref(n)))
        return [arr[i] for i in range(min(n.value, max_peaks))], n.value


def kiwi_decode(engine: "Engine", raw: bytes) -> np.ndarray:
    out = np.empty(len(raw) // 2, np.float32)
    engine._ck(engine.L.sdr_kiwi_decode_iq_bytes(engine.h, raw, len(raw), out.ctypes.data_as(_f32p)))
    return out


class GoertzelBank:
    def __init__(self, pitches, sample_rate: int, max_blocks: int = 4096, blocksize_ratio: float = 0.005, device: int = 0):
        self.L = lib()
        self.pitches = np.ascontiguousarray(np.asarray(pitches, dtype=np.float64))
        cfg = GoertzelConfig()
        cfg.device, cfg.sample_rate, cfg.n_filters = device, sample_rate, self.pitches.size
        cfg.pitch = self.pitches.ctypes.data_as(_f64p)
        cfg.blocksize_ratio, cfg.max_blocks = blocksize_ratio, max_blocks
        h = C.c_void_p()
        rc = self.L.sdr_goertzel_create(C.byref(cfg), C.byref(h))
        if rc != OK:
            raise SdrError(rc, self.L.sdr_goertzel_last_error(None).decode())
        self.h = h
        self.n_filters = self.pitches.size
        self.max_blocks = max_blocks

    def close(self):
        if getattr(self, "h", None):
            self.L.sdr_goertzel_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != OK:
            raise SdrError(rc, self.L.sdr_goertzel_last_error(self.h).decode())

    def blocksize(self, i: int) -> int:
        return int(self.L.sdr_goertzel_blocksize(self.h, i))

    def process_audio(self, audio_list, scale=None, max_scale: float = 12.0):
        """audio_list[i]: float32 samples for filter i (whole blocks).  Returns (magnitude, state) lists."""
        nf = self.n_filters
        arrs = [np.ascontiguousarray(a, np.float32) for a in audio_list]
        nb = np.array([a.size // self.blocksize(i) for i, a in enumerate(arrs)], dtype=np.int32)
        ptrs = (_f32p * nf)(*[a.ctypes.data_as(_f32p) for a in arrs])
        stride = max(int(nb.max()), 1)
        mag = np.zeros((nf, stride), np.float64)
        st = np.zeros((nf, stride), np.uint8)
        sc = None if scale is None else np.ascontiguousarray(np.asarray(scale, dtype=np.float32))
        self._ck(self.L.sdr_goertzel_process_audio(self.h, ptrs, nb.ctypes.data_as(_i32p),
                                                   None if sc is None else sc.ctypes.data_as(_f32p), max_scale,
                                                   mag.ctypes.data_as(_f64p), st.ctypes.data_as(_u8p), stride))
        return [mag[i, : nb[i]] for i in range(nf)], [st[i, : nb[i]] for i in range(nf)]

    def process_iq(self, iq, block_size: int, bins, n_blocks: int | None = None):
        bins = np.ascontiguousarray(np.asarray(bins, dtype=np.int32))
        if isinstance(iq, np.ndarray):
            iq = np.ascontiguousarray(iq, np.float32).reshape(-1)
            nb = iq.size // (2 * block_size)
            ptr, mem = iq.ctypes.data, MEM_HOST
        else:
            nb, ptr, mem = n_blocks, int(iq), MEM_DEVICE
        out = np.empty((nb, bins.size), np.float32)
        self._ck(self.L.sdr_goertzel_process_iq(self.h, ptr, mem, block_size, nb, bins.ctypes.data_as(_i32p), bins.size,
                                                out.ctypes.data_as(_f32p)))
        return out
