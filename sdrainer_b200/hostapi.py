"""ctypes binding of libsdrhost.so: the C++ host mirror of the reference's Go interface (dsp / cw / rx)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build, capi

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    capi.lib()  # libsdrgpu.so first (libsdrhost.so links against it)
    if not os.path.exists(_build.HOST_LIB):
        raise RuntimeError(f"{_build.HOST_LIB} is missing: run __graft_entry__.build()")
    L = C.CDLL(_build.HOST_LIB)
    vp, i, d, f, ll, cp = C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_longlong, C.c_char_p
    sig = {
        "sdrh_decoder_new": (vp, [i, i]), "sdrh_decoder_free": (None, [vp]), "sdrh_decoder_reset": (None, [vp]),
        "sdrh_decoder_tick": (None, [vp, i]), "sdrh_decoder_ticks": (None, [vp, C.c_char_p, i]),
        "sdrh_decoder_stop": (None, [vp]), "sdrh_decoder_text": (cp, [vp]), "sdrh_decoder_clear_text": (None, [vp]),
        "sdrh_debouncer_new": (vp, [i]), "sdrh_debouncer_free": (None, [vp]), "sdrh_debouncer_debounce": (i, [vp, i]),
        "sdrh_bin_to_frequency": (ll, [i, i, ll, i, d]), "sdrh_frequency_to_bin": (i, [i, i, ll, ll]),
        "sdrh_peak_signal_frequency": (ll, [i, i, ll, i, f, f, f]),
        "sdrh_peaks_new": (vp, [i]), "sdrh_peaks_free": (None, [vp]), "sdrh_peaks_make": (i, [vp, i, i]),
        "sdrh_peaks_put": (None, [vp, i, i]), "sdrh_peaks_activate": (None, [vp, i]), "sdrh_peaks_deactivate": (None, [vp, i]),
        "sdrh_peaks_cleanup": (None, [vp]), "sdrh_peaks_clock_add": (None, [vp, d]), "sdrh_peaks_find_next": (i, [vp]),
        "sdrh_peaks_bin": (i, [vp, i]), "sdrh_peaks_bin_state": (i, [vp, i]),
        "sdrh_pool_new": (vp, [i, cp]), "sdrh_pool_free": (None, [vp]), "sdrh_pool_bind_next": (i, [vp]),
        "sdrh_pool_made_id": (cp, [vp, i]), "sdrh_pool_release": (None, [vp, i]), "sdrh_pool_len": (i, [vp]),
        "sdrh_pool_active_id": (cp, [vp, i]),
        "sdrh_receiver_new": (vp, [vp, i, i]), "sdrh_receiver_free": (None, [vp]), "sdrh_receiver_start": (i, [vp, i, i]),
        "sdrh_receiver_stop": (None, [vp]), "sdrh_receiver_set": (None, [vp, f, i, d, d, i, ll]),
        "sdrh_receiver_set_find_next": (None, [vp, i, C.c_ulonglong]),
        "sdrh_receiver_set_device_debounce": (None, [vp, i]),
        "sdrh_receiver_iq_data": (i, [vp, i, C.POINTER(C.c_float), ll]), "sdrh_receiver_process": (i, [vp]),
        "sdrh_receiver_error": (cp, [vp]), "sdrh_receiver_attach_at_bin": (i, [vp, i]),
        "sdrh_receiver_skipped": (i, [vp]), "sdrh_receiver_rejected": (i, [vp]),
        "sdrh_receiver_listener_count": (i, [vp]), "sdrh_receiver_listener_bin": (i, [vp, i]),
        "sdrh_receiver_listener_text": (cp, [vp, i]),
        "sdrh_receiver_listener_keys": (C.POINTER(C.c_ubyte), [vp, i, C.POINTER(ll)]),
        "sdrh_receiver_attach_block": (ll, [vp, i]), "sdrh_receiver_n_reports": (i, [vp]),
        "sdrh_receiver_report": (None, [vp, i, C.POINTER(C.c_float), C.POINTER(d)]),
        "sdrh_receiver_n_events": (i, [vp]), "sdrh_receiver_event": (cp, [vp, i]),
        "sdrh_receiver_n_flushes": (i, [vp]),
        "sdrh_receiver_flush_peaks": (i, [vp, i, C.POINTER(i), C.POINTER(ll), i]),
        "sdrh_dispatcher_new": (vp, [vp, i, i]), "sdrh_dispatcher_free": (None, [vp]), "sdrh_dispatcher_add": (i, [vp, vp]),
        "sdrh_dispatcher_tick": (i, [vp]), "sdrh_dispatcher_submits": (i, [vp]), "sdrh_dispatcher_error": (cp, [vp]),
        "sdrh_rt_new": (vp, [vp, i, i, i, i, i, i, i, C.POINTER(C.c_float), i, i, C.POINTER(C.c_int)]),
        "sdrh_rt_free": (None, [vp]), "sdrh_rt_run": (i, [vp, i, i, i, C.POINTER(d)]),
        "sdrh_audio_new": (vp, [d, i]), "sdrh_audio_free": (None, [vp]), "sdrh_audio_blocksize": (i, [vp]),
        "sdrh_audio_set_scale": (None, [vp, d]), "sdrh_audio_write": (i, [vp, C.POINTER(C.c_float), i]),
        "sdrh_audio_close": (None, [vp]), "sdrh_audio_text": (cp, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class Decoder:
    """cw.Decoder (cw/decode.go) -- C++ host mirror."""

    def __init__(self, sample_rate=48000, block_size=512):
        self.L = lib()
        self.h = self.L.sdrh_decoder_new(sample_rate, block_size)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.sdrh_decoder_free(self.h)
            self.h = None

    def reset(self):
        self.L.sdrh_decoder_reset(self.h)
        self.L.sdrh_decoder_clear_text(self.h)

    def feed(self, keys):
        buf = bytes(bytearray(int(k) for k in keys))
        self.L.sdrh_decoder_ticks(self.h, buf, len(buf))

    def stop(self):
        self.L.sdrh_decoder_stop(self.h)

    @property
    def text(self) -> str:
        return self.L.sdrh_decoder_text(self.h).decode("utf-8")


class Receiver:
    """rx.Receiver (rx/receiver.go) over a capi.Engine."""

    def __init__(self, engine: capi.Engine, strain=True, pool_size=30):
        self.L = lib()
        self.engine = engine
        self.h = self.L.sdrh_receiver_new(engine.h, 1 if strain else 0, pool_size)

    def close(self):
        if getattr(self, "h", None):
            self.L.sdrh_receiver_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def configure(self, peak_threshold=15.0, edge_width=70, silence_s=20.0, attach_s=120.0, debounce=1, center=0):
        self.L.sdrh_receiver_set(self.h, peak_threshold, edge_width, silence_s, attach_s, debounce, center)

    def set_device_debounce(self, on=True):
        """dsp.BoolDebouncer on the device, packed key bits back (call before start)"""
        self.L.sdrh_receiver_set_device_debounce(self.h, 1 if on else 0)

    def set_find_next(self, deterministic=True, seed=1):
        self.L.sdrh_receiver_set_find_next(self.h, 1 if deterministic else 0, seed)

    def start(self, sample_rate, block_size):
        if self.L.sdrh_receiver_start(self.h, sample_rate, block_size) != 0:
            raise RuntimeError("Receiver.Start failed")
        self.block_size = block_size

    def stop(self):
        self.L.sdrh_receiver_stop(self.h)

    def iq_data(self, sample_rate, frame: np.ndarray) -> bool:
        frame = np.ascontiguousarray(frame, np.float32)
        return bool(self.L.sdrh_receiver_iq_data(self.h, sample_rate, frame.ctypes.data_as(C.POINTER(C.c_float)), frame.size))

    def process(self) -> int:
        n = self.L.sdrh_receiver_process(self.h)
        if n < 0:
            raise RuntimeError(self.L.sdrh_receiver_error(self.h).decode())
        return n

    def attach_at_bin(self, b):
        return self.L.sdrh_receiver_attach_at_bin(self.h, b)

    def listeners(self):
        out = []
        for i in range(self.L.sdrh_receiver_listener_count(self.h)):
            n = C.c_longlong()
            kp = self.L.sdrh_receiver_listener_keys(self.h, i, C.byref(n))
            keys = np.ctypeslib.as_array(kp, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint8)
            out.append(dict(bin=self.L.sdrh_receiver_listener_bin(self.h, i),
                            text=self.L.sdrh_receiver_listener_text(self.h, i).decode("utf-8"), keys=keys,
                            attach_block=self.L.sdrh_receiver_attach_block(self.h, i)))
        return out

    def reports(self):
        n = self.L.sdrh_receiver_n_reports(self.h)
        out = np.zeros((n, 5), np.float32)
        var = np.zeros(n, np.float64)
        for b in range(n):
            v = C.c_double()
            self.L.sdrh_receiver_report(self.h, b, out[b].ctypes.data_as(C.POINTER(C.c_float)), C.byref(v))
            var[b] = v.value
        return out, var

    def events(self):
        return [self.L.sdrh_receiver_event(self.h, i).decode() for i in range(self.L.sdrh_receiver_n_events(self.h))]

    def flush_peaks(self, f):
        cap = self.block_size
        bins = (C.c_int * cap)()
        fr = (C.c_longlong * cap)()
        n = self.L.sdrh_receiver_flush_peaks(self.h, f, bins, fr, cap)
        return [(bins[i], fr[i]) for i in range(min(n, cap))]

    def n_flushes(self):
        return self.L.sdrh_receiver_n_flushes(self.h)


class Dispatcher:
    """rx.Dispatcher (host/sdrhost.hpp): drains the queues of many Receivers into ONE sdr_submit per tick."""

    def __init__(self, engine: capi.Engine, block_size: int, max_receivers: int):
        self.L = lib()
        self.h = self.L.sdrh_dispatcher_new(engine.h, block_size, max_receivers)
        if not self.h:
            raise RuntimeError("Dispatcher: pinned arena allocation failed")
        self._rx = []

    def add(self, rx: Receiver):
        if self.L.sdrh_dispatcher_add(self.h, rx.h) != 0:
            raise RuntimeError("dispatcher is full")
        self._rx.append(rx)

    def tick(self) -> int:
        n = self.L.sdrh_dispatcher_tick(self.h)
        if n < 0:
            raise RuntimeError(self.L.sdrh_dispatcher_error(self.h).decode())
        return n

    @property
    def submits(self) -> int:
        return self.L.sdrh_dispatcher_submits(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.sdrh_dispatcher_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RealtimeHarness:
    """host/realtime.hpp: S streams x L listeners through ring copy -> sdr_submit -> sdr_collect -> Decoder.Tick"""

    FIELDS = ("batch_s", "copy_s", "submit_s", "collect_wait_s", "decode_s", "gpu_ms", "ticks", "chars", "key_downs")

    def __init__(self, engine: capi.Engine, sample_rate, block_size, listeners, s_cap, blocks_per_batch, threads, src: np.ndarray,
                 bins: np.ndarray, debounce=1):
        self.L = lib()
        self.src = np.ascontiguousarray(src, np.float32)       # [templates, src_blocks, 2N]
        self.bins = np.ascontiguousarray(bins, np.int32)       # [templates, listeners]
        nt, sb = self.src.shape[0], self.src.shape[1]
        self.h = self.L.sdrh_rt_new(engine.h, sample_rate, block_size, listeners, s_cap, blocks_per_batch, threads, debounce,
                                    self.src.ctypes.data_as(C.POINTER(C.c_float)), nt, sb, self.bins.ctypes.data_as(C.POINTER(C.c_int)))
        if not self.h:
            raise RuntimeError("realtime harness: allocation failed")

    def run(self, n_streams, n_batches, ring_copy=True):
        out = (C.c_double * 9)()
        if self.L.sdrh_rt_run(self.h, n_streams, n_batches, 1 if ring_copy else 0, out) != 0:
            # free the harness while its engine is still alive: a later __del__ would close streams through a dangling handle
            self.close()
            raise RuntimeError("realtime harness: run failed (the engine's message is on stderr)")
        return dict(zip(self.FIELDS, [float(x) for x in out]))

    def close(self):
        if getattr(self, "h", None):
            self.L.sdrh_rt_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
