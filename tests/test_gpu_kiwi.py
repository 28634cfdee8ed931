"""KiwiSDR wire format (kiwi/client.go:284-308) ingested directly by the GPU path: big-endian int16 I,Q with the
division by 32767 fused into K1's load."""
import ctypes as C

import numpy as np
import pytest

from sdrainer_b200 import synth

pytestmark = pytest.mark.gpu


def _oracle_decode(oracle, raw: bytes) -> np.ndarray:
    out = np.empty(len(raw) // 2, np.float32)
    oracle.lib().orc_kiwi_decode_iq_bytes(raw, len(raw), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def test_decode_is_bit_exact_for_every_int16(capi, oracle):
    """float32(int16)/float32(32767) for all 65536 inputs: the reciprocal+FMA division must round like IEEE"""
    vals = np.arange(-32768, 32768, dtype=np.int16)
    raw = vals.astype(">i2").tobytes()
    with capi.Engine(512) as eng:
        got = capi.kiwi_decode(eng, raw)
    ref = _oracle_decode(oracle, raw)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(ref, vals.astype(np.float32) / np.float32(32767))


def test_kiwi_stream_equals_float_stream_and_oracle(capi, oracle):
    """12 kS/s Kiwi shape (kiwi/kiwi.go:13): N=512 blocks of wire bytes vs the same samples as float32"""
    n, fs, nb = 512, 12000, 230
    rng = np.random.default_rng(12)
    tones = synth.make_tones(rng, 4, n, 70, amp_range=(0.02, 0.3))
    spec = synth.StreamSpec(sample_rate=fs, block_size=n, n_blocks=nb, seed=12, tones=tones, noise_sigma=2e-3)
    x = synth.generate(spec)
    q = np.clip(np.rint(x * 32767), -32768, 32767).astype(np.int16)
    raw = np.frombuffer(q.astype(">i2").tobytes(), np.uint8).copy()
    as_float = _oracle_decode(oracle, raw.tobytes())
    bins = [t.bin for t in tones]
    outs = []
    for fmt, data in ((capi.FMT_KIWI_I16BE, raw), (capi.FMT_F32, as_float)):
        with capi.Engine(n, max_listeners=8, max_blocks_per_batch=nb, max_peaks_per_flush=257) as eng:
            s = eng.open_stream(fs)
            outs.append(eng.collect(eng.submit([dict(stream=s, iq=data, listener_bins=bins, format=fmt)],
                                               capi.WANT_FLUSH_CUM)))
    a, b = outs
    for name in ("psd_noise_floor", "noise_variance", "thresholds", "taps", "keys", "flush_cum", "flush_n_peaks"):
        assert np.array_equal(getattr(a, name), getattr(b, name), equal_nan=True), name  # same floats in -> same bits out
    r = oracle.process_stream(as_float, n, listener_bins=bins, sample_rate=fs)
    rel = np.abs(a.psd_noise_floor.astype(np.float64) - r.noise[:, 0]) / r.noise[:, 0]
    assert rel.max() < 1e-4
    listen = r.thresholds[:, 0] + r.thresholds[:, 1]
    flips = np.argwhere(a.keys[:, :len(bins)] != (r.taps > listen[:, None]))
    for blk, l in flips:
        assert abs(float(r.taps[blk, l]) - float(listen[blk])) < 1e-3
    for f in range(r.n_flush):
        got = [(int(p["from"]), int(p["to"]), int(p["signal_bin"])) for p in a.peaks(f)]
        assert got == [p.key() for p in r.peaks[f]]


def test_mixed_formats_in_one_submit_are_rejected(capi):
    n = 512
    with capi.Engine(n, max_streams=2, max_blocks_per_batch=8) as eng:
        s0, s1 = eng.open_stream(12000), eng.open_stream(12000)
        with pytest.raises(capi.SdrError):
            eng.submit([dict(stream=s0, iq=np.zeros(2 * n, np.float32)),
                        dict(stream=s1, iq=np.zeros(4 * n, np.uint8), format=capi.FMT_KIWI_I16BE)])
