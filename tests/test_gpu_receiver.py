"""rx.Receiver end to end: the C++ host mirror over the GPU engine against the oracle's Receiver.run driver
(identical IQ, manual block clock, deterministic FindNext): same attach sequence, same key streams, same text."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from sdrainer_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host():
    from sdrainer_b200 import _build, hostapi
    _build.build_host()
    hostapi.lib()
    return hostapi


def _oracle_receiver(oracle, spec, iq, strain=True, pool=30, force_bins=(), force_at=0):
    L = oracle.lib()
    cfg = oracle.ReceiverConfig()
    L.orc_receiver_config_default(C.byref(cfg), spec.sample_rate, spec.block_size)
    cfg.strain_mode = 1 if strain else 0
    cfg.listener_pool_size = pool
    rx = L.orc_receiver_new(C.byref(cfg))
    n = spec.block_size
    fp = C.POINTER(C.c_float)
    reports = []
    for b in range(spec.n_blocks):
        if b == force_at:
            for fb in force_bins:
                assert L.orc_receiver_force_attach(rx, fb) >= 0
        blk = np.ascontiguousarray(iq[b * 2 * n:(b + 1) * 2 * n])
        assert L.orc_receiver_process_block(rx, blk.ctypes.data_as(fp)) == 0
        r = L.orc_receiver_last_report(rx).contents
        reports.append((r.psd_noise_floor, r.noise_floor, r.noise_deviation, r.peak_threshold, r.listen_threshold))
    out = []
    for i in range(L.orc_receiver_listener_count(rx)):
        nk = C.c_int64()
        kp = L.orc_receiver_listener_keys(rx, i, C.byref(nk))
        out.append(dict(bin=L.orc_receiver_listener_bin(rx, i), text=L.orc_receiver_listener_text(rx, i).decode("utf-8"),
                        keys=np.array([kp[j] for j in range(nk.value)], np.uint8),
                        attach_block=L.orc_receiver_listener_attach_block(rx, i)))
    L.orc_receiver_free(rx)
    return out, np.asarray(reports, np.float32)


def test_strain_mode_receiver_matches_oracle(capi, oracle, host):
    """config 1 (48 kS/s, N=512, 5 keyed tones at 20 WPM) through rx.Receiver in strain mode"""
    spec = synth.config(1, seconds=14.0)
    iq = synth.generate(spec)
    n = spec.block_size
    ref, ref_reports = _oracle_receiver(oracle, spec, iq)
    with capi.Engine(n, max_streams=1, max_listeners=32, max_blocks_per_batch=100, max_peaks_per_flush=n // 2 + 1) as eng:
        rx = host.Receiver(eng, strain=True, pool_size=30)
        rx.start(spec.sample_rate, n)
        # feed like tci/tci.go:264: 4 frames per TCI packet, process whenever the queue holds a packet
        for b in range(spec.n_blocks):
            assert rx.iq_data(spec.sample_rate, iq[b * 2 * n:(b + 1) * 2 * n])
            if b % 4 == 3:
                rx.process()
        rx.process()
        got = rx.listeners()
        reports, _ = rx.reports()
        events = rx.events()
        rx.close()
    assert len(got) == len(ref) >= 5  # one listener attached per cumulation window
    assert np.abs(reports[:, 1:] - ref_reports[:, 1:]).max() < 1e-4
    for g, r in zip(got, ref):
        assert g["attach_block"] == r["attach_block"]
        assert np.array_equal(g["keys"], r["keys"])
        assert g["text"] == r["text"]
    # the listeners attached to the synthetic tones, and the activation events carry the interpolated frequency
    tone_bins = {t.bin for t in spec.tones}
    assert {r["bin"] for r in ref} <= tone_bins
    plus = [e for e in events if e.startswith("+")]
    for e, r in zip(plus, ref):
        freq = int(e.split("@")[1])
        assert abs(freq - (-spec.sample_rate // 2 + int(r["bin"] * spec.sample_rate / n))) <= 47
    assert sum(1 for e in events if e.startswith("+")) == len(got)
    decoded = " ".join(g["text"] for g in got)
    assert "dl1abc" in decoded


def test_decode_mode_golden_text_through_the_gpu(capi, host):
    """a recorded golden key stream re-synthesised as IQ (bounded-PSD background, on-bin
    tone) must come back as the golden text through FFT -> thresholds -> decoder on the GPU path."""
    with open(os.path.join(GOLDEN, "cw_keystreams.json"), encoding="utf-8") as f:
        s = json.load(f)["streams"][4]  # ii3wwa
    bits, cur = [], s["first"]
    for r in s["runs"]:
        bits.extend([cur] * r)
        cur ^= 1
    n, fs, kbin, warm = 512, 48000, 300, 70
    rng = np.random.default_rng(11)
    total = warm + len(bits) + 40
    nidx = np.arange(n)
    tone = 0.01 * np.exp(2j * np.pi * (kbin - n // 2) * nidx / n)
    with capi.Engine(n, max_streams=1, max_listeners=4, max_blocks_per_batch=100) as eng:
        rx = host.Receiver(eng, strain=False)
        rx.start(fs, n)
        for b in range(total):
            P = rng.uniform(0.5, 1.5, n) * (n * 2e-8)
            x = np.fft.ifft(np.sqrt(P) * np.exp(2j * np.pi * rng.uniform(0, 1, n)))
            if warm <= b < warm + len(bits) and bits[b - warm]:
                x = x + tone
            if b == warm:
                rx.process()
                assert rx.attach_at_bin(kbin) == 0
            frame = np.empty(2 * n, np.float32)
            frame[0::2], frame[1::2] = x.real, x.imag
            assert rx.iq_data(fs, frame)
            if b % 7 == 6:
                rx.process()
        rx.process()
        lst = rx.listeners()
        rx.close()
    assert list(lst[0]["keys"][:len(bits)]) == bits
    assert lst[0]["text"] == s["expected"]


def test_receiver_drops_when_queue_is_full_and_rejects_bad_frames(capi, host):
    """rx/receiver.go:319-333"""
    n = 512
    with capi.Engine(n, max_blocks_per_batch=100) as eng:
        rx = host.Receiver(eng, strain=True)
        rx.start(48000, n)
        frame = np.zeros(2 * n, np.float32)
        assert not rx.iq_data(44100, frame)          # wrong sample rate
        assert not rx.iq_data(48000, frame[:100])    # wrong block size
        oks = [rx.iq_data(48000, frame + 1e-4) for _ in range(105)]
        assert sum(oks) == 100 and not any(oks[100:])  # iqBufferSize = 100, then "IQ data skipped"
        assert rx.process() == 100
        rx.close()


def test_audio_demodulator_decodes_keyed_tone(capi, host):
    """cw.AudioDemodulator (cw/audio.go) over the GPU Goertzel bank: keyed 700 Hz audio -> text"""
    fs, pitch = 48000, 700.0
    L = host.lib()
    h = L.sdrh_audio_new(pitch, fs)
    assert h
    try:
        assert L.sdrh_audio_blocksize(h) == 207
        L.sdrh_audio_set_scale(h, 0.0)  # autoscale, as `sdrainer decode pulse` runs it (cw/audio.go:184-188)
        n = fs * 9
        t = np.arange(n) / fs
        env = synth.keying("cq de dl1abc k", 20.0, fs, n, 0.3)
        audio = (0.6 * np.cos(2 * np.pi * pitch * t) * env).astype(np.float32)
        for i in range(0, n, 4800):  # PulseAudio-sized writes
            chunk = np.ascontiguousarray(audio[i:i + 4800])
            assert L.sdrh_audio_write(h, chunk.ctypes.data_as(C.POINTER(C.c_float)), chunk.size) == chunk.size
        L.sdrh_audio_close(h)
        assert "cq de dl1abc k" in L.sdrh_audio_text(h).decode("utf-8")
    finally:
        L.sdrh_audio_free(h)
