"""Round-2 GPU parity tests: the reference-held cw/testdata fixtures through every spectral kernel (SURVEY appendix B
re-blocked at N = 512 ... 65536), the exact dsp.FindNoiseFloor replay for narrow windows, stream-state isolation,
Receiver stop/start, and the multi-receiver dispatcher against the oracle's Receiver.run driver."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import parity_util as pu
from conftest import GOLDEN
from sdrainer_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host():
    from sdrainer_b200 import _build, hostapi
    _build.build_host()
    hostapi.lib()
    return hostapi


def _golden_streams():
    with open(os.path.join(GOLDEN, "cw_keystreams.json"), encoding="utf-8") as f:
        out = []
        for s in json.load(f)["streams"]:
            bits, cur = [], s["first"]
            for r in s["runs"]:
                bits.extend([cur] * r)
                cur ^= 1
            out.append((s["name"], np.asarray(bits, np.uint8), s["expected"]))
        return out


def _resynth(bits, n, kbin, warm, tail, seed):
    """one block per tick, bounded-PSD background built in the frequency domain, a phase-continuous
    on-bin tone on key-down ticks.  The tone amplitude scales with 1/sqrt(N) so that tone / background stays at the
    64 dB of the N = 512 fixture at every block size."""
    rng = np.random.default_rng(seed)
    total = warm + len(bits) + tail
    a = 0.01 * np.sqrt(512.0 / n)
    tone = a * np.exp(2j * np.pi * (kbin - n // 2) * np.arange(n) / n)
    iq = np.empty((total, n, 2), np.float32)
    for b in range(total):
        P = rng.uniform(0.5, 1.5, n) * (n * 2e-8)
        x = np.fft.ifft(np.sqrt(P) * np.exp(2j * np.pi * rng.uniform(0, 1, n)))
        if warm <= b < warm + len(bits) and bits[b - warm]:
            x = x + tone
        iq[b, :, 0], iq[b, :, 1] = x.real, x.imag
    return iq.reshape(-1)


# which kernel serves which block size: tests/ asserts on results only; the forcing switches make single-stream
# submits take the many-stream kernels (k1_mid8k / k1_wide) that the engine would otherwise reserve for full launches
KERNEL_ENV = {512: {}, 1024: {}, 2048: {}, 4096: {}, 8192: {"SDR_K1_MID8K": "force"}}


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096, 8192])
def test_all_nine_goldens_keys_and_text_through_the_gpu(capi, host, monkeypatch, n):
    """cw/decode_test.go:177-213: the nine recorded key streams, re-synthesised as IQ at block size n (tick = n/fs kept
    at 10.67 ms), must come back bit-exact as key states from FFT -> noise floor -> thresholds -> value > threshold on
    the GPU, and as the golden text from ONE decoder instance Reset() between files (as the reference test does)."""
    for k, v in KERNEL_ENV[n].items():
        monkeypatch.setenv(k, v)
    fs = 48000 * n // 512
    warm, tail = 70, 40
    kbin = n // 2 + n // 5 + 3
    dec = host.Decoder(fs, n)
    streams = _golden_streams()
    longest = max(len(b) for _, b, _ in streams) + warm + tail
    with capi.Engine(n, max_streams=1, max_listeners=4, max_blocks_per_batch=longest, max_peaks_per_flush=64) as eng:
        sid = eng.open_stream(fs)
        for i, (name, bits, expected) in enumerate(streams):
            iq = _resynth(bits, n, kbin, warm, tail, seed=100 + i)
            eng.reset_stream(sid)
            res = eng.collect(eng.submit([dict(stream=sid, iq=iq, listener_bins=[kbin])], capi.NO_PEAKS))
            keys = res.keys[warm:warm + len(bits), 0]
            assert np.array_equal(keys, bits), f"{name}: {int((keys != bits).sum())} key states differ at N={n}"
            assert not res.keys[warm + len(bits):, 0].any()  # (the first ~59 warm-up blocks key down: zero-initialised rolling means)
            dec.reset()  # Decoder.Reset keeps lastState: files run in the listed order on one instance
            dec.feed(keys)
            dec.stop()
            assert dec.text == expected, name


def test_golden_through_the_wideband_kernel(capi, host, monkeypatch):
    """the same fixture at N = 65536 (BASELINE config 5 shape) through the single-pass wideband kernel"""
    monkeypatch.setenv("SDR_K1_WIDE", "force")
    n = 65536
    fs = 48000 * n // 512
    name, bits, expected = _golden_streams()[5]  # ly2px_1: 213 ticks
    warm, tail, kbin = 70, 5, n // 2 + 9001
    iq = _resynth(bits, n, kbin, warm, tail, seed=65)
    with capi.Engine(n, max_streams=1, max_listeners=4, max_blocks_per_batch=warm + len(bits) + tail) as eng:
        sid = eng.open_stream(fs)
        res = eng.collect(eng.submit([dict(stream=sid, iq=iq, listener_bins=[kbin])], capi.NO_PEAKS))
    keys = res.keys[warm:warm + len(bits), 0]
    assert np.array_equal(keys, bits), f"key states differ at ticks {np.flatnonzero(keys != bits)[:20]}"
    dec = host.Decoder(fs, n)
    dec.feed(keys)
    dec.stop()
    assert dec.text == expected


@pytest.mark.parametrize("n,edge", [(512, 212), (512, 220), (512, 251), (512, 256), (512, 300), (2048, 1000), (8192, 4070)])
def test_narrow_noise_windows_take_the_exact_replay(capi, oracle, n, edge):
    """dsp.FindNoiseFloor (dsp/fft.go:215-252) with (N-2e)/10 < 9: the reference closes a window every `windowSize`
    bins (up to 19 of them), or none at all when windowSize <= 0; the engine replays the loop literally"""
    spec = synth.StreamSpec(sample_rate=48000, block_size=n, n_blocks=12, seed=n + edge,
                            tones=synth.make_tones(np.random.default_rng(edge), 3, n, 70))
    iq = synth.generate(spec)
    with capi.Engine(n, max_streams=1, max_listeners=4, max_blocks_per_batch=16) as eng:
        sid = eng.open_stream(48000)
        res = eng.collect(eng.submit([dict(stream=sid, iq=iq, edge_width=edge, listener_bins=[n // 2])], capi.WANT_SPECTRUM))
        psd = res.psd
        for b in range(3):  # the single call takes the same route
            mn, var = eng.find_noise_floor(psd[b], edge)
            L = oracle.lib()
            omn, ovar = C.c_float(), C.c_double()
            L.orc_find_noise_floor(np.ascontiguousarray(psd[b]).ctypes.data_as(C.POINTER(C.c_float)), n, edge, C.byref(omn), C.byref(ovar))
            # identical PSD in, identical arithmetic: bit-exact (NaN/Inf compare as equal bit patterns)
            assert np.float32(mn).tobytes() == np.float32(omn.value).tobytes()
            assert np.float64(var).tobytes() == np.float64(ovar.value).tobytes()
            assert np.float32(res.psd_noise_floor[b]).tobytes() == np.float32(omn.value).tobytes()
            assert np.float64(res.noise_variance[b]).tobytes() == np.float64(ovar.value).tobytes()
    with capi.Engine(n, max_streams=1) as eng:
        sid = eng.open_stream(48000)
        with pytest.raises(capi.SdrError):
            eng.submit([dict(stream=sid, iq=iq, edge_width=-1)])


def test_reset_of_one_stream_leaves_the_others_partial_window_alone(capi):
    """opening / resetting stream B must not touch stream A's saved partial cumulation"""
    n = 1024
    spec = synth.StreamSpec(sample_rate=96000, block_size=n, n_blocks=100, seed=5,
                            tones=synth.make_tones(np.random.default_rng(5), 4, n, 70))
    iq = synth.generate(spec)
    bins = [t.bin for t in spec.tones]
    half = 50 * 2 * n
    with capi.Engine(n, max_streams=4, max_listeners=8, max_blocks_per_batch=200, max_peaks_per_flush=n // 2 + 1) as eng:
        a = eng.open_stream(96000)
        whole = eng.collect(eng.submit([dict(stream=a, iq=iq, listener_bins=bins)], capi.WANT_FLUSH_CUM))
        eng.reset_stream(a)
        eng.collect(eng.submit([dict(stream=a, iq=iq[:half], listener_bins=bins)]))
        b = eng.open_stream(96000)           # r1 bug: wiped row `b` = A's second state row
        c = eng.open_stream(96000)
        eng.collect(eng.submit([dict(stream=b, iq=iq[:half // 2], listener_bins=bins)]))
        eng.reset_stream(b)
        eng.reset_stream(c)
        rest = eng.collect(eng.submit([dict(stream=a, iq=iq[half:], listener_bins=bins)], capi.WANT_FLUSH_CUM))
        assert rest.n_flushes == 1
        assert np.array_equal(rest.flush_cum[0], whole.flush_cum[0])
        assert pu.peak_keys(rest.peaks(0)) == pu.peak_keys(whole.peaks(0))


def test_receiver_stop_start_in_the_middle_of_a_window(capi, host):
    """Stop/Start after 30 blocks -- run() restarts with cumulationCount = 0 on both sides of the boundary"""
    n, fs = 512, 48000
    spec = synth.config(1, seconds=3.0)
    iq = synth.generate(spec)
    with capi.Engine(n, max_streams=2, max_listeners=32, max_blocks_per_batch=100, max_peaks_per_flush=n // 2 + 1) as eng:
        rx = host.Receiver(eng, strain=True)
        rx.start(fs, n)
        for b in range(30):
            assert rx.iq_data(fs, iq[b * 2 * n:(b + 1) * 2 * n])
        assert rx.process() == 30
        rx.stop()
        rx.start(fs, n)
        done = 0
        for b in range(30, 250):
            assert rx.iq_data(fs, iq[b * 2 * n:(b + 1) * 2 * n])
            if b % 10 == 9:
                done += rx.process()
        done += rx.process()
        assert done == 220
        assert rx.n_flushes() == 2  # 100 and 200 blocks after the restart
        rx.close()


def test_dispatcher_64_receivers_one_engine_match_64_oracle_receivers(capi, oracle, host):
    """SURVEY 8(f1) / rx/receiver.go:315-364: 64 rx.Receiver mirrors share one engine; the dispatcher drains their
    queues into ONE sdr_submit per tick.  Every receiver must reproduce its own oracle Receiver.run exactly (attach
    blocks, key streams, text)."""
    from test_gpu_receiver import _oracle_receiver
    n, fs, n_rx = 512, 48000, 64
    specs = [synth.config(1, seconds=5.0, stream=900 + i) for i in range(4)]
    iqs = [synth.generate(sp) for sp in specs]
    refs = [_oracle_receiver(oracle, sp, iq)[0] for sp, iq in zip(specs, iqs)]
    nb = specs[0].n_blocks
    with capi.Engine(n, max_streams=n_rx, max_listeners=32, max_blocks_per_batch=n_rx * 100, max_peaks_per_flush=n // 2 + 1) as eng:
        disp = host.Dispatcher(eng, n, n_rx)
        rxs = []
        for i in range(n_rx):
            rx = host.Receiver(eng, strain=True, pool_size=30)
            rx.start(fs, n)
            disp.add(rx)
            rxs.append(rx)
        total = 0
        for b in range(nb):
            for i, rx in enumerate(rxs):
                assert rx.iq_data(fs, iqs[i % 4][b * 2 * n:(b + 1) * 2 * n])
            if b % 8 == 7:  # ~85 ms of signal per tick
                total += disp.tick()
        total += disp.tick()
        assert total == n_rx * nb
        # one submit per tick (plus one when a tick crosses a flush boundary), not one per receiver
        assert disp.submits <= (nb // 8 + 1) + nb // 100 + 2
        for i, rx in enumerate(rxs):
            got, ref = rx.listeners(), refs[i % 4]
            assert len(got) == len(ref) >= 2
            for g, r in zip(got, ref):
                assert g["attach_block"] == r["attach_block"]
                assert np.array_equal(g["keys"], r["keys"])
                assert g["text"] == r["text"]
        disp.close()
        for rx in rxs:
            rx.close()


def test_concurrent_submitters_on_one_engine(capi):
    """safe for concurrent calls on different streams (one goroutine per receiver in the reference)"""
    import threading
    n, n_thr = 1024, 8
    spec = synth.StreamSpec(sample_rate=96000, block_size=n, n_blocks=40, seed=77,
                            tones=synth.make_tones(np.random.default_rng(77), 4, n, 70))
    iq = synth.generate(spec)
    bins = [t.bin for t in spec.tones]
    with capi.Engine(n, max_streams=n_thr, max_listeners=8, max_blocks_per_batch=64, n_slots=n_thr) as eng:
        sids = [eng.open_stream(96000) for _ in range(n_thr)]
        ref = eng.collect(eng.submit([dict(stream=sids[0], iq=iq, listener_bins=bins)]))
        eng.reset_stream(sids[0])
        outs, errs = [None] * n_thr, []

        def work(i):
            try:
                for rep in range(5):
                    while True:
                        try:
                            t = eng.submit([dict(stream=sids[i], iq=iq, listener_bins=bins)])
                            break
                        except capi.SdrError as ex:
                            if ex.code != capi.EBUSY:
                                raise
                    r = eng.collect(t)
                    if rep == 0:
                        outs[i] = r
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)

        ths = [threading.Thread(target=work, args=(i,)) for i in range(n_thr)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        assert not errs, errs
        for o in outs:
            assert np.array_equal(o.keys, ref.keys) and np.array_equal(o.thresholds, ref.thresholds)


def _py_debounce(raw, thr, state=None):
    """dsp.BoolDebouncer.Debounce (dsp/dsp.go:164-182) in plain Python"""
    eff, last, count = state or (False, False, 0)
    out = []
    for r in raw:
        r = bool(r)
        if thr < 2:
            out.append(r)
            continue
        count = 1 if r != last else count + 1
        last = r
        if count >= thr and r != eff:
            eff = r
        out.append(eff)
    return np.asarray(out, np.uint8), (eff, last, count)


@pytest.mark.parametrize("n,thr", [(512, 1), (512, 3), (2048, 2), (8192, 4)])
def test_packed_key_bits_and_device_debouncer(capi, n, thr):
    """one bit per listener and block, debounced on the device with the state carried per position across
    ragged submits (dsp/dsp.go:139-182, cw/spectral.go:48-54); inactive positions stay 0 and keep their state; a reset
    position starts from zero"""
    fs = 48000 * n // 512
    spec = synth.StreamSpec(sample_rate=fs, block_size=n, n_blocks=260, seed=n + thr,
                            tones=synth.make_tones(np.random.default_rng(n + thr), 37, n, 70, wpm_range=(25.0, 40.0)))
    iq = synth.generate(spec)
    bins = [t.bin for t in spec.tones]
    L = len(bins)
    with capi.Engine(n, max_streams=2, max_listeners=40, max_blocks_per_batch=260, max_peaks_per_flush=64) as eng:
        sid = eng.open_stream(fs)
        whole = eng.collect(eng.submit([dict(stream=sid, iq=iq, listener_bins=bins, signal_debounce=thr)]))
        raw = whole.keys[:, :L]
        got = whole.debounced_keys(L)
        for l in range(L):
            want, _ = _py_debounce(raw[:, l], thr)
            assert np.array_equal(got[:, l], want), (l, thr)
        assert not whole.debounced_keys(whole.key_bits.shape[1] * 32)[:, L:].any()  # unused positions
        # ragged submits: identical bits, no raw keys copied back
        eng.reset_stream(sid)
        parts, pos = [], 0
        for c in (1, 59, 3, 100, 97):
            r = eng.collect(eng.submit([dict(stream=sid, iq=iq[pos * 2 * n:(pos + c) * 2 * n], listener_bins=bins, signal_debounce=thr)],
                                       capi.NO_RAW_KEYS | capi.NO_TAPS))
            assert r.keys is None
            parts.append(r.debounced_keys(L))
            pos += c
        assert np.array_equal(np.concatenate(parts), got)
        # flags: position 1 inactive in the middle part (bit 0, state kept), position 2 reset at the start of the last part
        eng.reset_stream(sid)
        cuts = (0, 80, 170, 260)
        outs = []
        for k in range(3):
            fl = np.full(L, capi.LISTENER_ACTIVE, np.uint8)
            if k == 1:
                fl[1] = 0
            if k == 2:
                fl[2] |= capi.LISTENER_RESET
            r = eng.collect(eng.submit([dict(stream=sid, iq=iq[cuts[k] * 2 * n:cuts[k + 1] * 2 * n], listener_bins=bins,
                                             signal_debounce=thr, listener_flags=fl)]))
            outs.append(r.debounced_keys(L))
        flagged = np.concatenate(outs)
        others = [l for l in range(L) if l not in (1, 2)]
        assert np.array_equal(flagged[:, others], got[:, others])
        a, st = _py_debounce(raw[:80, 1], thr)
        c, _ = _py_debounce(raw[170:, 1], thr, st)  # the debouncer did not see blocks 80..169
        assert np.array_equal(flagged[:, 1], np.concatenate([a, np.zeros(90, np.uint8), c]))
        a, _ = _py_debounce(raw[:170, 2], thr)
        c, _ = _py_debounce(raw[170:, 2], thr)      # fresh debouncer from block 170
        assert np.array_equal(flagged[:, 2], np.concatenate([a, c]))


@pytest.mark.parametrize("debounce", [1, 2, 3])
def test_receiver_with_device_debounce_matches_oracle(capi, oracle, host, debounce):
    """rx.Receiver in strain mode with SetSignalDebounce(d): the debouncer runs on the device, the host mirror only ticks
    the decoder -- same attach blocks, debounced key streams and text as the oracle's Receiver.run"""
    from test_gpu_receiver import _oracle_receiver  # noqa: F401
    L = oracle.lib()
    spec = synth.config(1, seconds=9.0)
    iq = synth.generate(spec)
    n = spec.block_size
    cfg = oracle.ReceiverConfig()
    L.orc_receiver_config_default(C.byref(cfg), spec.sample_rate, n)
    cfg.strain_mode, cfg.listener_pool_size, cfg.signal_debounce = 1, 30, debounce
    orx = L.orc_receiver_new(C.byref(cfg))
    fp = C.POINTER(C.c_float)
    for b in range(spec.n_blocks):
        blk = np.ascontiguousarray(iq[b * 2 * n:(b + 1) * 2 * n])
        assert L.orc_receiver_process_block(orx, blk.ctypes.data_as(fp)) == 0
    ref = []
    for i in range(L.orc_receiver_listener_count(orx)):
        nk = C.c_int64()
        kp = L.orc_receiver_listener_keys(orx, i, C.byref(nk))
        ref.append(dict(text=L.orc_receiver_listener_text(orx, i).decode("utf-8"), keys=np.array([kp[j] for j in range(nk.value)], np.uint8),
                        attach_block=L.orc_receiver_listener_attach_block(orx, i)))
    L.orc_receiver_free(orx)
    with capi.Engine(n, max_streams=1, max_listeners=32, max_blocks_per_batch=100, max_peaks_per_flush=n // 2 + 1) as eng:
        rx = host.Receiver(eng, strain=True, pool_size=30)
        rx.set_device_debounce(True)
        rx.configure(debounce=debounce)
        rx.start(spec.sample_rate, n)
        for b in range(spec.n_blocks):
            assert rx.iq_data(spec.sample_rate, iq[b * 2 * n:(b + 1) * 2 * n])
            if b % 5 == 4:
                rx.process()
        rx.process()
        got = rx.listeners()
        rx.close()
    assert len(got) == len(ref) >= 3
    for g, r in zip(got, ref):
        assert g["attach_block"] == r["attach_block"]
        assert np.array_equal(g["keys"], r["keys"])
        assert g["text"] == r["text"]


@pytest.mark.parametrize("n", [16384, 32768])
def test_stockham_block_sizes_against_the_oracle(capi, oracle, n):
    """N = 16384 / 32768 (64 x 256 and 128 x 256: fast_cols64_kernel + fast_rows256_kernel; round 1 ran them through the
    shared-memory Stockham four-step kernels): noise floor, thresholds, key states, cumulation and peak list against the
    oracle; tolerances from profiles/r2_error_table.md"""
    import test_gpu_parity as tp
    fs = 48000 * n // 512
    rng = np.random.default_rng(n)
    tones = synth.make_tones(rng, 24, n, 70, wpm_range=(18.0, 28.0))
    spec = synth.StreamSpec(sample_rate=fs, block_size=n, n_blocks=130, seed=n + 1, tones=tones)
    iq = synth.generate(spec)
    bins = [t.bin for t in tones]
    outs = tp._run_batch(capi, spec, iq, bins, chunks=[70, 60])
    r = oracle.process_stream(iq, n, edge_width=70, peak_threshold=15.0, listener_bins=bins, sample_rate=fs)
    pu.check_scalars(tp._concat(outs, "psd_noise_floor"), r.noise[:, 0], what="psdNoiseFloor")
    pu.check_scalars(tp._concat(outs, "noise_variance"), r.noise[:, 1], what="noise variance")
    thr = tp._concat(outs, "thresholds")
    assert np.abs(thr[:, :3] - r.thresholds).max() < 1e-4
    taps = tp._concat(outs, "taps")[:, :len(bins)]
    strong = r.taps > float(np.median(r.thresholds[60:, 0])) + 15
    assert strong.any() and np.abs(taps[strong] - r.taps[strong]).max() < 2e-3  # signal bins: <= 1.4e-3 dB at N = 32768
    pu.check_keys(tp._concat(outs, "keys")[:, :len(bins)], r.taps, r.thresholds[:, 0] + r.thresholds[:, 1])
    cum = outs[1].flush_cum[0]
    loud = r.flush_cum[0] > np.median(r.flush_cum[0]) + 1000.0
    assert loud.any() and np.abs(cum - r.flush_cum[0])[loud].max() < 0.1
    assert np.abs(cum - r.flush_cum[0]).max() < 2.0  # noise-level bins next to 24 carriers, summed over 100 blocks
    pu.check_peaks(pu.peak_keys(outs[1].peaks(0)), [p.key() for p in r.peaks[0]], r.flush_cum[0], r.thresholds[99, 2])


def test_realtime_harness_runs_the_whole_loop(capi, host):
    """host/realtime.hpp at toy size: ring copy -> one sdr_submit for all streams -> collect -> Decoder.Tick per key bit.
    Keyed tones on the listener bins must produce key-downs and decoded characters, with and without the ring copy."""
    n, fs, L, S, B = 2048, 192000, 6, 24, 10
    spec = synth.StreamSpec(sample_rate=fs, block_size=n, n_blocks=120, seed=31,
                            tones=synth.make_tones(np.random.default_rng(31), L, n, 70, wpm_range=(28.0, 36.0)))
    iq = synth.generate(spec)
    src = np.stack([iq.reshape(120, 2 * n), iq.reshape(120, 2 * n)[::-1].copy()])  # two templates
    bins = np.array([[t.bin for t in spec.tones]] * 2, np.int32)
    with capi.Engine(n, max_streams=S, max_listeners=L, max_blocks_per_batch=S * B, max_peaks_per_flush=64, n_slots=2) as eng:
        h = host.RealtimeHarness(eng, fs, n, L, S, B, 3, src, bins, debounce=1)
        a = h.run(S, 14, ring_copy=True)
        b = h.run(S // 2, 14, ring_copy=False)
        h.close()
    assert a["ticks"] == S * B * L * 12 and b["ticks"] == (S // 2) * B * L * 12  # 12 steady-state batches of 14
    assert a["key_downs"] > 0.05 * a["ticks"] and a["chars"] > 0
    assert a["batch_s"] > 0 and a["gpu_ms"] > 0 and b["copy_s"] < a["copy_s"] + 1e-3
