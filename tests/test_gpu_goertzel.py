"""K3 Goertzel / envelope bank against the oracle (dsp.Goertzel dsp/dsp.go:34-136, cw/audio.go:184-203)."""
import ctypes as C
import math

import numpy as np
import pytest

from sdrainer_b200 import synth

pytestmark = pytest.mark.gpu


def _oracle_audio(oracle, pitch, fs, audio, scale):
    L = oracle.lib()
    d = oracle.AudioDemod()
    L.orc_audio_demod_init(C.byref(d), pitch, fs)
    d.scale = scale
    bs = d.filter.blocksize
    mags, states = [], []
    for b in range(audio.size // bs):
        blk = np.ascontiguousarray(audio[b * bs:(b + 1) * bs]).copy()
        m, s, deb = C.c_double(), C.c_int(), C.c_int()
        assert L.orc_audio_demod_block(C.byref(d), blk.ctypes.data_as(C.POINTER(C.c_float)), bs, C.byref(m), C.byref(s),
                                       C.byref(deb)) == 0
        mags.append(m.value)
        states.append(s.value)
    return np.asarray(mags), np.asarray(states, np.uint8), bs


def _keyed_audio(pitch, fs, n, amp, wpm, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    env = synth.keying("cq de dl1abc k", wpm, fs, n, 0.1)
    x = amp * np.cos(2 * np.pi * pitch * t) * env + rng.standard_normal(n) * 0.02
    return x.astype(np.float32)


def test_audio_bank_bit_exact(capi, oracle):
    """float64 recurrence with the reference's operation order: magnitudes and states are bit-identical"""
    fs = 48000
    pitches = [700.0, 600.0, 850.0, 350.0]
    scales = [1.0, 0.0, 3.0, 1.0]  # none, autoscale, fixed gain with clipping, none
    bank = capi.GoertzelBank(pitches, fs, max_blocks=600)
    audios = []
    for i, p in enumerate(pitches):
        bs = bank.blocksize(i)
        audios.append(_keyed_audio(p, fs, bs * (400 + 13 * i), 0.4 + 0.2 * i, 18 + 3 * i, seed=i))
    assert bank.blocksize(0) == 207
    mag, st = bank.process_audio(audios, scale=scales)
    for i, p in enumerate(pitches):
        omag, ost, bs = _oracle_audio(oracle, p, fs, audios[i], scales[i])
        assert bs == bank.blocksize(i)
        assert np.array_equal(mag[i], omag), f"filter {i}"
        assert np.array_equal(st[i], ost)
    # the running magnitudeLimit is carried across calls (dsp/dsp.go:115-120): two halves == one call
    bank2 = capi.GoertzelBank(pitches[:1], fs, max_blocks=600)
    half = (audios[0].size // 207 // 2) * 207
    m1, s1 = bank2.process_audio([audios[0][:half]])
    m2, s2 = bank2.process_audio([audios[0][half:]])
    omag, ost, _ = _oracle_audio(oracle, 700.0, fs, audios[0], 1.0)
    assert np.array_equal(np.concatenate([m1[0], m2[0]]), omag)
    assert np.array_equal(np.concatenate([s1[0], s2[0]]), ost)


def test_audio_bank_reference_signal_state_table(capi):
    """dsp/dsp_test.go:25-149 deterministic rows through the GPU bank"""
    fs = 48000

    def sine(n, amp, f):
        out = np.empty(n, np.float32)
        t = 0.0
        for i in range(n):
            out[i] = np.float32(amp * math.cos(2 * math.pi * f * t))
            t += 1.0 / fs
        return out

    bank = capi.GoertzelBank([700.0, 350.0, 700.0, 700.0], fs, max_blocks=16)
    bs = [bank.blocksize(i) for i in range(4)]
    for blocks in (1, 10):
        audio = [sine(blocks * bs[0], 1, 700.0), sine(blocks * bs[1], 1, 700.0), np.zeros(blocks * bs[2], np.float32),
                 np.full(blocks * bs[3], 0.8, np.float32)]
        b = capi.GoertzelBank([700.0, 350.0, 700.0, 700.0], fs, max_blocks=16)
        _, st = b.process_audio(audio)
        assert [bool(s.any()) for s in st] == [True, False, False, False]


@pytest.mark.parametrize("n", [512, 2048, 8192])
def test_iq_bank_equals_fft_bin(capi, oracle, n):
    """a block-length Goertzel at a bin centre equals spectrum[l.SignalBin()] (rx/receiver.go:393)"""
    rng = np.random.default_rng(n)
    tones = synth.make_tones(rng, 12, n, 70, keyed=False)
    spec = synth.StreamSpec(sample_rate=int(93.75 * n), block_size=n, n_blocks=6, seed=n, tones=tones)
    iq = synth.generate(spec)
    bins = [t.bin for t in tones] + [80, n // 2 + 1, n - 90]
    bank = capi.GoertzelBank([700.0], 48000)
    got = bank.process_iq(iq, n, bins)
    r = oracle.process_stream(iq, n, listener_bins=bins, sample_rate=spec.sample_rate)
    strong = np.zeros_like(r.taps, dtype=bool)
    strong[:, :len(tones)] = True
    assert np.abs(got[strong] - r.taps[strong]).max() < 1e-3
    lin_g, lin_r = 10 ** (got.astype(np.float64) / 10), 10 ** (r.taps.astype(np.float64) / 10)
    assert (np.abs(lin_g - lin_r) <= 1e-5 * lin_r.max(axis=1, keepdims=True)).all()
    if n <= 4096:
        with capi.Engine(n, max_listeners=16, max_blocks_per_batch=8) as eng:
            s = eng.open_stream(spec.sample_rate)
            k1 = eng.collect(eng.submit([dict(stream=s, iq=iq, listener_bins=bins)]))
        assert np.abs(got[strong] - k1.taps[:, :len(bins)][strong]).max() < 1e-3


@pytest.mark.parametrize("n_bins", [1, 31, 32, 33, 64, 65, 96, 97, 128, 129, 200, 257])
def test_iq_bank_listener_counts_around_the_pass_boundaries(capi, n_bins):
    """goertzel_iq_kernel<LPT>: 32*LPT listeners per pass, the last pass runs 1..LPT chains -- every listener must get the
    DFT bin of ITS frequency whatever the count (checked against numpy's float64 FFT of the same float32 block)"""
    n, nb = 1024, 3
    rng = np.random.default_rng(n_bins)
    iq = (rng.standard_normal(nb * 2 * n) * 1e-2).astype(np.float32)
    bins = rng.choice(np.arange(n), size=n_bins, replace=False).astype(np.int32)
    bank = capi.GoertzelBank([700.0], 48000)
    got = bank.process_iq(iq, n, bins)
    x = iq.astype(np.float64).reshape(nb, n, 2)
    spec = np.fft.fftshift(np.fft.fft(x[..., 0] + 1j * x[..., 1], axis=1), axes=1)
    psd = np.abs(spec) ** 2
    want = 10 * np.log10(20 * psd[:, bins] / n ** 2) + 120
    assert got.shape == want.shape
    lin_g, lin_w = 10 ** (got.astype(np.float64) / 10), 10 ** (want / 10)
    assert (np.abs(lin_g - lin_w) <= 2e-5 * lin_w.max(axis=1, keepdims=True)).all()
    assert np.abs(got - want)[lin_w > 0.05 * lin_w.max()].max() < 2e-3
