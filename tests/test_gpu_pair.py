"""The warp-pair spectral kernel (csrc/k1_pair.cuh, N = 2048, selected with SDR_K1_PAIR=1) against the oracle and
against the three-pass kernel: same tolerances as tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest

import parity_util as pu
import test_gpu_parity as tp
from sdrainer_b200 import synth

pytestmark = pytest.mark.gpu

N = 2048
FS = 192000


@pytest.fixture()
def pair_env():
    """engine creation reads SDR_K1_PAIR / SDR_K1_PAIR_STAGES; restore the environment afterwards"""
    old = {k: os.environ.get(k) for k in ("SDR_K1_PAIR", "SDR_K1_PAIR_STAGES")}

    def select(pair, stages=1):
        os.environ["SDR_K1_PAIR"] = "1" if pair else "0"
        os.environ["SDR_K1_PAIR_STAGES"] = str(stages)

    yield select
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("stages", [1, 2])
def test_pair_spectrum_and_psd_parity(capi, oracle, pair_env, stages):
    """dsp.FFT.IQToSpectrumAndPSD (dsp/fft.go:23-37) through the radix-2 split + two 32x32 transforms"""
    pair_env(True, stages)
    spec = tp._spec(N, FS, 24, seed=77, k=10, off_center=0.3)
    iq = synth.generate(spec)
    with capi.Engine(N, max_blocks_per_batch=64) as eng:
        s_gpu, p_gpu = eng.iq_to_spectrum_and_psd(iq)
    r = oracle.process_stream(iq, N, want_spectrum=True)
    m = pu.check_spectrum(s_gpu, p_gpu, r.spectrum, r.psd)
    assert m["n_signal_bins"] > 0


def test_pair_analytic_tone_lands_on_its_bin(capi, pair_env):
    pair_env(True)
    for k in (0, 1, 517, 1023, 1024, 1025, 2047):
        t = np.arange(N)
        x = np.exp(2j * np.pi * k * t / N)
        iq = np.empty(2 * N, np.float32)
        iq[0::2], iq[1::2] = x.real, x.imag
        with capi.Engine(N, max_blocks_per_batch=4) as eng:
            _, psd = eng.iq_to_spectrum_and_psd(iq)
        kk = (k + N // 2) % N  # fftshift, dsp/fft.go:54-57
        assert int(np.argmax(psd[0])) == kk
        assert abs(psd[0, kk] / float(N) ** 2 - 1.0) < 1e-5
        rest = np.delete(psd[0], kk)
        assert rest.max() < 1e-6 * psd[0, kk]


@pytest.mark.parametrize("stages", [1, 2])
def test_pair_cfg2_batch_against_oracle(capi, oracle, pair_env, stages):
    """config 2: 192 kS/s, N=2048, 50 CW signals at varied SNR -- noise floor, thresholds, taps, keys, peaks"""
    pair_env(True, stages)
    spec = synth.config(2, seconds=3.3)
    iq = synth.generate(spec)
    bins = [t.bin for t in spec.tones]
    outs = tp._run_batch(capi, spec, iq, bins)
    tp._compare_with_oracle(oracle, spec, iq, bins, outs)


@pytest.mark.parametrize("edge", [0, 1, 24, 69, 70, 200, 300, 686, 689])
def test_pair_edge_widths(capi, oracle, pair_env, edge):
    """dsp.FindNoiseFloor (dsp/fft.go:215-252) for window sizes down to the pair kernel's limit (ws >= 67);
    (N-2e)%10 == 0 and != 0 both occur (9 or 10 evaluated windows)"""
    pair_env(True)
    spec = tp._spec(N, FS, 120, seed=1000 + edge, k=6)
    iq = synth.generate(spec)
    lo, hi = edge + 5, N - edge - 5
    bins = sorted({min(max(t.bin, lo), hi - 1) for t in spec.tones})
    outs = tp._run_batch(capi, spec, iq, bins, edge=edge)
    r = oracle.process_stream(iq, N, edge_width=edge, peak_threshold=15.0, listener_bins=bins, sample_rate=spec.sample_rate)
    pu.check_scalars(tp._concat(outs, "psd_noise_floor"), r.noise[:, 0], what="psdNoiseFloor")
    pu.check_scalars(tp._concat(outs, "noise_variance"), r.noise[:, 1], rel=2e-3, what="noise variance")
    keys = tp._concat(outs, "keys")[:, :len(bins)]
    pu.check_keys(keys, r.taps, r.thresholds[:, 0] + r.thresholds[:, 1])


def test_pair_falls_back_below_its_window_limit(capi, oracle, pair_env):
    """edge widths with (N-2e)/10 < 67 go through the three-pass kernel inside the same engine"""
    pair_env(True)
    spec = tp._spec(N, FS, 30, seed=5, k=3)
    iq = synth.generate(spec)
    for edge in (690, 800, 979):  # 979: (N-2e)/10 = 9, the narrowest supported window
        outs = tp._run_batch(capi, spec, iq, [1024], edge=edge)
        r = oracle.process_stream(iq, N, edge_width=edge, peak_threshold=15.0, listener_bins=[1024], sample_rate=spec.sample_rate)
        pu.check_scalars(tp._concat(outs, "psd_noise_floor"), r.noise[:, 0], what="psdNoiseFloor")
        pu.check_scalars(tp._concat(outs, "noise_variance"), r.noise[:, 1], rel=2e-3, what="noise variance")
    # narrower windows: dsp.FindNoiseFloor would evaluate more than 10 of them -> rejected, never computed wrongly
    for edge in (980, 1000, 1024, -1):
        with pytest.raises(capi.SdrError):
            tp._run_batch(capi, spec, iq, [1024], edge=edge)


def test_pair_ragged_batches_are_bit_identical_to_one_batch(capi, pair_env):
    """cumulation carried through cum_state across submits: not a bit may change (rx/receiver.go:404-407)"""
    pair_env(True)
    spec = tp._spec(N, FS, 337, seed=21, k=6)
    iq = synth.generate(spec)
    bins = [t.bin for t in spec.tones]
    one = tp._run_batch(capi, spec, iq, bins)
    ragged = tp._run_batch(capi, spec, iq, bins, chunks=[37, 1, 63, 100, 29, 7, 100], n_slots=1)
    for name in ("psd_noise_floor", "noise_variance", "thresholds", "taps", "keys", "flush_cum", "flush_n_peaks"):
        a, b = tp._concat(one, name), tp._concat(ragged, name)
        assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True), name


def test_pair_agrees_with_three_pass_kernel(capi, pair_env):
    """two independent factorizations of the same DFT: decisions identical, floats within fp32 FFT error"""
    spec = synth.config(2, seconds=2.2)
    iq = synth.generate(spec)
    bins = [t.bin for t in spec.tones] + list(range(100, 170))  # > 64 listeners: the re-read path of the taps
    res = []
    for pair in (False, True):
        pair_env(pair)
        res.append(tp._run_batch(capi, spec, iq, bins))
    a, b = res
    fa, fb = tp._concat(a, "psd_noise_floor"), tp._concat(b, "psd_noise_floor")
    assert np.abs(fa - fb).max() <= 2e-6 * np.abs(fa).max()
    ka, kb = tp._concat(a, "keys"), tp._concat(b, "keys")
    ta, tb = tp._concat(a, "taps"), tp._concat(b, "taps")
    thr = tp._concat(a, "thresholds")[:, 3]
    flips = np.argwhere(ka != kb)
    for blk, l in flips:  # only values sitting on the threshold may differ
        assert abs(float(ta[blk, l]) - float(thr[blk])) < 1e-3
    assert len(flips) <= 3
    assert np.array_equal(tp._concat(a, "flush_n_peaks"), tp._concat(b, "flush_n_peaks"))
    assert np.abs(tp._concat(a, "flush_cum") - tp._concat(b, "flush_cum")).max() < 0.5
    strong = ta > np.median(ta) + 15
    assert np.abs(ta[strong] - tb[strong]).max() < 1e-3


def test_pair_multi_stream_and_kiwi_wire_format(capi, oracle, pair_env):
    """several streams in one launch (segments walked with stride grid) and the fused int16 decode (kiwi/client.go:298-308)"""
    pair_env(True)
    rng = np.random.default_rng(9)
    nb = 130
    specs = [synth.StreamSpec(sample_rate=FS, block_size=N, n_blocks=nb, seed=40 + i,
                              tones=synth.make_tones(rng, 5, N, 70, amp_range=(0.02, 0.3)), noise_sigma=2e-3) for i in range(5)]
    xs = [synth.generate(s) for s in specs]
    raws = [np.frombuffer(np.clip(np.rint(x * 32767), -32768, 32767).astype(">i2").tobytes(), np.uint8).copy() for x in xs]
    floats = [(np.frombuffer(r.tobytes(), ">i2").astype(np.float32) / np.float32(32767)) for r in raws]
    outs = {}
    for fmt, datas in ((capi.FMT_KIWI_I16BE, raws), (capi.FMT_F32, floats)):
        with capi.Engine(N, max_streams=5, max_listeners=8, max_blocks_per_batch=5 * nb, max_peaks_per_flush=N // 2 + 1) as eng:
            ss = [eng.open_stream(FS) for _ in specs]
            works = [dict(stream=s, iq=d, listener_bins=[t.bin for t in sp.tones], format=fmt) for s, d, sp in zip(ss, datas, specs)]
            outs[fmt] = eng.collect(eng.submit(works, capi.WANT_FLUSH_CUM))
    a, b = outs[capi.FMT_KIWI_I16BE], outs[capi.FMT_F32]
    for name in ("psd_noise_floor", "noise_variance", "thresholds", "taps", "keys", "flush_cum", "flush_n_peaks"):
        assert np.array_equal(getattr(a, name), getattr(b, name), equal_nan=True), name
    for i, (sp, x) in enumerate(zip(specs, floats)):
        r = oracle.process_stream(x, N, listener_bins=[t.bin for t in sp.tones], sample_rate=FS)
        rel = np.abs(b.psd_noise_floor[i * nb:(i + 1) * nb].astype(np.float64) - r.noise[:, 0]) / r.noise[:, 0]
        assert rel.max() < 1e-4
