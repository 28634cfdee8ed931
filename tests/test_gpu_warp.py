"""The warp-per-block kernel for N = 512 (csrc/k1_warp.cuh, the TCI / Kiwi block shape) against the oracle and against
the three-pass kernel it replaces (SDR_K1_WARP=0)."""
import numpy as np
import pytest

import parity_util as pu
import test_gpu_parity as tp
from sdrainer_b200 import synth

pytestmark = pytest.mark.gpu

N, FS = 512, 48000
NAMES = ("psd_noise_floor", "noise_variance", "thresholds", "taps", "keys", "flush_cum", "flush_n_peaks")


def test_warp_analytic_tones(capi, monkeypatch):
    monkeypatch.delenv("SDR_K1_WARP", raising=False)
    for k in (0, 1, 2, 129, 255, 256, 257, 511):
        t = np.arange(N)
        x = np.exp(2j * np.pi * k * t / N)
        iq = np.empty(2 * N, np.float32)
        iq[0::2], iq[1::2] = x.real, x.imag
        with capi.Engine(N, max_blocks_per_batch=4) as eng:
            _, psd = eng.iq_to_spectrum_and_psd(iq)
        kk = (k + N // 2) % N
        assert int(np.argmax(psd[0])) == kk
        assert abs(psd[0, kk] / float(N) ** 2 - 1.0) < 1e-5
        assert np.delete(psd[0], kk).max() < 1e-6 * psd[0, kk]


@pytest.mark.parametrize("edge", [0, 1, 6, 70, 100, 170, 171, 172, 200])
def test_warp_edge_widths_ragged_identity_and_oracle(capi, oracle, monkeypatch, edge):
    """window sizes down to the warp kernel's limit (ws >= 17, e <= 171); 172 and 200 fall back to the three-pass kernel"""
    monkeypatch.delenv("SDR_K1_WARP", raising=False)
    spec = tp._spec(N, FS, 237, seed=edge + 50, k=5)
    iq = synth.generate(spec)
    lo, hi = edge + 5, N - edge - 5
    bins = sorted({min(max(t.bin, lo), hi - 1) for t in spec.tones})
    one = tp._run_batch(capi, spec, iq, bins, edge=edge)
    ragged = tp._run_batch(capi, spec, iq, bins, edge=edge, chunks=[37, 1, 63, 100, 29, 7], n_slots=1)
    for name in NAMES:
        a, b = tp._concat(one, name), tp._concat(ragged, name)
        assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True), name
    r = oracle.process_stream(iq, N, edge_width=edge, peak_threshold=15.0, listener_bins=bins, sample_rate=FS)
    pu.check_scalars(tp._concat(one, "psd_noise_floor"), r.noise[:, 0], what="psdNoiseFloor")
    pu.check_scalars(tp._concat(one, "noise_variance"), r.noise[:, 1], rel=2e-4, what="noise variance")  # narrow windows (17 bins at the limit): 1.1e-4 measured
    pu.check_keys(tp._concat(one, "keys")[:, :len(bins)], r.taps, r.thresholds[:, 0] + r.thresholds[:, 1])


def test_warp_agrees_with_three_pass_many_streams_and_listeners(capi, monkeypatch):
    """two factorizations of the same DFT; 37 streams (ragged warp/CTA tail), 40 listeners (> 32: the re-read path)"""
    nb, ns = 120, 37
    rng = np.random.default_rng(5)
    specs = [synth.StreamSpec(sample_rate=FS, block_size=N, n_blocks=nb, seed=500 + i,
                              tones=synth.make_tones(rng, 5, N, 70)) for i in range(ns)]
    iqs = [synth.generate(sp) for sp in specs]
    binss = [[t.bin for t in sp.tones] + list(range(90 + i, 90 + i + 35)) for i, sp in enumerate(specs)]
    res = []
    for sel in ("1", "0"):
        monkeypatch.setenv("SDR_K1_WARP", sel)
        with capi.Engine(N, max_streams=ns, max_listeners=40, max_blocks_per_batch=ns * nb, max_peaks_per_flush=257) as eng:
            ss = [eng.open_stream(FS) for _ in specs]
            works = [dict(stream=s, iq=x, listener_bins=b) for s, x, b in zip(ss, iqs, binss)]
            res.append(eng.collect(eng.submit(works, capi.WANT_FLUSH_CUM)))
    a, b = res
    assert np.abs(a.psd_noise_floor - b.psd_noise_floor).max() <= 3e-6 * np.abs(b.psd_noise_floor).max()
    flips = np.argwhere(a.keys != b.keys)
    thr = a.thresholds[:, 3]
    for blk, l in flips:
        assert abs(float(a.taps[blk, l]) - float(thr[blk])) < 1e-3
    assert len(flips) <= 6
    assert np.array_equal(a.flush_n_peaks, b.flush_n_peaks)
    assert np.abs(a.flush_cum - b.flush_cum).max() < 0.5
    strong = a.taps[:, :5] > np.median(a.taps[:, :5]) + 15
    assert np.abs(a.taps[:, :5][strong] - b.taps[:, :5][strong]).max() < 1e-3
