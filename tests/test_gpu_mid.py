"""The fused single-pass kernels for N = 4096 / 8192 (csrc/k1_mid4k.cuh, k1_mid8k.cuh) against the oracle and against the
kernels that serve the same sizes without them (the three-pass K1 at 4096, the two-kernel large-block path at 8192;
selected with SDR_K1_MID4K=0 / SDR_K1_MID8K=0)."""
import numpy as np
import pytest

import parity_util as pu
import test_gpu_parity as tp
from sdrainer_b200 import synth

pytestmark = pytest.mark.gpu

NAMES = ("psd_noise_floor", "noise_variance", "thresholds", "taps", "keys", "flush_cum", "flush_n_peaks")


def test_mid_4096_batch_against_oracle(capi, oracle, monkeypatch):
    monkeypatch.delenv("SDR_K1_MID4K", raising=False)
    n, fs = 4096, 384000
    rng = np.random.default_rng(4096)
    tones = synth.make_tones(rng, 30, n, 70, wpm_range=(18.0, 28.0))
    spec = synth.StreamSpec(sample_rate=fs, block_size=n, n_blocks=230, seed=41, tones=tones)
    iq = synth.generate(spec)
    bins = [t.bin for t in tones]
    outs = tp._run_batch(capi, spec, iq, bins)
    tp._compare_with_oracle(oracle, spec, iq, bins, outs)


@pytest.mark.parametrize("edge", [0, 70, 300, 1000])
def test_mid_4096_edge_widths_and_ragged_bit_identity(capi, oracle, monkeypatch, edge):
    monkeypatch.delenv("SDR_K1_MID4K", raising=False)
    n, fs = 4096, 384000
    spec = tp._spec(n, fs, 237, seed=edge + 5, k=6)
    iq = synth.generate(spec)
    lo, hi = edge + 5, n - edge - 5
    bins = sorted({min(max(t.bin, lo), hi - 1) for t in spec.tones})
    one = tp._run_batch(capi, spec, iq, bins, edge=edge)
    ragged = tp._run_batch(capi, spec, iq, bins, edge=edge, chunks=[37, 1, 63, 100, 29, 7], n_slots=1)
    for name in NAMES:
        a, b = tp._concat(one, name), tp._concat(ragged, name)
        assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True), name
    r = oracle.process_stream(iq, n, edge_width=edge, peak_threshold=15.0, listener_bins=bins, sample_rate=fs)
    pu.check_scalars(tp._concat(one, "psd_noise_floor"), r.noise[:, 0], what="psdNoiseFloor")
    pu.check_scalars(tp._concat(one, "noise_variance"), r.noise[:, 1], rel=2e-4, what="noise variance")  # narrow windows (17 bins at the limit): 1.1e-4 measured
    pu.check_keys(tp._concat(one, "keys")[:, :len(bins)], r.taps, r.thresholds[:, 0] + r.thresholds[:, 1])


def _agree(a, b, n_listen):
    fa, fb = a.psd_noise_floor, b.psd_noise_floor
    assert np.abs(fa - fb).max() <= 3e-6 * np.abs(fa).max()
    flips = np.argwhere(a.keys != b.keys)
    thr = a.thresholds[:, 3]
    for blk, l in flips:
        assert abs(float(a.taps[blk, l]) - float(thr[blk])) < 1e-3
    assert len(flips) <= 3
    assert np.array_equal(a.flush_n_peaks, b.flush_n_peaks)
    assert np.abs(a.flush_cum - b.flush_cum).max() < 0.5
    strong = a.taps[:, :n_listen] > np.median(a.taps[:, :n_listen]) + 15
    if strong.any():
        assert np.abs(a.taps[:, :n_listen][strong] - b.taps[:, :n_listen][strong]).max() < 1e-3


def test_mid_4096_kernels_agree(capi, monkeypatch):
    """N = 4096 has two spectral kernels: the TMA-staged k1_mid4k_kernel (default) and the three-pass kernel
    (SDR_K1_MID4K=0)"""
    n, fs, nb = 4096, 384000, 120
    rng = np.random.default_rng(77)
    specs = [synth.StreamSpec(sample_rate=fs, block_size=n, n_blocks=nb, seed=70 + i,
                              tones=synth.make_tones(rng, 10, n, 70)) for i in range(4)]
    iqs = [synth.generate(sp) for sp in specs]
    res, names = [], []
    for env in ({}, {"SDR_K1_MID4K": "0"}):
        monkeypatch.delenv("SDR_K1_MID4K", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with capi.Engine(n, max_streams=4, max_listeners=16, max_blocks_per_batch=4 * nb, max_peaks_per_flush=n // 2 + 1) as eng:
            ss = [eng.open_stream(fs) for _ in specs]
            works = [dict(stream=s, iq=x, listener_bins=[t.bin for t in sp.tones]) for s, x, sp in zip(ss, iqs, specs)]
            res.append(eng.collect(eng.submit(works, capi.WANT_FLUSH_CUM)))
            names.append(eng.last_kernel())
    assert names == ["k1_mid4k_kernel", "k1_spectral_kernel<4096>"]
    _agree(res[0], res[1], 10)


def test_mid_8192_many_streams_agrees_with_two_kernel_path_and_oracle(capi, oracle, monkeypatch):
    """k1_mid8k2_kernel takes N = 8192 launches with >= 49 segments, the two-kernel path the others (and all of them with
    SDR_K1_MID8K=0); 320 streams x 104 blocks cross a cumulation boundary (two segments per stream, state rows ping-pong)"""
    n, fs, nb, ns = 8192, 768000, 104, 320
    rng = np.random.default_rng(8)
    base = [synth.generate(synth.StreamSpec(sample_rate=fs, block_size=n, n_blocks=nb, seed=800 + i,
                                            tones=synth.make_tones(rng, 8, n, 70))) for i in range(4)]
    binss = [np.sort(rng.choice(np.arange(80, n - 80), size=6, replace=False)).astype(np.int32) for _ in range(ns)]
    res, names = [], []
    for mid in ("1", "0"):
        if mid == "1":
            monkeypatch.delenv("SDR_K1_MID8K", raising=False)
        else:
            monkeypatch.setenv("SDR_K1_MID8K", "0")
        with capi.Engine(n, max_streams=ns, max_listeners=8, max_blocks_per_batch=ns * nb, max_peaks_per_flush=256) as eng:
            ss = [eng.open_stream(fs) for _ in range(ns)]
            works = [dict(stream=ss[i], iq=base[i % 4], listener_bins=binss[i]) for i in range(ns)]
            first = eng.collect(eng.submit([dict(w, iq=w["iq"][:2 * n * 50]) for w in works]))  # 50 blocks, then 54
            keep = {k: np.array(getattr(first, k)) for k in ("psd_noise_floor", "keys", "taps")}
            second = eng.collect(eng.submit([dict(w, iq=w["iq"][2 * n * 50:]) for w in works], capi.WANT_FLUSH_CUM))
            res.append((keep, second))
            names.append(eng.last_kernel())
    assert names == ["k1_mid8k2_kernel", "fast_cols32_kernel + fast_rows256_kernel"]
    (k1, s1), (k0, s0) = res
    assert np.abs(k1["psd_noise_floor"] - k0["psd_noise_floor"]).max() <= 3e-6 * np.abs(k0["psd_noise_floor"]).max()
    assert (k1["keys"] != k0["keys"]).sum() <= 6
    _agree(s1, s0, 6)
    assert s1.n_flushes == ns
    for i in (0, 1, ns - 1):
        r = oracle.process_stream(base[i % 4], n, listener_bins=list(binss[i]), sample_rate=fs)
        lo, hi = s1.work_block_offset[i], s1.work_block_offset[i + 1]
        pu.check_scalars(s1.psd_noise_floor[lo:hi], r.noise[50:, 0], what="psdNoiseFloor")
        fl = s1.work_flush_offset[i]
        assert np.abs(s1.flush_cum[fl] - r.flush_cum[0]).max() < 0.5


def test_mid8k2_small_launch_matches_oracle(capi, oracle, monkeypatch):
    """k1_mid8k2_kernel (two decoupled 256-thread groups, window sums finished by large_nf_finish_kernel) forced onto a
    launch of 7 streams x 130 blocks in two submits, against the oracle"""
    n, fs, nb, ns = 8192, 768000, 130, 7
    rng = np.random.default_rng(82)
    specs = [synth.StreamSpec(sample_rate=fs, block_size=n, n_blocks=nb, seed=820 + i, tones=synth.make_tones(rng, 12, n, 70)) for i in range(ns)]
    iqs = [synth.generate(sp) for sp in specs]
    monkeypatch.setenv("SDR_K1_MID8K", "force")
    with capi.Engine(n, max_streams=ns, max_listeners=16, max_blocks_per_batch=ns * nb, max_peaks_per_flush=256) as eng:
        ss = [eng.open_stream(fs) for _ in specs]
        works = [dict(stream=s, iq=x, listener_bins=[t.bin for t in sp.tones]) for s, x, sp in zip(ss, iqs, specs)]
        eng.collect(eng.submit([dict(w, iq=w["iq"][:2 * n * 40]) for w in works]))
        b1 = eng.collect(eng.submit([dict(w, iq=w["iq"][2 * n * 40:]) for w in works], capi.WANT_FLUSH_CUM))
        assert eng.last_kernel() == "k1_mid8k2_kernel"
    for i in (0, ns - 1):
        r = oracle.process_stream(iqs[i], n, listener_bins=[t.bin for t in specs[i].tones], sample_rate=fs)
        lo, hi = b1.work_block_offset[i], b1.work_block_offset[i + 1]
        pu.check_scalars(b1.psd_noise_floor[lo:hi], r.noise[40:, 0], what="psdNoiseFloor")
        pu.check_scalars(b1.noise_variance[lo:hi], r.noise[40:, 1], what="noise variance")
        pu.check_keys(b1.keys[lo:hi, :12], r.taps[40:], (r.thresholds[:, 0] + r.thresholds[:, 1])[40:])
        fl = b1.work_flush_offset[i]
        pu.check_peaks(pu.peak_keys(b1.peaks(fl)), [p.key() for p in r.peaks[0]], r.flush_cum[0], r.thresholds[99, 2])
