"""The single-pass wideband kernel for N = 65536 (csrc/k1_wide.cuh: teams of 16 CTAs, tensor-map TMA tiles, the four-step
intermediate in an L2-resident ring) against the two-kernel large-block path (SDR_K1_WIDE=0) and the oracle."""
import numpy as np
import pytest

import parity_util as pu
from sdrainer_b200 import synth

pytestmark = pytest.mark.gpu


def test_wide_65536_agrees_with_two_kernel_path_and_oracle(capi, oracle, monkeypatch):
    """9 streams x 104 blocks: every stream crosses a cumulation boundary (two segments, state rows ping-pong), the
    batch is submitted in two parts (50 + 54 blocks) so that the cumulation is carried through cum_state"""
    n, fs, nb, ns = 65536, 24576000, 104, 9
    rng = np.random.default_rng(65)
    tones = synth.make_tones(rng, 40, n, 70, keyed=False)
    base = [synth.generate(synth.StreamSpec(sample_rate=fs, block_size=n, n_blocks=nb, seed=650 + i, tones=tones)) for i in range(2)]
    binss = [np.sort(rng.choice(np.arange(80, n - 80), size=5, replace=False)).astype(np.int32) for _ in range(ns)]
    binss[0] = np.array(sorted(t.bin for t in tones[:5]), np.int32)
    res = []
    for sel in ("force", "0"):
        monkeypatch.setenv("SDR_K1_WIDE", sel)
        with capi.Engine(n, max_streams=ns, max_listeners=8, max_blocks_per_batch=ns * nb, max_peaks_per_flush=512) as eng:
            ss = [eng.open_stream(fs) for _ in range(ns)]
            works = [dict(stream=ss[i], iq=base[i % 2], listener_bins=binss[i]) for i in range(ns)]
            first = eng.collect(eng.submit([dict(w, iq=w["iq"][:2 * n * 50]) for w in works]))
            keep = {k: np.array(getattr(first, k)) for k in ("psd_noise_floor", "noise_variance", "keys", "taps")}
            second = eng.collect(eng.submit([dict(w, iq=w["iq"][2 * n * 50:]) for w in works], capi.WANT_FLUSH_CUM))
            res.append((keep, second))
    (k1, s1), (k0, s0) = res
    for name in ("psd_noise_floor", "noise_variance"):
        assert np.abs(k1[name] - k0[name]).max() <= 3e-6 * np.abs(k0[name]).max(), name
    assert (k1["keys"] != k0["keys"]).sum() <= 4
    assert np.abs(s1.psd_noise_floor - s0.psd_noise_floor).max() <= 3e-6 * np.abs(s0.psd_noise_floor).max()
    assert np.array_equal(s1.flush_n_peaks, s0.flush_n_peaks)
    assert s1.n_flushes == ns
    # two fp32 evaluations of the same transform (the wide kernel forms W_N^(c k) as a product of two table values):
    # noise-level bins next to 40 carriers differ by < 0.01 dB per block; each path is held to the oracle separately
    assert np.abs(s1.flush_cum - s0.flush_cum).max() < 1.5
    strong = s0.taps[:, :5] > np.median(s0.taps[:, :5]) + 15
    assert strong.any() and np.abs(s1.taps[:, :5][strong] - s0.taps[:, :5][strong]).max() < 1e-3
    for i in (0, ns - 1):
        r = oracle.process_stream(base[i % 2], n, listener_bins=list(binss[i]), sample_rate=fs)
        lo, hi = s1.work_block_offset[i], s1.work_block_offset[i + 1]
        pu.check_scalars(s1.psd_noise_floor[lo:hi], r.noise[50:, 0], what="psdNoiseFloor")
        pu.check_scalars(s1.noise_variance[lo:hi], r.noise[50:, 1], what="noise variance")
        fl = s1.work_flush_offset[i]
        # sums of 100 dB values: bins at the noise level next to 40 carriers carry the fp32-FFT error of a 65536-point
        # transform (both GPU paths agree with each other to < 0.5 above); bins >= 10 dB over the median are tight
        d = np.abs(s1.flush_cum[fl] - r.flush_cum[0])
        assert d.max() < 5.0
        loud = r.flush_cum[0] > np.median(r.flush_cum[0]) + 1000.0
        assert loud.any() and d[loud].max() < 0.05
        pu.check_peaks(pu.peak_keys(s1.peaks(fl)), [p.key() for p in r.peaks[0]], r.flush_cum[0], r.thresholds[99, 2])


def test_segment_sequential_rows_kernel_equals_block_parallel_path(capi, monkeypatch):
    """N = 65536 with >= 37 segments: fast_rows256_seg_kernel (cumulation in registers) against fast_rows256_kernel +
    large_round_cum_kernel (SDR_LARGE_NO_SEGROWS=1): the same arithmetic in the same order -> identical bits.  101 blocks
    per stream in two submits (60 + 41): the second closes the window (flush) and leaves one block in the next one"""
    monkeypatch.setenv("SDR_K1_WIDE", "0")
    n, fs, nb, ns = 65536, 24576000, 101, 40
    rng = np.random.default_rng(66)
    base = [(rng.standard_normal(nb * 2 * n) * 1e-3).astype(np.float32) for _ in range(3)]
    binss = [np.sort(rng.choice(np.arange(80, n - 80), size=4, replace=False)).astype(np.int32) for _ in range(ns)]
    names = ("psd_noise_floor", "noise_variance", "taps", "keys", "thresholds", "flush_cum", "flush_n_peaks")
    res = []
    for sel in (None, "1"):
        if sel is None:
            monkeypatch.delenv("SDR_LARGE_NO_SEGROWS", raising=False)
        else:
            monkeypatch.setenv("SDR_LARGE_NO_SEGROWS", sel)
        with capi.Engine(n, max_streams=ns, max_listeners=4, max_blocks_per_batch=ns * 60, max_peaks_per_flush=64) as eng:
            ss = [eng.open_stream(fs) for _ in range(ns)]
            works = [dict(stream=ss[i], iq=base[i % 3], listener_bins=binss[i]) for i in range(ns)]
            a = eng.collect(eng.submit([dict(w, iq=w["iq"][:2 * n * 60]) for w in works], capi.WANT_FLUSH_CUM))
            keep = {k: np.array(getattr(a, k)) for k in names}
            b = eng.collect(eng.submit([dict(w, iq=w["iq"][2 * n * 60:]) for w in works], capi.WANT_FLUSH_CUM))
            cum = [eng.cumulation_count(s) for s in ss]
            res.append((keep, {k: np.array(getattr(b, k)) for k in names}, cum, b.n_flushes))
    (k1, b1, c1, f1), (k0, b0, c0, f0) = res
    assert c1 == c0 == [1] * ns and f1 == f0 == ns
    for d1, d0 in ((k1, k0), (b1, b0)):
        for name in names:
            assert np.array_equal(d1[name], d0[name], equal_nan=True), name


@pytest.mark.parametrize("lookahead", ["3", "4", "6"])
def test_wide_many_segments_few_teams_ragged_bit_identity(capi, monkeypatch, lookahead):
    """more segments than co-resident teams (every team walks several segments, of unequal length), listeners on every
    row tile, any lookahead depth: results must not depend on how the batch is cut into submits"""
    monkeypatch.setenv("SDR_K1_WIDE", "force")
    monkeypatch.setenv("SDR_K1_WIDE_LOOKAHEAD", lookahead)
    n, fs, ns = 65536, 24576000, 23
    rng = np.random.default_rng(67)
    lens = [int(x) for x in rng.integers(1, 9, size=ns)]
    base = (rng.standard_normal(8 * 2 * n) * 1e-3).astype(np.float32)
    binss = [np.sort(rng.choice(np.arange(80, n - 80), size=6, replace=False)).astype(np.int32) for _ in range(ns)]
    names = ("psd_noise_floor", "noise_variance", "taps", "keys", "thresholds")
    outs = []
    for split in (False, True):
        with capi.Engine(n, max_streams=ns, max_listeners=8, max_blocks_per_batch=ns * 8, max_peaks_per_flush=64) as eng:
            ss = [eng.open_stream(fs) for _ in range(ns)]
            if not split:
                works = [dict(stream=ss[i], iq=base[:2 * n * lens[i]], listener_bins=binss[i]) for i in range(ns)]
                r = eng.collect(eng.submit(works))
                outs.append({k: [np.array(getattr(r, k))[r.work_block_offset[i]:r.work_block_offset[i + 1]] for i in range(ns)] for k in names})
            else:
                per = {k: [[] for _ in range(ns)] for k in names}
                for part in range(2):  # first block alone, then the rest: the cumulation is carried through cum_state
                    idx = [i for i in range(ns) if (part == 0 or lens[i] > 1)]
                    works = [dict(stream=ss[i], iq=(base[:2 * n] if part == 0 else base[2 * n:2 * n * lens[i]]), listener_bins=binss[i]) for i in idx]
                    r = eng.collect(eng.submit(works))
                    for j, i in enumerate(idx):
                        for k in names:
                            per[k][i].append(np.array(getattr(r, k))[r.work_block_offset[j]:r.work_block_offset[j + 1]])
                outs.append({k: [np.concatenate(v, axis=0) for v in per[k]] for k in names})
    for k in names:
        for i in range(ns):
            assert np.array_equal(outs[0][k][i], outs[1][k][i], equal_nan=True), (k, i)


@pytest.mark.parametrize("n,env,edge", [
    (65536, {"SDR_K1_WIDE": "force"}, 0), (65536, {"SDR_K1_WIDE": "force"}, 1000), (65536, {"SDR_K1_WIDE": "force"}, 10000),
    (65536, {"SDR_K1_WIDE": "force"}, 30100),   # 533-bin windows: two whole positions per row
    (65536, {"SDR_K1_WIDE": "force"}, 31000),   # 353-bin windows: one or two positions per row (clamped body)
    (65536, {"SDR_K1_WIDE": "force"}, 32700),   # 13-bin windows: at most one position per row
    (8192, {"SDR_K1_MID8K": "force"}, 0), (8192, {"SDR_K1_MID8K": "force"}, 1000), (8192, {"SDR_K1_MID8K": "force"}, 3500),
    (8192, {"SDR_K1_MID8K": "force"}, 4000),    # 19-bin windows
    (4096, {}, 1500), (4096, {}, 1950), (4096, {}, 2003),  # k1_mid4k: 109-, 19- and 9-bin windows
])
def test_noise_window_geometry_over_edge_widths(capi, oracle, monkeypatch, n, env, edge):
    """The single-pass kernels split every noise window into per-row shares read from a common start position, the upper
    half-warp rotated by one element (k1_large.cuh: nf_row_share); the geometry depends on the edge width.  Noise floor and
    variance against the oracle from the widest to the narrowest windows that do not take the exact replay (>= 9 bins)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    fs = {65536: 24576000, 8192: 768000, 4096: 384000}[n]
    nb = 6 if n == 65536 else 12
    ws = (n - 2 * edge) // 10
    assert ws >= 9
    rng = np.random.default_rng(n + edge)
    tones = synth.make_tones(rng, max(2, min(12, (n - 2 * edge - 26) // 14)), n, edge + 8, keyed=False)
    iq = synth.generate(synth.StreamSpec(sample_rate=fs, block_size=n, n_blocks=nb, seed=n + edge, tones=tones))
    bins = sorted({min(max(t.bin, edge + 5), n - edge - 6) for t in tones})[:6]
    with capi.Engine(n, max_streams=1, max_listeners=8, max_blocks_per_batch=nb, max_peaks_per_flush=256) as eng:
        sid = eng.open_stream(fs)
        res = eng.collect(eng.submit([dict(stream=sid, iq=iq, edge_width=edge, listener_bins=bins)]))
        kernel = eng.last_kernel()
    assert kernel == {65536: "k1_wide_kernel", 8192: "k1_mid8k2_kernel", 4096: "k1_mid4k_kernel"}[n]
    r = oracle.process_stream(iq, n, edge_width=edge, peak_threshold=15.0, listener_bins=bins, sample_rate=fs)
    pu.check_scalars(res.psd_noise_floor, r.noise[:, 0], what=f"psdNoiseFloor, {ws}-bin windows")
    # narrow windows: the fp32 error of the few bins of a window does not average out (1.1e-4 measured with 17 bins)
    pu.check_scalars(res.noise_variance, r.noise[:, 1], rel=1e-4 if ws >= 100 else 5e-4, what=f"noise variance, {ws}-bin windows")


@pytest.mark.parametrize("n,env,kernel", [
    (512, {}, "k1_warp_kernel"), (2048, {}, "k1_spectral_kernel<2048>"), (4096, {}, "k1_mid4k_kernel"),
    (8192, {"SDR_K1_MID8K": "force"}, "k1_mid8k2_kernel"), (8192, {"SDR_K1_MID8K": "0"}, "fast_cols32_kernel + fast_rows256_kernel"),
    (16384, {}, "fast_cols64_kernel + fast_rows256_kernel"), (32768, {}, "fast_cols64_kernel + fast_rows256_kernel"),
    (65536, {"SDR_K1_WIDE": "force"}, "k1_wide_kernel"), (65536, {"SDR_K1_WIDE": "0"}, "fast_cols256_kernel + fast_rows256_kernel"),
])
def test_window_table_through_every_spectral_kernel(capi, oracle, monkeypatch, n, env, kernel):
    """The optional window table (an extension: the reference applies none) multiplies in float32 before the transform in
    every spectral kernel of the submit path: noise floor against the oracle run with the same window"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    fs = 48000 * n // 512
    nb = 8 if n >= 16384 else 24
    rng = np.random.default_rng(n)
    tones = synth.make_tones(rng, 8, n, 70, keyed=False)
    iq = synth.generate(synth.StreamSpec(sample_rate=fs, block_size=n, n_blocks=nb, seed=n + 9, tones=tones))
    win = (np.hanning(n) + 0.01).astype(np.float32)
    bins = [t.bin for t in tones][:6]
    with capi.Engine(n, window=win, max_streams=1, max_listeners=8, max_blocks_per_batch=nb, max_peaks_per_flush=256) as eng:
        sid = eng.open_stream(fs)
        res = eng.collect(eng.submit([dict(stream=sid, iq=iq, listener_bins=bins)]))
        assert eng.last_kernel() == kernel
    r = oracle.process_stream(iq, n, window=win, listener_bins=bins, sample_rate=fs)
    pu.check_scalars(res.psd_noise_floor, r.noise[:, 0], what="psdNoiseFloor with a window")
    strong = r.taps > r.thresholds[:, :1] + 15  # the listeners sit on the tones: well above the mean noise floor (dB)
    assert strong.any() and np.abs(res.taps[:, :len(bins)][strong] - r.taps[strong]).max() < 2e-3
