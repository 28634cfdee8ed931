"""The 16-CTA cluster kernel for N = 65536 (csrc/k1_cluster.cuh: the whole block in distributed shared memory) against the
two-kernel large-block path (SDR_K1_CLUSTER=0) and the oracle."""
import numpy as np
import pytest

import parity_util as pu
from sdrainer_b200 import synth

pytestmark = pytest.mark.gpu


def test_cluster_65536_agrees_with_two_kernel_path_and_oracle(capi, oracle, monkeypatch):
    """9 streams x 104 blocks: every stream crosses a cumulation boundary (two segments, state rows ping-pong), the
    batch is submitted in two parts (50 + 54 blocks) so that the cumulation is carried through cum_state"""
    n, fs, nb, ns = 65536, 24576000, 104, 9
    rng = np.random.default_rng(65)
    tones = synth.make_tones(rng, 40, n, 70, keyed=False)
    base = [synth.generate(synth.StreamSpec(sample_rate=fs, block_size=n, n_blocks=nb, seed=650 + i, tones=tones)) for i in range(2)]
    binss = [np.sort(rng.choice(np.arange(80, n - 80), size=5, replace=False)).astype(np.int32) for _ in range(ns)]
    binss[0] = np.array(sorted(t.bin for t in tones[:5]), np.int32)
    res = []
    for sel in ("1", "0"):
        monkeypatch.setenv("SDR_K1_CLUSTER", sel)
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
    assert np.abs(s1.flush_cum - s0.flush_cum).max() < 0.5
    strong = s0.taps[:, :5] > np.median(s0.taps[:, :5]) + 15
    assert strong.any() and np.abs(s1.taps[:, :5][strong] - s0.taps[:, :5][strong]).max() < 1e-3
    for i in (0, ns - 1):
        r = oracle.process_stream(base[i % 2], n, listener_bins=list(binss[i]), sample_rate=fs)
        lo, hi = s1.work_block_offset[i], s1.work_block_offset[i + 1]
        pu.check_scalars(s1.psd_noise_floor[lo:hi], r.noise[50:, 0], what="psdNoiseFloor")
        pu.check_scalars(s1.noise_variance[lo:hi], r.noise[50:, 1], rel=2e-3, what="noise variance")
        fl = s1.work_flush_offset[i]
        # sums of 100 dB values: bins at the noise level next to 40 carriers carry the fp32-FFT error of a 65536-point
        # transform (both GPU paths agree with each other to < 0.5 above); bins >= 10 dB over the median are tight
        d = np.abs(s1.flush_cum[fl] - r.flush_cum[0])
        assert d.max() < 5.0
        loud = r.flush_cum[0] > np.median(r.flush_cum[0]) + 1000.0
        assert loud.any() and d[loud].max() < 0.05
        pu.check_peaks(pu.peak_keys(s1.peaks(fl)), [p.key() for p in r.peaks[0]], r.flush_cum[0], r.thresholds[99, 2])
