"""Randomised parity sweeps: dsp.FindPeaks / dsp.FindNoiseFloor drop-ins on adversarial vectors (ties, plateaus,
NaN, +-Inf, thresholds equal to values) and rx.Receiver with the seeded random FindNext probe."""
import ctypes as C

import numpy as np
import pytest

from sdrainer_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [512, 2048])
def test_find_peaks_random_vectors(capi, oracle, n):
    fm = oracle.freqmap(48000, n, 7000000)
    with capi.Engine(n) as eng:
        for seed in range(40):
            rng = np.random.default_rng(1000 * n + seed)
            # quantised levels make exact ties with the threshold and plateaus common
            cum = (rng.integers(0, 12, n) * 250.0).astype(np.float32)
            if seed % 3 == 0:
                cum[rng.integers(0, n, 6)] = np.nan
            if seed % 4 == 0:
                cum[rng.integers(0, n, 4)] = np.inf
                cum[rng.integers(0, n, 4)] = -np.inf
            if seed % 5 == 0:
                cum[:3] = 5000.0   # run open at bin 0
                cum[-3:] = 5000.0  # run open at the end
            thr = float(rng.choice([2.5, 5.0, 7.5, 12.5, 15.0, 27.5]))  # cum/100 takes the values 0, 2.5, 5, ...
            got, cnt = eng.find_peaks(cum, thr, max_peaks=n)
            ref = oracle.find_peaks(cum, thr, fm)
            assert cnt == len(ref), seed
            assert [p.key() for p in got] == [p.key() for p in ref], seed
            a = np.array([p.signal_value for p in got], np.float32)
            b = np.array([p.signal_value for p in ref], np.float32)
            assert np.array_equal(a, b, equal_nan=True), seed


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096])
def test_find_noise_floor_random_vectors(capi, oracle, n):
    with capi.Engine(n) as eng:
        for seed in range(24):
            rng = np.random.default_rng(77 * n + seed)
            e_max = (n - 10 * 16) // 2
            e = int(rng.integers(0, min(e_max, n // 4)))
            psd = (rng.exponential(1.0, n) * 4e-5).astype(np.float32)
            for _ in range(int(rng.integers(0, 6))):  # a few strong carriers
                psd[rng.integers(0, n)] = np.float32(rng.uniform(1.0, 4000.0))
            mn, var = eng.find_noise_floor(psd, e)
            omn, ovar = oracle.find_noise_floor(psd, e)
            assert abs(float(mn) - float(omn)) <= 2e-7 * float(omn), (seed, e)
            assert abs(var - ovar) <= 1e-9 * ovar, (seed, e)


def test_receiver_with_seeded_random_find_next(capi, oracle):
    """rx/peaks.go:183-207: the random probe, with the PRNG shared by oracle and host mirror"""
    from sdrainer_b200 import _build, hostapi
    _build.build_host()
    spec = synth.config(1, seconds=9.0)
    iq = synth.generate(spec)
    n = spec.block_size
    L = oracle.lib()
    cfg = oracle.ReceiverConfig()
    L.orc_receiver_config_default(C.byref(cfg), spec.sample_rate, n)
    cfg.deterministic_find_next = 0
    cfg.rng_seed = 12345
    orx = L.orc_receiver_new(C.byref(cfg))
    fp = C.POINTER(C.c_float)
    for b in range(spec.n_blocks):
        blk = np.ascontiguousarray(iq[b * 2 * n:(b + 1) * 2 * n])
        assert L.orc_receiver_process_block(orx, blk.ctypes.data_as(fp)) == 0
    ref = [(L.orc_receiver_listener_bin(orx, i), L.orc_receiver_listener_attach_block(orx, i),
            L.orc_receiver_listener_text(orx, i).decode("utf-8")) for i in range(L.orc_receiver_listener_count(orx))]
    L.orc_receiver_free(orx)
    with capi.Engine(n, max_listeners=32, max_blocks_per_batch=100, max_peaks_per_flush=n // 2 + 1) as eng:
        rx = hostapi.Receiver(eng, strain=True)
        rx.set_find_next(deterministic=False, seed=12345)
        rx.start(spec.sample_rate, n)
        for b in range(spec.n_blocks):
            rx.iq_data(spec.sample_rate, iq[b * 2 * n:(b + 1) * 2 * n])
            if b % 10 == 9:
                rx.process()
        rx.process()
        got = rx.listeners()
        rx.close()
    assert len(got) == len(ref) >= 4
    # the random probe makes the attach ORDER differ from bin order; both sides must agree on it
    assert [g["attach_block"] for g in got] == [r[1] for r in ref]
    assert [g["text"] for g in got] == [r[2] for r in ref]
    assert len({r[0] for r in ref}) == len(ref)
