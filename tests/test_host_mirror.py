"""The C++ host mirror (sdrainer_b200/host) against the reference's own unit tests -- no GPU needed.

Mirrors cw/decode_test.go:177-213 and :35-56, dsp/dsp_test.go:13-23, dsp/fft_test.go:31-50,
rx/peaks_test.go:12-143, rx/listener_test.go:10-67."""
import json
import os

import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def host():
    from sdrainer_b200 import _build, hostapi
    _build.build_host()
    hostapi.lib()
    return hostapi


def _streams():
    with open(os.path.join(GOLDEN, "cw_keystreams.json"), encoding="utf-8") as f:
        data = json.load(f)
    for s in data["streams"]:
        bits, cur = [], s["first"]
        for r in s["runs"]:
            bits.extend([cur] * r)
            cur ^= 1
        yield s["name"], bits, s["expected"]


def test_decoder_recorded_streams(host):
    """cw/decode_test.go:177-213 through the product's decoder: one instance, Reset() between files"""
    d = host.Decoder(48000, 512)
    for name, bits, expected in _streams():
        d.reset()
        d.feed(bits)
        d.stop()
        assert d.text == expected, name


def test_decoder_code_table(host):
    """cw/decode_test.go:35-56: every table entry survives encode -> decode at 20 WPM (5 ticks per dit)"""
    from sdrainer_b200 import synth
    d = host.Decoder(48000, 512)
    for ch, code in synth.MORSE.items():
        keys = []
        for i, s in enumerate(code):
            if i:
                keys += [0] * 5
            keys += [1] * (5 if s == "." else 15)
        keys += [0] * (3 * 7 * 5)
        d.reset()
        d.feed(keys)
        d.stop()
        assert d.text == ch


def test_decoder_speed_range(host):
    """cw/decode_test.go:137-175 in spirit: the adaptive decoder follows 10..40 WPM without a preset"""
    from sdrainer_b200 import synth
    tick = 512 / 48000
    for wpm in (10, 15, 20, 25, 30, 40):
        dit = max(2, int((1.2 / wpm) / tick))
        units = synth.morse_units("paris paris paris paris")
        keys = [int(u) for u in units for _ in range(dit)]
        d = host.Decoder(48000, 512)
        d.feed(keys + [0] * 100)
        d.stop()
        assert d.text.strip().endswith("paris paris"), (wpm, d.text)


def test_bool_debouncer(host):
    L = host.lib()
    h = L.sdrh_debouncer_new(3)
    try:
        seq = [(1, 0), (1, 0), (1, 1), (1, 1), (0, 1), (0, 1), (0, 0)]
        for raw, exp in seq:
            assert L.sdrh_debouncer_debounce(h, raw) == exp
    finally:
        L.sdrh_debouncer_free(h)


def test_frequency_mapping(host):
    L = host.lib()
    fs, n, fc = 48000, 512, 7020000
    for b, center in ((0, fc - fs // 2), (256, fc)):
        assert L.sdrh_frequency_to_bin(fs, n, fc, center) == b
        assert L.sdrh_bin_to_frequency(fs, n, fc, b, 0.0) == center
    # PeakCenterCorrection (dsp/fft.go:292-309): symmetric neighbours -> 0; edges -> 0; skewed -> shifted
    assert L.sdrh_peak_signal_frequency(fs, n, fc, 300, 10.0, 20.0, 10.0) == fc - 24000 + int(300 * 93.75)
    assert L.sdrh_peak_signal_frequency(fs, n, fc, 0, 1.0, 20.0, 10.0) == fc - 24000
    corr = (15.0 - 10.0) / (2 * (2 * 20.0 - 10.0 - 15.0))
    assert L.sdrh_peak_signal_frequency(fs, n, fc, 300, 10.0, 20.0, 15.0) == fc - 24000 + int(300 * 93.75 + 93.75 * corr)


def test_peaks_table_put_into_empty_table(host):
    L = host.lib()
    t = L.sdrh_peaks_new(512)
    try:
        p = L.sdrh_peaks_make(t, 234, 235)
        L.sdrh_peaks_put(t, p, 0)
        assert L.sdrh_peaks_bin(t, 234) == p and L.sdrh_peaks_bin(t, 235) == p
        assert L.sdrh_peaks_bin_state(t, 234) == 1  # peakNew
        assert L.sdrh_peaks_bin(t, 233) == -1 and L.sdrh_peaks_bin(t, 236) == -1
    finally:
        L.sdrh_peaks_free(t)


def test_peaks_table_put_overlaps(host):
    """rx/peaks_test.go:30-78"""
    L = host.lib()
    t = L.sdrh_peaks_new(12)
    try:
        p1 = L.sdrh_peaks_make(t, 3, 4)
        p2 = L.sdrh_peaks_make(t, 5, 6)
        p3 = L.sdrh_peaks_make(t, 8, 8)
        p4 = L.sdrh_peaks_make(t, 10, 10)
        for p in (p1, p2, p3, p4):
            L.sdrh_peaks_put(t, p, 0)
        L.sdrh_peaks_activate(t, p3)
        L.sdrh_peaks_activate(t, p4)
        L.sdrh_peaks_deactivate(t, p4)  # inactive
        n1 = L.sdrh_peaks_make(t, 1, 2)
        n2 = L.sdrh_peaks_make(t, 4, 5)
        n3 = L.sdrh_peaks_make(t, 7, 8)
        n4 = L.sdrh_peaks_make(t, 10, 11)
        for p in (n1, n2, n3, n4):
            L.sdrh_peaks_put(t, p, 0)
        got = [L.sdrh_peaks_bin(t, i) for i in range(12)]
        assert got == [-1, n1, n1, -1, n2, n2, -1, -1, p3, -1, p4, -1]
    finally:
        L.sdrh_peaks_free(t)


def test_peaks_table_cleanup_and_find_next(host):
    """rx/peaks_test.go:80-143"""
    L = host.lib()
    t = L.sdrh_peaks_new(512)
    try:
        p = L.sdrh_peaks_make(t, 234, 235)
        L.sdrh_peaks_put(t, p, 0)
        L.sdrh_peaks_cleanup(t)
        assert L.sdrh_peaks_bin(t, 234) == p
        L.sdrh_peaks_clock_add(t, 121.0)
        L.sdrh_peaks_cleanup(t)
        assert L.sdrh_peaks_bin(t, 234) == -1 and L.sdrh_peaks_bin(t, 235) == -1
        # an active peak survives the timeout, an inactive one does not
        q = L.sdrh_peaks_make(t, 100, 101)
        L.sdrh_peaks_put(t, q, 0)
        L.sdrh_peaks_activate(t, q)
        L.sdrh_peaks_clock_add(t, 121.0)
        L.sdrh_peaks_cleanup(t)
        assert L.sdrh_peaks_bin(t, 100) == q
        L.sdrh_peaks_deactivate(t, q)
        L.sdrh_peaks_cleanup(t)
        assert L.sdrh_peaks_bin(t, 100) == -1
        # FindNext only returns new peaks
        r = L.sdrh_peaks_make(t, 300, 301)
        L.sdrh_peaks_put(t, r, 0)
        assert L.sdrh_peaks_find_next(t) == r
        L.sdrh_peaks_activate(t, r)
        assert L.sdrh_peaks_find_next(t) == -1
        L.sdrh_peaks_deactivate(t, r)
        assert L.sdrh_peaks_find_next(t) == -1
    finally:
        L.sdrh_peaks_free(t)


def test_listener_pool(host):
    """rx/listener_test.go:10-67"""
    L = host.lib()
    pool = L.sdrh_pool_new(3, b"test")
    try:
        made = []
        for i in range(1, 4):
            idx = L.sdrh_pool_bind_next(pool)
            assert idx >= 0
            made.append(idx)
            assert L.sdrh_pool_made_id(pool, idx) == f"test{i}".encode()
            assert L.sdrh_pool_active_id(pool, i - 1) == f"test{i}".encode()
        assert L.sdrh_pool_bind_next(pool) == -1
        L.sdrh_pool_release(pool, made[1])
        assert L.sdrh_pool_len(pool) == 2
        assert L.sdrh_pool_active_id(pool, 0) == b"test1" and L.sdrh_pool_active_id(pool, 1) == b"test3"
        L.sdrh_pool_release(pool, made[0])
        assert L.sdrh_pool_len(pool) == 1 and L.sdrh_pool_active_id(pool, 0) == b"test3"
        L.sdrh_pool_release(pool, made[2])
        assert L.sdrh_pool_len(pool) == 0
        idx = L.sdrh_pool_bind_next(pool)
        assert L.sdrh_pool_made_id(pool, idx) == b"test3"  # the id released last is reused first
    finally:
        L.sdrh_pool_free(pool)
