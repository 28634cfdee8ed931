import numpy as np

from sdrainer_b200 import synth


def test_generator_is_seeded_and_shaped():
    spec = synth.config(1, seconds=0.5)
    a = synth.generate(spec)
    b = synth.generate(synth.config(1, seconds=0.5))
    assert a.dtype == np.float32 and a.size == spec.n_blocks * 2 * 512
    assert np.array_equal(a, b)
    assert len(spec.tones) == 5
    lo, hi = 75, 512 - 75
    assert all(lo <= t.bin < hi for t in spec.tones)
    # never an all-zero block
    assert (np.abs(a.reshape(spec.n_blocks, -1)).max(axis=1) > 0).all()


def test_keying_timing():
    # 20 WPM: dit = 60 ms = 2880 samples at 48 kS/s; "e" = one dit then 7 units of gap
    env = synth.keying("e", 20.0, 48000, 8 * 2880, 0.0)
    assert env[:2880].all() and not env[2880:8 * 2880].any()


def test_config_shapes():
    for cfg, n, k in ((1, 512, 5), (2, 2048, 50), (3, 8192, 200)):
        s = synth.config(cfg, seconds=0.05)
        assert s.block_size == n and len(s.tones) == k
