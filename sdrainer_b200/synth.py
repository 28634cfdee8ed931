"""Seeded synthetic IQ for the BASELINE.json configs.

Complex white Gaussian noise (sigma = 1e-4 per component, never an all-zero block) plus K keyed
complex tones a*key(t)*exp(2*pi*i*(f*t + phi)), f on or near bin centres inside [edge+5, N-edge-5),
amplitudes log-uniform in [1e-3, 3e-2], hard on/off ITU morse keying of a fixed text at a given WPM.
Output: float32 interleaved I,Q -- the layout of tci/tci.go:264 and kiwi/client.go:298-308.
Host-side plumbing only (numpy); used by tests, bench.py and the oracle alike.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

MORSE = {
    "a": ".-", "b": "-...", "c": "-.-.", "d": "-..", "e": ".", "f": "..-.", "g": "--.", "h": "....", "i": "..",
    "j": ".---", "k": "-.-", "l": ".-..", "m": "--", "n": "-.", "o": "---", "p": ".--.", "q": "--.-", "r": ".-.",
    "s": "...", "t": "-", "u": "..-", "v": "...-", "w": ".--", "x": "-..-", "y": "-.--", "z": "--..",
    "0": "-----", "1": ".----", "2": "..---", "3": "...--", "4": "....-", "5": ".....", "6": "-....", "7": "--...",
    "8": "---..", "9": "----.", "/": "-..-.", "?": "..--..", "=": "-...-",
}

DEFAULT_TEXT = "cq de dl1abc dl1abc k"


def morse_units(text: str) -> np.ndarray:
    """0/1 per dit unit: dit 1, dah 3, symbol gap 1, char gap 3, word gap 7, trailing word gap."""
    out = []
    pending = 0
    for ch in text.lower():
        if ch == " ":
            if pending:
                pending = 7
            continue
        code = MORSE.get(ch)
        if code is None:
            continue
        out.extend([0] * pending)
        for i, s in enumerate(code):
            if i:
                out.append(0)
            out.extend([1] * (1 if s == "." else 3))
        pending = 3
    out.extend([0] * 7)
    return np.asarray(out, dtype=np.uint8)


def keying(text: str, wpm: float, fs: int, n_samples: int, start_s: float = 0.0) -> np.ndarray:
    """Sample-rate key envelope (uint8 0/1) of `text` repeated to fill n_samples, starting at start_s."""
    units = morse_units(text)
    dit_s = 1.2 / wpm
    t = (np.arange(n_samples, dtype=np.float64) / fs) - start_s
    idx = np.floor(t / dit_s).astype(np.int64)
    env = np.zeros(n_samples, dtype=np.uint8)
    valid = idx >= 0
    env[valid] = units[idx[valid] % units.size]
    return env


@dataclass
class Tone:
    bin: int               # fftshifted bin index (Listener.SignalBin convention)
    amplitude: float
    wpm: float = 20.0
    text: str = DEFAULT_TEXT
    phase: float = 0.0
    start_s: float = 0.0
    bin_offset: float = 0.0  # fraction of a bin off centre
    keyed: bool = True


@dataclass
class StreamSpec:
    sample_rate: int
    block_size: int
    n_blocks: int
    seed: int
    tones: list = field(default_factory=list)
    noise_sigma: float = 1e-4
    edge_width: int = 70


def make_tones(rng: np.random.Generator, n_tones: int, block_size: int, edge_width: int, wpm_range=(20.0, 20.0),
               amp_range=(1e-3, 3e-2), min_spacing: int = 6, keyed: bool = True, off_center: float = 0.0):
    lo, hi = edge_width + 5, block_size - edge_width - 5
    bins = []
    guard = 0
    while len(bins) < n_tones:
        b = int(rng.integers(lo, hi))
        guard += 1
        if guard > 100000:
            raise ValueError("cannot place tones with the requested spacing")
        if abs(b - block_size // 2) < 2:  # keep clear of DC
            continue
        if all(abs(b - o) >= min_spacing for o in bins):
            bins.append(b)
    tones = []
    for b in sorted(bins):
        amp = float(np.exp(rng.uniform(np.log(amp_range[0]), np.log(amp_range[1]))))
        wpm = float(rng.uniform(*wpm_range)) if wpm_range[1] > wpm_range[0] else float(wpm_range[0])
        tones.append(Tone(bin=b, amplitude=amp, wpm=wpm, phase=float(rng.uniform(0, 2 * np.pi)),
                          start_s=float(rng.uniform(0.0, 1.0)), keyed=keyed,
                          bin_offset=float(rng.uniform(-off_center, off_center)) if off_center else 0.0))
    return tones


def generate(spec: StreamSpec) -> np.ndarray:
    """Returns float32 [n_blocks * 2N] interleaved I,Q."""
    n, fs = spec.block_size, spec.sample_rate
    total = spec.n_blocks * n
    rng = np.random.default_rng(spec.seed)
    z = np.empty((total, 2), dtype=np.float32)
    z[:] = rng.standard_normal((total, 2), dtype=np.float32) * np.float32(spec.noise_sigma)
    sig = np.zeros(total, dtype=np.complex128)
    nidx = np.arange(n, dtype=np.float64)
    steady = np.zeros(n, dtype=np.complex128)  # unkeyed on-bin carriers repeat identically in every block
    for tn in spec.tones:
        k = tn.bin - n // 2  # baseband bin
        if tn.bin_offset == 0.0 and not tn.keyed:
            steady += tn.amplitude * np.exp(2j * np.pi * (k * nidx / n) + 1j * tn.phase)
            continue
        if tn.bin_offset == 0.0:
            one = np.exp(2j * np.pi * (k * nidx / n) + 1j * tn.phase)  # identical in every block
            carrier = np.tile(one, spec.n_blocks)
        else:
            f = (k + tn.bin_offset) / n  # cycles per sample
            ph = (f * np.arange(total, dtype=np.float64)) % 1.0
            carrier = np.exp(2j * np.pi * ph + 1j * tn.phase)
        if tn.keyed:
            carrier = carrier * keying(tn.text, tn.wpm, fs, total, tn.start_s)
        sig += tn.amplitude * carrier
    if steady.any():
        sig += np.tile(steady, spec.n_blocks)
    z[:, 0] += sig.real.astype(np.float32)
    z[:, 1] += sig.imag.astype(np.float32)
    return z.reshape(-1)


# ---- the BASELINE.json configs (sizes per SURVEY.md section 8) ------------------------------------
def config(cfg: int, seconds: float | None = None, stream: int = 0) -> StreamSpec:
    if cfg == 1:
        fs, n, k, wpm, dur = 48000, 512, 5, (20.0, 20.0), 30.0
    elif cfg in (2, 4):
        fs, n, k, wpm, dur = 192000, 2048, 50, (15.0, 30.0), 20.0
    elif cfg == 3:
        fs, n, k, wpm, dur = 768000, 8192, 200, (15.0, 30.0), 10.0
    elif cfg == 5:
        fs, n, k, wpm, dur = 24576000, 65536, 500, (20.0, 20.0), 0.5
    else:
        raise ValueError("cfg must be 1..5")
    if seconds is not None:
        dur = seconds
    n_blocks = max(1, int(dur * fs / n))
    seed = cfg * 1000 + stream
    rng = np.random.default_rng(seed + 7)
    tones = make_tones(rng, k, n, 70, wpm_range=wpm, keyed=(cfg != 5))
    return StreamSpec(sample_rate=fs, block_size=n, n_blocks=n_blocks, seed=seed, tones=tones)
