"""Parity helpers shared by the GPU tests (tolerances of SURVEY.md section 8(d))."""
import numpy as np

SIGNAL_REL = 1e-4       # (i) |dPSD|/PSD on signal bins (>= floor + 15 dB); BASELINE.json's criterion
ALLBIN_ABS = 1e-6       # (ii) |dPSD| <= 1e-6 * max_bin(PSD) on every bin
SCALAR_REL = 1e-4       # (iv) psdNoiseFloor / variance
TIE_EPS_DB = 1e-3       # (v) a decision may differ only if |value - threshold| < 1e-3 dB


def check_spectrum(spec_gpu, psd_gpu, spec_ref, psd_ref):
    """spec/psd: [blocks, N].  Returns a dict of measured maxima; asserts (i) and (ii)."""
    psd_ref64 = psd_ref.astype(np.float64)
    d = np.abs(psd_gpu.astype(np.float64) - psd_ref64)
    peak = psd_ref64.max(axis=1, keepdims=True)
    assert (d <= ALLBIN_ABS * peak).all(), f"(ii) violated: {(d / peak).max():.3e}"
    floor = np.median(psd_ref64, axis=1, keepdims=True)  # robust stand-in for the noise-floor mean
    sig = psd_ref64 >= floor * 10 ** 1.5
    out = {"allbin_over_peak": float((d / peak).max()), "n_signal_bins": int(sig.sum())}
    if sig.any():
        rel = d[sig] / psd_ref64[sig]
        assert rel.max() <= SIGNAL_REL, f"(i) violated on PSD: {rel.max():.3e}"
        ddb = np.abs(spec_gpu[sig].astype(np.float64) - spec_ref[sig].astype(np.float64))
        assert (ddb <= SIGNAL_REL * np.abs(spec_ref[sig].astype(np.float64))).all(), f"(i) violated on dB: {ddb.max():.3e}"
        out["signal_rel"] = float(rel.max())
        out["signal_ddb"] = float(ddb.max())
    return out


def check_scalars(a, b, rel=SCALAR_REL, what=""):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    assert err.max() <= rel, f"{what}: relative error {err.max():.3e} > {rel}"
    return float(err.max())


def check_keys(keys_gpu, taps_ref, thr_ref, keys_ref=None):
    """keys_gpu [blocks, L] vs oracle decisions taps_ref > thr_ref[:, None]; near-ties are listed and excused."""
    ref = (taps_ref > thr_ref[:, None]).astype(np.uint8) if keys_ref is None else keys_ref
    diff = np.argwhere(keys_gpu != ref)
    excused = []
    for b, l in diff:
        margin = abs(float(taps_ref[b, l]) - float(thr_ref[b]))
        assert margin < TIE_EPS_DB, f"key flip at block {b} listener {l} with margin {margin:.3e} dB"
        excused.append((int(b), int(l), margin))
    return excused


def peak_keys(peaks):
    return [(int(p["from"]), int(p["to"]), int(p["signal_bin"])) for p in peaks]


def check_peaks(gpu_keys, ref_keys, cum_ref, thr_ref):
    """identical (From, To, SignalBin) lists, except documented near-ties"""
    if gpu_keys == ref_keys:
        return []
    v = cum_ref.astype(np.float32) / np.float32(100)
    near = np.abs(v.astype(np.float64) - float(thr_ref)) < TIE_EPS_DB
    excused = []
    for k in sorted(set(gpu_keys) ^ set(ref_keys)):
        lo, hi = max(0, k[0] - 1), min(len(v) - 1, k[1] + 1)
        seg = v[lo:hi + 1]
        top2 = np.sort(seg)[-2:] if seg.size > 1 else np.array([0, 1])
        tie_max = abs(float(top2[-1]) - float(top2[0])) < TIE_EPS_DB
        assert near[lo:hi + 1].any() or tie_max, f"peak mismatch {k} not explained by a near-tie"
        excused.append(k)
    return excused
