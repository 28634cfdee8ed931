#!/usr/bin/env python
"""Per-bin-class error of the fp32 GPU path against the float64 oracle, per block size (SURVEY section 7 "hard parts",
section 8(d) tolerances (i)-(iv)).  Writes profiles/r2_error_table.{json,md}: the table DESIGN.md section 7 cites for the
tolerances the parity tests use.

Classes (by the ORACLE's PSD of the bin, per block):  signal = >= block noise floor + 15 dB;  floor = >= the block's
noise-floor mean;  deep = below it (fades / nulls next to strong carriers).
Usage (GPU box):  python tools/error_table.py [N ...]   (with sizes: only those rows are re-measured, the others are kept)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (checker only)
from sdrainer_b200 import capi, synth  # noqa: E402

ENV = {8192: {"SDR_K1_MID8K": "force"}, 65536: {"SDR_K1_WIDE": "force"}}


def one(n):
    fs = 48000 * n // 512
    nb = 104
    rng = np.random.default_rng(n)
    k = min(50, max(5, n // 64))
    tones = synth.make_tones(rng, k, n, 70, wpm_range=(18.0, 28.0))
    spec = synth.StreamSpec(sample_rate=fs, block_size=n, n_blocks=nb, seed=n + 3, tones=tones)
    iq = synth.generate(spec)
    bins = [t.bin for t in tones]
    old = {kk: os.environ.get(kk) for kk in ENV.get(n, {})}
    os.environ.update(ENV.get(n, {}))
    with capi.Engine(n, max_streams=1, max_listeners=max(len(bins), 1), max_blocks_per_batch=nb, max_peaks_per_flush=n // 2 + 1) as eng:
        s = eng.open_stream(fs)
        res = eng.collect(eng.submit([dict(stream=s, iq=iq, listener_bins=bins)], capi.WANT_SPECTRUM | capi.WANT_FLUSH_CUM))
        kernel = eng.last_kernel()
    for kk, v in old.items():
        if v is None:
            os.environ.pop(kk, None)
        else:
            os.environ[kk] = v
    ref = O.process_stream(iq, n, listener_bins=bins, sample_rate=fs, want_spectrum=True)
    pr = ref.psd.astype(np.float64)
    pg = res.psd.astype(np.float64)
    floor = ref.noise[:, 0].astype(np.float64)[:, None]
    rel = np.abs(pg - pr) / np.maximum(pr, 1e-300)
    ddb = np.abs(res.spectrum.astype(np.float64) - ref.spectrum.astype(np.float64))
    cls = {"signal": pr >= floor * 10 ** 1.5, "floor": (pr >= floor) & (pr < floor * 10 ** 1.5), "deep": pr < floor}
    row = {"block_size": n, "kernel": kernel, "blocks": nb, "tones": k,
           "allbin_abs_over_block_peak": float((np.abs(pg - pr) / pr.max(axis=1, keepdims=True)).max())}
    for name, m in cls.items():
        if m.any():
            row[name] = {"bins": int(m.sum()), "psd_rel_max": float(rel[m].max()), "psd_rel_p99": float(np.quantile(rel[m], 0.99)),
                         "db_abs_max": float(ddb[m].max())}
    row["psd_noise_floor_rel_max"] = float((np.abs(res.psd_noise_floor - ref.noise[:, 0]) / ref.noise[:, 0]).max())
    row["noise_variance_rel_max"] = float((np.abs(res.noise_variance - ref.noise[:, 1]) / ref.noise[:, 1]).max())
    row["thresholds_db_abs_max"] = float(np.abs(res.thresholds[:, :3] - ref.thresholds).max())
    row["flush_cum_abs_max"] = float(np.abs(res.flush_cum[0] - ref.flush_cum[0]).max())
    loud = ref.flush_cum[0] > np.median(ref.flush_cum[0]) + 1000.0
    row["flush_cum_abs_max_loud_bins"] = float(np.abs(res.flush_cum[0] - ref.flush_cum[0])[loud].max()) if loud.any() else None
    listen = ref.thresholds[:, 0] + ref.thresholds[:, 1]
    ref_keys = (ref.taps > listen[:, None]).astype(np.uint8)
    flips = np.argwhere(res.keys[:, :len(bins)] != ref_keys)
    row["key_flips"] = int(len(flips))
    row["key_flip_max_margin_db"] = float(max((abs(float(ref.taps[b, l]) - float(listen[b])) for b, l in flips), default=0.0))
    got = [(int(p["from"]), int(p["to"]), int(p["signal_bin"])) for p in res.peaks(0)]
    row["peak_list_identical"] = got == [p.key() for p in ref.peaks[0]]
    return row


def main():
    O.lib()
    capi.lib()
    sizes = [int(x) for x in sys.argv[1:]] or [512, 1024, 2048, 4096, 8192, 16384, 32768, 65536]
    rows = [one(n) for n in sizes]
    committed = os.path.join(ROOT, "profiles", "r2_error_table.json")
    if sys.argv[1:] and os.path.exists(committed):  # partial run: replace the measured rows, keep the rest
        with open(committed) as f:
            keep = {r["block_size"]: r for r in json.load(f)}
        keep.update({r["block_size"]: r for r in rows})
        rows = [keep[n] for n in sorted(keep)]
    # under gpurun only gpurun_out/ travels back: write there when it exists, else straight into profiles/
    out_dir = os.path.join(ROOT, "gpurun_out") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "r2_error_table.json"), "w") as f:
        json.dump(rows, f, indent=1)
    lines = ["# fp32 GPU path vs float64 oracle, per block size and bin class (tools/error_table.py, measured on B200)", "",
             "Synthetic stream per size: Gaussian noise sigma 1e-4 + keyed tones (amplitude log-uniform 1e-3..3e-2), 104 blocks, edge 70.",
             "signal = oracle PSD >= block noise floor + 15 dB; floor = >= noise floor; deep = below the noise floor.", "",
             "| N | kernel | signal: max rel PSD / max dB | floor: max rel / p99 rel / max dB | deep: max rel / p99 rel / max dB | all bins: abs / block peak | "
             "psdNoiseFloor rel | variance rel | thresholds dB | flush_cum abs (all / loud bins) | key flips (max margin dB) | peak list |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        def c(name):
            d = r.get(name)
            return "-" if not d else f"{d['psd_rel_max']:.1e} / {d['psd_rel_p99']:.1e} / {d['db_abs_max']:.1e}"
        sg = r.get("signal")
        sig = "-" if not sg else f"{sg['psd_rel_max']:.1e} / {sg['db_abs_max']:.1e}"
        loud = r["flush_cum_abs_max_loud_bins"]
        lines.append(f"| {r['block_size']} | {r['kernel']} | {sig} | {c('floor')} | {c('deep')} | {r['allbin_abs_over_block_peak']:.1e} | "
                     f"{r['psd_noise_floor_rel_max']:.1e} | {r['noise_variance_rel_max']:.1e} | {r['thresholds_db_abs_max']:.1e} | "
                     f"{r['flush_cum_abs_max']:.2e} / {'-' if loud is None else format(loud, '.1e')} | {r['key_flips']} ({r['key_flip_max_margin_db']:.1e}) | "
                     f"{'identical' if r['peak_list_identical'] else 'differs'} |")
    with open(os.path.join(out_dir, "r2_error_table.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
