"""Generates the committed fixtures under tests/golden/.

1. cw_keystreams.json -- the reference's own recorded decoder fixtures: the nine key-state streams
   of /root/reference/cw/testdata/*.txt (one 0/1 per 512/48000 s tick) with the expected strings of
   cw/decode_test.go:184-192, run-length packed.  These PIN the oracle's decoder.
2. oracle_vectors.npz -- SELF-GENERATED vectors (the reference holds no fixture for spectrum, noise
   floor, thresholds or peak lists, and cannot be run here: no Go toolchain).  They freeze the
   oracle's output on seeded synthetic IQ so that an accidental change to the oracle is caught, and
   give the GPU tests a fixture that does not need the oracle at all.

Run from the repo root in the build container:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

EXPECTED = [
    ("db100fk_1", "i100fk"),
    ("db100fk_2", "i100fk cq db1drfk"),
    ("db100fk_3", "i100fk cq db1drfk db 100fk"),
    ("gb4wwa", "rq gb4wwa gb4wwa up"),
    ("ii3wwa", "kde ii3wwa ii3wwa pse k"),
    ("ly2px_1", "q cq"),
    ("ly2px_2", "q cq cqde"),
    ("ly2px_3", "q cq cqde ly2px ly2px"),
    ("ly2px_4", "q cq cqde ly2px ly2px cqcq cqde ly2px ly2px ly2gx ä"),
]


def rle(bits):
    out, cur, cnt = [], bits[0], 0
    for b in bits:
        if b == cur:
            cnt += 1
        else:
            out.append(cnt)
            cur, cnt = b, 1
    out.append(cnt)
    return {"first": int(bits[0]), "runs": out}


def keystreams():
    ref = "/root/reference/cw/testdata"
    items = []
    for name, expected in EXPECTED:
        bits = [int(x) for x in open(os.path.join(ref, name + ".txt")).read().split()]
        items.append({"name": name, "expected": expected, "n_ticks": len(bits), **rle(bits)})
    with open(os.path.join(HERE, "cw_keystreams.json"), "w", encoding="utf-8") as f:
        json.dump({"source": "cw/testdata/*.txt + cw/decode_test.go:184-192 (sample_rate 48000, block 512)",
                   "streams": items}, f, ensure_ascii=False, indent=0)


def oracle_vectors():
    from oracle import oracle as O
    from sdrainer_b200 import synth
    out = {}
    for tag, cfg, nblk in (("c1", 1, 230), ("c2", 2, 120)):
        spec = synth.config(cfg)
        spec.n_blocks = nblk
        iq = synth.generate(spec)
        bins = [t.bin for t in spec.tones]
        r = O.process_stream(iq, spec.block_size, edge_width=70, peak_threshold=15.0, listener_bins=bins,
                             sample_rate=spec.sample_rate, want_spectrum=True)
        out[tag + "_bins"] = np.asarray(bins, np.int32)
        out[tag + "_noise"] = r.noise
        out[tag + "_thresholds"] = r.thresholds
        out[tag + "_taps"] = r.taps
        out[tag + "_flush_cum"] = r.flush_cum
        out[tag + "_spectrum0"] = r.spectrum[:2]
        out[tag + "_psd0"] = r.psd[:2]
        pk = np.asarray([[p.from_, p.to, p.signal_bin] for p in r.peaks[0]], np.int32).reshape(-1, 3)
        out[tag + "_peaks0"] = pk
        out[tag + "_iq_sha"] = np.frombuffer(__import__("hashlib").sha256(iq.tobytes()).digest(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)


if __name__ == "__main__":
    if os.path.isdir("/root/reference/cw/testdata"):
        keystreams()
    oracle_vectors()
    print("golden fixtures written")
