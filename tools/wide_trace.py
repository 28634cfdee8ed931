#!/usr/bin/env python
"""Phase timing of k1_wide_kernel from a measurement build (-DSDR_K1W_TRACE) run with SDR_K1_WIDE_TRACE=<file>.

The file holds [32 CTAs][512 steps][32] clock64 stamps of the last launch (k1_wide.cuh, K1W_TR).  Prints the median
number of SM cycles between consecutive stamps in steady state (every stamp itself costs about 120 cycles).

  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DSDR_K1W_TRACE -Xcompiler -fPIC -shared \
       -o sdrainer_b200/libsdrgpu_trace.so sdrainer_b200/csrc/engine.cu sdrainer_b200/csrc/goertzel.cu
  SDRGPU_LIB=$PWD/sdrainer_b200/libsdrgpu_trace.so SDR_K1_WIDE_TRACE=gpurun_out/wide_trace.bin \
       python tools/bench_configs.py --steps 5 --only "(4 whole rounds"       # add SDR_K1_WIDE_TEAMS=9 for one CTA per SM
  python tools/wide_trace.py gpurun_out/wide_trace.bin
"""
import sys

import numpy as np

NAMES = {0: "produce start", 1: "FULL_A seen", 2: "produce barrier passed", 3: "produce done (OUT_RDY arrive)",
         4: "consume start", 5: "FULL_B seen", 6: "planes stored", 7: "consume barrier passed", 8: "consume end (tid 0)",
         9: "consume end (tid 255)", 10: "DMA: OUT_RDY seen", 11: "DMA: tile read out of A", 12: "DMA: store complete",
         13: "DMA: published", 14: "DMA: B_FREE seen", 15: "DMA: ready[] complete, B requested"}


def main(path, lookahead=3):
    t = np.fromfile(path, dtype=np.int64).reshape(32, 512, 32).astype(np.float64)
    t[t == 0] = np.nan
    lo, hi = 40, 360  # steady state of the first segments
    d = lookahead - 1

    def med(x):
        return np.nanmedian(x)

    # consume step i is in the same loop iteration as produce step i + D - 1
    P = lambda k: t[:, lo + d:hi + d, k]
    C = lambda k: t[:, lo:hi, k]
    print("iteration period (consume end -> consume end): %.0f cycles" % med(t[:, lo + 1:hi + 1, 8] - t[:, lo:hi, 8]))
    rows = [("wait FULL_A", P(1) - P(0)), ("produce: tile -> registers, barrier", P(2) - P(1)),
            ("produce: transform + twiddle + tile stores", P(3) - P(2)), ("produce end -> consume start", C(4) - P(3)),
            ("wait FULL_B", C(5) - C(4)), ("consume: transform + planes", C(6) - C(5)), ("consume barrier (tid 0 waits)", C(7) - C(6)),
            ("window sums, taps (tid 0)", C(8) - C(7)), ("  barrier -> share sums done", C(16) - C(7)), ("  float64 half-warp reduction", C(17) - C(16)),
            ("  partial store", C(18) - C(17)), ("  x_to, taps", C(19) - C(18)), ("  flush check .. end", C(8) - C(19)),
            ("produce done -> segment record read", C(20) - P(3)), ("segment record -> consume start", C(4) - C(20)), ("tid 255 ends after tid 0 by", C(9) - C(8)),
            ("next produce start after consume end", t[:, lo + d + 1:hi + d + 1, 0] - C(8)),
            ("DMA: OUT_RDY seen after produce done", P(10) - P(3)), ("DMA: tile read out", P(11) - P(10)),
            ("DMA: store complete", P(12) - P(11)), ("DMA: publish", P(13) - P(12)),
            ("DMA: B_FREE seen after FULL_B seen", C(14) - C(5)), ("DMA: B_FREE seen after publish", C(14) - P(13)),
            ("DMA: ready complete after B_FREE", t[:, lo + 1:hi + 1, 15] - C(14)),
            ("B requested -> FULL_B seen (next step)", t[:, lo + 1:hi + 1, 5] - t[:, lo + 1:hi + 1, 15]),
            ("B requested -> next consume start", t[:, lo + 1:hi + 1, 4] - t[:, lo + 1:hi + 1, 15])]
    for name, x in rows:
        print("  %-46s median %7.0f   p10 %7.0f   p90 %7.0f" % (name, med(x), np.nanpercentile(x, 10), np.nanpercentile(x, 90)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
