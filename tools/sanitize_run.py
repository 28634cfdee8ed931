"""Small hot-path run for compute-sanitizer (memcheck / racecheck): a few streams through K1+K2 at several block
sizes, ragged batches so that state load/store, flush and the named-barrier groups are all exercised."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdrainer_b200 import capi

rng = np.random.default_rng(0)
for n in (int(a) for a in (sys.argv[1:] or ["512", "1024", "2048"])):
    with capi.Engine(n, max_streams=3, max_listeners=8, max_blocks_per_batch=400, max_peaks_per_flush=64) as eng:
        ss = [eng.open_stream(48000) for _ in range(3)]
        for nb in ((37, 120, 5), (90, 1, 101)):
            works = []
            for s, b in zip(ss, nb):
                iq = (rng.standard_normal(b * 2 * n) * 1e-3).astype(np.float32)
                works.append(dict(stream=s, iq=iq, listener_bins=[80, n // 2 + 3, n - 90]))
            r = eng.collect(eng.submit(works, capi.WANT_FLUSH_CUM))
            assert np.isfinite(r.psd_noise_floor).all()
    print("ok", n)
