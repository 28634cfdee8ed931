#!/usr/bin/env python
"""Device-resident throughput of the hot path for every BASELINE.json config shape (single GPU).

Not the contract bench (that is bench.py on configs[1]); this reports Msamples/s and the HBM-roofline fraction of
the dominant kernel stage for the other block sizes so that DESIGN.md can quote them.
Usage: python tools/bench_configs.py [--out profiles/r1_all_configs.json]
Multi-GPU (streams sharded by rank, no collective on the data path; barrier + max-over-ranks time only):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_configs.py --only cfg5
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdrainer_b200 import capi  # noqa: E402

# (name, N, fs, listeners, streams, blocks per stream[, {env overrides read at engine creation}])
CASES = [
    # name, N, fs, listeners, streams, blocks per stream
    ("cfg1 48 kS/s N=512 L=5", 512, 48000, 5, 4 * 148 * 12, 100),
    ("N=1024 L=25", 1024, 96000, 25, 2 * 148 * 12, 100),
    ("cfg2/cfg4 192 kS/s N=2048 L=50", 2048, 192000, 50, 148 * 12, 100),
    ("cfg4 literal N=2048 L=50, 64 streams x 2000 blocks", 2048, 192000, 50, 64, 2000),
    ("N=4096 L=100", 4096, 384000, 100, 148 * 4, 100),
    ("N=4096 L=100 (three-pass kernel)", 4096, 384000, 100, 148 * 4, 100, {"SDR_K1_MID4K": "0"}),
    # >= 2 GiB of blocks per batch like the other shapes; 444 streams = three segments per SM for the
    # fused N = 8192 kernel (k1_mid8k.cuh); below 49 segments the engine takes the block-parallel two-kernel path
    ("cfg3 768 kS/s N=8192 L=200 (k1_mid8k: TMA ring, 512 threads)", 8192, 768000, 200, 444, 100),
    ("cfg3 768 kS/s N=8192 L=200 (k1_mid8k, 1 stage)", 8192, 768000, 200, 444, 100, {"SDR_K1_MID8K_STAGES": "1"}),
    ("cfg3 768 kS/s N=8192 L=200 (two-kernel path)", 8192, 768000, 200, 444, 100, {"SDR_K1_MID8K": "0"}),
    ("cfg3 literal N=8192 L=200, 1 stream x 4000 blocks", 8192, 768000, 200, 1, 4000),
    ("cfg3-like N=8192 L=0 (no listeners)", 8192, 768000, 0, 444, 100),
    ("cfg3-like N=8192 L=50", 8192, 768000, 50, 444, 100),
    ("cfg3 768 kS/s N=8192 L=200, 64 streams", 8192, 768000, 200, 64, 100),
    ("cfg3 768 kS/s N=8192 L=200, 64 streams (r1 two-kernel large-block path)", 8192, 768000, 200, 64, 100, {"SDR_K1_MID8K": "0"}),
    ("N=16384 L=100 (64 x 256)", 16384, 1536000, 100, 222, 100),
    ("N=32768 L=100 (128 x 256)", 32768, 3072000, 100, 111, 100),
    ("cfg5 24.576 MS/s N=65536 peak scan, 64 streams", 65536, 24576000, 0, 64, 100),
    ("cfg5 24.576 MS/s N=65536 peak scan, 72 streams, lookahead 4", 65536, 24576000, 0, 72, 100, {"SDR_K1_WIDE_LOOKAHEAD": "4"}),
    ("cfg5 24.576 MS/s N=65536 peak scan, 72 streams, ring 8", 65536, 24576000, 0, 72, 100, {"SDR_K1_WIDE_RING": "8"}),
    ("cfg5 24.576 MS/s N=65536 peak scan, 72 streams, no discard", 65536, 24576000, 0, 72, 100, {"SDR_K1_WIDE_DISCARD": "0"}),
    ("cfg5 24.576 MS/s N=65536 peak scan, 72 streams (4 whole rounds of 18 teams)", 65536, 24576000, 0, 72, 100),
    ("cfg5 24.576 MS/s N=65536 peak scan, 64 streams (r1 two-kernel path)", 65536, 24576000, 0, 64, 100, {"SDR_K1_WIDE": "0"}),
    ("cfg5 24.576 MS/s N=65536 peak scan, 8 streams", 65536, 24576000, 0, 8, 100),
    ("cfg5 24.576 MS/s N=65536 peak scan, 8 streams (r1 two-kernel path)", 65536, 24576000, 0, 8, 100, {"SDR_K1_WIDE": "0"}),
    ("k2 cfg2 N=2048 L=50, K2 on a second stream, high priority", 2048, 192000, 50, 148 * 12, 100, {"SDR_K2_OVERLAP": "1"}),
    ("k2 cfg2 N=2048 L=50, K2 on a second stream, low priority", 2048, 192000, 50, 148 * 12, 100, {"SDR_K2_OVERLAP": "1", "SDR_K2_PRIO": "low"}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--only", default=None, help="comma-separated substring filters on the case name")
    ap.add_argument("--stream-scale", type=int, default=1, help="multiply the stream count of the selected cases")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    rows = []
    for case in CASES:
        name, n, fs, nl, n_streams, nb = case[:6]
        env = case[6] if len(case) > 6 else {}
        if args.only and not any(o in name for o in args.only.split(",")):
            continue
        saved = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        n_streams *= args.stream_scale
        g = torch.Generator(device=dev)
        g.manual_seed(n + 7919 * rank)
        iq = torch.randn((n_streams, nb * n * 2), generator=g, device=dev, dtype=torch.float32) * 1e-4
        rng = np.random.default_rng(n)
        bins = [np.sort(rng.choice(np.arange(80, n - 80), size=nl, replace=False)).astype(np.int32) for _ in range(n_streams)]
        # a keyed-looking carrier on every listener bin so that peaks/keys have work to do
        t = torch.arange(n, device=dev, dtype=torch.float32)
        for s in range(min(n_streams, 64)):
            for b in bins[s][:8]:
                ph = 2 * np.pi * float(b - n // 2) / n
                v = iq[s].view(nb, n, 2)
                v[:, :, 0] += 0.01 * torch.cos(ph * t)
                v[:, :, 1] += 0.01 * torch.sin(ph * t)
        eng = capi.Engine(n, max_streams=n_streams, max_listeners=max(nl, 1), max_blocks_per_batch=n_streams * nb,
                          max_peaks_per_flush=128, n_slots=2, device=local_rank, cuda_stream=stream.cuda_stream)
        for k, v in saved.items():  # the engine read its switches at creation
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        sids = [eng.open_stream(fs) for _ in range(n_streams)]
        per = nb * 2 * n * 4
        prepared = eng.prepare([dict(stream=sids[i], iq=iq.data_ptr() + i * per, n_blocks=nb, listener_bins=bins[i])
                                for i in range(n_streams)])
        for _ in range(3):
            tk = eng.submit_prepared(prepared, capi.NO_D2H)
            eng.collect_raw(tk)
            eng.release(tk)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k1, k2 = [], []
        e0.record(stream)
        pend = []
        for _ in range(args.steps):
            pend.append(eng.submit_prepared(prepared, capi.NO_D2H))
            if len(pend) == 2:
                r = eng.collect_raw(pend[0])
                k1.append(r.k1_ms)
                k2.append(r.k2_ms)
                eng.release(pend.pop(0))
        eng.fence()  # the last post kernels run on the engine's own stream
        e1.record(stream)
        for tk in pend:
            r = eng.collect_raw(tk)
            k1.append(r.k1_ms)
            k2.append(r.k2_ms)
            eng.release(tk)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:  # whole-job time = slowest rank
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        samples = n_streams * nb * n
        alg = (8 * n + 4 * n / 100.0 + 4 * nl + 16) * n_streams * nb
        k1_ms = float(np.mean(k1))
        rows.append({"config": name, "n_gpus": world, "block_size": n, "streams_per_gpu": n_streams, "streams": n_streams * world,
                     "listeners": nl, "batch_bytes": samples * 8, "msamples_per_s": world * samples / (ms * 1e-3) / 1e6,
                     "ms_per_step": ms,
                     "spectral_stage_ms": k1_ms, "spectral_stage_gbs": alg / (k1_ms * 1e-3) / 1e9,
                     "hbm_roofline_frac": alg / (k1_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "path_frac": alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "k2_ms": float(np.mean(k2)), "env": env,
                     "realtime_streams": world * samples / (ms * 1e-3) / fs})
        if rank == 0:
            print(json.dumps(rows[-1]))
        eng.close()
        del iq
        torch.cuda.empty_cache()
    if args.out and rank == 0:
        with open(args.out, "w") as f:
            json.dump({"peak_hbm_gbs": peaks["hbm_gbs"], "rows": rows}, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__" and "--goertzel" not in sys.argv:
    main()


def goertzel_stress():
    """config 3 'Goertzel bank stress': 200 listeners on 768 kS/s, N=8192 blocks through the IQ Goertzel bank (K3)
    next to the FFT-tap path that serves the same listeners."""
    import time
    n, nl, nb = 8192, 200, 16384
    dev = torch.device("cuda", 0)
    iq = torch.randn(nb * n * 2, device=dev, dtype=torch.float32) * 1e-4
    bins = np.sort(np.random.default_rng(3).choice(np.arange(80, n - 80), size=nl, replace=False)).astype(np.int32)
    bank = capi.GoertzelBank([700.0], 48000)
    bank.process_iq(iq.data_ptr(), n, bins, n_blocks=nb)  # warm-up (table upload, output buffers)
    torch.cuda.synchronize()
    dts = []
    for _ in range(3):
        t0 = time.perf_counter()
        bank.process_iq(iq.data_ptr(), n, bins, n_blocks=nb)
        torch.cuda.synchronize()
        dts.append(time.perf_counter() - t0)
    dt = min(dts)
    row = {"config": "cfg3 Goertzel bank stress (K3 IQ bank, 200 listeners, N=8192)", "blocks": nb,
           "msamples_per_s": nb * n / dt / 1e6, "listener_msamples_per_s": nb * n * nl / dt / 1e6,
           "gflops": 8.0 * nb * n * nl / dt / 1e9, "fp32_peak_frac": 8.0 * nb * n * nl / dt / 72e12,
           "note": "wall clock of one call, includes the D2H of the 200 dB values per block; "
           "compute-bound (200 flop/byte): reported against fp32 throughput, not HBM"}
    print(json.dumps(row))
    return row


if __name__ == "__main__" and "--goertzel" in sys.argv:
    goertzel_stress()
