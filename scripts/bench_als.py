#!/usr/bin/env python
"""ALS sweep (sfm_als_sweep) at BASELINE config 1 / 2 shapes: device time per sweep, the one-off
build of the transposed input + level schedule, and the fp64 CPU restatement beside it."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(name, n_rows, n_slots, k, mean_nnz, values, sweeps=4, graph=True):
    os.environ["SFM_ALS_GRAPH"] = "1" if graph else "0"
    from oracle.capi import OracleFM
    from sparkfm_b200 import Handle, synth
    rng = np.random.default_rng(1)
    rp, idx, val = synth.ragged_rows(n_rows, n_slots, mean_nnz, seed=7, values=values, max_nnz=64)
    dval = np.ones(len(idx)) if values == "ones" else val.astype(np.float64)
    gen = OracleFM(n_slots, k)
    gen.set_model(0.1, rng.normal(0, 0.3, n_slots), rng.normal(0, 0.2, (n_slots, k)))
    y = (gen.predict(rp, idx, dval, fast=True, threads=8) + rng.normal(0, 0.1, n_rows)).astype(np.float32)
    w, v = np.zeros(n_slots, np.float32), rng.normal(0, 0.05, (n_slots, k)).astype(np.float32)
    reg = (0.0, 0.01, 0.05)
    hd = Handle(n_slots, k, task=0, reg=reg)
    hd.set_model(0.0, w, v)
    hd.load_dataset(rp, idx, None if values == "ones" else val, y)
    t0 = time.perf_counter()
    r_first = hd.als_sweep()
    t_first = time.perf_counter() - t0
    hd.stats_reset()
    t0 = time.perf_counter()
    hist = [hd.als_sweep() for _ in range(sweeps)]
    t_gpu = (time.perf_counter() - t0) / sweeps
    launches = hd.stats()["kernel_launches"] / sweeps
    orc = OracleFM(n_slots, k, task=0, reg=reg)
    orc.set_model(0.0, w, v)
    t0 = time.perf_counter()
    want, _ = orc.als_sweep(rp, idx, dval, y.astype(np.float64), store_f32=True)
    t_cpu = time.perf_counter() - t0
    hd.close()
    return {"config": name, "cuda_graph": graph, "rows": n_rows, "n_slots": n_slots, "k": k, "nnz": int(rp[-1]),
            "first_sweep_s_incl_build": t_first, "sweep_s": t_gpu, "launches_per_sweep": launches,
            "levels": round((launches - 6 - k) / (k + 1)), "rmse": [r_first] + hist,
            "cpu_oracle_sweep_s": t_cpu, "first_sweep_rmse_vs_oracle": [r_first, want],
            "entries_per_s": int(rp[-1]) * (k + 1) / t_gpu}


if __name__ == "__main__":
    out = [run("C1 shape: 100k rows x 10k features, nnz~20, k=8, all ones", 100_000, 10_000, 8, 20, "ones",
               graph=False),
           run("C1 shape: 100k rows x 10k features, nnz~20, k=8, all ones", 100_000, 10_000, 8, 20, "ones"),
           run("C2 shape / 4: 250k rows x 100k features, nnz~50, k=16, N(0,1) values", 250_000, 100_000,
               16, 50, "normal", sweeps=2)]
    print(json.dumps(out, indent=1))
