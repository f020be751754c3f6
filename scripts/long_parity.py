#!/usr/bin/env python
"""Loss-curve parity over a LONG run: Criteo-shaped rows (39 one-hot fields, Zipf ids), k = 16,
logistic loss, N iterations of Bernoulli mini-batches on the GPU vs the fp64 CPU oracle fed the same
row lists.  Writes gpurun_out/long_parity.json (per-iteration losses and relative error)."""
import json
import os
import sys

os.environ.setdefault("OMP_WAIT_POLICY", "passive")
import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import capi as ocapi  # noqa: E402
from oracle.capi import OracleFM  # noqa: E402
from sparkfm_b200 import Handle, synth  # noqa: E402

N_SLOTS, K, N_ROWS, ITERS = 100_000, 16, 400_000, int(sys.argv[1]) if len(sys.argv) > 1 else 150
FRAC, STEP, REG = 0.1, 0.5, (0.0, 0.0, 1e-5)
rp, idx, _, label = synth.ctr_csr(0, N_ROWS, 39, N_SLOTS, 20260103)
ones = np.ones(len(idx))
out = {}
for mode, name in ((0, "bernoulli"), (1, "partition")):
    hd = Handle(N_SLOTS, K, task=1, reg=REG, step_size=STEP, mini_batch_fraction=FRAC,
                sampler_seed=42, sampler_mode=mode)
    hd.init_model(0.0, 0.01, 1)
    w0, w, v = hd.get_model()
    hd.load_dataset(rp, idx, None, label)
    orc = OracleFM(N_SLOTS, K, task=1, reg=tuple(float(np.float32(r)) for r in REG))
    orc.set_model(w0, w, v)
    f32, s32 = float(np.float32(FRAC)), float(np.float32(STEP))
    P = ocapi.n_parts_for(FRAC)
    gl = hd.train(1, ITERS)
    ol, worst = [], 0.0
    for it in range(1, ITERS + 1):
        ids = ocapi.sample_rows(42, it, f32, 0, N_ROWS) if mode == 0 else \
            ocapi.partition_rows(42, P, (it - 1) % P, 0, N_ROWS)
        lo = orc.train_step(rp, idx, ones, label, ids, it, s32, threads=8) / len(ids)
        ol.append(lo)
        worst = max(worst, abs(gl[it - 1] - lo) / lo)
    gm = hd.get_model()
    out[name] = {"iters": ITERS, "max_rel_loss_err": worst, "loss_gpu_first_last": [gl[0], gl[-1]],
                 "loss_oracle_first_last": [ol[0], ol[-1]],
                 "max_rel_v_err": float(np.max(np.abs(gm[2] - orc.v)) / np.abs(orc.v).max()),
                 "rel_err_every_10": [abs(gl[i] - ol[i]) / ol[i] for i in range(0, ITERS, 10)]}
    hd.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "long_parity.json"), "w"), indent=1)
print(json.dumps({k: {a: b for a, b in v.items() if a != "rel_err_every_10"} for k, v in out.items()}, indent=1))
