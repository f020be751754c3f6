#!/usr/bin/env python
"""BASELINE config 4 (Avazu-shaped: 24 one-hot fields, 10M hashed features, k = 64, logistic) under
torchrun: row-sharded V (SFM_FLAG_SHARD_V, all-to-all row gather) vs the replicated data-parallel
mode (dense all-reduce of the 2.6 GB gradient).  Weak scaling: --batch rows per GPU per step.

    python -m torch.distributed.run --nproc-per-node N scripts/bench_c4.py --steps 20
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_FIELDS, N_SLOTS, K = 24, 10_000_000, 64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--rows", type=int, default=8_000_000, help="rows per GPU")
    ap.add_argument("--batch", type=int, default=500_000, help="mini-batch rows per GPU")
    ap.add_argument("--modes", default="sharded,sharded_partition,replicated")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    saved = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from sparkfm_b200 import Handle, synth
    from sparkfm_b200.dist import init_comm

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    card = synth.ctr_field_log2_cards(N_FIELDS)
    cdf, off = synth.zipf_tables(card)
    out = {"config": "C4 Avazu-shaped: 24 one-hot fields, 10M hashed features, k=64, logistic",
           "n_gpus": world, "rows_per_gpu": args.rows, "batch_per_gpu": args.batch}
    m = N_FIELDS
    b_train = 8 * m * (K + 2) + 4
    b_step = 12 * (1 + N_SLOTS * (K + 1))
    for mode in args.modes.split(","):
        sharded = mode.startswith("sharded")
        if sharded and world == 1:
            continue
        part = mode.endswith("partition")
        hd = Handle(N_SLOTS, K, task=1, reg=(0.0, 0.0, 1e-5), step_size=0.1,
                    mini_batch_fraction=args.batch / args.rows, sampler_seed=42, device=local,
                    shard_v=sharded, sampler_mode=1 if part else 0)
        if world > 1:
            init_comm(hd, device=f"cuda:{local}")
        hd.init_model(0.0, 0.01, 1)
        if world > 1 and not sharded:
            hd.comm_broadcast_model()
        hd.synth_ctr_dataset(args.rows, rank * args.rows, card, cdf, off, 20260104)
        warm = max(args.warmup, round(args.rows / args.batch)) if part else args.warmup
        h0 = hd.train(1, warm)    # PARTITION: the first epoch builds every batch's plan
        hd.stats_reset()
        hd.synchronize()
        if world > 1:
            dist.barrier()
        hd.timer_start()
        hist = hd.train(warm + 1, args.steps)
        ms = hd.timer_stop()
        rows = hd.stats()["train_rows"]
        t = torch.tensor([ms, float(rows)], dtype=torch.float64, device="cuda")
        if world > 1:
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = t.clone()
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            ms, rows = float(tm[0]), float(ts[1])
        # sharded: the dense step bytes are spread over the ranks (each updates 1/N of the model)
        step_bytes = b_step * (1 if sharded else world)
        gbs = (rows * b_train + args.steps * step_bytes) / (ms * 1e-3) / 1e9
        out[mode] = {"ms_per_step": ms / args.steps, "samples_per_s": rows / (ms * 1e-3),
                     "roofline_step_frac": gbs / (6554.2 * world),
                     "loss_first_last": [float(h0[0]), float(hist[-1])]}
        hd.close()
    if rank == 0:
        os.write(saved, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
