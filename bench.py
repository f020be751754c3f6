#!/usr/bin/env python
"""bench.py -- FM SGD training throughput on the Criteo-shaped config (BASELINE.json configs[2]).

    python bench.py --gpus N --steps K --warmup W            # ours (libsparkfm_b200.so)
    python bench.py --impl reference --gpus N ...            # CPU arm (see below)

One "step" = one SGD iteration (FMLearn.learn) over one mini-batch: forward + loss, deterministic
reduce-by-feature, (all-reduce,) L2-regularised update.  Workload: synthetic CTR rows, 39 one-hot
fields, Zipf ids hashed into 1,000,000 features, k = 16, logistic loss, random-initialised model;
the data set is generated ON the device and stays resident in HBM, each iteration samples a
Bernoulli mini-batch of ~`--batch` rows per GPU (weak scaling: per-GPU batch fixed).

`value`   samples/s, whole job, inputs resident in HBM, device-timed (CUDA events on the library's
          stream, max over ranks).
`e2e`     the same metric through the C-ABI call a host makes for each mini-batch
          (sfm_train_step_csr): pinned HOST CSR buffers in, mean loss out, wall-clock around the
          calls (each call returns after its D2H read), H2D/D2H inside the timed region.
`roofline` dominant kernel's algorithmic bytes / its CUDA-event duration vs the measured HBM peak.
`cpu_baseline` the fp64 CPU oracle (OpenMP; "port" -- the reference is Scala/Spark and cannot run
          in this image) on a bounded sample of the same workload.

--impl reference: the reference's own implementation cannot be executed here (no JVM); per the
task contract this arm times the CPU oracle port of the path on all host threads, same config,
metric and unit, on a bounded sample per step.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FIELDS, N_SLOTS, K, ZIPF_S = 39, 1_000_000, 16, 1.1
DATA_SEED, INIT_SEED, SAMPLER_SEED = 20260103, 1, 42
STEP_SIZE, REG = 0.1, (0.0, 0.0, 1e-5)
METRIC, UNIT = "FM SGD train samples/sec", "samples/s"   # BASELINE.json "metric"
WORKLOAD = ("C3 Criteo-shaped CTR: 39 one-hot fields, 1M hashed features, Zipf 1.1, k=16, "
            "logistic loss, 45M rows resident in HBM")


def b_train(m, k):   # algorithmic bytes per sample
    return 8 * m * (k + 2) + 4


def b_step(n_slots, k):
    return 12 * (1 + n_slots * (k + 1))


def b_forward(m, k):  # forward kernel's share: idx+val, label, gather w_i + V_i
    return 8 * m + 4 + 4 * m * (k + 1)


def b_reduce(m, k, n_slots, rows):  # reduce-by-feature + update share, per launch
    return rows * 4 * m * (k + 1) + b_step(n_slots, k)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md).  The poller is started before the
    warm-up (nvidia-smi needs ~0.2 s to come up); only samples whose timestamp falls inside the
    timed region are kept -- if the region is shorter than one poll, the samples taken under load
    since the warm-up are used and `window` says so."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None
        self.nvml = None          # (module, handle) when NVML can be polled in-process
        self.samples = []         # (time, sm MHz, max MHz, reasons bitmask)
        self._stop = False

    def _nvml_open(self):
        import pynvml
        pynvml.nvmlInit()
        h = None
        try:
            import torch
            u = str(torch.cuda.get_device_properties(self.gpu).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + u if not u.startswith("GPU-") else u).encode())
        except Exception:
            h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)   # probe
        return pynvml, h

    def _poll(self):
        nv, h = self.nvml
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = None
        while not self._stop:
            try:
                if mx is None:
                    mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                self.samples.append((time.time(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                     float(mx), int(get_reasons(h))))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        # in-process NVML polling (every 5 ms; the bench thread sits in a ctypes call, GIL released);
        # nvidia-smi -lms as the fallback: its start-up can outlast a short timed region
        try:
            self.nvml = self._nvml_open()
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def _stop_nvml(self):
        self._stop = True
        self.t.join(timeout=1)
        inside = [x for x in self.samples if self.t0 is not None and self.t0 <= x[0] <= self.t1]
        window = "timed region"
        if not inside:
            inside = self.samples[-8:]
            window = "warm-up + timed region (timed region shorter than one poll)"
        bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20,
                "hw_thermal_slowdown": 0x40}
        reasons = [n for n, b in bits.items() if any(x[3] & b for x in inside)]
        sm = [x[1] for x in inside]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(x[2] for x in inside) if inside else None, "reasons": reasons,
                "samples": len(sm), "window": window, "source": "NVML polled in-process every 5 ms"}

    def stop(self):
        if self.nvml:
            return self._stop_nvml()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        ok = [(t, r) for t, r in self.rows if len(r) >= 10 and r[2].replace(".", "").isdigit()]
        inside = [r for t, r in ok if self.t0 is not None and self.t0 <= t <= self.t1 + 0.05]
        window = "timed region"
        if not inside:
            inside = [r for t, r in ok if self.t0 is None or t <= self.t1 + 0.05][-8:]
            window = "warm-up + timed region (timed region shorter than one poll)"
        sm = [float(r[2]) for r in inside]
        mx = [float(r[3]) for r in inside if r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names)
                   if any(r[6 + j].lower().startswith("active") for r in inside)]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "window": window}


def host_threads():
    """All host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return max(1, min(len(os.sched_getaffinity(0)), 64))
    except Exception:
        return max(1, min(os.cpu_count() or 1, 64))


def cpu_baselines(batch_rows, seconds_each, threads, rows_total=1_500_000):
    """Times the two CPU variants of BASELINE.md section 4 on a bounded sample of the workload
    (`batch_rows` rows per SGD step, same model shape / data distribution / hyper-parameters,
    about `seconds_each` seconds per variant):
      faithful_f64_kpass   fp64, forward = k passes over the row as FMModel.scala:48-51, dense
                           per-thread gradients summed in thread order (the treeAggregate analogue)
      tuned_f32_onepass    fp32, one pass per row, touched-only zeroing / reduction
    Returns {name: {"value": samples/s, "steps": n, "seconds": s}}, threads, description."""
    from oracle.capi import OracleFM, OracleFast32
    from sparkfm_b200 import synth
    rows_total = max(rows_total, batch_rows + batch_rows // 2)
    rp, idx, _, label = synth.ctr_csr(0, rows_total, N_FIELDS, N_SLOTS, DATA_SEED, ZIPF_S)
    rng = np.random.default_rng(0)

    def batch():
        return np.sort(rng.choice(rows_total, batch_rows, replace=False)).astype(np.int64)

    out = {}
    # (ii) tuned fp32
    orc = OracleFM(N_SLOTS, K, task=1, reg=REG)
    orc.init_v(0.0, 0.01, INIT_SEED)
    fast = OracleFast32(orc, threads)
    lab32 = label.astype(np.float32)
    fast.train_step(rp, idx, None, lab32, batch(), 1, STEP_SIZE)  # warm-up
    t0, n = time.perf_counter(), 0
    while True:
        fast.train_step(rp, idx, None, lab32, batch(), n + 2, STEP_SIZE)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds_each or n >= 400:
            break
    fast.close()
    out["tuned_f32_onepass"] = {"value": batch_rows * n / dt, "steps": n, "seconds": dt}
    # (i) faithful fp64
    val = np.ones(len(idx), dtype=np.float64)
    orc.train_step(rp, idx, val, label, batch(), 1, STEP_SIZE, threads=threads, faithful=True)
    t0, n = time.perf_counter(), 0
    while True:
        orc.train_step(rp, idx, val, label, batch(), n + 2, STEP_SIZE, threads=threads, faithful=True)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds_each or n >= 400:
            break
    out["faithful_f64_kpass"] = {"value": batch_rows * n / dt, "steps": n, "seconds": dt}
    desc = (f"SGD steps of {batch_rows} rows sampled from {rows_total} rows of the Criteo-shaped config "
            f"(n_slots={N_SLOTS}, k={K}), {threads} OpenMP threads, ~{seconds_each:.0f} s per variant; "
            "CPU restatement (oracle/fm_oracle.c), not Spark local[N]: the reference is Scala/Spark "
            "and no JVM exists in this image")
    return out, threads, desc


def cpu_baseline_block(batch_rows, seconds_each):
    variants, threads, desc = cpu_baselines(batch_rows, seconds_each, host_threads())
    best = max(variants, key=lambda k_: variants[k_]["value"])
    return {"value": variants[best]["value"], "unit": UNIT, "cores": threads, "kind": "port",
            "variant": best, "variants": variants, "sample": desc}


def run_reference(args, rank, world):
    """CPU arm: the reference cannot run here (Scala/Spark, no JVM), so the oracle port is timed
    on every host core (the inherited OMP_NUM_THREADS is ignored), same config / metric / unit.
    `value` is the FASTER of the two variants, so ratios against it are not against a strawman."""
    if rank != 0:
        return
    per_step = min(args.batch, 1_000_000)
    cpu = cpu_baseline_block(per_step, 12.0)
    val = cpu["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step / val * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if cpu["variant"].startswith("tuned") else "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows": args.rows, "global_batch_target": args.batch,
                   "global_batch": per_step, "n_slots": N_SLOTS, "k": K,
                   "cpu_arm": "each step is a bounded sample of the workload (cpu_baseline.sample)"},
        "cpu_baseline": cpu,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def guard_stdout():
    """stdout must carry exactly one JSON line: library chatter (NCCL prints its version banner to
    stdout) is diverted to stderr at the file-descriptor level; emit() writes to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def timed_train(hd, dist, world, it, k_steps, min_seconds, torch):
    """Device-timed training: exactly K steps per block (CUDA events on the library's stream,
    barrier + synchronize on both sides), repeated until the timed region is >= min_seconds.
    Returns (total ms = sum over blocks of the max over ranks, rows of all ranks, steps, it,
    loss history, ms of the first K-step block)."""
    def barrier():
        hd.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    total_ms, total_rows, total_steps, first_ms, hist_all = 0.0, 0.0, 0, None, []
    while True:
        hd.stats_reset()
        barrier()
        hd.timer_start()
        hist = hd.train(it, k_steps)
        ms = hd.timer_stop()
        barrier()
        it += k_steps
        rows = float(hd.stats()["train_rows"])
        t = torch.tensor([ms, rows], dtype=torch.float64, device="cuda")
        if world > 1:
            tm, ts = t.clone(), t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            ms, rows = float(tm[0]), float(ts[1])
        total_ms += ms
        total_rows += rows
        total_steps += k_steps
        hist_all += [float(x) for x in hist]
        if first_ms is None:
            first_ms = ms
        if total_ms >= min_seconds * 1e3 or total_steps >= 100000:
            break
    return total_ms, total_rows, total_steps, it, hist_all, first_ms


def model_digest(hd):
    import hashlib
    w0, w, v = hd.get_model()
    h = hashlib.sha256()
    h.update(np.float32(w0).tobytes())
    h.update(np.ascontiguousarray(w).tobytes())
    h.update(np.ascontiguousarray(v).tobytes())
    return h.hexdigest(), (float(w0), w, v)


def parity_n(args, rank, world, local_rank, torch, dist, card, cdf, off):
    """Driver-visible multi-GPU parity (N > 1): a small Criteo-shaped job trained for 3 steps with
    the peer-memory exchange and with the NCCL all-reduce; replicas must be bitwise identical
    across ranks in each mode, the two modes must agree (bitwise at 2 ranks: both add in rank
    order), and the first-step loss must match the fp64 CPU oracle on the same global batch."""
    from sparkfm_b200 import Handle, synth
    from sparkfm_b200.dist import init_comm, shard_range
    rows_total, frac, k_steps = 40_000 * world, 0.25, 3
    lo, hi = shard_range(rows_total, rank, world)
    out = {"world": world, "rows_total": rows_total, "fraction": frac, "steps": k_steps}
    models, losses = {}, {}
    for mode, env, sparse in (("peer", "1", "1"), ("peer_dense", "1", "0"), ("nccl", "0", "1")):
        os.environ["SFM_P2P"] = env
        os.environ["SFM_P2P_SPARSE"] = sparse
        h = Handle(N_SLOTS, K, task=1, reg=REG, step_size=STEP_SIZE, mini_batch_fraction=frac,
                   sampler_seed=SAMPLER_SEED, device=local_rank)
        h.init_model(0.0, 0.01, INIT_SEED)
        init_comm(h, device=f"cuda:{local_rank}")
        h.comm_broadcast_model()
        h.synth_ctr_dataset(hi - lo, lo, card, cdf, off, DATA_SEED)
        losses[mode] = [float(x) for x in h.train(1, k_steps)]
        dig, m = model_digest(h)
        models[mode] = m
        out[f"{mode}_comm_mode"] = h.comm_mode()
        gathered = [None] * world
        dist.all_gather_object(gathered, dig)
        out[f"{mode}_replicas_bitwise_equal"] = len(set(gathered)) == 1
        h.close()
    os.environ.pop("SFM_P2P", None)
    os.environ.pop("SFM_P2P_SPARSE", None)
    (_, wa, va), (_, wd, vd), (_, wb, vb) = models["peer"], models["peer_dense"], models["nccl"]
    # sparse (touched rows only) and dense peer-memory exchange add the same values in rank order
    out["peer_sparse_vs_dense_bitwise"] = bool(np.array_equal(wa, wd) and np.array_equal(va, vd))
    out["peer_vs_nccl_bitwise"] = bool(np.array_equal(wa, wb) and np.array_equal(va, vb))
    out["peer_vs_nccl_max_abs_diff"] = float(max(np.max(np.abs(wa - wb)), np.max(np.abs(va - vb))))
    out["loss_peer"], out["loss_nccl"] = losses["peer"], losses["nccl"]
    if rank == 0:   # fp64 oracle on the same global row list, same init
        from oracle.capi import OracleFM, sample_rows
        rp, idx, _, label = synth.ctr_csr(0, rows_total, N_FIELDS, N_SLOTS, DATA_SEED, ZIPF_S)
        orc = OracleFM(N_SLOTS, K, task=1, reg=REG)
        orc.init_v(0.0, 0.01, INIT_SEED)
        val = np.ones(len(idx), dtype=np.float64)
        ref = []
        for it in range(1, k_steps + 1):
            ids = sample_rows(SAMPLER_SEED, it, float(np.float32(frac)), 0, rows_total)
            ref.append(orc.train_step(rp, idx, val, label, ids, it, STEP_SIZE,
                                      threads=host_threads()) / max(len(ids), 1))
        out["loss_oracle"] = ref
        out["loss_rel_err_vs_oracle"] = float(max(abs(a - b) / abs(b) for a, b in zip(losses["peer"], ref)))
        out["ok"] = bool(out["peer_replicas_bitwise_equal"] and out["nccl_replicas_bitwise_equal"] and
                         out["peer_dense_replicas_bitwise_equal"] and
                         out["peer_sparse_vs_dense_bitwise"] and
                         out["loss_rel_err_vs_oracle"] <= 1e-4 and
                         (out["peer_vs_nccl_bitwise"] or world > 2) and
                         out["peer_vs_nccl_max_abs_diff"] <= 1e-6)
    return out


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=45_000_000, help="data set rows (whole job)")
    ap.add_argument("--batch", type=int, default=1_000_000,
                    help="GLOBAL mini-batch rows per SGD step (BASELINE configs[4]: 64k-1M); "
                         "each of the N GPUs samples batch/N rows of its shard: strong scaling")
    ap.add_argument("--weak-batch", type=int, default=1_000_000,
                    help="extra weak-scaling line at N > 1: mini-batch rows PER GPU (0 = skip)")
    ap.add_argument("--min-seconds", type=float, default=0.5, help="minimum timed region")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-partition", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch  # plumbing only: rendezvous + cross-rank max of the timings
    import torch.distributed as dist
    from sparkfm_b200 import Handle, pack_onehot, synth
    from sparkfm_b200.dist import init_comm, shard_range

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    rows_lo, rows_hi = shard_range(args.rows, rank, world)
    n_local = rows_hi - rows_lo
    batch_per_gpu = max(1, args.batch // world)
    frac = min(1.0, batch_per_gpu / max(n_local, 1))
    hd = Handle(N_SLOTS, K, task=1, reg=REG, step_size=STEP_SIZE, mini_batch_fraction=frac,
                sampler_seed=SAMPLER_SEED, device=local_rank)
    hd.init_model(0.0, 0.01, INIT_SEED)
    if world > 1:
        init_comm(hd, device=f"cuda:{local_rank}")
        hd.comm_broadcast_model()
    card = synth.ctr_field_log2_cards(N_FIELDS)
    cdf, off = synth.zipf_tables(card, ZIPF_S)
    hd.synth_ctr_dataset(n_local, rows_lo, card, cdf, off, DATA_SEED)   # generated in HBM

    def barrier():
        hd.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident arm: warm-up, then K-step blocks until >= min_seconds, device-timed
    it = 1
    clocks = ClockSampler(local_rank)
    clocks.start()
    hd.train(it, args.warmup)
    it += args.warmup
    barrier()
    clocks.mark_begin()
    ms, rows_all, steps_timed, it, hist, ms_k = timed_train(hd, dist, world, it, args.steps,
                                                            args.min_seconds, torch)
    clocks.mark_end()
    clk = clocks.stop()
    launches = hd.stats()["kernel_launches"] * (steps_timed // args.steps)
    value = rows_all / (ms * 1e-3)

    # ---- per-phase device times of the same steps (CUDA events per phase, extra pass)
    hd.set_phase_timing(True)
    hd.stats_reset()
    n_phase = 5
    hd.train(it, n_phase)
    it += n_phase
    ph = hd.stats()
    hd.set_phase_timing(False)
    rows_ph = ph["train_rows"] / n_phase
    nnz_ph = ph["train_nnz"] / n_phase
    phases = {k_: ph[k_] / n_phase for k_ in ("ms_forward", "ms_sort", "ms_reduce", "ms_allreduce",
                                              "ms_update")}
    peak, peak_src = measured_peaks()
    # kernels of one step, in launch order: (CUDA-event ms, algorithmic bytes per launch)
    kern = {
        "fm_forward_onehot16_kernel": (phases["ms_forward"], rows_ph * b_forward(N_FIELDS, K)),
        "bkt_scatter_kernel (+ bkt_count/offsets/plan: the transposition)": (phases["ms_sort"], 0.0),
        "bkt_pull_kernel (reduce-by-feature + update)": (phases["ms_reduce"],
                                                         b_reduce(N_FIELDS, K, N_SLOTS, rows_ph)),
    }
    if world > 1:
        kern["p2p_reduce_update_kernel (gradient exchange + update)"] = (phases["ms_allreduce"] + phases["ms_update"], 0.0)
    dom = max(kern, key=lambda n: kern[n][0])
    dom_ms, dom_bytes = kern[dom]
    ms_step = ms / steps_timed
    step_bytes = (rows_all / steps_timed) * b_train(N_FIELDS, K) + world * b_step(N_SLOTS, K)
    step_gbs = step_bytes / (ms_step * 1e-3) / 1e9
    traffic, tj = None, {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            tj = json.load(fh)
        traffic = tj[dom.split(" ")[0]]["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": dom, "unit": "GB/s", "peak": peak * world, "peak_source": peak_src,
        # whole SGD step: ALL algorithmic bytes (8m(k+2)+4 per sample + 12(1+n_slots(k+1)) per step
        # and GPU) / step time, against the measured HBM peak of the N GPUs; the transposition
        # carries no algorithmic bytes, so every microsecond of it lowers this fraction
        "achieved": step_gbs, "frac": step_gbs / (peak * world),
        "algorithmic_bytes_per_step": step_bytes,
        "traffic": traffic,
        "traffic_source": "ncu --set full capture of the dominant kernel, dram__bytes_read.sum + "
                          "dram__bytes_write.sum per launch at ~1 M rows (profiles/ncu_traffic.json)",
        "dominant": {"kernel": dom, "ms": dom_ms, "algorithmic_bytes_per_launch": dom_bytes,
                     "achieved": dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0,
                     "frac": dom_bytes / (dom_ms * 1e-3) / 1e9 / peak if dom_ms > 0 else 0.0,
                     "share_of_step": dom_ms / max(sum(v[0] for v in kern.values()), 1e-9)},
        "kernels": {n: {"ms": v[0], "algorithmic_bytes_per_launch": v[1],
                        "achieved": v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0} for n, v in kern.items()},
        "phase_ms": phases,
        "transposition": {"entries_per_step": nnz_ph, "bytes_per_step": nnz_ph * 16,
                          "note": "own traffic of bkt_count (4 B/entry read) + bkt_scatter (4 B read, "
                                  "4 B packed word written) + the read in bkt_pull (4 B): overhead, "
                                  "not part of the algorithmic bytes"},
    }

    # ---- weak-scaling line (N > 1): the per-GPU batch fixed at --weak-batch
    weak = None
    if world > 1 and args.weak_batch > 0:
        wfrac = min(1.0, args.weak_batch / max(n_local, 1))
        hd.set_hyper(REG[0], REG[1], REG[2], STEP_SIZE, wfrac)
        hd.train(it, 3)
        it += 3
        wms, wrows, wsteps, it, _, _ = timed_train(hd, dist, world, it, min(args.steps, 50), 0.3, torch)
        weak = {"batch_per_gpu": args.weak_batch, "value": wrows / (wms * 1e-3), "unit": UNIT,
                "ms_per_step": wms / wsteps, "steps": wsteps, "scaling": "weak",
                "roofline_step_frac": (wrows / wsteps * b_train(N_FIELDS, K) + world * b_step(N_SLOTS, K))
                / (wms / wsteps * 1e-3) / 1e9 / (peak * world)}
        hd.set_hyper(REG[0], REG[1], REG[2], STEP_SIZE, frac)

    # ---- e2e arm: host mini-batches through the C ABI, H2D + D2H inside the timed region.
    #      Headline: the compact one-hot staging format (sfm_stage_onehot: ids bit-packed at
    #      ID_BITS, 1-bit labels, unpacked by a device kernel); beside it the generic CSR form
    #      (sfm_stage_csr: int64 row_ptr, int32 ids, fp32 labels).
    e2e = None
    e2e_csr = None
    if not args.no_e2e:
        L = hd._L
        nb = 3
        id_bits = max(1, int(N_SLOTS - 1).bit_length())
        e2e_rows = batch_per_gpu
        csr_bufs, oh_bufs = [], []

        def pinned(a_):
            p_ = ctypes.c_void_p()
            assert L.sfm_host_alloc(ctypes.byref(p_), max(a_.nbytes, 4)) == 0
            ctypes.memmove(p_, a_.ctypes.data, a_.nbytes)
            return p_

        pack_s = 0.0
        for b in range(nb):
            idx, label = synth.ctr_rows(rows_lo + b * e2e_rows, rows_lo + (b + 1) * e2e_rows, card,
                                        cdf, off, N_SLOTS, DATA_SEED)
            arrs = (np.arange(e2e_rows + 1, dtype=np.int64) * N_FIELDS, idx.reshape(-1), label)
            csr_bufs.append([pinned(a_) for a_ in arrs])
            tp = time.perf_counter()
            packed, lbits = pack_onehot(idx, label, N_FIELDS, id_bits)
            pack_s += time.perf_counter() - tp
            oh_bufs.append([pinned(packed), pinned(lbits), packed.nbytes + lbits.nbytes])
        h2d_csr = (e2e_rows + 1) * 8 + e2e_rows * N_FIELDS * 4 + e2e_rows * 4
        h2d_oh = oh_bufs[0][2]

        def stage(kind, slot, b_):
            if kind == "onehot":
                q_ = oh_bufs[b_]
                hd.stage_onehot_raw(slot, q_[0], q_[1], None, e2e_rows, N_FIELDS, id_bits)
            else:
                q_ = csr_bufs[b_]
                hd.stage_csr_raw(slot, q_[0], q_[1], None, q_[2], e2e_rows)

        def run_pipelined(kind, n_steps, it0):
            # stage batch s+1 (async H2D on the copy stream) while batch s trains
            stage(kind, 0, 0)
            for s_ in range(n_steps):
                if s_ + 1 < n_steps:
                    stage(kind, (s_ + 1) & 1, (s_ + 1) % nb)
                hd.train_step_staged(s_ & 1, it0 + s_)
            return it0 + n_steps

        def time_e2e(kind, it0):
            it1 = run_pipelined(kind, 3, it0)
            n_steps = args.e2e_steps
            barrier()
            t0 = time.perf_counter()
            it1 = run_pipelined(kind, n_steps, it1)
            barrier()
            dt = time.perf_counter() - t0
            if world > 1:   # every rank must run the SAME number of steps (each step is a collective)
                td = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
                dt = float(td[0])
            if dt < args.min_seconds:   # short steps (small per-GPU batch): time a longer run
                n_steps = int(n_steps * args.min_seconds / max(dt, 1e-6)) + 1
                barrier()
                t0 = time.perf_counter()
                it1 = run_pipelined(kind, n_steps, it1)
                barrier()
                dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return e2e_rows * world * n_steps / float(tt[0]), n_steps, it1

        v_oh, n_oh, it = time_e2e("onehot", it)
        v_csr, n_csr, it = time_e2e("csr", it)
        e2e = {"value": v_oh, "unit": UNIT, "h2d_bytes_per_step": h2d_oh, "d2h_bytes_per_step": 36,
               "steps": n_oh, "rows_per_step_per_gpu": e2e_rows,
               "api": f"sfm_stage_onehot (pinned host buffer: {N_FIELDS} ids per row bit-packed at "
                      f"{id_bits} bits + 1 label bit per row -> device, copy stream, device unpack) "
                      "+ sfm_train_step_staged (returns the mean loss)",
               "timing": "wall clock around the C-ABI calls, every H2D/D2H inside the timed region",
               "host_packer": {"api": "sfm_pack_onehot (multi-threaded, outside the timed region like "
                                      "the CSR packing of the other form)",
                               "rows_per_s": nb * e2e_rows / max(pack_s, 1e-9)},
               "frac_of_resident_value": None}
        e2e_csr = {"value": v_csr, "unit": UNIT, "h2d_bytes_per_step": h2d_csr,
                   "d2h_bytes_per_step": 36, "steps": n_csr,
                   "api": "sfm_stage_csr (int64 row_ptr, int32 ids, fp32 labels) + sfm_train_step_staged"}
        for ptrs in csr_bufs:
            for p_ in ptrs:
                L.sfm_host_free(p_)
        for q_ in oh_bufs:
            L.sfm_host_free(q_[0])
            L.sfm_host_free(q_[1])

    # ---- PARTITION sampler (epoch-wise fixed mini-batches, transposition cached at first use):
    #      reported beside the Bernoulli headline, never instead of it
    part = None
    if not args.no_partition:
        hd.unload_dataset()
        hp = Handle(N_SLOTS, K, task=1, reg=REG, step_size=STEP_SIZE, mini_batch_fraction=frac,
                    sampler_seed=SAMPLER_SEED, device=local_rank, sampler_mode=1)
        hp.init_model(0.0, 0.01, INIT_SEED)
        if world > 1:
            init_comm(hp, device=f"cuda:{local_rank}")
            hp.comm_broadcast_model()
        hp.synth_ctr_dataset(n_local, rows_lo, card, cdf, off, DATA_SEED)
        n_parts = max(1, int(np.floor(1.0 / float(np.float32(frac)) + 0.5)))
        hp.synchronize()
        t0 = time.perf_counter()
        hp.train(1, n_parts)                       # first epoch: builds every batch's transposition
        hp.synchronize()
        first_epoch_s = time.perf_counter() - t0
        pms, prow, psteps, _, _, _ = timed_train(hp, dist, world, n_parts + 1, args.steps,
                                                 args.min_seconds, torch)
        pbytes = prow * b_train(N_FIELDS, K) + psteps * world * b_step(N_SLOTS, K)
        part = {"value": prow / (pms * 1e-3), "unit": UNIT, "ms_per_step": pms / psteps,
                "n_parts": n_parts, "first_epoch_s": first_epoch_s,
                "roofline_step_frac": pbytes / (pms * 1e-3) / 1e9 / (peak * world),
                "note": "SFM_SAMPLER_PARTITION: rows split once into n_parts disjoint random "
                        "mini-batches, iteration t uses batch (t-1) mod n_parts; each batch's "
                        "feature-sorted entry list is built at first use (first_epoch_s includes "
                        "all of them) and stays resident (+8 B per entry), like the reference's "
                        "cached transposeInput (DataSet.scala:48)"}
        hp.close()
        hd.synth_ctr_dataset(n_local, rows_lo, card, cdf, off, DATA_SEED)

    # ---- predict rows/s (FMModel.predict over resident rows; outputs copied back to the host)
    n_pred = min(n_local, 4_000_000)
    pbuf = ctypes.c_void_p()   # predictions land in pinned host memory (what a JVM host's direct buffer is)
    assert hd._L.sfm_host_alloc(ctypes.byref(pbuf), 4 * max(n_pred, 1)) == 0
    hd.predict_resident_raw(0, n_pred, pbuf)
    barrier()
    hd.timer_start()
    for _ in range(5):
        hd.predict_resident_raw(0, n_pred, pbuf)
    pms_ = hd.timer_stop() / 5
    hd._L.sfm_host_free(pbuf)
    pred_rows_s = n_pred / (pms_ * 1e-3)
    predict = {"value": pred_rows_s * world, "unit": "rows/s", "rows_per_call": n_pred,
               "ms_per_call": pms_,
               "roofline_frac": pred_rows_s * (4 * N_FIELDS * (K + 3) + 4) / 1e9 / peak,
               "note": "sfm_predict_resident incl. the D2H copy of the predictions into pinned host "
                       "memory; algorithmic bytes 4m(k+3)+4 per row"}
    comm_mode = hd.comm_mode()
    hd.close()

    # ---- multi-GPU parity (N > 1), CPU baseline (rank 0, N = 1 only)
    par = None
    if world > 1 and not args.no_parity:
        par = parity_n(args, rank, world, local_rank, torch, dist, card, cdf, off)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_block(min(args.batch, 1_000_000), 10.0)

    if rank == 0:
        if e2e:
            e2e["frac_of_resident_value"] = e2e["value"] / value
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "steps_timed": steps_timed, "timed_region_ms": ms,
            "first_k_steps": {"steps": args.steps, "ms": ms_k},
            "config": {"workload": WORKLOAD, "rows": args.rows, "global_batch_target": args.batch,
                       "global_batch": int(rows_all / steps_timed), "batch_per_gpu": batch_per_gpu,
                       "n_slots": N_SLOTS, "k": K, "sampler": "Bernoulli (a fresh random batch, "
                       "transposed on the device, every step)",
                       "parallelism": f"dp{world}", "gradient_exchange": comm_mode,
                       "l2": "inputs larger than L2: each step streams a fresh sampled batch "
                             "out of a 7 GB resident set; model 68 MB + per-batch scratch"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_csr": e2e_csr,
            "predict": predict,
            "partition_sampler": part, "weak_scaling": weak, "parity_n": par,
            "comm_mode": comm_mode, "gpu_launches": int(launches),
            "clocks": clk, "loss_first_last": [float(hist[0]), float(hist[-1])],
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
