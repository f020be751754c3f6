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


def cpu_oracle_throughput(batch_rows, steps, threads, rows_total=1_500_000):
    """Times the fp64 CPU oracle (OpenMP) on a bounded sample of the workload: `steps` SGD
    iterations of `batch_rows` rows each, same model shape / data distribution / hyper-parameters.
    Returns (samples/s, threads, description)."""
    from oracle.capi import OracleFM, max_threads
    from sparkfm_b200 import synth
    threads = min(threads or max_threads(), 64)
    rp, idx, _, label = synth.ctr_csr(0, rows_total, N_FIELDS, N_SLOTS, DATA_SEED, ZIPF_S)
    val = np.ones(len(idx), dtype=np.float64)
    orc = OracleFM(N_SLOTS, K, task=1, reg=REG)
    orc.init_v(0.0, 0.01, INIT_SEED)
    rng = np.random.default_rng(0)
    ids = [np.sort(rng.choice(rows_total, batch_rows, replace=False)).astype(np.int64)
           for _ in range(steps + 1)]
    orc.train_step(rp, idx, val, label, ids[0], 1, STEP_SIZE, threads=threads)  # warm-up
    t0 = time.perf_counter()
    for s in range(steps):
        orc.train_step(rp, idx, val, label, ids[s + 1], s + 2, STEP_SIZE, threads=threads)
    dt = time.perf_counter() - t0
    return batch_rows * steps / dt, threads, (
        f"{steps} SGD steps x {batch_rows} rows of the Criteo-shaped config (n_slots={N_SLOTS}, "
        f"k={K}), fp64 OpenMP oracle, {threads} threads, {dt:.1f} s")


def run_reference(args, rank, world):
    if rank != 0:
        return
    per_step = 1_000_000
    val, threads, sample = cpu_oracle_throughput(per_step, min(max(args.steps, 1), 120), None)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step / val * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C3 Criteo-shaped CTR: 39 one-hot fields, 1M hashed features, Zipf 1.1, "
                               "k=16, logistic, 45M rows (CPU arm: bounded sample)",
                   "per_step_rows": per_step},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample + "; CPU restatement (OpenMP), not Spark local[N]: the "
                                            "reference is Scala/Spark and no JVM exists in this image"},
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


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=45_000_000, help="data set rows (whole job)")
    ap.add_argument("--batch", type=int, default=1_000_000, help="mini-batch rows per GPU")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-partition", action="store_true")
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
    from sparkfm_b200 import Handle, synth
    from sparkfm_b200.dist import init_comm, shard_range

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    rows_lo, rows_hi = shard_range(args.rows, rank, world)
    n_local = rows_hi - rows_lo
    frac = min(1.0, args.batch / max(n_local, 1))
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

    # ---- resident arm: warm-up, then exactly K steps, device-timed
    it = 1
    clocks = ClockSampler(local_rank)
    clocks.start()
    hd.train(it, args.warmup)
    it += args.warmup
    hd.stats_reset()
    barrier()
    clocks.mark_begin()
    hd.timer_start()
    hist = hd.train(it, args.steps)
    ms = hd.timer_stop()
    barrier()
    clocks.mark_end()
    clk = clocks.stop()
    it += args.steps
    st = hd.stats()
    rows_done, nnz_done, launches = st["train_rows"], st["train_nnz"], st["kernel_launches"]
    t = torch.tensor([ms, float(rows_done)], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms, rows_all = float(tm[0]), float(ts[1])
    else:
        rows_all = float(rows_done)
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
    phases = {k_: ph[k_] / n_phase for k_ in ("ms_forward", "ms_sort", "ms_reduce", "ms_allreduce",
                                              "ms_update")}
    peak, peak_src = measured_peaks()
    kern = {"fm_forward_kernel": (phases["ms_forward"], rows_ph * b_forward(N_FIELDS, K)),
            "fm_pull_kernel": (phases["ms_reduce"], b_reduce(N_FIELDS, K, N_SLOTS, rows_ph))}
    dom = max(kern, key=lambda n: kern[n][0])
    dom_ms, dom_bytes = kern[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    step_bytes = (rows_all / args.steps) * b_train(N_FIELDS, K) + world * b_step(N_SLOTS, K)
    step_gbs = step_bytes / (ms / args.steps * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            tj = json.load(fh)
        key = {"fm_forward_kernel": "fm_forward_onehot16_kernel",
               "fm_pull_kernel": "fm_pull_chunks_kernel"}[dom]
        traffic = tj[key]["dram_bytes_per_launch"]
        if dom == "fm_pull_kernel":
            traffic += tj["fm_pull_finalize_kernel"]["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "ncu --set full capture of the same kernel(s), dram__bytes_read.sum"
                                  " + dram__bytes_write.sum per launch (profiles/ncu_traffic.json)",
                "algorithmic_bytes_per_launch": dom_bytes, "peak_source": peak_src,
                "step": {"achieved": step_gbs, "frac": step_gbs / (peak * world),
                         "note": "whole step incl. sort/scan overhead, algorithmic bytes "
                                 "8m(k+2)+4 per sample + 12(1+n_slots(k+1)) per step"},
                "phase_ms": phases}
    # the transposition sort (sfm_radix.cu) is overhead on top of the algorithmic bytes; its own
    # traffic: per 10-bit pass the keys are read twice (count + scatter) and key + row written once
    nnz_ph = ph["train_nnz"] / n_phase
    key_bits = max(1, int(N_SLOTS - 1).bit_length())
    passes = -(-key_bits // 10)
    sort_bytes = nnz_ph * (passes * (4 + 8 + 8) - 4)      # first pass derives the row: no payload read
    if phases["ms_sort"] > 0:
        sort_gbs = sort_bytes / (phases["ms_sort"] * 1e-3) / 1e9
        roofline["sort"] = {"kernels": "radix_count_kernel + radix_scatter_kernel (x%d passes)" % passes,
                            "bytes_per_step": sort_bytes, "achieved": sort_gbs,
                            "frac": sort_gbs / peak,
                            "note": "own stable LSD radix sort, not part of the algorithmic bytes"}

    # ---- e2e arm: host CSR mini-batches through sfm_train_step_csr
    e2e = None
    if not args.no_e2e:
        L = hd._L
        nb = 3
        bufs = []
        e2e_rows = args.batch
        for b in range(nb):
            idx, label = synth.ctr_rows(rows_lo + b * e2e_rows, rows_lo + (b + 1) * e2e_rows, card,
                                        cdf, off, N_SLOTS, DATA_SEED)
            arrs = (np.arange(e2e_rows + 1, dtype=np.int64) * N_FIELDS, idx.reshape(-1), label)
            ptrs = []
            for a in arrs:
                p = ctypes.c_void_p()
                assert L.sfm_host_alloc(ctypes.byref(p), a.nbytes) == 0
                ctypes.memmove(p, a.ctypes.data, a.nbytes)
                ptrs.append(p)
            bufs.append(ptrs)
        h2d = (e2e_rows + 1) * 8 + e2e_rows * N_FIELDS * 4 + e2e_rows * 4
        def run_pipelined(n_steps, it0):
            # stage batch s+1 (async H2D on the copy stream) while batch s trains
            p = bufs[0]
            hd.stage_csr_raw(0, p[0], p[1], None, p[2], e2e_rows)
            for s in range(n_steps):
                if s + 1 < n_steps:
                    q = bufs[(s + 1) % nb]
                    hd.stage_csr_raw((s + 1) & 1, q[0], q[1], None, q[2], e2e_rows)
                hd.train_step_staged(s & 1, it0 + s)
            return it0 + n_steps

        it = run_pipelined(3, it)
        barrier()
        t0 = time.perf_counter()
        it = run_pipelined(args.e2e_steps, it)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": e2e_rows * world * args.e2e_steps / float(tt[0]), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 36, "steps": args.e2e_steps,
               "api": "sfm_stage_csr (pinned host CSR -> device, copy stream) + "
                      "sfm_train_step_staged (returns the mean loss)",
               "timing": "wall clock around the C-ABI calls, every H2D/D2H inside the timed region"}
        for ptrs in bufs:
            for p in ptrs:
                L.sfm_host_free(p)

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
        hp.stats_reset()
        if world > 1:
            dist.barrier()
        hp.timer_start()
        hp.train(n_parts + 1, args.steps)
        pms = hp.timer_stop()
        stp = hp.stats()
        tp = torch.tensor([pms, float(stp["train_rows"])], dtype=torch.float64, device="cuda")
        if world > 1:
            tmx = tp.clone()
            dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
            tsm = tp.clone()
            dist.all_reduce(tsm, op=dist.ReduceOp.SUM)
            pms, prow = float(tmx[0]), float(tsm[1])
        else:
            prow = float(stp["train_rows"])
        pbytes = prow * b_train(N_FIELDS, K) + args.steps * world * b_step(N_SLOTS, K)
        part = {"value": prow / (pms * 1e-3), "unit": UNIT, "ms_per_step": pms / args.steps,
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
    hd.predict_resident(0, n_pred)
    barrier()
    hd.timer_start()
    for _ in range(3):
        hd.predict_resident(0, n_pred)
    pms = hd.timer_stop() / 3
    pred_rows_s = n_pred / (pms * 1e-3)
    predict = {"value": pred_rows_s * world, "unit": "rows/s", "rows_per_call": n_pred,
               "ms_per_call": pms,
               "roofline_frac": pred_rows_s * (4 * N_FIELDS * (K + 3) + 4) / 1e9 / peak,
               "note": "sfm_predict_resident incl. the D2H copy of the predictions; algorithmic "
                       "bytes 4m(k+3)+4 per row"}

    # ---- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, threads, sample = cpu_oracle_throughput(args.batch, 60, None)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": sample + "; CPU restatement (OpenMP), not Spark local[N]"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3 Criteo-shaped CTR: 39 one-hot fields, 1M hashed features, "
                                   "Zipf 1.1, k=16, logistic loss, 45M rows resident in HBM",
                       "rows": args.rows, "batch_per_gpu": args.batch,
                       "global_batch": int(rows_all / args.steps), "n_slots": N_SLOTS, "k": K,
                       "parallelism": f"dp{world}", "gradient_exchange": hd.comm_mode(),
                       "l2": "inputs larger than L2: each step streams a fresh sampled batch "
                             "(>=156 MB of indices out of a 7 GB resident set)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "predict": predict,
            "partition_sampler": part, "comm_mode": hd.comm_mode(),
            "gpu_launches": int(launches),
            "clocks": clk, "loss_first_last": [float(hist[0]), float(hist[-1])],
        }
        emit(line)
    hd.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
