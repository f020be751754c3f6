#!/usr/bin/env python
"""Runs BASELINE.json configs 1-4 at FULL size on one B200 through the public API, checks the
per-iteration loss against the fp64 CPU oracle where the oracle finishes in seconds (C1, C2, and
C3 through its sampled mini-batches) and size-independent properties elsewhere, and writes
gpurun_out/configs_r2.json.  (Config 3's throughput is bench.py; config 5 is bench.py under
torchrun at N = 1, 2, 4, 8.)"""
import json
import os
import sys
import time

os.environ.setdefault("OMP_WAIT_POLICY", "passive")  # the oracle's OpenMP team must not spin
import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import capi as ocapi  # noqa: E402
from oracle.capi import OracleFM  # noqa: E402
from sparkfm_b200 import FMUtils, Handle, synth  # noqa: E402

OUT = {}


def timed_steps(hd, it0, n):
    """Device-timed (CUDA events on the library's stream); best of two runs of n iterations."""
    best = None
    for rep in range(2):
        hd.synchronize()
        hd.timer_start()
        hist = hd.train(it0 + rep * n, n)
        ms = hd.timer_stop()
        best = ms if best is None else min(best, ms)
    return hist, best / n


def config1(n_check=8, n_timed=50):
    """100k x 10k LIBSVM-format binary classification, ~20 nnz/row, k=8, logistic, full batch."""
    row_ptr, idx, val, label = synth.classification_c1()
    t0 = time.perf_counter()
    text = synth.to_libfm_text(row_ptr, idx, val, label).encode()
    t_text = time.perf_counter() - t0
    path = "/tmp/c1.libfm"
    open(path, "wb").write(text)
    t0 = time.perf_counter()
    ds = FMUtils.loadLibFMFile(path)
    t_parse = time.perf_counter() - t0
    assert np.array_equal(ds.idx, idx) and np.array_equal(ds.row_ptr, row_ptr)      # bit-exact packing
    assert np.array_equal(ds.val, val.astype(np.float64)) and ds.dimension == int(idx.max())
    n_slots, k = ds.dimension + 1, 8
    reg, step = (0.0, 1e-4, 1e-4), 1.0
    hd = Handle(n_slots, k, task=1, reg=reg, step_size=step, mini_batch_fraction=1.0)
    hd.init_model(0.0, 0.01, 1)
    w0, w, v = hd.get_model()
    hd.load_dataset(ds.row_ptr, ds.idx, ds.val, ds.labels)
    orc = OracleFM(n_slots, k, task=1, reg=tuple(float(np.float32(r)) for r in reg))
    orc.set_model(w0, w, v)
    ids = np.arange(ds.size, dtype=np.int64)
    worst = 0.0
    losses = []
    for it in range(1, n_check + 1):
        lo = orc.train_step(ds.row_ptr, ds.idx, ds.val, ds.labels, ids, it, step, threads=8) / ds.size
        lg, batch = hd.train_step(it)
        assert batch == ds.size
        worst = max(worst, abs(lg - lo) / lo)
        losses.append(lg)
    assert worst < 1e-4, worst
    hist, ms = timed_steps(hd, n_check + 1, n_timed)
    pred = hd.predict_resident(0, ds.size)
    want = orc.predict(ds.row_ptr, ds.idx, ds.val, fast=True, threads=8)
    # models have drifted apart by fp32 rounding over 8+ steps; compare predictions of the SAME model
    o2 = OracleFM(n_slots, k, task=1)
    o2.set_model(*hd.get_model())
    want = o2.predict(ds.row_ptr, ds.idx, ds.val, fast=True, threads=8)
    perr = float(np.max(np.abs(pred - want) / np.maximum(np.abs(want), np.mean(np.abs(want)))))
    assert perr < 1e-5, perr
    OUT["C1"] = {"rows": ds.size, "n_slots": n_slots, "k": k, "nnz": int(ds.row_ptr[-1]),
                 "text_bytes": len(text), "parse_s": t_parse, "parse_MBps": len(text) / t_parse / 1e6,
                 "loss_rel_err_max": worst, "loss_iters_checked": n_check, "loss_first_last": [losses[0], float(hist[-1])],
                 "ms_per_full_batch_step": ms, "samples_per_s": ds.size / (ms * 1e-3),
                 "predict_rel_err_max": perr, "accuracy": hd.evaluate()["accuracy"]}
    hd.close()


def config2(n_check=10, n_timed=100):
    """1M x 100k regression, ~50 nnz/row, values N(0,1), k=16, squared loss, miniBatchFraction 0.1."""
    row_ptr, idx, val, y = synth.regression_c2()
    n_rows, n_slots, k = len(y), 100_000, 16
    reg, step, frac = (0.0, 1e-4, 1e-3), 0.02, 0.1
    hd = Handle(n_slots, k, task=0, reg=reg, step_size=step, mini_batch_fraction=frac, sampler_seed=42)
    hd.init_model(0.0, 0.01, 1)
    w0, w, v = hd.get_model()
    hd.load_dataset(row_ptr, idx, val, y)
    orc = OracleFM(n_slots, k, task=0, reg=tuple(float(np.float32(r)) for r in reg))
    orc.set_model(w0, w, v)
    v64 = val.astype(np.float64)
    worst = 0.0
    f32 = float(np.float32(frac))
    s32 = float(np.float32(step))
    losses = []
    for it in range(1, n_check + 1):
        ids = ocapi.sample_rows(42, it, f32, 0, n_rows)
        lo = orc.train_step(row_ptr, idx, v64, y, ids, it, s32, threads=8) / len(ids)
        lg, batch = hd.train_step(it)
        assert batch == len(ids)
        worst = max(worst, abs(lg - lo) / lo)
        losses.append(lg)
    assert worst < 1e-4, worst
    hist, ms = timed_steps(hd, n_check + 1, n_timed)
    rows = hd.stats()["train_rows"]
    ev = hd.evaluate()
    OUT["C2"] = {"rows": n_rows, "n_slots": n_slots, "k": k, "nnz": int(row_ptr[-1]),
                 "loss_rel_err_max": worst, "loss_iters_checked": n_check, "loss_first_last": [losses[0], float(hist[-1])],
                 "ms_per_step": ms, "batch_rows": int(round(n_rows * frac)),
                 "samples_per_s": n_rows * frac / (ms * 1e-3), "rmse_after": ev["rmse"],
                 "train_rows_total": rows}
    hd.close()


def config4(n_timed=20, modes=((0, "bernoulli"), (1, "partition"))):
    """Avazu-shaped: 24 one-hot fields, 10M hashed features, k=64, logistic -- model REPLICATED on
    one GPU here (the row-sharded form needs >= 2 GPUs: tests/test_gpu_multi.py, scripts/bench_c4.py)."""
    n_fields, n_slots, k, n_rows, batch = 24, 10_000_000, 64, 8_000_000, 500_000
    card = synth.ctr_field_log2_cards(n_fields)
    cdf, off = synth.zipf_tables(card)
    res = {}
    for mode, name in modes:
        hd = Handle(n_slots, k, task=1, reg=(0.0, 0.0, 1e-5), step_size=0.1,
                    mini_batch_fraction=batch / n_rows, sampler_seed=42, sampler_mode=mode)
        hd.init_model(0.0, 0.01, 1)
        hd.synth_ctr_dataset(n_rows, 0, card, cdf, off, 20260104)
        n_parts = round(n_rows / batch)
        warm = n_parts if mode == 1 else 3
        h0 = hd.train(1, warm)
        hist, ms = timed_steps(hd, warm + 1, n_timed)
        res[name] = {"ms_per_step": ms, "samples_per_s": batch / (ms * 1e-3),
                     "loss_first_last": [float(h0[0]), float(hist[-1])]}
        assert hist[-1] < h0[0]
        if mode == 0:
            # determinism at full size: rerun from the same state gives the same bits
            a = Handle(n_slots, k, task=1, reg=(0.0, 0.0, 1e-5), step_size=0.1,
                       mini_batch_fraction=batch / n_rows, sampler_seed=42)
            a.init_model(0.0, 0.01, 1)
            a.synth_ctr_dataset(n_rows, 0, card, cdf, off, 20260104)
            ha = a.train(1, warm)
            assert np.array_equal(ha, h0)
            a.close()
            res["bitwise_rerun"] = True
        hd.close()
    m = n_fields
    res.update({"rows": n_rows, "n_slots": n_slots, "k": k, "batch": batch,
                "algorithmic_bytes_per_sample": 8 * m * (k + 2) + 4,
                "roofline_samples_per_s": 6554.2e9 / (8 * m * (k + 2) + 4)})
    OUT["C4_replicated_1gpu"] = res


def config3(n_check=3, n_rows=45_000_000, batch=1_000_000):
    """Criteo-shaped, the HEADLINE config at full size: 39 one-hot fields, 1 M hashed features,
    Zipf 1.1, k=16, logistic loss, 45 M rows generated on the device, Bernoulli mini-batches of
    ~1 M rows.  The oracle cannot hold 45 M rows in seconds, but a row depends on its number only,
    so each iteration's sampled rows are rebuilt on the host (numpy twin of the device generator)
    and fed to the fp64 oracle: exact batch sizes, per-iteration loss at 1e-4, predictions of the
    trained model at 1e-5; plus size-independent properties: the device rows at the far end of the
    data set equal the twin's bit for bit, and a rerun reproduces loss and model bits."""
    import hashlib
    n_fields, n_slots, k = 39, 1_000_000, 16
    reg, step, frac, data_seed = (0.0, 0.0, 1e-5), 0.1, batch / n_rows, 20260103
    card = synth.ctr_field_log2_cards(n_fields)
    cdf, off = synth.zipf_tables(card)

    def run(n_iters):
        hd = Handle(n_slots, k, task=1, reg=reg, step_size=step, mini_batch_fraction=frac, sampler_seed=42)
        hd.init_model(0.0, 0.01, 1)
        hd.synth_ctr_dataset(n_rows, 0, card, cdf, off, data_seed)
        return hd

    def digest(hd):
        w0, w, v = hd.get_model()
        h = hashlib.sha256(np.float32(w0).tobytes())
        h.update(np.ascontiguousarray(w).tobytes())
        h.update(np.ascontiguousarray(v).tobytes())
        return h.hexdigest()

    hd = run(n_check)
    assert hd.dataset_info()[:2] == (n_rows, n_rows * n_fields)
    # the device generator at the far end of the data set (64-bit row arithmetic) vs the numpy twin
    lo = n_rows - 4096
    _, d_idx, _, d_label = hd.get_dataset_rows(lo, n_rows, with_val=False)
    t_idx, t_label = synth.ctr_rows(lo, n_rows, card, cdf, off, n_slots, data_seed)
    assert np.array_equal(d_idx.reshape(-1, n_fields), t_idx) and np.array_equal(d_label, t_label)
    w0, w, v = hd.get_model()
    orc = OracleFM(n_slots, k, task=1, reg=tuple(float(np.float32(r)) for r in reg))
    orc.set_model(w0, w, v)
    f32, s32 = float(np.float32(frac)), float(np.float32(step))
    worst, losses, batches = 0.0, [], []
    for it in range(1, n_check + 1):
        ids = ocapi.sample_rows(42, it, f32, 0, n_rows)
        idx, label = synth.ctr_rows_at(ids, card, cdf, off, n_slots, data_seed)
        rp = np.arange(len(ids) + 1, dtype=np.int64) * n_fields
        lo_ = orc.train_step(rp, idx.reshape(-1), np.ones(idx.size), label,
                             np.arange(len(ids), dtype=np.int64), it, s32, threads=8) / len(ids)
        lg, nb = hd.train_step(it)
        assert nb == len(ids), (nb, len(ids))      # the sampler's batch, exactly
        worst = max(worst, abs(lg - lo_) / lo_)
        losses.append(lg)
        batches.append(int(nb))
    assert worst < 1e-4, worst
    # predictions of the trained model on rows from both ends, against the oracle holding the
    # SAME (device) model
    o2 = OracleFM(n_slots, k, task=1)
    o2.set_model(*hd.get_model())
    perr = 0.0
    for a, b in ((0, 50_000), (n_rows - 50_000, n_rows)):
        pred = hd.predict_resident(a, b)
        idx, _ = synth.ctr_rows(a, b, card, cdf, off, n_slots, data_seed)
        rp = np.arange(b - a + 1, dtype=np.int64) * n_fields
        want = o2.predict(rp, idx.reshape(-1), np.ones(idx.size), fast=True, threads=8)
        perr = max(perr, float(np.max(np.abs(pred - want) / np.maximum(np.abs(want), np.mean(np.abs(want))))))
    assert perr < 1e-5, perr
    dig = digest(hd)
    hd.close()
    # rerun from the same seeds through the device-side loop (sfm_train): same loss and model bits
    hb = run(n_check)
    hist = hb.train(1, n_check)
    same = bool(np.array_equal(np.asarray(hist, dtype=np.float64), np.asarray(losses, dtype=np.float64))
                and digest(hb) == dig)
    hb.close()
    assert same
    OUT["C3"] = {"rows": n_rows, "n_slots": n_slots, "k": k, "batch_rows": batches,
                 "loss_rel_err_max": worst, "loss_iters_checked": n_check, "losses": losses,
                 "predict_rel_err_max": perr, "bitwise_rerun": same,
                 "far_end_rows_bit_identical_to_twin": True}


if __name__ == "__main__":
    for fn in (config1, config2, config3, config4):
        t0 = time.perf_counter()
        fn()
        print(fn.__name__, "ok", f"{time.perf_counter() - t0:.1f}s", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs_r2.json"), "w") as fh:
        json.dump(OUT, fh, indent=1)
    print(json.dumps(OUT, indent=1))
