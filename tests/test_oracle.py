"""CPU tests of the oracle itself (no GPU): the C oracle (oracle/fm_oracle.c) and the independent
numpy restatement (oracle/fm_numpy.py) against the builder-authored exact known answers
(tests/golden/predict_kat.json -- the reference has no golden vectors), against
each other, and against first principles (finite differences for the gradient spec)."""
import json
import math
import os

import numpy as np
import pytest

from oracle import capi, fm_numpy as fn
from oracle.capi import OracleFM
from sparkfm_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "predict_kat.json")


def _case_arrays(c):
    rp, idx, val = [0], [], []
    for r in c["rows"]:
        idx += r["idx"]
        val += r["val"]
        rp.append(len(idx))
    n, k = c["n_slots"], c["k"]
    return (np.array(rp, np.int64), np.array(idx, np.int32), np.array(val, np.float64),
            np.array(c["w"]), np.array(c["v"], np.float64).reshape(n, k))


def test_golden_file_is_reproducible():
    from tests.golden import make_predict_kat as mk  # noqa: F401  (regenerates CASES on import)
    disk = json.load(open(GOLDEN))["cases"]
    assert [c["name"] for c in disk] == [c["name"] for c in mk.CASES]
    for a, b in zip(disk, mk.CASES):
        assert a["predict"] == b["predict"] and a["predict_exact"] == b["predict_exact"]


def test_predict_known_answers_both_oracles():
    for c in json.load(open(GOLDEN))["cases"]:
        rp, idx, val, w, v = _case_arrays(c)
        o = OracleFM(c["n_slots"], c["k"], k0=c["k0"], k1=c["k1"])
        o.set_model(c["w0"], w, v)
        assert o.predict(rp, idx, val).tolist() == c["predict"], c["name"]
        assert o.predict(rp, idx, val, fast=True).tolist() == c["predict"], c["name"]
        assert fn.predict(c["w0"], w, v, rp, idx, val, c["k0"], c["k1"]).tolist() == c["predict"]
        assert fn.predict_vec(c["w0"], w, v, rp, idx, val, c["k0"], c["k1"]).tolist() == c["predict"]


def test_hand_computed_two_feature_case():
    """yhat = w0 + w1 x1 + w2 x2 + <v1, v2> x1 x2, worked by hand: 1/4 + 1/2 - 1/4 + (1/16 - 3/16)."""
    w = np.array([0, 0.5, -0.25, 0.375])
    v = np.array([[0, 0], [0.5, -0.25], [0.125, 0.75], [-0.5, 0.5]])
    o = OracleFM(4, 2)
    o.set_model(0.25, w, v)
    assert o.predict([0, 2], np.array([1, 2], np.int32), np.array([1.0, 1.0]))[0] == 0.375


@pytest.mark.parametrize("k", [1, 5, 16])
def test_c_oracle_matches_numpy_restatement(k):
    rng = np.random.default_rng(k)
    n_slots, n_rows = 400, 300
    row_ptr, idx, val = synth.ragged_rows(n_rows, n_slots, 9, seed=k, values="normal")
    row_ptr = row_ptr.copy()
    w0, w, v = 0.3, rng.normal(0, 1, n_slots), rng.normal(0, 0.5, (n_slots, k))
    o = OracleFM(n_slots, k)
    o.set_model(w0, w, v)
    a = o.predict(row_ptr, idx, val.astype(np.float64))
    b = fn.predict(w0, w, v, row_ptr, idx, val.astype(np.float64))
    assert np.array_equal(a, b)            # same fold order -> same bits
    c = o.predict(row_ptr, idx, val.astype(np.float64), fast=True, threads=2)
    assert np.allclose(a, c, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("task", [0, 1])
def test_gradient_is_the_derivative_of_the_loss(task):
    """The written spec against central finite differences of the loss; for
    regression the reported loss is (yhat-y)^2 and the gradient is that of HALF of it."""
    rng = np.random.default_rng(7 + task)
    n_slots, k, n_rows = 12, 3, 9
    row_ptr, idx, val = synth.ragged_rows(n_rows, n_slots, 4, seed=3, values="normal")
    val = val.astype(np.float64)
    label = np.where(rng.random(n_rows) < 0.5, 1.0, -1.0) if task else rng.normal(0, 1, n_rows)
    w0, w, v = 0.2, rng.normal(0, 0.5, n_slots), rng.normal(0, 0.5, (n_slots, k))
    ids = np.arange(n_rows)

    def loss_at(w0_, w_, v_):
        o = OracleFM(n_slots, k, task=task)
        o.set_model(w0_, w_, v_)
        tot = o.gradient(row_ptr, idx, val, label, ids)[3]
        return tot * (1.0 if task else 0.5)

    o = OracleFM(n_slots, k, task=task)
    o.set_model(w0, w, v)
    gv, gw, gw0, _ = o.gradient(row_ptr, idx, val, label, ids)
    nv, nw, nw0, _ = fn.gradient(task, w0, w, v, row_ptr, idx, val, label, ids)
    assert np.allclose(gv, nv, rtol=1e-12, atol=1e-13) and np.allclose(gw, nw, rtol=1e-12, atol=1e-13)
    assert math.isclose(gw0, nw0, rel_tol=1e-12)
    eps = 1e-6
    assert abs((loss_at(w0 + eps, w, v) - loss_at(w0 - eps, w, v)) / (2 * eps) - gw0) < 1e-6
    for i in range(n_slots):
        wp, wm = w.copy(), w.copy()
        wp[i] += eps
        wm[i] -= eps
        assert abs((loss_at(w0, wp, v) - loss_at(w0, wm, v)) / (2 * eps) - gw[i]) < 1e-6
        for f in range(k):
            vp, vm = v.copy(), v.copy()
            vp[i, f] += eps
            vm[i, f] -= eps
            assert abs((loss_at(w0, w, vp) - loss_at(w0, w, vm)) / (2 * eps) - gv[i, f]) < 1e-6


def test_loss_mult_extremes_are_finite():
    for yhat in (-800.0, -30.0, 0.0, 30.0, 800.0):
        for lab in (1.0, 0.0, -1.0):
            ls, mu = fn.loss_mult(1, yhat, lab)
            assert math.isfinite(ls) and math.isfinite(mu) and ls >= 0.0 and abs(mu) <= 1.0


def test_update_rule_by_hand():
    """theta <- theta - (step/sqrt(t)) * (g/|B| + lambda*theta), all slots, t = 4 -> eta = step/2."""
    o = OracleFM(2, 1, task=0, reg=(0.5, 0.25, 0.125))
    o.set_model(1.0, np.array([2.0, -4.0]), np.array([[8.0], [16.0]]))
    # one row: feature 0, x = 1, label 0  ->  yhat = w0 + w_0 = 3, mult = 3 (pair term is 0)
    loss = o.train_step([0, 1], np.array([0], np.int32), np.array([1.0]), np.array([0.0]),
                        np.array([0]), it=4, step_size=1.0, batch_count=1)
    assert loss == 9.0
    eta = 0.5
    assert o.w0.value == 1.0 - eta * (3.0 + 0.5 * 1.0)
    assert o.w.tolist() == [2.0 - eta * (3.0 + 0.25 * 2.0), -4.0 - eta * (0.25 * -4.0)]
    # dV_00 = (x*s - v*x^2)*mult = 0 for a single-feature row; only the L2 decay moves V
    assert o.v.reshape(-1).tolist() == [8.0 - eta * 0.125 * 8.0, 16.0 - eta * 0.125 * 16.0]


def test_mt_step_matches_single_thread():
    rng = np.random.default_rng(3)
    n_slots, k, n_rows = 300, 4, 500
    row_ptr, idx, val = synth.ragged_rows(n_rows, n_slots, 8, seed=9, values="normal")
    label = rng.normal(0, 1, n_rows)
    w0, w, v = 0.1, rng.normal(0, 0.1, n_slots), rng.normal(0, 0.1, (n_slots, k))
    a = OracleFM(n_slots, k, reg=(0.0, 0.01, 0.01))
    b = OracleFM(n_slots, k, reg=(0.0, 0.01, 0.01))
    ids = np.arange(0, n_rows, 2)
    for o in (a, b):
        o.set_model(w0, w, v)
    la = a.train_step(row_ptr, idx, val, label, ids, 2, 0.1)
    lb = b.train_step(row_ptr, idx, val, label, ids, 2, 0.1, threads=3)
    assert math.isclose(la, lb, rel_tol=1e-12)
    assert np.allclose(a.v, b.v, rtol=1e-12, atol=1e-14) and np.allclose(a.w, b.w, rtol=1e-12, atol=1e-14)


def test_sampler_is_the_digit_comparison_it_is_defined_as():
    """DESIGN.md 2.5 from first principles, one row at a time and with no early exit: the 53
    digit words of the row's block give its uniform variate u (bit j of word i = digit i of row
    64 q + j); the row is selected iff u < thr.  The bit-sliced samplers must select exactly
    these rows (ragged block edges, thresholds with trailing zero digits, p close to 1)."""
    gamma, m64 = 0x9E3779B97F4A7C15, (1 << 64) - 1

    def by_definition(seed, it, frac, lo, hi):
        thr = int(math.floor(frac * 2.0 ** 53))
        key = fn.mix64((seed + it) & m64)
        out = []
        for r in range(lo, hi):
            q, j, u = r >> 6, r & 63, 0
            for i in range(1, 54):
                u = (u << 1) | ((fn.mix64((key + ((q << 6) + i - 1) * gamma) & m64) >> j) & 1)
            if u < thr:
                out.append(r)
        return np.asarray(out, dtype=np.int64)

    for frac in (0.5, 0.3, float(np.float32(1 / 45)), 0.75, float(np.float32(0.999)), 2.0 ** -20):
        want = by_definition(42, 3, frac, 100, 1100)
        assert np.array_equal(capi.sample_rows(42, 3, frac, 100, 1100), want)
        assert np.array_equal(fn.sample_rows(42, 3, frac, 100, 1100), want)


def test_sampler_statistics_are_bernoulli():
    n, p = 4_000_000, float(np.float32(1 / 45))
    a = capi.sample_rows(42, 1, p, 0, n)
    sd = math.sqrt(n * p * (1 - p))
    assert abs(len(a) - n * p) < 5 * sd                            # count
    cnt = np.bincount(a & 63, minlength=64)                         # every position of a block alike
    assert ((cnt - len(a) / 64) ** 2 / (len(a) / 64)).sum() < 130   # chi^2, 63 dof (p ~ 1e-6)
    gaps = np.diff(a)                                               # geometric gaps
    assert abs(gaps.mean() - 1 / p) < 0.5 and abs(gaps.std() - math.sqrt(1 - p) / p) < 0.5
    hit = np.zeros(n, dtype=np.int8)
    hit[a] = 1
    pair_sd = math.sqrt(n) * p
    for lag in (1, 7, 32, 64):                                      # no correlation between rows
        assert abs(int((hit[lag:] & hit[:-lag]).sum()) - n * p * p) < 6 * pair_sd
    other = np.zeros(n, dtype=np.int8)
    other[capi.sample_rows(42, 2, p, 0, n)] = 1                     # nor between iterations
    assert abs(int((hit & other).sum()) - n * p * p) < 6 * pair_sd


def test_sampler_c_vs_numpy_and_shard_union():
    for frac in (0.0, 0.01, 0.3, float(np.float32(0.1)), 1.0, 1.5):
        for it in (1, 2, 50):
            a = capi.sample_rows(42, it, frac, 0, 5000)
            b = fn.sample_rows(42, it, frac, 0, 5000)
            assert np.array_equal(a, b)
            # shards sample GLOBAL row numbers: the union over ranks is the single-node batch
            parts = [capi.sample_rows(42, it, frac, lo, hi) for lo, hi in ((0, 1234), (1234, 3000), (3000, 5000))]
            assert np.array_equal(np.concatenate(parts), a)
    assert len(capi.sample_rows(42, 1, 1.0, 0, 100)) == 100
    n = len(capi.sample_rows(42, 3, 0.25, 0, 200_000))
    assert abs(n / 200_000 - 0.25) < 0.005
    assert not np.array_equal(capi.sample_rows(42, 1, 0.5, 0, 1000), capi.sample_rows(42, 2, 0.5, 0, 1000))
    assert capi.lib().fmo_mix64(0) == fn.mix64(0) == 0xE220A8397B1DCDAF  # splitmix64 first output


def test_init_v_statistics_and_determinism():
    o = OracleFM(2000, 8)
    o.init_v(0.0, 0.01, 5)
    a = o.v.copy()
    o.init_v(0.0, 0.01, 5)
    assert np.array_equal(a, o.v)
    assert abs(a.mean()) < 3e-4 and abs(a.std() - 0.01) < 3e-4
    assert np.array_equal(a, a.astype(np.float32).astype(np.float64))  # fp32-representable
    o.init_v(0.0, 0.01, 6)
    assert not np.array_equal(a, o.v)


# ------------------------------------------------------------------------------ ALS (the reference's trainer)
def _als_problem(seed, n_rows=60, n_slots=23, k=3, mean_nnz=5, noise=0.05):
    rng = np.random.default_rng(seed)
    rp, idx, val = synth.ragged_rows(n_rows, n_slots, mean_nnz, seed=seed, values="normal")
    val = val.astype(np.float64)
    tw0, tw, tv = 0.3, rng.normal(0, 0.5, n_slots), rng.normal(0, 0.4, (n_slots, k))
    y = fn.predict(tw0, tw, tv, rp, idx, val) + rng.normal(0, noise, n_rows)
    model = (0.0, np.zeros(n_slots),
             rng.normal(0, 0.1, (n_slots, k)).astype(np.float32).astype(np.float64))
    return rp, idx, val, y, model


@pytest.mark.parametrize("quirks,f32", [(False, False), (True, False), (False, True), (True, True)])
def test_als_c_oracle_matches_scala_transliteration(quirks, f32):
    """fmo_als_sweep (arrays, CSC) against the dict-based transliteration of fm/lib/ALS.scala:15-75:
    same folds in the same order, so two sweeps agree to the last bit."""
    rp, idx, val, y, (w0, w, v) = _als_problem(11)
    n_slots, k = v.shape
    reg = (0.0, 0.01, 0.02)
    orc = OracleFM(n_slots, k, task=0, reg=reg)
    orc.set_model(w0, w, v)
    pw0, pw, pv = w0, w.copy(), v.copy()
    for _ in range(2):
        rmse, e = orc.als_sweep(rp, idx, val, y, ref_quirks=quirks, store_f32=f32)
        pw0, pw, pv, pe = fn.als_sweep(pw0, pw, pv, rp, idx, val, y, reg=reg, ref_quirks=quirks,
                                       store_f32=f32)
        assert orc.w0.value == pw0
        assert np.array_equal(orc.w, pw) and np.array_equal(orc.v, pv)
        assert np.array_equal(e, np.array(pe))
        assert abs(rmse - math.sqrt(np.mean(np.square(pe)))) < 1e-15
    if f32:   # every stored parameter is exactly representable in fp32
        assert np.array_equal(orc.v, orc.v.astype(np.float32).astype(np.float64))


def test_als_sweeps_fit_a_planted_model_and_keep_residuals_consistent():
    rp, idx, val, y, (w0, w, v) = _als_problem(5, n_rows=400, n_slots=40, k=4, mean_nnz=6)
    orc = OracleFM(40, 4, task=0, reg=(0.0, 1e-3, 1e-3))
    orc.set_model(w0, w, v)
    hist = []
    for _ in range(12):
        rmse, e = orc.als_sweep(rp, idx, val, y)
        hist.append(rmse)
        # the cached residuals ARE yhat - y of the updated model (what makes ALS O(nnz) per sweep)
        assert np.allclose(e, orc.predict(rp, idx, val) - y, rtol=0, atol=1e-9)
    assert hist[-1] < 0.5 * hist[0] and all(b <= a + 1e-9 for a, b in zip(hist, hist[1:]))


def test_als_reference_quirk_is_reproduced_on_request():
    rp, idx, val, y, (w0, w, v) = _als_problem(7)
    n_slots, k = v.shape
    last = n_slots - 1
    assert last in set(idx.tolist())
    a = OracleFM(n_slots, k, task=0, reg=(0.0, 0.01, 0.01))
    b = OracleFM(n_slots, k, task=0, reg=(0.0, 0.01, 0.01))
    a.set_model(w0, w, v)
    b.set_model(w0, w, v)
    _, ea = a.als_sweep(rp, idx, val, y)
    _, eb = b.als_sweep(rp, idx, val, y, ref_quirks=True)
    # (i) `0 until num_attribute` never trains the last slot (ALS.scala:38,52)
    assert b.w[last] == w[last] and np.array_equal(b.v[last], v[last])
    assert a.w[last] != w[last]
    # the residual cache is yhat - y of the model the sweep left behind in BOTH modes: the
    # reference's lazy `error` RDD is re-evaluated with the new w0 (ALS.scala:27,31,142-144)
    assert np.allclose(ea, a.predict(rp, idx, val) - y, atol=1e-9)
    assert np.allclose(eb, b.predict(rp, idx, val) - y, atol=1e-9)


def test_als_rejects_a_repeated_feature_in_a_row():
    orc = OracleFM(5, 2, task=0)
    with pytest.raises(ValueError):
        orc.als_sweep([0, 2], np.array([3, 3], np.int32), [1.0, 2.0], [0.5])


def test_als_update_is_the_coordinate_minimiser():
    """Independent of any transliteration: the closed form of ALS.scala:167-176 minimises
    sum_r (yhat_r - y_r)^2 + lambda * theta^2 over ONE coordinate.  After a sweep the last coordinate
    touched (last factor of the highest feature id) is therefore stationary:
    sum_r e_r * d yhat_r / d theta + lambda * theta = 0, with the derivative taken numerically."""
    rp, idx, val, y, (w0, w, v) = _als_problem(21, n_rows=200, n_slots=15, k=3, mean_nnz=4)
    n_slots, k = v.shape
    reg = (0.0, 0.05, 0.3)
    orc = OracleFM(n_slots, k, task=0, reg=reg)
    orc.set_model(w0, w, v)
    for _ in range(2):
        _, e = orc.als_sweep(rp, idx, val, y)
    last = int(idx.max())
    f = k - 1
    theta = orc.v[last, f]
    h = 1e-6
    vp, vm = orc.v.copy(), orc.v.copy()
    vp[last, f] += h
    vm[last, f] -= h
    dy = (fn.predict(orc.w0.value, orc.w, vp, rp, idx, val)
          - fn.predict(orc.w0.value, orc.w, vm, rp, idx, val)) / (2 * h)
    grad = float(np.dot(e, dy) + reg[2] * theta)
    scale = float(np.abs(e).dot(np.abs(dy)) + abs(reg[2] * theta))
    assert abs(grad) <= 1e-7 * scale
    # and a coordinate that was updated earlier in the sweep is generally NOT stationary any more
    dy0 = np.zeros(len(y))
    for r in range(len(y)):
        seg = slice(rp[r], rp[r + 1])
        dy0[r] = val[seg][idx[seg] == idx[0]].sum()
    assert abs(float(np.dot(e, dy0) + reg[1] * orc.w[idx[0]])) > 1e-9


def test_als_known_answers_exact_rational():
    """tests/golden/als_kat.json: one sweep computed in exact rational arithmetic from the formulas of
    fm/lib/ALS.scala:15-75 (tests/golden/make_als_kat.py, independent of oracle/); both fp64
    restatements must land on it up to rounding, with and without the reference quirks."""
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "als_kat.json")))
    rp = np.cumsum([0] + [len(r) for r in kat["rows"]]).astype(np.int64)
    idx = np.array([i for r in kat["rows"] for i, _ in r], np.int32)
    val = np.array([x for r in kat["rows"] for _, x in r], np.float64)
    y = np.array(kat["y"])
    for case in kat["cases"]:
        orc = OracleFM(kat["n_slots"], kat["k"], task=0, reg=kat["reg"])
        orc.set_model(kat["w0"], kat["w"], kat["v"])
        _, e = orc.als_sweep(rp, idx, val, y, ref_quirks=case["ref_quirks"])
        pw0, pw, pv, pe = fn.als_sweep(kat["w0"], kat["w"], kat["v"], rp, idx, val, y, reg=kat["reg"],
                                       ref_quirks=case["ref_quirks"])
        for got_w0, got_w, got_v, got_e in ((orc.w0.value, orc.w, orc.v, e), (pw0, pw, pv, pe)):
            assert abs(got_w0 - case["w0"]) <= 1e-12
            assert np.allclose(got_w, case["w"], rtol=1e-12, atol=1e-13)
            assert np.allclose(got_v, case["v"], rtol=1e-12, atol=1e-13)
            assert np.allclose(got_e, case["e"], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("task", [0, 1])
def test_cpu_baseline_variants_agree_with_the_oracle(task):
    """bench.py times two CPU variants: the faithful k-pass fp64 one must
    equal the single-pass oracle up to fp64 summation order, the tuned fp32 one within fp32
    round-off -- so both are baselines of the SAME computation."""
    from oracle.capi import OracleFast32
    n_slots, k, n_rows = 400, 6, 900
    row_ptr, idx, val = synth.ragged_rows(n_rows, n_slots, 9, seed=3, values="normal")
    rng = np.random.default_rng(5)
    label = rng.normal(0, 1, n_rows) if task == 0 else np.where(rng.random(n_rows) < 0.4, 1.0, -1.0)
    w, v = rng.normal(0, 0.1, n_slots), rng.normal(0, 0.1, (n_slots, k))
    reg = (0.01, 0.02, 0.03)
    a, b, c = (OracleFM(n_slots, k, task=task, reg=reg) for _ in range(3))
    for o in (a, b, c):
        o.set_model(0.05, w, v)
    fast = OracleFast32(c, threads=3)
    for it in (1, 2, 3):
        ids = np.sort(rng.choice(n_rows, 500, replace=False)).astype(np.int64)
        la = a.train_step(row_ptr, idx, val, label, ids, it, 0.2, threads=2)
        lb = b.train_step(row_ptr, idx, val, label, ids, it, 0.2, threads=3, faithful=True)
        lc = fast.train_step(row_ptr, idx, val.astype(np.float32), label.astype(np.float32), ids, it, 0.2)
        assert abs(la - lb) <= 1e-11 * abs(la)
        assert abs(la - lc) <= 2e-5 * abs(la)
    assert np.allclose(a.v, b.v, rtol=0, atol=1e-12) and np.allclose(a.w, b.w, rtol=0, atol=1e-12)
    w0c, wc, vc = fast.get_model()
    assert np.allclose(a.v, vc, rtol=0, atol=2e-6) and np.allclose(a.w, wc, rtol=0, atol=2e-6)
    assert abs(a.w0.value - w0c) <= 2e-6
    fast.close()
