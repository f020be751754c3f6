"""ALS on the GPU (sfm_als_sweep) against the fp64 restatement of fm/lib/ALS.scala:15-75
(oracle.fmo_als_sweep with fp32 parameter storage, which is what the device model does).  The
device runs the reference's sequential coordinate sweep level by level (columns of a level share
no row), so parameters, residuals and RMSE must agree up to fp64 summation order: parameters to
1e-5 relative, RMSE to 1e-6."""
import numpy as np
import pytest

from oracle.capi import OracleFM
from sparkfm_b200 import ALS, DataSet, FM, Handle, synth
from sparkfm_b200._lib import SfmError

pytestmark = pytest.mark.gpu


def _problem(kind, n_rows, n_slots, k, fields, seed):
    rng = np.random.default_rng(seed)
    if kind == "binary":   # all-ones data, no value array: the 4-byte-payload kernels
        rp, idx, _ = synth.ragged_rows(n_rows, n_slots, fields, seed=seed, values="ones")
        val = None
        dval = np.ones(len(idx))
    else:
        rp, idx, val = synth.ragged_rows(n_rows, n_slots, fields, seed=seed, values="normal")
        dval = val.astype(np.float64)
    tw, tv = rng.normal(0, 0.3, n_slots), rng.normal(0, 0.3, (n_slots, k))
    oracle_gen = OracleFM(n_slots, k)
    oracle_gen.set_model(0.2, tw, tv)
    y = (oracle_gen.predict(rp, idx, dval) + rng.normal(0, 0.05, n_rows)).astype(np.float32)
    model = (0.0, np.zeros(n_slots, np.float32), rng.normal(0, 0.1, (n_slots, k)).astype(np.float32))
    return rp, idx, val, dval, y, model


@pytest.mark.parametrize("kind,n_rows,n_slots,k,fields,quirks", [
    ("ragged", 3000, 300, 4, 6, False),
    ("ragged", 3000, 300, 4, 6, True),
    ("binary", 20000, 5000, 8, 9, False),
    ("ragged", 40, 12, 3, 3, False),
])
def test_als_sweeps_match_the_scala_restatement(kind, n_rows, n_slots, k, fields, quirks):
    rp, idx, val, dval, y, (w0, w, v) = _problem(kind, n_rows, n_slots, k, fields, seed=n_rows % 89)
    reg = (0.0, 0.01, 0.05)
    orc = OracleFM(n_slots, k, task=0, reg=reg)
    orc.set_model(w0, w, v)
    hd = Handle(n_slots, k, task=0, reg=reg)
    hd.set_model(w0, w, v)
    hd.load_dataset(rp, idx, val, y)
    hist = []
    for _ in range(3):
        want, e_want = orc.als_sweep(rp, idx, dval, y.astype(np.float64), ref_quirks=quirks,
                                     store_f32=True)
        got = hd.als_sweep(ref_quirks=quirks)
        hist.append(got)
        assert abs(got - want) <= 1e-6 * want
        gw0, gw, gv = hd.get_model()
        assert abs(gw0 - orc.w0.value) <= 1e-5 * max(abs(orc.w0.value), 1e-3)
        assert np.max(np.abs(gw - orc.w)) <= 1e-5 * np.abs(orc.w).max()
        assert np.max(np.abs(gv - orc.v)) <= 1e-5 * np.abs(orc.v).max()
        e = hd.als_residuals(n_rows)
        assert np.max(np.abs(e - e_want)) <= 1e-5 * max(np.abs(e_want).max(), 1e-3)
    assert hist[-1] < hist[0]
    # the residual cache is yhat - y of the model the sweep left behind, quirk mode or not
    pred = hd.predict_resident(0, n_rows)
    assert np.max(np.abs(e - (pred.astype(np.float64) - y))) < 1e-4
    hd.close()


def test_als_learner_plugs_into_learnwith_and_is_reproducible():
    """FM(dataset, k).learnWith(ALS.run()) -- the reference's only runnable scenario
    (driver.scala:105-110) -- on a small planted problem; reruns give the same bits."""
    rp, idx, val, dval, y, _ = _problem("ragged", 2000, 120, 4, 5, seed=3)
    ds = DataSet(y, rp, idx, val)
    out = []
    for _ in range(2):
        als = ALS.run()
        fm = FM(ds, 4, maxIteration=4).learnWith(als)
        out.append((als.rmseHistory, fm.w0, fm.w.copy(), fm.v.copy()))
    assert out[0][0] == out[1][0] and np.array_equal(out[0][3], out[1][3])
    assert out[0][0][-1] < out[0][0][0]


def test_als_rejects_duplicate_feature_in_a_row():
    hd = Handle(8, 2, task=0)
    hd.load_dataset([0, 2, 3], np.array([3, 3, 1], np.int32), [1.0, 2.0, 1.0], [0.5, 1.0])
    with pytest.raises(SfmError):
        hd.als_sweep()
    hd.close()
