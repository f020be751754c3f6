"""The Python mirror of the reference's operator interface on the GPU: FMUtils.loadLibFMFile ->
DataSet -> FMWithSGD.train / FM(...).learnWith(SGD.run(...)) -> FMModel.predict / computeRMSE /
save / load (reference: fm/FMUtils.scala:23-53, fm/FM.scala:25-33,
fm/impl/FactorizationMachines.scala:30-51, fm/FMModel.scala:34-55, Model.scala:13-30)."""
import numpy as np
import pytest

from oracle.capi import OracleFM
from sparkfm_b200 import FM, FMModel, FMUtils, FMWithSGD, SGD, SparseVector, Task, synth

pytestmark = pytest.mark.gpu


def test_libfm_file_to_trained_model_roundtrip(tmp_path):
    row_ptr, idx, val, label = synth.classification_c1(4000, 2000, 15, 8, seed=11)
    path = tmp_path / "train.libfm"
    path.write_text(synth.to_libfm_text(row_ptr, idx, val, label))
    ds = FMUtils.loadLibFMFile(path)
    assert ds.size == 4000 and ds.dimension == int(idx.max())
    assert np.array_equal(ds.idx, idx) and np.array_equal(ds.row_ptr, row_ptr)   # packing bit-exact
    model, hist = FMWithSGD.train(ds, task=Task.Classification, numIterations=25, stepSize=1.0,
                                  miniBatchFraction=0.5, dim=(True, True, 8),
                                  regParam=(0.0, 1e-4, 1e-4), initStd=0.05, seed=3,
                                  return_history=True)
    assert hist[-1] < hist[0] * 0.99                       # it learns
    acc = model.computeAccuracy(ds)
    assert 0.5 < acc <= 1.0
    # single-vector predict (fm/FMModel.scala:34) == batched predict == oracle on the trained model
    sv = ds.inputs(5)
    p1 = model.predict(sv)
    pb = model.predict_dataset(ds)
    assert p1 == pytest.approx(float(pb[5]), rel=1e-6, abs=1e-7)
    orc = OracleFM(ds.dimension + 1, 8, task=1)
    orc.set_model(model.w0, model.w, model.v)
    want = orc.predict(ds.row_ptr, ds.idx, ds.val)
    assert np.max(np.abs(pb - want)) <= 1e-5 * max(np.mean(np.abs(want)), 1e-3) + 1e-5 * np.max(np.abs(want))
    rmse = model.computeRMSE(ds)
    assert rmse == pytest.approx(np.sqrt(np.mean((ds.labels - want) ** 2)), rel=1e-5)
    # save / load (north_star; absent upstream)
    f = tmp_path / "model.sfm"
    model.save(f)
    again = FMModel.load(f)
    assert again.num_attribute == model.num_attribute and again.num_factor == 8
    assert np.array_equal(again.v, model.v) and np.array_equal(again.w, model.w)
    assert again.predict(SparseVector(sv.index, sv.data, sv.length)) == p1
    # saveAsLibFMFile: index + 1 and 3-decimal rounding, like the reference (:58-74)
    out = tmp_path / "export.libfm"
    FMUtils.saveAsLibFMFile(ds, out)
    first = out.read_text().split("\n")[0].split(" ")
    assert first[1].split(":")[0] == str(int(idx[0]) + 1)


def test_learnwith_loop_regression_rmse_goes_down():
    row_ptr, idx, val, y = synth.regression_c2(6000, 3000, 20, 8, seed=5)
    from sparkfm_b200 import DataSet
    ds = DataSet(y, row_ptr, idx, val, "c2.small")
    fm = FM(ds, 8, Task.Regression, maxIteration=30)
    model = fm.learnWith(SGD.run(stepSize=0.05, regParam=(0.0, 1e-3, 1e-3), miniBatchFraction=0.3))
    assert len(fm.rmseHistory) == 30 and fm.rmseHistory[-1] < fm.rmseHistory[0]
    assert model.computeRMSE(ds) < fm.rmseHistory[0]
