"""Host-side mirror of the reference's operator interface for the hot path.

Same names, argument meaning and error behaviour as the Scala (paths relative to
/root/reference/src/main/scala/io/edstud/spark/):

    Task                     Task.scala:3-6
    SparseVector             breeze.linalg.SparseVector as the reference uses it (index, data, length)
    DataSet                  DataSet.scala:9-62         (size, dimension, inputs, targets, cache)
    Model / FMModel          Model.scala:9-31, fm/FMModel.scala:9-65
    FMLearn                  fm/FMLearn.scala:10-16     (the learner plugin boundary)
    FM / FactorizationMachines   fm/FM.scala:15-33, fm/impl/FactorizationMachines.scala:10-53
    FMUtils                  fm/FMUtils.scala:23-74     (LibFM text I/O)

and the SGD pieces BASELINE.json north_star names, which the reference does NOT have (SURVEY.md
F1) and which plug into that interface: `SGD extends FMLearn` and the `FMWithSGD.train` facade
(spark-libFM style: task / numIterations / stepSize / miniBatchFraction / dim / regParam /
initStd), `FMModel.save` / `FMModel.load`.

All numerics run on the GPU through the C ABI (handle.py); nothing here computes a prediction
or a gradient.  The Scala twin of this file (what a maintainer would add to the reference) is
scala/ and INTEGRATION.md.
"""
from __future__ import annotations

import logging
import math

import numpy as np

from . import handle as _h

log = logging.getLogger("io.edstud.spark")  # log4j.properties:2 of the reference


class Task:
    """Task.scala:3-6"""
    Regression = _h.REGRESSION
    Classification = _h.CLASSIFICATION


class SparseVector:
    """breeze SparseVector(index, data, length): stored entries in stored order; duplicates and
    explicit zeros are kept (activeIterator semantics, fm/FMModel.scala:45,58)."""

    def __init__(self, index, data, length):
        self.index = np.asarray(index, dtype=np.int32)
        self.data = np.asarray(data, dtype=np.float64)
        if self.index.shape != self.data.shape:
            raise ValueError("index and data differ in length")
        self.length = int(length)

    @property
    def used(self):
        return len(self.index)


class LabeledPoint:
    """org.apache.spark.mllib.regression.LabeledPoint(label, features) -- the input element type
    north_star names for FMWithSGD.train."""

    def __init__(self, label, features: SparseVector):
        self.label = float(label)
        self.features = features


class DataSet:
    """DataSet.scala:42-62 over CSR arrays instead of an RDD[(Double, SparseVector[Double])].

    Packing is bit-exact with respect to input order: row r's entries are
    idx[row_ptr[r]:row_ptr[r+1]] in stored order."""

    def __init__(self, labels, row_ptr, idx, val, name="dataset", length=None):
        self.labels = np.ascontiguousarray(labels, dtype=np.float64)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        self.idx = np.ascontiguousarray(idx, dtype=np.int32)
        self.val = None if val is None else np.ascontiguousarray(val, dtype=np.float64)
        if len(self.row_ptr) != len(self.labels) + 1:
            raise ValueError("row_ptr must have size + 1 entries")
        self.name = name
        self.length = length
        self._cached_on = None

    @classmethod
    def from_rows(cls, rows, name="dataset"):
        """rows: iterable of (label, SparseVector) or LabeledPoint -- the RDD element types."""
        labels, row_ptr, idx, val, length = [], [0], [], [], None
        for r in rows:
            lab, sv = (r.label, r.features) if isinstance(r, LabeledPoint) else r
            labels.append(float(lab))
            idx.append(sv.index)
            val.append(sv.data)
            row_ptr.append(row_ptr[-1] + sv.used)
            length = sv.length if length is None else max(length, sv.length)
        idx_a = np.concatenate(idx) if idx else np.zeros(0, np.int32)
        val_a = np.concatenate(val) if val else np.zeros(0, np.float64)
        return cls(labels, row_ptr, idx_a, val_a, name, length)

    # Features.isEmpty / size / dimension, DataSet.scala:11-29
    @property
    def isEmpty(self):
        return len(self.labels) == 0

    @property
    def size(self):
        return 0 if self.isEmpty else len(self.labels)

    @property
    def dimension(self):
        """max feature index over the rows (NOT max+1; DataSet.scala:27-29).  Like the Scala's
        `_.index.max`, a row without entries makes this fail."""
        if self.isEmpty:
            return 0
        if np.any(np.diff(self.row_ptr) == 0):
            raise ValueError("empty.max: a row has no features")  # UnsupportedOperationException
        return int(self.idx.max())

    @property
    def targets(self):
        return self.labels

    def inputs(self, r):
        b, e = int(self.row_ptr[r]), int(self.row_ptr[r + 1])
        length = self.length if self.length is not None else self.dimension + 1
        return SparseVector(self.idx[b:e],
                            np.ones(e - b) if self.val is None else self.val[b:e], length)

    def cache(self):
        return self  # the device copy is made by the model that trains / scores on it

    def unpersist(self):
        if self._cached_on is not None:
            self._cached_on._uncache(self)
        return self


class Model:
    """Model.scala:9-31."""

    def predict(self, features: SparseVector) -> float:
        raise NotImplementedError

    def computeRMSE(self, dataset: DataSet) -> float:
        raise NotImplementedError


class FMModel(Model):
    """fm/FMModel.scala:9-65 with the parameters resident on one GPU.

    `seed` IS used here (the reference ignores it, :14): V ~ N(init_mean, init_stdev^2) from the
    documented counter-based generator."""

    def __init__(self, num_attribute, num_factor, init_mean=0.0, init_stdev=0.01, seed=0,
                 task=Task.Regression, k0=True, k1=True, device=0, _handle=None):
        self.num_attribute = int(num_attribute)
        self.num_factor = int(num_factor)
        self.init_mean, self.init_stdev, self.seed = init_mean, init_stdev, seed
        self.k0, self.k1 = bool(k0), bool(k1)          # :25-26
        self.task = task
        self._reg = [0.0, 0.0, 10.0]                   # reg0, regw, regv defaults :29-31
        if _handle is not None:
            self._hd = _handle
        else:
            self._hd = _h.Handle(self.num_attribute + 1, self.num_factor, task=task, k0=k0, k1=k1,
                                 reg=tuple(self._reg), device=device)
            self._hd.init_model(init_mean, init_stdev, seed)
        self._resident = None

    # ---- parameters (:17-19); reads copy device -> host
    @property
    def w0(self):
        return self._hd.get_model()[0]

    @property
    def w(self):
        return self._hd.get_model()[1]

    @property
    def v(self):
        """[num_attribute+1][num_factor] -- Breeze's DenseMatrix(k, n+1) column-major memory."""
        return self._hd.get_model()[2]

    def set_parameters(self, w0, w, v):
        self._hd.set_model(w0, w, v)

    reg0 = property(lambda s: s._reg[0], lambda s, x: s._set_reg(0, x))
    regw = property(lambda s: s._reg[1], lambda s, x: s._set_reg(1, x))
    regv = property(lambda s: s._reg[2], lambda s, x: s._set_reg(2, x))

    def _set_reg(self, i, x):
        self._reg[i] = float(x)

    # ---- scorer
    def predict(self, features: SparseVector) -> float:
        """fm/FMModel.scala:34-55 for one vector (a 1-row batch on the device)."""
        return float(self.predict_batch(np.array([0, features.used], np.int64), features.index,
                                        features.data)[0])

    def predict_batch(self, row_ptr, idx, val):
        return self._hd.predict(row_ptr, idx, val)

    def _cache(self, dataset: DataSet):
        if self._resident is not dataset:
            self._hd.load_dataset(dataset.row_ptr, dataset.idx, dataset.val, dataset.labels)
            self._resident = dataset
            dataset._cached_on = self

    def _uncache(self, dataset):
        if self._resident is dataset:
            self._hd.unload_dataset()
            self._resident = None
            dataset._cached_on = None

    def predict_dataset(self, dataset: DataSet):
        """dataset.rdd.mapValues(predict) (Model.scala:14)."""
        self._cache(dataset)
        return self._hd.predict_resident(0, dataset.size)

    def computeRMSE(self, dataset: DataSet) -> float:
        """Model.scala:13-19."""
        self._cache(dataset)
        rmse = self._hd.evaluate()["rmse"]
        log.info("%s RMSE = %s", dataset.name, rmse)
        return rmse

    def computeMAE(self, dataset: DataSet) -> float:
        """Model.scala:21-26 -- the reference's 'MAE' is the mean SIGNED error (no abs); kept."""
        self._cache(dataset)
        return self._hd.evaluate()["mean_error"]

    def computeAccuracy(self, dataset: DataSet) -> float:
        """Model.scala:28-30 without its Long/Int integer division (which yields 0 or 1)."""
        self._cache(dataset)
        return self._hd.evaluate()["accuracy"]

    def computeAUC(self, dataset: DataSet) -> float:
        """Area under the ROC curve (not in the reference; SURVEY.md 8f 'next')."""
        self._cache(dataset)
        return self._hd.evaluate_auc()["auc"]

    # ---- persistence (north_star; absent from the reference)
    def save(self, path):
        self._hd.save(path)

    @classmethod
    def load(cls, path, device=0):
        hd = _h.Handle.load(path, device)
        cfg = hd.config()
        m = cls(cfg.n_slots - 1, cfg.k, task=cfg.task, k0=bool(cfg.k0), k1=bool(cfg.k1),
                device=device, _handle=hd)
        m._reg = [cfg.reg0, cfg.regw, cfg.regv]
        return m


class FMLearn:
    """fm/FMLearn.scala:10-16 -- the plugin boundary: called once per iteration by FM.learnWith,
    may mutate and return the same model (as ALS does, fm/lib/ALS.scala:27,40,64,74)."""

    def learn(self, fm: FMModel, dataset: DataSet) -> FMModel:
        raise NotImplementedError


class SGD(FMLearn):
    """Mini-batch SGD learner (FMGradient + FMUpdater of north_star).  One `learn`
    call is one iteration t = 1, 2, ...: sample a Bernoulli(miniBatchFraction) batch with seed
    `seed + t`, sum the gradient, step with eta = stepSize / sqrt(t) and L2 regParam
    (r0, r1, r2) on (w0, w, V)."""

    def __init__(self, stepSize=0.1, regParam=(0.0, 0.0, 0.0), miniBatchFraction=1.0, seed=42):
        self.stepSize = float(stepSize)
        self.regParam = tuple(float(x) for x in regParam)
        self.miniBatchFraction = float(miniBatchFraction)
        self.seed = int(seed)
        self.iteration = 0
        self.lossHistory = []

    @staticmethod
    def run(stepSize=0.1, regParam=(0.0, 0.0, 0.0), miniBatchFraction=1.0, seed=42):
        """Companion factory, like ALS.run() (fm/lib/ALS.scala:202-208)."""
        return SGD(stepSize, regParam, miniBatchFraction, seed)

    def learn(self, fm: FMModel, dataset: DataSet) -> FMModel:
        fm._cache(dataset)
        fm.reg0, fm.regw, fm.regv = self.regParam
        hd = fm._hd
        hd.set_hyper(self.regParam[0], self.regParam[1], self.regParam[2], self.stepSize,
                     self.miniBatchFraction)
        self.iteration += 1
        if self.seed != hd.config().sampler_seed:
            raise ValueError("sampler seed is fixed when the model handle is created")
        loss, batch = hd.train_step(self.iteration)
        self.lossHistory.append(loss)
        log.debug("SGD iteration %d: batch %d, mean loss %.6g", self.iteration, batch, loss)
        return fm


class ALS(FMLearn):
    """fm/lib/ALS.scala:11-200 -- the learner the reference ships: one `learn` call is one sweep of
    closed-form coordinate updates over w0, w and V for squared loss, using the model's own
    reg0 / regw / regv (FMModel.scala:29-31).  `refQuirks=True` reproduces the reference's one bug
    (the last slot is never trained, ALS.scala:38,52; see `sfm_als_sweep` in include/sparkfm_b200.h)."""

    def __init__(self, refQuirks=False):
        self.refQuirks = bool(refQuirks)
        self.rmseHistory = []

    @staticmethod
    def run(refQuirks=False):
        """ALS.run() (fm/lib/ALS.scala:202-208)."""
        return ALS(refQuirks)

    def learn(self, fm: FMModel, dataset: DataSet) -> FMModel:
        fm._cache(dataset)
        cfg = fm._hd.config()
        fm._hd.set_hyper(fm.reg0, fm.regw, fm.regv, cfg.step_size, cfg.mini_batch_fraction)
        self.rmseHistory.append(fm._hd.als_sweep(self.refQuirks))
        return fm


class FM:
    """fm/FM.scala:15-33."""

    def __new__(cls, dataset, numFactor, task=Task.Regression, maxIteration=100, timeout=0):
        return FactorizationMachines(dataset, numFactor, task, maxIteration, timeout)


class FactorizationMachines:
    """fm/impl/FactorizationMachines.scala:10-53."""

    def __init__(self, dataset, numFactor=8, task=Task.Regression, maxIteration=100, timeout=0,
                 device=0, seed=0, sampler_seed=42):
        self.dataset, self.numFactor, self.task = dataset, numFactor, task
        self.maxIteration, self.timeout = maxIteration, timeout
        self.device, self.seed, self.sampler_seed = device, seed, sampler_seed
        self.relations = []
        self.rmseHistory = []

    def withRelation(self, relation):
        # block-structure scaffolding is out of scope (unfinished upstream)
        raise NotImplementedError("relations (fm/bs/) are outside the SGD hot path")

    def learnWith(self, fml: FMLearn) -> FMModel:
        """:30-51 -- cache, build FMModel(dataset.dimension, numFactor), then per iteration:
        computeRMSE on the training set (logged), fm = fml.learn(fm, dataset)."""
        ds = self.dataset.cache()
        log.info("Initializing FM Model...")
        hd = _h.Handle(ds.dimension + 1, self.numFactor, task=self.task,
                       sampler_seed=getattr(fml, "seed", self.sampler_seed), device=self.device)
        fm = FMModel(ds.dimension, self.numFactor, task=self.task, seed=self.seed,
                     device=self.device, _handle=hd)
        hd.init_model(fm.init_mean, fm.init_stdev, self.seed)
        log.info("Starting Learning Process...")
        for i in range(1, self.maxIteration + 1):
            self.rmseHistory.append(fm.computeRMSE(ds))
            log.info("Iteration %d in progress...", i)
            fm = fml.learn(fm, ds)
        ds.unpersist()
        return fm


class FMWithSGD:
    """spark-libFM style facade named by north_star (not in the reference)."""

    @staticmethod
    def train(input, task=Task.Classification, numIterations=100, stepSize=0.1,
              miniBatchFraction=1.0, dim=(True, True, 8), regParam=(0.0, 0.0, 0.0), initStd=0.01,
              seed=0, sampler_seed=42, device=0, return_history=False):
        """input: DataSet or an iterable of LabeledPoint / (label, SparseVector).
        dim = (k0, k1, k).  Returns the trained FMModel (and the per-iteration mean loss)."""
        ds = input if isinstance(input, DataSet) else DataSet.from_rows(input)
        k0, k1, k = dim
        hd = _h.Handle(ds.dimension + 1, int(k), task=task, k0=k0, k1=k1, reg=regParam,
                       step_size=stepSize, mini_batch_fraction=miniBatchFraction,
                       sampler_seed=sampler_seed, device=device)
        hd.init_model(0.0, initStd, seed)
        fm = FMModel(ds.dimension, int(k), init_stdev=initStd, seed=seed, task=task, k0=k0, k1=k1,
                     device=device, _handle=hd)
        fm._reg = list(regParam)
        fm._cache(ds)
        hist = hd.train(1, int(numIterations))
        return (fm, hist) if return_history else fm


class FMUtils:
    """fm/FMUtils.scala:12-74 (text I/O only; Kryo registration is JVM plumbing)."""

    @staticmethod
    def loadLibFMFile(path, numFeatures=-1, name=None) -> DataSet:
        """:23-53.  Parsed by the library's C++ parser (sfm_parse_libfm)."""
        with open(path, "rb") as fh:
            text = fh.read()
        return FMUtils.parseLibFM(text, numFeatures, name or str(path))

    @staticmethod
    def parseLibFM(text: bytes, numFeatures=-1, name="libfm") -> DataSet:
        label, row_ptr, idx, val, d = _h.parse_libfm(text, numFeatures)
        return DataSet(label, row_ptr, idx, val, name, length=d + 1)

    @staticmethod
    def saveAsLibFMFile(data: DataSet, path):
        """:58-69 -- writes index+1 and rounds to 3 decimals, like the reference."""
        val = np.ones(len(data.idx)) if data.val is None else data.val
        with open(path, "wb") as fh:
            fh.write(_h.format_libfm(data.labels, data.row_ptr, data.idx, val))
