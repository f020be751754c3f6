"""BASELINE.json configs 1, 2, 3 and 4 at their FULL sizes on one GPU (scripts/run_configs.py with
fewer iterations, so that the driver's GPU test run carries them): bit-exact LibFM ingest + packing
(C1), per-iteration loss against the fp64 CPU oracle at 1e-4 relative (C1 full batch, C2 sampled
mini-batches), predictions at 1e-5, and for the 10 M-feature k = 64 model of C4 -- too large for the
oracle in seconds -- size-independent properties: the loss goes down and a rerun from the same
state reproduces the same bits.  Config 3 (the headline, 45 M rows): the oracle checks the sampled
mini-batches, rebuilt on the host from their global row numbers.  (Config 3's throughput is
bench.py; config 5 is bench.py under torchrun.)"""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mod():
    spec = importlib.util.spec_from_file_location("run_configs", os.path.join(ROOT, "scripts", "run_configs.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_config1_full_size_libsvm_classification():
    m = _mod()
    m.config1(n_check=4, n_timed=5)
    r = m.OUT["C1"]
    assert r["rows"] == 100_000 and r["k"] == 8
    assert r["loss_rel_err_max"] < 1e-4 and r["predict_rel_err_max"] < 1e-5
    assert r["loss_first_last"][1] < r["loss_first_last"][0]


def test_config2_full_size_regression_minibatch():
    m = _mod()
    m.config2(n_check=3, n_timed=5)
    r = m.OUT["C2"]
    assert r["rows"] == 1_000_000 and r["n_slots"] == 100_000 and r["k"] == 16
    assert r["loss_rel_err_max"] < 1e-4


def test_config3_full_size_criteo_shape():
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs 45 M rows (7.5 GB) resident twice over")
    m = _mod()
    m.config3(n_check=3)
    r = m.OUT["C3"]
    assert r["rows"] == 45_000_000 and r["n_slots"] == 1_000_000 and r["k"] == 16
    assert all(900_000 < b < 1_100_000 for b in r["batch_rows"])
    assert r["loss_rel_err_max"] < 1e-4 and r["predict_rel_err_max"] < 1e-5
    assert r["bitwise_rerun"] is True


def test_config4_full_size_avazu_shape_replicated():
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs a 10 M x 64 model + 8 M rows resident")
    m = _mod()
    m.config4(n_timed=3, modes=((0, "bernoulli"),))
    r = m.OUT["C4_replicated_1gpu"]
    assert r["bitwise_rerun"] is True
    assert r["bernoulli"]["loss_first_last"][1] < r["bernoulli"]["loss_first_last"][0]
