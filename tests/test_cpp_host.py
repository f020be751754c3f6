"""The compiled-language host layer (sparkfm_b200/host/sparkfm.hpp: FM / FMModel / FMLearn / SGD /
DataSet mirroring the Scala) drives the same C ABI: it must build everywhere, fail loudly without
a GPU, and on a GPU reproduce the numbers of the Python mirror and the CPU oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import fm_numpy as fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_smoke.cpp")
LIBDIR = os.path.join(ROOT, "sparkfm_b200")


def _build(tmp_path):
    exe = str(tmp_path / "host_smoke")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-o", exe, SRC,
                           "-L" + LIBDIR, "-lsparkfm_b200", "-Wl,-rpath," + LIBDIR])
    return exe


def test_cpp_host_layer_builds_and_fails_loudly_without_gpu(tmp_path):
    exe = _build(tmp_path)
    from sparkfm_b200 import device_count
    if device_count() > 0:
        pytest.skip("a GPU is visible")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 2 and "CUDA error" in p.stdout   # no CPU fallback


def _dataset():
    n_rows, n_feat = 3000, 500
    rp, idx, val, lab = [0], [], [], []
    for r in range(n_rows):
        m = 3 + fn.mix64(r) % 9
        for j in range(m):
            idx.append(fn.mix64(r * 131 + j) % n_feat)
            val.append(1.0 + (fn.mix64(r * 977 + j) % 4) * 0.25)
        rp.append(len(idx))
        lab.append(1.0 if fn.mix64(r ^ 0xABCD) & 1 else -1.0)
    return (np.array(rp, np.int64), np.array(idx, np.int32), np.array(val, np.float64),
            np.array(lab, np.float64))


@pytest.mark.gpu
def test_cpp_host_layer_matches_python_mirror_and_oracle(tmp_path):
    from oracle import capi
    from oracle.capi import OracleFM
    from sparkfm_b200 import DataSet, FM, SGD, Task
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    rows = re.findall(r"iter (\d+) rmse_before (\S+) loss (\S+)", out)
    assert len(rows) == 5 and "index error status -5" in out
    rp, idx, val, lab = _dataset()
    ds = DataSet(lab, rp, idx, val, "py.train")
    assert f"dimension {ds.dimension} size {ds.size}" in out
    # same program through the Python mirror: identical bits (same library, same inputs)
    fm = FM(ds, 8, Task.Classification, 5)
    fm.seed = 7
    sgd = SGD.run(0.3, (0.0, 1e-4, 1e-3), 0.5)
    model = fm.learnWith(sgd)
    for (it, rmse, loss), prm, pl in zip(rows, fm.rmseHistory, sgd.lossHistory):
        assert float(rmse) == pytest.approx(prm, rel=1e-8) and float(loss) == pytest.approx(pl, rel=1e-8)
    # and the CPU oracle on the same batches: 1e-4 relative per iteration
    orc = OracleFM(ds.dimension + 1, 8, task=1, reg=(0.0, float(np.float32(1e-4)), float(np.float32(1e-3))))
    orc.init_v(0.0, 0.01, 7)
    for it in range(1, 6):
        ids = capi.sample_rows(42, it, 0.5, 0, ds.size)
        lo = orc.train_step(rp, idx, val, lab, ids, it, float(np.float32(0.3))) / len(ids)
        assert abs(float(rows[it - 1][2]) - lo) <= 1e-4 * lo
    p0 = float(re.search(r"predict0 (\S+)", out).group(1))
    want = orc.predict([0, 2], idx[:2], np.array([1.0, 2.0]))[0]
    assert abs(p0 - want) <= 1e-5 * max(abs(want), 0.1)
    del model
