"""The transposition sort written for this path (sfm_radix.cu: stable LSD radix sort in 10/11-bit
digits) against the library sort it replaces (SFM_SORT=cub).  Both are stable, so the sorted entry
lists -- and therefore every bit of the trained model and of the loss -- must be identical.  Covers
one, two and three digit passes, both payload widths (4-byte row for all-ones data, 8-byte
{row, x} otherwise), partial last tiles and batches smaller than a warp."""
import os

import numpy as np
import pytest

from sparkfm_b200 import Handle, synth

pytestmark = pytest.mark.gpu


def _case(kind, n_slots, k, n_rows, fields, seed):
    rng = np.random.default_rng(seed)
    if kind == "onehot":
        rp, idx, _, label = synth.ctr_csr(0, n_rows, fields, n_slots, seed)
        val = None
    else:
        rp, idx, val = synth.ragged_rows(n_rows, n_slots, fields, seed=seed, values="normal")
        label = np.where(rng.random(n_rows) < 0.4, 1.0, -1.0).astype(np.float32)
    model = (0.01, rng.normal(0, 0.05, n_slots).astype(np.float32),
             rng.normal(0, 0.05, (n_slots, k)).astype(np.float32))
    return rp, idx, val, label, model


def _run(sort, n_slots, k, data, frac, iters=3, mode=0):
    rp, idx, val, label, (w0, w, v) = data
    if sort:
        os.environ["SFM_SORT"] = sort
    else:
        os.environ.pop("SFM_SORT", None)
    try:
        hd = Handle(n_slots, k, task=1, reg=(0.0, 1e-4, 1e-4), step_size=0.2,
                    mini_batch_fraction=frac, sampler_seed=9, sampler_mode=mode)
        hd.set_model(w0, w, v)
        hd.load_dataset(rp, idx, val, label)
        losses = [hd.train_step(it) for it in range(1, iters + 1)]
        m = hd.get_model()
        hd.close()
    finally:
        os.environ.pop("SFM_SORT", None)
    return losses, m


@pytest.mark.parametrize("kind,n_slots,k,n_rows,fields,frac", [
    ("onehot", 1_000_000, 16, 30_000, 39, 1.0),     # 20 bits: 10 + 10, 143 tiles, partial last
    ("onehot", 50_000, 16, 40_000, 13, 0.5),        # 16 bits: 8 + 8
    ("onehot", 1_500, 8, 9_000, 7, 1.0),            # 11 bits: one pass
    ("onehot", 100, 4, 3, 5, 1.0),                  # 7 bits, 15 entries: less than a warp
    ("onehot", 6_000_000, 4, 20_000, 21, 1.0),      # 23 bits: 8 + 8 + 7
    ("ragged", 3_001, 5, 20_000, 11, 0.7),          # 12 bits, {row, x} payload
    ("ragged", 300_000, 8, 30_000, 30, 1.0),        # 19 bits: 10 + 9, {row, x} payload
])
def test_own_radix_sort_is_bit_identical_to_library_sort(kind, n_slots, k, n_rows, fields, frac):
    data = _case(kind, n_slots, k, n_rows, fields, seed=n_rows % 97)
    la, ma = _run(None, n_slots, k, data, frac)
    lb, mb = _run("cub", n_slots, k, data, frac)
    assert la == lb
    assert ma[0] == mb[0] and np.array_equal(ma[1], mb[1]) and np.array_equal(ma[2], mb[2])
    assert la[-1][0] < la[0][0]          # and it trains


def test_own_radix_sort_partition_cache_path():
    """The cached transposition of the PARTITION sampler is built by the same sort."""
    data = _case("onehot", 200_000, 16, 50_000, 20, seed=5)
    la, ma = _run(None, 200_000, 16, data, 0.25, iters=6, mode=1)
    lb, mb = _run("cub", 200_000, 16, data, 0.25, iters=6, mode=1)
    assert la == lb and np.array_equal(ma[2], mb[2]) and np.array_equal(ma[1], mb[1])
