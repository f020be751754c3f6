"""The transposition of the SGD step, three implementations against each other:

  * default: bucket form (sfm_bucket.cu) -- one global stable partition by the top bits of the
    feature id, the rest of the sort done per tile in shared memory inside the reduce kernel;
  * SFM_BUCKET=0: the two-pass wide-digit radix sort (sfm_radix.cu) + chunked reduce;
  * SFM_SORT=cub: the library radix sort + chunked reduce.

The last two are both stable sorts feeding the same reduce, so every bit of the trained model and
of the loss must be identical.  The bucket form adds the same terms in a different (but fixed)
tree shape: it must agree within fp32 round-off and be bitwise reproducible from run to run.
Covers one, two and three digit passes, bucket counts from 2 to 2048, both payload kinds (row for
all-ones data, {row, x} otherwise), partial tiles, multi-item buckets and batches smaller than a
warp."""
import os

import numpy as np
import pytest

from sparkfm_b200 import Handle, synth

pytestmark = pytest.mark.gpu

KNOBS = ("SFM_SORT", "SFM_BUCKET", "SFM_BUCKET_CACHE")


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


def _run(env, n_slots, k, data, frac, iters=3, mode=0):
    """env: dict of knob -> value; the library reads the knobs when the handle is created."""
    rp, idx, val, label, (w0, w, v) = data
    for kn in KNOBS:
        os.environ.pop(kn, None)
    os.environ.update(env)
    try:
        hd = Handle(n_slots, k, task=1, reg=(0.0, 1e-4, 1e-4), step_size=0.2,
                    mini_batch_fraction=frac, sampler_seed=9, sampler_mode=mode)
        hd.set_model(w0, w, v)
        hd.load_dataset(rp, idx, val, label)
        losses = [hd.train_step(it) for it in range(1, iters + 1)]
        m = hd.get_model()
        hd.close()
    finally:
        for kn in KNOBS:
            os.environ.pop(kn, None)
    return losses, m


CASES = [
    ("onehot", 1_000_000, 16, 30_000, 39, 1.0),     # 20 bits: 2048 buckets x 512 features
    ("onehot", 50_000, 16, 40_000, 13, 0.5),        # 16 bits: 128 buckets
    ("onehot", 1_500, 8, 9_000, 7, 1.0),            # 11 bits: 4 buckets, multi-item buckets
    ("onehot", 100, 4, 3, 5, 1.0),                  # 7 bits, 15 entries: less than a warp
    ("onehot", 6_000_000, 4, 20_000, 21, 1.0),      # 23 bits: too many buckets -> sorted path
    ("ragged", 3_001, 5, 20_000, 11, 0.7),          # 12 bits, {row, x} payload
    ("ragged", 300_000, 8, 30_000, 30, 1.0),        # 19 bits, {row, x} payload
    ("onehot", 2, 1, 5_000, 3, 1.0),                # 1 bit: two buckets of one feature
    ("ragged", 40_000, 64, 8_000, 12, 1.0),         # k = 64: 128 features per bucket
    ("onehot", 700, 16, 120_000, 20, 1.0),          # 2 buckets x 2.4 M entries: 37 items each
]


@pytest.mark.parametrize("kind,n_slots,k,n_rows,fields,frac", CASES[:7])
def test_own_radix_sort_is_bit_identical_to_library_sort(kind, n_slots, k, n_rows, fields, frac):
    data = _case(kind, n_slots, k, n_rows, fields, seed=n_rows % 97)
    la, ma = _run({"SFM_BUCKET": "0"}, n_slots, k, data, frac)
    lb, mb = _run({"SFM_SORT": "cub"}, n_slots, k, data, frac)
    assert la == lb
    assert ma[0] == mb[0] and np.array_equal(ma[1], mb[1]) and np.array_equal(ma[2], mb[2])
    assert la[-1][0] < la[0][0]          # and it trains


@pytest.mark.parametrize("kind,n_slots,k,n_rows,fields,frac", CASES)
def test_bucket_form_matches_sorted_form_and_is_reproducible(kind, n_slots, k, n_rows, fields, frac):
    data = _case(kind, n_slots, k, n_rows, fields, seed=n_rows % 97)
    la, ma = _run({}, n_slots, k, data, frac)
    lb, mb = _run({"SFM_SORT": "cub"}, n_slots, k, data, frac)
    lc, mc = _run({}, n_slots, k, data, frac)
    # bitwise reproducible
    assert la == lc
    assert ma[0] == mc[0] and np.array_equal(ma[1], mc[1]) and np.array_equal(ma[2], mc[2])
    # same sums in a different tree shape: fp32 round-off only
    assert [b for _, b in la] == [b for _, b in lb]
    for (x, _), (y, _) in zip(la, lb):
        assert abs(x - y) <= 2e-6 * abs(y)
    scale_w = max(float(np.max(np.abs(mb[1]))), 1e-30)
    scale_v = max(float(np.max(np.abs(mb[2]))), 1e-30)
    assert abs(ma[0] - mb[0]) <= 1e-5 * max(abs(mb[0]), 1e-3)
    assert float(np.max(np.abs(ma[1] - mb[1]))) <= 2e-5 * scale_w
    assert float(np.max(np.abs(ma[2] - mb[2]))) <= 2e-5 * scale_v
    assert la[-1][0] < la[0][0]


@pytest.mark.parametrize("env", [{"SFM_BUCKET_CACHE": "1"}, {}])
def test_partition_cache_path(env):
    """The cached transposition of the PARTITION sampler: bucket form (entries grouped by bucket
    once, ranked per tile every step) and sorted form, against the library sort."""
    data = _case("onehot", 200_000, 16, 50_000, 20, seed=5)
    la, ma = _run(env, 200_000, 16, data, 0.25, iters=6, mode=1)
    lb, mb = _run({"SFM_SORT": "cub"}, 200_000, 16, data, 0.25, iters=6, mode=1)
    if not env:   # default: cached batches keep the fully sorted form (own radix sort)
        assert la == lb and np.array_equal(ma[2], mb[2]) and np.array_equal(ma[1], mb[1])
    else:
        for (x, bx), (y, by) in zip(la, lb):
            assert bx == by and abs(x - y) <= 2e-6 * abs(y)
        assert float(np.max(np.abs(ma[2] - mb[2]))) <= 2e-5 * float(np.max(np.abs(mb[2])))
