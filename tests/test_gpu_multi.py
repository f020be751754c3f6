"""2/4/8-GPU data-parallel tests (skipped with fewer GPUs): one process per GPU, rows block-
sharded, gradient all-reduced by the library over NCCL (sfm_comm_init).  Checks (SURVEY.md
section 4 item 4): per-iteration loss on 2 GPUs == 1 GPU within the 1e-4 tolerance, replicas stay
bitwise identical across ranks, and a rerun at the same GPU count reproduces the same bits."""
import os
import socket

import numpy as np
import pytest

from sparkfm_b200 import device_count

pytestmark = pytest.mark.gpu

N_SLOTS, K, N_ROWS, FIELDS, ITERS = 30_000, 16, 40_000, 13, 6
KW = dict(task=1, reg=(0.0, 1e-4, 1e-4), step_size=0.3, mini_batch_fraction=0.25, sampler_seed=42)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    from sparkfm_b200 import synth
    rp, idx, _, label = synth.ctr_csr(0, N_ROWS, FIELDS, N_SLOTS, 77)
    rng = np.random.default_rng(3)
    return rp, idx, label, (0.02, rng.normal(0, 0.05, N_SLOTS).astype(np.float32),
                            rng.normal(0, 0.05, (N_SLOTS, K)).astype(np.float32))


def _worker(rank, world, port, out_dir, p2p="1", sparse="1", bad_rank=-1):
    import torch.distributed as dist
    os.environ["SFM_P2P"] = p2p
    os.environ["SFM_P2P_SPARSE"] = sparse
    os.environ["SFM_P2P_TIMEOUT_S"] = "20"
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # rendezvous only
    try:
        from sparkfm_b200 import Handle
        from sparkfm_b200._lib import SfmError
        from sparkfm_b200.dist import init_comm, shard_range
        rp, idx, label, (w0, w, v) = _data()
        lo, hi = shard_range(N_ROWS, rank, world)
        hd = Handle(N_SLOTS, K, device=rank, **KW)
        if rank == 0:
            hd.set_model(w0, w, v)
        init_comm(hd, device="cpu")
        hd.comm_broadcast_model()
        mode = hd.comm_mode()
        assert mode == ("nccl" if p2p == "0" else mode)      # SFM_P2P=0 must force NCCL
        sub = idx[rp[lo]:rp[hi]]
        hd.load_dataset(rp[lo:hi + 1] - rp[lo], sub, None, label[lo:hi], global_row_offset=lo)
        losses = [hd.train_step(it) for it in range(1, ITERS + 1)]
        m = hd.get_model()
        bad_seen = -1
        if bad_rank >= 0:
            # one rank feeds a host batch with an out-of-range index: EVERY rank must report the
            # step as failed and leave its replica untouched (no silent divergence)
            brp = np.arange(0, 5 * FIELDS + 1, FIELDS, dtype=np.int64)
            bidx = sub[:5 * FIELDS].copy()
            if rank == bad_rank:
                bidx[7] = N_SLOTS + 3
            try:
                hd.train_step_csr(ITERS + 1, brp, bidx, None, label[lo:lo + 5])
                bad_seen = 0
            except SfmError:
                bad_seen = 1
            m2 = hd.get_model()
            assert m2[0] == m[0] and np.array_equal(m2[1], m[1]) and np.array_equal(m2[2], m[2])
            # and the job carries on afterwards
            losses.append(hd.train_step(ITERS + 2))
            m = hd.get_model()
        ev = hd.evaluate()
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), loss=np.array([l for l, _ in losses]),
                 batch=np.array([b for _, b in losses]), w0=m[0], w=m[1], v=m[2], rmse=ev["rmse"],
                 n=ev["n"], mode=mode, bad_seen=bad_seen)
        hd.close()
    finally:
        dist.destroy_process_group()


def _need(world):
    if device_count() < world:
        pytest.skip(f"needs {world} GPUs")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_n_gpus_match_one_gpu(tmp_path, world):
    """Runs 0 and 1: default gradient exchange (sparse sum + update kernel over NVLink peer memory
    when the GPUs can map each other, see `mode`); run 2: the dense peer-memory kernel
    (SFM_P2P_SPARSE=0); run 3: SFM_P2P=0, the NCCL all-reduce path.  Both peer-memory kernels add the
    ranks' rows in rank order, so runs 0-2 agree bit for bit at every world size; NCCL's order is
    its own, so run 3 is bitwise only at two ranks (a + b) and within 1e-6 otherwise."""
    _need(world)
    import torch.multiprocessing as mp
    from sparkfm_b200 import Handle
    runs = []
    for rep, (p2p, sparse) in enumerate([("1", "1"), ("1", "1"), ("1", "0"), ("0", "1")]):
        d = tmp_path / f"rep{rep}"
        d.mkdir()
        mp.spawn(_worker, args=(world, _free_port(), str(d), p2p, sparse), nprocs=world, join=True)
        runs.append([np.load(d / f"r{r}.npz") for r in range(world)])
    a = runs[0][0]
    print("gradient exchange mode:", str(a["mode"]))
    assert str(runs[3][0]["mode"]) == "nccl"
    for key in ("loss", "batch", "w0", "w", "v"):
        for r in range(1, world):                                  # replicas identical across ranks
            assert np.array_equal(a[key], runs[0][r][key]), (key, r)
            assert np.array_equal(runs[3][0][key], runs[3][r][key]), (key, r)
        assert np.array_equal(a[key], runs[1][0][key]), key         # reproducible across reruns
        assert np.array_equal(a[key], runs[2][0][key]), key         # sparse == dense peer-memory kernel
        if world == 2:
            assert np.array_equal(a[key], runs[3][0][key]), key     # == NCCL path
        else:
            assert np.max(np.abs(a[key] - runs[3][0][key])) <= 1e-6, key
    assert a["n"] == N_ROWS
    # same trajectory as one GPU holding every row
    rp, idx, label, (w0, w, v) = _data()
    hd = Handle(N_SLOTS, K, device=0, **KW)
    hd.set_model(w0, w, v)
    hd.load_dataset(rp, idx, None, label)
    one = [hd.train_step(it) for it in range(1, ITERS + 1)]
    assert [bt for _, bt in one] == a["batch"].tolist()      # global batch = union of shard batches
    for (l1, _), l2 in zip(one, a["loss"]):
        assert abs(l1 - l2) <= 1e-4 * abs(l1)
    m1 = hd.get_model()
    assert np.max(np.abs(m1[2] - a["v"])) <= 1e-4 * np.abs(m1[2]).max()
    assert abs(hd.evaluate()["rmse"] - float(a["rmse"])) < 1e-5
    hd.close()


@pytest.mark.parametrize("p2p,sparse", [("1", "1"), ("1", "0"), ("0", "1")])
def test_bad_index_on_one_rank_stops_every_rank(tmp_path, p2p, sparse):
    """A feature index outside [0, n_slots) seen by ONE rank: the error travels with the summed
    scalars, every rank skips that update and reports SFM_ERR_INDEX; replicas stay identical."""
    _need(2)
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), p2p, sparse, 1), nprocs=2, join=True)
    a, b = [np.load(tmp_path / f"r{r}.npz") for r in range(2)]
    assert int(a["bad_seen"]) == 1 and int(b["bad_seen"]) == 1
    for key in ("loss", "w0", "w", "v"):
        assert np.array_equal(a[key], b[key]), key


# ------------------------------------------------------------------------------ row-sharded V
def _shard_data(kind):
    from sparkfm_b200 import synth
    rng = np.random.default_rng(17)
    if kind == "onehot":          # uniform all-ones rows, k = 16: forward fast path on the compact model
        n_slots, k, n_rows = 20_001, 16, 24_000
        rp, idx, _, label = synth.ctr_csr(0, n_rows, 13, n_slots, 5)
        val = None
    else:                          # ragged valued rows, k = 5 (padded): generic kernels
        n_slots, k, n_rows = 3_001, 5, 9_000
        rp, idx, val = synth.ragged_rows(n_rows, n_slots, 11, seed=3, values="normal")
        label = np.where(rng.random(n_rows) < 0.4, 1.0, -1.0).astype(np.float32)
    model = (0.01, rng.normal(0, 0.05, n_slots).astype(np.float32),
             rng.normal(0, 0.05, (n_slots, k)).astype(np.float32))
    return n_slots, k, n_rows, rp, idx, val, label, model


def _shard_worker(rank, world, port, out_dir, kind, mode=0):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sparkfm_b200 import Handle
        from sparkfm_b200._lib import SfmError
        from sparkfm_b200.dist import init_comm, shard_range
        n_slots, k, n_rows, rp, idx, val, label, (w0, w, v) = _shard_data(kind)
        lo, hi = shard_range(n_rows, rank, world)
        hd = Handle(n_slots, k, device=rank, shard_v=True, sampler_mode=mode, **KW)
        try:
            hd.set_model(w0, w, v)
            raise AssertionError("model calls must fail before sfm_comm_init")
        except SfmError:
            pass
        init_comm(hd, device="cpu")
        hd.set_model(w0, w, v)                       # every rank keeps its own rows
        g = hd.get_model()                           # collective all-gather
        assert g[0] == np.float32(w0) and np.array_equal(g[1], w) and np.array_equal(g[2], v)
        sub_val = None if val is None else val[rp[lo]:rp[hi]]
        hd.load_dataset(rp[lo:hi + 1] - rp[lo], idx[rp[lo]:rp[hi]], sub_val, label[lo:hi],
                        global_row_offset=lo)
        pred0 = hd.predict_resident(0, hi - lo)
        losses = [hd.train_step(it) for it in range(1, ITERS + 1)]
        m = hd.get_model()
        ev = hd.evaluate()
        np.savez(os.path.join(out_dir, f"s{rank}.npz"), loss=np.array([l for l, _ in losses]),
                 batch=np.array([b for _, b in losses]), w0=m[0], w=m[1], v=m[2], rmse=ev["rmse"],
                 pred0=pred0, lo=lo, hi=hi)
        hd.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kind,mode", [("onehot", 0), ("ragged", 0), ("onehot", 1), ("ragged", 1)])
def test_row_sharded_model_matches_replicated(tmp_path, kind, mode):
    """SFM_FLAG_SHARD_V on 2 GPUs (each owns half of V / w, touched rows travel by all-to-all)
    against ONE GPU holding the whole model and every row: same predictions, same per-iteration
    loss (1e-4), same final model; reruns are bitwise identical.  mode 1 = PARTITION sampler: the
    row-dependent exchange plan of each fixed mini-batch is cached (ITERS > P, so it is reused)."""
    import torch.multiprocessing as mp
    from sparkfm_b200 import Handle
    runs = []
    for rep in range(2):
        d = tmp_path / f"rep{rep}"
        d.mkdir()
        mp.spawn(_shard_worker, args=(2, _free_port(), str(d), kind, mode), nprocs=2, join=True)
        runs.append([np.load(d / f"s{r}.npz") for r in range(2)])
    a, b = runs[0]
    for key in ("loss", "batch", "w0", "w", "v"):
        assert np.array_equal(a[key], b[key]), key              # get_model is the same on both ranks
        assert np.array_equal(a[key], runs[1][0][key]), key      # and reproducible
    n_slots, k, n_rows, rp, idx, val, label, (w0, w, v) = _shard_data(kind)
    hd = Handle(n_slots, k, device=0, sampler_mode=mode, **KW)
    hd.set_model(w0, w, v)
    hd.load_dataset(rp, idx, val, label)
    p1 = hd.predict_resident(0, n_rows)
    p2 = np.concatenate([a["pred0"], b["pred0"]])
    assert np.max(np.abs(p1 - p2)) <= 1e-6 * max(np.abs(p1).max(), 1e-3)
    one = [hd.train_step(it) for it in range(1, ITERS + 1)]
    assert [bt for _, bt in one] == a["batch"].tolist()
    for (l1, _), l2 in zip(one, a["loss"]):
        assert abs(l1 - l2) <= 1e-4 * abs(l1)
    m1 = hd.get_model()
    assert np.max(np.abs(m1[2] - a["v"])) <= 1e-4 * np.abs(m1[2]).max()
    assert np.max(np.abs(m1[1] - a["w"])) <= 1e-4 * max(np.abs(m1[1]).max(), 1e-3)
    assert abs(hd.evaluate()["rmse"] - float(a["rmse"])) < 1e-5
    hd.close()
