"""2-GPU data-parallel test (skipped with fewer than 2 GPUs): one process per GPU, rows block-
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


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # rendezvous only
    try:
        from sparkfm_b200 import Handle
        from sparkfm_b200.dist import init_comm, shard_range
        rp, idx, label, (w0, w, v) = _data()
        lo, hi = shard_range(N_ROWS, rank, world)
        hd = Handle(N_SLOTS, K, device=rank, **KW)
        if rank == 0:
            hd.set_model(w0, w, v)
        init_comm(hd, device="cpu")
        hd.comm_broadcast_model()
        sub = idx[rp[lo]:rp[hi]]
        hd.load_dataset(rp[lo:hi + 1] - rp[lo], sub, None, label[lo:hi], global_row_offset=lo)
        losses = [hd.train_step(it) for it in range(1, ITERS + 1)]
        m = hd.get_model()
        ev = hd.evaluate()
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), loss=np.array([l for l, _ in losses]),
                 batch=np.array([b for _, b in losses]), w0=m[0], w=m[1], v=m[2], rmse=ev["rmse"],
                 n=ev["n"])
        hd.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_match_one_gpu(tmp_path):
    import torch.multiprocessing as mp
    from sparkfm_b200 import Handle
    runs = []
    for rep in range(2):
        d = tmp_path / f"rep{rep}"
        d.mkdir()
        mp.spawn(_worker, args=(2, _free_port(), str(d)), nprocs=2, join=True)
        runs.append([np.load(d / f"r{r}.npz") for r in range(2)])
    a, b = runs[0]
    # replicas identical across ranks, and reproducible across reruns
    for key in ("loss", "batch", "w0", "w", "v"):
        assert np.array_equal(a[key], b[key]), key
        assert np.array_equal(a[key], runs[1][0][key]), key
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
