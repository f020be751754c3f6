"""World-size-2 CPU test (gloo) of the N>1 host logic: row sharding, the unique-id broadcast
plumbing that sfm_comm_init relies on, the sampler's shard-union property, and the data-parallel
identity the NCCL all-reduce implements (sum over ranks of per-shard gradients == the global
gradient; same update on every rank).  The arithmetic here is the CPU oracle's -- the CUDA path
needs GPUs and is covered by tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import capi
        from oracle.capi import OracleFM
        from sparkfm_b200 import synth
        from sparkfm_b200.dist import broadcast_bytes, shard_range

        # 1. the 128-byte id travels from rank 0 to everyone
        payload = bytes(range(128)) if rank == 0 else b""
        got = broadcast_bytes(payload, 128, 0, "cpu")
        assert got == bytes(range(128))

        # 2. block sharding covers the rows exactly once
        n_rows, n_slots, k = 3001, 400, 4
        lo, hi = shard_range(n_rows, rank, world)
        spans = [None] * world
        dist.all_gather_object(spans, (lo, hi))
        assert spans[0][0] == 0 and spans[-1][1] == n_rows
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))

        # 3. every rank samples GLOBAL row numbers of its own shard: the union is the 1-GPU batch
        row_ptr, idx, val = synth.ragged_rows(n_rows, n_slots, 8, seed=5, values="normal")
        rng = np.random.default_rng(1)
        label = rng.normal(0, 1, n_rows)
        w0, w, v = 0.1, rng.normal(0, 0.1, n_slots), rng.normal(0, 0.1, (n_slots, k))
        it, frac, step = 3, 0.3, 0.05
        mine = capi.sample_rows(42, it, frac, lo, hi)
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        whole = capi.sample_rows(42, it, frac, 0, n_rows)
        assert np.array_equal(np.concatenate(parts), whole)

        # 4. all-reduce(sum) of the per-shard gradients and losses == the global gradient
        orc = OracleFM(n_slots, k, task=0, reg=(0.0, 0.01, 0.02))
        orc.set_model(w0, w, v)
        gv, gw, gw0, loss = orc.gradient(row_ptr, idx, val.astype(np.float64), label, mine)
        buf = torch.from_numpy(np.concatenate([gv.reshape(-1), gw, [gw0, loss, float(len(mine))]]))
        dist.all_reduce(buf)
        tot = buf.numpy()
        ref = OracleFM(n_slots, k, task=0, reg=(0.0, 0.01, 0.02))
        ref.set_model(w0, w, v)
        rv, rw, rw0, rloss = ref.gradient(row_ptr, idx, val.astype(np.float64), label, whole)
        assert np.allclose(tot[:n_slots * k], rv.reshape(-1), rtol=1e-12, atol=1e-13)
        assert np.allclose(tot[n_slots * k:n_slots * k + n_slots], rw, rtol=1e-12, atol=1e-13)
        assert abs(tot[-3] - rw0) < 1e-10 and abs(tot[-2] - rloss) < 1e-9 and tot[-1] == len(whole)

        # 5. the same update from the summed gradient on every rank keeps the replicas identical
        import ctypes as C
        g = np.ascontiguousarray(tot[:-2].copy())
        g[-1] = tot[-3]
        capi.lib().fmo_update(C.byref(orc.p), C.byref(orc.w0), orc.w.ctypes.data_as(C.POINTER(C.c_double)),
                              orc.v.ctypes.data_as(C.POINTER(C.c_double)),
                              g.ctypes.data_as(C.POINTER(C.c_double)), it, step, int(tot[-1]))
        ref.train_step(row_ptr, idx, val.astype(np.float64), label, whole, it, step)
        assert np.allclose(orc.v, ref.v, rtol=1e-12, atol=1e-14) and np.allclose(orc.w, ref.w, rtol=1e-12, atol=1e-14)
        mine_v = torch.from_numpy(orc.v.copy())
        other = [torch.zeros_like(mine_v) for _ in range(world)]
        dist.all_gather(other, mine_v)
        assert all(torch.equal(other[0], o) for o in other)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
