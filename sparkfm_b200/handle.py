"""Thin numpy-facing wrapper over the C ABI (one Handle == one sfm_handle == one GPU).

Only marshals arguments; every number is computed by libsparkfm_b200.so on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import SfmConfig, SfmStats, check

REGRESSION = 0
CLASSIFICATION = 1


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


class Handle:
    def __init__(self, n_slots, k, task=REGRESSION, k0=True, k1=True, reg=(0.0, 0.0, 0.0),
                 step_size=0.1, mini_batch_fraction=1.0, sampler_seed=42, device=0,
                 sampler_mode=0, shard_v=False):
        self._L = _lib.load()
        cfg = SfmConfig(_lib.SFM_ABI_VERSION, int(task), int(k), int(bool(k0)), int(bool(k1)),
                        int(device), int(n_slots), float(reg[0]), float(reg[1]), float(reg[2]),
                        float(step_size), float(mini_batch_fraction),
                        int(sampler_mode) | (0x100 if shard_v else 0),
                        int(sampler_seed))
        self._h = C.c_void_p()
        check(self._L.sfm_create(C.byref(cfg), C.byref(self._h)))
        self.n_slots, self.k = int(n_slots), int(k)

    @classmethod
    def _from_raw(cls, raw):
        self = cls.__new__(cls)
        self._L = _lib.load()
        self._h = raw
        cfg = self.config()
        self.n_slots, self.k = int(cfg.n_slots), int(cfg.k)
        return self

    @classmethod
    def load(cls, path, device=0):
        L = _lib.load()
        raw = C.c_void_p()
        check(L.sfm_load(str(path).encode(), int(device), C.byref(raw)))
        return cls._from_raw(raw)

    def close(self):
        if getattr(self, "_h", None):
            self._L.sfm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        check(st, self._h)

    # ------------------------------------------------------------------ config / model
    def config(self):
        cfg = SfmConfig()
        self._ck(self._L.sfm_get_config(self._h, C.byref(cfg)))
        return cfg

    def set_hyper(self, reg0, regw, regv, step_size, mini_batch_fraction):
        self._ck(self._L.sfm_set_hyper(self._h, reg0, regw, regv, step_size, mini_batch_fraction))

    def init_model(self, mean=0.0, stdev=0.01, seed=1):
        self._ck(self._L.sfm_init_model(self._h, mean, stdev, seed))

    def set_model(self, w0, w, v):
        w = _arr(w, np.float32)
        v = _arr(v, np.float32)
        if w is not None and w.size != self.n_slots:
            raise ValueError("w has the wrong length")
        if v is not None and v.size != self.n_slots * self.k:
            raise ValueError("v has the wrong size")
        self._ck(self._L.sfm_set_model(self._h, float(w0), _p(w, C.c_float),
                                       _p(v if self.k > 0 else None, C.c_float)))

    def get_model(self):
        w0 = C.c_float()
        w = np.empty(self.n_slots, dtype=np.float32)
        v = np.empty((self.n_slots, self.k), dtype=np.float32)
        self._ck(self._L.sfm_get_model(self._h, C.byref(w0), _p(w, C.c_float),
                                       _p(v if self.k > 0 else None, C.c_float)))
        return float(w0.value), w, v

    def set_model_f64(self, w0, w, v):
        w = _arr(w, np.float64)
        v = _arr(v, np.float64)
        self._ck(self._L.sfm_set_model_f64(self._h, float(w0), _p(w, C.c_double),
                                           _p(v if self.k > 0 else None, C.c_double)))

    def get_model_f64(self):
        w0 = C.c_double()
        w = np.empty(self.n_slots, dtype=np.float64)
        v = np.empty((self.n_slots, self.k), dtype=np.float64)
        self._ck(self._L.sfm_get_model_f64(self._h, C.byref(w0), _p(w, C.c_double),
                                           _p(v if self.k > 0 else None, C.c_double)))
        return float(w0.value), w, v

    def save(self, path):
        self._ck(self._L.sfm_save(self._h, str(path).encode()))

    # ------------------------------------------------------------------ scorer
    def predict(self, row_ptr, idx, val):
        row_ptr = _arr(row_ptr, np.int64)
        idx = _arr(idx, np.int32)
        val = _arr(val, np.float32)
        n = len(row_ptr) - 1
        out = np.empty(max(n, 0), dtype=np.float32)
        self._ck(self._L.sfm_predict(self._h, _p(row_ptr, C.c_int64), _p(idx, C.c_int32),
                                     _p(val, C.c_float), n, _p(out, C.c_float)))
        return out

    # ------------------------------------------------------------------ resident data set
    def load_dataset(self, row_ptr, idx, val, label, global_row_offset=0):
        row_ptr = _arr(row_ptr, np.int64)
        idx = _arr(idx, np.int32)
        val = _arr(val, np.float32)
        label = _arr(label, np.float32)
        n = len(row_ptr) - 1
        self._ck(self._L.sfm_load_dataset(self._h, _p(row_ptr, C.c_int64), _p(idx, C.c_int32),
                                          _p(val, C.c_float), _p(label, C.c_float), n,
                                          int(global_row_offset)))

    def unload_dataset(self):
        self._ck(self._L.sfm_unload_dataset(self._h))

    def synth_ctr_dataset(self, n_rows, global_row_offset, field_log2_card, zipf_cdf, zipf_cdf_off,
                          seed):
        card = _arr(field_log2_card, np.int32)
        cdf = _arr(zipf_cdf, np.uint32)
        off = _arr(zipf_cdf_off, np.int64)
        self._ck(self._L.sfm_synth_ctr_dataset(self._h, int(n_rows), int(global_row_offset),
                                               len(card), _p(card, C.c_int32),
                                               _p(cdf, C.c_uint32), _p(off, C.c_int64), int(seed)))

    def dataset_info(self):
        n, nnz, mx = C.c_int64(), C.c_int64(), C.c_int32()
        self._ck(self._L.sfm_dataset_info(self._h, C.byref(n), C.byref(nnz), C.byref(mx)))
        return int(n.value), int(nnz.value), int(mx.value)

    def get_dataset_rows(self, row_lo, row_hi, with_val=True):
        n = row_hi - row_lo
        row_ptr = np.empty(n + 1, dtype=np.int64)
        # first call fetches the pointers so the entry count is known
        self._ck(self._L.sfm_get_dataset_rows(self._h, row_lo, row_hi, _p(row_ptr, C.c_int64),
                                              None, None, None))
        cnt = int(row_ptr[-1])
        idx = np.empty(cnt, dtype=np.int32)
        val = np.empty(cnt, dtype=np.float32) if with_val else None
        label = np.empty(n, dtype=np.float32)
        self._ck(self._L.sfm_get_dataset_rows(self._h, row_lo, row_hi, _p(row_ptr, C.c_int64),
                                              _p(idx, C.c_int32), _p(val, C.c_float),
                                              _p(label, C.c_float)))
        return row_ptr, idx, val, label

    def predict_resident(self, row_lo, row_hi):
        out = np.empty(row_hi - row_lo, dtype=np.float32)
        self._ck(self._L.sfm_predict_resident(self._h, row_lo, row_hi, _p(out, C.c_float)))
        return out

    def predict_resident_raw(self, row_lo, row_hi, out_p):
        """Same into a caller-owned buffer given by address (e.g. pinned memory from host_alloc,
        which makes the device->host copy of the predictions a DMA at PCIe speed)."""
        self._ck(self._L.sfm_predict_resident(self._h, int(row_lo), int(row_hi),
                                              C.cast(out_p, C.POINTER(C.c_float))))

    def evaluate(self):
        m = np.zeros(5, dtype=np.float64)
        self._ck(self._L.sfm_evaluate(self._h, _p(m, C.c_double)))
        return {"rmse": m[0], "mean_error": m[1], "accuracy": m[2], "logloss": m[3], "n": int(m[4])}

    def evaluate_auc(self):
        m = np.zeros(3, dtype=np.float64)
        self._ck(self._L.sfm_evaluate_auc(self._h, _p(m, C.c_double)))
        return {"auc": m[0], "n_pos": int(m[1]), "n_neg": int(m[2])}

    # ------------------------------------------------------------------ learner
    def train_step(self, it, row_ids=None):
        """One SGD iteration on resident rows; row_ids None -> built-in sampler.
        Returns (mean loss over the global batch, global batch size)."""
        loss, batch = C.c_double(), C.c_int64()
        if row_ids is None:
            self._ck(self._L.sfm_train_step(self._h, None, -1, int(it), C.byref(loss),
                                            C.byref(batch)))
        else:
            ids = _arr(row_ids, np.int64)
            self._ck(self._L.sfm_train_step(self._h, _p(ids, C.c_int64), len(ids), int(it),
                                            C.byref(loss), C.byref(batch)))
        return float(loss.value), int(batch.value)

    def train_step_csr(self, it, row_ptr, idx, val, label):
        row_ptr = _arr(row_ptr, np.int64)
        idx = _arr(idx, np.int32)
        val = _arr(val, np.float32)
        label = _arr(label, np.float32)
        loss, batch = C.c_double(), C.c_int64()
        self._ck(self._L.sfm_train_step_csr(self._h, _p(row_ptr, C.c_int64), _p(idx, C.c_int32),
                                            _p(val, C.c_float), _p(label, C.c_float),
                                            len(row_ptr) - 1, int(it), C.byref(loss),
                                            C.byref(batch)))
        return float(loss.value), int(batch.value)

    def train_step_csr_raw(self, it, row_ptr_p, idx_p, val_p, label_p, n_rows):
        """Same with raw addresses (pinned buffers from host_alloc): no numpy conversion."""
        loss, batch = C.c_double(), C.c_int64()
        self._ck(self._L.sfm_train_step_csr(
            self._h, C.cast(row_ptr_p, C.POINTER(C.c_int64)), C.cast(idx_p, C.POINTER(C.c_int32)),
            C.cast(val_p, C.POINTER(C.c_float)) if val_p else None,
            C.cast(label_p, C.POINTER(C.c_float)), int(n_rows), int(it), C.byref(loss),
            C.byref(batch)))
        return float(loss.value), int(batch.value)

    def stage_csr_raw(self, slot, row_ptr_p, idx_p, val_p, label_p, n_rows):
        """Queue the H2D copy of a (pinned) host CSR batch into staging slot 0/1."""
        self._ck(self._L.sfm_stage_csr(
            self._h, int(slot), C.cast(row_ptr_p, C.POINTER(C.c_int64)),
            C.cast(idx_p, C.POINTER(C.c_int32)),
            C.cast(val_p, C.POINTER(C.c_float)) if val_p else None,
            C.cast(label_p, C.POINTER(C.c_float)), int(n_rows)))

    def stage_csr(self, slot, row_ptr, idx, val, label):
        """numpy variant; the arrays must stay alive until the staged step has run."""
        self._staged = getattr(self, "_staged", {})
        arrs = (_arr(row_ptr, np.int64), _arr(idx, np.int32), _arr(val, np.float32),
                _arr(label, np.float32))
        self._staged[slot] = arrs
        self._ck(self._L.sfm_stage_csr(self._h, int(slot), _p(arrs[0], C.c_int64),
                                       _p(arrs[1], C.c_int32), _p(arrs[2], C.c_float),
                                       _p(arrs[3], C.c_float), len(arrs[0]) - 1))

    def stage_onehot_raw(self, slot, packed_p, label_bits_p, label_f32_p, n_rows, m, id_bits):
        """Queue the H2D copy + device unpack of a compact one-hot batch (raw pinned addresses)."""
        self._ck(self._L.sfm_stage_onehot(
            self._h, int(slot), C.cast(packed_p, C.POINTER(C.c_uint32)),
            C.cast(label_bits_p, C.POINTER(C.c_uint32)) if label_bits_p else None,
            C.cast(label_f32_p, C.POINTER(C.c_float)) if label_f32_p else None,
            int(n_rows), int(m), int(id_bits)))

    def stage_onehot(self, slot, packed, label_bits, n_rows, m, id_bits, label_f32=None):
        """numpy form; the arrays must stay alive and unchanged until the slot is consumed."""
        packed = _arr(packed, np.uint32)
        label_bits = _arr(label_bits, np.uint32)
        label_f32 = _arr(label_f32, np.float32)
        self._keep = getattr(self, "_keep", {})
        self._keep[slot] = (packed, label_bits, label_f32)
        self._ck(self._L.sfm_stage_onehot(self._h, int(slot), _p(packed, C.c_uint32),
                                          _p(label_bits, C.c_uint32), _p(label_f32, C.c_float),
                                          int(n_rows), int(m), int(id_bits)))

    def train_step_staged(self, slot, it):
        loss, batch = C.c_double(), C.c_int64()
        self._ck(self._L.sfm_train_step_staged(self._h, int(slot), int(it), C.byref(loss),
                                               C.byref(batch)))
        return float(loss.value), int(batch.value)

    def train(self, first_iter, n_iters):
        hist = np.zeros(max(n_iters, 0), dtype=np.float64)
        self._ck(self._L.sfm_train(self._h, int(first_iter), int(n_iters), _p(hist, C.c_double)))
        return hist

    def gradient(self, row_ids=None):
        gv = np.empty((self.n_slots, self.k), dtype=np.float32)
        gw = np.empty(self.n_slots, dtype=np.float32)
        gw0, loss, batch = C.c_float(), C.c_double(), C.c_int64()
        ids = _arr(row_ids, np.int64)
        self._ck(self._L.sfm_gradient(self._h, _p(ids, C.c_int64), -1 if ids is None else len(ids),
                                      _p(gv if self.k > 0 else None, C.c_float),
                                      _p(gw, C.c_float), C.byref(gw0), C.byref(loss),
                                      C.byref(batch)))
        return gv, gw, float(gw0.value), float(loss.value), int(batch.value)

    # ------------------------------------------------------------------ ALS
    def als_sweep(self, ref_quirks: bool = False) -> float:
        """ALS.learn (fm/lib/ALS.scala:15-75): one sweep over w0, w, V on the resident data set.
        Returns the RMSE of the residuals after the sweep."""
        out = C.c_double()
        self._ck(self._L.sfm_als_sweep(self._h, 1 if ref_quirks else 0, C.byref(out)))
        return out.value

    def als_residuals(self, n_rows: int) -> np.ndarray:
        e = np.empty(int(n_rows), dtype=np.float64)
        self._ck(self._L.sfm_als_residuals(self._h, e.ctypes.data_as(C.POINTER(C.c_double)), len(e)))
        return e

    # ------------------------------------------------------------------ multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * _lib.SFM_UNIQUE_ID_BYTES)()
        check(_lib.load().sfm_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, world_size: int):
        buf = (C.c_uint8 * _lib.SFM_UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        self._ck(self._L.sfm_comm_init(self._h, buf, int(rank), int(world_size)))

    def comm_mode(self) -> str:
        """'none' | 'nccl' | 'peer' (fused sum + update over NVLink peer memory) | 'sharded'."""
        m = C.c_int32()
        self._ck(self._L.sfm_comm_mode(self._h, C.byref(m)))
        return ("none", "nccl", "peer", "sharded")[m.value]

    def comm_broadcast_model(self):
        self._ck(self._L.sfm_comm_broadcast_model(self._h))

    # ------------------------------------------------------------------ stats / timing
    def stats(self):
        s = SfmStats()
        self._ck(self._L.sfm_stats_get(self._h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in SfmStats._fields_}

    def stats_reset(self):
        self._ck(self._L.sfm_stats_reset(self._h))

    def set_phase_timing(self, enabled):
        self._ck(self._L.sfm_set_phase_timing(self._h, int(bool(enabled))))

    def synchronize(self):
        self._ck(self._L.sfm_synchronize(self._h))

    def timer_start(self):
        self._ck(self._L.sfm_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self._L.sfm_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)


def pack_onehot(idx, label, m, id_bits):
    """Host packer of the compact one-hot staging format: (packed uint32 words, label bit words)."""
    idx = _arr(np.asarray(idx).reshape(-1), np.int32)
    n_rows = len(idx) // m
    assert n_rows * m == len(idx)
    label = _arr(label, np.float32)
    words = (n_rows * m * id_bits + 31) // 32 + 1
    packed = np.empty(words, dtype=np.uint32)
    lbits = np.zeros(max(1, (n_rows + 31) // 32), dtype=np.uint32)
    check(_lib.load().sfm_pack_onehot(_p(idx, C.c_int32), _p(label, C.c_float), n_rows, int(m),
                                 int(id_bits), _p(packed, C.c_uint32),
                                 _p(lbits, C.c_uint32) if label is not None else None))
    return packed, lbits


def sample_rows(seed, it, fraction, row_lo, row_hi):
    """Host twin of the device sampler through the C ABI."""
    L = _lib.load()
    out = np.empty(max(row_hi - row_lo, 0), dtype=np.int64)
    n = C.c_int64()
    check(L.sfm_sample_rows(int(seed), int(it), float(fraction), int(row_lo), int(row_hi),
                            _p(out, C.c_int64), C.byref(n)))
    return out[:n.value].copy()


def partition_rows(seed, n_parts, part, row_lo, row_hi):
    """Host twin of the PARTITION sampler through the C ABI."""
    L = _lib.load()
    out = np.empty(max(row_hi - row_lo, 0), dtype=np.int64)
    n = C.c_int64()
    check(L.sfm_partition_rows(int(seed), int(n_parts), int(part), int(row_lo), int(row_hi),
                               _p(out, C.c_int64), C.byref(n)))
    return out[:n.value].copy()


def parse_libfm(text: bytes, num_features: int = -1):
    """FMUtils.loadLibFMFile on a bytes buffer -> (label f64, row_ptr i64, idx i32, val f64, d).
    One library call: the output arrays are sized from what the grammar allows per byte (a row
    needs a label and a terminator, an entry " i:x" four bytes; untouched pages of the
    over-allocation are never resident), the library counts exactly and fills, the arrays are
    trimmed.  Falls back to count-then-fill if the bounds cannot be allocated."""
    L = _lib.load()
    n, nnz, d, el = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int64(-1)
    try:
        max_rows, max_nnz = len(text) // 2 + 1, len(text) // 4 + 1
        label = np.empty(max_rows, dtype=np.float64)
        row_ptr = np.empty(max_rows + 1, dtype=np.int64)
        idx = np.empty(max_nnz, dtype=np.int32)
        val = np.empty(max_nnz, dtype=np.float64)
    except MemoryError:
        st = L.sfm_parse_libfm(text, len(text), int(num_features), C.byref(n), C.byref(nnz),
                               C.byref(d), None, None, None, None, C.byref(el))
        if st != _lib.SFM_OK:
            raise ValueError(f"LibFM parse error at line {el.value}")
        label = np.empty(n.value, dtype=np.float64)
        row_ptr = np.empty(n.value + 1, dtype=np.int64)
        idx = np.empty(max(nnz.value, 1), dtype=np.int32)
        val = np.empty(max(nnz.value, 1), dtype=np.float64)
    st = L.sfm_parse_libfm(text, len(text), int(num_features), C.byref(n), C.byref(nnz),
                           C.byref(d), _p(label, C.c_double), _p(row_ptr, C.c_int64),
                           _p(idx, C.c_int32), _p(val, C.c_double), C.byref(el))
    if st != _lib.SFM_OK:
        raise ValueError(f"LibFM parse error at line {el.value}")
    return (label[:n.value], row_ptr[:n.value + 1], idx[:nnz.value], val[:nnz.value], int(d.value))


def format_libfm(label, row_ptr, idx, val) -> bytes:
    L = _lib.load()
    label = _arr(label, np.float64)
    row_ptr = _arr(row_ptr, np.int64)
    idx = _arr(idx, np.int32)
    val = _arr(val, np.float64)
    need = C.c_uint64()
    n = len(row_ptr) - 1
    check(L.sfm_format_libfm(_p(label, C.c_double), _p(row_ptr, C.c_int64), _p(idx, C.c_int32),
                             _p(val, C.c_double), n, None, 0, C.byref(need)))
    buf = C.create_string_buffer(int(need.value) + 1)
    check(L.sfm_format_libfm(_p(label, C.c_double), _p(row_ptr, C.c_int64), _p(idx, C.c_int32),
                             _p(val, C.c_double), n, buf, need.value, C.byref(need)))
    return buf.raw[:need.value]


def device_count() -> int:
    return int(_lib.load().sfm_device_count())
