"""Seeded synthetic data for the BASELINE.json configs.

`ctr_rows` is the numpy twin of the device generator `sfm_synth_ctr_dataset`
(csrc/sfm_kernels.cu synth_ctr_kernel): integer arithmetic only, bit-identical.
"""
from __future__ import annotations

import numpy as np

_U64 = np.uint64
SALT_FEATURE = _U64(0x5FE14BD7A3C1E5B3)
SALT_A = _U64(0x1B873593CC9E2D51)
SALT_B = _U64(0x85EBCA6BC2B2AE35)
SALT_NOISE = _U64(0x27D4EB2F165667C5)
LABEL_THRESHOLD = 110000  # ~25 % positives on the Criteo- and Avazu-shaped configs


def mix64(x):
    """splitmix64 finaliser on uint64 arrays / scalars (wraps mod 2^64)."""
    with np.errstate(over="ignore"):
        z = np.asarray(x, dtype=np.uint64) + _U64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
        return z ^ (z >> _U64(31))


# ------------------------------------------------------------------ CTR-shaped (C3, C4)
def ctr_field_log2_cards(n_fields: int) -> np.ndarray:
    """Criteo-like mix: the first third of the fields are small 'bucketised numeric' fields
    (64 ids), the rest spread log-uniformly from 2^4 to 2^20 ids."""
    n_small = n_fields // 3
    rest = n_fields - n_small
    spread = [4 + round(16 * j / max(rest - 1, 1)) for j in range(rest)]
    return np.asarray([6] * n_small + spread, dtype=np.int32)


def zipf_tables(field_log2_card, s: float = 1.1):
    """Per-field inverse-CDF tables for Zipf(s) over 2^c ids: cdf[j] = floor(2^32 * P(rank <= j))
    clipped to 2^32-1; fields with the same cardinality share one table."""
    tabs, offs, where, pos = [], [], {}, 0
    for c in np.asarray(field_log2_card).tolist():
        if c not in where:
            n = 1 << c
            p = np.arange(1, n + 1, dtype=np.float64) ** (-s)
            cdf = np.cumsum(p)
            cdf /= cdf[-1]
            t = np.minimum(np.floor(cdf * 4294967296.0), 4294967295.0).astype(np.uint32)
            t[-1] = np.uint32(4294967295)
            where[c] = pos
            tabs.append(t)
            pos += n
        offs.append(where[c])
    return np.concatenate(tabs), np.asarray(offs, dtype=np.int64)


def ctr_rows(row_lo: int, row_hi: int, field_log2_card, cdf, cdf_off, n_slots: int, seed: int):
    """Rows [row_lo, row_hi) of the synthetic CTR data set: idx int32 [n][F] (one id per field,
    values all 1), label float32 [n] in {0, 1}."""
    return ctr_rows_at(np.arange(row_lo, row_hi, dtype=np.uint64), field_log2_card, cdf, cdf_off,
                       n_slots, seed)


def ctr_rows_at(rows, field_log2_card, cdf, cdf_off, n_slots: int, seed: int):
    """The same for an arbitrary list of GLOBAL row numbers (a row depends on its number only), e.g.
    one sampled mini-batch of a 45 M-row data set without materialising the data set on the host."""
    card = np.asarray(field_log2_card, dtype=np.int64)
    F = len(card)
    gr = np.ascontiguousarray(rows).astype(np.uint64)
    n = len(gr)
    seed_key = mix64(_U64(seed))
    idx = np.empty((n, F), dtype=np.int32)
    lin = np.zeros(n, dtype=np.int64)
    sb = np.zeros(n, dtype=np.int64)
    sb2 = np.zeros(n, dtype=np.int64)
    with np.errstate(over="ignore"):
        for f in range(F):
            h = mix64(seed_key + gr * _U64(F) + _U64(f))
            u = (h >> _U64(32)).astype(np.uint32)
            tab = cdf[int(cdf_off[f]): int(cdf_off[f]) + (1 << int(card[f]))]
            rank = np.searchsorted(tab, u, side="left").astype(np.uint64)  # first j: u <= tab[j]
            fid = mix64((_U64(f) << _U64(40)) ^ rank ^ SALT_FEATURE) % _U64(n_slots)
            idx[:, f] = fid.astype(np.int32)
            a = (mix64(fid ^ SALT_A) & _U64(0xFFFF)).astype(np.int64) - 32768
            b = (mix64(fid ^ SALT_B) & _U64(0xFF)).astype(np.int64) - 128
            lin += a
            sb += b
            sb2 += b * b
        noise = (mix64(seed_key ^ mix64(gr ^ SALT_NOISE)) & _U64(0x7FFFF)).astype(np.int64) - 262144
    score = lin + (sb * sb - sb2) // 2 + noise
    label = (score > LABEL_THRESHOLD).astype(np.float32)
    return idx, label


def ctr_csr(row_lo, row_hi, n_fields, n_slots, seed, zipf_s=1.1):
    """Convenience: CSR arrays (row_ptr, idx flat, val=None, label) of CTR rows."""
    card = ctr_field_log2_cards(n_fields)
    cdf, off = zipf_tables(card, zipf_s)
    idx, label = ctr_rows(row_lo, row_hi, card, cdf, off, n_slots, seed)
    n = row_hi - row_lo
    row_ptr = np.arange(n + 1, dtype=np.int64) * n_fields
    return row_ptr, idx.reshape(-1), None, label


# ------------------------------------------------------------------ ragged sparse (C1, C2)
def _planted_fm(rng, n_features, k, v_std):
    return (rng.normal(0.0, 0.3), rng.normal(0.0, 0.3, n_features),
            rng.normal(0.0, v_std, (n_features, k)))


def _planted_score(w0, w, v, row_ptr, idx, val):
    n = len(row_ptr) - 1
    out = np.full(n, w0, dtype=np.float64)
    lens = np.diff(row_ptr)
    nz = lens > 0
    starts = row_ptr[:-1][nz]
    x = val.astype(np.float64)
    vx = v[idx] * x[:, None]
    s = np.add.reduceat(vx, starts, axis=0)
    q = np.add.reduceat(vx * vx, starts, axis=0)
    out[nz] += np.add.reduceat(w[idx] * x, starts) + 0.5 * (s * s - q).sum(axis=1)
    return out


def ragged_rows(n_rows, n_features, mean_nnz, seed, max_nnz=None, values="ones"):
    """Rows with nnz ~ Poisson(mean_nnz) clipped to [1, max_nnz], indices uniform without
    replacement per row (sorted), values 1.0 / N(0,1) / U(0,1]."""
    rng = np.random.default_rng(seed)
    max_nnz = max_nnz or max(4 * mean_nnz, 8)
    lens = np.clip(rng.poisson(mean_nnz, n_rows), 1, min(max_nnz, n_features)).astype(np.int64)
    row_ptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(lens, out=row_ptr[1:])
    nnz = int(row_ptr[-1])
    # uniform ids, sorted within the row; duplicates within a row are re-drawn until none remain
    rows = np.repeat(np.arange(n_rows, dtype=np.int64), lens)
    idx = rng.integers(0, n_features, nnz, dtype=np.int64)
    while True:
        key = rows * n_features + idx
        key.sort()
        idx = key % n_features          # rows is already sorted, so it is unchanged by the sort
        dup = np.zeros(nnz, dtype=bool)
        dup[1:] = key[1:] == key[:-1]
        if not dup.any():
            break
        idx[dup] = rng.integers(0, n_features, int(dup.sum()), dtype=np.int64)
    idx = idx.astype(np.int32)
    if values == "ones":
        val = np.ones(nnz, dtype=np.float32)
    elif values == "normal":
        val = rng.normal(0.0, 1.0, nnz).astype(np.float32)
    else:
        val = (1.0 - rng.random(nnz)).astype(np.float32)
    return row_ptr, idx, val


def classification_c1(n_rows=100_000, n_features=10_000, mean_nnz=20, k=8, seed=20260101):
    """BASELINE config 1: binary labels (+1/-1) from a planted FM + logistic noise."""
    row_ptr, idx, val = ragged_rows(n_rows, n_features, mean_nnz, seed, max_nnz=64)
    rng = np.random.default_rng(seed + 7)
    w0, w, v = _planted_fm(rng, n_features, k, 0.1)
    score = _planted_score(w0, w, v, row_ptr, idx, val)
    p = 1.0 / (1.0 + np.exp(-score))
    label = np.where(rng.random(n_rows) < p, 1.0, -1.0).astype(np.float32)
    return row_ptr, idx, val, label


def regression_c2(n_rows=1_000_000, n_features=100_000, mean_nnz=50, k=16, seed=20260102):
    """BASELINE config 2: y = planted FM(x) + N(0, 0.1^2), values N(0,1)."""
    row_ptr, idx, val = ragged_rows(n_rows, n_features, mean_nnz, seed, values="normal")
    rng = np.random.default_rng(seed + 7)
    w0, w, v = _planted_fm(rng, n_features, k, 0.05)
    y = _planted_score(w0, w, v, row_ptr, idx, val) + rng.normal(0.0, 0.1, n_rows)
    return row_ptr, idx, val, y.astype(np.float32)


def to_libfm_text(row_ptr, idx, val, label) -> str:
    """LibFM text in the LOADER's convention (index verbatim, FMUtils.scala:32), full-precision
    values -- for feeding the parser, not the reference's lossy saveAsLibFMFile."""
    lines = []
    for r in range(len(row_ptr) - 1):
        b, e = int(row_ptr[r]), int(row_ptr[r + 1])
        toks = [repr(float(label[r]))] + [f"{int(i)}:{float(x)!r}" for i, x in zip(idx[b:e], val[b:e])]
        lines.append(" ".join(toks))
    return "\n".join(lines) + "\n"
