"""numpy / pure-Python restatement of the SparkFM hot path (second, independent oracle).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs, never by the product package.  PARITY UNPINNED: the reference cannot run in this
image and ships no tests or golden vectors; `predict` follows the reference
SOURCE, the SGD pieces follow the written spec in DESIGN.md section 2.

Citations are relative to /root/reference/src/main/scala/io/edstud/spark/.
"""
from __future__ import annotations

import math
import re

import numpy as np

REGRESSION = 0      # Task.Regression      Task.scala:5
CLASSIFICATION = 1  # Task.Classification  Task.scala:5

_M64 = (1 << 64) - 1


# ----------------------------------------------------------------------------- hashing
def mix64(x: int) -> int:
    """splitmix64 finaliser on a Python int (mod 2^64)."""
    z = (x + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def mix64_np(x: np.ndarray) -> np.ndarray:
    """Vectorised mix64 on uint64 arrays (wraps mod 2^64)."""
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


# ----------------------------------------------------------------------------- predict
def predict_row(w0, w, v, idx, val, k0=True, k1=True) -> float:
    """fm/FMModel.scala:34-55 with computeFactorComponents :57-63, literally.

    `v` is [n_slots][k] (feature-major).  Python floats are IEEE doubles, the folds run in
    stored order, duplicates and explicit zeros are visited like Breeze's activeIterator."""
    k = v.shape[1]
    result = 0.0
    if k0:
        result += float(w0)
    if len(idx) > 0:
        if k1:
            terms = [float(w[i]) * float(x) for i, x in zip(idx, val)]
            acc = terms[0]
            for t in terms[1:]:
                acc = acc + t
            result += acc
        for f in range(k):
            fl = [float(v[i, f]) * float(x) for i, x in zip(idx, val)]
            sum_f = fl[0]
            for t in fl[1:]:
                sum_f = sum_f + t
            sq = [t * t for t in fl]
            sum_sqr = sq[0]
            for t in sq[1:]:
                sum_sqr = sum_sqr + t
            result += 0.5 * (sum_f * sum_f - sum_sqr)
    return result


def predict(w0, w, v, row_ptr, idx, val, k0=True, k1=True) -> np.ndarray:
    n = len(row_ptr) - 1
    out = np.empty(n, dtype=np.float64)
    for r in range(n):
        b, e = int(row_ptr[r]), int(row_ptr[r + 1])
        out[r] = predict_row(w0, w, v, idx[b:e], val[b:e], k0, k1)
    return out


def predict_vec(w0, w, v, row_ptr, idx, val, k0=True, k1=True) -> np.ndarray:
    """Vectorised fp64 predict (segment sums with np.add.reduceat); same formula, numpy's
    summation order.  For mid-sized parity checks where the Python loop is too slow."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    n = len(row_ptr) - 1
    out = np.full(n, float(w0) if k0 else 0.0, dtype=np.float64)
    if n == 0 or len(idx) == 0:
        return out
    lens = np.diff(row_ptr)
    nz = lens > 0
    starts = row_ptr[:-1][nz]
    x = np.asarray(val, dtype=np.float64)
    vx = np.asarray(v, dtype=np.float64)[idx] * x[:, None]
    s = np.add.reduceat(vx, starts, axis=0)
    q = np.add.reduceat(vx * vx, starts, axis=0)
    pair = 0.5 * (s * s - q).sum(axis=1)
    if k1:
        lin = np.add.reduceat(np.asarray(w, dtype=np.float64)[idx] * x, starts)
        out[nz] += lin
    out[nz] += pair
    return out


# ----------------------------------------------------------------------------- SGD spec
def loss_mult(task: int, yhat: float, label: float):
    """DESIGN.md section 2.2."""
    if task == CLASSIFICATION:
        y = 1.0 if label > 0.0 else -1.0
        m = y * yhat
        e = math.exp(-abs(m))
        loss = (0.0 if m > 0.0 else -m) + math.log1p(e)
        sig = e / (1.0 + e) if m > 0.0 else 1.0 / (1.0 + e)
        return loss, -y * sig
    d = yhat - label
    return d * d, d


def gradient(task, w0, w, v, row_ptr, idx, val, label, row_ids, k0=True, k1=True):
    """Dense gradient (gV [n_slots][k], gw [n_slots], gw0) and loss sum over row_ids."""
    n_slots, k = v.shape
    gv = np.zeros((n_slots, k), dtype=np.float64)
    gw = np.zeros(n_slots, dtype=np.float64)
    gw0 = 0.0
    loss_sum = 0.0
    vd = np.asarray(v, dtype=np.float64)
    wd = np.asarray(w, dtype=np.float64)
    for r in row_ids:
        b, e = int(row_ptr[r]), int(row_ptr[r + 1])
        ii = np.asarray(idx[b:e], dtype=np.int64)
        xx = np.asarray(val[b:e], dtype=np.float64)
        yhat = float(w0) if k0 else 0.0
        s = np.zeros(k)
        if e > b:
            vx = vd[ii] * xx[:, None]
            s = vx.sum(axis=0)
            q = (vx * vx).sum(axis=0)
            if k1:
                yhat += float((wd[ii] * xx).sum())
            yhat += float(0.5 * (s * s - q).sum())
        loss, mult = loss_mult(task, yhat, float(label[r]))
        loss_sum += loss
        if k0:
            gw0 += mult
        if e > b:
            if k1:
                np.add.at(gw, ii, xx * mult)
            np.add.at(gv, ii, (xx[:, None] * s[None, :] - vd[ii] * (xx * xx)[:, None]) * mult)
    return gv, gw, gw0, loss_sum


def update(w0, w, v, gv, gw, gw0, it, step_size, batch_count, reg0, regw, regv,
           k0=True, k1=True):
    """theta <- theta - eta*(g/|batch| + lambda*theta), eta = step/sqrt(it)."""
    if batch_count <= 0:
        return w0, w, v
    eta = step_size / math.sqrt(it)
    inv = 1.0 / batch_count
    v = v - eta * (gv * inv + regv * v)
    if k1:
        w = w - eta * (gw * inv + regw * w)
    if k0:
        w0 = w0 - eta * (gw0 * inv + reg0 * w0)
    return w0, w, v


def als_sweep(w0, w, v, row_ptr, idx, val, label, k0=True, k1=True, reg=(0.0, 0.0, 0.0),
              ref_quirks=False, store_f32=False):
    """ALS.learn, fm/lib/ALS.scala:15-75, transliterated with the Scala's own data structures
    (`e` and `q` as maps keyed by row, `features` as a map id -> column) -- independent of the C
    oracle's arrays.  Pure Python: small cases only.  Returns (w0, w, v, e list)."""
    w = np.array(w, dtype=np.float64)
    v = np.array(v, dtype=np.float64)
    n_slots, k = v.shape
    n = len(row_ptr) - 1
    r0, rw, rv = (float(x) for x in reg)

    def rnd(x):
        return float(np.float32(x)) if store_f32 else x

    def updatable(nv, old):                                   # :178-180
        return not math.isnan(nv) and not math.isinf(nv) and nv != old

    def compute_theta(theta, lam, sum_e_h, sum_h_sqr):         # :167-176
        den = lam + sum_h_sqr
        num = -(sum_e_h - theta * sum_h_sqr)
        if den == 0.0:
            nv = math.nan if num == 0.0 or math.isnan(num) else math.copysign(math.inf, num)
        else:
            nv = num / den
        return nv if updatable(nv, theta) else theta

    e = {}
    for r in range(n):                                        # precomputeTermE :142-144
        a, b = int(row_ptr[r]), int(row_ptr[r + 1])
        e[r] = predict_row(w0, w, v, idx[a:b], val[a:b], k0, k1) - float(label[r])
    if k0:                                                    # :19-28
        total = 0.0
        for r in range(n):
            total = e[r] if r == 0 else total + e[r]
        nw0 = rnd(compute_theta(w0, r0, total, float(n)))
        # :24 is lazily evaluated to "+ 0" at :31, but the same re-evaluation runs fm.predict with
        # the new w0 (:27, :142-144): the materialised residuals are shifted either way
        if updatable(nw0, w0):
            for r in range(n):
                e[r] = e[r] + (nw0 - w0)
        w0 = nw0
    features = {}                                             # transposeInput, DataSet.scala:31-38
    for r in range(n):
        for j in range(int(row_ptr[r]), int(row_ptr[r + 1])):
            features.setdefault(int(idx[j]), []).append((r, float(val[j])))
    id_end = n_slots - 1 if ref_quirks else n_slots           # the one quirk: `0 until num_attribute`

    def draw_theta(theta, lam, h):                            # :156-165, h = [(row, value)]
        sum_h_sqr = sum_e_h = 0.0
        for j, (r, x) in enumerate(h):                        # :183-190 left folds
            sum_h_sqr = x * x if j == 0 else sum_h_sqr + x * x
            sum_e_h = e[r] * x if j == 0 else sum_e_h + e[r] * x
        nt = rnd(compute_theta(theta, lam, sum_e_h, sum_h_sqr))
        if updatable(nt, theta):
            for r, x in h:                                    # updateError :194-198
                e[r] += x * (nt - theta)
        return nt

    if k1:                                                    # :36-43
        for i in range(id_end):
            if i in features:
                w[i] = draw_theta(w[i], rw, features[i])
    for f in range(k):                                        # :45-70
        q = {}
        for r in range(n):                                    # precomputeTermQ :146-150
            s = 0.0
            for j in range(int(row_ptr[r]), int(row_ptr[r + 1])):
                s += v[int(idx[j]), f] * float(val[j])
            q[r] = s
        for i in range(id_end):
            if i in features:
                old = float(v[i, f])
                h = [(r, x * q[r] - x * x * old) for r, x in features[i]]
                new = draw_theta(old, rv, h)
                for r, x in features[i]:
                    q[r] += x * (new - old)
                v[i, f] = new
    return w0, w, v, [e[r] for r in range(n)]


def sample_rows(seed: int, it: int, fraction: float, row_lo: int, row_hi: int) -> np.ndarray:
    """DESIGN.md section 2.5 (vectorised over the 64-row blocks): bit-sliced Bernoulli(thr / 2^53)."""
    rows = np.arange(row_lo, row_hi, dtype=np.int64)
    if fraction >= 1.0:
        return rows
    if not fraction > 0.0 or row_hi <= row_lo:
        return rows[:0]
    thr = int(math.floor(fraction * 2.0 ** 53))
    if thr == 0:
        return rows[:0]
    key = mix64((seed + it) & _M64)
    gamma = 0x9E3779B97F4A7C15
    last = 53 - ((thr & -thr).bit_length() - 1)      # 1-based position of the lowest set digit
    q = np.arange(row_lo >> 6, ((row_hi - 1) >> 6) + 1, dtype=np.uint64)
    und = np.full(len(q), _M64, dtype=np.uint64)
    hit = np.zeros(len(q), dtype=np.uint64)
    with np.errstate(over="ignore"):
        ctr = np.uint64(key) + (q << np.uint64(6)) * np.uint64(gamma)
        for i in range(1, last + 1):
            if not und.any():
                break
            w = mix64_np(ctr)                        # SplitMix64(seed = key), output number 64 q + i - 1
            ctr = ctr + np.uint64(gamma)
            if (thr >> (53 - i)) & 1:
                hit |= und & ~w
                und &= w
            else:
                und &= ~w
    bits = (hit[:, None] >> np.arange(64, dtype=np.uint64)[None, :]) & np.uint64(1)
    g = (q.astype(np.int64)[:, None] << 6) + np.arange(64, dtype=np.int64)[None, :]
    sel = g[bits.astype(bool)]
    return sel[(sel >= row_lo) & (sel < row_hi)]


# ----------------------------------------------------------------------------- LibFM text
def _java_trim(s: str) -> str:
    """java.lang.String.trim: strips every char <= U+0020 from both ends."""
    b, e = 0, len(s)
    while b < e and s[b] <= " ":
        b += 1
    while e > b and s[e - 1] <= " ":
        e -= 1
    return s[b:e]


def _java_split(s: str, ch: str):
    """String.split(char): trailing empty strings are removed, leading ones are kept."""
    parts = s.split(ch)
    while parts and parts[-1] == "":
        parts.pop()
    return parts if parts else [s] if s == "" else parts


_JAVA_DOUBLE = re.compile(
    r"^[\x00-\x20]*[+-]?(NaN|Infinity|((\d+\.?\d*|\.\d+)([eE][+-]?\d+)?)[fFdD]?)[\x00-\x20]*$")
_JAVA_INT = re.compile(r"^[+-]?\d+$")


def java_double(s: str) -> float:
    """java.lang.Double.parseDouble for the decimal grammar: surrounding chars <= U+0020 are
    trimmed, optional sign, `NaN` / `Infinity`, digits with optional fraction and exponent and
    an optional f/F/d/D suffix.  Hex float literals are NOT supported (documented deviation)."""
    m = _JAVA_DOUBLE.match(s)
    if not m:
        raise ValueError(f"NumberFormatException: {s!r}")
    body = _java_trim(s).rstrip("fFdD") if m.group(1) not in ("NaN", "Infinity") else _java_trim(s)
    return float(body.replace("Infinity", "inf").replace("NaN", "nan"))


def java_int(s: str) -> int:
    """java.lang.Integer.parseInt: optional sign, decimal digits, no whitespace, 32-bit range."""
    if not _JAVA_INT.match(s):
        raise ValueError(f"NumberFormatException: {s!r}")
    v = int(s)
    if not -(1 << 31) <= v < (1 << 31):
        raise ValueError(f"NumberFormatException: {s!r} out of int range")
    return v


def parse_libfm_lines(lines, num_features: int = -1):
    """fm/FMUtils.scala:23-53 loadLibFMFile, on an iterable of text lines.

    Returns (labels f64[N], row_ptr i64[N+1], idx i32[nnz], val f64[nnz], d) where the vectors
    have length d+1 (:50).  Indices are kept verbatim -- no shift (:32); order and duplicates
    are preserved; lines that are empty after trim or start with '#' are skipped (:25-26);
    empty tokens from repeated spaces are dropped (:30).  Raises ValueError where the Scala
    would throw (bad number, token without ':value', row with no features when the dimension
    has to be inferred -- `indices.max` on an empty array, :45)."""
    labels, row_ptr, idx, val = [], [0], [], []
    for ln, raw in enumerate(lines, 1):
        line = _java_trim(raw.rstrip("\n").rstrip("\r") if raw.endswith(("\n", "\r")) else raw)
        if line == "" or line.startswith("#"):
            continue
        items = _java_split(line, " ")
        try:
            labels.append(java_double(items[0]))
        except ValueError as ex:
            raise ValueError(f"line {ln}: bad label {items[0]!r}") from ex
        for item in items[1:]:
            if item == "":
                continue
            iv = _java_split(item, ":")
            if len(iv) < 2:
                raise ValueError(f"line {ln}: token {item!r} has no ':value'")
            try:
                i = java_int(iv[0])
                x = java_double(iv[1])
            except ValueError as ex:
                raise ValueError(f"line {ln}: bad token {item!r}") from ex
            idx.append(i)
            val.append(x)
        row_ptr.append(len(idx))
    labels_a = np.asarray(labels, dtype=np.float64)
    row_ptr_a = np.asarray(row_ptr, dtype=np.int64)
    idx_a = np.asarray(idx, dtype=np.int64)
    val_a = np.asarray(val, dtype=np.float64)
    if num_features > 0:
        d = num_features
    else:
        if len(labels) == 0:
            raise ValueError("empty collection: cannot infer the dimension")  # reduce on empty RDD
        if np.any(np.diff(row_ptr_a) == 0):
            raise ValueError("row without features: indices.max on an empty array")
        d = int(idx_a.max())
    return labels_a, row_ptr_a, idx_a.astype(np.int32), val_a, d
