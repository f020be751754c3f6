"""Generates tests/golden/predict_kat.json -- BUILDER-AUTHORED known-answer vectors for
FMModel.predict (reference: src/main/scala/io/edstud/spark/fm/FMModel.scala:34-63).

The reference ships no golden vectors and cannot run here, so these answers
are computed with exact rational arithmetic (fractions.Fraction) straight from the formula in
the Scala source -- independently of oracle/ -- and rounded once to the nearest double.  All
inputs are dyadic rationals, so every fp64 evaluation order gives the exact answer too.

Run:  python tests/golden/make_predict_kat.py
"""
import json
import os
from fractions import Fraction as F


def predict_exact(w0, w, v, idx, val, k, k0=True, k1=True):
    r = F(0)
    if k0:
        r += w0
    if idx:                                     # features.used > 0   (:42)
        if k1:
            r += sum(w[i] * x for i, x in zip(idx, val))
        for f in range(k):                      # (:48-51)
            t = [v[i][f] * x for i, x in zip(idx, val)]
            s = sum(t)
            q = sum(u * u for u in t)
            r += F(1, 2) * (s * s - q)
    return r


def loss_mult_exact_regression(yhat, y):
    d = yhat - y
    return d * d, d


CASES = []


def add(name, w0, w, v, rows, k, k0=True, k1=True):
    preds = [predict_exact(w0, w, v, i, x, k, k0, k1) for i, x in rows]
    CASES.append({
        "name": name, "k": k, "k0": k0, "k1": k1, "n_slots": len(w),
        "w0": float(w0), "w": [float(a) for a in w], "v": [[float(a) for a in r] for r in v],
        "rows": [{"idx": list(i), "val": [float(a) for a in x]} for i, x in rows],
        "predict": [float(p) for p in preds],
        "predict_exact": [f"{p.numerator}/{p.denominator}" for p in preds],
    })


# 1. the two-feature textbook case: yhat = w0 + w1 x1 + w2 x2 + <v1,v2> x1 x2
w = [F(0), F(1, 2), F(-1, 4), F(3, 8)]
v = [[F(0), F(0)], [F(1, 2), F(-1, 4)], [F(1, 8), F(3, 4)], [F(-1, 2), F(1, 2)]]
add("two_features_k2", F(1, 4), w, v,
    [([1, 2], [F(1), F(1)]),            # 1/4 + 1/2 - 1/4 + (1/16 - 3/16) = 3/8
     ([1, 2], [F(2), F(-3)]),
     ([1, 2, 3], [F(1), F(1), F(1)]),
     ([3], [F(5, 2)]),                  # single feature: pairwise term is exactly 0
     ([], [])],                         # empty row -> w0 only (:42)
    k=2)

# 2. stored order, duplicates and explicit zeros are all visited (activeIterator semantics)
add("duplicates_and_zeros", F(-1, 2), w, v,
    [([2, 2], [F(1), F(1)]),            # duplicate index: contributes <v2,v2> x x to the pair term
     ([3, 1, 2], [F(1, 2), F(0), F(4)]),
     ([0, 3], [F(7), F(1)])],           # feature 0 is a valid slot (no index shift, FMUtils:32)
    k=2)

# 3. bias flags
add("no_bias_no_linear", F(9), w, v, [([1, 3], [F(1), F(2)]), ([], [])], k=2, k0=False, k1=False)
add("no_linear", F(9), w, v, [([1, 3], [F(1), F(2)]), ([], [])], k=2, k0=True, k1=False)

# 4. k = 3 (not a multiple of 4: exercises factor padding on the device) and a longer row
w3 = [F(i - 3, 16) for i in range(8)]
v3 = [[F((i * 3 + f * 5) % 7 - 3, 8) for f in range(3)] for i in range(8)]
add("k3_long_row", F(1, 8), w3, v3,
    [(list(range(8)), [F((j % 3) + 1, 2) for j in range(8)]),
     ([7, 0, 5, 5, 2], [F(-1), F(1, 4), F(2), F(-2), F(3)])],
    k=3)

# 5. k = 0: linear model only
add("k0_factors", F(1, 2), w, [[] for _ in w], [([1, 2, 3], [F(1), F(2), F(4)])], k=0)

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "predict_kat.json")
    with open(out, "w") as fh:
        json.dump({"provenance": "builder-authored, exact rational arithmetic; NOT a reference fixture",
                   "cases": CASES}, fh, indent=1)
    print("wrote", out, len(CASES), "cases")
