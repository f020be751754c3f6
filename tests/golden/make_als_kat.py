"""Generates tests/golden/als_kat.json -- BUILDER-AUTHORED known answers for one ALS sweep
(reference: src/main/scala/io/edstud/spark/fm/lib/ALS.scala:15-75).

The reference ships no golden vectors and cannot run here, so the answers are computed in exact
rational arithmetic (fractions.Fraction) straight from the formulas in the Scala source --
e = predict - y, theta* = -(sum e*h - theta*sum h^2) / (lambda + sum h^2) with h = x for w and
h = x*q - x^2*v for V, ids ascending, residuals and q corrected after every accepted update --
independently of oracle/, and rounded once to the nearest double.  fp64 implementations agree
with them up to rounding (the test allows 1e-12 relative).

Run:  python tests/golden/make_als_kat.py
"""
import json
import os
from fractions import Fraction as F


def predict(w0, w, v, row, k):
    r = w0 + sum(w[i] * x for i, x in row)
    for f in range(k):
        t = [v[i][f] * x for i, x in row]
        r += F(1, 2) * (sum(t) ** 2 - sum(u * u for u in t))
    return r


def als_sweep(w0, w, v, rows, y, k, reg, quirks):
    n_slots = len(w)
    r0, rw, rv = reg
    e = [predict(w0, w, v, row, k) - yy for row, yy in zip(rows, y)]
    new = -(sum(e) - w0 * len(rows)) / (r0 + len(rows))
    if new != w0:   # the reference's lazily re-evaluated residuals carry the shift too (ALS.scala:27,31,142-144)
        e = [x + (new - w0) for x in e]
    w0 = new
    cols = {}
    for r, row in enumerate(rows):
        for i, x in row:
            cols.setdefault(i, []).append((r, x))
    id_end = n_slots - 1 if quirks else n_slots
    for i in range(id_end):
        if i in cols:
            sh = sum(x * x for _, x in cols[i])
            se = sum(e[r] * x for r, x in cols[i])
            new = -(se - w[i] * sh) / (rw + sh)
            for r, x in cols[i]:
                e[r] += x * (new - w[i])
            w[i] = new
    for f in range(k):
        q = [sum(v[i][f] * x for i, x in row) for row in rows]
        for i in range(id_end):
            if i in cols:
                old = v[i][f]
                h = [(r, x * q[r] - x * x * old) for r, x in cols[i]]
                sh = sum(hh * hh for _, hh in h)
                se = sum(e[r] * hh for r, hh in h)
                new = -(se - old * sh) / (rv + sh)
                for r, hh in h:
                    e[r] += hh * (new - old)
                for r, x in cols[i]:
                    q[r] += x * (new - old)
                v[i][f] = new
    return w0, w, v, e


def main():
    k, n_slots = 2, 6
    rows = [[(0, F(1)), (3, F(1, 2))], [(1, F(2)), (3, F(1)), (5, F(-1))], [(0, F(-1, 2)), (2, F(1))],
            [(2, F(3, 2)), (4, F(1)), (5, F(1, 4))], [(1, F(1)), (4, F(-2))], [(0, F(1)), (1, F(1)), (2, F(1))],
            [(3, F(2)), (5, F(1))]]
    y = [F(3, 4), F(-1, 2), F(1), F(5, 4), F(-3, 2), F(2), F(1, 4)]
    w0 = F(1, 8)
    w = [F(0), F(1, 4), F(-1, 4), F(1, 2), F(0), F(1, 8)]
    v = [[F(1, 4), F(-1, 8)], [F(1, 8), F(1, 2)], [F(-1, 2), F(1, 4)], [F(1, 16), F(-1, 4)],
         [F(3, 8), F(1, 8)], [F(-1, 8), F(-3, 8)]]
    reg = (F(0), F(1, 8), F(1, 4))
    cases = []
    for quirks in (False, True):
        a_w0, a_w, a_v, a_e = als_sweep(w0, list(w), [list(r) for r in v], rows, y, k, reg, quirks)
        cases.append({"ref_quirks": quirks, "w0": float(a_w0), "w": [float(x) for x in a_w],
                      "v": [[float(x) for x in r] for r in a_v], "e": [float(x) for x in a_e]})
    out = {"k": k, "n_slots": n_slots, "reg": [float(x) for x in reg],
           "rows": [[[i, float(x)] for i, x in row] for row in rows], "y": [float(x) for x in y],
           "w0": float(w0), "w": [float(x) for x in w], "v": [[float(x) for x in r] for r in v],
           "cases": cases}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "als_kat.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
        fh.write("\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
