"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/sparkfm_b200.h
declares; host-side logic of the boundary (LibFM text ingest/export, CSR packing, sampler,
synthetic generators, the Python mirror's DataSet) -- bit-exact against the oracle's pure-Python
restatement of the Scala (fm/FMUtils.scala:23-74, DataSet.scala:23-29)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import capi as ocapi, fm_numpy as fn
from sparkfm_b200 import DataSet, LabeledPoint, SparseVector, _lib, format_libfm, parse_libfm, \
    sample_rows, synth
from sparkfm_b200._lib import SFM_ERR_CUDA, SfmConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sparkfm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sfm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 40
    L = ctypes.CDLL(_lib.SO_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} is declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.load().sfm_abi_version() == 1
    assert _lib.load().sfm_status_string(-5).decode().startswith("feature index")


def test_config_struct_layout_matches_header():
    # int32 x6, int64, float x5, int32, uint64  -> 64 bytes, n_slots at 24, seed at 56
    assert ctypes.sizeof(SfmConfig) == 64
    assert SfmConfig.n_slots.offset == 24 and SfmConfig.sampler_seed.offset == 56


def test_no_cpu_fallback_without_a_gpu():
    L = _lib.load()
    if L.sfm_device_count() > 0:
        pytest.skip("a GPU is visible")
    cfg = SfmConfig(1, 0, 8, 1, 1, 0, 100, 0, 0, 0, 0.1, 1.0, 0, 42)
    h = ctypes.c_void_p()
    assert L.sfm_create(ctypes.byref(cfg), ctypes.byref(h)) == SFM_ERR_CUDA
    assert not h.value
    from sparkfm_b200 import Handle
    with pytest.raises(_lib.SfmError):
        Handle(100, 8)


# ------------------------------------------------------------------------------ LibFM text
TEXT = b"""# a comment line
1 3:1.5 7:2 1:0.25

   -1 0:1   5:1e-2  5:3
0.5
+2.5e0 2147483647:1 4:-0d 9:7f\r
\t 3 1:1 \t
"""


def test_parser_matches_scala_rules_bit_exact():
    """Index verbatim (no shift, FMUtils:32), order + duplicates kept, trim, '#', repeated spaces,
    CRLF, Java number grammar (d/f suffix), a row without features when numFeatures is given."""
    lab, rp, idx, val, d = parse_libfm(TEXT, num_features=100)
    olab, orp, oidx, oval, od = fn.parse_libfm_lines(TEXT.decode().split("\n"), 100)
    assert d == od == 100
    assert lab.tolist() == olab.tolist() == [1.0, -1.0, 0.5, 2.5, 3.0]
    assert rp.tolist() == orp.tolist() == [0, 3, 6, 6, 9, 10]
    assert idx.tolist() == oidx.tolist() == [3, 7, 1, 0, 5, 5, 2147483647, 4, 9, 1]
    assert val.tobytes() == oval.tobytes()
    assert val.tolist()[:6] == [1.5, 2.0, 0.25, 1.0, 0.01, 3.0] and np.signbit(val[7])


def test_parser_dimension_inference_and_errors():
    lab, rp, idx, val, d = parse_libfm(b"1 4:1 9:2\n0 2:1\n")
    assert d == 9                                    # max index; vector length is d+1 (:50)
    for bad, line in ((b"1 4:1\n0\n", 0),            # indices.max on an empty row (:45)
                      (b"", 0),                      # reduce on an empty collection
                      (b"1 4:1\nx 2:1\n", 2),        # label.toDouble
                      (b"1 4\n", 1),                 # indexAndValue(1) missing
                      (b"1 4:\n", 1),                # "4:".split(':') has one element
                      (b"1 4.0:1\n", 1),             # toInt
                      (b"1 :3\n", 1),
                      (b"1 4:1_0\n", 1),
                      (b"1\t4:1\n", 1),              # only ' ' separates tokens (:28)
                      (b"1 99999999999:1\n", 1)):
        with pytest.raises(ValueError) as ei:
            parse_libfm(bad)
        assert f"line {line}" in str(ei.value), bad
        with pytest.raises(ValueError):
            fn.parse_libfm_lines(bad.decode().split("\n"))
    # "3:4:5" -> split(':') -> (3, 4): extra parts ignored, like the Scala
    _, _, idx, val, _ = parse_libfm(b"1 3:4:5\n")
    assert idx.tolist() == [3] and val.tolist() == [4.0]
    _, _, oidx, oval, _ = fn.parse_libfm_lines(["1 3:4:5"])
    assert oidx.tolist() == [3] and oval.tolist() == [4.0]


def test_parser_config1_text_roundtrip_is_bit_exact():
    """BASELINE config 1 is LIBSVM-format: generator -> text -> C++ parser == generator arrays
    == pure-Python parser, bit for bit (labels, row_ptr, indices, values)."""
    row_ptr, idx, val, label = synth.classification_c1(3000, 10_000, 20, 8)
    text = synth.to_libfm_text(row_ptr, idx, val, label).encode()
    lab, rp, i2, v2, d = parse_libfm(text)
    assert np.array_equal(rp, row_ptr) and np.array_equal(i2, idx)
    assert np.array_equal(v2, val.astype(np.float64)) and np.array_equal(lab, label.astype(np.float64))
    assert d == int(idx.max())
    olab, orp, oidx, oval, od = fn.parse_libfm_lines(text.decode().split("\n"))
    assert np.array_equal(orp, rp) and np.array_equal(oidx, i2) and oval.tobytes() == v2.tobytes()
    assert od == d and olab.tobytes() == lab.tobytes()


def test_formatter_matches_decimalformat_rules():
    """fm/FMUtils.scala:58-74: index + 1, "#" for integral values, "#.###" HALF_EVEN otherwise
    (no leading zero: 0.5 -> ".5")."""
    out = format_libfm([1.0, -0.5, 2.0], [0, 3, 4, 4], np.array([0, 9, 4, 2], np.int32),
                       [1.0, 0.12345, -2.5, 0.0625]).decode()
    assert out == "1 1:1 10:.123 5:-2.5\n-.5 3:.062\n2\n"
    out = format_libfm([1234567.0], [0, 2], np.array([0, 1], np.int32), [0.0004, 1e3]).decode()
    assert out == "1234567 1:0 2:1000\n"


# ------------------------------------------------------------------------------ sampler
def test_c_abi_sampler_equals_oracle_sampler():
    for frac in (0.0, 0.05, float(np.float32(0.1)), 0.5, 1.0):
        for it in (1, 7):
            a = sample_rows(42, it, frac, 100, 9000)
            assert np.array_equal(a, ocapi.sample_rows(42, it, frac, 100, 9000))
            assert np.array_equal(a, fn.sample_rows(42, it, frac, 100, 9000))


def test_sampler_twins_agree_on_random_seeds_fractions_and_shards():
    """Property test over the three host twins of DESIGN.md 2.5 (C ABI, C oracle, numpy): random
    seeds, iterations, fp32 fractions from 2^-24 to 1 - 2^-24, shard ranges that start and end
    anywhere inside a 64-row block, far-away row numbers; and a random 3-way split of the range
    samples exactly the rows the whole range does."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(seed=st.integers(0, 2 ** 64 - 1), it=st.integers(1, 10 ** 9),
           frac=st.floats(2.0 ** -24, 1.0 - 2.0 ** -24, width=32),
           lo=st.integers(0, 2 ** 40), n=st.integers(0, 3000), cuts=st.tuples(st.floats(0, 1), st.floats(0, 1)))
    def check(seed, it, frac, lo, n, cuts):
        hi = lo + n
        a = sample_rows(seed, it, frac, lo, hi)
        assert np.array_equal(a, ocapi.sample_rows(seed, it, frac, lo, hi))
        assert np.array_equal(a, fn.sample_rows(seed, it, frac, lo, hi))
        assert np.all(np.diff(a) > 0) and (len(a) == 0 or (a[0] >= lo and a[-1] < hi))
        c1, c2 = sorted(lo + int(c * n) for c in cuts)
        parts = [sample_rows(seed, it, frac, x, y) for x, y in ((lo, c1), (c1, c2), (c2, hi))]
        assert np.array_equal(np.concatenate(parts), a)

    check()


# ------------------------------------------------------------------------------ generators
def test_ctr_generator_shape_and_determinism():
    card = synth.ctr_field_log2_cards(39)
    assert len(card) == 39 and card.max() == 20 and card.min() == 4
    cdf, off = synth.zipf_tables(card)
    assert cdf.dtype == np.uint32 and all(cdf[o + (1 << c) - 1] == 2 ** 32 - 1 for o, c in zip(off, card))
    a, la = synth.ctr_rows(1000, 3000, card, cdf, off, 1_000_000, 20260103)
    b, lb = synth.ctr_rows(0, 3000, card, cdf, off, 1_000_000, 20260103)
    assert np.array_equal(a, b[1000:]) and np.array_equal(la, lb[1000:])   # counter-based
    assert a.min() >= 0 and a.max() < 1_000_000
    assert 0.2 < lb.mean() < 0.3
    # Zipf: the most frequent id of a wide field covers far more than a uniform share
    col = b[:, 38]
    assert np.bincount(col).max() / len(col) > 0.05


def test_ragged_generator_has_no_duplicates_and_is_sorted():
    rp, idx, val = synth.ragged_rows(500, 50, 20, seed=1, max_nnz=40)
    for r in range(500):
        seg = idx[rp[r]:rp[r + 1]]
        assert len(seg) >= 1 and np.all(np.diff(seg) > 0)


# ------------------------------------------------------------------------------ DataSet
def test_dataset_size_dimension_and_packing():
    rows = [(1.0, SparseVector([5, 2, 2], [1.0, 0.0, 3.0], 10)),
            LabeledPoint(0.0, SparseVector([7], [2.0], 10))]
    ds = DataSet.from_rows(rows, "t")
    assert ds.size == 2 and not ds.isEmpty
    assert ds.dimension == 7                       # max index, not max+1 (DataSet.scala:27-29)
    assert ds.row_ptr.tolist() == [0, 3, 4]
    assert ds.idx.tolist() == [5, 2, 2, 7]         # stored order and duplicates preserved
    assert ds.val.tolist() == [1.0, 0.0, 3.0, 2.0] and ds.targets.tolist() == [1.0, 0.0]
    sv = ds.inputs(0)
    assert sv.index.tolist() == [5, 2, 2] and sv.used == 3 and sv.length == 10
    empty = DataSet([], [0], [], [])
    assert empty.isEmpty and empty.size == 0 and empty.dimension == 0
    with pytest.raises(ValueError):
        DataSet([1.0], [0, 0], [], []).dimension   # `_.index.max` on an empty row throws


def test_parser_multithreaded_chunks_are_seamless():
    """A text above the 1 MB-per-chunk threshold is cut at line boundaries and parsed by several
    threads (count pass + fill pass): same arrays as the pure-Python parser, CRLF / CR / comment /
    blank lines at arbitrary places, and the reported error line counts physical lines across
    chunks."""
    rng = np.random.default_rng(4)
    lines = []
    for r in range(60_000):
        if r % 997 == 0:
            lines.append("# comment %d" % r)
        if r % 1301 == 0:
            lines.append("   ")
        m = int(rng.integers(1, 12))
        toks = ["%d:%s" % (int(rng.integers(0, 5000)), repr(float(np.float32(rng.normal()))))
                for _ in range(m)]
        lines.append(("%d " % (r % 3 - 1)) + " ".join(toks) + ("  " if r % 5 == 0 else ""))
    seps = ["\n", "\r\n", "\r"]
    text = "".join(l + seps[i % 3] for i, l in enumerate(lines)).encode()
    assert len(text) > 3 * (1 << 20)
    lab, rp, idx, val, d = parse_libfm(text)
    py = "".join(l + "\n" for l in lines)
    olab, orp, oidx, oval, od = fn.parse_libfm_lines(py.split("\n"))
    assert d == od and np.array_equal(rp, orp) and np.array_equal(idx, oidx)
    assert val.tobytes() == oval.tobytes() and lab.tobytes() == olab.tobytes()
    # an error deep inside a later chunk: line number = physical line (1-based)
    bad_at = len(lines) * 3 // 4
    broken = list(lines)
    broken[bad_at] = "1 7:oops"
    text2 = "".join(l + "\n" for l in broken).encode()
    with pytest.raises(ValueError) as ei:
        parse_libfm(text2)
    assert f"line {bad_at + 1}" in str(ei.value)


def test_parser_number_grammar_fuzz_against_correctly_rounded_python():
    """30,000 random decimal tokens (long mantissas, huge / tiny exponents, halfway cases, subnormals,
    Java d/f suffixes, leading zeros, signs) through the C++ parser's exact fast path + from_chars
    fallback: every value must have the same 64 bits as Python's correctly rounded float()."""
    import random
    rnd = random.Random(20260105)
    special = ["1e400", "1e-400", "4.9e-324", "2.4703282292062327e-324", "2.4703282292062328e-324",
               "1.7976931348623157e308", "1.7976931348623159e308", "00012", "1E+5", "9007199254740993",
               "9007199254740992.5", "123456789012345678901234567890", "5e-324", "3e-324", "2e-324",
               "1.00000000000000011102230246251565404236316680908203125",
               "1.00000000000000011102230246251565404236316680908203124",
               "1.00000000000000011102230246251565404236316680908203126", "1e23", "8.5e22",
               "9.999999999999999e22", "8.41e21", "2.2250738585072011e-308", "7.2057594037927933e16"]
    digits = "0123456789"

    def tok():
        r = rnd.random()
        sign = rnd.choice(["", "-", "+"]) if rnd.random() < 0.5 else ""
        if r < 0.25:
            s = str(rnd.randint(0, 10 ** rnd.randint(1, 25)))
        elif r < 0.5:
            s = "%d.%s" % (rnd.randint(0, 10 ** rnd.randint(0, 20)),
                           "".join(rnd.choice(digits) for _ in range(rnd.randint(0, 25))))
        elif r < 0.75:
            s = "%d.%se%s%d" % (rnd.randint(0, 10 ** rnd.randint(0, 18)),
                                "".join(rnd.choice(digits) for _ in range(rnd.randint(1, 20))),
                                rnd.choice(["", "-", "+"]), rnd.randint(0, 330))
        elif r < 0.8:
            s = "." + "".join(rnd.choice(digits) for _ in range(rnd.randint(1, 30)))
        elif r < 0.88:
            s = repr(abs(rnd.uniform(-1, 1)) * 10.0 ** rnd.randint(-320, 308))
        elif r < 0.94:
            s = str(rnd.randint(0, 999)) + rnd.choice(["d", "D", "f", "F", ".0d", ".5f", "e3f", "E-2D"])
        else:
            s = rnd.choice(special)
        return sign + s

    toks = [tok() for _ in range(30_000)]
    lines = ["1 %d:%s" % (i % 100, t) for i, t in enumerate(toks)]
    _, _, _, val, _ = parse_libfm(("\n".join(lines) + "\n").encode(), num_features=200)
    _, _, _, oval, _ = fn.parse_libfm_lines(lines, 200)
    diff = np.nonzero(val.view(np.uint64) != oval.view(np.uint64))[0]
    assert len(diff) == 0, [(toks[j], val[j].hex(), oval[j].hex()) for j in diff[:5]]


def test_parser_fast_token_path_edges():
    """The one-scan "<digits>:<digits[.digits]>" path and its hand-over to the general grammar:
    digit-count limits (10 for the id, 15 for the value), leading zeros, bare '.', '5.' and '.5',
    values whose nearest double needs the correctly rounded division."""
    toks = ["0:0", "007:1.", "5:.5", "5:00000000000000.1", "5:0.000000000000001", "5:123456789012345",
            "5:1234567890123456", "5:1.23456789012345", "5:1.234567890123456", "5:0.1", "5:0.3",
            "5:4.35", "5:1e5", "2147483647:1", "0000000005:2", "00000000005:2", "5:0.000000", "5:000",
            "5:99999999999999.9", "5:9.99999999999999", "5:1.0", "9:3.14159", "5:2.5d", "5:7f",
            "5:0.1:9", "+5:1", "5:+1", "5:-0.0", "5:1.7976931348623157e308"]
    lines = ["1 " + t for t in toks] + ["0.5 " + " ".join(toks[:12])]
    lab, rp, idx, val, d = parse_libfm(("\n".join(lines) + "\n").encode(), num_features=10)
    olab, orp, oidx, oval, od = fn.parse_libfm_lines(lines, 10)
    assert np.array_equal(rp, orp) and np.array_equal(idx, oidx) and lab.tobytes() == olab.tobytes()
    assert val.tobytes() == oval.tobytes()
    for bad in ("1 5:.", "1 5:", "1 :5", "1 5:1x", "1 2147483648:1", "1 5:1..2", "1 5 :1", "1 5: 1"):
        with pytest.raises(ValueError):
            parse_libfm((bad + "\n").encode())
        with pytest.raises(ValueError):
            fn.parse_libfm_lines([bad])


def test_pack_onehot_matches_bitstream_definition():
    """sfm_pack_onehot: entry e occupies bits [e*id_bits, (e+1)*id_bits) of the little-endian
    uint32 stream, labels one bit per row; ids wider than id_bits are SFM_ERR_INDEX."""
    from sparkfm_b200 import pack_onehot
    from sparkfm_b200._lib import SFM_ERR_INDEX, SfmError
    rng = np.random.default_rng(0)
    for m, bits, n in [(39, 20, 1000), (13, 15, 77), (3, 32, 50), (64, 1, 33), (39, 20, 100_000), (1, 7, 0)]:
        lim = (1 << bits) if bits < 32 else (1 << 31)
        idx = rng.integers(0, lim, size=(n, m)).astype(np.int32)
        lab = (rng.random(n) < 0.3).astype(np.float32)
        pk, lb = pack_onehot(idx, lab, m, bits)
        flat = idx.reshape(-1).astype(np.uint64)
        stream = np.zeros(len(pk) * 32, dtype=np.uint8)
        b = ((flat[:, None] >> np.arange(bits, dtype=np.uint64)) & 1).astype(np.uint8).reshape(-1)
        stream[:len(b)] = b
        want = np.packbits(stream.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)
        assert np.array_equal(pk, want), (m, bits, n)
        wl = np.zeros(len(lb) * 32, dtype=np.uint8)
        wl[:n] = lab > 0
        assert np.array_equal(lb, np.packbits(wl.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1))
    with pytest.raises(SfmError) as ei:
        pack_onehot(np.array([[1, 2, 1 << 12]], np.int32), np.array([1.0], np.float32), 3, 12)
    assert ei.value.status == SFM_ERR_INDEX


def test_jni_glue_binds_every_export_and_matches_the_scala_declarations():
    """jni/sfm_jni.c (the binding for the reference's Scala 2.10 / Java 7-8 toolchain): valid C
    against a stand-in jni.h (no JDK in the image), one call of every export the header declares,
    and one `@native def` in SfmJni.scala per JNI function (and vice versa)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "jni", "sfm_jni.c")
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-Wno-unused-parameter",
                           "-DSFM_JNI_COMPILE_CHECK_ONLY", "-I" + os.path.join(root, "jni", "compile_check"),
                           "-I" + os.path.join(root, "include"), src])
    text = open(src).read()
    header = open(os.path.join(root, "include", "sparkfm_b200.h")).read()
    declared = set(re.findall(r"\b(sfm_[a-z0-9_]+)\s*\(", header))
    called = set(re.findall(r"\b(sfm_[a-z0-9_]+)\s*\(", text))
    assert declared <= called, sorted(declared - called)
    c_names = set(re.findall(r"FN\((\w+)\)", text)) - {"name"}
    scala = open(os.path.join(root, "scala", "io", "edstud", "spark", "fm", "gpu", "SfmJni.scala")).read()
    s_names = set(re.findall(r"@native def (\w+)", scala))
    assert c_names == s_names, (sorted(c_names - s_names), sorted(s_names - c_names))


def test_model_file_header_is_validated_before_anything_is_allocated(tmp_path):
    """sfm_load treats the header as untrusted: absurd sizes, a payload that does not match the
    header, bad flags -> SFM_ERR_IO, never an exception through the C ABI (no GPU needed: the
    checks come before the handle is created)."""
    import struct
    from sparkfm_b200._lib import SFM_ERR_IO
    L = _lib.load()

    def header(n_slots, k, task=0, k0=1, k1=1, version=1):
        return struct.pack("@8sIiiiiqfffffIQ", b"SFMB200\0", version, task, k, k0, k1, n_slots,
                           0.0, 0.0, 0.0, 0.1, 1.0, 0, 42)

    assert len(header(1, 1)) == 72
    cases = {
        "huge": header(1 << 40, 128),                       # would need 2^49 bytes
        "negative": header(-5, 4),
        "truncated": header(1000, 8) + b"\0" * 100,       # payload shorter than the header says
        "trailing": header(2, 1) + b"\0" * (4 * (1 + 2 * 2)) + b"junk",
        "bad_task": header(2, 1, task=7) + b"\0" * (4 * (1 + 2 * 2)),
        "bad_version": header(2, 1, version=9) + b"\0" * (4 * (1 + 2 * 2)),
    }
    for name, blob in cases.items():
        p = tmp_path / f"{name}.sfm"
        p.write_bytes(blob)
        out = ctypes.c_void_p()
        assert L.sfm_load(str(p).encode(), 0, ctypes.byref(out)) == SFM_ERR_IO, name
        assert not out.value
