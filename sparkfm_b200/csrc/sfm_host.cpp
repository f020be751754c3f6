// sfm_host.cpp -- host-side pieces of the hot path's boundary: LibFM text ingest / export
// (fm/FMUtils.scala:23-74 of the reference), the mini-batch sampler and the seeded model
// initialiser.  No CUDA here.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <charconv>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include "sfm_common.h"

namespace sfm {

uint64_t mix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// DESIGN.md 2.1.  Box-Muller in fp64 with glibc's log/cos, rounded once to fp32.
void init_gaussian_f32(float* v, int64_t first, int64_t count, double mean, double stdev,
                       uint64_t seed) {
    const uint64_t s = mix64(seed);
    const double two_pi = 6.283185307179586476925286766559;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 32) nt = 32;
    if (count < (1 << 16)) nt = 1;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) {
        const int64_t lo = count * t / nt, hi = count * (t + 1) / nt;
        th.emplace_back([=] {
            for (int64_t i = lo; i < hi; ++i) {
                const uint64_t e = (uint64_t)(first + i);
                const double u1 = (double)((mix64(s + 2ULL * e) >> 11) + 1ULL) * 0x1.0p-53;
                const double u2 = (double)(mix64(s + 2ULL * e + 1ULL) >> 11) * 0x1.0p-53;
                const double z = sqrt(-2.0 * log(u1)) * cos(two_pi * u2);
                v[i] = (float)(mean + stdev * z);
            }
        });
    }
    for (auto& x : th) x.join();
}

// ---- java.lang number grammar ------------------------------------------------------------
// Double.parseDouble: chars <= ' ' trimmed at both ends, [sign] (NaN | Infinity | decimal
// [fFdD]).  Hex float literals are not supported (documented deviation).
static bool java_double(const char* b, const char* e, double* out) {
    while (b < e && (unsigned char)*b <= ' ') ++b;
    while (e > b && (unsigned char)e[-1] <= ' ') --e;
    if (b >= e) return false;
    const char* p = b;
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = *p == '-'; ++p; }
    const size_t rest = (size_t)(e - p);
    if (rest == 3 && memcmp(p, "NaN", 3) == 0) { *out = NAN; return true; }
    if (rest == 8 && memcmp(p, "Infinity", 8) == 0) { *out = neg ? -INFINITY : INFINITY; return true; }
    const char* q = p;
    int digits = 0;
    while (q < e && *q >= '0' && *q <= '9') { ++q; ++digits; }
    if (q < e && *q == '.') {
        ++q;
        while (q < e && *q >= '0' && *q <= '9') { ++q; ++digits; }
    }
    if (digits == 0) return false;
    if (q < e && (*q == 'e' || *q == 'E')) {
        ++q;
        if (q < e && (*q == '+' || *q == '-')) ++q;
        int ed = 0;
        while (q < e && *q >= '0' && *q <= '9') { ++q; ++ed; }
        if (ed == 0) return false;
    }
    const char* num_end = q;
    if (q < e && (*q == 'f' || *q == 'F' || *q == 'd' || *q == 'D')) ++q;
    if (q != e) return false;
    {
        // Clinger's fast path: <= 15 significant digits and |10-exponent| <= 22 -> the mantissa
        // and the power of ten are exact doubles, one correctly rounded multiply / divide gives the
        // same bits as strtod
        static const double P10[] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,
                                     1e8,  1e9,  1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                     1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
        uint64_t mant = 0;
        int nd = 0, frac = 0, ex = 0;
        bool seen_dot = false, exneg = false, ok = true;
        const char* z = p;
        for (; z < num_end; ++z) {
            const char c = *z;
            if (c >= '0' && c <= '9') {
                if (nd > 0 || c != '0') {
                    if (++nd > 15) { ok = false; break; }
                    mant = mant * 10 + (uint64_t)(c - '0');
                }
                if (seen_dot) ++frac;
            } else if (c == '.') {
                seen_dot = true;
            } else {  // exponent part
                ++z;
                if (z < num_end && (*z == '+' || *z == '-')) { exneg = *z == '-'; ++z; }
                for (; z < num_end; ++z) {
                    ex = ex * 10 + (*z - '0');
                    if (ex > 400) { ok = false; break; }
                }
                break;
            }
        }
        if (ok) {
            const int e10 = (exneg ? -ex : ex) - frac;
            if (e10 >= -22 && e10 <= 22) {
                double v = (double)mant;
                v = e10 < 0 ? v / P10[-e10] : v * P10[e10];
                *out = neg ? -v : v;
                return true;
            }
        }
    }
    // general case: std::from_chars is correctly rounded (same bits as strtod) and locale-free
    double v = 0.0;
    const std::from_chars_result r = std::from_chars(p, num_end, v, std::chars_format::general);
    if (r.ec == std::errc::result_out_of_range) {
        // overflow -> +-Infinity, underflow -> +-0 like Double.parseDouble; let strtod decide
        std::string tmp(p, num_end);
        v = strtod(tmp.c_str(), nullptr);
    } else if (r.ec != std::errc() || r.ptr != num_end) {
        return false;
    }
    *out = neg ? -v : v;
    return true;
}

// Integer.parseInt: [sign] digits, no whitespace, 32-bit range.
static bool java_int(const char* b, const char* e, int32_t* out) {
    if (b >= e) return false;
    bool neg = false;
    if (*b == '+' || *b == '-') { neg = *b == '-'; ++b; }
    if (b >= e) return false;
    int64_t v = 0;
    for (; b < e; ++b) {
        if (*b < '0' || *b > '9') return false;
        v = v * 10 + (*b - '0');
        if (v > 2147483648LL) return false;
    }
    if (neg) v = -v;
    if (v > 2147483647LL || v < -2147483648LL) return false;
    *out = (int32_t)v;
    return true;
}

}  // namespace sfm

using namespace sfm;

namespace {

// First '\n' or '\r' at or after p (or end): two library scans instead of a byte loop.
static inline const char* line_end(const char* p, const char* end) {
    const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
    const char* stop = nl ? nl : end;
    const char* cr = static_cast<const char*>(memchr(p, '\r', (size_t)(stop - p)));
    return cr ? cr : stop;
}

static const double kPow10[] = {1e0, 1e1, 1e2,  1e3,  1e4,  1e5,  1e6,  1e7,
                                1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};

struct ChunkStat {
    int64_t rows = 0, ents = 0, lines = 0, err_line = -1;   // err_line: 1-based inside the chunk
    int32_t max_index = INT32_MIN;
    bool empty_row = false;
};

// Parses the physical lines of [p, end) (a chunk that starts at a line start and ends after a
// line terminator or at the end of the text).  Counting pass: out arrays null.  Filling pass:
// rows / entries are written starting at row0 / ent0.
static void parse_chunk(const char* p, const char* end, ChunkStat* st, double* label,
                        int64_t* row_ptr, int32_t* idx, double* val, int64_t row0, int64_t ent0) {
    int64_t rows = 0, ents = 0, line_no = 0;
    int32_t max_index = INT32_MIN;
    bool empty_row = false;
    while (p < end) {
        // one physical line: terminated by \n, \r\n or \r (Hadoop LineReader, used by sc.textFile)
        const char* le = line_end(p, end);
        const char* next = le;
        if (next < end) next += (*next == '\r' && next + 1 < end && next[1] == '\n') ? 2 : 1;
        ++line_no;
        const char* b = p;
        const char* e = le;
        p = next;
        while (b < e && (unsigned char)*b <= ' ') ++b;        // .map(_.trim)          FMUtils:25
        while (e > b && (unsigned char)e[-1] <= ' ') --e;
        if (b == e || *b == '#') continue;                    // isEmpty || startsWith("#") :26
        // items = line.split(' ')                                                     :28
        const char* t = b;
        const char* te = t;
        while (te < e && *te != ' ') ++te;
        double lab;
        if (!java_double(t, te, &lab)) {                      // items.head.toDouble         :29
            st->err_line = line_no;
            return;
        }
        if (label) label[row0 + rows] = lab;
        int64_t row_ents = 0;
        t = te;
        while (t < e) {
            ++t;  // skip the separating space
            {
                // fast path, one scan: "<digits>:<digits>" followed by a space or the line end --
                // the shape of every token of one-hot / count data.  Anything else (signs, '.',
                // exponents, suffixes, extra ':' parts, out-of-range ids) takes the general path
                // below, which restates the Scala's split / toInt / toDouble literally.
                const char* z = t;
                uint64_t id64 = 0;
                while (z < e && (unsigned)(*z - '0') <= 9u && z - t < 10) id64 = id64 * 10 + (uint64_t)(*z++ - '0');
                if (z > t && z < e && *z == ':' && id64 <= 2147483647ull) {
                    // value: digits [. digits], at most 15 digits in all: the mantissa and the
                    // power of ten are exact doubles, one division is correctly rounded (Clinger)
                    const char* y = z + 1;
                    uint64_t mant = 0;
                    int nd = 0, frac = 0;
                    while (y < e && (unsigned)(*y - '0') <= 9u && nd < 15) { mant = mant * 10 + (uint64_t)(*y++ - '0'); ++nd; }
                    if (y < e && *y == '.') {
                        ++y;
                        while (y < e && (unsigned)(*y - '0') <= 9u && nd < 15) { mant = mant * 10 + (uint64_t)(*y++ - '0'); ++nd; ++frac; }
                    }
                    if (nd > 0 && (y == e || *y == ' ')) {
                        if (idx) idx[ent0 + ents] = (int32_t)id64;
                        if (val) val[ent0 + ents] = frac ? (double)mant / kPow10[frac] : (double)mant;
                        if ((int32_t)id64 > max_index) max_index = (int32_t)id64;
                        ++ents;
                        ++row_ents;
                        t = y;
                        continue;
                    }
                }
            }
            te = t;
            while (te < e && *te != ' ') ++te;
            if (te == t) continue;                            // .filter(_.nonEmpty)         :30
            // item.split(':') -> indexAndValue(0).toInt, indexAndValue(1).toDouble  :31-34
            const char* c0 = t;
            while (c0 < te && *c0 != ':') ++c0;
            if (c0 == te) {                                   // no ':' -> indexAndValue(1) throws
                st->err_line = line_no;
                return;
            }
            const char* v0 = c0 + 1;
            const char* v1 = v0;
            while (v1 < te && *v1 != ':') ++v1;               // extra ":..." parts are ignored
            int32_t id;
            double x;
            // "3:" -> split drops the trailing empty string -> only one part -> throws
            bool rest_empty = true;
            for (const char* z = v0; z < te; ++z) if (*z != ':') { rest_empty = false; break; }
            bool ok_tok = !rest_empty && java_int(t, c0, &id);
            if (ok_tok) {
                // "id:1", "id:37": up to 15 plain digits are an exact double, no grammar walk needed
                const size_t vl = (size_t)(v1 - v0);
                uint64_t mant = 0;
                size_t z = 0;
                if (vl >= 1 && vl <= 15)
                    for (; z < vl && (unsigned)(v0[z] - '0') <= 9u; ++z) mant = mant * 10 + (uint64_t)(v0[z] - '0');
                if (z == vl && vl >= 1 && vl <= 15) x = (double)mant;
                else ok_tok = java_double(v0, v1, &x);
            }
            if (!ok_tok) {
                st->err_line = line_no;
                return;
            }
            if (idx) idx[ent0 + ents] = id;
            if (val) val[ent0 + ents] = x;
            if (id > max_index) max_index = id;
            ++ents;
            ++row_ents;
            t = te;
        }
        if (row_ents == 0) empty_row = true;
        ++rows;
        if (row_ptr) row_ptr[row0 + rows] = ent0 + ents;
    }
    st->rows = rows;
    st->ents = ents;
    st->lines = line_no;
    st->max_index = max_index;
    st->empty_row = empty_row;
}

// Counting pass of a filling call: rows, entries and physical lines of a chunk from the line /
// token structure alone (no number is parsed; malformed input is reported by the filling pass,
// which walks the same structure).
static void count_chunk(const char* p, const char* end, ChunkStat* st) {
    int64_t rows = 0, ents = 0, line_no = 0;
    while (p < end) {
        const char* le = line_end(p, end);
        const char* next = le;
        if (next < end) next += (*next == '\r' && next + 1 < end && next[1] == '\n') ? 2 : 1;
        ++line_no;
        const char* b = p;
        const char* e = le;
        p = next;
        while (b < e && (unsigned char)*b <= ' ') ++b;
        while (e > b && (unsigned char)e[-1] <= ' ') --e;
        if (b == e || *b == '#') continue;
        ++rows;
        const char* t = static_cast<const char*>(memchr(b, ' ', (size_t)(e - b)));   // the label
        if (!t) continue;
        // non-empty items = characters that are not a space and follow a space (branch-free)
        int64_t c = 0;
        for (const char* z = t + 1; z < e; ++z) c += (int64_t)((z[-1] == ' ') & (z[0] != ' '));
        ents += c;
    }
    st->rows = rows;
    st->ents = ents;
    st->lines = line_no;
}

}  // namespace

// Multi-threaded: the text is cut at line boundaries into one chunk per thread; a counting pass
// gives every chunk its first row / entry, a filling pass writes the arrays.  The result is
// identical to a sequential parse (text ingest at speed).
extern "C" int32_t sfm_parse_libfm(const char* text, uint64_t len, int32_t num_features,
                                   int64_t* n_rows, int64_t* nnz, int32_t* dimension_out,
                                   double* label, int64_t* row_ptr, int32_t* idx, double* val,
                                   int64_t* err_line) {
    if (!text && len) return SFM_ERR_ARG;
    if (!n_rows || !nnz) return SFM_ERR_ARG;
    const bool fill = idx != nullptr || label != nullptr || row_ptr != nullptr || val != nullptr;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    const uint64_t min_chunk = 1u << 20;
    if (len / min_chunk < nt) nt = (unsigned)(len / min_chunk);
    if (nt < 1) nt = 1;
    // chunk boundaries: advance to just after the next line terminator
    std::vector<const char*> cut(nt + 1);
    const char* end = text + len;
    cut[0] = text;
    cut[nt] = end;
    for (unsigned t = 1; t < nt; ++t) {
        const char* p = text + len * t / nt;
        if (p < cut[t - 1]) p = cut[t - 1];
        while (p < end && *p != '\n' && *p != '\r') ++p;
        if (p < end) p += (*p == '\r' && p + 1 < end && p[1] == '\n') ? 2 : 1;
        cut[t] = p;
    }
    std::vector<ChunkStat> st(nt);
    auto run = [&](bool filling, const std::vector<int64_t>* row0, const std::vector<int64_t>* ent0) {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t] {
                if (filling)
                    parse_chunk(cut[t], cut[t + 1], &st[t], label, row_ptr, idx, val, (*row0)[t],
                                (*ent0)[t]);
                else if (fill)
                    count_chunk(cut[t], cut[t + 1], &st[t]);      // sizes only; the fill validates
                else
                    parse_chunk(cut[t], cut[t + 1], &st[t], nullptr, nullptr, nullptr, nullptr, 0, 0);
            });
        for (auto& x : th) x.join();
    };
    run(false, nullptr, nullptr);
    int64_t rows = 0, ents = 0, lines_before = 0;
    int32_t max_index = INT32_MIN;
    bool empty_row = false;
    std::vector<int64_t> row0(nt), ent0(nt);
    for (unsigned t = 0; t < nt; ++t) {
        if (st[t].err_line >= 0) {  // the first failing chunk holds the first failing line
            if (err_line) *err_line = lines_before + st[t].err_line;
            return SFM_ERR_IO;
        }
        row0[t] = rows;
        ent0[t] = ents;
        rows += st[t].rows;
        ents += st[t].ents;
        lines_before += st[t].lines;
        if (st[t].max_index > max_index) max_index = st[t].max_index;
        empty_row = empty_row || st[t].empty_row;
    }
    if (fill) {
        if (row_ptr) row_ptr[0] = 0;
        run(true, &row0, &ent0);
        lines_before = 0;
        max_index = INT32_MIN;
        empty_row = false;
        for (unsigned t = 0; t < nt; ++t) {
            if (st[t].err_line >= 0) {
                if (err_line) *err_line = lines_before + st[t].err_line;
                return SFM_ERR_IO;
            }
            lines_before += st[t].lines;
            if (st[t].max_index > max_index) max_index = st[t].max_index;
            empty_row = empty_row || st[t].empty_row;
        }
    }
    *n_rows = rows;
    *nnz = ents;
    int32_t d;
    if (num_features > 0) {
        d = num_features;                                     // :40-41
    } else {
        // parsed.map(indices.max).reduce(math.max): throws on an empty RDD and on a row
        // without features                                                         :43-46
        if (rows == 0 || empty_row) {
            if (err_line) *err_line = 0;
            return SFM_ERR_IO;
        }
        d = max_index;
    }
    if (dimension_out) *dimension_out = d;
    return SFM_OK;
}

// DecimalFormat("#") / DecimalFormat("#.###"), RoundingMode.HALF_EVEN  (FMUtils.scala:71-74)
static void minimize_string(double v, std::string& out) {
    char buf[400];
    if (v != v) { out += "NaN"; return; }
    if (isinf(v)) { out += v < 0 ? "-Infinity" : "Infinity"; return; }
    if (v == floor(v)) {
        snprintf(buf, sizeof buf, "%.0f", v);
        if (strcmp(buf, "-0") == 0 || strcmp(buf, "0") == 0) { out += (signbit(v) ? "-0" : "0"); return; }
        out += buf;
        return;
    }
    snprintf(buf, sizeof buf, "%.3f", v);  // glibc rounds the exact binary value half-to-even
    std::string s(buf);
    while (!s.empty() && s.back() == '0') s.pop_back();
    if (!s.empty() && s.back() == '.') s.pop_back();
    // "#.###" has no mandatory integer digit: 0.5 -> ".5", -0.25 -> "-.25", 0.0004 -> "0"
    bool neg = !s.empty() && s[0] == '-';
    std::string body = neg ? s.substr(1) : s;
    if (body == "0" || body.empty()) { out += neg ? "-0" : "0"; return; }
    if (body.size() > 1 && body[0] == '0' && body[1] == '.') body.erase(0, 1);
    if (neg) out += '-';
    out += body;
}

extern "C" int32_t sfm_format_libfm(const double* label, const int64_t* row_ptr,
                                    const int32_t* idx, const double* val, int64_t n_rows,
                                    char* out, uint64_t cap, uint64_t* needed) {
    if (n_rows < 0 || (n_rows > 0 && (!label || !row_ptr))) return SFM_ERR_ARG;
    std::string s;
    char buf[32];
    for (int64_t r = 0; r < n_rows; ++r) {
        minimize_string(label[r], s);                          // :60
        for (int64_t j = row_ptr[r]; j < row_ptr[r + 1]; ++j) {
            s += ' ';
            snprintf(buf, sizeof buf, "%lld:", (long long)idx[j] + 1);  // i + 1   :63
            s += buf;
            minimize_string(val[j], s);
        }
        s += '\n';
    }
    if (needed) *needed = s.size();
    if (out && cap >= s.size()) memcpy(out, s.data(), s.size());
    else if (out) return SFM_ERR_ARG;
    return SFM_OK;
}

extern "C" int32_t sfm_sample_rows(uint64_t seed, int64_t iter, double fraction, int64_t row_lo,
                                   int64_t row_hi, int64_t* out, int64_t* n_out) {
    if (!n_out || (row_hi > row_lo && !out)) return SFM_ERR_ARG;
    int64_t n = 0;
    if (fraction >= 1.0) {
        for (int64_t r = row_lo; r < row_hi; ++r) out[n++] = r;
    } else if (fraction > 0.0 && row_hi > row_lo) {
        // bit-sliced Bernoulli(thr / 2^53) over aligned blocks of 64 global rows (DESIGN.md 2.5;
        // device twin: BernoulliBlock in sfm_scan.cu)
        const uint64_t thr = (uint64_t)floor(fraction * 9007199254740992.0);
        const uint64_t key = mix64(seed + (uint64_t)iter);
        const uint64_t gamma = 0x9E3779B97F4A7C15ULL;
        int last = 0;
        if (thr) {
            int tz = 0;
            while (!((thr >> tz) & 1ULL)) ++tz;
            last = 53 - tz;
        }
        for (int64_t q = row_lo >> 6; thr && q <= ((row_hi - 1) >> 6); ++q) {
            uint64_t und = ~0ULL, hit = 0ULL, ctr = key + ((uint64_t)q << 6) * gamma;
            for (int i = 1; i <= last && und; ++i) {
                const uint64_t w = mix64(ctr);
                ctr += gamma;
                if ((thr >> (53 - i)) & 1ULL) {
                    hit |= und & ~w;
                    und &= w;
                } else {
                    und &= ~w;
                }
            }
            while (hit) {
                const int j = __builtin_ctzll(hit);
                hit &= hit - 1ULL;
                const int64_t r = (q << 6) + j;
                if (r >= row_lo && r < row_hi) out[n++] = r;
            }
        }
    }
    *n_out = n;
    return SFM_OK;
}

extern "C" int32_t sfm_partition_rows(uint64_t seed, int64_t n_parts, int64_t part, int64_t row_lo,
                                      int64_t row_hi, int64_t* out, int64_t* n_out) {
    if (!n_out || (row_hi > row_lo && !out) || n_parts < 1 || part < 0 || part >= n_parts)
        return SFM_ERR_ARG;
    const uint64_t key = mix64(seed);
    int64_t n = 0;
    for (int64_t r = row_lo; r < row_hi; ++r)
        if ((int64_t)((mix64(key ^ mix64((uint64_t)r)) >> 11) % (uint64_t)n_parts) == part) out[n++] = r;
    *n_out = n;
    return SFM_OK;
}

// Host-side packer of the compact one-hot staging format (include/sparkfm_b200.h,
// sfm_stage_onehot): entry e -> bits [e*id_bits, (e+1)*id_bits) of the little-endian uint32
// stream.  Threads own disjoint entry ranges whose first bit is word-aligned (ranges start at
// multiples of 32 entries), so no two threads touch the same word.
extern "C" int32_t sfm_pack_onehot(const int32_t* idx, const float* label, int64_t n_rows, int32_t m,
                                   int32_t id_bits, uint32_t* packed_idx, uint32_t* label_bits) {
    if (n_rows < 0 || m < 1 || id_bits < 1 || id_bits > 32 || (n_rows > 0 && (!idx || !packed_idx)))
        return SFM_ERR_ARG;
    const int64_t n = n_rows * (int64_t)m;
    const int64_t words = (int64_t)(((uint64_t)n * (uint64_t)id_bits + 31) / 32) + 1;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 32) nt = 32;
    if (n < (int64_t)1 << 16) nt = 1;
    const int64_t groups = (n + 31) / 32;   // 32 entries = id_bits whole words
    std::vector<int> bad(nt, 0);
    auto work = [&](unsigned t) {
        const int64_t e0 = groups * t / nt * 32, e1 = std::min<int64_t>(n, groups * (t + 1) / nt * 32);
        if (e0 >= e1) return;
        const uint64_t lim = id_bits >= 32 ? 0x100000000ull : (1ull << id_bits);
        uint32_t* out = packed_idx + (uint64_t)e0 * (uint64_t)id_bits / 32;
        uint64_t acc = 0;
        int fill = 0;
        for (int64_t e = e0; e < e1; ++e) {
            const int32_t v = idx[e];
            if (v < 0 || (uint64_t)(uint32_t)v >= lim) bad[t] = 1;
            acc |= ((uint64_t)(uint32_t)v & (lim - 1)) << fill;
            fill += id_bits;
            if (fill >= 32) {
                *out++ = (uint32_t)acc;
                acc >>= 32;
                fill -= 32;
            }
        }
        if (fill > 0) *out++ = (uint32_t)acc;
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    // the trailing slack word(s) the unpack kernel may touch
    const int64_t used = (int64_t)(((uint64_t)n * (uint64_t)id_bits + 31) / 32);
    for (int64_t w = used; w < words; ++w) packed_idx[w] = 0u;
    if (label && label_bits) {
        const int64_t lw = (n_rows + 31) / 32;
        for (int64_t w = 0; w < lw; ++w) {
            uint32_t bits = 0;
            const int64_t r1 = std::min<int64_t>(n_rows, (w + 1) * 32);
            for (int64_t r = w * 32; r < r1; ++r)
                if (label[r] > 0.f) bits |= 1u << (r & 31);
            label_bits[w] = bits;
        }
    }
    for (int b : bad)
        if (b) return SFM_ERR_INDEX;
    return SFM_OK;
}
