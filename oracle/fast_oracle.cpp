// CPU ORACLE, fast restatement (test infrastructure + the timed CPU baseline; never linked into or
// called by the product path -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference leg may load it).
//
// A single-threaded C++17 restatement of relp's two-phase simplex over exact, always-normalised
// big rationals, i.e. of `Carry<RationalBig, BasisInverseRows<RationalBig>>` driven by
// `phase_one::primal` / `phase_two::primal` (reference paths relative to src/algorithm/two_phase/):
//   - Rational  : stand-in for relp-num 0.1.13 `RationalBig` (un-vendored dependency): sign +
//                 magnitude numerator, positive denominator, gcd-normalised after every operation
//   - Carry     : tableau/inverse_maintenance/carry/mod.rs:46-66,295-349,561-604 with sparse rows as
//                 in carry/basis_inverse_rows.rs:43-99,147-195
//   - Tableau   : tableau/mod.rs:106-138,287-313; kinds tableau/kind/{artificial,non_artificial}
//   - rules     : strategy/pivot_rule.rs:86-305
//   - loops     : phase_one.rs:123-179,232-278; phase_two.rs:22-58; two_phase/mod.rs:25-109
// It mirrors oracle/relp_oracle.py function by function; tests/test_fast_oracle.py pins it to the
// Python oracle (which is pinned to the reference's golden fixtures).  The Rust reference itself
// cannot be built in this image (no cargo/rustc, nightly features, un-vendored crates).
//
// Storage: the provider's columns are held compactly (CSC with machine-integer numerators and optional
// denominators, plus an optional dense int8 block with implicit row indices), so that config 5
// (16384 x 32768 dense coefficients) fits; rationals are formed on the fly.  Dense columns are indexed
// directly in the sparse-row dots (what an O(1) `get` gives the reference).  Columns are independent in
// pricing, in the steepest-edge initialisation and in its update, so those loops run over OpenMP
// threads (fo_set_threads; results are exact rationals and do not depend on the thread count).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef __int128 i128;
typedef uint64_t u64;
typedef int64_t i64;

// ------------------------------------------------------------------------------------------------
// Magnitude: little-endian u64 limbs, no leading zeros (zero = empty)
// ------------------------------------------------------------------------------------------------
struct Mag {
    std::vector<u64> d;
    Mag() {}
    explicit Mag(u64 v) { if (v) d.push_back(v); }
    bool zero() const { return d.empty(); }
    size_t size() const { return d.size(); }
    void trim() { while (!d.empty() && d.back() == 0) d.pop_back(); }
    bool is_one() const { return d.size() == 1 && d[0] == 1; }
    bool fits64() const { return d.size() <= 1; }
    u64 low() const { return d.empty() ? 0 : d[0]; }
    size_t bits() const { return d.empty() ? 0 : 64 * (d.size() - 1) + (64 - __builtin_clzll(d.back())); }
};

static int cmp(const Mag& a, const Mag& b) {
    if (a.size() != b.size()) return a.size() < b.size() ? -1 : 1;
    for (size_t i = a.size(); i-- > 0;)
        if (a.d[i] != b.d[i]) return a.d[i] < b.d[i] ? -1 : 1;
    return 0;
}
static Mag add(const Mag& a, const Mag& b) {
    const Mag& x = a.size() >= b.size() ? a : b;
    const Mag& y = a.size() >= b.size() ? b : a;
    Mag r; r.d.resize(x.size() + 1);
    u64 c = 0;
    for (size_t i = 0; i < x.size(); ++i) {
        u128 s = (u128)x.d[i] + (i < y.size() ? y.d[i] : 0) + c;
        r.d[i] = (u64)s; c = (u64)(s >> 64);
    }
    r.d[x.size()] = c;
    r.trim();
    return r;
}
// a - b, requires a >= b
static Mag sub(const Mag& a, const Mag& b) {
    Mag r; r.d.resize(a.size());
    u64 bw = 0;
    for (size_t i = 0; i < a.size(); ++i) {
        u64 bi = i < b.size() ? b.d[i] : 0;
        u128 t = (u128)a.d[i] - bi - bw;
        r.d[i] = (u64)t; bw = (t >> 64) ? 1 : 0;
    }
    r.trim();
    return r;
}
static Mag mul(const Mag& a, const Mag& b) {
    Mag r;
    if (a.zero() || b.zero()) return r;
    r.d.assign(a.size() + b.size(), 0);
    for (size_t i = 0; i < a.size(); ++i) {
        u64 c = 0;
        for (size_t j = 0; j < b.size(); ++j) {
            u128 t = (u128)a.d[i] * b.d[j] + r.d[i + j] + c;
            r.d[i + j] = (u64)t; c = (u64)(t >> 64);
        }
        r.d[i + b.size()] = c;
    }
    r.trim();
    return r;
}
static Mag shl(const Mag& a, unsigned s) {
    if (a.zero()) return a;
    unsigned w = s / 64, b = s % 64;
    Mag r; r.d.assign(a.size() + w + 1, 0);
    for (size_t i = 0; i < a.size(); ++i) {
        r.d[i + w] |= a.d[i] << b;
        if (b) r.d[i + w + 1] |= a.d[i] >> (64 - b);
    }
    r.trim();
    return r;
}
static Mag shr(const Mag& a, unsigned s) {
    unsigned w = s / 64, b = s % 64;
    Mag r;
    if (w >= a.size()) return r;
    r.d.assign(a.size() - w, 0);
    for (size_t i = 0; i < r.d.size(); ++i) {
        r.d[i] = a.d[i + w] >> b;
        if (b && i + w + 1 < a.size()) r.d[i] |= a.d[i + w + 1] << (64 - b);
    }
    r.trim();
    return r;
}
// Knuth algorithm D.  q = a / b, r = a % b (b != 0)
static void divmod(const Mag& a, const Mag& b, Mag& q, Mag& r) {
    if (cmp(a, b) < 0) { q = Mag(); r = a; return; }
    if (b.size() == 1) {
        u64 dv = b.d[0]; u128 rem = 0;
        q.d.assign(a.size(), 0);
        for (size_t i = a.size(); i-- > 0;) {
            u128 cur = (rem << 64) | a.d[i];
            q.d[i] = (u64)(cur / dv); rem = cur % dv;
        }
        q.trim(); r = Mag((u64)rem);
        return;
    }
    unsigned s = __builtin_clzll(b.d.back());
    Mag v = shl(b, s), u = shl(a, s);
    size_t n = v.size(), m = a.size() + 1 - n;
    u.d.resize(a.size() + 1, 0);
    q.d.assign(m, 0);
    for (size_t jj = m; jj-- > 0;) {
        u128 num = ((u128)u.d[jj + n] << 64) | u.d[jj + n - 1];
        u128 qhat = num / v.d[n - 1], rhat = num % v.d[n - 1];
        while ((qhat >> 64) || (u128)(u64)qhat * v.d[n - 2] > ((rhat << 64) | u.d[jj + n - 2])) {
            --qhat; rhat += v.d[n - 1];
            if (rhat >> 64) break;
        }
        // multiply-subtract
        u64 borrow = 0, carry = 0;
        for (size_t i = 0; i < n; ++i) {
            u128 p = (u128)(u64)qhat * v.d[i] + carry;
            carry = (u64)(p >> 64);
            u128 t = (u128)u.d[i + jj] - (u64)p - borrow;
            u.d[i + jj] = (u64)t; borrow = (t >> 64) ? 1 : 0;
        }
        u128 t = (u128)u.d[jj + n] - carry - borrow;
        u.d[jj + n] = (u64)t;
        if (t >> 64) {   // add back
            --qhat;
            u64 c = 0;
            for (size_t i = 0; i < n; ++i) {
                u128 sum = (u128)u.d[i + jj] + v.d[i] + c;
                u.d[i + jj] = (u64)sum; c = (u64)(sum >> 64);
            }
            u.d[jj + n] += c;
        }
        q.d[jj] = (u64)qhat;
    }
    q.trim();
    u.trim();
    r = shr(u, s);
}
static u64 gcd64(u64 a, u64 b) { while (b) { u64 t = a % b; a = b; b = t; } return a; }
static Mag gcd(Mag a, Mag b) {
    // Euclid with Knuth division until both fit a word
    while (!b.zero()) {
        if (a.fits64() && b.fits64()) return Mag(gcd64(a.low(), b.low()));
        Mag q, r;
        divmod(a, b, q, r);
        a = std::move(b); b = std::move(r);
    }
    return a;
}
static Mag divexact(const Mag& a, const Mag& g) {
    if (g.is_one()) return a;
    Mag q, r;
    divmod(a, g, q, r);
    return q;
}

// ------------------------------------------------------------------------------------------------
// Rational: normalised, den > 0.  Small fast path: both parts below 2^63 stored in (sn, sd).
// ------------------------------------------------------------------------------------------------
struct Rational {
    bool small = true;
    i64 sn = 0; i64 sd = 1;      // valid when small
    bool neg = false; Mag n, dn; // valid when !small (dn > 0, gcd(n, dn) = 1)

    Rational() {}
    Rational(i64 num) : sn(num) {}
    Rational(i64 num, i64 den) {
        if (den < 0) { num = -num; den = -den; }
        u64 g = gcd64(num < 0 ? (u64)(-num) : (u64)num, (u64)den);
        if (g > 1) { num /= (i64)g; den /= (i64)g; }
        sn = num; sd = den;
    }
    bool is_zero() const { return small ? sn == 0 : n.zero(); }
    int sign() const { return small ? (sn > 0) - (sn < 0) : (n.zero() ? 0 : (neg ? -1 : 1)); }
    void to_big(bool& ng, Mag& nn, Mag& dd) const {
        if (small) { ng = sn < 0; nn = Mag(sn < 0 ? (u64)(-(i128)sn) : (u64)sn); dd = Mag((u64)sd); }
        else { ng = neg; nn = n; dd = dn; }
    }
    static Rational from_big(bool ng, Mag nn, Mag dd, bool reduce = true) {
        Rational r;
        if (nn.zero()) return r;
        if (reduce) {
            Mag g = gcd(nn, dd);
            if (!g.is_one()) { nn = divexact(nn, g); dd = divexact(dd, g); }
        }
        if (nn.fits64() && dd.fits64() && nn.low() < (1ull << 62) && dd.low() < (1ull << 62)) {
            r.sn = ng ? -(i64)nn.low() : (i64)nn.low(); r.sd = (i64)dd.low();
            return r;
        }
        r.small = false; r.neg = ng; r.n = std::move(nn); r.dn = std::move(dd);
        return r;
    }
};

static Rational neg(const Rational& a) {
    Rational r = a;
    if (r.small) r.sn = -r.sn; else r.neg = !r.neg;
    return r;
}
static Rational from_i128(i128 num, i128 den) {   // den > 0
    u128 an = num < 0 ? (u128)(-num) : (u128)num, ad = (u128)den;
    u128 x = an, y = ad;
    while (y) { u128 t = x % y; x = y; y = t; }
    if (x > 1) { an /= x; ad /= x; }
    if (an < ((u128)1 << 62) && ad < ((u128)1 << 62)) {
        Rational r; r.sn = num < 0 ? -(i64)an : (i64)an; r.sd = (i64)ad; return r;
    }
    Mag nn, dd;
    nn.d = {(u64)an, (u64)(an >> 64)}; nn.trim();
    dd.d = {(u64)ad, (u64)(ad >> 64)}; dd.trim();
    return Rational::from_big(num < 0, nn, dd, false);
}
static Rational add(const Rational& a, const Rational& b) {
    if (a.small && b.small) {
        i128 num = (i128)a.sn * b.sd + (i128)b.sn * a.sd;
        i128 den = (i128)a.sd * b.sd;
        return from_i128(num, den);
    }
    bool an, bn; Mag ann, ad, bnn, bd;
    a.to_big(an, ann, ad); b.to_big(bn, bnn, bd);
    Mag x = mul(ann, bd), y = mul(bnn, ad), den = mul(ad, bd);
    if (an == bn) return Rational::from_big(an, add(x, y), den);
    int c = cmp(x, y);
    if (c == 0) return Rational();
    return c > 0 ? Rational::from_big(an, sub(x, y), den) : Rational::from_big(bn, sub(y, x), den);
}
static Rational sub(const Rational& a, const Rational& b) { return add(a, neg(b)); }
static Rational mul(const Rational& a, const Rational& b) {
    if (a.is_zero() || b.is_zero()) return Rational();
    if (a.small && b.small) return from_i128((i128)a.sn * b.sn, (i128)a.sd * b.sd);
    bool an, bn; Mag ann, ad, bnn, bd;
    a.to_big(an, ann, ad); b.to_big(bn, bnn, bd);
    // cross-cancel first (cheaper gcds)
    Mag g1 = gcd(ann, bd), g2 = gcd(bnn, ad);
    if (!g1.is_one()) { ann = divexact(ann, g1); bd = divexact(bd, g1); }
    if (!g2.is_one()) { bnn = divexact(bnn, g2); ad = divexact(ad, g2); }
    return Rational::from_big(an != bn, mul(ann, bnn), mul(ad, bd), false);
}
static Rational inv(const Rational& a) {
    if (a.small) return a.sn < 0 ? Rational(-a.sd, -a.sn) : Rational(a.sd, a.sn);
    Rational r; r.small = false; r.neg = a.neg; r.n = a.dn; r.dn = a.n;
    return Rational::from_big(r.neg, r.n, r.dn, false);
}
static Rational div(const Rational& a, const Rational& b) { return mul(a, inv(b)); }
static int cmp(const Rational& a, const Rational& b) {
    if (a.small && b.small) {
        i128 x = (i128)a.sn * b.sd, y = (i128)b.sn * a.sd;
        return (x > y) - (x < y);
    }
    int sa = a.sign(), sb = b.sign();
    if (sa != sb) return sa < sb ? -1 : 1;
    if (sa == 0) return 0;
    bool an, bn; Mag ann, ad, bnn, bd;
    a.to_big(an, ann, ad); b.to_big(bn, bnn, bd);
    int c = cmp(mul(ann, bd), mul(bnn, ad));
    return sa > 0 ? c : -c;
}
static std::string to_hex(const Mag& m) {
    if (m.zero()) return "0";
    static const char* hx = "0123456789abcdef";
    std::string s;
    for (size_t i = m.size(); i-- > 0;)
        for (int k = 60; k >= 0; k -= 4) s.push_back(hx[(m.d[i] >> k) & 15]);
    size_t p = s.find_first_not_of('0');
    return s.substr(p);
}
static std::string to_string(const Rational& a) {   // "[-]hexnum/hexden"
    bool ng; Mag nn, dd;
    a.to_big(ng, nn, dd);
    return std::string(ng ? "-" : "") + to_hex(nn) + "/" + to_hex(dd);
}

// ------------------------------------------------------------------------------------------------
// problem
// ------------------------------------------------------------------------------------------------
typedef std::vector<std::pair<int, Rational>> SparseCol;

// A column of the constraint matrix: a view of compact storage (no per-entry rationals are kept).
struct Col {
    const int* idx = nullptr; const i64* num = nullptr; const i64* den = nullptr; int len = 0;   // CSC slice
    const int8_t* d8 = nullptr; int m = 0;                                                        // dense int8
    const SparseCol* sc = nullptr;                                                                // materialised
    bool dense() const { return d8 != nullptr; }
    size_t size() const { return sc ? sc->size() : (d8 ? (size_t)m : (size_t)len); }
    template <class F> void for_each(F f) const {     // f(row, value) in ascending row order, zeros skipped
        if (sc) { for (auto& e : *sc) f(e.first, e.second); return; }
        if (d8) { for (int i = 0; i < m; ++i) if (d8[i]) f(i, Rational((i64)d8[i])); return; }
        for (int k = 0; k < len; ++k) f(idx[k], den ? Rational(num[k], den[k]) : Rational(num[k]));
    }
};

struct Provider {   // MatrixProvider (matrix_provider/mod.rs:37-134)
    int m = 0, n = 0;
    const i64* colptr = nullptr; const int* rowidx = nullptr; const i64* vnum = nullptr; const i64* vden = nullptr;
    int nd = 0; const int8_t* dense = nullptr;      // columns [0, nd): column-major nd x m, CSC ranges empty
    std::vector<Rational> cost, rhs;
    bool partial = false, full = false;
    std::vector<std::pair<int, int>> pivots;
    Col column(int j) const {
        Col c;
        if (j < nd) { c.d8 = dense + (size_t)j * m; c.m = m; return c; }
        c.idx = rowidx + colptr[j]; c.num = vnum + colptr[j]; c.den = vden ? vden + colptr[j] : nullptr;
        c.len = (int)(colptr[j + 1] - colptr[j]);
        return c;
    }
};

typedef std::map<int, Rational> SparseRow;   // column -> value, ordered (SparseVector semantics)

static Rational dot_dense(const std::vector<Rational>& dense, const std::vector<int>* nz, const Col& col) {
    // DenseVector::sparse_inner_product (data/linear_algebra/vector/dense.rs:101-112).  `nz` (optional):
    // the indices of the non-zero entries of `dense`, used to walk a dense column by index.
    Rational s;
    if (col.dense() && nz) {
        for (int i : *nz) if (col.d8[i]) s = add(s, mul(dense[i], Rational((i64)col.d8[i])));
        return s;
    }
    col.for_each([&](int i, const Rational& v) {
        const Rational& dv = dense[i];
        if (!dv.is_zero()) s = add(s, mul(dv, v));
    });
    return s;
}
static Rational dot_row(const SparseRow& row, const Col& col) {
    // SparseVector::sparse_inner_product (data/linear_algebra/vector/sparse.rs:105-128)
    Rational s;
    if (col.dense()) {      // implicit indices: direct lookup per row entry
        for (auto& kv : row) if (col.d8[kv.first]) s = add(s, mul(kv.second, Rational((i64)col.d8[kv.first])));
        return s;
    }
    if (col.sc) {
        const SparseCol& c = *col.sc;
        if (row.size() < 4 * c.size()) {
            auto it = row.begin();
            size_t k = 0;
            while (it != row.end() && k < c.size()) {
                if (it->first < c[k].first) ++it;
                else if (it->first > c[k].first) ++k;
                else { s = add(s, mul(it->second, c[k].second)); ++it; ++k; }
            }
            return s;
        }
        for (auto& e : c) {
            auto it = row.find(e.first);
            if (it != row.end()) s = add(s, mul(it->second, e.second));
        }
        return s;
    }
    auto val = [&](int k) { return col.den ? Rational(col.num[k], col.den[k]) : Rational(col.num[k]); };
    if (row.size() < 4 * (size_t)col.len) {
        // merge join over two sorted sequences
        auto it = row.begin();
        int k = 0;
        while (it != row.end() && k < col.len) {
            if (it->first < col.idx[k]) ++it;
            else if (it->first > col.idx[k]) ++k;
            else { s = add(s, mul(it->second, val(k))); ++it; ++k; }
        }
        return s;
    }
    for (int k = 0; k < col.len; ++k) {
        auto it = row.find(col.idx[k]);
        if (it != row.end()) s = add(s, mul(it->second, val(k)));
    }
    return s;
}

struct Info {   // BasisChangeComputationInfo (tableau/mod.rs:205-234)
    int p, q, leaving;
    SparseRow column, work, row_p;
};

struct Carry {   // carry/mod.rs:46-66
    Rational minus_objective;
    std::vector<Rational> minus_pi, b;
    std::vector<int> basis;
    std::vector<SparseRow> rows;
    int m() const { return (int)b.size(); }

    std::vector<int> pi_nz;      // indices of the non-zero entries of minus_pi (refreshed by the pricing loops)
    bool pi_nz_valid = false;
    void refresh_pi_nz() {
        pi_nz.clear();
        for (int i = 0; i < m(); ++i) if (!minus_pi[i].is_zero()) pi_nz.push_back(i);
        pi_nz_valid = true;
    }
    Rational cost_difference(const Col& c) const { return dot_dense(minus_pi, pi_nz_valid ? &pi_nz : nullptr, c); }   // :606-611
    SparseRow generate_column(const Col& c) const {                                          // :613-621
        SparseRow out;
        for (int i = 0; i < m(); ++i) {
            Rational v = dot_row(rows[i], c);
            if (!v.is_zero()) out.emplace_hint(out.end(), i, std::move(v));
        }
        return out;
    }
    Rational generate_element(int i, const Col& c) const { return dot_row(rows[i], c); }

    Info change_basis(int p, int q, SparseRow column, const Rational& relative_cost) {       // :561-604
        Info info;
        pi_nz_valid = false;
        info.p = p; info.q = q;
        // work vector = column^T B^-1 (basis_inverse_rows.rs:162-177)
        for (auto& ia : column)
            for (auto& kv : rows[ia.first]) {
                Rational t = mul(ia.second, kv.second);
                auto it = info.work.find(kv.first);
                if (it == info.work.end()) info.work.emplace(kv.first, std::move(t));
                else it->second = add(it->second, t);
            }
        for (auto it = info.work.begin(); it != info.work.end();)
            it = it->second.is_zero() ? info.work.erase(it) : std::next(it);
        // update_b (:295-325)
        const Rational pivot_value = column.at(p);
        b[p] = div(b[p], pivot_value);
        for (auto& ia : column)
            if (ia.first != p) b[ia.first] = sub(b[ia.first], mul(ia.second, b[p]));
        info.leaving = basis[p];
        basis[p] = q;
        // BasisInverseRows::change_basis (basis_inverse_rows.rs:43-99)
        SparseRow& rowp = rows[p];
        if (!(pivot_value.small && pivot_value.sn == 1 && pivot_value.sd == 1))
            for (auto& kv : rowp) kv.second = div(kv.second, pivot_value);
        for (auto& ia : column) {
            if (ia.first == p) continue;
            SparseRow& row = rows[ia.first];
            for (auto& kv : rowp) {
                Rational t = mul(ia.second, kv.second);
                auto it = row.find(kv.first);
                if (it == row.end()) row.emplace(kv.first, neg(t));
                else {
                    it->second = sub(it->second, t);
                    if (it->second.is_zero()) row.erase(it);
                }
            }
        }
        // update_minus_pi_and_obj (:338-349)
        for (auto& kv : rowp) minus_pi[kv.first] = sub(minus_pi[kv.first], mul(relative_cost, kv.second));
        minus_objective = sub(minus_objective, mul(relative_cost, b[p]));
        info.row_p = rowp;
        info.column = std::move(column);
        return info;
    }
};

struct Tableau {   // tableau/mod.rs:25-39
    const Provider* prov = nullptr;
    Carry im;
    std::unordered_set<int> basis_columns;
    bool artificial = false;
    std::vector<int> column_to_row;   // Partially::column_to_row
    std::vector<int> row_map;         // phase-two row -> original row (after RemoveRows)
    std::vector<SparseCol> filtered;  // columns with rows removed (RemoveRows::column)
    bool use_filtered = false;

    int nr_artificial() const { return artificial ? (int)column_to_row.size() : 0; }
    int start_index() const { return nr_artificial(); }
    int nr_rows() const { return im.m(); }
    int nr_columns() const { return nr_artificial() + prov->n; }
    Rational initial_cost(int j) const {
        if (!artificial) return prov->cost[j];
        return j < nr_artificial() ? Rational(1) : Rational(0);
    }
    SparseCol identity_col(int j) const { return SparseCol{{column_to_row[j], Rational(1)}}; }
    Col provider_col(int j) const {
        if (use_filtered) { Col c; c.sc = &filtered[j]; return c; }
        return prov->column(j);
    }
    static Col view(const SparseCol& sc) { Col c; c.sc = &sc; return c; }
    bool in_basis(int j) const { return basis_columns.count(j) != 0; }
    Rational relative_cost(int j) const {   // tableau/mod.rs:106-112
        int na = nr_artificial();
        if (j < na) { SparseCol ic = identity_col(j); return add(im.cost_difference(view(ic)), initial_cost(j)); }
        return add(im.cost_difference(provider_col(j - na)), initial_cost(j));
    }
    SparseRow generate_column(int j) const {
        int na = nr_artificial();
        if (j < na) { SparseCol ic = identity_col(j); return im.generate_column(view(ic)); }
        return im.generate_column(provider_col(j - na));
    }
    Rational generate_element(int i, int j) const {
        int na = nr_artificial();
        if (j < na) { SparseCol ic = identity_col(j); return im.generate_element(i, view(ic)); }
        return im.generate_element(i, provider_col(j - na));
    }
    Col original_column_real(int j) const { return provider_col(j - nr_artificial()); }
    // ratio test with Bland tie-break (tableau/mod.rs:287-313)
    int select_primal_pivot_row(const SparseRow& column) const {
        int best = -1, best_leaving = 0;
        Rational best_ratio;
        for (auto& e : column) {
            if (e.second.sign() <= 0) continue;
            Rational ratio = div(im.b[e.first], e.second);
            int leaving = im.basis[e.first];
            if (best < 0) { best = e.first; best_ratio = ratio; best_leaving = leaving; continue; }
            int c = cmp(ratio, best_ratio);
            if (c == 0 && leaving < best_leaving) { best = e.first; best_leaving = leaving; }
            else if (c < 0) { best = e.first; best_ratio = ratio; best_leaving = leaving; }
        }
        return best;
    }
    Info bring_into_basis(int q, int p, SparseRow column, const Rational& cost) {
        Info info = im.change_basis(p, q, std::move(column), cost);
        basis_columns.erase(info.leaving);
        basis_columns.insert(q);
        return info;
    }
};

// ------------------------------------------------------------------------------------------------
// pivot rules (strategy/pivot_rule.rs)
// ------------------------------------------------------------------------------------------------
static int g_threads = 1;           // OpenMP team size of the column-parallel loops (fo_set_threads)
static double g_time_limit = 0;     // seconds; > 0: a solve stops (as at a pivot limit) once it has run this long

struct Rule {
    int kind;                       // 0 FirstProfitable, 1 ..WithMemory, 2 Dantzig, 3 steepest edge
    int last_selected = -1;
    std::vector<Rational> gamma;    // valid where has[j]
    std::vector<char> has;

    static Rational initial_gamma(int j, const Tableau& t) {   // :299-305
        Rational s(1);
        for (auto& e : t.generate_column(j)) s = add(s, mul(e.second, e.second));
        return s;
    }
    Rule(int k, const Tableau& t) : kind(k) {
        if (kind == 3) {   // :202-219
            gamma.resize(t.nr_columns()); has.assign(t.nr_columns(), 0);
            const int lo = t.start_index(), hi = t.nr_columns();
#pragma omp parallel for schedule(dynamic, 16) num_threads(g_threads)
            for (int j = lo; j < hi; ++j)
                if (!t.in_basis(j)) { gamma[j] = initial_gamma(j, t); has[j] = 1; }
        }
    }
    bool find_first(const Tableau& t, int lo, int hi, int& q, Rational& cost) const {
        for (int j = lo; j < hi; ++j) {
            if (t.in_basis(j)) continue;
            Rational c = t.relative_cost(j);
            if (c.sign() < 0) { q = j; cost = c; return true; }
        }
        return false;
    }
    bool select(Tableau& t, int& q, Rational& cost) {
        t.im.refresh_pi_nz();
        if (kind == 0) return find_first(t, t.start_index(), t.nr_columns(), q, cost);   // :95-108
        if (kind == 1) {   // :126-149
            bool ok;
            if (last_selected < 0) ok = find_first(t, t.start_index(), t.nr_columns(), q, cost);
            else ok = find_first(t, last_selected + 1, t.nr_columns(), q, cost) ||
                      find_first(t, t.start_index(), last_selected, q, cost);
            last_selected = ok ? q : -1;
            return ok;
        }
        // the relative costs (and steepest-edge keys) of all columns are independent: computed by the
        // thread team, then scanned in index order exactly like the reference's iterator chain
        const int lo = t.start_index(), hi = t.nr_columns();
        std::vector<Rational> costs(hi - lo), keys(kind == 3 ? hi - lo : 0);
        std::vector<char> neg(hi - lo, 0);
        const Tableau& ct = t;
#pragma omp parallel for schedule(dynamic, 64) num_threads(g_threads)
        for (int j = lo; j < hi; ++j) {
            if (ct.in_basis(j)) continue;
            Rational c = ct.relative_cost(j);
            if (c.sign() >= 0) continue;
            neg[j - lo] = 1;
            if (kind == 3) keys[j - lo] = div(mul(c, c), gamma[j]);
            costs[j - lo] = std::move(c);
        }
        bool found = false;
        Rational best_key;
        for (int j = lo; j < hi; ++j) {
            if (!neg[j - lo]) continue;
            const Rational& c = costs[j - lo];
            if (kind == 2) {   // Dantzig :163-186: strict < => lowest index on ties
                if (!found || cmp(c, cost) < 0) { q = j; cost = c; found = true; }
            } else {           // steepest edge :221-241: max_by_key => last maximum wins
                const Rational& key = keys[j - lo];
                if (!found || cmp(key, best_key) >= 0) { q = j; cost = c; best_key = key; found = true; }
            }
        }
        return found;
    }
    void after_basis_update(const Info& info, const Tableau& t) {   // :243-296
        if (kind != 3) return;
        has[info.q] = 0;
        Rational gamma_q(1);
        for (auto& e : info.column) gamma_q = add(gamma_q, mul(e.second, e.second));
        Rational one(1);
        const int lo = t.start_index(), hi = (int)gamma.size();
#pragma omp parallel for schedule(dynamic, 16) num_threads(g_threads)
        for (int j = lo; j < hi; ++j) {
            if (!has[j]) continue;
            const Col original = t.original_column_real(j);
            Rational alpha = dot_row(info.row_p, original);
            Rational alternative = one;
            Rational g = gamma[j];
            if (!alpha.is_zero()) {
                Rational sq = mul(alpha, alpha);
                Rational inner = dot_row(info.work, original);
                if (!inner.is_zero()) {
                    Rational first = mul(alpha, inner);
                    g = sub(g, first); g = sub(g, first);
                }
                g = add(g, mul(sq, gamma_q));
                alternative = add(one, sq);
            }
            if (cmp(g, alternative) < 0) g = alternative;
            gamma[j] = g;
        }
        const Rational& wp = info.column.at(info.p);
        gamma[info.leaving] = div(gamma_q, mul(wp, wp));
        has[info.leaving] = 1;
    }
};

// ------------------------------------------------------------------------------------------------
// loops
// ------------------------------------------------------------------------------------------------
struct TraceEntry { int phase, q, p, leaving; };

struct Solver {
    const Provider& prov;
    int rule_kind;
    long long max_pivots;
    std::vector<TraceEntry> trace;
    bool limit_hit = false;
    Solver(const Provider& p, int rk, long long mp) : prov(p), rule_kind(rk), max_pivots(mp) {}

    std::chrono::steady_clock::time_point t_start = std::chrono::steady_clock::now();
    bool budget() {
        if (max_pivots > 0 && (long long)trace.size() >= max_pivots) { limit_hit = true; return false; }
        if (g_time_limit > 0 &&
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > g_time_limit) {
            limit_hit = true; return false;
        }
        return true;
    }

    // returns 0 optimal (no entering), 1 unbounded, 2 limit
    int loop(Tableau& t, int phase) {
        Rule rule(rule_kind, t);
        for (;;) {
            if (!budget()) return 2;
            int q; Rational cost;
            if (!rule.select(t, q, cost)) return 0;
            SparseRow column = t.generate_column(q);
            int p = t.select_primal_pivot_row(column);
            if (p < 0) return 1;
            Info info = t.bring_into_basis(q, p, std::move(column), cost);
            rule.after_basis_update(info, t);
            int prow = t.row_map.empty() ? p : t.row_map[p];
            trace.push_back({phase, q, prow, info.leaving});
        }
    }

    // remove_artificial_basis_variables (phase_one.rs:232-278)
    std::vector<int> remove_artificial(Tableau& t) {
        std::vector<int> rows_to_remove;
        int na = t.nr_artificial();
        std::vector<std::pair<int, int>> arts;
        for (int i = 0; i < t.nr_rows(); ++i) if (t.im.basis[i] < na) arts.push_back({i, t.im.basis[i]});
        for (auto& pa : arts) {
            t.im.refresh_pi_nz();
            int pivot_row = pa.first;
            bool nonzero = !t.im.b[pivot_row].is_zero();
            int found = -1; Rational fcost;
            for (int j = na; j < t.nr_columns(); ++j) {
                if (t.in_basis(j)) continue;
                Rational cost = t.relative_cost(j);
                if (nonzero) {
                    if (!cost.is_zero()) continue;
                    if (t.generate_element(pivot_row, j).sign() > 0) { found = j; fcost = cost; break; }
                } else {
                    if (!t.generate_element(pivot_row, j).is_zero()) { found = j; fcost = cost; break; }
                }
            }
            if (found >= 0) {
                SparseRow column = t.generate_column(found);
                Info info = t.bring_into_basis(found, pivot_row, std::move(column), fcost);
                trace.push_back({0, found, pivot_row, info.leaving});
            } else rows_to_remove.push_back(pivot_row);
        }
        return rows_to_remove;
    }

    // status: 0 optimal 1 unbounded 2 infeasible -1 limit
    int solve(Tableau& final_tableau, std::vector<int>& rows_removed, int& nr_art) {
        const int m = prov.m;
        Tableau t;
        t.prov = &prov;
        nr_art = 0;
        if (prov.full) {   // two_phase/mod.rs:80-109, Carry::from_basis_pivots with an identity basis
            t.im.basis.assign(m, 0);
            for (auto& rc : prov.pivots) t.im.basis[rc.first] = rc.second;
            t.im.b = prov.rhs;
            t.im.rows.resize(m);
            for (int i = 0; i < m; ++i) t.im.rows[i].emplace(i, Rational(1));
            from_costs(t, prov.cost);
            for (int j : t.im.basis) t.basis_columns.insert(j);
        } else {
            // Fully (fully.rs:82-97) / Partially (partially.rs:125-205)
            std::vector<int> real_col(m, -1);
            if (prov.partial) for (auto& rc : prov.pivots) real_col[rc.first] = rc.second;
            t.artificial = true;
            for (int i = 0; i < m; ++i) if (real_col[i] < 0) t.column_to_row.push_back(i);
            int na = (int)t.column_to_row.size();
            nr_art = na;
            t.im.basis.resize(m);
            int a = 0;
            for (int i = 0; i < m; ++i) t.im.basis[i] = real_col[i] >= 0 ? na + real_col[i] : a++;
            t.im.b = prov.rhs;
            t.im.rows.resize(m);
            t.im.minus_pi.assign(m, Rational(0));
            Rational obj;
            for (int i = 0; i < m; ++i) {
                t.im.rows[i].emplace(i, Rational(1));
                if (real_col[i] < 0) { t.im.minus_pi[i] = Rational(-1); obj = add(obj, prov.rhs[i]); }
            }
            t.im.minus_objective = neg(obj);
            for (int j : t.im.basis) t.basis_columns.insert(j);
            int r = loop(t, 1);
            if (r == 2) { final_tableau = std::move(t); return -1; }
            if (r == 1) return -2;   // "Artificial cost can not be unbounded." (phase_one.rs:151)
            if (!t.im.minus_objective.is_zero()) return 2;
            bool any_art = false;
            for (int j : t.im.basis) if (j < na) any_art = true;
            if (any_art) rows_removed = remove_artificial(t);
            // from_artificial / from_artificial_removing_rows (non_artificial.rs:151-226, carry/mod.rs:499-559,673-712)
            std::unordered_set<int> basis_set;
            if (!rows_removed.empty()) {
                std::vector<char> skip(m, 0);
                for (int r2 : rows_removed) skip[r2] = 1;
                std::vector<int> new_index(m, -1);
                for (int i = 0, k = 0; i < m; ++i) if (!skip[i]) { new_index[i] = k++; t.row_map.push_back(i); }
                Carry c2;
                for (int i = 0; i < m; ++i) {
                    if (skip[i]) continue;
                    c2.basis.push_back(t.im.basis[i] - na);
                    c2.b.push_back(t.im.b[i]);
                    SparseRow row;
                    for (auto& kv : t.im.rows[i]) if (new_index[kv.first] >= 0) row.emplace(new_index[kv.first], kv.second);
                    c2.rows.push_back(std::move(row));
                }
                t.im = std::move(c2);
                t.filtered.resize(prov.n);
                for (int j = 0; j < prov.n; ++j)
                    prov.column(j).for_each([&](int i, const Rational& v) {
                        if (new_index[i] >= 0) t.filtered[j].push_back({new_index[i], v});
                    });
                t.use_filtered = true;
            } else {
                for (int& j : t.im.basis) j -= na;
            }
            t.artificial = false;
            t.column_to_row.clear();
            t.basis_columns.clear();
            for (int j : t.im.basis) t.basis_columns.insert(j);
            from_costs(t, prov.cost);
        }
        int r = loop(t, 2);
        final_tableau = std::move(t);
        return r == 0 ? 0 : (r == 1 ? 1 : -1);
    }

    // create_minus_pi_from_artificial / create_minus_obj_from_artificial (carry/mod.rs:226-283)
    void from_costs(Tableau& t, const std::vector<Rational>& cost) {
        int m = t.im.m();
        std::vector<Rational> pi(m);
        Rational obj;
        for (int i = 0; i < m; ++i) {
            const Rational& c = cost[t.im.basis[i]];
            if (c.is_zero()) continue;
            for (auto& kv : t.im.rows[i]) pi[kv.first] = add(pi[kv.first], mul(kv.second, c));
            obj = add(obj, mul(t.im.b[i], c));
        }
        t.im.minus_pi.resize(m);
        for (int i = 0; i < m; ++i) t.im.minus_pi[i] = neg(pi[i]);
        t.im.minus_objective = neg(obj);
    }
};

// ------------------------------------------------------------------------------------------------
// C ABI (ctypes): see oracle/fast_oracle.py
// ------------------------------------------------------------------------------------------------
struct fo_result {
    int status = 0;
    std::vector<int> trace;          // 4 ints per pivot
    std::vector<int> rows_removed;
    std::vector<int> bfs_cols;
    std::string objective;           // hex "num/den"
    std::string bfs_values;          // ';' separated
    int nr_artificial = 0;
    double seconds = 0;
};

extern "C" {

void fo_set_time_limit(double seconds) { g_time_limit = seconds; }   // 0: none
int fo_set_threads(int n) {      // n <= 0: all hardware threads
#ifdef _OPENMP
    g_threads = n > 0 ? n : omp_get_num_procs();
#else
    g_threads = 1;
#endif
    return g_threads;
}

// The arrays must stay alive for the duration of the call (they are viewed, not copied).
// dense (may be null): provider columns [0, n_dense) as int8, column-major n_dense x m; their CSC ranges are empty.
int fo_solve(int m, int n, const i64* colptr, const int* rowidx, const i64* val_num, const i64* val_den,
             const i64* cost_num, const i64* cost_den, const i64* rhs_num, const i64* rhs_den,
             int n_pivots, const int* prow, const int* pcol, int full_basis, int rule, long long max_pivots,
             int n_dense, const int8_t* dense, fo_result** out) {
    Provider p;
    p.m = m; p.n = n;
    p.colptr = colptr; p.rowidx = rowidx; p.vnum = val_num; p.vden = val_den;
    p.nd = dense ? n_dense : 0; p.dense = dense;
    p.cost.resize(n); p.rhs.resize(m);
    for (int j = 0; j < n; ++j) p.cost[j] = Rational(cost_num[j], cost_den ? cost_den[j] : 1);
    for (int i = 0; i < m; ++i) p.rhs[i] = Rational(rhs_num[i], rhs_den ? rhs_den[i] : 1);
    p.partial = n_pivots >= 0; p.full = full_basis != 0;
    for (int k = 0; k < n_pivots; ++k) p.pivots.push_back({prow[k], pcol[k]});
    fo_result* r = new fo_result();
    *out = r;
    auto t0 = std::chrono::steady_clock::now();
    Solver s(p, rule, max_pivots);
    Tableau fin;
    r->status = s.solve(fin, r->rows_removed, r->nr_artificial);
    r->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto& e : s.trace) { r->trace.push_back(e.phase); r->trace.push_back(e.q); r->trace.push_back(e.p); r->trace.push_back(e.leaving); }
    // optimum, or the intermediate basic solution of a phase-two prefix (pivot limit): same exports
    if (r->status == 0 || (r->status == -1 && fin.prov && !fin.artificial)) {
        r->objective = to_string(neg(fin.im.minus_objective));
        std::vector<std::pair<int, Rational>> bfs;   // Carry::current_bfs (carry/mod.rs:636-645)
        for (int i = 0; i < fin.im.m(); ++i)
            if (!fin.im.b[i].is_zero()) bfs.push_back({fin.im.basis[i], fin.im.b[i]});
        std::sort(bfs.begin(), bfs.end(), [](auto& a, auto& b) { return a.first < b.first; });
        for (auto& e : bfs) {
            r->bfs_cols.push_back(e.first);
            if (!r->bfs_values.empty()) r->bfs_values.push_back(';');
            r->bfs_values += to_string(e.second);
        }
    }
    return 0;
}
int fo_status(const fo_result* r) { return r->status; }
long long fo_trace_len(const fo_result* r) { return (long long)r->trace.size() / 4; }
const int* fo_trace(const fo_result* r) { return r->trace.data(); }
int fo_rows_removed_len(const fo_result* r) { return (int)r->rows_removed.size(); }
const int* fo_rows_removed(const fo_result* r) { return r->rows_removed.data(); }
int fo_bfs_len(const fo_result* r) { return (int)r->bfs_cols.size(); }
const int* fo_bfs_cols(const fo_result* r) { return r->bfs_cols.data(); }
const char* fo_bfs_values(const fo_result* r) { return r->bfs_values.c_str(); }
const char* fo_objective(const fo_result* r) { return r->objective.c_str(); }
int fo_nr_artificial(const fo_result* r) { return r->nr_artificial; }
double fo_seconds(const fo_result* r) { return r->seconds; }
void fo_free(fo_result* r) { delete r; }
// arithmetic self-check hook: op 0 add, 1 sub, 2 mul, 3 div, 4 cmp on rationals given as i64 pairs
// raised to small powers to reach the multi-limb code paths
const char* fo_arith(int op, i64 an, i64 ad, int apow, i64 bn, i64 bd, int bpow) {
    static std::string out;
    Rational a(1), b(1), x(an, ad), y(bn, bd);
    for (int i = 0; i < apow; ++i) a = mul(a, x);
    for (int i = 0; i < bpow; ++i) b = mul(b, y);
    Rational r;
    if (op == 0) r = add(a, b); else if (op == 1) r = sub(a, b); else if (op == 2) r = mul(a, b);
    else if (op == 3) r = div(a, b); else r = Rational(cmp(a, b));
    out = to_string(r);
    return out.c_str();
}
}
