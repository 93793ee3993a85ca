// Fixed-capacity multi-limb integer arithmetic for sm_100a.
//
// Numbers are little-endian arrays of 64-bit limbs in two's complement.  Two families:
//   * compile-time widths (`template<int W>`), fully unrolled, register resident -- used by the hot
//     carry kernels (rank-1 pivot update, FTRAN, work vector);
//   * run-time widths on local arrays -- used by the O(n) / O(1) bookkeeping kernels (pricing
//     compare, steepest-edge recurrence, pivot scalars) where generality matters more than speed.
//
// The pivot update never forms a wide product: the Bareiss exact division by the previous
// determinant D is folded into the multiply-subtract by working modulo 2^(64 W) with the 2-adic
// inverse of the odd part of D (Jebelean exact division), so only LOW products are ever needed.
#pragma once
#include <cstdint>

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;

#define RG_MAXL 16            // max limbs of a carry entry (north-star: 2/4/8/16 x 64 bit)
#define RG_MAXW 80            // max limbs of any run-time-width temporary (>= 4*RG_MAXL + 9)

namespace rg {

// ------------------------------------------------------------------------------------------
// 3-word column accumulator:  (c2,c1,c0) += a*b   (full 128-bit product)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mac3(u64& c0, u64& c1, u64& c2, u64 a, u64 b) {
    asm("mad.lo.cc.u64 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u64 %1, %3, %4, %1;\n\t"
        "addc.u64 %2, %2, 0;"
        : "+l"(c0), "+l"(c1), "+l"(c2) : "l"(a), "l"(b));
}
// (c1,c0) += a*b, dropping anything beyond 128 bits (last-but-one column of a low product)
__device__ __forceinline__ void mac2(u64& c0, u64& c1, u64 a, u64 b) {
    asm("mad.lo.cc.u64 %0, %2, %3, %0;\n\t"
        "madc.hi.u64 %1, %2, %3, %1;"
        : "+l"(c0), "+l"(c1) : "l"(a), "l"(b));
}
// c0 += lo(a*b)
__device__ __forceinline__ void mac1(u64& c0, u64 a, u64 b) { c0 += a * b; }

// ------------------------------------------------------------------------------------------
// compile-time width helpers
// ------------------------------------------------------------------------------------------
// r = a*b + c*d  (mod 2^(64 W)); product scanning, both products share the column accumulators.
template <int W>
__device__ __forceinline__ void mul2_lo(u64 (&r)[W], const u64 (&a)[W], const u64 (&b)[W],
                                        const u64 (&c)[W], const u64 (&d)[W]) {
    u64 c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) {
#pragma unroll
        for (int i = 0; i <= k; ++i) {
            if (k == W - 1) {
                mac1(c0, a[i], b[k - i]);
                mac1(c0, c[i], d[k - i]);
            } else if (k == W - 2) {
                mac2(c0, c1, a[i], b[k - i]);
                mac2(c0, c1, c[i], d[k - i]);
            } else {
                mac3(c0, c1, c2, a[i], b[k - i]);
                mac3(c0, c1, c2, c[i], d[k - i]);
            }
        }
        r[k] = c0;
        c0 = c1; c1 = c2; c2 = 0;
    }
}

// r = a*b (mod 2^(64 W))
template <int W>
__device__ __forceinline__ void mul_lo(u64 (&r)[W], const u64 (&a)[W], const u64 (&b)[W]) {
    u64 c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) {
#pragma unroll
        for (int i = 0; i <= k; ++i) {
            if (k == W - 1) mac1(c0, a[i], b[k - i]);
            else if (k == W - 2) mac2(c0, c1, a[i], b[k - i]);
            else mac3(c0, c1, c2, a[i], b[k - i]);
        }
        r[k] = c0;
        c0 = c1; c1 = c2; c2 = 0;
    }
}

// acc (WA limbs, two's complement) += a (LA limbs, two's complement, sign-extended) * s (signed 64)
// Exact as long as the true value fits WA limbs; intermediate wrap-around is harmless.
template <int WA, int LA>
__device__ __forceinline__ void mac_small(u64 (&acc)[WA], const u64 (&a)[LA], i64 s) {
    // multiply the sign-extended operand limb by limb: unsigned product scanning with the
    // sign corrections of two's complement folded in afterwards.
    u64 us = (u64)s;
    u64 carry = 0;       // running high word
    u64 cf = 0;          // carry flag chain emulated
#pragma unroll
    for (int k = 0; k < WA; ++k) {
        u64 ak = (k < LA) ? a[k] : ((i64)a[LA - 1] < 0 ? ~0ull : 0ull);
        u64 lo = ak * us;
        u64 hi = __umul64hi(ak, us);
        // t = lo + carry
        u64 t = lo + carry;
        u64 c1 = t < lo;
        // acc[k] += t + cf
        u64 v = acc[k] + t;
        u64 c2 = v < t;
        u64 v2 = v + cf;
        u64 c3 = v2 < v;
        acc[k] = v2;
        cf = c2 + c3;            // 0..1 (cannot both be set)
        carry = hi + c1;         // hi <= 2^64-2 so no overflow
    }
    // two's complement correction for negative s: (a * us) - (a << 64) when s < 0 (mod 2^(64 WA))
    if (s < 0) {
        u64 bf = 0;
#pragma unroll
        for (int k = 1; k < WA; ++k) {
            u64 ak = (k - 1 < LA) ? a[k - 1] : ((i64)a[LA - 1] < 0 ? ~0ull : 0ull);
            u64 v = acc[k] - ak;
            u64 b1 = acc[k] < ak;
            u64 v2 = v - bf;
            u64 b2 = v < bf;
            acc[k] = v2;
            bf = b1 + b2;
        }
    }
}

// a (W limbs) += b (W limbs)
template <int W>
__device__ __forceinline__ void add_n(u64 (&a)[W], const u64 (&b)[W]) {
    u64 cf = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        u64 v = a[k] + b[k];
        u64 c1 = v < b[k];
        u64 v2 = v + cf;
        u64 c2 = v2 < v;
        a[k] = v2;
        cf = c1 + c2;
    }
}

// bit length of |x| for a two's complement W-limb value (0 for x == 0; for negative x the bit
// length of -x, rounded up by at most one for exact powers of two -- a safe upper bound).
template <int W>
__device__ __forceinline__ int bitlen_signed(const u64 (&x)[W]) {
    u64 sign = (i64)x[W - 1] < 0 ? ~0ull : 0ull;
    int bits = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        u64 v = x[k] ^ sign;
        if (v) bits = 64 * k + 64 - __clzll(v);
    }
    return bits + (sign ? 1 : 0);
}

// ------------------------------------------------------------------------------------------
// run-time width helpers (local arrays, lengths in limbs)
// ------------------------------------------------------------------------------------------
__device__ inline bool rt_is_neg(const u64* x, int n) { return (i64)x[n - 1] < 0; }
__device__ inline bool rt_is_zero(const u64* x, int n) {
    u64 o = 0;
    for (int k = 0; k < n; ++k) o |= x[k];
    return o == 0;
}
__device__ inline void rt_neg(u64* x, int n) {
    u64 c = 1;
    for (int k = 0; k < n; ++k) {
        u64 v = ~x[k] + c;
        c = (c && v == 0) ? 1 : 0;
        x[k] = v;
    }
}
// r = |x| ; returns sign (-1, 0, 1)
__device__ inline int rt_abs(u64* r, const u64* x, int n) {
    for (int k = 0; k < n; ++k) r[k] = x[k];
    if (rt_is_neg(x, n)) { rt_neg(r, n); return -1; }
    return rt_is_zero(x, n) ? 0 : 1;
}
__device__ inline int rt_bitlen_u(const u64* x, int n) {
    for (int k = n - 1; k >= 0; --k)
        if (x[k]) return 64 * k + 64 - __clzll(x[k]);
    return 0;
}
// sign-extend / truncate x (nx limbs) into r (nr limbs)
__device__ inline void rt_sext(u64* r, int nr, const u64* x, int nx) {
    u64 s = rt_is_neg(x, nx) ? ~0ull : 0ull;
    for (int k = 0; k < nr; ++k) r[k] = k < nx ? x[k] : s;
}
// unsigned full product r[na+nb] = a[na]*b[nb]
__device__ inline void rt_mul_full(u64* r, const u64* a, int na, const u64* b, int nb) {
    for (int k = 0; k < na + nb; ++k) r[k] = 0;
    for (int i = 0; i < na; ++i) {
        u64 carry = 0;
        u64 ai = a[i];
        if (ai == 0) continue;
        for (int j = 0; j < nb; ++j) {
            u64 lo = ai * b[j];
            u64 hi = __umul64hi(ai, b[j]);
            u64 v = r[i + j] + lo;
            u64 c1 = v < lo;
            u64 v2 = v + carry;
            u64 c2 = v2 < v;
            r[i + j] = v2;
            carry = hi + c1 + c2;
        }
        r[i + nb] += carry;   // r[i+nb] was zero or small enough: classic schoolbook invariant
    }
}
// r = a*b mod 2^(64 n); r must not alias a or b.  Product scanning: column k is summed into a 3-word accumulator
// and written once -- no read-modify-write of r through local memory, and the operand loads of a column do not
// depend on the accumulation chain (the operand-scanning form was bound by its load-add-store chain on r).
__device__ inline void rt_mul_lo(u64* r, const u64* a, const u64* b, int n) {
    u64 c0 = 0, c1 = 0, c2 = 0;
    for (int k = 0; k < n; ++k) {
        int i = 0;
        for (; i + 1 <= k; i += 2) {
            const u64 a0 = a[i], b0 = b[k - i], a1 = a[i + 1], b1 = b[k - i - 1];
            mac3(c0, c1, c2, a0, b0);
            mac3(c0, c1, c2, a1, b1);
        }
        if (i <= k) mac3(c0, c1, c2, a[i], b[k - i]);
        r[k] = c0;
        c0 = c1; c1 = c2; c2 = 0;
    }
}
__device__ inline void rt_add(u64* a, const u64* b, int n) {
    u64 cf = 0;
    for (int k = 0; k < n; ++k) {
        u64 v = a[k] + b[k];
        u64 c1 = v < b[k];
        u64 v2 = v + cf;
        u64 c2 = v2 < v;
        a[k] = v2;
        cf = c1 + c2;
    }
}
__device__ inline void rt_sub(u64* a, const u64* b, int n) {
    u64 bf = 0;
    for (int k = 0; k < n; ++k) {
        u64 v = a[k] - b[k];
        u64 b1 = a[k] < b[k];
        u64 v2 = v - bf;
        u64 b2 = v < bf;
        a[k] = v2;
        bf = b1 + b2;
    }
}
// unsigned compare: -1, 0, 1
__device__ inline int rt_cmp_u(const u64* a, const u64* b, int n) {
    for (int k = n - 1; k >= 0; --k) {
        if (a[k] != b[k]) return a[k] < b[k] ? -1 : 1;
    }
    return 0;
}
// signed compare of two's complement values of equal width
__device__ inline int rt_cmp_s(const u64* a, const u64* b, int n) {
    bool na = rt_is_neg(a, n), nb = rt_is_neg(b, n);
    if (na != nb) return na ? -1 : 1;
    return rt_cmp_u(a, b, n);
}
// logical shift right by `sh` bits (0 <= sh < 64 n), in place, zero fill
__device__ inline void rt_shr(u64* x, int n, int sh) {
    int w = sh >> 6, b = sh & 63;
    for (int k = 0; k < n; ++k) {
        u64 lo = (k + w < n) ? x[k + w] : 0;
        u64 hi = (k + w + 1 < n) ? x[k + w + 1] : 0;
        x[k] = b ? ((lo >> b) | (hi << (64 - b))) : lo;
    }
}
__device__ inline int rt_ctz(const u64* x, int n) {
    for (int k = 0; k < n; ++k)
        if (x[k]) return 64 * k + (__ffsll((long long)x[k]) - 1);
    return 64 * n;
}
// inverse of odd d modulo 2^(64 n) by Newton iteration: x <- x (2 - d x).
// `ws` is caller-provided scratch of 3*RG_MAXW limbs.  (Helpers never declare large local arrays
// of their own: every kernel slices ONE buffer by hand, because NVVM 12.9 was observed to merge the
// stack slots of simultaneously live arrays of inlined helpers -- see DESIGN.md "toolchain notes".)
__device__ inline void rt_inv_odd(u64* x, const u64* d, int nd, int n, u64* ws) {
    u64 d0 = d[0];
    u64 y = d0;                        // correct to 3 bits for odd d0
    for (int it = 0; it < 6; ++it) y *= 2 - d0 * y;
    u64* t = ws;
    u64* de = ws + RG_MAXW;
    u64* xn = ws + 2 * RG_MAXW;
    for (int k = 0; k < n; ++k) { x[k] = 0; de[k] = k < nd ? d[k] : 0; }
    x[0] = y;
    for (int have = 1; have < n; have *= 2) {
        int want = have * 2 < n ? have * 2 : n;
        rt_mul_lo(t, de, x, want);     // t = d x
        rt_neg(t, want);               // t = -d x
        u64 v = t[0] + 2;              // t += 2
        u64 c = v < t[0];
        t[0] = v;
        for (int k = 1; k < want && c; ++k) { t[k] += 1; c = t[k] == 0; }
        rt_mul_lo(xn, x, t, want);
        for (int k = 0; k < want; ++k) x[k] = xn[k];
    }
}

}  // namespace rg
