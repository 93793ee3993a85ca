// 32-bit-limb multi-precision multiply-accumulate for the hot kernels.
//
// ptxas fuses each `mad.lo.cc.u32 / madc.hi.cc.u32` pair on the same product into one
// IMAD.WIDE.U32 with a carry chain, so a low N x N-limb product costs N(N+1)/2 IMADs plus one merge
// chain -- the 64-bit `mad.lo.cc.u64` forms the compiler lowers ~4x worse (see profiles/).
// Products whose limb index sum is even accumulate into `ev`, odd ones into `od` (which is shifted
// by one limb), so every carry chain walks consecutive registers; the two are merged once at the end.
// Everything is fully unrolled: all indices are compile-time constants.
#pragma once
#include "bigint.cuh"

namespace rg {

// one row of a truncated (low N limbs) product:  acc += a[0..] * bi, shifted by I limbs
//   ev[k] holds limb k, od[k] holds limb k+1.
template <int N, int I>
__device__ __forceinline__ void mp_row_lo(u32 (&ev)[N], u32 (&od)[N], const u32 (&a)[N], u32 bi) {
    // even chain: j = I%2, I%2+2, ...  position p = I + j (even), pairs (ev[p], ev[p+1])
    {
        constexpr int j0 = I & 1;
        bool first = true;
#pragma unroll
        for (int j = j0; I + j < N; j += 2) {
            const int p = I + j;
            if (p + 1 < N) {
                if (first)
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                                 : "+r"(ev[p]), "+r"(ev[p + 1 < N ? p + 1 : p]) : "r"(a[j]), "r"(bi));
                else
                    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                                 : "+r"(ev[p]), "+r"(ev[p + 1 < N ? p + 1 : p]) : "r"(a[j]), "r"(bi));
            } else {   // top limb: low half only
                if (first) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(ev[p]) : "r"(a[j]), "r"(bi));
                else asm volatile("madc.lo.u32 %0, %1, %2, %0;" : "+r"(ev[p]) : "r"(a[j]), "r"(bi));
            }
            first = false;
        }
    }
    // odd chain: position p = I + j odd, lo -> od[p-1], hi -> od[p]
    {
        constexpr int j0 = (I & 1) ^ 1;
        bool first = true;
#pragma unroll
        for (int j = j0; I + j < N; j += 2) {
            const int p = I + j;
            if (p + 1 < N) {
                if (first)
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                                 : "+r"(od[p - 1]), "+r"(od[p]) : "r"(a[j]), "r"(bi));
                else
                    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                                 : "+r"(od[p - 1]), "+r"(od[p]) : "r"(a[j]), "r"(bi));
            } else {
                if (first) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(od[p - 1]) : "r"(a[j]), "r"(bi));
                else asm volatile("madc.lo.u32 %0, %1, %2, %0;" : "+r"(od[p - 1]) : "r"(a[j]), "r"(bi));
            }
            first = false;
        }
    }
}

template <int N, int I>
struct MpRows {
    __device__ static __forceinline__ void run(u32 (&ev)[N], u32 (&od)[N], const u32 (&a)[N],
                                               const u32* __restrict__ b) {
        mp_row_lo<N, I>(ev, od, a, b[I]);
        MpRows<N, I + 1>::run(ev, od, a, b);
    }
};
template <int N>
struct MpRows<N, N> {
    __device__ static __forceinline__ void run(u32 (&)[N], u32 (&)[N], const u32 (&)[N], const u32* __restrict__) {}
};

// r = x*a + y*b  (mod 2^(32 N));  a, b are read limb by limb (shared / uniform memory friendly)
template <int N>
__device__ __forceinline__ void mp_mul2_lo(u32 (&r)[N], const u32 (&x)[N], const u32* __restrict__ a,
                                           const u32 (&y)[N], const u32* __restrict__ b) {
    u32 ev[N], od[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { ev[k] = 0; od[k] = 0; }
    MpRows<N, 0>::run(ev, od, x, a);
    MpRows<N, 0>::run(ev, od, y, b);
    // merge: r[0] = ev[0]; r[k] = ev[k] + od[k-1] + carry
    r[0] = ev[0];
    if (N > 1) {
        asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r[1]) : "r"(ev[1]), "r"(od[0]));
#pragma unroll
        for (int k = 2; k < N; ++k)
            asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r[k]) : "r"(ev[k]), "r"(od[k - 1]));
    }
}

// r = x*a (mod 2^(32 N))
template <int N>
__device__ __forceinline__ void mp_mul_lo(u32 (&r)[N], const u32 (&x)[N], const u32* __restrict__ a) {
    u32 ev[N], od[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { ev[k] = 0; od[k] = 0; }
    MpRows<N, 0>::run(ev, od, x, a);
    r[0] = ev[0];
    if (N > 1) {
        asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r[1]) : "r"(ev[1]), "r"(od[0]));
#pragma unroll
        for (int k = 2; k < N; ++k)
            asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r[k]) : "r"(ev[k]), "r"(od[k - 1]));
    }
}

}  // namespace rg
