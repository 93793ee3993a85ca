// Hand-written sm_100a kernels of the exact simplex iteration.  See DESIGN.md for the maths.
//
// Layout: the carry is limb-planar -- plane l holds limb l of every entry, entry (i,k) at
// i*ld + k -- so a warp walking k reads 256 B (one column per thread) or 512 B (two columns per
// thread, 128-bit loads) per plane, fully coalesced.  Vectors (pivot column u, staged pivot row,
// work vector, per-column pricing data) use the same planar scheme.
#pragma once
#include "engine.cuh"
#include "mp32.cuh"
#include "k1_update.cuh"

namespace rg {

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
template <int LN>
__device__ __forceinline__ void load_planar(u64 (&x)[LN], const u64* __restrict__ base, size_t stride,
                                            size_t idx) {
#pragma unroll
    for (int l = 0; l < LN; ++l) x[l] = base[l * stride + idx];
}
template <int LN>
__device__ __forceinline__ void store_planar(u64* __restrict__ base, size_t stride, size_t idx,
                                             const u64 (&x)[LN]) {
#pragma unroll
    for (int l = 0; l < LN; ++l) base[l * stride + idx] = x[l];
}
__device__ inline void rt_load_planar(u64* x, int nl, const u64* base, size_t stride, size_t idx) {
    for (int l = 0; l < nl; ++l) x[l] = base[l * stride + idx];
}
__device__ inline void rt_store_planar(u64* base, size_t stride, size_t idx, const u64* x, int nl) {
    for (int l = 0; l < nl; ++l) base[l * stride + idx] = x[l];
}
// number of significant limbs of an unsigned magnitude
__device__ inline int rt_trim(const u64* x, int n) {
    while (n > 0 && x[n - 1] == 0) --n;
    return n;
}
// sign of (a*b - c*d) for two's complement operands; ws: 4*RG_MAXW limbs of scratch
__device__ inline int rt_cmp_prod(const u64* a, int na, const u64* b, int nb, const u64* c, int nc,
                                  const u64* d, int nd, u64* ws) {
    u64* ma = ws;
    u64* mb = ws + RG_MAXW;
    u64* p1 = ws + 2 * RG_MAXW;
    u64* p2 = ws + 3 * RG_MAXW;
    int s1 = rt_abs(ma, a, na) * rt_abs(mb, b, nb);
    int la = rt_trim(ma, na), lb = rt_trim(mb, nb);
    rt_mul_full(p1, ma, la, mb, lb);
    int l1 = la + lb;
    int s2 = rt_abs(ma, c, nc) * rt_abs(mb, d, nd);
    la = rt_trim(ma, nc); lb = rt_trim(mb, nd);
    rt_mul_full(p2, ma, la, mb, lb);
    int l2 = la + lb;
    if (s1 != s2) return s1 < s2 ? -1 : 1;
    if (s1 == 0) return 0;
    int n = l1 > l2 ? l1 : l2;
    for (int k = l1; k < n; ++k) p1[k] = 0;
    for (int k = l2; k < n; ++k) p2[k] = 0;
    int c0 = rt_cmp_u(p1, p2, n);
    return s1 > 0 ? c0 : -c0;
}

// ---------------------------------------------------------------------------------------------
// init: identity carry  (Carry::create_for_*_artificial, carry/mod.rs:374-442)
// ---------------------------------------------------------------------------------------------
__global__ void k_zero(u64* p, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t s = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) p[i] = 0;
}

// local rows: b in column 0, 1 on the diagonal (global column index).
// (packed active block: only column 0 exists at the start, the unit diagonal is implicit: diag = 0)
__global__ void k_init_identity(u64* C, size_t ps, int ld, int nloc, int row_lo, int L, const long long* rhs,
                                int diag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;   // local constraint row
    if (i >= nloc) return;
    int g = row_lo + i;
    size_t r = (size_t)(i + 1) * ld;
    long long b = rhs[g];
    C[r + 0] = (u64)b;
    for (int l = 1; l < L; ++l) C[l * ps + r + 0] = b < 0 ? ~0ull : 0ull;
    if (diag) C[r + (g + 1)] = 1;
}
// cost row (replicated on every rank): -1 under artificial rows
__global__ void k_init_row0(u64* C, size_t ps, int m, int L, const int* basis, const long long* artcost) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= m) return;
    if (basis[g] < 0) {
        long long c = artcost ? -artcost[g] : -1;    // -(phase-one cost of the artificial of row g)
        C[g + 1] = (u64)c;
        for (int l = 1; l < L; ++l) C[l * ps + (g + 1)] = c < 0 ? ~0ull : 0ull;
    }
}
// factor of the variable basic in each row (weighted problems): W / weight
__global__ void k_init_rowf(const int* basis, const long long* wf, const long long* artf, long long* rowf, int m) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= m) return;
    int j = basis[g];
    rowf[g] = j >= 0 ? wf[j] : artf[g];
}
// (0,0) = -sum_{artificial rows} b   and scalar state
__global__ void k_init_scalars(u64* C, size_t ps, int m, int L, const long long* rhs, const int* basis,
                               const long long* artcost, int row_lo, int nloc, int rank, int world, Scalars* sc) {
    if (threadIdx.x || blockIdx.x) return;
    u64 buf[3 * RG_MAXL];
    u64* acc = buf; u64* t = buf + RG_MAXL; u64* mg = buf + 2 * RG_MAXL;
    const int LA = L < 3 ? 3 : L;          // the objective is accumulated in at least 3 limbs
    for (int l = 0; l < LA; ++l) acc[l] = 0;
    int maxbits = 1;
    for (int i = 0; i < m; ++i) {
        long long b = rhs[i];
        u64 mag = b < 0 ? (u64)(-b) : (u64)b;
        int bl = mag ? 64 - __clzll(mag) : 0;
        if (bl > maxbits) maxbits = bl;
        if (basis[i] < 0) {
            // acc -= cost_i * b_i   (both non-negative, product up to 126 bits)
            u64 c = artcost ? (u64)artcost[i] : 1ull;
            int cb = c ? 64 - __clzll(c) : 0;
            if (cb > maxbits) maxbits = cb;
            u64 lo = c * mag, hi = __umul64hi(c, mag);
            for (int l = 0; l < LA; ++l) t[l] = l == 0 ? lo : (l == 1 ? hi : 0);
            rt_sub(acc, t, LA);
        }
    }
    rt_abs(mg, acc, LA);
    int bl = rt_bitlen_u(mg, LA);
    if (bl > maxbits) maxbits = bl;
    for (int l = 0; l < L; ++l) C[l * ps] = acc[l];   // the host re-initialises at a wider L if maxbits does not fit
    sc->status = ST_RUN;
    sc->q = -1; sc->p = -1; sc->pg = -1; sc->leaving = 0; sc->sgn = 1;
    sc->row_lo = row_lo; sc->nloc = nloc; sc->rank = rank; sc->world = world; sc->nk = 1;
    sc->t = 0; sc->E = 0; sc->t2 = 0; sc->E2 = 0;
    sc->maxbits_carry = maxbits; sc->maxbits_new = 0; sc->maxbits_u = 0; sc->maxbits_rowp = 0;
    sc->bits_D = 1; sc->predicted = 0; sc->last_selected = -1; sc->found = -1; sc->fatal = 0; sc->row0_ticket = 0;
    for (int l = 0; l < RG_MAXL; ++l) { sc->D[l] = l == 0; sc->Dnew[l] = 0; }
}

__global__ void k_set_inbasis(unsigned char* inbasis, const int* basis, int m) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m && basis[i] >= 0) inbasis[basis[i]] = 1;
}

// ---------------------------------------------------------------------------------------------
// K9: limb-width promotion.  Planar two's complement => the old planes are kept verbatim (one
// device copy) and the new planes are pure sign extension.
// ---------------------------------------------------------------------------------------------
__global__ void k_sign_extend(u64* dst, size_t ps, size_t count, int Lold, int Lnew) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t s = (size_t)gridDim.x * blockDim.x;
    for (; i < count; i += s) {
        u64 sg = (i64)dst[(size_t)(Lold - 1) * ps + i] < 0 ? ~0ull : 0ull;
        for (int l = Lold; l < Lnew; ++l) dst[(size_t)l * ps + i] = sg;
    }
}

// ---------------------------------------------------------------------------------------------
// K4 pricing / generic column dots:  out_j = cmul * cost_j * D + vec[1 + row] . a_j
// (Tableau::relative_cost, tableau/mod.rs:106-112 with cmul = 1; row dots with cmul = 0)
// One thread per provider column; `vec` is an (m+1)-vector whose entry k+1 pairs with row k.
// ---------------------------------------------------------------------------------------------
template <int LV, int LO>
__global__ void __launch_bounds__(256)
k_coldot(const u64* __restrict__ vec, size_t vs, int n, int j0, int j1, const long long* __restrict__ colptr,
         const int* __restrict__ rowidx, const long long* __restrict__ vals,
         const unsigned char* __restrict__ inbasis, const long long* __restrict__ cost, int cmul,
         int LD, u64* __restrict__ out, const Scalars* __restrict__ sc) {
    if (sc->status != ST_RUN) return;
    int j = j0 + blockIdx.x * blockDim.x + threadIdx.x;     // this rank's CSC columns [j0, j1)
    if (j >= j1) return;
    u64 acc[LO];
#pragma unroll
    for (int l = 0; l < LO; ++l) acc[l] = 0;
    if (!inbasis[j]) {
        if (cmul) {
            long long c = cost[j];
            if (c) {
                u64 d[LV];
#pragma unroll
                for (int l = 0; l < LV; ++l) d[l] = l < LD ? sc->D[l] : 0;
                mac_small<LO, LV>(acc, d, c);
            }
        }
        long long k0 = colptr[j], k1 = colptr[j + 1];
        for (long long k = k0; k < k1; ++k) {
            int r = rowidx[k];
            u64 x[LV];
            load_planar<LV>(x, vec, vs, (size_t)(r + 1));
            mac_small<LO, LV>(acc, x, vals[k]);
        }
    }
    store_planar<LO>(out, (size_t)n, (size_t)j, acc);
}


// ---------------------------------------------------------------------------------------------
// Dense int8 column block (config 5: every coefficient stored column-major, implicit row indices; the
// pivot column a_q is contiguous for FTRAN and the dots below read 16 contiguous rows per lane).
// Tensor-core form of the dense dots (exact): the multi-limb vector is cut into BYTE slices, and
//     R[s][j] = sum_i slice_s(vec_i) * a_ij        (u8 x s8 -> s32, mma.sync.m16n8k32)
// is an integer GEMM [slices x rows] x [rows x columns]; dot_j = sum_s R[s][j] 2^(8 s) is recombined with
// carries afterwards.  vec_i is taken in two's complement over nb bytes (nb from the tracked bit-length
// maximum): vec_i = sum_{s<nb} byte_s 2^(8s) - neg_i 2^(8 nb), so one extra 0/1 slice row carries the sign
// and every slice row is unsigned.  |R| <= rows * 255 * 128 < 2^31 for rows <= 65536 per k-slice.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int eff_bytes(const int* bits, int LV) {
    int need = (*bits + 1 + 7) >> 3;               // +1: sign bit
    return need < 1 ? 1 : (need > 8 * LV ? 8 * LV : need);
}
// slice rows Sl[s][i] (row pitch mp, a multiple of 64; rows i >= m are zero) and one flag per 64-row chunk
template <int LV>
__global__ void __launch_bounds__(64)
k_dense_slices(const u64* __restrict__ vec, size_t vs, int m, const int* bits, unsigned char* __restrict__ Sl,
               size_t mp, int* __restrict__ chunknz, const Scalars* sc,
               const unsigned char* __restrict__ triv = nullptr, int keep_trivial = 0) {
    // triv != nullptr: only the entries whose carry column (i + 1) is trivial (keep_trivial = 1) or listed
    // (keep_trivial = 0) are kept, the others count as zero (split sigma dot, launch_sigma_split)
    if (sc->status != ST_RUN) return;
    const int nb = eff_bytes(bits, LV);
    const int nrows = ((nb + 1 + 7) >> 3) << 3;    // byte rows + sign row, padded to an n-tile of 8
    const int i = blockIdx.x * 64 + threadIdx.x;
    const bool in = i < m && (!triv || (int)(triv[i + 1] != 0) == keep_trivial);
    const bool neg = in && (i64)vec[(size_t)(LV - 1) * vs + 1 + i] < 0;
    u64 any = 0;
    const int nl = (nb + 7) >> 3;
    for (int l = 0; l < nl; ++l) {
        u64 v = in ? vec[(size_t)l * vs + 1 + i] : 0;
        any |= v;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            int sidx = 8 * l + b;
            if (sidx < nb) Sl[(size_t)sidx * mp + i] = (unsigned char)(v >> (8 * b));
        }
    }
    for (int sidx = nb; sidx < nrows; ++sidx) Sl[(size_t)sidx * mp + i] = (sidx == nb && neg) ? 1 : 0;
    int nz = __syncthreads_or(any != 0);
    if (threadIdx.x == 0) chunknz[blockIdx.x] = nz;
}

__device__ __forceinline__ void mma_s8u8(int (&d)[4], u32 a0, u32 a1, u32 a2, u32 a3, u32 b0, u32 b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// CTA = 8 warps x 16 matrix columns (MMA M dimension); N dimension = NTC tiles of 8 slice rows; K = rows.
// A fragments come straight from the column-major int8 block (16 contiguous rows per lane, two MMAs per
// 128-bit load: the k order inside a chunk is a fixed permutation applied to both operands); the slice
// chunk is staged once per CTA in shared memory.  Chunks whose vector entries are all zero are skipped.
template <int NTC>
__global__ void __launch_bounds__(256)
k_dense_mma(const signed char* __restrict__ Acm, size_t ldc, int jd0, int jd1, const unsigned char* __restrict__ Sl,
            size_t mp, const int* __restrict__ chunknz, const int* bits, int LV, int rows_per_kslice,
            int* __restrict__ R, size_t rstride_k, int rpitch, const Scalars* sc) {
    __shared__ __align__(16) unsigned char sB[NTC * 8][64];
    if (sc->status != ST_RUN) return;
    const int nb = eff_bytes(bits, LV);
    const int nt_eff = (nb + 1 + 7) >> 3;
    const int tile0 = blockIdx.z * NTC;
    if (tile0 >= nt_eff) return;
    const int ntl = min(NTC, nt_eff - tile0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int c0 = jd0 + blockIdx.x * 128 + warp * 16 + g, c1 = c0 + 8;
    const bool v0 = c0 < jd1, v1 = c1 < jd1;
    int acc[NTC][4];
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) { acc[nt][0] = 0; acc[nt][1] = 0; acc[nt][2] = 0; acc[nt][3] = 0; }
    const int r_begin = blockIdx.y * rows_per_kslice;
    const int r_end = min((int)mp, r_begin + rows_per_kslice);
    const signed char* p0 = Acm + (size_t)(v0 ? c0 : jd0) * ldc + 16 * t;
    const signed char* p1 = Acm + (size_t)(v1 ? c1 : jd0) * ldc + 16 * t;
    for (int R0 = r_begin; R0 < r_end; R0 += 64) {
        if (!chunknz[R0 >> 6]) continue;           // uniform over the CTA
        const uint4 X0 = *reinterpret_cast<const uint4*>(p0 + R0);
        const uint4 X1 = *reinterpret_cast<const uint4*>(p1 + R0);
        __syncthreads();
        for (int e = threadIdx.x; e < ntl * 32; e += 256) {
            int row = e >> 2, part = e & 3;
            *reinterpret_cast<uint4*>(&sB[row][16 * part]) =
                *reinterpret_cast<const uint4*>(Sl + (size_t)(tile0 * 8 + row) * mp + R0 + 16 * part);
        }
        __syncthreads();
#pragma unroll
        for (int nt = 0; nt < NTC; ++nt) {
            if (nt < ntl) {
                const uint4 B = *reinterpret_cast<const uint4*>(&sB[nt * 8 + g][16 * t]);
                mma_s8u8(acc[nt], X0.x, X1.x, X0.y, X1.y, B.x, B.y);
                mma_s8u8(acc[nt], X0.z, X1.z, X0.w, X1.w, B.z, B.w);
            }
        }
    }
    int* base = R + (size_t)blockIdx.y * rstride_k;
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) {
        if (nt < ntl) {
            const size_t s0 = (size_t)(tile0 + nt) * 8 + 2 * t;
            if (v0) { base[s0 * rpitch + (c0 - jd0)] = acc[nt][0]; base[(s0 + 1) * rpitch + (c0 - jd0)] = acc[nt][1]; }
            if (v1) { base[s0 * rpitch + (c1 - jd0)] = acc[nt][2]; base[(s0 + 1) * rpitch + (c1 - jd0)] = acc[nt][3]; }
        }
    }
}
// recombination: dot_j = sum_{s<nb} R[s][j] 2^(8s) - R[nb][j] 2^(8 nb)  (+ cmul * cost_j * D), k-slices summed
template <int LV, int LO>
__global__ void __launch_bounds__(128)
k_dense_combine(const int* __restrict__ R, size_t rstride_k, int kslices, int rpitch, int n, int jd0, int jd1,
                const int* bits, const unsigned char* __restrict__ inbasis, const long long* __restrict__ cost,
                int cmul, int LD, u64* __restrict__ out, const Scalars* sc) {
    if (sc->status != ST_RUN) return;
    int j = jd0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= jd1) return;
    const int nb = eff_bytes(bits, LV);
    u64 res[LO];
#pragma unroll
    for (int l = 0; l < LO; ++l) res[l] = 0;
    if (!inbasis[j]) {
        long long carry = 0;
        // not unrolled over the limbs: 8 LO copies of the slice loop cost minutes of compile time at LO = 39.
        // The 8 slice sums of the NEXT limb are fetched while the current limb's carry chain runs (the loop is
        // otherwise one memory latency per limb).
        long long cur[8], nxt[8];
        // the 8 loads of a limb are unconditional (slice index clamped, value masked afterwards): guarded loads are
        // issued one after the other, and this loop is nothing but their latency
        auto fetch = [&](long long (&dst)[8], int l) {
            long long v[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) v[b] = 0;
            for (int k = 0; k < kslices; ++k) {
                int w[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int sidx = min(8 * l + b, nb);
                    w[b] = R[(size_t)k * rstride_k + (size_t)sidx * rpitch + (j - jd0)];
                }
#pragma unroll
                for (int b = 0; b < 8; ++b) v[b] += w[b];
            }
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int sidx = 8 * l + b;
                dst[b] = sidx < nb ? v[b] : (sidx == nb ? -v[b] : 0);
            }
        };
        fetch(cur, 0);
#pragma unroll 1
        for (int l = 0; l < LO; ++l) {
            if (l + 1 < LO && 8 * (l + 1) <= nb) fetch(nxt, l + 1);
            else {
#pragma unroll
                for (int b = 0; b < 8; ++b) nxt[b] = 0;
            }
            u64 limb = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                long long tt = carry + cur[b];
                limb |= (u64)(tt & 0xff) << (8 * b);
                carry = tt >> 8;                    // arithmetic: keeps the sign
            }
            res[l] = limb;
#pragma unroll
            for (int b = 0; b < 8; ++b) cur[b] = nxt[b];
        }
        if (cmul) {
            long long c = cost[j];
            if (c) {
                u64 d[LO];
#pragma unroll
                for (int l = 0; l < LO; ++l) d[l] = l < LD ? sc->D[l] : 0;
                mac_small<LO, LO>(res, d, c);
            }
        }
    }
    store_planar<LO>(out, (size_t)n, (size_t)j, res);
}


template <int LA, int LB>
__device__ __forceinline__ void mul_full_ct(u64 (&r)[LA + LB], const u64 (&a)[LA], const u64 (&b)[LB]);

// split sigma dot (list mode): the work vector restricted to the trivial carry columns is D * s (s = the factor
// vector, u or its weighted form), so   sigma_j = omega_listed . a_j  +  D * (s_trivial . a_j).
// This kernel adds the second term: sigma[j] += D * tau[j] for the dense columns j in [jd0, jd1).
template <int LT, int L, int LS>
__global__ void __launch_bounds__(128)
k_sigma_add_dtau(const u64* __restrict__ tau, int n, int jd0, int jd1, const unsigned char* __restrict__ inbasis,
                 u64* __restrict__ sigma, const Scalars* sc) {
    static_assert(LT + L <= LS, "D * tau must fit the sigma width");
    if (sc->status != ST_RUN) return;
    const int j = jd0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= jd1 || inbasis[j]) return;
    u64 t[LT], d[L];
    load_planar<LT>(t, tau, (size_t)n, (size_t)j);
    u64 any = 0;
#pragma unroll
    for (int l = 0; l < LT; ++l) any |= t[l];
    if (any == 0) return;
    const bool neg = (i64)t[LT - 1] < 0;
    if (neg) {
        u64 c = 1;
#pragma unroll
        for (int l = 0; l < LT; ++l) { u64 v = ~t[l] + c; c = (c && v == 0) ? 1 : 0; t[l] = v; }
    }
#pragma unroll
    for (int l = 0; l < L; ++l) d[l] = sc->Dold[l];     // the denominator the work vector was built with
    u64 p[LT + L];
    mul_full_ct<LT, L>(p, t, d);          // |tau| * D, D > 0
    u64 sg[LS];
    load_planar<LS>(sg, sigma, (size_t)n, (size_t)j);
    if (!neg) {
        u64 cf = 0;
#pragma unroll
        for (int l = 0; l < LS; ++l) {
            u64 b = l < LT + L ? p[l] : 0;
            u64 v = sg[l] + b; u64 c1 = v < b; u64 v2 = v + cf; u64 c2 = v2 < v;
            sg[l] = v2; cf = c1 + c2;
        }
    } else {
        u64 bf = 0;
#pragma unroll
        for (int l = 0; l < LS; ++l) {
            u64 b = l < LT + L ? p[l] : 0;
            u64 v = sg[l] - b; u64 b1 = sg[l] < b; u64 v2 = v - bf; u64 b2 = v < bf;
            sg[l] = v2; bf = b1 + b2;
        }
    }
    store_planar<LS>(sigma, (size_t)n, (size_t)j, sg);
}

}  // namespace rg
#include "dense_umma.cuh"
namespace rg {

// FTRAN of a dense column q (column-major copy): warp per carry row, lanes over the rows of a_q
template <int L>
__global__ void __launch_bounds__(256)
k_ftran_dense(const u64* __restrict__ C, size_t ps, int ld, int nrows, int m, const signed char* __restrict__ Acm,
              size_t ldc, const long long* __restrict__ cost, int qarg, int nd, u64* __restrict__ u, size_t us,
              Scalars* sc) {
    constexpr int LU = L + 2, NV = 2 * L;
    if (sc->status != ST_RUN) return;
    int q = qarg >= 0 ? qarg : sc->q;
    if (q >= nd) return;                      // sparse column: k_ftran handles it
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    size_t row = (size_t)warp * ld;
    const signed char* aq = Acm + (size_t)q * ldc;
    long long acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0;
    for (int k = lane; k < m; k += 32) {
        long long a = aq[k];
        u64 x[L];
        load_planar<L>(x, C, ps, row + 1 + k);
        u64 any = 0;
#pragma unroll
        for (int l = 0; l < L; ++l) any |= x[l];
        if (a == 0 || any == 0) continue;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            acc[2 * l] += a * (long long)(x[l] & 0xffffffffull);
            if (l < L - 1) acc[2 * l + 1] += a * (long long)(x[l] >> 32);
            else acc[2 * l + 1] += a * (long long)(int)(x[l] >> 32);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
    if (lane == 0) {
        u64 res[LU];
        long long carry = 0;
        u32 limbs[2 * LU];
#pragma unroll
        for (int k = 0; k < 2 * LU; ++k) {
            long long a = carry + (k < NV ? acc[k] : 0);
            limbs[k] = (u32)a;
            carry = a >> 32;
        }
#pragma unroll
        for (int l = 0; l < LU; ++l) res[l] = (u64)limbs[2 * l] | ((u64)limbs[2 * l + 1] << 32);
        if (warp == 0) {
            long long c = cost[q];
            if (c) {
                u64 d[L];
#pragma unroll
                for (int l = 0; l < L; ++l) d[l] = sc->D[l];
                mac_small<LU, L>(res, d, c);
            }
        }
        store_planar<LU>(u, us, (size_t)warp, res);
        atomicMax(&sc->maxbits_u, bitlen_signed<LU>(res));
    }
}

// steepest-edge weights of the dense columns on an identity carry: Ghat_j = (W/w_j)^2 + sum_i a_ij^2
__global__ void __launch_bounds__(128)
k_gamma_init_identity_dense(int nd, int n, int m, const signed char* __restrict__ Acm, size_t ldc,
                            const unsigned char* __restrict__ inbasis, u64* __restrict__ G, int LG) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);    // one warp per dense column
    if (j >= nd) return;
    u64 acc = 0;
    if (!inbasis[j]) {
        for (int i = lane; i < m; i += 32) { long long a = Acm[(size_t)j * ldc + i]; acc += (u64)(a * a); }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
        acc += 1;
    }
    if (lane == 0)
        for (int l = 0; l < LG; ++l) G[(size_t)l * n + j] = l == 0 ? acc : 0;
}

// ---------------------------------------------------------------------------------------------
// index reductions (pivot column choice, ratio test, artificial-removal search)
// ---------------------------------------------------------------------------------------------
// Comparators expose eligible(j) and better(j,k) (strict total order: key then index rule).

// columns a rank prices: a slice of the dense block plus a slice of the CSC columns (balanced separately)
struct ColOwn {
    int nd, d0, d1, s0, s1;
    __host__ __device__ bool owned(int j) const { return j < nd ? (j >= d0 && j < d1) : (j >= s0 && j < s1); }
    // the owned columns enumerated densely: t in [0, count()) -> column (kernels that do real work per column are
    // launched over this range, so a rank's threads are all busy in column-sharded runs)
    __host__ __device__ int count() const { return (d1 - d0) + (s1 - s0); }
    __host__ __device__ int at(int t) const { return t < d1 - d0 ? d0 + t : s0 + (t - (d1 - d0)); }
};
struct PriceView {
    const u64* kappa; int LU; int n;
    const unsigned char* inbasis; ColOwn own;
    __device__ bool negative(int j) const { return (i64)kappa[(size_t)(LU - 1) * n + j] < 0; }
};

// FirstProfitable (pivot_rule.rs:95-108): lowest j with negative cost
struct CmpFirst {
    PriceView v;
    __device__ bool eligible(int j) const { return v.own.owned(j) && !v.inbasis[j] && v.negative(j); }
    __device__ bool better(int j, int k) const { return j < k; }
};
// FirstProfitableWithMemory (pivot_rule.rs:126-149): first after `last`, wrapping, never `last`
struct CmpFirstMem {
    PriceView v; const Scalars* sc;
    __device__ bool eligible(int j) const { return v.own.owned(j) && j != sc->last_selected && !v.inbasis[j] && v.negative(j); }
    __device__ int key(int j) const { int last = sc->last_selected; return j > last ? j - last : j + v.n - last; }
    __device__ bool better(int j, int k) const { return key(j) < key(k); }
};
// Dantzig (pivot_rule.rs:163-186): most negative, strict < => lowest index on ties
struct CmpDantzig {
    PriceView v; const long long* wcol;     // true reduced cost = w_j * kappa_j / (D * cost scale)
    __device__ bool eligible(int j) const { return !v.inbasis[j] && v.negative(j); }
    __device__ bool better(int j, int k) const {
        u64 buf[4 * (RG_MAXL + 3)];
        u64* a = buf; u64* b = buf + (RG_MAXL + 3); u64* x = buf + 2 * (RG_MAXL + 3); u64* y = buf + 3 * (RG_MAXL + 3);
        rt_load_planar(a, v.LU, v.kappa, v.n, j);
        rt_load_planar(b, v.LU, v.kappa, v.n, k);
        int c;
        if (wcol) {     // both negative: compare magnitudes |kappa_j| w_j vs |kappa_k| w_k
            u64 wj = (u64)wcol[j], wk = (u64)wcol[k];
            rt_neg(a, v.LU); rt_neg(b, v.LU);
            rt_mul_full(x, a, v.LU, &wj, 1);
            rt_mul_full(y, b, v.LU, &wk, 1);
            c = -rt_cmp_u(x, y, v.LU + 1);
        } else {
            c = rt_cmp_s(a, b, v.LU);
        }
        return c < 0 || (c == 0 && j < k);
    }
};
// Steepest edge (pivot_rule.rs:221-241): max cost^2/gamma, max_by_key => highest index on ties.
// cost_j^2/gamma_j = kappa_j^2 W^2 / Ghat_j  =>  compare kappa_j^2 Ghat_k with kappa_k^2 Ghat_j.
struct CmpSteepest {
    PriceView v; const u64* G; int LG;
    __device__ bool eligible(int j) const { return !v.inbasis[j] && v.negative(j); }
    __device__ bool better(int j, int k) const {
        u64 buf[6 * RG_MAXW];
        u64* x = buf; u64* y = buf + RG_MAXW; u64* sj = buf + 2 * RG_MAXW; u64* sk = buf + 3 * RG_MAXW;
        u64* pj = buf + 4 * RG_MAXW; u64* pk = buf + 5 * RG_MAXW;
        rt_load_planar(x, v.LU, v.kappa, v.n, j);
        rt_abs(y, x, v.LU);
        int ly = rt_trim(y, v.LU);
        rt_mul_full(sj, y, ly, y, ly);
        int lsj = rt_trim(sj, 2 * ly);
        rt_load_planar(x, v.LU, v.kappa, v.n, k);
        rt_abs(y, x, v.LU);
        ly = rt_trim(y, v.LU);
        rt_mul_full(sk, y, ly, y, ly);
        int lsk = rt_trim(sk, 2 * ly);
        rt_load_planar(x, LG, G, v.n, k);
        int lg = rt_trim(x, LG);
        rt_mul_full(pj, sj, lsj, x, lg);
        int lpj = lsj + lg;
        rt_load_planar(x, LG, G, v.n, j);
        lg = rt_trim(x, LG);
        rt_mul_full(pk, sk, lsk, x, lg);
        int lpk = lsk + lg;
        int nn = lpj > lpk ? lpj : lpk;
        for (int i = lpj; i < nn; ++i) pj[i] = 0;
        for (int i = lpk; i < nn; ++i) pk[i] = 0;
        int c = rt_cmp_u(pj, pk, nn);
        return c > 0 || (c == 0 && j > k);
    }
};
// ratio test (tableau/mod.rs:287-313): rows with u_i > 0, min b_i/u_i, ties -> lowest leaving column.
// Index space: carry rows 1..m mapped to 0..m-1.
struct CmpRatio {
    const u64* C; size_t ps; int ld; int L;
    const u64* u; size_t us; int LU;
    const int* basis; int row_lo;
    __device__ bool eligible(int r) const {
        size_t i = (size_t)r + 1;
        if ((i64)u[(size_t)(LU - 1) * us + i] < 0) return false;
        u64 o = 0;
        for (int l = 0; l < LU; ++l) o |= u[(size_t)l * us + i];
        return o != 0;
    }
    __device__ bool better(int r, int s) const {
        u64 buf[4 * (RG_MAXL + 2) + 4 * RG_MAXW];
        u64* bi = buf; u64* bk = buf + (RG_MAXL + 2); u64* ui = buf + 2 * (RG_MAXL + 2);
        u64* uk = buf + 3 * (RG_MAXL + 2); u64* ws = buf + 4 * (RG_MAXL + 2);
        rt_load_planar(bi, L, C, ps, (size_t)(r + 1) * ld);
        rt_load_planar(bk, L, C, ps, (size_t)(s + 1) * ld);
        rt_load_planar(ui, LU, u, us, (size_t)r + 1);
        rt_load_planar(uk, LU, u, us, (size_t)s + 1);
        int c = rt_cmp_prod(bi, L, uk, LU, bk, L, ui, LU, ws);   // b_r/u_r ? b_s/u_s
        return c < 0 || (c == 0 && basis[row_lo + r] < basis[row_lo + s]);
    }
};
// remove_artificial_basis_variables search (phase_one.rs:245-261): first j (ascending)
struct CmpArtificial {
    PriceView v; const u64* nu; const Scalars* sc;
    __device__ bool eligible(int j) const {
        if (!v.own.owned(j) || v.inbasis[j]) return false;
        u64 o = 0;
        for (int l = 0; l < v.LU; ++l) o |= nu[(size_t)l * v.n + j];
        bool neg = (i64)nu[(size_t)(v.LU - 1) * v.n + j] < 0;
        if (sc->bp_nonzero) {
            u64 c = 0;
            for (int l = 0; l < v.LU; ++l) c |= v.kappa[(size_t)l * v.n + j];
            return c == 0 && o != 0 && !neg;
        }
        return o != 0;
    }
    __device__ bool better(int j, int k) const { return j < k; }
};

// ---------------------------------------------------------------------------------------------
// floating-point pre-filter for the index reductions.  Every candidate gets a log2-domain score
// accurate to ~1e-12; only candidates within SCORE_EPS of the best score are compared exactly, so
// the exact multi-limb cross-multiplications run on a handful of columns / rows instead of all.
// The exact winner is always inside the candidate set (|score error| << SCORE_EPS / 2).
// ---------------------------------------------------------------------------------------------
#define SCORE_EPS 1e-8
#define SCORE_NONE (-1.0e308)

// log2 of the magnitude of a planar two's complement number (sign returned separately)
__device__ inline double planar_log2_abs(const u64* base, size_t stride, size_t idx, int nl, int* sign) {
    bool neg = (i64)base[(size_t)(nl - 1) * stride + idx] < 0;
    u64 hi = 0, lo = 0, prev = 0, c = neg ? 1 : 0;
    int k = -1;
    for (int l = 0; l < nl; ++l) {
        u64 v = base[(size_t)l * stride + idx];
        if (neg) { v = ~v + c; c = (c && v == 0) ? 1 : 0; }   // |x| = ~x + 1, limb by limb
        if (v) { k = l; hi = v; lo = prev; }
        prev = v;
    }
    if (k < 0) { *sign = 0; return SCORE_NONE; }
    int sh = __clzll((long long)hi);
    u64 top = sh ? ((hi << sh) | (lo >> (64 - sh))) : hi;
    *sign = neg ? -1 : 1;
    return (double)(64 * k - sh) + log2((double)top);
}

// mode 2: Dantzig  score = log2|kappa|;  mode 3: steepest edge  score = 2 log2|kappa| - log2 Ghat
__global__ void __launch_bounds__(256)
k_score_columns(int n, ColOwn own, int mode, const u64* __restrict__ kappa, int LU, const u64* __restrict__ G, int LG,
                const unsigned char* __restrict__ inbasis, const long long* __restrict__ wcol,
                double* __restrict__ score, const Scalars* sc) {
    if (sc->status != ST_RUN) return;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = SCORE_NONE;
    if (own.owned(j) && !inbasis[j]) {
        int sg;
        double lk = planar_log2_abs(kappa, n, j, LU, &sg);
        if (sg < 0) {
            if (mode == 2) s = wcol ? lk + log2((double)wcol[j]) : lk;
            else { int sg2; s = 2.0 * lk - planar_log2_abs(G, n, j, LG, &sg2); }
        }
    }
    score[j] = s;
}
// ratio test: rows with u_i > 0; score = -(log2 b_i - log2 u_i) so that "best" is the maximum
__global__ void __launch_bounds__(256)
k_score_rows(int m, const u64* __restrict__ C, size_t ps, int ld, int L, const u64* __restrict__ u,
             size_t us, int LU, double* __restrict__ score, const Scalars* sc) {
    if (sc->status != ST_RUN) return;
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    int su, sb;
    double lu = planar_log2_abs(u, us, (size_t)r + 1, LU, &su);
    double s = SCORE_NONE;
    if (su > 0) {
        double lb = planar_log2_abs(C, ps, (size_t)(r + 1) * ld, L, &sb);
        s = sb == 0 ? 1.0e300 : -(lb - lu);
        if (sb < 0) s = 1.0e300;   // b >= 0 always holds for a basic feasible solution
    }
    score[r] = s;
}

// single block: max score, then exact comparison among the candidates within SCORE_EPS
template <class Cmp>
__global__ void __launch_bounds__(1024) k_select_scored(int off, int count, Cmp cmp,
                                                        const double* __restrict__ score, int mode, Scalars* sc,
                                                        const u64* __restrict__ take_u = nullptr,
                                                        size_t take_us = 0, int take_LU = 0) {
    __shared__ double sd[1024];
    __shared__ int sm[1024];
    if (sc->status != ST_RUN) return;
    int tid = threadIdx.x;
    double best = SCORE_NONE;
    for (int j = off + tid; j < off + count; j += blockDim.x) best = fmax(best, score[j]);
    sd[tid] = best;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (tid < s) sd[tid] = fmax(sd[tid], sd[tid + s]);
        __syncthreads();
    }
    double top = sd[0];
    int cand = -1;
    if (top > SCORE_NONE) {
        double thr = top >= 1.0e299 ? 1.0e299 : top - SCORE_EPS;
        for (int j = off + tid; j < off + count; j += blockDim.x) {
            if (score[j] >= thr && (cand < 0 || cmp.better(j, cand))) cand = j;
        }
    }
    sm[tid] = cand;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (tid < s) {
            int a = sm[tid], b = sm[tid + s];
            if (b >= 0 && (a < 0 || cmp.better(b, a))) sm[tid] = b;
        }
        __syncthreads();
    }
    if (tid == 0) {
        int bestj = sm[0];
        if (mode == 0) {
            sc->q = bestj;
            if (bestj < 0) sc->status = ST_OPTIMAL; else sc->last_selected = bestj;
        } else if (mode == 1) {
            sc->p = bestj < 0 ? -1 : bestj + 1;
            sc->pg = bestj < 0 ? -1 : sc->row_lo + bestj + 1;
            if (bestj < 0) sc->status = ST_UNBOUNDED;
            else if (take_u)     // the pivot element a = u[p] (what k_take_a does)
                for (int l = 0; l < take_LU; ++l) sc->a[l] = take_u[(size_t)l * take_us + bestj + 1];
        } else if (mode == 3) {          // local candidate of a row-sharded ratio test
            sc->p = bestj < 0 ? -1 : bestj + 1;
            sc->pg = bestj < 0 ? -1 : sc->row_lo + bestj + 1;
        } else if (mode == 4) {          // local candidate of a column-sharded pricing
            sc->q = bestj;
        } else {
            sc->found = bestj;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// row-sharded ratio test: every rank packs its local candidate, an all-gather exchanges them and
// every rank reduces them identically (deterministic Bland tie-break), SURVEY section 8e.
// candidate layout (u64 words): 0 valid, 1 global carry row, 2 basis id, 3 maxbits_u, 4 maxbits_carry,
// 5.. b numerator (RG_MAXL), then pivot column entry (RG_MAXL + 2)
// ---------------------------------------------------------------------------------------------
#define RG_CAND_WORDS (5 + RG_MAXL + RG_MAXL + 2)
__global__ void k_ratio_pack(const u64* __restrict__ C, size_t ps, int ld, int L, const u64* __restrict__ u,
                             size_t us, const int* __restrict__ basis, u64* __restrict__ send, const Scalars* sc) {
    if (threadIdx.x || blockIdx.x) return;
    for (int k = 0; k < RG_CAND_WORDS; ++k) send[k] = 0;
    if (sc->status != ST_RUN) return;
    const int LU = L + 2;
    send[3] = (u64)sc->maxbits_u;
    send[4] = (u64)sc->maxbits_carry;
    int p = sc->p;
    if (p < 1) return;
    send[0] = 1;
    send[1] = (u64)sc->pg;
    send[2] = (u64)(i64)basis[sc->pg - 1];
    for (int l = 0; l < L; ++l) send[5 + l] = C[(size_t)l * ps + (size_t)p * ld];
    for (int l = 0; l < LU; ++l) send[5 + RG_MAXL + l] = u[(size_t)l * us + p];
}
// The candidates of all ranks are ordered by exact cross-multiplied comparisons.  One thread per PAIR of ranks
// (block of 128 threads, world <= 16): a serial scan is world - 1 dependent multi-limb comparisons in one thread,
// which at eight ranks was a visible share of the pivot; the order is strict and total (distinct basis ids break
// ties), so the rank that wins all its pairs is the one the scan would keep.
__device__ __forceinline__ void pair_of(int t, int world, int& i, int& k) {
    // t-th pair (i < k) in row-major order of the strict upper triangle
    i = 0;
    int rowlen = world - 1;
    while (t >= rowlen && rowlen > 0) { t -= rowlen; ++i; --rowlen; }
    k = i + 1 + t;
}
__global__ void __launch_bounds__(128) k_ratio_merge(const u64* __restrict__ recv, int world, int L, Scalars* sc) {
    __shared__ unsigned char sBeats[16][16];     // sBeats[i][k]: candidate i precedes candidate k
    if (sc->status != ST_RUN) return;
    const int LU = L + 2;
    const int npairs = world * (world - 1) / 2;
    if ((int)threadIdx.x < npairs) {
        int i, k;
        pair_of(threadIdx.x, world, i, k);
        const u64* c = recv + (size_t)i * RG_CAND_WORDS;
        const u64* b = recv + (size_t)k * RG_CAND_WORDS;
        bool ik = false, ki = false;
        if (c[0] && b[0]) {
            u64 ws[4 * RG_MAXW];
            // ratio_i ? ratio_k :  b_i * u_k  vs  b_k * u_i
            int cmpv = rt_cmp_prod(c + 5, L, b + 5 + RG_MAXL, LU, b + 5, L, c + 5 + RG_MAXL, LU, ws);
            ik = cmpv < 0 || (cmpv == 0 && (i64)c[2] < (i64)b[2]);
            ki = !ik;
        } else { ik = c[0] != 0; ki = b[0] != 0 && !ik; }
        sBeats[i][k] = ik; sBeats[k][i] = ki;
    }
    __syncthreads();
    if (threadIdx.x) return;
    int best = -1;
    int mu = 0, mc = 0;
    for (int r = 0; r < world; ++r) {
        const u64* c = recv + (size_t)r * RG_CAND_WORDS;
        mu = max(mu, (int)c[3]);
        mc = max(mc, (int)c[4]);
        if (!c[0]) continue;
        bool all = true;
        for (int o = 0; o < world; ++o) if (o != r && !sBeats[r][o]) all = false;
        if (all) best = r;
    }
    sc->maxbits_u = mu;
    sc->maxbits_carry = mc;
    if (best < 0) { sc->status = ST_UNBOUNDED; sc->p = -1; sc->pg = -1; return; }
    const u64* b = recv + (size_t)best * RG_CAND_WORDS;
    int pg = (int)b[1];
    sc->pg = pg;
    int lo = sc->row_lo;
    sc->p = (pg - 1 >= lo && pg - 1 < lo + sc->nloc) ? pg - lo : -1;
    for (int l = 0; l < LU; ++l) sc->a[l] = b[5 + RG_MAXL + l];
}
// single GPU: the pivot element is local
__global__ void k_take_a(const u64* __restrict__ u, size_t us, int L, Scalars* sc) {
    if (threadIdx.x || blockIdx.x) return;
    if (sc->status != ST_RUN || sc->p < 1) return;
    for (int l = 0; l < L + 2; ++l) sc->a[l] = u[(size_t)l * us + sc->p];
}

template <class Cmp>
__device__ __forceinline__ int block_best(int best, const Cmp& cmp, int* sm) {
    int tid = threadIdx.x;
    sm[tid] = best;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (tid < s) {
            int a = sm[tid], b = sm[tid + s];
            if (b >= 0 && (a < 0 || cmp.better(b, a))) sm[tid] = b;
        }
        __syncthreads();
    }
    return sm[0];
}

template <class Cmp>
__global__ void __launch_bounds__(256) k_argbest1(int off, int n, Cmp cmp, int* cand, const Scalars* sc) {
    __shared__ int sm[256];
    if (sc->status != ST_RUN) return;
    int best = -1;
    for (int j = off + blockIdx.x * blockDim.x + threadIdx.x; j < off + n; j += gridDim.x * blockDim.x) {
        if (cmp.eligible(j) && (best < 0 || cmp.better(j, best))) best = j;
    }
    best = block_best(best, cmp, sm);
    if (threadIdx.x == 0) cand[blockIdx.x] = best;
}
// mode 0: entering column -> sc->q (none => ST_OPTIMAL); mode 1: pivot row -> sc->p (none =>
// ST_UNBOUNDED); mode 2: generic search -> sc->found
template <class Cmp>
__global__ void __launch_bounds__(256) k_argbest2(int nblocks, Cmp cmp, const int* cand, int mode,
                                                  Scalars* sc) {
    __shared__ int sm[256];
    if (sc->status != ST_RUN) return;
    int best = -1;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
        int j = cand[b];
        if (j >= 0 && (best < 0 || cmp.better(j, best))) best = j;
    }
    best = block_best(best, cmp, sm);
    if (threadIdx.x == 0) {
        if (mode == 0) {
            sc->q = best;
            if (best < 0) sc->status = ST_OPTIMAL; else sc->last_selected = best;
        } else if (mode == 1) {
            sc->p = best + 1;
            if (best < 0) sc->status = ST_UNBOUNDED;
        } else if (mode == 4) {
            sc->q = best;
        } else {
            sc->found = best;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// column-sharded pricing: every rank prices its own block of columns and keeps
// their steepest-edge weights; the local candidates are all-gathered and reduced identically on
// every rank with the rule's own exact order.
// candidate layout (u64 words): 0 valid, 1 column, 2 weight w_j, 3.. kappa (RG_MAXL+2), then Ghat (2 RG_MAXL+6)
// ---------------------------------------------------------------------------------------------
#define RG_COLCAND_WORDS (3 + (RG_MAXL + 2) + (2 * RG_MAXL + 6))
__global__ void k_column_pack(const u64* __restrict__ kappa, int LU, const u64* __restrict__ G, int LG, int n,
                              const long long* __restrict__ wcol, int use_found, u64* __restrict__ send,
                              const Scalars* sc) {
    if (threadIdx.x || blockIdx.x) return;
    for (int k = 0; k < RG_COLCAND_WORDS; ++k) send[k] = 0;
    if (sc->status != ST_RUN) return;
    int j = use_found ? sc->found : sc->q;
    if (j < 0) return;
    send[0] = 1; send[1] = (u64)j; send[2] = wcol ? (u64)wcol[j] : 1ull;
    for (int l = 0; l < LU; ++l) send[3 + l] = kappa[(size_t)l * n + j];
    for (int l = 0; l < LG; ++l) send[3 + (RG_MAXL + 2) + l] = G[(size_t)l * n + j];
}
// does candidate c (column j) precede candidate b (column k) under the rule?  (strict total order)
__device__ inline bool column_cand_better(const u64* c, const u64* b, int rule, int want_found, int L, int n, int last,
                                          u64* buf) {
    const int LU = L + 2, LG = 2 * L + 6;
    u64* x = buf; u64* y = buf + RG_MAXW; u64* sj = buf + 2 * RG_MAXW; u64* sk = buf + 3 * RG_MAXW;
    u64* pj = buf + 4 * RG_MAXW; u64* pk = buf + 5 * RG_MAXW;
    const int j = (int)c[1], k = (int)b[1];
    if (rule == 0 || want_found) return j < k;                                  // first profitable / first hit
    if (rule == 1) {                                                             // ... with memory
        int kj = j > last ? j - last : j + n - last, kk = k > last ? k - last : k + n - last;
        return kj < kk;
    }
    if (rule == 2) {                                                             // Dantzig: |kappa| w, ties lowest j
        for (int l = 0; l < LU; ++l) { x[l] = c[3 + l]; y[l] = b[3 + l]; }
        rt_neg(x, LU); rt_neg(y, LU);
        u64 wj = c[2], wk = b[2];
        rt_mul_full(sj, x, LU, &wj, 1);
        rt_mul_full(sk, y, LU, &wk, 1);
        int cmpv = rt_cmp_u(sj, sk, LU + 1);
        return cmpv > 0 || (cmpv == 0 && j < k);
    }
    // steepest edge, ties highest j
    for (int l = 0; l < LU; ++l) x[l] = c[3 + l];
    rt_abs(y, x, LU); int ly = rt_trim(y, LU);
    rt_mul_full(sj, y, ly, y, ly); int lsj = rt_trim(sj, 2 * ly);
    for (int l = 0; l < LU; ++l) x[l] = b[3 + l];
    rt_abs(y, x, LU); ly = rt_trim(y, LU);
    rt_mul_full(sk, y, ly, y, ly); int lsk = rt_trim(sk, 2 * ly);
    const u64* gk = b + 3 + (RG_MAXL + 2); const u64* gj = c + 3 + (RG_MAXL + 2);
    int lgk = rt_trim(gk, LG), lgj = rt_trim(gj, LG);
    rt_mul_full(pj, sj, lsj, gk, lgk); int lpj = lsj + lgk;
    rt_mul_full(pk, sk, lsk, gj, lgj); int lpk = lsk + lgj;
    int nn = lpj > lpk ? lpj : lpk;
    for (int i = lpj; i < nn; ++i) pj[i] = 0;
    for (int i = lpk; i < nn; ++i) pk[i] = 0;
    int cmpv = rt_cmp_u(pj, pk, nn);
    return cmpv > 0 || (cmpv == 0 && j > k);
}
// one thread per pair of ranks (see k_ratio_merge); the winner beats every other valid candidate
__global__ void __launch_bounds__(128) k_column_merge(const u64* __restrict__ recv, int world, int rule, int L, int n,
                                                      int want_found, Scalars* sc) {
    __shared__ unsigned char sBeats[16][16];
    if (sc->status != ST_RUN) return;
    const int LG = 2 * L + 6;
    const int npairs = world * (world - 1) / 2;
    if ((int)threadIdx.x < npairs) {
        int i, k;
        pair_of(threadIdx.x, world, i, k);
        const u64* c = recv + (size_t)i * RG_COLCAND_WORDS;
        const u64* b = recv + (size_t)k * RG_COLCAND_WORDS;
        bool ik = false, ki = false;
        if (c[0] && b[0]) {
            u64 buf[6 * RG_MAXW];
            ik = column_cand_better(c, b, rule, want_found, L, n, sc->last_selected, buf);
            ki = !ik;
        } else { ik = c[0] != 0; ki = b[0] != 0 && !ik; }
        sBeats[i][k] = ik; sBeats[k][i] = ki;
    }
    __syncthreads();
    if (threadIdx.x) return;
    int best = -1;
    for (int r = 0; r < world; ++r) {
        if (!recv[(size_t)r * RG_COLCAND_WORDS]) continue;
        bool all = true;
        for (int o = 0; o < world; ++o) if (o != r && !sBeats[r][o]) all = false;
        if (all) best = r;
    }
    if (want_found) { sc->found = best < 0 ? -1 : (int)recv[(size_t)best * RG_COLCAND_WORDS + 1]; return; }
    if (best < 0) { sc->q = -1; sc->status = ST_OPTIMAL; return; }
    const u64* b = recv + (size_t)best * RG_COLCAND_WORDS;
    sc->q = (int)b[1];
    sc->last_selected = sc->q;
    for (int l = 0; l < LG; ++l) sc->Gq[l] = b[3 + (RG_MAXL + 2) + l];
}

// ---------------------------------------------------------------------------------------------
// K2 FTRAN: u_i = sum_k a_kq C[i][1 + row_k]   for every carry row i (row 0: + c_q D = kappa_q)
// (Carry::generate_column, carry/mod.rs:613-621).  One warp per row, lanes over the column's
// nonzeros, multi-limb warp reduction.
// ---------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(256)
k_ftran(const u64* __restrict__ C, size_t ps, int ld, int nrows, const long long* __restrict__ colptr,
        const int* __restrict__ rowidx, const long long* __restrict__ vals,
        const long long* __restrict__ cost, int qarg, int nd, u64* __restrict__ u, size_t us, Scalars* sc) {
    constexpr int LU = L + 2;
    if (sc->status != ST_RUN) return;
    int q = qarg >= 0 ? qarg : sc->q;
    if (q < nd) return;                       // dense column: k_ftran_dense handles it
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    size_t row = (size_t)warp * ld;
    u64 acc[LU];
#pragma unroll
    for (int l = 0; l < LU; ++l) acc[l] = 0;
    long long k0 = colptr[q], k1 = colptr[q + 1];
    for (long long k = k0 + lane; k < k1; k += 32) {
        int r = rowidx[k];
        u64 x[L];
        load_planar<L>(x, C, ps, row + 1 + r);
        mac_small<LU, L>(acc, x, vals[k]);
    }
    if (warp == 0 && lane == 0) {
        long long c = cost[q];
        if (c) {
            u64 d[L];
#pragma unroll
            for (int l = 0; l < L; ++l) d[l] = sc->D[l];
            mac_small<LU, L>(acc, d, c);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u64 other[LU];
#pragma unroll
        for (int l = 0; l < LU; ++l) other[l] = __shfl_down_sync(0xffffffffu, acc[l], o);
        add_n<LU>(acc, other);
    }
    if (lane == 0) {
        store_planar<LU>(u, us, (size_t)warp, acc);
        int bl = bitlen_signed<LU>(acc);
        atomicMax(&sc->maxbits_u, bl);
    }
}


// ---------------------------------------------------------------------------------------------
// Active columns.  Carry column k (1..m) is `trivial` while it equals D e_k on
// rows 1..m: true for every column of an identity start, and a pivot in row p can only destroy it for
// k = p (the pivot row's own column, because row p of a trivial column k != p is zero).  Trivial
// columns are never read or written; the others are listed in klist (column 0, the right-hand side,
// is always listed).  Row 0 (the cost row) is maintained densely for all columns.
// ---------------------------------------------------------------------------------------------
__global__ void k_init_active(unsigned char* triv, int* klist, int* kpos, int ld, int m, int all_trivial) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ld) return;
    triv[k] = (all_trivial && k >= 1 && k <= m) ? 1 : 0;
    kpos[k] = 0;
    if (k == 0) klist[0] = 0;
}
// before the pivot row is staged: materialise the pivot row's own column if it is still trivial
// (packed block P: the column gets the next free slot, position sc->nk; k_activate_pivot_column lists it)
__global__ void __launch_bounds__(256)
k_materialise_pivot_column(u64* __restrict__ P, size_t ps, int cap, int L, const unsigned char* __restrict__ triv,
                           const Scalars* sc) {
    if (sc->status != ST_RUN) return;
    const int k = sc->pg;
    if (k < 1 || !triv[k]) return;
    int li = blockIdx.x * blockDim.x + threadIdx.x + 1;      // local carry row
    if (li > sc->nloc) return;
    bool diag = sc->row_lo + li == k;
    const int pos = sc->nk;
    for (int l = 0; l < L; ++l) P[(size_t)l * ps + (size_t)li * cap + pos] = diag ? sc->D[l] : 0ull;
}
// leave list mode: scatter the packed columns into the full carry (thread = list position, block row = carry row)
__global__ void __launch_bounds__(128)
k_unpack(u64* __restrict__ C, size_t ps, int ld, const u64* __restrict__ P, size_t pps, int cap, int L,
         const int* __restrict__ klist, const Scalars* sc) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= sc->nk) return;
    const int k = klist[t];
    for (int li = 1 + blockIdx.y; li <= sc->nloc; li += gridDim.y)
        for (int l = 0; l < L; ++l) C[(size_t)l * ps + (size_t)li * ld + k] = P[(size_t)l * pps + (size_t)li * cap + t];
}
__global__ void k_activate_pivot_column(unsigned char* triv, int* klist, int* kpos, Scalars* sc) {
    if (threadIdx.x || blockIdx.x) return;
    if (sc->status != ST_RUN) return;
    const int k = sc->pg;
    if (k >= 1 && triv[k]) { triv[k] = 0; klist[sc->nk] = k; kpos[k] = sc->nk; sc->nk = sc->nk + 1; }
}
// leave list mode: materialise every trivial column (thread = column, loops the local rows)
__global__ void __launch_bounds__(128)
k_materialise_all(u64* __restrict__ C, size_t ps, int ld, int L, int m, unsigned char* __restrict__ triv,
                  const Scalars* sc) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 1 || k > m || !triv[k]) return;
    for (int li = 1 + blockIdx.y; li <= sc->nloc; li += gridDim.y) {
        bool diag = sc->row_lo + li == k;
        for (int l = 0; l < L; ++l) C[(size_t)l * ps + (size_t)li * ld + k] = diag ? sc->D[l] : 0ull;
    }
}
__global__ void k_clear_trivial(unsigned char* triv, int ld) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < ld) triv[k] = 0;
}
// entering column scattered densely: aq[r] = a_rq (list-mode FTRAN indexes it by carry column)
__global__ void k_scatter_col(long long* __restrict__ aq, int m, int nd, const long long* __restrict__ colptr,
                              const int* __restrict__ rowidx, const long long* __restrict__ vals,
                              const signed char* __restrict__ Acm, size_t ldc, int qarg, const Scalars* sc) {
    if (sc->status != ST_RUN) return;
    int q = qarg >= 0 ? qarg : sc->q;
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nd) { if (r < m) aq[r] = Acm[(size_t)q * ldc + r]; return; }
    if (r < m) aq[r] = 0;
}
__global__ void k_scatter_col2(long long* __restrict__ aq, int nd, const long long* __restrict__ colptr,
                               const int* __restrict__ rowidx, const long long* __restrict__ vals, int qarg,
                               const Scalars* sc) {
    if (sc->status != ST_RUN) return;
    int q = qarg >= 0 ? qarg : sc->q;
    if (q < nd) return;
    long long k = colptr[q] + blockIdx.x * blockDim.x + threadIdx.x;
    if (k < colptr[q + 1]) aq[rowidx[k]] = vals[k];
}
// cost-row entry of the pivot column: u_0 = c_q D + sum_k aq[k-1] C[0][k] over ALL columns.  gridDim.x blocks each
// reduce a slice of the columns into `scratch`; the last block to finish (ticket in sc->row0_ticket, zeroed by
// k_reset_iter) adds the slices and the cost term.
// cost == nullptr: plain dot of the (m+1)-vector C[.][1..m] with the scattered column (rg_get_element)
template <int L>
__global__ void __launch_bounds__(256)
k_ftran_row0(const u64* __restrict__ C, size_t ps, int m, const long long* __restrict__ aq,
             const long long* __restrict__ cost, int qarg, u64* __restrict__ u, size_t us,
             u64* __restrict__ scratch, Scalars* sc) {
    constexpr int LU = L + 2;
    __shared__ u64 sAcc[8][LU];
    __shared__ int sLast;
    if (sc->status != ST_RUN) return;
    int q = qarg >= 0 ? qarg : sc->q;
    u64 acc[LU];
#pragma unroll
    for (int l = 0; l < LU; ++l) acc[l] = 0;
    for (int k = 1 + blockIdx.x * blockDim.x + threadIdx.x; k <= m; k += gridDim.x * blockDim.x) {
        long long a = aq[k - 1];
        if (!a) continue;
        u64 x[L];
        load_planar<L>(x, C, ps, (size_t)k);
        mac_small<LU, L>(acc, x, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u64 other[LU];
#pragma unroll
        for (int l = 0; l < LU; ++l) other[l] = __shfl_down_sync(0xffffffffu, acc[l], o);
        add_n<LU>(acc, other);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int l = 0; l < LU; ++l) sAcc[warp][l] = acc[l];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (blockDim.x >> 5); ++w) {
            u64 other[LU];
#pragma unroll
            for (int l = 0; l < LU; ++l) other[l] = sAcc[w][l];
            add_n<LU>(acc, other);
        }
#pragma unroll
        for (int l = 0; l < LU; ++l) scratch[(size_t)blockIdx.x * LU + l] = acc[l];
        __threadfence();
        sLast = atomicAdd(&sc->row0_ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!sLast || threadIdx.x != 0) return;
    __threadfence();
#pragma unroll
    for (int l = 0; l < LU; ++l) acc[l] = 0;
    for (int g = 0; g < (int)gridDim.x; ++g) {
        u64 other[LU];
#pragma unroll
        for (int l = 0; l < LU; ++l) other[l] = __ldcg(&scratch[(size_t)g * LU + l]);
        add_n<LU>(acc, other);
    }
    if (cost) {
        long long c = cost[q];
        if (c) {
            u64 d[L];
#pragma unroll
            for (int l = 0; l < L; ++l) d[l] = sc->D[l];
            mac_small<LU, L>(acc, d, c);
        }
    }
    store_planar<LU>(u, us, (size_t)0, acc);
    if (cost) atomicMax(&sc->maxbits_u, bitlen_signed<LU>(acc));
    sc->row0_ticket = 0;                  // ready for the next launch (rg_get_element uses the kernel too)
}
// list-mode FTRAN: u_i = sum_{k in klist, k >= 1} aq[k-1] C[i][k]  +  [column i trivial] D aq[i-1]
// (row 0 is dense: its warp walks all columns).  One warp per local carry row.
template <int L>
__global__ void __launch_bounds__(256)
k_ftran_list(const u64* __restrict__ C, size_t ps, int ld, int nrows, int m, const long long* __restrict__ aq,
             const int* __restrict__ klist, const unsigned char* __restrict__ triv,
             const long long* __restrict__ cost, int qarg, u64* __restrict__ u, size_t us, Scalars* sc) {
    constexpr int LU = L + 2;
    if (sc->status != ST_RUN) return;
    int q = qarg >= 0 ? qarg : sc->q;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    size_t row = (size_t)warp * ld;
    u64 acc[LU];
#pragma unroll
    for (int l = 0; l < LU; ++l) acc[l] = 0;
    if (warp == 0) return;      // the dense cost row is handled by k_ftran_row0
    {
        const int nk = sc->nk;
        for (int t = 1 + lane; t < nk; t += 32) {        // klist[0] is column 0 (b): not part of B^-1
            int k = klist[t];
            long long a = aq[k - 1];
            if (!a) continue;
            u64 x[L];
            load_planar<L>(x, C, ps, row + t);           // packed block: column klist[t] lives at position t
            mac_small<LU, L>(acc, x, a);
        }
        if (lane == 0) {
            int g = sc->row_lo + warp;                   // global carry row = its own column index
            if (g >= 1 && g <= m && triv[g]) {
                long long a = aq[g - 1];
                if (a) {
                    u64 d[L];
#pragma unroll
                    for (int l = 0; l < L; ++l) d[l] = sc->D[l];
                    mac_small<LU, L>(acc, d, a);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u64 other[LU];
#pragma unroll
        for (int l = 0; l < LU; ++l) other[l] = __shfl_down_sync(0xffffffffu, acc[l], o);
        add_n<LU>(acc, other);
    }
    if (lane == 0) {
        store_planar<LU>(u, us, (size_t)warp, acc);
        atomicMax(&sc->maxbits_u, bitlen_signed<LU>(acc));
    }
}

__global__ void k_reset_iter(Scalars* sc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { sc->maxbits_u = 0; sc->maxbits_rowp = 0; sc->maxbits_new = 0; sc->maxbits_tmp = 0; sc->nnz_s = 0; sc->maxbits_s = 0; sc->row0_ticket = 0;
        // the denominator this iteration starts from: side-stream kernels (split sigma dot) read it while the main
        // stream's bookkeeping may already have installed the new one
        for (int l = 0; l < RG_MAXL; ++l) sc->Dold[l] = sc->D[l]; }
}
__global__ void k_set_pq(Scalars* sc, int q, int p) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (q >= -1) sc->q = q;
        if (p >= 0) sc->p = p;
    }
}
__global__ void k_set_rows(Scalars* sc, int p_local, int pg) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { sc->p = p_local; sc->pg = pg; }
}
__global__ void k_set_status(Scalars* sc, int st) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { sc->status = st; if (st == ST_RUN) sc->fatal = 0; }
}

// stage the pivot row (old values) so the update can run in place.  Row-sharded: the owner copies,
// the other ranks zero-fill and a sum all-reduce (exact: one non-zero contributor) replicates it.
template <int L>
__global__ void __launch_bounds__(256)
k_copyrow(const u64* __restrict__ C, size_t ps, int ld, int stride, int m, const unsigned char* __restrict__ triv,
          const int* __restrict__ kpos, u64* __restrict__ rowp, size_t rs, Scalars* sc) {
    // dense mode: C = carry, stride = ld, triv = kpos = nullptr.  List mode: C = packed block, stride = cap;
    // a non-trivial column k sits at position kpos[k]; columns beyond m (padding) are zero.
    if (sc->status != ST_RUN) return;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < ld) {
        u64 x[L];
        // trivial columns hold D e_k implicitly; the pivot row's own column was materialised before
        if (sc->p >= 1 && !(triv && triv[k]) && !(kpos && k > m))
            load_planar<L>(x, C, ps, (size_t)sc->p * stride + (kpos ? kpos[k] : k));
        else {
            // implicit entry of a trivial column: D on its own row, 0 elsewhere
            const bool diag = sc->p >= 1 && triv && triv[k] && k == sc->pg;
#pragma unroll
            for (int l = 0; l < L; ++l) x[l] = diag ? sc->D[l] : 0;
        }
        store_planar<L>(rowp, rs, (size_t)k, x);
    }
}
template <int L>
__global__ void __launch_bounds__(256) k_rowbits(const u64* __restrict__ rowp, size_t rs, int ld, Scalars* sc) {
    if (sc->status != ST_RUN) return;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int bl = 0;
    if (k < ld) {
        u64 x[L];
        load_planar<L>(x, rowp, rs, (size_t)k);
        bl = bitlen_signed<L>(x);
    }
    bl = warp_max(bl);
    if ((threadIdx.x & 31) == 0 && bl) atomicMax(&sc->maxbits_rowp, bl);
}

// ---------------------------------------------------------------------------------------------
// pivot scalars (one thread): overflow prediction, 2-adic inverse of D, A = |a|/D, u_p' = a - D,
// steepest-edge scalars.  Sets ST_PROMOTE / ST_FATAL when the update would not fit L limbs.
// ---------------------------------------------------------------------------------------------
__global__ void k_scalars(int L, int E_host, Scalars* sc) {
    if (threadIdx.x || blockIdx.x) return;
    if (sc->status != ST_RUN) return;
    const int LU = L + 2;
    u64 buf[11 * RG_MAXW];     // ONE local buffer, sliced by hand (see bigint.cuh rt_inv_odd)
    u64* a = buf;               u64* am = buf + RG_MAXW;        u64* dodd = buf + 2 * RG_MAXW;
    u64* inv = buf + 3 * RG_MAXW;  u64* tmp = buf + 4 * RG_MAXW;   u64* ext = buf + 5 * RG_MAXW;
    u64* upv = buf + 6 * RG_MAXW;  u64* dext = buf + 7 * RG_MAXW;  u64* ws = buf + 8 * RG_MAXW;   // 3 slots
    for (int l = 0; l < LU; ++l) a[l] = sc->a[l];
    int sgn = rt_abs(am, a, LU);
    sc->sgn = sgn;
    if (sgn == 0) { sc->status = ST_FATAL; sc->fatal = 3; return; }   // zero pivot element: the caller's row is invalid
    int bits_a = rt_bitlen_u(am, LU);
    int bits_D = rt_bitlen_u(sc->D, L);
    sc->bits_D = bits_D;
    int bu = max(sc->maxbits_u, bits_D) + 1;
    int pred = max(bits_a + sc->maxbits_carry, bu + sc->maxbits_rowp) + 2 - bits_D;
    sc->predicted = pred;
    if (pred > 64 * L - 1) {
        sc->status = (L >= RG_MAXL) ? ST_FATAL : ST_PROMOTE;
        sc->fatal = 1;
        return;
    }
    int t = rt_ctz(sc->D, L);
    int E = E_host;                     // extra limbs of the kernel variant the host launches
    int W = L + E;
    sc->t = t; sc->E = E;
    if (E < ((t + 63) >> 6)) { sc->status = ST_FATAL; sc->fatal = 2; return; }   // variant too narrow for ctz(D)
    for (int l = 0; l < L; ++l) dodd[l] = sc->D[l];
    rt_shr(dodd, L, t);
    // two limbs beyond the K1 width: the reduced-cost recurrence (k_kappa_update) divides (L+2)-limb numbers; the
    // low W limbs are the W-limb inverse / factor (2-adic truncation), so K1 reads the same arrays
    const int W2 = W + 2;
    rt_inv_odd(inv, dodd, L, W2, ws);
#ifdef RG_DEBUG_PRINT
    printf("scalars: right after inv: inv0=%llu dodd0=%llu L=%d W=%d\n", inv[0], dodd[0], L, W);
#endif
    for (int l = 0; l < W2; ++l) sc->Dinv[l] = inv[l];
    for (int l = 0; l < W2; ++l) ext[l] = l < LU ? am[l] : 0;
    rt_mul_lo(tmp, ext, inv, W2);
    for (int l = 0; l < W2; ++l) sc->A[l] = tmp[l];
#ifdef RG_DEBUG_PRINT
    printf("scalars: L=%d t=%d E=%d W=%d D0=%llu dodd0=%llu inv0=%llu am0=%llu A0=%llu\n", L, t, E, W, sc->D[0], dodd[0], inv[0], am[0], tmp[0]);
#endif
    for (int l = 0; l < L; ++l) sc->Dnew[l] = am[l];
    // up = a - D  (two's complement, max(W, LU) limbs kept)
    int WU = W > LU ? W : LU;
    rt_sext(upv, WU, a, LU);
    for (int l = 0; l < WU; ++l) dext[l] = l < L ? sc->D[l] : 0;
    rt_sub(upv, dext, WU);
    for (int l = 0; l < WU; ++l) sc->up[l] = upv[l];
    sc->maxbits_new = 0;
}

// warp-cooperative low product in shared memory: r = a*b mod 2^(64 n).  Lane k sums column k into a
// 3-word accumulator, lane 0 propagates the carries.  r must not alias a or b; cols: 3*n words.
__device__ inline void warp_mul_lo(u64* r, const u64* a, const u64* b, int n, u64* cols) {
    const int lane = threadIdx.x & 31;
    for (int k = lane; k < n; k += 32) {
        u64 c0 = 0, c1 = 0, c2 = 0;
        for (int i = 0; i <= k; ++i) mac3(c0, c1, c2, a[i], b[k - i]);
        cols[3 * k] = c0; cols[3 * k + 1] = c1; cols[3 * k + 2] = c2;
    }
    __syncwarp();
    if (lane == 0) {
        u64 carry = 0;
        for (int k = 0; k < n; ++k) {
            // limb k = c0[k] + c1[k-1] + c2[k-2] + carry
            u64 v = cols[3 * k], c = 0;
            u64 t = v + carry; c += t < v; v = t;
            if (k >= 1) { t = v + cols[3 * (k - 1) + 1]; c += t < v; v = t; }
            if (k >= 2) { t = v + cols[3 * (k - 2) + 2]; c += t < v; v = t; }
            r[k] = v; carry = c;
        }
    }
    __syncwarp();
}

// steepest-edge scalars of the pivot (one warp, runs on a side stream concurrently with K1):
// Ghat' = [a^2 Ghat - 2 a nu sigma + nu^2 Gq] / D^2 is evaluated mod 2^(64 WX) with the 2-adic inverse of
// odd(D)^2 and a final shift by 2t (pivot_rule.rs:243-296 in integer form).  WX = LG + max(E2, 4) so that
// the fixed-width update kernel can be used whenever D^2 has at most 256 trailing zero bits.
__global__ void __launch_bounds__(32) k_scalars_se(int L, const u64* __restrict__ G, int n, const int* __restrict__ basis,
                                                   Scalars* sc) {
    __shared__ u64 am[RG_MAXW], dodd[RG_MAXW], x[RG_MAXW], tt[RG_MAXW], xn[RG_MAXW], i2[RG_MAXW],
        ax[RG_MAXW], tmp[RG_MAXW], cols[3 * RG_MAXW];
    if (sc->status != ST_RUN) return;
    const int lane = threadIdx.x;
    const int LU = L + 2, LG = 2 * L + 6;
    // self-contained (reads only D and the pivot element a): it runs on its own stream from the moment the
    // pivot row is chosen, concurrently with k_scalars, the work vector and K1
    const int t = rt_ctz(sc->D, L);
    const int t2 = 2 * t;
    int E2 = (t2 + 63) >> 6;
    if (E2 < 4) E2 = 4;
    const int WX = LG + E2;
    if (lane == 0) {
        // the leaving column is known as soon as the row is: the weight recurrence (side stream) skips it and runs
        // before k_finalize's bookkeeping
        if (sc->pg >= 1) sc->leaving = basis[sc->pg - 1];
        sc->t2 = t2; sc->E2 = E2;
        for (int l = 0; l < WX; ++l) { tmp[l] = l < LU ? sc->a[l] : 0; dodd[l] = l < L ? sc->D[l] : 0; }
        rt_abs(am, tmp, LU);                           // |a| (LU limbs; the new denominator)
        for (int l = LU; l < WX; ++l) am[l] = 0;
        rt_shr(dodd, L, t);
        u64 d0 = dodd[0], y = d0;
        for (int it = 0; it < 6; ++it) y *= 2 - d0 * y;
        for (int l = 0; l < WX; ++l) x[l] = 0;
        x[0] = y;
    }
    __syncwarp();
    // Newton: x <- x (2 - d x), doubling the number of correct limbs
    for (int have = 1; have < WX; have *= 2) {
        int want = have * 2 < WX ? have * 2 : WX;
        warp_mul_lo(tt, dodd, x, want, cols);
        if (lane == 0) {
            rt_neg(tt, want);
            u64 v = tt[0] + 2; u64 c = v < tt[0]; tt[0] = v;
            for (int k = 1; k < want && c; ++k) { tt[k] += 1; c = tt[k] == 0; }
        }
        __syncwarp();
        warp_mul_lo(xn, x, tt, want, cols);
        for (int k = lane; k < want; k += 32) x[k] = xn[k];
        __syncwarp();
    }
    warp_mul_lo(i2, x, x, WX, cols);                 // 1 / odd(D)^2
    for (int k = lane; k < WX; k += 32) ax[k] = k < LU ? am[k] : 0;
    __syncwarp();
    warp_mul_lo(xn, ax, ax, WX, cols);               // a^2
    warp_mul_lo(tmp, xn, i2, WX, cols);
    for (int k = lane; k < WX; k += 32) sc->S1[k] = tmp[k];
    __syncwarp();
    if (lane == 0) rt_add(ax, ax, WX);               // 2a
    __syncwarp();
    warp_mul_lo(tmp, ax, i2, WX, cols);
    for (int k = lane; k < WX; k += 32) sc->S2[k] = tmp[k];
    // Ghat of the entering column: local array, or (column-sharded pricing) the value the selection merge left in Gq
    for (int k = lane; k < WX; k += 32) { u64 v = k < LG ? (G ? G[(size_t)k * n + sc->q] : sc->Gq[k]) : 0; xn[k] = v; if (k < LG) sc->Gq[k] = v; }
    __syncwarp();
    warp_mul_lo(tmp, xn, i2, WX, cols);
    for (int k = lane; k < WX; k += 32) sc->S3[k] = tmp[k];
}

// Generic-width fallback of K1 for E > 2 (D divisible by 2^129 or more): run-time widths.
__global__ void __launch_bounds__(128)
k_update_generic(u64* __restrict__ C, size_t ps, int ld, int nrows, int L, const u64* __restrict__ u,
                 size_t us, const u64* __restrict__ rowp, size_t rs, Scalars* sc) {
    if (sc->status != ST_RUN) return;
    const int E = sc->E;
    const int W = L + E, LU = L + 2;
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    int maxb = 0;
    if (col < ld && i < nrows) {
        u64 buf[8 * RG_MAXW];
        u64* ui = buf; u64* bn = buf + RG_MAXW; u64* cv = buf + 2 * RG_MAXW; u64* rp = buf + 3 * RG_MAXW;
        u64* x1 = buf + 4 * RG_MAXW; u64* x2 = buf + 5 * RG_MAXW; u64* raw = buf + 6 * RG_MAXW; u64* mg = buf + 7 * RG_MAXW;
        if (i == sc->p) { for (int l = 0; l < W; ++l) ui[l] = sc->up[l]; }
        else { rt_load_planar(raw, LU, u, us, i); rt_sext(ui, W, raw, LU); }
        rt_mul_lo(bn, ui, sc->Dinv, W);
        if (sc->sgn > 0) rt_neg(bn, W);
        rt_load_planar(raw, L, C, ps, (size_t)i * ld + col); rt_sext(cv, W, raw, L);
        rt_load_planar(raw, L, rowp, rs, col); rt_sext(rp, W, raw, L);
        rt_mul_lo(x1, sc->A, cv, W);
        rt_mul_lo(x2, bn, rp, W);
        rt_add(x1, x2, W);
        rt_shr(x1, W, sc->t);
        rt_store_planar(C, ps, (size_t)i * ld + col, x1, L);
        rt_abs(mg, x1, L);
        maxb = rt_bitlen_u(mg, L) + 1;
    }
    maxb = warp_max(maxb);
    if ((threadIdx.x & 31) == 0 && maxb) atomicMax(&sc->maxbits_new, maxb);
}

// after the update: basis bookkeeping, D <- |a|, steepest-edge weight of the leaving column
__global__ void k_finalize(int* basis, unsigned char* inbasis, int L, u64* G, int n, int LG,
                           int want_se, const long long* wf, long long* rowf, Scalars* sc, HostMirror* hm) {
    if (threadIdx.x || blockIdx.x) return;
    if (sc->status == ST_RUN) {
        int r = sc->pg - 1;
        int leaving = basis[r];
        sc->leaving = leaving;
        basis[r] = sc->q;
        inbasis[sc->q] = 1;
        if (rowf) rowf[r] = wf[sc->q];
        if (leaving >= 0) {
            inbasis[leaving] = 0;
            if (want_se) for (int l = 0; l < LG; ++l) G[(size_t)l * n + leaving] = sc->Gq[l];
        }
        for (int l = 0; l < L; ++l) sc->D[l] = sc->Dnew[l];
        sc->bits_D = rt_bitlen_u(sc->D, L);
        sc->maxbits_carry = max(sc->maxbits_new, sc->bits_D);     // implicit diagonals hold D
        hm->pivoted = 1; hm->q_done = sc->q; hm->p_done = sc->pg; hm->leaving_done = leaving;
    }
}

__global__ void k_mirror(Scalars* sc, HostMirror* hm, int L) {
    if (threadIdx.x || blockIdx.x) return;
    hm->status = sc->status; hm->q = sc->q; hm->p = sc->pg; hm->leaving = sc->leaving;
    hm->t_next = rt_ctz(sc->D, L);
    hm->bits_D = sc->bits_D; hm->maxbits_carry = sc->maxbits_carry; hm->predicted = sc->predicted;
    hm->found = sc->found; hm->sgn = sc->sgn; hm->maxbits_tmp = sc->maxbits_tmp; hm->nk = sc->nk;
    hm->fatal = sc->fatal;
}

// ---------------------------------------------------------------------------------------------
// K6 work vector  omega_k = sum_{i=1..m} s_i C[i][k]    (w = alpha^T B^-1, carry/mod.rs:575-576;
// with s = basic costs it is the phase switch -pi = -c_B^T B^-1, carry/mod.rs:226-283)
// Stage 1: thread = column, loops a chunk of rows, skips rows with s_i = 0; stage 2 sums chunks.
// ---------------------------------------------------------------------------------------------
template <int LA, int LB>
__device__ __forceinline__ void mul_full_ct(u64 (&r)[LA + LB], const u64 (&a)[LA], const u64 (&b)[LB]) {
    u64 c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
    for (int k = 0; k < LA + LB - 1; ++k) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            int j = k - i;
            if (j >= 0 && j < LB) mac3(c0, c1, c2, a[i], b[j]);
        }
        r[k] = c0;
        c0 = c1; c1 = c2; c2 = 0;
    }
    r[LA + LB - 1] = c0;
}

template <int LN>
__device__ __forceinline__ void neg_n(u64 (&x)[LN]) {
    u64 c = 1;
#pragma unroll
    for (int l = 0; l < LN; ++l) { u64 v = ~x[l] + c; c = (c && v == 0) ? 1 : 0; x[l] = v; }
}

// Stage 1 in unsigned form (IMAD.WIDE product scanning, 1152 multiply-adds per 16-limb entry instead of the 2485
// of a two's complement accumulation at the output width):
//   * the carry entry is biased, x' = C[i][k] + 2^(64L-1) (its top bit flipped): 0 <= x' < 2^(64L);
//   * the factor is taken as magnitude |s_i| and the rows of a chunk are walked in two passes, s_i > 0 then
//     s_i < 0, each into its own partial sum  P = sum x'_ik |s_i| - 2^(64L-1) sum |s_i|  (the bias removed with
//     the chunk's magnitude sum, computed once per block); stage 2 adds the positive and subtracts the negative
//     partials, so no signed product is ever formed;
//   * one column of the product at a time: (c2:c1:c0) += x'_i * |s|_j for i + j = k, then limb k is final.
// Partial slabs: chunk c writes slab 2c (positive rows) and 2c+1 (negative rows).
template <int NA, int NB, int K, int I>
struct ColTerms {
    __device__ static __forceinline__ void run(u32& c0, u32& c1, u32& c2, const u32 (&x)[NA], const u32 (&m)[NB]) {
        if constexpr (I < NA && K - I >= 0 && K - I < NB)
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                         : "+r"(c0), "+r"(c1), "+r"(c2) : "r"(x[I]), "r"(m[K - I]));
        if constexpr (I + 1 < NA && I + 1 <= K) ColTerms<NA, NB, K, I + 1>::run(c0, c1, c2, x, m);
    }
};
template <int NA, int NB, int K, int I>
struct ColPair {     // term I of column K (x_I m_{K-I}) and term I of column K+1 (x_I m_{K+1-I}), then I+1
    __device__ static __forceinline__ void run(u32& a0, u32& a1, u32& a2, u32& b0, u32& b1, u32& b2, const u32 (&x)[NA],
                                               const u32 (&m)[NB]) {
        if constexpr (I < NA && K - I >= 0 && K - I < NB)
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                         : "+r"(a0), "+r"(a1), "+r"(a2) : "r"(x[I]), "r"(m[K - I]));
        if constexpr (I < NA && K + 1 - I >= 0 && K + 1 - I < NB)
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                         : "+r"(b0), "+r"(b1), "+r"(b2) : "r"(x[I]), "r"(m[K + 1 - I]));
        if constexpr (I + 1 < NA && I + 1 <= K + 1) ColPair<NA, NB, K, I + 1>::run(a0, a1, a2, b0, b1, b2, x, m);
    }
};

// Two columns at a time: columns K and K+1 are summed into independent 3-word accumulators (two IMAD.WIDE chains
// that ptxas interleaves: a single chain is bound by the latency of the dependent accumulate), then folded:
// limb K = a0, limb K+1 = a1 + b0, and (a2 + b1 + carry, b2 + carry) seed the next pair.
template <int NA, int NB, int NW, int K>
struct ColScan {
    __device__ static __forceinline__ void run(u32 (&acc)[NW], u32& s0, u32& s1, u32& s2, const u32 (&x)[NA],
                                               const u32 (&m)[NB]) {
        static_assert(NW % 2 == 0 && K % 2 == 0, "columns are taken in pairs");
        u32 a0 = s0, a1 = s1, a2 = s2, b0 = 0, b1 = 0, b2 = 0;
        // the running sum's limbs K and K+1 join the column accumulators
        asm volatile("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.u32 %2, %2, 0;"
                     : "+r"(a0), "+r"(a1), "+r"(a2) : "r"(acc[K]));
        b0 = acc[K + 1];
        // the terms of the two columns are emitted alternately (the inline-asm statements keep their order)
        ColPair<NA, NB, K, (K - NB + 1 > 0 ? K - NB + 1 : 0)>::run(a0, a1, a2, b0, b1, b2, x, m);
        acc[K] = a0;
        u32 t0, t1, t2;
        asm volatile("add.cc.u32 %0, %3, %4;\n\taddc.cc.u32 %1, %5, %6;\n\taddc.u32 %2, %7, 0;"
                     : "=r"(t0), "=r"(t1), "=r"(t2) : "r"(a1), "r"(b0), "r"(a2), "r"(b1), "r"(b2));
        acc[K + 1] = t0;
        s0 = t1; s1 = t2; s2 = 0;
        if constexpr (K + 2 < NW) ColScan<NA, NB, NW, K + 2>::run(acc, s0, s1, s2, x, m);
    }
};

// Four columns at a time (K .. K+3, four independent IMAD.WIDE chains per thread): with three warps per sub-partition
// two chains per warp still leave the multiply pipe waiting on the dependent accumulate.  The four 96-bit column sums
// are folded as A + B 2^32 + C 2^64 + D 2^96 into six words: the low four are limbs K .. K+3, the upper two seed
// the next group (every column sum is < NB 2^64, so word 6 stays zero).
template <int NA, int NB, int K, int I>
struct ColQuadTerms {
    __device__ static __forceinline__ void run(u32 (&a)[3], u32 (&b)[3], u32 (&c)[3], u32 (&d)[3], const u32 (&x)[NA],
                                               const u32 (&m)[NB]) {
        if constexpr (I < NA && K - I >= 0 && K - I < NB)
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                         : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]) : "r"(x[I]), "r"(m[K - I]));
        if constexpr (I < NA && K + 1 - I >= 0 && K + 1 - I < NB)
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                         : "+r"(b[0]), "+r"(b[1]), "+r"(b[2]) : "r"(x[I]), "r"(m[K + 1 - I]));
        if constexpr (I < NA && K + 2 - I >= 0 && K + 2 - I < NB)
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                         : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]) : "r"(x[I]), "r"(m[K + 2 - I]));
        if constexpr (I < NA && K + 3 - I >= 0 && K + 3 - I < NB)
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                         : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]) : "r"(x[I]), "r"(m[K + 3 - I]));
        if constexpr (I + 1 < NA && I + 1 <= K + 3) ColQuadTerms<NA, NB, K, I + 1>::run(a, b, c, d, x, m);
    }
};
template <int NA, int NB, int NW, int K>
struct ColScan4 {
    __device__ static __forceinline__ void run(u32 (&acc)[NW], u32& s0, u32& s1, u32& s2, const u32 (&x)[NA],
                                               const u32 (&m)[NB]) {
        if constexpr (K + 4 <= NW) {
            u32 a[3] = {s0, s1, s2}, b[3] = {acc[K + 1], 0, 0}, c[3] = {acc[K + 2], 0, 0}, d[3] = {acc[K + 3], 0, 0};
            asm volatile("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.u32 %2, %2, 0;"
                         : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]) : "r"(acc[K]));
            ColQuadTerms<NA, NB, K, (K - NB + 1 > 0 ? K - NB + 1 : 0)>::run(a, b, c, d, x, m);
            u32 r1 = a[1], r2 = a[2], r3 = 0, r4 = 0, r5 = 0;
            asm volatile("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\t"
                         "addc.cc.u32 %3, %3, 0;\n\taddc.u32 %4, %4, 0;"
                         : "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5) : "r"(b[0]), "r"(b[1]), "r"(b[2]));
            asm volatile("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %6;\n\taddc.u32 %3, %3, 0;"
                         : "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5) : "r"(c[0]), "r"(c[1]), "r"(c[2]));
            asm volatile("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, %5;"
                         : "+r"(r3), "+r"(r4), "+r"(r5) : "r"(d[0]), "r"(d[1]), "r"(d[2]));
            acc[K] = a[0]; acc[K + 1] = r1; acc[K + 2] = r2; acc[K + 3] = r3;
            s0 = r4; s1 = r5; s2 = 0;
            ColScan4<NA, NB, NW, K + 4>::run(acc, s0, s1, s2, x, m);
        } else if constexpr (K + 2 <= NW) {
            ColScan<NA, NB, NW, K>::run(acc, s0, s1, s2, x, m);     // NW = 2 (mod 4): the last two columns as a pair
        }
    }
};

// Thread mapping: a block is 32 column slots x 4 row groups (one warp each: a warp reads 32 consecutive entries of
// one row, 256 B per limb plane); the four groups split the rows of the chunk and their partial sums are added
// through shared memory at the end of a pass, so a chunk still produces one slab per pass.
template <int L, int LSRC, int LOUT>
__global__ void __launch_bounds__(128, 3)
k_colsum1(const u64* __restrict__ C, size_t ps, int ld, int m, int rows_per_chunk, const int* __restrict__ klist,
          const u64* __restrict__ s, size_t ss, u64* __restrict__ part, int pcols, const Scalars* sc) {
    constexpr int RB = 64;                 // rows whose factors are staged in shared memory at a time
    constexpr int WS = (L + LSRC + 1) < LOUT ? (L + LSRC + 1) : LOUT;   // 64-bit limbs of a partial sum
    constexpr int NA = 2 * L, NB = 2 * LSRC, NW = 2 * WS;
    constexpr int LQ = LSRC + 1;           // 64-bit limbs of a chunk's magnitude sum (up to 2^32 rows)
    constexpr int NU = 2 * LQ;
    static_assert(NW >= NA + NB + 1 || WS == LOUT, "partial sums need one limb of headroom");
    __shared__ u32 sMag[RB][NB];
    __shared__ signed char sSgn[RB];
    __shared__ u64 sPart[2][2][LQ];        // [staging warp][pass]: magnitude sums of one batch
    __shared__ u64 sUsum[2][LQ];           // per pass: sum of |s_i| over the chunk's rows of that sign
    __shared__ u32 sRed[3][NW][32];        // partial sums of row groups 1..3
    __shared__ unsigned sMask[2][2];       // [staging warp][pass]: rows of the batch whose factor has that sign
    if (sc->status != ST_RUN) return;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int kidx = blockIdx.x * 32 + lane;
    // list mode (klist != nullptr): C is the packed block, `ld` its capacity, column slot = list position
    const int k = klist ? (kidx < sc->nk ? kidx : ld) : kidx;
    const int r0 = 1 + blockIdx.y * rows_per_chunk;
    const int r1 = min(m + 1, r0 + rows_per_chunk);
    if (threadIdx.x < 2 * LQ) sUsum[threadIdx.x / LQ][threadIdx.x % LQ] = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int want = pass == 0 ? 1 : -1;
        u32 acc[NW];
#pragma unroll
        for (int l = 0; l < NW; ++l) acc[l] = 0;
        for (int base = r0; base < r1; base += RB) {
            // a chunk of a single batch (list mode: 64 rows) keeps its staged factors for the second pass
            const bool stage = pass == 0 || r1 - r0 > RB;
            if (stage) {
            __syncthreads();
            if (threadIdx.x < RB) {        // warps 0 and 1 stage the factors of rows base .. base+63
                int i = base + threadIdx.x;
                int sg = 0;
                u64 x[LSRC];
#pragma unroll
                for (int l = 0; l < LSRC; ++l) x[l] = 0;
                if (i < r1) {
                    load_planar<LSRC>(x, s, ss, (size_t)i);
                    const bool neg = (i64)x[LSRC - 1] < 0;
                    u64 o = 0;
                    if (neg) {
                        u64 c = 1;
#pragma unroll
                        for (int l = 0; l < LSRC; ++l) { u64 v = ~x[l] + c; c = (c && v == 0) ? 1 : 0; x[l] = v; }
                    }
#pragma unroll
                    for (int l = 0; l < LSRC; ++l) {
                        sMag[threadIdx.x][2 * l] = (u32)x[l]; sMag[threadIdx.x][2 * l + 1] = (u32)(x[l] >> 32); o |= x[l];
                    }
                    sg = o == 0 ? 0 : (neg ? -1 : 1);
                }
                sSgn[threadIdx.x] = (signed char)sg;
                {   // the rows of either sign as bit masks: the row groups below take every fourth row OF THE PASS'S
                    // SIGN, so they finish together (every fourth row of the batch left them up to 30 % apart)
                    const unsigned mp = __ballot_sync(0xffffffffu, sg == 1), mn = __ballot_sync(0xffffffffu, sg == -1);
                    if (lane == 0) { sMask[grp][0] = mp; sMask[grp][1] = mn; }
                }
                if (pass == 0) {           // magnitude sums of both signs, once per batch (bias removal below)
#pragma unroll
                    for (int p2 = 0; p2 < 2; ++p2) {
                        u64 q[LQ];
                        const bool mine = sg == (p2 == 0 ? 1 : -1);
#pragma unroll
                        for (int l = 0; l < LQ; ++l) q[l] = (mine && l < LSRC) ? x[l] : 0;
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) {
                            u64 o2[LQ];
#pragma unroll
                            for (int l = 0; l < LQ; ++l) o2[l] = __shfl_down_sync(0xffffffffu, q[l], off);
                            add_n<LQ>(q, o2);
                        }
                        if (lane == 0) {
#pragma unroll
                            for (int l = 0; l < LQ; ++l) sPart[grp][p2][l] = q[l];
                        }
                    }
                }
            }
            __syncthreads();
            if (pass == 0 && threadIdx.x < 2) {      // thread p2 folds the two staging warps' sums into the chunk sum
                u64 a[LQ], b0[LQ], b1[LQ];
#pragma unroll
                for (int l = 0; l < LQ; ++l) { a[l] = sUsum[threadIdx.x][l]; b0[l] = sPart[0][threadIdx.x][l]; b1[l] = sPart[1][threadIdx.x][l]; }
                add_n<LQ>(a, b0); add_n<LQ>(a, b1);
#pragma unroll
                for (int l = 0; l < LQ; ++l) sUsum[threadIdx.x][l] = a[l];
            }
            }
            if (k >= ld) continue;
            const unsigned m0 = sMask[0][pass], m1 = sMask[1][pass];
            const int n0 = __popc(m0), cnt = n0 + __popc(m1);
            for (int idx = grp; idx < cnt; idx += 4) {     // rows with a zero factor are never read
                const int r = idx < n0 ? (int)__fns(m0, 0, idx + 1) : 32 + (int)__fns(m1, 0, idx - n0 + 1);
                u32 x[NA], mg[NB];
                {
                    u64 xl[L];
                    load_planar<L>(xl, C, ps, (size_t)(base + r) * ld + k);
                    xl[L - 1] ^= 0x8000000000000000ull;                      // + 2^(64L-1)
#pragma unroll
                    for (int l = 0; l < L; ++l) { x[2 * l] = (u32)xl[l]; x[2 * l + 1] = (u32)(xl[l] >> 32); }
                }
#pragma unroll
                for (int l = 0; l < NB; ++l) mg[l] = sMag[r][l];
                u32 c0 = 0, c1 = 0, c2 = 0;
                ColScan4<NA, NB, NW, 0>::run(acc, c0, c1, c2, x, mg);
            }
        }
        // add the four row groups' partial sums (group 0 keeps the result)
        __syncthreads();
        if (grp > 0) {
#pragma unroll
            for (int l = 0; l < NW; ++l) sRed[grp - 1][l][lane] = acc[l];
        }
        __syncthreads();                   // (sUsum[pass] is complete here too)
        if (grp == 0 && k < ld) {
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(acc[0]) : "r"(sRed[g][0][lane]));
#pragma unroll
                for (int l = 1; l < NW; ++l) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(acc[l]) : "r"(sRed[g][l][lane]));
            }
            // remove the bias: acc -= (sum |s_i|) << (64L - 1)   (word offset 2L-1, bit offset 31)
            const u32* us = reinterpret_cast<const u32*>(sUsum[pass]);
            u32 bw = 0;
#pragma unroll
            for (int q = NA - 1; q < NW; ++q) {
                const int j = q - (NA - 1);
                const u32 lo = (j - 1 >= 0 && j - 1 < NU) ? us[(j - 1 >= 0 && j - 1 < NU) ? j - 1 : 0] : 0u;
                const u32 hi = j < NU ? us[j < NU ? j : 0] : 0u;
                const u32 sub = (hi << 31) | (lo >> 1);
                u32 v = acc[q] - sub; u32 b1 = acc[q] < sub; u32 v2 = v - bw; u32 b2 = v < bw;
                acc[q] = v2; bw = b1 + b2;
            }
            u64 out[LOUT];
            const u64 sgo = (int)acc[NW - 1] < 0 ? ~0ull : 0ull;
#pragma unroll
            for (int l = 0; l < LOUT; ++l) out[l] = l < WS ? ((u64)acc[2 * l] | ((u64)acc[2 * l + 1] << 32)) : sgo;
            // partial sums are indexed by column (dense mode) or by list position (list mode), row stride pcols
            store_planar<LOUT>(part + (size_t)(2 * blockIdx.y + pass) * LOUT * pcols, (size_t)pcols, (size_t)k, out);
        }
    }
}

// Stage 2 for the listed columns (list mode): one block per list position; the threads stride over the partial
// slabs (odd slabs subtract), the block's sums are folded with warp shuffles and shared memory.
template <int LOUT>
__global__ void __launch_bounds__(128)
k_colsum2_list(const u64* __restrict__ part, int nslabs, int pcols, const int* __restrict__ klist, int ld,
               int negate, u64* __restrict__ out, Scalars* sc, int packed_stride = 0) {
    // packed_stride != 0 (row-sharded split mode): the sums are written by list position, plane stride packed_stride,
    // for the small all-gather of the listed columns' partial sums (k_list_sum_scatter finishes them)
    __shared__ u64 sW[4][LOUT];
    if (sc->status != ST_RUN) return;
    const int t = blockIdx.x;
    if (t >= sc->nk) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 acc[LOUT];
#pragma unroll
    for (int l = 0; l < LOUT; ++l) acc[l] = 0;
    for (int c = threadIdx.x; c < nslabs; c += blockDim.x) {
        u64 x[LOUT];
        load_planar<LOUT>(x, part + (size_t)c * LOUT * pcols, (size_t)pcols, (size_t)t);
        if (c & 1) neg_n<LOUT>(x);
        add_n<LOUT>(acc, x);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        u64 o[LOUT];
#pragma unroll
        for (int l = 0; l < LOUT; ++l) o[l] = __shfl_down_sync(0xffffffffu, acc[l], off);
        add_n<LOUT>(acc, o);
    }
    if (lane == 0) {
#pragma unroll
        for (int l = 0; l < LOUT; ++l) sW[warp][l] = acc[l];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 4; ++w) {
            u64 o[LOUT];
#pragma unroll
            for (int l = 0; l < LOUT; ++l) o[l] = sW[w][l];
            add_n<LOUT>(acc, o);
        }
        if (negate) neg_n<LOUT>(acc);
        if (packed_stride) store_planar<LOUT>(out, (size_t)packed_stride, (size_t)t, acc);
        else {
            store_planar<LOUT>(out, (size_t)ld, (size_t)klist[t], acc);
            atomicMax(&sc->maxbits_tmp, bitlen_signed<LOUT>(acc));
        }
    }
}
// row-sharded split mode: sum the ranks' packed partial sums of the listed columns and scatter them into the
// work vector (thread = list position)
template <int LOUT>
__global__ void __launch_bounds__(128)
k_list_sum_scatter(const u64* __restrict__ recv, int world, int packed_stride, const int* __restrict__ klist, int ld,
                   u64* __restrict__ out, Scalars* sc) {
    if (sc->status != ST_RUN) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int bl = 0;
    if (t < sc->nk) {
        u64 acc[LOUT];
#pragma unroll
        for (int l = 0; l < LOUT; ++l) acc[l] = 0;
        for (int r = 0; r < world; ++r) {
            u64 x[LOUT];
            load_planar<LOUT>(x, recv + (size_t)r * LOUT * packed_stride, (size_t)packed_stride, (size_t)t);
            add_n<LOUT>(acc, x);
        }
        store_planar<LOUT>(out, (size_t)ld, (size_t)klist[t], acc);
        bl = bitlen_signed<LOUT>(acc);
    }
    bl = warp_max(bl);
    if ((threadIdx.x & 31) == 0 && bl) atomicMax(&sc->maxbits_tmp, bl);
}

// List mode, stage 0: compact the local rows whose factor s_i is non-zero (order is irrelevant: the sums
// below are exact integers).  sc->nnz_s is zeroed by k_reset_iter.
template <int LSRC>
__global__ void __launch_bounds__(256)
k_nzrows(const u64* __restrict__ s, size_t ss, int nloc, int* __restrict__ nzrows, Scalars* sc) {
    if (sc->status != ST_RUN) return;
    int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    bool nz = false;
    if (i <= nloc) {
        u64 o = 0;
#pragma unroll
        for (int l = 0; l < LSRC; ++l) o |= s[(size_t)l * ss + i];
        nz = o != 0;
    }
    unsigned mask = __ballot_sync(0xffffffffu, nz);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(&sc->nnz_s, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (nz) nzrows[base + __popc(mask & ((1u << lane) - 1))] = i;
}

// List mode, stage 1+2 in one: one WARP per non-trivial column; lanes stride over the compacted non-zero
// rows, multiply in sign-magnitude form, and the 32 partial sums are combined with a shuffle tree.
template <int L, int LSRC, int LOUT>
__global__ void __launch_bounds__(128)
k_colsum_list(const u64* __restrict__ C, size_t ps, int ld, const int* __restrict__ klist,
              const int* __restrict__ nzrows, const u64* __restrict__ s, size_t ss, u64* __restrict__ part,
              int pcols, const Scalars* sc) {
    if (sc->status != ST_RUN) return;
    const int lane = threadIdx.x & 31;
    const int kidx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (kidx >= sc->nk) return;
    const int k = klist[kidx];
    // blockIdx.y: segment of the non-zero row list (sparse factors: most segments are a handful of rows;
    // dense factors: enough warps in flight to hide the latency of the scattered loads)
    const int nnz = sc->nnz_s;
    const int seg_len = (nnz + gridDim.y - 1) / gridDim.y;
    const int e0 = blockIdx.y * seg_len, e1 = min(nnz, e0 + seg_len);
    u64 acc[LOUT];
#pragma unroll
    for (int l = 0; l < LOUT; ++l) acc[l] = 0;
    for (int e = e0 + lane; e < e1; e += 32) {
        const int i = nzrows[e];
        u64 x[L];
        load_planar<L>(x, C, ps, (size_t)i * ld + k);
        u64 any = 0;
#pragma unroll
        for (int l = 0; l < L; ++l) any |= x[l];
        if (any == 0) continue;
        u64 sm[LSRC];
        load_planar<LSRC>(sm, s, ss, (size_t)i);
        int sgn = 1;
        if ((i64)sm[LSRC - 1] < 0) {
            u64 c = 1;
#pragma unroll
            for (int l = 0; l < LSRC; ++l) { u64 v = ~sm[l] + c; c = (c && v == 0) ? 1 : 0; sm[l] = v; }
            sgn = -sgn;
        }
        if ((i64)x[L - 1] < 0) {
            u64 c = 1;
#pragma unroll
            for (int l = 0; l < L; ++l) { u64 v = ~x[l] + c; c = (c && v == 0) ? 1 : 0; x[l] = v; }
            sgn = -sgn;
        }
        u64 pr[LSRC + L];
        mul_full_ct<LSRC, L>(pr, sm, x);
        if (sgn > 0) {
            u64 cf = 0;
#pragma unroll
            for (int l = 0; l < LOUT; ++l) {
                u64 b = l < LSRC + L ? pr[l] : 0;
                u64 v = acc[l] + b; u64 c1 = v < b; u64 v2 = v + cf; u64 c2 = v2 < v;
                acc[l] = v2; cf = c1 + c2;
            }
        } else {
            u64 bf = 0;
#pragma unroll
            for (int l = 0; l < LOUT; ++l) {
                u64 b = l < LSRC + L ? pr[l] : 0;
                u64 v = acc[l] - b; u64 b1 = acc[l] < b; u64 v2 = v - bf; u64 b2 = v < bf;
                acc[l] = v2; bf = b1 + b2;
            }
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        u64 o[LOUT];
#pragma unroll
        for (int l = 0; l < LOUT; ++l) o[l] = __shfl_down_sync(0xffffffffu, acc[l], off);
        add_n<LOUT>(acc, o);
    }
    // partial sums are indexed by list position, one slab per segment (summed by k_colsum2)
    if (lane == 0) store_planar<LOUT>(part + (size_t)blockIdx.y * LOUT * pcols, (size_t)pcols, (size_t)kidx, acc);
}

// `triv` (may be null): column k is D e_k implicitly, so its column sum is s_k * D, contributed by the rank
// that owns row k (s: the factor vector, LSRC limbs, local row index k - row_lo).
// alt != 0: the partial slabs alternate in sign (slab 2c: rows with a positive factor, slab 2c+1: negative ones,
// see k_colsum1): odd slabs are subtracted.
template <int LOUT, int LSRC = 1, int LDT = 0>
__global__ void __launch_bounds__(64)
k_colsum2(const u64* __restrict__ part, int ld, int chunks, int negate, u64* __restrict__ out,
          Scalars* sc, const unsigned char* __restrict__ triv = nullptr, const u64* __restrict__ s = nullptr,
          size_t ss = 0, int LD = 0, const int* __restrict__ kpos = nullptr, int pcols = 0, int alt = 0,
          int global_s = 0) {
    // global_s != 0 (row-sharded split mode): `s` is the WHOLE factor vector (all-gathered), so every rank writes
    // the implicit columns' sums s_k * D itself and no exchange of them is needed
    if (sc->status != ST_RUN) return;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int bl = 0;
    if (k < ld) {
        bool listed = false;
        u64 acc[LOUT];
#pragma unroll
        for (int l = 0; l < LOUT; ++l) acc[l] = 0;
        if (triv && triv[k]) {
            int li = global_s ? k : k - sc->row_lo;   // (local) carry row of the diagonal entry
            if (li >= 1 && (global_s || li <= sc->nloc)) {
                u64 x[LSRC];
                load_planar<LSRC>(x, s, ss, (size_t)li);
                u64 any = 0;
#pragma unroll
                for (int l = 0; l < LSRC; ++l) any |= x[l];
                if (any && LDT > 0) {
                    // acc = x * D in sign-magnitude form: |x| (LSRC limbs) times D (LDT limbs), full product
                    constexpr int LDC = LDT > 0 ? LDT : 1;
                    u64 dd[LDC];
#pragma unroll
                    for (int l = 0; l < LDC; ++l) dd[l] = sc->D[l];
                    const bool neg = (i64)x[LSRC - 1] < 0;
                    if (neg) {
                        u64 c = 1;
#pragma unroll
                        for (int l = 0; l < LSRC; ++l) { u64 v = ~x[l] + c; c = (c && v == 0) ? 1 : 0; x[l] = v; }
                    }
                    u64 pr[LSRC + LDC];
                    mul_full_ct<LSRC, LDC>(pr, x, dd);
#pragma unroll
                    for (int l = 0; l < LOUT; ++l) acc[l] = l < LSRC + LDC ? pr[l] : 0;
                    if (neg) {
                        u64 c = 1;
#pragma unroll
                        for (int l = 0; l < LOUT; ++l) { u64 v = ~acc[l] + c; c = (c && v == 0) ? 1 : 0; acc[l] = v; }
                    }
                } else if (any) {
                    // acc = x * D (x signed LSRC limbs, D positive LD limbs): limb-by-limb small products
                    u64 xe[LOUT];
                    u64 sg = (i64)x[LSRC - 1] < 0 ? ~0ull : 0ull;
#pragma unroll
                    for (int l = 0; l < LOUT; ++l) xe[l] = l < LSRC ? x[l] : sg;
                    u64 de[LOUT];
#pragma unroll
                    for (int l = 0; l < LOUT; ++l) de[l] = l < LD ? sc->D[l] : 0;
                    mul_lo<LOUT>(acc, xe, de);
                }
            }
        } else if (kpos) {
            // list mode: the listed columns (column 0 and every k with a list position) are summed and stored by
            // k_colsum2_list; what remains here is padding beyond m: zero
            listed = k == 0 || kpos[k] > 0;
        } else {
            const int pc = kpos ? pcols : ld;             // list mode: partials live at the list position
            const size_t idx = kpos ? (size_t)kpos[k] : (size_t)k;
            int c = 0;
            for (; c + 4 <= chunks; c += 4) {             // four chunks of loads in flight per thread
                u64 x0[LOUT], x1[LOUT], x2[LOUT], x3[LOUT];
                load_planar<LOUT>(x0, part + (size_t)(c + 0) * LOUT * pc, (size_t)pc, idx);
                load_planar<LOUT>(x1, part + (size_t)(c + 1) * LOUT * pc, (size_t)pc, idx);
                load_planar<LOUT>(x2, part + (size_t)(c + 2) * LOUT * pc, (size_t)pc, idx);
                load_planar<LOUT>(x3, part + (size_t)(c + 3) * LOUT * pc, (size_t)pc, idx);
                if (alt) { neg_n<LOUT>(x1); neg_n<LOUT>(x3); }     // c is a multiple of 4: odd slabs are c+1, c+3
                add_n<LOUT>(x0, x1); add_n<LOUT>(x2, x3); add_n<LOUT>(acc, x0); add_n<LOUT>(acc, x2);
            }
            for (; c < chunks; ++c) {
                u64 x[LOUT];
                load_planar<LOUT>(x, part + (size_t)c * LOUT * pc, (size_t)pc, idx);
                if (alt && (c & 1)) neg_n<LOUT>(x);
                add_n<LOUT>(acc, x);
            }
        }
        if (negate) {
            u64 c = 1;
#pragma unroll
            for (int l = 0; l < LOUT; ++l) { u64 v = ~acc[l] + c; c = (c && v == 0) ? 1 : 0; acc[l] = v; }
        }
        if (k < ld && !listed) {
            store_planar<LOUT>(out, (size_t)ld, (size_t)k, acc);
            bl = bitlen_signed<LOUT>(acc);
        }
    }
    bl = warp_max(bl);
    if ((threadIdx.x & 31) == 0 && bl) atomicMax(&sc->maxbits_tmp, bl);
}

// weighted problems: s_i = u_i * (W / w_B(i))^2, the factor vector of the work-vector column sum
template <int L>
__global__ void __launch_bounds__(256)
k_scale_u(const u64* __restrict__ u, size_t us, int nloc, const long long* __restrict__ rowf,
          u64* __restrict__ out, size_t os, Scalars* sc) {
    constexpr int LU = L + 2;
    if (sc->status != ST_RUN) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;    // local carry row
    if (i > nloc) i = 0;                              // (keeps the warp converged for the bit-length reduction)
    u64 x[LU], acc[LU + 1];
    load_planar<LU>(x, u, us, (size_t)i);
#pragma unroll
    for (int l = 0; l <= LU; ++l) acc[l] = 0;
    if (i >= 1) {
        long long f = rowf[sc->row_lo + i - 1];
        mac_small<LU + 1, LU>(acc, x, f * f);
    }
    store_planar<LU + 1>(out, os, (size_t)i, acc);
    int bl = warp_max(i >= 1 ? bitlen_signed<LU + 1>(acc) : 0);
    if ((threadIdx.x & 31) == 0 && bl) atomicMax(&sc->maxbits_s, bl);
    // row-sharded: the other ranks' rows are not seen here; maxbits_u is already the maximum over all ranks
    // (k_ratio_merge) and the row factors are below 2^31, so this bounds every rank's entries
    if (sc->world > 1 && blockIdx.x == 0 && threadIdx.x == 0) atomicMax(&sc->maxbits_s, sc->maxbits_u + 62);
}

// basic cost of every local row as an (nloc+1)-vector of LSRC = 1 limb (artificial / inert rows: 0)
__global__ void k_basic_costs(const int* basis, const long long* cost, int nloc, int row_lo, u64* s) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nloc) return;
    long long c = 0;
    if (i >= 1) { int j = basis[row_lo + i - 1]; c = j >= 0 ? cost[j] : 0; }
    s[i] = (u64)c;
}
// row 0 of the carry <- tmprow (truncated to L limbs; caller has checked it fits)
__global__ void k_store_row0(u64* C, size_t ps, int ld, int L, const u64* tmp, int LT) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ld) return;
    for (int l = 0; l < L; ++l) C[(size_t)l * ps + k] = tmp[(size_t)l * ld + k];
}
__global__ void k_max_into_carrybits(Scalars* sc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) sc->maxbits_carry = max(sc->maxbits_carry, sc->maxbits_tmp);
}
__global__ void k_reset_tmpbits(Scalars* sc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) sc->maxbits_tmp = 0;
}

// ---------------------------------------------------------------------------------------------
// K5 steepest-edge weights.
//   init (identity carry):  Ghat_j = 1 + |a_j|^2                       (initial_gamma, pivot_rule.rs:299-305)
//   init (general carry):   Ghat_j = D^2 + sum_i (C[i][1..m] . a_j)^2  block per column
//   update: Ghat'_j = [a^2 Ghat_j - 2 a nu_j sigma_j + nu_j^2 Ghat_q] / D^2   (after_basis_update :243-296)
// ---------------------------------------------------------------------------------------------
__global__ void k_gamma_init_identity(int n, int j0, const long long* colptr, const int* rowidx, const long long* vals,
                                      const unsigned char* inbasis, const long long* wf, const long long* rowf,
                                      u64* G, int LG) {
    int j = j0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    u64 acc[6] = {0, 0, 0, 0, 0, 0};
    if (!inbasis[j]) {
        // Ghat_j = (W/w_j)^2 + sum_i (a_ij W / w_B(i))^2     (D = 1)
        u64 f = wf ? (u64)wf[j] : 1ull;
        mac3(acc[0], acc[1], acc[2], f, f);
        for (long long k = colptr[j]; k < colptr[j + 1]; ++k) {
            long long v = vals[k];
            u64 mg = v < 0 ? (u64)(-v) : (u64)v;
            u64 rf = rowf ? (u64)rowf[rowidx[k]] : 1ull;
            // t = mg * rf (up to 94 bits), acc += t^2 (up to 188 bits)
            u64 t0 = mg * rf, t1 = __umul64hi(mg, rf);
            u64 sq[4] = {0, 0, 0, 0};
            u64 c0 = 0, c1 = 0, c2 = 0;
            mac3(c0, c1, c2, t0, t0); sq[0] = c0; c0 = c1; c1 = c2; c2 = 0;
            mac3(c0, c1, c2, t0, t1); mac3(c0, c1, c2, t1, t0); sq[1] = c0; c0 = c1; c1 = c2; c2 = 0;
            mac3(c0, c1, c2, t1, t1); sq[2] = c0; sq[3] = c1;
            u64 cf = 0;
            for (int l = 0; l < 6; ++l) {
                u64 b = l < 4 ? sq[l] : 0;
                u64 x = acc[l] + b; u64 k1 = x < b; u64 x2 = x + cf; u64 k2 = x2 < x;
                acc[l] = x2; cf = k1 + k2;
            }
        }
    }
    for (int l = 0; l < LG; ++l) G[(size_t)l * n + j] = l < 6 ? acc[l] : 0;
}

template <int L>
__global__ void __launch_bounds__(128)
k_gamma_init_general(const u64* __restrict__ C, size_t ps, int ld, int m, int n,
                     const long long* __restrict__ colptr, const int* __restrict__ rowidx,
                     const long long* __restrict__ vals, const unsigned char* __restrict__ inbasis,
                     u64* __restrict__ G, int add_d2, const long long* __restrict__ wf,
                     const long long* __restrict__ rowf, const unsigned char* __restrict__ triv,
                     const int* __restrict__ kpos, const Scalars* sc) {
    // list mode (triv != nullptr): C is the packed block, ld its capacity, column ck at position kpos[ck]
    constexpr int LU = L + 2, LG = 2 * L + 6;
    __shared__ u64 sAcc[4][LG];
    int j = blockIdx.x;
    if (j >= n) return;
    if (inbasis[j]) {
        if (threadIdx.x < LG) G[(size_t)threadIdx.x * n + j] = 0;
        return;
    }
    long long k0 = colptr[j], k1 = colptr[j + 1];
    u64 acc[LG];
#pragma unroll
    for (int l = 0; l < LG; ++l) acc[l] = 0;
    for (int i = 1 + threadIdx.x; i <= m; i += blockDim.x) {
        u64 nu[LU];
#pragma unroll
        for (int l = 0; l < LU; ++l) nu[l] = 0;
        for (long long k = k0; k < k1; ++k) {
            u64 x[L];
            const int ck = 1 + rowidx[k];
            if (triv && triv[ck]) {                  // implicit D e_ck
                if (sc->row_lo + i != ck) continue;
#pragma unroll
                for (int l = 0; l < L; ++l) x[l] = sc->D[l];
            } else {
                load_planar<L>(x, C, ps, (size_t)i * ld + (triv ? kpos[ck] : ck));
            }
            mac_small<LU, L>(nu, x, vals[k]);
        }
        if ((i64)nu[LU - 1] < 0) {
            u64 c = 1;
#pragma unroll
            for (int l = 0; l < LU; ++l) { u64 v = ~nu[l] + c; c = (c && v == 0) ? 1 : 0; nu[l] = v; }
        }
        // scale by the row factor W / w_B(i) (1 for integer problems)
        u64 nus[LU + 1];
        {
            u64 rf = rowf ? (u64)rowf[sc->row_lo + i - 1] : 1ull;
            u64 carry = 0;
#pragma unroll
            for (int l = 0; l < LU; ++l) {
                u64 lo = nu[l] * rf, hi = __umul64hi(nu[l], rf);
                u64 v = lo + carry; u64 c1 = v < lo;
                nus[l] = v; carry = hi + c1;
            }
            nus[LU] = carry;
        }
        u64 sq[2 * LU + 2];
        mul_full_ct<LU + 1, LU + 1>(sq, nus, nus);
        u64 cf = 0;
#pragma unroll
        for (int l = 0; l < LG; ++l) {
            u64 b = l < 2 * LU + 2 ? sq[l] : 0;
            u64 v = acc[l] + b; u64 c1 = v < b; u64 v2 = v + cf; u64 c2 = v2 < v;
            acc[l] = v2; cf = c1 + c2;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u64 other[LG];
#pragma unroll
        for (int l = 0; l < LG; ++l) other[l] = __shfl_down_sync(0xffffffffu, acc[l], o);
        add_n<LG>(acc, other);
    }
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int l = 0; l < LG; ++l) sAcc[warp][l] = acc[l];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (blockDim.x >> 5); ++w) {
            u64 other[LG];
#pragma unroll
            for (int l = 0; l < LG; ++l) other[l] = sAcc[w][l];
            add_n<LG>(acc, other);
        }
        if (add_d2) {    // row-sharded: only rank 0 contributes the D^2 term to the sum of parts
            // + (D W / w_j)^2
            u64 d[L + 1], d2[2 * L + 2];
            {
                u64 f = wf ? (u64)wf[j] : 1ull;
                u64 carry = 0;
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    u64 x = sc->D[l];
                    u64 lo = x * f, hi = __umul64hi(x, f);
                    u64 v = lo + carry; u64 c1 = v < lo;
                    d[l] = v; carry = hi + c1;
                }
                d[L] = carry;
            }
            mul_full_ct<L + 1, L + 1>(d2, d, d);
            u64 cf = 0;
#pragma unroll
            for (int l = 0; l < LG; ++l) {
                u64 b = l < 2 * L + 2 ? d2[l] : 0;
                u64 v = acc[l] + b; u64 c1 = v < b; u64 v2 = v + cf; u64 c2 = v2 < v;
                acc[l] = v2; cf = c1 + c2;
            }
        }
        store_planar<LG>(G, (size_t)n, (size_t)j, acc);
    }
}

// General-basis initialisation, row by row (any carry mode, CSC and dense-block columns alike): with
// nu_ij = (row i of the carry) . a_j staged by the column-dot machinery,
//     Ghat_j = (D W / w_j)^2 + sum_i (nu_ij W / w_B(i))^2          (initial_gamma, pivot_rule.rs:299-305)
// k_gamma_seed writes the first term, k_gamma_accum adds one row's squares.
template <int L>
__global__ void __launch_bounds__(128)
k_gamma_seed(int n, ColOwn own, const unsigned char* __restrict__ inbasis, const long long* __restrict__ wf,
             u64* __restrict__ G, const Scalars* sc) {
    constexpr int LG = 2 * L + 6;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || !own.owned(j)) return;
    u64 acc[LG];
#pragma unroll
    for (int l = 0; l < LG; ++l) acc[l] = 0;
    if (!inbasis[j]) {
        u64 d[L + 1], d2[2 * L + 2];
        u64 f = wf ? (u64)wf[j] : 1ull;
        u64 carry = 0;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            u64 x = sc->D[l];
            u64 lo = x * f, hi = __umul64hi(x, f);
            u64 v = lo + carry; u64 c1 = v < lo;
            d[l] = v; carry = hi + c1;
        }
        d[L] = carry;
        mul_full_ct<L + 1, L + 1>(d2, d, d);
#pragma unroll
        for (int l = 0; l < 2 * L + 2; ++l) acc[l] = d2[l];
    }
    store_planar<LG>(G, (size_t)n, (size_t)j, acc);
}
template <int L>
__global__ void __launch_bounds__(128)
k_gamma_accum(int n, ColOwn own, const unsigned char* __restrict__ inbasis, const u64* __restrict__ nu,
              const long long* __restrict__ rowf, int row, u64* __restrict__ G, const Scalars* sc) {
    constexpr int LU = L + 2, LG = 2 * L + 6;
    if (sc->status != ST_RUN) return;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || !own.owned(j) || inbasis[j]) return;
    u64 x[LU];
    load_planar<LU>(x, nu, (size_t)n, (size_t)j);
    u64 any = 0;
#pragma unroll
    for (int l = 0; l < LU; ++l) any |= x[l];
    if (any == 0) return;
    if ((i64)x[LU - 1] < 0) {
        u64 c = 1;
#pragma unroll
        for (int l = 0; l < LU; ++l) { u64 v = ~x[l] + c; c = (c && v == 0) ? 1 : 0; x[l] = v; }
    }
    u64 xs[LU + 1];
    {
        u64 rf = rowf ? (u64)rowf[row] : 1ull;
        u64 carry = 0;
#pragma unroll
        for (int l = 0; l < LU; ++l) {
            u64 lo = x[l] * rf, hi = __umul64hi(x[l], rf);
            u64 v = lo + carry; u64 c1 = v < lo;
            xs[l] = v; carry = hi + c1;
        }
        xs[LU] = carry;
    }
    u64 sq[2 * LU + 2];
    mul_full_ct<LU + 1, LU + 1>(sq, xs, xs);
    u64 g[LG];
    load_planar<LG>(g, G, (size_t)n, (size_t)j);
    u64 cf = 0;
#pragma unroll
    for (int l = 0; l < LG; ++l) {
        u64 b = l < 2 * LU + 2 ? sq[l] : 0;
        u64 v = g[l] + b; u64 c1 = v < b; u64 v2 = v + cf; u64 c2 = v2 < v;
        g[l] = v2; cf = c1 + c2;
    }
    store_planar<LG>(G, (size_t)n, (size_t)j, g);
}
// general-basis constructor: rows of the carry gathered through a permutation (new row i = old row perm[i-1]+1)
__global__ void __launch_bounds__(256)
k_permute_rows(u64* __restrict__ dst, const u64* __restrict__ src, size_t ps, int ld, int m, int L,
               const int* __restrict__ perm) {
    const int i = blockIdx.y;                       // carry row 0..m
    const int from = i == 0 ? 0 : perm[i - 1] + 1;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ld; k += gridDim.x * blockDim.x)
        for (int l = 0; l < L; ++l) dst[(size_t)l * ps + (size_t)i * ld + k] = src[(size_t)l * ps + (size_t)from * ld + k];
}

// per-column steepest-edge recurrence, fixed width WX = LG + 4 limbs (register resident, 32-bit limb
// IMAD chains); used whenever D^2 has at most 256 trailing zero bits (E2 <= 4).
template <int L>
__global__ void __launch_bounds__(128)
k_gamma_update_t(int n, ColOwn own, const unsigned char* __restrict__ inbasis, const u64* __restrict__ nu,
                 const u64* __restrict__ sigma, u64* __restrict__ G, const Scalars* sc) {
    constexpr int LU = L + 2, LS = 2 * L + 7, LG = 2 * L + 6, WX = LG + 4, N = 2 * WX;
    if (sc->status != ST_RUN) return;
    const int tix = blockIdx.x * blockDim.x + threadIdx.x;
    if (tix >= own.count()) return;
    const int j = own.at(tix);
    if (inbasis[j] || j == sc->leaving) return;   // entering: None; leaving: set by k_finalize
    u32 nv[N], x[N];
    {
        u64 top = nu[(size_t)(LU - 1) * n + j];
        u64 sgx = (i64)top < 0 ? ~0ull : 0ull;
        u64 any = 0;
#pragma unroll
        for (int l = 0; l < WX; ++l) {
            u64 v = l < LU ? nu[(size_t)l * n + j] : sgx;
            if (l < LU) any |= v;
            nv[2 * l] = (u32)v; nv[2 * l + 1] = (u32)(v >> 32);
        }
        u32 g[N];
#pragma unroll
        for (int l = 0; l < WX; ++l) {
            u64 v = l < LG ? G[(size_t)l * n + j] : 0;
            g[2 * l] = (u32)v; g[2 * l + 1] = (u32)(v >> 32);
        }
        // S1 is staged in registers first: with two or three warps per SM nothing hides the load-use
        // latency of an operand fetched inside the carry chains
        u32 s1[N];
#pragma unroll
        for (int k = 0; k < N; ++k) s1[k] = reinterpret_cast<const u32*>(sc->S1)[k];
        mp_mul_lo<N>(x, g, s1);                                               // a^2/D^2 Ghat
        if (any != 0) {
            u32 sg[N], y[N], z[N];
            u64 tops = sigma[(size_t)(LS - 1) * n + j];
            u64 sgs = (i64)tops < 0 ? ~0ull : 0ull;
#pragma unroll
            for (int l = 0; l < WX; ++l) {
                u64 v = l < LS ? sigma[(size_t)l * n + j] : sgs;
                sg[2 * l] = (u32)v; sg[2 * l + 1] = (u32)(v >> 32);
            }
            mp_mul_lo<N>(y, nv, sg);                                          // nu sigma
            mp_mul_lo<N>(z, y, reinterpret_cast<const u32*>(sc->S2));       // 2a/D^2 nu sigma
            {
                u32 bf = 0;
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    u32 v = x[k] - z[k]; u32 b1 = x[k] < z[k]; u32 v2 = v - bf; u32 b2 = v < bf;
                    x[k] = v2; bf = b1 + b2;
                }
            }
            mp_mul_lo<N>(y, nv, nv);                                          // nu^2
            mp_mul_lo<N>(z, y, reinterpret_cast<const u32*>(sc->S3));       // Gq/D^2 nu^2
            {
                u32 cf = 0;
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    u32 v = x[k] + z[k]; u32 c1 = v < z[k]; u32 v2 = v + cf; u32 c2 = v2 < v;
                    x[k] = v2; cf = c1 + c2;
                }
            }
        }
    }
    // shift right by t2 (run-time word offset: go through a small local array)
    const int t2 = sc->t2;
    const int tw = t2 >> 5, tb = t2 & 31;
    u32 xs[N + 1];
#pragma unroll
    for (int k = 0; k < N; ++k) xs[k] = x[k];
    xs[N] = 0;
#pragma unroll
    for (int l = 0; l < LG; ++l) {
        u32 a0 = xs[min(2 * l + tw, N)], a1 = xs[min(2 * l + tw + 1, N)], a2 = xs[min(2 * l + tw + 2, N)];
        u32 lo = __funnelshift_r(a0, a1, tb), hi = __funnelshift_r(a1, a2, tb);
        G[(size_t)l * n + j] = (u64)lo | ((u64)hi << 32);
    }
}

// per-column steepest-edge recurrence (run-time widths: O(n) work, not the hot spot)
__global__ void __launch_bounds__(128)
k_gamma_update(int n, ColOwn own, int L, const unsigned char* __restrict__ inbasis, const u64* __restrict__ nu,
               const u64* __restrict__ sigma, u64* __restrict__ G, const Scalars* sc) {
    if (sc->status != ST_RUN) return;
    const int tix = blockIdx.x * blockDim.x + threadIdx.x;
    if (tix >= own.count()) return;
    const int j = own.at(tix);
    if (inbasis[j] || j == sc->leaving) return;   // entering: None; leaving: set by k_finalize
    const int LU = L + 2, LS = 2 * L + 7, LG = 2 * L + 6;
    const int WX = LG + sc->E2;
    u64 buf[7 * RG_MAXW];
    u64* raw = buf; u64* nv = buf + RG_MAXW; u64* sg = buf + 2 * RG_MAXW; u64* g = buf + 3 * RG_MAXW;
    u64* x = buf + 4 * RG_MAXW; u64* y = buf + 5 * RG_MAXW; u64* z = buf + 6 * RG_MAXW;
    rt_load_planar(raw, LU, nu, n, j);
    if (rt_is_zero(raw, LU)) {
        // alpha_j_bar == 0: gamma unchanged, so Ghat' = Ghat * a^2 / D^2
        rt_load_planar(raw, LG, G, n, j);
        for (int l = 0; l < WX; ++l) g[l] = l < LG ? raw[l] : 0;
        rt_mul_lo(x, sc->S1, g, WX);
        rt_shr(x, WX, sc->t2);
        rt_store_planar(G, n, j, x, LG);
        return;
    }
    rt_sext(nv, WX, raw, LU);
    rt_load_planar(raw, LS, sigma, n, j);
    rt_sext(sg, WX, raw, LS);
    rt_load_planar(raw, LG, G, n, j);
    for (int l = 0; l < WX; ++l) g[l] = l < LG ? raw[l] : 0;
    rt_mul_lo(x, sc->S1, g, WX);          // a^2/D^2 Ghat
    rt_mul_lo(y, nv, sg, WX);             // nu sigma
    rt_mul_lo(z, sc->S2, y, WX);          // 2a/D^2 nu sigma
    rt_sub(x, z, WX);
    rt_mul_lo(y, nv, nv, WX);             // nu^2
    rt_mul_lo(z, sc->S3, y, WX);          // Gq/D^2 nu^2
    rt_add(x, z, WX);
    rt_shr(x, WX, sc->t2);
    rt_store_planar(G, n, j, x, LG);
}

// The same recurrence with the five products of a column spread over three warps of the block (role = warp):
//   warp 0: S1 Ghat          warp 1: (nu sigma) S2          warp 2: (nu nu) S3
// for the block's 32 columns (thread = column within each warp), exchanged through shared memory and folded by
// warp 0.  One thread per column makes the launch as long as five dependent 40-limb products of a single thread
// whatever the number of columns (it did not shrink under column sharding); this form is two products long.
__global__ void __launch_bounds__(96)
k_gamma_update3(int n, ColOwn own, int L, const unsigned char* __restrict__ inbasis, const u64* __restrict__ nu,
                const u64* __restrict__ sigma, u64* __restrict__ G, const Scalars* sc) {
    extern __shared__ u64 sX[];            // [2][WX][32]: results of warps 1 and 2
    if (sc->status != ST_RUN) return;
    const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tix = blockIdx.x * 32 + lane;
    const int LU = L + 2, LS = 2 * L + 7, LG = 2 * L + 6;
    const int WX = LG + sc->E2;
    bool act = tix < own.count();
    const int j = act ? own.at(tix) : 0;
    if (act && (inbasis[j] || j == sc->leaving)) act = false;   // entering: None; leaving: set by k_finalize
    u64 buf[4 * RG_MAXW];                  // ONE local buffer, sliced by hand (see bigint.cuh)
    u64* raw = buf; u64* v = buf + RG_MAXW; u64* y = buf + 2 * RG_MAXW; u64* z = buf + 3 * RG_MAXW;
    bool nz = false;
    if (act) {
        rt_load_planar(raw, LU, nu, n, j);
        nz = !rt_is_zero(raw, LU);         // alpha_j_bar == 0: gamma unchanged, Ghat' = Ghat a^2 / D^2
        if (role == 0) {
            rt_load_planar(raw, LG, G, n, j);
            for (int l = 0; l < WX; ++l) v[l] = l < LG ? raw[l] : 0;
            rt_mul_lo(z, sc->S1, v, WX);                  // a^2/D^2 Ghat
        } else if (nz) {
            rt_sext(v, WX, raw, LU);
            if (role == 1) {
                rt_load_planar(raw, LS, sigma, n, j);
                rt_sext(y, WX, raw, LS);
                rt_mul_lo(z, v, y, WX);                   // nu sigma
                rt_mul_lo(y, sc->S2, z, WX);              // 2a/D^2 nu sigma
            } else {
                rt_mul_lo(z, v, v, WX);                   // nu^2
                rt_mul_lo(y, sc->S3, z, WX);              // Gq/D^2 nu^2
            }
            u64* dst = sX + (size_t)(role - 1) * WX * 32 + lane;
            for (int l = 0; l < WX; ++l) dst[(size_t)l * 32] = y[l];
        }
    }
    __syncthreads();
    if (act && role == 0) {
        if (nz) {
            for (int l = 0; l < WX; ++l) y[l] = sX[(size_t)l * 32 + lane];
            rt_sub(z, y, WX);
            for (int l = 0; l < WX; ++l) y[l] = sX[((size_t)WX + l) * 32 + lane];
            rt_add(z, y, WX);
        }
        rt_shr(z, WX, sc->t2);
        rt_store_planar(G, n, j, z, LG);
    }
}

// ---------------------------------------------------------------------------------------------
// The weight recurrence with a WARP per column (run-time widths, 32-bit limbs).
// Every low product P = A B mod 2^(32 W) is spread over the lanes by OUTPUT limb: lane l owns the CH consecutive
// limbs k = CH l .. CH l + CH - 1 and scans their columns, (c2:c1:c0)_k += A_i B_{k-i} for i = 0 .. k, with A_i
// broadcast from shared memory and a CH-word window of B sliding through registers (one new shared-memory word
// per i for CH multiply-adds).  The three products of the final sum share one set of column accumulators; the
// carries are resolved in registers (neighbour words by shuffle, then a ripple that almost always ends after one
// round).  Latency per column: ~3 W multiply-add steps instead of 5 W^2 / 2 in one thread.
// ---------------------------------------------------------------------------------------------
template <int CH>
__device__ __forceinline__ void wm_cols(u32 (&c0)[CH], u32 (&c1)[CH], u32 (&c2)[CH], const u32* __restrict__ A,
                                        const u32* __restrict__ B, int W32, int lane) {
    const int base = CH * lane;
    if (base >= W32) return;
    u32 w[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) w[c] = base + c < W32 ? B[base + c] : 0u;
    const int iend = min(W32, base + CH);
    for (int i = 0; i < iend; ++i) {
        const u32 a = A[i];
#pragma unroll
        for (int c = 0; c < CH; ++c)
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                         : "+r"(c0[c]), "+r"(c1[c]), "+r"(c2[c]) : "r"(a), "r"(w[c]));
#pragma unroll
        for (int c = CH - 1; c >= 1; --c) w[c] = w[c - 1];
        const int idx = base - i - 1;
        w[0] = idx >= 0 ? B[idx] : 0u;
    }
}
// limbs of the column sums: r_k = c0_k + c1_{k-1} + c2_{k-2} + carry; every lane of the warp must call this
template <int CH>
__device__ __forceinline__ void wm_resolve(u32 (&r)[CH], const u32 (&c0)[CH], const u32 (&c1)[CH], const u32 (&c2)[CH],
                                           int lane) {
    static_assert(CH >= 2, "a chunk takes words from the previous lane only");
    u32 p_c1 = __shfl_up_sync(0xffffffffu, c1[CH - 1], 1);
    u32 p_c2a = __shfl_up_sync(0xffffffffu, c2[CH - 1], 1);     // joins limb base + 1
    u32 p_c2b = __shfl_up_sync(0xffffffffu, c2[CH - 2], 1);     // joins limb base
    if (lane == 0) { p_c1 = 0; p_c2a = 0; p_c2b = 0; }
    u64 carry = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        u64 t = (u64)c0[c] + carry;
        t += c >= 1 ? c1[c - 1] : p_c1;
        t += c >= 2 ? c2[c - 2] : (c == 1 ? p_c2a : p_c2b);
        r[c] = (u32)t; carry = t >> 32;
    }
    u32 cout = (u32)carry;
    for (;;) {
        u32 cin = __shfl_up_sync(0xffffffffu, cout, 1);
        if (lane == 0) cin = 0;
        if (!__any_sync(0xffffffffu, cin != 0)) break;
        u64 cc = cin;
#pragma unroll
        for (int c = 0; c < CH; ++c) { u64 t = (u64)r[c] + cc; r[c] = (u32)t; cc = t >> 32; }
        cout = (u32)cc;
    }
}
template <int CH>
__global__ void __launch_bounds__(128)
k_gamma_update_w(int n, ColOwn own, int L, const unsigned char* __restrict__ inbasis, const u64* __restrict__ nu,
                 const u64* __restrict__ sigma, u64* __restrict__ G, Scalars* sc) {
    extern __shared__ u32 smw[];
    if (sc->status != ST_RUN) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int LU = L + 2, LS = 2 * L + 7, LG = 2 * L + 6;
    const int W32 = 2 * (LG + sc->E2);
    if (W32 > 32 * CH) {           // the host sized the lanes' chunks for a narrower division width
        if (threadIdx.x == 0) { sc->status = ST_FATAL; sc->fatal = 2; }
        return;
    }
    const int WP = W32 + 2;        // two zero words behind every number (the final shift reads past the end)
    u32* S1 = smw; u32* S2n = smw + WP; u32* S3 = smw + 2 * WP;
    u32* mine = smw + 3 * WP + warp * 6 * WP;
    u32* NV = mine; u32* SG = mine + WP; u32* GG = mine + 2 * WP; u32* Y1 = mine + 3 * WP; u32* Y2 = mine + 4 * WP;
    u32* X = mine + 5 * WP;
    // block-wide: the three uniform scalars, S2 negated (x = S1 Ghat + (-S2) (nu sigma) + S3 (nu nu))
    for (int k = threadIdx.x; k < W32; k += blockDim.x) {
        S1[k] = reinterpret_cast<const u32*>(sc->S1)[k];
        S3[k] = reinterpret_cast<const u32*>(sc->S3)[k];
    }
    if (threadIdx.x == 0) {
        u32 c = 1;
        for (int k = 0; k < W32; ++k) {
            u32 v = ~reinterpret_cast<const u32*>(sc->S2)[k] + c;
            c = (c && v == 0) ? 1u : 0u;
            S2n[k] = v;
        }
    }
    for (int k = lane; k < 2; k += 32) { X[W32 + k] = 0; }
    __syncthreads();
    const int tix = blockIdx.x * 4 + warp;
    if (tix >= own.count()) return;
    const int j = own.at(tix);
    if (inbasis[j] || j == sc->leaving) return;   // entering: None; leaving: set by k_finalize
    // operands: nu (sign-extended), sigma (sign-extended), Ghat (zero-extended)
    const u64 nu_top = nu[(size_t)(LU - 1) * n + j], sg_top = sigma[(size_t)(LS - 1) * n + j];
    const u32 nu_sign = (i64)nu_top < 0 ? ~0u : 0u, sg_sign = (i64)sg_top < 0 ? ~0u : 0u;
    u32 any = 0;
    for (int k = lane; k < W32; k += 32) {
        const int l = k >> 1, hi = k & 1;
        u32 a = nu_sign, b = sg_sign, g = 0;
        if (l < LU) { u64 v = nu[(size_t)l * n + j]; a = hi ? (u32)(v >> 32) : (u32)v; any |= a; }
        if (l < LS) { u64 v = sigma[(size_t)l * n + j]; b = hi ? (u32)(v >> 32) : (u32)v; }
        if (l < LG) { u64 v = G[(size_t)l * n + j]; g = hi ? (u32)(v >> 32) : (u32)v; }
        NV[k] = a; SG[k] = b; GG[k] = g;
    }
    const bool nz = __any_sync(0xffffffffu, any != 0);     // alpha_j_bar == 0: Ghat' = Ghat a^2 / D^2
    __syncwarp();
    u32 c0[CH], c1[CH], c2[CH], r[CH];
    const int base = CH * lane;
    if (nz) {
#pragma unroll
        for (int c = 0; c < CH; ++c) { c0[c] = 0; c1[c] = 0; c2[c] = 0; }
        wm_cols<CH>(c0, c1, c2, NV, SG, W32, lane);             // nu sigma
        wm_resolve<CH>(r, c0, c1, c2, lane);
#pragma unroll
        for (int c = 0; c < CH; ++c) if (base + c < W32) Y1[base + c] = r[c];
#pragma unroll
        for (int c = 0; c < CH; ++c) { c0[c] = 0; c1[c] = 0; c2[c] = 0; }
        wm_cols<CH>(c0, c1, c2, NV, NV, W32, lane);             // nu^2
        wm_resolve<CH>(r, c0, c1, c2, lane);
#pragma unroll
        for (int c = 0; c < CH; ++c) if (base + c < W32) Y2[base + c] = r[c];
        __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) { c0[c] = 0; c1[c] = 0; c2[c] = 0; }
    wm_cols<CH>(c0, c1, c2, S1, GG, W32, lane);                 // a^2/D^2 Ghat
    if (nz) {
        wm_cols<CH>(c0, c1, c2, S2n, Y1, W32, lane);            // - 2a/D^2 nu sigma
        wm_cols<CH>(c0, c1, c2, S3, Y2, W32, lane);             // + Gq/D^2 nu^2
    }
    wm_resolve<CH>(r, c0, c1, c2, lane);
#pragma unroll
    for (int c = 0; c < CH; ++c) if (base + c < W32) X[base + c] = r[c];
    __syncwarp();
    // shift right by t2 and store the low LG limbs
    const int t2 = sc->t2;
    const int tw = t2 >> 5, tb = t2 & 31;
    for (int l = lane; l < LG; l += 32) {
        const int wi = 2 * l + tw;
        const u32 a0 = wi < WP ? X[wi] : 0u, a1 = wi + 1 < WP ? X[wi + 1] : 0u, a2 = wi + 2 < WP ? X[wi + 2] : 0u;
        const u32 lo = __funnelshift_r(a0, a1, tb), hi = __funnelshift_r(a1, a2, tb);
        G[(size_t)l * n + j] = (u64)lo | ((u64)hi << 32);
    }
}

// b_p != 0 ?  (remove_artificial_basis_variables, phase_one.rs:250) -- read from the staged row
__global__ void k_bp_nonzero(const u64* rowp, size_t rs, int L, Scalars* sc) {
    if (threadIdx.x || blockIdx.x) return;
    u64 o = 0;
    for (int l = 0; l < L; ++l) o |= rowp[(size_t)l * rs];
    sc->bp_nonzero = o != 0;
}

}  // namespace rg
