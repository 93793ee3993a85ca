// K1, the rank-1 integer-preserving pivot of the carry, in its own header: its (L, E, CP, RT) variants are the
// bulk of the compile time and are instantiated in one translation unit per limb width (k1_variants.cu).
#pragma once
#include "engine.cuh"
#include "mp32.cuh"

namespace rg {

__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// K1: rank-1 integer-preserving pivot of the whole carry, in place:
//        C'[i][k] = ( |a| C[i][k] - sgn(a) u_i C[p][k] ) / D          (u_p replaced by a - D)
//               = ( A C[i][k] + Bn_i rowp[k]  mod 2^(64 (L+E)) ) >> t
// with A = |a| inv(odd D), Bn_i = -sgn(a) u_i inv(odd D): the exact division is fused into two low
// products.  (Carry::change_basis + update_b + update_minus_pi_and_obj, carry/mod.rs:561-604,295-349;
// BasisInverseRows::{normalize_pivot_row,row_reduce}, basis_inverse_rows.rs:43-84.)
// Thread = CP adjacent columns (CP = 2: 128-bit loads/stores); block = blockDim*CP columns x RT rows
// (dense mode: 256 threads x 32 rows; active-column mode: 128 threads x 8 rows over the packed block, the
// list is short and the serial row loop is what bounds the launch).
// ---------------------------------------------------------------------------------------------
template <int L, int E, int CP, int RT = 32>
__global__ void __launch_bounds__(256)
k_update(u64* __restrict__ C, size_t ps, int ld, int row_first, int nrows, const int* __restrict__ klist,
         const u64* __restrict__ u, size_t us, const u64* __restrict__ rowp, size_t rs, Scalars* sc) {
    constexpr int W = L + E;
    constexpr int N = 2 * W;          // 32-bit limbs of the working width
    constexpr int LU = L + 2;
    __shared__ u32 sBn[RT][N];
    __shared__ u32 sA[N];
    __shared__ unsigned char sBnz[RT];     // row factor Bn_i != 0 (u_i != 0)
    if (sc->status != ST_RUN) return;
    if (sc->E != E) return;
    const int tid = threadIdx.x;
    const int row0 = row_first + blockIdx.y * RT;
    if (tid < N) sA[tid] = reinterpret_cast<const u32*>(sc->A)[tid];
    if (tid < RT) {
        int i = row0 + tid;
        if (i < nrows) {
            u32 ui[N], bn[N];
            if (i == sc->p) {
#pragma unroll
                for (int l = 0; l < W; ++l) { u64 v = sc->up[l]; ui[2 * l] = (u32)v; ui[2 * l + 1] = (u32)(v >> 32); }
            } else {
                u64 top = u[(size_t)(LU - 1) * us + i];
                u64 sg = (i64)top < 0 ? ~0ull : 0ull;
#pragma unroll
                for (int l = 0; l < W; ++l) {
                    u64 v = l < LU ? u[(size_t)l * us + i] : sg;
                    ui[2 * l] = (u32)v; ui[2 * l + 1] = (u32)(v >> 32);
                }
            }
            mp_mul_lo<N>(bn, ui, reinterpret_cast<const u32*>(sc->Dinv));
            if (sc->sgn > 0) {   // Bn = -u Dinv
                u32 c = 1;
#pragma unroll
                for (int k = 0; k < N; ++k) { u32 v = ~bn[k] + c; c = (c && v == 0) ? 1u : 0u; bn[k] = v; }
            }
            u32 any = 0;
#pragma unroll
            for (int k = 0; k < N; ++k) { sBn[tid][k] = bn[k]; any |= bn[k]; }
            sBnz[tid] = any != 0;
        }
    }
    __syncthreads();
    // dense mode: CP adjacent columns per thread.  List mode (klist != nullptr): C is the PACKED active block
    // (`ld` = its column capacity) and the thread owns list positions idx .. idx+CP-1, whose pivot-row
    // entries are gathered once from the densely staged row through klist -- trivial columns are never
    // touched and every access of the row loop below is a coalesced 64/128-bit one
    const int idx = (blockIdx.x * blockDim.x + tid) * CP;
    const int ncol = klist ? sc->nk : ld;
    int maxb = 0;
    if (idx < ncol) {
        const int col = idx;
        const int t = sc->t;
        const int tw = t >> 5, tb = t & 31;
        u32 rp[CP][N];
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            const int rc = klist ? (idx + c < ncol ? klist[idx + c] : -1) : idx + c;
#pragma unroll
            for (int l = 0; l < L; ++l) {
                u64 v = rc >= 0 ? rowp[(size_t)l * rs + rc] : 0ull;
                rp[c][2 * l] = (u32)v; rp[c][2 * l + 1] = (u32)(v >> 32);
            }
            u32 sg = (int)rp[c][2 * L - 1] < 0 ? ~0u : 0u;
#pragma unroll
            for (int k = 2 * L; k < N; ++k) rp[c][k] = sg;
        }
        u32 rpnz = 0;
#pragma unroll
        for (int c = 0; c < CP; ++c)
#pragma unroll
            for (int k = 0; k < 2 * L; ++k) rpnz |= rp[c][k];
        const int rend = min(RT, nrows - row0);
        // software prefetch (L <= 8): the next row's entry is in flight while this one is processed -- most
        // entries are skipped as zeros, so the loop is a dependent load chain without it
        constexpr bool PF = L <= 8;
        u64 nx[PF ? CP : 1][PF ? L : 1];
        if (PF && rend > 0) {
            const size_t off0 = (size_t)row0 * ld + col;
#pragma unroll
            for (int l = 0; l < (PF ? L : 0); ++l) {
                if (CP == 2) {
                    ulonglong2 v = *reinterpret_cast<const ulonglong2*>(C + (size_t)l * ps + off0);
                    nx[0][l] = v.x; nx[(PF ? CP : 1) - 1][l] = v.y;
                } else nx[0][l] = C[(size_t)l * ps + off0];
            }
        }
        for (int r = 0; r < rend; ++r) {
            size_t off = (size_t)(row0 + r) * ld + col;
            u32 cv[CP][N];
            if (PF) {
#pragma unroll
                for (int c = 0; c < CP; ++c)
#pragma unroll
                    for (int l = 0; l < L; ++l) {
                        u64 v = nx[PF ? c : 0][PF ? l : 0];
                        cv[c][2 * l] = (u32)v; cv[c][2 * l + 1] = (u32)(v >> 32);
                    }
                if (r + 1 < rend) {
                    const size_t offn = off + ld;
#pragma unroll
                    for (int l = 0; l < (PF ? L : 0); ++l) {
                        if (CP == 2) {
                            ulonglong2 v = *reinterpret_cast<const ulonglong2*>(C + (size_t)l * ps + offn);
                            nx[0][l] = v.x; nx[(PF ? CP : 1) - 1][l] = v.y;
                        } else nx[0][l] = C[(size_t)l * ps + offn];
                    }
                }
            } else if (CP == 2) {
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    ulonglong2 v = *reinterpret_cast<const ulonglong2*>(C + (size_t)l * ps + off);
                    cv[0][2 * l] = (u32)v.x; cv[0][2 * l + 1] = (u32)(v.x >> 32);
                    cv[CP - 1][2 * l] = (u32)v.y; cv[CP - 1][2 * l + 1] = (u32)(v.y >> 32);
                }
            } else {
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    u64 v = C[(size_t)l * ps + off];
                    cv[0][2 * l] = (u32)v; cv[0][2 * l + 1] = (u32)(v >> 32);
                }
            }
            // zero skip: C[i][k] == 0 and (C[p][k] == 0 or u_i == 0)  =>  C'[i][k] == 0: nothing to compute
            // or store; rows with u_i == 0 (uniform over the block) need one product instead of two
            const bool two = sBnz[r] != 0;
            u32 nzc = two ? rpnz : 0u;
#pragma unroll
            for (int c = 0; c < CP; ++c)
#pragma unroll
                for (int k = 0; k < 2 * L; ++k) nzc |= cv[c][k];
            if (nzc == 0) continue;
            u64 res[CP][L];
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                u32 sg = (int)cv[c][2 * L - 1] < 0 ? ~0u : 0u;
#pragma unroll
                for (int k = 2 * L; k < N; ++k) cv[c][k] = sg;
                u32 X[N];
                if (two) mp_mul2_lo<N>(X, cv[c], sA, rp[c], sBn[r]);
                else mp_mul_lo<N>(X, cv[c], sA);
                u32 o[2 * L];
                if (E == 0) {
#pragma unroll
                    for (int k = 0; k < 2 * L; ++k) o[k] = X[k];
                } else {
#pragma unroll
                    for (int w = 0; w <= 2 * E; ++w) {
                        if (tw == w) {
#pragma unroll
                            for (int k = 0; k < 2 * L; ++k) {
                                u32 lo = X[k + w < N ? k + w : N - 1];
                                u32 hi = (k + w + 1 < N) ? X[k + w + 1 < N ? k + w + 1 : N - 1] : 0u;
                                o[k] = __funnelshift_r(lo, hi, tb);
                            }
                        }
                    }
                }
                // bit length of |result|
                u32 sgn = (int)o[2 * L - 1] < 0 ? ~0u : 0u;
                int bl = 0;
#pragma unroll
                for (int k = 0; k < 2 * L; ++k) {
                    u32 v = o[k] ^ sgn;
                    if (v) bl = 32 * k + 32 - __clz(v);
                }
                maxb = max(maxb, bl + (sgn ? 1 : 0));
#pragma unroll
                for (int l = 0; l < L; ++l) res[c][l] = (u64)o[2 * l] | ((u64)o[2 * l + 1] << 32);
            }
            if (CP == 2) {
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    ulonglong2 v; v.x = res[0][l]; v.y = res[CP - 1][l];
                    *reinterpret_cast<ulonglong2*>(C + (size_t)l * ps + off) = v;
                }
            } else {
#pragma unroll
                for (int l = 0; l < L; ++l) C[(size_t)l * ps + off] = res[0][l];
            }
        }
    }
    maxb = warp_max(maxb);
    if ((tid & 31) == 0 && maxb) atomicMax(&sc->maxbits_new, maxb);
}


// ---------------------------------------------------------------------------------------------
// K1 on the packed active block, warp-granular (list mode, L >= 8).
//
// k_bn_rows: the row factors Bn_i = -sgn(a) u_i inv(odd D) mod 2^(32 N) of all local rows, once per pivot (one
// thread per row), as rows of NP = N + 1 words -- word N flags Bn_i != 0.  k_update computes them per block in a
// prologue that leaves most of the block idle; here they are a table every work item copies its rows from.
//
// k_update_items: one WARP per work item.  An item is (tile of RT rows) x (32 consecutive list positions); the
// nk mod 32 remainder positions are not given 32-lane items of their own (one lane in 32 busy when nk = 161):
// a remainder item maps its lanes to (row, position) pairs -- rc positions x floor(32 / rc) rows per pass -- and
// covers correspondingly more rows.  No block-level synchronisation, no idle warps inside a block, and the
// 64-thread blocks let the register-limited occupancy move in steps of two warps instead of four.
// ---------------------------------------------------------------------------------------------
template <int L, int E>
__global__ void __launch_bounds__(128)
k_bn_rows(const u64* __restrict__ u, size_t us, int nloc, u32* __restrict__ bn, Scalars* sc) {
    constexpr int W = L + E, N = 2 * W, LU = L + 2, NP = N + 1;
    if (sc->status != ST_RUN) return;
    if (sc->E != E) return;
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;     // local carry row
    if (i > nloc) return;
    u32 ui[N], b[N];
    if (i == sc->p) {
#pragma unroll
        for (int l = 0; l < W; ++l) { u64 v = sc->up[l]; ui[2 * l] = (u32)v; ui[2 * l + 1] = (u32)(v >> 32); }
    } else {
        u64 top = u[(size_t)(LU - 1) * us + i];
        u64 sg = (i64)top < 0 ? ~0ull : 0ull;
#pragma unroll
        for (int l = 0; l < W; ++l) {
            u64 v = l < LU ? u[(size_t)l * us + i] : sg;
            ui[2 * l] = (u32)v; ui[2 * l + 1] = (u32)(v >> 32);
        }
    }
    u32 anyu = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) anyu |= ui[k];
    u32* out = bn + (size_t)i * NP;
    if (anyu == 0) {            // u_i == 0 (most rows of a sparse problem): Bn_i = 0 without the product
#pragma unroll
        for (int k = 0; k < NP; ++k) out[k] = 0;
        return;
    }
    mp_mul_lo<N>(b, ui, reinterpret_cast<const u32*>(sc->Dinv));
    if (sc->sgn > 0) {   // Bn = -u Dinv
        u32 c = 1;
#pragma unroll
        for (int k = 0; k < N; ++k) { u32 v = ~b[k] + c; c = (c && v == 0) ? 1u : 0u; b[k] = v; }
    }
    u32 any = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) { out[k] = b[k]; any |= b[k]; }
    out[N] = any != 0;
}

template <int L, int E>
__global__ void __launch_bounds__(64)
k_update_items(u64* __restrict__ C, size_t ps, int ld, int nloc, int RT, int prefetch, const int* __restrict__ klist,
               const u32* __restrict__ bn, const u64* __restrict__ rowp, size_t rs, Scalars* sc) {
    constexpr int W = L + E, N = 2 * W, NP = N + 1;
    __shared__ u32 sB[2][32 * NP];      // per warp: the Bn rows of its item (at most 32)
    __shared__ u32 sA[N];
    if (sc->status != ST_RUN) return;
    if (sc->E != E) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < N) sA[tid] = reinterpret_cast<const u32*>(sc->A)[tid];
    __syncthreads();
    const int nk = sc->nk;
    const int full = nk >> 5, rc = nk & 31;
    const int tiles = (nloc + RT - 1) / RT;
    const int n_full_items = tiles * full;
    const int rpp = rc ? 32 / rc : 0;                    // rows per pass of a remainder item
    const int rows_rem = rpp ? (32 / rpp) * rpp : 1;     // rows of a remainder item (<= 32)
    const int tiles_rem = rc ? (nloc + rows_rem - 1) / rows_rem : 0;
    const int w = blockIdx.x * 2 + warp;
    int row_begin, nrows_item, rstep, rsub, pos;
    bool active;
    if (w < n_full_items) {
        const int tile = w / full, chunk = w - tile * full;
        row_begin = 1 + tile * RT; nrows_item = min(RT, nloc - tile * RT);
        rstep = 1; rsub = 0; pos = chunk * 32 + lane; active = true;
    } else {
        const int w2 = w - n_full_items;
        if (w2 >= tiles_rem) return;
        row_begin = 1 + w2 * rows_rem; nrows_item = min(rows_rem, nloc - w2 * rows_rem);
        rstep = rpp; rsub = lane / rc; pos = full * 32 + lane - rsub * rc; active = lane < rpp * rc;
    }
    // the item's rows of the factor table are contiguous: one coalesced copy
    u32* sBw = sB[warp];
    {
        const u32* src = bn + (size_t)row_begin * NP;
        const int words = nrows_item * NP;
        for (int j = lane; j < words; j += 32) sBw[j] = src[j];
    }
    __syncwarp();
    int maxb = 0;
    if (active) {
        const int t = sc->t;
        const int tw = t >> 5, tb = t & 31;
        u32 rp[N];
        {
            const int rcol = klist[pos];
#pragma unroll
            for (int l = 0; l < L; ++l) {
                u64 v = rowp[(size_t)l * rs + rcol];
                rp[2 * l] = (u32)v; rp[2 * l + 1] = (u32)(v >> 32);
            }
            u32 sg = (int)rp[2 * L - 1] < 0 ? ~0u : 0u;
#pragma unroll
            for (int k = 2 * L; k < N; ++k) rp[k] = sg;
        }
        u32 rpnz = 0;
#pragma unroll
        for (int k = 0; k < 2 * L; ++k) rpnz |= rp[k];
        for (int k = rsub; k < nrows_item; k += rstep) {
            const size_t off = (size_t)(row_begin + k) * ld + pos;
            u32 cv[N];
#pragma unroll
            for (int l = 0; l < L; ++l) {
                u64 v = C[(size_t)l * ps + off];
                cv[2 * l] = (u32)v; cv[2 * l + 1] = (u32)(v >> 32);
            }
            if (prefetch && k + rstep < nrows_item) {      // next row's entry on its way to L2 / L1 while this one is computed
                const size_t offn = off + (size_t)rstep * ld;
#pragma unroll
                for (int l = 0; l < L; ++l)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(__cvta_generic_to_global(C + (size_t)l * ps + offn)));
            }
            const u32* bptr = sBw + k * NP;
            const bool two = bptr[N] != 0;
            // zero skip: C[i][k] == 0 and (C[p][k] == 0 or u_i == 0)  =>  C'[i][k] == 0
            u32 nzc = two ? rpnz : 0u;
#pragma unroll
            for (int q = 0; q < 2 * L; ++q) nzc |= cv[q];
            if (nzc == 0) continue;
            {
                u32 sg = (int)cv[2 * L - 1] < 0 ? ~0u : 0u;
#pragma unroll
                for (int q = 2 * L; q < N; ++q) cv[q] = sg;
            }
            u32 X[N];
            if (two) mp_mul2_lo<N>(X, cv, sA, rp, bptr);
            else mp_mul_lo<N>(X, cv, sA);
            u32 o[2 * L];
            if (E == 0) {
#pragma unroll
                for (int q = 0; q < 2 * L; ++q) o[q] = X[q];
            } else {
#pragma unroll
                for (int ww = 0; ww <= 2 * E; ++ww) {
                    if (tw == ww) {
#pragma unroll
                        for (int q = 0; q < 2 * L; ++q) {
                            u32 lo = X[q + ww < N ? q + ww : N - 1];
                            u32 hi = (q + ww + 1 < N) ? X[q + ww + 1 < N ? q + ww + 1 : N - 1] : 0u;
                            o[q] = __funnelshift_r(lo, hi, tb);
                        }
                    }
                }
            }
            u32 sgn = (int)o[2 * L - 1] < 0 ? ~0u : 0u;
            int bl = 0;
#pragma unroll
            for (int q = 0; q < 2 * L; ++q) {
                u32 v = o[q] ^ sgn;
                if (v) bl = 32 * q + 32 - __clz(v);
            }
            maxb = max(maxb, bl + (sgn ? 1 : 0));
#pragma unroll
            for (int l = 0; l < L; ++l) C[(size_t)l * ps + off] = (u64)o[2 * l] | ((u64)o[2 * l + 1] << 32);
        }
    }
    maxb = warp_max(maxb);
    if (lane == 0 && maxb) atomicMax(&sc->maxbits_new, maxb);
}


// ---------------------------------------------------------------------------------------------
// Reduced costs by recurrence (steepest edge): kappa_j = D cbar_j of every priced column follows the cost row of
// the carry through the pivot,
//        kappa'_j = ( |a| kappa_j - sgn(a) kappa_q nu_j ) / D,        nu_j = rowp . a_j,
// i.e. K1 applied to the vector kappa with nu as its pivot row and kappa_q = u[0] as the row factor.  nu is
// computed for the weight update anyway, so the pricing dot over the constraint matrix (one more pass over the
// int8 block, its byte slices and their recombination) is not needed after a steepest-edge pivot.  Values are
// those of Tableau::relative_cost (tableau/mod.rs:106-112) -- exact arithmetic makes the two routes identical.
// Leaving column: nu is not computed for basic columns; kappa' = -sgn(a) kappa_q (its nu is D).
// Needs sc->A and sc->Dinv two limbs wider than K1 does (k_scalars provides them).
// ---------------------------------------------------------------------------------------------
template <int L, int E>
__global__ void __launch_bounds__(128)
k_kappa_update(int n, int d0, int d1, int s0, int s1, const unsigned char* __restrict__ inbasis,
               const u64* __restrict__ nu, u64* __restrict__ kappa, const u64* __restrict__ u, size_t us,
               Scalars* sc) {
    constexpr int LU = L + 2, W = LU + E, N = 2 * W;
    __shared__ u32 sA[N], sB[N];
    __shared__ u64 sLeave[LU];
    if (sc->status != ST_RUN) return;
    if (sc->E != E) return;
    const int tid = threadIdx.x;
    if (tid < N) sA[tid] = reinterpret_cast<const u32*>(sc->A)[tid];
    if (tid == 96) {       // Bn0 = -sgn(a) kappa_q inv(odd D) mod 2^(32 N), and -sgn(a) kappa_q itself
        u64 kq[LU];
#pragma unroll
        for (int l = 0; l < LU; ++l) kq[l] = u[(size_t)l * us];
        const u64 sgq = (i64)kq[LU - 1] < 0 ? ~0ull : 0ull;
        u32 ui[N], b[N];
#pragma unroll
        for (int l = 0; l < W; ++l) {
            u64 v = l < LU ? kq[l] : sgq;
            ui[2 * l] = (u32)v; ui[2 * l + 1] = (u32)(v >> 32);
        }
        mp_mul_lo<N>(b, ui, reinterpret_cast<const u32*>(sc->Dinv));
        if (sc->sgn > 0) {
            u32 c = 1;
#pragma unroll
            for (int k = 0; k < N; ++k) { u32 v = ~b[k] + c; c = (c && v == 0) ? 1u : 0u; b[k] = v; }
            u64 c2 = 1;
#pragma unroll
            for (int l = 0; l < LU; ++l) { u64 v = ~kq[l] + c2; c2 = (c2 && v == 0) ? 1 : 0; kq[l] = v; }
        }
#pragma unroll
        for (int k = 0; k < N; ++k) sB[k] = b[k];
#pragma unroll
        for (int l = 0; l < LU; ++l) sLeave[l] = kq[l];
    }
    __syncthreads();
    const int tix = blockIdx.x * blockDim.x + tid;
    if (tix >= (d1 - d0) + (s1 - s0)) return;
    const int j = tix < d1 - d0 ? d0 + tix : s0 + (tix - (d1 - d0));
    if (j == sc->leaving) {
#pragma unroll
        for (int l = 0; l < LU; ++l) kappa[(size_t)l * n + j] = sLeave[l];
        return;
    }
    if (inbasis[j]) return;        // basic columns keep kappa = 0 (the entering column's comes out 0 below)
    u32 kv[N], nv[N];
    {
        const u64 tk = kappa[(size_t)(LU - 1) * n + j], tn = nu[(size_t)(LU - 1) * n + j];
        const u64 sk = (i64)tk < 0 ? ~0ull : 0ull, sn = (i64)tn < 0 ? ~0ull : 0ull;
#pragma unroll
        for (int l = 0; l < W; ++l) {
            u64 a = l < LU ? kappa[(size_t)l * n + j] : sk;
            u64 b = l < LU ? nu[(size_t)l * n + j] : sn;
            kv[2 * l] = (u32)a; kv[2 * l + 1] = (u32)(a >> 32);
            nv[2 * l] = (u32)b; nv[2 * l + 1] = (u32)(b >> 32);
        }
    }
    u32 X[N];
    mp_mul2_lo<N>(X, kv, sA, nv, sB);
    const int t = sc->t;
    const int tw = t >> 5, tb = t & 31;
    u32 o[2 * LU];
    if (E == 0) {
#pragma unroll
        for (int q = 0; q < 2 * LU; ++q) o[q] = X[q];
    } else {
#pragma unroll
        for (int ww = 0; ww <= 2 * E; ++ww) {
            if (tw == ww) {
#pragma unroll
                for (int q = 0; q < 2 * LU; ++q) {
                    u32 lo = X[q + ww < N ? q + ww : N - 1];
                    u32 hi = (q + ww + 1 < N) ? X[q + ww + 1 < N ? q + ww + 1 : N - 1] : 0u;
                    o[q] = __funnelshift_r(lo, hi, tb);
                }
            }
        }
    }
#pragma unroll
    for (int l = 0; l < LU; ++l) kappa[(size_t)l * n + j] = (u64)o[2 * l] | ((u64)o[2 * l + 1] << 32);
}

}  // namespace rg
