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

}  // namespace rg
