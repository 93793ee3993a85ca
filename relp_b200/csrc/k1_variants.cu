// One translation unit per limb width (compiled with -DRG_K1_L=1|2|4|8|10|12|14|16): the (E, CP, RT) variants of K1.
// They are the bulk of the library's compile time, so build.py compiles these units in parallel.
// rg::k1_launch_<L>(ctx, E) enqueues K1 for the current carry mode on ctx->stream and returns false when no
// fixed-width variant covers E (the caller falls back to the run-time-width kernel).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "k1_update.cuh"

#ifndef RG_K1_L
#error "compile with -DRG_K1_L=<limb width>"
#endif

using namespace rg;

namespace {

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

template <int L, int E>
void launch_le(rg_context* ctx) {
    constexpr int CP = L <= 4 ? 2 : 1;
    // L = 16 is bound by the multiply pipe in either mode: one instantiation (8-row blocks) serves both
    constexpr int RTD = L > 8 ? 8 : 32;
    if (ctx->list_mode) {
        // cost row: dense over all columns; rows 1..nloc: the non-trivial columns only
        dim3 g0(cdiv(ctx->ld, 256 * CP), 1);
        k_update<L, E, CP, RTD><<<g0, 256, 0, ctx->stream>>>(ctx->carry, ctx->plane, ctx->ld, 0, 1, (const int*)nullptr,
                                                             ctx->u, (size_t)ctx->ld, ctx->rowp, (size_t)ctx->ld,
                                                             ctx->sc);
        ctx->launches++;
        if (ctx->nloc > 0 && L >= ctx->k1_items_min_limbs) {
            // rows 1..nloc of the packed active block, warp-granular work items (k_update_items)
            k_bn_rows<L, E><<<cdiv(ctx->nloc, 128), 128, 0, ctx->stream>>>(ctx->u, (size_t)ctx->ld, ctx->nloc, ctx->bn,
                                                                          ctx->sc);
            const int RT = ctx->k1_items_rows;
            long long items = 0;
            if (ctx->capturing) {
                // a captured launch is replayed for every list length of its graph key: bound by the key's nk_grid
                // (a remainder item covers at least 17 rows; RT <= 16)
                items = (long long)cdiv(ctx->nloc, RT < 16 ? RT : 16) * (ctx->nk_grid >> 5) + cdiv(ctx->nloc, 16);
            } else {
                for (int nk = ctx->nk_host; nk <= ctx->nk_host + 1; ++nk) {     // this pivot may list one more column
                    const int full = nk >> 5, rc = nk & 31;
                    const int rpp = rc ? 32 / rc : 0, rows_rem = rpp ? (32 / rpp) * rpp : 1;
                    const long long it = (long long)cdiv(ctx->nloc, RT) * full + (rc ? cdiv(ctx->nloc, rows_rem) : 0);
                    items = it > items ? it : items;
                }
            }
            k_update_items<L, E><<<cdiv(items, 2), 64, 0, ctx->stream>>>(ctx->pk, ctx->pplane, ctx->cap, ctx->nloc, RT,
                                                                         ctx->k1_items_prefetch, (const int*)ctx->klist,
                                                                         ctx->bn, ctx->rowp, (size_t)ctx->ld, ctx->sc);
            ctx->launches += 2;
        } else if (ctx->nloc > 0) {
            // rows 1..nloc: the packed active block (column slot = list position, stride = capacity)
            dim3 g1(cdiv(ctx->nk_grid, 128 * CP), cdiv(ctx->nloc, 8));
            k_update<L, E, CP, 8><<<g1, 128, 0, ctx->stream>>>(ctx->pk, ctx->pplane, ctx->cap, 1, ctx->nloc + 1,
                                                               (const int*)ctx->klist, ctx->u, (size_t)ctx->ld,
                                                               ctx->rowp, (size_t)ctx->ld, ctx->sc);
            ctx->launches++;
        }
        return;
    }
    dim3 grid(cdiv(ctx->ld, 256 * CP), cdiv(ctx->nloc + 1, RTD));
    k_update<L, E, CP, RTD><<<grid, 256, 0, ctx->stream>>>(ctx->carry, ctx->plane, ctx->ld, 0, ctx->nloc + 1,
                                                           (const int*)nullptr, ctx->u, (size_t)ctx->ld, ctx->rowp,
                                                           (size_t)ctx->ld, ctx->sc);
    ctx->launches++;
}

template <int L, int E>
void launch_kappa_le(rg_context* ctx) {
    const int cnt = (ctx->d1 - ctx->d0) + (ctx->s1 - ctx->s0);
    if (cnt <= 0) return;
    k_kappa_update<L, E><<<cdiv(cnt, 128), 128, 0, ctx->stream>>>(ctx->n, ctx->d0, ctx->d1, ctx->s0, ctx->s1, ctx->inbasis,
                                                                  ctx->nu, ctx->kappa, ctx->u, (size_t)ctx->ld, ctx->sc);
    ctx->launches++;
}
template <int L>
bool launch_kappa_t(rg_context* ctx, int E) {
    switch (E) {
        case 0: launch_kappa_le<L, 0>(ctx); return true;
        case 1: launch_kappa_le<L, 1>(ctx); return true;
        case 2: if constexpr (L >= 2) { launch_kappa_le<L, 2>(ctx); return true; } break;
        case 3: if constexpr (L >= 4) { launch_kappa_le<L, 3>(ctx); return true; } break;
        case 4: if constexpr (L >= 4) { launch_kappa_le<L, 4>(ctx); return true; } break;
        case 6: if constexpr (L == 8 || L == 16) { launch_kappa_le<L, 6>(ctx); return true; } break;
        case 8: if constexpr (L == 8 || L == 16) { launch_kappa_le<L, 8>(ctx); return true; } break;
        default: break;
    }
    return false;
}

template <int L>
bool launch_t(rg_context* ctx, int E) {
    switch (E) {
        case 0: launch_le<L, 0>(ctx); return true;
        case 1: launch_le<L, 1>(ctx); return true;
        case 2: if constexpr (L >= 2) { launch_le<L, 2>(ctx); return true; } break;
        case 3: if constexpr (L >= 4) { launch_le<L, 3>(ctx); return true; } break;
        case 4: if constexpr (L >= 4) { launch_le<L, 4>(ctx); return true; } break;
        case 6: if constexpr (L == 8 || L == 16) { launch_le<L, 6>(ctx); return true; } break;
        case 8: if constexpr (L == 8 || L == 16) { launch_le<L, 8>(ctx); return true; } break;
        default: break;
    }
    return false;
}

}  // namespace

#define RG_K1_CAT2(a, b) a##b
#define RG_K1_CAT(a, b) RG_K1_CAT2(a, b)

namespace rg {
bool RG_K1_CAT(k1_launch_, RG_K1_L)(rg_context* ctx, int E) {
    bool ok = launch_t<RG_K1_L>(ctx, E);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess && ctx->launch_err.empty())
        ctx->launch_err = std::string("launch of k_update failed: ") + cudaGetErrorString(e);
    return ok;
}
// the reduced-cost recurrence for the same (L, E) variant, on ctx->stream; false: no fixed-width variant covers E
bool RG_K1_CAT(k1_kappa_launch_, RG_K1_L)(rg_context* ctx, int E) {
    bool ok = launch_kappa_t<RG_K1_L>(ctx, E);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess && ctx->launch_err.empty())
        ctx->launch_err = std::string("launch of k_kappa_update failed: ") + cudaGetErrorString(e);
    return ok;
}
}  // namespace rg
