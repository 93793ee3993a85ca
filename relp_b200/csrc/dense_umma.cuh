// tcgen05 form of the exact dense dots: the byte-slice GEMM
//     R[s][j] = sum_i slice_s(vec_i) * a_ij          (u8 x s8 -> s32)
// on the 5th-generation tensor cores of sm_100a.  Per CTA: 128 matrix columns (UMMA M) x 160 slice rows (UMMA N);
// K = constraint rows in blocks of 128.  TMA (cp.async.bulk.tensor, 128-byte swizzle) stages the int8 block and the
// slice rows in a 4-deep shared-memory ring, one elected thread issues tcgen05.mma.kind::i8 with the accumulators
// in TMEM, tcgen05.commit releases the ring slots, and the four warps read the accumulators back with tcgen05.ld
// for the epilogue.  K blocks whose vector entries are all zero are skipped by producer and issuer alike.
// The mma.sync kernel (k_dense_mma) remains for small blocks and as the cross-check of the parity tests.
#pragma once
#include <cuda.h>          // CUtensorMap (types only; cuTensorMapEncodeTiled is resolved at run time)
#include "engine.cuh"

namespace rg {

constexpr int UM_M = 128;          // matrix columns per CTA (rows of D = TMEM lanes)
constexpr int UM_N = 160;          // slice rows per CTA (columns of D)
constexpr int UM_KB = 128;         // constraint rows per pipeline stage (one 128-byte swizzle row of int8)
constexpr int UM_STAGES = 4;
constexpr int UM_A_BYTES = UM_M * UM_KB;
constexpr int UM_B_BYTES = UM_N * UM_KB;
constexpr int UM_TMEM_COLS = 256;  // power of two >= UM_N
constexpr size_t UM_SMEM = (size_t)UM_STAGES * (UM_A_BYTES + UM_B_BYTES) + 1024 /* alignment */ + 256 /* barriers */;

__device__ __forceinline__ unsigned um_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void um_bar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(um_smem(bar)), "r"(count));
}
__device__ __forceinline__ void um_bar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(um_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void um_bar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(um_smem(bar)), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();     // a lost arrival must not hang the device
    } while (!ok);
}
__device__ __forceinline__ void um_tma_2d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(um_smem(dst)), "l"(map), "r"(um_smem(bar)), "r"(c0), "r"(c1) : "memory");
}
// shared-memory matrix descriptor, K-major, 128-byte swizzle (cute::UMMA::SmemDescriptor): start address >> 4,
// leading byte offset 1 (unused inside one swizzle atom), stride byte offset = 8 rows x 128 B = 1024 >> 4,
// version 1 (Blackwell), layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ unsigned long long um_desc(unsigned smem_addr) {
    return (unsigned long long)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor of tcgen05.mma.kind::i8 (cute::UMMA::InstrDescriptor): D = s32, A = signed 8 bit (the matrix),
// B = unsigned 8 bit (the slices), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr unsigned UM_IDESC = (2u << 4) | (1u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((unsigned)(UM_N >> 3) << 17) |
                              ((unsigned)(UM_M >> 4) << 24);

__global__ void __launch_bounds__(128, 1)
k_dense_umma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int jd0, int jd1,
             size_t mp, const int* __restrict__ chunknz, const int* bits, int LV, int rows_per_kslice,
             int* __restrict__ R, size_t rstride_k, int rpitch, const Scalars* sc) {
    extern __shared__ unsigned char um_raw[];
    __shared__ unsigned tmem_base_s;
    if (sc->status != ST_RUN) return;
    const int nb = eff_bytes(bits, LV);
    const int rows_needed = nb + 1;                               // byte rows + the sign row
    const int row0 = blockIdx.z * UM_N;                            // slice-row layer of this CTA
    if (row0 >= rows_needed) return;                              // uniform over the CTA
    unsigned char* tiles = (unsigned char*)(((size_t)um_raw + 1023) & ~(size_t)1023);
    unsigned long long* full_bar = (unsigned long long*)(tiles + (size_t)UM_STAGES * (UM_A_BYTES + UM_B_BYTES));
    unsigned long long* empty_bar = full_bar + UM_STAGES;
    unsigned long long* done_bar = empty_bar + UM_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col0 = jd0 + blockIdx.x * UM_M;
    const int kb0 = blockIdx.y * (rows_per_kslice / UM_KB);
    const int kb1 = min((int)((mp + UM_KB - 1) / UM_KB), kb0 + rows_per_kslice / UM_KB);
    if (threadIdx.x == 0) {
        for (int s = 0; s < UM_STAGES; ++s) { um_bar_init(&full_bar[s], 1); um_bar_init(&empty_bar[s], 1); }
        um_bar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                              // one warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem(&tmem_base_s)),
                     "r"((unsigned)UM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_base_s;
    const int nchunk = (int)(mp / 64);
    auto block_nonzero = [&](int kb) {
        const int c = 2 * kb;
        return chunknz[c] != 0 || (c + 1 < nchunk && chunknz[c + 1] != 0);
    };
    int issued = 0;
    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        int it = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            if (!block_nonzero(kb)) continue;
            const int s = it % UM_STAGES;
            um_bar_wait(&empty_bar[s], ((it / UM_STAGES) & 1) ^ 1);
            um_bar_expect_tx(&full_bar[s], UM_A_BYTES + UM_B_BYTES);
            unsigned char* sa = tiles + (size_t)s * (UM_A_BYTES + UM_B_BYTES);
            um_tma_2d(sa, &mapA, &full_bar[s], kb * UM_KB, col0);
            um_tma_2d(sa + UM_A_BYTES, &mapB, &full_bar[s], kb * UM_KB, row0);
            ++it;
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        int it = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            if (!block_nonzero(kb)) continue;
            const int s = it % UM_STAGES;
            um_bar_wait(&full_bar[s], (it / UM_STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned sa = um_smem(tiles + (size_t)s * (UM_A_BYTES + UM_B_BYTES));
            const unsigned long long da = um_desc(sa), db = um_desc(sa + UM_A_BYTES);
#pragma unroll
            for (int k = 0; k < UM_KB / 32; ++k) {                 // UMMA K = 32 bytes of int8: +2 in the address field
                const unsigned acc = (it > 0 || k > 0) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(da + 2ull * k), "l"(db + 2ull * k), "r"(UM_IDESC), "r"(acc) : "memory");
            }
            // frees the ring slot once the MMAs above have read it
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(um_smem(&empty_bar[s])) : "memory");
            ++it;
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(um_smem(done_bar)) : "memory");
        issued = it;
    }
    // every thread learns whether anything was accumulated (all-zero vector: the results are zero)
    issued = __shfl_sync(0xffffffffu, issued, 0);
    __shared__ int issued_s;
    if (threadIdx.x == 32) issued_s = issued;
    __syncthreads();
    issued = issued_s;
    // ===== epilogue: TMEM -> registers -> R[k-slice][slice row][column] =====
    if (issued > 0) {
        um_bar_wait(done_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int j = col0 + warp * 32 + lane;                          // TMEM lane = matrix column of the tile
    int* base = R + (size_t)blockIdx.y * rstride_k;
    const int nrows = min(UM_N, rows_needed - row0);                // slice rows of this layer that are consumed
    for (int n0 = 0; n0 < nrows; n0 += 16) {
        unsigned v[16];
        if (issued > 0) {
            const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + (unsigned)n0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0;
        }
        if (j < jd1) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (n0 + i < nrows) base[(size_t)(row0 + n0 + i) * rpitch + (j - jd0)] = (int)v[i];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)UM_TMEM_COLS) : "memory");
}

}  // namespace rg
