/* relp_gpu_test.h -- test hooks of librelp_gpu.so.  NOT part of the drop-in boundary (include/relp_gpu.h):
 * no reference counterpart exists; only tests/ and scripts/ call these. */
#ifndef RELP_GPU_TEST_H
#define RELP_GPU_TEST_H

#include "relp_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

int rg_debug_scalars(rg_context* ctx, void* out, int64_t bytes);
int rg_debug_vector(rg_context* ctx, int32_t which, uint64_t* out);
/* runs one device big-integer primitive on W-limb operands: see relp_gpu.cu */
int rg_selftest(int32_t op, int32_t W, const uint64_t* a, const uint64_t* b, const uint64_t* c,
                const uint64_t* d, int64_t s, uint64_t* out);

/* measurement hook: peak rate of the IMAD.WIDE carry-chain mix of the K1 products (the integer-pipe roofline
 * denominator of SURVEY section 8d), measured on `device` for about `seconds`; IMAD.WIDE instructions per second */
int rg_measure_imad_peak(int32_t device, double seconds, double* imad_per_s);

#ifdef __cplusplus
}
#endif
#endif /* RELP_GPU_TEST_H */
