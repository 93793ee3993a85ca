// Device-side state shared by all kernels of the exact simplex engine.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/relp_gpu.h"
#include "bigint.cuh"

namespace rg {

enum DevStatus { ST_RUN = 0, ST_OPTIMAL = 1, ST_UNBOUNDED = 2, ST_PROMOTE = 3, ST_FATAL = 4 };

// Scalars of the iteration in flight; written by 1-thread kernels, read by all others.
struct Scalars {
    int status;          // DevStatus
    int q;               // entering provider column id
    int p;               // pivot row as LOCAL carry row index (1..nloc), -1 when another rank owns it
    int leaving;         // column id leaving the basis
    int sgn;             // sign of the pivot element numerator a = u[p]
    int t, E;            // ctz(D) and ceil(t/64): extra limbs the exact division needs
    int t2, E2;          // same for D^2 (steepest-edge recurrence)
    int maxbits_carry;   // max |numerator| bit length in the carry (all rows known to this rank)
    int maxbits_new;     // being accumulated by the update kernel
    int maxbits_u;       // of the current pivot column (incl. the cost row entry)
    int maxbits_rowp;    // of the staged pivot row
    int maxbits_tmp;     // scratch (phase switch)
    int maxbits_s;       // of the weighted factor vector s_i = u_i (W / w_B(i))^2 (weighted problems)
    int bits_D;
    int predicted;       // predicted bit length of the update's results
    int last_selected;   // FirstProfitableWithMemory state (-1 = none)
    int found;           // generic "search hit" index (-1 = none)
    int bp_nonzero;      // remove_artificial: is b_p != 0
    int pg;              // pivot row as global carry row index (1..m)
    int row_lo, nloc;    // this rank's block of constraint rows [row_lo, row_lo + nloc)
    int rank, world;
    int nk;              // number of non-trivial carry columns (entries of klist), column 0 included
    int nnz_s;           // list mode: local rows with a non-zero work-vector factor (entries of nzrows)
    int row0_ticket;     // k_ftran_row0: blocks finished (last one folds the slices)
    int fatal;           // why status became ST_FATAL: 1 overflow beyond 16 limbs, 2 kernel variant too narrow, 3 zero pivot element
    u64 D[RG_MAXL];          // current denominator (positive)
    u64 a[RG_MAXL + 2];      // pivot element numerator u[p] (replicated on every rank)
    u64 Dnew[RG_MAXL];       // |a|: denominator after the pivot
    u64 Dold[RG_MAXL];       // denominator before the last pivot (BasisChangeComputationInfo export)
    u64 Dinv[2 * RG_MAXL + 2];   // inverse of odd(D) mod 2^(64 (L+E)), E <= L (+2 spare)
    u64 A[2 * RG_MAXL + 2];      // |a| * Dinv mod 2^(64 (L+E))
    u64 up[2 * RG_MAXL + 2];     // replacement for u[p]: a - D (makes the pivot row uniform)
    u64 Gq[RG_MAXW];         // steepest edge: Ghat of the entering column
    u64 S1[RG_MAXW], S2[RG_MAXW], S3[RG_MAXW];   // a^2/D^2, 2a/D^2, Gq/D^2 (2-adic)
};

// what the host reads back after every iteration (pinned)
struct HostMirror {
    int status, q, p, leaving;          // device status; NEXT entering column; last pivot row; last leaving
    int t_next, bits_D, maxbits_carry, predicted;
    int found, sgn, maxbits_tmp, pivoted;   // pivoted: a basis change happened since the host cleared it
    int q_done, p_done, leaving_done, nk;   // the pivot that was performed (written by k_finalize); active columns
    int fatal, pad0, pad1, pad2;            // Scalars::fatal
};

struct Csc {
    long long* colptr = nullptr;   // n+1
    int* rowidx = nullptr;         // nnz
    long long* vals = nullptr;     // nnz
    long long nnz = 0;
};

// widths derived from the carry limb count L
__host__ __device__ constexpr int LU_of(int L) { return L + 2; }          // pivot column, costs, row dots
__host__ __device__ constexpr int LW_of(int L) { return 2 * L + 5; }      // work vector
__host__ __device__ constexpr int LS_of(int L) { return 2 * L + 7; }      // work vector . column
__host__ __device__ constexpr int LG_of(int L) { return 2 * L + 6; }      // Ghat = gamma * D^2 W^2 / w_j^2

}  // namespace rg

struct rg_context {
    int device = 0;
    int rank = 0, world = 1;
    int row_lo = 0, nloc = 0;          // constraint rows owned by this rank
    int d0 = 0, d1 = 0, s0 = 0, s1 = 0;  // provider columns priced by this rank: dense-block slice [d0,d1), CSC slice [s0,s1)
    void* nccl_comm = nullptr;         // ncclComm_t (world > 1)
    bool own_comm = false;             // communicator created for this context only (RG_NCCL_NO_CACHE)
    u64* xsend = nullptr;              // exchange buffers (world > 1)
    u64* xrecv = nullptr;
    size_t xbytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;          // side stream: steepest-edge scalars overlap the K1 update
    int* nzrows = nullptr;             // list mode: compacted local rows with s_i != 0 (nloc entries)
    cudaEvent_t ev_side0 = nullptr, ev_side1 = nullptr, ev_side2 = nullptr, ev_work = nullptr, ev_side3 = nullptr, ev_nu = nullptr, ev_ft0 = nullptr, ev_ft1 = nullptr;
    cudaStream_t side3 = nullptr;      // third side stream: steepest-edge scalars (k_scalars_se)
    cudaStream_t side2 = nullptr;      // second side stream: nu / sigma column dots, concurrent with K1
    int m = 0, n = 0;
    int ld = 0;                 // carry leading dimension in entries (multiple of 16)
    int L = 2;                  // current limb count
    size_t plane = 0;           // entries per carry plane: (nloc+1) * ld in dense mode, ld (cost row only) in list mode
    u64* carry = nullptr;       // L planes
    // packed active block (list mode): rows 1..nloc of the non-trivial carry columns,
    // entry (i, t) = C[i][klist[t]] at i * cap + t in each plane -- every access of the hot kernels is coalesced
    u64* pk = nullptr;          // L planes x (nloc+1) x cap
    int cap = 0;                // column capacity (multiple of 32), grown by doubling
    size_t pplane = 0;          // entries per packed plane = (nloc+1) * cap
    u64* u = nullptr;           // LU planes x ld      current pivot column (rows 0..m)
    u64* rowp = nullptr;        // L  planes x ld      staged pivot row
    u64* omega = nullptr;       // LW planes x ld      work vector
    u64* omega_part = nullptr;  // chunks x LW planes x ld
    int work_chunks = 0;
    u64* tmprow = nullptr;      // LU planes x ld      phase-switch row
    u64* svec = nullptr;        // 1 plane x ld        basic costs (phase switch)
    // active-column bookkeeping: carry column k is `trivial` while it equals
    // D * e_k on rows 1..m; trivial columns are never read or written, the others are listed in klist
    unsigned char* triv = nullptr;   // ld flags
    int* klist = nullptr;            // ld column indices
    int* kpos = nullptr;             // ld: position of a column in klist
    int list_chunks = 0;             // row chunks of the column sums in list mode
    int list_pcols = 0;              // column slots of the list-mode partial sums
    long long* aq = nullptr;         // m: the entering column scattered densely (list-mode FTRAN)
    u64* row0_part = nullptr;        // k_ftran_row0 slices: 32 x (RG_MAXL + 2) words
    bool list_mode = false;
    int nk_host = 1;                 // upper bound of sc->nk known to the host
    int dense_carry_opt = 0;         // rg_options: 1 = never use the active-column list
    u64* us2 = nullptr;         // LU+1 planes x ld    pivot column times row factors^2 (weighted problems)
    u64* ufull = nullptr;       // LU+1 planes x ld    row-sharded runs: the whole factor vector (all-gathered blocks)
    bool ufull_valid = false;
    // weights of a prescaled rational problem; all 1 for integer problems
    bool weighted = false;
    long long* wf = nullptr;      // n: W / w_j
    long long* wcol = nullptr;    // n: w_j
    long long* artf = nullptr;    // m: W / (weight of the artificial of row i)
    long long* artcost = nullptr; // m: phase-one cost numerator of the artificial of row i
    long long* rowf = nullptr;    // m: factor of the variable currently basic in row i
    rg::Csc A;
    std::vector<long long> h_colptr, h_vals;   // host copy of the CSC structure (argument validation)
    std::vector<int> h_rowidx;
    // optional dense int8 block holding provider columns [0, nd): row-major and column-major copies
    int nd = 0;
    signed char* Acm = nullptr; size_t ldc = 0;    // [nd][ldc]
    int* dR = nullptr; size_t dR_words = 0;        // tensor-core dense dots: slice-by-column s32 products
    unsigned char* dSl = nullptr; int* dchunk = nullptr; size_t dmp = 0;   // byte slices of the vector, chunk flags
    int* dR2 = nullptr; unsigned char* dSl2 = nullptr; int* dchunk2 = nullptr;   // second scratch set: pricing dot
    // tcgen05 path of the dense dots (dense_umma.cuh): TMA descriptors of the int8 block and of the two slice buffers
    CUtensorMap mapA, mapB, mapB2;
    bool umma_ok = false;
    unsigned char* dSl_primary = nullptr;      // the first set's slice buffer (tells the sets apart while they are swapped)
    long long* cost = nullptr;  // n
    long long* rhs = nullptr;   // m
    int* basis = nullptr;       // m column ids
    unsigned char* inbasis = nullptr;  // n
    u64* kappa = nullptr;       // LU planes x n  relative cost numerators
    u64* nu = nullptr;          // LU planes x n  pivot-row . column
    u64* sigma = nullptr;       // LS planes x n  work-vector . column
    u64* tau = nullptr;         // LU+3 planes x n  factor vector (trivial part) . column: split sigma dot
    u64* G = nullptr;           // LG planes x n  Ghat
    int* cand = nullptr;        // block winners scratch
    double* score = nullptr;    // max(n, m) filter scores
    rg::Scalars* sc = nullptr;
    rg::HostMirror* hm = nullptr;      // pinned host
    rg::HostMirror* hm_dev = nullptr;  // device alias of hm
    int rule = 2;
    bool rule_ready = false;
    bool have_column = false;
    bool selected = false;             // sc->q holds the entering column chosen for the current basis
    bool identity_carry = false;       // B^-1 == I and D == 1 (fresh init)
    bool work_valid = false;           // omega holds the work vector of the last basis change
    int t_cur = 0;                     // ctz(D) of the current denominator (host copy)
    long long pivots = 0, promotions = 0, launches = 0;
    long long pivots_at[RG_NWIDTHS] = {};
    long long demotions = 0;
    int demote_need = 0;               // upper bound of the carry's bit length after the last pivot (0 = unknown)
    u32* bn = nullptr;                 // K1 row factors Bn_i (k_bn_rows): (nloc + 2) rows of 2 (L + 8) + 1 words
    int k1_items_min_limbs = 10;       // list mode: widths from here on run the warp-granular K1 (RG_K1_ITEMS_MINL);
                                       // at 8 limbs k_update with its register prefetch measured 3 % faster on config 4
    bool kappa_valid = false;          // kappa holds the reduced costs of the current carry (k_kappa_update may run)
    bool ftran_overlap = true;         // RG_NO_FTRAN_OVERLAP=1: cost-row FTRAN on the main stream
    bool kappa_recur = true;           // RG_NO_KAPPA_RECUR=1: always price from the cost row
    int k1_items_prefetch = 1;         // RG_K1_NOPF=1: no L1 prefetch of the next row's entry
    int k1_items_rows = 8;             // rows per full work item (RG_K1_ITEMS_ROWS, <= 32)
    bool pow2_only = false;            // RG_WIDTH_LADDER=pow2: widths 1, 2, 4, 8, 16 only
    int demote_margin = 16;            // bits of slack a narrower width must leave (RG_DEMOTE_MARGIN); numerators
                                       // move by ~6 bits per pivot, a width change costs about half a pivot
    int demote_floor = 8;              // narrowest width a demotion may reach (RG_DEMOTE_FLOOR)
    int profile = 0;                   // 0 off, 1 events around K1 only, 2 events around every phase
    // CUDA graphs of one fused iteration, keyed by everything that shapes the launch sequence
    struct GraphEntry { long long key; cudaGraphExec_t exec; long long launches; };
    std::vector<GraphEntry> graphs;
    bool use_graphs = true;
    bool graph_nccl = false;           // replay graphs in row-sharded runs too (RG_GRAPH_NCCL=1)
    bool capturing = false;
    int nk_grid = 128;                 // list-mode grid bound (multiple of 128, >= nk + 1)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evt0 = nullptr, evt1 = nullptr;
    cudaEvent_t evp[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double phase_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // ftran+ratio+copyrow, work, scalars, K1, finalize+SE update, price+select, mirror
    long long k1_launches[RG_NWIDTHS] = {};
    double k1_ms[RG_NWIDTHS] = {};
    double k1_bytes[RG_NWIDTHS] = {};   // algorithmic bytes / IMAD.WIDE of the timed K1 launches (roofline accounting)
    double k1_imads[RG_NWIDTHS] = {};
    double k1_cur_bytes = 0, k1_cur_imads = 0;   // of the launch in flight
    double timer_ms = 0;
    std::string err;
    std::string launch_err;            // first failed kernel launch since the last synchronising call
};
