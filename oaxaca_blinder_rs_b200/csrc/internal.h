// csrc/internal.h -- host-side declarations shared between the obboot translation units.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <string>
#include <vector>

#include "../../include/obboot.h"

namespace ob {

// ---- tile geometry of the Gram contraction (gram.cu) ----
constexpr int BM = 128;        // replicate slots per panel (M tile)
constexpr int BN = 128;        // sufficient-statistic columns per N tile
constexpr int KT = 32;         // rows (contraction) per pipeline stage
constexpr int GRAM_THREADS = 256;

struct CudaError { cudaError_t code; const char* what; const char* file; int line; };

#define OB_CUDA(expr)                                                          \
    do {                                                                       \
        cudaError_t e__ = (expr);                                              \
        if (e__ != cudaSuccess) throw ::ob::CudaError{e__, #expr, __FILE__, __LINE__}; \
    } while (0)

struct StatusError { ob_status code; std::string msg; };

// Number of variables V = K + 1 (design columns + outcome); sufficient-statistic columns
// P' = V(V+1)/2 (upper triangle of [x|y][x|y]^T, row-major); P = P' - 1 are algorithmic
// (the (y,y) cell rides along in the padding).
inline int64_t num_pairs(int V) { return (int64_t)V * (V + 1) / 2; }
// row stride (in doubles) of the packed design: >= V, congruent to 4 mod 8 so that the
// B-fragment shared-memory loads of gram.cu are bank-conflict free
inline int design_ldx(int V) { int l = V; while ((l & 7) != 4) ++l; return l; }
// index of pair (j,l), j <= l < V, in the row-major upper triangle
inline int64_t pair_index(int V, int j, int l) { return (int64_t)j * V - (int64_t)j * (j - 1) / 2 + (l - j); }

struct GroupData {               // one group's packed design in HBM
    int64_t n = 0;               // valid rows
    int64_t n_pad = 0;           // rows padded to a multiple of KT (zero rows)
    double* X = nullptr;         // [n_pad][ldx]: cols 0..K-1 design (intercept first), col K outcome, rest 0
    double* w = nullptr;         // [n_pad] sample weights (0 on padding) or nullptr
    double* Xs = nullptr;        // weighted only: sqrt(w_i) * X[i][:] (ols.rs:68-78), the operand of the Gram contraction
    const double* gram_operand() const { return Xs ? Xs : X; }
};

// ---- resample.cu ----
// counts layout per group: [panels][n_pad][BM] of count_t (uint8_t or uint16_t); slot 0 of panel 0
// is the point estimate (all ones on valid rows).
struct CountsArgs {
    void* C; int count_bytes; int64_t n; int64_t n_pad; int panels;
    int64_t slots;                // valid slots in this batch
    int first_slot;               // 1: slot 0 is the point estimate, replicates start at slot 1; 0: replicates from slot 0
    int64_t rep0;                 // global replicate id of local slot 0 (-1.. for the point slot)
    int group; uint64_t seed;
};
void counts_clear(const CountsArgs& a, cudaStream_t st);
// index-stream mode: idx [reps][n] u32 on device; overflow_flag (device int) set when a count saturates
void counts_from_indices(const CountsArgs& a, const uint32_t* d_idx, int* d_overflow, cudaStream_t st);
// native mode: Poisson(lambda) body + exact fix-up draws; d_colsum [slots] int64 scratch; d_flags[2] ints
void counts_philox(const CountsArgs& a, long long* d_colsum, int* d_flags, cudaStream_t st);

// ---- gram.cu ----
struct GramPlan {
    int V, ldx, panels, ntiles;
    int64_t n_pad[2];
    int segs[2], seg_rows[2];     // fixed row segmentation per group (function of n_pad only)
    int64_t units[2];             // panels * ntiles * segs
    int grid;
    int64_t num_partials;         // = units[0] + units[1]
    size_t smem_bytes; int stages;
    int tile_variant;             // 0: 1x8 warps, 128x16 warp tiles (default); 1: 2x4 warps, 64x32 warp tiles
};
GramPlan gram_make_plan(int V, int panels, const int64_t n_pad[2], int count_bytes, bool weighted, int num_sms);
struct GramArgs {
    const double* X[2]; const void* C[2];     // X: the (sqrt(w)-scaled when weighted) design
    int count_bytes;
    double* partials;            // [num_partials][BM*BN]
    const uint16_t* d_pairs;     // [ntiles*BN][2] column offsets (j,l) within a design row
    double* gram;                // out: [2][panels*BM][ntiles*BN]
};
void gram_launch(const GramPlan& plan, const GramArgs& args, cudaStream_t st, cudaEvent_t ev_main_begin = nullptr,
                 cudaEvent_t ev_main_end = nullptr);
std::vector<uint16_t> gram_pair_table(int V, int ntiles);

// ---- solve.cu ----
struct SolveArgs {
    const double* gram;      // [2][slots_pad][Pld]
    int64_t slots_pad; int Pld;
    int64_t slots;           // valid slots in this batch (slot 0 = point estimate when has_point)
    int K, n_cont, ref_kind;
    int n_norm; const int* d_norm_m; const int* d_norm_off; const int* d_norm_idx; const int* d_norm_has_base;
    int n_base; int S;
    int weighted;
    double na, nb;           // group row counts (n_obs)
    // outputs
    double* stats;           // [slots][S]
    int* status;             // [slots]
    double* beta_a; double* beta_b;    // [slots][K] (post-Yun) or nullptr
    double* point_extra;     // slot 0 only: [xa_mean K | xb_mean K | beta_star K | raw beta_a K | raw beta_b K | total_gap] or nullptr
};
void solve_launch(const SolveArgs& a, cudaStream_t st);
size_t solve_smem_bytes(int K, bool pooled, int n_norm);

// ---- reduce_stats.cu ----
// stats [reps][S] + status [reps] (device) -> se,p,lo,hi,t [S] each (device, contiguous 5*S) and n_ok
void reduce_stats_launch(const double* stats, const int* status, int64_t reps, int S,
                         const double* point_stats, double* out5S, long long* n_ok, cudaStream_t st);
constexpr int64_t REDUCE_MAX_REPS = 16384;

// ---- pack.cu ----
struct PackArgs {
    int64_t n; int n_cont, n_cat;
    const double* const* d_cont;     // device array of device column pointers
    const int32_t* const* d_cat;     // device array of device code-column pointers
    const int32_t* d_cat_levels;     // [n_cat]
    const int32_t* d_dummy_start;    // [n_cat] first design column of each categorical
    const double* d_y; const double* d_w; const uint8_t* d_group;
    int K, ldx;
};
// pass 1+2: per-block group counts and their exclusive scan (d_block_counts [nblocks][2] becomes the
// per-block base rank; d_totals[2] = group sizes); flags[0] = negative weight seen, flags[1] = bad code
int pack_num_blocks(int64_t n);
void pack_count_scan(const PackArgs& a, long long* d_block_counts, long long* d_totals, int* d_flags, cudaStream_t st);
// pass 3: staged transpose/scatter of the rows into the packed per-group designs
void pack_scatter(const PackArgs& a, const long long* d_block_base, GroupData ga, GroupData gb, int* d_flags,
                  cudaStream_t st);
// Xs[i][:] = sqrt(w[i]) * X[i][:] for all V = K+1 columns (WLS as OLS on sqrt(w)-scaled data, ols.rs:68-78)
void scale_rows_launch(const GroupData& g, int ldx, cudaStream_t st);
// residuals of the point estimate: r = y - X beta (ols.rs:118-119) for one group
void residuals_launch(const GroupData& g, int K, int ldx, const double* d_beta, double* d_out, cudaStream_t st);

// ---- rif.cu ----
// in-place RIF transform (math/rif.rs:14-88) of the outcome column (col K) of a packed group
void rif_transform(const GroupData& g, int K, int ldx, double tau, void* d_scratch, size_t scratch_bytes,
                   cudaStream_t st);
size_t rif_scratch_bytes(int64_t n);

}  // namespace ob
