// csrc/ingest.cu -- device side of the ingest step that precedes the pack: the cleaning and
// coding the reference does on the host before run() reaches the group split --
//   clean_dataframe          builder.rs:760-784   drop every row with a null in any used column
//   create_dummies_manual    builder.rs:380-418   levels = sorted unique values of the CLEANED frame, first = base
//   split_groups             builder.rs:61-102    sorted unique group values; A = first value that is not the reference
// for a frame whose string columns arrive dictionary-encoded (Arrow DictionaryArray / polars Categorical physical
// codes / pandas Categorical: int32 code per row, < 0 = null, dictionary in arbitrary order on the host).
//
// Two passes over the staged columns, both HBM-bound:
//   scan   row validity (AND over the used columns) + which dictionary entries occur among the valid rows
//          -> the host sorts the few present strings and derives group map / level remaps (no O(n) host work)
//   apply  group byte (0 = A, 1 = reference, 255 = dropped / other group) and remapped level codes, in place
// after which the ordinary pack kernels (pack.cu) run on the cleaned columns.
#include "common.cuh"
#include "internal.h"

#include <algorithm>

namespace ob {

__global__ void __launch_bounds__(256) ingest_scan_kernel(const IngestScanArgs a) {
    __shared__ int kept;
    if (threadIdx.x == 0) kept = 0;
    __syncthreads();
    int mine = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        bool ok = true;
        for (int c = 0; c < a.n_valid; ++c) ok &= a.valid[c][i] != 0;
        for (int c = 0; c < a.n_nan; ++c) { const double v = a.nan_cols[c][i]; ok &= v == v; }
        for (int c = 0; c < a.n_dict; ++c) {
            const int code = a.codes[c][i];
            if (code >= a.dict_size[c]) atomicOr(a.flags, 1);       // malformed dictionary column
            ok &= code >= 0 && code < a.dict_size[c];
        }
        a.row_valid[i] = ok ? 1 : 0;
        if (ok) {
            ++mine;
            // benign race: every writer stores the same byte
            for (int c = 0; c < a.n_dict; ++c) a.present[c][a.codes[c][i]] = 1;
        }
    }
    if (mine) atomicAdd(&kept, mine);
    __syncthreads();
    if (threadIdx.x == 0 && kept) atomicAdd(reinterpret_cast<unsigned long long*>(a.kept), (unsigned long long)kept);
}

__global__ void __launch_bounds__(256) ingest_apply_kernel(const IngestApplyArgs a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        const bool ok = a.row_valid[i] != 0;
        uint8_t g = 255;
        if (ok) {
            const int m = a.group_map[a.group_codes[i]];
            g = m == 0 ? 0 : (m == 1 ? 1 : 255);
        }
        a.group_out[i] = g;
        for (int q = 0; q < a.n_cat; ++q) {
            int32_t* col = a.cat_codes[q];
            int v = 0;
            if (ok) {
                v = a.remap[a.remap_off[q] + col[i]];
                if (v < 0) { atomicOr(a.flags, 2); v = 0; }          // a present value the host left unmapped
            }
            col[i] = v;
        }
    }
}

void ingest_scan_launch(const IngestScanArgs& a, cudaStream_t st) {
    if (a.n == 0) return;
    const unsigned blocks = (unsigned)std::min<long long>((a.n + 255) / 256, 148 * 16);
    ingest_scan_kernel<<<blocks, 256, 0, st>>>(a);
    OB_CUDA(cudaGetLastError());
}

void ingest_apply_launch(const IngestApplyArgs& a, cudaStream_t st) {
    if (a.n == 0) return;
    const unsigned blocks = (unsigned)std::min<long long>((a.n + 255) / 256, 148 * 16);
    ingest_apply_kernel<<<blocks, 256, 0, st>>>(a);
    OB_CUDA(cudaGetLastError());
}

}  // namespace ob
