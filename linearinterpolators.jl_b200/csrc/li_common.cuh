// li_common.cuh -- internal declarations shared by the translation units of libli_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <map>
#include <vector>

#include "li_b200.h"

namespace li {

void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Process-wide switches of the dispatch (li_set_option / li_get_option).  Read from the environment ONCE, when the
// library is loaded (LI_B200_NO_FUSED, LI_B200_NO_STREAM_DN, LI_B200_NO_BANDS, LI_B200_AFFINE_PLAIN, LI_B200_TRACE,
// LI_B200_ADJ_DEBUG, LI_B200_NO_MULTI, LI_B200_EXP); the hot path only reads these atomics, never getenv().
enum Option {
    OPT_NO_FUSED = 0,      // fused 2-D kernels off: per-dimension passes (cross-check of the fused kernels)
    OPT_NO_STREAM_DN,      // streaming / window kernels of li_tab_apply off: generic gather kernels
    OPT_NO_BANDS,          // 1: host-buffer pipeline off (upload, apply, download); 2..8: that many bands (default 8;
                           // measured on config #2: 2 / 3 / 4 / 6 / 8 bands 7.9 / 7.75 / 7.6 / 8.1 / 7.8 ms per step, 16: 8.2)
    OPT_AFFINE_PLAIN,      // affine operators: one thread per pixel on global memory
    OPT_TRACE,             // print dispatch decisions to stderr
    OPT_ADJ_DEBUG,         // timing probes of the fused adjoint (see li_sep2d.cu)
    OPT_NO_MULTI,          // multi-frame accumulate kernel off: one launch per frame
    OPT_EXP,               // development: bit mask of A/B switches (2: 3-D applies without the fused 3-D kernels,
                           // 4: affine operators with aligned axes without the separable shortcut,
                           // 8: all-reduces through NCCL only (no peer-memory all-reduce);
                           // bits 8..11: waves of CTAs of the fused forward kernel, 0 = the built-in choice)
    OPT_COUNT
};
extern std::atomic<int> g_options[OPT_COUNT];
inline int opt(Option o) { return g_options[o].load(std::memory_order_relaxed); }

#define LI_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            li::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                          __LINE__);                                                       \
            return e__ == cudaErrorMemoryAllocation ? LI_ERR_NOMEM : LI_ERR_CUDA;          \
        }                                                                                  \
    } while (0)

#define LI_CHECK_LAUNCH()                                                                  \
    do {                                                                                   \
        li::count_launch();                                                                \
        LI_CUDA(cudaGetLastError());                                                       \
    } while (0)

#define LI_REQUIRE(cond, status, ...)                                                      \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            li::set_error(__VA_ARGS__);                                                    \
            return (status);                                                               \
        }                                                                                  \
    } while (0)

inline int current_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}
constexpr int LI_HEAVY_COLUMN = 512;
inline size_t dtype_size(int dtype) { return dtype == LI_F32 ? 4 : 8; }
inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Tile granularities for which per-table footprint bounds are cached.
struct AdjointIndex;

}  // namespace li

// Device-resident table.  J is kept as int32, 0-based; the reference layout
// (Int64, 1-based, S x nrows) is produced by li_table_export.
struct li_table {
    int32_t dtype = 0, S = 0;
    int64_t nrows = 0, ncols = 0;
    uint64_t uid = 0;         // unique per table of this process (plan caches key on it: pointers get reused)
    int32_t *J = nullptr;     // [nrows*S], 0-based, clamped
    void *W = nullptr;        // [nrows*S], T
    int32_t *base = nullptr;  // [nrows] canonical window base (0-based tap s is clamp(base+s)); valid iff windowed
    bool windowed = false;
    bool monotone = false;
    int max_step = 1;         // max(base[i+1] - base[i], 1) for monotone tables (1 <=> up-sampling or unit rate)
    // lazily built adjoint helpers (guarded by mu)
    mutable std::mutex mu;
    // generic transposed (CSC) table: entries sorted by (column, row, tap)
    mutable int32_t *colptr = nullptr;  // [ncols+1]
    mutable int32_t *tent = nullptr;    // [nrows*S] entry id e = row*S + tap
    // columns with more than LI_HEAVY_COLUMN entries (Flat clamping piles every row hanging over an edge onto the
    // first / last column): summed by a whole CTA each in the gather adjoint (li_tab_apply.cu), not by one thread
    mutable int32_t *heavy = nullptr;   // [nheavy] column indices
    mutable int nheavy = 0;
    // monotone windowed tables: bptr[j] = first row with base >= j-(S-1), j in [0, ncols+S-1] (see li_table.cu)
    mutable int32_t *rowstart = nullptr;
    // adjoint of the fused kernels: windows folded into [0, ncols) (clamped taps merged into the first/last tap)
    mutable void *Wfold = nullptr;      // [nrows*S] folded weights
    mutable int32_t *bfold = nullptr;   // [nrows] folded window bases, in [0, ncols - S]
    mutable int32_t *bpf = nullptr;     // bptr table over bfold, [ncols + S]
    mutable int32_t *irrp = nullptr;    // [ncols+1] prefix count of dst rows fed by a number of src rows != reg_R
    mutable int reg_R = 0;              // the regular number of src rows per dst row (1 or 2), 0 if none
    // per tile height: device list of tile rows, regular ones first; second = number of regular tile rows
    mutable std::map<int, std::pair<int32_t *, int>> tile_rows;
    mutable std::map<int, int> adjf_rows;
    mutable std::map<int, int> feed_rows;
    // streaming adjoint: cost-balanced chunks of columns, key = number of chunks:
    // (device array of nchunks+1 boundaries, longest chunk, most rows feeding one chunk)
    struct ChunkPlan {
        int32_t *starts = nullptr;
        int crmax = 0, nrmax = 0;
        std::vector<int32_t> starts_host, src_end_host;   // per chunk: first dst row; src rows needed (exclusive)
    };
    mutable std::vector<int32_t> base_host;   // lazily: host copy of `base` (band planning of the host pipeline)
    mutable std::map<int, ChunkPlan> chunk_plans;
    // cached footprint bounds: key = tile rows G -> max over aligned tiles of (max base - min base)
    mutable std::map<int, int> fwd_span;
    // cached: key = tile cols G -> max over aligned column tiles of number of contributing rows
    mutable std::map<int, int> adj_rows;
};

namespace li {
int table_ensure_csc(const li_table *t, cudaStream_t st);
int table_ensure_rowstart(const li_table *t, cudaStream_t st);
int table_fwd_span(const li_table *t, int G, cudaStream_t st, int *span);
int table_adj_rows(const li_table *t, int G, cudaStream_t st, int *rows);
int table_ensure_fold(const li_table *t, cudaStream_t st);            // needs ncols >= S
int table_adjf_rows(const li_table *t, int G, cudaStream_t st, int *rows);  // like table_adj_rows, folded bases
// max over aligned tiles of G columns of the number of rows with a tap landing in the tile (clamped taps included:
// the first / last column collect every row hanging over that edge)
int table_feed_rows(const li_table *t, int G, cudaStream_t st, int *rows);
// tile rows of height TD2 over the columns of t, regular ones (all dst rows fed by exactly reg_R rows, none clamped)
// first: *list (device) has ntiles entries, the first *nregular of them regular
int table_tile_rows(const li_table *t, int TD2, cudaStream_t st, const int32_t **list, int *nregular);
// `nchunks` chunks of columns of about equal cost for the streaming adjoint: a chunk holding a column that is not
// fed by exactly reg_R rows (or touches a folded window) runs the slower generic loop and is made shorter by `penalty`
// (fraction of a chunk).  *starts (device) has nchunks+1 boundaries.
// `weights` (optional, one per chunk; `weights_key` < 2^14 identifies them in the plan cache): relative shares of the
// chunks instead of equal ones.
int table_chunk_plan(const li_table *t, int nchunks, double penalty, cudaStream_t st, const int32_t **starts, int *crmax,
                     int *nrmax, const li_table::ChunkPlan **plan = nullptr, const std::vector<double> *weights = nullptr,
                     int weights_key = 0);
int table_base_host(const li_table *t, cudaStream_t st, const std::vector<int32_t> **base);
int table_identity(int dtype, int64_t n, cudaStream_t st, const li_table **out);   // cached S = 1 identity operator
// builds a table from device positions into preallocated device arrays (used by the OTF operators)
int table_create_impl(const li_kernel_desc *ker, const void *pos_dev, int64_t nrows, int64_t ncols,
                      cudaStream_t st, li_table **out);
int tab_apply_impl(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha,
                   const void *src, double beta, void *dst, int variant, cudaStream_t st);
int vscale_impl(int dtype, void *y, int64_t n, double beta, cudaStream_t st);
int staging(int slot, size_t bytes, void **out, void *stream_key = nullptr);  // grow-only device scratch per (thread, stream)
int sep2d_fused_supported(const li_table *t1, const li_table *t2);
int sep2d_host_banded(int adjoint, const li_table *t1, const li_table *t2, double alpha, const void *src_host, void *dst_host,
                      void *dsrc, void *ddst, cudaStream_t st, bool *done);
int sep2d_fused_apply(int adjoint, const li_table *t1, const li_table *t2, int64_t nbatch,
                      double alpha, const void *src, double beta, void *dst, cudaStream_t st,
                      bool *done);
}  // namespace li
