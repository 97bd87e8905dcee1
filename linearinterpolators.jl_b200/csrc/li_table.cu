// li_table.cu -- device-side table generation (K1), analysis and transposition (K2).
//
// Replaces __maketables (reference src/tabulated.jl:137-189) and _sparsecoefs
// (src/sparse.jl:140-164): getcoefs (src/boundaries.jl:77-156) for every row.
#include <math_constants.h>

#include <climits>
#include <functional>
#include <map>
#include <tuple>
#include <vector>

#include <cub/cub.cuh>

#include "li_common.cuh"
#include "li_kerdev.cuh"
#include "li_weights.cuh"

namespace li {

// ---------------------------------------------------------------------------
// K1: one thread per row.
template <typename T, int S>
__global__ void gen_table_kernel(KerDev<T> k, const T *__restrict__ pos, int64_t nrows, int ncols,
                                 int32_t *__restrict__ J, T *__restrict__ W,
                                 int32_t *__restrict__ base, int *__restrict__ err)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    T w[S];
    int b;
    if (!getcoefs<T, S>(k, ncols, pos[i], b, w)) {
        atomicExch(err, 1);
        b = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) w[s] = (T)0;
    }
    base[i] = b;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        int j = b + s;
        J[i * S + s] = j < 0 ? 0 : (j > ncols - 1 ? ncols - 1 : j);
        W[i * S + s] = w[s];
    }
}

template <typename T, int S>
static int launch_gen(const KerDev<T> &k, const void *pos, int64_t nrows, int64_t ncols,
                      li_table *t, int *err, cudaStream_t st)
{
    const int threads = 128;
    const unsigned blocks = (unsigned)((nrows + threads - 1) / threads);
    gen_table_kernel<T, S><<<blocks, threads, 0, st>>>(k, (const T *)pos, nrows, (int)ncols, t->J,
                                                       (T *)t->W, t->base, err);
    LI_CHECK_LAUNCH();
    return LI_OK;
}

template <typename T>
static int dispatch_gen(const li_kernel_desc *ker, const void *pos, int64_t nrows, int64_t ncols,
                        li_table *t, int *err, cudaStream_t st)
{
    KerDev<T> k;
    int rc = make_kerdev<T>(ker, ncols, &k);
    if (rc != LI_OK) return rc;
    switch (ker->S) {
#define LI_CASE(SS) \
    case SS:        \
        return launch_gen<T, SS>(k, pos, nrows, ncols, t, err, st);
        LI_CASE(1)
        LI_CASE(2)
        LI_CASE(3)
        LI_CASE(4)
        LI_CASE(6)
        LI_CASE(8)
        LI_CASE(10)
        LI_CASE(12)
        LI_CASE(14)
        LI_CASE(16)
#undef LI_CASE
    default:
        set_error("unsupported kernel size S = %d", ker->S);
        return LI_ERR_UNSUPPORTED;
    }
}

// ---------------------------------------------------------------------------
// Analysis kernels.

// monotone <=> base[i] <= base[i+1] for all i
// flag[0] <- 0 unless monotone; flag[1] <- max(base[i+1] - base[i])
__global__ void monotone_kernel(const int32_t *__restrict__ base, int64_t nrows, int *flag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 < nrows) {
        const int d = base[i + 1] - base[i];
        if (d < 0) atomicExch(flag, 0);
        if (d > 1) atomicMax(flag + 1, d);
    }
}

// For imported tables: recover the canonical window base of every row, or flag the table
// as not windowed.  Also performs the bound check of __fullcheck (src/tabulated.jl:206-215).
__global__ void analyse_import_kernel(const int64_t *__restrict__ J1, int S, int64_t nrows, int ncols,
                                      int32_t *__restrict__ J, int32_t *__restrict__ base,
                                      int *__restrict__ flags /* [0]=windowed, [1]=out of bounds */)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    bool oob = false;
    for (int s = 0; s < S; ++s) {
        int64_t j = J1[i * S + s];
        if (j < 1 || j > ncols) {
            oob = true;
            j = 1;
        }
        J[i * S + s] = (int32_t)(j - 1);
    }
    if (oob) {
        atomicExch(&flags[1], 1);
        base[i] = 0;
        return;
    }
    int b;
    if (J[i * S] > 0) {
        b = J[i * S];
    } else {
        b = -(S - 1);
        for (int s = 1; s < S; ++s) {
            if (J[i * S + s] > 0) {
                b = J[i * S + s] - s;
                break;
            }
        }
    }
    bool ok = true;
    for (int s = 0; s < S; ++s) {
        int j = b + s;
        j = j < 0 ? 0 : (j > ncols - 1 ? ncols - 1 : j);
        ok = ok && (j == J[i * S + s]);
    }
    if (!ok) atomicExch(&flags[0], 0);
    base[i] = b;
}

__global__ void export_kernel(const int32_t *__restrict__ J, int64_t n, int64_t *__restrict__ J1)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) J1[i] = (int64_t)J[i] + 1;
}

// max over aligned tiles of G rows of (max base - min base)
__global__ void fwd_span_kernel(const int32_t *__restrict__ base, int64_t nrows, int G, int *out)
{
    int64_t r0 = (int64_t)blockIdx.x * G;
    int lo = INT_MAX, hi = INT_MIN;
    for (int64_t r = r0 + threadIdx.x; r < r0 + G && r < nrows; r += blockDim.x) {
        int b = base[r];
        lo = min(lo, b);
        hi = max(hi, b);
    }
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && hi >= lo) atomicMax(out, hi - lo);
}

// bptr[j] = first row i with base[i] >= j - (S-1), j in [0, ncols + S - 1]  (monotone base; canonical bases
// lie in [-(S-1), ncols-1], so bptr[0] = 0 and bptr[ncols+S-1] = nrows).
//   rows whose window base equals v      : [bptr[v+S-1], bptr[v+S])
//   rows contributing to column k        : [bptr[k], bptr[k+S])   (clamped taps included for k = 0, ncols-1)
__global__ void bptr_kernel(const int32_t *__restrict__ base, int64_t nrows, int S, int ncols,
                            int32_t *__restrict__ bptr)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > ncols + S - 1) return;
    const int target = j - (S - 1);
    int64_t lo = 0, hi = nrows;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (base[mid] >= target)
            hi = mid;
        else
            lo = mid + 1;
    }
    bptr[j] = (int32_t)lo;
}

// number of rows feeding an aligned tile of G columns
__global__ void adj_rows_kernel(const int32_t *__restrict__ bptr, int S, int ncols, int G, int *out)
{
    int tile = blockIdx.x * blockDim.x + threadIdx.x;
    int k0 = tile * G;
    if (k0 >= ncols) return;
    int k1 = min(k0 + G, ncols);  // exclusive
    atomicMax(out, (int)(bptr[k1 - 1 + S] - bptr[k0]));
}

// rows [lo, hi) with a tap landing in an aligned tile of G columns, clamped taps included
__global__ void feed_rows_kernel(const int32_t *__restrict__ rowstart, int S, int ncols, int nrows, int G, int *out)
{
    int tile = blockIdx.x * blockDim.x + threadIdx.x;
    int k0 = tile * G;
    if (k0 >= ncols) return;
    int k1 = min(k0 + G, ncols);  // exclusive
    const int lo = k0 == 0 ? 0 : rowstart[k0];
    const int hi = k1 == ncols ? nrows : rowstart[k1 - 1 + S];
    atomicMax(out, hi - lo);
}

// Fold the clamped taps of a window into the window shifted inside [0, ncols): bc = clamp(base, 0, ncols - S),
// tap s lands in slot clamp(base + s, 0, ncols-1) - bc.  (Sums of weights that hit the same column.)
template <typename T>
__global__ void fold_kernel(const int32_t *__restrict__ base, const T *__restrict__ W, int64_t nrows, int S, int ncols,
                            int32_t *__restrict__ bfold, T *__restrict__ Wfold)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const int b = base[i];
    const int bc = min(max(b, 0), ncols - S);
    T w[LI_MAX_S];
    for (int s = 0; s < S; ++s) w[s] = (T)0;
    for (int s = 0; s < S; ++s) w[min(max(b + s, 0), ncols - 1) - bc] += W[i * S + s];
    bfold[i] = bc;
    for (int s = 0; s < S; ++s) Wfold[i * S + s] = w[s];
}

__global__ void iota_keys_kernel(const int32_t *__restrict__ J, int64_t n, int32_t *__restrict__ keys,
                                 int32_t *__restrict__ vals)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) {
        keys[e] = J[e];
        vals[e] = (int32_t)e;
    }
}

__global__ void colptr_kernel(const int32_t *__restrict__ sorted_keys, int64_t n, int ncols,
                              int32_t *__restrict__ colptr)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > ncols) return;
    int64_t lo = 0, hi = n;  // first e with key >= k
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] >= k)
            hi = mid;
        else
            lo = mid + 1;
    }
    colptr[k] = (int32_t)lo;
}

// ---------------------------------------------------------------------------
static int alloc_table(int dtype, int S, int64_t nrows, int64_t ncols, li_table **out)
{
    LI_REQUIRE(nrows >= 1, LI_ERR_ARG, "number of rows too small");      // src/tabulated.jl:83,95
    LI_REQUIRE(ncols >= 1, LI_ERR_ARG, "number of columns too small");   // src/tabulated.jl:84,96
    LI_REQUIRE(ncols <= 0x7fffffffLL && nrows * S <= 0x7fffffffLL, LI_ERR_UNSUPPORTED,
               "table too large for 32-bit device indices");
    static std::atomic<uint64_t> next_uid{1};
    li_table *t = new li_table();
    t->uid = next_uid.fetch_add(1);
    t->dtype = dtype;
    t->S = S;
    t->nrows = nrows;
    t->ncols = ncols;
    cudaError_t e;
    if ((e = cudaMalloc(&t->J, (size_t)nrows * S * sizeof(int32_t))) != cudaSuccess ||
        (e = cudaMalloc(&t->W, (size_t)nrows * S * dtype_size(dtype))) != cudaSuccess ||
        (e = cudaMalloc(&t->base, (size_t)nrows * sizeof(int32_t))) != cudaSuccess) {
        set_error("cudaMalloc failed: %s", cudaGetErrorString(e));
        li_table_destroy(t);
        return e == cudaErrorMemoryAllocation ? LI_ERR_NOMEM : LI_ERR_CUDA;
    }
    *out = t;
    return LI_OK;
}

int table_create_impl(const li_kernel_desc *ker, const void *pos_dev, int64_t nrows, int64_t ncols,
                      cudaStream_t st, li_table **out)
{
    LI_REQUIRE(ker && pos_dev && out, LI_ERR_ARG, "null argument");
    LI_REQUIRE(ker->dtype == LI_F32 || ker->dtype == LI_F64, LI_ERR_ARG, "bad dtype");
    li_table *t = nullptr;
    int rc = alloc_table(ker->dtype, ker->S, nrows, ncols, &t);
    if (rc != LI_OK) return rc;
    int *flags = nullptr;  // [0] inexact, [1] monotone, [2] max step of the window base
    int h_flags[3] = {0, 1, 1};
    cudaError_t e = cudaMalloc(&flags, sizeof h_flags);
    if (e == cudaSuccess) e = cudaMemcpyAsync(flags, h_flags, sizeof h_flags, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
        set_error("cudaMalloc/cudaMemcpy failed: %s", cudaGetErrorString(e));
        li_table_destroy(t);
        return LI_ERR_CUDA;
    }
    rc = ker->dtype == LI_F32 ? dispatch_gen<float>(ker, pos_dev, nrows, ncols, t, flags, st)
                              : dispatch_gen<double>(ker, pos_dev, nrows, ncols, t, flags, st);
    if (rc == LI_OK) {
        monotone_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(t->base, nrows, flags + 1);
        count_launch();
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_flags, flags, sizeof h_flags, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_error("table generation failed: %s", cudaGetErrorString(e));
            rc = LI_ERR_CUDA;
        }
    }
    cudaFree(flags);
    if (rc == LI_OK && h_flags[0]) {
        set_error("InexactError: trunc(Int64, f) for a NaN or out-of-range position");
        rc = LI_ERR_INEXACT;
    }
    if (rc != LI_OK) {
        li_table_destroy(t);
        return rc;
    }
    t->windowed = true;
    t->monotone = h_flags[1] != 0;
    t->max_step = h_flags[2];
    *out = t;
    return LI_OK;
}

__global__ void heavy_columns_kernel(const int32_t *__restrict__ colptr, int ncols, int threshold, int32_t *count, int32_t *list)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < ncols && colptr[j + 1] - colptr[j] > threshold) list[atomicAdd(count, 1)] = j;   // (list order is irrelevant)
}

int table_ensure_csc(const li_table *t, cudaStream_t st)
{
    std::lock_guard<std::mutex> lock(t->mu);
    if (t->colptr) return LI_OK;
    const int64_t n = t->nrows * t->S;
    int32_t *keys = nullptr, *vals = nullptr, *keys_out = nullptr, *vals_out = nullptr, *colptr = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    int rc = LI_OK;
    auto cleanup = [&]() {
        cudaFree(keys);
        cudaFree(vals);
        cudaFree(keys_out);
        cudaFree(tmp);
    };
#define LI_TRY(expr)                                                               \
    do {                                                                           \
        cudaError_t e__ = (expr);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            set_error("%s failed: %s", #expr, cudaGetErrorString(e__));            \
            cleanup();                                                             \
            cudaFree(vals_out);                                                    \
            cudaFree(colptr);                                                      \
            return LI_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)
    LI_TRY(cudaMalloc(&keys, n * sizeof(int32_t)));
    LI_TRY(cudaMalloc(&vals, n * sizeof(int32_t)));
    LI_TRY(cudaMalloc(&keys_out, n * sizeof(int32_t)));
    LI_TRY(cudaMalloc(&vals_out, n * sizeof(int32_t)));
    LI_TRY(cudaMalloc(&colptr, (t->ncols + 1) * sizeof(int32_t)));
    iota_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t->J, n, keys, vals);
    count_launch();
    int end_bit = 1;
    while (end_bit < 31 && (1LL << end_bit) < t->ncols) ++end_bit;
    // stable LSD radix sort: within a column the entries stay in ascending (row, tap) order,
    // i.e. the order of the reference's sequential scatter (src/tabulated.jl:455-462)
    LI_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, vals, vals_out, (int)n, 0,
                                           end_bit, st));
    LI_TRY(cudaMalloc(&tmp, tmp_bytes));
    LI_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, vals, vals_out, (int)n, 0,
                                           end_bit, st));
    count_launch(4);
    colptr_kernel<<<(unsigned)((t->ncols + 1 + 255) / 256), 256, 0, st>>>(keys_out, n, (int)t->ncols, colptr);
    count_launch();
    LI_TRY(cudaGetLastError());
    // heavy columns (more than LI_HEAVY_COLUMN entries): listed once, for the segmented sums of the gather adjoint
    int32_t *heavy = nullptr;
    int nheavy = 0;
    const int64_t cap = n / LI_HEAVY_COLUMN + 1;
    LI_TRY(cudaMalloc(&heavy, (cap + 1) * sizeof(int32_t)));
    LI_TRY(cudaMemsetAsync(heavy, 0, sizeof(int32_t), st));
    heavy_columns_kernel<<<(unsigned)((t->ncols + 255) / 256), 256, 0, st>>>(colptr, (int)t->ncols, LI_HEAVY_COLUMN, heavy, heavy + 1);
    count_launch();
    LI_TRY(cudaMemcpyAsync(&nheavy, heavy, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    LI_TRY(cudaStreamSynchronize(st));
#undef LI_TRY
    cleanup();
    t->tent = vals_out;
    t->colptr = colptr;
    t->heavy = heavy;        // ([0] is the counter, the list follows)
    t->nheavy = nheavy;
    return rc;
}

int table_ensure_rowstart(const li_table *t, cudaStream_t st)
{
    std::lock_guard<std::mutex> lock(t->mu);
    if (t->rowstart) return LI_OK;
    LI_REQUIRE(t->windowed && t->monotone, LI_ERR_ARG, "rowstart needs a monotone windowed table");
    int32_t *rs = nullptr;
    LI_CUDA(cudaMalloc(&rs, (t->ncols + t->S) * sizeof(int32_t)));
    bptr_kernel<<<(unsigned)((t->ncols + t->S + 255) / 256), 256, 0, st>>>(t->base, t->nrows, t->S, (int)t->ncols, rs);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFree(rs);
        set_error("bptr kernel failed: %s", cudaGetErrorString(e));
        return LI_ERR_CUDA;
    }
    t->rowstart = rs;
    return LI_OK;
}

// irr[v] = 1 if the number of rows whose folded window starts at column v differs from R (inclusive prefix sums
// are taken afterwards): tile rows with irr-free ranges take the statically scheduled adjoint kernel
__global__ void irregular_kernel(const int32_t *__restrict__ bpf, int S, int ncols, int R, int32_t *__restrict__ irr)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < ncols) irr[v] = (bpf[v + S] - bpf[v + S - 1]) != R;
}

int table_ensure_fold(const li_table *t, cudaStream_t st)
{
    std::lock_guard<std::mutex> lock(t->mu);
    if (t->bpf) return LI_OK;
    LI_REQUIRE(t->windowed && t->monotone && t->ncols >= t->S, LI_ERR_ARG, "folding needs a monotone windowed table with ncols >= S");
    void *wf = nullptr;
    int32_t *bf = nullptr, *bp = nullptr;
    cudaError_t e = cudaMalloc(&wf, (size_t)t->nrows * t->S * dtype_size(t->dtype));
    if (e == cudaSuccess) e = cudaMalloc(&bf, (size_t)t->nrows * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&bp, (size_t)(t->ncols + t->S) * sizeof(int32_t));
    if (e == cudaSuccess) {
        const unsigned blocks = (unsigned)((t->nrows + 127) / 128);
        if (t->dtype == LI_F32)
            fold_kernel<float><<<blocks, 128, 0, st>>>(t->base, (const float *)t->W, t->nrows, t->S, (int)t->ncols, bf, (float *)wf);
        else
            fold_kernel<double><<<blocks, 128, 0, st>>>(t->base, (const double *)t->W, t->nrows, t->S, (int)t->ncols, bf, (double *)wf);
        bptr_kernel<<<(unsigned)((t->ncols + t->S + 255) / 256), 256, 0, st>>>(bf, t->nrows, t->S, (int)t->ncols, bp);
        count_launch(2);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFree(wf);
        cudaFree(bf);
        cudaFree(bp);
        set_error("fold kernel failed: %s", cudaGetErrorString(e));
        return LI_ERR_CUDA;
    }
    // regularity: R src rows per dst row (R = 1: same-rate resampling, R = 2: 2x up-sampling)
    const int R = (int)((t->nrows + t->ncols / 2) / t->ncols);
    int32_t *irr = nullptr, *irrp = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    if (R >= 1 && R <= 2) {
        e = cudaMalloc(&irr, (size_t)t->ncols * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMalloc(&irrp, (size_t)(t->ncols + 1) * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMemsetAsync(irrp, 0, sizeof(int32_t), st);
        if (e == cudaSuccess) {
            irregular_kernel<<<(unsigned)((t->ncols + 255) / 256), 256, 0, st>>>(bp, t->S, (int)t->ncols, R, irr);
            count_launch();
            e = cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, irr, irrp + 1, (int)t->ncols, st);
        }
        if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes);
        if (e == cudaSuccess) e = cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, irr, irrp + 1, (int)t->ncols, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(irr);
        cudaFree(tmp);
        if (e != cudaSuccess) {
            cudaFree(irrp);
            irrp = nullptr;
        }
    }
    t->Wfold = wf;
    t->bfold = bf;
    t->bpf = bp;
    t->irrp = irrp;
    t->reg_R = irrp ? R : 0;
    return LI_OK;
}

static int reduce_to_host(cudaStream_t st, int init, int *result,
                          const std::function<void(int *)> &launch)
{
    int *d = nullptr;
    LI_CUDA(cudaMalloc(&d, sizeof(int)));
    cudaError_t e = cudaMemcpyAsync(d, &init, sizeof(int), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        launch(d);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(result, d, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) {
        set_error("table analysis failed: %s", cudaGetErrorString(e));
        return LI_ERR_CUDA;
    }
    return LI_OK;
}

int table_fwd_span(const li_table *t, int G, cudaStream_t st, int *span)
{
    {
        std::lock_guard<std::mutex> lock(t->mu);
        auto it = t->fwd_span.find(G);
        if (it != t->fwd_span.end()) {
            *span = it->second;
            return LI_OK;
        }
    }
    LI_REQUIRE(t->windowed, LI_ERR_ARG, "footprints need a windowed table");
    int v = 0;
    int rc = reduce_to_host(st, 0, &v, [&](int *d) {
        fwd_span_kernel<<<(unsigned)((t->nrows + G - 1) / G), 32, 0, st>>>(t->base, t->nrows, G, d);
    });
    if (rc != LI_OK) return rc;
    std::lock_guard<std::mutex> lock(t->mu);
    t->fwd_span[G] = v;
    *span = v;
    return LI_OK;
}

int table_adj_rows(const li_table *t, int G, cudaStream_t st, int *rows)
{
    {
        std::lock_guard<std::mutex> lock(t->mu);
        auto it = t->adj_rows.find(G);
        if (it != t->adj_rows.end()) {
            *rows = it->second;
            return LI_OK;
        }
    }
    int rc = table_ensure_rowstart(t, st);
    if (rc != LI_OK) return rc;
    int v = 0;
    const int ntiles = (int)((t->ncols + G - 1) / G);
    rc = reduce_to_host(st, 0, &v, [&](int *d) {
        adj_rows_kernel<<<(ntiles + 127) / 128, 128, 0, st>>>(t->rowstart, t->S, (int)t->ncols, G, d);
    });
    if (rc != LI_OK) return rc;
    std::lock_guard<std::mutex> lock(t->mu);
    t->adj_rows[G] = v;
    *rows = v;
    return LI_OK;
}

int table_feed_rows(const li_table *t, int G, cudaStream_t st, int *rows)
{
    {
        std::lock_guard<std::mutex> lock(t->mu);
        auto it = t->feed_rows.find(G);
        if (it != t->feed_rows.end()) {
            *rows = it->second;
            return LI_OK;
        }
    }
    int rc = table_ensure_rowstart(t, st);
    if (rc != LI_OK) return rc;
    int v = 0;
    const int ntiles = (int)((t->ncols + G - 1) / G);
    rc = reduce_to_host(st, 0, &v, [&](int *d) {
        feed_rows_kernel<<<(ntiles + 127) / 128, 128, 0, st>>>(t->rowstart, t->S, (int)t->ncols, (int)t->nrows, G, d);
    });
    if (rc != LI_OK) return rc;
    std::lock_guard<std::mutex> lock(t->mu);
    t->feed_rows[G] = v;
    *rows = v;
    return LI_OK;
}

int table_adjf_rows(const li_table *t, int G, cudaStream_t st, int *rows)
{
    {
        std::lock_guard<std::mutex> lock(t->mu);
        auto it = t->adjf_rows.find(G);
        if (it != t->adjf_rows.end()) {
            *rows = it->second;
            return LI_OK;
        }
    }
    int rc = table_ensure_fold(t, st);
    if (rc != LI_OK) return rc;
    int v = 0;
    const int ntiles = (int)((t->ncols + G - 1) / G);
    rc = reduce_to_host(st, 0, &v, [&](int *d) {
        adj_rows_kernel<<<(ntiles + 127) / 128, 128, 0, st>>>(t->bpf, t->S, (int)t->ncols, G, d);
    });
    if (rc != LI_OK) return rc;
    std::lock_guard<std::mutex> lock(t->mu);
    t->adjf_rows[G] = v;
    *rows = v;
    return LI_OK;
}

int table_tile_rows(const li_table *t, int TD2, cudaStream_t st, const int32_t **list, int *nregular)
{
    int rc = table_ensure_fold(t, st);
    if (rc != LI_OK) return rc;
    std::lock_guard<std::mutex> lock(t->mu);
    auto it = t->tile_rows.find(TD2);
    if (it == t->tile_rows.end()) {
        const int ntiles = (int)((t->ncols + TD2 - 1) / TD2);
        std::vector<int32_t> irrp(t->ncols + 1, 0), order;
        if (t->irrp) {
            LI_CUDA(cudaMemcpyAsync(irrp.data(), t->irrp, (t->ncols + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            LI_CUDA(cudaStreamSynchronize(st));
        }
        std::vector<int32_t> irregular;
        for (int k = 0; k < ntiles; ++k) {
            const int d0 = k * TD2, d1 = d0 + TD2;   // only full tile rows can be regular
            const bool regular = t->irrp && d1 <= t->ncols && d0 >= t->S - 1 && irrp[d1] == irrp[d0 - (t->S - 1)];
            (regular ? order : irregular).push_back(k);
        }
        const int nreg = (int)order.size();
        order.insert(order.end(), irregular.begin(), irregular.end());
        int32_t *d = nullptr;
        LI_CUDA(cudaMalloc(&d, ntiles * sizeof(int32_t)));
        cudaError_t e = cudaMemcpyAsync(d, order.data(), ntiles * sizeof(int32_t), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            cudaFree(d);
            set_error("tile list upload failed: %s", cudaGetErrorString(e));
            return LI_ERR_CUDA;
        }
        it = t->tile_rows.emplace(TD2, std::make_pair(d, nreg)).first;
    }
    *list = it->second.first;
    *nregular = it->second.second;
    return LI_OK;
}

// Identity operator as a table (S = 1, row i reads column i with weight 1): the second factor that lets a 1-D apply
// along the contiguous dimension run through the fused 2-D kernels.  Cached per (device, dtype, n), never freed.
template <typename T>
__global__ void identity_table_kernel(int32_t *__restrict__ J, T *__restrict__ W, int32_t *__restrict__ base, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        J[i] = i;
        base[i] = i;
        W[i] = (T)1;
    }
}

int table_identity(int dtype, int64_t n, cudaStream_t st, const li_table **out)
{
    static std::mutex mu;
    static std::map<std::tuple<int, int, int64_t>, li_table *> cache;
    int dev = 0;
    LI_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple(dev, dtype, n);
    auto it = cache.find(key);
    if (it == cache.end()) {
        li_table *t = nullptr;
        int rc = alloc_table(dtype, 1, n, n, &t);
        if (rc != LI_OK) return rc;
        const unsigned blocks = (unsigned)((n + 255) / 256);
        if (dtype == LI_F32)
            identity_table_kernel<float><<<blocks, 256, 0, st>>>(t->J, (float *)t->W, t->base, (int)n);
        else
            identity_table_kernel<double><<<blocks, 256, 0, st>>>(t->J, (double *)t->W, t->base, (int)n);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_error("identity table failed: %s", cudaGetErrorString(e));
            li_table_destroy(t);
            return LI_ERR_CUDA;
        }
        t->windowed = t->monotone = true;
        t->max_step = 1;
        it = cache.emplace(key, t).first;
    }
    *out = it->second;
    return LI_OK;
}

int table_base_host(const li_table *t, cudaStream_t st, const std::vector<int32_t> **base)
{
    std::lock_guard<std::mutex> lock(t->mu);
    if (t->base_host.empty()) {
        t->base_host.resize(t->nrows);
        LI_CUDA(cudaMemcpyAsync(t->base_host.data(), t->base, t->nrows * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        LI_CUDA(cudaStreamSynchronize(st));
    }
    *base = &t->base_host;
    return LI_OK;
}

int table_chunk_plan(const li_table *t, int nchunks, double penalty, cudaStream_t st, const int32_t **starts, int *crmax,
                     int *nrmax, const li_table::ChunkPlan **plan_out, const std::vector<double> *weights, int weights_key)
{
    int rc = table_ensure_fold(t, st);
    if (rc != LI_OK) return rc;
    std::lock_guard<std::mutex> lock(t->mu);
    const int key = weights ? (nchunks | (weights_key << 12)) : nchunks;   // (nchunks < 4096: CR >= 8 rows)
    auto it = t->chunk_plans.find(key);
    if (it == t->chunk_plans.end()) {
        const int m = (int)t->ncols, S = t->S;
        std::vector<int32_t> bpf(m + S), irrp(m + 1, 0);
        LI_CUDA(cudaMemcpyAsync(bpf.data(), t->bpf, (m + S) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (t->irrp) LI_CUDA(cudaMemcpyAsync(irrp.data(), t->irrp, (m + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        LI_CUDA(cudaStreamSynchronize(st));
        // cost of a column: 1, plus `penalty` chunks' worth for the first column of every run of columns that force
        // their chunk onto the generic loop (not fed by exactly reg_R rows, or next to the array edges)
        const double base = (double)m / nchunks;
        std::vector<double> cost(m + 1, 0.0);
        double total = 0.0;
        bool prev = false;
        for (int v = 0; v < m; ++v) {
            const bool odd = !t->irrp || irrp[v + 1] != irrp[v] || v < S || v >= m - S;
            total += 1.0 + ((odd && !prev) ? penalty * base : 0.0);
            cost[v + 1] = total;
            prev = odd;
        }
        // share of chunk k: equal, or proportional to weights[k] (CTAs that the hardware favours get longer chunks)
        std::vector<double> cum(nchunks + 1, 0.0);
        for (int k = 0; k < nchunks; ++k) cum[k + 1] = cum[k] + (weights ? (*weights)[k] : 1.0);
        std::vector<int32_t> b(nchunks + 1, 0);
        int v = 0;
        for (int k = 1; k < nchunks; ++k) {
            const double want = total * cum[k] / cum[nchunks];
            while (v < m && cost[v] < want) ++v;
            b[k] = std::max(v, b[k - 1]);
        }
        b[nchunks] = m;
        li_table::ChunkPlan plan;
        plan.starts_host = b;
        plan.src_end_host.assign(nchunks, 0);
        for (int k = 0; k < nchunks; ++k) {
            const int va = b[k], vb = b[k + 1];
            plan.crmax = std::max(plan.crmax, vb - va);
            if (vb > va) plan.nrmax = std::max(plan.nrmax, (int)(bpf[vb - 1 + S] - bpf[va]));
            plan.src_end_host[k] = vb > va ? bpf[vb - 1 + S] : (k ? plan.src_end_host[k - 1] : 0);
        }
        LI_CUDA(cudaMalloc(&plan.starts, (nchunks + 1) * sizeof(int32_t)));
        cudaError_t e = cudaMemcpyAsync(plan.starts, b.data(), (nchunks + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            cudaFree(plan.starts);
            set_error("chunk plan upload failed: %s", cudaGetErrorString(e));
            return LI_ERR_CUDA;
        }
        it = t->chunk_plans.emplace(key, plan).first;
    }
    *starts = it->second.starts;
    *crmax = it->second.crmax;
    *nrmax = it->second.nrmax;
    if (plan_out) *plan_out = &it->second;
    return LI_OK;
}

}  // namespace li

// ---------------------------------------------------------------------------
extern "C" {

int li_table_create(const li_kernel_desc *ker, const void *pos, int pos_on_device, int64_t nrows,
                    int64_t ncols, void *stream, li_table **out)
{
    LI_REQUIRE(ker && pos && out, LI_ERR_ARG, "null argument");
    LI_REQUIRE(nrows >= 1, LI_ERR_ARG, "number of positions too small");
    cudaStream_t st = li::as_stream(stream);
    if (pos_on_device) return li::table_create_impl(ker, pos, nrows, ncols, st, out);
    void *d = nullptr;
    const size_t bytes = (size_t)nrows * li::dtype_size(ker->dtype);
    LI_CUDA(cudaMalloc(&d, bytes));
    cudaError_t e = cudaMemcpyAsync(d, pos, bytes, cudaMemcpyHostToDevice, st);
    int rc = LI_OK;
    if (e != cudaSuccess) {
        li::set_error("cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
        rc = LI_ERR_CUDA;
    } else {
        rc = li::table_create_impl(ker, d, nrows, ncols, st, out);
    }
    cudaStreamSynchronize(st);
    cudaFree(d);
    return rc;
}

int li_table_import(int dtype, int S, int64_t nrows, int64_t ncols, const int64_t *J_host,
                    const void *W_host, void *stream, li_table **out)
{
    LI_REQUIRE(J_host && W_host && out, LI_ERR_ARG, "null argument");
    LI_REQUIRE(dtype == LI_F32 || dtype == LI_F64, LI_ERR_ARG, "bad dtype");
    LI_REQUIRE(S >= 1 && S <= LI_MAX_S, LI_ERR_UNSUPPORTED, "unsupported kernel size S = %d", S);
    cudaStream_t st = li::as_stream(stream);
    li_table *t = nullptr;
    int rc = li::alloc_table(dtype, S, nrows, ncols, &t);
    if (rc != LI_OK) return rc;
    const int64_t n = nrows * S;
    int64_t *dJ = nullptr;
    int *flags = nullptr;
    int h_flags[4] = {1, 0, 1, 1};  // windowed, out-of-bounds, monotone, max step
    cudaError_t e = cudaMalloc(&dJ, n * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMalloc(&flags, sizeof h_flags);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dJ, J_host, n * sizeof(int64_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t->W, W_host, n * li::dtype_size(dtype), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(flags, h_flags, sizeof h_flags, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        li::analyse_import_kernel<<<(unsigned)((nrows + 127) / 128), 128, 0, st>>>(dJ, S, nrows, (int)ncols, t->J,
                                                                                 t->base, flags);
        li::monotone_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(t->base, nrows, flags + 2);
        li::count_launch(2);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_flags, flags, sizeof h_flags, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(dJ);
    cudaFree(flags);
    if (e != cudaSuccess) {
        li::set_error("table import failed: %s", cudaGetErrorString(e));
        li_table_destroy(t);
        return LI_ERR_CUDA;
    }
    if (h_flags[1]) {
        li::set_error("out of bound index in interpolation table");  // src/tabulated.jl:212
        li_table_destroy(t);
        return LI_ERR_BOUNDS;
    }
    t->windowed = h_flags[0] != 0;
    t->monotone = t->windowed && h_flags[2] != 0;
    t->max_step = h_flags[3];
    *out = t;
    return LI_OK;
}

int li_table_export(const li_table *t, int64_t *J_host, void *W_host, void *stream)
{
    LI_REQUIRE(t, LI_ERR_ARG, "null table");
    cudaStream_t st = li::as_stream(stream);
    const int64_t n = t->nrows * t->S;
    if (J_host) {
        int64_t *dJ = nullptr;
        LI_CUDA(cudaMalloc(&dJ, n * sizeof(int64_t)));
        li::export_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t->J, n, dJ);
        li::count_launch();
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(J_host, dJ, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(dJ);
        if (e != cudaSuccess) {
            li::set_error("table export failed: %s", cudaGetErrorString(e));
            return LI_ERR_CUDA;
        }
    }
    if (W_host) {
        LI_CUDA(cudaMemcpyAsync(W_host, t->W, n * li::dtype_size(t->dtype), cudaMemcpyDeviceToHost, st));
        LI_CUDA(cudaStreamSynchronize(st));
    }
    return LI_OK;
}

int li_table_get_info(const li_table *t, li_table_info *info)
{
    LI_REQUIRE(t && info, LI_ERR_ARG, "null argument");
    info->dtype = t->dtype;
    info->S = t->S;
    info->nrows = t->nrows;
    info->ncols = t->ncols;
    info->windowed = t->windowed;
    info->monotone = t->monotone;
    return LI_OK;
}

int li_table_destroy(li_table *t)
{
    if (!t) return LI_OK;
    cudaFree(t->J);
    cudaFree(t->W);
    cudaFree(t->base);
    cudaFree(t->colptr);
    cudaFree(t->tent);
    cudaFree(t->heavy);
    cudaFree(t->rowstart);
    cudaFree(t->Wfold);
    cudaFree(t->bfold);
    cudaFree(t->bpf);
    cudaFree(t->irrp);
    for (auto &kv : t->tile_rows) cudaFree(kv.second.first);
    for (auto &kv : t->chunk_plans) cudaFree(kv.second.starts);
    delete t;
    return LI_OK;
}

}  // extern "C"
