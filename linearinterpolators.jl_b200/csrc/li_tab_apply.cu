// li_tab_apply.cu -- table-backed operators along one dimension (K3-K6, generic form).
//
// Direct  : dst[p,i,q] = alpha * sum_s W[s,i] * src[p, J[s,i], q] + beta * dst[p,i,q]
//           reference __direct! (src/tabulated.jl:394-439), SparseInterpolator Direct
//           (src/sparse.jl:201-244), _apply_direct! (src/sparse.jl:719-774)
// Adjoint : dst[p,j,q] = beta * dst[p,j,q] + sum_{(i,s): J[s,i]=j} W[s,i] * (alpha * src[p,i,q])
//           reference __adjoint! (src/tabulated.jl:441-478), src/sparse.jl:246-275, :776-800
//
// These kernels work for ANY table (arbitrary J, e.g. SparseInterpolator with N-d positions):
// they gather through L1/L2.  Windowed tables of separable operators take the fused, staged
// kernels of li_sep2d.cu instead.  Threads always run along the contiguous dimension (the
// opposite of the reference's CPU loop nest, which puts the interpolated index innermost).
#include <cstdlib>

#include "li_common.cuh"

namespace li {

template <typename T>
__device__ __forceinline__ T fma_(T a, T b, T c);
template <>
__device__ __forceinline__ float fma_<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <>
__device__ __forceinline__ double fma_<double>(double a, double b, double c) { return fma(a, b, c); }

template <typename T>
__device__ __forceinline__ T finish(T a, T alpha, T beta, const T *dst)
{
    // beta == 0: dst is never read (src/tabulated.jl:408-425)
    if (beta == (T)0) return alpha == (T)1 ? a : alpha * a;
    return fma_(alpha, a, beta * (*dst));
}

// ---- Direct, interpolated dimension is the contiguous one (npre == 1) ------
// thread <-> output row i; J/W of the row are loaded once and reused for QB planes q.
template <typename T, int SC>
__global__ void __launch_bounds__(256)
tab_direct_d1_kernel(const int32_t *__restrict__ J, const T *__restrict__ W, int Srt, int nrows, int ncols,
                     int64_t npost, T alpha, T beta, const T *__restrict__ src, T *__restrict__ dst)
{
    const int S = SC ? SC : Srt;
    constexpr int SMAX = SC ? SC : LI_MAX_S;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    int j[SMAX];
    T w[SMAX];
#pragma unroll
    for (int s = 0; s < SMAX; ++s) {
        if (s < S) {
            j[s] = J[(int64_t)i * S + s];
            w[s] = W[(int64_t)i * S + s];
        }
    }
    for (int64_t q = blockIdx.y; q < npost; q += gridDim.y) {
        const T *sp = src + q * ncols;
        T a = (T)0;
#pragma unroll
        for (int s = 0; s < SMAX; ++s)
            if (s < S) a = fma_(sp[j[s]], w[s], a);
        T *d = dst + q * nrows + i;
        *d = finish(a, alpha, beta, d);
    }
}

// ---- Direct, strided dimension (npre > 1) -----------------------------------
// thread <-> p (contiguous); a block handles RB consecutive rows i; J/W are warp-uniform.
template <typename T, int SC, int RB>
__global__ void __launch_bounds__(256)
tab_direct_dn_kernel(const int32_t *__restrict__ J, const T *__restrict__ W, int Srt, int nrows, int ncols,
                     int64_t npre, int64_t npost, T alpha, T beta, const T *__restrict__ src,
                     T *__restrict__ dst)
{
    const int S = SC ? SC : Srt;
    constexpr int SMAX = SC ? SC : LI_MAX_S;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npre) return;
    for (int i0 = blockIdx.y * RB; i0 < nrows; i0 += gridDim.y * RB)
    for (int64_t q = blockIdx.z; q < npost; q += gridDim.z) {
        const T *sp = src + p + q * npre * ncols;
        T *dp = dst + p + q * npre * nrows;
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const int i = i0 + r;
            if (i < nrows) {
                T a = (T)0;
#pragma unroll
                for (int s = 0; s < SMAX; ++s)
                    if (s < S) a = fma_(sp[(int64_t)J[(int64_t)i * S + s] * npre], W[(int64_t)i * S + s], a);
                T *d = dp + (int64_t)i * npre;
                *d = finish(a, alpha, beta, d);
            }
        }
    }
}

// ---- Adjoint, deterministic gather over the transposed (CSC) table ------------
// Entries of a column are visited in ascending (row, tap) order == the order in which the
// reference's sequential scatter adds them, so the rounding sequence is the same.
template <typename T>
__global__ void __launch_bounds__(256)
tab_adjoint_gather_d1_kernel(const int32_t *__restrict__ colptr, const int32_t *__restrict__ tent,
                             const T *__restrict__ W, int S, int nrows, int ncols, int64_t npost, T alpha,
                             T beta, const T *__restrict__ src, T *__restrict__ dst)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    const int e0 = colptr[j], e1 = colptr[j + 1];
    if (e1 - e0 > LI_HEAVY_COLUMN) return;   // (a heavy column: tab_adjoint_heavy_d1_kernel)
    for (int64_t q = blockIdx.y; q < npost; q += gridDim.y) {
        const T *sp = src + q * nrows;
        T *d = dst + q * ncols + j;
        T acc = beta == (T)0 ? (T)0 : beta * (*d);
        for (int e = e0; e < e1; ++e) {
            const int ent = tent[e];
            const T x = alpha == (T)1 ? sp[ent / S] : alpha * sp[ent / S];
            acc = fma_(W[ent], x, acc);
        }
        *d = acc;
    }
}

// Heavy columns of the same operator (more than LI_HEAVY_COLUMN entries: with Flat clamping the first / last column
// collect every row hanging over that edge): one CTA per (heavy column, slice).  Thread t sums the t-th contiguous
// segment of the column's entries in ascending order, the 256 partial sums are combined by a fixed tree: deterministic
// (the order of the sums is fixed by the table alone), O(entries / 256 + 8) steps instead of O(entries) on one thread.
template <typename T>
__global__ void __launch_bounds__(256)
tab_adjoint_heavy_d1_kernel(const int32_t *__restrict__ heavy, const int32_t *__restrict__ colptr, const int32_t *__restrict__ tent,
                            const T *__restrict__ W, int S, int nrows, int ncols, int64_t npost, T alpha, T beta,
                            const T *__restrict__ src, T *__restrict__ dst)
{
    __shared__ T part[256];
    const int j = heavy[blockIdx.x];
    const int e0 = colptr[j], e1 = colptr[j + 1];
    const int seg = (e1 - e0 + 255) / 256;
    const int a = min(e0 + (int)threadIdx.x * seg, e1), b = min(a + seg, e1);
    for (int64_t q = blockIdx.y; q < npost; q += gridDim.y) {
        const T *sp = src + q * nrows;
        T acc = (T)0;
        for (int e = a; e < b; ++e) {
            const int ent = tent[e];
            const T x = alpha == (T)1 ? sp[ent / S] : alpha * sp[ent / S];
            acc = fma_(W[ent], x, acc);
        }
        part[threadIdx.x] = acc;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) part[threadIdx.x] = part[threadIdx.x] + part[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            T *d = dst + q * ncols + j;
            *d = beta == (T)0 ? part[0] : fma_(beta, *d, part[0]);
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
tab_adjoint_gather_dn_kernel(const int32_t *__restrict__ colptr, const int32_t *__restrict__ tent,
                             const T *__restrict__ W, int S, int nrows, int ncols, int64_t npre, int64_t npost,
                             T alpha, T beta, const T *__restrict__ src, T *__restrict__ dst)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npre) return;
    for (int j = blockIdx.y; j < ncols; j += gridDim.y) {
    const int e0 = colptr[j], e1 = colptr[j + 1];
    for (int64_t q = blockIdx.z; q < npost; q += gridDim.z) {
        const T *sp = src + p + q * npre * nrows;
        T *d = dst + p + ((int64_t)j + q * ncols) * npre;
        T acc = beta == (T)0 ? (T)0 : beta * (*d);
        for (int e = e0; e < e1; ++e) {
            const int ent = tent[e];
            const T v = sp[(int64_t)(ent / S) * npre];
            const T x = alpha == (T)1 ? v : alpha * v;
            acc = fma_(W[ent], x, acc);
        }
        *d = acc;
    }
    }
}

// ---- Streaming kernels for a strided dimension (npre > 1) of a windowed, monotone table ------------------------
// Along a dimension other than the first the threads share nothing: a thread owns V consecutive elements of the
// contiguous index (one 16-byte vector) and walks along the interpolated dimension with the window in registers,
// so every src element is read once per chunk (plus S-1 halo planes) and every dst element written once, with
// 16-byte coalesced accesses.  Same scheme as dim 2 of the fused kernels in li_sep2d.cu.
template <typename T, int V>
struct VecN {
    T v[V];
};
template <typename T, int V>
__device__ __forceinline__ VecN<T, V> ldv(const T *p)
{
    VecN<T, V> r;
    if constexpr (sizeof(T) * V == 16)
        *reinterpret_cast<uint4 *>(r.v) = *reinterpret_cast<const uint4 *>(p);
    else if constexpr (sizeof(T) * V == 8)
        *reinterpret_cast<uint2 *>(r.v) = *reinterpret_cast<const uint2 *>(p);
    else
        r.v[0] = p[0];
    return r;
}
template <typename T, int V>
__device__ __forceinline__ void stv(T *p, const VecN<T, V> &r)
{
    if constexpr (sizeof(T) * V == 16)
        *reinterpret_cast<uint4 *>(p) = *reinterpret_cast<const uint4 *>(r.v);
    else if constexpr (sizeof(T) * V == 8)
        *reinterpret_cast<uint2 *>(p) = *reinterpret_cast<const uint2 *>(r.v);
    else
        p[0] = r.v[0];
}
__device__ __forceinline__ int clampi_(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// Direct: the src planes base[i0] .. base[i1-1]+S-1 of the chunk stream through a window of S vectors whose rotation
// is static (the plane loop is unrolled by S); after plane r every row whose window ends at r is emitted.
// Flat / SafeFlat clamping = the plane index is clamped when loading (replicated edge planes).
template <typename T, int S, int V>
__global__ void __launch_bounds__(256)
tab_direct_stream_kernel(const int32_t *__restrict__ base, const T *__restrict__ W, int nrows, int ncols, int64_t npv,
                         int64_t npost, int CR, T alpha, T beta, const T *__restrict__ src, T *__restrict__ dst)
{
    const int64_t pv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pv >= npv) return;
    const int i0 = blockIdx.y * CR, i1 = min(i0 + CR, nrows);
    const int lo = base[i0], hi = base[i1 - 1] + S - 1;
    const int64_t npre = npv * V;
    for (int64_t q = blockIdx.z; q < npost; q += gridDim.z) {
        const T *sp = src + pv * V + q * npre * ncols;
        T *dp = dst + pv * V + (q * nrows + i0) * npre;
        VecN<T, V> h[S];
        int i = i0;
        for (int r0 = lo; r0 <= hi; r0 += S) {
#pragma unroll
            for (int u = 0; u < S; ++u) {
                const int r = r0 + u;
                if (r <= hi) {   // block-uniform
                    h[u] = ldv<T, V>(sp + (int64_t)clampi_(r, ncols - 1) * npre);
                    while (i < i1 && base[i] + S - 1 == r) {   // rows completed by plane r (block-uniform)
                        VecN<T, V> o;
#pragma unroll
                        for (int e = 0; e < V; ++e) o.v[e] = (T)0;
#pragma unroll
                        for (int t = 0; t < S; ++t) {
                            const T w = W[(int64_t)i * S + t];
#pragma unroll
                            for (int e = 0; e < V; ++e) o.v[e] = fma_(h[(u + 1 + t) % S].v[e], w, o.v[e]);
                        }
                        if (beta == (T)0) {   // dst is never read
#pragma unroll
                            for (int e = 0; e < V; ++e) o.v[e] = alpha == (T)1 ? o.v[e] : alpha * o.v[e];
                        } else {
                            const VecN<T, V> d = ldv<T, V>(dp);
#pragma unroll
                            for (int e = 0; e < V; ++e) o.v[e] = fma_(alpha, o.v[e], beta * d.v[e]);
                        }
                        stv<T, V>(dp, o);
                        dp += npre;
                        ++i;
                    }
                }
            }
        }
    }
}

// Adjoint (gather form, deterministic): the src rows feeding the chunk of dst planes [va, vb) stream in ascending
// order; S accumulators hold the dst planes v .. v+S-1 (static rotation); dst plane v is complete once every row
// whose folded window starts at v has been added -- ascending (row, tap) order, the order of the reference's
// sequential scatter.  Clamped taps use the folded weights of the table (li_table.cu: fold_kernel); slots that no
// tap of an edge row lands on are skipped, not multiplied by zero.
template <typename T, int S, int V>
__global__ void __launch_bounds__(256)
tab_adjoint_stream_kernel(const int32_t *__restrict__ base, const int32_t *__restrict__ bpf, const T *__restrict__ Wf,
                          int nrows, int ncols, int64_t npv, int64_t npost, int CR, T alpha, T beta,
                          const T *__restrict__ src, T *__restrict__ dst)
{
    const int64_t pv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pv >= npv) return;
    const int va = blockIdx.y * CR, vb = min(va + CR, ncols);
    const int64_t npre = npv * V;
    for (int64_t q = blockIdx.z; q < npost; q += gridDim.z) {
        int i = bpf[va];                                   // first row with folded base >= va-(S-1)
        const T *sp = src + pv * V + (q * nrows + i) * npre;
        T *dp = dst + pv * V + (q * ncols + va) * npre;
        VecN<T, V> acc[S];
#pragma unroll
        for (int t = 0; t < S; ++t)
#pragma unroll
            for (int e = 0; e < V; ++e) acc[t].v[e] = (T)0;
        for (int v0 = va - (S - 1); v0 < vb; v0 += S) {
#pragma unroll
            for (int u = 0; u < S; ++u) {
                const int vv = v0 + u;
                if (vv < vb) {   // block-uniform
                    const int iend = bpf[vv + S];          // rows with folded base <= vv end here
                    const bool interior = vv > 0 && vv < ncols - S;
                    for (; i < iend; ++i) {
                        const VecN<T, V> y = ldv<T, V>(sp);
                        sp += npre;
                        int slo = 0, shi = S - 1;
                        if (!interior) {   // first / last window of the array: only the slots the clamped taps land on
                            const int b = base[i];
                            slo = clampi_(b, ncols - 1) - vv, shi = clampi_(b + S - 1, ncols - 1) - vv;
                        }
#pragma unroll
                        for (int t = 0; t < S; ++t) {
                            if (t >= slo && t <= shi) {
                                const T w = Wf[(int64_t)i * S + t];
#pragma unroll
                                for (int e = 0; e < V; ++e) {
                                    const T x = alpha == (T)1 ? y.v[e] : alpha * y.v[e];
                                    acc[(u + t) % S].v[e] = fma_(w, x, acc[(u + t) % S].v[e]);
                                }
                            }
                        }
                    }
                    if (vv >= va) {   // dst plane vv is complete
                        VecN<T, V> o = acc[u];
                        if (beta != (T)0) {
                            const VecN<T, V> d = ldv<T, V>(dp);
#pragma unroll
                            for (int e = 0; e < V; ++e) o.v[e] = fma_(beta, d.v[e], o.v[e]);
                        }
                        stv<T, V>(dp, o);
                        dp += npre;
                    }
#pragma unroll
                    for (int e = 0; e < V; ++e) acc[u].v[e] = (T)0;
                }
            }
        }
    }
}

// ---- Adjoint along the contiguous dimension (npre == 1) of a windowed, monotone table -------------------------
// thread <-> dst column j; the rows feeding it are the contiguous range [rowstart[j], rowstart[j+S]) (plus every row
// hanging over the edge for the first / last column), row i contributes its tap j - base[i]: no transposed table, no
// per-entry indirection, ascending (row, tap) order like the reference's sequential scatter.  Adjacent threads read
// overlapping ranges of src (L1 hits), dst is written coalesced.
template <typename T, int S>
__global__ void __launch_bounds__(256)
tab_adjoint_window_d1_kernel(const int32_t *__restrict__ base, const int32_t *__restrict__ rowstart, const T *__restrict__ W,
                             int nrows, int ncols, int64_t npost, T alpha, T beta, const T *__restrict__ src,
                             T *__restrict__ dst)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    const int e0 = j == 0 ? 0 : rowstart[j], e1 = j == ncols - 1 ? nrows : rowstart[j + S];
    const bool edge = j == 0 || j == ncols - 1;
    for (int64_t q = blockIdx.y; q < npost; q += gridDim.y) {
        const T *sp = src + q * nrows;
        T *d = dst + q * ncols + j;
        T acc = beta == (T)0 ? (T)0 : beta * (*d);
        for (int i = e0; i < e1; ++i) {
            const T x = alpha == (T)1 ? sp[i] : alpha * sp[i];
            const int b = base[i];
            if (!edge) {
                acc = fma_(W[(int64_t)i * S + (j - b)], x, acc);
            } else {   // every tap of the row that clamps onto this column, in tap order
#pragma unroll
                for (int t = 0; t < S; ++t)
                    if (clampi_(b + t, ncols - 1) == j) acc = fma_(W[(int64_t)i * S + t], x, acc);
            }
        }
        *d = acc;
    }
}

// ---- Adjoint, atomic scatter (benchmarked variant; not run-to-run deterministic) ----
template <typename T>
__global__ void __launch_bounds__(256)
tab_adjoint_atomic_kernel(const int32_t *__restrict__ J, const T *__restrict__ W, int S, int nrows, int ncols,
                          int64_t npre, int64_t npost, T alpha, const T *__restrict__ src, T *__restrict__ dst)
{
    const int64_t total = npre * nrows * npost;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int64_t p = idx % npre;
        const int64_t r = idx / npre;
        const int i = (int)(r % nrows);
        const int64_t q = r / nrows;
        const T x = alpha == (T)1 ? src[idx] : alpha * src[idx];
        if (x == (T)0) continue;  // src/tabulated.jl:457
        T *dp = dst + p + q * npre * ncols;
        for (int s = 0; s < S; ++s) atomicAdd(dp + (int64_t)J[(int64_t)i * S + s] * npre, W[(int64_t)i * S + s] * x);
    }
}

// the streaming kernels need a windowed monotone table with a compile-time S; returns the widest vector (elements per
// thread, 16 bytes at most) that the length of the contiguous index and the alignment of the arrays allow, 0 = none
template <typename T, int SC>
static int stream_vec(const li_table *t, int64_t npre, const void *src, const void *dst)
{
    if (SC == 0 || opt(OPT_NO_STREAM_DN) || !t->windowed || !t->monotone) return 0;
    for (int V = 16 / (int)sizeof(T); V > 1; V >>= 1)
        if (npre % V == 0 && ((uintptr_t)src) % (V * sizeof(T)) == 0 && ((uintptr_t)dst) % (V * sizeof(T)) == 0) return V;
    return 1;
}

// chunks along the interpolated dimension: enough CTAs to fill the machine (about 8 per SM with the other two grid
// dimensions), but at least 8*S planes per chunk (each chunk re-reads S-1 halo planes)
static void stream_chunks(int n, int64_t other_ctas, int S, int *CR, int *chunks)
{
    const int64_t want = (148 * 8 + other_ctas - 1) / std::max<int64_t>(other_ctas, 1);
    int c = (int)std::max<int64_t>(1, std::min<int64_t>(want, n / std::max(8 * std::max(S, 1), 1)));
    c = std::max(1, std::min(c, 65535));
    *CR = (n + c - 1) / c;
    *chunks = (n + *CR - 1) / *CR;
}

// ---------------------------------------------------------------------------
template <typename T, int SC>
static int launch_direct(const li_table *t, int64_t npre, int64_t npost, T alpha, T beta, const T *src, T *dst,
                         cudaStream_t st)
{
    const int nrows = (int)t->nrows, ncols = (int)t->ncols;
    if (npre == 1) {
        dim3 grid((nrows + 255) / 256, (unsigned)std::min<int64_t>(npost, 65535));
        tab_direct_d1_kernel<T, SC><<<grid, 256, 0, st>>>(t->J, (const T *)t->W, t->S, nrows, ncols, npost, alpha,
                                                         beta, src, dst);
    } else if (const int V = stream_vec<T, SC>(t, npre, src, dst)) {
        const int64_t npv = npre / V;
        const int threads = npv >= 256 ? 256 : (npv >= 128 ? 128 : (npv >= 64 ? 64 : 32));
        const unsigned gx = (unsigned)((npv + threads - 1) / threads);
        int CR, chunks;
        stream_chunks(nrows, (int64_t)gx * std::min<int64_t>(npost, 65535), SC, &CR, &chunks);
        dim3 grid(gx, chunks, (unsigned)std::min<int64_t>(npost, 65535));
        constexpr int S_ = SC ? SC : 1, VMAX = 16 / (int)sizeof(T);
#define LI_DSTREAM(VV)                                                                                                  \
    tab_direct_stream_kernel<T, S_, VV><<<grid, threads, 0, st>>>(t->base, (const T *)t->W, nrows, ncols, npv, npost, CR, \
                                                                 alpha, beta, src, dst)
        if (V == VMAX)
            LI_DSTREAM(VMAX);
        else if (V == 2)
            LI_DSTREAM(2);
        else
            LI_DSTREAM(1);
#undef LI_DSTREAM
    } else {
        constexpr int RB = 4;
        const int threads = npre >= 256 ? 256 : (npre >= 128 ? 128 : (npre >= 64 ? 64 : 32));
        dim3 grid((unsigned)((npre + threads - 1) / threads), std::min((nrows + RB - 1) / RB, 65535),
                  (unsigned)std::min<int64_t>(npost, 65535));
        tab_direct_dn_kernel<T, SC, RB><<<grid, threads, 0, st>>>(t->J, (const T *)t->W, t->S, nrows, ncols, npre,
                                                                npost, alpha, beta, src, dst);
    }
    LI_CHECK_LAUNCH();
    return LI_OK;
}

template <typename T>
static int direct_T(const li_table *t, int64_t npre, int64_t npost, double alpha, double beta, const void *src,
                    void *dst, cudaStream_t st)
{
    const T a = (T)alpha, b = (T)beta;
    switch (t->S) {
    case 1: return launch_direct<T, 1>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    case 2: return launch_direct<T, 2>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    case 3: return launch_direct<T, 3>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    case 4: return launch_direct<T, 4>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    case 6: return launch_direct<T, 6>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    case 8: return launch_direct<T, 8>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    default: return launch_direct<T, 0>(t, npre, npost, a, b, (const T *)src, (T *)dst, st);
    }
}

template <typename T, int S>
static int launch_adjoint_stream(const li_table *t, int64_t npre, int64_t npost, T alpha, T beta, const T *src, T *dst,
                                 cudaStream_t st, bool *done)
{
    *done = false;
    const int V = stream_vec<T, S>(t, npre, src, dst);
    if (!V) return LI_OK;
    int rc = table_ensure_fold(t, st);
    if (rc != LI_OK) return rc;
    constexpr int VMAX = 16 / (int)sizeof(T);
    const int nrows = (int)t->nrows, ncols = (int)t->ncols;
    const int64_t npv = npre / V;
    const int threads = npv >= 256 ? 256 : (npv >= 128 ? 128 : (npv >= 64 ? 64 : 32));
    const unsigned gx = (unsigned)((npv + threads - 1) / threads);
    int CR, chunks;
    stream_chunks(ncols, (int64_t)gx * std::min<int64_t>(npost, 65535), S, &CR, &chunks);
    dim3 grid(gx, chunks, (unsigned)std::min<int64_t>(npost, 65535));
#define LI_ASTREAM(VV)                                                                                                   \
    tab_adjoint_stream_kernel<T, S, VV><<<grid, threads, 0, st>>>(t->base, t->bpf, (const T *)t->Wfold, nrows, ncols, npv, \
                                                                 npost, CR, alpha, beta, src, dst)
    if (V == VMAX)
        LI_ASTREAM(VMAX);
    else if (V == 2)
        LI_ASTREAM(2);
    else
        LI_ASTREAM(1);
#undef LI_ASTREAM
    LI_CHECK_LAUNCH();
    *done = true;
    return LI_OK;
}

template <typename T>
static int adjoint_T(const li_table *t, int64_t npre, int64_t npost, double alpha, double beta, const void *src,
                     void *dst, int variant, cudaStream_t st)
{
    const int nrows = (int)t->nrows, ncols = (int)t->ncols;
    const T a = (T)alpha, b = (T)beta;
    if (variant == LI_VARIANT_ATOMIC) {
        int rc = vscale_impl(t->dtype, dst, npre * ncols * npost, beta, st);
        if (rc != LI_OK) return rc;
        const int64_t total = npre * nrows * npost;
        const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 32);
        tab_adjoint_atomic_kernel<T><<<blocks, 256, 0, st>>>(t->J, (const T *)t->W, t->S, nrows, ncols, npre, npost,
                                                            a, (const T *)src, (T *)dst);
        LI_CHECK_LAUNCH();
        return LI_OK;
    }
    // strided dimension of a windowed monotone table: streaming gather (no transposed table needed)
    if (npre > 1 && ncols >= t->S) {
        bool done = false;
        int rcs = LI_OK;
        switch (t->S) {
#define LI_ADJS(SS)                                                                                            \
    case SS:                                                                                                   \
        rcs = launch_adjoint_stream<T, SS>(t, npre, npost, a, b, (const T *)src, (T *)dst, st, &done);          \
        break;
            LI_ADJS(1) LI_ADJS(2) LI_ADJS(3) LI_ADJS(4) LI_ADJS(6) LI_ADJS(8)
#undef LI_ADJS
        default: break;
        }
        if (rcs != LI_OK || done) return rcs;
    }
    // contiguous dimension of a windowed monotone table: gather over the row ranges of the window table
    if (npre == 1 && t->windowed && t->monotone && ncols >= 2 && !opt(OPT_NO_STREAM_DN)) {
        int rcw = table_ensure_rowstart(t, st);
        if (rcw != LI_OK) return rcw;
        dim3 grid((ncols + 255) / 256, (unsigned)std::min<int64_t>(npost, 65535));
        bool launched = true;
        switch (t->S) {
#define LI_ADJW(SS)                                                                                                  \
    case SS:                                                                                                         \
        tab_adjoint_window_d1_kernel<T, SS><<<grid, 256, 0, st>>>(t->base, t->rowstart, (const T *)t->W, nrows, ncols, npost, \
                                                                 a, b, (const T *)src, (T *)dst);                    \
        break;
            LI_ADJW(1) LI_ADJW(2) LI_ADJW(3) LI_ADJW(4) LI_ADJW(6) LI_ADJW(8)
#undef LI_ADJW
        default: launched = false; break;
        }
        if (launched) {
            LI_CHECK_LAUNCH();
            return LI_OK;
        }
    }
    int rc = table_ensure_csc(t, st);
    if (rc != LI_OK) return rc;
    if (npre == 1) {
        dim3 grid((ncols + 255) / 256, (unsigned)std::min<int64_t>(npost, 65535));
        tab_adjoint_gather_d1_kernel<T><<<grid, 256, 0, st>>>(t->colptr, t->tent, (const T *)t->W, t->S, nrows,
                                                             ncols, npost, a, b, (const T *)src, (T *)dst);
        if (t->nheavy > 0) {
            count_launch();
            dim3 hgrid((unsigned)t->nheavy, (unsigned)std::min<int64_t>(npost, 65535));
            tab_adjoint_heavy_d1_kernel<T><<<hgrid, 256, 0, st>>>(t->heavy + 1, t->colptr, t->tent, (const T *)t->W, t->S, nrows, ncols,
                                                                 npost, a, b, (const T *)src, (T *)dst);
        }
    } else {
        const int threads = npre >= 256 ? 256 : (npre >= 128 ? 128 : (npre >= 64 ? 64 : 32));
        dim3 grid((unsigned)((npre + threads - 1) / threads), std::min(ncols, 65535),
                  (unsigned)std::min<int64_t>(npost, 65535));
        tab_adjoint_gather_dn_kernel<T><<<grid, threads, 0, st>>>(t->colptr, t->tent, (const T *)t->W, t->S, nrows,
                                                                 ncols, npre, npost, a, b, (const T *)src, (T *)dst);
    }
    LI_CHECK_LAUNCH();
    return LI_OK;
}

int tab_apply_impl(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha, const void *src,
                   double beta, void *dst, int variant, cudaStream_t st)
{
    LI_REQUIRE(t && dst, LI_ERR_ARG, "null argument");
    LI_REQUIRE(npre >= 0 && npost >= 0, LI_ERR_DIM, "negative dimension");
    const int64_t nout = npre * npost * (adjoint ? t->ncols : t->nrows);
    if (nout == 0) return LI_OK;
    if (alpha == 0.0) return vscale_impl(t->dtype, dst, nout, beta, st);  // src/tabulated.jl:352-353
    LI_REQUIRE(src, LI_ERR_ARG, "null source");
    LI_REQUIRE(npre * npost * std::max(t->ncols, t->nrows) < (1LL << 62), LI_ERR_DIM, "array too large");
    // Contiguous dimension of a sorted table over many columns: the fused 2-D kernels with the identity as second
    // factor (staged rows, TMA ring) instead of the gather kernels below.  [len, npost] is viewed as
    // [len, n2, nbatch] with n2 a divisor of npost, so the identity table stays small.
    // (short vectors leave most of a 256-thread strip idle: measured break-even at a length of about 96)
    if (npre == 1 && npost >= 64 && std::max(t->nrows, t->ncols) >= 96 && t->windowed && t->monotone &&
        variant != LI_VARIANT_ATOMIC &&
        (t->S == 2 || t->S == 3 || t->S == 4 || t->S == 6) && !opt(OPT_NO_STREAM_DN)) {
        int64_t n2 = 0;
        if (npost <= 4096) {
            n2 = npost;
        } else {
            for (int64_t c = 4096; c >= 256; --c)
                if (npost % c == 0) {
                    n2 = c;
                    break;
                }
            // (no divisor: the gather kernels below -- an identity table of npost rows would cost device memory
            // per distinct batch size)
        }
        if (n2) {
            const li_table *ident = nullptr;
            int rc = table_identity(t->dtype, n2, st, &ident);
            if (rc != LI_OK) return rc;
            bool done = false;
            rc = sep2d_fused_apply(adjoint, t, ident, npost / n2, alpha, src, beta, dst, st, &done);
            if (rc != LI_OK || done) return rc;
        }
    }
    if (!adjoint)
        return t->dtype == LI_F32 ? direct_T<float>(t, npre, npost, alpha, beta, src, dst, st)
                                  : direct_T<double>(t, npre, npost, alpha, beta, src, dst, st);
    return t->dtype == LI_F32 ? adjoint_T<float>(t, npre, npost, alpha, beta, src, dst, variant, st)
                              : adjoint_T<double>(t, npre, npost, alpha, beta, src, dst, variant, st);
}

}  // namespace li

extern "C" {

int li_tab_apply(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha, const void *src_dev,
                 double beta, void *dst_dev, int variant, void *stream)
{
    return li::tab_apply_impl(adjoint, t, npre, npost, alpha, src_dev, beta, dst_dev, variant, li::as_stream(stream));
}

}  // extern "C"

static int tab_apply_host_impl(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha,
                               const void *src_host, double beta, void *dst_host, int variant, void *stream, bool sync)
{
    LI_REQUIRE(t && src_host && dst_host, LI_ERR_ARG, "null argument");
    cudaStream_t st = li::as_stream(stream);
    const size_t es = li::dtype_size(t->dtype);
    const size_t nin = (size_t)(npre * npost * (adjoint ? t->nrows : t->ncols)) * es;
    const size_t nout = (size_t)(npre * npost * (adjoint ? t->ncols : t->nrows)) * es;
    void *dsrc = nullptr, *ddst = nullptr;
    int rc = li::staging(0, nin ? nin : 1, &dsrc, stream);
    if (rc == LI_OK) rc = li::staging(1, nout ? nout : 1, &ddst, stream);
    if (rc != LI_OK) return rc;
    cudaError_t e = cudaMemcpyAsync(dsrc, src_host, nin, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && beta != 0.0) e = cudaMemcpyAsync(ddst, dst_host, nout, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = li::tab_apply_impl(adjoint, t, npre, npost, alpha, dsrc, beta, ddst, variant, st);
    if (e == cudaSuccess && rc == LI_OK) e = cudaMemcpyAsync(dst_host, ddst, nout, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && sync) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        li::set_error("host-staged apply failed: %s", cudaGetErrorString(e));
        return LI_ERR_CUDA;
    }
    return rc;
}

extern "C" {

int li_tab_apply_host(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha,
                      const void *src_host, double beta, void *dst_host, int variant, void *stream)
{
    return tab_apply_host_impl(adjoint, t, npre, npost, alpha, src_host, beta, dst_host, variant, stream, true);
}

int li_tab_apply_host_async(int adjoint, const li_table *t, int64_t npre, int64_t npost, double alpha,
                            const void *src_host, double beta, void *dst_host, int variant, void *stream)
{
    return tab_apply_host_impl(adjoint, t, npre, npost, alpha, src_host, beta, dst_host, variant, stream, false);
}

}  // extern "C"
