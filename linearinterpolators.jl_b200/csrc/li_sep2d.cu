// li_sep2d.cu -- fused 2-D separable interpolation, forward and adjoint (K7).
//
// H = A1*A2 with A1 / A2 table-backed interpolators along dimensions 1 / 2 (reference: the
// LazyAlgebra composition of test/benchmarks.jl:27-29, whose factors are the loops of
// src/tabulated.jl:394-478 / src/sparse.jl:719-800; same mathematics as src/separable.jl:56-162).
// The reference makes two full passes over memory with an intermediate array; here one CTA
// produces one output tile from one input tile, the intermediate lives in shared memory, and each
// element of src and dst crosses HBM once (plus tile halos, which hit in L2).
//
// Both tables must be "windowed": row i reads columns clamp(base[i] + s, 0, ncols-1), s = 0..S-1,
// which is what getcoefs (src/boundaries.jl:87-156) always produces.  The clamp is applied while
// staging the input tile, so the inner loops use un-clamped contiguous windows.
//
//   forward : stage src tile (clamped) -> pass 1 along dim 1 (smem -> smem, per-thread taps in
//             registers) -> pass 2 along dim 2 (threads along the contiguous dim, a sliding window of
//             S2 rows in registers, block-uniform weights) -> coalesced vector stores.
//   adjoint : the exact transpose: threads along the contiguous dim read y rows with coalesced vector
//             loads and slide S2 accumulators along dim 2 (the reference's sequential scatter order),
//             the dim-2 result goes to smem, then a gather along dim 1 with per-thread taps in registers
//             writes the owned dst tile.  No atomics; the result is deterministic.
#include <algorithm>

#include "li_common.cuh"

namespace li {

// ---------------------------------------------------------------------------------------------
template <typename T, int V>
struct VecT;
template <>
struct VecT<float, 4> { using type = float4; };
template <>
struct VecT<float, 2> { using type = float2; };
template <>
struct VecT<float, 1> { using type = float; };
template <>
struct VecT<double, 2> { using type = double2; };
template <>
struct VecT<double, 1> { using type = double; };

template <typename T, int V>
struct Pack {
    T v[V];
};

template <typename T, int V>
__device__ __forceinline__ Pack<T, V> ld_pack(const T *p)
{
    using VT = typename VecT<T, V>::type;
    Pack<T, V> r;
    *reinterpret_cast<VT *>(r.v) = *reinterpret_cast<const VT *>(p);
    return r;
}

template <typename T, int V>
__device__ __forceinline__ void st_pack(T *p, const Pack<T, V> &r)
{
    using VT = typename VecT<T, V>::type;
    *reinterpret_cast<VT *>(p) = *reinterpret_cast<const VT *>(r.v);
}

template <typename T>
__device__ __forceinline__ T fmaT(T a, T b, T c);
template <>
__device__ __forceinline__ float fmaT<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <>
__device__ __forceinline__ double fmaT<double>(double a, double b, double c) { return fma(a, b, c); }

__device__ __forceinline__ int clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

template <typename T>
struct SepArgs {
    const int32_t *base1, *base2;  // window bases (canonical, 0-based)
    const T *W1, *W2;              // [n*S]
    const int32_t *rs1, *re1;      // adjoint only: rowstart / rowend of dim 1 ([m1+1])
    const int32_t *rs2, *re2;      // adjoint only: dim 2
    int n1, n2;                    // "rows" side (fine grid for an upsampling operator)
    int m1, m2;                    // "columns" side
    int tiles1, tiles2;
    int F1MAX, F2MAX;              // forward: footprint bounds; adjoint: max contributing rows per tile
    int vec_ok;                    // rows of the big array are vector-aligned
    T alpha, beta;
    const T *src;
    T *dst;
};

// ---------------------------------------------------------------------------------------------
// Forward.  Output tile TO1 x TO2 of dst[n1, n2]; NT threads.
constexpr int FWD_TO1 = 128;
constexpr int FWD_TO2 = 64;
constexpr int FWD_NT = 256;

template <typename T, int S1, int S2, int V>
__global__ void __launch_bounds__(FWD_NT)
sep2d_fwd_kernel(const SepArgs<T> a)
{
    constexpr int TO1 = FWD_TO1, TO2 = FWD_TO2, NT = FWD_NT;
    constexpr int TPG = TO1 / V;    // threads per row group in pass 2
    constexpr int NG = NT / TPG;    // row groups
    constexpr int RG = TO2 / NG;    // output rows per group
    static_assert(TO1 + TO2 <= NT, "footprint scan needs TO1 + TO2 threads");
    static_assert(RG >= 1 && NG * RG == TO2, "bad tiling");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P1 = a.F1MAX;
    T *sT = reinterpret_cast<T *>(smem_raw);           // [F2MAX][TO1]  pass-1 output (vector-aligned rows)
    T *sS = sT + (size_t)a.F2MAX * TO1;                // [F2MAX][P1]   staged src tile
    T *sW2 = sS + (size_t)a.F2MAX * P1;                // [TO2][S2]
    int *sB2 = reinterpret_cast<int *>(sW2 + TO2 * S2);  // [TO2]
    int *sLim = sB2 + TO2;                             // lo1, hi1, lo2, hi2

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int bid = blockIdx.x;
    const int tile1 = bid % a.tiles1;
    bid /= a.tiles1;
    const int tile2 = bid % a.tiles2;
    const int64_t batch = bid / a.tiles2;
    const int i1_0 = tile1 * TO1, i2_0 = tile2 * TO2;
    const T *src = a.src + batch * (int64_t)a.m1 * a.m2;
    T *dst = a.dst + batch * (int64_t)a.n1 * a.n2;

    if (tid < 4) sLim[tid] = (tid & 1) ? INT_MIN : INT_MAX;
    __syncthreads();
    {   // footprint of the tile: [min base, max base + S) in each dimension
        int b = 0;
        const bool d1 = tid < TO1, d2 = tid >= TO1 && tid < TO1 + TO2;
        if (d1) b = a.base1[min(i1_0 + tid, a.n1 - 1)];
        if (d2) {
            b = a.base2[min(i2_0 + tid - TO1, a.n2 - 1)];
            sB2[tid - TO1] = b;
        }
        int lo = (d1 || d2) ? b : INT_MAX, hi = (d1 || d2) ? b : INT_MIN;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0 && (d1 || d2)) {   // TO1 is a multiple of 32: a warp is entirely dim 1 or dim 2
            atomicMin(&sLim[d1 ? 0 : 2], lo);
            atomicMax(&sLim[d1 ? 1 : 3], hi);
        }
        for (int k = tid; k < TO2 * S2; k += NT) {
            const int64_t g = (int64_t)i2_0 * S2 + k;
            sW2[k] = g < (int64_t)a.n2 * S2 ? a.W2[g] : (T)0;
        }
    }
    __syncthreads();
    const int lo1 = sLim[0], F1 = sLim[1] + S1 - lo1;
    const int lo2 = sLim[2], F2 = sLim[3] + S2 - lo2;

    // stage the (clamped) source tile: row r <-> source column-2 index clamp(lo2 + r)
    for (int r = warp; r < F2; r += NT / 32) {
        const T *row = src + (int64_t)clampi(lo2 + r, a.m2 - 1) * a.m1;
        for (int c = lane; c < F1; c += 32) sS[r * P1 + c] = row[clampi(lo1 + c, a.m1 - 1)];
    }

    // pass 1: interpolate along dim 1, thread <-> output column i1 (taps and weights in registers)
    const int c1 = tid % TO1;
    const int i1g = min(i1_0 + c1, a.n1 - 1);
    const int off1 = a.base1[i1g] - lo1;
    T w1[S1];
#pragma unroll
    for (int s = 0; s < S1; ++s) w1[s] = a.W1[(int64_t)i1g * S1 + s];
    __syncthreads();
    for (int r = tid / TO1; r < F2; r += NT / TO1) {
        const T *p = sS + r * P1 + off1;
        T acc = p[0] * w1[0];
#pragma unroll
        for (int s = 1; s < S1; ++s) acc = fmaT(p[s], w1[s], acc);
        sT[r * TO1 + c1] = acc;
    }
    __syncthreads();

    // pass 2: interpolate along dim 2.  thread <-> V adjacent columns, a group of TPG threads walks RG
    // consecutive output rows keeping the S2 source rows of the current window in registers.
    const int g = tid / TPG;
    const int cv = (tid % TPG) * V;
    const int col = i1_0 + cv;
    if (col >= a.n1) return;
    const bool vec = a.vec_ok && (col + V <= a.n1);
    Pack<T, V> win[S2];
    int cur = INT_MIN / 2;
    const int rend = min((g + 1) * RG, a.n2 - i2_0);
    for (int r = g * RG; r < rend; ++r) {
        const int b = sB2[r] - lo2;
        int adv = b - cur;
        if (adv < 0 || adv >= S2) {
#pragma unroll
            for (int s = 0; s < S2; ++s) win[s] = ld_pack<T, V>(sT + (b + s) * TO1 + cv);
        } else {
            while (adv > 0) {
#pragma unroll
                for (int s = 0; s + 1 < S2; ++s) win[s] = win[s + 1];
                win[S2 - 1] = ld_pack<T, V>(sT + (cur + S2) * TO1 + cv);
                ++cur;
                --adv;
            }
        }
        cur = b;
        T w2[S2];
#pragma unroll
        for (int s = 0; s < S2; ++s) w2[s] = sW2[r * S2 + s];
        Pack<T, V> out;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            T acc = win[0].v[v] * w2[0];
#pragma unroll
            for (int s = 1; s < S2; ++s) acc = fmaT(win[s].v[v], w2[s], acc);
            out.v[v] = acc;
        }
        T *d = dst + (int64_t)(i2_0 + r) * a.n1 + col;
        if (vec) {
            if (a.beta != (T)0) {
                const Pack<T, V> old = ld_pack<T, V>(d);
#pragma unroll
                for (int v = 0; v < V; ++v) out.v[v] = fmaT(a.alpha, out.v[v], a.beta * old.v[v]);
            } else if (a.alpha != (T)1) {
#pragma unroll
                for (int v = 0; v < V; ++v) out.v[v] *= a.alpha;
            }
            st_pack<T, V>(d, out);
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                if (col + v < a.n1) {
                    T o = out.v[v];
                    if (a.beta != (T)0)
                        o = fmaT(a.alpha, o, a.beta * d[v]);
                    else if (a.alpha != (T)1)
                        o *= a.alpha;
                    d[v] = o;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Adjoint.  The CTA owns the dst tile [c0, c0+TD1) x [d0, d0+TD2) of dst[m1, m2].
//   phase 1 (dim 2): thread <-> V adjacent src columns i1; group g owns RG dst rows and walks the src
//            rows that feed them in ascending order with S2 accumulators in registers; completed dst
//            rows go to sT[k2][i1].
//   phase 2 (dim 1): thread <-> dst column k1; its taps (row offset, weight) over i1 are loaded into
//            registers once and reused for all TD2 rows.
constexpr int ADJ_TD2 = 32;
constexpr int ADJ_NT = 256;
constexpr int ADJ_MAXE = 12;  // taps per dst column kept in registers; more go through a slow loop

template <typename T, int S1, int S2, int V, int TPG>
__global__ void __launch_bounds__(ADJ_NT)
sep2d_adj_kernel(const SepArgs<T> a, const int TD1)
{
    constexpr int TD2 = ADJ_TD2, NT = ADJ_NT;
    constexpr int NG = NT / TPG;   // row groups of phase 1
    constexpr int RG = TD2 / NG;   // dst rows owned by a group
    static_assert(RG >= 1 && NG * RG == TD2, "bad tiling");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int PT = TPG * V;                            // pitch of sT (columns of the src range, vector-aligned)
    T *sT = reinterpret_cast<T *>(smem_raw);           // [TD2][PT]
    T *sW1 = sT + (size_t)TD2 * PT;                    // [PT][S1] weights of the staged src columns
    int *sB1 = reinterpret_cast<int *>(sW1 + (size_t)PT * S1);  // [PT]

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int tile1 = bid % a.tiles1;
    bid /= a.tiles1;
    const int tile2 = bid % a.tiles2;
    const int64_t batch = bid / a.tiles2;
    const int c0 = tile1 * TD1, c1 = min(c0 + TD1, a.m1);   // owned dst columns [c0, c1)
    const int d0 = tile2 * TD2, d1 = min(d0 + TD2, a.m2);   // owned dst rows    [d0, d1)
    const T *src = a.src + batch * (int64_t)a.n1 * a.n2;
    T *dst = a.dst + batch * (int64_t)a.m1 * a.m2;

    // src columns feeding this tile: [r1lo, r1hi), start aligned down to a vector boundary
    const int r1lo_raw = a.rs1[c0], r1hi = a.re1[c1 - 1];
    const int r1lo = a.vec_ok ? (r1lo_raw / V) * V : r1lo_raw;
    const int ncol = r1hi - r1lo;   // <= PT by construction (host)

    // stage dim-1 bases/weights of the src columns
    for (int k = tid; k < ncol; k += NT) sB1[k] = a.base1[r1lo + k];
    for (int k = tid; k < ncol * S1; k += NT) sW1[k] = a.W1[(int64_t)r1lo * S1 + k];

    // ---- phase 1
    {
        const int g = tid / TPG;
        const int cv = (tid % TPG) * V;
        const int ka = d0 + g * RG, kb = min(ka + RG, d1);   // owned dst rows of this group
        // zero the rows this group owns (clamped taps accumulate into edge rows)
        for (int k = ka; k < kb; ++k) {
            Pack<T, V> z;
#pragma unroll
            for (int v = 0; v < V; ++v) z.v[v] = (T)0;
            st_pack<T, V>(sT + (k - d0) * PT + cv, z);
        }
        if (ka < kb && cv < ncol) {
            const int i2lo = a.rs2[ka], i2hi = a.re2[kb - 1];
            const int col = r1lo + cv;
            const bool vec = a.vec_ok && (col + V <= a.n1);
            Pack<T, V> acc[S2];
#pragma unroll
            for (int s = 0; s < S2; ++s)
#pragma unroll
                for (int v = 0; v < V; ++v) acc[s].v[v] = (T)0;
            int cur = INT_MIN / 2;   // dst row of acc[0]
            auto emit = [&](int k2, const Pack<T, V> &val) {
                const int kk = clampi(k2, a.m2 - 1);
                if (kk >= ka && kk < kb) {
                    T *p = sT + (kk - d0) * PT + cv;
                    if (k2 == kk && kk != 0 && kk != a.m2 - 1) {
                        st_pack<T, V>(p, val);
                    } else {   // edge rows collect every clamped tap
                        Pack<T, V> old = ld_pack<T, V>(p);
#pragma unroll
                        for (int v = 0; v < V; ++v) old.v[v] += val.v[v];
                        st_pack<T, V>(p, old);
                    }
                }
            };
            for (int i2 = i2lo; i2 < i2hi; ++i2) {
                const int b = a.base2[i2];
                int adv = b - cur;
                if (adv < 0 || adv >= S2) {   // disjoint window: flush everything
                    if (cur > INT_MIN / 4) {
#pragma unroll
                        for (int s = 0; s < S2; ++s) emit(cur + s, acc[s]);
                    }
#pragma unroll
                    for (int s = 0; s < S2; ++s)
#pragma unroll
                        for (int v = 0; v < V; ++v) acc[s].v[v] = (T)0;
                } else {
                    while (adv > 0) {
                        emit(cur, acc[0]);
#pragma unroll
                        for (int s = 0; s + 1 < S2; ++s) acc[s] = acc[s + 1];
#pragma unroll
                        for (int v = 0; v < V; ++v) acc[S2 - 1].v[v] = (T)0;
                        ++cur;
                        --adv;
                    }
                }
                cur = b;
                Pack<T, V> y;
                const T *p = src + (int64_t)i2 * a.n1 + col;
                if (vec) {
                    y = ld_pack<T, V>(p);
                } else {
#pragma unroll
                    for (int v = 0; v < V; ++v) y.v[v] = (col + v < a.n1) ? p[v] : (T)0;
                }
                if (a.alpha != (T)1) {
#pragma unroll
                    for (int v = 0; v < V; ++v) y.v[v] *= a.alpha;
                }
#pragma unroll
                for (int s = 0; s < S2; ++s) {
                    const T w = a.W2[(int64_t)i2 * S2 + s];
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[s].v[v] = fmaT(w, y.v[v], acc[s].v[v]);
                }
            }
            if (cur > INT_MIN / 4) {
#pragma unroll
                for (int s = 0; s < S2; ++s) emit(cur + s, acc[s]);
            }
        }
    }
    __syncthreads();

    // ---- phase 2: gather along dim 1
    for (int kc = tid; kc < (c1 - c0) * 1; kc += NT) {
        // (TD1 <= NT: each thread handles at most one dst column; the loop is for generality)
        const int k1 = c0 + kc;
        const int e0 = a.rs1[k1] - r1lo, e1 = a.re1[k1] - r1lo;   // contributing src columns (staged indices)
        // interior columns: one tap per contributing src column, kept in registers (static indexing);
        // edge columns (clamped taps, several per src column) and very long lists use the slow loop
        int off[ADJ_MAXE];
        T w[ADJ_MAXE];
        const bool interior = k1 > 0 && k1 < a.m1 - 1;
        const int cnt = interior ? min(e1 - e0, ADJ_MAXE) : 0;
#pragma unroll
        for (int k = 0; k < ADJ_MAXE; ++k) {
            off[k] = 0;
            w[k] = (T)0;
            if (k < cnt) {
                off[k] = e0 + k;
                w[k] = sW1[(e0 + k) * S1 + (k1 - sB1[e0 + k])];
            }
        }
        const int e_reg_end = e0 + cnt;
        for (int r = 0; r < d1 - d0; ++r) {
            const T *row = sT + r * PT;
            T acc = (T)0;
#pragma unroll
            for (int k = 0; k < ADJ_MAXE; ++k)
                if (k < cnt) acc = fmaT(w[k], row[off[k]], acc);
            for (int e = e_reg_end; e < e1; ++e) {   // slow path: very long tap lists (heavily clamped edges)
                const int b = sB1[e];
                int slo = k1 - b, shi = k1 - b;
                if (k1 == 0) slo = 0;
                if (k1 == a.m1 - 1) shi = S1 - 1;
                slo = max(slo, 0);
                shi = min(shi, S1 - 1);
                for (int s = slo; s <= shi; ++s) acc = fmaT(sW1[e * S1 + s], row[e], acc);
            }
            T *d = dst + (int64_t)(d0 + r) * a.m1 + k1;
            *d = a.beta == (T)0 ? acc : fmaT(a.beta, *d, acc);
        }
    }
}

// ---------------------------------------------------------------------------------------------
static const int kSmemLimit = 200 * 1024;

template <typename T, int S1, int S2>
static int launch_fwd(const SepArgs<T> &a, int64_t nbatch, cudaStream_t st, bool *done)
{
    constexpr int V = sizeof(T) == 4 ? 4 : 2;
    const size_t smem = ((size_t)a.F2MAX * FWD_TO1 + (size_t)a.F2MAX * a.F1MAX + FWD_TO2 * S2) * sizeof(T) +
                        (FWD_TO2 + 4) * sizeof(int);
    if (smem > (size_t)kSmemLimit) return LI_OK;   // footprint too large for staging: decline
    auto kern = sep2d_fwd_kernel<T, S1, S2, V>;
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        LI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        configured = kSmemLimit;
    }
    const int64_t blocks = (int64_t)a.tiles1 * a.tiles2 * nbatch;
    LI_REQUIRE(blocks <= 0x7fffffffLL, LI_ERR_UNSUPPORTED, "too many tiles");
    kern<<<(unsigned)blocks, FWD_NT, smem, st>>>(a);
    LI_CHECK_LAUNCH();
    *done = true;
    return LI_OK;
}

template <typename T, int S1, int S2>
static int launch_adj(const SepArgs<T> &a, int TD1, int64_t nbatch, cudaStream_t st, bool *done)
{
    constexpr int V = sizeof(T) == 4 ? 4 : 2;
    constexpr int TPG = 64;
    const int PT = TPG * V;
    const size_t smem = ((size_t)ADJ_TD2 * PT + (size_t)PT * S1) * sizeof(T) + (size_t)PT * sizeof(int);
    if (smem > (size_t)kSmemLimit) return LI_OK;
    auto kern = sep2d_adj_kernel<T, S1, S2, V, TPG>;
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        LI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        configured = kSmemLimit;
    }
    const int64_t blocks = (int64_t)a.tiles1 * a.tiles2 * nbatch;
    LI_REQUIRE(blocks <= 0x7fffffffLL, LI_ERR_UNSUPPORTED, "too many tiles");
    kern<<<(unsigned)blocks, ADJ_NT, smem, st>>>(a, TD1);
    LI_CHECK_LAUNCH();
    *done = true;
    return LI_OK;
}

template <typename T>
static int fused_T(int adjoint, const li_table *t1, const li_table *t2, int64_t nbatch, double alpha,
                   const void *src, double beta, void *dst, cudaStream_t st, bool *done)
{
    constexpr int V = sizeof(T) == 4 ? 4 : 2;
    SepArgs<T> a;
    memset(&a, 0, sizeof a);
    a.base1 = t1->base, a.base2 = t2->base;
    a.W1 = (const T *)t1->W, a.W2 = (const T *)t2->W;
    a.n1 = (int)t1->nrows, a.n2 = (int)t2->nrows, a.m1 = (int)t1->ncols, a.m2 = (int)t2->ncols;
    a.alpha = (T)alpha, a.beta = (T)beta;
    a.src = (const T *)src, a.dst = (T *)dst;
    int TD1 = 0;
    if (!adjoint) {
        int span1 = 0, span2 = 0;
        int rc = table_fwd_span(t1, FWD_TO1, st, &span1);
        if (rc == LI_OK) rc = table_fwd_span(t2, FWD_TO2, st, &span2);
        if (rc != LI_OK) return rc;
        a.F1MAX = span1 + t1->S, a.F2MAX = span2 + t2->S;
        a.tiles1 = (a.n1 + FWD_TO1 - 1) / FWD_TO1, a.tiles2 = (a.n2 + FWD_TO2 - 1) / FWD_TO2;
        a.vec_ok = (a.n1 % V == 0) && (((uintptr_t)dst) % (V * sizeof(T)) == 0);
    } else {
        if (!t1->monotone || !t2->monotone) return LI_OK;   // decline: the gather needs contiguous row ranges
        int rc = table_ensure_rowstart(t1, st);
        if (rc == LI_OK) rc = table_ensure_rowstart(t2, st);
        if (rc != LI_OK) return rc;
        a.rs1 = t1->rowstart, a.re1 = t1->rowstart + (t1->ncols + 1);
        a.rs2 = t2->rowstart, a.re2 = t2->rowstart + (t2->ncols + 1);
        a.vec_ok = (a.n1 % V == 0) && (((uintptr_t)src) % (V * sizeof(T)) == 0);
        // widest dst tile whose contributing src columns fit the staged width (64 threads x V columns)
        const int PT = 64 * V;
        const int cands[] = {256, 192, 160, 128, 120, 112, 96, 80, 64, 56, 48, 32, 24, 16, 8};
        for (int c : cands) {
            int rows = 0;
            rc = table_adj_rows(t1, c, st, &rows);
            if (rc != LI_OK) return rc;
            if (rows + (V - 1) <= PT) {
                TD1 = c;
                break;
            }
        }
        if (TD1 == 0) return LI_OK;   // decline
        a.tiles1 = (a.m1 + TD1 - 1) / TD1, a.tiles2 = (a.m2 + ADJ_TD2 - 1) / ADJ_TD2;
    }
    const int S1 = t1->S, S2 = t2->S;
#define LI_SEP(A1, A2)                                                                   \
    if (S1 == A1 && S2 == A2)                                                            \
        return adjoint ? launch_adj<T, A1, A2>(a, TD1, nbatch, st, done) : launch_fwd<T, A1, A2>(a, nbatch, st, done);
    LI_SEP(2, 2) LI_SEP(2, 4) LI_SEP(4, 2) LI_SEP(4, 4) LI_SEP(6, 6) LI_SEP(4, 6) LI_SEP(6, 4) LI_SEP(1, 1) LI_SEP(3, 3)
    LI_SEP(8, 8)
#undef LI_SEP
    return LI_OK;   // size pair not built: decline
}

int sep2d_fused_supported(const li_table *t1, const li_table *t2)
{
    return t1 && t2 && t1->windowed && t2->windowed && t1->dtype == t2->dtype;
}

int sep2d_fused_apply(int adjoint, const li_table *t1, const li_table *t2, int64_t nbatch, double alpha,
                      const void *src, double beta, void *dst, cudaStream_t st, bool *done)
{
    *done = false;
    if (!sep2d_fused_supported(t1, t2) || nbatch <= 0) return LI_OK;
    if (getenv("LI_B200_NO_FUSED")) return LI_OK;
    return t1->dtype == LI_F32 ? fused_T<float>(adjoint, t1, t2, nbatch, alpha, src, beta, dst, st, done)
                               : fused_T<double>(adjoint, t1, t2, nbatch, alpha, src, beta, dst, st, done);
}

}  // namespace li
