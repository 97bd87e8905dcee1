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
    const int32_t *bp1, *bp2;      // adjoint only: bptr tables (see li_table.cu), [m+S]
    int n1, n2;                    // "rows" side (fine grid for an upsampling operator)
    int m1, m2;                    // "columns" side
    int tiles1, tiles2;
    int F1MAX, F2MAX;              // forward: footprint bounds; adjoint: F2MAX = max src rows feeding one row group
    int vec_ok;                    // rows of the big array are 2-element-vector aligned
    T alpha, beta;
    const T *src;
    T *dst;
};

// ---------------------------------------------------------------------------------------------
// Forward: a register pipeline.  A thread owns two adjacent output columns (i1, i1+1) of a tile of
// TO1 = 2*TPG columns x TO2 rows; NG row groups split the TO2 rows.  For every source row of its group, in
// ascending order, the thread interpolates along dim 1 from the staged tile (taps and weights in
// registers; the two columns share their loads when the table never advances by more than one column
// per row, i.e. for any up-sampling), pushes the value into a window of S2 registers whose rotation is
// static (the loop is unrolled by S2), and emits every output row whose dim-2 window has just become
// complete (block-uniform count).  The dim-1 result never touches shared memory.
constexpr int FWD_TO2 = 64;
constexpr int FWD_NT = 256;

template <typename T, int S1, int S2, int NG, bool PAIR>
__global__ void __launch_bounds__(FWD_NT)
sep2d_fwd_kernel(const SepArgs<T> a)
{
    constexpr int NT = FWD_NT, TPG = NT / NG, TO1 = 2 * TPG, TO2 = FWD_TO2, RG = TO2 / NG;
    static_assert(NG * RG == TO2 && TO2 <= NT, "bad tiling");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P1 = a.F1MAX + 1;                          // one spare column: the paired load reads S1+1 taps
    T *sS = reinterpret_cast<T *>(smem_raw);             // [F2MAX][P1] staged src tile
    T *sW2 = sS + (size_t)a.F2MAX * P1;                  // [TO2][S2]
    int *sB2 = reinterpret_cast<int *>(sW2 + TO2 * S2);  // [TO2]
    int *sCnt = sB2 + TO2;                               // [NG][F2MAX] outputs completed by each source row

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int bid = blockIdx.x;
    const int tile1 = bid % a.tiles1;
    bid /= a.tiles1;
    const int tile2 = bid % a.tiles2;
    const int64_t batch = bid / a.tiles2;
    const int i1_0 = tile1 * TO1, i2_0 = tile2 * TO2;
    const T *src = a.src + batch * (int64_t)a.m1 * a.m2;
    T *dst = a.dst + batch * (int64_t)a.n1 * a.n2;

    // footprint of the tile (tables are monotone): [lo, lo + F) in each dimension
    const int lo1 = a.base1[i1_0], F1 = a.base1[min(i1_0 + TO1, a.n1) - 1] + S1 - lo1;
    const int rows = min(TO2, a.n2 - i2_0);
    const int lo2 = a.base2[i2_0], F2 = a.base2[i2_0 + rows - 1] + S2 - lo2;

    for (int k = tid; k < NG * a.F2MAX; k += NT) sCnt[k] = 0;
    __syncthreads();
    if (tid < rows) {
        const int b = a.base2[i2_0 + tid];
        sB2[tid] = b;
        atomicAdd(&sCnt[(tid / RG) * a.F2MAX + (b - lo2 + S2 - 1)], 1);
    }
    for (int k = tid; k < rows * S2; k += NT) sW2[k] = a.W2[(int64_t)i2_0 * S2 + k];

    // stage the (clamped) source tile: row r <-> source index clamp(lo2 + r) of dim 2
    for (int r = warp; r < F2; r += NT / 32) {
        const T *row = src + (int64_t)clampi(lo2 + r, a.m2 - 1) * a.m1;
        for (int c = lane; c < F1; c += 32) sS[r * P1 + c] = row[clampi(lo1 + c, a.m1 - 1)];
    }

    // per-thread dim-1 taps
    const int g = tid / TPG, cA = 2 * (tid % TPG);
    const int colA = i1_0 + cA;
    const bool haveA = colA < a.n1, haveB = colA + 1 < a.n1;
    const int iA = min(colA, a.n1 - 1), iB = min(colA + 1, a.n1 - 1);
    const int offA = a.base1[iA] - lo1, offB = a.base1[iB] - lo1;
    T w1A[S1], w1B[S1];
#pragma unroll
    for (int s = 0; s < S1; ++s) {
        w1A[s] = a.W1[(int64_t)iA * S1 + s];
        w1B[s] = a.W1[(int64_t)iB * S1 + s];
    }
    const bool shift = offB != offA;   // PAIR: offB - offA is 0 or 1
    __syncthreads();

    const int r0 = g * RG;
    if (!haveA || r0 >= rows) return;
    const int r1 = min(r0 + RG, rows);
    const int rb = sB2[r0] - lo2, re = sB2[r1 - 1] - lo2 + S2;   // source rows of this group
    const bool vec = a.vec_ok && haveB;
    int next = r0;                                              // next output row (tile-local) to emit
    T hA[S2], hB[S2];
#pragma unroll
    for (int s = 0; s < S2; ++s) hA[s] = hB[s] = (T)0;

    for (int r = rb; r < re; r += S2) {
#pragma unroll
        for (int u = 0; u < S2; ++u) {
            const int rr = r + u;
            if (rr < re) {
                // pass 1: dim 1, from the staged tile
                const T *p = sS + rr * P1 + offA;
                T accA, accB;
                if constexpr (PAIR) {
                    T v[S1 + 1];
#pragma unroll
                    for (int s = 0; s <= S1; ++s) v[s] = p[s];
                    accA = v[0] * w1A[0];
                    accB = (shift ? v[1] : v[0]) * w1B[0];
#pragma unroll
                    for (int s = 1; s < S1; ++s) {
                        accA = fmaT(v[s], w1A[s], accA);
                        accB = fmaT(shift ? v[s + 1] : v[s], w1B[s], accB);
                    }
                } else {
                    const T *q = sS + rr * P1 + offB;
                    accA = p[0] * w1A[0];
                    accB = q[0] * w1B[0];
#pragma unroll
                    for (int s = 1; s < S1; ++s) {
                        accA = fmaT(p[s], w1A[s], accA);
                        accB = fmaT(q[s], w1B[s], accB);
                    }
                }
                hA[u] = accA;
                hB[u] = accB;
                // pass 2: emit the output rows whose window [rr-S2+1, rr] is now complete
                const int n = sCnt[g * a.F2MAX + rr];
                for (int k = 0; k < n; ++k, ++next) {
                    T w2[S2];
#pragma unroll
                    for (int s = 0; s < S2; ++s) w2[s] = sW2[next * S2 + s];
                    T oA = hA[(u + 1) % S2] * w2[0], oB = hB[(u + 1) % S2] * w2[0];
#pragma unroll
                    for (int s = 1; s < S2; ++s) {
                        oA = fmaT(hA[(u + 1 + s) % S2], w2[s], oA);
                        oB = fmaT(hB[(u + 1 + s) % S2], w2[s], oB);
                    }
                    T *d = dst + (int64_t)(i2_0 + next) * a.n1 + colA;
                    if (a.beta != (T)0) {
                        oA = fmaT(a.alpha, oA, a.beta * d[0]);
                        if (haveB) oB = fmaT(a.alpha, oB, a.beta * d[1]);
                    } else if (a.alpha != (T)1) {
                        oA *= a.alpha;
                        oB *= a.alpha;
                    }
                    if (vec) {
                        Pack<T, 2> o;
                        o.v[0] = oA, o.v[1] = oB;
                        st_pack<T, 2>(d, o);
                    } else {
                        d[0] = oA;
                        if (haveB) d[1] = oB;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Adjoint: the transpose of the above.  The CTA owns the dst tile [c0, c0+TD1) x [d0, d0+TD2) of dst[m1, m2].
//   phase 1 (dim 2): a thread owns two adjacent src columns and streams the src rows feeding its row
//            group in ascending order (coalesced vector loads, a register queue PF rows ahead), adding each
//            row into S2 accumulators -- the order of the reference's sequential scatter.  When the window
//            base advances, the oldest accumulator is a finished dst row and goes to shared memory.
//   phase 2 (dim 1): a thread owns one dst column; its taps (one per contributing src column) sit in
//            registers and are reused for every row.  sT is stored in two parity planes so that the
//            stride-2 gather of a 2x operator is bank-conflict free.
constexpr int ADJ_TD2 = 32;
constexpr int ADJ_NT = 256;
constexpr int ADJ_MAXE = 12;  // taps per dst column kept in registers; longer lists use a slow loop
constexpr int ADJ_PF = 4;     // src rows in flight per thread

template <typename T, int S1, int S2, int NG>
__global__ void __launch_bounds__(ADJ_NT)
sep2d_adj_kernel(const SepArgs<T> a, const int TD1)
{
    constexpr int TD2 = ADJ_TD2, NT = ADJ_NT, TPG = NT / NG, PT = 2 * TPG, RG = TD2 / NG, PF = ADJ_PF;
    constexpr int PH = PT / 2;               // columns per parity plane
    constexpr int PLANE = TD2 * PH + 16;     // plane stride: +16 banks so that unit-stride gathers do not collide
    static_assert(NG * RG == TD2, "bad tiling");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sT = reinterpret_cast<T *>(smem_raw);                 // [2][TD2][PH] (+skew)
    T *sW1 = sT + 2 * PLANE;                                 // [PT][S1]
    T *sW2 = sW1 + (size_t)PT * S1;                          // [NG][F2MAX][S2]
    int *sB1 = reinterpret_cast<int *>(sW2 + (size_t)NG * a.F2MAX * S2);   // [PT]
    int *sB2 = sB1 + PT;                                     // [NG][F2MAX]

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

    // src columns feeding this tile: [r1lo, r1hi), start aligned down to an even column
    const int r1lo_raw = a.bp1[c0], r1hi = a.bp1[c1 - 1 + S1];
    const int r1lo = a.vec_ok ? (r1lo_raw & ~1) : r1lo_raw;
    const int ncol = r1hi - r1lo;   // <= PT (host)

    const int g = tid / TPG, ct = tid % TPG;
    const int ka = d0 + g * RG, kb = min(ka + RG, d1);      // dst rows owned by this group
    const int i2lo = ka < kb ? a.bp2[ka] : 0, i2hi = ka < kb ? a.bp2[kb - 1 + S2] : 0;

    for (int k = tid; k < ncol; k += NT) sB1[k] = a.base1[r1lo + k];
    for (int k = tid; k < ncol * S1; k += NT) sW1[k] = a.W1[(int64_t)r1lo * S1 + k];
    for (int k = ct; k < i2hi - i2lo; k += TPG) sB2[g * a.F2MAX + k] = a.base2[i2lo + k];
    for (int k = ct; k < (i2hi - i2lo) * S2; k += TPG) sW2[(size_t)g * a.F2MAX * S2 + k] = a.W2[(int64_t)i2lo * S2 + k];
    __syncthreads();

    // ---- phase 1
    if (ka < kb) {
        const int cA = 2 * ct;
        const int colA = r1lo + cA;
        const bool haveA = cA < ncol, haveB = cA + 1 < ncol && colA + 1 < a.n1;
        const bool vec = a.vec_ok && haveB;
        T *tA = sT + (ka - d0) * PH + ct;   // plane 0 (even staged columns), plane 1 at +PLANE
        auto load = [&](int i2) {
            Pack<T, 2> y;
            y.v[0] = y.v[1] = (T)0;
            if (haveA && i2 < i2hi) {
                const T *p = src + (int64_t)i2 * a.n1 + colA;
                if (vec) {
                    y = ld_pack<T, 2>(p);
                } else {
                    y.v[0] = p[0];
                    if (haveB) y.v[1] = p[1];
                }
            }
            return y;
        };
        Pack<T, 2> q[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) q[u] = load(i2lo + u);
        T A0[S2], A1[S2];
#pragma unroll
        for (int s = 0; s < S2; ++s) A0[s] = A1[s] = (T)0;
        int cur = ka - (S2 - 1);   // dst row of accumulator slot 0
        int done = 0;              // owned rows already stored
        // retire accumulator slot 0 (dst row `cur`), then shift
        auto retire = [&]() {
            if (cur < 0) {   // clamped taps: everything below row 0 belongs to row 0
                if constexpr (S2 > 1) {
                    A0[1] += A0[0];
                    A1[1] += A1[0];
                }
            } else if (cur >= ka) {
                if (cur == a.m2 - 1) {   // ... and everything above the last row to the last row
#pragma unroll
                    for (int s = 1; s < S2; ++s) {
                        A0[0] += A0[s];
                        A1[0] += A1[s];
                    }
                }
                tA[done * PH] = A0[0];
                tA[done * PH + PLANE] = A1[0];
                ++done;
            }
#pragma unroll
            for (int s = 0; s + 1 < S2; ++s) {
                A0[s] = A0[s + 1];
                A1[s] = A1[s + 1];
            }
            A0[S2 - 1] = A1[S2 - 1] = (T)0;
            ++cur;
        };
        const int *gb2 = sB2 + g * a.F2MAX;
        const T *gw2 = sW2 + (size_t)g * a.F2MAX * S2;
        for (int i2 = i2lo; i2 < i2hi; i2 += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int ii = i2 + u;
                if (ii < i2hi) {
                    Pack<T, 2> y = q[u];
                    q[u] = load(ii + PF);
                    const int b = gb2[ii - i2lo];
                    while (cur < b && cur < kb) retire();
                    if (a.alpha != (T)1) {
                        y.v[0] *= a.alpha;
                        y.v[1] *= a.alpha;
                    }
#pragma unroll
                    for (int s = 0; s < S2; ++s) {
                        const T w = gw2[(ii - i2lo) * S2 + s];
                        A0[s] = fmaT(w, y.v[0], A0[s]);
                        A1[s] = fmaT(w, y.v[1], A1[s]);
                    }
                }
            }
        }
        while (cur < kb) retire();
    }
    __syncthreads();

    // ---- phase 2: gather along dim 1; NT threads = (NT/128) row slices x 128 dst columns
    {
        constexpr int CT = 128;
        constexpr int NSL = NT / CT;
        const int kc = tid % CT, sl = tid / CT;
        for (int kk = kc; kk < c1 - c0; kk += CT) {
            const int k1 = c0 + kk;
            const int e0 = a.bp1[k1] - r1lo, e1 = a.bp1[k1 + S1] - r1lo;   // contributing staged src columns
            // interior columns: one tap per src column, in registers (static indexing); edge columns (clamped
            // taps, several per src column) and very long lists go through the slow loop
            int off[ADJ_MAXE];
            T w[ADJ_MAXE];
            const bool interior = k1 > 0 && k1 < a.m1 - 1;
            const int cnt = interior ? min(e1 - e0, ADJ_MAXE) : 0;
#pragma unroll
            for (int k = 0; k < ADJ_MAXE; ++k) {
                off[k] = 0;
                w[k] = (T)0;
                if (k < cnt) {
                    const int c = e0 + k;
                    off[k] = (c & 1) * PLANE + (c >> 1);
                    w[k] = sW1[c * S1 + (k1 - sB1[c])];
                }
            }
            const int e_reg_end = e0 + cnt;
            for (int r = sl; r < d1 - d0; r += NSL) {
                const T *row = sT + r * PH;
                T acc = (T)0;
#pragma unroll
                for (int k = 0; k < ADJ_MAXE; ++k)
                    if (k < cnt) acc = fmaT(w[k], row[off[k]], acc);
                for (int e = e_reg_end; e < e1; ++e) {
                    const int b = sB1[e];
                    int slo = k1 - b, shi = k1 - b;   // taps s with clamp(b + s) == k1
                    if (k1 == 0) slo = 0;
                    if (k1 == a.m1 - 1) shi = S1 - 1;
                    slo = max(slo, 0);
                    shi = min(shi, S1 - 1);
                    const T t = row[(e & 1) * PLANE + (e >> 1)];
                    for (int s = slo; s <= shi; ++s) acc = fmaT(sW1[e * S1 + s], t, acc);
                }
                T *d = dst + (int64_t)(d0 + r) * a.m1 + k1;
                *d = a.beta == (T)0 ? acc : fmaT(a.beta, *d, acc);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
static const int kSmemLimit = 200 * 1024;

template <typename KernT>
static int set_smem(KernT kern, size_t smem)
{
    if (smem > 48 * 1024) LI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return LI_OK;
}

template <typename T, int S1, int S2>
static int launch_fwd(SepArgs<T> &a, const li_table *t1, const li_table *t2, int64_t nbatch, cudaStream_t st,
                      bool *done)
{
    constexpr int NG = 2, TO1 = 2 * (FWD_NT / NG);
    int span1 = 0, span2 = 0;
    int rc = table_fwd_span(t1, TO1, st, &span1);
    if (rc == LI_OK) rc = table_fwd_span(t2, FWD_TO2, st, &span2);
    if (rc != LI_OK) return rc;
    a.F1MAX = span1 + S1, a.F2MAX = span2 + S2;
    a.tiles1 = (a.n1 + TO1 - 1) / TO1, a.tiles2 = (a.n2 + FWD_TO2 - 1) / FWD_TO2;
    const size_t smem = ((size_t)a.F2MAX * (a.F1MAX + 1) + FWD_TO2 * S2) * sizeof(T) +
                        (FWD_TO2 + (size_t)NG * a.F2MAX) * sizeof(int);
    if (smem > (size_t)kSmemLimit) return LI_OK;   // footprint too large for staging: decline
    const int64_t blocks = (int64_t)a.tiles1 * a.tiles2 * nbatch;
    LI_REQUIRE(blocks <= 0x7fffffffLL, LI_ERR_UNSUPPORTED, "too many tiles");
    if (t1->max_step <= 1) {
        auto kern = sep2d_fwd_kernel<T, S1, S2, NG, true>;
        rc = set_smem(kern, smem);
        if (rc != LI_OK) return rc;
        kern<<<(unsigned)blocks, FWD_NT, smem, st>>>(a);
    } else {
        auto kern = sep2d_fwd_kernel<T, S1, S2, NG, false>;
        rc = set_smem(kern, smem);
        if (rc != LI_OK) return rc;
        kern<<<(unsigned)blocks, FWD_NT, smem, st>>>(a);
    }
    LI_CHECK_LAUNCH();
    *done = true;
    return LI_OK;
}

template <typename T, int S1, int S2>
static int launch_adj(SepArgs<T> &a, const li_table *t1, const li_table *t2, int64_t nbatch, cudaStream_t st,
                      bool *done)
{
    constexpr int NG = 2, PT = 2 * (ADJ_NT / NG), RG = ADJ_TD2 / NG;
    int rc = table_ensure_rowstart(t1, st);
    if (rc == LI_OK) rc = table_ensure_rowstart(t2, st);
    if (rc != LI_OK) return rc;
    a.bp1 = t1->rowstart, a.bp2 = t2->rowstart;
    // widest dst tile whose contributing src columns fit the staged width
    int TD1 = 0;
    const int cands[] = {128, 120, 112, 96, 80, 64, 56, 48, 32, 24, 16, 8};
    for (int c : cands) {
        int rows = 0;
        rc = table_adj_rows(t1, c, st, &rows);
        if (rc != LI_OK) return rc;
        if (rows + 1 <= PT) {
            TD1 = c;
            break;
        }
    }
    if (TD1 == 0) return LI_OK;   // decline
    int r2 = 0;
    rc = table_adj_rows(t2, RG, st, &r2);
    if (rc != LI_OK) return rc;
    a.F2MAX = r2;
    a.tiles1 = (a.m1 + TD1 - 1) / TD1, a.tiles2 = (a.m2 + ADJ_TD2 - 1) / ADJ_TD2;
    const size_t smem = (2 * ((size_t)ADJ_TD2 * (PT / 2) + 16) + (size_t)PT * S1 + (size_t)NG * a.F2MAX * S2) * sizeof(T) +
                        ((size_t)PT + (size_t)NG * a.F2MAX) * sizeof(int);
    if (smem > (size_t)kSmemLimit) return LI_OK;
    const int64_t blocks = (int64_t)a.tiles1 * a.tiles2 * nbatch;
    LI_REQUIRE(blocks <= 0x7fffffffLL, LI_ERR_UNSUPPORTED, "too many tiles");
    auto kern = sep2d_adj_kernel<T, S1, S2, NG>;
    rc = set_smem(kern, smem);
    if (rc != LI_OK) return rc;
    kern<<<(unsigned)blocks, ADJ_NT, smem, st>>>(a, TD1);
    LI_CHECK_LAUNCH();
    *done = true;
    return LI_OK;
}

template <typename T>
static int fused_T(int adjoint, const li_table *t1, const li_table *t2, int64_t nbatch, double alpha,
                   const void *src, double beta, void *dst, cudaStream_t st, bool *done)
{
    SepArgs<T> a;
    memset(&a, 0, sizeof a);
    a.base1 = t1->base, a.base2 = t2->base;
    a.W1 = (const T *)t1->W, a.W2 = (const T *)t2->W;
    a.n1 = (int)t1->nrows, a.n2 = (int)t2->nrows, a.m1 = (int)t1->ncols, a.m2 = (int)t2->ncols;
    a.alpha = (T)alpha, a.beta = (T)beta;
    a.src = (const T *)src, a.dst = (T *)dst;
    // the big ("rows"-side) array is accessed with 2-element vectors when its rows are aligned
    const void *big = adjoint ? src : (const void *)dst;
    a.vec_ok = (a.n1 % 2 == 0) && (((uintptr_t)big) % (2 * sizeof(T)) == 0);
    const int S1 = t1->S, S2 = t2->S;
#define LI_SEP(A1, A2)                                                                   \
    if (S1 == A1 && S2 == A2)                                                            \
        return adjoint ? launch_adj<T, A1, A2>(a, t1, t2, nbatch, st, done)             \
                       : launch_fwd<T, A1, A2>(a, t1, t2, nbatch, st, done);
    LI_SEP(1, 1) LI_SEP(2, 2) LI_SEP(3, 3) LI_SEP(4, 4) LI_SEP(6, 6) LI_SEP(8, 8)
    LI_SEP(2, 4) LI_SEP(4, 2) LI_SEP(4, 6) LI_SEP(6, 4)
#undef LI_SEP
    return LI_OK;   // size pair not built: decline
}

int sep2d_fused_supported(const li_table *t1, const li_table *t2)
{
    // the fused kernels walk monotone windowed tables (what getcoefs produces for sorted positions)
    return t1 && t2 && t1->windowed && t2->windowed && t1->monotone && t2->monotone && t1->dtype == t2->dtype;
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
