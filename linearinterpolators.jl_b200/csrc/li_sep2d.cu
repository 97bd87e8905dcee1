// li_sep2d.cu -- fused 2-D separable interpolation, forward and adjoint (K7).
//
// H = A1*A2 with A1 / A2 table-backed interpolators along dimensions 1 / 2 (reference: the
// LazyAlgebra composition of test/benchmarks.jl:27-29, whose factors are the loops of
// src/tabulated.jl:394-478 / src/sparse.jl:719-800; same mathematics as src/separable.jl:56-162).
// The reference makes two full passes over memory with an intermediate array; here each element of src and dst
// crosses HBM once (plus a few halo rows per chunk).
//
// Both tables must be "windowed" and monotone: row i reads columns clamp(base[i] + s, 0, ncols-1), s = 0..S-1, with
// non-decreasing bases -- what getcoefs (src/boundaries.jl:87-156) produces for sorted positions.
//
//   forward : sep2d_fwd_kernel (li_sep2d_fwd.cuh): a streaming register pipeline fed by a ring of TMA bulk copies.
//   adjoint : sep2d_adj_stream_kernel (li_sep2d_adjs.cuh): the same data flow run backwards; it is tried first and
//             declines what it cannot stage.  The tile-owned kernels of this file (sep2d_adj_regular_kernel,
//             sep2d_adj_kernel: dim-2 accumulation from global memory into a shared-memory intermediate, then a
//             dim-1 gather) are the fallback for unaligned sources and for down-sampling along dim 1.
//   No kernel uses atomics; every dst element is written once by its owner, in a fixed order (deterministic).
//   This file: the tile-owned adjoint kernels, the dispatch over the kernel-size pairs (the eight original pairs + the
//   identity pairs here, the other pairs of {2,3,4,6,8}^2 in li_sep2d_more_*.cu) and the banded host-buffer pipeline.
//   Three dimensions: li_sep3d.cuh (the same two streaming kernels with a third factor applied while staging).
#include <algorithm>
#include <map>

#include "li_common.cuh"
#include "li_sep2d.cuh"
#include "li_sep2d_fwd.cuh"

namespace li {

// ---------------------------------------------------------------------------------------------
// Adjoint: the transpose of the forward operator, tile-owned and atomic-free.
//
// The CTA owns the dst tile [c0, c0+TD1) x [d0, d0+TD2) of dst[m1, m2]; the big src array is read exactly once
// per tile (plus halos) straight from global memory, many small CTAs keep every SM busy.
//   phase 1 (dim 2): a thread owns two adjacent src columns; NG row groups each own RG dst rows.  The src rows
//            feeding the group arrive through a thread-private ring of cp.async copies (depth D, no barrier, no
//            register staging), in ascending order.  dst rows are finished one at a time: for dst row v the
//            thread adds every src row whose window starts at v into S2 accumulators (static slot rotation:
//            the row loop is unrolled by S2), then slot v is complete -- the order of the reference's
//            sequential scatter (src/tabulated.jl:455-462) -- and goes to shared memory (sT).  The clamped taps
//            of the first/last dst rows are folded into the staged copy of the weights (no edge cases).
//   phase 2 (dim 1): a thread owns one dst column and RS rows; its taps (one per contributing src column) sit
//            in registers.  sT is stored in two parity planes so that the stride-2 gather of a 2x operator is
//            bank-conflict free.
// No atomics: every dst element is written once by its owner, in a fixed order -> deterministic.
constexpr int ADJ_TD2 = 32;
constexpr int ADJ_NT = 256;
constexpr int ADJ_NG = 2;
constexpr int ADJ_D = 8;      // src rows in flight per thread (cp.async ring depth)
constexpr int ADJ_MAXE = 12;  // taps per dst column kept in registers; longer lists use a slow loop

template <int BYTES>
__device__ __forceinline__ void cp_async(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int S1, int S2, bool BETA>
__global__ void __launch_bounds__(ADJ_NT, (sizeof(T) == 4 && S1 <= 4 && S2 <= 4) ? 4 : 2)
sep2d_adj_kernel(const SepArgs<T> a, const int TD1, const int R2 /* staged dim-2 rows per group */)
{
    constexpr int TD2 = ADJ_TD2, NT = ADJ_NT, NG = ADJ_NG, TPG = NT / NG, PT = 2 * TPG, RG = TD2 / NG, D = ADJ_D;
    constexpr int PH = PT / 2;               // columns per parity plane
    constexpr int PLANE = TD2 * PH + 16;     // plane stride: +16 banks so that unit-stride gathers do not collide
    static_assert(NG * RG == TD2, "bad tiling");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sT = reinterpret_cast<T *>(smem_raw);                 // [2][TD2][PH] (+skew)
    T *sRing = sT + 2 * PLANE;                               // [D][NT][2] thread-private ring of src row pairs
    T *sW1 = sRing + (size_t)D * NT * 2;                     // [PT][S1]
    T *sW2 = sW1 + (size_t)PT * S1;                          // [NG][R2][S2]  alpha * W2, edge taps folded
    int *sB1 = reinterpret_cast<int *>(sW2 + (size_t)NG * R2 * S2);   // [PT]
    int *sNv = sB1 + PT;                                     // [NG][RG + S2] src rows whose (folded) window starts at a dst row

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int tile1 = bid % a.tiles1;
    bid /= a.tiles1;
    const int tile2 = a.rowlist ? a.rowlist[bid % a.tiles2] : bid % a.tiles2;   // (a.tiles2 = tile rows of this launch)
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
    const bool active = ka < kb;
    const int i2lo = active ? a.bp2[ka] : 0, i2hi = active ? a.bp2[kb - 1 + S2] : 0;
    const int nrow = i2hi - i2lo;                            // <= R2 (host)
    const int v0 = ka - (S2 - 1);                            // first dst row an accumulator slot can stand for

    // ---- start streaming this thread's two src columns right away.  Ring slot of row k: byte offset
    // (k * SLOT) & (D * SLOT - 1) from ring_u32 (D is a power of two): no wrap branch, no pointer table.
    static_assert((D & (D - 1)) == 0, "ring depth must be a power of two");
    constexpr uint32_t SLOT = 2u * NT * (uint32_t)sizeof(T);   // bytes between consecutive ring slots
    constexpr uint32_t RMASK = D * SLOT - 1u;
    const int cA = 2 * ct;
    const int colA = r1lo + cA;
    const bool haveA = active && cA < ncol, haveB = haveA && cA + 1 < ncol && colA + 1 < a.n1;
    const bool vec = a.vec_ok && haveB;
    const T *gp = src + (int64_t)i2lo * a.n1 + colA;         // next row to request
    const uint32_t ring_u32 = smem_u32(sRing + 2 * tid);
    const int64_t gstep = (int64_t)a.n1 * (int64_t)sizeof(T);
    int togo = haveA ? nrow : 0;                             // rows still to be requested
    // request the next row into the slot at byte offset `off` (predicated, branch-free in the vector case)
    auto request = [&](uint32_t off) {
        if (vec) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %2, 0;\n\t"
                "@p cp.async.ca.shared.global [%0], [%1], %3;\n\t"
                "cp.async.commit_group;\n\t}" ::"r"(ring_u32 + off),
                "l"(gp), "r"(togo), "n"(2 * sizeof(T))
                : "memory");
        } else {
            if (togo > 0) {
                cp_async<sizeof(T)>(sRing + 2 * tid + off / sizeof(T), gp);
                if (haveB) cp_async<sizeof(T)>(sRing + 2 * tid + off / sizeof(T) + 1, gp + 1);
            }
            cp_async_commit();
        }
        --togo;
        gp = reinterpret_cast<const T *>(reinterpret_cast<const char *>(gp) + gstep);
    };
#pragma unroll
    for (int d = 0; d < D; ++d) request((uint32_t)d * SLOT);

    // ---- stage the small tables (plain copies: the folding of the clamped dim-2 taps is precomputed per table)
    for (int k = tid; k < ncol; k += NT) sB1[k] = a.base1[r1lo + k];
    for (int k = tid; k < ncol * S1; k += NT) sW1[k] = a.W1[(int64_t)r1lo * S1 + k];
    for (int k = ct; k < nrow * S2; k += TPG) sW2[(size_t)g * R2 * S2 + k] = a.W2f[(int64_t)i2lo * S2 + k];
    // src rows whose (folded) window starts at dst row v0 + k; rows above the last full window have none
    for (int k = ct; k < RG + S2; k += TPG) {
        const int v = v0 + k;
        sNv[g * (RG + S2) + k] = (active && v >= 0 && v < kb) ? a.bp2[v + S2] - a.bp2[v + S2 - 1] : 0;
    }
    __syncthreads();

    // ---- phase 1: dst rows v0, v0+1, ... of the group are finished one at a time (rows below ka are discarded)
    if (active) {
        T A0[S2], A1[S2];
#pragma unroll
        for (int s = 0; s < S2; ++s) A0[s] = A1[s] = (T)0;
        const int *pn = sNv + g * (RG + S2);
        const T *pw = sW2 + (size_t)g * R2 * S2;
        T *tp = sT + (ka - d0) * PH + ct;   // plane 0 (even staged columns); plane 1 at +PLANE
        uint32_t roff = 0;                  // ring byte offset of the next row to consume
        uint32_t woff = 0;                  // byte offset of the next row's weights
        bool first = true;
        for (int v = v0; v < kb; v += S2) {
#pragma unroll
            for (int u = 0; u < S2; ++u) {
                const int vv = v + u;
                if (vv < kb) {
                    // add every src row whose window starts at dst row vv: tap s goes to slot (u + s) % S2
                    for (int n = *pn++; n > 0; --n) {
                        cp_async_wait<D - 2>();
                        Pack<T, 2> y;
                        if constexpr (sizeof(T) == 4)
                            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(y.v[0]), "=f"(y.v[1]) : "r"(ring_u32 + roff));
                        else
                            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(y.v[0]), "=d"(y.v[1]) : "r"(ring_u32 + roff));
                        // refill the slot consumed one row ago (its load has long completed)
                        if (!first) request((roff - SLOT) & RMASK);
                        first = false;
                        roff = (roff + SLOT) & RMASK;
                        T w[S2];
                        ld_weights<T, S2>(reinterpret_cast<const T *>(reinterpret_cast<const char *>(pw) + woff), w);
                        if (vv > 0 && vv < a.m2 - S2) {   // interior windows: every slot is fed
#pragma unroll
                            for (int s = 0; s < S2; ++s) {
                                A0[(u + s) % S2] = fmaT(w[s], y.v[0], A0[(u + s) % S2]);
                                A1[(u + s) % S2] = fmaT(w[s], y.v[1], A1[(u + s) % S2]);
                            }
                        } else {   // first / last window of the array: only the slots the clamped taps land on
                            const int b = a.base2[i2lo + (int)(woff / (S2 * sizeof(T)))];
                            const int lo = clampi(b, a.m2 - 1) - vv, hi = clampi(b + S2 - 1, a.m2 - 1) - vv;
#pragma unroll
                            for (int s = 0; s < S2; ++s)
                                if (s >= lo && s <= hi) {
                                    A0[(u + s) % S2] = fmaT(w[s], y.v[0], A0[(u + s) % S2]);
                                    A1[(u + s) % S2] = fmaT(w[s], y.v[1], A1[(u + s) % S2]);
                                }
                        }
                        woff += S2 * sizeof(T);
                    }
                    if (vv >= ka) {   // dst row vv is complete
                        tp[0] = A0[u];
                        tp[PLANE] = A1[u];
                        tp += PH;
                    }
                    A0[u] = A1[u] = (T)0;
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- phase 2: gather along dim 1; NT threads = NSL row slices x CT dst columns
    {
        constexpr int CT = 128, NSL = NT / CT, RS = TD2 / NSL;   // RS rows per slice
        const int kc = tid % CT, sl = tid / CT;
        const int nr = min(RS, d1 - d0 - sl * RS);               // valid rows of this slice
        for (int kk0 = kc; kk0 < c1 - c0 && nr > 0; kk0 += CT) {
            const int k1 = c0 + kk0;
            const int e0 = a.bp1[k1] - r1lo, e1 = a.bp1[k1 + S1] - r1lo;   // contributing staged src columns
            const T *rowbase = sT + sl * RS * PH;
            T acc[RS];
#pragma unroll
            for (int r = 0; r < RS; ++r) acc[r] = (T)0;
            // interior columns: one tap per src column; tap k reads staged column c = e0 + k, stored in plane
            // (c & 1) at index (c >> 1): with q = e0 & 1 the even taps live at pE[k/2], the odd ones at pO[(k-1)/2]
            const bool interior = k1 > 0 && k1 < a.m1 - 1;
            const int cnt = interior ? min(e1 - e0, ADJ_MAXE) : 0;
            const int q0 = e0 & 1;
            const T *pE = rowbase + q0 * PLANE + (e0 >> 1);
            const T *pO = rowbase + (1 - q0) * PLANE + (e0 >> 1) + q0;
#pragma unroll
            for (int k = 0; k < ADJ_MAXE; ++k) {
                if (k < cnt) {
                    const T wk = sW1[(e0 + k) * S1 + (k1 - sB1[e0 + k])];
                    const T *p = (k & 1) ? pO + (k - 1) / 2 : pE + k / 2;
#pragma unroll
                    for (int r = 0; r < RS; ++r) acc[r] = fmaT(wk, p[r * PH], acc[r]);
                }
            }
            // edge columns carry the clamped taps of every out-of-range src column; they and very long tap
            // lists take this generic loop
            for (int e = e0 + cnt; e < e1; ++e) {
                const int b = sB1[e];
                int slo = k1 - b, shi = k1 - b;   // taps s with clamp(b + s) == k1
                if (k1 == 0) slo = 0;
                if (k1 == a.m1 - 1) shi = S1 - 1;
                slo = max(slo, 0);
                shi = min(shi, S1 - 1);
                const T *p = rowbase + (e & 1) * PLANE + (e >> 1);
                for (int s = slo; s <= shi; ++s) {
                    const T ws = sW1[e * S1 + s];
#pragma unroll
                    for (int r = 0; r < RS; ++r) acc[r] = fmaT(ws, p[r * PH], acc[r]);
                }
            }
            T *d = dst + (int64_t)(d0 + sl * RS) * a.m1 + k1;
#pragma unroll
            for (int r = 0; r < RS; ++r) {
                if (r < nr) {
                    T *dr = d + (int64_t)r * a.m1;
                    if constexpr (BETA)
                        *dr = fmaT(a.beta, *dr, a.alpha * acc[r]);
                    else
                        *dr = a.alpha * acc[r];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Adjoint, statically scheduled kernel for REGULAR tile rows: full tile rows whose dst rows are each fed by exactly
// R src rows with no clamped tap (the interior of a same-rate (R = 1) or 2x (R = 2) resampling; the host finds them
// from the table, li_table.cu: table_tile_rows).  Same tiling, same phases and same summation order as the generic
// kernel above, but phase 1 has no per-row control flow at all: row j of an unrolled block of NB rows belongs to dst
// row v + j / R, the src rows are plain vector loads kept NB rows ahead in registers, and the only index lookups
// (bptr entries, the dim-1 taps of phase 2) are issued in one batch at the start so the prologue costs one memory
// round trip.  (A micro-benchmark of this data flow, scratch/mb/adj_v9.cu, runs config #2 in 85 us.)
template <typename T, int S1, int S2, bool BETA, int R>
__global__ void __launch_bounds__(ADJ_NT, (sizeof(T) == 4 && S1 <= 4 && S2 <= 4) ? 4 : 2)
sep2d_adj_regular_kernel(const SepArgs<T> a, const int TD1)
{
    constexpr int TD2 = ADJ_TD2, NT = ADJ_NT, NG = ADJ_NG, TPG = NT / NG, PT = 2 * TPG, RG = TD2 / NG;
    constexpr int PH = PT / 2, PLANE = TD2 * PH + 16;
    constexpr int NBM = (S2 * R >= 4) ? 1 : (4 + S2 * R - 1) / (S2 * R);
    constexpr int NB = S2 * R * NBM;                         // rows per unrolled block == register queue depth
    constexpr int NROW = R * (RG + S2 - 1);                  // src rows of a group
    constexpr int CT = 128, NSL = NT / CT, RS = TD2 / NSL;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sT = reinterpret_cast<T *>(smem_raw);                 // [2][TD2][PH] (+skew)
    T *sW2 = sT + 2 * PLANE;                                 // [NG][NROW][S2] folded dim-2 weights

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int tile1 = bid % a.tiles1;
    bid /= a.tiles1;
    const int tile2 = a.rowlist[bid % a.tiles2];
    const int64_t batch = bid / a.tiles2;
    const int c0 = tile1 * TD1, c1 = min(c0 + TD1, a.m1);
    const int d0 = tile2 * TD2;                              // (regular tile rows are full: d1 = d0 + TD2)
    const T *src = a.src + batch * (int64_t)a.n1 * a.n2;
    T *dst = a.dst + batch * (int64_t)a.m1 * a.m2;
    const int g = tid / TPG, ct = tid % TPG;
    const int ka = d0 + g * RG, kb = ka + RG, v0 = ka - (S2 - 1);

    // round 1: every index lookup of the tile
    const int r1lo_raw = a.bp1[c0], r1hi = a.bp1[c1 - 1 + S1];
    const int i2lo = a.bp2[ka];
    const int kc = tid % CT, sl = tid / CT;
    const int k1a = min(c0 + kc, a.m1 - 1);
    int pe0 = a.bp1[k1a], pe1 = a.bp1[k1a + S1];
    const int r1lo = r1lo_raw & ~1, ncol = r1hi - r1lo;      // (launched on vector-aligned sources only)

    // round 2: the first NB src rows of this thread's column pair, the dim-2 weights, the dim-1 taps
    const bool have = 2 * ct < ncol;
    const T *gp = src + (int64_t)i2lo * a.n1 + r1lo + 2 * ct;
    const int64_t gstep = (int64_t)a.n1 * (int64_t)sizeof(T);
    int togo = have ? NROW : 0;
    auto load = [&]() {
        Pack<T, 2> y;
        y.v[0] = y.v[1] = (T)0;
        if (togo > 0) y = ld_pack<T, 2>(gp);   // (the column after the last staged one is inside the row)
        --togo;
        gp = reinterpret_cast<const T *>(reinterpret_cast<const char *>(gp) + gstep);
        return y;
    };
    Pack<T, 2> q[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) q[j] = load();
    for (int k = ct; k < NROW * S2; k += TPG) sW2[g * NROW * S2 + k] = a.W2f[(int64_t)i2lo * S2 + k];
    // dim-1 taps of dst column k1 (interior columns: one tap per contributing src column, straight from global --
    // they are only needed in phase 2); edge columns and very long lists take the generic loop of phase 2
    auto taps = [&](int k1, int e_first, int e_end, T (&tw)[ADJ_MAXE]) {
        const bool interior = k1 > 0 && k1 < a.m1 - 1 && k1 < c1;
        const int cnt = interior ? min(e_end - e_first, ADJ_MAXE) : 0;
#pragma unroll
        for (int k = 0; k < ADJ_MAXE; ++k) {
            tw[k] = (T)0;
            if (k < cnt) {
                const int i1 = e_first + k;
                tw[k] = a.W1[(int64_t)i1 * S1 + (k1 - a.base1[i1])];
            }
        }
        return cnt;
    };
    T tw[ADJ_MAXE];
    int cnt = taps(c0 + kc, pe0, pe1, tw);
    __syncthreads();

    // ---- phase 1: static schedule
    {
        T A0[S2], A1[S2];
#pragma unroll
        for (int s = 0; s < S2; ++s) A0[s] = A1[s] = (T)0;
        const T *pw = sW2 + g * NROW * S2;
        T *tp = sT + (ka - d0) * PH + ct;   // plane 0 (even staged columns); plane 1 at +PLANE
        for (int v = v0; v < kb; v += NB / R) {
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const int vv = v + j / R;
                if (vv < kb) {   // (block-uniform) only the last block is partial
                    const Pack<T, 2> y = q[j];
                    q[j] = load();
                    T w[S2];
                    ld_weights<T, S2>(pw, w);
                    pw += S2;
#pragma unroll
                    for (int s = 0; s < S2; ++s) {
                        A0[(j / R + s) % S2] = fmaT(w[s], y.v[0], A0[(j / R + s) % S2]);
                        A1[(j / R + s) % S2] = fmaT(w[s], y.v[1], A1[(j / R + s) % S2]);
                    }
                    if (j % R == R - 1) {   // dst row vv is complete
                        if (vv >= ka) {
                            tp[0] = A0[(j / R) % S2];
                            tp[PLANE] = A1[(j / R) % S2];
                            tp += PH;
                        }
                        A0[(j / R) % S2] = A1[(j / R) % S2] = (T)0;
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 2: gather along dim 1
    for (int kk0 = kc; kk0 < c1 - c0; kk0 += CT) {
        const int k1 = c0 + kk0;
        if (kk0 != kc) {   // TD1 > CT: further passes look their taps up now
            pe0 = a.bp1[k1], pe1 = a.bp1[k1 + S1];
            cnt = taps(k1, pe0, pe1, tw);
        }
        const int e0 = pe0 - r1lo, e1 = pe1 - r1lo;
        const T *rowbase = sT + sl * RS * PH;
        T acc[RS];
#pragma unroll
        for (int r = 0; r < RS; ++r) acc[r] = (T)0;
        const int q0 = e0 & 1;
        const T *pE = rowbase + q0 * PLANE + (e0 >> 1);
        const T *pO = rowbase + (1 - q0) * PLANE + (e0 >> 1) + q0;
#pragma unroll
        for (int k = 0; k < ADJ_MAXE; ++k) {
            if (k < cnt) {
                const T *p = (k & 1) ? pO + (k - 1) / 2 : pE + k / 2;
#pragma unroll
                for (int r = 0; r < RS; ++r) acc[r] = fmaT(tw[k], p[r * PH], acc[r]);
            }
        }
        for (int e = e0 + cnt; e < e1; ++e) {   // edge columns (clamped taps) and very long lists
            const int b = a.base1[r1lo + e];
            int slo = k1 - b, shi = k1 - b;   // taps s with clamp(b + s) == k1
            if (k1 == 0) slo = 0;
            if (k1 == a.m1 - 1) shi = S1 - 1;
            slo = max(slo, 0);
            shi = min(shi, S1 - 1);
            const T *p = rowbase + (e & 1) * PLANE + (e >> 1);
            for (int s = slo; s <= shi; ++s) {
                const T ws = a.W1[(int64_t)(r1lo + e) * S1 + s];
#pragma unroll
                for (int r = 0; r < RS; ++r) acc[r] = fmaT(ws, p[r * PH], acc[r]);
            }
        }
        T *d = dst + (int64_t)(d0 + sl * RS) * a.m1 + k1;
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            T *dr = d + (int64_t)r * a.m1;
            if constexpr (BETA)
                *dr = fmaT(a.beta, *dr, a.alpha * acc[r]);
            else
                *dr = a.alpha * acc[r];
        }
    }
}


// ---------------------------------------------------------------------------------------------
BandCtl *&sep2d_band()
{
    static thread_local BandCtl *ctl = nullptr;
    return ctl;
}

int sep2d_num_sms()
{
    static thread_local std::map<int, int> cache;   // per device
    const int dev = current_device();
    int &n = cache[dev];
    if (!n && (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)) n = 148;
    return n;
}

template <typename T, int S1, int S2>
static int launch_adj(SepArgs<T> &a, const li_table *t1, const li_table *t2, int64_t nbatch, cudaStream_t st,
                      bool *done)
{
    constexpr int NG = ADJ_NG, PT = 2 * (ADJ_NT / NG), RG = ADJ_TD2 / NG;
    if (a.m2 < S2 || a.m1 < 2) return LI_OK;   // decline: the edge folding needs a full window inside the array
    // LI_B200_ADJ_DEBUG (timing probes): 1 = no side stream, 2 = skip the generic tile rows, 3 = tile kernels only
    const int dbg_mode = opt(OPT_ADJ_DEBUG);
    if (dbg_mode != 3) {   // first choice: the streaming kernel (li_sep2d_adjs.cu); it declines what it cannot stage
        int rcs = sep2d_adj_stream<T>(a, t1, t2, nbatch, st, done);
        if (rcs != LI_OK || *done) return rcs;
    }
    if (sep2d_band()) return LI_OK;   // (the tile kernels have no banded form: the host pipeline falls back)
    int rc = table_ensure_rowstart(t1, st);
    if (rc == LI_OK) rc = table_ensure_fold(t2, st);
    if (rc != LI_OK) return rc;
    a.bp1 = t1->rowstart, a.bp2 = t2->bpf;
    a.W2f = (const T *)t2->Wfold;
    // widest dst tile whose contributing src columns fit the staged width
    int TD1 = 0;
    const int cands[] = {248, 240, 224, 208, 192, 176, 160, 144, 128, 120, 112, 96, 80, 64, 56, 48, 32, 24, 16, 8};
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
    rc = table_adjf_rows(t2, RG, st, &r2);
    if (rc != LI_OK) return rc;
    if (r2 > 2048) return LI_OK;   // decline (heavy down-sampling along dim 2)
    const int R2 = std::max(r2, 1);
    a.tiles1 = (a.m1 + TD1 - 1) / TD1, a.tiles2 = (a.m2 + ADJ_TD2 - 1) / ADJ_TD2;
    a.vec_ok = (a.n1 % 2 == 0) && (((uintptr_t)a.src) % (2 * sizeof(T)) == 0);
    const size_t smem = (2 * ((size_t)ADJ_TD2 * (PT / 2) + 16) + (size_t)ADJ_D * ADJ_NT * 2 + (size_t)PT * S1 +
                         (size_t)NG * R2 * S2) * sizeof(T) +
                        ((size_t)PT + (size_t)NG * (RG + S2)) * sizeof(int);
    if (smem > (size_t)kSmemLimit) return LI_OK;
    const int64_t blocks = (int64_t)a.tiles1 * a.tiles2 * nbatch;
    LI_REQUIRE(blocks <= 0x7fffffffLL, LI_ERR_UNSUPPORTED, "too many tiles");
    // regular tile rows (li_table.cu: table_tile_rows) go to the statically scheduled kernel, the others (array
    // edges, irregular rates) to the generic one: two launches over disjoint lists of tile rows
    const int32_t *rowlist = nullptr;
    int nreg = 0;
    rc = table_tile_rows(t2, ADJ_TD2, st, &rowlist, &nreg);
    if (rc != LI_OK) return rc;
    const int R = t2->reg_R;
    const bool vec4 = (a.n1 % 2 == 0) && (((uintptr_t)a.src) % (2 * sizeof(T)) == 0);
    if ((R != 1 && R != 2) || !vec4) nreg = 0;
    const int ntile_rows = a.tiles2;
    const bool beta = a.beta != (T)0;
    const size_t smem_reg = (2 * ((size_t)ADJ_TD2 * (PT / 2) + 16) + (size_t)NG * 2 * (RG + S2 - 1) * S2) * sizeof(T);
    struct Side {
        cudaStream_t side = nullptr;
        cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    };
    static thread_local std::map<int, Side> sides;   // per device
    Side &sd = sides[current_device()];
    cudaStream_t &side = sd.side;
    cudaEvent_t &ev_fork = sd.ev_fork, &ev_join = sd.ev_join;
    const bool fork = nreg > 0 && ntile_rows - nreg > 0 && dbg_mode == 0;
    if (ntile_rows - nreg > 0 && dbg_mode != 2) {
        // the few irregular tile rows run concurrently with the regular ones on a side stream (fork / join with
        // events), otherwise their short launch would serialise behind the main kernel
        cudaStream_t gst = st;
        if (fork) {
            if (!side) {
                // highest priority: its few long-running CTAs must be scheduled as soon as slots free up, not
                // behind the thousands of queued CTAs of the regular kernel
                int prio_lo = 0, prio_hi = 0;
                LI_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
                LI_CUDA(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, prio_hi));
                LI_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
                LI_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
            }
            LI_CUDA(cudaEventRecord(ev_fork, st));   // (before the regular launch: the two kernels overlap)
            LI_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
            gst = side;
        }
        a.rowlist = rowlist + nreg;
        a.tiles2 = ntile_rows - nreg;
        const unsigned gblocks = (unsigned)((int64_t)a.tiles1 * a.tiles2 * nbatch);
        if (beta) {
            auto kern = sep2d_adj_kernel<T, S1, S2, true>;
            rc = set_smem(kern, smem);
            if (rc != LI_OK) return rc;
            kern<<<gblocks, ADJ_NT, smem, gst>>>(a, TD1, R2);
        } else {
            auto kern = sep2d_adj_kernel<T, S1, S2, false>;
            rc = set_smem(kern, smem);
            if (rc != LI_OK) return rc;
            kern<<<gblocks, ADJ_NT, smem, gst>>>(a, TD1, R2);
        }
        count_launch();
        if (fork) LI_CUDA(cudaEventRecord(ev_join, side));
    }
#define LI_ADJ_REG(B, RR)                                                                     \
    if (beta == B && R == RR && nreg > 0) {                                                   \
        auto kern = sep2d_adj_regular_kernel<T, S1, S2, B, RR>;                               \
        rc = set_smem(kern, smem_reg);                                                        \
        if (rc != LI_OK) return rc;                                                           \
        a.rowlist = rowlist;                                                                  \
        a.tiles2 = nreg;                                                                      \
        kern<<<(unsigned)((int64_t)a.tiles1 * nreg * nbatch), ADJ_NT, smem_reg, st>>>(a, TD1); \
        count_launch();                                                                       \
    }
    LI_ADJ_REG(false, 1) LI_ADJ_REG(false, 2) LI_ADJ_REG(true, 1) LI_ADJ_REG(true, 2)
#undef LI_ADJ_REG
    if (fork) LI_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
    LI_CUDA(cudaGetLastError());
    *done = true;
    return LI_OK;
}

#define LI_MORE_DECL(NAME)                                                                                          \
    template <typename T>                                                                                           \
    int NAME(int adjoint, SepArgs<T> &a, const li_table *t1, const li_table *t2, int64_t nbatch, cudaStream_t st,   \
             bool *done, bool *mine);
LI_MORE_DECL(sep2d_more_a) LI_MORE_DECL(sep2d_more_b) LI_MORE_DECL(sep2d_more_c) LI_MORE_DECL(sep2d_more_d)
#undef LI_MORE_DECL

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
    const int S1 = t1->S, S2 = t2->S;
#define LI_SEP(A1, A2)                                                                   \
    if (S1 == A1 && S2 == A2)                                                            \
        return adjoint ? launch_adj<T, A1, A2>(a, t1, t2, nbatch, st, done)             \
                       : launch_fwd<T, A1, A2>(a, t1, t2, nbatch, st, done);
    LI_SEP(2, 2) LI_SEP(4, 4) LI_SEP(6, 6) LI_SEP(2, 4) LI_SEP(4, 2) LI_SEP(4, 6) LI_SEP(6, 4) LI_SEP(3, 3)
    LI_SEP(2, 1) LI_SEP(3, 1) LI_SEP(4, 1) LI_SEP(6, 1)   // (second factor = identity: 1-D applies along dim 1)
#undef LI_SEP
    // the other pairs of {2, 3, 4, 6, 8}^2: streaming kernels only (li_sep2d_more.cuh)
    bool mine = false;
    int rc = sep2d_more_a<T>(adjoint, a, t1, t2, nbatch, st, done, &mine);
    if (!mine) rc = sep2d_more_b<T>(adjoint, a, t1, t2, nbatch, st, done, &mine);
    if (!mine) rc = sep2d_more_c<T>(adjoint, a, t1, t2, nbatch, st, done, &mine);
    if (!mine) rc = sep2d_more_d<T>(adjoint, a, t1, t2, nbatch, st, done, &mine);
    if (!mine && opt(OPT_TRACE))
        fprintf(stderr, "[li_b200] fused 2-D: no kernel for the size pair (%d, %d): per-dimension passes\n", S1, S2);
    return rc;
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
    if (opt(OPT_NO_FUSED)) return LI_OK;
    return t1->dtype == LI_F32 ? fused_T<float>(adjoint, t1, t2, nbatch, alpha, src, beta, dst, st, done)
                               : fused_T<double>(adjoint, t1, t2, nbatch, alpha, src, beta, dst, st, done);
}


// Host-buffer pipeline of one fused 2-D apply (beta = 0, one image): the chunks of the launch are grouped into bands
// along dim 2; the src rows a band needs go up on a private copy stream while earlier bands compute, and every band's
// dst rows go down as soon as its kernel has run, so uploads, kernels and downloads of one apply overlap (and two
// applies on two caller streams keep both copy engines busy).  dsrc / ddst are full-size device staging buffers.
// *done stays false when the problem has no banded form (the caller then does upload - apply - download).
int sep2d_host_banded(int adjoint, const li_table *t1, const li_table *t2, double alpha, const void *src_host, void *dst_host,
                      void *dsrc, void *ddst, cudaStream_t st, bool *done)
{
    *done = false;
    constexpr int NBMAX = 8;
    if (opt(OPT_NO_BANDS) == 1 || !sep2d_fused_supported(t1, t2)) return LI_OK;   // (values 2..8: that many bands)
    struct Guard {
        ~Guard() { sep2d_band() = nullptr; }
    } guard;
    BandCtl bc;
    bc.mode = 1;
    sep2d_band() = &bc;
    bool planned = false;
    int rc = sep2d_fused_apply(adjoint, t1, t2, 1, alpha, dsrc, 0.0, ddst, st, &planned);
    if (rc != LI_OK || !planned) return rc;
    const int nchunks = (int)bc.src_end.size();
    const int NB = std::min(opt(OPT_NO_BANDS) >= 2 ? std::min(opt(OPT_NO_BANDS), NBMAX) : NBMAX, nchunks);
    if (NB < 2 || (int)bc.row_begin.size() != nchunks + 1) return LI_OK;
    const size_t es = t1->dtype == LI_F32 ? 4 : 8;
    const int64_t src_rows = adjoint ? t2->nrows : t2->ncols, src_w = adjoint ? t1->nrows : t1->ncols;
    const int64_t dst_w = adjoint ? t1->ncols : t1->nrows;
    const size_t src_rb = (size_t)src_w * es, dst_rb = (size_t)dst_w * es;

    struct Aux {
        cudaStream_t sh = nullptr;
        cudaEvent_t start = nullptr, ev[NBMAX] = {};
    };
    static thread_local std::map<cudaStream_t, Aux> auxmap;
    Aux &ax = auxmap[st];
    if (!ax.sh) {
        LI_CUDA(cudaStreamCreateWithFlags(&ax.sh, cudaStreamNonBlocking));
        LI_CUDA(cudaEventCreateWithFlags(&ax.start, cudaEventDisableTiming));
        for (int b = 0; b < NBMAX; ++b) LI_CUDA(cudaEventCreateWithFlags(&ax.ev[b], cudaEventDisableTiming));
    }
    // uploads: after everything already queued on the caller's stream (the staging buffers may still be in use)
    LI_CUDA(cudaEventRecord(ax.start, st));
    LI_CUDA(cudaStreamWaitEvent(ax.sh, ax.start, 0));
    int64_t copied = 0;
    for (int b = 0; b < NB; ++b) {
        const int c1 = (int)((int64_t)nchunks * (b + 1) / NB);
        const int64_t need = std::min<int64_t>(std::max<int64_t>(bc.src_end[c1 - 1], copied), src_rows);
        if (need > copied)
            LI_CUDA(cudaMemcpyAsync((char *)dsrc + copied * src_rb, (const char *)src_host + copied * src_rb,
                                    (size_t)(need - copied) * src_rb, cudaMemcpyHostToDevice, ax.sh));
        copied = need;
        LI_CUDA(cudaEventRecord(ax.ev[b], ax.sh));
    }
    // bands: kernel, then download, in order on the caller's stream
    for (int b = 0; b < NB; ++b) {
        const int c0 = (int)((int64_t)nchunks * b / NB), c1 = (int)((int64_t)nchunks * (b + 1) / NB);
        LI_CUDA(cudaStreamWaitEvent(st, ax.ev[b], 0));
        bc.mode = 2, bc.chunk0 = c0, bc.nchunks = c1 - c0;
        bool ok = false;
        rc = sep2d_fused_apply(adjoint, t1, t2, 1, alpha, dsrc, 0.0, ddst, st, &ok);
        if (rc == LI_OK && !ok) {
            set_error("internal: the fused kernel declined a band it had planned");
            rc = LI_ERR_UNSUPPORTED;
        }
        if (rc != LI_OK) return rc;
        const int64_t r0 = bc.row_begin[c0], r1 = bc.row_begin[c1];
        if (r1 > r0)
            LI_CUDA(cudaMemcpyAsync((char *)dst_host + r0 * dst_rb, (const char *)ddst + r0 * dst_rb, (size_t)(r1 - r0) * dst_rb,
                                    cudaMemcpyDeviceToHost, st));
    }
    *done = true;
    return LI_OK;
}

}  // namespace li
