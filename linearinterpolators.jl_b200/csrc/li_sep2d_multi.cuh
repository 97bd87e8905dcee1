// li_sep2d_multi.cuh -- adjoint of MANY fused 2-D separable operators accumulated into ONE model (K11):
//     model <- beta * model + alpha * sum_k (A1_k * A2_k)' * y_k
// the inner step of iterative fitting (README.md:36-37 "exploited for iterative fitting of data given an interpolated
// model"; the per-frame operator is the composition of test/benchmarks.jl:27-29, its factors the scatter loops of
// src/tabulated.jl:441-478).  The reference runs one `apply!(1, Adjoint, H_k, y_k, false, 1, model)` per frame.
//
// One persistent launch instead of one launch per frame.  A CTA owns a tile of the model -- a strip of NT columns x a
// chunk of rows -- for ALL frames of its frame group and keeps the tile's accumulators in shared memory.  For every
// frame it runs the data flow of sep2d_adj_stream_kernel (li_sep2d_adjs.cuh): the src rows of that frame that feed the
// tile stream through a ring of TMA bulk copies, a thread gathers along dim 1 from the staged row with its folded
// taps in registers, adds the value into S2 rotating accumulators with the row's folded dim-2 weights, and every model
// row that is complete for this frame is added to the shared-memory tile (one LDS + FADD + STS per row, owner thread
// only: no atomics).  The ring does not drain between frames: while the last stage of frame k is consumed the first
// stage of frame k+1 is already in flight, and the per-frame scalars (window of src rows / columns of the tile) are
// loaded one frame ahead.  The tile leaves the CTA once, at the end.  Frames are visited in ascending order and frame
// groups (used when the tiles alone cannot fill the machine) are combined in ascending order by the last CTA of a
// tile to finish, so the result is run-to-run deterministic.
#pragma once
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>
#include <vector>

#include "li_sep2d.cuh"
#include "li_sep2d_adjs.cuh"

namespace li {

template <typename T>
struct MultiFrame {                       // tables of one frame (device array, cached per list of tables)
    const int32_t *base1, *bp1, *bfold1;  // dim-1 table: canonical bases, bptr over them (rowstart), folded bases
    const T *W1f;                         // folded dim-1 weights
    const int32_t *base2, *bp2, *irrp2;   // dim-2 table: canonical bases, bptr over FOLDED bases, irregularity sums
    const T *W2f;                         // folded dim-2 weights
    int32_t n1, n2;
};

template <typename T>
struct MultiArgs {
    const MultiFrame<T> *frames;
    const T *const *srcs;                 // [nframes] the frames themselves, [n1, n2] each
    const int4 *tiles;                    // {c0, va, vb, 0}: columns [c0, c0 + NT) x rows [va, vb) of the model
    int ntiles, nframes, G;               // grid = ntiles * G; group g owns frames [g*nframes/G, (g+1)*nframes/G)
    int m1, m2;
    int P1, CRMAX, NRMAX;                 // staged row pitch; longest tile; most src rows feeding one tile
    T alpha, beta;
    T *dst;
    T *partial;                           // G > 1: [G][m1*m2] partial models of the groups
    unsigned *counters;                   // G > 1: [ntiles] arrival counters (zero on entry, zero again on exit)
};

constexpr int ADM_NT = 256;

template <typename T, int S1, int S2, int R>
__global__ void __launch_bounds__(ADM_NT, sizeof(T) == 4 ? 3 : 2)
sep2d_adj_multi_kernel(const MultiArgs<T> a)
{
    constexpr int NT = ADM_NT, K = AdsK<S2>::value, NSTAGE = 2;
    constexpr int V = 2;                                         // LDS.64 (float) / LDS.128 (double)
    constexpr int NU_S = (S1 + 2 + 1) & ~1, NU_L = 2 * S1 + 2;   // staged values per thread: usual / wide tap lists
    static_assert(R == 0 || (K % R == 0 && (K / R) % S2 == 0), "bad stage size");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P1 = a.P1;
    T *sS = reinterpret_cast<T *>(smem_raw);                       // [NSTAGE][K][P1] ring of src rows
    T *sAcc = sS + (size_t)NSTAGE * K * P1;                        // [CRMAX][NT] the tile's accumulators
    T *sW2 = sAcc + (size_t)a.CRMAX * NT;                          // [NRMAX][S2] alpha * folded dim-2 weights
    int *sNv = reinterpret_cast<int *>(sW2 + (size_t)a.NRMAX * S2);   // [CRMAX + 2*S2] src rows starting at a dst row
    uint64_t *bar = reinterpret_cast<uint64_t *>(sNv + ((a.CRMAX + 2 * S2 + 1) & ~1));
    __shared__ unsigned s_ticket;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x % a.ntiles, grp = blockIdx.x / a.ntiles;
    const int4 tl = a.tiles[tile];
    const int c0 = tl.x, va = tl.y, vb = tl.z;
    const int c1 = min(c0 + NT, a.m1);
    const int v0 = va - (S2 - 1);
    const int NV = vb - v0;
    const int f0 = (int)((int64_t)a.nframes * grp / a.G), f1 = (int)((int64_t)a.nframes * (grp + 1) / a.G);
    constexpr int AL = 16 / (int)sizeof(T);

    for (int k = tid; k < (vb - va) * NT; k += NT) sAcc[k] = (T)0;
    if (tid == 0) {
#pragma unroll
        for (int sgi = 0; sgi < NSTAGE; ++sgi) mbar_init(&bar[sgi], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // per-frame scalars of this tile (block-uniform)
    struct Geo {
        const T *src;
        int n1, ulo_a, cv1, i2lo, NR, NJ;
    };
    auto load_geo = [&](int f) {
        Geo g;
        const MultiFrame<T> &fd = a.frames[f];
        g.src = a.srcs[f];
        g.n1 = fd.n1;
        const int u_lo = c0 == 0 ? 0 : fd.bp1[c0];               // first src column with a tap in the strip
        g.ulo_a = u_lo & ~(AL - 1);
        g.cv1 = min(P1, fd.n1 - g.ulo_a);
        g.i2lo = fd.bp2[va];
        g.NR = fd.bp2[vb - 1 + S2] - g.i2lo;
        g.NJ = (g.NR + K - 1) / K;
        return g;
    };
    unsigned gj = 0;                                             // stages consumed so far, over all frames
    auto issue = [&](const Geo &g, int j, unsigned gidx) {       // one thread: stage j of the frame -> ring slot gidx
        uint64_t *b = &bar[gidx % NSTAGE];
        T *buf = sS + (size_t)(gidx % NSTAGE) * K * P1;
        const uint32_t bytes = (uint32_t)(g.cv1 * sizeof(T));
        const int nr = min(K, g.NR - j * K);
        mbar_expect_tx(b, bytes * nr);
        for (int r = 0; r < nr; ++r)
            bulk_g2s(buf + r * P1, g.src + (int64_t)(g.i2lo + j * K + r) * g.n1 + g.ulo_a, bytes, b);
    };

    using Bits = typename std::conditional<sizeof(T) == 4, uint32_t, uint64_t>::type;
    auto keep = [](T x, Bits m) {
        Bits b;
        memcpy(&b, &x, sizeof b);
        b &= m;
        memcpy(&x, &b, sizeof b);
        return x;
    };

    Geo geo;
    geo.NJ = 0;
    if (f0 < f1) geo = load_geo(f0);
    bool issued = false;                                         // stage 0 of the current frame is already in flight
    const int c = c0 + tid;
    const bool have = c < c1;
    const int cc = min(c, a.m1 - 1);

    for (int f = f0; f < f1; ++f) {
        const MultiFrame<T> &fd = a.frames[f];
        if (tid == 0 && !issued && geo.NJ > 0) issue(geo, 0, gj);
        Geo nxt;
        nxt.NJ = 0;
        if (f + 1 < f1) nxt = load_geo(f + 1);                   // (loads in flight while this frame computes)

        // ---- per-thread dim-1 taps of this frame: src columns [e0, e1) feed dst column c
        const int e0 = cc == 0 ? 0 : fd.bp1[cc], e1 = cc == a.m1 - 1 ? fd.n1 : fd.bp1[cc + S1];
        const int a0 = e0 & ~(V - 1);
        T w1[NU_L];
        unsigned vm = 0;
#pragma unroll
        for (int t = 0; t < NU_L; ++t) {
            const int i = a0 + t;
            w1[t] = (T)0;
            if (have && i >= e0 && i < e1) {
                const int b = fd.base1[i];
                if (cc >= clampi(b, a.m1 - 1) && cc <= clampi(b + S1 - 1, a.m1 - 1)) {
                    w1[t] = fd.W1f[(int64_t)i * S1 + (cc - fd.bfold1[i])];
                    vm |= 1u << t;
                }
            }
        }
        // narrow warps: every lane's taps fit the first NU_S staged values (the host guarantees e1 - a0 <= NU_L)
        const bool narrow = __all_sync(0xffffffffu, !have || e1 - a0 <= NU_S);
        const int off = a0 - geo.ulo_a;
        const int i2lo = geo.i2lo, NR = geo.NR, NJ = geo.NJ;
        bool regular = false;
        if constexpr (R > 0) regular = v0 > 0 && vb <= a.m2 - S2 && fd.irrp2[vb] == fd.irrp2[v0];

        __syncthreads();   // everybody is done with the previous frame's sW2 / sNv
        if (!regular)
            for (int k = tid; k < a.CRMAX + 2 * S2; k += NT) sNv[k] = k < NV ? fd.bp2[va + k + 1] - fd.bp2[va + k] : 0;
        for (int k = tid; k < NR * S2; k += NT) sW2[k] = a.alpha * fd.W2f[(int64_t)i2lo * S2 + k];
        __syncthreads();

        T acc[S2];
#pragma unroll
        for (int s = 0; s < S2; ++s) acc[s] = (T)0;
        T *ap = sAcc + tid;                                       // accumulator of the next model row to complete
        int rowstep = P1;
        asm volatile("" : "+r"(vm), "+r"(rowstep));   // (opaque: keeps vm a bit mask instead of NU flags)

        auto next_stage = [&](int j) {   // block-uniform; returns this thread's window in the first row of stage j
            __syncthreads();             // every thread is done with stage j-1: its buffer may be refilled
            if (tid == 0) {
                if (j + 1 < NJ)
                    issue(geo, j + 1, gj + j + 1);
                else if (nxt.NJ > 0)
                    issue(nxt, 0, gj + NJ);                      // first stage of the next frame
            }
            const unsigned gidx = gj + j;
            mbar_wait(&bar[gidx % NSTAGE], (gidx / NSTAGE) & 1u);
            return (const T *)(sS + (size_t)(gidx % NSTAGE) * K * P1 + off);
        };
        auto emit = [&](int vv, T value) {   // model row vv is complete for this frame
            if (vv >= va) {
                *ap += value;
                ap += NT;
            }
        };

        // The row loops, instantiated for (NU = staged values per thread, CLEAN = padding cleared with bit masks at
        // the two ends of the window / with one select per tap).
        auto run = [&](auto nu_tag, auto clean_tag) {
            constexpr int NU = decltype(nu_tag)::value;
            constexpr bool CLEAN = decltype(clean_tag)::value;
            constexpr int NTAIL = NU <= 6 ? NU - 1 : NU / 2;
            Bits k0 = (vm & 1u) ? ~(Bits)0 : (Bits)0;
            Bits kt[NTAIL];
#pragma unroll
            for (int t = 0; t < NTAIL; ++t) kt[t] = ((vm >> (NU - NTAIL + t)) & 1u) ? ~(Bits)0 : (Bits)0;
            if constexpr (sizeof(T) == 4) {
                asm volatile("" : "+r"(k0));
#pragma unroll
                for (int t = 0; t < NTAIL; ++t) asm volatile("" : "+r"(kt[t]));
            } else {
                asm volatile("" : "+l"(k0));
#pragma unroll
                for (int t = 0; t < NTAIL; ++t) asm volatile("" : "+l"(kt[t]));
            }
            auto gather = [&](const T *p) {
                T x[NU];
#pragma unroll
                for (int t = 0; t < NU; t += V) {
                    const Pack<T, V> q = ld_pack<T, V>(p + t);
#pragma unroll
                    for (int e = 0; e < V; ++e) x[t + e] = q.v[e];
                }
                if constexpr (CLEAN) {
                    x[0] = keep(x[0], k0);
#pragma unroll
                    for (int t = 0; t < NTAIL; ++t) x[NU - NTAIL + t] = keep(x[NU - NTAIL + t], kt[t]);
                } else {
                    unsigned m = vm;
                    asm volatile("" : "+r"(m));
#pragma unroll
                    for (int t = 0; t < NU; ++t) x[t] = ((m >> t) & 1u) ? x[t] : (T)0;
                }
                if constexpr (sizeof(T) == 4) {
                    float2 acc2 = __fmul2_rn(make_float2(x[0], x[1]), make_float2(w1[0], w1[1]));
#pragma unroll
                    for (int t = 2; t < NU; t += 2)
                        acc2 = __ffma2_rn(make_float2(x[t], x[t + 1]), make_float2(w1[t], w1[t + 1]), acc2);
                    return (T)(acc2.x + acc2.y);
                } else {
                    T va0 = x[0] * w1[0], va1 = x[1] * w1[1];
#pragma unroll
                    for (int t = 2; t < NU; t += 2) {
                        va0 = fmaT(x[t], w1[t], va0);
                        va1 = fmaT(x[t + 1], w1[t + 1], va1);
                    }
                    return va0 + va1;
                }
            };
            if (regular) {
                if constexpr (R > 0) {   // stage j holds the R rows of each of the model rows v0 + j*DR .. + DR-1
                    constexpr int DR = K / R;
                    const T *pw = sW2;
                    for (int j = 0; j < NJ; ++j) {
                        const T *prow = next_stage(j);
#pragma unroll
                        for (int d = 0; d < DR; ++d) {
                            const int vv = v0 + j * DR + d;
                            if (vv < vb) {   // block-uniform: only the last stage is partial
#pragma unroll
                                for (int r = 0; r < R; ++r) {
                                    const T val = gather(prow);
                                    prow += rowstep;
                                    T w2[S2];
                                    ld_weights<T, S2>(pw + (d * R + r) * S2, w2);
#pragma unroll
                                    for (int s = 0; s < S2; ++s) acc[(d + s) % S2] = fmaT(w2[s], val, acc[(d + s) % S2]);
                                }
                                emit(vv, acc[d % S2]);
                                acc[d % S2] = (T)0;
                            }
                        }
                        pw += K * S2;
                    }
                }
            } else {   // generic schedule: for model row vv, every src row whose folded window starts at vv
                const T *pw = sW2;
                const int *pn = sNv;
                const T *prow = sS + off;
                int left = 0, j = -1, k = 0;
                auto row = [&]() {
                    if (left == 0) {
                        prow = next_stage(++j);
                        left = K;
                    }
                    --left;
                    const T val = gather(prow);
                    prow += rowstep;
                    return val;
                };
                for (int v = v0; v < vb; v += S2) {
#pragma unroll
                    for (int u = 0; u < S2; ++u) {
                        const int vv = v + u;
                        if (vv < vb) {   // block-uniform
                            int n = pn[u];
                            if (vv > 0 && vv < a.m2 - S2) {   // interior windows: every slot is fed
                                k += n;
#pragma unroll 1
                                for (; n > 0; --n) {
                                    const T val = row();
                                    T w2[S2];
                                    ld_weights<T, S2>(pw, w2);
                                    pw += S2;
#pragma unroll
                                    for (int s = 0; s < S2; ++s) acc[(u + s) % S2] = fmaT(w2[s], val, acc[(u + s) % S2]);
                                }
                            } else {   // first / last window of the array: only the slots the clamped taps land on
#pragma unroll 1
                                for (; n > 0; --n) {
                                    const T val = row();
                                    T w2[S2];
                                    ld_weights<T, S2>(pw, w2);
                                    pw += S2;
                                    const int b = fd.base2[i2lo + k];
                                    ++k;
                                    const int lo = clampi(b, a.m2 - 1) - vv, hi = clampi(b + S2 - 1, a.m2 - 1) - vv;
#pragma unroll
                                    for (int s = 0; s < S2; ++s)
                                        if (s >= lo && s <= hi) acc[(u + s) % S2] = fmaT(w2[s], val, acc[(u + s) % S2]);
                                }
                            }
                            emit(vv, acc[u]);
                            acc[u] = (T)0;
                        }
                    }
                    pn += S2;
                }
            }
        };
        // (every copy runs the same sequence of barriers, so the warps of a CTA may take different copies)
        if (narrow) {
            constexpr int NTAIL = NU_S <= 6 ? NU_S - 1 : NU_S / 2;
            constexpr unsigned MID = ((1u << (NU_S - NTAIL)) - 1u) & ~1u;
            if (__all_sync(0xffffffffu, (vm & MID) == MID))
                run(std::integral_constant<int, NU_S>{}, std::true_type{});
            else
                run(std::integral_constant<int, NU_S>{}, std::false_type{});
        } else {
            constexpr int NTAIL = NU_L / 2;
            constexpr unsigned MID = ((1u << (NU_L - NTAIL)) - 1u) & ~1u;
            if (__all_sync(0xffffffffu, (vm & MID) == MID))
                run(std::integral_constant<int, NU_L>{}, std::true_type{});
            else
                run(std::integral_constant<int, NU_L>{}, std::false_type{});
        }
        gj += (unsigned)NJ;
        issued = NJ > 0 && nxt.NJ > 0;
        geo = nxt;
    }
    __syncthreads();

    // ---- the tile leaves the CTA
    const int64_t plane = (int64_t)a.m1 * a.m2;
    if (a.G == 1) {
        if (have) {
            T *dp = a.dst + (int64_t)va * a.m1 + c;
            const T *ap = sAcc + tid;
            if (a.beta != (T)0) {
                for (int r = 0; r < vb - va; ++r) dp[(int64_t)r * a.m1] = fmaT(a.beta, dp[(int64_t)r * a.m1], ap[r * NT]);
            } else {   // dst is never read
                for (int r = 0; r < vb - va; ++r) dp[(int64_t)r * a.m1] = ap[r * NT];
            }
        }
        return;
    }
    if (have) {
        T *pp = a.partial + grp * plane + (int64_t)va * a.m1 + c;
        const T *ap = sAcc + tid;
        for (int r = 0; r < vb - va; ++r) pp[(int64_t)r * a.m1] = ap[r * NT];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&a.counters[tile], 1u);
    __syncthreads();
    if (s_ticket != (unsigned)(a.G - 1)) return;
    __threadfence();   // the last group to arrive adds the partial tiles up, in group order
    if (have) {
        T *dp = a.dst + (int64_t)va * a.m1 + c;
        const T *pp = a.partial + (int64_t)va * a.m1 + c;
        for (int r = 0; r < vb - va; ++r) {
            T s = __ldcg(pp + (int64_t)r * a.m1);
            for (int g = 1; g < a.G; ++g) s += __ldcg(pp + g * plane + (int64_t)r * a.m1);
            dp[(int64_t)r * a.m1] = a.beta != (T)0 ? fmaT(a.beta, dp[(int64_t)r * a.m1], s) : s;
        }
    }
    if (tid == 0) a.counters[tile] = 0u;
}

// ---------------------------------------------------------------------------------------------
// Host side.

// per frame, over the strips / tiles of the launch: longest dim-1 tap list of a column, widest span of src columns
// feeding a strip, most src rows feeding a tile (one launch for the whole batch; results by atomicMax)
template <typename T>
__global__ void multi_plan_kernel(const MultiFrame<T> *__restrict__ frames, const int4 *__restrict__ tiles, int ntiles, int m1,
                                  int S1, int S2, int NT, int *__restrict__ out /* [3] */)
{
    const MultiFrame<T> fd = frames[blockIdx.x];
    int maxcount = 0, span = 0, nr = 0;
    for (int k = threadIdx.x; k < m1; k += blockDim.x) {
        const int lo = k == 0 ? 0 : fd.bp1[k], hi = k == m1 - 1 ? fd.n1 : fd.bp1[k + S1];
        maxcount = max(maxcount, hi - lo);
    }
    for (int k0 = threadIdx.x * NT; k0 < m1; k0 += blockDim.x * NT) {
        const int k1 = min(k0 + NT, m1);
        const int lo = k0 == 0 ? 0 : fd.bp1[k0], hi = k1 == m1 ? fd.n1 : fd.bp1[k1 - 1 + S1];
        span = max(span, hi - lo);
    }
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
        const int4 tl = tiles[t];
        nr = max(nr, fd.bp2[tl.z - 1 + S2] - fd.bp2[tl.y]);
    }
    for (int o = 16; o; o >>= 1) {
        maxcount = max(maxcount, __shfl_xor_sync(0xffffffffu, maxcount, o));
        span = max(span, __shfl_xor_sync(0xffffffffu, span, o));
        nr = max(nr, __shfl_xor_sync(0xffffffffu, nr, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out + 0, maxcount);
        atomicMax(out + 1, span);
        atomicMax(out + 2, nr);
    }
}

struct MultiPlan {
    void *frames = nullptr;       // device MultiFrame<T>[n]
    int4 *tiles = nullptr;
    unsigned *counters = nullptr;
    int ntiles = 0, G = 1, P1 = 0, CRMAX = 0, NRMAX = 0, R = 0;
    size_t smem = 0;
    bool ok = false;              // false: the kernel cannot take this batch (cached too)
    uint64_t stamp = 0;
};
struct MultiKey {
    int dev, dtype, row_lo, row_hi;
    int64_t n;
    uint64_t h1, h2;
    bool operator<(const MultiKey &o) const
    {
        return std::tie(dev, dtype, row_lo, row_hi, n, h1, h2) < std::tie(o.dev, o.dtype, o.row_lo, o.row_hi, o.n, o.h1, o.h2);
    }
};
std::mutex &multi_mutex();
std::map<MultiKey, MultiPlan> &multi_plans();

template <typename T, int S1, int S2>
static int multi_build_plan(int64_t n, const li_table *const *tabs, int row_lo, int row_hi, cudaStream_t st, MultiPlan *plan)
{
    constexpr int NT = ADM_NT, AL = 16 / (int)sizeof(T), K = AdsK<S2>::value, NU_L = 2 * S1 + 2, V = 2;
    const int m1 = (int)tabs[0]->ncols, m2 = (int)tabs[1]->ncols;
    plan->ok = false;
    if (m1 < S1 || m2 < S2) return LI_OK;
    int R = -1;
    std::vector<MultiFrame<T>> fr((size_t)n);
    for (int64_t k = 0; k < n; ++k) {
        const li_table *t1 = tabs[2 * k], *t2 = tabs[2 * k + 1];
        if (t1->nrows % AL != 0) return LI_OK;   // bulk copies need 16-byte aligned row segments
        int rc = table_ensure_rowstart(t1, st);
        if (rc == LI_OK) rc = table_ensure_fold(t1, st);
        if (rc == LI_OK) rc = table_ensure_fold(t2, st);
        if (rc != LI_OK) return rc;
        const int r = (t2->irrp && (t2->reg_R == 1 || t2->reg_R == 2)) ? t2->reg_R : 0;
        R = R < 0 ? r : (R == r ? R : 0);
        MultiFrame<T> &f = fr[(size_t)k];
        f.base1 = t1->base, f.bp1 = t1->rowstart, f.bfold1 = t1->bfold, f.W1f = (const T *)t1->Wfold;
        f.base2 = t2->base, f.bp2 = t2->bpf, f.irrp2 = t2->irrp, f.W2f = (const T *)t2->Wfold;
        f.n1 = (int32_t)t1->nrows, f.n2 = (int32_t)t2->nrows;
    }
    plan->R = std::max(R, 0);
    // tiles: strips of NT columns x chunks of rows; one resident wave of CTAs; chunks of at least 4*S2 rows (each
    // chunk re-reads S2-1 halo rows per frame); frame groups when the tiles alone cannot fill the machine
    const int strips = (m1 + NT - 1) / NT, rows = row_hi - row_lo;
    const int p1_est = (NT + 4 * S1 + 2 * AL + NU_L + 7) & ~(AL - 1);
    auto smem_of = [&](int CR, int P1, int NRMAX) {
        return ((size_t)2 * K * P1 + (size_t)CR * NT + (size_t)NRMAX * S2) * sizeof(T) +
               (size_t)((CR + 2 * S2 + 1) & ~1) * sizeof(int) + 2 * sizeof(uint64_t);
    };
    int CR = 0, chunks = 0;
    for (int occ = sizeof(T) == 4 ? 3 : 2; occ >= 1; --occ) {
        const int target = sep2d_num_sms() * occ;
        chunks = std::max(1, std::min(target / strips, rows / std::max(8, 4 * S2)));
        CR = (rows + chunks - 1) / chunks;
        chunks = (rows + CR - 1) / CR;
        const size_t est = smem_of(CR, p1_est, CR + 4 * S2);
        if (est > (size_t)kSmemLimit) {   // accumulators too large: shorter chunks (more than one wave)
            const int cap = std::max(8, (int)(((size_t)kSmemLimit - smem_of(0, p1_est, 8 * S2) - 4096) / (NT * sizeof(T) + 4 + S2 * sizeof(T))));
            chunks = (rows + cap - 1) / cap;
            CR = (rows + chunks - 1) / chunks;
            chunks = (rows + CR - 1) / CR;
        }
        int got = 0;
#define LI_OCCQ(RR)                                                                                          \
    if (plan->R == RR) {                                                                                     \
        auto kern = sep2d_adj_multi_kernel<T, S1, S2, RR>;                                                   \
        const size_t sm = smem_of(CR, p1_est, CR + 4 * S2);                                                  \
        if (sm > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&got, kern, NT, sm) != cudaSuccess) got = 0;      \
    }
        LI_OCCQ(0) LI_OCCQ(1) LI_OCCQ(2)
#undef LI_OCCQ
        cudaGetLastError();
        if (got >= occ || occ == 1) {
            const int tiles = strips * chunks;
            plan->G = (int)std::max<int64_t>(1, std::min<int64_t>((target + tiles / 2) / tiles, n / 4));
            break;
        }
    }
    std::vector<int4> tl;
    for (int ch = 0; ch < chunks; ++ch)
        for (int sx = 0; sx < strips; ++sx)
            tl.push_back(make_int4(sx * NT, row_lo + ch * CR, std::min(row_lo + (ch + 1) * CR, row_hi), 0));
    plan->ntiles = (int)tl.size();
    plan->CRMAX = CR;
    int *d_out = nullptr;
    LI_CUDA(cudaMalloc(&plan->frames, (size_t)n * sizeof(MultiFrame<T>)));
    LI_CUDA(cudaMalloc(&plan->tiles, tl.size() * sizeof(int4)));
    LI_CUDA(cudaMalloc(&plan->counters, tl.size() * sizeof(unsigned)));
    LI_CUDA(cudaMalloc(&d_out, 3 * sizeof(int)));
    LI_CUDA(cudaMemcpyAsync(plan->frames, fr.data(), (size_t)n * sizeof(MultiFrame<T>), cudaMemcpyHostToDevice, st));
    LI_CUDA(cudaMemcpyAsync(plan->tiles, tl.data(), tl.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    LI_CUDA(cudaMemsetAsync(plan->counters, 0, tl.size() * sizeof(unsigned), st));
    LI_CUDA(cudaMemsetAsync(d_out, 0, 3 * sizeof(int), st));
    multi_plan_kernel<T><<<(unsigned)n, 128, 0, st>>>((const MultiFrame<T> *)plan->frames, plan->tiles, plan->ntiles, m1, S1, S2, NT, d_out);
    count_launch();
    int h_out[3] = {0, 0, 0};
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, d_out, sizeof h_out, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // (also: fr / tl may go out of scope)
    cudaFree(d_out);
    if (e != cudaSuccess) {
        set_error("multi-frame plan failed: %s", cudaGetErrorString(e));
        return LI_ERR_CUDA;
    }
    const int maxcount = h_out[0], span = h_out[1];
    if (maxcount + (V - 1) > NU_L) {   // (down-sampling along dim 1, or many columns hanging over an edge)
        if (opt(OPT_TRACE)) fprintf(stderr, "[li_b200] adj multi declines: %d taps per model column > %d\n", maxcount + V - 1, NU_L);
        return LI_OK;
    }
    plan->P1 = (span + (AL - 1) + NU_L + (AL - 1)) & ~(AL - 1);
    plan->NRMAX = (std::max(h_out[2], 1) + 1) & ~1;
    plan->smem = smem_of(CR, plan->P1, plan->NRMAX);
    if (opt(OPT_TRACE))
        fprintf(stderr, "[li_b200] adj multi: %lld frames, tiles %d x %d (CR %d) x G %d, P1 %d NRMAX %d smem %zu R %d\n", (long long)n,
                strips, chunks, CR, plan->G, plan->P1, plan->NRMAX, plan->smem, plan->R);
    if (plan->smem > (size_t)kSmemLimit) return LI_OK;
    plan->ok = true;
    return LI_OK;
}

template <typename T, int S1, int S2>
static int multi_T(int64_t n, const li_table *const *tabs, const void *const *src_dev, double alpha, double beta, void *dst,
                   int row_lo, int row_hi, cudaStream_t st, bool *done)
{
    int dev = 0;
    LI_CUDA(cudaGetDevice(&dev));
    MultiKey key{dev, tabs[0]->dtype, row_lo, row_hi, n, 1469598103934665603ull, 0x9e3779b97f4a7c15ull};
    for (int64_t k = 0; k < 2 * n; ++k) {   // two independent hashes of the table identities
        key.h1 = (key.h1 ^ tabs[k]->uid) * 1099511628211ull;
        key.h2 = (key.h2 + tabs[k]->uid) * 0xff51afd7ed558ccdull + (key.h2 >> 29);
    }
    MultiPlan plan;
    {
        std::lock_guard<std::mutex> lock(multi_mutex());
        auto &plans = multi_plans();
        static uint64_t clock = 0;
        auto it = plans.find(key);
        if (it == plans.end()) {
            if (plans.size() >= 16) {   // drop the least recently used plan
                auto old = plans.begin();
                for (auto jt = plans.begin(); jt != plans.end(); ++jt)
                    if (jt->second.stamp < old->second.stamp) old = jt;
                cudaStreamSynchronize(st);
                cudaFree(old->second.frames), cudaFree(old->second.tiles), cudaFree(old->second.counters);
                plans.erase(old);
            }
            MultiPlan p;
            int rc = multi_build_plan<T, S1, S2>(n, tabs, row_lo, row_hi, st, &p);
            if (rc != LI_OK || !p.ok) {
                cudaFree(p.frames), cudaFree(p.tiles), cudaFree(p.counters);
                p.frames = nullptr, p.tiles = nullptr, p.counters = nullptr;
                if (rc != LI_OK) return rc;
            }
            it = plans.emplace(key, p).first;
        }
        it->second.stamp = ++clock;
        plan = it->second;
    }
    if (!plan.ok) return LI_OK;
    for (int64_t k = 0; k < n; ++k)
        if (((uintptr_t)src_dev[k]) % 16 != 0) return LI_OK;
    const int m1 = (int)tabs[0]->ncols, m2 = (int)tabs[1]->ncols;
    // per call: the frame pointers (and the partial models of the frame groups)
    void *d_srcs = nullptr, *d_part = nullptr;
    int rc = staging(4, (size_t)n * sizeof(void *), &d_srcs, (void *)st);
    if (rc == LI_OK && plan.G > 1) rc = staging(5, (size_t)plan.G * m1 * m2 * sizeof(T), &d_part, (void *)st);
    if (rc != LI_OK) return rc;
    LI_CUDA(cudaMemcpyAsync(d_srcs, src_dev, (size_t)n * sizeof(void *), cudaMemcpyHostToDevice, st));   // (pageable: staged before return)
    MultiArgs<T> a;
    a.frames = (const MultiFrame<T> *)plan.frames;
    a.srcs = (const T *const *)d_srcs;
    a.tiles = plan.tiles;
    a.ntiles = plan.ntiles, a.nframes = (int)n, a.G = plan.G;
    a.m1 = m1, a.m2 = m2;
    a.P1 = plan.P1, a.CRMAX = plan.CRMAX, a.NRMAX = plan.NRMAX;
    a.alpha = (T)alpha, a.beta = (T)beta;
    a.dst = (T *)dst;
    a.partial = (T *)d_part;
    a.counters = plan.counters;
    const unsigned blocks = (unsigned)(plan.ntiles * plan.G);
#define LI_ADM(RR)                                                      \
    if (plan.R == RR) {                                                 \
        auto kern = sep2d_adj_multi_kernel<T, S1, S2, RR>;              \
        rc = set_smem(kern, plan.smem);                                 \
        if (rc != LI_OK) return rc;                                     \
        kern<<<blocks, ADM_NT, plan.smem, st>>>(a);                     \
    }
    LI_ADM(0) LI_ADM(1) LI_ADM(2)
#undef LI_ADM
    LI_CHECK_LAUNCH();
    *done = true;
    return LI_OK;
}

// model rows [row_lo, row_hi) <- beta * model + alpha * sum_k H_k' y_k; *done stays false when the kernel declines
// the batch (the caller then applies the frames one by one)
template <typename T>
int sep2d_adj_multi(int64_t n, const li_table *const *tabs, const void *const *src_dev, double alpha, double beta,
                    void *dst, int row_lo, int row_hi, cudaStream_t st, bool *done)
{
    const int S1 = tabs[0]->S, S2 = tabs[1]->S;
#define LI_SEP(A1, A2) \
    if (S1 == A1 && S2 == A2) return multi_T<T, A1, A2>(n, tabs, src_dev, alpha, beta, dst, row_lo, row_hi, st, done);
    LI_SEP(2, 2) LI_SEP(3, 3) LI_SEP(4, 4) LI_SEP(6, 6)
#undef LI_SEP
    return LI_OK;
}

}  // namespace li
