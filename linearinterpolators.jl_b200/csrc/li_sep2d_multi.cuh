// li_sep2d_multi.cuh -- adjoint of MANY fused 2-D separable operators accumulated into ONE model (K11):
//     model <- beta * model + alpha * sum_k (A1_k * A2_k)' * y_k
// the inner step of iterative fitting (README.md:36-37 "exploited for iterative fitting of data given an interpolated
// model"; the per-frame operator is the composition of test/benchmarks.jl:27-29, its factors the scatter loops of
// src/tabulated.jl:441-478).  The reference runs one `apply!(1, Adjoint, H_k, y_k, false, 1, model)` per frame.
//
// One persistent, warp-specialised launch instead of one launch per frame.  A CTA owns a tile of the model -- a strip
// of 512 columns x a chunk of rows -- for ALL frames of its frame group and keeps the tile's accumulators in shared
// memory.  Eight consumer warps: a thread owns TWO adjacent model columns (its staged window along dim 1 is shared by
// both: half the shared-memory reads of a one-column thread, and the pairs are what the packed FP32x2 instructions
// take).  One producer warp runs ahead of them: for every frame it writes the frame's dim-2 weights and row counts
// of the tile into a double-buffered block of shared memory and streams the src rows of that frame that feed the tile
// through a ring of TMA bulk copies (full / empty mbarriers per ring slot; the ring never drains between frames).
// Per src row a consumer thread gathers along dim 1 from the staged row (taps of its two columns in registers,
// transposed and folded once per plan, prefetched one frame ahead), adds the pair into S2 rotating pair accumulators
// with the row's folded dim-2 weights, and every model row that is complete for this frame is added to the tile in
// shared memory (LDS.64 + FADD2 + STS.64, owner thread only: no atomics).  The tile leaves the CTA once, at the end.
// Frames are visited in ascending order and frame groups (used when the tiles alone cannot fill the machine) are
// combined in ascending order by the last CTA of a tile to finish, so the result is run-to-run deterministic.
// Values outside a column's tap list are cleared with bit masks / selects, never multiplied by zero: a non-finite
// sample reaches exactly the model elements it feeds in the reference.
// The first and the last column of the model collect every src column hanging over that edge (Flat / SafeFlat
// clamping), so the pair holding them can have a long tap list.  A ninth consumer warp (the overflow warp) takes the
// part of those lists beyond the usual window, cut into windows of the usual width ("virtual pairs": one lane each,
// private accumulators in shared memory, merged into the tile before it leaves the CTA); the eight regular warps all
// run the same narrow loops and advance through the ring at the same pace.
#pragma once
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>
#include <vector>

#include "li_sep2d.cuh"

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
    const int4 *tiles;                    // {c0, va, vb, 0}: columns [c0, c0 + 512) x rows [va, vb) of the model
    const int4 *geo;                      // [nframes][ntiles] {first src row, src rows, first staged src column, flags}
    const T *w1t;                         // [nframes][NU/2][npairs][4]: taps of the column pairs (see multi_taps_kernel)
    const int2 *info;                     // [nframes][npairs] {first src column of the pair's window, validity bits}
    const T *ovw;                         // [nframes][NU/2][32][4]: taps of the overflow windows of the two edge pairs
    const int2 *ovinfo;                   // [nframes][32]: lanes 0..15 first pair of the model, 16..31 last pair
    int novl, novr;                       // overflow windows ever needed by the first / last pair (<= 16 each)
    int ntiles, nframes, G, npairs;       // grid = ntiles * G; group g owns frames [g*nframes/G, (g+1)*nframes/G)
    int m1, m2;
    int P1, CRMAX, NRMAX, NS;             // staged row pitch; longest tile; most src rows feeding one tile; ring slots
    T alpha, beta;
    T *dst;
    T *partial;                           // G > 1: [G][m1*m2] partial models of the groups
    unsigned *counters;                   // G > 1: [ntiles] arrival counters (zero on entry, zero again on exit)
    long long *dbg;                       // timing probe (li_set_option("adj_debug", 9)): [grid] cycles per CTA, or null
};

constexpr int ADM_NC = 256;               // threads of the regular consumer warps
constexpr int ADM_NCO = ADM_NC + 32;      // + the overflow warp
constexpr int ADM_NT = ADM_NCO + 64;      // + the copy warp and the metadata warp
constexpr int ADM_NM = 4;                 // frames of metadata in flight
constexpr int ADM_SW = 2 * ADM_NC;        // columns of a strip
constexpr int ADM_NSMAX = 8;              // ring slots (upper bound)

template <typename T, int S2>
struct AdmW2 {   // elements per row of staged dim-2 weights: whole 16-byte vectors (8 bytes for two floats)
    static constexpr int value = S2 * sizeof(T) <= 8 ? 8 / (int)sizeof(T) : (int)(((S2 * sizeof(T) + 15) & ~(size_t)15) / sizeof(T));
};

template <int S2, int R>
struct AdmK {   // src rows per ring slot: K/R model rows must be whole rotations of the S2 accumulators
    static constexpr int value = (S2 == 6 || S2 == 3) ? (R == 2 ? 12 : 6) : 8;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }
__device__ __forceinline__ double4a ldg4(const double4a *p)
{
    const double2 lo = __ldg(reinterpret_cast<const double2 *>(p)), hi = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    double4a r;
    r.x = lo.x, r.y = lo.y, r.z = hi.x, r.w = hi.y;
    return r;
}

template <typename T>
struct __align__(16) MultiHdr {           // per-frame scalars of a tile, written by the metadata warp
    int4 g;                               // {first src row, src rows, first staged src column, flags (1: regular)}
    const T *src;
    const int32_t *base2;
    int n1, pad[3];
};

template <int S1, bool WIDE>
struct AdmNU {   // staged values per thread = window of a column pair: same-rate operators / up to 2x up-sampling
    static constexpr int value = WIDE ? 2 * S1 + 4 : ((S1 + 3) & ~1);
};

template <typename T, int S1, int S2, int R, bool WIDE>
__global__ void __launch_bounds__(ADM_NT, sizeof(T) == 4 ? 2 : 1)
sep2d_adj_multi_kernel(const MultiArgs<T> a)
{
    constexpr int NC = ADM_NC, SW = ADM_SW, K = AdmK<S2, R>::value, NM = ADM_NM;
    constexpr int NCO = ADM_NCO;
    constexpr int NU_S = AdmNU<S1, WIDE>::value;                  // staged values per thread (window of a column pair)
    constexpr int NQ_S = NU_S / 2;
    static_assert(NU_S <= 16, "validity bits are 16 per column");
    constexpr int W2P = AdmW2<T, S2>::value;
    static_assert(R == 0 || (K % R == 0 && (K / R) % S2 == 0), "bad stage size");
    using V2 = typename VecT<T, 2>::type;
    using V4 = typename VecT<T, 4>::type;
    using Bits = typename std::conditional<sizeof(T) == 4, uint32_t, uint64_t>::type;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P1 = a.P1, NS = a.NS, NVMAX = a.CRMAX + 2 * S2;
    T *sS = reinterpret_cast<T *>(smem_raw);                          // [NS][K][P1] ring of src rows
    T *sAcc = sS + (size_t)NS * K * P1;                               // [CRMAX][SW] the tile's accumulators
    T *sOvf = sAcc + (size_t)a.CRMAX * SW;                            // [CRMAX][NOV][2] accumulators of the overflow windows
    const int NOV = max(a.novl + a.novr, 1);
    T *sW2 = sOvf + (((size_t)a.CRMAX * NOV * 2 + 3) & ~(size_t)3);      // [NM][NRMAX][W2P] alpha * folded dim-2 weights
    int *sNv = reinterpret_cast<int *>(sW2 + (size_t)NM * a.NRMAX * W2P);   // [NM][NVMAX] src rows starting at a model row
    MultiHdr<T> *sHdr = reinterpret_cast<MultiHdr<T> *>(sNv + NM * ((NVMAX + 3) & ~3));   // [NM]
    uint64_t *bar = reinterpret_cast<uint64_t *>(sHdr + NM);
    uint64_t *bfull = bar, *bempty = bar + ADM_NSMAX, *mfull = bar + 2 * ADM_NSMAX, *mempty = mfull + NM;
    __shared__ unsigned s_ticket;

    const int tid = threadIdx.x, lane = tid & 31;
    const int tile = blockIdx.x % a.ntiles, grp = blockIdx.x / a.ntiles;
    const int4 tl = a.tiles[tile];
    const int c0 = tl.x, va = tl.y, vb = tl.z;
    const int v0 = va - (S2 - 1);
    const int NV = vb - v0;
    const int f0 = (int)((int64_t)a.nframes * grp / a.G), f1 = (int)((int64_t)a.nframes * (grp + 1) / a.G);
    const long long t_begin = a.dbg ? clock64() : 0;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&bfull[s], 1);
            mbar_init(&bempty[s], NCO / 32);
        }
        for (int s = 0; s < NM; ++s) {
            mbar_init(&mfull[s], 1);
            mbar_init(&mempty[s], NCO / 32 + 1);   // the consumer warps and the copy warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= NCO + 32) {
        // ------------------------------------------------------------------ metadata warp: up to NM frames ahead
        int ms = 0;
        unsigned mph = 0;
        struct Pre {
            int4 g;
            const T *src, *W2f;
            const int32_t *bp2, *base2;
            int n1;
        };
        auto fetch = [&](int f) {
            Pre p;
            p.g = a.geo[(size_t)f * a.ntiles + tile];
            const MultiFrame<T> &fd = a.frames[f];
            p.src = a.srcs[f], p.W2f = fd.W2f, p.bp2 = fd.bp2, p.base2 = fd.base2, p.n1 = fd.n1;
            return p;
        };
        Pre nx{};
        if (f0 < f1) nx = fetch(f0);
        for (int f = f0; f < f1; ++f) {
            const Pre p = nx;
            if (f + 1 < f1) nx = fetch(f + 1);                    // (in flight while this frame's weights are staged)
            const int i2lo = p.g.x, NR = p.g.y;
            mbar_wait(&mempty[ms], mph ^ 1u);                     // everybody is done with the frame NM frames back
            T *w2 = sW2 + (size_t)ms * a.NRMAX * W2P;
            for (int k = lane; k < NR * S2; k += 32) w2[(k / S2) * W2P + k % S2] = a.alpha * p.W2f[(int64_t)i2lo * S2 + k];
            if (!(p.g.w & 1)) {
                int *nv = sNv + ms * ((NVMAX + 3) & ~3);
                for (int k = lane; k < NVMAX; k += 32) nv[k] = k < NV ? p.bp2[va + k + 1] - p.bp2[va + k] : 0;
            }
            if (lane == 0) {
                MultiHdr<T> h;
                h.g = p.g, h.src = p.src, h.base2 = p.base2, h.n1 = p.n1;
                h.pad[0] = h.pad[1] = h.pad[2] = 0;
                sHdr[ms] = h;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&mfull[ms]);
            if (++ms == NM) ms = 0, mph ^= 1u;
        }
    } else if (tid >= NCO) {
        // ------------------------------------------------------------------ copy warp: the ring of src rows
        int slot = 0, ms = 0;
        unsigned phase = 0, mph = 0;
        for (int f = f0; f < f1; ++f) {
            mbar_wait(&mfull[ms], mph);
            const int4 g = sHdr[ms].g;
            const T *src = sHdr[ms].src;
            const int n1 = sHdr[ms].n1;
            __syncwarp();
            if (lane == 0) mbar_arrive(&mempty[ms]);              // (the header is in registers now)
            if (++ms == NM) ms = 0, mph ^= 1u;
            const int i2lo = g.x, NR = g.y, ulo_a = g.z;
            const int cv1 = min(P1, n1 - ulo_a);
            const uint32_t bytes = (uint32_t)(cv1 * sizeof(T));
            for (int r0 = 0; r0 < NR; r0 += K) {
                const int nr = min(K, NR - r0);
                mbar_wait(&bempty[slot], phase ^ 1u);
                if (lane == 0) mbar_expect_tx(&bfull[slot], bytes * nr);
                __syncwarp();
                if (lane < nr)
                    bulk_g2s(sS + ((size_t)slot * K + lane) * P1, src + (int64_t)(i2lo + r0 + lane) * n1 + ulo_a, bytes, &bfull[slot]);
                if (++slot == NS) slot = 0, phase ^= 1u;
            }
        }
    } else {
        // ------------------------------------------------------------------ consumer warps
        // regular warps: thread t owns columns c0 + 2t, c0 + 2t + 1.  Overflow warp: lanes 0..15 / 16..31 own the
        // overflow windows of the model's first / last pair, when that pair is in this strip.
        const bool ovf = tid >= NC;
        const int c = c0 + 2 * tid;
        const int side = lane >> 4, ow = lane & 15;
        const bool have0 = !ovf ? c < a.m1 : (side == 0 ? (c0 == 0 && ow < a.novl) : (c0 + SW >= a.m1 && ow < a.novr));
        const int pair = !ovf ? c >> 1 : lane;                     // (c0 is a multiple of SW)
        const int npw = !ovf ? a.npairs : 32;                      // pairs per frame in the tap arrays
        const T *const w1t = !ovf ? a.w1t : a.ovw;
        const int2 *const infos = !ovf ? a.info : a.ovinfo;
        V2 *const accbase = !ovf ? reinterpret_cast<V2 *>(sAcc) + tid
                                 : reinterpret_cast<V2 *>(sOvf) + (side == 0 ? ow : a.novl + ow) % NOV;
        const int accstep = !ovf ? SW / 2 : NOV;
        const bool wr = !ovf || have0;                             // (idle overflow lanes share accumulator slots: no writes)
        if (wr) {
            V2 z;
            z.x = (T)0, z.y = (T)0;
            for (int r = 0; r < vb - va; ++r) accbase[(size_t)r * accstep] = z;
        }
        auto keep = [](T x, Bits m) {
            Bits b;
            memcpy(&b, &x, sizeof b);
            b &= m;
            memcpy(&x, &b, sizeof b);
            return x;
        };
        // taps of the thread's column pair: chunk q = (w0[2q], w0[2q+1], w1[2q], w1[2q+1]); the lines of the next
        // frame are pulled into L2 while this frame is processed
        auto taps_of = [&](int f) { return reinterpret_cast<const V4 *>(w1t) + (size_t)f * NQ_S * npw + pair; };
        auto prefetch = [&](int f) {
            if (have0) {
                const V4 *p = taps_of(f);
#pragma unroll
                for (int q = 0; q < NQ_S; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (size_t)q * npw));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(infos + (size_t)f * npw + pair));
            }
        };

        int slot = 0, ms = 0;
        unsigned phase = 0, mph = 0;
        for (int f = f0; f < f1; ++f) {
            V4 wqs[NQ_S];
            int2 info = make_int2(0, 0);
#pragma unroll
            for (int q = 0; q < NQ_S; ++q) wqs[q] = V4{};
            if (have0) {
                const V4 *p = taps_of(f);
#pragma unroll
                for (int q = 0; q < NQ_S; ++q) wqs[q] = ldg4(p + (size_t)q * npw);
                info = __ldg(infos + (size_t)f * npw + pair);
            }
            if (f + 1 < f1) prefetch(f + 1);
            const unsigned vm = have0 ? (unsigned)info.y : 0u;
            const unsigned vm0 = vm & 0xffffu, vm1 = vm >> 16;
            const bool idle = __all_sync(0xffffffffu, vm == 0u);   // (a warp without any tap this frame: it only walks the ring)
            // exact windows: both columns have exactly S1 consecutive taps, the second column's one position later
            constexpr unsigned TAPS = (1u << S1) - 1u;
            const bool ex0 = __all_sync(0xffffffffu, vm0 == TAPS && vm1 == TAPS << 1);
            const bool ex1 = __all_sync(0xffffffffu, vm0 == TAPS << 1 && vm1 == TAPS << 2);
            mbar_wait(&mfull[ms], mph);
            const int4 g = sHdr[ms].g;
            const int32_t *const base2 = sHdr[ms].base2;
            const int i2lo = g.x, NR = g.y;
            const bool regular = R > 0 && (g.w & 1);
            const int off = vm != 0u ? info.x - g.z : 0;           // (lanes without taps read the start of the row)
            const T *const w2base = sW2 + (size_t)ms * a.NRMAX * W2P;
            const int *const nvbase = sNv + ms * ((NVMAX + 3) & ~3);

            V2 acc[S2];
#pragma unroll
            for (int s = 0; s < S2; ++s) acc[s].x = (T)0, acc[s].y = (T)0;
            V2 *ap = accbase;                                         // accumulator pair of the next model row to complete
            int rowstep = P1;
            asm volatile("" : "+r"(rowstep));

            int held = -1;                                            // ring slot being consumed
            auto release = [&]() {
                if (held >= 0) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bempty[held]);
                    held = -1;
                }
            };
            auto next_stage = [&]() {   // warp-uniform; returns this thread's window in the first row of the next slot
                release();
                mbar_wait(&bfull[slot], phase);
                held = slot;
                const T *p = sS + (size_t)slot * K * P1 + off;
                if (++slot == NS) slot = 0, phase ^= 1u;
                return p;
            };
            auto emit = [&](int vv, int s) {   // model row vv is complete for this frame
                if (vv >= va) {
                    if (wr) *ap = add2(*ap, acc[s]);
                    ap += accstep;
                }
                acc[s].x = (T)0, acc[s].y = (T)0;
            };
            auto ld_w2 = [&](const T *p, V2(&w2)[S2]) {   // the row's weights, each for both columns (block-uniform loads)
                T w[W2P];
                if constexpr (W2P * sizeof(T) >= 16) {
                    constexpr int VV = 16 / (int)sizeof(T);
#pragma unroll
                    for (int k = 0; k < W2P / VV; ++k) {
                        if (k * VV < S2) {
                            const Pack<T, VV> t = ld_pack<T, VV>(p + k * VV);
#pragma unroll
                            for (int v = 0; v < VV; ++v) w[k * VV + v] = t.v[v];
                        }
                    }
                } else {
                    const Pack<T, W2P> t = ld_pack<T, W2P>(p);
#pragma unroll
                    for (int v = 0; v < W2P; ++v) w[v] = t.v[v];
                }
#pragma unroll
                for (int s = 0; s < S2; ++s) w2[s].x = w[s], w2[s].y = w[s];
            };

            // The row loops, instantiated for NU = staged values per thread and MODE = how a column's taps are picked
            // from the staged window: 0 = one select per tap and column, 1 = bit masks on the first two / last two
            // positions (all others are taps of both columns), 2 / 3 = exact windows starting at position 0 / 1 (no
            // masks: only taps are touched).
            auto run = [&](auto nu_tag, auto mode_tag, const auto &wq) {
                constexpr int NU = decltype(nu_tag)::value;
                constexpr int MODE = decltype(mode_tag)::value;
                constexpr int NMK = NU >= 4 ? 4 : NU;                 // masked positions: 0, 1, NU-2, NU-1
                Bits km[2][NMK];
                if constexpr (MODE == 1) {
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int t = 0; t < NMK; ++t) {
                            const int pos = t < 2 ? t : NU - 4 + t;
                            km[e][t] = ((vm >> (16 * e + pos)) & 1u) ? ~(Bits)0 : (Bits)0;
                            if constexpr (sizeof(T) == 4)
                                asm volatile("" : "+r"(km[e][t]));
                            else
                                asm volatile("" : "+l"(km[e][t]));
                        }
                }
                auto wsel = [&](int q, int e) {   // weights of window positions (2q, 2q+1) for column e of the pair
                    V2 w;
                    w.x = e == 0 ? wq[q].x : wq[q].z, w.y = e == 0 ? wq[q].y : wq[q].w;
                    return w;
                };
                auto load_x = [&](const T *p, V2(&x)[NU / 2]) {
#pragma unroll
                    for (int q = 0; q < NU / 2; ++q) x[q] = *reinterpret_cast<const V2 *>(p + 2 * q);
                };
                auto dot = [&](const V2(&x)[NU / 2]) {   // dim 1 of the staged row: one value per column of the pair
                    V2 h;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        T r;
                        if constexpr (MODE >= 2) {
                            constexpr int O = MODE - 2;
                            const int st = O + e;                    // taps at positions st .. st + S1 - 1
                            V2 s;
                            s.x = (T)0, s.y = (T)0;
                            bool first = true;
#pragma unroll
                            for (int q = 0; q < NU / 2; ++q)
                                if (2 * q >= st && 2 * q + 1 < st + S1) {
                                    s = first ? mul2(x[q], wsel(q, e)) : fma2(x[q], wsel(q, e), s);
                                    first = false;
                                }
                            r = s.x + s.y;
#pragma unroll
                            for (int q = 0; q < NU / 2; ++q) {
                                const bool in0 = 2 * q >= st && 2 * q < st + S1, in1 = 2 * q + 1 >= st && 2 * q + 1 < st + S1;
                                if (in0 != in1) r = in0 ? fmaT(x[q].x, wsel(q, e).x, r) : fmaT(x[q].y, wsel(q, e).y, r);
                            }
                        } else {
                            V2 y[NU / 2];
#pragma unroll
                            for (int q = 0; q < NU / 2; ++q) y[q] = x[q];
                            if constexpr (MODE == 1) {
                                y[0].x = keep(y[0].x, km[e][0]);
                                y[0].y = keep(y[0].y, km[e][1]);
                                y[NU / 2 - 1].x = keep(y[NU / 2 - 1].x, km[e][NMK - 2]);
                                y[NU / 2 - 1].y = keep(y[NU / 2 - 1].y, km[e][NMK - 1]);
                            } else {
                                unsigned m = vm >> (16 * e);
                                asm volatile("" : "+r"(m));
#pragma unroll
                                for (int q = 0; q < NU / 2; ++q) {
                                    y[q].x = ((m >> (2 * q)) & 1u) ? y[q].x : (T)0;
                                    y[q].y = ((m >> (2 * q + 1)) & 1u) ? y[q].y : (T)0;
                                }
                            }
                            V2 s = mul2(y[0], wsel(0, e));
#pragma unroll
                            for (int q = 1; q < NU / 2; ++q) s = fma2(y[q], wsel(q, e), s);
                            r = s.x + s.y;
                        }
                        if (e == 0)
                            h.x = r;
                        else
                            h.y = r;
                    }
                    return h;
                };
                if (regular) {
                    if constexpr (R > 0) {   // a slot holds the R rows of each of DR consecutive model rows
                        constexpr int DR = K / R;
                        const T *pw = w2base;
                        // row i of a slot belongs to model row d = i / R: it adds into accumulators (d .. d+S2-1) mod S2;
                        // the last of them was emitted by the previous model row (or is still zero), so the first row
                        // of a model row writes it with a multiplication: no zeroing
                        auto dim2 = [&](const V2 h, const T *pwr, int d, int r) {
                            V2 w2[S2];
                            ld_w2(pwr, w2);
#pragma unroll
                            for (int s = 0; s < S2 - 1; ++s) acc[(d + s) % S2] = fma2(w2[s], h, acc[(d + s) % S2]);
                            if (r == 0)
                                acc[(d + S2 - 1) % S2] = mul2(w2[S2 - 1], h);
                            else
                                acc[(d + S2 - 1) % S2] = fma2(w2[S2 - 1], h, acc[(d + S2 - 1) % S2]);
                        };
                        for (int vs = v0; vs < vb; vs += DR) {
                            const T *prow = next_stage();
                            if (vs >= va && vs + DR <= vb) {
                                // a full slot inside the tile: no per-row tests; the staged values of row i+1 are loaded
                                // before the arithmetic of row i
                                V2 xc[NU / 2];
                                load_x(prow, xc);
#pragma unroll
                                for (int i = 0; i < K; ++i) {
                                    V2 xn[NU / 2];
                                    if (i + 1 < K) {
                                        prow += rowstep;
                                        load_x(prow, xn);
                                    }
                                    dim2(dot(xc), pw + i * W2P, i / R, i % R);
                                    if (i % R == R - 1) {
                                        if (wr) *ap = add2(*ap, acc[(i / R) % S2]);
                                        ap += accstep;
                                    }
                                    if (i + 1 < K) {
#pragma unroll
                                        for (int q = 0; q < NU / 2; ++q) xc[q] = xn[q];
                                    }
                                }
                            } else {
#pragma unroll
                                for (int d = 0; d < DR; ++d) {
                                    const int vv = vs + d;
                                    if (vv < vb) {   // block-uniform: only the last slot of a frame is partial
#pragma unroll
                                        for (int r = 0; r < R; ++r) {
                                            V2 x[NU / 2];
                                            load_x(prow, x);
                                            prow += rowstep;
                                            dim2(dot(x), pw + (d * R + r) * W2P, d, r);
                                        }
                                        if (vv >= va) {
                                            if (wr) *ap = add2(*ap, acc[d % S2]);
                                            ap += accstep;
                                        }
                                    }
                                }
                            }
                            pw += K * W2P;
                        }
                    }
                } else if constexpr (MODE < 2) {   // generic schedule: for model row vv, every src row whose folded window starts at vv
                    const T *pw = w2base;
                    const int *pn = nvbase;
                    const T *prow = sS + off;
                    int left = 0, k = 0;
                    auto row = [&]() {
                        if (left == 0) {
                            prow = next_stage();
                            left = K;
                        }
                        --left;
                        V2 x[NU / 2];
                        load_x(prow, x);
                        prow += rowstep;
                        return dot(x);
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
                                        const V2 h = row();
                                        V2 w2[S2];
                                        ld_w2(pw, w2);
                                        pw += W2P;
#pragma unroll
                                        for (int s = 0; s < S2; ++s) acc[(u + s) % S2] = fma2(w2[s], h, acc[(u + s) % S2]);
                                    }
                                } else {   // first / last window of the array: only the slots the clamped taps land on
#pragma unroll 1
                                    for (; n > 0; --n) {
                                        const V2 h = row();
                                        V2 w2[S2];
                                        ld_w2(pw, w2);
                                        pw += W2P;
                                        const int b = base2[i2lo + k];
                                        ++k;
                                        const int lo = clampi(b, a.m2 - 1) - vv, hi = clampi(b + S2 - 1, a.m2 - 1) - vv;
#pragma unroll
                                        for (int s = 0; s < S2; ++s)
                                            if (s >= lo && s <= hi) acc[(u + s) % S2] = fma2(w2[s], h, acc[(u + s) % S2]);
                                    }
                                }
                                emit(vv, u);
                            }
                        }
                        pn += S2;
                    }
                }
                release();
            };
            if (NR > 0) {
                // (warps of one CTA may take different copies: they only meet at the mbarriers)
                using TagS = std::integral_constant<int, NU_S>;
                if (idle) {
                    for (int r0 = 0; r0 < NR; r0 += K) next_stage();
                    release();
                } else if (!WIDE && regular && ex0) {
                    if constexpr (!WIDE) run(TagS{}, std::integral_constant<int, 2>{}, wqs);
                } else if (!WIDE && regular && ex1) {
                    if constexpr (!WIDE) run(TagS{}, std::integral_constant<int, 3>{}, wqs);
                } else {
                    constexpr unsigned MID = NU_S > 4 ? (((1u << (NU_S - 2)) - 1u) & ~3u) * 0x10001u : 0u;
                    if (__all_sync(0xffffffffu, !have0 || (vm & MID) == MID))
                        run(TagS{}, std::integral_constant<int, 1>{}, wqs);
                    else
                        run(TagS{}, std::integral_constant<int, 0>{}, wqs);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&mempty[ms]);
            if (++ms == NM) ms = 0, mph ^= 1u;
        }
    }
    __syncthreads();
    if (a.dbg && tid == 0) a.dbg[blockIdx.x] = clock64() - t_begin;
    // ---- overflow windows: added to the pair they belong to, in window order (fixed order: deterministic)
    if (a.novl + a.novr > 0) {
        for (int k = tid; k < (vb - va) * 2; k += ADM_NT) {
            const int r = k >> 1, e = k & 1;
            if (c0 == 0) {
                T s = sAcc[(size_t)r * SW + e];
                for (int w = 0; w < a.novl; ++w) s += sOvf[((size_t)r * NOV + w) * 2 + e];
                sAcc[(size_t)r * SW + e] = s;
            }
            if (c0 + SW >= a.m1) {
                const int col = ((a.npairs - 1) * 2 - c0) + e;
                T s = sAcc[(size_t)r * SW + col];
                for (int w = 0; w < a.novr; ++w) s += sOvf[((size_t)r * NOV + a.novl + w) * 2 + e];
                sAcc[(size_t)r * SW + col] = s;
            }
        }
        __syncthreads();
    }

    // ---- the tile leaves the CTA (consumer threads: their own pair of columns)
    const int64_t plane = (int64_t)a.m1 * a.m2;
    const int c = c0 + 2 * tid;
    const bool have0 = tid < NC && c < a.m1, have1 = tid < NC && c + 1 < a.m1;
    const T *accp = sAcc + 2 * (tid < NC ? tid : 0);
    if (a.G == 1) {
        if (have0) {
            T *dp = a.dst + (int64_t)va * a.m1 + c;
            const bool hb = a.beta != (T)0;   // beta = 0: dst is never read
            for (int r = 0; r < vb - va; ++r) {
                T *d = dp + (int64_t)r * a.m1;
                d[0] = hb ? fmaT(a.beta, d[0], accp[(size_t)r * SW]) : accp[(size_t)r * SW];
                if (have1) d[1] = hb ? fmaT(a.beta, d[1], accp[(size_t)r * SW + 1]) : accp[(size_t)r * SW + 1];
            }
        }
        return;
    }
    if (have0) {
        T *pp = a.partial + grp * plane + (int64_t)va * a.m1 + c;
        for (int r = 0; r < vb - va; ++r) {
            pp[(int64_t)r * a.m1] = accp[(size_t)r * SW];
            if (have1) pp[(int64_t)r * a.m1 + 1] = accp[(size_t)r * SW + 1];
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&a.counters[tile], 1u);
    __syncthreads();
    if (s_ticket != (unsigned)(a.G - 1)) return;
    __threadfence();   // the last group to arrive adds the partial tiles up, in group order
    if (have0) {
        T *dp = a.dst + (int64_t)va * a.m1 + c;
        const T *pp = a.partial + (int64_t)va * a.m1 + c;
        for (int r = 0; r < vb - va; ++r)
            for (int e = 0; e < (have1 ? 2 : 1); ++e) {
                const int64_t o = (int64_t)r * a.m1 + e;
                T s = __ldcg(pp + o);
                for (int g = 1; g < a.G; ++g) s += __ldcg(pp + g * plane + o);
                dp[o] = a.beta != (T)0 ? fmaT(a.beta, dp[o], s) : s;
            }
    }
    if (tid == 0) a.counters[tile] = 0u;
}

// ---------------------------------------------------------------------------------------------
// Plan kernels (once per list of tables, cached).

// Taps of the model's column pairs, transposed and folded: for frame f and pair p = columns (2p, 2p+1) the staged
// window starts at src column a0 = (first src column feeding column 2p) & ~1; position t of the window carries the
// weight W1f[a0+t][column - bfold[a0+t]] when src column a0+t has a (possibly clamped) tap on the column, else 0 with
// its validity bit clear.  A pair stores NU positions; the first / last pair of the model (which collect the src
// columns hanging over the edges) continue in up to 16 overflow windows of NU positions each (ovw / ovinfo, lanes
// 0..15 / 16..31).  out[1] = widest span of src columns feeding a strip, out[3] = 1 when some other pair needs more
// than NU positions or an edge pair more than 17 windows, out[4] / out[5] = most overflow windows of the first / last pair.
template <typename T>
__global__ void multi_taps_kernel(const MultiFrame<T> *__restrict__ frames, int m1, int S1, int NU, int npairs, int SW,
                                  T *__restrict__ w1t, int2 *__restrict__ info, T *__restrict__ ovw, int2 *__restrict__ ovinfo,
                                  int *__restrict__ out)
{
    const int f = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
    const MultiFrame<T> fd = frames[f];
    int span = 0;
    if (p < npairs) {
        const int c = 2 * p, cl = min(c + 1, m1 - 1);
        const int e0 = c == 0 ? 0 : fd.bp1[c], e1 = cl == m1 - 1 ? fd.n1 : fd.bp1[cl + S1];
        const int a0 = e0 & ~1;
        const int need = e1 - a0;
        int nwin = max((need + NU - 1) / NU, 1);                  // windows of NU positions (the first one is the pair's own)
        const bool left = p == 0, right = !left && p == npairs - 1;
        if (nwin > 1 && !(left || right)) atomicMax(out + 3, 1);
        if (nwin > 17) atomicMax(out + 3, 1), nwin = 17;
        if (!(left || right)) nwin = 1;
        if (left) atomicMax(out + 4, nwin - 1);
        if (right) atomicMax(out + 5, nwin - 1);
        for (int k = 0; k < nwin; ++k) {
            unsigned vm = 0;
            const int lane = (left ? 0 : 16) + k - 1;             // (k >= 1: slot of the overflow window)
            for (int q = 0; q < NU / 2; ++q) {
                T w[4] = {(T)0, (T)0, (T)0, (T)0};
                for (int e = 0; e < 2; ++e) {
                    const int t = 2 * q + e, i = a0 + k * NU + t;
                    if (i < fd.n1) {
                        const int b = fd.base1[i], bf = fd.bfold1[i];
                        const int lo = min(max(b, 0), m1 - 1), hi = min(max(b + S1 - 1, 0), m1 - 1);
                        if (c >= lo && c <= hi) {
                            w[e] = fd.W1f[(int64_t)i * S1 + (c - bf)];
                            vm |= 1u << t;
                        }
                        if (c + 1 < m1 && c + 1 >= lo && c + 1 <= hi) {
                            w[2 + e] = fd.W1f[(int64_t)i * S1 + (c + 1 - bf)];
                            vm |= 1u << (16 + t);
                        }
                    }
                }
                T *o = k == 0 ? w1t + (((size_t)f * (NU / 2) + q) * npairs + p) * 4 : ovw + (((size_t)f * (NU / 2) + q) * 32 + lane) * 4;
                o[0] = w[0], o[1] = w[1], o[2] = w[2], o[3] = w[3];
            }
            if (k == 0)
                info[(size_t)f * npairs + p] = make_int2(a0, (int)vm);
            else
                ovinfo[(size_t)f * 32 + lane] = make_int2(a0 + k * NU, (int)vm);
        }
        if (c % SW == 0) {
            const int k1 = min(c + SW, m1);
            const int lo = c == 0 ? 0 : fd.bp1[c], hi = k1 == m1 ? fd.n1 : fd.bp1[k1 - 1 + S1];
            span = hi - lo;
        }
    }
    for (int o = 16; o; o >>= 1) span = max(span, __shfl_xor_sync(0xffffffffu, span, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out + 1, span);
}

// per (frame, tile): src rows feeding the tile, first staged src column, regularity; out[2] = most src rows of a tile
template <typename T>
__global__ void multi_geo_kernel(const MultiFrame<T> *__restrict__ frames, const int4 *__restrict__ tiles, int ntiles, int m2,
                                 int S2, int R, int AL, int4 *__restrict__ geo, int *__restrict__ out)
{
    const int f = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    const MultiFrame<T> fd = frames[f];
    int nr = 0;
    if (t < ntiles) {
        const int4 tl = tiles[t];
        const int c0 = tl.x, va = tl.y, vb = tl.z, v0 = va - (S2 - 1);
        const int i2lo = fd.bp2[va];
        nr = fd.bp2[vb - 1 + S2] - i2lo;
        const int u_lo = c0 == 0 ? 0 : fd.bp1[c0];
        const bool regular = R > 0 && fd.irrp2 && v0 > 0 && vb <= m2 - S2 && fd.irrp2[vb] == fd.irrp2[v0];
        geo[(size_t)f * ntiles + t] = make_int4(i2lo, nr, u_lo & ~(AL - 1), regular ? 1 : 0);
    }
    for (int o = 16; o; o >>= 1) nr = max(nr, __shfl_xor_sync(0xffffffffu, nr, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out + 2, nr);
}

// ---------------------------------------------------------------------------------------------
// Host side.

struct MultiPlan {
    void *frames = nullptr;       // device MultiFrame<T>[n]
    int4 *tiles = nullptr, *geo = nullptr;
    void *w1t = nullptr, *ovw = nullptr;
    int2 *info = nullptr, *ovinfo = nullptr;
    unsigned *counters = nullptr;
    int ntiles = 0, G = 1, P1 = 0, CRMAX = 0, NRMAX = 0, R = 0, NS = 2, npairs = 0, novl = 0, novr = 0, wide = 0;
    size_t smem = 0;
    bool ok = false;              // false: the kernel cannot take this batch (cached too)
    uint64_t stamp = 0;
    void release()
    {
        cudaFree(frames), cudaFree(tiles), cudaFree(geo), cudaFree(w1t), cudaFree(info), cudaFree(counters), cudaFree(ovw), cudaFree(ovinfo);
        frames = nullptr, tiles = nullptr, geo = nullptr, w1t = nullptr, info = nullptr, counters = nullptr, ovw = nullptr, ovinfo = nullptr;
    }
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
    constexpr int SW = ADM_SW, AL = 16 / (int)sizeof(T);
    const int m1 = (int)tabs[0]->ncols, m2 = (int)tabs[1]->ncols;
    const bool trace = opt(OPT_TRACE) != 0;
    plan->ok = false;
    if (m1 < S1 || m2 < S2) return LI_OK;
    int R = -1;
    int64_t n2sum = 0;
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
        n2sum += t2->nrows;
    }
    plan->R = std::max(R, 0);
    const int K = (S2 == 6 || S2 == 3) ? (plan->R == 2 ? 12 : 6) : 8;   // = AdmK<S2, R>::value
    const int npairs = (m1 + 1) / 2;
    plan->npairs = npairs;
    int h_out[6] = {0, 0, 0, 0, 0, 0};
    int *d_out = nullptr;
    LI_CUDA(cudaMalloc(&plan->frames, (size_t)n * sizeof(MultiFrame<T>)));
    LI_CUDA(cudaMalloc(&d_out, 6 * sizeof(int)));
    LI_CUDA(cudaMemcpyAsync(plan->frames, fr.data(), (size_t)n * sizeof(MultiFrame<T>), cudaMemcpyHostToDevice, st));
    int NU = 0;
    for (int wide = 0; wide < 2; ++wide) {
    NU = wide ? AdmNU<S1, true>::value : AdmNU<S1, false>::value;
    plan->wide = wide;
    cudaFree(plan->w1t), cudaFree(plan->info), cudaFree(plan->ovw), cudaFree(plan->ovinfo);
    plan->w1t = plan->ovw = nullptr, plan->info = plan->ovinfo = nullptr;

    // ---- pass 1: taps of the column pairs (+ longest window, widest strip span)
    LI_CUDA(cudaMalloc(&plan->w1t, (size_t)n * (NU / 2) * npairs * 4 * sizeof(T)));
    LI_CUDA(cudaMalloc(&plan->info, (size_t)n * npairs * sizeof(int2)));
    LI_CUDA(cudaMalloc(&plan->ovw, (size_t)n * (NU / 2) * 32 * 4 * sizeof(T)));
    LI_CUDA(cudaMalloc(&plan->ovinfo, (size_t)n * 32 * sizeof(int2)));
    LI_CUDA(cudaMemsetAsync(plan->ovw, 0, (size_t)n * (NU / 2) * 32 * 4 * sizeof(T), st));
    LI_CUDA(cudaMemsetAsync(plan->ovinfo, 0, (size_t)n * 32 * sizeof(int2), st));
    LI_CUDA(cudaMemsetAsync(d_out, 0, 6 * sizeof(int), st));
    multi_taps_kernel<T><<<dim3((unsigned)((npairs + 127) / 128), (unsigned)n), 128, 0, st>>>(
        (const MultiFrame<T> *)plan->frames, m1, S1, NU, npairs, SW, (T *)plan->w1t, plan->info, (T *)plan->ovw, plan->ovinfo, d_out);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, d_out, sizeof h_out, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // (also: fr may go out of scope)
    if (e != cudaSuccess) {
        cudaFree(d_out);
        set_error("multi-frame plan failed: %s", cudaGetErrorString(e));
        return LI_ERR_CUDA;
    }
    if (!h_out[3]) break;   // (else: the wide windows of up-sampling operators)
    }
    cudaError_t e = cudaSuccess;
    if (h_out[3]) {   // (down-sampling along dim 1, or more than 17 windows of src columns hanging over an edge)
        cudaFree(d_out);
        if (trace) fprintf(stderr, "[li_b200] adj multi declines: a column pair is fed by more than %d src columns\n", NU);
        return LI_OK;
    }
    plan->novl = h_out[4], plan->novr = h_out[5];
    plan->P1 = (h_out[1] + (AL - 1) + 2 * NU + (AL - 1)) & ~(AL - 1);

    // ---- tiles: strips of SW columns x chunks of rows, one resident wave of CTAs; the chunks at the array edges
    // always run the generic row loop and are made shorter; frame groups when the tiles alone cannot fill the machine
    int dev = 0, smem_sm = 0, smem_optin = 0;
    LI_CUDA(cudaGetDevice(&dev));
    LI_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    LI_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int occ = sizeof(T) == 4 ? 2 : 1;
    const size_t budget = std::min<size_t>((size_t)smem_optin, (size_t)smem_sm / occ - 1024) & ~(size_t)15;
    const int strips = (m1 + SW - 1) / SW, rows = row_hi - row_lo;
    const double rate2 = (double)n2sum / (double)n / (double)m2;        // src rows per model row
    auto nrmax_of = [&](int CR) { return ((int)((CR + S2) * rate2 * 1.05) + 2 * S2 + 2) & ~1; };
    auto smem_of = [&](int CR, int NS, int NRMAX) {
        const int NVMAX = CR + 2 * S2;
        return ((size_t)NS * K * plan->P1 + (size_t)CR * SW + (((size_t)CR * std::max(plan->novl + plan->novr, 1) * 2 + 3) & ~(size_t)3) +
                (size_t)ADM_NM * NRMAX * AdmW2<T, S2>::value) * sizeof(T) +
               (size_t)ADM_NM * ((NVMAX + 3) & ~3) * sizeof(int) + ADM_NM * sizeof(MultiHdr<T>) +
               (2 * ADM_NSMAX + 2 * ADM_NM) * sizeof(uint64_t);
    };
    const int slots = sep2d_num_sms() * occ;
    const double edge_w = 0.4;
    // chunk boundaries for interior chunks of CR rows: the first / last chunk (when it touches the array edge and
    // there are more than two chunks) holds edge_w * CR rows
    auto cut = [&](int CRi, std::vector<int> &cs) {
        const int ne = (row_lo == 0 ? 1 : 0) + (row_hi == m2 ? 1 : 0);
        const int E = std::max(1, (int)(CRi * edge_w));
        const bool edges = ne > 0 && rows > ne * E + CRi;
        cs.assign(1, row_lo);
        int pos = row_lo;
        const int stop = edges && row_hi == m2 ? row_hi - E : row_hi;   // (the last edge chunk is cut off first)
        if (edges && row_lo == 0) cs.push_back(pos += E);
        while (pos < stop) cs.push_back(pos = std::min(pos + CRi, stop));
        if (stop < row_hi) cs.push_back(row_hi);
    };
    int chunks = std::max(1, std::min(slots / strips, rows / std::max(8, 2 * S2)));
    int CR = 0, NS = 4;
    std::vector<int> cs;
    for (;; ++chunks) {
        // the shortest interior chunk length that needs at most `chunks` chunks
        int lo = 1, hi = rows;
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            cut(mid, cs);
            if ((int)cs.size() - 1 <= chunks)
                hi = mid;
            else
                lo = mid + 1;
        }
        cut(lo, cs);
        CR = 0;
        for (size_t k = 0; k + 1 < cs.size(); ++k) CR = std::max(CR, cs[k + 1] - cs[k]);
        bool fits = false;
        for (NS = 4; NS >= 2; --NS)
            if (smem_of(CR, NS, nrmax_of(CR)) <= budget) {
                fits = true;
                break;
            }
        if (fits || CR <= 2 * S2) break;   // (more chunks than resident slots: the frames are read once per wave)
    }
    NS = std::max(NS, 2);
    if (smem_of(CR, NS, nrmax_of(CR)) > budget) {
        cudaFree(d_out);
        if (trace) fprintf(stderr, "[li_b200] adj multi declines: a tile of %d rows does not fit shared memory\n", CR);
        return LI_OK;
    }
    std::vector<int4> tl;
    for (size_t ch = 0; ch + 1 < cs.size(); ++ch)
        for (int sx = 0; sx < strips; ++sx) tl.push_back(make_int4(sx * SW, cs[ch], cs[ch + 1], 0));
    plan->ntiles = (int)tl.size();
    plan->CRMAX = CR;
    plan->NS = NS;
    // frame groups: enough CTAs to fill the machine, then several times as many as resident slots when every group keeps
    // at least 8 frames -- tiles differ in cost (edge strips and edge chunks run slower loops), and with more CTAs than
    // slots the hardware scheduler evens that out (measured on config #5: 10.0 -> 7.9 us per frame)
    const int64_t g_fill = std::max<int64_t>(1, (slots + plan->ntiles / 2) / plan->ntiles);
    // (measured, frames -> best G with g_fill = 1: 64 -> 3, 128 -> 6, 512 -> 12; the groups of a tile are combined at the end,
    // so short batches want fewer of them)
    const int64_t g_over = n >= 256 ? 12 : (n >= 96 ? 6 : 3);
    plan->G = (int)std::max<int64_t>(1, std::max(std::min<int64_t>(g_fill, n / 4), std::min<int64_t>(g_over * g_fill, n / 8)));
    if (opt(OPT_ADJ_DEBUG) >= 11 && opt(OPT_ADJ_DEBUG) <= 26) plan->G = opt(OPT_ADJ_DEBUG) - 10;   // (experiments)

    // ---- pass 2: per (frame, tile) geometry (+ most src rows feeding a tile)
    LI_CUDA(cudaMalloc(&plan->tiles, tl.size() * sizeof(int4)));
    LI_CUDA(cudaMalloc(&plan->geo, (size_t)n * tl.size() * sizeof(int4)));
    LI_CUDA(cudaMalloc(&plan->counters, tl.size() * sizeof(unsigned)));
    LI_CUDA(cudaMemcpyAsync(plan->tiles, tl.data(), tl.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    LI_CUDA(cudaMemsetAsync(plan->counters, 0, tl.size() * sizeof(unsigned), st));
    multi_geo_kernel<T><<<dim3((unsigned)((tl.size() + 127) / 128), (unsigned)n), 128, 0, st>>>(
        (const MultiFrame<T> *)plan->frames, plan->tiles, plan->ntiles, m2, S2, plan->R, AL, plan->geo, d_out);
    count_launch();
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, d_out, sizeof h_out, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // (also: tl may go out of scope)
    cudaFree(d_out);
    if (e != cudaSuccess) {
        set_error("multi-frame plan failed: %s", cudaGetErrorString(e));
        return LI_ERR_CUDA;
    }
    plan->NRMAX = (std::max(h_out[2], 1) + 1) & ~1;
    while (plan->NS > 2 && smem_of(CR, plan->NS, plan->NRMAX) > budget) --plan->NS;
    plan->smem = smem_of(CR, plan->NS, plan->NRMAX);
    if (trace)
        fprintf(stderr, "[li_b200] adj multi: %lld frames, tiles %d x %zu (CR %d) x G %d, P1 %d NRMAX %d NS %d K %d smem %zu R %d overflow windows %d + %d, NU %d\n",
                (long long)n, strips, cs.size() - 1, CR, plan->G, plan->P1, plan->NRMAX, plan->NS, K, plan->smem, plan->R, plan->novl, plan->novr, NU);
    if (plan->smem > budget) return LI_OK;
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
                old->second.release();
                plans.erase(old);
            }
            MultiPlan p;
            int rc = multi_build_plan<T, S1, S2>(n, tabs, row_lo, row_hi, st, &p);
            if (rc != LI_OK || !p.ok) {
                p.release();
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
    a.geo = plan.geo;
    a.w1t = (const T *)plan.w1t;
    a.info = plan.info;
    a.ovw = (const T *)plan.ovw;
    a.ovinfo = plan.ovinfo;
    a.novl = plan.novl, a.novr = plan.novr;
    a.ntiles = plan.ntiles, a.nframes = (int)n, a.G = plan.G, a.npairs = plan.npairs;
    a.m1 = m1, a.m2 = m2;
    a.P1 = plan.P1, a.CRMAX = plan.CRMAX, a.NRMAX = plan.NRMAX, a.NS = plan.NS;
    a.alpha = (T)alpha, a.beta = (T)beta;
    a.dst = (T *)dst;
    a.partial = (T *)d_part;
    a.counters = plan.counters;
    const unsigned blocks = (unsigned)(plan.ntiles * plan.G);
    a.dbg = nullptr;
    if (opt(OPT_ADJ_DEBUG) == 9) LI_CUDA(cudaMalloc(&a.dbg, blocks * sizeof(long long)));
#define LI_ADM(RR, WW)                                                  \
    if (plan.R == RR && plan.wide == (WW ? 1 : 0)) {                    \
        auto kern = sep2d_adj_multi_kernel<T, S1, S2, RR, WW>;          \
        rc = set_smem(kern, plan.smem);                                 \
        if (rc != LI_OK) return rc;                                     \
        kern<<<blocks, ADM_NT, plan.smem, st>>>(a);                     \
    }
    LI_ADM(0, false) LI_ADM(1, false) LI_ADM(2, false) LI_ADM(0, true) LI_ADM(1, true) LI_ADM(2, true)
#undef LI_ADM
    LI_CHECK_LAUNCH();
    if (a.dbg) {   // timing probe: cycles per CTA, tile by tile
        std::vector<long long> h(blocks);
        std::vector<int4> tl((size_t)plan.ntiles);
        cudaMemcpy(h.data(), a.dbg, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaMemcpy(tl.data(), plan.tiles, tl.size() * sizeof(int4), cudaMemcpyDeviceToHost);
        cudaFree(a.dbg);
        for (unsigned b = 0; b < blocks; ++b)
            fprintf(stderr, "[li_b200] multi cta %u tile c0 %d rows %d..%d group %u: %lld cycles\n", b, tl[b % plan.ntiles].x,
                    tl[b % plan.ntiles].y, tl[b % plan.ntiles].z, b / plan.ntiles, h[b]);
    }
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
