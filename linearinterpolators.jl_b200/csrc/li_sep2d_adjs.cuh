// li_sep2d_adjs.cuh -- fused 2-D separable interpolation, adjoint, streaming form (K7).
//
// Reference: the adjoint of the LazyAlgebra composition A1*A2 of test/benchmarks.jl:27-29, whose factors are the
// scatter loops of src/tabulated.jl:438-478 / src/sparse.jl:760-800 (same mathematics as src/separable.jl:113-162).
// Included by li_sep2d_adjs_f32.cu / li_sep2d_adjs_f64.cu, which instantiate it for one element type each.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "li_sep2d.cuh"

namespace li {

// ---------------------------------------------------------------------------------------------
// Adjoint, streaming form: the forward kernel's data flow run backwards.
//
// A CTA owns a strip of NT dst columns and a chunk of CR dst rows (about one wave of CTAs); a thread owns ONE dst
// column.  The src rows feeding the chunk stream, in ascending order, through the same ring of TMA bulk copies as
// in the forward kernel (no per-thread global loads: the LSU pipe only sees shared-memory reads).  For every src
// row the thread
//   1. gathers along dim 1 from the staged row: the NU staged values around its column (LDS.64, lanes stride one
//      vector: conflict free for a 2x operator) times its NU folded taps, kept in registers for the whole kernel;
//   2. adds the value into S2 accumulators (dst rows v .. v+S2-1) with the row's folded dim-2 weights.
// dst rows are finished one at a time -- for dst row v every src row whose folded window starts at v, like the
// tile kernels above -- and stored straight from the accumulator (coalesced along dim 1): no intermediate in
// shared memory, no barrier besides the ring's, no atomics; the result is deterministic.
// Values outside a thread's tap list are discarded with selects, not multiplied by zero, so a non-finite src
// element only reaches the dst elements it feeds in the reference.
//
// Chunks whose dst rows are each fed by exactly R src rows with no clamped tap (R = reg_R of the dim-2 table: the
// interior of a same-rate or 2x resampling; found from the table's irregularity prefix sums) run a statically
// scheduled loop: one stage of K rows is K/R dst rows, no per-row control flow.  Any other chunk (array edges,
// irregular rates) runs the generic loop over the per-row counts.  Both loops exist in two copies, chosen per warp:
// one for warps whose threads all have the usual tap pattern (padding only at the ends of the staged window, cleared
// with three bit masks) and one with a select per tap.
#ifndef ADS_NT_
#define ADS_NT_ 256
#endif
#ifndef ADS_NSTAGE_
#define ADS_NSTAGE_ 2
#endif
#ifndef ADS_K_
#define ADS_K_ 8
#endif
#ifndef ADS_WAVES_
#define ADS_WAVES_ 1
#endif
constexpr int ADS_NT = ADS_NT_;
constexpr int ADS_NSTAGE = ADS_NSTAGE_;
#ifndef ADS_OCC_
#define ADS_OCC_ 4
#endif
constexpr int ADS_PER_SM = ADS_OCC_ * 256 / ADS_NT;   // CTAs per SM of the 4-byte, <= 10-tap kernels

template <int S2, bool D3 = false>
struct AdsK {   // src rows per stage: a multiple of 2*S2 so that the static schedule has whole slot rotations
    static constexpr int value = D3 ? ((S2 == 6 || S2 == 3) ? 6 : (S2 == 8 ? 8 : 4))
                                    : ((S2 == 6 || S2 == 3) ? 12 : (S2 == 8 ? 16 : ADS_K_));
};
#ifndef ADS_TAPER_
#define ADS_TAPER_ 8   // chunk taper of single-wave launches, in eighths of the measured imbalance (0 = equal chunks)
#endif
#ifndef ADS3_NSR_
#define ADS3_NSR_ 2
#endif
constexpr int ADS3_NSR = ADS3_NSR_;   // D3: stages of the ring of raw src rows
constexpr int ADS3_NC = 4;            // D3: most src planes feeding one dst plane

// D3 = true: a third table along dim 3, (A1*A2*A3)'.  The "batch" index of the grid is a dst plane j3; its staged rows
// are the combinations  sum_i W3f[i, j3 - bfold3[i]] * src[:, row, i]  over the src planes i feeding that dst plane (a
// block-uniform list of at most ADS3_NC planes, ascending = the order of the reference's scatter along dim 3): the raw
// rows land in a second ring (TMA bulk copies, one per plane and row) and every thread combines a few 16-byte vectors
// of stage j+1 into the row buffer while stage j is consumed.  Only the generic schedule (R = 0) is built.
template <typename T, int S1, int S2, int NU, int R, bool D3>
#ifndef ADS3_OCC_
#define ADS3_OCC_ 3
#endif
__global__ void __launch_bounds__(ADS_NT, D3 ? ((sizeof(T) == 4 && NU <= 10 && S2 <= 4) ? ADS3_OCC_ : 2)
                                             : ((sizeof(T) == 4 && NU <= 10 && S2 <= 4) ? ADS_PER_SM : ADS_PER_SM / 2))
sep2d_adj_stream_kernel(const SepArgs<T> a, const int CR, const int NRMAX)
{
    constexpr int NT = ADS_NT, K = AdsK<S2, D3>::value;
    constexpr int NSTAGE = D3 ? 2 : ADS_NSTAGE;              // row buffers the pipeline reads
    constexpr int NBAR = D3 ? ADS3_NSR : ADS_NSTAGE;         // stages filled by TMA
    static_assert(!D3 || R == 0, "the 3-D form runs the generic schedule");
    constexpr int V = 2;   // elements per shared-memory load of the gather: LDS.64 for float, LDS.128 for double
    static_assert(NU % V == 0 && NU >= 4 && NU % 2 == 0, "bad NU");
    static_assert(R == 0 || (K % R == 0 && (K / R) % S2 == 0), "bad stage size");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P1 = a.F1MAX;                                  // elements per staged row (multiple of 16 bytes)
    T *sS = reinterpret_cast<T *>(smem_raw);                 // [NSTAGE][K][P1]
    T *sR = sS + (size_t)NSTAGE * K * P1;                    // D3: [NBAR][nc3][K][P1] ring of raw rows
    T *sW2 = sR + (size_t)(D3 ? NBAR * a.nc3 : 0) * K * P1;  // [NRMAX][S2] alpha * folded dim-2 weights
    int *sNv = reinterpret_cast<int *>(sW2 + (size_t)NRMAX * S2);   // [NRMAX] generic loop: dst rows completed before a src row
    uint64_t *bar = reinterpret_cast<uint64_t *>(sNv + NRMAX);      // [NBAR] (NRMAX is even)

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int strip = bid % a.tiles1;
    bid /= a.tiles1;
    const int chunk = bid % a.tiles2 + a.chunk0;
    const int64_t batch = bid / a.tiles2;
    const int c0 = strip * NT, c1 = min(c0 + NT, a.m1);      // owned dst columns
#ifdef ADS_PROBE
    unsigned long long probe_t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(probe_t0));
#endif
    const int va = a.cstart[chunk], vb = a.cstart[chunk + 1];   // owned dst rows (chunks of about equal cost, <= CR rows)
    if (va >= vb) return;                                    // (empty chunk: block-uniform, before any barrier)
    const int v0 = va - (S2 - 1);                            // first dst row the chunk's src rows reach
    const int64_t plane = (int64_t)a.n1 * a.n2;
    const T *src = a.src + (D3 ? (batch / a.m3) * a.n3 : batch) * plane;
    T *dst = a.dst + batch * (int64_t)a.m1 * a.m2;
    // D3: the src planes feeding dst plane j3 and their folded weights (block-uniform, every thread its own copy)
    int nc = 0;
    int pl3[ADS3_NC];
    T w3[ADS3_NC];
    if constexpr (D3) {
        const int j3 = (int)(batch % a.m3);
        const int f0 = j3 == 0 ? 0 : a.bp3[j3], f1 = j3 == a.m3 - 1 ? a.n3 : a.bp3[j3 + a.S3];
#pragma unroll
        for (int t = 0; t < ADS3_NC; ++t) pl3[t] = 0, w3[t] = (T)0;
        for (int i = f0; i < f1; ++i) {
            const int b = a.base3[i];
            if (j3 >= clampi(b, a.m3 - 1) && j3 <= clampi(b + a.S3 - 1, a.m3 - 1)) {
                const T w = a.W3f[(int64_t)i * a.S3 + (j3 - a.bfold3[i])];
#pragma unroll
                for (int t = 0; t < ADS3_NC; ++t)
                    if (t == nc) pl3[t] = i, w3[t] = w;
                ++nc;
            }
        }
        nc = min(nc, ADS3_NC);   // (the launcher checked the table: never more)
    }

    const int u_lo = c0 == 0 ? 0 : a.bp1[c0];                // first src column with a tap in the strip
    constexpr int AL = 16 / (int)sizeof(T);
    const int ulo_a = u_lo & ~(AL - 1);                      // first staged src column (16-byte aligned)
    const int cv1 = min(P1, a.n1 - ulo_a);                   // staged columns that exist (n1 is a multiple of AL)
    const int i2lo = a.bp2[va], NR = a.bp2[vb - 1 + S2] - i2lo;   // src rows [i2lo, i2lo + NR)
    const int NJ = (NR + K - 1) / K;
    bool regular = false;
    if constexpr (R > 0) regular = v0 > 0 && vb <= a.m2 - S2 && a.irrp2[vb] == a.irrp2[v0];

    if (tid == 0) {
#pragma unroll
        for (int sgi = 0; sgi < NBAR; ++sgi) mbar_init(&bar[sgi], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // D3: the copies of a stage are shared out over the warps (lane 0 of warp w sends copy q = w, w + NT/32, ..: one or
    // two per stage, so that no warp falls behind the others); what does not change from stage to stage is kept in
    // registers.  A dst plane that nothing feeds still runs the pipeline, on plane 0.
    constexpr int NQ = D3 ? (ADS3_NC * K + NT / 32 - 1) / (NT / 32) : 1;
    const T *gq[NQ];
    int rq[NQ];
    uint32_t sq[NQ];
    const uint32_t bar0 = smem_u32(bar);
    const int ncc = max(nc, 1);
    if constexpr (D3) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            const int q = (tid >> 5) + i * (NT / 32);
            const int t = q / K, r = q - t * K;
            rq[i] = q < ncc * K ? r : K;   // (K: never sent)
            int pl = pl3[0];
#pragma unroll
            for (int u = 1; u < ADS3_NC; ++u) pl = t == u ? pl3[u] : pl;
            gq[i] = src + pl * plane + (int64_t)(i2lo + r) * a.n1 + ulo_a;
            sq[i] = smem_u32(sR + (size_t)q * P1);
        }
    }
    auto issue = [&](int j) {   // producer: stage j <- src rows i2lo + j*K ..
        const uint32_t bytes = (uint32_t)(cv1 * sizeof(T));
        const int nr = min(K, NR - j * K);
        if constexpr (D3) {     // (lane 0 of every warp)
            const uint32_t slot = (uint32_t)(j % NBAR);
            const uint32_t b = bar0 + slot * 8u, so = slot * (uint32_t)(a.nc3 * K * P1 * sizeof(T));
            if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes * nr * ncc) : "memory");
#pragma unroll
            for (int i = 0; i < NQ; ++i)
                if (rq[i] < nr)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     sq[i] + so),
                                 "l"(gq[i] + (int64_t)j * K * a.n1), "r"(bytes), "r"(b)
                                 : "memory");
        } else {                // (one thread)
            uint64_t *b = &bar[j % NBAR];
            T *buf = sS + (size_t)(j % NSTAGE) * K * P1;
            mbar_expect_tx(b, bytes * nr);
            for (int r = 0; r < nr; ++r)
                bulk_g2s(buf + r * P1, src + (int64_t)(i2lo + j * K + r) * a.n1 + ulo_a, bytes, b);
        }
    };
    // D3: row buffer j <- sum over the feeding planes of raw stage j, ascending plane order; a dst plane without
    // feeders gets zeros (never 0 * src: a non-finite sample reaches only the elements it feeds in the reference)
    const int KP = K * P1;
    auto combine_nc = [&](int j, auto NCt) {
        constexpr int NC = decltype(NCt)::value;
        const T *raw = sR + (size_t)(j % NBAR) * a.nc3 * KP;
        T *out = sS + (size_t)(j % NSTAGE) * KP;
        for (int idx = 4 * tid; idx < KP; idx += 4 * NT) {
            Pack<T, 4> acc;
            if constexpr (NC > 0) {
                acc = ld_pack<T, 4>(raw + idx);
#pragma unroll
                for (int e = 0; e < 4; ++e) acc.v[e] *= w3[0];
#pragma unroll
                for (int t = 1; t < NC; ++t) {
                    const Pack<T, 4> q = ld_pack<T, 4>(raw + t * KP + idx);
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc.v[e] = fmaT(w3[t], q.v[e], acc.v[e]);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) acc.v[e] = (T)0;
            }
            st_pack<T, 4>(out + idx, acc);
        }
    };
    auto combine = [&](int j) {
        switch (nc) {
        case 0: combine_nc(j, std::integral_constant<int, 0>{}); break;
        case 1: combine_nc(j, std::integral_constant<int, 1>{}); break;
        case 2: combine_nc(j, std::integral_constant<int, 2>{}); break;
        case 3: combine_nc(j, std::integral_constant<int, 3>{}); break;
        default: combine_nc(j, std::integral_constant<int, 4>{}); break;
        }
    };
    if (D3 ? (tid & 31) == 0 : tid == 0)
        for (int j = 0; j < (D3 ? NBAR : NSTAGE - 1) && j < NJ; ++j) issue(j);

    // chunk metadata: rows per dst row (generic loop only), dim-2 weights
    // (generic loop: for every src row of the chunk, the number of dst rows its folded window starts after the one of
    // the previous row -- dst rows that are complete before the row is added; bit 30 marks a window folded at an
    // edge of the array, where only the slots the clamped taps land on are fed)
    if (!regular)
        for (int k = tid; k < NR; k += NT) {
            const int b = a.base2[i2lo + k], bf = clampi(b, a.m2 - S2);
            const int prev = k == 0 ? v0 : clampi(a.base2[i2lo + k - 1], a.m2 - S2);
            sNv[k] = (bf - prev) | (b != bf ? (1 << 30) : 0);
        }
    for (int k = tid; k < NR * S2; k += NT) sW2[k] = a.alpha * a.W2f[(int64_t)i2lo * S2 + k];

    // per-thread dim-1 taps: src columns [e0, e1) feed dst column c; staged window starts at a0 <= e0 (vector aligned)
    const int c = c0 + tid;
    const bool have = c < c1;
    const int cc = min(c, a.m1 - 1);
    const int e0 = cc == 0 ? 0 : a.bp1[cc], e1 = cc == a.m1 - 1 ? a.n1 : a.bp1[cc + S1];
    const int a0 = e0 & ~(V - 1);
    T w1[NU];
    unsigned vm = 0;
#pragma unroll
    for (int t = 0; t < NU; ++t) {
        const int i = a0 + t;
        w1[t] = (T)0;
        if (have && i >= e0 && i < e1) {
            // (a folded window can hold slots that none of the row's clamped taps lands on: those are not fed)
            const int b = a.base1[i];
            if (cc >= clampi(b, a.m1 - 1) && cc <= clampi(b + S1 - 1, a.m1 - 1)) {
                w1[t] = a.W1f[(int64_t)i * S1 + (cc - a.bfold1[i])];
                vm |= 1u << t;
            }
        }
    }
    // Padding of the staged window is cleared with bit masks (one LOP3 per masked position).  Level 1, the pattern of
    // an integer up-sampling rate: only the first and the last two values can be padding (3 masks, static loop).
    // Level 2, any rate: the first and the last NTAIL values (generic loop).  Warps holding any other pattern (array
    // edges) take the copy with a select per tap (level 0).
    constexpr int NTAIL = NU <= 6 ? NU - 1 : NU / 2;
    constexpr unsigned MID = ((1u << (NU - 2)) - 1u) & ~1u, MID2 = ((1u << (NU - NTAIL)) - 1u) & ~1u;
    const bool warp_clean = __all_sync(0xffffffffu, (vm & MID) == MID);
    const bool warp_clean2 = __all_sync(0xffffffffu, (vm & MID2) == MID2);
    using Bits = typename std::conditional<sizeof(T) == 4, uint32_t, uint64_t>::type;
    Bits k0 = (vm & 1u) ? ~(Bits)0 : (Bits)0;
    Bits kt[NTAIL];
#pragma unroll
    for (int t = 0; t < NTAIL; ++t) kt[t] = ((vm >> (NU - NTAIL + t)) & 1u) ? ~(Bits)0 : (Bits)0;
    int rowstep = P1;
    int dstep = a.m1;
    // (opaque to the optimiser: otherwise the masks turn back into compare + select pairs and vm into NU flags)
    if constexpr (sizeof(T) == 4) {
        asm volatile("" : "+r"(k0), "+r"(vm), "+r"(rowstep), "+r"(dstep));
#pragma unroll
        for (int t = 0; t < NTAIL; ++t) asm volatile("" : "+r"(kt[t]));
    } else {
        asm volatile("" : "+l"(k0), "+r"(vm), "+r"(rowstep), "+r"(dstep));
#pragma unroll
        for (int t = 0; t < NTAIL; ++t) asm volatile("" : "+l"(kt[t]));
    }
    auto keep = [](T x, Bits m) {
        Bits b;
        memcpy(&b, &x, sizeof b);
        b &= m;
        memcpy(&x, &b, sizeof b);
        return x;
    };
    const int off = a0 - ulo_a;
    const bool beta = a.beta != (T)0;

    T acc[S2];
#pragma unroll
    for (int s = 0; s < S2; ++s) acc[s] = (T)0;
    T *dp = dst + (int64_t)va * a.m1 + c;
    __syncthreads();

    // dim 1 of the staged row at p
    auto gather = [&](const T *p, auto clean) {
#ifdef ADS_ABLATE
        if (a.vec_ok == 77) return p[0];   // timing probe: the ring without the dim-1 gather
#endif
        T x[NU];
#pragma unroll
        for (int t = 0; t < NU; t += V) {
            const Pack<T, V> q = ld_pack<T, V>(p + t);
#pragma unroll
            for (int e = 0; e < V; ++e) x[t + e] = q.v[e];
        }
        if constexpr (decltype(clean)::value == 1) {
            x[0] = keep(x[0], k0);
            x[NU - 2] = keep(x[NU - 2], kt[NTAIL - 2]);
            x[NU - 1] = keep(x[NU - 1], kt[NTAIL - 1]);
        } else if constexpr (decltype(clean)::value == 2) {
            x[0] = keep(x[0], k0);
#pragma unroll
            for (int t = 0; t < NTAIL; ++t) x[NU - NTAIL + t] = keep(x[NU - NTAIL + t], kt[t]);
        } else {
            unsigned m = vm;
            asm volatile("" : "+r"(m));   // (keeps the NU bit tests out of registers)
#pragma unroll
            for (int t = 0; t < NU; ++t) x[t] = ((m >> t) & 1u) ? x[t] : (T)0;
        }
        if constexpr (sizeof(T) == 4) {
            // two chains (even / odd staged values) as packed FP32x2 operations (FMUL2 / FFMA2, sm_100): the pairs
            // are exactly what the 8-byte shared-memory loads deliver
            float2 acc2 = __fmul2_rn(make_float2(x[0], x[1]), make_float2(w1[0], w1[1]));
#pragma unroll
            for (int t = 2; t < NU; t += 2)
                acc2 = __ffma2_rn(make_float2(x[t], x[t + 1]), make_float2(w1[t], w1[t + 1]), acc2);
            return (T)(acc2.x + acc2.y);
        } else {
            T va0 = x[0] * w1[0], va1 = x[1] * w1[1];   // two independent chains
#pragma unroll
            for (int t = 2; t < NU; t += 2) {
                va0 = fmaT(x[t], w1[t], va0);
                va1 = fmaT(x[t + 1], w1[t + 1], va1);
            }
            return va0 + va1;
        }
    };
    auto next_stage = [&](int j) {   // block-uniform; returns the first staged row of stage j (this thread's window)
        if constexpr (D3) {
            // stage j is combined into its row buffer now (the buffer's last readers passed the previous barrier);
            // its raw slot is refilled after the barrier.  The copies run NBAR stages ahead of the arithmetic.
            mbar_wait(&bar[j % NBAR], (uint32_t)((j / NBAR) & 1));
            combine(j);
            __syncthreads();
            if ((tid & 31) == 0 && j + NBAR < NJ) issue(j + NBAR);
        } else {
            __syncthreads();             // every thread is done with stage j-1: its buffer may be refilled
            if (tid == 0 && j + NSTAGE - 1 < NJ) issue(j + NSTAGE - 1);
            mbar_wait(&bar[j % NSTAGE], (uint32_t)((j / NSTAGE) & 1));
        }
        return (const T *)(sS + (size_t)(j % NSTAGE) * K * P1 + off);
    };
    // beta != 0: the old values of the thread's dst column are read S2 rows ahead of their use (slot = row mod S2,
    // like the accumulators), so that the read-modify-write of a dst row never waits for memory
    const bool ld_old = beta && have;
    const T betav = a.beta;
    T oldv[S2];
#pragma unroll
    for (int i = 0; i < S2; ++i) {
        oldv[(S2 - 1 + i) % S2] = (T)0;
        if (ld_old && va + i < vb) oldv[(S2 - 1 + i) % S2] = dp[(int64_t)i * dstep];   // (beta == 0: dst is never read)
    }
    // (compiled once for beta == 0 and once for beta != 0: `hasbeta` is a compile-time tag)
    auto emit = [&](int vv, T value, int slot, auto hasbeta) {   // dst row vv is complete; slot = (vv - v0) % S2, static
        if (vv >= va) {
            if constexpr (!decltype(hasbeta)::value) {   // dst is never read
                if (have) *dp = value;
            } else {
                const T o = fmaT(betav, oldv[slot], value);
                if (have) *dp = o;
                if (ld_old && vv + S2 < vb) oldv[slot] = dp[(int64_t)S2 * dstep];
            }
            dp += dstep;
        }
    };

    // ---- static schedule: stage j holds the R rows of each of the dst rows v0 + j*DR .. + DR-1
    auto run_regular = [&](auto clean, auto hasbeta) {
        if constexpr (R > 0) {
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
                            const T val = gather(prow, clean);
                            prow += rowstep;
                            T w2[S2];
                            ld_weights<T, S2>(pw + (d * R + r) * S2, w2);
#pragma unroll
                            for (int s = 0; s < S2; ++s) acc[(d + s) % S2] = fmaT(w2[s], val, acc[(d + s) % S2]);
                        }
                        emit(vv, acc[d % S2], d % S2, hasbeta);
                        acc[d % S2] = (T)0;
                    }
                }
                pw += K * S2;
            }
        }
    };

    // ---- generic schedule: the src rows in order; acc[0] belongs to dst row `vcur`, and the window slides (register
    // moves) by as many dst rows as the folded window of the next src row starts later -- those rows are complete and
    // are stored.  Per dst row the contributions arrive in ascending src-row order, as in the reference's scatter.
    auto run_generic = [&](auto clean, auto hasbeta) {
        const T *pw = sW2;
        const int *pa = sNv;
        int vcur = v0;
        // (dp follows vcur: the leading rows v0 .. va-1 of the window belong to the previous chunk, or lie before the array,
        // and are never stored -- a predicate, not a branch)
        T *dq = dp - (int64_t)(va - v0) * dstep;
        T oldq[S2];             // beta != 0: old values of the next S2 rows to be stored, read S2 rows ahead
        if constexpr (decltype(hasbeta)::value) {
#pragma unroll
            for (int i = 0; i < S2; ++i) oldq[i] = oldv[(S2 - 1 + i) % S2];
        }
        auto advance = [&]() {   // dst row vcur (slot 0) is complete
            const bool owned = vcur >= va;
            if constexpr (!decltype(hasbeta)::value) {   // dst is never read
                if (have && owned) *dq = acc[0];
            } else {
                if (owned) {
                    const T o = fmaT(betav, oldq[0], acc[0]);
                    if (have) *dq = o;
#pragma unroll
                    for (int i = 0; i + 1 < S2; ++i) oldq[i] = oldq[i + 1];
                    oldq[S2 - 1] = (T)0;
                    if (ld_old && vcur + S2 < vb) oldq[S2 - 1] = dq[(int64_t)S2 * dstep];
                }
            }
            dq += dstep;
#pragma unroll
            for (int s = 0; s + 1 < S2; ++s) acc[s] = acc[s + 1];
            acc[S2 - 1] = (T)0;
            ++vcur;
        };
        for (int j = 0; j < NJ; ++j) {
            const T *prow = next_stage(j);
            const int nr = min(K, NR - j * K);
#pragma unroll
            for (int r = 0; r < K; ++r) {
                if (r < nr) {   // block-uniform: only the last stage is partial
                    const int ad = pa[r];
                    int n = ad & 0xffff;
                    if (n) {               // (mostly one row)
                        advance();
#pragma unroll 1
                        for (--n; n > 0; --n) advance();
                    }
                    const T val = gather(prow, clean);
                    prow += rowstep;
                    T w2[S2];
                    ld_weights<T, S2>(pw, w2);
                    pw += S2;
                    if (ad < (1 << 30)) {   // interior window: every slot is fed
#pragma unroll
                        for (int s = 0; s < S2; ++s) acc[s] = fmaT(w2[s], val, acc[s]);
                    } else {                // first / last window of the array: only the slots the clamped taps land on
                        const int b = a.base2[i2lo + j * K + r];
                        const int lo = clampi(b, a.m2 - 1) - vcur, hi = clampi(b + S2 - 1, a.m2 - 1) - vcur;
#pragma unroll
                        for (int s = 0; s < S2; ++s)
                            if (s >= lo && s <= hi) acc[s] = fmaT(w2[s], val, acc[s]);
                    }
                }
            }
            pa += K;
        }
#pragma unroll 1
        while (vcur < vb) advance();   // the last (at most S2) rows of the chunk
    };

    // (every copy runs the same sequence of barriers, so warps of one CTA may take different copies)
    auto run = [&](auto hasbeta) {
        if (regular) {
            if (warp_clean)
                run_regular(std::integral_constant<int, 1>{}, hasbeta);
            else
                run_regular(std::integral_constant<int, 0>{}, hasbeta);
        } else {
            if (warp_clean2)
                run_generic(std::integral_constant<int, 2>{}, hasbeta);
            else
                run_generic(std::integral_constant<int, 0>{}, hasbeta);
        }
    };
    if (beta)
        run(std::true_type{});
    else
        run(std::false_type{});
#ifdef ADS_PROBE   // timing probe (development builds): start / end of every CTA on the global timer
    if (a.rowlist && tid == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        unsigned long long *p = (unsigned long long *)a.rowlist + 2 * (size_t)blockIdx.x;
        p[0] = probe_t0, p[1] = t1;
    }
#endif
}

// ---------------------------------------------------------------------------------------------
template <typename T, int S1, int S2, int NU, int R, bool D3>
static int launch_stream(SepArgs<T> &a, int CR, int NRMAX, size_t smem, int64_t blocks, cudaStream_t st)
{
    auto kern = sep2d_adj_stream_kernel<T, S1, S2, NU, R, D3>;
    int rc = set_smem(kern, smem);
    if (rc != LI_OK) return rc;
    kern<<<(unsigned)blocks, ADS_NT, smem, st>>>(a, CR, NRMAX);
    count_launch();
    LI_CUDA(cudaGetLastError());
    return LI_OK;
}

// resident CTAs per SM of one kernel instance with `smem` bytes of dynamic shared memory (registers decide for the
// wide instances); cached per instance
template <typename T, int S1, int S2, int NU, int R, bool D3>
static int stream_occupancy(size_t smem)
{
    static thread_local int cached = 0, cached_dev = -1;
    static thread_local size_t cached_smem = 0;
    const int dev = current_device();
    if (cached && cached_smem == smem && cached_dev == dev) return cached;
    cached_dev = dev;
    auto kern = sep2d_adj_stream_kernel<T, S1, S2, NU, R, D3>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, ADS_NT, smem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 1;
    }
    cached = occ, cached_smem = smem;
    return occ;
}

// (D3: nbatch counts dst planes, i.e. m3 times the number of 3-D arrays; a.base3 / bp3 / bfold3 / W3f / n3 / m3 / S3 /
// nc3 are set)
template <typename T, int S1, int S2, bool D3 = false>
static int stream_T(SepArgs<T> &a, const li_table *t1, const li_table *t2, int64_t nbatch, cudaStream_t st, bool *done)
{
    constexpr int V = 2, AL = 16 / (int)sizeof(T), K = AdsK<S2, D3>::value;
    constexpr int NROWBUF = D3 ? 2 : ADS_NSTAGE, NBAR = D3 ? ADS3_NSR : ADS_NSTAGE;
    const size_t stage_elems = (size_t)(NROWBUF + (D3 ? NBAR * a.nc3 : 0)) * K;   // staged rows of all rings
    constexpr int NU_S = (S1 + 2 + 1) & ~1, NU_L = 2 * S1 + 2;   // taps kept per thread: same-rate / 2x operators
    if (a.m1 < S1 || a.m2 < S2) return LI_OK;                    // folding needs a full window inside the array
    if (a.n1 % AL != 0 || ((uintptr_t)a.src) % 16 != 0) return LI_OK;   // bulk copies need 16-byte aligned row segments
    int rc = table_ensure_fold(t1, st);
    if (rc == LI_OK) rc = table_ensure_fold(t2, st);
    if (rc != LI_OK) return rc;
    int maxcount = 0, span = 0;
    rc = table_feed_rows(t1, 1, st, &maxcount);
    if (rc == LI_OK) rc = table_feed_rows(t1, ADS_NT, st, &span);
    if (rc != LI_OK) return rc;
    const int need = maxcount + (V - 1);
    const bool trace = opt(OPT_TRACE) != 0;
    if (need > NU_L) {   // decline (down-sampling along dim 1, or many rows hanging over an edge)
        if (trace) fprintf(stderr, "[li_b200] adj stream declines: %d taps per dst column > %d\n", need, NU_L);
        return LI_OK;
    }
    const int NU = need <= NU_S ? NU_S : NU_L;
    const int P1 = (span + (AL - 1) + NU + (AL - 1)) & ~(AL - 1);
    a.tiles1 = (a.m1 + ADS_NT - 1) / ADS_NT;
    // chunks of dst rows: about ADS_PER_SM CTAs per SM (a thread walks its rows one after the other, so a small problem
    // cut into few long chunks is latency-bound), but at least 4*S2 rows per chunk (each chunk re-reads S2-1 halo rows)
    // (the grid is sized for the CTAs that are really resident: the wide instances hold 2 or 3 per SM, not 4)
    const int R = (!D3 && t2->irrp && (t2->reg_R == 1 || t2->reg_R == 2)) ? t2->reg_R : 0;
    int occ = ADS_PER_SM;
    {
        const size_t smem_est = (stage_elems * P1) * sizeof(T) + 8 * 1024;
#define LI_OCC(NUU, RR) \
    if (NU == NUU && R == RR) occ = stream_occupancy<T, S1, S2, NUU, RR, D3>(smem_est);
        LI_OCC(NU_S, 0) LI_OCC(NU_L, 0)
        if constexpr (!D3) {
            LI_OCC(NU_S, 1) LI_OCC(NU_S, 2) LI_OCC(NU_L, 1) LI_OCC(NU_L, 2)
        }
#undef LI_OCC
    }
    occ = std::min(occ, ADS_PER_SM);
    // (a kernel that holds fewer CTAs per SM gets as many rounds as needed to keep the chunks short)
    const int rounds = std::max(1, ADS_PER_SM / occ);
    const int64_t target = (int64_t)sep2d_num_sms() * occ * rounds * ADS_WAVES_;
    // (rounded DOWN: a grid slightly larger than the resident slots would put its last CTAs into a second round and
    // nearly double the kernel time -- measured on 8192^2 at rate 1: 608 CTAs on 592 slots)
    int chunks = (int)std::max<int64_t>(1, target / (a.tiles1 * nbatch));
    // (many planes: chunks of at most 256 dst rows keep the per-chunk tables in shared memory small)
    if (D3) chunks = std::max(chunks, (a.m2 + 255) / 256);
    // (at least 4*S2 rows per chunk: each chunk re-reads S2-1 halo rows)
    const int nchunks = std::max(1, std::min(chunks, a.m2 / std::max(8, 4 * S2)));
    a.tiles2 = nchunks;
    int CR = 0, nrmax = 0;
    // chunks that must run the generic loop (array edges, irregular rates) are charged 30 % of a chunk, i.e. made
    // shorter: all CTAs of the single wave then finish together
    const li_table::ChunkPlan *plan = nullptr;
    // One image in one wave: the CTAs of a grid are handed to the SMs in index order, one per SM and round, and an SM
    // favours its older CTAs (measured with the ADS_PROBE build on config #2: with equal chunks the first CTA of an SM
    // finishes at 48 us, the second at 50.5, the third at 54, the fourth at 60.5, and the SM runs under-occupied from
    // 48 us on).  The chunks are therefore tapered: longer for the CTAs scheduled first, shorter for the last ones.
    std::vector<double> taper;
    int taper_key = 0;
    {
        const int strength = ((opt(OPT_EXP) >> 12) & 15) ? ((opt(OPT_EXP) >> 12) & 15) - 1 : ADS_TAPER_;   // (x/8; exp bits 12..15: value + 1)
        const int nsm = sep2d_num_sms();
        if (!D3 && nbatch == 1 && occ >= 2 && strength > 0 && (int64_t)nchunks * a.tiles1 <= (int64_t)nsm * occ && a.tiles1 < 128) {
            static const double shape[4] = {0.18, 0.10, 0.0, -0.14};   // (second iteration of the probe)
            taper.resize((size_t)nchunks);
            for (int c = 0; c < nchunks; ++c) {
                double w = 0.0;
                for (int sx = 0; sx < a.tiles1; ++sx) {
                    const int slot = std::min((c * a.tiles1 + sx) / nsm, occ - 1);
                    // position of the slot within the rounds, mapped onto the four measured ones
                    const double pos = occ > 1 ? 3.0 * slot / (occ - 1) : 0.0;
                    const int lo = (int)pos;
                    const double f = pos - lo;
                    w += 1.0 + strength / 8.0 * (shape[lo] * (1.0 - f) + shape[std::min(lo + 1, 3)] * f);
                }
                taper[(size_t)c] = w / a.tiles1;
            }
            // (the chunks at the array edges run the slower generic loop: no bonus on top of their penalty)
            taper[0] = std::min(taper[0], 1.0);
            taper[(size_t)nchunks - 1] = std::min(taper[(size_t)nchunks - 1], 1.0);
            taper_key = a.tiles1 | (strength << 7) | (occ << 11);   // (< 2^14)
        }
    }
    rc = table_chunk_plan(t2, nchunks, 0.3, st, &a.cstart, &CR, &nrmax, &plan, taper.empty() ? nullptr : &taper, taper_key);
    if (rc != LI_OK) return rc;
    if (CR > 4096) return LI_OK;
    const int NRMAX = (std::max(nrmax, 1) + 1) & ~1;
    const size_t smem = (stage_elems * P1 + (size_t)NRMAX * S2) * sizeof(T) +
                        (size_t)NRMAX * sizeof(int) + NBAR * sizeof(uint64_t);
    if (trace)
        fprintf(stderr, "[li_b200] adj stream%s: NU %d P1 %d CR %d NRMAX %d smem %zu tiles %d x %d x %lld R %d occ %d\n",
                D3 ? " 3-D" : "", NU, P1, CR, NRMAX, smem, a.tiles1, a.tiles2, (long long)nbatch, R, occ);
    if (smem > (size_t)kSmemLimit) return LI_OK;
    if (BandCtl *bc = D3 ? nullptr : sep2d_band()) {   // host-buffer pipeline: describe the chunks, or launch a band of them
        if (bc->mode == 1) {
            bc->row_begin.assign(plan->starts_host.begin(), plan->starts_host.end());
            bc->src_end.assign(plan->src_end_host.begin(), plan->src_end_host.end());
            *done = true;
            return LI_OK;
        }
        a.chunk0 = bc->chunk0;
        a.tiles2 = bc->nchunks;
    }
    const int64_t blocks = (int64_t)a.tiles1 * a.tiles2 * nbatch;
    LI_REQUIRE(blocks <= 0x7fffffffLL, LI_ERR_UNSUPPORTED, "too many tiles");
    a.bp1 = t1->rowstart, a.bfold1 = t1->bfold, a.W1f = (const T *)t1->Wfold;
    a.bp2 = t2->bpf, a.W2f = (const T *)t2->Wfold, a.irrp2 = t2->irrp;
    a.F1MAX = P1;
#ifdef ADS_ABLATE
    if (opt(OPT_ADJ_DEBUG) == 4) a.vec_ok = 77;
#endif
#ifdef ADS_PROBE
    unsigned long long *probe = nullptr;
    if (opt(OPT_ADJ_DEBUG) == 5) {
        LI_CUDA(cudaMalloc((void **)&probe, (size_t)blocks * 2 * sizeof(unsigned long long)));
        a.rowlist = (const int32_t *)probe;   // (unused by this kernel: carries the probe buffer)
    }
#endif
#define LI_ADS(NUU, RR) \
    if (NU == NUU && R == RR) rc = launch_stream<T, S1, S2, NUU, RR, D3>(a, CR, NRMAX, smem, blocks, st);
    LI_ADS(NU_S, 0) LI_ADS(NU_L, 0)
    if constexpr (!D3) {
        LI_ADS(NU_S, 1) LI_ADS(NU_S, 2) LI_ADS(NU_L, 1) LI_ADS(NU_L, 2)
    }
#undef LI_ADS
#ifdef ADS_PROBE
    if (probe) {
        std::vector<unsigned long long> h((size_t)blocks * 2);
        cudaMemcpy(h.data(), probe, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        cudaFree(probe);
        a.rowlist = nullptr;
        unsigned long long t0 = ~0ull, t1 = 0;
        for (int64_t b = 0; b < blocks; ++b) t0 = std::min(t0, h[2 * b]), t1 = std::max(t1, h[2 * b + 1]);
        fprintf(stderr, "[li_b200] adj probe: %lld CTAs, kernel span %.1f us\n", (long long)blocks, (t1 - t0) * 1e-3);
        for (int c = 0; c < a.tiles2; ++c) {
            double s0 = 0, s1 = 0, mx = 0;
            for (int sx = 0; sx < a.tiles1; ++sx) {
                const int64_t b = (int64_t)c * a.tiles1 + sx;
                s0 += (h[2 * b] - t0) * 1e-3, s1 += (h[2 * b + 1] - t0) * 1e-3, mx = std::max(mx, (h[2 * b + 1] - t0) * 1e-3);
            }
            fprintf(stderr, "[li_b200] adj probe chunk %3d: start %.1f end %.1f (latest %.1f) us\n", c, s0 / a.tiles1, s1 / a.tiles1, mx);
        }
        for (int sx = 0; sx < a.tiles1; ++sx) {
            double s1 = 0, mx = 0;
            for (int c = 0; c < a.tiles2; ++c) {
                const int64_t b = (int64_t)c * a.tiles1 + sx;
                s1 += (h[2 * b + 1] - h[2 * b]) * 1e-3, mx = std::max(mx, (h[2 * b + 1] - t0) * 1e-3);
            }
            fprintf(stderr, "[li_b200] adj probe strip %3d: mean duration %.1f us, latest end %.1f us\n", sx, s1 / a.tiles2, mx);
        }
    }
#endif
    if (rc == LI_OK) *done = true;
    return rc;
}

template <typename T>
int sep2d_adj_stream(SepArgs<T> &a, const li_table *t1, const li_table *t2, int64_t nbatch, cudaStream_t st, bool *done)
{
    const int S1 = t1->S, S2 = t2->S;
#define LI_SEP(A1, A2) \
    if (S1 == A1 && S2 == A2) return stream_T<T, A1, A2>(a, t1, t2, nbatch, st, done);
#ifdef LI_DEV_FAST   // development builds: the headline pair only
    LI_SEP(4, 4)
#else
    LI_SEP(2, 2) LI_SEP(4, 4) LI_SEP(6, 6) LI_SEP(2, 4) LI_SEP(4, 2) LI_SEP(4, 6) LI_SEP(6, 4) LI_SEP(3, 3)
    LI_SEP(2, 1) LI_SEP(3, 1) LI_SEP(4, 1) LI_SEP(6, 1)
#endif
#undef LI_SEP
    return LI_OK;
}

}  // namespace li
