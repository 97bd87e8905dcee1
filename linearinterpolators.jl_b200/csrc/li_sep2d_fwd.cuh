// li_sep2d_fwd.cuh -- fused separable interpolation, forward, streaming form (K7): the kernel and its launcher.
//
// Reference: the LazyAlgebra composition A1*A2[*A3] of test/benchmarks.jl:27-29, whose factors are the loops of
// src/tabulated.jl:394-439 / src/sparse.jl:719-759 (same mathematics as src/separable.jl:56-111).
// Included by li_sep2d.cu (two dimensions) and li_sep3d_f32.cu / li_sep3d_f64.cu (three dimensions, D3 = true).
#pragma once
#include <algorithm>
#include <type_traits>

#include "li_sep2d.cuh"

namespace li {

// ---------------------------------------------------------------------------------------------
// Forward: a streaming register pipeline.
//
// A CTA owns a strip of TO1 = 2*NT output columns and a chunk of output rows; a thread owns two adjacent
// output columns (i1, i1+1).  The source rows the chunk needs are streamed, in ascending order, through a
// ring of shared-memory stages filled by TMA bulk copies (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP)
// issued by one thread NSTAGE-1 stages ahead.  For every source row the thread
//   1. interpolates along dim 1 from the staged row (taps and weights in registers; the two columns share
//      their loads whenever the table never advances by more than one column per row, i.e. any up-sampling),
//   2. pushes the value into a window of S2 registers whose rotation is static (the row loop is unrolled by
//      a multiple of S2),
//   3. emits every output row whose dim-2 window has just become complete (block-uniform count).
// The dim-1 result never touches shared memory, no source row is re-computed, and src/dst cross HBM once.
// Flat/SafeFlat clamping is applied while staging (clamped row index for the copy, replicated edge columns).
constexpr int FWD_NT = 256;
#ifndef FWD_NSTAGE_
#define FWD_NSTAGE_ 3
#endif
#ifndef FWD_WAVES_
#define FWD_WAVES_ 2
#endif
constexpr int FWD_NSTAGE = FWD_NSTAGE_;
constexpr int FWD_MAXCR = 256;   // output rows per chunk (one thread per row when counting completions)

template <int S2, bool D3>
struct FwdK {   // source rows per stage: a multiple of S2 (static window rotation)
    static constexpr int value = D3 ? (S2 >= 3 ? S2 : 4) : (S2 == 4 ? 8 : (S2 > 4 ? S2 : (S2 == 3 ? 6 : 8)));
};
#ifndef FWD3_NSR_
#define FWD3_NSR_ 2
#endif
constexpr int FWD3_NSR = FWD3_NSR_;   // D3: stages of the ring of raw source rows
constexpr int SEP3_NC = 4;            // D3: most planes combined into one staged row

// D3 = true: a third table along dim 3 (A1*A2*A3).  The "batch" index of the grid is an output plane i3; the staged
// rows of that plane are combinations  sum_t W3[i3, t] * src[:, row, clamp(base3[i3] + t)]  of the rows of S3 source
// planes: the raw rows land in a second ring (TMA bulk copies, one per plane and row) and every thread combines a few
// 16-byte vectors of stage j+1 into the row buffer while stage j is consumed (one barrier per stage, as before).
// The planes shared by neighbouring output planes are re-read from L2 (CTAs of consecutive planes run side by side).
template <typename T, int S1, int S2, bool PAIR, bool BETA, bool D3>
__global__ void __launch_bounds__(FWD_NT, (sizeof(T) == 4 && S1 <= 4 && S2 <= 4) ? 4 : (sizeof(T) == 4 ? 3 : 2))
sep2d_fwd_kernel(const SepArgs<T> a, const int CR /* output rows per chunk */, const int NCNT /* staged-row counters */)
{
    constexpr int NT = FWD_NT, TO1 = 2 * NT, K = FwdK<S2, D3>::value;
    constexpr int NSTAGE = D3 ? 2 : FWD_NSTAGE;              // row buffers the pipeline reads
    constexpr int NBAR = D3 ? FWD3_NSR : FWD_NSTAGE;         // stages filled by TMA

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P1 = a.F1MAX;                                  // elements per staged row (multiple of 4)
    const int nc = D3 ? a.nc3 : 0;                           // D3: planes per staged row
    T *sS = reinterpret_cast<T *>(smem_raw);                 // [NSTAGE][K][P1] ring of source rows
    T *sR = sS + (size_t)NSTAGE * K * P1;                    // D3: [NBAR][nc][K][P1] ring of raw rows
    T *sW2 = sR + (size_t)(D3 ? NBAR : 0) * nc * K * P1;     // [CR][S2]  alpha * W2
    int *sCnt = reinterpret_cast<int *>(sW2 + (size_t)CR * S2);   // [NCNT] outputs completed by each source row
    uint64_t *bar = reinterpret_cast<uint64_t *>(sCnt + ((NCNT + 1) & ~1));   // [NBAR]

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int strip = bid % a.tiles1;
    bid /= a.tiles1;
    const int chunk = bid % a.tiles2 + a.chunk0;
    const int64_t batch = bid / a.tiles2;
    const int i1_0 = strip * TO1;
    const int i2a = chunk * CR, rows = min(CR, a.n2 - i2a);   // output rows [i2a, i2a + rows)
    const int64_t plane = (int64_t)a.m1 * a.m2;
    const int i3 = D3 ? (int)(batch % a.n3) : 0;
    const T *src = a.src + (D3 ? (batch / a.n3) * a.m3 : batch) * plane;
    T *dst = a.dst + batch * (int64_t)a.n1 * a.n2;
    int b3 = 0;
    T w3[SEP3_NC];
    if constexpr (D3) {
        b3 = a.base3[i3];
#pragma unroll
        for (int t = 0; t < SEP3_NC; ++t) w3[t] = t < nc ? a.W3[(int64_t)i3 * nc + t] : (T)0;
    }

    // footprint (tables are monotone): staged column c <-> source column lo1a + c, staged row r <-> lo2 + r
    const int lo1 = a.base1[i1_0], F1 = a.base1[min(i1_0 + TO1, a.n1) - 1] + S1 - lo1;
    const int lo1a = (a.vec_ok & 1) ? (lo1 & ~3) : lo1;
    const int lo2 = a.base2[i2a];
    const int NR = a.base2[i2a + rows - 1] + S2 - lo2;        // source rows needed
    const int NJ = (NR + K - 1) / K;                         // stages
    // columns of the staged row that exist in the source; the rest are replicated edge columns
    const int cv0 = max(0, -lo1a), cv1 = min(P1, a.m1 - lo1a);
    const bool bulk = (a.vec_ok & 1) != 0;                   // TMA bulk copies need 16-byte aligned row segments
    const int need1 = min(P1, F1 + (lo1 - lo1a) + 1);        // staged columns actually read (incl. the spare tap)
    const bool patch = cv0 > 0 || cv1 < need1;

    for (int k = tid; k < NCNT; k += NT) sCnt[k] = 0;
    if (tid == 0) {
#pragma unroll
        for (int sgi = 0; sgi < NBAR; ++sgi) mbar_init(&bar[sgi], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // D3: the copies of a stage are shared out over the warps (lane 0 of warp w sends copy q = w, w + NT/32, ..: one or
    // two per stage, so that no warp falls behind the others); what does not change from stage to stage is kept in
    // registers: the plane's first staged element, the row within the stage, the slot's address in the raw ring
    constexpr int NQ = D3 ? (SEP3_NC * K + NT / 32 - 1) / (NT / 32) : 1;
    const T *gq[NQ];
    int rq[NQ];
    uint32_t sq[NQ];
    const uint32_t bar0 = smem_u32(bar);
    if constexpr (D3) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            const int q = (tid >> 5) + i * (NT / 32);
            const int t = q / K;
            rq[i] = q < nc * K ? q - t * K : -1;
            gq[i] = src + clampi(b3 + min(t, nc - 1), a.m3 - 1) * plane + lo1a + cv0;
            sq[i] = smem_u32(sR + (size_t)q * P1 + cv0);
        }
    }
    auto issue = [&](int j) {   // producer: stage j <- source rows lo2 + j*K .. + K-1, clamped
        const uint32_t bytes = (uint32_t)((cv1 - cv0) * sizeof(T));
        if constexpr (D3) {     // (lane 0 of every warp)
            const uint32_t slot = (uint32_t)(j % NBAR);
            const uint32_t b = bar0 + slot * 8u, so = slot * (uint32_t)(nc * K * P1 * sizeof(T));
            if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes * K * nc) : "memory");
#pragma unroll
            for (int i = 0; i < NQ; ++i)
                if (rq[i] >= 0) {
                    const int row = clampi(lo2 + j * K + rq[i], a.m2 - 1);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     sq[i] + so),
                                 "l"(gq[i] + (int64_t)row * a.m1), "r"(bytes), "r"(b)
                                 : "memory");
                }
        } else {                // (one thread)
            uint64_t *b = &bar[j % NBAR];
            T *buf = sS + (size_t)(j % NSTAGE) * K * P1;
            mbar_expect_tx(b, bytes * K);
#pragma unroll
            for (int r = 0; r < K; ++r) {
                const int row = clampi(lo2 + j * K + r, a.m2 - 1);
                bulk_g2s(buf + r * P1 + cv0, src + (int64_t)row * a.m1 + lo1a + cv0, bytes, b);
            }
        }
    };
    // D3: row buffer j <- sum over the planes of raw stage j (ascending taps, like the reference's dim-3 pass), one
    // 16-byte vector at a time; the vectors of an edge strip that hold out-of-range columns read the clamped column
    // instead (the replicated first / last source column)
    const int KP = K * P1;
    // (the four elements idx .. idx+3 of a row, idx a multiple of 4, into the de-interleaved row buffer: even elements
    // to the first half of the row, odd ones to the second)
    auto st_deint = [&](T *out, int r, int c, const Pack<T, 4> &acc) {   // row r, elements c .. c+3 (c % 4 == 0)
        Pack<T, 2> ev, od;
        ev.v[0] = acc.v[0], ev.v[1] = acc.v[2], od.v[0] = acc.v[1], od.v[1] = acc.v[3];
        st_pack<T, 2>(out + r * P1 + (c >> 1), ev);
        st_pack<T, 2>(out + r * P1 + (c >> 1) + (P1 >> 1), od);
    };
    const int nv1 = P1 >> 2;                                 // vectors per staged row
    const int vr0 = tid / nv1, vc0 = tid - vr0 * nv1;        // this thread's first vector (row, vector in the row)
    const int dvr = NT / nv1, dvc = NT - dvr * nv1;          // ... and the step to its next one
    auto combine_nc = [&](int j, auto NCt) {
        constexpr int NC = decltype(NCt)::value;
        const T *raw = sR + (size_t)(j % NBAR) * NC * KP;
        T *out = sS + (size_t)(j % NSTAGE) * KP;
        for (int r = vr0, c = vc0; r < K;) {
            const int c4 = 4 * c, idx = r * P1 + c4;
            if (!patch || (c4 >= cv0 && c4 + 4 <= cv1)) {
                Pack<T, 4> acc = ld_pack<T, 4>(raw + idx);
#pragma unroll
                for (int e = 0; e < 4; ++e) acc.v[e] *= w3[0];
#pragma unroll
                for (int t = 1; t < NC; ++t) {
                    const Pack<T, 4> q = ld_pack<T, 4>(raw + t * KP + idx);
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc.v[e] = fmaT(w3[t], q.v[e], acc.v[e]);
                }
                st_deint(out, r, c4, acc);
            } else {   // (edge strips: a few vectors per row)
                for (int e = 0; e < 4; ++e) {
                    const int cs = r * P1 + min(max(c4 + e, cv0), cv1 - 1);
                    T acc = raw[cs] * w3[0];
#pragma unroll
                    for (int t = 1; t < NC; ++t) acc = fmaT(w3[t], raw[t * KP + cs], acc);
                    out[r * P1 + ((c4 + e) >> 1) + ((c4 + e) & 1) * (P1 >> 1)] = acc;
                }
            }
            r += dvr, c += dvc;
            if (c >= nv1) c -= nv1, ++r;
        }
    };
    auto combine = [&](int j) {
        if (nc == 2)
            combine_nc(j, std::integral_constant<int, 2>{});
        else if (nc == 3)
            combine_nc(j, std::integral_constant<int, 3>{});
        else
            combine_nc(j, std::integral_constant<int, 4>{});
    };
    auto stage_manual = [&](int j) {   // unaligned sources: plain loads by every thread
        T *buf = sS + (size_t)(j % NSTAGE) * K * P1;
        for (int idx = tid; idx < K * P1; idx += NT) {
            const int r = idx / P1, c = idx - r * P1;
            const int row = clampi(lo2 + j * K + r, a.m2 - 1);
            buf[idx] = src[(int64_t)row * a.m1 + clampi(lo1a + c, a.m1 - 1)];
        }
    };
    if constexpr (D3) {   // (the launcher guarantees aligned sources)
        if ((tid & 31) == 0)
            for (int j = 0; j < NBAR && j < NJ; ++j) issue(j);
    } else if (bulk) {
        if (tid == 0)
            for (int j = 0; j < NSTAGE - 1 && j < NJ; ++j) issue(j);
    } else {
        for (int j = 0; j < NSTAGE - 1 && j < NJ; ++j) stage_manual(j);
    }

    // completion counters and dim-2 weights of the chunk
    for (int k = tid; k < rows; k += NT) atomicAdd(&sCnt[a.base2[i2a + k] - lo2 + S2 - 1], 1);
    for (int k = tid; k < rows * S2; k += NT) sW2[k] = a.alpha * a.W2[(int64_t)i2a * S2 + k];

    // per-thread dim-1 taps
    const int colA = i1_0 + 2 * tid;
    const bool haveA = colA < a.n1, haveB = colA + 1 < a.n1;
    const int iA = min(colA, a.n1 - 1), iB = min(colA + 1, a.n1 - 1);
    const int offA = a.base1[iA] - lo1a, offB = a.base1[iB] - lo1a;
    T w1A[S1], w1B[S1];
#pragma unroll
    for (int s = 0; s < S1; ++s) {
        w1A[s] = a.W1[(int64_t)iA * S1 + s];
        w1B[s] = a.W1[(int64_t)iB * S1 + s];
    }
    const bool shift = offB != offA;   // PAIR: offB - offA is 0 or 1
    // D3: the row buffers are de-interleaved -- element e of a staged row sits at (e >> 1) + (e & 1) * P1/2 -- so that
    // lanes whose windows start two elements apart (a unit-rate table: two output columns per thread) read consecutive
    // words.  posX0 / posX1: where the elements offX and offX + 1 sit; tap s reads posX{s & 1}[s >> 1].
    const int H1 = P1 >> 1;
    const int posA0 = (offA >> 1) + (offA & 1) * H1, posA1 = ((offA + 1) >> 1) + ((offA + 1) & 1) * H1;
    const int posB0 = (offB >> 1) + (offB & 1) * H1, posB1 = ((offB + 1) >> 1) + ((offB + 1) & 1) * H1;
    const bool vec = (a.vec_ok & 2) && haveB;
    const T *wp = sW2;                                  // weights of the next output row
    T *dp = dst + (int64_t)i2a * a.n1 + colA;           // next output element
    T hA[S2], hB[S2];
#pragma unroll
    for (int s = 0; s < S2; ++s) hA[s] = hB[s] = (T)0;

    if constexpr (D3) {
        mbar_wait(&bar[0], 0);
        combine(0);
    }

    for (int j = 0; j < NJ; ++j) {
        __syncthreads();   // every thread is done with stage j-1: its buffer may be refilled
        T *buf = sS + (size_t)(j % NSTAGE) * K * P1;
        if constexpr (D3) {
            // raw stage j was combined before this barrier: its slot takes stage j + NBAR; stage j + 1 is combined
            // into the other row buffer (read after the next barrier)
            if ((tid & 31) == 0 && j + NBAR < NJ) issue(j + NBAR);
        } else {
            if (bulk) {
                if (tid == 0 && j + NSTAGE - 1 < NJ) issue(j + NSTAGE - 1);
                mbar_wait(&bar[j % NSTAGE], (uint32_t)((j / NSTAGE) & 1));
            } else if (j + NSTAGE - 1 < NJ) {
                stage_manual(j + NSTAGE - 1);   // (visible after the next barrier; stage j was written two barriers ago)
            }
            if (bulk && patch) {   // replicate the first / last source column into the out-of-range staged columns
                const int warp = tid >> 5, lane = tid & 31;   // warp <-> staged row, lane <-> column to patch
                if (warp < K) {
                    T *row = buf + warp * P1;
                    if (lane < cv0) row[lane] = row[cv0];
                    for (int c = cv1 + lane; c < need1; c += 32) row[c] = row[cv1 - 1];
                }
                __syncthreads();
            }
        }
        if (haveA) {
        const int *cp = sCnt + j * K;
        // pass 1: dim 1, for the K staged rows at once (independent FMA chains)
        T nA[K], nB[K];
#pragma unroll
        for (int u = 0; u < K; ++u) {
            if constexpr (PAIR) {
                T v[S1 + 1];
                if constexpr (D3) {   // de-interleaved row buffer (see combine): conflict-free for unit-rate tables too
                    const T *pe = buf + u * P1 + posA0, *po = buf + u * P1 + posA1;
#pragma unroll
                    for (int s = 0; s <= S1; ++s) v[s] = (s & 1) ? po[s >> 1] : pe[s >> 1];
                } else {
                    const T *p = buf + u * P1 + offA;
#pragma unroll
                    for (int s = 0; s <= S1; ++s) v[s] = p[s];
                }
                T accA = v[0] * w1A[0];
                T accB = (shift ? v[1] : v[0]) * w1B[0];
#pragma unroll
                for (int s = 1; s < S1; ++s) {
                    accA = fmaT(v[s], w1A[s], accA);
                    accB = fmaT(shift ? v[s + 1] : v[s], w1B[s], accB);
                }
                nA[u] = accA, nB[u] = accB;
            } else {
                T accA, accB;
                if constexpr (D3) {
                    const T *pe = buf + u * P1 + posA0, *po = buf + u * P1 + posA1;
                    const T *qe = buf + u * P1 + posB0, *qo = buf + u * P1 + posB1;
                    accA = pe[0] * w1A[0];
                    accB = qe[0] * w1B[0];
#pragma unroll
                    for (int s = 1; s < S1; ++s) {
                        accA = fmaT((s & 1) ? po[s >> 1] : pe[s >> 1], w1A[s], accA);
                        accB = fmaT((s & 1) ? qo[s >> 1] : qe[s >> 1], w1B[s], accB);
                    }
                } else {
                    const T *p = buf + u * P1 + offA, *q = buf + u * P1 + offB;
                    accA = p[0] * w1A[0];
                    accB = q[0] * w1B[0];
#pragma unroll
                    for (int s = 1; s < S1; ++s) {
                        accA = fmaT(p[s], w1A[s], accA);
                        accB = fmaT(q[s], w1B[s], accB);
                    }
                }
                nA[u] = accA, nB[u] = accB;
            }
        }
#pragma unroll
        for (int u = 0; u < K; ++u) {
            hA[u % S2] = nA[u];
            hB[u % S2] = nB[u];
            // pass 2: emit the output rows whose window [row-S2+1, row] is now complete
#pragma unroll 1
            for (int n = cp[u]; n > 0; --n) {
                T w2[S2];
                ld_weights<T, S2>(wp, w2);
                wp += S2;
                T oA = hA[(u + 1) % S2] * w2[0], oB = hB[(u + 1) % S2] * w2[0];
#pragma unroll
                for (int s = 1; s < S2; ++s) {
                    oA = fmaT(hA[(u + 1 + s) % S2], w2[s], oA);
                    oB = fmaT(hB[(u + 1 + s) % S2], w2[s], oB);
                }
                if constexpr (BETA) {
                    oA = fmaT(a.beta, dp[0], oA);
                    if (haveB) oB = fmaT(a.beta, dp[1], oB);
                }
                if (vec) {
                    Pack<T, 2> o;
                    o.v[0] = oA, o.v[1] = oB;
                    st_pack<T, 2>(dp, o);
                } else {
                    dp[0] = oA;
                    if (haveB) dp[1] = oB;
                }
                dp += a.n1;
            }
        }
        }   // haveA
        if constexpr (D3) {
            if (j + 1 < NJ) {
                mbar_wait(&bar[(j + 1) % NBAR], (uint32_t)(((j + 1) / NBAR) & 1));
                combine(j + 1);
            }
        }
    }
}

// (D3: nbatch counts output planes, i.e. n3 times the number of 3-D arrays; a.base3 / W3 / n3 / m3 / nc3 are set)
template <typename T, int S1, int S2, bool D3 = false>
static int launch_fwd(SepArgs<T> &a, const li_table *t1, const li_table *t2, int64_t nbatch, cudaStream_t st,
                      bool *done)
{
    constexpr int TO1 = 2 * FWD_NT, K = FwdK<S2, D3>::value;
    constexpr int NROWBUF = D3 ? 2 : FWD_NSTAGE, NBAR = D3 ? FWD3_NSR : FWD_NSTAGE;
    a.tiles1 = (a.n1 + TO1 - 1) / TO1;
    // chunks of output rows: FWD_WAVES_ whole waves of the CTAs resident per SM (the kernel's launch bound), chunk
    // length a multiple of 8 and at most FWD_MAXCR
    constexpr int PER_SM = (sizeof(T) == 4 && S1 <= 4 && S2 <= 4) ? 4 : (sizeof(T) == 4 ? 3 : 2);
    // (measured, scratch/waves_probe.py: the narrow Float32 kernels -- four CTAs per SM -- on integer resampling rates run
    // fastest in four waves of shorter chunks, 65 -> 59 us on config #2: less idle time at the end of the grid; wide
    // kernels and fractional rates in two)
    int waves = FWD_WAVES_;
    if (!D3 && PER_SM == 4 && a.m2 > 0 && a.n2 % a.m2 == 0) waves = 4;
    if (const int o = (opt(OPT_EXP) >> 8) & 15) waves = o;   // (development: li_set_option("exp", waves << 8))
    const int64_t target = (int64_t)sep2d_num_sms() * PER_SM * waves;
    int chunks = (int)std::max<int64_t>(1, target / (a.tiles1 * nbatch));   // (rounded down: whole waves)
    int CR = (a.n2 + chunks - 1) / chunks;
    CR = std::min(FWD_MAXCR, std::max(8, (CR + 7) & ~7));
    a.tiles2 = (a.n2 + CR - 1) / CR;
    int span1 = 0, span2 = 0;
    int rc = table_fwd_span(t1, TO1, st, &span1);
    if (rc == LI_OK) rc = table_fwd_span(t2, CR, st, &span2);
    if (rc != LI_OK) return rc;
    // staged row: footprint + 1 spare tap (paired load) + up to 3 columns of alignment shift, multiple of 4
    a.F1MAX = (span1 + S1 + 1 + 3 + 3) & ~3;
    const int NCNT = (span2 + S2 + K - 1) / K * K;   // staged source rows of a chunk, padded to whole stages
    // 16-byte bulk copies need aligned source rows; 8-byte stores need aligned destination rows
    // (bit 0: sources -- 16-byte bulk copies need aligned source rows; bit 1: destinations -- 8 / 16-byte stores need
    // aligned destination rows)
    a.vec_ok = (((a.m1 % 4 == 0) && (((uintptr_t)a.src) % 16 == 0)) ? 1 : 0) |
               (((a.n1 % 2 == 0) && (((uintptr_t)a.dst) % (2 * sizeof(T)) == 0)) ? 2 : 0);
    const size_t smem = ((size_t)(NROWBUF + (D3 ? NBAR * a.nc3 : 0)) * K * a.F1MAX + (size_t)CR * S2) * sizeof(T) +
                        (size_t)((NCNT + 1) & ~1) * sizeof(int) + NBAR * sizeof(uint64_t);
    if (smem > (size_t)kSmemLimit) return LI_OK;   // footprint too large for staging: decline
    if (D3 && !(a.vec_ok & 1)) return LI_OK;   // the raw ring is TMA only
    if (opt(OPT_TRACE))
        fprintf(stderr, "[li_b200] fwd%s: P1 %d CR %d K %d smem %zu tiles %d x %d x %lld\n", D3 ? " 3-D" : "", a.F1MAX, CR, K,
                smem, a.tiles1, a.tiles2, (long long)nbatch);
    if (BandCtl *bc = D3 ? nullptr : sep2d_band()) {   // host-buffer pipeline: describe the chunks, or launch a band of them
        if (bc->mode == 1) {
            const std::vector<int32_t> *bh = nullptr;
            rc = table_base_host(t2, st, &bh);
            if (rc != LI_OK) return rc;
            bc->row_begin.assign(a.tiles2 + 1, a.n2);
            bc->src_end.assign(a.tiles2, 0);
            for (int k = 0; k < a.tiles2; ++k) {
                bc->row_begin[k] = k * CR;
                const int last = std::min((k + 1) * CR, a.n2) - 1;
                bc->src_end[k] = std::min(std::max((*bh)[last] + S2 - 1, 0), a.m2 - 1) + 1;
            }
            *done = true;
            return LI_OK;
        }
        a.chunk0 = bc->chunk0;
        a.tiles2 = bc->nchunks;
    }
    const int64_t blocks = (int64_t)a.tiles1 * a.tiles2 * nbatch;
    LI_REQUIRE(blocks <= 0x7fffffffLL, LI_ERR_UNSUPPORTED, "too many tiles");
    const bool pair = t1->max_step <= 1, beta = a.beta != (T)0;
#define LI_FWD(P, B)                                                          \
    if (pair == P && beta == B) {                                             \
        auto kern = sep2d_fwd_kernel<T, S1, S2, P, B, D3>;                    \
        rc = set_smem(kern, smem);                                            \
        if (rc != LI_OK) return rc;                                           \
        kern<<<(unsigned)blocks, FWD_NT, smem, st>>>(a, CR, NCNT);            \
    }
    LI_FWD(true, false) LI_FWD(true, true) LI_FWD(false, false) LI_FWD(false, true)
#undef LI_FWD
    LI_CHECK_LAUNCH();
    *done = true;
    return LI_OK;
}

}  // namespace li
