#pragma once
// li_affine_gather.cuh -- adjoint of the 2-D affine-transform operator in gather form (K10): deterministic, no atomics.
//
// Reference: apply!(alpha, Adjoint, ker1, ker2, R, src, beta, dst), src/nonseparable.jl:204-264 -- a sequential scatter
//     dst[j1,j2] += w1[s1] * (w2[s2] * (alpha*src[i1,i2]))        for every pixel (i1,i2) of src and tap (s1,s2),
// (j1, j2) the clamped tap indices of the pixel's position x = R.xx*i1 + (R.xy*i2 + R.x), y = R.yx*i1 + (R.yy*i2 + R.y).
//
// Two kernels, both owner-computes (every dst element is written once, by one thread, from a sum taken in a fixed order):
//
//  * INTERIOR elements (not on the first / last row or column of dst) only receive unclamped taps, i.e. taps of pixels
//    whose window base (b1, b2) lies in [j1-S1+1, j1] x [j2-S2+1, j2].  A CTA owns a tile of 32 x TD2 dst elements.  It
//    inverse-maps the tile (plus the tap margin) to a box of candidate src pixels, evaluates the coefficients of every
//    candidate ONCE with the forward arithmetic (same un-fused operations as the direct kernel: the same bases and
//    weights bit for bit), files the candidates by window base into cells, and every thread collects, for its element,
//    the pixels of the S1*S2 cells that reach it: a list of (offset in the staged box, w1*w2) that stays in REGISTERS
//    for all images of the batch.  Per image the box of src pixels is staged in shared memory (tensor-map TMA, ring of
//    boxes) and an element costs one shared-memory load and one FMA per contributing pixel.
//
//  * BORDER elements collect every pixel hanging over that edge (Flat / SafeFlat clamping): whole regions of a rotated
//    frame.  Their contributions form a sparse matrix (border element x src pixel) that depends only on the operator:
//    it is built once per operator (every pixel emits its folded border terms, a stable radix sort groups them by
//    element: ascending pixel order = the order of the reference's scatter), cached, and applied per image as
//    fixed-size work items (a warp sums one item with a fixed lane assignment and a fixed shuffle tree; the items of
//    an element are added in order).
//
// Declines (the caller then uses the atomic scatter kernel): kernels with S1*S2 > 16, transforms that put more than 2
// src pixels into one dst cell (minification), candidate boxes that do not fit shared memory, dst smaller than 3 x 3.
#include <cub/cub.cuh>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "li_affine.cuh"

namespace li {

constexpr int AG_TD1 = 32;       // tile width (dst elements along dim 1 = lanes of a warp)
constexpr int AG_KC = 2;         // pixels a cell holds in the per-element register lists (more go to a small overflow list)
constexpr int AG_NOVF = 16;      // overflow entries per tile (cells that received a third pixel: position rounding)
constexpr int AG_NSTG = 4;       // boxes (images) in flight per CTA
constexpr int AG_ITEM = 2048;    // border entries per work item
constexpr int AG_SLOTB = 12288;  // bytes per ring slot (128 bytes of zero pad + the staged box)

struct AGParams {
    double ixx, ixy, iyx, iyy;   // inverse of the linear part of R (positions -> pixel indices)
    int BW, BH;                  // candidate box of a tile (src pixels), upper bounds
    int PB, HB;                  // staged box (TMA): PB >= BW + alignment slack, HB >= BH; PB = row pitch
    int bulk;                    // boxes arrive by tensor-map TMA (else plain loads)
    int ring_off;                // byte offset of the ring of boxes in dynamic shared memory
};

template <typename T, int S1, int S2, int TD2>
__global__ void __launch_bounds__(AG_TD1 *TD2)
affine_adj_gather_kernel(AffineArgs<T> a, AGParams g, const T *__restrict__ src, T *__restrict__ dst, int *__restrict__ err,
                         const __grid_constant__ CUtensorMap tmap)
{
    constexpr int TD1 = AG_TD1, NT = TD1 * TD2, KC = AG_KC, CAP = 2 * S1 * S2;
    constexpr int CW = TD1 + S1 - 1, CH = TD2 + S2 - 1;        // cells (window bases) that reach the tile
    constexpr int AL = 16 / (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // phase 1: [cand weights][cells]   phase 2: [zero pad | boxes ...] over the same bytes
    const int NCAND = (g.BW + AL - 1) * g.BH;                   // (a TMA box starts at an aligned column: up to AL-1 more)
    T *sw1 = reinterpret_cast<T *>(smem_raw);                   // [NCAND + 1][S1] (the last one: the sentinel, zeros)
    T *sw2 = sw1 + (size_t)(NCAND + 1) * S1;                    // [NCAND + 1][S2]
    int *ccnt = reinterpret_cast<int *>(sw2 + (size_t)(NCAND + 1) * S2);   // [CW*CH]
    ushort2 *cpix = reinterpret_cast<ushort2 *>(ccnt + CW * CH);   // [CW*CH][KC] {candidate (index of its weights), offset in the staged box}
    struct Ovf {                                                // a pixel that found its cell full
        int u1, u2, off;
        T w1[S1], w2[S2];
    };
    __shared__ Ovf sovf[AG_NOVF];
    __shared__ int novf;
    __shared__ int sbox[4];                                     // I1a, I2a (first candidate pixel), nb1, nb2
    __shared__ __align__(8) uint64_t bar[AG_NSTG];

    const int tid = threadIdx.x, lx = tid % TD1, ly = tid / TD1;
    const int J1a = blockIdx.x * TD1, J2a = blockIdx.y * TD2;
    const int j1 = J1a + lx, j2 = J2a + ly;
    if (tid == 0) {
        // candidate pixels: every position (x, y) (1-based coordinates of dst) whose window reaches the tile lies in
        // [J1a - S1/2, J1a + TD1 + S1/2 + 1] x [...] (even and odd S; one more on each side for safety); that rectangle
        // is mapped back through the inverse transform
        const double x0 = J1a - 0.5 * S1 - 1.0, x1 = J1a + TD1 + 0.5 * S1 + 2.0, y0 = J2a - 0.5 * S2 - 1.0, y1 = J2a + TD2 + 0.5 * S2 + 2.0;
        double lo1 = 1e300, hi1 = -1e300, lo2 = 1e300, hi2 = -1e300;
        for (int k = 0; k < 4; ++k) {
            const double xx = ((k & 1) ? x1 : x0) - (double)a.x, yy = ((k & 2) ? y1 : y0) - (double)a.y;
            const double p1 = g.ixx * xx + g.ixy * yy, p2 = g.iyx * xx + g.iyy * yy;   // 1-based pixel position
            lo1 = fmin(lo1, p1), hi1 = fmax(hi1, p1), lo2 = fmin(lo2, p2), hi2 = fmax(hi2, p2);
        }
        int I1a = (int)fmax(fmin(floor(lo1) - 2.0, 2e9), -2e9), I2a = (int)fmax(fmin(floor(lo2) - 2.0, 2e9), -2e9);
        int I1b = (int)fmax(fmin(ceil(hi1), 2e9), -2e9), I2b = (int)fmax(fmin(ceil(hi2), 2e9), -2e9);   // (0-based, inclusive)
        I1a = max(I1a, 0), I2a = max(I2a, 0), I1b = min(I1b, a.n1 - 1), I2b = min(I2b, a.n2 - 1);
        if (g.bulk) I1a &= ~(AL - 1);            // the innermost TMA coordinate must be 16-byte aligned
        sbox[0] = I1a, sbox[1] = I2a;
        const int w1max = g.BW + (g.bulk ? AL - 1 : 0);
        sbox[2] = max(0, min(I1b - I1a + 1, w1max)), sbox[3] = max(0, min(I2b - I2a + 1, g.BH));
        // (the box bounds are upper bounds by construction; a box cut by them would lose contributors: flagged)
        if (I1b - I1a + 1 > w1max || I2b - I2a + 1 > g.BH) atomicExch(err, 3);
        novf = 0;
#pragma unroll
        for (int k = 0; k < AG_NSTG; ++k) mbar_init(&bar[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Empty cell entries point at a sentinel candidate (index NCAND: zero weights) and at offset 0 of the ring slot, a
    // zero word in front of the box (box offsets are stored + PADN): the list building below needs no branches.
    constexpr int PADN = 128 / (int)sizeof(T);
    for (int k = tid; k < CW * CH; k += NT) ccnt[k] = 0;
    for (int k = tid; k < CW * CH * KC; k += NT) cpix[k] = make_ushort2((unsigned short)NCAND, 0);
    for (int k = tid; k < S1 + S2; k += NT) (k < S1 ? sw1[NCAND * S1 + k] : sw2[NCAND * S2 + (k - S1)]) = (T)0;
    __syncthreads();
    const int I1a = sbox[0], I2a = sbox[1], nb1 = sbox[2], nb2 = sbox[3];
    const int P = g.bulk ? g.PB : g.BW;          // row pitch of the staged box
    const unsigned inv1 = nb1 > 1 ? (unsigned)((1ull << 32) / (unsigned)nb1 + 1ull) : 0u;   // p / nb1 = (p * inv1) >> 32 while p * nb1 < 2^32

    // The ring of src boxes (phase 2) has its own shared memory, so the first boxes are requested NOW and arrive while
    // phase 1 runs.  Ring slots have a compile-time stride and the image loop below is unrolled over the slots: a
    // contributor costs one LDS with an immediate offset and one FMA.
    constexpr int SLOT = AG_SLOTB / (int)sizeof(T);   // elements per ring slot (PADN of them: the zero pad)
    T *ring = reinterpret_cast<T *>(smem_raw + g.ring_off);
    const int NS = g.bulk ? AG_NSTG : 1;
    for (int k = tid; k < NS * PADN; k += NT) ring[(size_t)(k / PADN) * SLOT + k % PADN] = (T)0;
    const int64_t splane = (int64_t)a.n1 * a.n2, dplane = (int64_t)a.m1 * a.m2;
    const bool mine = j1 >= 1 && j1 < a.m1 - 1 && j2 >= 1 && j2 < a.m2 - 1;   // interior element of dst
    const bool hb = a.beta != (T)0;
    auto stage = [&](int64_t b, int k) {
        T *buf = ring + (size_t)k * SLOT + PADN;
        if (g.bulk) {
            if (tid == 0) {
                mbar_expect_tx(&bar[k], (uint32_t)(g.PB * g.HB * sizeof(T)));
                asm volatile(
                    "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
                        "r"(smem_u32(buf)), "l"(&tmap), "r"(I1a), "r"(I2a), "r"((int)b), "r"(smem_u32(&bar[k]))
                    : "memory");
            }
        } else {
            const T *sp = src + b * splane + (int64_t)I2a * a.n1 + I1a;
            for (int r = tid / 32; r < nb2; r += NT / 32)
                for (int c = tid % 32; c < nb1; c += 32) buf[r * P + c] = sp[(int64_t)r * a.n1 + c];
        }
    };
    const int64_t bstep = gridDim.z;
    const int64_t b0 = blockIdx.z;
    const int nimg = b0 < a.nbatch ? (int)((a.nbatch - b0 + bstep - 1) / bstep) : 0;
    const bool any = nb1 > 0 && nb2 > 0;
    if (any && g.bulk)
        for (int k = 0; k < NS - 1 && k < nimg; ++k) stage(b0 + k * bstep, k);

    // ---- phase 1a: coefficients of the candidate pixels, filed by window base
    for (int p = tid; p < nb1 * nb2; p += NT) {
        const int pr = nb1 > 1 ? (int)(((unsigned long long)(unsigned)p * inv1) >> 32) : p, pc = p - pr * nb1;   // row, column in the box
        const int i1 = I1a + pc, i2 = I2a + pr;
        T w1[S1], w2[S2];
        int b1, b2;
        {   // the pixel's position, with the operations of affine_coefs; most of the bounding box lies outside the
            // (rotated) rectangle of positions that reach the tile: those candidates stop here
            using A = Ar<T>;
            const T pos1 = (T)(i1 + 1), pos2 = (T)(i2 + 1);
            const T px = A::add(A::mul(a.xx, pos1), A::add(A::mul(a.xy, pos2), a.x));
            const T py = A::add(A::mul(a.yx, pos1), A::add(A::mul(a.yy, pos2), a.y));
            if (px < (T)(J1a - 0.5 * S1 - 1.0) || px > (T)(J1a + TD1 + 0.5 * S1 + 2.0) || py < (T)(J2a - 0.5 * S2 - 1.0) ||
                py > (T)(J2a + TD2 + 0.5 * S2 + 2.0))
                continue;   // (NaN positions fall through to affine_coefs, which flags them)
        }
        if (!affine_coefs<T, S1, S2>(a, i1, i2, b1, b2, w1, w2)) {
            atomicExch(err, 1);
            continue;
        }
#pragma unroll
        for (int s = 0; s < S1; ++s) sw1[p * S1 + s] = w1[s];
#pragma unroll
        for (int s = 0; s < S2; ++s) sw2[p * S2 + s] = w2[s];
        const int u1 = b1 - (J1a - (S1 - 1)), u2 = b2 - (J2a - (S2 - 1));
        // (canonical bases -(S-1) and len-1 stand for every window hanging completely over an edge: those pixels only
        // feed border elements, and any number of them shares that base)
        const bool sat = b1 <= -(S1 - 1) || b1 >= a.m1 - 1 || b2 <= -(S2 - 1) || b2 >= a.m2 - 1;
        if (!sat && u1 >= 0 && u1 < CW && u2 >= 0 && u2 < CH) {
            const int k = atomicAdd(&ccnt[u2 * CW + u1], 1);
            if (k < KC) cpix[(u2 * CW + u1) * KC + k] = make_ushort2((unsigned short)p, (unsigned short)(pr * P + pc + PADN));
        }
    }
    __syncthreads();
    // Cells in ascending pixel order (the atomics above filled them in any order).  A cell that received more than KC
    // pixels keeps its KC smallest; the others go to the tile's overflow list, also in ascending order: one thread
    // re-scans the candidates of such a cell (rare: it takes a position rounded across a cell boundary).
    for (int c = tid; c < CW * CH; c += NT) {
        const int n = ccnt[c];
        ushort2 *e = cpix + c * KC;
        if (n <= KC) {
            if (n == 2 && e[0].x > e[1].x) {
                const ushort2 t = e[0];
                e[0] = e[1], e[1] = t;
            }
            continue;
        }
        const int u1 = c % CW, u2 = c / CW;
        int kept = 0;
        for (int p = 0; p < nb1 * nb2; ++p) {   // ascending pixel order
            T w1[S1], w2[S2];
            int b1, b2;
            if (!affine_coefs<T, S1, S2>(a, I1a + p % nb1, I2a + p / nb1, b1, b2, w1, w2)) continue;
            if (b1 - (J1a - (S1 - 1)) != u1 || b2 - (J2a - (S2 - 1)) != u2) continue;
            if (kept < KC) {
                e[kept++] = make_ushort2((unsigned short)p, (unsigned short)((p / nb1) * P + p % nb1 + PADN));
            } else {
                const int o = atomicAdd(&novf, 1);
                if (o < AG_NOVF) {
                    sovf[o].u1 = u1, sovf[o].u2 = u2, sovf[o].off = (p / nb1) * P + p % nb1 + PADN;
#pragma unroll
                    for (int s = 0; s < S1; ++s) sovf[o].w1[s] = w1[s];
#pragma unroll
                    for (int s = 0; s < S2; ++s) sovf[o].w2[s] = w2[s];
                } else {
                    atomicExch(err, 4);
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 1b: this element's contributors: the pixels of cell (j1 - s1, j2 - s2) reach it with tap (s1, s2).
    // Register slot k*S1*S2 + s2*S1 + s1 = k-th pixel of that cell (first pixels first: the second half is mostly empty);
    // empty slots carry weight 0 and the offset of the zero word.  Offsets are relative to the ring slot.
    T w[CAP];
    int off[CAP];
    unsigned second = 0;                            // bit q: slots S1*S2 + 4q .. + 4q+3 hold something
    {
        const ushort2 *cbase = cpix + ((ly + (S2 - 1)) * CW + (lx + (S1 - 1))) * KC;
#pragma unroll
        for (int s2 = 0; s2 < S2; ++s2)
#pragma unroll
            for (int s1 = 0; s1 < S1; ++s1) {
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const int slot = k * S1 * S2 + s2 * S1 + s1;
                    const ushort2 e = cbase[-(s2 * CW + s1) * KC + k];
                    w[slot] = sw1[e.x * S1 + s1] * sw2[e.x * S2 + s2];
                    off[slot] = e.y;
                    if (k > 0) second |= (e.y != 0 ? 1u : 0u) << ((slot - S1 * S2) / 4);
                }
            }
    }
    const int nov = min(novf, AG_NOVF);

    T *dp = dst + b0 * dplane + (int64_t)j2 * a.m1 + j1;
    const T *pk[CAP];                              // the contributors in ring slot 0
#pragma unroll
    for (int k = 0; k < CAP; ++k) pk[k] = ring + off[k];
    auto one = [&](int it, auto slot_tag, uint32_t par) {   // image `it` from ring slot KS (static)
        constexpr int KS = decltype(slot_tag)::value;
        if (any) {
            if (it + NS - 1 < nimg) stage(b0 + (int64_t)(it + NS - 1) * bstep, (KS + AG_NSTG - 1) % AG_NSTG % (g.bulk ? AG_NSTG : 1));
            if (g.bulk)
                mbar_wait(&bar[KS], par);
            else
                __syncthreads();
        }
        T acc = (T)0;
#pragma unroll
        for (int k = 0; k < S1 * S2; ++k) acc = fmaT(w[k], pk[k][KS * SLOT], acc);
#pragma unroll
        for (int k0 = S1 * S2; k0 < CAP; k0 += 4) {
            if (second >> ((k0 - S1 * S2) / 4) & 1u) {
#pragma unroll
                for (int k = k0; k < k0 + 4 && k < CAP; ++k) acc = fmaT(w[k], pk[k][KS * SLOT], acc);
            }
        }
        for (int o = 0; o < nov; ++o) {   // (pixels that found their cell full)
            const int s1 = lx + (S1 - 1) - sovf[o].u1, s2 = ly + (S2 - 1) - sovf[o].u2;
            if (s1 >= 0 && s1 < S1 && s2 >= 0 && s2 < S2)
                acc = fmaT(sovf[o].w1[s1] * sovf[o].w2[s2], ring[(size_t)KS * SLOT + sovf[o].off], acc);
        }
        if (mine) {
            const T v = a.alpha == (T)1 ? acc : a.alpha * acc;
            *dp = hb ? fmaT(a.beta, *dp, v) : v;     // beta = 0: dst is never read
        }
        dp += dplane * bstep;
        __syncthreads();   // every thread is done with this box before it is refilled
    };
    __syncthreads();   // (the zero pads of the ring)
    if (g.bulk) {
        uint32_t par = 0;
        for (int it0 = 0; it0 < nimg; it0 += AG_NSTG, par ^= 1u) {
            if (it0 + 0 < nimg) one(it0 + 0, std::integral_constant<int, 0>{}, par);
            if (it0 + 1 < nimg) one(it0 + 1, std::integral_constant<int, 1>{}, par);
            if (it0 + 2 < nimg) one(it0 + 2, std::integral_constant<int, 2>{}, par);
            if (it0 + 3 < nimg) one(it0 + 3, std::integral_constant<int, 3>{}, par);
        }
    } else {
        for (int it = 0; it < nimg; ++it) one(it, std::integral_constant<int, 0>{}, 0u);
    }
}

// ---------------------------------------------------------------------------------------------
// Border elements: the sparse matrix (border element x src pixel), built once per operator.
//
// border id of dst element (j1, j2): first row j2 = 0: j1; last row: m1 + j1; first column (1 <= j2 <= m2-2):
// 2*m1 + j2 - 1; last column: 2*m1 + (m2 - 2) + j2 - 1.
__device__ __forceinline__ int border_id(int j1, int j2, int m1, int m2)
{
    if (j2 == 0) return j1;
    if (j2 == m2 - 1) return m1 + j1;
    if (j1 == 0) return 2 * m1 + j2 - 1;
    if (j1 == m1 - 1) return 2 * m1 + (m2 - 2) + j2 - 1;
    return -1;
}

// pass 0: number of border terms of every pixel; pass 1: the terms (key = border id, pixel, folded weight) at ofs[pixel]
template <typename T, int S1, int S2, int PASS>
__global__ void __launch_bounds__(256)
affine_border_terms_kernel(AffineArgs<T> a, int *__restrict__ counts, const int64_t *__restrict__ ofs, int *__restrict__ keys,
                           int *__restrict__ pix, T *__restrict__ wts, int *__restrict__ err)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (int64_t)a.n1 * a.n2) return;
    const int i1 = (int)(p % a.n1), i2 = (int)(p / a.n1);
    T w1[S1], w2[S2];
    int b1, b2;
    if (!affine_coefs<T, S1, S2>(a, i1, i2, b1, b2, w1, w2)) {
        atomicExch(err, 1);
        if (PASS == 0) counts[p] = 0;
        return;
    }
    const int t1 = clampi(b1, a.m1 - 1), n1d = clampi(b1 + S1 - 1, a.m1 - 1) - t1 + 1;
    const int t2 = clampi(b2, a.m2 - 1), n2d = clampi(b2 + S2 - 1, a.m2 - 1) - t2 + 1;
    // (interior pixels: no target on the border)
    if (t1 > 0 && t1 + n1d - 1 < a.m1 - 1 && t2 > 0 && t2 + n2d - 1 < a.m2 - 1) {
        if (PASS == 0) counts[p] = 0;
        return;
    }
    int n = 0;
    int64_t o = PASS ? ofs[p] : 0;
    for (int k2 = 0; k2 < n2d; ++k2) {
        T f2 = (T)0;
        if (PASS) {
#pragma unroll
            for (int s = 0; s < S2; ++s)
                if (clampi(b2 + s, a.m2 - 1) - t2 == k2) f2 += w2[s];   // taps clamped onto one row, summed first
        }
        for (int k1 = 0; k1 < n1d; ++k1) {
            const int id = border_id(t1 + k1, t2 + k2, a.m1, a.m2);
            if (id < 0) continue;
            if (PASS) {
                T f1 = (T)0;
#pragma unroll
                for (int s = 0; s < S1; ++s)
                    if (clampi(b1 + s, a.m1 - 1) - t1 == k1) f1 += w1[s];
                keys[o + n] = id;
                pix[o + n] = (int)p;
                wts[o + n] = f1 * f2;
            }
            ++n;
        }
    }
    if (PASS == 0) counts[p] = n;
}

template <typename T>
__global__ void affine_border_permute_kernel(const int *__restrict__ order, const int *__restrict__ pix, const T *__restrict__ wts,
                                             int64_t n, int *__restrict__ spix, T *__restrict__ swts)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) {
        spix[e] = pix[order[e]];
        swts[e] = wts[order[e]];
    }
}

static __global__ void affine_border_rowptr_kernel(const int *__restrict__ sorted_keys, int64_t n, int nb, int64_t *__restrict__ rowptr)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nb) return;
    int64_t lo = 0, hi = n;   // first entry with key >= k
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] >= k)
            hi = mid;
        else
            lo = mid + 1;
    }
    rowptr[k] = lo;
}

// one warp per work item (a run of at most AG_ITEM entries of one border element): lane l sums entries l, l+32, ..
// in order, then a fixed shuffle tree; partial[image][item]
template <typename T>
__global__ void __launch_bounds__(256)
affine_border_items_kernel(const int64_t *__restrict__ item_lo, const int *__restrict__ item_n, int nitems, const int *__restrict__ spix,
                           const T *__restrict__ swts, const T *__restrict__ src, int64_t splane, T *__restrict__ partial)
{
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (item >= nitems) return;
    const int64_t lo = item_lo[item];
    const int n = item_n[item];
    const T *sp = src + blockIdx.y * splane;
    T s = (T)0;
    for (int k = lane; k < n; k += 32) s = fmaT(swts[lo + k], sp[spix[lo + k]], s);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) partial[(int64_t)blockIdx.y * nitems + item] = s;
}

// dst[border element] = beta * dst + alpha * (sum of its items, in order)
template <typename T>
__global__ void affine_border_final_kernel(const int *__restrict__ first_item, int nb, int nitems, const T *__restrict__ partial, int m1,
                                           int m2, T alpha, T beta, T *__restrict__ dst)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nb) return;
    int j1, j2;
    if (id < m1)
        j1 = id, j2 = 0;
    else if (id < 2 * m1)
        j1 = id - m1, j2 = m2 - 1;
    else if (id < 2 * m1 + (m2 - 2))
        j1 = 0, j2 = id - 2 * m1 + 1;
    else
        j1 = m1 - 1, j2 = id - 2 * m1 - (m2 - 2) + 1;
    const T *pp = partial + (int64_t)blockIdx.y * nitems;
    T s = (T)0;
    for (int k = first_item[id]; k < first_item[id + 1]; ++k) s += pp[k];
    T *d = dst + (int64_t)blockIdx.y * m1 * m2 + (int64_t)j2 * m1 + j1;
    const T v = alpha == (T)1 ? s : alpha * s;
    *d = beta != (T)0 ? fmaT(beta, *d, v) : v;
}

struct AffineBorderPlan {
    int *spix = nullptr;
    void *swts = nullptr;
    int64_t *item_lo = nullptr;
    int *item_n = nullptr, *first_item = nullptr;
    int nitems = 0, nb = 0;
    int64_t nterms = 0;
    uint64_t stamp = 0;
    void release()
    {
        cudaFree(spix), cudaFree(swts), cudaFree(item_lo), cudaFree(item_n), cudaFree(first_item);
        spix = nullptr, swts = nullptr, item_lo = nullptr, item_n = nullptr, first_item = nullptr;
    }
};
struct AffineBorderKey {
    int dev, dtype, f1, f2, S1, S2, b1, b2;
    int64_t n1, n2, m1, m2;
    double p[8], R[6];
    bool operator<(const AffineBorderKey &o) const { return memcmp(this, &o, sizeof *this) < 0; }
};
std::mutex &affine_plan_mutex();
std::map<AffineBorderKey, AffineBorderPlan> &affine_plans();

static __global__ void affine_iota_kernel(int *__restrict__ v, int64_t n)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) v[e] = (int)e;
}

template <typename T, int S1, int S2>
static int affine_border_build(const AffineArgs<T> &a, int *err, cudaStream_t st, AffineBorderPlan *plan)
{
    const int64_t npix = (int64_t)a.n1 * a.n2;
    const int nb = 2 * a.m1 + 2 * (a.m2 - 2);
    plan->nb = nb;
    int *counts = nullptr, *keys = nullptr, *pix = nullptr, *order = nullptr, *keys_s = nullptr, *order_s = nullptr;
    int64_t *ofs = nullptr, *rowptr = nullptr;
    T *wts = nullptr;
    void *tmp = nullptr;
    auto cleanup = [&]() {
        cudaFree(counts), cudaFree(keys), cudaFree(pix), cudaFree(order), cudaFree(keys_s), cudaFree(order_s), cudaFree(ofs), cudaFree(rowptr),
            cudaFree(wts), cudaFree(tmp);
    };
#define LI_TRYB(expr)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            set_error("%s failed: %s", #expr, cudaGetErrorString(e__));                      \
            cleanup();                                                                       \
            plan->release();                                                                 \
            return e__ == cudaErrorMemoryAllocation ? LI_ERR_NOMEM : LI_ERR_CUDA;            \
        }                                                                                    \
    } while (0)
    const unsigned blocks = (unsigned)((npix + 255) / 256);
    LI_TRYB(cudaMalloc(&counts, (size_t)npix * sizeof(int)));
    LI_TRYB(cudaMalloc(&ofs, (size_t)(npix + 1) * sizeof(int64_t)));
    affine_border_terms_kernel<T, S1, S2, 0><<<blocks, 256, 0, st>>>(a, counts, nullptr, nullptr, nullptr, nullptr, err);
    count_launch();
    size_t tb = 0;
    LI_TRYB(cudaMemsetAsync(ofs, 0, sizeof(int64_t), st));
    {   // inclusive sums of the counts (64-bit): ofs[p+1] = terms of pixels 0..p
        auto in = cub::TransformInputIterator<int64_t, cub::CastOp<int64_t>, const int *>(counts, cub::CastOp<int64_t>());
        LI_TRYB(cub::DeviceScan::InclusiveSum(nullptr, tb, in, ofs + 1, npix, st));
        LI_TRYB(cudaMalloc(&tmp, tb));
        LI_TRYB(cub::DeviceScan::InclusiveSum(tmp, tb, in, ofs + 1, npix, st));
        count_launch();
    }
    int64_t nterms = 0;
    LI_TRYB(cudaMemcpyAsync(&nterms, ofs + npix, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    LI_TRYB(cudaStreamSynchronize(st));
    cudaFree(tmp), tmp = nullptr;
    cudaFree(counts), counts = nullptr;
    plan->nterms = nterms;
    if (nterms >= (1LL << 31)) {   // (32-bit entry indices in the sort: the caller falls back to the atomic kernel)
        cleanup();
        plan->nterms = -1;
        return LI_OK;
    }
    const size_t ne = (size_t)std::max<int64_t>(nterms, 1);
    LI_TRYB(cudaMalloc(&keys, ne * sizeof(int)));
    LI_TRYB(cudaMalloc(&pix, ne * sizeof(int)));
    LI_TRYB(cudaMalloc(&wts, ne * sizeof(T)));
    affine_border_terms_kernel<T, S1, S2, 1><<<blocks, 256, 0, st>>>(a, nullptr, ofs, keys, pix, wts, err);
    count_launch();
    LI_TRYB(cudaMalloc(&order, ne * sizeof(int)));
    LI_TRYB(cudaMalloc(&keys_s, ne * sizeof(int)));
    LI_TRYB(cudaMalloc(&order_s, ne * sizeof(int)));
    affine_iota_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(order, nterms);
    count_launch();
    {   // STABLE sort by border id: the entries of an element stay in ascending pixel order (the reference's order)
        int end_bit = 1;
        while (end_bit < 31 && (1LL << end_bit) < nb) ++end_bit;
        LI_TRYB(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys_s, order, order_s, (int)nterms, 0, end_bit, st));
        LI_TRYB(cudaMalloc(&tmp, std::max<size_t>(tb, 16)));
        LI_TRYB(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys_s, order, order_s, (int)nterms, 0, end_bit, st));
        count_launch(4);
    }
    LI_TRYB(cudaMalloc(&plan->spix, ne * sizeof(int)));
    LI_TRYB(cudaMalloc(&plan->swts, ne * sizeof(T)));
    affine_border_permute_kernel<T><<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(order_s, pix, wts, nterms, plan->spix, (T *)plan->swts);
    count_launch();
    LI_TRYB(cudaMalloc(&rowptr, (size_t)(nb + 1) * sizeof(int64_t)));
    affine_border_rowptr_kernel<<<(nb + 1 + 255) / 256, 256, 0, st>>>(keys_s, nterms, nb, rowptr);
    count_launch();
    std::vector<int64_t> h_row((size_t)nb + 1);
    LI_TRYB(cudaMemcpyAsync(h_row.data(), rowptr, (size_t)(nb + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    LI_TRYB(cudaStreamSynchronize(st));
    // work items: runs of at most AG_ITEM entries of one element
    std::vector<int64_t> item_lo;
    std::vector<int> item_n, first_item((size_t)nb + 1);
    for (int k = 0; k < nb; ++k) {
        first_item[(size_t)k] = (int)item_lo.size();
        for (int64_t e = h_row[(size_t)k]; e < h_row[(size_t)k + 1]; e += AG_ITEM) {
            item_lo.push_back(e);
            item_n.push_back((int)std::min<int64_t>(AG_ITEM, h_row[(size_t)k + 1] - e));
        }
    }
    first_item[(size_t)nb] = (int)item_lo.size();
    plan->nitems = (int)item_lo.size();
    const size_t ni = std::max<size_t>(item_lo.size(), 1);
    LI_TRYB(cudaMalloc(&plan->item_lo, ni * sizeof(int64_t)));
    LI_TRYB(cudaMalloc(&plan->item_n, ni * sizeof(int)));
    LI_TRYB(cudaMalloc(&plan->first_item, (size_t)(nb + 1) * sizeof(int)));
    LI_TRYB(cudaMemcpyAsync(plan->item_lo, item_lo.data(), item_lo.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    LI_TRYB(cudaMemcpyAsync(plan->item_n, item_n.data(), item_n.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    LI_TRYB(cudaMemcpyAsync(plan->first_item, first_item.data(), (size_t)(nb + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    LI_TRYB(cudaStreamSynchronize(st));
#undef LI_TRYB
    cleanup();
    return LI_OK;
}

// most points of the lattice {L * (i1, i2)} that one half-open unit cell can hold (L = linear part of R); exact for
// the small vectors that matter: a set of lattice points shares a cell iff its extent is < 1 in x and in y
static int affine_cell_capacity(double xx, double xy, double yx, double yy)
{
    std::vector<std::pair<double, double>> d;   // differences with both components in (-1, 1), one of each +- pair
    for (int u = -6; u <= 6; ++u)
        for (int v = 0; v <= 6; ++v) {
            if (v == 0 && u <= 0) continue;
            const double dx = xx * u + xy * v, dy = yx * u + yy * v;
            if (std::fabs(dx) < 1.0 && std::fabs(dy) < 1.0) d.emplace_back(dx, dy);
        }
    if (d.size() > 12) return 99;
    int best = 1;
    const int n = (int)d.size();
    for (int mask = 1; mask < (1 << n); ++mask) {   // points {0} + subset: extent < 1 in both coordinates?
        double lox = 0, hix = 0, loy = 0, hiy = 0;
        int cntp = 1;
        // (differences are taken from one base point, with either sign: try the subset as given and negated members
        // via the +- closure below)
        for (int k = 0; k < n; ++k)
            if (mask >> k & 1) {
                lox = std::min(lox, d[k].first), hix = std::max(hix, d[k].first);
                loy = std::min(loy, d[k].second), hiy = std::max(hiy, d[k].second);
                ++cntp;
            }
        if (hix - lox < 1.0 && hiy - loy < 1.0) best = std::max(best, cntp);
    }
    return best;
}

// wavefronts of a warp-wide load whose lane l reads column floor(l*xx + u), row floor(l*yx + v) of a box of pitch P
// (averaged over a few offsets): the cost model of best_pitch_mod (li_affine.cuh) for one pitch
static double pitch_cost(double xx, double yx, int P)
{
    double cost = 0;
    for (int trial = 0; trial < 8; ++trial) {
        const double u = 0.137 * trial + 0.05, v = 0.291 * trial + 0.11;
        int addr[32];
        for (int l = 0; l < 32; ++l) addr[l] = (int)std::floor(l * xx + u + 64) + (int)std::floor(l * yx + v + 64) * P;
        int worst = 0;
        for (int bank = 0; bank < 32; ++bank) {
            int seen[32], ns = 0;
            for (int l = 0; l < 32; ++l) {
                if (((addr[l] % 32) + 32) % 32 != bank) continue;
                bool dup = false;
                for (int k = 0; k < ns; ++k) dup |= seen[k] == addr[l];
                if (!dup) seen[ns++] = addr[l];
            }
            worst = std::max(worst, ns);
        }
        cost += worst;
    }
    return cost;
}

template <typename T, int S1, int S2, int TD2>
static int launch_affine_gather_interior(const AffineArgs<T> &a, const T *src, T *dst, int *err, cudaStream_t st, bool *done)
{
    constexpr int AL = 16 / (int)sizeof(T), TD1 = AG_TD1, NT = TD1 * TD2, KC = AG_KC;
    constexpr int CW = TD1 + S1 - 1, CH = TD2 + S2 - 1;
    const double det = (double)a.xx * a.yy - (double)a.xy * a.yx;
    if (!(std::fabs(det) > 1e-12) || !std::isfinite(det)) return LI_OK;
    AGParams g;
    g.ixx = (double)a.yy / det, g.ixy = -(double)a.xy / det, g.iyx = -(double)a.yx / det, g.iyy = (double)a.xx / det;
    // candidate box of a tile: extent of the (TD1 + 2*S1 + 3) x (TD2 + 2*S2 + 3) dst rectangle under the inverse map
    const double ex = TD1 + S1 + 3.0, ey = TD2 + S2 + 3.0;
    const double w1 = std::fabs(g.ixx) * ex + std::fabs(g.ixy) * ey, w2 = std::fabs(g.iyx) * ex + std::fabs(g.iyy) * ey;
    if (!(w1 < 400 && w2 < 400)) return LI_OK;
    g.BW = (int)std::ceil(w1) + 6;
    g.BH = (int)std::ceil(w2) + 6;
    if ((int64_t)(g.BW + AL) * g.BH > 60000) return LI_OK;     // (candidate ids are 16-bit)
    g.bulk = (a.n1 % AL == 0) && (((uintptr_t)src) % 16 == 0) && encode_tiled() != nullptr && a.nbatch < (1LL << 31);
    // staged box: pitch with the fewest bank conflicts for lanes walking along the inverse image of dim 1
    const int need1 = g.BW + (g.bulk ? AL - 1 : 0);
    const int slot_elems = AG_SLOTB / (int)sizeof(T) - 128 / (int)sizeof(T);
    g.HB = g.BH;
    {   // the narrowest pitch whose bank conflicts stay within 25 % of the best one's (boxes are re-read per image: bytes first)
        const int step = g.bulk ? AL : 1, first = (need1 + step - 1) / step * step;
        double best = 1e30;
        for (int pb = first; pb < first + 32; pb += step) best = std::min(best, pitch_cost(g.ixx, g.iyx, pb));
        g.PB = first;
        for (int pb = first; pb < first + 32; pb += step)
            if (pitch_cost(g.ixx, g.iyx, pb) <= 1.25 * best) {
                g.PB = pb;
                break;
            }
        if (g.PB * g.HB > slot_elems) g.PB = first;
    }
    if (g.bulk && (g.PB > 256 || g.HB > 256)) g.bulk = 0;
    if (!g.bulk) g.PB = g.BW, g.HB = g.BH;
    if ((int64_t)g.PB * g.HB > slot_elems || (int64_t)g.PB * g.HB + 64 > 65535) return LI_OK;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    if (g.bulk) {
        const cuuint64_t gdim[3] = {(cuuint64_t)a.n1, (cuuint64_t)a.n2, (cuuint64_t)a.nbatch};
        const cuuint64_t gstr[2] = {(cuuint64_t)a.n1 * sizeof(T), (cuuint64_t)a.n1 * a.n2 * sizeof(T)};
        const cuuint32_t box[3] = {(cuuint32_t)g.PB, (cuuint32_t)g.HB, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode_tiled()(&tmap, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
                                          (void *)src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            g.bulk = 0;
            g.PB = g.BW, g.HB = g.BH;
        }
    }
    const int P = g.PB;
    const size_t phase1 = (((size_t)(g.BW + AL - 1) * g.BH + 1) * (S1 + S2) * sizeof(T) + (size_t)CW * CH * (4 + 4 * KC) + 15) & ~(size_t)15;
    g.ring_off = (int)((phase1 + 127) & ~(size_t)127);
    const size_t smem = (size_t)g.ring_off + (size_t)AG_SLOTB * (g.bulk ? AG_NSTG : 1) + 128;
    if (smem > (size_t)kSmemLimit) return LI_OK;
    dim3 grid((a.m1 + TD1 - 1) / TD1, (a.m2 + TD2 - 1) / TD2, 1);
    if (grid.y > 65535) return LI_OK;
    const int64_t tiles = (int64_t)grid.x * grid.y;
    grid.z = (unsigned)std::max<int64_t>(1, std::min<int64_t>(a.nbatch, (148 * 4 + tiles - 1) / tiles));
    auto kern = affine_adj_gather_kernel<T, S1, S2, TD2>;
    LI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
    if (opt(OPT_TRACE))
        fprintf(stderr, "[li_b200] affine gather adjoint: tiles %u x %u x %u, candidate box %d x %d (pitch %d), smem %zu, tma %d\n", grid.x,
                grid.y, grid.z, g.BW, g.BH, P, smem, g.bulk);
    kern<<<grid, NT, smem, st>>>(a, g, src, dst, err, tmap);
    LI_CHECK_LAUNCH();
    *done = true;
    return LI_OK;
}

// dst <- beta * dst + alpha * R' src, deterministic; *done stays false when the gather form declines the operator
template <typename T, int S1, int S2>
static int affine_gather_T(const AffineArgs<T> &a, const li_kernel_desc *ker1, const li_kernel_desc *ker2, const double R[6], const T *src,
                           T *dst, int *err, cudaStream_t st, bool *done)
{
    *done = false;
    if (S1 * S2 > 16 || a.m1 < 3 || a.m2 < 3) return LI_OK;
    if (affine_cell_capacity((double)a.xx, (double)a.xy, (double)a.yx, (double)a.yy) > 2) return LI_OK;
    // ---- the border plan (cached per operator)
    AffineBorderKey key;
    memset(&key, 0, sizeof key);
    LI_CUDA(cudaGetDevice(&key.dev));
    key.dtype = ker1->dtype, key.f1 = ker1->family, key.f2 = ker2->family, key.S1 = S1, key.S2 = S2, key.b1 = ker1->boundary,
    key.b2 = ker2->boundary;
    key.n1 = a.n1, key.n2 = a.n2, key.m1 = a.m1, key.m2 = a.m2;
    for (int k = 0; k < 4; ++k) key.p[k] = ker1->params[k], key.p[4 + k] = ker2->params[k];
    for (int k = 0; k < 6; ++k) key.R[k] = R[k];
    AffineBorderPlan plan;
    {
        std::lock_guard<std::mutex> lock(affine_plan_mutex());
        auto &plans = affine_plans();
        static uint64_t clock = 0;
        auto it = plans.find(key);
        if (it == plans.end()) {
            if (plans.size() >= 4) {   // drop the least recently used plan
                auto old = plans.begin();
                for (auto jt = plans.begin(); jt != plans.end(); ++jt)
                    if (jt->second.stamp < old->second.stamp) old = jt;
                cudaStreamSynchronize(st);
                old->second.release();
                plans.erase(old);
            }
            AffineBorderPlan p;
            int rc = affine_border_build<T, S1, S2>(a, err, st, &p);
            if (rc != LI_OK) return rc;
            it = plans.emplace(key, p).first;
        }
        it->second.stamp = ++clock;
        plan = it->second;
    }
    if (plan.nterms < 0) return LI_OK;
    // ---- interior elements
    bool ok = false;
    int rc = launch_affine_gather_interior<T, S1, S2, (sizeof(T) == 4 ? 8 : 4)>(a, src, dst, err, st, &ok);
    if (rc != LI_OK || !ok) return rc;
    // ---- border elements
    void *partial = nullptr;
    rc = staging(7, (size_t)std::max(plan.nitems, 1) * a.nbatch * sizeof(T), &partial, (void *)st);
    if (rc != LI_OK) return rc;
    if (plan.nitems > 0) {
        dim3 grid((unsigned)((plan.nitems * 32 + 255) / 256), (unsigned)a.nbatch);
        affine_border_items_kernel<T><<<grid, 256, 0, st>>>(plan.item_lo, plan.item_n, plan.nitems, plan.spix, (const T *)plan.swts, src,
                                                            (int64_t)a.n1 * a.n2, (T *)partial);
        LI_CHECK_LAUNCH();
    }
    dim3 gridf((unsigned)((plan.nb + 255) / 256), (unsigned)a.nbatch);
    affine_border_final_kernel<T><<<gridf, 256, 0, st>>>(plan.first_item, plan.nb, plan.nitems, (const T *)partial, a.m1, a.m2, a.alpha, a.beta, dst);
    LI_CHECK_LAUNCH();
    *done = true;
    return LI_OK;
}

template <typename T>
int affine_adjoint_gather(const li_kernel_desc *ker1, const li_kernel_desc *ker2, const double R[6], int64_t n1, int64_t n2, int64_t m1,
                          int64_t m2, int64_t nbatch, double alpha, const void *src, double beta, void *dst, int *err, cudaStream_t st,
                          bool *done)
{
    *done = false;
    if (nbatch > 65535) return LI_OK;
    AffineArgs<T> a;
    int rc = make_kerdev<T>(ker1, m1, &a.k1);
    if (rc == LI_OK) rc = make_kerdev<T>(ker2, m2, &a.k2);
    if (rc != LI_OK) return rc;
    a.xx = (T)R[0], a.xy = (T)R[1], a.x = (T)R[2], a.yx = (T)R[3], a.yy = (T)R[4], a.y = (T)R[5];
    a.n1 = (int)n1, a.n2 = (int)n2, a.m1 = (int)m1, a.m2 = (int)m2;
    a.nbatch = nbatch;
    a.alpha = (T)alpha, a.beta = (T)beta;
    const int S1 = ker1->S, S2 = ker2->S;
#define LI_AFG(A1, A2) \
    if (S1 == A1 && S2 == A2) return affine_gather_T<T, A1, A2>(a, ker1, ker2, R, (const T *)src, (T *)dst, err, st, done);
    LI_AFG(1, 1) LI_AFG(2, 2) LI_AFG(3, 3) LI_AFG(4, 4) LI_AFG(2, 4) LI_AFG(4, 2) LI_AFG(1, 2) LI_AFG(2, 1) LI_AFG(1, 4) LI_AFG(4, 1)
    LI_AFG(3, 4) LI_AFG(4, 3) LI_AFG(2, 3) LI_AFG(3, 2) LI_AFG(1, 3) LI_AFG(3, 1)
#undef LI_AFG
    return LI_OK;
}

}  // namespace li
