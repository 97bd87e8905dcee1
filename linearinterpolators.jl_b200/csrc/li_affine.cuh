#pragma once
// li_affine.cuh -- 2-D interpolation through an affine coordinate transform (K9/K10).
//
// Reference: apply!(alpha, P, ker1, ker2, R, src, beta, dst), src/nonseparable.jl:147-264.
// For output pixel (i1, i2) (1-based):
//     x = R.xx*T(i1) + (R.xy*T(i2) + R.x),  y = R.yx*T(i1) + (R.yy*T(i2) + R.y)
//     (j1, w1) = getcoefs(ker1, lim1, x),   (j2, w2) = getcoefs(ker2, lim2, y)
//     Direct : dst[i1,i2] = sum_{s2} (sum_{s1} src[j1,j2]*w1)*w2           (:185,195)
//     Adjoint: dst[j1,j2] += w1 * (w2 * (alpha*src[i1,i2]))                  (:225-232)
// Coordinates and coefficients are evaluated with un-fused IEEE operations so the indices and
// weights are the reference's bit for bit; they are computed once per pixel and reused for all
// images of the batch (the reference loops over images outside the operator).
#include <cuda.h>   // CUtensorMap (the encoder is reached through cudaGetDriverEntryPoint: no libcuda link)

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "li_common.cuh"
#include "li_kerdev.cuh"
#include "li_sep2d.cuh"
#include "li_weights.cuh"

namespace li {

template <typename T>
struct AffineArgs {
    KerDev<T> k1, k2;
    T xx, xy, x, yx, yy, y;
    int n1, n2, m1, m2;
    int64_t nbatch;
    T alpha, beta;
};

template <typename T, int S1, int S2>
__device__ __forceinline__ bool affine_coefs(const AffineArgs<T> &a, int i1, int i2, int &b1, int &b2, T (&w1)[S1],
                                             T (&w2)[S2])
{
    using A = Ar<T>;
    const T pos1 = (T)(i1 + 1), pos2 = (T)(i2 + 1);
    const T off1 = A::add(A::mul(a.xy, pos2), a.x);
    const T off2 = A::add(A::mul(a.yy, pos2), a.y);
    const T px = A::add(A::mul(a.xx, pos1), off1);
    const T py = A::add(A::mul(a.yx, pos1), off2);
    bool ok = getcoefs<T, S1>(a.k1, a.m1, px, b1, w1);
    ok = getcoefs<T, S2>(a.k2, a.m2, py, b2, w2) && ok;
    return ok;
}


// Direct: one thread per output pixel, looping over the images of its batch slice.
template <typename T, int S1, int S2>
__global__ void __launch_bounds__(256)
affine_direct_kernel(AffineArgs<T> a, const T *__restrict__ src, T *__restrict__ dst, int *__restrict__ err)
{
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i2 = blockIdx.y * blockDim.y + threadIdx.y;
    if (i1 >= a.n1 || i2 >= a.n2) return;
    T w1[S1], w2[S2];
    int b1, b2;
    if (!affine_coefs<T, S1, S2>(a, i1, i2, b1, b2, w1, w2)) {
        atomicExch(err, 1);
        return;
    }
    int j1[S1];
    int64_t o2[S2];
#pragma unroll
    for (int s = 0; s < S1; ++s) j1[s] = clampi(b1 + s, a.m1 - 1);
#pragma unroll
    for (int s = 0; s < S2; ++s) o2[s] = (int64_t)clampi(b2 + s, a.m2 - 1) * a.m1;
    const int64_t splane = (int64_t)a.m1 * a.m2, dplane = (int64_t)a.n1 * a.n2;
    const bool plain = a.alpha == (T)1 && a.beta == (T)0;
    for (int64_t b = blockIdx.z; b < a.nbatch; b += gridDim.z) {
        const T *sp = src + b * splane;
        T e = (T)0;
#pragma unroll
        for (int s2 = 0; s2 < S2; ++s2) {
            const T *col = sp + o2[s2];
            T sub = col[j1[0]] * w1[0];
#pragma unroll
            for (int s1 = 1; s1 < S1; ++s1) sub += col[j1[s1]] * w1[s1];
            e = s2 == 0 ? sub * w2[0] : e + sub * w2[s2];
        }
        T *d = dst + b * dplane + (int64_t)i2 * a.n1 + i1;
        // :185 / :195 -- the general branch reads dst even when beta == 0, like the reference
        *d = plain ? e : e * a.alpha + (*d) * a.beta;
    }
}

// Adjoint, atomic scatter variant (dst pre-scaled by beta).
template <typename T, int S1, int S2>
__global__ void __launch_bounds__(256)
affine_adjoint_atomic_kernel(AffineArgs<T> a, const T *__restrict__ src, T *__restrict__ dst, int *__restrict__ err)
{
    // a warp covers an 8 x 4 patch of pixels (block = 64 x 4 pixels = 8 patches side by side): whatever the rotation,
    // the 32 atomics of one instruction land in a compact 2-D neighbourhood, i.e. on few cache lines
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int i1 = blockIdx.x * 64 + w * 8 + (lane & 7);
    const int i2 = blockIdx.y * 4 + (lane >> 3);
    if (i1 >= a.n1 || i2 >= a.n2) return;
    T w1[S1], w2[S2];
    int b1, b2;
    if (!affine_coefs<T, S1, S2>(a, i1, i2, b1, b2, w1, w2)) {
        atomicExch(err, 1);
        return;
    }
    const int64_t splane = (int64_t)a.n1 * a.n2, dplane = (int64_t)a.m1 * a.m2;
    const bool inner = b1 >= 0 && b1 + S1 <= a.m1 && b2 >= 0 && b2 + S2 <= a.m2;
    if (inner) {   // no clamped tap: S1*S2 distinct targets
        const int64_t o0 = (int64_t)b2 * a.m1 + b1;
        for (int64_t b = blockIdx.z; b < a.nbatch; b += gridDim.z) {
            T val = src[b * splane + (int64_t)i2 * a.n1 + i1];
            if (a.alpha != (T)1) val = a.alpha * val;
            T *dp = dst + b * dplane + o0;
#pragma unroll
            for (int s2 = 0; s2 < S2; ++s2) {
                const T tmp = w2[s2] * val;   // :225-232
#pragma unroll
                for (int s1 = 0; s1 < S1; ++s1) atomicAdd(dp + (int64_t)s2 * a.m1 + s1, w1[s1] * tmp);
            }
        }
        return;
    }
    // Pixels with clamped taps (at or beyond the border of the destination): many of them land on the same few
    // border elements -- whole regions of a rotated frame do -- and plain atomics would serialise there.  The taps
    // that clamp onto one element are summed first (per dimension), then lanes of the warp that target the same
    // element are combined with shuffles and only one of them issues the atomic.
    const int t1 = clampi(b1, a.m1 - 1), n1d = clampi(b1 + S1 - 1, a.m1 - 1) - t1 + 1;
    const int t2 = clampi(b2, a.m2 - 1), n2d = clampi(b2 + S2 - 1, a.m2 - 1) - t2 + 1;
    T f1[S1], f2[S2];
#pragma unroll
    for (int k = 0; k < S1; ++k) {
        f1[k] = (T)0;
#pragma unroll
        for (int s = 0; s < S1; ++s)
            if (clampi(b1 + s, a.m1 - 1) - t1 == k) f1[k] += w1[s];
    }
#pragma unroll
    for (int k = 0; k < S2; ++k) {
        f2[k] = (T)0;
#pragma unroll
        for (int s = 0; s < S2; ++s)
            if (clampi(b2 + s, a.m2 - 1) - t2 == k) f2[k] += w2[s];
    }
    for (int64_t b = blockIdx.z; b < a.nbatch; b += gridDim.z) {
        T val = src[b * splane + (int64_t)i2 * a.n1 + i1];
        if (a.alpha != (T)1) val = a.alpha * val;
        T *dp = dst + b * dplane;
#pragma unroll
        for (int k2 = 0; k2 < S2; ++k2) {
#pragma unroll
            for (int k1 = 0; k1 < S1; ++k1) {
                if (k2 < n2d && k1 < n1d) {
                    const long long tgt = (long long)(t2 + k2) * a.m1 + (t1 + k1);
                    T v = f1[k1] * (f2[k2] * val);
                    const unsigned act = __activemask();
                    const unsigned peers = __match_any_sync(act, tgt);
                    const int leader = __ffs(peers) - 1;
                    for (unsigned m = peers & (peers - 1); m; m &= m - 1) {   // the other lanes of the group, in lane order
                        const T o = __shfl_sync(peers, v, __ffs(m) - 1);
                        if ((int)(threadIdx.x + threadIdx.y * blockDim.x) % 32 == leader) v += o;
                    }
                    if ((int)(threadIdx.x + threadIdx.y * blockDim.x) % 32 == leader) atomicAdd(dp + tgt, v);
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Tiled forward.  A CTA owns a TX x TY tile of output pixels (PP pixels per thread); the coefficients of its pixels are
// evaluated once and kept in registers for all images of the batch.  The clamped tap indices of the tile are reduced
// to their bounding box in the source image, which is staged in shared memory image by image (one tensor-map TMA copy
// per box, several images in flight; plain loads when the rows are not 16-byte aligned), then each pixel reads its
// S1*S2 taps from shared memory instead of global memory.  The row pitch of the box is chosen on the host (mod 32) so
// that the lanes of a warp, which walk through the box along the direction (R.xx, R.yx), fall into different banks.
// Tiles whose box does not fit (extreme magnification) gather from global memory inside the same kernel.
// (A shared-memory privatised ADJOINT of the same shape was measured at half the speed of the global atomics of
// affine_adjoint_atomic_kernel -- float atomics on shared memory are CAS loops -- and was removed.)
constexpr int AFT_TX = 32, AFT_TY = 32, AFT_NT = 512, AFT_PP = AFT_TX * AFT_TY / AFT_NT;
constexpr int AFT_NSTG = 6;   // direct: boxes (images) in flight per CTA -- the per-image work of a tile is far shorter
                              // than a DRAM round trip, so the copies run several images ahead

template <typename T, int S1, int S2>
struct TilePix {
    T w1[S1], w2[S2];
    int b1, b2;
    bool ok;      // inside the image and coefficients valid
};

template <typename T, int S1, int S2>
__global__ void __launch_bounds__(AFT_NT, 2)
affine_tiled_kernel(AffineArgs<T> a, const T *__restrict__ src, T *__restrict__ dst, int *__restrict__ err,
                    const int pmod, const int cap /* elements per box buffer */,
                    const int bulk /* direct: boxes arrive by TMA (tmap) */, const int PB, const int HB /* TMA box */,
                    const __grid_constant__ CUtensorMap tmap)
{
    constexpr int TX = AFT_TX, TY = AFT_TY, NT = AFT_NT, PP = AFT_PP, RY = TY / PP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *box = reinterpret_cast<T *>(smem_raw);   // direct: [AFT_NSTG][cap]; adjoint: [cap]
    __shared__ int sbb[4];                      // lo1, hi1, lo2, hi2 of the clamped taps
    __shared__ __align__(8) uint64_t bar[AFT_NSTG];   // direct: one mbarrier per box buffer (TMA bulk copies)

    const int tid = threadIdx.x, lx = tid % TX, ly = tid / TX;
    const int i1 = blockIdx.x * TX + lx;
    if (tid == 0) {
        sbb[0] = sbb[2] = 0x7fffffff;
        sbb[1] = sbb[3] = -1;
#pragma unroll
        for (int k = 0; k < AFT_NSTG; ++k) mbar_init(&bar[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    TilePix<T, S1, S2> px[PP];
    int lo1 = 0x7fffffff, hi1 = -1, lo2 = 0x7fffffff, hi2 = -1;
#pragma unroll
    for (int p = 0; p < PP; ++p) {
        const int i2 = blockIdx.y * TY + ly + p * RY;
        px[p].ok = i1 < a.n1 && i2 < a.n2;
        px[p].b1 = px[p].b2 = 0;
        if (px[p].ok) {
            if (!affine_coefs<T, S1, S2>(a, i1, i2, px[p].b1, px[p].b2, px[p].w1, px[p].w2)) {
                atomicExch(err, 1);
                px[p].ok = false;
            }
        }
        if (px[p].ok) {
            lo1 = min(lo1, clampi(px[p].b1, a.m1 - 1));
            hi1 = max(hi1, clampi(px[p].b1 + S1 - 1, a.m1 - 1));
            lo2 = min(lo2, clampi(px[p].b2, a.m2 - 1));
            hi2 = max(hi2, clampi(px[p].b2 + S2 - 1, a.m2 - 1));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo1 = min(lo1, __shfl_xor_sync(0xffffffffu, lo1, o));
        hi1 = max(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
        lo2 = min(lo2, __shfl_xor_sync(0xffffffffu, lo2, o));
        hi2 = max(hi2, __shfl_xor_sync(0xffffffffu, hi2, o));
    }
    if ((tid & 31) == 0) {
        atomicMin(&sbb[0], lo1);
        atomicMax(&sbb[1], hi1);
        atomicMin(&sbb[2], lo2);
        atomicMax(&sbb[3], hi2);
    }
    __syncthreads();
    lo1 = sbb[0], hi1 = sbb[1], lo2 = sbb[2], hi2 = sbb[3];
    if (hi1 < 0) return;                         // no valid pixel in the tile (block-uniform)
    constexpr int AL = 16 / (int)sizeof(T);
    if (bulk) lo1 &= ~(AL - 1);   // the innermost TMA coordinate must be 16-byte aligned (else: illegal instruction)
    const int BW = hi1 - lo1 + 1, BH = hi2 - lo2 + 1;
    // row pitch of the staged box: the TMA box width (chosen on the host), or the smallest P >= BW with P % 32 == pmod
    const int P = bulk ? PB : BW + ((pmod - BW) & 31);
    const bool staged = bulk ? (BW <= PB && BH <= HB) : ((int64_t)P * BH <= cap);   // block-uniform

    const int64_t splane = (int64_t)a.m1 * a.m2, dplane = (int64_t)a.n1 * a.n2;   // planes of src / dst
    const bool plain = a.alpha == (T)1 && a.beta == (T)0;

    // per pixel: offset of tap (0,0) in the box when no tap is clamped, else -1 (taps are clamped one by one)
    int off[PP];
#pragma unroll
    for (int p = 0; p < PP; ++p) {
        const bool inner = px[p].b1 >= 0 && px[p].b1 + S1 <= a.m1 && px[p].b2 >= 0 && px[p].b2 + S2 <= a.m2;
        off[p] = inner ? (px[p].b2 - lo2) * P + (px[p].b1 - lo1) : -1;
    }
    auto tap_off = [&](int p, int s1, int s2) {   // box offset of tap (s1, s2) of pixel p, any pixel
        return (clampi(px[p].b2 + s2, a.m2 - 1) - lo2) * P + (clampi(px[p].b1 + s1, a.m1 - 1) - lo1);
    };
    const int warp = tid >> 5, lane = tid & 31;

    {
        // box of image b -> buffer k: ONE tensor-map TMA copy of the PB x HB box at (lo1, lo2, b) issued by thread 0
        // (elements outside the image are zero-filled and never read), or plain loads by everybody when the rows
        // of the image are not 16-byte aligned
        auto stage = [&](int64_t b, int k) {
            T *buf = box + (size_t)k * cap;
            if (bulk) {
                if (tid == 0) {
                    mbar_expect_tx(&bar[k], (uint32_t)(PB * HB * sizeof(T)));
                    asm volatile(
                        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
                            "r"(smem_u32(buf)), "l"(&tmap), "r"(lo1), "r"(lo2), "r"((int)b), "r"(smem_u32(&bar[k]))
                        : "memory");
                }
            } else {
                const T *sp = src + b * splane + (int64_t)lo2 * a.m1 + lo1;
                for (int r = warp; r < BH; r += NT / 32)
                    for (int c = lane; c < BW; c += 32) buf[r * P + c] = sp[(int64_t)r * a.m1 + c];
            }
        };
        const int64_t bstep = gridDim.z;
        int64_t b = blockIdx.z;
        const int NS = bulk ? AFT_NSTG : 1;   // (plain loads are staged one image at a time)
        if (staged)
            for (int k = 0; k < NS - 1; ++k)
                if (b + k * bstep < a.nbatch) stage(b + k * bstep, k);

        // interior tiles of the plain operator (alpha = 1, beta = 0, no clamped tap, every pixel valid): a loop
        // over the images with nothing but the S1*S2 shared-memory loads, the arithmetic and one store per pixel
        bool fast = staged && bulk && plain;
#pragma unroll
        for (int p = 0; p < PP; ++p) fast = fast && px[p].ok && off[p] >= 0;
        if (__syncthreads_and(fast)) {
            T *d[PP];
#pragma unroll
            for (int p = 0; p < PP; ++p)
                d[p] = dst + b * dplane + (int64_t)(blockIdx.y * TY + ly + p * RY) * a.n1 + i1;
            const int64_t dstride = dplane * bstep;
            const int nimg = (int)((a.nbatch - b + bstep - 1) / bstep);
            int k = 0, kn = NS - 1;         // buffer of image `it`, buffer refilled at image `it`
            uint32_t par = 0;               // parity of buffer k's current use
            for (int it = 0; it < nimg; ++it) {
                if (it + NS - 1 < nimg) stage(b + (int64_t)(it + NS - 1) * bstep, kn);
                mbar_wait(&bar[k], par);
                const T *buf = box + (size_t)k * cap;
#pragma unroll
                for (int p = 0; p < PP; ++p) {
                    const T *q = buf + off[p];
                    T e = (T)0;
#pragma unroll
                    for (int s2 = 0; s2 < S2; ++s2) {
                        T sub = q[0] * px[p].w1[0];
#pragma unroll
                        for (int s1 = 1; s1 < S1; ++s1) sub += q[s1] * px[p].w1[s1];
                        e = s2 == 0 ? sub * px[p].w2[0] : e + sub * px[p].w2[s2];
                        q += P;
                    }
                    *d[p] = e;
                    d[p] += dstride;
                }
                if (++k == NS) k = 0, par ^= 1u;
                if (++kn == NS) kn = 0;
                __syncthreads();   // every thread is done with this buffer before it is refilled
            }
            return;
        }

        for (int it = 0; b < a.nbatch; b += bstep, ++it) {
            const T *buf = box + (size_t)(it % NS) * cap;
            if (staged) {
                // (the buffer refilled now was last read one image ago; the barrier at the end of that image
                // orders those reads before the refill)
                if (b + (NS - 1) * bstep < a.nbatch) stage(b + (NS - 1) * bstep, (it + NS - 1) % NS);
                if (bulk)
                    mbar_wait(&bar[it % NS], (uint32_t)((it / NS) & 1));
                else
                    __syncthreads();
            }
            const T *sp = src + b * splane;
#pragma unroll
            for (int p = 0; p < PP; ++p) {
                if (!px[p].ok) continue;
                T e = (T)0;
                if (staged && off[p] >= 0) {
                    const T *q = buf + off[p];
#pragma unroll
                    for (int s2 = 0; s2 < S2; ++s2) {
                        T sub = q[s2 * P] * px[p].w1[0];
#pragma unroll
                        for (int s1 = 1; s1 < S1; ++s1) sub += q[s2 * P + s1] * px[p].w1[s1];
                        e = s2 == 0 ? sub * px[p].w2[0] : e + sub * px[p].w2[s2];
                    }
                } else {
#pragma unroll
                    for (int s2 = 0; s2 < S2; ++s2) {
                        T sub = (T)0;
#pragma unroll
                        for (int s1 = 0; s1 < S1; ++s1) {
                            const T v = staged ? buf[tap_off(p, s1, s2)]
                                               : sp[(int64_t)clampi(px[p].b2 + s2, a.m2 - 1) * a.m1 + clampi(px[p].b1 + s1, a.m1 - 1)];
                            sub = s1 == 0 ? v * px[p].w1[0] : sub + v * px[p].w1[s1];
                        }
                        e = s2 == 0 ? sub * px[p].w2[0] : e + sub * px[p].w2[s2];
                    }
                }
                T *d = dst + b * dplane + (int64_t)(blockIdx.y * TY + ly + p * RY) * a.n1 + i1;
                *d = plain ? e : e * a.alpha + (*d) * a.beta;   // :185 / :195
            }
            if (staged) __syncthreads();   // every thread is done with this buffer before it is refilled
        }
    }
}

// row pitch (mod 32) of the staged box that spreads the lanes of a warp over the banks: lane l of a warp reads
// column floor(l*xx + u) and row floor(l*yx + v) of the box; pick the residue with the fewest wavefronts on average
static int best_pitch_mod(double xx, double yx, int step = 1)
{
    int best = 1;
    double best_cost = 1e30;
    for (int pm = 0; pm < 32; pm += step) {
        double cost = 0;
        for (int trial = 0; trial < 8; ++trial) {
            const double u = 0.137 * trial + 0.05, v = 0.291 * trial + 0.11;
            int addr[32];
            for (int l = 0; l < 32; ++l)
                addr[l] = (int)std::floor(l * xx + u + 64) + (int)std::floor(l * yx + v + 64) * (64 + pm);
            int worst = 0;
            for (int bank = 0; bank < 32; ++bank) {   // distinct addresses per bank = wavefronts
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
        if (cost < best_cost - 1e-9) best_cost = cost, best = pm;
    }
    return best;
}

template <typename T, int S1, int S2>
static int launch_affine(int adjoint, const AffineArgs<T> &a, const T *src, T *dst, int *err, cudaStream_t st);

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

template <typename T, int S1, int S2>
static int launch_affine_tiled(int adjoint, const AffineArgs<T> &a, const T *src, T *dst, int *err, cudaStream_t st)
{
    constexpr int AL = 16 / (int)sizeof(T);
    dim3 grid((a.n1 + AFT_TX - 1) / AFT_TX, (a.n2 + AFT_TY - 1) / AFT_TY, 1);
    LI_REQUIRE(grid.y <= 65535, LI_ERR_UNSUPPORTED, "image too tall for the affine kernel");
    const int64_t tiles = (int64_t)grid.x * grid.y;
    // split the batch over grid.z only when the tiles alone cannot fill the machine (coefficients are per CTA)
    grid.z = (unsigned)std::max<int64_t>(1, std::min<int64_t>(a.nbatch, (148 * 4 + tiles - 1) / tiles));
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    if (adjoint) return launch_affine<T, S1, S2>(adjoint, a, src, dst, err, st);   // (the adjoint has no tiled form)
    // direct: source box of a tile = extent of TX x TY pixels under the linear part of R, plus the taps
    const double ex = (AFT_TX - 1) * std::fabs((double)a.xx) + (AFT_TY - 1) * std::fabs((double)a.xy);
    const double ey = (AFT_TX - 1) * std::fabs((double)a.yx) + (AFT_TY - 1) * std::fabs((double)a.yy);
    const int need1 = (int)std::ceil(std::min(ex, 1e6)) + S1 + 2 + (AL - 1);   // (+ alignment of the first column)
    const int HB = (int)std::ceil(std::min(ey, 1e6)) + S2 + 2;
    int bulk = (a.m1 % AL == 0) && (((uintptr_t)src) % 16 == 0) && encode_tiled() != nullptr && a.nbatch < (1LL << 31);
    int pmod = best_pitch_mod((double)a.xx, (double)a.yx, bulk ? AL : 1);
    int PB = need1 + ((pmod - need1) & 31);          // smallest width >= need1 with PB % 32 == pmod (a multiple of AL)
    int cap = (int)(14 * 1024 / sizeof(T));
    if (bulk && (PB > 256 || HB > 256 || (int64_t)PB * HB > cap)) bulk = 0;   // (huge magnification: plain staging / global taps)
    if (bulk) {
        cap = (PB * HB + 31) & ~31;                  // buffers stay 128-byte aligned
        const cuuint64_t gdim[3] = {(cuuint64_t)a.m1, (cuuint64_t)a.m2, (cuuint64_t)a.nbatch};
        const cuuint64_t gstr[2] = {(cuuint64_t)a.m1 * sizeof(T), (cuuint64_t)a.m1 * a.m2 * sizeof(T)};
        const cuuint32_t box[3] = {(cuuint32_t)PB, (cuuint32_t)HB, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode_tiled()(&tmap, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                                          3, (void *)src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            bulk = 0;
            cap = (int)(14 * 1024 / sizeof(T));
            pmod = best_pitch_mod((double)a.xx, (double)a.yx, 1);
        }
    }
    const size_t smem = (size_t)cap * sizeof(T) * (bulk ? AFT_NSTG : 1);
    auto kern = affine_tiled_kernel<T, S1, S2>;
    LI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
    kern<<<grid, AFT_NT, smem, st>>>(a, src, dst, err, pmod, cap, bulk, PB, HB, tmap);
    LI_CHECK_LAUNCH();
    return LI_OK;
}

template <typename T, int S1, int S2>
static int launch_affine(int adjoint, const AffineArgs<T> &a, const T *src, T *dst, int *err, cudaStream_t st)
{
    dim3 block(64, 4);
    const int64_t pix_blocks = (int64_t)((a.n1 + 63) / 64) * ((a.n2 + 3) / 4);
    // enough z-slices to fill the machine for small images, few for large ones (weights are reused per slice)
    int64_t z = std::max<int64_t>(1, std::min<int64_t>(a.nbatch, (148 * 16 + pix_blocks - 1) / pix_blocks));
    dim3 grid((a.n1 + 63) / 64, (a.n2 + 3) / 4, (unsigned)std::min<int64_t>(z, 65535));
    LI_REQUIRE(grid.y <= 65535, LI_ERR_UNSUPPORTED, "image too tall for the affine kernel");
    if (adjoint)
        affine_adjoint_atomic_kernel<T, S1, S2><<<grid, block, 0, st>>>(a, src, dst, err);
    else
        affine_direct_kernel<T, S1, S2><<<grid, block, 0, st>>>(a, src, dst, err);
    LI_CHECK_LAUNCH();
    return LI_OK;
}

template <typename T>
int affine_T(int adjoint, const li_kernel_desc *ker1, const li_kernel_desc *ker2, const double R[6], int64_t n1,
                    int64_t n2, int64_t m1, int64_t m2, int64_t nbatch, double alpha, const void *src, double beta,
                    void *dst, int *err, bool tiled, cudaStream_t st)
{
    AffineArgs<T> a;
    int rc = make_kerdev<T>(ker1, m1, &a.k1);
    if (rc == LI_OK) rc = make_kerdev<T>(ker2, m2, &a.k2);
    if (rc != LI_OK) return rc;
    // convert(AffineTransform2D{T}, R): every coefficient rounded to T (src/nonseparable.jl:58)
    a.xx = (T)R[0], a.xy = (T)R[1], a.x = (T)R[2], a.yx = (T)R[3], a.yy = (T)R[4], a.y = (T)R[5];
    a.n1 = (int)n1, a.n2 = (int)n2, a.m1 = (int)m1, a.m2 = (int)m2;
    a.nbatch = nbatch;
    a.alpha = (T)alpha, a.beta = (T)beta;
    const int S1 = ker1->S, S2 = ker2->S;
#define LI_AFF(A1, A2)                                                                                              \
    if (S1 == A1 && S2 == A2)                                                                                       \
        return tiled ? launch_affine_tiled<T, A1, A2>(adjoint, a, (const T *)src, (T *)dst, err, st)               \
                     : launch_affine<T, A1, A2>(adjoint, a, (const T *)src, (T *)dst, err, st);
    LI_AFF(1, 1) LI_AFF(1, 2) LI_AFF(1, 3) LI_AFF(1, 4) LI_AFF(1, 6) LI_AFF(1, 8)
    LI_AFF(2, 1) LI_AFF(2, 2) LI_AFF(2, 3) LI_AFF(2, 4) LI_AFF(2, 6) LI_AFF(2, 8)
    LI_AFF(3, 1) LI_AFF(3, 2) LI_AFF(3, 3) LI_AFF(3, 4) LI_AFF(3, 6) LI_AFF(3, 8)
    LI_AFF(4, 1) LI_AFF(4, 2) LI_AFF(4, 3) LI_AFF(4, 4) LI_AFF(4, 6) LI_AFF(4, 8)
    LI_AFF(6, 1) LI_AFF(6, 2) LI_AFF(6, 3) LI_AFF(6, 4) LI_AFF(6, 6) LI_AFF(6, 8)
    LI_AFF(8, 1) LI_AFF(8, 2) LI_AFF(8, 3) LI_AFF(8, 4) LI_AFF(8, 6) LI_AFF(8, 8)
#undef LI_AFF
    set_error("affine operator: kernel sizes (%d, %d) are not built", S1, S2);
    return LI_ERR_UNSUPPORTED;
}

}  // namespace li
