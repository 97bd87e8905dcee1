// li_sep3d.cuh -- fused 3-D separable interpolation A1*A2*A3, forward and adjoint, one pass over HBM.
//
// Reference: the LazyAlgebra composition of three per-dimension operators (the shape of test/benchmarks.jl:27-29 with a
// third factor), whose factors are the loops of src/tabulated.jl:394-478; the reference makes three full passes with
// two intermediate arrays.  Here the dim-3 factor is applied to the rows of the source planes while they are staged
// for the fused dims-1-2 pipelines (li_sep2d_fwd.cuh, li_sep2d_adjs.cuh, D3 = true): every element of src and dst
// crosses HBM once; source planes shared by neighbouring output planes are re-read from L2.
// Included by li_sep3d_f32.cu / li_sep3d_f64.cu, which instantiate it for one element type each.
#pragma once
#include "li_sep2d_adjs.cuh"
#include "li_sep2d_fwd.cuh"

namespace li {

template <typename T>
int sep3d_fused_T(int adjoint, const li_table *t1, const li_table *t2, const li_table *t3, int64_t nb, double alpha,
                  const void *src, double beta, void *dst, cudaStream_t st, bool *done)
{
    SepArgs<T> a;
    memset(&a, 0, sizeof a);
    a.base1 = t1->base, a.base2 = t2->base, a.base3 = t3->base;
    a.W1 = (const T *)t1->W, a.W2 = (const T *)t2->W, a.W3 = (const T *)t3->W;
    a.n1 = (int)t1->nrows, a.n2 = (int)t2->nrows, a.n3 = (int)t3->nrows;
    a.m1 = (int)t1->ncols, a.m2 = (int)t2->ncols, a.m3 = (int)t3->ncols;
    a.S3 = t3->S;
    a.alpha = (T)alpha, a.beta = (T)beta;
    a.src = (const T *)src, a.dst = (T *)dst;
    int64_t nbatch = 0;
    if (!adjoint) {
        if (t3->S > SEP3_NC || t3->S < 2) return LI_OK;   // decline
        a.nc3 = t3->S;
        nbatch = nb * a.n3;
    } else {
        if (a.m3 < t3->S) return LI_OK;      // folding needs a full window inside the array
        int rc = table_ensure_fold(t3, st);
        int feeders = 0;
        if (rc == LI_OK) rc = table_feed_rows(t3, 1, st, &feeders);
        if (rc != LI_OK) return rc;
        if (feeders > ADS3_NC) {   // down-sampling along dim 3, or many rows hanging over an edge
            if (opt(OPT_TRACE)) fprintf(stderr, "[li_b200] adj 3-D declines: %d planes feed one dst plane > %d\n", feeders, ADS3_NC);
            return LI_OK;
        }
        a.nc3 = std::max(feeders, 1);
        a.bp3 = t3->rowstart, a.bfold3 = t3->bfold, a.W3f = (const T *)t3->Wfold;
        nbatch = nb * a.m3;
    }
    const int S1 = t1->S, S2 = t2->S;
#define LI_SEP(A1, A2)                                                                        \
    if (S1 == A1 && S2 == A2)                                                                 \
        return adjoint ? stream_T<T, A1, A2, true>(a, t1, t2, nbatch, st, done)              \
                       : launch_fwd<T, A1, A2, true>(a, t1, t2, nbatch, st, done);
#ifdef LI_DEV_FAST   // development builds: one pair only
    LI_SEP(4, 4)
#else
    LI_SEP(2, 2) LI_SEP(4, 4) LI_SEP(6, 6) LI_SEP(2, 4) LI_SEP(4, 2) LI_SEP(4, 6) LI_SEP(6, 4) LI_SEP(3, 3)
#endif
#undef LI_SEP
    return LI_OK;   // size pair not built: decline
}

}  // namespace li
