// li_sep2d_more.cuh -- the fused 2-D kernels (Float32) for the kernel-size pairs (S1, S2) of {2, 3, 4, 6, 8}^2 that
// li_sep2d.cu does not build: streaming forward (li_sep2d_fwd.cuh) and streaming adjoint (li_sep2d_adjs.cuh) only -- what the streaming
// adjoint declines (unaligned sources, down-sampling along dim 1) takes the per-dimension passes for these pairs.
// Split over li_sep2d_more_{a,b,c,d}.cu (LI_MORE_GROUP) to keep the build parallel.
#pragma once
#include "li_sep2d_adjs.cuh"
#include "li_sep2d_fwd.cuh"

namespace li {

template <typename T, int S1, int S2>
static int more_pair(int adjoint, SepArgs<T> &a, const li_table *t1, const li_table *t2, int64_t nbatch, cudaStream_t st,
                     bool *done)
{
    if (adjoint) {
        if (a.m2 < S2 || a.m1 < 2) return LI_OK;
        return stream_T<T, S1, S2>(a, t1, t2, nbatch, st, done);
    }
    return launch_fwd<T, S1, S2>(a, t1, t2, nbatch, st, done);
}

// *mine = the pair (t1->S, t2->S) belongs to this group (whether or not the kernels then declined the problem)
#define LI_MORE_GROUP_FN(NAME, ...)                                                                                    \
    template <typename T>                                                                                              \
    int NAME(int adjoint, SepArgs<T> &a, const li_table *t1, const li_table *t2, int64_t nbatch, cudaStream_t st,      \
             bool *done, bool *mine)                                                                                   \
    {                                                                                                                  \
        const int S1 = t1->S, S2 = t2->S;                                                                              \
        *mine = false;                                                                                                 \
        __VA_ARGS__                                                                                                    \
        return LI_OK;                                                                                                  \
    }                                                                                                                  \
    template int NAME<float>(int, SepArgs<float> &, const li_table *, const li_table *, int64_t, cudaStream_t, bool *, \
                             bool *);                                                                                  \
    template int NAME<double>(int, SepArgs<double> &, const li_table *, const li_table *, int64_t, cudaStream_t,       \
                              bool *, bool *);
// (Float32 only: the Float64 instances of these wide kernels would double the build time of the library for pairs that
// are rarely used in double precision; they take the per-dimension passes)
#define LI_MORE(A1, A2)                                                          \
    if (S1 == A1 && S2 == A2) {                                                  \
        if constexpr (sizeof(T) == 4) {                                          \
            *mine = true;                                                        \
            return more_pair<T, A1, A2>(adjoint, a, t1, t2, nbatch, st, done);   \
        }                                                                        \
    }

}  // namespace li
