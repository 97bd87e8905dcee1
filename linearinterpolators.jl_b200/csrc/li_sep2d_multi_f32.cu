// explicit instantiation of the multi-frame accumulate kernel for Float32 (separate unit: parallel nvcc)
#include "li_sep2d_multi.cuh"
namespace li {
template int sep2d_adj_multi<float>(int64_t, const li_table *const *, const void *const *, double, double, void *, int, int,
                                    cudaStream_t, bool *);
}
