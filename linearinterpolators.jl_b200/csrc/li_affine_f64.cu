// li_affine_f64.cu -- the affine-transform kernels of li_affine.cuh for double data (one file per type: build time).
#include "li_affine.cuh"

namespace li {
template int affine_T<double>(int, const li_kernel_desc *, const li_kernel_desc *, const double[6], int64_t, int64_t, int64_t, int64_t,
                          int64_t, double, const void *, double, void *, int *, bool, cudaStream_t);
}  // namespace li
