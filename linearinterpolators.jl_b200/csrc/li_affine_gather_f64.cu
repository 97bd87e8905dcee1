// li_affine_gather_f64.cu -- the gather-form affine adjoint of li_affine_gather.cuh for double data.
#include "li_affine_gather.cuh"

namespace li {
template int affine_adjoint_gather<double>(const li_kernel_desc *, const li_kernel_desc *, const double[6], int64_t, int64_t, int64_t, int64_t,
                                           int64_t, double, const void *, double, void *, int *, cudaStream_t, bool *);
}  // namespace li
