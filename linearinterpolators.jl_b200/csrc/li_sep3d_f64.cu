// li_sep3d_f64.cu -- the fused 3-D kernels of li_sep3d.cuh for double data (one file per type: build time).
#include "li_sep3d.cuh"

namespace li {
template int sep3d_fused_T<double>(int, const li_table *, const li_table *, const li_table *, int64_t, double, const void *,
                                  double, void *, cudaStream_t, bool *);
}  // namespace li
