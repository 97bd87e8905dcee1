// li_sep2d.cu -- fused 2-D separable kernels (placeholder: declines, callers fall back to passes).
#include "li_common.cuh"

namespace li {
int sep2d_fused_supported(const li_table *, const li_table *) { return 0; }
int sep2d_fused_apply(int, const li_table *, const li_table *, int64_t, double, const void *, double, void *,
                      cudaStream_t, bool *done)
{
    *done = false;
    return LI_OK;
}
}  // namespace li
