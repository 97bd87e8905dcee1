// li_affine_gather_f32.cu -- the gather-form affine adjoint of li_affine_gather.cuh for float data (one file per type: build time).
#include "li_affine_gather.cuh"

namespace li {
template int affine_adjoint_gather<float>(const li_kernel_desc *, const li_kernel_desc *, const double[6], int64_t, int64_t, int64_t, int64_t,
                                          int64_t, double, const void *, double, void *, int *, cudaStream_t, bool *);
std::mutex &affine_plan_mutex()
{
    static std::mutex mu;
    return mu;
}
std::map<AffineBorderKey, AffineBorderPlan> &affine_plans()
{
    static std::map<AffineBorderKey, AffineBorderPlan> plans;
    return plans;
}
}  // namespace li
