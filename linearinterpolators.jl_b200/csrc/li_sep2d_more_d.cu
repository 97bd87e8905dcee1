// li_sep2d_more_d.cu -- group d of the additional kernel-size pairs of the fused 2-D kernels (li_sep2d_more.cuh).
#include "li_sep2d_more.cuh"

namespace li {
LI_MORE_GROUP_FN(sep2d_more_d, LI_MORE(2, 3) LI_MORE(3, 2) LI_MORE(3, 6) LI_MORE(6, 3))
}  // namespace li
