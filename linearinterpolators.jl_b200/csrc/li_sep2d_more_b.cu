// li_sep2d_more_b.cu -- group b of the additional kernel-size pairs of the fused 2-D kernels (li_sep2d_more.cuh).
#include "li_sep2d_more.cuh"

namespace li {
LI_MORE_GROUP_FN(sep2d_more_b, LI_MORE(8, 6) LI_MORE(6, 8) LI_MORE(8, 2) LI_MORE(3, 8))
}  // namespace li
