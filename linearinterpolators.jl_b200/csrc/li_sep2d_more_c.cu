// li_sep2d_more_c.cu -- group c of the additional kernel-size pairs of the fused 2-D kernels (li_sep2d_more.cuh).
#include "li_sep2d_more.cuh"

namespace li {
LI_MORE_GROUP_FN(sep2d_more_c, LI_MORE(2, 6) LI_MORE(6, 2) LI_MORE(3, 4) LI_MORE(4, 3) LI_MORE(8, 3))
}  // namespace li
