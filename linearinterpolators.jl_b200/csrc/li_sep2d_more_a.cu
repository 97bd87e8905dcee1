// li_sep2d_more_a.cu -- group a of the additional kernel-size pairs of the fused 2-D kernels (li_sep2d_more.cuh).
#include "li_sep2d_more.cuh"

namespace li {
LI_MORE_GROUP_FN(sep2d_more_a, LI_MORE(8, 8) LI_MORE(8, 4) LI_MORE(4, 8) LI_MORE(2, 8))
}  // namespace li
