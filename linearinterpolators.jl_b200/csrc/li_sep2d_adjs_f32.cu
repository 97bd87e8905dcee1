// li_sep2d_adjs_f32.cu -- the streaming adjoint of li_sep2d_adjs.cuh for float data (one file per type: build time).
#include "li_sep2d_adjs.cuh"

namespace li {
template int sep2d_adj_stream<float>(SepArgs<float> &, const li_table *, const li_table *, int64_t, cudaStream_t, bool *);
}  // namespace li
