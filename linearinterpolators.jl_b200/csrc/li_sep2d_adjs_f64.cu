// li_sep2d_adjs_f64.cu -- the streaming adjoint of li_sep2d_adjs.cuh for double data (one file per type: build time).
#include "li_sep2d_adjs.cuh"

namespace li {
template int sep2d_adj_stream<double>(SepArgs<double> &, const li_table *, const li_table *, int64_t, cudaStream_t, bool *);
}  // namespace li
