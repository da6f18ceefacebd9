// tcgen05 scan, split-bf16 operands (exact-fp32 search on the tensor pipe), cta_group::2.
#include "bb_scan_tc.cuh"

namespace bb {
cudaError_t launch_scan_tc_split(const ScanParams& P, const void* tq, const void* te, const void* tql, const void* tel, bool row_cons,
                                 int n_groups, cudaStream_t stream) {
  return row_cons ? launch_scan_tc_variant<2, true, true, false>(P, tq, te, tql, tel, n_groups, stream)
                  : launch_scan_tc_variant<2, false, true, false>(P, tq, te, tql, tel, n_groups, stream);
}
}  // namespace bb
