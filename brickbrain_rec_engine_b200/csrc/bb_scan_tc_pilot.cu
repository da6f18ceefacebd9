// tcgen05 scan, pilot pass (threshold seeding: group maxima instead of candidates), bf16 and split-bf16 operands.
#include "bb_scan_tc.cuh"

namespace bb {
cudaError_t launch_scan_tc_pilot(const ScanParams& P, const void* tq, const void* te, const void* tql, const void* tel, bool split,
                                 int n_groups, cudaStream_t stream) {
  return split ? launch_scan_tc_variant<2, false, true, true>(P, tq, te, tql, tel, n_groups, stream)
               : launch_scan_tc_variant<2, false, false, true>(P, tq, te, nullptr, nullptr, n_groups, stream);
}
}  // namespace bb
