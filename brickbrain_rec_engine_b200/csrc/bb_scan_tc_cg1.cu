// tcgen05 scan with single-CTA groups (cta_group::1; BB_CTA_GROUP=1): kept as a debugging aid for the 2-CTA protocol.
#include "bb_scan_tc.cuh"

namespace bb {
cudaError_t launch_scan_tc_cg1(const ScanParams& P, const void* tq, const void* te, bool row_cons, int n_groups, cudaStream_t stream) {
  return row_cons ? launch_scan_tc_variant<1, true, false, false>(P, tq, te, nullptr, nullptr, n_groups, stream)
                  : launch_scan_tc_variant<1, false, false, false>(P, tq, te, nullptr, nullptr, n_groups, stream);
}
}  // namespace bb
