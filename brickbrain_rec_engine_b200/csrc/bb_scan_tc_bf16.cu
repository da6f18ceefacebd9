// tcgen05 scan, bf16 operands, cta_group::2 (the production variants) + the dispatcher of all variants.
#include "bb_scan_tc.cuh"

namespace bb {

cudaError_t launch_scan_tc_split(const ScanParams& P, const void* tq, const void* te, const void* tql, const void* tel, bool row_cons,
                                 int n_groups, cudaStream_t stream);
cudaError_t launch_scan_tc_pilot(const ScanParams& P, const void* tq, const void* te, const void* tql, const void* tel, bool split,
                                 int n_groups, cudaStream_t stream);
cudaError_t launch_scan_tc_cg1(const ScanParams& P, const void* tq, const void* te, bool row_cons, int n_groups, cudaStream_t stream);

cudaError_t launch_scan_tc(const ScanParams& P, const void* tmap_q, const void* tmap_e, int cta_group, int n_groups,
                           cudaStream_t stream, const void* tmap_ql, const void* tmap_el) {
  const bool split = tmap_ql != nullptr && tmap_el != nullptr;
  const bool row_cons = P.cons != nullptr && P.class_valid == nullptr;
  if (cta_group != 2) {                    // cta_group::1 (BB_CTA_GROUP=1, a debugging aid): plain bf16 scan only
    if (split || P.pilot_mode) return cudaErrorNotSupported;
    return launch_scan_tc_cg1(P, tmap_q, tmap_e, row_cons, n_groups, stream);
  }
  if (P.pilot_mode) return launch_scan_tc_pilot(P, tmap_q, tmap_e, tmap_ql, tmap_el, split, n_groups, stream);
  if (split) return launch_scan_tc_split(P, tmap_q, tmap_e, tmap_ql, tmap_el, row_cons, n_groups, stream);
  return row_cons ? launch_scan_tc_variant<2, true, false, false>(P, tmap_q, tmap_e, nullptr, nullptr, n_groups, stream)
                  : launch_scan_tc_variant<2, false, false, false>(P, tmap_q, tmap_e, nullptr, nullptr, n_groups, stream);
}

}  // namespace bb
