// detect_stage2.cu -- kernels that start at the SoaEstimator's stage boundary (thr_soa_batch): shifted spectrum in
#define THR_MULTI 0
#define THR_STAGES 2
#include "variants_impl.cuh"

namespace thr {
bool pick_variant_stage1(int n, Variant *out);
bool pick_variant_stage(int n, int stages, Variant *out) {
    return stages == 1 ? pick_variant_stage1(n, out) : (stages == 2 ? pick_variant_stage2(n, out) : false);
}
}  // namespace thr
