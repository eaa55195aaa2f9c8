// detect_stage1.cu -- kernels that stop at the Synchronizer's stage boundary (thr_sync_batch): shifted spectrum out
#define THR_MULTI 0
#define THR_STAGES 1
#include "variants_impl.cuh"
