// detect_multi.cu -- multi-template instantiations of the fused detect kernel
#define THR_MULTI 1
#include "variants_impl.cuh"
