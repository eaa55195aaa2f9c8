// detect_single.cu -- single-template instantiations of the fused detect kernel
#define THR_MULTI 0
#include "variants_impl.cuh"
