// detect_fastdet.cu -- fastdet-semantics instantiations of the fused detect kernel (2 transforms per block)
#define THR_MULTI 0
#define THR_FASTDET 1
#include "variants_impl.cuh"
