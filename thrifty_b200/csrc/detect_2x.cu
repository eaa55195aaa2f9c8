// detect_2x.cu -- block_len = 32768 as two interleaved 16384-point transforms (detect_kernel_2x.cuh)
#include "detect_kernel_2x.cuh"
#include "variants.h"

namespace thr {

bool pick_variant_2x(int n, bool multi, Variant *out) {
    if (n != Cfg2x::NB) return false;
    using H = Cfg2x::Half;
    Variant v;
    v.log2n = 15;
    v.threads = Cfg2x::T;
    v.gmem = false;
    v.two_halves = true;
    v.r2 = Cfg2x::R2;
    v.r3 = Cfg2x::R3;
    v.i3 = 2;
    v.p3_item = [](int tid, int it) { return H::p3_item(tid, it); };
    v.launch_threads = Cfg2x::LAUNCH_THREADS;
    v.worker_regs = 112;               // detect_kernel_2x.cuh: setmaxnreg.inc 112 / dec 32
    v.smem = Cfg2x::smem_bytes();
    v.fn = multi ? (const void *)&detect2x_kernel<true> : (const void *)&detect2x_kernel<false>;
    v.name = multi ? "detect2x_kernel<N=32768 as 2x16384,T=512,smem+L2 park,multi>"
                   : "detect2x_kernel<N=32768 as 2x16384,T=512,smem+L2 park>";
    *out = v;
    return true;
}

}  // namespace thr
