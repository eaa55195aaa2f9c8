// variants_impl.cuh -- included by detect_single.cu / detect_multi.cu / detect_fastdet.cu with THR_MULTI
// (and THR_FASTDET) defined.
#pragma once

#include "detect_kernel.cuh"
#include "variants.h"

#ifndef THR_FASTDET
#define THR_FASTDET 0
#endif
#ifndef THR_STAGES
#define THR_STAGES 0
#endif

namespace thr {

template <int LOG2N, int T, bool GMEM>
static Variant make_variant(const char *name) {
    using C = Cfg<LOG2N, T, GMEM, (THR_FASTDET != 0)>;
    Variant v;
    v.log2n = LOG2N;
    v.threads = T;
    v.gmem = GMEM;
    v.r2 = C::R2;
    v.r3 = C::R3;
    v.i3 = C::I3;
    v.p3_item = [](int tid, int it) { return C::p3_item(tid, it); };
    v.launch_threads = C::LAUNCH_THREADS;
    v.worker_regs = C::SERVICE ? C::WORKER_REGS : 0;
    v.smem = C::smem_bytes(THR_MULTI != 0);
    v.fn = (const void *)&detect_kernel<LOG2N, T, GMEM, (THR_MULTI != 0), (THR_FASTDET != 0), THR_STAGES>;
    v.name = name;
    return v;
}

#if THR_STAGES == 1
#define THR_PICK pick_variant_stage1
#define THR_SUFFIX ",sync>"
#elif THR_STAGES == 2
#define THR_PICK pick_variant_stage2
#define THR_SUFFIX ",soa>"
#elif THR_FASTDET
#define THR_PICK pick_variant_fastdet
#define THR_SUFFIX ",fastdet>"
#elif THR_MULTI
#define THR_PICK pick_variant_multi
#define THR_SUFFIX ",multi>"
#else
#define THR_PICK pick_variant_single
#define THR_SUFFIX ">"
#endif

bool THR_PICK(int n, Variant *out) {
    switch (n) {
#ifndef THR_ONLY_N16384     // experiment builds (tools/variants.sh) carry the headline size only
        case 1024:  *out = make_variant<10, 32, false>("detect_kernel<N=1024,T=32,smem" THR_SUFFIX); return true;
        case 2048:  *out = make_variant<11, 64, false>("detect_kernel<N=2048,T=64,smem" THR_SUFFIX); return true;
        case 4096:  *out = make_variant<12, 128, false>("detect_kernel<N=4096,T=128,smem" THR_SUFFIX); return true;
        case 8192:  *out = make_variant<13, 256, false>("detect_kernel<N=8192,T=256,smem" THR_SUFFIX); return true;
        case 32768: *out = make_variant<15, 512, true>("detect_kernel<N=32768,T=512,gmem" THR_SUFFIX); return true;
#endif
#if !(THR_MULTI && defined(THR_ONLY_N16384) && !defined(THR_KEEP_MULTI))   // (experiment builds drop the multi-template kernel unless asked)
#ifdef THR_N16384_T256      // experiment: 8 fat worker warps (2 items per thread and pass, 232 registers)
        case 16384: *out = make_variant<14, 256, false>("detect_kernel<N=16384,T=256,smem" THR_SUFFIX); return true;
#else
        case 16384: *out = make_variant<14, 512, false>("detect_kernel<N=16384,T=512,smem" THR_SUFFIX); return true;
#endif
#endif
        default: return false;
    }
}

}  // namespace thr
