// variants.h -- table of compiled detect_kernel instantiations (one per block length), built in two
// translation units so that they compile in parallel: detect_single.cu (one template, no template
// loop), detect_multi.cu (several templates per detector), detect_fastdet.cu (native-twin semantics).
#pragma once

#include <stddef.h>

namespace thr {

struct Variant {
    int log2n;
    int threads;
    bool gmem;
    bool two_halves = false;           // detect2x_kernel: template spectrum stored as [k < F][k >= F]
    int r2, r3, i3;
    int (*p3_item)(int tid, int it);   // pass-3 item owned by (thread, iteration): fixes the template order
    int launch_threads;
    int worker_regs = 0;               // setmaxnreg target of the workers (0: the kernel does not re-split registers)
    size_t smem;
    const void *fn;
    const char *name;
};

bool pick_variant_single(int block_len, Variant *out);   // n_templates == 1
bool pick_variant_multi(int block_len, Variant *out);    // n_templates >= 1
bool pick_variant_fastdet(int block_len, Variant *out);  // fastdet semantics (one template)
bool pick_variant_stage(int block_len, int stages, Variant *out);   // stage-boundary kernels (one template): stages = 1
                                                                    // (thr_sync_batch) or 2 (thr_soa_batch)
bool pick_variant_2x(int block_len, bool multi, Variant *out);   // 32768 = 2 x 16384 (one / several templates)

}  // namespace thr
