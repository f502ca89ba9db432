// Shared between the fused-attention kernels (attention.cu: any head_dim <= 192, one query tile per CTA, V^T;
// attention3.cu / attention4.cu: head_dim <= 128, V row-major, several query tiles per CTA).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/mobi_b200.h"

namespace mobi {

struct AttnParams {
    int heads, head_dim, tq, tk;
    int nch;        // 64-wide chunks of the head dim
    int dk16;       // k-steps of QK^T  (ceil(d/16))
    int dn;         // N of the PV MMA  (d rounded up to 16)
    int kv_stages;  // 1 or 2            (attention.cu only)
    int p_bufs;     // 1 or 2            (attention.cu only)
    long long ld_out;
    __nv_bfloat16* out;
    float* lse;     // optional [BH, tq]: row max + log2(row sum) in the log2 domain (attention3 only)
};

int attention3_dispatch(const mobi_attn_args* a, const AttnParams& p, cudaStream_t stream);  // V row-major
int attention4_dispatch(const mobi_attn_args* a, const AttnParams& p, cudaStream_t stream);  // V row-major, P in TMEM
bool attention4_supports(int head_dim);

}  // namespace mobi
