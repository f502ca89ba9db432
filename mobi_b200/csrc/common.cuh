// Host-side helpers shared by all translation units: error reporting across the C ABI and
// CUtensorMap construction through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mobi {

void set_error(const char* fmt, ...);

#define MOBI_CHECK(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            mobi::set_error(__VA_ARGS__); \
            return 1;                    \
        }                                \
    } while (0)

#define MOBI_CUDA(expr)                                                                      \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            mobi::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 2;                                                                        \
        }                                                                                    \
    } while (0)

// Encodes a bf16 tiled tensor map with the 128-byte swizzle. dims/box are innermost-first; strides are in
// bytes for dims 1..rank-1. Returns 0 on success.
int make_tensor_map_2d_plain(CUtensorMap* map, const void* base, int elem_bytes, uint64_t cols, uint64_t rows,
                             uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows);
int make_tensor_map_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box);

int sm_count();

// Programmatic dependent launch (griddepcontrol): kernels that execute pdl_wait() before their first global access are
// launched with cudaLaunchAttributeProgrammaticStreamSerialization, so their CTAs are scheduled and run their prologue
// (barrier initialisation, TMEM allocation, descriptor prefetch) while the previous kernel of the stream drains; inside a
// captured CUDA graph the edge becomes a programmatic dependency.  OPT-IN (MOBI_PDL=1): the A/B of round 2 on the 444-node
// UNet graph (profiles/r02/pdl_ab.md) found it 1 % SLOWER than plain graph edges (95.96 vs 94.80 ms per 64-row step,
// 49.51 vs 49.06 ms at 32 rows) although the eager per-kernel sum improved (91.2 vs 92.9 ms): graph kernel-to-kernel edges
// already cost ~1 us, and early-scheduled dependents hold SM slots next to the persistent GEMM CTAs.
bool pdl_enabled();
bool res_prefetch_enabled();   // L2 prefetch of the GEMM epilogue's residual boxes (MOBI_RES_PREFETCH=0 turns it off)

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace mobi
