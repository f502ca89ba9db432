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
int make_tensor_map_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box);

int sm_count();

}  // namespace mobi
