#include <stdlib.h>
#include "common.cuh"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/mobi_b200.h"

namespace mobi {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    return fn;
}

int make_tensor_map_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box) {
    auto enc = get_encode();
    MOBI_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
    MOBI_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base %p is not 16-byte aligned", base);
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        MOBI_CHECK(box[i] >= 1 && box[i] <= 256, "tensor map box dim %d = %u out of range", i, box[i]);
    }
    for (int i = 0; i + 1 < rank; ++i) {
        gstr[i] = strides_bytes[i];
        MOBI_CHECK((strides_bytes[i] & 15) == 0, "tensor map stride %d = %llu bytes is not a multiple of 16", i,
                   (unsigned long long)strides_bytes[i]);
    }
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                     bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MOBI_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu box %u %u)",
               (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
               rank > 1 ? box[1] : 0);
    return 0;
}

// Unswizzled 2-D map over a row-major matrix of 2- or 4-byte elements, for L2 prefetches of epilogue operands.
int make_tensor_map_2d_plain(CUtensorMap* map, const void* base, int elem_bytes, uint64_t cols, uint64_t rows,
                             uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows) {
    auto enc = get_encode();
    MOBI_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
    MOBI_CHECK(elem_bytes == 2 || elem_bytes == 4, "plain tensor map: 2- or 4-byte elements");
    MOBI_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (row_stride_bytes & 15) == 0 &&
                   (box_cols * (uint64_t)elem_bytes) % 16 == 0 && box_cols >= 1 && box_cols <= 256 && box_rows >= 1 && box_rows <= 256,
               "plain tensor map: base / stride / box not 16-byte granular or box out of range");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {row_stride_bytes};
    cuuint32_t bdim[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MOBI_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (plain) failed with CUresult %d", (int)r);
    return 0;
}

int sm_count() {
    static int n = 0;
    if (n) return n;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    return n;
}

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("MOBI_PDL");
        on = (e && e[0] == '1') ? 1 : 0;   // opt-in: measured 1 % slower inside the captured UNet graph (profiles/r02/)
    }
    return on != 0;
}

bool res_prefetch_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("MOBI_RES_PREFETCH");
        on = (e && e[0] == '0') ? 0 : 1;   // on by default; MOBI_RES_PREFETCH=0 is the A/B arm
    }
    return on != 0;
}

}  // namespace mobi

extern "C" const char* mobi_last_error(void) { return mobi::g_err; }
extern "C" int mobi_version(void) { return 100; }
