#include <cudaTypedefs.h>

#include <map>
#include <mutex>
#include <utility>

#include "state.h"
#include "tmap.h"

namespace sb {

cudaError_t make_tmap_2d(CUtensorMap* out, const double* base, int ldG, int n_alloc, int box_genes, int box_cells) {
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    }
    cuuint64_t gdim[2] = {cuuint64_t(ldG), cuuint64_t(n_alloc)};
    cuuint64_t gstride[1] = {cuuint64_t(ldG) * 8};
    cuuint32_t box[2] = {cuuint32_t(box_genes), cuuint32_t(box_cells)};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = (box_genes * 8 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

void ensure_dyn_smem(const void* kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return;
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    size_t& cur = done[std::make_pair(kernel, dev)];
    if (bytes > cur) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
        cur = bytes;
    }
}

}  // namespace sb
