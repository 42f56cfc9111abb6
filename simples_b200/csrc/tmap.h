// TMA tensor-map helper (driver entry point fetched at run time: no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace sb {

// 2-D FP64 tensor map over Y[n_alloc][ldG] (genes contiguous): box = {box_genes, box_cells},
// SWIZZLE_128B when box_genes * 8 == 128.  Returns cudaSuccess / an error code.
cudaError_t make_tmap_2d(CUtensorMap* out, const double* base, int ldG, int n_alloc, int box_genes, int box_cells);

}  // namespace sb
