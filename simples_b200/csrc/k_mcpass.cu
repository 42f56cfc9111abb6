// Dispatch of the fused MC pass (k_mcpass.cuh) over its template variants.  The variants are instantiated in
// k_mcpass_k12.cu / _k24.cu / _k32.cu (one translation unit per sweep-GEMM depth) so they compile in parallel.
#include "rng.cuh"
#include "state.h"

namespace sb {

void launch_mcpass_k12(const McArgs& a, const CUtensorMap& tm, cudaStream_t st);   // K0 <= 12
void launch_mcpass_k24(const McArgs& a, const CUtensorMap& tm, cudaStream_t st);   // K0 <= 24
void launch_mcpass_k32(const McArgs& a, const CUtensorMap& tm, cudaStream_t st);   // K0 <= 32

// (sweep-GEMM K-steps, projection n-tiles) pairs that exist
bool mcpass_supported(const Dims& d) {
    if (d.NS != 1) return false;
    const int kst = d.NFP / 4, nt = d.NQP / 8;
    if (d.QS != d.NQP + 4 || d.NQP % 8) return false;
    if (kst <= 3) return nt >= 2 && nt <= 6;
    if (kst <= 6) return nt >= 2 && nt <= 7;
    return kst <= 8 && nt >= 4 && nt <= 8;
}

void launch_mcpass(McArgs a, const CUtensorMap& tmapY8, cudaStream_t st) {
    cudaMemsetAsync(a.counter, 0, sizeof(unsigned), st);
    philox_round_keys(a.seed, a.rk);
    const int kst = a.d.NFP / 4;
    if (kst <= 3) launch_mcpass_k12(a, tmapY8, st);
    else if (kst <= 6) launch_mcpass_k24(a, tmapY8, st);
    else launch_mcpass_k32(a, tmapY8, st);
}

}  // namespace sb
