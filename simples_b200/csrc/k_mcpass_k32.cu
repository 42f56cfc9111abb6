// Instantiations of the fused MC pass for sweep-GEMM depth <= 32 (see k_mcpass.cuh).
#include "k_mcpass.cuh"

namespace sb {

void launch_mcpass_k32(const McArgs& a, const CUtensorMap& tm, cudaStream_t st) {
    switch (a.d.NQP / 8) {
        case 4: mcp::launch_mcpass_variant<8, 4>(a, tm, st); break;
        case 5: mcp::launch_mcpass_variant<8, 5>(a, tm, st); break;
        case 6: mcp::launch_mcpass_variant<8, 6>(a, tm, st); break;
        case 7: mcp::launch_mcpass_variant<8, 7>(a, tm, st); break;
        case 8: mcp::launch_mcpass_variant<8, 8>(a, tm, st); break;
        default: break;   // excluded by mcpass_supported()
    }
}

}  // namespace sb
