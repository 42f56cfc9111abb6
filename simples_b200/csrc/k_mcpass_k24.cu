// Instantiations of the fused MC pass for sweep-GEMM depth <= 24 (see k_mcpass.cuh).
#include "k_mcpass.cuh"

namespace sb {

void launch_mcpass_k24(const McArgs& a, const CUtensorMap& tm, cudaStream_t st) {
    switch (a.d.NQP / 8) {
        case 2: mcp::launch_mcpass_variant<6, 2>(a, tm, st); break;
        case 3: mcp::launch_mcpass_variant<6, 3>(a, tm, st); break;
        case 4: mcp::launch_mcpass_variant<6, 4>(a, tm, st); break;
        case 5: mcp::launch_mcpass_variant<6, 5>(a, tm, st); break;
        case 6: mcp::launch_mcpass_variant<6, 6>(a, tm, st); break;
        case 7: mcp::launch_mcpass_variant<6, 7>(a, tm, st); break;
        default: break;   // excluded by mcpass_supported()
    }
}

}  // namespace sb
