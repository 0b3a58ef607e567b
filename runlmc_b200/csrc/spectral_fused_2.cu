// Explicit instantiations of the fused spectral kernel for D = 5, 6 (see spectral_fused.cuh).
#include "spectral_fused.cuh"

namespace lmc {
LMC_FUSED_INSTANTIATE(5) LMC_FUSED_INSTANTIATE(6)
}  // namespace lmc
