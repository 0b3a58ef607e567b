// Explicit instantiations of the fused spectral kernel for D = 1, 2 (see spectral_fused.cuh).
#include "spectral_fused.cuh"

namespace lmc {
LMC_FUSED_INSTANTIATE(1) LMC_FUSED_INSTANTIATE(2)
}  // namespace lmc
