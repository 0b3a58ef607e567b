// Explicit instantiations of the fused spectral kernel for D = 9, 10 (see spectral_fused.cuh).
#include "spectral_fused.cuh"

namespace lmc {
LMC_FUSED_INSTANTIATE(9) LMC_FUSED_INSTANTIATE(10)
}  // namespace lmc
