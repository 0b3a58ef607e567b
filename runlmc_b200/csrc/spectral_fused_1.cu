// Explicit instantiations of the fused spectral kernel for D = 3, 4 (see spectral_fused.cuh).
#include "spectral_fused.cuh"

namespace lmc {
LMC_FUSED_INSTANTIATE(3) LMC_FUSED_INSTANTIATE(4)
}  // namespace lmc
