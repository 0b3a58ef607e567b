// Explicit instantiations of the fused spectral kernel for D = 7, 8 (see spectral_fused.cuh).
#include "spectral_fused.cuh"

namespace lmc {
LMC_FUSED_INSTANTIATE(7) LMC_FUSED_INSTANTIATE(8)
}  // namespace lmc
