// Explicit instantiations of the fused spectral kernel for D = 13, 14 (see spectral_fused.cuh).
#include "spectral_fused.cuh"

namespace lmc {
LMC_FUSED_INSTANTIATE(13) LMC_FUSED_INSTANTIATE(14)
}  // namespace lmc
