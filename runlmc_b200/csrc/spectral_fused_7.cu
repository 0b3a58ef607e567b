// Explicit instantiations of the fused spectral kernel for D = 15, 16 (see spectral_fused.cuh).
#include "spectral_fused.cuh"

namespace lmc {
LMC_FUSED_INSTANTIATE(15) LMC_FUSED_INSTANTIATE(16)
}  // namespace lmc
