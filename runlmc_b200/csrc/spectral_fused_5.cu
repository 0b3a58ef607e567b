// Explicit instantiations of the fused spectral kernel for D = 11, 12 (see spectral_fused.cuh).
#include "spectral_fused.cuh"

namespace lmc {
LMC_FUSED_INSTANTIATE(11) LMC_FUSED_INSTANTIATE(12)
}  // namespace lmc
