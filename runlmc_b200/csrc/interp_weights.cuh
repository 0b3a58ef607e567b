// Keys cubic-convolution weights (reference approx/interpolation.py:21-53), evaluated operation by
// operation like numpy does (no FMA contraction), so they are bit-identical to the reference's.
#pragma once
#include "common.cuh"

namespace lmc {

__device__ __forceinline__ double keys_near(double x) {  // |x| <= 1
    double t = __dadd_rn(__dmul_rn(1.5, x), -2.5);
    t = __dmul_rn(__dmul_rn(t, x), x);
    return __dadd_rn(t, 1.0);
}
__device__ __forceinline__ double keys_far(double x) {  // 1 < |x| <= 2
    double t = __dadd_rn(__dmul_rn(-0.5, x), 2.5);
    t = __dadd_rn(__dmul_rn(t, x), -4.0);
    return __dadd_rn(__dmul_rn(t, x), 2.0);
}
__device__ __forceinline__ void keys_weights(double u, double* w) {
    const double x0 = __dadd_rn(u, 1.0);
    w[0] = (x0 <= 1.0) ? keys_near(x0) : keys_far(x0);
    w[1] = keys_near(u);
    w[2] = keys_near(fabs(__dadd_rn(u, -1.0)));
    w[3] = keys_far(fabs(__dadd_rn(u, -2.0)));
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

}  // namespace lmc
