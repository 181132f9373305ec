// ark::Ed25519 instantiation of the point kernels, part 0: dispatch table, linear gates, mul, mul_gen, sum.
// The slow-to-compile kernels are spread over curve_ed25519_beaver.cu, _msm.cu and _shares.cu so they build in parallel.
#define ARK_CURVE_IMPL
#define ARK_CURVE_PART 0
#include "curve_launch.cuh"
namespace arkctx {
template struct CurveLaunch<ark::Ed25519>;
const CurveOps* curve_ops_ed25519() { return CurveLaunch<ark::Ed25519>::ops(); }
}  // namespace arkctx
