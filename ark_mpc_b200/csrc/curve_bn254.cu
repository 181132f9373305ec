// ark::Bn254G1 instantiation of the point kernels, part 0: dispatch table, linear gates, mul, mul_gen, sum.
// The slow-to-compile kernels are spread over curve_bn254_beaver.cu, _msm.cu and _shares.cu so they build in parallel.
#define ARK_CURVE_IMPL
#define ARK_CURVE_PART 0
#include "curve_launch.cuh"
namespace arkctx {
template struct CurveLaunch<ark::Bn254G1>;
const CurveOps* curve_ops_bn254() { return CurveLaunch<ark::Bn254G1>::ops(); }
}  // namespace arkctx
