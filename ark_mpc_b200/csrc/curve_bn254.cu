// BN254 G1 instantiation of the point kernels (one curve per translation unit: they compile in parallel).
#define ARK_CURVE_IMPL
#include "curve_launch.cuh"
namespace arkctx {
const CurveOps* curve_ops_bn254() { return CurveLaunch<ark::Bn254G1>::ops(); }
}  // namespace arkctx
