// Curve25519 (twisted Edwards) instantiation of the point kernels.
#define ARK_CURVE_IMPL
#include "curve_launch.cuh"
namespace arkctx {
const CurveOps* curve_ops_ed25519() { return CurveLaunch<ark::Ed25519>::ops(); }
}  // namespace arkctx
