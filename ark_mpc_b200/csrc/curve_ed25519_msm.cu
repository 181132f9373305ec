// ark::Ed25519 instantiation of the point kernels, part 2: bucket-method MSM (see curve_launch.cuh).
#define ARK_CURVE_IMPL
#define ARK_CURVE_PART 2
#include "curve_launch.cuh"
namespace arkctx {
template struct CurveLaunch<ark::Ed25519>;
}  // namespace arkctx
