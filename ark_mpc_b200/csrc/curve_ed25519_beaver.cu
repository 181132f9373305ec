// ark::Ed25519 instantiation of the point kernels, part 1: point Beaver mask / recombine (see curve_launch.cuh).
#define ARK_CURVE_IMPL
#define ARK_CURVE_PART 1
#include "curve_launch.cuh"
namespace arkctx {
template struct CurveLaunch<ark::Ed25519>;
}  // namespace arkctx
