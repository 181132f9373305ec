// ark::Ed25519 instantiation of the point kernels, part 3: mul_auth, share_add_public, mac_check (see curve_launch.cuh).
#define ARK_CURVE_IMPL
#define ARK_CURVE_PART 3
#include "curve_launch.cuh"
namespace arkctx {
template struct CurveLaunch<ark::Ed25519>;
}  // namespace arkctx
