// cartpole: kernel instantiations (Philox / MT19937 x RewardNoise x Logging accumulators) and host path.
#include <cstring>

#include "bsb_dispatch.cuh"

namespace bsb {
int run_cartpole(bsb_env* e, const LaunchArgs& a, cudaStream_t stream) { return run_family<Cartpole>(e, a, stream); }
}  // namespace bsb
