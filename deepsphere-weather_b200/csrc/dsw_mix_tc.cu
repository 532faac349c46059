// tcgen05 (5th-gen tensor core) channel mix — placeholder until the split-bf16 kernel lands.
#include "dsw_internal.cuh"

namespace dsw {
int launch_mix_tc(const MixArgs&, cudaStream_t) { return DSW_ERR_UNSUPPORTED; }
}  // namespace dsw
