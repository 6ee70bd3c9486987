// placeholder until the fused tcgen05 kernel lands
#include "common.cuh"
namespace pasn {
bool sm100_supported(const pasn_dims&) { return false; }
size_t sm100_workspace_bytes(const pasn_dims&) { return 0; }
size_t sm100_packed_bytes(const pasn_dims&) { return 0; }
int sm100_pack_weights(const pasn_weights&, const pasn_dims&, void*, cudaStream_t) { return PASN_ERR_UNSUPPORTED; }
int sm100_head_forward(const void*, const pasn_weights&, const void*, const pasn_dims&, float*, float*, void*, float*,
                       float*, const pasn_push_args*, void*, size_t, cudaStream_t) { return PASN_ERR_UNSUPPORTED; }
}
