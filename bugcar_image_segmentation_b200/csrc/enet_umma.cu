// tcgen05 path of the regular / dilated / asymmetric bottlenecks (placeholder until the
// fused kernel lands: reports "not supported" so the CUDA-core kernels run).
#include "internal.h"

namespace bc {

bool umma_supported(const Bottleneck&) { return false; }
void umma_pack(Bottleneck&, const std::vector<float>&, const std::vector<float>&, const std::vector<float>&) {}
void launch_umma_bottleneck(const bf16*, bf16*, const Bottleneck&, int, int, int, int, cudaStream_t) {}

}  // namespace bc
