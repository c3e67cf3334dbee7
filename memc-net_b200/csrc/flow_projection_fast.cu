// flow_projection_fast.cu -- fused FlowProjection forward (placeholder: "not applicable").
#include "memc_common.cuh"
namespace memc {
struct FpArgs;
int fp_forward_fast(cudaStream_t, const FpArgs&, bool) { return 0; }
}  // namespace memc
