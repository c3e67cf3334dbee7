// flow_projection.cuh -- argument block shared by the generic and the fast FlowProjection paths.
#pragma once
#include "memc_common.cuh"

namespace memc {

struct FpArgs {
    int B, H, W, fillhole;
    View flow, count, out;  // out = output (fwd) / gradoutput (bwd)
    View gi;                // bwd
    const float* flowp;
    float* countp;          // fwd: written; bwd: read
    float* outp;
    const float* goutp;
    float* gip;
};

// fast path (flow_projection_fast.cu): 1 = handled, 0 = layout preconditions not met (caller
// runs the generic kernels), -1 = error
int fp_forward_fast(cudaStream_t stream, const FpArgs& a, bool overwrite, bool no_zero, int variant);
// frame-by-frame fast driver (flow_projection_fast.cu): [zero fills] -> splat(b) -> average + occupancy masks per frame, then
// one mask-based fill-hole launch (O(1) per hole).  `splat` launches the caller's splat kernel for frame b (0 = ok);
// signed_counts: count is an accumulated weight that may be <= 0 / NaN where something landed (DepthFlowProjection).
// 1 = handled, 0 = layout preconditions not met, -1 = error
typedef int (*FpSplatFn)(cudaStream_t stream, const void* ctx, int b);
// extra: one more dense [B,1,H,W] plane (batch stride extra_b) that the average pass divides by count as well, or null
int fp_frames_fast(cudaStream_t stream, const FpArgs& a, bool overwrite, bool no_zero, bool signed_counts, FpSplatFn splat,
                   const void* ctx, float* extra, int64_t extra_b);
// average (+ fill-hole) over frames [b0, b0 + nb) with the generic kernels (flow_projection.cu)
int fp_average_fill(cudaStream_t stream, const FpArgs& a, int b0, int nb, bool do_average);

}  // namespace memc
