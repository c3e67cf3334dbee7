// filter_interpolation.cuh -- argument block shared by the generic and the TMA kernels.
#pragma once
#include "memc_common.cuh"

namespace memc {

struct FiArgs {
    int B, C, H, W, fs;
    View in1, flow, filt, out;  // `out` = output (fwd) or gradoutput (bwd)
    View gi1, gi2, gi3;         // bwd only
    const float* in1p;
    const float* flowp;
    const float* filtp;
    float* outp;          // fwd
    const float* goutp;   // bwd
    float* gi1p;
    float* gi2p;
    float* gi3p;
    int flags;            // MEMC_B200_* flags of the call
};

// fast path (filter_interpolation_tma.cu): returns 1 if it took the call, 0 if its layout
// preconditions do not hold (the caller then runs the generic kernels), -1 on error
int fi_forward_fast(cudaStream_t stream, const FiArgs& a);
int fi_backward_fast(cudaStream_t stream, const FiArgs& a, bool overwrite);
// C > 4, C % 4 == 0 (filter_interpolation_bwd_chunked.cu): 1 / 0 / -1 as above
int fi_backward_chunked(cudaStream_t stream, const FiArgs& a, bool overwrite);
// two images warped with ONE flow / filter in one pass (filter_interpolation_fwd_cols.cu): 1 / 0 / -1 as above
int fi_forward_cols_pair(cudaStream_t stream, const FiArgs& a, const float* in2, View v_in2, float* out2, View v_out2, int C2);
// fused FilterInterpolation pair + occlusion blend (forward, fs == 4, C == 3); a0.outp = blended output
int fi_blend_forward_fast(cudaStream_t stream, const FiArgs& a0, const FiArgs& a1, const float* occ0, View v_occ0,
                          const float* occ1, View v_occ1);

}  // namespace memc
