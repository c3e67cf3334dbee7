// separable_conv.cu -- per-pixel separable fs x fs local convolution (SeparableConv).
//
// Semantics: reference my_package/src/my_lib_kernel.cu:285-337 (forward), :339-389
// (backward); CPU twin my_lib.c:250-439.
//   out[b,c,h,w] = sum_{y,x<fs} in1[b,c,h+y,w+x] * vert[b,y,h,w] * horiz[b,x,h,w]
// over the valid region Ho x Wo = (H-fs+1) x (W-fs+1).
//
// Backward here needs NO atomics (the legacy kernel issues 3*fs^2*C per pixel):
//   gi2 / gi3 are sums at the thread's own pixel -> registers, one write;
//   gi1 is a scatter with a FIXED footprint, so it is rewritten as the equivalent gather
//   gi1[c,Y,X] = sum_{y,x} go[c,Y-y,X-x] * vert[y,Y-y,X-x] * horiz[x,Y-y,X-x]
//   (terms whose (Y-y, X-x) fall outside the output extent dropped) -- deterministic.
#include "memc_common.cuh"

namespace memc {

struct ScArgs {
    int B, C, H, W, fs;  // H, W = input extent
    View in1, vert, horiz, out;  // out = output (fwd) / gradoutput (bwd)
    View gi1, gi2, gi3;
    const float* in1p;
    const float* vertp;
    const float* horizp;
    float* outp;
    const float* goutp;
    float* gi1p;
    float* gi2p;
    float* gi3p;
};

constexpr int BX = 32, BY = 8;

// Two output pixels per thread (rows h and h + BY): twice the independent loads in flight.
// FS = 4 (the size every shipped model uses): loops unrolled, the 2 x 4 filter values of a pixel
// are loaded once into registers instead of once per channel and tap.
template <int FS>
__global__ void __launch_bounds__(BX* BY) sc_fwd_kernel(const ScArgs p) {
    constexpr int MAXFS = FS > 0 ? FS : 1;
    const int w = blockIdx.x * BX + threadIdx.x;
    const int h0 = blockIdx.y * (2 * BY) + threadIdx.y;
    const int b = blockIdx.z;
    const int fs = FS > 0 ? FS : p.fs, Ho = p.H - fs + 1, Wo = p.W - fs + 1;
    if (w >= Wo || h0 >= Ho) return;
    const int h1 = min(h0 + BY, Ho - 1);          // second row (clamped: recomputed, not stored, if out of range)
    const bool two = h0 + BY < Ho;
    const float* vp0 = p.vertp + b * p.vert.b + h0 * p.vert.h + w;
    const float* hp0 = p.horizp + b * p.horiz.b + h0 * p.horiz.h + w;
    const float* vp1 = p.vertp + b * p.vert.b + h1 * p.vert.h + w;
    const float* hp1 = p.horizp + b * p.horiz.b + h1 * p.horiz.h + w;
    const float* img0 = p.in1p + b * p.in1.b + h0 * p.in1.h + w;
    const float* img1 = p.in1p + b * p.in1.b + h1 * p.in1.h + w;
    float* ob0 = p.outp + b * p.out.b + h0 * p.out.h + w;
    float* ob1 = p.outp + b * p.out.b + h1 * p.out.h + w;
    float v0[MAXFS], v1[MAXFS], z0[MAXFS], z1[MAXFS];
    if (FS > 0) {
#pragma unroll
        for (int k = 0; k < FS; ++k) {
            v0[k] = ldg_stream(vp0 + k * p.vert.c); v1[k] = ldg_stream(vp1 + k * p.vert.c);
            z0[k] = ldg_stream(hp0 + k * p.horiz.c); z1[k] = ldg_stream(hp1 + k * p.horiz.c);
        }
    }
    for (int c = 0; c < p.C; ++c, img0 += p.in1.c, img1 += p.in1.c) {
        float acc0 = 0.f, acc1 = 0.f;
        if (FS > 0) {
#pragma unroll
            for (int y = 0; y < FS; ++y) {
                const float* row0 = img0 + y * p.in1.h;
                const float* row1 = img1 + y * p.in1.h;
#pragma unroll
                for (int x = 0; x < FS; ++x) {  // (t1*t2)*t3 as the reference
                    acc0 += __ldg(row0 + x) * v0[y] * z0[x];
                    acc1 += __ldg(row1 + x) * v1[y] * z1[x];
                }
            }
        } else {
            for (int y = 0; y < fs; ++y) {
                const float vy0 = __ldg(vp0 + y * p.vert.c), vy1 = __ldg(vp1 + y * p.vert.c);
                const float* row0 = img0 + y * p.in1.h;
                const float* row1 = img1 + y * p.in1.h;
                for (int x = 0; x < fs; ++x) {
                    acc0 += __ldg(row0 + x) * vy0 * __ldg(hp0 + x * p.horiz.c);
                    acc1 += __ldg(row1 + x) * vy1 * __ldg(hp1 + x * p.horiz.c);
                }
            }
        }
        stg_stream(ob0 + c * p.out.c, acc0);
        if (two) stg_stream(ob1 + c * p.out.c, acc1);
    }
}

// filter gradients: own pixel, registers.  FS = 4: one pass over the 16 taps of a channel feeds
// all eight accumulators (each accumulator still sums in the generic kernel's order).
template <bool OVERWRITE, int FS>
__global__ void __launch_bounds__(BX* BY) sc_bwd_filters_kernel(const ScArgs p) {
    const int w = blockIdx.x * BX + threadIdx.x;
    const int h = blockIdx.y * BY + threadIdx.y;
    const int b = blockIdx.z;
    const int fs = FS > 0 ? FS : p.fs, Ho = p.H - fs + 1, Wo = p.W - fs + 1;
    if (w >= Wo || h >= Ho) return;
    const float* vp = p.vertp + b * p.vert.b + h * p.vert.h + w;
    const float* hp = p.horizp + b * p.horiz.b + h * p.horiz.h + w;
    const float* img0 = p.in1p + b * p.in1.b + h * p.in1.h + w;
    const float* go = p.goutp + b * p.out.b + h * p.out.h + w;
    float* g2 = p.gi2p + b * p.gi2.b + h * p.gi2.h + w;
    float* g3 = p.gi3p + b * p.gi3.b + h * p.gi3.h + w;
    if (FS > 0) {
        constexpr int N = FS > 0 ? FS : 1;
        float v[N], z[N], a2[N], a3[N];
#pragma unroll
        for (int k = 0; k < N; ++k) {
            v[k] = ldg_stream(vp + k * p.vert.c);
            z[k] = ldg_stream(hp + k * p.horiz.c);
            a2[k] = a3[k] = 0.f;
        }
        for (int c = 0; c < p.C; ++c) {
            const float gov = ldg_stream(go + c * p.out.c);
            float t[N][N];
#pragma unroll
            for (int y = 0; y < N; ++y)
#pragma unroll
                for (int x = 0; x < N; ++x) t[y][x] = __ldg(img0 + c * p.in1.c + y * p.in1.h + x);
#pragma unroll
            for (int y = 0; y < N; ++y)  // d/d vert[y] = sum_c go_c sum_x in1[c,h+y,w+x] horiz[x]
#pragma unroll
                for (int x = 0; x < N; ++x) a2[y] += gov * t[y][x] * z[x];
#pragma unroll
            for (int x = 0; x < N; ++x)  // d/d horiz[x]
#pragma unroll
                for (int y = 0; y < N; ++y) a3[x] += gov * t[y][x] * v[y];
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {
            if (OVERWRITE) { stg_stream(g2 + k * p.gi2.c, a2[k]); stg_stream(g3 + k * p.gi3.c, a3[k]); }
            else { g2[k * p.gi2.c] += a2[k]; g3[k * p.gi3.c] += a3[k]; }
        }
        return;
    }
    for (int y = 0; y < fs; ++y) {  // d/d vert[y] = sum_c go_c sum_x in1[c,h+y,w+x] horiz[x]
        float acc = 0.f;
        for (int c = 0; c < p.C; ++c) {
            const float gov = __ldg(go + c * p.out.c);
            const float* row = img0 + c * p.in1.c + y * p.in1.h;
            for (int x = 0; x < fs; ++x) acc += gov * __ldg(row + x) * __ldg(hp + x * p.horiz.c);
        }
        if (OVERWRITE) stg_stream(g2 + y * p.gi2.c, acc);
        else g2[y * p.gi2.c] += acc;
    }
    for (int x = 0; x < fs; ++x) {  // d/d horiz[x]
        float acc = 0.f;
        for (int c = 0; c < p.C; ++c) {
            const float gov = __ldg(go + c * p.out.c);
            const float* col = img0 + c * p.in1.c + x;
            for (int y = 0; y < fs; ++y) acc += gov * __ldg(col + y * p.in1.h) * __ldg(vp + y * p.vert.c);
        }
        if (OVERWRITE) stg_stream(g3 + x * p.gi3.c, acc);
        else g3[x * p.gi3.c] += acc;
    }
}

// image gradient as a gather over the INPUT extent.  FS = 4: unrolled; the product
// vert[y] * horiz[x] of a source pixel cannot be hoisted (the reference multiplies
// (go * vert) * horiz), but its two factors are loaded once for all channels.
template <bool OVERWRITE, int FS>
__global__ void __launch_bounds__(BX* BY) sc_bwd_image_kernel(const ScArgs p) {
    const int X = blockIdx.x * BX + threadIdx.x;
    const int Y = blockIdx.y * BY + threadIdx.y;
    const int b = blockIdx.z;
    if (X >= p.W || Y >= p.H) return;
    const int fs = FS > 0 ? FS : p.fs, Ho = p.H - fs + 1, Wo = p.W - fs + 1;
    const int y0 = max(0, Y - Ho + 1), y1 = min(fs - 1, Y);  // 0 <= Y-y <= Ho-1
    const int x0 = max(0, X - Wo + 1), x1 = min(fs - 1, X);
    const float* vb = p.vertp + b * p.vert.b;
    const float* hb = p.horizp + b * p.horiz.b;
    const float* gob = p.goutp + b * p.out.b;
    float* g1 = p.gi1p + b * p.gi1.b + Y * p.gi1.h + X;
    if (FS > 0 && p.C <= 4) {
        constexpr int N = FS > 0 ? FS : 1;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        // base pointers at source pixel (Y, X); tap (y, x) sits y rows up and x columns left
        const float* vq = vb + (int64_t)Y * p.vert.h + X;
        const float* zq = hb + (int64_t)Y * p.horiz.h + X;
        const float* gq = gob + (int64_t)Y * p.out.h + X;
        const int vh = (int)p.vert.h, zh = (int)p.horiz.h, gh = (int)p.out.h;
        const int vc = (int)p.vert.c, zc = (int)p.horiz.c, gc = (int)p.out.c;  // plane strides fit 31 bits per frame
        const bool interior = y0 == 0 && y1 == N - 1 && x0 == 0 && x1 == N - 1;
        if (interior) {
#pragma unroll
            for (int y = 0; y < N; ++y)
#pragma unroll
                for (int x = 0; x < N; ++x) {
                    const float vv = __ldg(vq + y * vc - y * vh - x);
                    const float zz = __ldg(zq + x * zc - y * zh - x);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < p.C) acc[c] += __ldg(gq + c * gc - y * gh - x) * vv * zz;
                }
        } else {
#pragma unroll
            for (int y = 0; y < N; ++y)
#pragma unroll
                for (int x = 0; x < N; ++x) {
                    if (y < y0 || y > y1 || x < x0 || x > x1) continue;
                    const float vv = __ldg(vq + y * vc - y * vh - x);
                    const float zz = __ldg(zq + x * zc - y * zh - x);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < p.C) acc[c] += __ldg(gq + c * gc - y * gh - x) * vv * zz;
                }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < p.C) {
                if (OVERWRITE) stg_stream(g1 + c * p.gi1.c, acc[c]);
                else g1[c * p.gi1.c] += acc[c];
            }
        return;
    }
    for (int c = 0; c < p.C; ++c) {
        float acc = 0.f;
        for (int y = y0; y <= y1; ++y)
            for (int x = x0; x <= x1; ++x) {
                const int hh = Y - y, ww = X - x;
                acc += __ldg(gob + c * p.out.c + hh * p.out.h + ww) * __ldg(vb + y * p.vert.c + hh * p.vert.h + ww) *
                       __ldg(hb + x * p.horiz.c + hh * p.horiz.h + ww);
            }
        if (OVERWRITE) stg_stream(g1 + c * p.gi1.c, acc);
        else g1[c * p.gi1.c] += acc;
    }
}

static int sc_forward(cudaStream_t stream, const ScArgs& a, int flags) {
    (void)flags;
    if (a.fs <= 0) return -1;
    const int Ho = a.H - a.fs + 1, Wo = a.W - a.fs + 1;
    if (a.B <= 0 || a.C <= 0 || Ho <= 0 || Wo <= 0) return 0;
    DeviceGuard guard(a.in1p);
    if (!guard.ok) return -1;
    dim3 block(BX, BY, 1), grid((Wo + BX - 1) / BX, (Ho + 2 * BY - 1) / (2 * BY), a.B);
    if (a.fs == 4) sc_fwd_kernel<4><<<grid, block, 0, stream>>>(a);
    else sc_fwd_kernel<0><<<grid, block, 0, stream>>>(a);
    count_launch();
    return check_launch("SeparableConv forward");
}

static int sc_backward(cudaStream_t stream, const ScArgs& a, int flags) {
    if (a.fs <= 0) return -1;
    const int Ho = a.H - a.fs + 1, Wo = a.W - a.fs + 1;
    if (a.B <= 0 || a.C <= 0 || Ho <= 0 || Wo <= 0) return 0;
    DeviceGuard guard(a.in1p);
    if (!guard.ok) return -1;
    const bool ow = (flags & MEMC_B200_OVERWRITE) != 0;
    dim3 block(BX, BY, 1);
    dim3 gout((Wo + BX - 1) / BX, (Ho + BY - 1) / BY, a.B), gin((a.W + BX - 1) / BX, (a.H + BY - 1) / BY, a.B);
    if (a.fs == 4) {
        if (ow) {
            sc_bwd_filters_kernel<true, 4><<<gout, block, 0, stream>>>(a);
            sc_bwd_image_kernel<true, 4><<<gin, block, 0, stream>>>(a);
        } else {
            sc_bwd_filters_kernel<false, 4><<<gout, block, 0, stream>>>(a);
            sc_bwd_image_kernel<false, 4><<<gin, block, 0, stream>>>(a);
        }
    } else if (ow) {
        sc_bwd_filters_kernel<true, 0><<<gout, block, 0, stream>>>(a);
        sc_bwd_image_kernel<true, 0><<<gin, block, 0, stream>>>(a);
    } else {
        sc_bwd_filters_kernel<false, 0><<<gout, block, 0, stream>>>(a);
        sc_bwd_image_kernel<false, 0><<<gin, block, 0, stream>>>(a);
    }
    count_launch(2);
    return check_launch("SeparableConv backward");
}

}  // namespace memc

using namespace memc;

extern "C" int memc_b200_separable_conv_forward(
    memc_stream_t stream, int batch, int channel, int h, int w, int filter_size,
    memc_strides s_in1, memc_strides s_vert, memc_strides s_horiz, memc_strides s_out,
    const float* input1, const float* vertical, const float* horizontal, float* output, int flags) {
    ScArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w; a.fs = filter_size;
    a.in1 = mk_view(s_in1); a.vert = mk_view(s_vert); a.horiz = mk_view(s_horiz); a.out = mk_view(s_out);
    a.in1p = input1; a.vertp = vertical; a.horizp = horizontal; a.outp = output;
    return sc_forward(stream, a, flags);
}

extern "C" int memc_b200_separable_conv_backward(
    memc_stream_t stream, int batch, int channel, int h, int w, int filter_size,
    memc_strides s_in1, memc_strides s_vert, memc_strides s_horiz, memc_strides s_gout,
    memc_strides s_gi1, memc_strides s_gi2, memc_strides s_gi3,
    const float* input1, const float* vertical, const float* horizontal, const float* gradoutput,
    float* gradinput1, float* gradinput2, float* gradinput3, int flags) {
    ScArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w; a.fs = filter_size;
    a.in1 = mk_view(s_in1); a.vert = mk_view(s_vert); a.horiz = mk_view(s_horiz); a.out = mk_view(s_gout);
    a.gi1 = mk_view(s_gi1); a.gi2 = mk_view(s_gi2); a.gi3 = mk_view(s_gi3);
    a.in1p = input1; a.vertp = vertical; a.horizp = horizontal; a.goutp = gradoutput;
    a.gi1p = gradinput1; a.gi2p = gradinput2; a.gi3p = gradinput3;
    return sc_backward(stream, a, flags);
}

// Reference-named launchers (my_lib_kernel.h:37-64): gradients use their input's strides
// (my_lib_kernel.cu:376-381), gradoutput uses the "output" strides.
extern "C" int SeparableConvLayer_gpu_forward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel,
    const int batch, const int filter_size,
    const int i1b, const int i1c, const int i1h, const int i1w,
    const int i2b, const int i2c, const int i2h, const int i2w,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int ob, const int oc, const int oh, const int ow_,
    const float* input1, const float* input2, const float* input3, float* output) {
    (void)nElement;
    if (i1w != 1 || i2w != 1 || i3w != 1 || ow_ != 1) return -1;
    ScArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w; a.fs = filter_size;
    a.in1 = mk_view(i1b, i1c, i1h); a.vert = mk_view(i2b, i2c, i2h); a.horiz = mk_view(i3b, i3c, i3h);
    a.out = mk_view(ob, oc, oh);
    a.in1p = input1; a.vertp = input2; a.horizp = input3; a.outp = output;
    return sc_forward(stream, a, 0);
}

extern "C" int SeparableConvLayer_gpu_backward_kernel(
    memc_stream_t stream, const int nElement, const int w, const int h, const int channel,
    const int batch, const int filter_size,
    const int i1b, const int i1c, const int i1h, const int i1w,
    const int i2b, const int i2c, const int i2h, const int i2w,
    const int i3b, const int i3c, const int i3h, const int i3w,
    const int ob, const int oc, const int oh, const int ow_,
    const float* input1, const float* input2, const float* input3, const float* gradoutput,
    float* gradinput1, float* gradinput2, float* gradinput3) {
    (void)nElement;
    if (i1w != 1 || i2w != 1 || i3w != 1 || ow_ != 1) return -1;
    ScArgs a{};
    a.B = batch; a.C = channel; a.H = h; a.W = w; a.fs = filter_size;
    a.in1 = mk_view(i1b, i1c, i1h); a.vert = mk_view(i2b, i2c, i2h); a.horiz = mk_view(i3b, i3c, i3h);
    a.out = mk_view(ob, oc, oh);
    a.gi1 = a.in1; a.gi2 = a.vert; a.gi3 = a.horiz;
    a.in1p = input1; a.vertp = input2; a.horizp = input3; a.goutp = gradoutput;
    a.gi1p = gradinput1; a.gi2p = gradinput2; a.gi3p = gradinput3;
    return sc_backward(stream, a, 0);
}
