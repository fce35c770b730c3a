// LoFTR fine stage (match type 'coarse_fine') SIMT kernels: FPN bilinear x2 upsample + lateral add,
// 5x5 fine-window gather at the coarse matches, coarse-token gather.
#pragma once
#include "tc_common.cuh"

namespace dfsfm {

// out = lateral + F.interpolate(low, scale_factor=2, mode='bilinear', align_corners=True)   (resnet_fpn.py:110-116)
// low: split-fp16 flat-halo [ (h+1)*(w+1) ][C];  lateral: fp32 flat-halo [ (2h+1)*(2w+1) ][C];  out: split-fp16, lateral's geometry.
// One thread = one output pixel x 8 channels.  Source index = dst * (in-1)/(out-1), the ATen align_corners formula.
static __global__ void upsample2x_add_kernel(const __half* __restrict__ low_hi, const __half* __restrict__ low_lo, int h, int w, int C,
                                             const float* __restrict__ lateral, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                             long long total) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    const int c8 = C / 8;
    const int cg = static_cast<int>(idx % c8);
    long long t = idx / c8;
    const int H = 2 * h, W = 2 * w;
    const int ox = static_cast<int>(t % W);
    const int oy = static_cast<int>(t / W);
    const float sy = static_cast<float>(h - 1) / static_cast<float>(H - 1), sx = static_cast<float>(w - 1) / static_cast<float>(W - 1);
    const float fy = sy * static_cast<float>(oy), fx = sx * static_cast<float>(ox);
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const int wp = w + 1, Wp = W + 1;
    auto load8 = [&](int yy, int xx, float* o) {
        const long long r = (static_cast<long long>(yy) * wp + xx) * C + cg * 8;
        const uint4 uh = *reinterpret_cast<const uint4*>(low_hi + r);
        const uint4 ul = *reinterpret_cast<const uint4*>(low_lo + r);
        const __half* a = reinterpret_cast<const __half*>(&uh);
        const __half* b = reinterpret_cast<const __half*>(&ul);
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = __half2float(a[q]) + __half2float(b[q]);
    };
    float v00[8], v01[8], v10[8], v11[8];
    load8(y0, x0, v00); load8(y0, x1, v01); load8(y1, x0, v10); load8(y1, x1, v11);
    const long long orow = (static_cast<long long>(oy) * Wp + ox) * C + cg * 8;
    const float4 l0 = *reinterpret_cast<const float4*>(lateral + orow);
    const float4 l1 = *reinterpret_cast<const float4*>(lateral + orow + 4);
    const float lat[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
    __half oh[8], ol[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        // ATen upsample_bilinear2d: h0lambda*(w0lambda*v00 + w1lambda*v01) + h1lambda*(w0lambda*v10 + w1lambda*v11)
        const float up = hy * (hx * v00[q] + lx * v01[q]) + ly * (hx * v10[q] + lx * v11[q]);
        split_f16(lat[q] + up, oh[q], ol[q]);
    }
    *reinterpret_cast<uint4*>(out_hi + orow) = *reinterpret_cast<uint4*>(oh);
    *reinterpret_cast<uint4*>(out_lo + orow) = *reinterpret_cast<uint4*>(ol);
}

// fine_preprocess.py:40-46: F.unfold(feat_f, 5x5, stride, padding 2) rows selected at the coarse matches -- restated as a
// gather: window m of image side s (0: i_ids on feat_f0, 1: j_ids on feat_f1), cell (ky,kx) = feat_f[(cy*stride + ky - 2, cx*stride + kx - 2)]
// with zero padding.  feat_f: dense fp32 [Hf*Wf][128].  out: split-fp16 [2*M*25][128] (side-major: all feat0 windows first).
static __global__ void __launch_bounds__(128) gather_windows_kernel(const float* __restrict__ f0, int Hf0, int Wf0, int wc0,
                                                                    const float* __restrict__ f1, int Hf1, int Wf1, int wc1, int stride,
                                                                    const int* __restrict__ i_ids, const int* __restrict__ j_ids, int M,
                                                                    __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
    const int wm = blockIdx.x;  // window index in [0, 2M)
    const int side = wm >= M;
    const int m = side ? wm - M : wm;
    const float* f = side ? f1 : f0;
    const int Hf = side ? Hf1 : Hf0, Wf = side ? Wf1 : Wf0, wc = side ? wc1 : wc0;
    const int id = side ? j_ids[m] : i_ids[m];
    const int cy = (id / wc) * stride, cx = (id % wc) * stride;
    const int c = threadIdx.x;  // 128 channels
    for (int cell = 0; cell < 25; ++cell) {
        const int y = cy + cell / 5 - 2, x = cx + cell % 5 - 2;
        float v = 0.f;
        if (y >= 0 && y < Hf && x >= 0 && x < Wf) v = f[(static_cast<long long>(y) * Wf + x) * 128 + c];
        __half h, l;
        split_f16(v, h, l);
        const long long o = (static_cast<long long>(wm) * 25 + cell) * 128 + c;
        out_hi[o] = h;
        out_lo[o] = l;
    }
}

// rows [m] = feat_c0[i_ids[m]], rows [M + m] = feat_c1[j_ids[m]]  -> split-fp16 [2M][256]   (fine_preprocess.py:50-51)
static __global__ void __launch_bounds__(256) gather_coarse_kernel(const float* __restrict__ c0, const float* __restrict__ c1,
                                                                   const int* __restrict__ i_ids, const int* __restrict__ j_ids, int M,
                                                                   __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
    const int r = blockIdx.x;
    const int side = r >= M;
    const int m = side ? r - M : r;
    const float* src = (side ? c1 : c0) + static_cast<long long>(side ? j_ids[m] : i_ids[m]) * 256;
    __half h, l;
    split_f16(src[threadIdx.x], h, l);
    out_hi[static_cast<long long>(r) * 256 + threadIdx.x] = h;
    out_lo[static_cast<long long>(r) * 256 + threadIdx.x] = l;
}

}  // namespace dfsfm
