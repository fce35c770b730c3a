// Host image pipeline, device part: the PIL-LANCZOS resize of read_grayscale
// (src/dataset/utils.py:121-177: cv2 decode -> PIL.Image.resize(size, LANCZOS) on uint8 -> /255 -> float32 [1,h,w]).
//
// Pillow's 8-bit resampler (src/libImaging/Resample.c, Pillow 12.2: precompute_coeffs, normalize_coeffs_8bpc,
// ImagingResampleHorizontal_8bpc / Vertical_8bpc) is integer arithmetic once the coefficients exist: two separable passes,
// horizontal first, each   out = clip8((2^21 + sum_k in[min + k] * coef[k]) >> 22)   with the intermediate image rounded to
// uint8.  The host builds the fixed-point coefficient tables exactly as Pillow does (libm sin; image_pipeline.py) -- the
// kernels below do the integer part, bit-exactly, and fuse the /255 float conversion of grayscale2tensor into the last pass.
// HBM-bound byte work: h*w bytes in, 4*out_h*out_w bytes out, one h*out_w byte intermediate.
#include "../../include/dfsfm_b200.h"
#include "engine_common.h"

namespace dfsfm {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Resample.c: PRECISION_BITS

__device__ __forceinline__ int clip8(int acc) {
    const int v = acc >> kPrecisionBits;  // arithmetic shift, as the C code's
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}
template <bool kFinal>
__device__ __forceinline__ void put(int v, long long idx, uint8_t* out_u8, float* out_f32) {
    if (kFinal) out_f32[idx] = __fdiv_rn(static_cast<float>(v), 255.f);  // grayscale2tensor: float32(image) / 255.
    else out_u8[idx] = static_cast<uint8_t>(v);
}

// one thread per (row y, output column xx)
template <bool kFinal>
static __global__ void __launch_bounds__(128) lanczos_h_kernel(const uint8_t* __restrict__ in, int W_in, long long ld, const int2* __restrict__ bounds,
                                                               const int* __restrict__ coef, int ksize, int out_w, uint8_t* __restrict__ out_u8,
                                                               float* __restrict__ out_f32) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (xx >= out_w) return;
    const int2 b = bounds[xx];
    const uint8_t* line = in + static_cast<long long>(y) * ld + b.x;
    const int* k = coef + static_cast<long long>(xx) * ksize;
    int acc = 1 << (kPrecisionBits - 1);
    for (int x = 0; x < b.y; ++x) acc += static_cast<int>(line[x]) * k[x];
    put<kFinal>(clip8(acc), static_cast<long long>(y) * out_w + xx, out_u8, out_f32);
}
// one thread per (output row yy, column xx); `in` has row pitch ld and W columns
template <bool kFinal>
static __global__ void __launch_bounds__(128) lanczos_v_kernel(const uint8_t* __restrict__ in, long long ld, int W, const int2* __restrict__ bounds,
                                                               const int* __restrict__ coef, int ksize, uint8_t* __restrict__ out_u8,
                                                               float* __restrict__ out_f32) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y;
    if (xx >= W) return;
    const int2 b = bounds[yy];
    const int* k = coef + static_cast<long long>(yy) * ksize;
    int acc = 1 << (kPrecisionBits - 1);
    for (int y = 0; y < b.y; ++y) acc += static_cast<int>(in[static_cast<long long>(b.x + y) * ld + xx]) * k[y];
    put<kFinal>(clip8(acc), static_cast<long long>(yy) * W + xx, out_u8, out_f32);
}
static __global__ void u8_to_unit_float_kernel(const uint8_t* __restrict__ in, long long ld, int W, int H, float* __restrict__ out) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<long long>(W) * H) return;
    const int y = static_cast<int>(i / W), x = static_cast<int>(i % W);
    out[i] = __fdiv_rn(static_cast<float>(in[static_cast<long long>(y) * ld + x]), 255.f);
}

}  // namespace dfsfm

extern "C" int dfsfm_resize_lanczos_gray(const uint8_t* img_dev, int h, int w, int64_t ld, const int32_t* xbounds_dev, const int32_t* xcoef_dev,
                                         int xksize, const int32_t* ybounds_dev, const int32_t* ycoef_dev, int yksize, int out_h, int out_w,
                                         uint8_t* tmp_dev, float* out_dev, void* stream) {
    using namespace dfsfm;
    return guard([&] {
        DFSFM_CHECK(h >= 1 && w >= 1 && out_h >= 1 && out_w >= 1 && ld >= w, "bad image geometry");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const bool need_h = out_w != w, need_v = out_h != h;  // Pillow skips a pass whose size does not change (ImagingResample)
        DFSFM_CHECK(!need_h || (xbounds_dev && xcoef_dev && xksize > 0), "horizontal coefficients missing");
        DFSFM_CHECK(!need_v || (ybounds_dev && ycoef_dev && yksize > 0), "vertical coefficients missing");
        DFSFM_CHECK(!(need_h && need_v) || tmp_dev, "two passes need the h x out_w intermediate");
        const int2* xb = reinterpret_cast<const int2*>(xbounds_dev);
        const int2* yb = reinterpret_cast<const int2*>(ybounds_dev);
        const dim3 blk(128);
        if (need_h && need_v) {
            { LaunchScope ls("resize_h", st);
              lanczos_h_kernel<false><<<dim3((out_w + 127) / 128, h), blk, 0, st>>>(img_dev, w, ld, xb, xcoef_dev, xksize, out_w, tmp_dev, nullptr); }
            { LaunchScope ls("resize_v", st);
              lanczos_v_kernel<true><<<dim3((out_w + 127) / 128, out_h), blk, 0, st>>>(tmp_dev, out_w, out_w, yb, ycoef_dev, yksize, nullptr, out_dev); }
        } else if (need_h) {
            LaunchScope ls("resize_h", st);
            lanczos_h_kernel<true><<<dim3((out_w + 127) / 128, h), blk, 0, st>>>(img_dev, w, ld, xb, xcoef_dev, xksize, out_w, nullptr, out_dev);
        } else if (need_v) {
            LaunchScope ls("resize_v", st);
            lanczos_v_kernel<true><<<dim3((w + 127) / 128, out_h), blk, 0, st>>>(img_dev, ld, w, yb, ycoef_dev, yksize, nullptr, out_dev);
        } else {
            LaunchScope ls("resize_cvt", st);
            const long long n = static_cast<long long>(w) * h;
            u8_to_unit_float_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(img_dev, ld, w, h, out_dev);
        }
        DFSFM_CUDA(cudaGetLastError());
    });
}
