// HP-2 (multi-view refinement) SIMT kernels: fused RoIAlign gather + normalise + conv1_1, 3x3/2 max-pool on
// split-fp16 planes, bicubic merge + centre crop, window correlation / soft-argmax / reference-point search.
#pragma once
#include "tc_common.cuh"

namespace dfsfm {

constexpr int kCrop = 35;          // multiview_transform.crop_size (multiview_refinement_matching.yaml:49)
constexpr int kCropP = kCrop + 1;  // flat-halo pitch of the 35x35 patch maps

struct PatchRec {
    const float* image;  // [3][H][W] fp32, RGB in [0,1]
    int H, W;
    float y1, x1, y2, x2;  // normalised box (y1,x1,y2,x2) exactly as RoIAlign.forward builds it (roi_align.py:39-44)
};

// RoIAlign 35x35x3 bilinear gather (crop_and_resize_kernel.cu:10-82 semantics) -> ImageNet mean/std normalise
// (s2dnet.py:131-133) -> VGG conv1_1 3x3 pad 1 (3 -> 64) + bias + ReLU -> flat-halo split-fp16 [P*36*36][64].
// CTA per patch, 256 threads: thread = (channel, pixel phase).
static __global__ void __launch_bounds__(256) patch_conv11_kernel(const PatchRec* __restrict__ recs, const float* __restrict__ w /*[64][27]*/,
                                                                  const float* __restrict__ bias, __half* __restrict__ out_hi,
                                                                  __half* __restrict__ out_lo, float* __restrict__ patches_out /*optional [P][3][35][35]*/) {
    __shared__ float tile[3][kCrop + 2][kCrop + 3];
    const PatchRec r = recs[blockIdx.x];
    for (int i = threadIdx.x; i < 3 * (kCrop + 2) * (kCrop + 3); i += 256) (&tile[0][0][0])[i] = 0.f;
    __syncthreads();
    // explicit round-to-nearest mul/add/div (no FMA contraction): bit-identical to the reference's CPU op
    const float height_scale = __fdiv_rn(__fmul_rn(r.y2 - r.y1, r.H - 1), kCrop - 1);
    const float width_scale = __fdiv_rn(__fmul_rn(r.x2 - r.x1, r.W - 1), kCrop - 1);
    const float mean[3] = {0.485f, 0.456f, 0.406f};
    const float stdv[3] = {0.229f, 0.224f, 0.225f};
    for (int i = threadIdx.x; i < 3 * kCrop * kCrop; i += 256) {
        const int x = i % kCrop, y = (i / kCrop) % kCrop, d = i / (kCrop * kCrop);
        const float in_y = __fadd_rn(__fmul_rn(r.y1, r.H - 1), __fmul_rn(y, height_scale));
        const float in_x = __fadd_rn(__fmul_rn(r.x1, r.W - 1), __fmul_rn(x, width_scale));
        float v = 0.f;  // extrapolation_value
        if (!(in_y < 0 || in_y > r.H - 1 || in_x < 0 || in_x > r.W - 1)) {
            const int top = static_cast<int>(floorf(in_y)), bottom = static_cast<int>(ceilf(in_y));
            const int left = static_cast<int>(floorf(in_x)), right = static_cast<int>(ceilf(in_x));
            const float y_lerp = in_y - top, x_lerp = in_x - left;
            const float* p = r.image + static_cast<long long>(d) * r.H * r.W;
            const float tl = __ldg(p + static_cast<long long>(top) * r.W + left), tr = __ldg(p + static_cast<long long>(top) * r.W + right);
            const float bl = __ldg(p + static_cast<long long>(bottom) * r.W + left), br = __ldg(p + static_cast<long long>(bottom) * r.W + right);
            const float t = __fadd_rn(tl, __fmul_rn(tr - tl, x_lerp));
            const float b = __fadd_rn(bl, __fmul_rn(br - bl, x_lerp));
            v = __fadd_rn(t, __fmul_rn(b - t, y_lerp));
        }
        if (patches_out) patches_out[static_cast<long long>(blockIdx.x) * 3 * kCrop * kCrop + i] = v;
        tile[d][y + 1][x + 1] = (v - mean[d]) / stdv[d];
    }
    __syncthreads();
    const int c = threadIdx.x & 63, phase = threadIdx.x >> 6;
    float wr[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) wr[i] = w[c * 27 + i];
    const float b = bias[c];
    const long long base = static_cast<long long>(blockIdx.x) * kCropP * kCropP;
    // four adjacent output pixels per pass: a 3 x 6 input window per colour feeds 4 x 27 FMAs (18 shared loads instead of 4 x 27 --
    // the one-pixel version was bound by the shared-memory pipe, one broadcast load per FMA)
    constexpr int kGroups = (kCrop + 3) / 4;
    for (int g = phase; g < kCrop * kGroups; g += 4) {
        const int y = g / kGroups, x0 = (g - y * kGroups) * 4;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                float in[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) in[i] = tile[d][y + ky][x0 + i];
#pragma unroll
                for (int px = 0; px < 4; ++px)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) acc[px] = fmaf(wr[d * 9 + ky * 3 + kx], in[px + kx], acc[px]);
            }
#pragma unroll
        for (int px = 0; px < 4; ++px) {
            const int x = x0 + px;
            if (x < kCrop) {
                const float a = fmaxf(acc[px] + b, 0.f);
                __half h, l;
                split_f16(a, h, l);
                const long long o = (base + y * kCropP + x) * 64 + c;
                out_hi[o] = h;
                out_lo[o] = l;
            }
        }
    }
}

// MaxPool2d(3, stride 2, padding 1) (s2dnet.py:89-92) on flat-halo split-fp16 maps: [P][(Hi+1)*(Wi+1)][C] -> [P][(Ho+1)*(Wo+1)][C].
// One thread = one output pixel x 8 channels; the (hi,lo) pair of the arg-max input is copied (max is exact).
static __global__ void maxpool3s2_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, int Hi, int Wi, int C,
                                         __half* __restrict__ out_hi, __half* __restrict__ out_lo, int Ho, int Wo, long long total) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    const int c8 = C / 8;
    const int cg = static_cast<int>(idx % c8);
    long long t = idx / c8;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho);
    const long long p = t / Ho;
    const int Wip = Wi + 1, Wop = Wo + 1;
    const long long ibase = p * (Hi + 1) * Wip;
    float best[8];
    __half bh[8], bl[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { best[q] = -INFINITY; bh[q] = __float2half(0.f); bl[q] = __float2half(0.f); }
    for (int dy = -1; dy <= 1; ++dy) {
        const int iy = oy * 2 + dy;
        if (iy < 0 || iy >= Hi) continue;
        for (int dx = -1; dx <= 1; ++dx) {
            const int ix = ox * 2 + dx;
            if (ix < 0 || ix >= Wi) continue;
            const long long o = (ibase + static_cast<long long>(iy) * Wip + ix) * C + cg * 8;
            const uint4 uh = *reinterpret_cast<const uint4*>(in_hi + o);
            const uint4 ul = *reinterpret_cast<const uint4*>(in_lo + o);
            const __half* hh = reinterpret_cast<const __half*>(&uh);
            const __half* hl = reinterpret_cast<const __half*>(&ul);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float v = __half2float(hh[q]) + __half2float(hl[q]);
                if (v > best[q]) { best[q] = v; bh[q] = hh[q]; bl[q] = hl[q]; }
            }
        }
    }
    const long long oo = (p * (Ho + 1) * Wop + static_cast<long long>(oy) * Wop + ox) * C + cg * 8;
    *reinterpret_cast<uint4*>(out_hi + oo) = *reinterpret_cast<uint4*>(bh);
    *reinterpret_cast<uint4*>(out_lo + oo) = *reinterpret_cast<uint4*>(bl);
}

// fmap = adap0 + Upsample(size 35x35, bicubic, align_corners=True)(adap1), centre-cropped to WxW (s2dnet.py:164-171,193):
//   tokens[p][y*W+x][c] = a0[p][y*W+x][c] + sum_{i,j} wy[y][i] wx[x][j] a1[p][(iy[y][i])*pitch1 + ix[x][j]][c]
// a0: dense fp32 [P*W*W][128]; a1: fp32 flat [P*pitch1*pitch1][128] (9x9 valid at origin); tap tables from the host.
struct BicubicTab {
    int idx[15][4];
    float w[15][4];
};
static __global__ void __launch_bounds__(128) bicubic_merge_kernel(const float* __restrict__ a0, const float* __restrict__ a1, int pitch1,
                                                                   BicubicTab tab, int W, float* __restrict__ out_f32,
                                                                   __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
    const long long p = blockIdx.x;
    const int c = threadIdx.x;  // 128 channels
    const float* s = a1 + p * pitch1 * pitch1 * 128 + c;
    for (int y = 0; y < W; ++y) {
        float rowv[9];
        // separable: interpolate along y first for the 9 source columns
#pragma unroll
        for (int sx = 0; sx < 9; ++sx) {
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) v = fmaf(tab.w[y][i], s[(tab.idx[y][i] * pitch1 + sx) * 128], v);
            rowv[sx] = v;
        }
        for (int x = 0; x < W; ++x) {
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float sv = 0.f;
#pragma unroll
                for (int sx = 0; sx < 9; ++sx) sv = (tab.idx[x][j] == sx) ? rowv[sx] : sv;
                v = fmaf(tab.w[x][j], sv, v);
            }
            const long long o = (p * W * W + y * W + x) * 128 + c;
            const float r = a0[o] + v;
            out_f32[o] = r;
            __half h, l;
            split_f16(r, h, l);
            out_hi[o] = h;
            out_lo[o] = l;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// FineMatching.forward (src/MultiviewMatcher/utils/fine_matching.py:36-98): for every candidate reference position l
// (LW x LW centre window of the reference patch) and every query view n: heat = softmax(<ref[l], qry[n][r]> / sqrt(C))
// over the W*W window, expectation (x,y) on the [-1,1] grid, std = sum sqrt(clamp(var, 1e-10)); score[l] = mean over
// views of std; best l = first arg-min (centre if the reference point is not movable).  CTA per track.
struct TrackRec {
    int tok0;      // first token row of the reference patch
    int qtok0;     // first token row of query view 0; view n starts at qtok0 + n*W*W
    int n_views;   // valid query views
    int movable;
    float qx, qy;        // query_points (orig px)
    float sqx, sqy;      // scales_origin_to_fine_query (w, h)
};
struct ViewRec {
    float rx, ry;   // reference_points_coarse (orig px)
    float sx, sy;   // scales_origin_to_fine_reference (w, h)
};
constexpr int kMaxViews = 16;
constexpr int kFmThreads = 256;
constexpr int kFmLB = 4;  // candidate reference points handled together by a warp (register tile)
inline __host__ __device__ int fm_lpad(int L) { return ((L + kFmLB - 1) / kFmLB) * kFmLB; }
inline size_t fine_match_smem_bytes(int W, int LW) {
    const int WW = W * W, L = LW * LW;
    return (static_cast<size_t>(128) * fm_lpad(L) + static_cast<size_t>(WW) * 129 + static_cast<size_t>(L) * kMaxViews * 3) * sizeof(float);
}
// dynamic smem: refT[C][Lpad] + qry[WW][C+1] + res[L][kMaxViews][3]
static __global__ void __launch_bounds__(kFmThreads) fine_match_kernel(const float* __restrict__ tokens /*[T][128]*/,
                                                                       const TrackRec* __restrict__ tracks, const ViewRec* __restrict__ views,
                                                                       int Nq, int W, int LW, float* __restrict__ query_out /*[M][2]*/,
                                                                       float* __restrict__ ref_out /*[Nq][M][2]*/, float* __restrict__ std_out /*[Nq][M]*/,
                                                                       int M) {
    constexpr int C = 128;
    extern __shared__ __align__(16) float sm[];
    const int WW = W * W, L = LW * LW, Lp = fm_lpad(L);
    float* refT = sm;                    // [C][Lp]: candidate features, transposed so that 4 candidates are one float4
    float* qry = refT + C * Lp;          // [WW][C+1]
    float* res = qry + WW * (C + 1);     // [L][kMaxViews][3] = (cx, cy, std)
    __shared__ int s_best;
    const int t = blockIdx.x;
    const TrackRec tr = tracks[t];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = W / 2 - LW / 2;
    for (int i = threadIdx.x; i < Lp * C; i += kFmThreads) {
        const int l = i / C, c = i - l * C;
        float v = 0.f;
        if (l < L) {
            const int ly = l / LW, lx = l - ly * LW;
            v = tokens[(static_cast<long long>(tr.tok0) + (r0 + ly) * W + (r0 + lx)) * C + c];
        }
        refT[c * Lp + l] = v;
    }
    const float inv_sqrt_c = 1.f / sqrtf(static_cast<float>(C));
    for (int n = 0; n < tr.n_views; ++n) {
        __syncthreads();
        const float* q = tokens + (static_cast<long long>(tr.qtok0) + static_cast<long long>(n) * WW) * C;
        for (int i = threadIdx.x; i < WW * C; i += kFmThreads) {
            const int r = i / C, c = i - r * C;
            qry[r * (C + 1) + c] = q[i];
        }
        __syncthreads();
        for (int l0 = warp * kFmLB; l0 < L; l0 += (kFmThreads / 32) * kFmLB) {
            float sim[kFmLB][8];  // W*W <= 225 -> at most 8 window cells per lane
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int r = lane + 32 * k;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                if (r < WW) {
                    const float* qr = qry + r * (C + 1);
                    const float* rf = refT + l0;
#pragma unroll 8
                    for (int c = 0; c < C; ++c) {
                        const float qv = qr[c];
                        const float4 rv = *reinterpret_cast<const float4*>(rf + c * Lp);
                        a0 = fmaf(rv.x, qv, a0); a1 = fmaf(rv.y, qv, a1); a2 = fmaf(rv.z, qv, a2); a3 = fmaf(rv.w, qv, a3);
                    }
                }
                sim[0][k] = a0 * inv_sqrt_c; sim[1][k] = a1 * inv_sqrt_c; sim[2][k] = a2 * inv_sqrt_c; sim[3][k] = a3 * inv_sqrt_c;
            }
#pragma unroll
            for (int li = 0; li < kFmLB; ++li) {
                const int l = l0 + li;
                if (l >= L) break;
                float mx = -INFINITY;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (lane + 32 * k < WW) mx = fmaxf(mx, sim[li][k]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                float se = 0.f, ex = 0.f, ey = 0.f, exx = 0.f, eyy = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int r = lane + 32 * k;
                    if (r < WW) {
                        const float e = expf(sim[li][k] - mx);
                        const int ry = r / W, rx = r - ry * W;
                        const float gx = (static_cast<float>(rx) / static_cast<float>(W - 1) - 0.5f) * 2.f;
                        const float gy = (static_cast<float>(ry) / static_cast<float>(W - 1) - 0.5f) * 2.f;
                        se += e;
                        ex = fmaf(e, gx, ex);
                        ey = fmaf(e, gy, ey);
                        exx = fmaf(e, gx * gx, exx);
                        eyy = fmaf(e, gy * gy, eyy);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    se += __shfl_xor_sync(0xffffffffu, se, o);
                    ex += __shfl_xor_sync(0xffffffffu, ex, o);
                    ey += __shfl_xor_sync(0xffffffffu, ey, o);
                    exx += __shfl_xor_sync(0xffffffffu, exx, o);
                    eyy += __shfl_xor_sync(0xffffffffu, eyy, o);
                }
                if (lane == 0) {
                    const float cx = ex / se, cy = ey / se;
                    const float vx = exx / se - cx * cx, vy = eyy / se - cy * cy;
                    const float sd = sqrtf(fmaxf(vx, 1e-10f)) + sqrtf(fmaxf(vy, 1e-10f));
                    float* o = res + (l * kMaxViews + n) * 3;
                    o[0] = cx; o[1] = cy; o[2] = sd;
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = L / 2;
        if (tr.movable) {
            float bs = INFINITY;
            best = 0;
            for (int l = 0; l < L; ++l) {
                float s = 0.f;
                for (int n = 0; n < tr.n_views; ++n) s += res[(l * kMaxViews + n) * 3 + 2];
                s = s / fmaxf(static_cast<float>(tr.n_views), 1.f);
                if (s < bs) { bs = s; best = l; }
            }
        }
        s_best = best;
        const int bx = best % LW, by = best / LW;
        const float den = static_cast<float>(LW > 1 ? LW - 1 : 1);
        const float ox = (static_cast<float>(bx) / den) * 2.f - 1.f;
        const float oy = (static_cast<float>(by) / den) * 2.f - 1.f;
        query_out[t * 2 + 0] = tr.qx + ox * static_cast<float>(LW / 2) * tr.sqx;
        query_out[t * 2 + 1] = tr.qy + oy * static_cast<float>(LW / 2) * tr.sqy;
    }
    __syncthreads();
    const int best = s_best;
    for (int n = threadIdx.x; n < tr.n_views; n += kFmThreads) {
        const float* o = res + (best * kMaxViews + n) * 3;
        const ViewRec v = views[static_cast<long long>(n) * M + t];
        ref_out[(static_cast<long long>(n) * M + t) * 2 + 0] = v.rx + o[0] * static_cast<float>(W / 2) * v.sx;
        ref_out[(static_cast<long long>(n) * M + t) * 2 + 1] = v.ry + o[1] * static_cast<float>(W / 2) * v.sy;
        std_out[static_cast<long long>(n) * M + t] = o[2];
    }
}

}  // namespace dfsfm
