// Library-wide state (last error, launch counter), the standalone RoIAlign op and the engine test hook.
#include "../../include/dfsfm_b200.h"
#include <mutex>

#include "engine_common.h"

namespace dfsfm {
constexpr int kTlCtas = 148;
static unsigned long long* g_tl_buf = nullptr;
static int g_tl_cap = 0, g_tl_used = 0;
static std::vector<int32_t> g_tl_info;
unsigned long long* timeline_next_launch(int grid_ctas, int tiles) {
    if (g_tl_used >= g_tl_cap || grid_ctas > kTlCtas) return nullptr;
    g_tl_info.push_back(grid_ctas);
    g_tl_info.push_back(tiles);
    return g_tl_buf + static_cast<size_t>(g_tl_used++) * kTlCtas * 16;
}

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
std::atomic<long long>& launch_counter() {
    static std::atomic<long long> c{0};
    return c;
}

static int g_engine = -1;
int engine_version() {
    if (g_engine < 0) {
        const char* e = getenv("DFSFM_ENGINE");
        g_engine = (e && e[0] == '1') ? 1 : 2;
    }
    return g_engine;
}
void set_engine_version(int v) { g_engine = (v == 1) ? 1 : 2; }

// ------------------------------------------------------------------------------------------------ profiler
namespace {
struct ProfRec { std::string label; cudaEvent_t a, b; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::mutex g_prof_mu;                      // several host threads may drive engines concurrently (one engine handle per thread)
thread_local size_t g_prof_open = 0;       // index of this thread's open record
}  // namespace
bool profiling_enabled() { return g_prof_on; }
void prof_begin(const char* label, cudaStream_t st) {
    ProfRec r;
    r.label = label;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_open = g_prof.size();
    g_prof.push_back(r);
}
void prof_end(cudaStream_t st) {
    cudaEvent_t b;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        b = g_prof[g_prof_open].b;
    }
    cudaEventRecord(b, st);
}

// ---------------------------------------------------------------------------------------------------------
// TensorFlow-style crop_and_resize forward (the reference's L0 native op):
//   third_party/RoIAlign.pytorch/roi_align/src/cuda/crop_and_resize_kernel.cu:10-82 (semantics),
//   roi_align/src/crop_and_resize.cpp:7-113 (CPU twin the oracle is checked against).
// One CTA per box; consecutive threads take consecutive output elements (x fastest), so the NCHW crop is written fully
// coalesced and the four bilinear taps of neighbouring threads fall into the same cache lines of the source rows.
__global__ void __launch_bounds__(256) crop_and_resize_kernel(const float* __restrict__ image, int batch, int depth, int ih, int iw,
                                                              const float* __restrict__ boxes, const int* __restrict__ box_index,
                                                              float extrapolation, int ch, int cw, float* __restrict__ crops) {
    const int b = blockIdx.x;
    const float y1 = boxes[b * 4 + 0], x1 = boxes[b * 4 + 1], y2 = boxes[b * 4 + 2], x2 = boxes[b * 4 + 3];
    const int b_in = box_index[b];
    const long long crop_elems = static_cast<long long>(depth) * ch * cw;
    float* out = crops + b * crop_elems;
    if (b_in < 0 || b_in >= batch) {  // the reference silently skips such boxes on the GPU (kernel.cu:36-39)
        return;
    }
    const float height_scale = (ch > 1) ? __fdiv_rn(__fmul_rn(y2 - y1, ih - 1), ch - 1) : 0.f;
    const float width_scale = (cw > 1) ? __fdiv_rn(__fmul_rn(x2 - x1, iw - 1), cw - 1) : 0.f;
    const float* img = image + static_cast<long long>(b_in) * depth * ih * iw;
    for (long long idx = threadIdx.x; idx < crop_elems; idx += blockDim.x) {
        const int x = static_cast<int>(idx % cw);
        const int y = static_cast<int>((idx / cw) % ch);
        const int d = static_cast<int>(idx / (static_cast<long long>(cw) * ch));
        // explicit round-to-nearest mul/add (no FMA contraction): bit-identical to the reference's CPU op (gcc, SSE)
        const float in_y = (ch > 1) ? __fadd_rn(__fmul_rn(y1, ih - 1), __fmul_rn(y, height_scale)) : 0.5f * (y1 + y2) * (ih - 1);
        const float in_x = (cw > 1) ? __fadd_rn(__fmul_rn(x1, iw - 1), __fmul_rn(x, width_scale)) : 0.5f * (x1 + x2) * (iw - 1);
        float v = extrapolation;
        if (!(in_y < 0 || in_y > ih - 1 || in_x < 0 || in_x > iw - 1)) {
            const int top = static_cast<int>(floorf(in_y)), bottom = static_cast<int>(ceilf(in_y));
            const int left = static_cast<int>(floorf(in_x)), right = static_cast<int>(ceilf(in_x));
            const float y_lerp = in_y - top, x_lerp = in_x - left;
            const float* p = img + static_cast<long long>(d) * ih * iw;
            const float tl = __ldg(p + static_cast<long long>(top) * iw + left);
            const float tr = __ldg(p + static_cast<long long>(top) * iw + right);
            const float bl = __ldg(p + static_cast<long long>(bottom) * iw + left);
            const float br = __ldg(p + static_cast<long long>(bottom) * iw + right);
            const float t = __fadd_rn(tl, __fmul_rn(tr - tl, x_lerp));
            const float bt = __fadd_rn(bl, __fmul_rn(br - bl, x_lerp));
            v = __fadd_rn(t, __fmul_rn(bt - t, y_lerp));
        }
        out[idx] = v;
    }
}

}  // namespace dfsfm

extern "C" {

const char* dfsfm_last_error(void) { return dfsfm::g_last_error.c_str(); }
int dfsfm_version(void) { return 1; }
int64_t dfsfm_launch_count(void) { return dfsfm::launch_counter().load(); }
void dfsfm_thread_set_pdl(int mode) { dfsfm::pdl_thread_override() = mode < 0 ? -1 : (mode ? 1 : 0); }

void dfsfm_set_engine(int version) { dfsfm::set_engine_version(version); }
int dfsfm_get_engine(void) { return dfsfm::engine_version(); }

int dfsfm_debug_timeline_arm(int max_launches) {
    using namespace dfsfm;
    return guard([&] {
        if (g_tl_buf) { DFSFM_CUDA(cudaFree(g_tl_buf)); g_tl_buf = nullptr; }
        g_tl_cap = g_tl_used = 0;
        g_tl_info.clear();
        if (max_launches > 0) {
            const size_t bytes = static_cast<size_t>(max_launches) * kTlCtas * 16 * sizeof(unsigned long long);
            DFSFM_CUDA(cudaMalloc(&g_tl_buf, bytes));
            DFSFM_CUDA(cudaMemset(g_tl_buf, 0, bytes));
            g_tl_cap = max_launches;
        }
    });
}
int dfsfm_debug_timeline_read(uint64_t* stamps, int32_t* info, int max_launches) {
    using namespace dfsfm;
    int n = 0;
    const int rc = guard([&] {
        DFSFM_CUDA(cudaDeviceSynchronize());
        n = g_tl_used < max_launches ? g_tl_used : max_launches;
        if (n > 0) {
            DFSFM_CUDA(cudaMemcpy(stamps, g_tl_buf, static_cast<size_t>(n) * kTlCtas * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            for (int i = 0; i < 2 * n; ++i) info[i] = g_tl_info[i];
        }
    });
    return rc ? -1 : n;
}

void dfsfm_profile_enable(int on) {
    dfsfm::g_prof_on = on != 0;
    if (on) {
        for (auto& r : dfsfm::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        dfsfm::g_prof.clear();
    }
}
// Writes "label count total_ms\n" lines (labels aggregated) into buf; returns the number of bytes needed.
int dfsfm_profile_report(char* buf, int cap) {
    cudaDeviceSynchronize();
    std::map<std::string, std::pair<long long, double>> agg;
    for (auto& r : dfsfm::g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto& e = agg[r.label];
            e.first += 1;
            e.second += ms;
        }
    }
    std::string out;
    for (auto& kv : agg) out += kv.first + " " + std::to_string(kv.second.first) + " " + std::to_string(kv.second.second) + "\n";
    if (buf && cap > 0) {
        const int n = static_cast<int>(out.size()) < cap - 1 ? static_cast<int>(out.size()) : cap - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return static_cast<int>(out.size()) + 1;
}

int dfsfm_crop_and_resize_forward(const float* image_dev, int batch, int depth, int image_h, int image_w, const float* boxes_dev,
                                  const int32_t* box_index_dev, int num_boxes, float extrapolation_value, int crop_h, int crop_w,
                                  float* crops_dev, void* stream) {
    return dfsfm::guard([&] {
        if (num_boxes <= 0) return;
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        // crops.zero_() of the reference (crop_and_resize_gpu.cpp:42-43): boxes with an invalid index keep zeros
        DFSFM_CUDA(cudaMemsetAsync(crops_dev, 0, static_cast<size_t>(num_boxes) * depth * crop_h * crop_w * sizeof(float), st));
        dfsfm::LaunchScope ls("roialign", st);
        dfsfm::crop_and_resize_kernel<<<num_boxes, 256, 0, st>>>(image_dev, batch, depth, image_h, image_w, boxes_dev, box_index_dev,
                                                                extrapolation_value, crop_h, crop_w, crops_dev);
        DFSFM_CUDA(cudaGetLastError());
    });
}

int dfsfm_debug_gemm(const void* a_dev, int64_t a_rows, int C, const void* w_dev, int64_t w_rows, int taps, const int32_t* shifts, int cpad,
                     int bn, int split, float* out_dev, int M, int N, void* stream) {
    using namespace dfsfm;
    return guard([&] {
        DFSFM_CHECK(taps >= 1 && taps <= kMaxTaps, "taps out of range");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const __half* a = static_cast<const __half*>(a_dev);
        const __half* w = static_cast<const __half*>(w_dev);
        TmapPack maps;
        for (int i = 0; i < kMaxAMaps; ++i) maps.a[i] = make_tmap(a, C, a_rows, a_rows * C, kBM);
        const long long ktot = static_cast<long long>(taps) * cpad;
        maps.b = make_tmap(w, static_cast<int>(ktot), w_rows, w_rows * ktot, bbox(bn));
        GemmCore c;
        memset(&c, 0, sizeof(c));
        c.M = M;
        c.num_taps = taps;
        set_k(c, cpad);
        for (int t = 0; t < taps; ++t) { c.tap_map[t] = 0; c.tap_shift[t] = shifts[t]; }
        ConvEpiParams e;
        memset(&e, 0, sizeof(e));
        e.M = M; e.N = N; e.out_mode = OUT_FLAT; e.out_f32 = out_dev; e.out_f32_ld = N;
#define DFSFM_CASE(BN_)                                                                  \
    if (bn == BN_) {                                                                     \
        if (split) launch_gemm_counted<BN_, true, ConvEpi>(maps, c, e, N, st);           \
        else launch_gemm_counted<BN_, false, ConvEpi>(maps, c, e, N, st);                \
        return;                                                                          \
    }
        DFSFM_CASE(64) DFSFM_CASE(128) DFSFM_CASE(208) DFSFM_CASE(256)
#undef DFSFM_CASE
        throw Error("unsupported bn");
    });
}

// Tap-group ("slab") variant of the engine test hook: groups of 3 taps with consecutive shifts share one activation slab.
int dfsfm_debug_gemm_slab(const void* a_dev, int64_t a_rows, int C, const void* w_dev, int64_t w_rows, int taps, const int32_t* shifts, int cpad,
                          int bn, int bo_mode, float* out_dev, int M, int N, void* stream) {
    using namespace dfsfm;
    return guard([&] {
        DFSFM_CHECK(taps >= 3 && taps <= kMaxTaps && taps % 3 == 0, "taps must be a multiple of 3");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const __half* a = static_cast<const __half*>(a_dev);
        const __half* w = static_cast<const __half*>(w_dev);
        TmapPack maps;
        for (int i = 0; i < kMaxAMaps; ++i) maps.a[i] = make_tmap(a, C, a_rows, a_rows * C, kSlabRows);
        const long long ktot = static_cast<long long>(taps) * cpad;
        maps.b = make_tmap(w, static_cast<int>(ktot), w_rows, w_rows * ktot, bn / 2);
        GemmCore c;
        memset(&c, 0, sizeof(c));
        c.M = M;
        c.num_taps = taps;
        c.bo_mode = bo_mode;
        set_k(c, cpad);
        for (int t = 0; t < taps; ++t) { c.tap_map[t] = 0; c.tap_shift[t] = shifts[t]; }
        ConvEpiParams e;
        memset(&e, 0, sizeof(e));
        e.M = M; e.N = N; e.out_mode = OUT_FLAT; e.out_f32 = out_dev; e.out_f32_ld = N;
        LaunchScope ls("gemm_slab", st);
        if (bn == 128) launch_gemm2<128, true, ConvEpi, 3>(maps, c, e, N, st);
        else if (bn == 64) launch_gemm2<64, true, ConvEpi, 3>(maps, c, e, N, st);
        else throw Error("unsupported bn for the slab variant");
    });
}

}  // extern "C"
