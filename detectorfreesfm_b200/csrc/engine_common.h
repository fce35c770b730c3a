// Shared by the engine translation units: parameter store, C-ABI error guard, launch counter.
#pragma once
#include <type_traits>
#include <atomic>
#include <functional>
#include <map>
#include <string>

#include "host_utils.h"
#include "encoder_fused.cuh"
#include "mlp_fused.cuh"
#include "simt_kernels.cuh"

namespace dfsfm {

void set_last_error(const std::string& s);
std::atomic<long long>& launch_counter();
inline void count_launch(int n = 1) { launch_counter().fetch_add(n, std::memory_order_relaxed); }

template <class F>
inline int guard(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return 1;
    } catch (...) {
        set_last_error("unknown error");
        return 2;
    }
}

// Optional per-launch CUDA-event timing (dfsfm_profile_*): bench.py turns it on for a few extra steps to attribute device
// time to kernel families; it is off in every timed region.
bool profiling_enabled();
void prof_begin(const char* label, cudaStream_t st);
void prof_end(cudaStream_t st);
struct LaunchScope {  // counts one kernel launch; brackets it with events when profiling
    cudaStream_t st;
    bool on;
    LaunchScope(const char* label, cudaStream_t s) : st(s), on(profiling_enabled()) {
        if (on) prof_begin(label, st);
    }
    ~LaunchScope() {
        if (on) prof_end(st);
        count_launch();
    }
};

// Engine selection: 2 = persistent CTA pairs (default), 1 = one CTA per tile.  DFSFM_ENGINE overrides (tests compare both).
int engine_version();
void set_engine_version(int v);
inline int bbox(int bn) { return engine_version() == 2 ? bn / 2 : bn; }  // box rows of the weight tensor map

template <int BN, bool kSplit, class Epi>
inline void launch_gemm_counted(const TmapPack& maps, const GemmCore& core, const typename Epi::Params& ep, int n_total, cudaStream_t st,
                                const char* label = "gemm") {
    LaunchScope ls(label, st);
    if (engine_version() == 2) {
        if constexpr (std::is_same<Epi, LinEpi>::value) {
            // one kernel per epilogue flavour (see LinEpiS)
            if (ep.mode == LIN_F32_ELU) launch_gemm2<BN, kSplit, LinEpiS<LIN_F32_ELU, 0>>(maps, core, ep, n_total, st);
            else if (ep.mode == LIN_RELU_HL) launch_gemm2<BN, kSplit, LinEpiS<LIN_RELU_HL, 0>>(maps, core, ep, n_total, st);
            else if (ep.mode == LIN_QZ) launch_gemm2<BN, kSplit, LinEpiS<LIN_QZ, 0>>(maps, core, ep, n_total, st);
            else if (ep.resid != nullptr) launch_gemm2<BN, kSplit, LinEpiS<LIN_LN, 1>>(maps, core, ep, n_total, st);
            else if (ep.res_hi != nullptr) launch_gemm2<BN, kSplit, LinEpiS<LIN_LN, 2>>(maps, core, ep, n_total, st);
            else launch_gemm2<BN, kSplit, LinEpiS<LIN_LN, 0>>(maps, core, ep, n_total, st);
        } else {
            launch_gemm2<BN, kSplit, Epi>(maps, core, ep, n_total, st);
        }
    } else {
        if constexpr (std::is_same<Epi, LinEpi>::value) DFSFM_CHECK(ep.mode != LIN_QZ, "LIN_QZ needs engine 2");
        launch_gemm<BN, kSplit, Epi>(maps, core, ep, n_total, st);
    }
}

// Fused mlp.0 + ReLU + mlp.2 + norm2 + residual for d_model 128 (mlp_fused.cuh); DFSFM_FUSED_MLP=0 selects the two-GEMM path.
inline bool fused_mlp_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DFSFM_FUSED_MLP");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1 && engine_version() == 2;
}
inline void launch_mlp128_fused(const HL& x, const HL& m1, long long row0, long long T, const HL& w0, const HL& w2, const float* gamma,
                                const float* beta, float* xf, cudaStream_t st) {
    static PerDeviceOnce once;
    once.run([&] { DFSFM_CUDA(cudaFuncSetAttribute(mlp128_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMlpSmemBytes)); });
    MlpMaps maps;
    maps.x = make_tmap(x.hi + row0 * 128, 128, T, x.plane_elems(), 128);
    maps.m1 = make_tmap(m1.hi + row0 * 128, 128, T, m1.plane_elems(), 128);
    maps.w0 = make_tmap(w0, 128);
    maps.w2 = make_tmap(w2, 64);
    MlpParams p;
    p.T = static_cast<int>(T);
    p.gamma = gamma; p.beta = beta;
    p.xf = xf + row0 * 128;
    p.x_hi = x.hi + row0 * 128;
    p.x_lo = x.lo() + row0 * 128;
    const int tiles = static_cast<int>((T + 2 * kBM - 1) / (2 * kBM));
    const int max_clusters = sm_count() / 2;
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    LaunchScope ls("mlp_fused", st);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kMlpThreads);
    cfg.dynamicSmemBytes = kMlpSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DFSFM_CUDA(cudaLaunchKernelEx(&cfg, mlp128_fused_kernel, maps, p, tiles));
}

// Fused encoder layer (encoder_fused.cuh): q projection + normaliser, attention/merge GEMM, LayerNorm1, mlp, LayerNorm2 + residual
// for the token rows [row0, row0 + T) of `x`.  `wqkv` rows [0,256) are Wq; `g` holds the folded matrix of each segment.
inline void launch_enc256_fused(const HL& x, long long row0, long long T, const HL& wqkv, const HL& g, const HL& w0, const HL& w2, EncParams p,
                                cudaStream_t st) {
    static PerDeviceOnce once;
    once.run([&] { DFSFM_CUDA(cudaFuncSetAttribute(enc256_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncSmemBytes)); });
    EncMaps maps;
    maps.x = make_tmap(x.hi + row0 * 256, 256, T, x.plane_elems(), 128);
    maps.wq = make_tmap(wqkv.hi, 256, 256, wqkv.plane_elems(), 128);
    maps.g = make_tmap(g, 128);
    maps.w0 = make_tmap(w0, 128);
    maps.w2 = make_tmap(w2, 128);
    p.T = static_cast<int>(T);
    const int tiles = static_cast<int>((T + 2 * kBM - 1) / (2 * kBM));
    const int max_clusters = sm_count() / 2;
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    p.tl = timeline_next_launch(2 * clusters, tiles);
    LaunchScope ls("enc_fused", st);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kEncThreads);
    cfg.dynamicSmemBytes = kEncSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DFSFM_CUDA(cudaLaunchKernelEx(&cfg, enc256_fused_kernel, maps, p, tiles));
}

// 3x3 stride-1 convolutions with BN <= 128: the three dx taps of a kernel row share one activation slab (engine 2 only).
inline bool slab_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DFSFM_SLAB");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1 && engine_version() == 2;
}
inline bool slab_applicable(const GemmCore& c, int n_in) {
    if (!slab_enabled() || n_in != 1 || c.num_taps != 9) return false;
    for (int t = 0; t < 9; ++t)
        if (c.tap_map[t] != 0 || c.tap_shift[t] != c.tap_shift[(t / 3) * 3] + t % 3) return false;
    return true;
}
template <int BN, class Epi>
inline void launch_gemm_slab_counted(const TmapPack& maps, const GemmCore& core, const typename Epi::Params& ep, int n_total, cudaStream_t st,
                                     const char* label) {
    LaunchScope ls(label, st);
    launch_gemm2<BN, true, Epi, 3>(maps, core, ep, n_total, st);
}

// Packed parameters on the device: GEMM operands as split-fp16 [rows][cols], everything else as fp32.
class ParamStore {
  public:
    ~ParamStore() {
        for (auto& kv : mats_) hl_free(kv.second);
        for (auto& kv : vecs_) cudaFree(kv.second);
    }
    void set(const std::string& name, const float* host, long long rows, long long cols, int kind) {
        const long long n = rows * cols;
        DFSFM_CHECK(n > 0, "empty parameter " + name);
        if (kind == 0) {
            DFSFM_CHECK(cols % 8 == 0, "GEMM operand K must be a multiple of 8: " + name);
            std::vector<__half> tmp(static_cast<size_t>(2 * n));
            for (long long i = 0; i < n; ++i) {
                const __half h = __float2half_rn(host[i]);
                tmp[i] = h;
                tmp[n + i] = __float2half_rn(host[i] - __half2float(h));
            }
            auto it = mats_.find(name);
            if (it != mats_.end()) { hl_free(it->second); mats_.erase(it); }
            HL b = hl_alloc(rows, static_cast<int>(cols));
            DFSFM_CUDA(cudaMemcpy(b.hi, tmp.data(), tmp.size() * sizeof(__half), cudaMemcpyHostToDevice));
            DFSFM_CUDA(cudaStreamSynchronize(cudaStreamLegacy));  // pageable source: the DMA may still be in flight, and non-blocking streams do not wait for it
            mats_[name] = b;
        } else {
            auto it = vecs_.find(name);
            if (it != vecs_.end()) { cudaFree(it->second); vecs_.erase(it); }
            float* d = nullptr;
            DFSFM_CUDA(cudaMalloc(&d, static_cast<size_t>(n) * sizeof(float)));
            DFSFM_CUDA(cudaMemcpy(d, host, static_cast<size_t>(n) * sizeof(float), cudaMemcpyHostToDevice));
            DFSFM_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
            vecs_[name] = d;
        }
    }
    const HL& mat(const std::string& name) const {
        auto it = mats_.find(name);
        if (it == mats_.end()) throw Error("missing GEMM parameter '" + name + "' (upload weights first)");
        return it->second;
    }
    const float* vec(const std::string& name) const {
        auto it = vecs_.find(name);
        if (it == vecs_.end()) throw Error("missing fp32 parameter '" + name + "' (upload weights first)");
        return it->second;
    }
    bool has_vec(const std::string& name) const { return vecs_.count(name) != 0; }

  private:
    std::map<std::string, HL> mats_;
    std::map<std::string, float*> vecs_;
};

}  // namespace dfsfm
