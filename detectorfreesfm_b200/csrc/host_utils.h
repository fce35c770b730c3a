// Host-side helpers: error handling, split-fp16 device buffers, TMA tensor-map encoding, engine launch.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "gemm_engine.cuh"

namespace dfsfm {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define DFSFM_CUDA(expr)                                                                                   \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            throw ::dfsfm::Error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                                 std::to_string(__LINE__));                                                \
    } while (0)

#define DFSFM_CHECK(cond, msg)                                                   \
    do {                                                                         \
        if (!(cond)) throw ::dfsfm::Error(std::string("check failed: ") + (msg)); \
    } while (0)

// A split-fp16 activation / weight buffer: plane 0 = hi, plane 1 = lo, each [rows][C] row-major.
struct HL {
    __half* hi = nullptr;
    long long rows = 0;
    int C = 0;
    __half* lo() const { return hi + rows * C; }
    long long plane_elems() const { return rows * C; }
    size_t bytes() const { return static_cast<size_t>(rows) * C * 2 * sizeof(__half); }
};

// Allocation-time zero fill (halo cells, flags): cudaMemset runs in the legacy default stream, which does NOT order against the non-blocking
// streams the engines are driven on (a pair worker's side stream: its first kernels could read a halo before the fill had run) -- wait for it.
inline void zero_device_sync(void* p, int value, size_t bytes) {
    DFSFM_CUDA(cudaMemset(p, value, bytes));
    DFSFM_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
}
inline HL hl_alloc(long long rows, int C) {
    HL b;
    b.rows = rows;
    b.C = C;
    DFSFM_CUDA(cudaMalloc(&b.hi, b.bytes()));
    zero_device_sync(b.hi, 0, b.bytes());
    return b;
}
inline void hl_free(HL& b) {
    if (b.hi) cudaFree(b.hi);
    b.hi = nullptr;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        DFSFM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        DFSFM_CHECK(q == cudaDriverEntryPointSuccess && p, "cuTensorMapEncodeTiled not available");
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// dims {C (inner), rows, 2 planes}; box {64, box_rows, 1}; 128-byte swizzle; OOB -> zeros.
inline CUtensorMap make_tmap(const __half* base, int C, long long rows, long long plane_elems, int box_rows) {
    CUtensorMap m;
    memset(&m, 0, sizeof(m));
    DFSFM_CHECK(C % 8 == 0, "tensor-map inner dim must be a multiple of 8 halves");
    DFSFM_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor-map base must be 16-byte aligned");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(rows), 2};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(plane_elems) * 2};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
    return m;
}
inline CUtensorMap make_tmap(const HL& b, int box_rows) { return make_tmap(b.hi, b.C, b.rows, b.plane_elems(), box_rows); }

// Function attributes (max dynamic shared memory) and the SM count are PER DEVICE while the C ABI takes a device index per handle:
// "configured once" state is therefore kept per device ordinal.
inline int current_device() {
    int dev = 0;
    DFSFM_CUDA(cudaGetDevice(&dev));
    return dev;
}
struct PerDeviceOnce {
    std::mutex mu;
    bool done[64] = {};
    // runs f exactly once per device; other host threads launching the same kernel for the first time wait until it has finished (one
    // engine handle per thread may drive the GPU concurrently: a launch must not overtake the cudaFuncSetAttribute of its own kernel)
    template <class F>
    void run(F&& f) {
        const int d = current_device() & 63;
        std::lock_guard<std::mutex> lk(mu);
        if (!done[d]) {
            f();
            done[d] = true;
        }
    }
};

template <int BN, bool kSplit, class Epi>
inline void launch_gemm(const TmapPack& maps, const GemmCore& core, const typename Epi::Params& ep, int n_total, cudaStream_t st) {
    using Cfg = GemmCfg<BN, kSplit>;
    auto kern = gemm_tc_kernel<BN, kSplit, Epi>;
    static PerDeviceOnce once;
    once.run([&] { DFSFM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes)); });
    dim3 grid((core.M + kBM - 1) / kBM, (n_total + BN - 1) / BN);
    kern<<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(maps, core, ep);
    DFSFM_CUDA(cudaGetLastError());
}

inline int sm_count() {
    static std::atomic<int> n[64];
    const int dev = current_device();
    int v = n[dev & 63].load(std::memory_order_relaxed);
    if (!v) {
        DFSFM_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        n[dev & 63].store(v, std::memory_order_relaxed);
    }
    return v;
}
// Programmatic dependent launch: on by default (the next kernel's prologue overlaps the tail of the previous one: +3 % with one pair in
// flight).  With several pair workers per GPU it costs throughput -- the early-resident CTAs of a dependent launch hold SMs (200+ KB of
// shared memory each) that another worker's ready kernel could use: 233 -> 250 pairs/s with it off at two workers -- so a worker thread
// of a pool turns it off for its own launches (dfsfm_thread_set_pdl).  -1 = the process default (DFSFM_PDL, on).
inline int& pdl_thread_override() {
    static thread_local int v = -1;
    return v;
}
// DFSFM_TILE_STEAL=1: one cluster per tile and cluster-launch-control work stealing (gemm_engine.cuh TileSched) instead of the static
// round-robin tile order on min(tiles, SM pairs) persistent clusters.  Off by default: measured neutral on B200 (three pair workers 255.5 /
// 256.0 vs 256.9 / 256.1 pairs/s, conv launches 66.8-67.1 ms per step either way; profiles/r02_worker_pool.txt) -- uniform tiles leave
// nothing to balance and a late cluster costs little next to 9+ tiles per cluster.
inline bool tile_steal_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DFSFM_TILE_STEAL");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
inline bool pdl_enabled() {
    const int o = pdl_thread_override();
    if (o >= 0) return o == 1;
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DFSFM_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}
#ifdef DFSFM_LIN_DEBUG  // epilogue ablation switches (tuning builds only: results are wrong by design when set)
inline int lin_debug_flags() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DFSFM_LIN_DBG");
        v = e ? atoi(e) : 0;
    }
    return v;
}
inline void set_debug(LinEpiParams& e) { e.dbg = lin_debug_flags(); }
#endif
template <class T>
inline void set_debug(T&) {}
// debugging (common.cu): when a timeline capture is armed, the stamp buffer of the next engine-2 launch, else null
unsigned long long* timeline_next_launch(int grid_ctas, int tiles);
// engine v2 (persistent CTA pairs).  maps.b must have been built with box rows BN/2.
template <int BN, bool kSplit, class Epi, int G = 1>
inline void launch_gemm2(const TmapPack& maps, const GemmCore& core, const typename Epi::Params& ep, int n_total, cudaStream_t st) {
    using Cfg = Gemm2Cfg<BN, kSplit, Epi::kSeparateCorr, G, Epi::kEpiStageBytes>;
    auto kern = gemm_tc2_kernel<BN, kSplit, Epi, G>;
    if (G > 1) {
        DFSFM_CHECK(core.num_taps % G == 0, "tap groups need num_taps divisible by the group size");
        for (int t = 0; t < core.num_taps; ++t)
            DFSFM_CHECK(core.tap_map[t] == core.tap_map[(t / G) * G] && core.tap_shift[t] == core.tap_shift[(t / G) * G] + t % G,
                        "taps of a group must read the same map at consecutive row shifts");
    }
    static PerDeviceOnce once;
    once.run([&] { DFSFM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes)); });
    const int m_pairs = (core.M + 2 * kBM - 1) / (2 * kBM);
    const int n_tiles = (n_total + BN - 1) / BN;
    const int tiles = m_pairs * n_tiles;
    const int max_clusters = sm_count() / 2;
    // stateless epilogues: one cluster per tile, running clusters steal the tiles of clusters not yet launched (gemm_engine.cuh TileSched)
    const bool steal = !Epi::kHasState && tile_steal_enabled() && tiles > max_clusters;
    const int clusters = (steal || tiles < max_clusters) ? tiles : max_clusters;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kGemm2Threads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: overlap this kernel's prologue with the previous kernel's tail
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GemmCore core_l = core;
    core_l.tl = timeline_next_launch(2 * clusters < 148 ? 2 * clusters : 148, tiles);
    typename Epi::Params ep_l = ep;
    set_debug(ep_l);
    DFSFM_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, core_l, ep_l, tiles, n_tiles, steal ? 1 : 0));
}

// Fill the tap table of a stride-1 k x k convolution on a flat-halo geometry with row pitch Wp.
inline void conv_taps_s1(GemmCore& c, int k, int Wp) {
    const int r = k / 2;
    c.num_taps = k * k;
    int t = 0;
    for (int dy = -r; dy <= r; ++dy)
        for (int dx = -r; dx <= r; ++dx, ++t) {
            c.tap_map[t] = 0;
            c.tap_shift[t] = dy * Wp + dx;
        }
}
// 3x3 stride-2 convolution reading the four parity planes (map index = (dy&1)*2 + (dx&1)); Wp = OUTPUT pitch.
inline void conv_taps_s2(GemmCore& c, int Wp) {
    c.num_taps = 9;
    int t = 0;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx, ++t) {
            c.tap_map[t] = static_cast<int8_t>((dy & 1) * 2 + (dx & 1));
            c.tap_shift[t] = (dy < 0 ? -Wp : 0) + (dx < 0 ? -1 : 0);
        }
}
inline void set_k(GemmCore& c, int cpad) {
    c.cpad = cpad;
    c.kchunks = (cpad + 63) / 64;
    const int rem = cpad - (c.kchunks - 1) * 64;
    c.k16_last = (rem + 15) / 16;
}

}  // namespace dfsfm
