// Experiment (GPU box): tcgen05.mma with the A operand in TENSOR MEMORY, CTA pair (cta_group::2), written by the epilogue warps with
// tcgen05.st as packed fp16 -- the mechanism the fused encoder-layer kernel uses to keep Q*Z and relu(hidden) on the SM.
//   D[256 x N] = A[256 x 64] . B[N x 64]^T,  A: lane = row, 32-bit column c holds K elements (2c, 2c+1); B: shared memory, K-major SW128.
// Self-checking: prints max |D - ref|; exit code 0 iff exact (inputs are small integers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I detectorfreesfm_b200/csrc tools/exp_tmem_a.cu -o tools/exp_tmem_a
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace dfsfm;

constexpr int kN = 64;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_f16_2sm_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// mode 0: A written in place over the low half of a 64-column region that first held fp32 data (the in-place pattern);
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
exp_kernel(const __half* __restrict__ A /*[256][64]*/, const __half* __restrict__ B /*[kN][64]*/, float* __restrict__ D /*[256][kN]*/) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* bt = smem;                                            // this CTA's half of B: (kN/2) rows x 128 B, SW128
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 8192);
    uint64_t* a_ready = bars;      // 8 epilogue warps (4 per CTA) -> leader
    uint64_t* d_full = bars + 1;   // commit, multicast to both CTAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (warp == 0 && lane == 0) {
        mbar_init(a_ready, 8);
        mbar_init(d_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2sm<128>(tmem_slot);
    // B half -> shared memory in the K-major 128-byte-swizzled layout (what TMA would write)
    for (int i = threadIdx.x; i < (kN / 2) * 8; i += blockDim.x) {
        const int r = i >> 3, j = i & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(B + (static_cast<int>(rank) * (kN / 2) + r) * 64 + j * 8);
        *reinterpret_cast<uint4*>(bt + r * 128 + ((j ^ (r & 7)) << 4)) = v;
    }
    fence_proxy_async();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kAcc = 0, kA = 64;   // accumulator columns [0,64), A operand columns [64, 96)
    if (warp >= 2) {
        const int quad = warp & 3;
        const int row = static_cast<int>(rank) * 128 + quad * 32 + lane;
        const uint32_t tw = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        uint32_t r[32];
        const uint4* src = reinterpret_cast<const uint4*>(A + row * 64);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint4 v = src[q];
            r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;   // element 2c in the low half of column c
        }
        tmem_st16(tw + kA, r);
        tmem_st16(tw + kA + 16, r + 16);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(a_ready, 0);
    }
    if (warp == 1 && lane == 0 && rank == 0) {
        mbar_wait(a_ready, 0);
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_f16(256, kN);
        for (int k = 0; k < 4; ++k) {
            const uint64_t db = make_smem_desc_sw128(smem_u32(bt) + k * 32);
            umma_f16_2sm_ts(tmem_base + kAcc, tmem_base + kA + 8 * k, db, idesc, k > 0);
        }
        umma_commit_2sm(d_full);
    }
    __syncwarp();
    if (warp >= 2) {
        const int quad = warp & 3;
        const int row = static_cast<int>(rank) * 128 + quad * 32 + lane;
        const uint32_t tw = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        mbar_wait(d_full, 0);
        tc_fence_after();
        for (int c0 = 0; c0 < kN; c0 += 32) {
            float v[32];
            tmem_ld32(tw + kAcc + c0, v);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) D[row * kN + c0 + j] = v[j];
        }
        tc_fence_before();
    }
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<128>(tmem_base);
    }
}

int main() {
    std::vector<__half> hA(256 * 64), hB(kN * 64);
    std::vector<float> fA(256 * 64), fB(kN * 64), ref(256 * kN), out(256 * kN);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = static_cast<float>(rand() % 7 - 3); hA[i] = __float2half(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = static_cast<float>(rand() % 5 - 2); hB[i] = __float2half(fB[i]); }
    for (int m = 0; m < 256; ++m)
        for (int n = 0; n < kN; ++n) {
            float s = 0;
            for (int k = 0; k < 64; ++k) s += fA[m * 64 + k] * fB[n * 64 + k];
            ref[m * kN + n] = s;
        }
    __half *dA, *dB;
    float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, out.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, out.size() * 4);
    cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    exp_kernel<<<2, 192, 16384>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tmem-A experiment: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    double md = 0;
    int bad = 0;
    for (size_t i = 0; i < out.size(); ++i) {
        const double d = fabs(static_cast<double>(out[i]) - ref[i]);
        if (!(d <= md)) md = d;
        if (d != 0) ++bad;
    }
    printf("tmem-A experiment (cta_group::2, M=256, N=%d, K=64): max |D - ref| = %g, mismatches = %d / %zu\n", kN, md, bad, out.size());
    if (bad) for (int m = 0; m < 256; m += 37) printf("  row %3d: got %g %g %g %g   want %g %g %g %g\n", m, out[m * kN], out[m * kN + 1], out[m * kN + 33],
                                                    out[m * kN + 63], ref[m * kN], ref[m * kN + 1], ref[m * kN + 33], ref[m * kN + 63]);
    return bad ? 1 : 0;
}
