// Experiment (GPU box): the A operand of a CTA-pair tcgen05.mma staged shared memory -> TENSOR MEMORY with tcgen05.cp (128x256b = one K=16
// slice of 128 rows per CTA) from the K-major 128-byte-swizzled layout TMA writes, with the ROW-SHIFTED start addresses the 3x3 convolutions
// use (one (128+8)-row slab per kernel row, tap i reads rows [i, i+128)); then .ts MMAs read A from tensor memory.  Question answered: does
// tcgen05.cp take the same shared-memory descriptor as the MMA's A operand (swizzle resolved from the address bits)?
//   D[256 x N] = A_shift[256 x 64] . B[N x 64]^T ; self-checking for shift = 0, 1, 2 (small integers: exact).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I detectorfreesfm_b200/csrc tools/exp_tmem_cp.cu -o /tmp/exp_tmem_cp
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace dfsfm;

constexpr int kN = 64;
constexpr int kSlab = 136;

__device__ __forceinline__ void umma_f16_2sm_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b_2sm(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
exp_kernel(const __half* __restrict__ A /*[2][kSlab][64]*/, const __half* __restrict__ B /*[kN][64]*/, float* __restrict__ D /*[256][kN]*/, int shift) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* at = smem;                                            // this CTA's slab: kSlab rows x 128 B, SW128 (17 KB -> 18 KB region)
    uint8_t* bt = smem + 18 * 1024;                                // this CTA's half of B
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 18 * 1024 + 8192);
    uint64_t* d_full = bars;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (warp == 0 && lane == 0) {
        mbar_init(d_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2sm<128>(tmem_slot);
    for (int i = threadIdx.x; i < kSlab * 8; i += blockDim.x) {
        const int r = i >> 3, j = i & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(A + (static_cast<int>(rank) * kSlab + r) * 64 + j * 8);
        *reinterpret_cast<uint4*>(at + r * 128 + ((j ^ (r & 7)) << 4)) = v;
    }
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
    if (warp == 1 && lane == 0 && rank == 0) {
        constexpr uint32_t idesc = make_idesc_f16(256, kN);
        for (int k = 0; k < 4; ++k) tmem_cp_128x256b_2sm(tmem_base + kA + 8 * k, make_smem_desc_sw128(smem_u32(at) + shift * 128 + k * 32));
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
    std::vector<__half> hA(2 * kSlab * 64), hB(kN * 64);
    std::vector<float> fA(hA.size()), fB(hB.size()), ref(256 * kN), out(256 * kN);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = static_cast<float>(rand() % 7 - 3); hA[i] = __float2half(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = static_cast<float>(rand() % 5 - 2); hB[i] = __float2half(fB[i]); }
    __half *dA, *dB;
    float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, out.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
    int rc = 0;
    for (int shift = 0; shift < 3; ++shift) {
        for (int m = 0; m < 256; ++m)
            for (int n = 0; n < kN; ++n) {
                float s = 0;
                const int cta = m / 128, r = m % 128 + shift;
                for (int k = 0; k < 64; ++k) s += fA[(cta * kSlab + r) * 64 + k] * fB[n * 64 + k];
                ref[m * kN + n] = s;
            }
        cudaMemset(dD, 0xff, out.size() * 4);
        exp_kernel<<<2, 192, 40 * 1024>>>(dA, dB, dD, shift);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("tmem-cp experiment: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
        cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        double md = 0;
        for (size_t i = 0; i < out.size(); ++i) {
            const double d = fabs(static_cast<double>(out[i]) - ref[i]);
            if (!(d <= md)) md = d;
            if (d != 0) ++bad;
        }
        printf("tmem-cp experiment (cta_group::2, 128x256b, row shift %d): max |D - ref| = %g, mismatches = %d / %zu\n", shift, md, bad, out.size());
        if (bad) {
            rc = 1;
            for (int m = 0; m < 256; m += 37) printf("  row %3d: got %g %g %g %g   want %g %g %g %g\n", m, out[m * kN], out[m * kN + 1], out[m * kN + 33],
                                                    out[m * kN + 63], ref[m * kN], ref[m * kN + 1], ref[m * kN + 33], ref[m * kN + 63]);
        }
    }
    return rc;
}
