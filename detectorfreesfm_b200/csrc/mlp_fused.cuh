// Fused feed-forward block of a d_model = 128 LoFTR encoder layer (multi-view transformer of HP-2, loftr_fine of HP-1):
//
//     hid = relu(cat[x, m1] . W0^T)              (mlp.0, 256 -> 256)
//     out = x + LayerNorm(hid . W2^T)            (mlp.2, 256 -> 128; norm2; residual)       matcher_module/transformer.py:88-95
//
// in ONE kernel per 256-token tile: the 256-wide hidden activation never leaves the SM.  A CTA pair (cta_group::2) owns the
// tile; warp 0 of each CTA TMA-loads its 128 rows of x and m1 (split-fp16 planes) and its half of the weight tiles through a
// 2-stage ring; the leader issues the M=256 tcgen05.mma for mlp.0 into TMEM; eight epilogue warps per CTA read the accumulator,
// apply ReLU, split to (hi, lo) and write it back to shared memory IN THE UMMA OPERAND LAYOUT (128-byte swizzle), over the x / m1
// tiles that are dead by then; the leader then issues mlp.2 from that operand, and the epilogue finishes with LayerNorm +
// residual straight to HBM.  Per token-layer this removes 2 KB of hid traffic and one kernel boundary.
#pragma once
#include "gemm_engine.cuh"

namespace dfsfm {

constexpr int kMlpThreads = 64 + 32 * 8;
constexpr int kMlpActBytes = 4 * 32 * 1024;   // 4 K-chunks x (hi 16 KB + lo 16 KB): x, x, m1, m1 -> later hid chunks 0..3
constexpr int kMlpRingStage = 32 * 1024;      // one weight chunk: this CTA's half, hi + lo
constexpr int kMlpSmemBytes = kMlpActBytes + 2 * kMlpRingStage + 1024 + 256;

struct MlpMaps {
    CUtensorMap x;    // {128, T, 2} box {64, 128, 1}
    CUtensorMap m1;   // same geometry
    CUtensorMap w0;   // {256 (K), 256 (rows), 2} box {64, 128, 1}: this CTA's 128 of the 256 output channels
    CUtensorMap w2;   // {256 (K), 128 (rows), 2} box {64, 64, 1}
};
struct MlpParams {
    int T;                 // token rows
    const float* gamma;    // norm2
    const float* beta;
    float* xf;             // fp32 residual stream [T][128], updated in place
    __half* x_hi;          // split planes of the updated tokens
    __half* x_lo;
};

static __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
mlp128_fused_kernel(const __grid_constant__ MlpMaps maps, const MlpParams p, const int num_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* act = smem;
    uint8_t* ring = smem + kMlpActBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + 2 * kMlpRingStage);
    uint64_t* in_full = bars + 0;
    uint64_t* w_full = bars + 1;    // [2]
    uint64_t* w_empty = bars + 3;   // [2]
    uint64_t* s2_done = bars + 5;
    uint64_t* hid_full = bars + 6;
    uint64_t* s3_done = bars + 7;
    uint64_t* e3_done = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.x); tma_prefetch_desc(&maps.m1); tma_prefetch_desc(&maps.w0); tma_prefetch_desc(&maps.w2);
        mbar_init(in_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        mbar_init(s2_done, 1);
        mbar_init(hid_full, 16);  // 8 epilogue warps x 2 CTAs
        mbar_init(s3_done, 1);
        mbar_init(e3_done, 16);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2sm<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    constexpr uint32_t kAcc2 = 0, kAcc3 = 256;

    if (warp == 0) {
        if (elect_one()) {
            int ws = 0;
            uint32_t wphase = 0, tphase = 0;
            bool first = true;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m0 = (tile * 2 + static_cast<int>(rank)) * kBM;
                if (!first) { mbar_wait(s3_done, tphase); tphase ^= 1; }  // mlp.2 of the previous tile has consumed hid: act is free
                first = false;
                if (rank == 0) mbar_arrive_expect_tx(in_full, 2 * kMlpActBytes);
                for (int c = 0; c < 4; ++c) {
                    const CUtensorMap* am = c < 2 ? &maps.x : &maps.m1;
                    tma_load_3d_2sm(act + c * 32768, am, in_full, (c & 1) * 64, m0, 0);
                    tma_load_3d_2sm(act + c * 32768 + 16384, am, in_full, (c & 1) * 64, m0, 1);
                }
                for (int i = 0; i < 8; ++i) {  // 4 K-chunks of W0 (this CTA's 128 rows), then 4 of W2 (this CTA's 64 rows)
                    mbar_wait(&w_empty[ws], wphase ^ 1);
                    uint8_t* st = ring + ws * kMlpRingStage;
                    if (i < 4) {
                        if (rank == 0) mbar_arrive_expect_tx(&w_full[ws], 2 * 32768);
                        tma_load_3d_2sm(st, &maps.w0, &w_full[ws], i * 64, static_cast<int>(rank) * 128, 0);
                        tma_load_3d_2sm(st + 16384, &maps.w0, &w_full[ws], i * 64, static_cast<int>(rank) * 128, 1);
                    } else {
                        if (rank == 0) mbar_arrive_expect_tx(&w_full[ws], 2 * 16384);
                        tma_load_3d_2sm(st, &maps.w2, &w_full[ws], (i - 4) * 64, static_cast<int>(rank) * 64, 0);
                        tma_load_3d_2sm(st + 8192, &maps.w2, &w_full[ws], (i - 4) * 64, static_cast<int>(rank) * 64, 1);
                    }
                    if (++ws == 2) { ws = 0; wphase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc2 = make_idesc_f16(256, 256);
            constexpr uint32_t idesc3 = make_idesc_f16(256, 128);
            int ws = 0;
            uint32_t wphase = 0, tphase = 0;
            bool first = true;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                mbar_wait(in_full, tphase);
                tc_fence_after();
                uint32_t acc = 0;
                for (int c = 0; c < 4; ++c) {  // mlp.0: K = [x | m1]
                    mbar_wait(&w_full[ws], wphase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(act + c * 32768), b_hi = smem_u32(ring + ws * kMlpRingStage);
                    const uint64_t da0 = make_smem_desc_sw128(a_hi), db0 = make_smem_desc_sw128(b_hi);   // + constants per K step / lo plane
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = da0 + 2u * k, db = db0 + 2u * k;
                        umma_f16_2sm(tmem_base + kAcc2, da, db, idesc2, acc);
                        umma_f16_2sm(tmem_base + kAcc2, da, db + (16384 >> 4), idesc2, 1);
                        umma_f16_2sm(tmem_base + kAcc2, da + (16384 >> 4), db, idesc2, 1);
                        acc = 1;
                    }
                    umma_commit_2sm(&w_empty[ws]);
                    if (++ws == 2) { ws = 0; wphase ^= 1; }
                }
                umma_commit_2sm(s2_done);
                mbar_wait(hid_full, tphase);                 // both CTAs have written relu(hid) as the next A operand
                if (!first) mbar_wait(e3_done, tphase ^ 1);  // the previous tile's LayerNorm epilogue has drained acc3
                first = false;
                tc_fence_after();
                acc = 0;
                for (int c = 0; c < 4; ++c) {  // mlp.2: K = hid
                    mbar_wait(&w_full[ws], wphase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(act + c * 32768), b_hi = smem_u32(ring + ws * kMlpRingStage);
                    const uint64_t da0 = make_smem_desc_sw128(a_hi), db0 = make_smem_desc_sw128(b_hi);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = da0 + 2u * k, db = db0 + 2u * k;
                        umma_f16_2sm(tmem_base + kAcc3, da, db, idesc3, acc);
                        umma_f16_2sm(tmem_base + kAcc3, da, db + (8192 >> 4), idesc3, 1);
                        umma_f16_2sm(tmem_base + kAcc3, da + (16384 >> 4), db, idesc3, 1);
                        acc = 1;
                    }
                    umma_commit_2sm(&w_empty[ws]);
                    if (++ws == 2) { ws = 0; wphase ^= 1; }
                }
                umma_commit_2sm(s3_done);
                tphase ^= 1;
            }
        }
        __syncwarp();
    } else {
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const uint32_t tmem_w = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const int r = quad * 32 + lane;  // row of this thread inside the CTA's 128-row tile
        uint32_t tphase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const long long row = static_cast<long long>(tile * 2 + static_cast<int>(rank)) * kBM + r;
            const bool valid = row < p.T;
            // ---- E2: relu(acc2) -> (hi, lo) -> UMMA A-operand tiles in shared memory (this warp: 128 of the 256 columns)
            mbar_wait(s2_done, tphase);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_w + kAcc2 + c0, v);
                tmem_ld_wait();
                uint8_t* chunk = act + (c0 >> 6) * 32768 + r * 128;  // K-chunk c0/64, row r (8-row atoms are contiguous: SBO = 1024)
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint4 uh, ul;
                    __half* hh = reinterpret_cast<__half*>(&uh);
                    __half* hl = reinterpret_cast<__half*>(&ul);
#pragma unroll
                    for (int q = 0; q < 8; ++q) split_f16(valid ? fmaxf(v[j + q], 0.f) : 0.f, hh[q], hl[q]);
                    const int j16 = ((c0 & 63) + j) >> 3;                 // 16-byte chunk index inside the 128-byte row
                    const int phys = (j16 ^ (r & 7)) << 4;                // 128-byte swizzle: chunk index XOR (row mod 8)
                    *reinterpret_cast<uint4*>(chunk + phys) = uh;
                    *reinterpret_cast<uint4*>(chunk + 16384 + phys) = ul;
                }
            }
            fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(hid_full, 0);
            // ---- E3: LayerNorm(acc3) + residual -> HBM (stats over the full 128-wide row, this warp writes 64 columns)
            mbar_wait(s3_done, tphase);
            tc_fence_after();
            float s1 = 0.f, s2 = 0.f, pivot = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_w + kAcc3 + c0, v);
                tmem_ld_wait();
                if (c0 == 0) pivot = v[0];
#pragma unroll
                for (int j = 0; j < 32; ++j) { const float d = v[j] - pivot; s1 += d; s2 = fmaf(d, d, s2); }
            }
            const float md = s1 * (1.f / 128.f);
            const float mean = pivot + md;
            const float rstd = rsqrtf(fmaxf(s2 * (1.f / 128.f) - md * md, 0.f) + 1e-5f);
#pragma unroll 1
            for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_w + kAcc3 + c0, v);
                tmem_ld_wait();
                if (!valid) continue;
                float* xr = p.xf + row * 128 + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + c0 + j));
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.beta + c0 + j));
                    const float4 r4 = *reinterpret_cast<const float4*>(xr + j);
                    v[j] = r4.x + ((v[j] - mean) * rstd * g4.x + b4.x);
                    v[j + 1] = r4.y + ((v[j + 1] - mean) * rstd * g4.y + b4.y);
                    v[j + 2] = r4.z + ((v[j + 2] - mean) * rstd * g4.z + b4.z);
                    v[j + 3] = r4.w + ((v[j + 3] - mean) * rstd * g4.w + b4.w);
                    *reinterpret_cast<float4*>(xr + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint4 uh, ul;
                    __half* hh = reinterpret_cast<__half*>(&uh);
                    __half* hl = reinterpret_cast<__half*>(&ul);
#pragma unroll
                    for (int q = 0; q < 8; ++q) split_f16(v[j + q], hh[q], hl[q]);
                    *reinterpret_cast<uint4*>(p.x_hi + row * 128 + c0 + j) = uh;
                    *reinterpret_cast<uint4*>(p.x_lo + row * 128 + c0 + j) = ul;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(e3_done, 0);
            tphase ^= 1;
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<512>(tmem_base);
    }
}

}  // namespace dfsfm
