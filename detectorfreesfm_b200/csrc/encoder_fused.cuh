// Fused LoFTR encoder layer, attending-token side (d_model 256, 8 heads; third_party/LoFTR/src/loftr/loftr_module/
// transformer.py:35-58 after the k/v state of the source has been reduced -- KvEpi + attn_fold_merge_kernel):
//
//     q   = x . Wq^T ;  Q = elu(q)+1 ;  Z = 1/(Q . Ksum_head + eps)                         (linear_attention.py:32,43)
//     mg  = (Q*Z*len) . G^T                  G = KV . Wm^T : attention + merge in one GEMM           (:44-45, transformer.py:48)
//     m   = LayerNorm1(mg)
//     hid = relu([x | m] . W0^T)             512 wide                                           (transformer.py:52)
//     out = x + LayerNorm2(hid . W2^T)                                                          (:53-55)
//
// ONE kernel per 256-token tile (CTA pair, cta_group::2, 128 rows per CTA): nothing but x in and x out touches HBM
// (1 KB + 1 KB per token as split-fp16 planes; x is re-read once more from L2 for the first half of mlp.0 and once for the residual).
//
// On-chip residency (per CTA):
//   shared memory   act  128 KB : x tile (TMA, 4 K-chunks of [hi 16 KB | lo 16 KB]) -> later m, written by the epilogue warps in the
//                                 UMMA operand layout (128-byte swizzle) -> chunk 3 doubles as the store-transposition staging of the
//                                 last epilogue
//                   ring  96 KB : 3 stages x 32 KB of streamed operand chunks (weights; x chunks for mlp.0)
//   tensor memory   R0 [0,256)  : acc(q) -> Q*Z as packed fp16 (hi | lo interleaved per 32-column block) = the A OPERAND of the next
//                                 GEMM read straight from TMEM (tcgen05.mma .ts form) -> acc(hid chunk) -> relu(hid) as A operand
//                   R1 [256,512): acc(mg) -> acc(out)
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA) + TMEM allocator, warps 2..9 = epilogues (two per TMEM lane
// quadrant, each owning 128 of the 256 columns).  All steps of a tile are dependent, so within a tile the tensor pipe and the
// epilogue warps alternate; the first GEMM of the next tile overlaps the last epilogue of this one.
#pragma once
#include "gemm_engine.cuh"

namespace dfsfm {

constexpr int kEncThreads = 64 + 32 * 8;
constexpr int kEncChunk = 32 * 1024;            // one 64-wide K chunk of a 128-row operand: hi 16 KB + lo 16 KB
constexpr int kEncAct = 4 * kEncChunk;
constexpr int kEncRingStages = 3;
constexpr int kEncBarOff = kEncAct + kEncRingStages * kEncChunk;
constexpr int kEncLnxOff = kEncBarOff + 512;    // LayerNorm-1 statistics exchange: 8 warps x 256 B
constexpr int kEncSmemBytes = kEncLnxOff + 8 * 256;
static_assert(kEncSmemBytes <= 227 * 1024, "fused encoder layer: shared memory budget");

struct EncMaps {
    CUtensorMap x;    // token planes {256, T, 2}, box {64, 128, 1}
    CUtensorMap wq;   // {256 (K), 256 rows, 2}, box {64, 128, 1}: this CTA's half of the output channels
    CUtensorMap g;    // {256, 512 rows (2 segments), 2}, box {64, 128, 1}
    CUtensorMap w0;   // {512, 512 rows, 2}, box {64, 128, 1}
    CUtensorMap w2;   // {512, 256 rows, 2}, box {64, 128, 1}
};
struct EncParams {
    int T;                 // token rows of this launch
    int seg_tile0;         // tiles >= seg_tile0 belong to segment 1 (its G rows and Ksum); <= 0: one segment
    const float* ksum;     // [2][256]
    float qz_scale[2];     // per segment: source length (Q*Z*len is O(1); see LinEpiParams::qz_scale)
    const float* ln1_g;
    const float* ln1_b;
    LinEpiParams e4;       // LayerNorm2 + residual + stores (M, N = 256, gamma/beta, res_hi/lo, out_hi/lo[, out_f32])
    unsigned long long* tl;  // debugging: when set, [CTA][16] globaltimer stamps (dfsfm_debug_timeline): 0 entry, 1 set-up done, 2 dependency
                             //   wait over, 3 first x chunk landed, 4 GEMM1 issued; first tile, epilogue warp 2: 5/6 E1 begin/end, 7/8 E2, 9 E3(0)
                             //   begin, 10 E3(1) end, 12/13 E4 begin/end; 14 last tile done, 11 exit
};
__device__ __forceinline__ void enc_stamp(const EncParams& p, int ev) {
    if (p.tl != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.tl[blockIdx.x * 16 + ev] = t;
    }
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem, both CTAs] (+)= A[TMEM of each CTA: lane = row, 32-bit column c = K elements 2c, 2c+1] * B[smem halves]^T
__device__ __forceinline__ void umma_f16_2sm_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// TMA store shared -> global (bulk async group), rows past the tensor's extent are clipped
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 32 fp32 values of one row -> the in-place A-operand image of their 32-column block: columns [0,16) = hi halves (elements 2c, 2c+1 in
// column c), columns [16,32) = lo halves.  K step s (16 elements) of the block reads columns 8s.. (hi) and 16+8s.. (lo).
__device__ __forceinline__ void pack_block_hl(const float* v, uint32_t* r) {
#pragma unroll
    for (int c = 0; c < 16; ++c) split_f16x2(v[2 * c], v[2 * c + 1], r[c], r[16 + c]);
}

static __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kEncThreads, 1)
enc256_fused_kernel(const __grid_constant__ EncMaps maps, const EncParams p, const int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* act = smem;
    uint8_t* ring = smem + kEncAct;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kEncBarOff);
    uint64_t* x_full = bars;            // [4]  TMA -> MMA (leader)
    uint64_t* r_full = bars + 4;        // [3]  ring
    uint64_t* r_empty = bars + 7;       // [3]
    uint64_t* q_done = bars + 10;       // MMA commits (multicast to both CTAs)
    uint64_t* mg_done = bars + 11;
    uint64_t* h_done = bars + 12;       // [2]
    uint64_t* out_done = bars + 14;
    uint64_t* g4_done = bars + 15;
    uint64_t* g3m_done = bars + 16;
    // epilogue -> MMA hand-offs, one barrier per 64-column K chunk of the consuming GEMM: a chunk belongs to one column half, i.e. to the
    // four quad warps of that half in both CTAs (8 arrivals on the leader's barrier).  The consuming GEMM takes the chunks in the order
    // 0, 2, 1, 3 -- both halves finish their first chunk at the same time -- so its first K steps run under the rest of the epilogue.
    uint64_t* qp_ready = bars + 17;     // [4]
    uint64_t* m_ready = bars + 21;      // [4]
    uint64_t* hid_ready = bars + 25;    // [2][4]
    uint64_t* e4_done = bars + 33;      // epilogue warps of both CTAs -> BOTH CTAs (16 arrivals each)
    uint64_t* xr_full = bars + 34;      // residual x tile landed in act (per CTA, local TMA)
    uint64_t* x0_done = bars + 35;      // MMA commit: the x part of mlp.0 chunk 0 has read x from act (the epilogue may overwrite it with m)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);
    auto chunk_order = [](int i) { return ((i & 1) << 1) | (i >> 1); };   // 0, 2, 1, 3
    float* lnx = reinterpret_cast<float*>(smem + kEncLnxOff);

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) enc_stamp(p, 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        if ((smem_u32(smem) & 1023u) != 0) asm volatile("trap;");  // the swizzled operand layout needs a 1 KB aligned base
        tma_prefetch_desc(&maps.x); tma_prefetch_desc(&maps.wq); tma_prefetch_desc(&maps.g); tma_prefetch_desc(&maps.w0); tma_prefetch_desc(&maps.w2);
        for (int i = 0; i < 4; ++i) mbar_init(&x_full[i], 1);
        for (int i = 0; i < kEncRingStages; ++i) { mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], 1); }
        mbar_init(q_done, 1); mbar_init(mg_done, 1); mbar_init(&h_done[0], 1); mbar_init(&h_done[1], 1);
        mbar_init(out_done, 1); mbar_init(g4_done, 1); mbar_init(g3m_done, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&qp_ready[i], 8); mbar_init(&m_ready[i], 8); mbar_init(&hid_ready[i], 8); mbar_init(&hid_ready[4 + i], 8); }
        mbar_init(e4_done, 16);
        mbar_init(xr_full, 1);
        mbar_init(x0_done, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2sm<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) enc_stamp(p, 1);
    // Programmatic dependent launch: everything that reads what the previous kernels wrote (x, G, Ksum) waits here -- except the TMA
    // producer, which first puts the first weight chunks (they depend on nothing) in flight and waits just before its first x load.
    if (warp != 0) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 32) enc_stamp(p, 2);
    constexpr uint32_t R0 = 0, R1 = 256;
    const bool tma_e4 = p.e4.out_f32 == nullptr;   // the last layer also writes an fp32 copy: it keeps the register-path epilogue

    if (warp == 0) {
        // ===================================================================================== TMA producer (one thread per CTA)
        if (elect_one()) {
            uint32_t it = 0;   // ring items issued so far
            uint32_t tp = 0;   // parity of the per-tile barriers
            bool first = true;
            auto ring_load = [&](const CUtensorMap* m, int col, int row) {
                const uint32_t s = it % kEncRingStages, ph = (it / kEncRingStages) & 1u;
                mbar_wait(&r_empty[s], ph ^ 1u);
                if (rank == 0) mbar_arrive_expect_tx(&r_full[s], 2 * kEncChunk);
                uint8_t* st = ring + s * kEncChunk;
                tma_load_3d_2sm(st, m, &r_full[s], col, row, 0);
                tma_load_3d_2sm(st + 16384, m, &r_full[s], col, row, 1);
                ++it;
            };
            for (int c = 0; c < 3; ++c) ring_load(&maps.wq, c * 64, static_cast<int>(rank) * 128);   // ahead of the dependency wait
            asm volatile("griddepcontrol.wait;" ::: "memory");
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m0 = (tile * 2 + static_cast<int>(rank)) * kBM;
                const int seg = (p.seg_tile0 > 0 && tile >= p.seg_tile0) ? 1 : 0;
                const int wrow = static_cast<int>(rank) * 128;
                // x tile -> act (free once the previous tile's last epilogue is done with it: residual tile / store staging),
                // interleaved with the Wq chunks so that GEMM1 can start early
                if (!first) mbar_wait(e4_done, tp ^ 1u);
                for (int c = 0; c < 4; ++c) {
                    if (rank == 0) mbar_arrive_expect_tx(&x_full[c], 2 * kEncChunk);
                    tma_load_3d_2sm(act + c * kEncChunk, &maps.x, &x_full[c], c * 64, m0, 0);
                    tma_load_3d_2sm(act + c * kEncChunk + 16384, &maps.x, &x_full[c], c * 64, m0, 1);
                    if (!(first && c < 3)) ring_load(&maps.wq, c * 64, wrow);
                }
                for (int i = 0; i < 4; ++i) ring_load(&maps.g, chunk_order(i) * 64, seg * 256 + wrow);
                for (int j = 0; j < 2; ++j) {
                    for (int c = 0; c < 4; ++c) {
                        if (j == 1) ring_load(&maps.x, c * 64, m0);   // chunk 0's x part runs while x is still resident in act
                        ring_load(&maps.w0, c * 64, j * 256 + wrow);
                    }
                    for (int i = 0; i < 4; ++i) ring_load(&maps.w0, 256 + chunk_order(i) * 64, j * 256 + wrow);
                    for (int i = 0; i < 4; ++i) ring_load(&maps.w2, j * 256 + chunk_order(i) * 64, wrow);
                }
                if (tma_e4) {
                    // the residual: this CTA's x rows once more (L2), into act as soon as mlp.0 is done reading m from it; the last
                    // epilogue adds LayerNorm2 on top IN PLACE and the tile leaves through a TMA store
                    mbar_wait(g3m_done, tp);
                    mbar_arrive_expect_tx(xr_full, kEncAct);
                    for (int c = 0; c < 4; ++c) {
                        tma_load_3d(act + c * kEncChunk, &maps.x, xr_full, c * 64, m0, 0);
                        tma_load_3d(act + c * kEncChunk + 16384, &maps.x, xr_full, c * 64, m0, 1);
                    }
                }
                first = false;
                tp ^= 1u;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================================================== MMA issuer (leader CTA, one thread)
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_f16(256, 256);
            uint32_t it = 0, tp = 0;
            bool first = true;
            const uint32_t ring_a = smem_u32(ring), act_a = smem_u32(act);
            auto ring_wait = [&]() -> uint32_t {   // -> shared address of the stage holding the next item
                const uint32_t s = it % kEncRingStages, ph = (it / kEncRingStages) & 1u;
                mbar_wait(&r_full[s], ph);
                tc_fence_after();
                ++it;
                return ring_a + s * kEncChunk;
            };
            auto ring_release = [&](uint32_t stage_addr) { umma_commit_2sm(&r_empty[(stage_addr - ring_a) / kEncChunk]); };
            // one 64-wide K chunk, both operands in shared memory: 4 K steps x (hi*hi, hi*lo, lo*hi)
            // (descriptors: one per operand and chunk, plus a constant in the 16-byte address field per K step / lo plane)
            auto mma_ss = [&](uint32_t d, uint32_t a, uint32_t b, bool fresh) {
                const uint64_t da0 = make_smem_desc_sw128(a), db0 = make_smem_desc_sw128(b);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t da = da0 + 2u * k, db = db0 + 2u * k;
                    umma_f16_2sm(tmem_base + d, da, db, idesc, (fresh && k == 0) ? 0u : 1u);
                    umma_f16_2sm(tmem_base + d, da, db + (16384 >> 4), idesc, 1u);
                    umma_f16_2sm(tmem_base + d, da + (16384 >> 4), db, idesc, 1u);
                }
            };
            // the same with the A operand in tensor memory: chunk c of a 256-wide packed activation at TMEM columns a0 ..
            auto mma_ts = [&](uint32_t d, uint32_t a0, int c, uint32_t b, bool fresh) {
                const uint64_t db0 = make_smem_desc_sw128(b);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t blk = tmem_base + a0 + 32u * static_cast<uint32_t>(2 * c + (k >> 1)) + 8u * static_cast<uint32_t>(k & 1);
                    const uint64_t db = db0 + 2u * k;
                    umma_f16_2sm_ts(tmem_base + d, blk, db, idesc, (fresh && k == 0) ? 0u : 1u);
                    umma_f16_2sm_ts(tmem_base + d, blk, db + (16384 >> 4), idesc, 1u);
                    umma_f16_2sm_ts(tmem_base + d, blk + 16u, db, idesc, 1u);
                }
            };
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                // ---- GEMM1: acc_q (R0) = x . Wq^T        (R0 held relu(hid) of the previous tile until its last GEMM4 finished)
                if (!first) { mbar_wait(out_done, tp ^ 1u); tc_fence_after(); }
                for (int c = 0; c < 4; ++c) {
                    mbar_wait(&x_full[c], tp);
                    const uint32_t b = ring_wait();
                    if (first && c == 0) enc_stamp(p, 3);
                    mma_ss(R0, act_a + c * kEncChunk, b, c == 0);
                    ring_release(b);
                }
                umma_commit_2sm(q_done);
                if (first) enc_stamp(p, 4);
                // ---- GEMM2: acc_mg (R1) = (Q*Z)[TMEM R0] . G^T   (R1 = acc_out of the previous tile until its last epilogue has read it)
                if (!first) mbar_wait(e4_done, tp ^ 1u);
                for (int i = 0; i < 4; ++i) {
                    const int c = chunk_order(i);
                    mbar_wait(&qp_ready[c], tp);
                    tc_fence_after();
                    const uint32_t b = ring_wait();
                    mma_ts(R1, R0, c, b, i == 0);
                    ring_release(b);
                }
                umma_commit_2sm(mg_done);
                // ---- mlp.0, x part of hidden chunk 0, while x is still resident in act: it runs under the LayerNorm1 epilogue, which
                //      waits for x0_done before it overwrites act with m.  R0 is the A operand of GEMM2 until that has completed.
                mbar_wait(mg_done, tp);
                tc_fence_after();
                for (int c = 0; c < 4; ++c) {
                    const uint32_t b = ring_wait();
                    mma_ss(R0, act_a + c * kEncChunk, b, c == 0);
                    ring_release(b);
                }
                umma_commit_2sm(x0_done);
                // ---- mlp: two 256-wide chunks of the hidden layer
                for (int j = 0; j < 2; ++j) {
                    if (j == 1) {
                        mbar_wait(g4_done, tp);                                  // relu(hid chunk 0) in R0 is no longer being read
                        tc_fence_after();
                        for (int c = 0; c < 4; ++c) {                            // x part of chunk 1 (x re-streamed through the ring)
                            const uint32_t a = ring_wait();
                            const uint32_t b = ring_wait();
                            mma_ss(R0, a, b, c == 0);
                            ring_release(a);
                            ring_release(b);
                        }
                    }
                    for (int i = 0; i < 4; ++i) {                                // m part (m resident in act, chunk by chunk as LayerNorm1 writes it)
                        const int c = chunk_order(i);
                        if (j == 0) { mbar_wait(&m_ready[c], tp); tc_fence_after(); }
                        const uint32_t b = ring_wait();
                        mma_ss(R0, act_a + c * kEncChunk, b, false);
                        ring_release(b);
                    }
                    if (j == 1) umma_commit_2sm(g3m_done);                       // act is free for the next tile's x
                    umma_commit_2sm(&h_done[j]);
                    for (int i = 0; i < 4; ++i) {                                // mlp.2 partial: acc_out (R1) += relu(hid_j)[TMEM R0] . W2_j^T
                        const int c = chunk_order(i);
                        mbar_wait(&hid_ready[4 * j + c], tp);
                        tc_fence_after();
                        const uint32_t b = ring_wait();
                        mma_ts(R1, R0, c, b, j == 0 && i == 0);
                        ring_release(b);
                    }
                    if (j == 0) umma_commit_2sm(g4_done);
                }
                umma_commit_2sm(out_done);
                first = false;
                tp ^= 1u;
            }
        }
        __syncwarp();
    } else {
        // ===================================================================================== epilogue warps
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const uint32_t tw = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const int r = quad * 32 + lane;        // row inside this CTA's 128-row tile
        const int cb = half * 128;             // this warp's columns of a 256-wide region
        uint32_t tp = 0;
        float* lnx_own = lnx + (warp - 2) * 64;
        float* lnx_partner = lnx + ((warp - 2) ^ 4) * 64;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const int seg = (p.seg_tile0 > 0 && tile >= p.seg_tile0) ? 1 : 0;
            const int row0 = (tile * 2 + static_cast<int>(rank)) * kBM + quad * 32;
            // ---- E1: q -> Q*Z, packed in place as the A operand of the merge GEMM (four heads per warp)
            const bool stamp = warp == 2 && lane == 0 && tile == cluster_id;
            mbar_wait(q_done, tp);
            tc_fence_after();
            if (stamp) enc_stamp(p, 5);
            {
                float va[128];   // all four heads of this warp in one TMEM round trip
#pragma unroll
                for (int b = 0; b < 4; ++b) tmem_ld32(tw + R0 + cb + 32 * b, va + 32 * b);
                tmem_ld_wait();
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    float* v = va + 32 * b;
                    const float* ks = p.ksum + seg * 256 + cb + 32 * b;
                    float dot = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 k4 = __ldg(reinterpret_cast<const float4*>(ks + j));
                        v[j] = v[j] > 0.f ? v[j] + 1.f : fast_ex2(v[j] * 1.4426950408889634f);          // elu(x) + 1
                        v[j + 1] = v[j + 1] > 0.f ? v[j + 1] + 1.f : fast_ex2(v[j + 1] * 1.4426950408889634f);
                        v[j + 2] = v[j + 2] > 0.f ? v[j + 2] + 1.f : fast_ex2(v[j + 2] * 1.4426950408889634f);
                        v[j + 3] = v[j + 3] > 0.f ? v[j + 3] + 1.f : fast_ex2(v[j + 3] * 1.4426950408889634f);
                        dot = fmaf(v[j], k4.x, dot); dot = fmaf(v[j + 1], k4.y, dot); dot = fmaf(v[j + 2], k4.z, dot); dot = fmaf(v[j + 3], k4.w, dot);
                    }
                    const float z = p.qz_scale[seg] / (dot + 1e-6f);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= z;
                    uint32_t pk[32];
                    pack_block_hl(v, pk);
                    tmem_st32(tw + R0 + cb + 32 * b, pk);
                    if (b & 1) {   // one 64-column K chunk of the merge GEMM is complete for this warp's rows
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote(&qp_ready[2 * half + (b >> 1)], 0);
                    }
                }
            }
            if (stamp) enc_stamp(p, 6);
            // ---- E2: LayerNorm1(acc_mg) -> m, written to act as the (hi, lo) A operand of mlp.0's second half
            mbar_wait(mg_done, tp);
            tc_fence_after();
            if (stamp) enc_stamp(p, 7);
            {
                float v[128];
#pragma unroll
                for (int b = 0; b < 4; ++b) tmem_ld32(tw + R1 + cb + 32 * b, v + 32 * b);
                tmem_ld_wait();
                const float pivot = v[0];
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int j = 0; j < 128; ++j) { const float d = v[j] - pivot; s1 += d; s2 = fmaf(d, d, s2); }
                const float m_own = pivot + s1 * (1.f / 128.f);
                const float q_own = fmaxf(s2 - s1 * s1 * (1.f / 128.f), 0.f);
                *reinterpret_cast<float2*>(lnx_own + 2 * lane) = make_float2(m_own, q_own);
                named_bar_sync(1 + quad, 64);
                const float2 o = *reinterpret_cast<const float2*>(lnx_partner + 2 * lane);
                named_bar_sync(1 + quad, 64);
                const float mean = 0.5f * (m_own + o.x);
                const float dm = m_own - o.x;
                const float var = (q_own + o.y + 64.f * dm * dm) * (1.f / 256.f);   // Chan: n_a n_b / n * dm^2 = 64 dm^2
                const float scale = rsqrtf(var + 1e-5f), shift = -mean * scale;
                mbar_wait(x0_done, tp);   // the tensor core has finished reading x from act (x part of mlp.0 chunk 0)
#pragma unroll
                for (int j8 = 0; j8 < 16; ++j8) {   // 8 columns = one 16-byte unit of the operand row
                    const int col = cb + 8 * j8;
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln1_g + col)), g1 = __ldg(reinterpret_cast<const float4*>(p.ln1_g + col + 4));
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ln1_b + col)), b1 = __ldg(reinterpret_cast<const float4*>(p.ln1_b + col + 4));
                    const float* w = v + 8 * j8;
                    const float y0 = fmaf(fmaf(w[0], scale, shift), g0.x, b0.x), y1 = fmaf(fmaf(w[1], scale, shift), g0.y, b0.y);
                    const float y2 = fmaf(fmaf(w[2], scale, shift), g0.z, b0.z), y3 = fmaf(fmaf(w[3], scale, shift), g0.w, b0.w);
                    const float y4 = fmaf(fmaf(w[4], scale, shift), g1.x, b1.x), y5 = fmaf(fmaf(w[5], scale, shift), g1.y, b1.y);
                    const float y6 = fmaf(fmaf(w[6], scale, shift), g1.z, b1.z), y7 = fmaf(fmaf(w[7], scale, shift), g1.w, b1.w);
                    uint4 uh, ul;
                    split_f16x2(y0, y1, uh.x, ul.x); split_f16x2(y2, y3, uh.y, ul.y);
                    split_f16x2(y4, y5, uh.z, ul.z); split_f16x2(y6, y7, uh.w, ul.w);
                    uint8_t* chunk = act + (col >> 6) * kEncChunk + r * 128;
                    const int phys = ((((col & 63) >> 3)) ^ (r & 7)) << 4;   // 128-byte swizzle: 16-byte unit index XOR (row mod 8)
                    *reinterpret_cast<uint4*>(chunk + phys) = uh;
                    *reinterpret_cast<uint4*>(chunk + 16384 + phys) = ul;
                    if ((j8 & 7) == 7) {   // one 64-column chunk of m is complete for this warp's rows
                        fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote(&m_ready[2 * half + (j8 >> 3)], 0);
                    }
                }
            }
            if (stamp) enc_stamp(p, 8);
            // ---- E3: relu(hid chunk) packed in place as the A operand of mlp.2
#pragma unroll 1
            for (int j = 0; j < 2; ++j) {
                mbar_wait(&h_done[j], tp);
                tc_fence_after();
                if (stamp && j == 0) enc_stamp(p, 9);
                {
                    float va[128];
#pragma unroll
                    for (int b = 0; b < 4; ++b) tmem_ld32(tw + R0 + cb + 32 * b, va + 32 * b);
                    tmem_ld_wait();
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        float* v = va + 32 * b;
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                        uint32_t pk[32];
                        pack_block_hl(v, pk);
                        tmem_st32(tw + R0 + cb + 32 * b, pk);
                        if (b & 1) {
                            tmem_st_wait();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_remote(&hid_ready[4 * j + 2 * half + (b >> 1)], 0);
                        }
                    }
                }
                if (stamp && j == 1) enc_stamp(p, 10);
            }
            // ---- E4: x + LayerNorm2(acc_out) -> HBM
            mbar_wait(out_done, tp);
            tc_fence_after();
            if (stamp) enc_stamp(p, 12);
            if (tma_e4) {
                // residual tile (hi, lo) sits in act in the operand layout; normalise, add, split and write back in place, then one
                // thread hands the 4 x 2 boxes to the TMA (no global load / store instruction in this epilogue)
                float v[128];
#pragma unroll
                for (int b = 0; b < 4; ++b) tmem_ld32(tw + R1 + cb + 32 * b, v + 32 * b);
                tmem_ld_wait();
                const float pivot = v[0];
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int j = 0; j < 128; ++j) { const float d = v[j] - pivot; s1 += d; s2 = fmaf(d, d, s2); }
                const float m_own = pivot + s1 * (1.f / 128.f);
                const float q_own = fmaxf(s2 - s1 * s1 * (1.f / 128.f), 0.f);
                *reinterpret_cast<float2*>(lnx_own + 2 * lane) = make_float2(m_own, q_own);
                named_bar_sync(1 + quad, 64);
                const float2 o = *reinterpret_cast<const float2*>(lnx_partner + 2 * lane);
                named_bar_sync(1 + quad, 64);
                const float mean = 0.5f * (m_own + o.x);
                const float dm = m_own - o.x;
                const float var = (q_own + o.y + 64.f * dm * dm) * (1.f / 256.f);
                const float scale = rsqrtf(var + 1e-5f), shift = -mean * scale;
                mbar_wait(xr_full, tp);
#pragma unroll
                for (int j8 = 0; j8 < 16; ++j8) {
                    const int col = cb + 8 * j8;
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.e4.gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(p.e4.gamma + col + 4));
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.e4.beta + col)), b1 = __ldg(reinterpret_cast<const float4*>(p.e4.beta + col + 4));
                    uint8_t* unit = act + (col >> 6) * kEncChunk + r * 128 + (((((col & 63) >> 3)) ^ (r & 7)) << 4);
                    const uint4 rh = *reinterpret_cast<const uint4*>(unit), rl = *reinterpret_cast<const uint4*>(unit + 16384);
                    const float* w = v + 8 * j8;
                    float y0 = fmaf(fmaf(w[0], scale, shift), g0.x, b0.x), y1 = fmaf(fmaf(w[1], scale, shift), g0.y, b0.y);
                    float y2 = fmaf(fmaf(w[2], scale, shift), g0.z, b0.z), y3 = fmaf(fmaf(w[3], scale, shift), g0.w, b0.w);
                    float y4 = fmaf(fmaf(w[4], scale, shift), g1.x, b1.x), y5 = fmaf(fmaf(w[5], scale, shift), g1.y, b1.y);
                    float y6 = fmaf(fmaf(w[6], scale, shift), g1.z, b1.z), y7 = fmaf(fmaf(w[7], scale, shift), g1.w, b1.w);
                    add_f16x2(y0, y1, rl.x); add_f16x2(y2, y3, rl.y); add_f16x2(y4, y5, rl.z); add_f16x2(y6, y7, rl.w);   // same order as the
                    add_f16x2(y0, y1, rh.x); add_f16x2(y2, y3, rh.y); add_f16x2(y4, y5, rh.z); add_f16x2(y6, y7, rh.w);   // register-path epilogue
                    uint4 uh, ul;
                    split_f16x2(y0, y1, uh.x, ul.x); split_f16x2(y2, y3, uh.y, ul.y);
                    split_f16x2(y4, y5, uh.z, ul.z); split_f16x2(y6, y7, uh.w, ul.w);
                    *reinterpret_cast<uint4*>(unit) = uh;
                    *reinterpret_cast<uint4*>(unit + 16384) = ul;
                }
                fence_proxy_async();
                tc_fence_before();
                named_bar_sync(6, 32 * 8);   // the eight epilogue warps of this CTA
                if (warp == 2 && lane == 0) {
                    const int m0 = (tile * 2 + static_cast<int>(rank)) * kBM;
                    for (int c = 0; c < 4; ++c) {
                        tma_store_3d(&maps.x, act + c * kEncChunk, c * 64, m0, 0);
                        tma_store_3d(&maps.x, act + c * kEncChunk + 16384, c * 64, m0, 1);
                    }
                    tma_store_commit();
                    tma_store_wait_read();   // act may be refilled (next tile's x) once the stores have read it
                }
            } else {
                EpiCtx ctx;
                ctx.stg = act + 3 * kEncChunk + (warp - 2) * 4096;
                ctx.stg_partner = act + 3 * kEncChunk + ((warp - 2) ^ 4) * 4096;
                ctx.bar_id = 1 + quad;
                LinEpi::run_ln_staged<256, 2>(p.e4, tw + R1, row0, lane, 0, cb, ctx);
                tc_fence_before();
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive_remote(e4_done, 0); mbar_arrive_remote(e4_done, 1); }
            if (stamp) enc_stamp(p, 13);
            if (warp == 2 && lane == 0) enc_stamp(p, 14);
            tp ^= 1u;
        }
    }
    if (warp == 2 && lane == 0) tma_store_wait_all();
    tc_fence_before();
    cluster_sync_all();
    if (threadIdx.x == 0) enc_stamp(p, 11);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<512>(tmem_base);
    }
}

}  // namespace dfsfm
