// HP-2: multi-view track refinement matcher on the B200 -- chunk schedule and C ABI.
// Reference: src/MultiviewMatcher/MultiviewMatcher.py:59-405 (forward, chunk_backbone_img path), backbone/S2DNet/s2dnet.py:127-193,
// matcher_module/transformer.py:132-177, utils/fine_matching.py:36-284, third_party/RoIAlign.pytorch (crop_and_resize).
//
// Patches are laid out track-major: [track 0: reference node, valid query views...][track 1: ...]; the reference's per-image
// backbone loop and its un-permute (MultiviewMatcher.py:188-279) disappear because every kernel takes a per-patch image pointer.
#include <cmath>
#include <memory>

#include "../../include/dfsfm_b200.h"
#include "engine_common.h"
#include "refine_kernels.cuh"

namespace dfsfm {

namespace {
constexpr int kRefineKvTok = 256;
struct PGeom {  // per-patch flat geometry
    int H, W, Hp, Wp;
    int rows() const { return Hp * Wp; }
    FlatGeom flat() const { return FlatGeom{Hp, Wp, H, W}; }
};
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    void ensure(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        cap = n + n / 8;
        DFSFM_CUDA(cudaMalloc(&p, cap * sizeof(T)));
        zero_device_sync(p, 0, cap * sizeof(T));
    }
    ~DevBuf() { if (p) cudaFree(p); }
};
struct HLBuf {  // grow-only split-fp16 buffer; re-zeroed on growth so halo cells stay zero
    HL b;
    long long cap_rows = 0;
    void ensure(long long rows, int C) {
        if (rows <= cap_rows && b.C == C) { b.rows = cap_rows; return; }
        hl_free(b);
        cap_rows = rows + rows / 8;
        b = hl_alloc(cap_rows, C);
    }
    ~HLBuf() { hl_free(b); }
};
}  // namespace

class RefineEngine {
  public:
    RefineEngine(int device, int window, int left_window) : W_(window), LW_(left_window) {
        DFSFM_CUDA(cudaSetDevice(device));
        DFSFM_CUDA(cudaFuncSetAttribute(attn_apply_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes<16>()));
        DFSFM_CHECK(window % 2 == 1 && window >= 7 && window <= 15, "window must be odd, 7..15");
        DFSFM_CHECK(left_window % 2 == 1 && left_window >= 1 && left_window <= window, "left window must be odd and <= window");
        g35_ = PGeom{35, 35, 36, 36};
        g18_ = PGeom{18, 18, 19, 19};
        g9_ = PGeom{9, 9, 10, 10};
        const int wa = W_ + 4;  // adap0: the 5x5 conv needs a 2-px ring around the centre window
        ga0_ = PGeom{wa, wa, wa + 1, wa + 1};
        ga1_ = PGeom{9, 9, 11, 11};  // halo 2 for the 5x5 conv with exact zero padding
        build_bicubic();
    }
    ParamStore params;
    void chunk(int n_img, const float* const* images, const int32_t* H, const int32_t* W, const float* scales_hw, int M, int Nq,
               const float* query_pts, const float* ref_pts, const uint8_t* valid, const int32_t* q_img_idx, const int32_t* r_img_idx,
               const uint8_t* movable, float* query_refined, float* ref_refined, float* std_out, cudaStream_t st);

  private:
    int W_, LW_;
    PGeom g35_, g18_, g9_, ga0_, ga1_;
    BicubicTab tab_;
    HLBuf c11_, c12_, p1_, c21_, c22_, p2_, c31_, c32_, c33_, a0_, a1_;
    HLBuf x_, msg_, m1_, hid_;
    DevBuf<float> a0out_, a1out_, xf_, qkv_, kvstate_, kvpart_, d_query_, d_ref_, d_std_;
    DevBuf<PatchRec> d_recs_;
    DevBuf<Seg> d_segs_;
    DevBuf<TrackRec> d_tracks_;
    DevBuf<ViewRec> d_views_;

    void build_bicubic();
    template <int BN>
    void conv(const HL& in, const PGeom& g, long long P, int taps_k, int cpad, const std::string& wname, ConvEpiParams ep, cudaStream_t st);
    ConvEpiParams epi(const PGeom& g, int N, bool relu) const {
        ConvEpiParams e;
        memset(&e, 0, sizeof(e));
        e.N = N;
        e.g = g.flat();
        e.relu = relu ? 1 : 0;
        e.out_mode = OUT_FLAT;
        return e;
    }
    static void set_out(ConvEpiParams& e, const HL& out) { e.out_hi = out.hi; e.out_lo = out.lo(); e.out_ld = out.C; }
    void transformer(long long T, int n_segs, int max_count, cudaStream_t st);
};

// torch.nn.Upsample(size=35, mode='bicubic', align_corners=True) source taps for the destination coordinates of the centre
// window (ATen UpSampleBicubic2d: scale = (in-1)/(out-1), A = -0.75, indices clamped to [0, in-1]).
void RefineEngine::build_bicubic() {
    const float A = -0.75f;
    const float scale = static_cast<float>(9 - 1) / static_cast<float>(kCrop - 1);
    const int off = kCrop / 2 - W_ / 2;
    for (int i = 0; i < W_; ++i) {
        const float real = scale * static_cast<float>(off + i);
        const int ix = static_cast<int>(floorf(real));
        const float t = real - static_cast<float>(ix);
        auto c1 = [&](float x) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; };
        auto c2 = [&](float x) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; };
        const float w[4] = {c2(t + 1.f), c1(t), c1(1.f - t), c2(2.f - t)};
        for (int k = 0; k < 4; ++k) {
            int id = ix - 1 + k;
            id = id < 0 ? 0 : (id > 8 ? 8 : id);
            tab_.idx[i][k] = id;
            tab_.w[i][k] = w[k];
        }
    }
    for (int i = W_; i < 15; ++i)
        for (int k = 0; k < 4; ++k) { tab_.idx[i][k] = 0; tab_.w[i][k] = 0.f; }
}

template <int BN>
void RefineEngine::conv(const HL& in, const PGeom& g, long long P, int taps_k, int cpad, const std::string& wname, ConvEpiParams ep,
                        cudaStream_t st) {
    const HL& w = params.mat(wname + ".w");
    GemmCore c;
    memset(&c, 0, sizeof(c));
    const long long rows = P * g.rows();
    DFSFM_CHECK(rows < (1ll << 31), "chunk too large for 32-bit row indices");
    c.M = static_cast<int>(rows);
    set_k(c, cpad);
    conv_taps_s1(c, taps_k, g.Wp);
    DFSFM_CHECK(static_cast<long long>(c.num_taps) * cpad == w.C, "weight K mismatch for " + wname);
    TmapPack maps;
    const bool slab = BN <= 128 && slab_applicable(c, 1);
    // the tensor map covers exactly the rows in use (OOB rows read as zero)
    CUtensorMap am = make_tmap(in.hi, in.C, rows, in.plane_elems(), slab ? kSlabRows : kBM);
    for (int i = 0; i < kMaxAMaps; ++i) maps.a[i] = am;
    maps.b = make_tmap(w, bbox(BN));
    ep.M = c.M;
    ep.bias = params.has_vec(wname + ".b") ? params.vec(wname + ".b") : nullptr;
    if constexpr (BN <= 128) {
        if (slab) {
            launch_gemm_slab_counted<BN, ConvEpi>(maps, c, ep, ep.N, st, "pconv");
            return;
        }
    }
    launch_gemm_counted<BN, true, ConvEpi>(maps, c, ep, ep.N, st, "pconv");
}

// 4-layer multiview transformer on the flat token array (token-wise linears over all patches; attention per segment).
void RefineEngine::transformer(long long T, int n_segs, int max_count, cudaStream_t st) {
    DFSFM_CHECK(T < (1ll << 31), "too many tokens");
    const int Ti = static_cast<int>(T);
    const Seg* segs_self = d_segs_.p;
    const Seg* segs_cross = d_segs_.p + n_segs;
    GemmCore c;
    memset(&c, 0, sizeof(c));
    conv_taps_s1(c, 1, 0);
    c.M = Ti;
    auto amap = [&](const HL& b) { return make_tmap(b.hi, b.C, T, b.plane_elems(), kBM); };
    for (int li = 0; li < 4; ++li) {
        const std::string p = "tr." + std::to_string(li);
        const bool self = (li % 2) == 0;  // layer_names ['self','cross'] * 2 (yaml:57-58)
        LinEpiParams e;
        TmapPack maps;
        {   // q,k,v for every token (+ elu+1 on q,k)
            for (int i = 0; i < kMaxAMaps; ++i) maps.a[i] = amap(x_.b);
            maps.b = make_tmap(params.mat(p + ".qkv"), bbox(128));
            set_k(c, 128);
            memset(&e, 0, sizeof(e));
            e.M = Ti; e.N = 384; e.mode = LIN_F32_ELU; e.elu_cols = 256; e.out_f32 = qkv_.p; e.out_f32_ld = 384;
            launch_gemm_counted<128, true, LinEpi>(maps, c, e, 384, st, "lin");
        }
        // KV state of every segment: 256-token partials, then a fixed-order sum (deterministic)
        {
            const int chunks = (max_count + kRefineKvTok - 1) / kRefineKvTok;
            { LaunchScope ls("kv", st);
              kv_partial_kernel<16><<<dim3(chunks, n_segs), 128, 0, st>>>(qkv_.p + 128, qkv_.p + 256, 384, segs_self, chunks, kvpart_.p, kRefineKvTok); }
            { LaunchScope ls("kv", st);
              kv_final_kernel<16><<<dim3((8 * 16 * 17 + 63) / 64, n_segs), kKvFinalThreads, 0, st>>>(kvpart_.p, segs_self, chunks, kvstate_.p, kRefineKvTok); }
        }
        // self: a segment reads its own state; cross: its partner's -- both directions use the PRE-update tokens
        // (matcher_module/transformer.py:162-167), which is what a single q/k/v pass over the old tokens gives.
        { LaunchScope ls("attn", st);
          attn_apply_kernel<16><<<dim3((max_count + kAttnTokRefine - 1) / kAttnTokRefine, n_segs), 256, attn_smem_bytes<16>(), st>>>(qkv_.p, 384, self ? segs_self : segs_cross, kvstate_.p,
                                                                                  msg_.b.hi, msg_.b.lo(), 128, kAttnTokRefine); }
        DFSFM_CUDA(cudaGetLastError());
        {   // merge + norm1
            for (int i = 0; i < kMaxAMaps; ++i) maps.a[i] = amap(msg_.b);
            maps.b = make_tmap(params.mat(p + ".merge"), bbox(128));
            set_k(c, 128);
            memset(&e, 0, sizeof(e));
            e.M = Ti; e.N = 128; e.mode = LIN_LN; e.gamma = params.vec(p + ".ln1.g"); e.beta = params.vec(p + ".ln1.b");
            e.out_hi = m1_.b.hi; e.out_lo = m1_.b.lo(); e.out_ld = 128;
            launch_gemm_counted<128, true, LinEpi>(maps, c, e, 128, st, "lin");
        }
        if (fused_mlp_enabled()) {  // mlp.0 + relu + mlp.2 + norm2 + residual in one kernel: hid stays on the SM
            launch_mlp128_fused(x_.b, m1_.b, 0, T, params.mat(p + ".mlp0"), params.mat(p + ".mlp2"), params.vec(p + ".ln2.g"),
                                params.vec(p + ".ln2.b"), xf_.p, st);
            continue;
        }
        {   // mlp.0 on cat[x, message] + relu
            for (int i = 0; i < kMaxAMaps; ++i) maps.a[i] = amap(i == 1 ? m1_.b : x_.b);
            maps.b = make_tmap(params.mat(p + ".mlp0"), bbox(256));
            GemmCore c2 = c;
            set_k(c2, 128);
            c2.num_taps = 2; c2.tap_map[0] = 0; c2.tap_map[1] = 1; c2.tap_shift[0] = c2.tap_shift[1] = 0;
            memset(&e, 0, sizeof(e));
            e.M = Ti; e.N = 256; e.mode = LIN_RELU_HL; e.out_hi = hid_.b.hi; e.out_lo = hid_.b.lo(); e.out_ld = 256;
            launch_gemm_counted<256, true, LinEpi>(maps, c2, e, 256, st, "lin");
        }
        {   // mlp.2 + norm2 + residual
            for (int i = 0; i < kMaxAMaps; ++i) maps.a[i] = amap(hid_.b);
            maps.b = make_tmap(params.mat(p + ".mlp2"), bbox(128));
            GemmCore c3 = c;
            set_k(c3, 256);
            memset(&e, 0, sizeof(e));
            e.M = Ti; e.N = 128; e.mode = LIN_LN; e.gamma = params.vec(p + ".ln2.g"); e.beta = params.vec(p + ".ln2.b");
            e.resid = xf_.p; e.resid_ld = 128; e.out_f32 = xf_.p; e.out_f32_ld = 128;
            e.out_hi = x_.b.hi; e.out_lo = x_.b.lo(); e.out_ld = 128;
            launch_gemm_counted<128, true, LinEpi>(maps, c3, e, 128, st, "lin");
        }
    }
}

void RefineEngine::chunk(int n_img, const float* const* images, const int32_t* H, const int32_t* Wd, const float* scales_hw, int M, int Nq,
                         const float* query_pts, const float* ref_pts, const uint8_t* valid, const int32_t* q_img_idx,
                         const int32_t* r_img_idx, const uint8_t* movable, float* query_refined, float* ref_refined, float* std_out,
                         cudaStream_t st) {
    DFSFM_CHECK(M > 0 && Nq >= 1 && Nq <= kMaxViews && n_img > 0, "bad chunk shape");
    const int WW = W_ * W_;
    // ---- host prep: patch records (track-major), attention segments, per-track / per-view scalars
    std::vector<PatchRec> recs;
    std::vector<Seg> segs(static_cast<size_t>(4) * M);  // [self: ref_t, qry_t]*M then [cross: ...]*M
    std::vector<TrackRec> tracks(M);
    std::vector<ViewRec> views(static_cast<size_t>(Nq) * M);
    recs.reserve(static_cast<size_t>(M) * (Nq + 1));
    auto add_patch = [&](int img, float px, float py) {
        DFSFM_CHECK(img >= 0 && img < n_img, "image index out of range");
        // all_sample_points /= scales (MultiviewMatcher.py:103-106): scales = (w_ratio, h_ratio) per image
        const float sw = 1.0f * scales_hw[img * 2 + 1], sh = 1.0f * scales_hw[img * 2 + 0];
        const float x = px / sw, y = py / sh;
        PatchRec r;
        r.image = images[img];
        r.H = H[img];
        r.W = Wd[img];
        const float rad = static_cast<float>(kCrop / 2);
        // fine_preprocess.py:102-104 boxes = kp -+ radius; roi_align.py:39-44 normalisation by (W-1), (H-1)
        r.x1 = (x - rad) / static_cast<float>(r.W - 1);
        r.x2 = (x + rad) / static_cast<float>(r.W - 1);
        r.y1 = (y - rad) / static_cast<float>(r.H - 1);
        r.y2 = (y + rad) / static_cast<float>(r.H - 1);
        recs.push_back(r);
    };
    int max_count = WW;
    for (int t = 0; t < M; ++t) {
        const int p0 = static_cast<int>(recs.size());
        add_patch(q_img_idx[t], query_pts[t * 2 + 0], query_pts[t * 2 + 1]);
        int k = 0;
        for (int n = 0; n < Nq; ++n) {
            const bool v = valid[static_cast<size_t>(n) * M + t] != 0;
            if (v) {
                DFSFM_CHECK(k == n, "valid query views must form a prefix of the view axis (construct_matching_data.py:350-352)");
                const size_t o = (static_cast<size_t>(n) * M + t) * 2;
                add_patch(r_img_idx[static_cast<size_t>(n) * M + t], ref_pts[o], ref_pts[o + 1]);
                ++k;
            }
        }
        DFSFM_CHECK(k >= 1, "every track needs at least one valid query view");
        TrackRec& tr = tracks[t];
        tr.tok0 = p0 * WW;
        tr.qtok0 = (p0 + 1) * WW;
        tr.n_views = k;
        tr.movable = movable ? (movable[t] != 0) : 1;
        tr.qx = query_pts[t * 2 + 0];
        tr.qy = query_pts[t * 2 + 1];
        tr.sqx = 1.0f * scales_hw[q_img_idx[t] * 2 + 1];
        tr.sqy = 1.0f * scales_hw[q_img_idx[t] * 2 + 0];
        for (int n = 0; n < Nq; ++n) {
            ViewRec& vr = views[static_cast<size_t>(n) * M + t];
            const size_t o = (static_cast<size_t>(n) * M + t) * 2;
            int img = r_img_idx[static_cast<size_t>(n) * M + t];
            if (img < 0) img = n_img - 1;  // index -1 wraps to the last image in the reference (MultiviewMatcher.py:99-101)
            vr.rx = ref_pts[o];
            vr.ry = ref_pts[o + 1];
            vr.sx = 1.0f * scales_hw[img * 2 + 1];
            vr.sy = 1.0f * scales_hw[img * 2 + 0];
        }
        const Seg ref_seg = {p0 * WW, WW, WW, 2 * t};
        const Seg qry_seg = {(p0 + 1) * WW, k * WW, k * WW, 2 * t + 1};
        segs[2 * t] = ref_seg;
        segs[2 * t + 1] = qry_seg;
        Seg rc = ref_seg, qc = qry_seg;
        rc.state = 2 * t + 1;  // cross: the reference attends the query views' state and vice versa
        qc.state = 2 * t;
        segs[static_cast<size_t>(2) * M + 2 * t] = rc;
        segs[static_cast<size_t>(2) * M + 2 * t + 1] = qc;
        if (k * WW > max_count) max_count = k * WW;
    }
    const long long P = static_cast<long long>(recs.size());
    const long long T = P * WW;
    // ---- device buffers
    c11_.ensure(P * g35_.rows(), 64); c12_.ensure(P * g35_.rows(), 64);
    p1_.ensure(P * g18_.rows(), 64); c21_.ensure(P * g18_.rows(), 128); c22_.ensure(P * g18_.rows(), 128);
    p2_.ensure(P * g9_.rows(), 128); c31_.ensure(P * g9_.rows(), 256); c32_.ensure(P * g9_.rows(), 256); c33_.ensure(P * g9_.rows(), 256);
    a0_.ensure(P * ga0_.rows(), 64); a1_.ensure(P * ga1_.rows(), 64);
    x_.ensure(T, 128); msg_.ensure(T, 128); m1_.ensure(T, 128); hid_.ensure(T, 256);
    a0out_.ensure(static_cast<size_t>(T) * 128); a1out_.ensure(static_cast<size_t>(P) * ga1_.rows() * 128);
    xf_.ensure(static_cast<size_t>(T) * 128); qkv_.ensure(static_cast<size_t>(T) * 384);
    kvstate_.ensure(static_cast<size_t>(2) * M * 8 * 16 * 17);
    kvpart_.ensure(static_cast<size_t>(2) * M * ((max_count + kRefineKvTok - 1) / kRefineKvTok) * 8 * 16 * 17);
    d_query_.ensure(static_cast<size_t>(M) * 2); d_ref_.ensure(static_cast<size_t>(Nq) * M * 2); d_std_.ensure(static_cast<size_t>(Nq) * M);
    d_recs_.ensure(recs.size()); d_segs_.ensure(segs.size()); d_tracks_.ensure(tracks.size()); d_views_.ensure(views.size());
    DFSFM_CUDA(cudaMemcpyAsync(d_recs_.p, recs.data(), recs.size() * sizeof(PatchRec), cudaMemcpyHostToDevice, st));
    DFSFM_CUDA(cudaMemcpyAsync(d_segs_.p, segs.data(), segs.size() * sizeof(Seg), cudaMemcpyHostToDevice, st));
    DFSFM_CUDA(cudaMemcpyAsync(d_tracks_.p, tracks.data(), tracks.size() * sizeof(TrackRec), cudaMemcpyHostToDevice, st));
    DFSFM_CUDA(cudaMemcpyAsync(d_views_.p, views.data(), views.size() * sizeof(ViewRec), cudaMemcpyHostToDevice, st));
    DFSFM_CUDA(cudaMemsetAsync(d_ref_.p, 0, static_cast<size_t>(Nq) * M * 2 * sizeof(float), st));
    DFSFM_CUDA(cudaMemsetAsync(d_std_.p, 0, static_cast<size_t>(Nq) * M * sizeof(float), st));

    // ---- S2DNet on patches (s2dnet.py:127-175)
    { LaunchScope ls("patch_conv11", st);
      patch_conv11_kernel<<<static_cast<unsigned>(P), 256, 0, st>>>(d_recs_.p, params.vec("c11.w"), params.vec("c11.b"), c11_.b.hi, c11_.b.lo(),
                                                                 nullptr); }
    DFSFM_CUDA(cudaGetLastError());
    { ConvEpiParams e = epi(g35_, 64, true); set_out(e, c12_.b); conv<64>(c11_.b, g35_, P, 3, 64, "c12", e, st); }
    auto pool = [&](const HL& in, const PGeom& gi, const HL& out, const PGeom& go, int C) {
        const long long total = P * go.H * go.W * (C / 8);
        { LaunchScope ls("pool", st);
          maxpool3s2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in.hi, in.lo(), gi.H, gi.W, C, out.hi, out.lo(), go.H, go.W,
                                                                                     total); }
        DFSFM_CUDA(cudaGetLastError());
    };
    pool(c12_.b, g35_, p1_.b, g18_, 64);
    { ConvEpiParams e = epi(g18_, 128, true); set_out(e, c21_.b); conv<128>(p1_.b, g18_, P, 3, 64, "c21", e, st); }
    { ConvEpiParams e = epi(g18_, 128, true); set_out(e, c22_.b); conv<128>(c21_.b, g18_, P, 3, 128, "c22", e, st); }
    pool(c22_.b, g18_, p2_.b, g9_, 128);
    { ConvEpiParams e = epi(g9_, 256, true); set_out(e, c31_.b); conv<256>(p2_.b, g9_, P, 3, 128, "c31", e, st); }
    { ConvEpiParams e = epi(g9_, 256, true); set_out(e, c32_.b); conv<256>(c31_.b, g9_, P, 3, 256, "c32", e, st); }
    { ConvEpiParams e = epi(g9_, 256, true); set_out(e, c33_.b); conv<256>(c32_.b, g9_, P, 3, 256, "c33", e, st); }
    // adaptation layer 0 on relu1_2: 1x1 (+ReLU) only where the 5x5 needs it, then 5x5 + BN on the centre window
    {
        ConvEpiParams e = epi(g35_, 64, true);
        e.out_mode = OUT_WINDOW;
        e.wy0 = e.wx0 = kCrop / 2 - W_ / 2 - 2; e.wh = e.ww = W_ + 4; e.ohp = ga0_.Hp; e.owp = ga0_.Wp;
        set_out(e, a0_.b);
        conv<64>(c12_.b, g35_, P, 1, 64, "a0.0", e, st);
    }
    {
        ConvEpiParams e = epi(ga0_, 128, false);
        e.out_mode = OUT_WINDOW_DENSE;
        e.wy0 = e.wx0 = 2; e.wh = e.ww = W_;
        e.out_f32 = a0out_.p; e.out_f32_ld = 128;
        conv<128>(a0_.b, ga0_, P, 5, 64, "a0.2", e, st);
    }
    // adaptation layer 1 on relu3_3: 1x1 (+ReLU) re-packed with a 2-cell halo, then 5x5 + BN with exact zero padding
    {
        ConvEpiParams e = epi(g9_, 64, true);
        e.out_mode = OUT_WINDOW;
        e.wy0 = e.wx0 = 0; e.wh = e.ww = 9; e.ohp = ga1_.Hp; e.owp = ga1_.Wp;
        set_out(e, a1_.b);
        conv<64>(c33_.b, g9_, P, 1, 256, "a1.0", e, st);
    }
    {
        ConvEpiParams e = epi(ga1_, 128, false);
        e.out_f32 = a1out_.p; e.out_f32_ld = 128;
        conv<128>(a1_.b, ga1_, P, 5, 64, "a1.2", e, st);
    }
    { LaunchScope ls("bicubic", st);
      bicubic_merge_kernel<<<static_cast<unsigned>(P), 128, 0, st>>>(a0out_.p, a1out_.p, ga1_.Wp, tab_, W_, xf_.p, x_.b.hi, x_.b.lo()); }
    DFSFM_CUDA(cudaGetLastError());

    transformer(T, 2 * M, max_count, st);

    // ---- window correlation + soft-argmax + reference-point search
    {
        const int L = LW_ * LW_;
        const size_t smem = fine_match_smem_bytes(W_, LW_);
        static PerDeviceOnce once;
        once.run([&] { DFSFM_CUDA(cudaFuncSetAttribute(fine_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); });
        { LaunchScope ls("finematch", st);
          fine_match_kernel<<<M, kFmThreads, smem, st>>>(xf_.p, d_tracks_.p, d_views_.p, Nq, W_, LW_, d_query_.p, d_ref_.p, d_std_.p, M); }
        DFSFM_CUDA(cudaGetLastError());
    }
    // results: device pointers (the plugin path: they stay on the GPU, nothing synchronises here) or host pointers (copied back, stream
    // synchronised before returning)
    cudaPointerAttributes attr;
    bool dev_out = false;
    if (cudaPointerGetAttributes(&attr, query_refined) == cudaSuccess) dev_out = attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
    else cudaGetLastError();
    DFSFM_CUDA(cudaMemcpyAsync(query_refined, d_query_.p, static_cast<size_t>(M) * 2 * sizeof(float), cudaMemcpyDefault, st));
    DFSFM_CUDA(cudaMemcpyAsync(ref_refined, d_ref_.p, static_cast<size_t>(Nq) * M * 2 * sizeof(float), cudaMemcpyDefault, st));
    DFSFM_CUDA(cudaMemcpyAsync(std_out, d_std_.p, static_cast<size_t>(Nq) * M * sizeof(float), cudaMemcpyDefault, st));
    if (!dev_out) DFSFM_CUDA(cudaStreamSynchronize(st));
}

}  // namespace dfsfm

using dfsfm::RefineEngine;
struct dfsfm_refine { std::unique_ptr<RefineEngine> e; };

extern "C" {

int dfsfm_refine_create(dfsfm_refine_t** out, int device, int window, int left_window) {
    return dfsfm::guard([&] {
        auto* h = new dfsfm_refine;
        h->e.reset(new RefineEngine(device, window, left_window));
        *out = h;
    });
}
void dfsfm_refine_destroy(dfsfm_refine_t* h) { delete h; }
int dfsfm_refine_set_param(dfsfm_refine_t* h, const char* name, const float* host, int64_t rows, int64_t cols, int kind) {
    return dfsfm::guard([&] { h->e->params.set(name, host, rows, cols, kind); });
}
int dfsfm_refine_chunk(dfsfm_refine_t* h, int n_img, const float* const* images_dev, const int32_t* H, const int32_t* W, const float* scales_hw,
                       int M, int Nq, const float* query_pts, const float* ref_pts, const uint8_t* valid, const int32_t* q_img_idx,
                       const int32_t* r_img_idx, const uint8_t* movable, float* query_refined, float* ref_refined, float* std_out,
                       void* stream) {
    return dfsfm::guard([&] {
        h->e->chunk(n_img, images_dev, H, W, scales_hw, M, Nq, query_pts, ref_pts, valid, q_img_idx, r_img_idx, movable, query_refined,
                    ref_refined, std_out, static_cast<cudaStream_t>(stream));
    });
}

}  // extern "C"
