// Match -> keypoint -> index post-processing on the GPU.
//
// Replaces the single-threaded Python between the matcher and keypoints.h5 / matches.h5
// (src/coarse_match/coarse_match.py:203-237): Match2Kpts + keypoint_worker (np.unique / np.bincount / sorted),
// update_matches (dict look-ups per match) and transform_keypoints.  All integer / ordering work is bit-exact with the
// reference; the float64 score sums are accumulated in the reference's input order (a stable LSD radix sort keeps the
// observations of one key in match order, one thread then walks the run).
//
//   observation o = 2*t + side  (t = match row over all pairs, side = which image of the pair)
//   key(o) = image << (xb+yb) | trunc(x) << yb | trunc(y)                   -> sort #1 (stable) = np.unique order per image
//   unique u: run of equal keys; sum(u) = double sum of the run's confidences, in o order
//   sort #2 over uniques (stable): by ~bits(sum) (descending score), then by image -> rank r = global keypoint index
//   id(u) = r - first rank of u's image; match_ids[t][side] = id(u(o))
//
// Not reproduced: a pair of an image with itself (never produced by the pair generators) would accumulate its two sides
// row-interleaved here and side after side in the reference -- the same float64 sum unless it is inexact.
//
// HBM-bound integer work: every pass streams 12-byte (key, value) records; no tensor cores involved.
#include <algorithm>
#include <memory>

#include "../../include/dfsfm_b200.h"
#include "engine_common.h"

namespace dfsfm {

constexpr int kSeg = 2048;          // elements per warp segment of the radix sort
constexpr int kScanItems = 16;      // items per thread of the scan kernels
constexpr int kScanThreads = 256;
constexpr int kScanTile = kScanItems * kScanThreads;

// ------------------------------------------------------------------------------------------------ exclusive scan (int32)
static __global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int* __restrict__ in, long long n, int* __restrict__ block_sums) {
    __shared__ int wsum[kScanThreads / 32];
    const long long base = static_cast<long long>(blockIdx.x) * kScanTile + threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) s += (base + i < n) ? in[base + i] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += wsum[w];
        block_sums[blockIdx.x] = t;
    }
}
// one block: in-place exclusive scan of `n` block sums; total -> *total_out (may be null)
static __global__ void __launch_bounds__(1024) scan_sums_kernel(int* __restrict__ sums, int n, int* __restrict__ total_out) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? sums[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = wsum[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            wsum[threadIdx.x] = w;  // inclusive over warps
        }
        __syncthreads();
        const int warp_off = (threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0;
        const int carry = carry_s;
        if (i < n) sums[i] = carry + warp_off + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_off + x;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}
static __global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int* __restrict__ in, int* __restrict__ out, long long n,
                                                                         const int* __restrict__ block_offs) {
    __shared__ int wsum[kScanThreads / 32];
    const long long base = static_cast<long long>(blockIdx.x) * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
    int x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    int off = block_offs[blockIdx.x] + x - s;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) off += wsum[w];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = off;
        off += v[i];
    }
}

// ------------------------------------------------------------------------------------------------ LSD radix sort pass
// One warp owns a contiguous segment of kSeg records.  hist[digit][segment]; after the exclusive scan over that array (digit
// major) a segment knows where its records of each digit start, and scatters them in their original order -> stable.
static __global__ void __launch_bounds__(256) rs_hist_kernel(const unsigned long long* __restrict__ keys, long long n, int shift, int nseg,
                                                             int* __restrict__ hist) {
    __shared__ int h[8][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < 256; i += 32) h[warp][i] = 0;
    __syncwarp();
    const long long seg = static_cast<long long>(blockIdx.x) * 8 + warp;
    if (seg < nseg) {
        const long long b = seg * kSeg;
        const long long e = (b + kSeg < n) ? b + kSeg : n;
        for (long long i = b + lane; i < e; i += 32) atomicAdd(&h[warp][static_cast<int>((keys[i] >> shift) & 255ull)], 1);
        __syncwarp();
        for (int d = lane; d < 256; d += 32) hist[static_cast<long long>(d) * nseg + seg] = h[warp][d];
    }
}
static __global__ void __launch_bounds__(256) rs_scatter_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ vals,
                                                                long long n, int shift, int nseg, const int* __restrict__ offs,
                                                                unsigned long long* __restrict__ keys_out, unsigned int* __restrict__ vals_out) {
    __shared__ int pos[8][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long seg = static_cast<long long>(blockIdx.x) * 8 + warp;
    if (seg >= nseg) return;
    for (int d = lane; d < 256; d += 32) pos[warp][d] = offs[static_cast<long long>(d) * nseg + seg];
    __syncwarp();
    const long long b = seg * kSeg;
    const long long e = (b + kSeg < n) ? b + kSeg : n;
    const unsigned int lt = (1u << lane) - 1u;
    for (long long i0 = b; i0 < e; i0 += 32) {
        const long long i = i0 + lane;
        const bool valid = i < e;
        const unsigned int act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned long long k = keys[i];
            const int d = static_cast<int>((k >> shift) & 255ull);
            const unsigned int peers = __match_any_sync(act, d);
            const int base = pos[warp][d];
            const int rank = __popc(peers & lt);
            __syncwarp(act);
            if (rank == 0) pos[warp][d] = base + __popc(peers);
            keys_out[base + rank] = k;
            vals_out[base + rank] = vals[i];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ pipeline kernels
struct PostMeta {
    int n_pairs, n_images;
    int xb, yb;  // key bits of the truncated x / y coordinate
};

// max truncated coordinate (and a validity flag: every coordinate must be finite and >= 0)
static __global__ void coord_range_kernel(const float* __restrict__ rows, long long T, int* __restrict__ range /* [0]=max x, [1]=max y, [2]=bad */) {
    int mx = 0, my = 0, bad = 0;
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < T; t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float* r = rows + t * 5;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const float x = r[2 * s], y = r[2 * s + 1];
            if (!(x >= 0.f && x < 1.0e9f && y >= 0.f && y < 1.0e9f)) { bad = 1; continue; }
            mx = max(mx, static_cast<int>(x));
            my = max(my, static_cast<int>(y));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
        my = max(my, __shfl_down_sync(0xffffffffu, my, o));
        bad |= __shfl_down_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&range[0], mx);
        atomicMax(&range[1], my);
        if (bad) atomicOr(&range[2], 1);
    }
}

// keys / values of the 2T observations.  pair_off: [P+1] first match row of every pair; pair_img: [P][2] image indices.
static __global__ void obs_key_kernel(const float* __restrict__ rows, long long T, const long long* __restrict__ pair_off,
                                      const int* __restrict__ pair_img, PostMeta m, unsigned long long* __restrict__ keys,
                                      unsigned int* __restrict__ vals) {
    const long long o = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (o >= 2 * T) return;
    const long long t = o >> 1;
    const int side = static_cast<int>(o & 1);
    int lo = 0, hi = m.n_pairs;  // last pair with pair_off[pair] <= t
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (pair_off[mid] <= t) lo = mid; else hi = mid;
    }
    const unsigned long long img = static_cast<unsigned long long>(pair_img[2 * lo + side]);
    const unsigned long long x = static_cast<unsigned long long>(static_cast<int>(rows[t * 5 + 2 * side]));      // .astype(int): truncation
    const unsigned long long y = static_cast<unsigned long long>(static_cast<int>(rows[t * 5 + 2 * side + 1]));
    keys[o] = (img << (m.xb + m.yb)) | (x << m.yb) | y;
    vals[o] = static_cast<unsigned int>(o);
}

static __global__ void head_flag_kernel(const unsigned long long* __restrict__ keys, long long n, int* __restrict__ flag) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// One thread per run head: float64 sum of the run's confidences in sorted (= reference input) order; records the run's unique
// index for each of its observations.
static __global__ void run_reduce_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ vals, long long n,
                                         const int* __restrict__ flag, const int* __restrict__ uidx, const float* __restrict__ rows,
                                         unsigned long long* __restrict__ ukey, unsigned long long* __restrict__ uscore_key,
                                         double* __restrict__ usum, unsigned int* __restrict__ uval, int* __restrict__ obs_u) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const unsigned long long k = keys[i];
    const int u = uidx[i];
    double s = 0.0;
    for (long long j = i; j < n && keys[j] == k; ++j) {
        const unsigned int o = vals[j];
        s += static_cast<double>(rows[static_cast<long long>(o >> 1) * 5 + 4]);
        obs_u[o] = u;
    }
    ukey[u] = k;
    usum[u] = s;
    uscore_key[u] = ~static_cast<unsigned long long>(__double_as_longlong(s));  // ascending ~bits == descending positive score
    uval[u] = static_cast<unsigned int>(u);
}

// second sort, last stage: the image index as the key (stable -> image, then descending score, then (x, y))
static __global__ void image_key_kernel(const unsigned int* __restrict__ uval_sorted, const unsigned long long* __restrict__ ukey, int K, int shift,
                                        unsigned long long* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < K) out[r] = ukey[uval_sorted[r]] >> shift;
}

// rank r -> keypoint (x, y) float32, score float32, first rank of every image, rank of every unique
static __global__ void emit_kernel(const unsigned int* __restrict__ uval_sorted, const unsigned long long* __restrict__ ukey,
                                   const double* __restrict__ usum, int K, PostMeta m, float* __restrict__ kpt_xy, float* __restrict__ kpt_score,
                                   int* __restrict__ img_off, int* __restrict__ rank_of_u) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= K) return;
    const unsigned int u = uval_sorted[r];
    const unsigned long long k = ukey[u];
    const int sh = m.xb + m.yb;
    const int img = static_cast<int>(k >> sh);
    kpt_xy[2 * r] = static_cast<float>(static_cast<int>((k >> m.yb) & ((1ull << m.xb) - 1ull)));
    kpt_xy[2 * r + 1] = static_cast<float>(static_cast<int>(k & ((1ull << m.yb) - 1ull)));
    kpt_score[r] = static_cast<float>(usum[u]);
    rank_of_u[u] = r;
    const int prev = r == 0 ? -1 : static_cast<int>(ukey[uval_sorted[r - 1]] >> sh);
    for (int i = prev + 1; i <= img; ++i) img_off[i] = r;  // images without key points share the next image's first rank
    if (r == K - 1)
        for (int i = img + 1; i <= m.n_images; ++i) img_off[i] = K;
}

static __global__ void match_ids_kernel(const int* __restrict__ obs_u, const int* __restrict__ rank_of_u, const unsigned long long* __restrict__ ukey,
                                        const int* __restrict__ img_off, long long T, PostMeta m, int* __restrict__ match_ids) {
    const long long o = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (o >= 2 * T) return;
    const int u = obs_u[o];
    const int img = static_cast<int>(ukey[u] >> (m.xb + m.yb));
    match_ids[o] = rank_of_u[u] - img_off[img];  // o = 2*t + side: the [T][2] layout
}

// ------------------------------------------------------------------------------------------------ host side
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void* get(size_t bytes) {
        if (bytes > cap) {
            if (p) cudaFree(p);
            p = nullptr;
            cap = 0;
            DFSFM_CUDA(cudaMalloc(&p, bytes));
            cap = bytes;
        }
        return p;
    }
    ~DevBuf() { if (p) cudaFree(p); }
};

class PostEngine {
public:
    explicit PostEngine(int device) : device_(device) { DFSFM_CUDA(cudaSetDevice(device)); }

    long long merge(const float* rows, long long T, int n_pairs, const long long* pair_off_dev, const int* pair_img_dev, int n_images,
                    float* kpt_xy, float* kpt_score, int* img_off, int* match_ids, cudaStream_t st) {
        DFSFM_CUDA(cudaSetDevice(device_));
        DFSFM_CHECK(T >= 0 && 2 * T < (1ll << 31), "too many matches for 32-bit observation indices");
        DFSFM_CHECK(n_images >= 1 && n_images < (1 << 24), "n_images out of range");
        if (T == 0) {
            DFSFM_CUDA(cudaMemsetAsync(img_off, 0, sizeof(int) * (n_images + 1), st));
            DFSFM_CUDA(cudaStreamSynchronize(st));
            return 0;
        }
        const long long N = 2 * T;
        int* range = static_cast<int*>(small_.get(64));
        DFSFM_CUDA(cudaMemsetAsync(range, 0, 3 * sizeof(int), st));
        { LaunchScope ls("post_range", st);
          coord_range_kernel<<<sm_count() * 4, 256, 0, st>>>(rows, T, range); }
        int h_range[3];
        DFSFM_CUDA(cudaMemcpyAsync(h_range, range, sizeof(h_range), cudaMemcpyDeviceToHost, st));
        DFSFM_CUDA(cudaStreamSynchronize(st));
        DFSFM_CHECK(h_range[2] == 0, "match coordinates must be finite and non-negative");
        PostMeta m;
        m.n_pairs = n_pairs;
        m.n_images = n_images;
        m.xb = bits_for(h_range[0]);
        m.yb = bits_for(h_range[1]);
        const int ib = bits_for(n_images - 1);
        DFSFM_CHECK(m.xb + m.yb + ib <= 62, "key does not fit 64 bits");

        // workspace
        const size_t kb = static_cast<size_t>(N) * 8, vb = static_cast<size_t>(N) * 4;
        unsigned long long* keys[2] = {static_cast<unsigned long long*>(k0_.get(kb)), static_cast<unsigned long long*>(k1_.get(kb))};
        unsigned int* vals[2] = {static_cast<unsigned int*>(v0_.get(vb)), static_cast<unsigned int*>(v1_.get(vb))};
        int* flag = static_cast<int*>(flag_.get(vb));
        int* uidx = static_cast<int*>(uidx_.get(vb));
        int* obs_u = static_cast<int*>(obsu_.get(vb));
        unsigned long long* ukey = static_cast<unsigned long long*>(ukey_.get(kb));
        double* usum = static_cast<double*>(usum_.get(kb));
        int* rank_of_u = static_cast<int*>(rank_.get(vb));
        const int nseg = static_cast<int>((N + kSeg - 1) / kSeg);
        int* hist = static_cast<int*>(hist_.get(static_cast<size_t>(nseg) * 256 * sizeof(int)));
        const long long scan_max = std::max<long long>(N, static_cast<long long>(nseg) * 256);
        int* bsums = static_cast<int*>(bsum_.get(static_cast<size_t>((scan_max + kScanTile - 1) / kScanTile + 1) * sizeof(int)));

        const int tb = 256;
        { LaunchScope ls("post_keys", st);
          obs_key_kernel<<<static_cast<unsigned>((N + tb - 1) / tb), tb, 0, st>>>(rows, T, pair_off_dev, pair_img_dev, m, keys[0], vals[0]); }
        int cur = 0;
        cur = radix_sort(keys, vals, cur, N, 0, m.xb + m.yb + ib, hist, bsums, st, "post_scatter");
        { LaunchScope ls("post_flag", st);
          head_flag_kernel<<<static_cast<unsigned>((N + tb - 1) / tb), tb, 0, st>>>(keys[cur], N, flag); }
        int* total = range;  // reuse the small buffer
        exclusive_scan(flag, uidx, N, bsums, total, st);
        // the sorted observation arrays stay in keys[cur] / vals[cur]; the other pair of buffers serves the second sort
        const int oth = cur ^ 1;
        { LaunchScope ls("post_reduce", st);
          run_reduce_kernel<<<static_cast<unsigned>((N + tb - 1) / tb), tb, 0, st>>>(keys[cur], vals[cur], N, flag, uidx, rows, ukey, keys[oth], usum,
                                                                                     vals[oth], obs_u); }
        int h_total = 0;
        DFSFM_CUDA(cudaMemcpyAsync(&h_total, total, sizeof(int), cudaMemcpyDeviceToHost, st));
        DFSFM_CUDA(cudaStreamSynchronize(st));
        const int K = h_total;
        DFSFM_CHECK(K >= 1 && K <= N, "unique count out of range");
        // second sort: keys[oth]/vals[oth] hold (~score bits, u) in (image, x, y) order; ping-pong partner is a fresh pair
        unsigned long long* k2[2] = {keys[oth], static_cast<unsigned long long*>(k2_.get(static_cast<size_t>(K) * 8))};
        unsigned int* v2[2] = {vals[oth], static_cast<unsigned int*>(v2_.get(static_cast<size_t>(K) * 4))};
        int c2 = radix_sort(k2, v2, 0, K, 0, 64, hist, bsums, st, "post_scatter_rank");
        if (ib > 0) {
            { LaunchScope ls("post_imgkey", st);
              image_key_kernel<<<(K + tb - 1) / tb, tb, 0, st>>>(v2[c2], ukey, K, m.xb + m.yb, k2[c2]); }
            c2 = radix_sort(k2, v2, c2, K, 0, ib, hist, bsums, st, "post_scatter_rank");
        }
        { LaunchScope ls("post_emit", st);
          emit_kernel<<<(K + tb - 1) / tb, tb, 0, st>>>(v2[c2], ukey, usum, K, m, kpt_xy, kpt_score, img_off, rank_of_u); }
        { LaunchScope ls("post_ids", st);
          match_ids_kernel<<<static_cast<unsigned>((N + tb - 1) / tb), tb, 0, st>>>(obs_u, rank_of_u, ukey, img_off, T, m, match_ids); }
        DFSFM_CUDA(cudaGetLastError());
        DFSFM_CUDA(cudaStreamSynchronize(st));
        return K;
    }

private:
    static int bits_for(int v) {  // bits needed to represent values 0..v
        int b = 1;
        while ((1ll << b) <= v) ++b;
        return b;
    }
    void exclusive_scan(const int* in, int* out, long long n, int* bsums, int* total, cudaStream_t st) {
        const int nb = static_cast<int>((n + kScanTile - 1) / kScanTile);
        { LaunchScope ls("post_scan", st);
          scan_reduce_kernel<<<nb, kScanThreads, 0, st>>>(in, n, bsums); }
        { LaunchScope ls("post_scan", st);
          scan_sums_kernel<<<1, 1024, 0, st>>>(bsums, nb, total); }
        { LaunchScope ls("post_scan", st);
          scan_apply_kernel<<<nb, kScanThreads, 0, st>>>(in, out, n, bsums); }
    }
    // LSD passes over key bits [bit_lo, bit_hi); returns the index of the buffer pair holding the result
    int radix_sort(unsigned long long* keys[2], unsigned int* vals[2], int cur, long long n, int bit_lo, int bit_hi, int* hist, int* bsums,
                   cudaStream_t st, const char* scatter_label) {
        const int nseg = static_cast<int>((n + kSeg - 1) / kSeg);
        const int nblk = (nseg + 7) / 8;
        for (int shift = bit_lo; shift < bit_hi; shift += 8) {
            { LaunchScope ls("post_hist", st);
              rs_hist_kernel<<<nblk, 256, 0, st>>>(keys[cur], n, shift, nseg, hist); }
            exclusive_scan(hist, hist, static_cast<long long>(nseg) * 256, bsums, nullptr, st);
            { LaunchScope ls(scatter_label, st);
              rs_scatter_kernel<<<nblk, 256, 0, st>>>(keys[cur], vals[cur], n, shift, nseg, hist, keys[cur ^ 1], vals[cur ^ 1]); }
            cur ^= 1;
        }
        return cur;
    }

    int device_;
    DevBuf small_, k0_, k1_, v0_, v1_, flag_, uidx_, obsu_, ukey_, usum_, rank_, hist_, bsum_, k2_, v2_;
};

}  // namespace dfsfm

struct dfsfm_post {
    std::unique_ptr<dfsfm::PostEngine> e;
};

extern "C" {

int dfsfm_post_create(dfsfm_post_t** out, int device) {
    return dfsfm::guard([&] {
        auto* h = new dfsfm_post;
        h->e.reset(new dfsfm::PostEngine(device));
        *out = h;
    });
}
void dfsfm_post_destroy(dfsfm_post_t* h) { delete h; }
int dfsfm_post_merge_keypoints(dfsfm_post_t* h, const float* rows_dev, int64_t n_rows, int n_pairs, const int64_t* pair_offset_dev,
                               const int32_t* pair_images_dev, int n_images, float* kpt_xy_dev, float* kpt_score_dev, int32_t* image_offset_dev,
                               int32_t* match_ids_dev, int64_t* n_keypoints, void* stream) {
    return dfsfm::guard([&] {
        static_assert(sizeof(long long) == sizeof(int64_t), "int64_t layout");
        *n_keypoints = h->e->merge(rows_dev, n_rows, n_pairs, reinterpret_cast<const long long*>(pair_offset_dev), pair_images_dev, n_images,
                                   kpt_xy_dev, kpt_score_dev, image_offset_dev, match_ids_dev, static_cast<cudaStream_t>(stream));
    });
}

}  // extern "C"
