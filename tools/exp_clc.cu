// Probe of cluster launch control (work stealing): 1000 two-CTA clusters, every tile index must be processed exactly once by both CTA ranks.
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/exp_clc tools/exp_clc.cu && /tmp/exp_clc
#include <cstdint>
#include <cstdio>
__global__ void __cluster_dims__(2,1,1) k(int* out, int* per_cluster, int tiles) {
    __shared__ __align__(16) uint4 resp[2];
    __shared__ __align__(8) uint64_t bar[2];
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    int tile = blockIdx.x >> 1;
    int it = 0;
    while (true) {
        int q = it & 1; uint32_t ph = (it >> 1) & 1;
        uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[q]);
        uint32_t r = (uint32_t)__cvta_generic_to_shared(&resp[q]);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" :: "r"(b) : "memory");
            if (rank == 0)
                asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];" :: "r"(r), "r"(b) : "memory");
        }
        if (threadIdx.x == 0) atomicAdd(&out[tile], 1 + (int)rank * 1000);   // 1 from CTA 0 + 1001 from CTA 1 = 1002 per tile
        for (volatile int spin = 0; spin < 20000; ++spin) {}                // the "work" of a tile
        // wait
        uint32_t done = 0;
        while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(b), "r"(ph) : "memory");
        uint32_t valid, x, y, z;
        asm volatile("{.reg .pred p1; .reg .b128 c; ld.shared.b128 c, [%4]; clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, c; selp.u32 %3, 1, 0, p1; @p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, c;}"
                     : "=r"(x), "=r"(y), "=r"(z), "=r"(valid) : "r"(r) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (!valid) { if (threadIdx.x == 0 && rank == 0) per_cluster[blockIdx.x >> 1] = it + 1; break; }
        tile = x >> 1;
        ++it;
        __syncthreads();
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main() {
    int tiles = 1000; int* d; cudaMalloc(&d, tiles * 4); cudaMemset(d, 0, tiles * 4);
    int* pc; cudaMalloc(&pc, tiles * 4); cudaMemset(pc, 0, tiles * 4);
    k<<<2 * tiles, 128>>>(d, pc, tiles);
    cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
    int* h = new int[tiles]; cudaMemcpy(h, d, tiles * 4, cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < tiles; ++i) if (h[i] != 1002) ++bad;
    printf("tiles processed other than exactly once by both CTAs: %d of %d\n", bad, tiles);
    cudaMemcpy(h, pc, tiles * 4, cudaMemcpyDeviceToHost);
    int launched = 0, mx = 0, sum = 0; for (int i = 0; i < tiles; ++i) if (h[i]) { ++launched; sum += h[i]; if (h[i] > mx) mx = h[i]; }
    printf("clusters launched %d (the rest were cancelled), tiles per launched cluster: max %d, total %d\n", launched, mx, sum);
}
