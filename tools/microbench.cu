// Issue-rate microbenchmarks for the epilogue instruction mix (sm_100a): cycles per warp-instruction per scheduler with
// 2 warps per scheduler (256 threads, 1 CTA per SM), plus tcgen05.ld and the swizzled shared-memory transposition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITER 256
#define CHK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int OP>
__global__ void __launch_bounds__(256, 1) k_alu(float* out, long long* cyc, float seed) {
    float a[16];
    uint32_t u[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = seed + threadIdx.x * 0.001f + i; u[i] = threadIdx.x * 77u + i; }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) a[i] = fmaf(a[i], 1.0001f, 0.5f);
            if (OP == 1) { unsigned short h; asm volatile("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(a[i])); u[i] += h; }
            if (OP == 2) { float f; asm volatile("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"((unsigned short)u[i])); a[i] += f; }
            if (OP == 3) { uint32_t h2; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(a[i]), "f"(a[(i + 1) & 15])); u[i] ^= h2; }
            if (OP == 4) { asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(a[i]) : "h"((unsigned short)u[i])); }
            if (OP == 5) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
            if (OP == 6) { asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 15])); }
            if (OP == 7) { u[i] = u[i] * 3u + 7u; }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// swizzled 32x32 fp32 transposition through a 4 KB per-warp buffer: 8 STS.128 + syncwarp + 8 LDS.128 + syncwarp
__global__ void __launch_bounds__(256, 1) k_transpose(float* out, long long* cyc) {
    __shared__ __align__(128) uint8_t buf[8 * 4096];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* stg = buf + warp * 4096;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
    const int ch = lane & 7, rsub = lane >> 3;
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rl = 4 * i + rsub;
            const float4 w = *reinterpret_cast<const float4*>(stg + rl * 128 + ((ch ^ (rl & 7)) << 4));
            v[4 * i] += w.x; v[4 * i + 1] += w.y; v[4 * i + 2] += w.z; v[4 * i + 3] += w.w;
        }
        __syncwarp();
    }
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// tcgen05.ld.32x32b.x32 + wait, NW warps reading (warp w reads its lane quadrant w % 4)
template <int NW, int PER_WAIT>
__global__ void __launch_bounds__(256, 1) k_tmem(float* out, long long* cyc) {
    __shared__ uint32_t slot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    float acc = 0.f;
    const long long t0 = clock64();
    if (warp < NW) {
        const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
        for (int it = 0; it < ITER; ++it) {
            uint32_t r[PER_WAIT][32];
#pragma unroll
            for (int b = 0; b < PER_WAIT; ++b) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[b][0]), "=r"(r[b][1]), "=r"(r[b][2]), "=r"(r[b][3]), "=r"(r[b][4]), "=r"(r[b][5]), "=r"(r[b][6]), "=r"(r[b][7]), "=r"(r[b][8]),
                      "=r"(r[b][9]), "=r"(r[b][10]), "=r"(r[b][11]), "=r"(r[b][12]), "=r"(r[b][13]), "=r"(r[b][14]), "=r"(r[b][15]), "=r"(r[b][16]),
                      "=r"(r[b][17]), "=r"(r[b][18]), "=r"(r[b][19]), "=r"(r[b][20]), "=r"(r[b][21]), "=r"(r[b][22]), "=r"(r[b][23]), "=r"(r[b][24]),
                      "=r"(r[b][25]), "=r"(r[b][26]), "=r"(r[b][27]), "=r"(r[b][28]), "=r"(r[b][29]), "=r"(r[b][30]), "=r"(r[b][31])
                    : "r"(taddr + ((it * PER_WAIT + b) * 32) % 512 + ((warp >> 2) * 0))
                    : "memory");
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int b = 0; b < PER_WAIT; ++b) acc += __uint_as_float(r[b][0]) + __uint_as_float(r[b][31]);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}

int main() {
    float* out; long long* cyc;
    CHK(cudaMalloc(&out, 148 * 256 * 4)); CHK(cudaMalloc(&cyc, 148 * 8));
    long long h[148];
    const char* names[8] = {"FFMA", "cvt.rn.f16.f32 (F2F)", "cvt.f32.f16", "cvt.rn.f16x2.f32 (F2FP)", "add.rn.f32.f16 (FHADD)", "ex2.approx (MUFU)", "add.f16x2 (HADD2)", "IMAD"};
#define RUN_ALU(OP) { k_alu<OP><<<148, 256>>>(out, cyc, 1.f); k_alu<OP><<<148, 256>>>(out, cyc, 1.f); CHK(cudaDeviceSynchronize()); CHK(cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost)); \
    printf("%-28s %7.2f cycles per warp-instruction per scheduler (2 warps/scheduler -> x2 instr)\n", names[OP], (double)h[0] / (ITER * 16 * 2)); }
    RUN_ALU(0) RUN_ALU(1) RUN_ALU(2) RUN_ALU(3) RUN_ALU(4) RUN_ALU(5) RUN_ALU(6) RUN_ALU(7)
    k_transpose<<<148, 256>>>(out, cyc); k_transpose<<<148, 256>>>(out, cyc); CHK(cudaDeviceSynchronize()); CHK(cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost));
    printf("transpose 32x32 fp32 (8 warps) %7.1f cycles per block per warp\n", (double)h[0] / ITER);
#define RUN_TM(NW, PW) { k_tmem<NW, PW><<<148, 256>>>(out, cyc); k_tmem<NW, PW><<<148, 256>>>(out, cyc); CHK(cudaDeviceSynchronize()); CHK(cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost)); \
    printf("tcgen05.ld x32: %d warps, %d loads per wait: %7.1f cycles per load per warp; %6.1f B/clk/SM\n", NW, PW, (double)h[0] / (ITER * PW), 4096.0 * NW * ITER * PW / (double)h[0]); }
    RUN_TM(1, 1) RUN_TM(4, 1) RUN_TM(8, 1) RUN_TM(4, 4) RUN_TM(8, 4) RUN_TM(8, 2)
    return 0;
}
