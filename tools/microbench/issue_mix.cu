// issue_mix.cu -- does a packed f32x2 instruction (FADD2/FFMA2, 2 FP32-pipe cycles per warp) leave its second cycle
// free for another pipe's instruction?  Each loop body issues 8 independent FADD2 plus K independent
// integer (ALU) or shared-memory (LSU) instructions; if issue slots are shared, time(K=8) == time(K=0).
#include <cuda_runtime.h>
#include <cstdio>

template <int K_ALU, int K_LDS>
__global__ void k_mix(float *out, int iters) {
    __shared__ float2 sm[32 * 8 * 8 + 64];
    float2 a[8];
    unsigned b[8];
    float2 l[8];
    for (int i = 0; i < 8; i++) {
        a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
        b[i] = threadIdx.x * 7 + i;
        l[i] = make_float2(0.f, 0.f);
    }
    for (int i = threadIdx.x; i < 32 * 8 * 8 + 64; i += blockDim.x) sm[i] = make_float2(i, -i);
    __syncthreads();
    const float2 c = make_float2(0.001f, 0.002f);
    const float2 *p = sm + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            a[i] = __fadd2_rn(a[i], c);
            if (i < K_ALU) b[i] = (b[i] ^ (b[i] >> 3)) + 0x9e37u;  // 3 ALU ops, independent chains
            if (i < K_LDS) {
                float2 t = p[(i * 32 + (it & 7)) ];
                l[i].x += 0.f * t.x;  // keep the load alive cheaply? (adds FP work) -> use xor accumulate instead
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y + (float)b[i] + l[i].x;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// latency: one dependent chain per warp, one warp per SM sub-partition
__global__ void k_lat_fadd2(float *out, int iters) {
    float2 a = make_float2(threadIdx.x, 1.f);
    const float2 c = make_float2(0.001f, 0.002f);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a = __fadd2_rn(a, c);
    }
    long long t1 = clock64();
    out[threadIdx.x] = a.x + a.y;
    if (threadIdx.x == 0) out[64] = (float)(t1 - t0) / (16.f * iters);
}
__global__ void k_lat_fadd(float *out, int iters) {
    float a = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a = a + 0.001f;
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) out[64] = (float)(t1 - t0) / (16.f * iters);
}

template <int KA, int KL>
void run(const char *name, float *out, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    k_mix<KA, KL><<<148 * 4, 256>>>(out, 100);
    cudaEventRecord(e0);
    k_mix<KA, KL><<<148 * 4, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    // cycles per loop body per SMSP: 8 warps per SMSP (4 CTAs x 8 warps / 4)
    const double cyc = ms * 1e-3 * 1.965e9 / iters / 8.0;
    printf("%-28s %.3f ms  %.2f cycles per warp-iteration (8 FADD2 + %d ALU-chains x3 + %d LDS.64)\n", name, ms, cyc, KA, KL);
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000;
    run<0, 0>("fadd2 only", out, iters);
    run<2, 0>("fadd2 + 6 alu", out, iters);
    run<4, 0>("fadd2 + 12 alu", out, iters);
    run<8, 0>("fadd2 + 24 alu", out, iters);
    run<0, 4>("fadd2 + 4 lds", out, iters);
    run<0, 8>("fadd2 + 8 lds", out, iters);
    run<4, 4>("fadd2 + 12 alu + 4 lds", out, iters);
    float h[65];
    k_lat_fadd2<<<1, 32>>>(out, 1000);
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("FADD2 dependent latency: %.2f cycles\n", h[64]);
    k_lat_fadd<<<1, 32>>>(out, 1000);
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("FADD  dependent latency: %.2f cycles\n", h[64]);
    return 0;
}
