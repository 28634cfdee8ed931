#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_scalar(float* out, int iters) {
    float a[16];
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.001f + i;
    float w = 0.999f, c = 0.001f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], w, c);
    }
    float s = 0; for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, int iters) {
    float2 a[8];
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    float2 w = make_float2(0.999f, 0.998f), c = make_float2(0.001f, 0.002f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = __ffma2_rn(a[i], w, c);
    }
    float s = 0; for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed_add(float* out, int iters) {
    float2 a[8];
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    float2 c = make_float2(0.001f, 0.002f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = __fadd2_rn(a[i], c);
    }
    float s = 0; for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_scalar_add(float* out, int iters) {
    float a[16];
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.001f + i;
    float c = 0.001f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = a[i] + c;
    }
    float s = 0; for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000; float ms;
    for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0); k_scalar<<<148 * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("scalar ffma : %.3f ms  %.1f TFLOP/s\n", ms, 148.0 * 8 * 256 * 16 * 2.0 * iters / ms / 1e9);
    cudaEventRecord(e0); k_packed<<<148 * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("packed ffma2: %.3f ms  %.1f TFLOP/s\n", ms, 148.0 * 8 * 256 * 16 * 2.0 * iters / ms / 1e9);
    cudaEventRecord(e0); k_scalar_add<<<148 * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("scalar fadd : %.3f ms  %.1f Tadd/s\n", ms, 148.0 * 8 * 256 * 16 * 1.0 * iters / ms / 1e9);
    cudaEventRecord(e0); k_packed_add<<<148 * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("packed fadd2: %.3f ms  %.1f Tadd/s\n", ms, 148.0 * 8 * 256 * 16 * 1.0 * iters / ms / 1e9);
    }
    return 0;
}
