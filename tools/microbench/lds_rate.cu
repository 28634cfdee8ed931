// lds_rate.cu -- shared-memory instruction throughput per SM on B200: LDS.32 / LDS.64 / LDS.128 / STS.32 / STS.64,
// conflict-free, 32 warps per SM, and the same with 8 independent FADD2 per 4 loads (overlap test).
#include <cuda_runtime.h>
#include <cstdio>

template <int W, bool STORE, int NFADD2>
__global__ void k(float *out, int iters) {
    extern __shared__ float4 smem4[];
    float *sm = reinterpret_cast<float *>(smem4);
    for (int i = threadIdx.x; i < 12288; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    unsigned accu = 0;
    float2 a[8];
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, i);
    const float2 c = make_float2(0.001f, 0.002f);
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + warp * 2048 + lane * (4 * W);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const unsigned addr = base + ((j & 1) * 1024) + ((it & 1) << 14);
            if (STORE) {
                if (W == 1) asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(acc));
                if (W == 2) asm volatile("st.shared.v2.f32 [%0], {%1, %1};" ::"r"(addr), "f"(acc));
                if (W == 4) asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "f"(acc));
            } else {
                float x, y, z, w;
                if (W == 1) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
                if (W == 2) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr));
                if (W == 4) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr));
                accu += __float_as_uint(x);
            }
            if (j < NFADD2) a[j] = __fadd2_rn(a[j], c);
        }
    }
    float s = acc + (float)accu;
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int W, bool STORE, int NF>
void run(const char *name, float *out) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    cudaFuncSetAttribute(k<W, STORE, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    k<W, STORE, NF><<<148 * 2, 512, 49152>>>(out, 10);
    cudaEventRecord(e0);
    k<W, STORE, NF><<<148 * 2, 512, 49152>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    // per SM: 32 warps x 8 instructions per iteration
    const double cyc_per_inst = ms * 1e-3 * 1.965e9 / ((double)iters * 32 * 8);
    printf("%-26s %.3f ms  %.2f SM-cycles per warp instruction, %.1f B/clk/SM\n", name, ms, cyc_per_inst, 32.0 * 4 * W / cyc_per_inst);
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 2 * 512 * 4);
    run<1, false, 0>("LDS.32", out);
    run<2, false, 0>("LDS.64", out);
    run<4, false, 0>("LDS.128", out);
    run<1, true, 0>("STS.32", out);
    run<2, true, 0>("STS.64", out);
    run<4, true, 0>("STS.128", out);
    run<2, false, 8>("LDS.64 + 8 FADD2 per 8", out);
    run<2, false, 4>("LDS.64 + 4 FADD2 per 8", out);
    run<1, false, 8>("LDS.32 + 8 FADD2 per 8", out);
    run<4, false, 8>("LDS.128 + 8 FADD2 per 8", out);
    return 0;
}
