// fft_radix.cuh -- register-resident forward DFT butterflies (fp32) for the fused spectral kernel.
//
// Replaces the arithmetic of github.com/mjibson/go-dsp/fft.FFT as called from dsp/fft.go:26
// (forward DFT, sign -, unnormalised).  The reference runs a complex128 radix-2 loop; here each
// thread keeps 16 complex points in registers and applies radix-16/8/4/2 butterflies, exchanging
// through shared memory between passes (see k1_spectral.cuh).
//
// Register order: dftR leaves the result for frequency index OutIdx<R>::of(p) in register p
// (digit-reversed); callers index with that constexpr map, everything is fully unrolled.
#pragma once
#include <cuda_runtime.h>

namespace sdr {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// a * w
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}
// a * (-i) = (a.y, -a.x)
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

#define SDR_SQRT1_2 0.70710678118654752440f
#define SDR_COS_PI_8 0.92387953251128675613f
#define SDR_SIN_PI_8 0.38268343236508977173f

__device__ __forceinline__ void dft2(float2 &a, float2 &b) {
    float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// natural order in, natural order out
__device__ __forceinline__ void dft4(float2 &x0, float2 &x1, float2 &x2, float2 &x3) {
    float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = csub(x1, x3);
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    // X1 = t1 - i*t3, X3 = t1 + i*t3
    x1 = make_float2(t1.x + t3.y, t1.y - t3.x);
    x3 = make_float2(t1.x - t3.y, t1.y + t3.x);
}

template <int R>
struct OutIdx;
template <>
struct OutIdx<2> {
    __host__ __device__ static constexpr int of(int p) { return p; }
};
template <>
struct OutIdx<4> {
    __host__ __device__ static constexpr int of(int p) { return p; }
};
template <>
struct OutIdx<8> {  // position p = 4*ka + kb holds X[ka + 2*kb]
    __host__ __device__ static constexpr int of(int p) { return (p >> 2) + 2 * (p & 3); }
};
template <>
struct OutIdx<16> {  // position p = 4*ka + kb holds X[ka + 4*kb]
    __host__ __device__ static constexpr int of(int p) { return (p >> 2) + 4 * (p & 3); }
};

// 8-point: n = 4*na + nb; DFT2 over na, twiddle W8^(nb*ka), DFT4 over nb.
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
#pragma unroll
    for (int b = 0; b < 4; b++) dft2(v[b], v[b + 4]);
    // ka = 1 row: v[4+b] *= W8^b
    {
        float2 a = v[5];  // W8^1 = s(1 - i)
        v[5] = make_float2(SDR_SQRT1_2 * (a.x + a.y), SDR_SQRT1_2 * (a.y - a.x));
        v[6] = mul_mi(v[6]);  // W8^2 = -i
        a = v[7];             // W8^3 = s(-1 - i)
        v[7] = make_float2(SDR_SQRT1_2 * (a.y - a.x), -SDR_SQRT1_2 * (a.x + a.y));
    }
    dft4(v[0], v[1], v[2], v[3]);
    dft4(v[4], v[5], v[6], v[7]);
}

// 16-point: n = 4*na + nb; DFT4 over na (stride 4), twiddle W16^(nb*ka), DFT4 over nb.
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
#pragma unroll
    for (int b = 0; b < 4; b++) dft4(v[b], v[b + 4], v[b + 8], v[b + 12]);
    // position 4*ka + b holds y[ka][b]; multiply by W16^(b*ka)
    const float2 w1 = make_float2(SDR_COS_PI_8, -SDR_SIN_PI_8);   // W16^1
    const float2 w3 = make_float2(SDR_SIN_PI_8, -SDR_COS_PI_8);   // W16^3
    // ka = 1: exponents 0,1,2,3
    v[5] = cmul(v[5], w1);
    {
        float2 a = v[6];  // W16^2 = s(1 - i)
        v[6] = make_float2(SDR_SQRT1_2 * (a.x + a.y), SDR_SQRT1_2 * (a.y - a.x));
    }
    v[7] = cmul(v[7], w3);
    // ka = 2: exponents 0,2,4,6
    {
        float2 a = v[9];  // W16^2
        v[9] = make_float2(SDR_SQRT1_2 * (a.x + a.y), SDR_SQRT1_2 * (a.y - a.x));
        v[10] = mul_mi(v[10]);  // W16^4 = -i
        a = v[11];              // W16^6 = s(-1 - i)
        v[11] = make_float2(SDR_SQRT1_2 * (a.y - a.x), -SDR_SQRT1_2 * (a.x + a.y));
    }
    // ka = 3: exponents 0,3,6,9
    v[13] = cmul(v[13], w3);
    {
        float2 a = v[14];  // W16^6
        v[14] = make_float2(SDR_SQRT1_2 * (a.y - a.x), -SDR_SQRT1_2 * (a.x + a.y));
    }
    v[15] = cmul(v[15], make_float2(-SDR_COS_PI_8, SDR_SIN_PI_8));  // W16^9 = -W16^1
#pragma unroll
    for (int a = 0; a < 4; a++) dft4(v[4 * a], v[4 * a + 1], v[4 * a + 2], v[4 * a + 3]);
}

template <int R>
__device__ __forceinline__ void dftR(float2 *v);
template <>
__device__ __forceinline__ void dftR<2>(float2 *v) { dft2(v[0], v[1]); }
template <>
__device__ __forceinline__ void dftR<4>(float2 *v) { dft4(v[0], v[1], v[2], v[3]); }
template <>
__device__ __forceinline__ void dftR<8>(float2 *v) { dft8(*reinterpret_cast<float2(*)[8]>(v)); }
template <>
__device__ __forceinline__ void dftR<16>(float2 *v) { dft16(*reinterpret_cast<float2(*)[16]>(v)); }

}  // namespace sdr
