// fft_radix.cuh -- register-resident forward DFT butterflies (fp32) for the fused spectral kernel.
//
// Replaces the arithmetic of github.com/mjibson/go-dsp/fft.FFT as called from dsp/fft.go:26
// (forward DFT, sign -, unnormalised).  The reference runs a complex128 radix-2 loop; here each
// thread keeps 16 complex points in registers and applies radix-16/8/4/2 butterflies, exchanging
// through shared memory between passes (see k1_spectral.cuh).
//
// B200-specific: a complex value lives in one 64-bit register pair (re, im) and all arithmetic is
// issued as packed f32x2 instructions (sm_100a FADD2 / FMUL2 / FFMA2).  Their operand modifiers
// (half swap LO_HI, per-half negate, 32-bit scalar broadcast) make a complex add ONE instruction,
// a multiplication by +-i free, and a complex multiply TWO instructions -- half the issue slots of
// scalar code at the same FP32 pipe throughput (tools/microbench/packed_f32x2.cu).  The kernel is
// issue-bound, so this is where the time goes.
//
// Register order: dftR leaves the result for frequency index OutIdx<R>::of(p) in register p
// (digit-reversed); callers index with that constexpr map, everything is fully unrolled.
#pragma once
#include <cuda_runtime.h>

namespace sdr {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// a - i*b = (a.x + b.y, a.y - b.x)
__device__ __forceinline__ float2 csub_i(float2 a, float2 b) { return __fadd2_rn(a, make_float2(b.y, -b.x)); }
// a + i*b = (a.x - b.y, a.y + b.x)
__device__ __forceinline__ float2 cadd_i(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.y, b.x)); }
// a * w = (a.x*w.x - a.y*w.y, a.y*w.x + a.x*w.y): FMUL2 with w.x broadcast, FFMA2 with swapped a and (-w.y, +w.y)
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    const float2 t = __fmul2_rn(a, make_float2(w.x, w.x));
    return __ffma2_rn(make_float2(a.y, a.x), make_float2(-w.y, w.y), t);
}
// a * (-i) = (a.y, -a.x): folded into the consumer's operand modifiers
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

#define SDR_SQRT1_2 0.70710678118654752440f
#define SDR_COS_PI_8 0.92387953251128675613f
#define SDR_SIN_PI_8 0.38268343236508977173f

// a * W8^1 = a * s(1 - i) = s(a.x + a.y, a.y - a.x)
__device__ __forceinline__ float2 mul_w8_1(float2 a) {
    const float2 t = __fadd2_rn(a, make_float2(a.y, -a.x));
    return __fmul2_rn(t, make_float2(SDR_SQRT1_2, SDR_SQRT1_2));
}
// a * W8^3 = a * s(-1 - i) = s(a.y - a.x, -(a.x + a.y))
__device__ __forceinline__ float2 mul_w8_3(float2 a) {
    const float2 t = __fadd2_rn(make_float2(a.y, -a.x), make_float2(-a.x, -a.y));
    return __fmul2_rn(t, make_float2(SDR_SQRT1_2, SDR_SQRT1_2));
}

__device__ __forceinline__ void dft2(float2 &a, float2 &b) {
    const float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// natural order in, natural order out: 8 packed adds
__device__ __forceinline__ void dft4(float2 &x0, float2 &x1, float2 &x2, float2 &x3) {
    const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = csub(x1, x3);
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1 = csub_i(t1, t3);  // t1 - i*t3
    x3 = cadd_i(t1, t3);  // t1 + i*t3
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
    v[5] = mul_w8_1(v[5]);
    v[6] = mul_mi(v[6]);  // W8^2 = -i
    v[7] = mul_w8_3(v[7]);
    dft4(v[0], v[1], v[2], v[3]);
    dft4(v[4], v[5], v[6], v[7]);
}

// 16-point: n = 4*na + nb; DFT4 over na (stride 4), twiddle W16^(nb*ka), DFT4 over nb.
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
#pragma unroll
    for (int b = 0; b < 4; b++) dft4(v[b], v[b + 4], v[b + 8], v[b + 12]);
    // position 4*ka + b holds y[ka][b]; multiply by W16^(b*ka)
    const float2 w1 = make_float2(SDR_COS_PI_8, -SDR_SIN_PI_8);  // W16^1
    const float2 w3 = make_float2(SDR_SIN_PI_8, -SDR_COS_PI_8);  // W16^3
    v[5] = cmul(v[5], w1);            // ka=1: W^1
    v[6] = mul_w8_1(v[6]);            //       W^2
    v[7] = cmul(v[7], w3);            //       W^3
    v[9] = mul_w8_1(v[9]);            // ka=2: W^2
    v[10] = mul_mi(v[10]);            //       W^4 = -i
    v[11] = mul_w8_3(v[11]);          //       W^6
    v[13] = cmul(v[13], w3);          // ka=3: W^3
    v[14] = mul_w8_3(v[14]);          //       W^6
    v[15] = cmul(v[15], make_float2(-SDR_COS_PI_8, SDR_SIN_PI_8));  // W^9 = -W^1
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

// ---- 32-point transform (pass A of the N = 8192 kernel, step 1 of the N = 8192 large path) ----
// ---- W64 constants (compile-time) ---------------------------------------------------------------------------
__host__ __device__ constexpr double w64_quarter(int k) {  // cos(2 pi k / 64), k = 0..16
    constexpr double q[17] = {1.0,
                              0.99518472667219693,
                              0.98078528040323043,
                              0.95694033573220882,
                              0.92387953251128674,
                              0.88192126434835505,
                              0.83146961230254524,
                              0.77301045336273699,
                              0.70710678118654757,
                              0.63439328416364549,
                              0.55557023301960229,
                              0.47139673682599781,
                              0.38268343236508984,
                              0.29028467725446233,
                              0.19509032201612833,
                              0.09801714032956077,
                              0.0};
    return q[k];
}
__host__ __device__ constexpr double w64_cos(int k) {
    k &= 63;
    return k <= 16 ? w64_quarter(k) : k <= 32 ? -w64_quarter(32 - k) : k <= 48 ? -w64_quarter(k - 32) : w64_quarter(64 - k);
}
__host__ __device__ constexpr double w64_sin(int k) { return w64_cos(k - 16); }  // sin(t) = cos(t - pi/2)

// a * W64^E, E a compile-time exponent; the multiples of 8 cost at most two packed instructions
template <int E>
__device__ __forceinline__ float2 mul_w64(float2 a) {
    constexpr int e = E & 63;
    if constexpr (e == 0) return a;
    else if constexpr (e == 16) return mul_mi(a);
    else if constexpr (e == 32) return make_float2(-a.x, -a.y);
    else if constexpr (e == 48) return make_float2(-a.y, a.x);
    else if constexpr (e == 8) return mul_w8_1(a);
    else if constexpr (e == 24) return mul_w8_3(a);
    else {
        constexpr float c = (float)w64_cos(e), s = (float)(-w64_sin(e));
        return cmul(a, make_float2(c, s));
    }
}

template <>
struct OutIdx<32> {  // position p = 8*ka + q holds X[ka + 4*OutIdx<8>(q)]
    __host__ __device__ static constexpr int of(int p) { return (p >> 3) + 4 * OutIdx<8>::of(p & 7); }
};

template <int KA, int B>
__device__ __forceinline__ void dft32_twiddle_one(float2 (&v)[32]) {
    v[8 * KA + B] = mul_w64<2 * B * KA>(v[8 * KA + B]);  // W32^(b ka) = W64^(2 b ka)
}
template <int KA>
__device__ __forceinline__ void dft32_twiddle_row(float2 (&v)[32]) {
    dft32_twiddle_one<KA, 1>(v);
    dft32_twiddle_one<KA, 2>(v);
    dft32_twiddle_one<KA, 3>(v);
    dft32_twiddle_one<KA, 4>(v);
    dft32_twiddle_one<KA, 5>(v);
    dft32_twiddle_one<KA, 6>(v);
    dft32_twiddle_one<KA, 7>(v);
}

// 32-point forward DFT in registers: n = 8a + b; DFT4 over a, twiddle W32^(b ka), DFT8 over b.
// Natural order in; v[p] = X[OutIdx<32>::of(p)] out.  64 + 40 + 112 = 216 packed instructions.
__device__ __forceinline__ void dft32(float2 (&v)[32]) {
#pragma unroll
    for (int b = 0; b < 8; b++) dft4(v[b], v[b + 8], v[b + 16], v[b + 24]);
    dft32_twiddle_row<1>(v);
    dft32_twiddle_row<2>(v);
    dft32_twiddle_row<3>(v);
#pragma unroll
    for (int ka = 0; ka < 4; ka++) dft8(*reinterpret_cast<float2(*)[8]>(&v[8 * ka]));
}

// ---- 64-point transform (step 1 of the N = 16384 / 32768 large path) ----
template <>
struct OutIdx<64> {  // position p = 8*j + q holds X[OutIdx<8>(j) + 8*OutIdx<8>(q)]
    __host__ __device__ static constexpr int of(int p) { return OutIdx<8>::of(p >> 3) + 8 * OutIdx<8>::of(p & 7); }
};
template <int J, int B>
__device__ __forceinline__ void dft64_twiddle_one(float2 (&w)[64]) {
    w[8 * J + B] = mul_w64<B * OutIdx<8>::of(J)>(w[8 * J + B]);  // W64^(b ka), ka = OutIdx<8>(j)
}
template <int J>
__device__ __forceinline__ void dft64_twiddle_row(float2 (&w)[64]) {
    dft64_twiddle_one<J, 1>(w);
    dft64_twiddle_one<J, 2>(w);
    dft64_twiddle_one<J, 3>(w);
    dft64_twiddle_one<J, 4>(w);
    dft64_twiddle_one<J, 5>(w);
    dft64_twiddle_one<J, 6>(w);
    dft64_twiddle_one<J, 7>(w);
}
// 64-point forward DFT in registers: n = 8a + b; DFT8 over a, twiddle W64^(b ka), DFT8 over b.
// Natural order in; v[p] = X[OutIdx<64>::of(p)] out.
__device__ __forceinline__ void dft64(float2 (&v)[64]) {
    float2 w[64];
#pragma unroll
    for (int b = 0; b < 8; b++) {
        float2 t[8];
#pragma unroll
        for (int a = 0; a < 8; a++) t[a] = v[b + 8 * a];
        dft8(t);
#pragma unroll
        for (int j = 0; j < 8; j++) w[8 * j + b] = t[j];  // slot j holds ka = OutIdx<8>(j)
    }
    dft64_twiddle_row<1>(w);
    dft64_twiddle_row<2>(w);
    dft64_twiddle_row<3>(w);
    dft64_twiddle_row<4>(w);
    dft64_twiddle_row<5>(w);
    dft64_twiddle_row<6>(w);
    dft64_twiddle_row<7>(w);
#pragma unroll
    for (int j = 0; j < 8; j++) dft8(*reinterpret_cast<float2(*)[8]>(&w[8 * j]));
#pragma unroll
    for (int p = 0; p < 64; p++) v[p] = w[p];
}

}  // namespace sdr
