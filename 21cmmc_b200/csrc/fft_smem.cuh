// Shared-memory Stockham FFT building blocks (forward transform, exp(-2 pi i jk/n)).
//
// Layout convention of the "column engine": a tile holds `ncols` independent sequences of length
// n; element i of column c lives at buf[i*ld + c] (columns contiguous), so consecutive threads
// touch consecutive shared-memory words for every butterfly leg -> bank-conflict free.
// Stages ping-pong between two buffers (autosort, no bit reversal).
#pragma once
#ifdef PS21_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

namespace ps21 {

// exact n / d for 0 <= n < 65536 by multiply-shift (constants from the host: make_fastdiv in ps21_plan.cu)
struct FastDiv {
    unsigned mul, shift;
};
__device__ __forceinline__ int fast_div(int n, FastDiv f) {
    return (int)(((unsigned long long)(unsigned)n * f.mul) >> f.shift);
}

struct FftFactors {
    int n;
    int nstages;
    int radix[14];
    FastDiv div_ns[14];  // division by Ns = product of the radices before stage s
    FastDiv div_nb[14];  // division by n / radix[s]
};

// Complex helpers.  On sm_100 the adds/subs and the twiddle multiplies use the packed fp32x2
// instructions (FADD2 / FFMA2 / FMUL2, with register-broadcast and lane-swap operand selectors), which
// halves the instruction count of the butterflies; the emulated test build uses the scalar forms.
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(PS21_EMU)
#define PS21_PACKED_F32X2 1
#endif

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
#ifdef PS21_PACKED_F32X2
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// a + (-i) b  and  a - (-i) b
__device__ __forceinline__ float2 cadd_mi(float2 a, float2 b) {
    return __ffma2_rn(make_float2(b.y, b.x), make_float2(1.f, -1.f), a);
}
__device__ __forceinline__ float2 csub_mi(float2 a, float2 b) {
    return __ffma2_rn(make_float2(b.y, b.x), make_float2(-1.f, 1.f), a);
}
// a + s*b, s real
__device__ __forceinline__ float2 caxpy(float s, float2 b, float2 a) { return __ffma2_rn(b, make_float2(s, s), a); }
// s*a, s real
__device__ __forceinline__ float2 cscale(float s, float2 a) { return __fmul2_rn(a, make_float2(s, s)); }
// a * (c - i s)
__device__ __forceinline__ float2 crot(float2 a, float c, float s) {
    return __ffma2_rn(make_float2(a.y, a.x), make_float2(s, -s), __fmul2_rn(a, make_float2(c, c)));
}
// a * w with the twiddle stored as (w.x, w.y, -w.y, w.x)
__device__ __forceinline__ float2 cmul_tw(float2 a, float4 w) {
    return __ffma2_rn(make_float2(a.x, a.x), make_float2(w.x, w.y),
                      __fmul2_rn(make_float2(a.y, a.y), make_float2(w.z, w.w)));
}
// a * w with a plain (w.x, w.y) twiddle: three packed instructions, half the table bytes of cmul_tw
__device__ __forceinline__ float2 cmul_w2(float2 a, float2 w) {
    const float2 rot = __fmul2_rn(make_float2(a.y, a.x), make_float2(-1.f, 1.f));  // i*a
    return __ffma2_rn(rot, make_float2(w.y, w.y), __fmul2_rn(a, make_float2(w.x, w.x)));
}
#else
__device__ __forceinline__ float2 cmul_w2(float2 a, float2 w) { return cmul(a, w); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cadd_mi(float2 a, float2 b) { return make_float2(a.x + b.y, a.y - b.x); }
__device__ __forceinline__ float2 csub_mi(float2 a, float2 b) { return make_float2(a.x - b.y, a.y + b.x); }
__device__ __forceinline__ float2 caxpy(float s, float2 b, float2 a) { return make_float2(a.x + s * b.x, a.y + s * b.y); }
__device__ __forceinline__ float2 cscale(float s, float2 a) { return make_float2(s * a.x, s * a.y); }
__device__ __forceinline__ float2 crot(float2 a, float c, float s) {
    return make_float2(c * a.x + s * a.y, c * a.y - s * a.x);
}
__device__ __forceinline__ float2 cmul_tw(float2 a, float4 w) {
    return make_float2(a.x * w.x + a.y * w.z, a.x * w.y + a.y * w.w);
}
#endif
// multiply by -i
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

template <int R>
__device__ __forceinline__ void small_dft(float2 (&v)[R]);

template <>
__device__ __forceinline__ void small_dft<2>(float2 (&v)[2]) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}

template <>
__device__ __forceinline__ void small_dft<3>(float2 (&v)[3]) {
    const float c = 0.86602540378443864676f;  // sin(pi/3)
    float2 s = cadd(v[1], v[2]);
    float2 d = csub(v[1], v[2]);
    float2 m = caxpy(-0.5f, s, v[0]);
    v[0] = cadd(v[0], s);
    v[1] = cadd_mi(m, make_float2(c * d.x, c * d.y));
    v[2] = csub_mi(m, make_float2(c * d.x, c * d.y));
}

template <>
__device__ __forceinline__ void small_dft<4>(float2 (&v)[4]) {
    float2 t0 = cadd(v[0], v[2]);
    float2 t1 = csub(v[0], v[2]);
    float2 t2 = cadd(v[1], v[3]);
    float2 t3 = csub(v[1], v[3]);
    v[0] = cadd(t0, t2);
    v[2] = csub(t0, t2);
    v[1] = cadd_mi(t1, t3);
    v[3] = csub_mi(t1, t3);
}

template <>
__device__ __forceinline__ void small_dft<5>(float2 (&v)[5]) {
    const float c1 = 0.30901699437494742410f;   // cos(2pi/5)
    const float c2 = -0.80901699437494742410f;  // cos(4pi/5)
    const float s1 = 0.95105651629515357212f;   // sin(2pi/5)
    const float s2 = 0.58778525229247312917f;   // sin(4pi/5)
    float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    float2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    float2 r1 = caxpy(c2, a2, caxpy(c1, a1, v[0]));
    float2 r2 = caxpy(c1, a2, caxpy(c2, a1, v[0]));
    float2 q1 = caxpy(s2, b2, cscale(s1, b1));    // packed: two instructions each (ncu: the scalar form was 5.6 % of
    float2 q2 = caxpy(-s1, b2, cscale(s2, b1));   // the 200^3 pass-2 instructions)
    v[0] = cadd(v[0], cadd(a1, a2));
    v[1] = cadd_mi(r1, q1);
    v[4] = csub_mi(r1, q1);
    v[2] = cadd_mi(r2, q2);
    v[3] = csub_mi(r2, q2);
}

template <>
__device__ __forceinline__ void small_dft<8>(float2 (&v)[8]) {
    const float h = 0.70710678118654752440f;
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    small_dft<4>(e);
    small_dft<4>(o);
    // w8^1 = (1 - i)/sqrt2, w8^2 = -i, w8^3 = (-1 - i)/sqrt2
    const float2 u1 = cadd_mi(o[1], o[1]);  // (x + y, y - x) = sqrt2 * o1 * w8
    const float2 u3 = csub_mi(o[3], o[3]);  // (x - y, x + y) = -sqrt2 * o3 * w8^3
    v[0] = cadd(e[0], o[0]);
    v[4] = csub(e[0], o[0]);
    v[1] = caxpy(h, u1, e[1]);
    v[5] = caxpy(-h, u1, e[1]);
    v[2] = cadd_mi(e[2], o[2]);
    v[6] = csub_mi(e[2], o[2]);
    v[3] = caxpy(-h, u3, e[3]);
    v[7] = caxpy(h, u3, e[3]);
}

// 16-point DFT as 4x4 (used by the fused power-of-two kernels).
template <>
__device__ __forceinline__ void small_dft<16>(float2 (&v)[16]) {
    const float c1 = 0.92387953251128675613f;  // cos(pi/8)
    const float s1 = 0.38268343236508977173f;  // sin(pi/8)
    const float h = 0.70710678118654752440f;
    float2 a[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float2 t[4] = {v[j], v[j + 4], v[j + 8], v[j + 12]};
        small_dft<4>(t);
#pragma unroll
        for (int q = 0; q < 4; ++q) a[j][q] = t[q];
    }
    // a[j][q] *= w16^(j q), w16^m = cos(m pi/8) - i sin(m pi/8)
    a[1][1] = crot(a[1][1], c1, s1);    // w^1
    a[1][2] = crot(a[1][2], h, h);      // w^2
    a[1][3] = crot(a[1][3], s1, c1);    // w^3
    a[2][1] = crot(a[2][1], h, h);      // w^2
    a[2][2] = cmul_mi(a[2][2]);         // w^4 = -i
    a[2][3] = crot(a[2][3], -h, h);     // w^6
    a[3][1] = crot(a[3][1], s1, c1);    // w^3
    a[3][2] = crot(a[3][2], -h, h);     // w^6
    a[3][3] = crot(a[3][3], -c1, -s1);  // w^9
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float2 t[4] = {a[0][q], a[1][q], a[2][q], a[3][q]};
        small_dft<4>(t);
#pragma unroll
        for (int p = 0; p < 4; ++p) v[q + 4 * p] = t[p];
    }
}

// 25-point DFT as 5x5: i = j + 5 i2, k = q + 5 p;  w25^(ik) = w25^(jq) w5^(jp) w5^(i2 q).
// (200 = 25*8, 100 = 25*4, 50 = 25*2: box sizes in multiples of 50 are common, and a 25-point butterfly saves a stage.)
template <>
__device__ __forceinline__ void small_dft<25>(float2 (&v)[25]) {
    // cos / sin of 2 pi m / 25, m = 0..16
    constexpr float W[17][2] = {
        {1.00000000000000000000f, 0.00000000000000000000f},
        {0.96858316112863107605f, 0.24868988716485479484f},
        {0.87630668004386358394f, 0.48175367410171532345f},
        {0.72896862742141155245f, 0.68454710592868861507f},
        {0.53582679497899654564f, 0.84432792550201507531f},
        {0.30901699437494745126f, 0.95105651629515353118f},
        {0.06279051952931352654f, 0.99802672842827155897f},
        {-0.18738131458572460097f, 0.98228725072868872115f},
        {-0.42577929156507271502f, 0.90482705246601946580f},
        {-0.63742398974868974548f, 0.77051324277578925326f},
        {-0.80901699437494734024f, 0.58778525229247324813f},
        {-0.92977648588825134723f, 0.36812455268467814129f},
        {-0.99211470131447776488f, 0.12533323356430453588f},
        {-0.99211470131447787590f, -0.12533323356430428608f},
        {-0.92977648588825145826f, -0.36812455268467791925f},
        {-0.80901699437494778433f, -0.58778525229247269301f},
        {-0.63742398974868952344f, -0.77051324277578936428f},
    };
    float2 a[5][5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        float2 t[5];
#pragma unroll
        for (int i2 = 0; i2 < 5; ++i2) t[i2] = v[j + 5 * i2];
        small_dft<5>(t);
#pragma unroll
        for (int q = 0; q < 5; ++q) a[j][q] = (j * q == 0) ? t[q] : crot(t[q], W[j * q][0], W[j * q][1]);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        float2 t[5] = {a[0][q], a[1][q], a[2][q], a[3][q], a[4][q]};
        small_dft<5>(t);
#pragma unroll
        for (int p = 0; p < 5; ++p) v[q + 5 * p] = t[p];
    }
}

// 32-point DFT as 8x4: i = j + 4 i2, k = q + 8 p;  w32^(ik) = w32^(jq) w4^(jp) w8^(i2 q).
template <>
__device__ __forceinline__ void small_dft<32>(float2 (&v)[32]) {
    // cos / sin of m pi / 16
    constexpr float C1 = 0.98078528040323044913f, S1 = 0.19509032201612826785f;
    constexpr float C2 = 0.92387953251128675613f, S2 = 0.38268343236508977173f;
    constexpr float C3 = 0.83146961230254523708f, S3 = 0.55557023301960222474f;
    constexpr float H = 0.70710678118654752440f;
    float2 a[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float2 t[8];
#pragma unroll
        for (int i2 = 0; i2 < 8; ++i2) t[i2] = v[j + 4 * i2];
        small_dft<8>(t);
#pragma unroll
        for (int q = 0; q < 8; ++q) a[j][q] = t[q];
    }
    // a[j][q] *= w32^(j q);  crot(x, cos, sin) multiplies by cos - i sin
    a[1][1] = crot(a[1][1], C1, S1);    // 1
    a[1][2] = crot(a[1][2], C2, S2);    // 2
    a[1][3] = crot(a[1][3], C3, S3);    // 3
    a[1][4] = crot(a[1][4], H, H);      // 4
    a[1][5] = crot(a[1][5], S3, C3);    // 5
    a[1][6] = crot(a[1][6], S2, C2);    // 6
    a[1][7] = crot(a[1][7], S1, C1);    // 7
    a[2][1] = crot(a[2][1], C2, S2);    // 2
    a[2][2] = crot(a[2][2], H, H);      // 4
    a[2][3] = crot(a[2][3], S2, C2);    // 6
    a[2][4] = cmul_mi(a[2][4]);         // 8 = -i
    a[2][5] = crot(a[2][5], -S2, C2);   // 10
    a[2][6] = crot(a[2][6], -H, H);     // 12
    a[2][7] = crot(a[2][7], -C2, S2);   // 14
    a[3][1] = crot(a[3][1], C3, S3);    // 3
    a[3][2] = crot(a[3][2], S2, C2);    // 6
    a[3][3] = crot(a[3][3], -S1, C1);   // 9
    a[3][4] = crot(a[3][4], -H, H);     // 12
    a[3][5] = crot(a[3][5], -C1, S1);   // 15
    a[3][6] = crot(a[3][6], -C2, -S2);  // 18
    a[3][7] = crot(a[3][7], -S3, -C3);  // 21
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float2 t[4] = {a[0][q], a[1][q], a[2][q], a[3][q]};
        small_dft<4>(t);
#pragma unroll
        for (int p = 0; p < 4; ++p) v[q + 8 * p] = t[p];
    }
}



// Composite radices 6, 9, 10, 12, 15 (two-factor Cooley-Tukey in registers): i = j + B i2, k = q + A p,
// w_N^(ik) = w_N^(jq) w_B^(jp) w_A^(i2 q).  They keep 100 = 10 * 10, 50 = 5 * 10, 150 = 10 * 15 ... at two stages.
template <int N>
__device__ __forceinline__ void lit_roots(float (&C)[N], float (&S)[N]);   // cos / sin of 2 pi m / N
template <>
__device__ __forceinline__ void lit_roots<6>(float (&C)[6], float (&S)[6]) {
    C[0] = 1.00000000000000000000f; S[0] = 0.00000000000000000000f;
    C[1] = 0.50000000000000011102f; S[1] = 0.86602540378443859659f;
    C[2] = -0.49999999999999977796f; S[2] = 0.86602540378443870761f;
    C[3] = -1.00000000000000000000f; S[3] = 0.00000000000000012246f;
    C[4] = -0.50000000000000044409f; S[4] = -0.86602540378443837454f;
    C[5] = 0.50000000000000011102f; S[5] = -0.86602540378443859659f;
}
template <>
__device__ __forceinline__ void lit_roots<9>(float (&C)[9], float (&S)[9]) {
    C[0] = 1.00000000000000000000f; S[0] = 0.00000000000000000000f;
    C[1] = 0.76604444311897801345f; S[1] = 0.64278760968653925190f;
    C[2] = 0.17364817766693041445f; S[2] = 0.98480775301220802032f;
    C[3] = -0.49999999999999977796f; S[3] = 0.86602540378443870761f;
    C[4] = -0.93969262078590831688f; S[4] = 0.34202014332566887944f;
    C[5] = -0.93969262078590842791f; S[5] = -0.34202014332566865740f;
    C[6] = -0.50000000000000044409f; S[6] = -0.86602540378443837454f;
    C[7] = 0.17364817766692997036f; S[7] = -0.98480775301220813134f;
    C[8] = 0.76604444311897779141f; S[8] = -0.64278760968653958496f;
}
template <>
__device__ __forceinline__ void lit_roots<10>(float (&C)[10], float (&S)[10]) {
    C[0] = 1.00000000000000000000f; S[0] = 0.00000000000000000000f;
    C[1] = 0.80901699437494745126f; S[1] = 0.58778525229247313710f;
    C[2] = 0.30901699437494745126f; S[2] = 0.95105651629515353118f;
    C[3] = -0.30901699437494734024f; S[3] = 0.95105651629515364220f;
    C[4] = -0.80901699437494734024f; S[4] = 0.58778525229247324813f;
    C[5] = -1.00000000000000000000f; S[5] = 0.00000000000000012246f;
    C[6] = -0.80901699437494756229f; S[6] = -0.58778525229247302608f;
    C[7] = -0.30901699437494756229f; S[7] = -0.95105651629515353118f;
    C[8] = 0.30901699437494722922f; S[8] = -0.95105651629515364220f;
    C[9] = 0.80901699437494734024f; S[9] = -0.58778525229247335915f;
}
template <>
__device__ __forceinline__ void lit_roots<12>(float (&C)[12], float (&S)[12]) {
    C[0] = 1.00000000000000000000f; S[0] = 0.00000000000000000000f;
    C[1] = 0.86602540378443870761f; S[1] = 0.49999999999999994449f;
    C[2] = 0.50000000000000011102f; S[2] = 0.86602540378443859659f;
    C[3] = 0.00000000000000006123f; S[3] = 1.00000000000000000000f;
    C[4] = -0.49999999999999977796f; S[4] = 0.86602540378443870761f;
    C[5] = -0.86602540378443870761f; S[5] = 0.49999999999999994449f;
    C[6] = -1.00000000000000000000f; S[6] = 0.00000000000000012246f;
    C[7] = -0.86602540378443881863f; S[7] = -0.49999999999999972244f;
    C[8] = -0.50000000000000044409f; S[8] = -0.86602540378443837454f;
    C[9] = -0.00000000000000018370f; S[9] = -1.00000000000000000000f;
    C[10] = 0.50000000000000011102f; S[10] = -0.86602540378443859659f;
    C[11] = 0.86602540378443837454f; S[11] = -0.50000000000000044409f;
}
template <>
__device__ __forceinline__ void lit_roots<15>(float (&C)[15], float (&S)[15]) {
    C[0] = 1.00000000000000000000f; S[0] = 0.00000000000000000000f;
    C[1] = 0.91354545764260086660f; S[1] = 0.40673664307580015276f;
    C[2] = 0.66913060635885823757f; S[2] = 0.74314482547739413310f;
    C[3] = 0.30901699437494745126f; S[3] = 0.95105651629515353118f;
    C[4] = -0.10452846326765333207f; S[4] = 0.99452189536827340088f;
    C[5] = -0.49999999999999977796f; S[5] = 0.86602540378443870761f;
    C[6] = -0.80901699437494734024f; S[6] = 0.58778525229247324813f;
    C[7] = -0.97814760073380568883f; S[7] = 0.20791169081775931482f;
    C[8] = -0.97814760073380568883f; S[8] = -0.20791169081775906502f;
    C[9] = -0.80901699437494756229f; S[9] = -0.58778525229247302608f;
    C[10] = -0.50000000000000044409f; S[10] = -0.86602540378443837454f;
    C[11] = -0.10452846326765423413f; S[11] = -0.99452189536827328986f;
    C[12] = 0.30901699437494722922f; S[12] = -0.95105651629515364220f;
    C[13] = 0.66913060635885845961f; S[13] = -0.74314482547739402207f;
    C[14] = 0.91354545764260097762f; S[14] = -0.40673664307580015276f;
}
template <int A, int B>
__device__ __forceinline__ void small_dft_ct(float2 (&v)[A * B]) {
    constexpr int N = A * B;
    float C[N], S[N];   // literal tables: folded into immediates once the loops are unrolled
    lit_roots<N>(C, S);
    float2 a[B][A];
#pragma unroll
    for (int j = 0; j < B; ++j) {
        float2 t[A];
#pragma unroll
        for (int i2 = 0; i2 < A; ++i2) t[i2] = v[j + B * i2];
        small_dft<A>(t);
#pragma unroll
        for (int q = 0; q < A; ++q) a[j][q] = (j * q == 0) ? t[q] : crot(t[q], C[(j * q) % N], S[(j * q) % N]);
    }
#pragma unroll
    for (int q = 0; q < A; ++q) {
        float2 t[B];
#pragma unroll
        for (int j = 0; j < B; ++j) t[j] = a[j][q];
        small_dft<B>(t);
#pragma unroll
        for (int p = 0; p < B; ++p) v[q + A * p] = t[p];
    }
}
template <>
__device__ __forceinline__ void small_dft<6>(float2 (&v)[6]) { small_dft_ct<2, 3>(v); }
template <>
__device__ __forceinline__ void small_dft<9>(float2 (&v)[9]) { small_dft_ct<3, 3>(v); }
template <>
__device__ __forceinline__ void small_dft<10>(float2 (&v)[10]) { small_dft_ct<2, 5>(v); }
template <>
__device__ __forceinline__ void small_dft<12>(float2 (&v)[12]) { small_dft_ct<4, 3>(v); }
template <>
__device__ __forceinline__ void small_dft<15>(float2 (&v)[15]) { small_dft_ct<3, 5>(v); }

// Odd prime radices 7, 11, 13 (lightcone chunk lengths are arbitrary: 2 * 7 * 13 = 182, 2 * 11 * 13 = 286, ...):
// X_k = v_0 + sum_m (v_m + v_{P-m}) cos(2 pi m k / P) - i (v_m - v_{P-m}) sin(2 pi m k / P),  X_{P-k} = conj-pair form.
template <int P>
__device__ __forceinline__ void odd_roots(float (&C)[P], float (&S)[P]);   // cos / sin of 2 pi m / P
template <>
__device__ __forceinline__ void odd_roots<7>(float (&C)[7], float (&S)[7]) {
    C[0] = 1.00000000000000000000f; S[0] = 0.00000000000000000000f;
    C[1] = 0.62348980185873359439f; S[1] = 0.78183148246802980363f;
    C[2] = -0.22252093395631433737f; S[2] = 0.97492791218182361934f;
    C[3] = -0.90096886790241903498f; S[3] = 0.43388373911755823142f;
    C[4] = -0.90096886790241914600f; S[4] = -0.43388373911755800938f;
    C[5] = -0.22252093395631458717f; S[5] = -0.97492791218182361934f;
    C[6] = 0.62348980185873337234f; S[6] = -0.78183148246802991466f;
}
template <>
__device__ __forceinline__ void odd_roots<11>(float (&C)[11], float (&S)[11]) {
    C[0] = 1.00000000000000000000f; S[0] = 0.00000000000000000000f;
    C[1] = 0.84125353283118120551f; S[1] = 0.54064081745559755543f;
    C[2] = 0.41541501300188643508f; S[2] = 0.90963199535451833011f;
    C[3] = -0.14231483827328500480f; S[3] = 0.98982144188093279524f;
    C[4] = -0.65486073394528498959f; S[4] = 0.75574957435425826890f;
    C[5] = -0.95949297361449736865f; S[5] = 0.28173255684142967104f;
    C[6] = -0.95949297361449747967f; S[6] = -0.28173255684142939348f;
    C[7] = -0.65486073394528521163f; S[7] = -0.75574957435425815788f;
    C[8] = -0.14231483827328522684f; S[8] = -0.98982144188093268422f;
    C[9] = 0.41541501300188604651f; S[9] = -0.90963199535451855215f;
    C[10] = 0.84125353283118120551f; S[10] = -0.54064081745559744441f;
}
template <>
__device__ __forceinline__ void odd_roots<13>(float (&C)[13], float (&S)[13]) {
    C[0] = 1.00000000000000000000f; S[0] = 0.00000000000000000000f;
    C[1] = 0.88545602565320991051f; S[1] = 0.46472317204376850652f;
    C[2] = 0.56806474673115592289f; S[2] = 0.82298386589365635224f;
    C[3] = 0.12053668025532300601f; S[3] = 0.99270887409805397272f;
    C[4] = -0.35460488704253545489f; S[4] = 0.93501624268541483342f;
    C[5] = -0.74851074817110119231f; S[5] = 0.66312265824079519305f;
    C[6] = -0.97094181742605201180f; S[6] = 0.23931566428755768339f;
    C[7] = -0.97094181742605212282f; S[7] = -0.23931566428755743359f;
    C[8] = -0.74851074817110130333f; S[8] = -0.66312265824079497101f;
    C[9] = -0.35460488704253589898f; S[9] = -0.93501624268541472240f;
    C[10] = 0.12053668025532320029f; S[10] = -0.99270887409805397272f;
    C[11] = 0.56806474673115481266f; S[11] = -0.82298386589365701838f;
    C[12] = 0.88545602565321002153f; S[12] = -0.46472317204376839550f;
}
template <int P>
__device__ __forceinline__ void small_dft_odd(float2 (&v)[P]) {
    constexpr int H = (P - 1) / 2;
    float C[P], S[P];   // literal tables: folded into immediates once the loops below are unrolled
    odd_roots<P>(C, S);
    float2 a[H], b[H];
#pragma unroll
    for (int m = 1; m <= H; ++m) {
        a[m - 1] = cadd(v[m], v[P - m]);
        b[m - 1] = csub(v[m], v[P - m]);
    }
    const float2 v0 = v[0];
    float2 s0 = v0;
#pragma unroll
    for (int m = 0; m < H; ++m) s0 = cadd(s0, a[m]);
    v[0] = s0;
#pragma unroll
    for (int k = 1; k <= H; ++k) {
        float2 re = v0, im = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 1; m <= H; ++m) {
            re = caxpy(C[(m * k) % P], a[m - 1], re);
            im = caxpy(S[(m * k) % P], b[m - 1], im);
        }
        v[k] = cadd_mi(re, im);       // re - i im
        v[P - k] = csub_mi(re, im);   // re + i im
    }
}
template <>
__device__ __forceinline__ void small_dft<7>(float2 (&v)[7]) { small_dft_odd<7>(v); }
template <>
__device__ __forceinline__ void small_dft<11>(float2 (&v)[11]) { small_dft_odd<11>(v); }
template <>
__device__ __forceinline__ void small_dft<13>(float2 (&v)[13]) { small_dft_odd<13>(v); }

// One Stockham stage of radix R over 2^lc columns (lc = log2 of the tile width; ragged tiles still run all
// columns of the tile -- columns are independent, the surplus ones are never stored).  Ns = product of the
// radices already applied.  No integer division: the column split is a shift, j % Ns a multiply-shift.
template <int R>
__device__ __forceinline__ void stage_radix(const float2* __restrict__ in, float2* __restrict__ out, int n,
                                            int Ns, FastDiv dns, int ld, int lc, const float2* __restrict__ tw) {
    const int nb = n / R;
    const int tws = n / (Ns * R);
    const int cmask = (1 << lc) - 1;
    // byte strides between the R inputs / outputs of a butterfly: addresses advance by one add per element
    const int sin = nb * ld * (int)sizeof(float2), sout = Ns * ld * (int)sizeof(float2);
    for (int task = threadIdx.x; task < (nb << lc); task += blockDim.x) {
        const int c = task & cmask;
        const int j = task >> lc;
        const int k = j - fast_div(j, dns) * Ns;
        float2 v[R];
        const char* pi = reinterpret_cast<const char*>(in + (j * ld + c));
#pragma unroll
        for (int m = 0; m < R; ++m) {
            v[m] = *reinterpret_cast<const float2*>(pi);
            pi += sin;
        }
        if (Ns > 1) {
            const float2* pt = tw + k * tws;
            const int step = k * tws;
#pragma unroll
            for (int m = 1; m < R; ++m) {
                v[m] = cmul_w2(v[m], __ldg(pt));
                pt += step;
            }
        }
        small_dft<R>(v);
        char* po = reinterpret_cast<char*>(out + (((j - k) * R + k) * ld + c));
#pragma unroll
        for (int m = 0; m < R; ++m) {
            *reinterpret_cast<float2*>(po) = v[m];
            po += sout;
        }
    }
}

// Any radix p (prime factors > 5, e.g. chunk lengths such as 334 = 2 * 167): one output per task.
__device__ __forceinline__ void stage_generic(const float2* __restrict__ in, float2* __restrict__ out, int n,
                                              int Ns, FastDiv dns, FastDiv dnb, int p, int ld, int lc,
                                              const float2* __restrict__ tw) {
    const int nb = n / p;
    const int tws = n / (Ns * p);
    const int cmask = (1 << lc) - 1;
    for (int task = threadIdx.x; task < (n << lc); task += blockDim.x) {
        const int c = task & cmask;
        const int q = task >> lc;
        const int mo = fast_div(q, dnb);
        const int j = q - mo * nb;
        const int k = j - fast_div(j, dns) * Ns;
        // twiddle of term mi: w_n^(mi*k*tws) * w_p^(mi*mo) = w_n^(mi * inc) with inc = k*tws + mo*nb (n = p*nb):
        // one running index, reduced mod n by a compare and subtract
        int inc = k * tws + mo * nb;
        if (inc >= n) inc -= n;
        float2 acc = make_float2(0.f, 0.f);
        int idx = 0;
        const char* pi = reinterpret_cast<const char*>(in + (j * ld + c));
        const int sin = nb * ld * (int)sizeof(float2);
        for (int mi = 0; mi < p; ++mi) {
            acc = cadd(acc, cmul_w2(*reinterpret_cast<const float2*>(pi), __ldg(&tw[idx])));
            pi += sin;
            idx += inc;
            if (idx >= n) idx -= n;
        }
        out[((j - k) * p + k + mo * Ns) * ld + c] = acc;
    }
}

// Full column FFT over 2^lc columns.  Returns the buffer that holds the result (a or b).  Ends with a barrier.
// R32: also instantiate the radix-32 and radix-25 stages (512 = 32*16, 200 = 25*8).  They need ~100 registers, so the kernels
// that may meet it are compiled separately and the common ones keep their occupancy.
template <bool R32 = false>
__device__ __forceinline__ float2* column_fft(float2* a, float2* b, const FftFactors& f, int ld, int lc,
                                              const float2* __restrict__ tw) {
    float2* in = a;
    float2* out = b;
    int Ns = 1;
    for (int s = 0; s < f.nstages; ++s) {
        const int r = f.radix[s];
        if (R32 && r == 32) {   // compile-time dead in the common kernels: their code does not change
            stage_radix<32>(in, out, f.n, Ns, f.div_ns[s], ld, lc, tw);
        } else if (R32 && r == 25) {
            stage_radix<25>(in, out, f.n, Ns, f.div_ns[s], ld, lc, tw);
        } else {
            switch (r) {
                case 2: stage_radix<2>(in, out, f.n, Ns, f.div_ns[s], ld, lc, tw); break;
                case 3: stage_radix<3>(in, out, f.n, Ns, f.div_ns[s], ld, lc, tw); break;
                case 4: stage_radix<4>(in, out, f.n, Ns, f.div_ns[s], ld, lc, tw); break;
                case 5: stage_radix<5>(in, out, f.n, Ns, f.div_ns[s], ld, lc, tw); break;
                case 8: stage_radix<8>(in, out, f.n, Ns, f.div_ns[s], ld, lc, tw); break;
                case 16: stage_radix<16>(in, out, f.n, Ns, f.div_ns[s], ld, lc, tw); break;
                default: stage_generic(in, out, f.n, Ns, f.div_ns[s], f.div_nb[s], r, ld, lc, tw); break;
            }
        }
        __syncthreads();
        float2* t = in;
        in = out;
        out = t;
        Ns *= r;
    }
    return in;
}

}  // namespace ps21
