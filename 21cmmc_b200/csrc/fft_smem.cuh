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

struct FftFactors {
    int n;
    int nstages;
    int radix[14];
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
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
    float2 m = make_float2(v[0].x - 0.5f * s.x, v[0].y - 0.5f * s.y);
    float2 jd = make_float2(c * d.y, -c * d.x);  // -i*c*d
    v[0] = cadd(v[0], s);
    v[1] = cadd(m, jd);
    v[2] = csub(m, jd);
}

template <>
__device__ __forceinline__ void small_dft<4>(float2 (&v)[4]) {
    float2 t0 = cadd(v[0], v[2]);
    float2 t1 = csub(v[0], v[2]);
    float2 t2 = cadd(v[1], v[3]);
    float2 t3 = cmul_mi(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
}

template <>
__device__ __forceinline__ void small_dft<5>(float2 (&v)[5]) {
    const float c1 = 0.30901699437494742410f;   // cos(2pi/5)
    const float c2 = -0.80901699437494742410f;  // cos(4pi/5)
    const float s1 = 0.95105651629515357212f;   // sin(2pi/5)
    const float s2 = 0.58778525229247312917f;   // sin(4pi/5)
    float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    float2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    float2 r1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    float2 r2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    float2 q1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
    float2 q2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
    float2 x0 = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
    float2 jq1 = cmul_mi(q1), jq2 = cmul_mi(q2);
    v[0] = x0;
    v[1] = cadd(r1, jq1);
    v[4] = csub(r1, jq1);
    v[2] = cadd(r2, jq2);
    v[3] = csub(r2, jq2);
}

template <>
__device__ __forceinline__ void small_dft<8>(float2 (&v)[8]) {
    const float h = 0.70710678118654752440f;
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    small_dft<4>(e);
    small_dft<4>(o);
    // w8^1 = (1 - i)/sqrt2, w8^2 = -i, w8^3 = (-1 - i)/sqrt2
    float2 o1 = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
    float2 o2 = cmul_mi(o[2]);
    float2 o3 = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
    v[0] = cadd(e[0], o[0]);
    v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);
    v[5] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);
    v[6] = csub(e[2], o2);
    v[3] = cadd(e[3], o3);
    v[7] = csub(e[3], o3);
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
    // twiddle a[j][q] *= w16^(j*q)
    // j=1: q=1 w^1, q=2 w^2, q=3 w^3 ; j=2: w^2, w^4, w^6 ; j=3: w^3, w^6, w^9
    auto w1 = [&](float2 x) { return make_float2(c1 * x.x + s1 * x.y, c1 * x.y - s1 * x.x); };
    auto w2 = [&](float2 x) { return make_float2(h * (x.x + x.y), h * (x.y - x.x)); };
    auto w3 = [&](float2 x) { return make_float2(s1 * x.x + c1 * x.y, s1 * x.y - c1 * x.x); };
    auto w6 = [&](float2 x) { return make_float2(h * (x.y - x.x), -h * (x.x + x.y)); };
    auto w9 = [&](float2 x) { return make_float2(-c1 * x.x - s1 * x.y, s1 * x.x - c1 * x.y); };
    a[1][1] = w1(a[1][1]);
    a[1][2] = w2(a[1][2]);
    a[1][3] = w3(a[1][3]);
    a[2][1] = w2(a[2][1]);
    a[2][2] = cmul_mi(a[2][2]);
    a[2][3] = w6(a[2][3]);
    a[3][1] = w3(a[3][1]);
    a[3][2] = w6(a[3][2]);
    a[3][3] = w9(a[3][3]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float2 t[4] = {a[0][q], a[1][q], a[2][q], a[3][q]};
        small_dft<4>(t);
#pragma unroll
        for (int p = 0; p < 4; ++p) v[q + 4 * p] = t[p];
    }
}

// One Stockham stage of radix R over `ncols` columns.  Ns = product of the radices already applied.
template <int R>
__device__ __forceinline__ void stage_radix(const float2* __restrict__ in, float2* __restrict__ out, int n,
                                            int Ns, int ld, int ncols, const float2* __restrict__ tw) {
    const int nb = n / R;
    const int tws = n / (Ns * R);
    for (int task = threadIdx.x; task < nb * ncols; task += blockDim.x) {
        const int c = task % ncols;
        const int j = task / ncols;
        const int k = j % Ns;
        float2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = in[(j + m * nb) * ld + c];
        if (Ns > 1) {
            const int step = k * tws;
#pragma unroll
            for (int m = 1; m < R; ++m) v[m] = cmul(v[m], __ldg(&tw[m * step]));
        }
        small_dft<R>(v);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int m = 0; m < R; ++m) out[(j0 + m * Ns) * ld + c] = v[m];
    }
}

// Any radix p (prime factors > 5, e.g. chunk lengths such as 334 = 2 * 167): one output per task.
__device__ __forceinline__ void stage_generic(const float2* __restrict__ in, float2* __restrict__ out, int n,
                                              int Ns, int p, int ld, int ncols,
                                              const float2* __restrict__ tw) {
    const int nb = n / p;
    const int tws = n / (Ns * p);
    for (int task = threadIdx.x; task < n * ncols; task += blockDim.x) {
        const int c = task % ncols;
        const int q = task / ncols;
        const int j = q % nb;
        const int mo = q / nb;
        const int k = j % Ns;
        const int step = k * tws;
        float2 acc = make_float2(0.f, 0.f);
        int rot = 0;  // (mi*mo) % p
        for (int mi = 0; mi < p; ++mi) {
            int idx = mi * step + rot * nb;
            if (idx >= n) idx -= n;
            acc = cadd(acc, cmul(in[(j + mi * nb) * ld + c], __ldg(&tw[idx])));
            rot += mo;
            if (rot >= p) rot -= p;
        }
        out[((j - k) * p + k + mo * Ns) * ld + c] = acc;
    }
}

// Full column FFT.  Returns the buffer that holds the result (a or b).  Ends with __syncthreads().
__device__ __forceinline__ float2* column_fft(float2* a, float2* b, const FftFactors& f, int ld, int ncols,
                                              const float2* __restrict__ tw) {
    float2* in = a;
    float2* out = b;
    int Ns = 1;
    for (int s = 0; s < f.nstages; ++s) {
        const int r = f.radix[s];
        switch (r) {
            case 2: stage_radix<2>(in, out, f.n, Ns, ld, ncols, tw); break;
            case 3: stage_radix<3>(in, out, f.n, Ns, ld, ncols, tw); break;
            case 4: stage_radix<4>(in, out, f.n, Ns, ld, ncols, tw); break;
            case 5: stage_radix<5>(in, out, f.n, Ns, ld, ncols, tw); break;
            case 8: stage_radix<8>(in, out, f.n, Ns, ld, ncols, tw); break;
            default: stage_generic(in, out, f.n, Ns, r, ld, ncols, tw); break;
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
