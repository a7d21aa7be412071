// Generic two-pass path ("g2"): the layout of the fused kernels (ps21_fused.cu) for any box whose axes factor into
// the radices of the in-place engine below -- 200^3 and 512^3 coeval cubes (BASELINE configs 2 and 4), lightcone chunks
// of any even length with factors up to 13 (250, 300, 182, ...), sky planes 64 ... 512.
//
// HBM traffic = 4N (read the box) + 8N_h (write the half spectrum) + 8N_h (read it) = the B_alg of SURVEY 8d; the
// three-pass general path (ps21_general.cu) moves 4N + 32N_h.
//
//   y = y_lo + R2 j,  k_y = ky_a + R1 ky_b   (n1 = R1 R2, R2 = 4 or 8, T = 32 / R2)
//
//   pass 1  k_g2_p1   CTA = x slabs x R2 units (y_lo) of R1 z-rows each.  Rows land in shared memory as packed
//                     (even, odd) complex pairs [j][z/2] (odd row stride); z: complex FFT of length nz/2 + untangle
//                     (real-FFT post-processing); y: FFT of length R1 over j; times w_n1^(y_lo ky_a); stored to HBM in
//                     tile order: tile = (ky_a position, T consecutive k_z rows), inner [x][y_lo][T] -> runs of T
//                     coefficients (64 bytes for R2 = 4).
//   pass 2  k_g2_p2   CTA loops over tiles: one bulk copy (TMA) per tile and one for its bin indices; first x stage
//                     fused with the R2-point FFT over y_lo (<= 32 values per thread), remaining x stages in place,
//                     the last one straight into |delta_k|^2 and the lane-private (<= 40 bins) or 4-lane-group
//                     (<= 254 bins) fp32 bins, flushed in a fixed order into fp64 (bitwise reproducible, no atomics).
//
// All transforms are decimation-in-frequency and IN PLACE (a thread writes the locations it read: one buffer, one
// barrier per stage).  Their outputs are therefore in digit-reversed order along every axis; nothing ever puts them
// back in natural order: the tile grid is defined on *positions*, and the bin-index table (one byte per mode, built
// once per plan in exactly the order pass 2 visits its outputs) knows which Fourier mode sits where.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <vector>

#include "ps21_internal.cuh"
#include "ptx_sm100.cuh"
#include "../../include/ps21.h"

namespace ps21 {
namespace g2 {

constexpr int kMaxStages = 8;
constexpr int MAX_BINS_LANE = 40;     // lane-private fp32 bin columns
constexpr int MAX_BINS = 254;         // byte LUT (255 = discarded mode)
constexpr int FLUSH_EVERY = 4;        // tiles between two fp32 -> fp64 flushes

struct IpStage {
    int radix, q, twmul;   // q = butterflies per sub-block (L / radix); twiddle of output m' at offset o: tw[o m' twmul]
    int nbf;               // butterflies of the stage per column (n / radix): from the host, a division per stage and unit
                           // and thread was 2.5 % of the pass-1 instructions (ncu)
    FastDiv div_q;
};
struct IpFft {
    int n, nstages;
    IpStage st[kMaxStages];
};

struct G2Dev {
    int n0, n1, nz, M, nh;      // M = length of the complex z transform, nh = nz / 2 + 1 half-spectrum rows
    int pair, zcols;            // even nz: M = nz / 2 on (even, odd)-packed rows, zcols = R1 columns (pair = 0);
                                // odd nz:  M = nz on row PAIRS (row j + i row j + R1/2), zcols = R1 / 2 columns (pair = 1)
    int R1, R2, T, lgT;         // n1 = R1 R2, T = 32 / R2 k_z rows per tile
    int ld;                     // pass 1: complex slots per shared-memory row [j][z slot] (odd -> conflict free both ways)
    int nzt, ntiles;            // z tiles per ky_a position, tiles per spectrum
    int slots;                  // pass-2 CTAs per spectrum
    int p2_threads;
    int nbuf;                   // pass 1: shared-memory row buffers (2: the next unit streams in during this one's FFTs)
    int slabs;                  // pass 1: consecutive x slabs per CTA
    int fuse_first, fuse_last;  // pass 1: untangle inside the first y stage / HBM store inside the last y stage
    int parts;                  // pass 1, units alone on an SM: pieces the last y stage is cut into for the row reload
    int lgMpad;                 // pass 1: log2 of M rounded up to a multiple of 32 when that is a power of two, else 0
    int mapb, lgR2, swshift;    // pass 2, last stage: lanes = (k_z, R2 consecutive x blocks) instead of (k_z, ky_b) so that
                                // the lanes of a warp meet their bin changes together; swshift: the ky_b slot of row x is
                                // XOR-ed with (x ^ (x >> swshift)) & (R2 - 1) (31: plain x & (R2 - 1))
    int dbg_skip;               // experiments only (PS21_G2_DBG_SKIP): bit 0 skip the z stages, bit 1 the first y stage
                                // (timing, results are garbage); bit 4: no two-half reload of the row buffer (correct)
    int bluestein, Lc;          // pass 1: z transform of length M as a chirp convolution of length Lc (M has a big prime)
    int zrows;                  // pass 1: rows per buffer = max(nh, Lc)
    IpFft fzb;                  // ... the convolution's transform of length Lc
    const float2 *tw_c, *chirp, *bhat;   // w_Lc^q [Lc]; exp(-i pi n^2 / M) [M]; FFT of the chirp / Lc, position order [Lc]
    IpFft fz, fy, fx;
    FastDiv div_M, div_R1, div_nh, div_zcols;
    const float2 *tw_z, *tw_y, *tw_x;   // exp(-2 pi i q / n) for n = nz, n1, n0
    const uint16_t* rowz_of_k;          // [M]   row that holds frequency k after the in-place z transform
    const uint16_t* kya_of_jp;          // [R1]  ky_a at position jp after the in-place y transform
    const uint16_t* prow_of_row;        // [nh]  pair mode: physical row of logical half-spectrum row r
    const uint8_t* zmult;               // [nzt T] Hermitian multiplicity of the k_z in row r (0: padding row)
};

struct G2Shape {
    G2Dev dev{};
    uint16_t *d_rowz = nullptr, *d_kya = nullptr, *d_kxpos = nullptr, *d_kzrow = nullptr, *d_prow = nullptr;
    uint8_t* d_zmult = nullptr;
    float2 *d_twc = nullptr, *d_chirp = nullptr, *d_bhat = nullptr;   // Bluestein tables
    int R0 = 0;                 // radix of the first x stage (fused with the y_lo transform)
    int nshare = 1;             // lanes per fp32 bin column: 1 (<= 40 bins) or 4
    size_t smem1 = 0, smem2 = 0;
    int p1_threads = 256;
};

// ------------------------------------------------------------------------------------------------
// in-place decimation-in-frequency column engine
// ------------------------------------------------------------------------------------------------
// One stage of radix R over `ncols` independent sequences: element i of column c lives at buf[addr(i, c)].
// Stage geometry: sub-blocks of L = q R consecutive indices; butterfly (b, o) takes i = b L + o + m q, m = 0..R-1,
// and leaves output m' (multiplied by w_L^(o m')) where input m' was.
// `pre(i, v)` transforms a value as it is read (index i along the transform axis):
//   PreSub    pass 1, first z stage: the chunk's offset estimate is subtracted from both halves of every packed pair
//             before any butterfly, so the fp32 transform never sees the large mean;
//   PreChirp  the same, times the chirp exp(-i pi i^2 / M), zero beyond M (Bluestein input);
//   PreBhat   times the chirp's spectrum, conjugated (Bluestein: first stage of the inverse transform).
struct PreSub {
    float2 s;
    __device__ __forceinline__ float2 operator()(int, float2 v) const { return csub(v, s); }
};
struct PreChirp {
    float2 s;
    const float2* chirp;
    int M;
    __device__ __forceinline__ float2 operator()(int i, float2 v) const {
        return i < M ? cmul_w2(csub(v, s), chirp[i]) : make_float2(0.f, 0.f);
    }
};
struct PreBhat {
    const float2* bhat;
    __device__ __forceinline__ float2 operator()(int i, float2 v) const {
        const float2 t = cmul_w2(v, bhat[i]);
        return make_float2(t.x, -t.y);
    }
};

// `addr.at(addr.col(c), i)` is the 32-bit shared-memory byte address of element i of column c.
template <int R, class Addr, class Pre>
__device__ __forceinline__ void ip_stage(const IpStage st, int nbf, int ncols, FastDiv div_cols,
                                         const Addr addr, const float2* __restrict__ tw, const Pre pre) {
    const int ntasks = nbf * ncols;
    for (int task = threadIdx.x; task < ntasks; task += blockDim.x) {
        const int bf = fast_div(task, div_cols), c = task - bf * ncols;
        const int b = fast_div(bf, st.div_q), o = bf - b * st.q;
        const int i0 = b * st.q * R + o;
        float2 v[R];
        const uint32_t cb = addr.col(c);
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = pre(i0 + m * st.q, lds64(addr.at(cb, i0 + m * st.q)));
        small_dft<R>(v);
        if (st.q > 1) {
            const int step = o * st.twmul;
            int idx = step;
#pragma unroll
            for (int m = 1; m < R; ++m) {
                v[m] = cmul_w2(v[m], tw[idx]);
                idx += step;
            }
        }
#pragma unroll
        for (int m = 0; m < R; ++m) sts64(addr.at(cb, i0 + m * st.q), v[m]);
    }
}

// The transposed stage (twiddle on the inputs, then the butterfly).  Running the stages of a forward plan transposed
// and in reverse order takes digit-reversed input to natural-order output: the DFT matrix is symmetric, so
// W = (P^T S_last ... S_0)^T = S_0^T ... S_last^T P.  Used by the Bluestein path for the inverse transform
// (on conjugated data).
template <int R, class Addr, class Pre>
__device__ __forceinline__ void ip_stage_T(const IpStage st, int nbf, int ncols, FastDiv div_cols,
                                           const Addr addr, const float2* __restrict__ tw, const Pre pre) {
    const int ntasks = nbf * ncols;
    for (int task = threadIdx.x; task < ntasks; task += blockDim.x) {
        const int bf = fast_div(task, div_cols), c = task - bf * ncols;
        const int b = fast_div(bf, st.div_q), o = bf - b * st.q;
        const int i0 = b * st.q * R + o;
        float2 v[R];
        const uint32_t cb = addr.col(c);
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = pre(i0 + m * st.q, lds64(addr.at(cb, i0 + m * st.q)));
        if (st.q > 1) {
            const int step = o * st.twmul;
            int idx = step;
#pragma unroll
            for (int m = 1; m < R; ++m) {
                v[m] = cmul_w2(v[m], tw[idx]);
                idx += step;
            }
        }
        small_dft<R>(v);
#pragma unroll
        for (int m = 0; m < R; ++m) sts64(addr.at(cb, i0 + m * st.q), v[m]);
    }
}

template <bool BIG, bool TRANSPOSED = false, class Addr, class Pre>
__device__ __forceinline__ void ip_stage_any(const IpStage st, int n, int ncols, FastDiv div_cols,
                                             const Addr addr, const float2* tw, const Pre pre) {
    const int nbf = st.nbf;
    (void)n;
#define PS21_IP_CASE(R)                                                        \
    case R:                                                                    \
        if (TRANSPOSED) ip_stage_T<R>(st, nbf, ncols, div_cols, addr, tw, pre);      \
        else ip_stage<R>(st, nbf, ncols, div_cols, addr, tw, pre);             \
        break
    switch (st.radix) {
        PS21_IP_CASE(2); PS21_IP_CASE(3); PS21_IP_CASE(4); PS21_IP_CASE(5); PS21_IP_CASE(6); PS21_IP_CASE(7);
        PS21_IP_CASE(8); PS21_IP_CASE(9); PS21_IP_CASE(10); PS21_IP_CASE(11); PS21_IP_CASE(12); PS21_IP_CASE(13);
        PS21_IP_CASE(15); PS21_IP_CASE(16);
        case 25: if (BIG) ip_stage<25>(st, nbf, ncols, div_cols, addr, tw, pre); break;
        case 32: if (BIG) ip_stage<32>(st, nbf, ncols, div_cols, addr, tw, pre); break;
        default: break;
    }
#undef PS21_IP_CASE
}

struct AddrStrided {   // pass 1, z transform: byte strides of the index (z pair / k_z: 8) and of the column (j: one row)
    uint32_t base, si, sc;
    __device__ __forceinline__ uint32_t col(int c) const { return base + c * sc; }
    __device__ __forceinline__ uint32_t at(uint32_t cb, int i) const { return cb + i * si; }
};
struct AddrRows {      // pass 1, y transform: index = j (one buffer row each), column = logical k_z row -> slot prow[c]
    uint32_t base, ldb;
    const uint16_t* prow;   // nullptr: the k_z rows are stored in logical order
    __device__ __forceinline__ uint32_t col(int c) const { return base + (prow ? (uint32_t)prow[c] : (uint32_t)c) * 8u; }
    __device__ __forceinline__ uint32_t at(uint32_t cb, int i) const { return cb + i * ldb; }
};
struct AddrTile {      // pass 2: row x holds 32 columns (ky_b, t); the ky_b slot is XOR-ed with the low bits of x
    uint32_t base;     // 256-byte aligned
    int swmask, sh;    // swmask = R2 - 1, sh = log2(T) + 3
    int swshift;       // G2Dev::swshift
    __device__ __forceinline__ uint32_t col(int c) const { return (uint32_t)(c << 3); }
    __device__ __forceinline__ uint32_t at(uint32_t cb, int i) const {
        return base + i * 256 + (cb ^ (((i ^ (i >> swshift)) & swmask) << sh));
    }
};
// slot swizzle of row x (see G2Dev::swshift)
__host__ __device__ __forceinline__ int sw_of_x(int x, int swshift, int R2) { return (x ^ (x >> swshift)) & (R2 - 1); }
// last x stage of pass 2: task -> (column c = (ky_b, t), butterfly bf).  mapb: lane = (t, low bits of bf), warp = (ky_b,
// high bits of bf): the lanes of a warp then differ by a few k_z and a few k_x only, so |k| and with it the bin changes
// along the walk nearly coincide (with ky_b in the lane index they differ by n1 / R2 rows of k_y and every warp took the
// slow path of the walk at almost every word: ncu, 27 % of all pass-2 instructions).
__host__ __device__ __forceinline__ void last_task_map(int task, int mapb, int lgT, int lgR2, int& c, int& bf) {
    if (!mapb) {
        c = task & 31;
        bf = task >> 5;
    } else {
        const int t = task & ((1 << lgT) - 1), bfl = (task >> lgT) & ((1 << lgR2) - 1), w = task >> 5;
        c = ((w & ((1 << lgR2) - 1)) << lgT) | t;
        bf = ((w >> lgR2) << lgR2) | bfl;
    }
}

// ------------------------------------------------------------------------------------------------
// LUT (one byte per mode, in the order pass 2 visits its outputs) + counts
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int visit_order(int v, int R) { return (v & 1) ? R - 1 - (v >> 1) : (v >> 1); }
// bin indices of one butterfly's RL outputs occupy RL rounded up to whole 32-bit words (25 -> 28): the walk reads words
__host__ __device__ __forceinline__ int lut_row(int RL) { return (RL + 3) & ~3; }
__host__ __device__ __forceinline__ long long lut_tile_bytes(int n0, int RL) { return (long long)(n0 / RL) * 32 * lut_row(RL); }

__global__ void k_g2_lut(uint8_t* __restrict__ lut, unsigned long long* __restrict__ counts, G2Dev g, int RL,
                         const uint16_t* __restrict__ kx_of_pos, const uint16_t* __restrict__ kz_of_row,
                         const double* __restrict__ k0sq, const double* __restrict__ k1sq,
                         const double* __restrict__ k2sq, const double* __restrict__ edges, int nbins, int ig_kperp,
                         int ig_kpar, int ig_k0) {
    const int RLP = lut_row(RL);
    const long long tile_bytes = lut_tile_bytes(g.n0, RL);
    const long long total = (long long)g.ntiles * tile_bytes;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int tile = (int)(i / tile_bytes), e = (int)(i % tile_bytes);
        // e = (bf * 32 + c) * RLP + v : butterfly bf of the last x stage, column c = (ky_b, t), output in visiting order
        const int v = e % RLP, task = e / RLP;
        int c, bf;
        last_task_map(task, g.mapb, g.lgT, g.lgR2, c, bf);
        if (v >= RL) {   // padding of the row
            lut[i] = 255;
            continue;
        }
        const int kx = kx_of_pos[bf * RL + visit_order(v, RL)];
        const int kyb = c >> g.lgT, t = c & (g.T - 1);
        const int jp = tile / g.nzt, zt = tile - jp * g.nzt;
        const int ky = g.kya_of_jp[jp] + g.R1 * kyb;
        const int r = zt * g.T + t;
        uint8_t out = 255;
        if (r < g.nh) {
            const int kz = kz_of_row[r];
            int w = (kz == 0 || 2 * kz == g.nz) ? 1 : 2;
            const bool perp0 = (kx == 0 && ky == 0);
            if (ig_kperp && perp0) w = 0;
            if (ig_kpar && kz == 0) w = 0;
            if (ig_k0 && perp0 && kz == 0) w = 0;
            const int b = digitize_mode(k0sq[kx], k1sq[ky], k2sq[kz], edges, nbins);
            if (b >= 0 && w > 0) {
                out = (uint8_t)b;
                atomicAdd(&counts[b], (unsigned long long)w);
            }
        }
        lut[i] = out;
    }
}

// ------------------------------------------------------------------------------------------------
// pass 1
// ------------------------------------------------------------------------------------------------
// NOTE: every shared-memory pointer below is formed as `smem + integer` inside the kernel and handed down by value.
// (A struct holding a dynamically indexed pointer array cost the address-space inference: ncu showed LD.E / ST.E
// generic accesses, 30 % excess wavefronts and long-scoreboard stalls where LDS / STS were meant.)

// untangle of the packed real FFT for pair p of one butterfly leg: rows r1 = row(p), r2 = row(M - p)
// BLU: the rows hold the conjugate of the chirp convolution; C_p = chirp[p] * conj(row p) is formed as they are read
template <bool BLU>
__device__ __forceinline__ float2 zfix(float2 v, const float2* __restrict__ chirp, int p) {
    return BLU ? cmul_w2(make_float2(v.x, -v.y), chirp[p]) : v;
}

template <int R, bool BLU>
__device__ __forceinline__ void p1_y_first(uint32_t buf, const G2Dev& g, const float2* __restrict__ twz,
                                           const float2* __restrict__ twf, const uint16_t* __restrict__ rowz,
                                           const float2* __restrict__ chirp) {
    // first y stage (radix R over j = o + m q) of the row pair (p, M - p), with the untangle of that pair done on
    // the values as they are read: one pass over shared memory and one barrier instead of two
    const IpStage st = g.fy.st[0];
    const int q = st.q, M = g.M;
    const uint32_t ldb = g.ld * 8, qb = q * ldb;   // element (k_z slot r, row j) lives at buf + j ldb + 8 r
    const int ntasks = (M / 2 + 1) * q;
    for (int task = threadIdx.x; task < ntasks; task += blockDim.x) {
        const int p = fast_div(task, st.div_q), o = task - p * q;   // o fastest: lanes along j (odd row stride)
        float2 a1[R], a2[R];
        uint32_t q1, q2;
        const bool self = (2 * p == M);
        if (p == 0) {
            q1 = buf + o * ldb;               // frequency 0 stays in slot 0
            q2 = buf + o * ldb + M * 8;       // the Nyquist slot M is created here
#pragma unroll
            for (int m = 0; m < R; ++m) {
                const float2 c0 = zfix<BLU>(lds64(q1 + m * qb), chirp, 0);
                a1[m] = make_float2(c0.x + c0.y, 0.f);
                a2[m] = make_float2(c0.x - c0.y, 0.f);
            }
        } else if (self) {
            q1 = q2 = buf + o * ldb + rowz[p] * 8;
#pragma unroll
            for (int m = 0; m < R; ++m) {
                const float2 c0 = zfix<BLU>(lds64(q1 + m * qb), chirp, p);
                a1[m] = make_float2(c0.x, -c0.y);
                a2[m] = a1[m];
            }
        } else {
            q1 = buf + o * ldb + rowz[p] * 8;
            q2 = buf + o * ldb + rowz[M - p] * 8;
            const float2 w = twz[p];   // w_nz^p; times -i/2 -> (wy/2, -wx/2)
            const float4 wu = make_float4(0.5f * w.y, -0.5f * w.x, 0.5f * w.x, 0.5f * w.y);
#pragma unroll
            for (int m = 0; m < R; ++m)
                untangle(zfix<BLU>(lds64(q1 + m * qb), chirp, p), zfix<BLU>(lds64(q2 + m * qb), chirp, M - p), wu,
                         a1[m], a2[m]);
        }
        small_dft<R>(a1);
        small_dft<R>(a2);
        if (q > 1) {
#pragma unroll
            for (int m = 1; m < R; ++m) {
                const float2 w = twf[(m - 1) * q + o];   // w_n1^(o m twmul), lanes along o: contiguous
                a1[m] = cmul_w2(a1[m], w);
                a2[m] = cmul_w2(a2[m], w);
            }
        }
#pragma unroll
        for (int m = 0; m < R; ++m) sts64(q1 + m * qb, a1[m]);
        if (!self) {
#pragma unroll
            for (int m = 0; m < R; ++m) sts64(q2 + m * qb, a2[m]);
        }
    }
}

// last y stage (radix R over R consecutive buffer rows) for every k_z slot, results times w_n1^(yl ky_a) straight to
// HBM in tile order; sub-blocks b0 .. b1-1 (the caller runs the stage in two halves and reloads the buffer rows of the
// first half while the second is being stored)
template <int R>
__device__ __forceinline__ void p1_y_last_store(uint32_t buf, float2* __restrict__ dst,
                                                const G2Dev& g, const float2* __restrict__ twu,
                                                const uint16_t* __restrict__ prow, int yl, long long inner,
                                                int b0, int b1) {
    const int nh = g.nh;
    const uint32_t ldb = g.ld * 8;
    const long long tile_elems = (long long)g.n0 * 32;
    for (int task = b0 * nh + threadIdx.x; task < b1 * nh; task += blockDim.x) {
        const int b = fast_div(task, g.div_nh), r = task - b * nh;   // r fastest: lanes along k_z slots
        float2 v[R];
        const uint32_t src = buf + (prow ? (uint32_t)prow[r] : (uint32_t)r) * 8u + (uint32_t)(b * R) * ldb;
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = lds64(src + m * ldb);
        small_dft<R>(v);
        const long long jstride = (long long)g.nzt * tile_elems;
        float2* out = dst + (long long)(r >> g.lgT) * tile_elems + inner + (r & (g.T - 1)) + (long long)(b * R) * jstride;
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const float2 o = yl ? cmul_w2(v[m], twu[b * R + m]) : v[m];
            *out = o;
            out += jstride;
        }
    }
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_g2_p1(const float* __restrict__ boxes, long long batch_stride, int n2_total, ItemMap im,
                        const ChunkDev* __restrict__ chunks, int nchunks, const double* __restrict__ mean_parts,
                        float2* __restrict__ scratch, G2Dev g, int aligned8) {
    PS21_DYN_SMEM(float2, smem);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int M = g.M, R1 = g.R1, ld = g.ld, nh = g.nh;
    const int bufstride = R1 * ld;
    const uint32_t smem32 = smem_u32(smem), ldb = (uint32_t)ld * 8u;
    float2* const twz = smem + g.nbuf * bufstride;                       // [nz]  w_nz^q
    float2* const twy = twz + g.nz;                                      // [n1]  w_n1^q
    float2* const twc = twy + g.n1;                                      // [Lc]  w_Lc^q            (Bluestein only)
    float2* const chirp = twc + g.Lc;                                    // [M]   exp(-i pi n^2/M)  (Bluestein only)
    float2* const bhat = chirp + (g.bluestein ? M : 0);                  // [Lc]  chirp spectrum    (Bluestein only)
    float2* const twu = bhat + g.Lc;                                     // [R1]  w_n1^(yl ky_a) of the current unit
    float2* const twf = twu + R1;                                        // [(R-1) q] twiddles of the first y stage by (m, o)
    uint16_t* const rowz = reinterpret_cast<uint16_t*>(twf + R1);        // [M]   copy of G2Dev::rowz_of_k
    uint16_t* const kya = rowz + M;                                      // [R1]  copy of G2Dev::kya_of_jp
    uint16_t* const prow = kya + R1;                                     // [nh]  copy of G2Dev::prow_of_row (pair mode)
    long long box;
    int cid;
    item_to_box_chunk(im, blockIdx.y, box, cid);
    const ChunkDev ch = chunks[cid];
    const float mean = chunk_offset(mean_parts, box * nchunks + cid, g.n1, ch.nz);
    const int x0 = blockIdx.x * g.slabs;
    const float* spec = boxes + box * batch_stride + ch.start;
    float2* dst = scratch + (long long)blockIdx.y * ((long long)g.ntiles * g.n0 * 32);

    for (int i = tid; i < g.nz; i += nthr) twz[i] = g.tw_z[i];
    for (int i = tid; i < g.n1; i += nthr) twy[i] = g.tw_y[i];
    {   // first y stage: w_n1^(o m twmul) stored [m - 1][o]; its lanes run along o, and read straight from the w_n1^q table
        // (stride m twmul) the R - 1 loads were 3- to 10-way bank conflicts: 15 % of all pass-1 wavefronts on 200^3 (ncu)
        const IpStage st0 = g.fy.st[0];
        for (int i = tid; i < (st0.radix - 1) * st0.q; i += nthr) {
            const int m = i / st0.q + 1, o = i - (m - 1) * st0.q;
            twf[i] = g.tw_y[o * m * st0.twmul];
        }
    }
    for (int i = tid; i < M; i += nthr) rowz[i] = g.rowz_of_k[i];
    for (int i = tid; i < R1; i += nthr) kya[i] = g.kya_of_jp[i];
    if (g.pair)
        for (int i = tid; i < nh; i += nthr) prow[i] = g.prow_of_row[i];
    if (g.bluestein) {
        for (int i = tid; i < g.Lc; i += nthr) {
            twc[i] = g.tw_c[i];
            bhat[i] = g.bhat[i];
        }
        for (int i = tid; i < M; i += nthr) chirp[i] = g.chirp[i];
    }
    // rows y = yl + R2 j of slab x as packed (even, odd) pairs: buf[j][zc]; asynchronous copies (LDGSTS), lanes along z:
    // coalesced reads, contiguous writes.  `load_rows` is called by whole warps: rows j0 .. j1-1 of unit u, one row
    // per warp turn starting at `w0` of `nw` warps.
    auto load_rows = [&](int u, uint32_t buf32, int j0, int j1, int w0, int nw) {
        const int x = x0 + (u >> g.lgR2), yl = u & (g.R2 - 1);
        const float* slab = spec + (long long)x * g.n1 * n2_total + (long long)yl * n2_total;
        const int lane = tid & 31;
        const long long rstride = (long long)g.R2 * n2_total;
        if (!g.pair && aligned8 == 2 && g.lgMpad > 0 && (1 << g.lgMpad) <= nthr) {
            // a row padded to Mpad = 2^lgMpad pair slots: thread -> (row jr of the turn, slot zc); zc is fixed for a thread,
            // only the row advances -- one LDGSTS + two pointer steps per copy, no per-row set-up (the warp-per-row
            // loop below was still 15 % of the pass-1 instructions on 200^3 after its first trimming).  Only for rows that
            // start on 32-byte sectors (aligned8 == 2): the 256-byte pieces of a row are then fetched by different warps,
            // and on 8-byte aligned rows their shared boundary sectors cost more than the instructions save
            // (250-slice chunks: pass 1 +7 %; 200^3: -2.3 %)
            const int zc = tid & ((1 << g.lgMpad) - 1), jr = tid >> g.lgMpad, rpt = nthr >> g.lgMpad;
            if (zc < M) {
                const float* src = slab + (long long)(j0 + jr) * rstride + 2 * zc;
                uint32_t d = buf32 + (uint32_t)(j0 + jr) * ldb + (uint32_t)zc * 8u;
                const long long sstep = (long long)rpt * rstride;
                const uint32_t dstep = (uint32_t)rpt * ldb;
#pragma unroll 4
                for (int j = j0 + jr; j < j1; j += rpt, src += sstep, d += dstep) cp_async8_s32(d, src);
            }
        } else if (!g.pair) {
            for (int j = j0 + w0; j < j1; j += nw) {
                const float* src = slab + j * rstride + 2 * lane;
                uint32_t d = buf32 + (uint32_t)j * ldb + (uint32_t)lane * 8u;
                if (aligned8) {
                    // M / 32 whole warp turns (uniform trip count: no divergence bookkeeping, unrollable) + one partial
                    // (ncu: the per-lane `zc < M` loop was 9 % of the pass-1 instructions on 200^3)
#pragma unroll 1
                    for (int i = 0; i < (M >> 5); ++i, d += 256u, src += 64) cp_async8_s32(d, src);
                    if (lane < (M & 31)) cp_async8_s32(d, src);
                } else {
                    for (int zc = lane; zc < M; zc += 32, d += 256u, src += 64) {
                        cp_async4_s32(d, src);
                        cp_async4_s32(d + 4, src + 1);
                    }
                }
            }
        } else {   // odd nz: column c = row j (real part) + i row j + R1/2 (imaginary part), one complex value per slice
            for (int j = j0 + w0; j < j1; j += nw) {
                const int c = j < g.zcols ? j : j - g.zcols;
                const float* src = slab + j * rstride + lane;
                uint32_t d = buf32 + (uint32_t)c * ldb + (j < g.zcols ? 0u : 4u) + (uint32_t)lane * 8u;
                for (int z = lane; z < M; z += 32, d += 256u, src += 32) cp_async4_s32(d, src);
            }
        }
        cp_async_commit();
    };
    auto issue_unit = [&](int u, uint32_t buf32) { load_rows(u, buf32, 0, R1, tid >> 5, nthr >> 5); };
    const int nunits = g.slabs * g.R2;
    issue_unit(0, smem32);
    for (int u = 0; u < nunits; ++u) {
        float2* buf = smem + (u & (g.nbuf - 1)) * bufstride;   // nbuf is 1 or 2
        const uint32_t buf32 = smem32 + (uint32_t)(u & (g.nbuf - 1)) * (uint32_t)bufstride * 8u;
        const AddrStrided az{buf32, 8u, ldb};    // z transform: index = z pair / k_z slot, column = j
        const AddrRows ay{buf32, ldb, g.pair ? prow : nullptr};   // y transform: index = j, column = row
        const int x = x0 + (u >> g.lgR2), yl = u & (g.R2 - 1);
        cp_async_wait<0>();
        __syncthreads();   // unit u has landed; everybody is done with unit u - 1 (tables ready on the first turn)
        if (g.nbuf == 2 && u + 1 < nunits) issue_unit(u + 1, smem32 + (uint32_t)((u + 1) % 2) * (uint32_t)bufstride * 8u);
        // L2 prefetch of the rows whose copies are issued next (the next unit of this CTA, or the first unit of the CTA
        // that will take its place): by then they come from L2, not DRAM.  Issued `pf_where` stages before the loads: too
        // early and the scratch stream has flushed them out of L2 again (ncu: +37 % DRAM reads when issued a whole
        // 512^3 unit = 37 us ahead).
        auto prefetch_next = [&]() {
            const int un = u + g.nbuf;
            const float* nslab = nullptr;
            if (un < nunits) {
                const int xn = x0 + (un >> g.lgR2), yn = un & (g.R2 - 1);
                nslab = spec + (long long)xn * g.n1 * n2_total + (long long)yn * n2_total;
            } else if ((im.pf & 2) && un == nunits) {
                const long long n = (long long)blockIdx.y * gridDim.x + blockIdx.x + im.pf_ahead;
                const long long it2 = n / gridDim.x;
                if (it2 < (long long)gridDim.y) {
                    long long b2;
                    int c2;
                    item_to_box_chunk(im, it2, b2, c2);
                    nslab = boxes + b2 * batch_stride + chunks[c2].start +
                            (long long)((int)(n % gridDim.x) * g.slabs) * g.n1 * n2_total;
                }
            }
            if (nslab)
                for (int j = tid; j < R1; j += nthr)
                    bulk_prefetch_l2(nslab + (long long)j * g.R2 * n2_total, (uint32_t)g.nz * 4u);
        };
        if (im.pf && im.pf_where == 0) prefetch_next();
        for (int jp = tid; jp < R1; jp += nthr) twu[jp] = twy[yl * kya[jp]];   // yl ky_a < n1: no reduction needed
        // ---- z: complex FFT of length M over the row index, in place; the offset is subtracted as stage 0 reads
        const float2 mean2 = make_float2(mean, mean);
        if (g.dbg_skip & 1) {   // timing experiment: rows in, tiles out (profiles/r02_g2_experiments.md)
        } else if (!g.bluestein) {
            for (int s = 0; s < g.fz.nstages; ++s) {
                ip_stage_any<false>(g.fz.st[s], M, g.zcols, g.div_zcols, az, twz, PreSub{s == 0 ? mean2 : make_float2(0.f, 0.f)});
                __syncthreads();
            }
        } else {
            // M has a prime factor the butterflies do not cover: X_k = c_k sum_n (x_n c_n) conj(c)_{k-n}, c_n = exp(-i pi n^2/M),
            // as a circular convolution of length Lc >= 2M - 1: forward transform (chirp and zero padding applied as
            // stage 0 reads), times the chirp's spectrum (applied, conjugated, as the inverse's first stage reads),
            // inverse transform = the forward plan transposed and reversed on conjugated data -> natural order;
            // the last factor c_k and the conjugation are applied where the untangle reads the rows.
            const int ns = g.fzb.nstages;
            for (int s = 0; s < ns; ++s) {
                if (s == 0) ip_stage_any<false>(g.fzb.st[0], g.Lc, g.zcols, g.div_zcols, az, twc, PreChirp{mean2, chirp, M});
                else ip_stage_any<false>(g.fzb.st[s], g.Lc, g.zcols, g.div_zcols, az, twc, PreSub{make_float2(0.f, 0.f)});
                __syncthreads();
            }
            for (int s = ns - 1; s >= 0; --s) {
                if (s == ns - 1) ip_stage_any<false, true>(g.fzb.st[s], g.Lc, g.zcols, g.div_zcols, az, twc, PreBhat{bhat});
                else ip_stage_any<false, true>(g.fzb.st[s], g.Lc, g.zcols, g.div_zcols, az, twc, PreSub{make_float2(0.f, 0.f)});
                __syncthreads();
            }
        }
        if (im.pf && im.pf_where == 1) prefetch_next();
        // ---- untangle (pairs (p, M - p) -> A_p, A_{M-p}; p = 0 also yields the Nyquist row M) + first y stage
        if (g.dbg_skip & 2) {
        } else if (g.fuse_first) {
#define PS21_G2_YFIRST(R)                                                        \
    do {                                                                         \
        if (g.bluestein) p1_y_first<R, true>(buf32, g, twz, twf, rowz, chirp);   \
        else p1_y_first<R, false>(buf32, g, twz, twf, rowz, chirp);              \
    } while (0)
            switch (g.fy.st[0].radix) {
                case 2: PS21_G2_YFIRST(2); break;
                case 3: PS21_G2_YFIRST(3); break;
                case 4: PS21_G2_YFIRST(4); break;
                case 5: PS21_G2_YFIRST(5); break;
                case 6: PS21_G2_YFIRST(6); break;
                case 8: PS21_G2_YFIRST(8); break;
                default: PS21_G2_YFIRST(10); break;
            }
#undef PS21_G2_YFIRST
            __syncthreads();
        } else if (g.pair) {
            // odd nz: Z = FFT(row_a + i row_b);  A_a(k) = (Z_k + conj Z_{nz-k}) / 2,  A_b(k) = (Z_k - conj Z_{nz-k}) / 2i.
            // A_a replaces Z_k, A_b goes to the same slot of buffer row c + R1/2 (unused so far); the slots of Z_{nz-k} are dead after.
            for (int e = tid; e < nh * g.zcols; e += nthr) {
                const int k = fast_div(e, g.div_zcols), c = e - k * g.zcols;
                auto fix = [&](float2 v, int row) {   // Bluestein: Z_k = chirp[k] conj(row k)
                    return g.bluestein ? cmul_w2(make_float2(v.x, -v.y), chirp[row]) : v;
                };
                float2* qk = buf + c * ld + rowz[k];
                const float2 zk = fix(*qk, k);
                const float2 zn = (k == 0) ? zk : fix(buf[c * ld + rowz[M - k]], M - k);
                qk[0] = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
                qk[g.zcols * ld] = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
            }
            __syncthreads();
            ip_stage_any<false>(g.fy.st[0], R1, nh, g.div_nh, ay, twy, PreSub{make_float2(0.f, 0.f)});
            __syncthreads();
        } else {
            const int npair = M / 2 + 1;
            for (int e = tid; e < npair * R1; e += nthr) {
                const int p = fast_div(e, g.div_R1), j = e - p * R1;
                auto fix = [&](float2 v, int row) {   // Bluestein: C_p = chirp[p] conj(row p)
                    return g.bluestein ? cmul_w2(make_float2(v.x, -v.y), chirp[row]) : v;
                };
                if (p == 0) {
                    const float2 c0 = fix(buf[j * ld], 0);
                    buf[j * ld] = make_float2(c0.x + c0.y, 0.f);
                    buf[j * ld + M] = make_float2(c0.x - c0.y, 0.f);
                } else if (2 * p == M) {
                    float2* q = buf + j * ld + rowz[p];
                    const float2 c0 = fix(*q, p);
                    *q = make_float2(c0.x, -c0.y);
                } else {
                    float2* q1 = buf + j * ld + rowz[p];
                    float2* q2 = buf + j * ld + rowz[M - p];
                    const float2 w = twz[p];
                    float2 a1, a2;
                    untangle(fix(*q1, p), fix(*q2, M - p), make_float4(0.5f * w.y, -0.5f * w.x, 0.5f * w.x, 0.5f * w.y), a1, a2);
                    *q1 = a1;
                    *q2 = a2;
                }
            }
            __syncthreads();
            ip_stage_any<false>(g.fy.st[0], R1, nh, g.div_nh, ay, twy, PreSub{make_float2(0.f, 0.f)});
            __syncthreads();
        }
        // ---- middle y stages, in place
        const int ylast = g.fy.nstages - 1;
        for (int s = 1; s < (g.fuse_last ? ylast : g.fy.nstages); ++s) {
            ip_stage_any<false>(g.fy.st[s], R1, nh, g.div_nh, ay, twy, PreSub{make_float2(0.f, 0.f)});
            __syncthreads();
        }
        if (im.pf && im.pf_where == 2) prefetch_next();
        // ---- last y stage, times w_n1^(yl ky_a), to HBM in tile order: runs of T consecutive rows
        const int slot = yl ^ sw_of_x(x, g.swshift, g.R2);
        const long long inner = ((long long)x * g.R2 + slot) << g.lgT;
        // A CTA that has the SM to itself (512^3 units) runs the last y stage in two halves of sub-blocks: after the
        // first half (one more barrier) its buffer rows are dead and ALL warps issue the copies of those rows of the
        // next unit, which then overlap the stores of the second half; the remaining rows follow after the stage.
        // Measured: 512^3 pass 1 19.6 -> 18.2 ms; with several resident CTAs the other CTAs already hide the loads and
        // the extra barrier costs (200^3: 4.36 -> 4.69 ms), so small units load after the stage.  (A finer scheme -- a
        // counter per sub-block, the warp that completes it issues the copies -- measured slower still:
        // profiles/r02_g2_experiments.md.)
        const int RL = g.fy.st[ylast].radix;
        const int nsub = g.fy.st[ylast].nbf;   // R1 / RL
        const bool halves = MINB == 1 && g.fuse_last && g.nbuf == 1 && !g.pair && nsub >= 2 && u + 1 < nunits &&
                            !(g.dbg_skip & 16);
        if (g.fuse_last) {
#define PS21_G2_YLAST(R, B0, B1) p1_y_last_store<R>(buf32, dst, g, twu, g.pair ? prow : nullptr, yl, inner, B0, B1)
#define PS21_G2_YLAST_ANY(B0, B1)                        \
    switch (RL) {                                        \
        case 2: PS21_G2_YLAST(2, B0, B1); break;         \
        case 3: PS21_G2_YLAST(3, B0, B1); break;         \
        case 4: PS21_G2_YLAST(4, B0, B1); break;         \
        case 5: PS21_G2_YLAST(5, B0, B1); break;         \
        case 6: PS21_G2_YLAST(6, B0, B1); break;         \
        case 8: PS21_G2_YLAST(8, B0, B1); break;         \
        default: PS21_G2_YLAST(10, B0, B1); break;       \
    }
            if (!halves) {
                PS21_G2_YLAST_ANY(0, nsub);
            } else {
                const int parts = (g.parts < nsub) ? g.parts : nsub;
                for (int pt = 0; pt < parts; ++pt) {
                    const int pb0 = nsub * pt / parts, pb1 = nsub * (pt + 1) / parts;
                    PS21_G2_YLAST_ANY(pb0, pb1);
                    __syncthreads();
                    load_rows(u + 1, smem32, pb0 * RL, pb1 * RL, tid >> 5, nthr >> 5);
                }
            }
#undef PS21_G2_YLAST_ANY
#undef PS21_G2_YLAST
        } else {
            const long long tile_elems = (long long)g.n0 * 32;
            for (int e = tid; e < R1 * nh; e += nthr) {
                const int jp = fast_div(e, g.div_nh), r = e - jp * nh;
                float2 v = buf[jp * ld + (g.pair ? (int)prow[r] : r)];
                if (yl) v = cmul_w2(v, twu[jp]);
                dst[((long long)jp * g.nzt + (r >> g.lgT)) * tile_elems + inner + (r & (g.T - 1))] = v;
            }
        }
        if (g.nbuf == 1 && !halves && u + 1 < nunits) {
            __syncthreads();
            issue_unit(u + 1, smem32);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pass 2
// ------------------------------------------------------------------------------------------------
// first x stage (radix R0, sub-block = the whole axis) fused with the R2-point transform over y_lo.
// `tile32` is 256-byte aligned: row x starts at tile32 + 256 x, and inside the row the (slot, t) offset
// ((yl ^ (x & (R2-1))) << (lgT+3)) | (t << 3) is a pure XOR of bit fields, so one row term per m and one XOR per value.
template <int R0, int R2>
__device__ __forceinline__ void p2_first_stage(uint32_t tile32, const G2Dev& g, const float2* __restrict__ tw) {
    constexpr int T = 32 / R2;
    constexpr int lgT = (T == 16) ? 4 : (T == 8) ? 3 : (T == 4) ? 2 : 1;
    constexpr int SH = lgT + 3;
    const int q0 = g.n0 / R0, swshift = g.swshift;
    const int ntasks = q0 * T;
    for (int task = threadIdx.x; task < ntasks; task += blockDim.x) {
        const int t = task & (T - 1), o = task >> lgT;
        float2 v[R0][R2];
        uint32_t row[R0];
#pragma unroll
        for (int m = 0; m < R0; ++m) {
            const int x = o + m * q0;
            row[m] = (tile32 + x * 256 + (t << 3)) ^ (sw_of_x(x, swshift, R2) << SH);
#pragma unroll
            for (int yl = 0; yl < R2; ++yl) v[m][yl] = lds64(row[m] ^ (yl << SH));
        }
#pragma unroll
        for (int m = 0; m < R0; ++m) small_dft<R2>(v[m]);
#pragma unroll
        for (int kyb = 0; kyb < R2; ++kyb) {
            float2 c[R0];
#pragma unroll
            for (int m = 0; m < R0; ++m) c[m] = v[m][kyb];
            small_dft<R0>(c);
#pragma unroll
            for (int m = 0; m < R0; ++m) v[m][kyb] = c[m];
        }
        const int step = o * g.fx.st[0].twmul;
        int idx = 0;
#pragma unroll
        for (int m = 0; m < R0; ++m) {
            const float2 w = tw[idx];
            idx += step;
#pragma unroll
            for (int kyb = 0; kyb < R2; ++kyb)
                sts64(row[m] ^ (kyb << SH), (m == 0) ? v[m][kyb] : cmul_w2(v[m][kyb], w));
        }
    }
}

// Bin state of a thread inside one tile: current bin, power summed in a register while the bin stays the same.
template <int NSHARE>
struct BinWalk {
    float* lbins;      // [nbins][ncol] fp32, private to a lane (NSHARE 1) or to 4 adjacent lanes (NSHARE 4)
    int ncol, col, member;
    uint32_t cur = 255u;
    float acc = 0.f;
    float mult = 1.f;
    __device__ __forceinline__ void put(uint32_t b, float val, bool on) {
        if (NSHARE == 1) {
            if (on && b != 255u) lbins[b * ncol + col] += val;
        } else {   // the lanes of a group take turns (warp-uniform call site)
#pragma unroll
            for (int mm = 0; mm < NSHARE; ++mm) {
                if (on && b != 255u && member == mm) lbins[b * ncol + col] += val;
                __syncwarp();
            }
        }
    }
    __device__ __forceinline__ void add(uint32_t b, float pw) {
        const bool chg = b != cur;
        if (NSHARE == 1) {
            if (chg) {
                put(cur, mult * acc, true);
                cur = b;
                acc = 0.f;
            }
        } else if (__any_sync(0xffffffffu, chg)) {
            put(cur, mult * acc, chg);
            if (chg) {
                cur = b;
                acc = 0.f;
            }
        }
        acc += pw;
    }
    __device__ __forceinline__ void flush() {   // warp-uniform call site
        put(cur, mult * acc, true);
        cur = 255u;
        acc = 0.f;
    }
};

// last x stage (radix RL, contiguous sub-blocks) -> |delta_k|^2 -> bins, straight from registers.
// `release`: called by every thread once its inputs and bin indices sit in registers when the whole stage is one task
// per thread (the usual case: (n0 / RL) * 32 tasks = block size) -- the tile buffers are dead from then on, and the
// bulk copy of the next tile overlaps the butterflies and the binning of this one.
template <int RL, int R2, int NSHARE, bool EARLY, class Release>
__device__ __forceinline__ void p2_last_stage(const uint8_t* __restrict__ lutb,
                                              const G2Dev& g, const AddrTile addr, BinWalk<NSHARE>& bw,
                                              float2* __restrict__ dc_out, Release release) {
    constexpr int NW = (RL + 3) / 4;
    const int ntasks = (g.n0 / RL) * 32;   // EARLY: == blockDim.x, the loop body runs once in every thread
    for (int task = threadIdx.x; task < ntasks; task += blockDim.x) {
        constexpr int lgR2 = (R2 == 16) ? 4 : (R2 == 8) ? 3 : 2;
        int c, bf;
        last_task_map(task, g.mapb, 5 - lgR2, lgR2, c, bf);
        float2 v[RL];
        uint32_t lw[NW];
        {   // x = bf RL + m: the XOR term of the address depends on x & (R2 - 1) only -> R2 row-relative offsets,
            // rotated by (bf RL) & (R2 - 1), and compile-time displacements m * 256.  (With the block index folded into
            // the swizzle -- swshift < 31, RL a multiple of R2 -- the rotation is 0 and the term is XOR-ed with bf.)
            const uint32_t base = addr.base + (uint32_t)(bf * RL) * 256u;
            const int rot = (bf * RL) & (R2 - 1);
            const int bx = (addr.swshift < 31) ? (bf & (R2 - 1)) : 0;
            uint32_t off[R2];
#pragma unroll
            for (int k = 0; k < R2; ++k)
                off[k] = base + (uint32_t)((c << 3) ^ ((((k + rot) & (R2 - 1)) ^ bx) << addr.sh));
#pragma unroll
            for (int m = 0; m < RL; ++m) v[m] = lds64(off[m & (R2 - 1)] + m * 256);
        }
        const uint32_t* lp = reinterpret_cast<const uint32_t*>(lutb) + task * NW;
#pragma unroll
        for (int w4 = 0; w4 < NW; ++w4) lw[w4] = lp[w4];
        if (EARLY) release();
        small_dft<RL>(v);
        if (dc_out && task == 0) *dc_out = v[0];
#pragma unroll
        for (int w4 = 0; w4 < NW; ++w4) {
            const uint32_t word = lw[w4];
            const int cnt = (4 * w4 + 4 <= RL) ? 4 : RL - 4 * w4;   // the last word of a 25-point row holds one output
            float pw[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < cnt) {
                    const float2 o = v[visit_order(4 * w4 + i, RL)];
                    pw[i] = o.x * o.x + o.y * o.y;
                }
            // the bin changes rarely along |k_x|: a whole word of four outputs usually stays in the current bin
            const bool same = cnt == 4 && word == bw.cur * 0x01010101u;
            if (NSHARE == 1 ? same : (bool)__all_sync(0xffffffffu, same)) {
                bw.acc += (pw[0] + pw[1]) + (pw[2] + pw[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < cnt) bw.add((word >> (8 * i)) & 255u, pw[i]);
            }
        }
        bw.flush();   // at most RL fp32 terms per register sum
    }
}

template <int R0, int R2, int NSHARE>
__global__ void __launch_bounds__(512)
k_g2_p2(const float2* __restrict__ scratch, ItemMap im, const ChunkDev* __restrict__ chunks,
        const uint16_t* const* __restrict__ luts, double* __restrict__ partials, long long out_per_box, int slots_max,
        float2* __restrict__ dc, int nchunks, G2Dev g, int nbins) {
    constexpr int T = 32 / R2;
    constexpr int lgT = (T == 16) ? 4 : (T == 8) ? 3 : (T == 4) ? 2 : 1;
    const int tile_elems = g.n0 * 32;
    PS21_DYN_SMEM(float2, smem_raw);
    // the tile starts on a 256-byte boundary of the shared window (the first-stage addressing XORs bit fields)
    float2* tile = smem_raw + ((256u - (smem_u32(smem_raw) & 255u)) & 255u) / 8u;   // [n0][32]
    const int RLk = g.fx.st[g.fx.nstages - 1].radix;
    const uint32_t lut_bytes = (uint32_t)lut_tile_bytes(g.n0, RLk);          // multiple of 128
    uint8_t* lutb = reinterpret_cast<uint8_t*>(tile + tile_elems);           // [lut_bytes]
    uint64_t* bar = reinterpret_cast<uint64_t*>(lutb + lut_bytes);
    double* cbins = reinterpret_cast<double*>(bar + 2);                      // [nbins] fp64 bins of the CTA
    float2* twx = reinterpret_cast<float2*>(cbins + MAX_BINS);               // [n0] w_n0^q
    float* lbins = reinterpret_cast<float*>(twx + g.n0);                     // [nbins][ncol]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int ncol = nthr / NSHARE;
    long long box;
    int cid;
    item_to_box_chunk(im, blockIdx.y, box, cid);
    const ChunkDev ch = chunks[cid];
    const uint8_t* __restrict__ lut = reinterpret_cast<const uint8_t*>(luts[cid]);
    const float2* src = scratch + (long long)blockIdx.y * ((long long)g.ntiles * tile_elems);
    for (int i = tid; i < nbins * ncol; i += nthr) lbins[i] = 0.f;
    for (int i = tid; i < nbins; i += nthr) cbins[i] = 0.0;
    for (int i = tid; i < g.n0; i += nthr) twx[i] = g.tw_x[i];
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        bar[1] = 0;   // address of the first tile of the CTA `pf_ahead` later in launch order (L2 prefetch), or 0
        if (im.pf & 2) {
            const long long n = (long long)blockIdx.y * gridDim.x + blockIdx.x + im.pf_ahead;
            const long long it2 = n / gridDim.x;
            if (it2 < (long long)gridDim.y)
                bar[1] = reinterpret_cast<uint64_t>(scratch + it2 * ((long long)g.ntiles * tile_elems) +
                                                    (n % gridDim.x) * tile_elems);
        }
    }
    __syncthreads();
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    auto issue_tile = [&](int tl) {
        mbar_expect_tx(bar, (uint32_t)tile_elems * 8u + lut_bytes);
        bulk_g2s_hint(tile, src + (long long)tl * tile_elems, (uint32_t)tile_elems * 8u, bar, pol_stream);
        bulk_g2s_hint(lutb, lut + (long long)tl * lut_bytes, lut_bytes, bar, pol_keep);
        if (im.pf) {   // the tile after it -> L2 (or the first tile of the CTA that will take this one's place: bar[1])
            const float2* nx = (tl + g.slots < g.ntiles) ? src + (long long)(tl + g.slots) * tile_elems
                                                         : reinterpret_cast<const float2*>(bar[1]);
            if (nx) bulk_prefetch_l2(nx, (uint32_t)tile_elems * 8u);
        }
    };
    const AddrTile addr{smem_u32(tile), R2 - 1, lgT + 3, g.swshift};
    FastDiv div32;
    div32.mul = 1u << 11;   // x / 32 for x < 65536: (x * 2^11) >> 16
    div32.shift = 16;
    BinWalk<NSHARE> bw;
    bw.lbins = lbins;
    bw.ncol = ncol;
    bw.col = tid / NSHARE;
    bw.member = tid % NSHARE;
    int tl = blockIdx.x;
    if (tid == 0 && tl < g.ntiles) issue_tile(tl);
    uint32_t phase = 0;
    int it = 0;
    // loop invariants (ncu: the two integer divisions below, redone for every tile, were 4 % of the 200^3 instructions):
    // one task per thread and whole warps busy or idle together (the group bin mode votes per warp) -> early hand-over
    const int ntasks_last = (g.n0 / RLk) * 32;
    const bool early = ntasks_last == nthr;
    int zt = tl % g.nzt;                  // z tile of `tl`, advanced with it
    const int zt_step = g.slots % g.nzt;
    for (; tl < g.ntiles; tl += g.slots, ++it, zt = (zt + zt_step >= g.nzt) ? zt + zt_step - g.nzt : zt + zt_step) {
        mbar_wait(bar, phase);
        phase ^= 1;
        p2_first_stage<R0, R2>(addr.base, g, twx);
        __syncthreads();
        for (int s = 1; s + 1 < g.fx.nstages; ++s) {
            ip_stage_any<true>(g.fx.st[s], g.n0, 32, div32, addr, twx, PreSub{make_float2(0.f, 0.f)});
            __syncthreads();
        }
        {
            bw.mult = (float)g.zmult[zt * T + (tid & (T - 1))];
            float2* dc_out = (tl == 0) ? dc + (box * nchunks + cid) : nullptr;
            const bool have_next = tl + g.slots < g.ntiles;
            auto release = [&]() {   // warps 1.. signal without waiting, warp 0 waits for them and starts the copy
                if (!have_next) return;
                if (tid >= 32) {
                    named_bar_arrive(3, nthr);
                } else {
                    named_bar_sync(3, nthr);
                    if (tid == 0) {
                        fence_proxy_async();
                        issue_tile(tl + g.slots);
                    }
                }
            };
#define PS21_G2_LAST(RL)                                                                                   \
    case RL:                                                                                               \
        if (early) p2_last_stage<RL, R2, NSHARE, true>(lutb, g, addr, bw, dc_out, release);                   \
        else p2_last_stage<RL, R2, NSHARE, false>(lutb, g, addr, bw, dc_out, release);                        \
        break
            switch (RLk) {
                PS21_G2_LAST(32);
                PS21_G2_LAST(25);
                PS21_G2_LAST(16);
                PS21_G2_LAST(8);
                PS21_G2_LAST(5);
                default:
                    if (early) p2_last_stage<4, R2, NSHARE, true>(lutb, g, addr, bw, dc_out, release);
                    else p2_last_stage<4, R2, NSHARE, false>(lutb, g, addr, bw, dc_out, release);
                    break;
            }
#undef PS21_G2_LAST
            if (!early) {
                __syncthreads();   // tile and table buffers are dead
                if (tid == 0 && have_next) {
                    fence_proxy_async();
                    issue_tile(tl + g.slots);
                }
            }
        }
        if ((it % FLUSH_EVERY) == FLUSH_EVERY - 1 || tl + g.slots >= g.ntiles) {
            // fixed-order fp32 -> fp64 flush; a lane bin then holds at most RL * tasks * FLUSH_EVERY fp32 terms
            __syncthreads();   // every thread's bins of this tile are complete
            const int warp = tid >> 5, lane = tid & 31;
            for (int b = warp; b < nbins; b += nthr >> 5) {
                double a = 0.0;
                for (int i = lane; i < ncol; i += 32) {
                    float* q = lbins + b * ncol + i;
                    a += (double)*q;
                    *q = 0.f;
                }
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
                if (lane == 0) cbins[b] += a;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    double* out = partials + (box * slots_max + blockIdx.x) * out_per_box + ch.out_off;
    for (int i = tid; i < nbins; i += nthr) out[i] = cbins[i];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Fewest stages first, then the smallest largest radix (balanced stages: 64 = 8 * 8 rather than 16 * 4), radices taken
// from `allowed`; false when n has a prime factor outside the set.
static bool split_radices(int n, const std::vector<int>& allowed, std::vector<int>& out) {
    out.clear();
    if (n == 1) return true;
    struct Best { int stages, maxr; std::vector<int> r; };
    std::vector<Best> memo(n + 1, Best{-1, 0, {}});
    std::vector<char> seen(n + 1, 0);
    std::function<const Best&(int)> go = [&](int m) -> const Best& {
        if (seen[m]) return memo[m];
        seen[m] = 1;
        Best best{1 << 20, 0, {}};
        if (m == 1) best = Best{0, 0, {}};
        else
            for (int r : allowed) {
                if (r > m || m % r) continue;
                const Best& sub = go(m / r);
                if (sub.stages >= (1 << 20)) continue;
                const int st = sub.stages + 1, mx = std::max(sub.maxr, r);
                if (st < best.stages || (st == best.stages && mx < best.maxr)) {
                    best.stages = st;
                    best.maxr = mx;
                    best.r = sub.r;
                    best.r.push_back(r);
                }
            }
        memo[m] = best;
        return memo[m];
    };
    const Best& b = go(n);
    if (b.stages >= (1 << 20) || b.stages > kMaxStages) return false;
    out = b.r;
    std::sort(out.begin(), out.end(), [](int x, int y) { return x > y; });
    return true;
}
static const std::vector<int> kRadZ = {16, 15, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};   // pass 1, z (<= 32 data registers)
static const std::vector<int> kRadY = {13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};           // pass 1, y (two rows per task)
static const std::vector<int> kRadX = {32, 25, 16, 15, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};   // pass 2
static bool is_fusable_y(int r) { return r == 2 || r == 3 || r == 4 || r == 5 || r == 6 || r == 8 || r == 10; }

static void make_ipfft(int n, const std::vector<int>& radices, int twscale, IpFft& f) {
    f = IpFft{};
    f.n = n;
    f.nstages = 0;
    int L = n;
    for (int r : radices) {
        if (r == 1) continue;
        IpStage& st = f.st[f.nstages++];
        st.radix = r;
        st.q = L / r;
        st.twmul = (n / L) * twscale;
        st.nbf = n / r;
        st.div_q = make_fastdiv(st.q);
        L = st.q;
    }
}

// frequency that ends up at position p after the in-place DIF stages
static std::vector<int> freq_of_pos(int n, const IpFft& f) {
    std::vector<int> out(n);
    for (int p = 0; p < n; ++p) {
        int rem = p, L = n, mul = 1, k = 0;
        for (int s = 0; s < f.nstages; ++s) {
            const int q = L / f.st[s].radix;
            k += (rem / q) * mul;
            rem %= q;
            mul *= f.st[s].radix;
            L = q;
        }
        out[p] = k;
    }
    return out;
}

template <class T>
static int upload(const std::vector<T>& v, T** d) {
    PS21_CUDA(cudaMalloc((void**)d, sizeof(T) * (v.empty() ? 1 : v.size())));
    PS21_CUDA(cudaMemcpy(*d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return PS21_OK;
}

static size_t p1_smem(const G2Dev& g) {
    return (size_t)g.nbuf * g.R1 * g.ld * sizeof(float2) +
           (size_t)(g.nz + g.n1 + 2 * g.R1 + 2 * g.Lc + (g.bluestein ? g.M : 0)) * sizeof(float2) +
           (size_t)(g.M + g.R1 + g.nh + 2) * 2 + 64;
}
static size_t p2_smem(const G2Dev& g, int nbins, int nshare) {
    const int RL = g.fx.st[g.fx.nstages - 1].radix;
    return 256 + (size_t)g.n0 * 32 * 8 + (size_t)lut_tile_bytes(g.n0, RL) + 16 + (size_t)MAX_BINS * 8 + (size_t)g.n0 * 8 +
           (size_t)nbins * (g.p2_threads / nshare) * 4;
}

template <int R0, int R2, int NSHARE>
static int set_attr(size_t smem) {
#ifndef PS21_EMU
    PS21_CUDA(cudaFuncSetAttribute(k_g2_p2<R0, R2, NSHARE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
    return PS21_OK;
}

#define PS21_G2_DISPATCH(R0, R2, NSHARE, WHAT)                                                  \
    do {                                                                                        \
        if (R2 == 4 && R0 == 8 && NSHARE == 1) { WHAT(8, 4, 1); }                               \
        else if (R2 == 4 && R0 == 8) { WHAT(8, 4, 4); }                                         \
        else if (R2 == 4 && R0 == 4 && NSHARE == 1) { WHAT(4, 4, 1); }                          \
        else if (R2 == 4 && R0 == 4) { WHAT(4, 4, 4); }                                         \
        else if (R2 == 4 && R0 == 2 && NSHARE == 1) { WHAT(2, 4, 1); }                          \
        else if (R2 == 4 && R0 == 2) { WHAT(2, 4, 4); }                                         \
        else if (R2 == 8 && R0 == 4 && NSHARE == 1) { WHAT(4, 8, 1); }                          \
        else if (R2 == 8 && R0 == 4) { WHAT(4, 8, 4); }                                         \
        else if (R2 == 8 && R0 == 2 && NSHARE == 1) { WHAT(2, 8, 1); }                          \
        else if (R2 == 8 && R0 == 2) { WHAT(2, 8, 4); }                                         \
        else if (R2 == 16 && NSHARE == 1) { WHAT(2, 16, 1); }                                   \
        else { WHAT(2, 16, 4); }                                                                \
    } while (0)

}  // namespace g2

int g2_try_setup(ps21_plan* p, ShapePlan& sp) {
    using namespace g2;
    auto no = [](const char* why) {
        set_error(std::string("two-pass path: ") + why);
        return -1;
    };
    if (p->mode != PS21_MODE_SPHERICAL) return no("spherical bins only");
    if (p->nbins > MAX_BINS) return no("n_psbins <= 254");
    const int n0 = p->n0, n1 = p->n1, nz = sp.nz;
    if (nz < 4) return no("chunk length >= 4 needed");
    if (n0 < 8 || n1 < 8 || n0 > 2048 || n1 > 2048 || nz > 4096) return no("axis length out of range");
    G2Dev g{};
    // even nz: (even, odd)-packed rows, complex transform of length nz / 2 + untangle; odd nz: complex transform of
    // length nz on pairs of rows + separation
    g.pair = nz % 2;
    if (std::getenv("PS21_G2_NOPAIR") && g.pair) return no("odd chunk lengths disabled");
    g.n0 = n0; g.n1 = n1; g.nz = nz; g.M = g.pair ? nz : nz / 2; g.nh = nz / 2 + 1;
    std::vector<int> rz, ry, rx;
    std::vector<int> rzb;   // Bluestein: radices of the convolution length
    if (!split_radices(g.M, kRadZ, rz) || std::getenv("PS21_G2_FORCE_BLUESTEIN")) {
        // nz / 2 has a prime factor above 13 (334 = 2 * 167, ...): chirp convolution of the smallest length >= 2M - 1
        // that the butterflies cover
        g.bluestein = 1;
        for (g.Lc = 2 * g.M - 1; !split_radices(g.Lc, kRadZ, rzb); ++g.Lc) {}
        if (g.Lc > 4096) return no("chunk too long for the chirp convolution");
    }
    g.zrows = std::max(g.pair ? g.M : g.nh, g.Lc);
    if (const char* e = std::getenv("PS21_G2_DBG_SKIP")) g.dbg_skip = std::atoi(e);
    g.parts = 2;
    if (const char* e = std::getenv("PS21_G2_PARTS")) g.parts = std::max(1, std::atoi(e));
    // y split n1 = R1 R2.  Pass 1 keeps nh x R1 coefficients per buffer; with two buffers the next unit streams in while
    // this one is transformed.  R2 = 4 gives 64-byte runs in the scratch, R2 = 8 halves the buffer (32-byte runs).
    int R2 = 0, nbuf = 0;
    int force_r2 = 0, force_nbuf = 0;
    if (const char* e = std::getenv("PS21_G2_R2")) force_r2 = std::atoi(e);
    if (const char* e = std::getenv("PS21_G2_NBUF")) force_nbuf = std::atoi(e);
    auto extra1 = [&](int R1) {
        return (size_t)(nz + n1 + 2 * R1 + 2 * g.Lc + (g.bluestein ? g.M : 0)) * sizeof(float2) +
               (size_t)(g.M + R1 + g.nh + 2) * 2 + 64;
    };
    auto buf1 = [&](int R1) { return (size_t)R1 * (g.zrows | 1) * sizeof(float2); };
    // measured on 200^3 (B200): one buffer with three CTAs per SM 4.83 ms, two buffers with two CTAs 5.24 ms -- occupancy
    // beats overlap inside a CTA, so: three resident CTAs first, then two with the prefetch buffer, then whatever fits
    struct Try { int r2, nb; size_t budget; };
    // R2 = 16 (T = 2: 16-byte runs) is the last resort: measured on 512^3, its three resident CTAs do not make up for
    // the half-sector writes (pass 1 44.5 ms against 20.6 ms with R2 = 8 and one 134 KB CTA per SM)
    const Try order[] = {{4, 1, 74 * 1024},  {8, 1, 74 * 1024},  {4, 2, 112 * 1024}, {8, 2, 112 * 1024},
                         {4, 1, 112 * 1024}, {8, 1, 112 * 1024}, {4, 1, 224 * 1024}, {8, 1, 224 * 1024},
                         {16, 1, 224 * 1024}};
    for (const Try& t : order) {
        if ((force_r2 && t.r2 != force_r2) || (force_nbuf && t.nb != force_nbuf) || n1 % t.r2) continue;
        const int R1 = n1 / t.r2;
        std::vector<int> tmp;
        if (R1 < 2 || !split_radices(R1, kRadY, tmp) || (g.pair && R1 % 2)) continue;
        if (t.nb * buf1(R1) + extra1(R1) <= t.budget) {
            R2 = t.r2;
            nbuf = t.nb;
            break;
        }
    }
    if (!R2) return no("no y split fits the shared memory");
    g.R2 = R2; g.R1 = n1 / R2; g.T = 32 / R2; g.lgT = (g.T == 8) ? 3 : (g.T == 4) ? 2 : 1;
    g.ld = g.zrows | 1;   // complex slots per buffer row (odd: lanes along j and lanes along k_z are both conflict free)
    g.nbuf = nbuf;
    split_radices(g.R1, kRadY, ry);
    // A unit that has the SM to itself (512^3: 134 KB) hides its latencies with threads instead of CTAs: 1024 threads at
    // 64 registers, which caps the butterflies at radix 8 (one more z stage for 256 = 8 * 8 * 4)
    bool big_unit = nbuf * buf1(g.R1) + extra1(g.R1) > 112 * 1024 && std::getenv("PS21_G2_P1_1024");
    if (big_unit) {
        static const std::vector<int> kRad8 = {8, 7, 6, 5, 4, 3, 2};
        std::vector<int> rz8, ry8;
        // (chirp convolution: only when its length also splits into the small radices, e.g. 336 = 8 * 7 * 6)
        if (split_radices(g.bluestein ? g.Lc : g.M, kRad8, rz8) && split_radices(g.R1, kRad8, ry8)) {
            (g.bluestein ? rzb : rz) = rz8;
            ry = ry8;
        } else {
            big_unit = false;
        }
    }
    if (ry.size() >= 2) std::swap(ry.front(), ry.back());   // smallest radix first: the fused first stage holds 2 R values
    g.fuse_first = (!g.pair && is_fusable_y(ry.front())) ? 1 : 0;   // the fused stage untangles packed rows only
    g.zcols = g.pair ? g.R1 / 2 : g.R1;
    g.fuse_last = (ry.size() >= 2 && is_fusable_y(ry.back())) ? 1 : 0;
    if (std::getenv("PS21_G2_NOFUSE")) g.fuse_first = g.fuse_last = 0;
    // x slabs per CTA and resident CTAs go together (200^3, pass 1, same box): 4 CTAs/SM x 2 slabs 4.41 ms, 3 x 1 4.52,
    // 3 x 2 4.67, 4 x 1 4.90
    g.slabs = (n0 % 2 == 0 && nbuf * buf1(g.R1) + extra1(g.R1) <= 55 * 1024) ? 2 : 1;
    if (const char* e = std::getenv("PS21_G2_SLABS")) {   // tuning sweeps only
        const int v = std::atoi(e);
        if (v >= 1 && n0 % v == 0) g.slabs = v;
    }
    // x: first stage R0 with R0 * R2 <= 32 values per thread, the largest remaining radix last (it feeds the bins)
    int R0 = 0;
    for (int cand : (R2 == 4 ? std::vector<int>{8, 4, 2} : R2 == 8 ? std::vector<int>{4, 2} : std::vector<int>{2})) {
        if (n0 % cand) continue;
        if (n0 / cand < 2) continue;
        // the last stage feeds the bins: it must be one of the instantiated ones, the larger the better
        int last = 0;
        std::vector<int> rest;
        for (int want : {32, 25, 16, 8, 5, 4})
            if (!last && (n0 / cand) % want == 0 && split_radices(n0 / cand / want, kRadX, rest)) last = want;
        if (!last) continue;
        rx.clear();
        rx.push_back(cand);
        for (int r : rest) rx.push_back(r);
        rx.push_back(last);
        R0 = cand;
        break;
    }
    if (!R0) return no("sky-plane size does not split into the x radices");
    {   // lane mapping of the last x stage (see last_task_map): needs whole groups of R2 butterflies, and a slot swizzle
        // that separates their rows: x & (R2 - 1) does for odd radices; for the power-of-two ones the block index is
        // folded in (RL a multiple of R2)
        const int RL = rx.back(), nbf = n0 / RL;
        g.lgR2 = (R2 == 16) ? 4 : (R2 == 8) ? 3 : 2;
        g.mapb = (nbf % R2 == 0 && (RL % 2 == 1 || RL % R2 == 0) && !std::getenv("PS21_G2_NOMAPB")) ? 1 : 0;
        g.swshift = 31;
        if (g.mapb && RL % 2 == 0)
            for (g.swshift = 0; (1 << g.swshift) < RL; ++g.swshift) {}
    }
    if ((long long)g.R1 * g.M >= 65536 || (long long)g.R1 * g.nh >= 65536 || (long long)(g.M / 2 + 1) * g.R1 >= 65536 ||
        (long long)g.R1 * g.Lc >= 65536)
        return no("unit too large for the index arithmetic");
    if (!g.bluestein) make_ipfft(g.M, rz, g.pair ? 1 : 2, g.fz);   // packed rows: w_M^q = w_nz^(2q)
    else make_ipfft(g.Lc, rzb, 1, g.fzb);
    make_ipfft(g.R1, ry, R2, g.fy);        // w_R1^q = w_n1^(R2 q)
    make_ipfft(n0, rx, 1, g.fx);
    g.div_M = make_fastdiv(g.M); g.div_R1 = make_fastdiv(g.R1); g.div_nh = make_fastdiv(g.nh);
    g.div_zcols = make_fastdiv(g.zcols);
    {
        const int Mpad = ((g.M + 31) / 32) * 32;
        g.lgMpad = 0;
        if ((Mpad & (Mpad - 1)) == 0 && !std::getenv("PS21_G2_NOFLATLOAD"))
            while ((1 << g.lgMpad) < Mpad) ++g.lgMpad;
    }
    g.nzt = (g.nh + g.T - 1) / g.T;
    g.ntiles = g.R1 * g.nzt;
    {   // pass-2 CTAs per spectrum: about 12 tiles each; 24 for tiles that have the SM to themselves (512-point axes: 4160
        // tiles per spectrum still give 174 CTAs; 512^3 pass 2 7.04 -> 6.95 ms: fewer bin flushes and tails per tile)
        int per = (n0 * 256 > 96 * 1024) ? 24 : 12;
        if (const char* e = std::getenv("PS21_G2_TILES_PER_CTA")) {   // tuning sweeps only
            const int v = std::atoi(e);
            if (v >= 1) per = v;
        }
        g.slots = (g.ntiles + per - 1) / per;
        if (g.slots < 1) g.slots = 1;
    }
    {   // one last-stage butterfly per thread when that gives a sensible block (enables the early tile hand-over)
        const int last_tasks = (n0 / rx.back()) * 32;
        g.p2_threads = (last_tasks >= 128 && last_tasks <= 512) ? last_tasks : ((n0 * 32 >= 12288) ? 512 : 256);
    }
    auto* gs = new G2Shape();
    gs->R0 = R0;
    gs->nshare = p->nbins <= MAX_BINS_LANE ? 1 : 4;
    g.tw_z = sp.tw_z; g.tw_y = sp.tw_y; g.tw_x = sp.tw_x;
    // tables
    std::vector<int> fz_pos = freq_of_pos(g.M, g.fz);
    if (g.bluestein)   // the inverse transform of the convolution leaves the rows in natural order
        for (int r = 0; r < g.M; ++r) fz_pos[r] = r;
    const std::vector<int> fy_pos = freq_of_pos(g.R1, g.fy), fx_pos = freq_of_pos(n0, g.fx);
    std::vector<uint16_t> rowz(g.M), kya(g.R1), kxpos(n0), kzrow((size_t)g.nzt * g.T, 0xFFFF), prow(g.nh);
    std::vector<uint8_t> zmult((size_t)g.nzt * g.T, 0);
    for (int r = 0; r < g.M; ++r) rowz[fz_pos[r]] = (uint16_t)r;
    if (!g.pair) {
        for (int r = 0; r < g.M; ++r) kzrow[r] = (uint16_t)fz_pos[r];
        kzrow[g.M] = (uint16_t)g.M;   // the Nyquist row
        for (int r = 0; r < g.nh; ++r) prow[r] = (uint16_t)r;
    } else {
        // the separated spectrum A(k), k = 0 .. nz/2, stays in the physical row of Z_k; logical rows in physical order
        std::vector<std::pair<int, int>> used;   // (physical row, k)
        for (int k = 0; k < g.nh; ++k) used.push_back({(int)rowz[k], k});
        std::sort(used.begin(), used.end());
        for (int r = 0; r < g.nh; ++r) {
            prow[r] = (uint16_t)used[r].first;
            kzrow[r] = (uint16_t)used[r].second;
        }
    }
    for (int r = 0; r < g.nh; ++r) zmult[r] = (kzrow[r] == 0 || 2 * kzrow[r] == nz) ? 1 : 2;
    for (int j = 0; j < g.R1; ++j) kya[j] = (uint16_t)fy_pos[j];
    for (int x = 0; x < n0; ++x) kxpos[x] = (uint16_t)fx_pos[x];
    int rc = PS21_OK;
    if ((rc = upload(rowz, &gs->d_rowz)) || (rc = upload(kya, &gs->d_kya)) || (rc = upload(kxpos, &gs->d_kxpos)) ||
        (rc = upload(kzrow, &gs->d_kzrow)) || (rc = upload(zmult, &gs->d_zmult)) || (rc = upload(prow, &gs->d_prow))) {
        sp.fused = gs;
        g2_destroy(sp);
        return no("table upload failed");
    }
    if (g.bluestein) {
        const int M = g.M, Lc = g.Lc;
        const double PI = 3.14159265358979323846;
        std::vector<float2> twc(Lc), chirp(M), bhat(Lc);
        std::vector<double> wr(Lc), wi(Lc);
        for (int q = 0; q < Lc; ++q) {
            wr[q] = std::cos(-2.0 * PI * q / Lc);
            wi[q] = std::sin(-2.0 * PI * q / Lc);
            twc[q] = make_float2((float)wr[q], (float)wi[q]);
        }
        // c_n = exp(-i pi n^2 / M); n^2 is reduced mod 2M in integers, so the angle stays small and exact
        std::vector<double> br(Lc, 0.0), bi(Lc, 0.0);   // b_ext = conj(c) wrapped around the convolution length
        for (int n = 0; n < M; ++n) {
            const long long r = ((long long)n * n) % (2LL * M);
            const double a = PI * (double)r / (double)M;
            chirp[n] = make_float2((float)std::cos(a), (float)-std::sin(a));
            br[n] = std::cos(a);
            bi[n] = std::sin(a);
            if (n) {
                br[Lc - n] = br[n];
                bi[Lc - n] = bi[n];
            }
        }
        // its spectrum (direct sum in double, roots by index), divided by Lc, in the position order of the forward plan
        const std::vector<int> fpos = freq_of_pos(Lc, g.fzb);
        for (int pos = 0; pos < Lc; ++pos) {
            const int f = fpos[pos];
            double sr = 0.0, si = 0.0;
            for (int m = 0; m < Lc; ++m) {
                if (br[m] == 0.0 && bi[m] == 0.0) continue;
                const int q = (int)(((long long)m * f) % Lc);
                sr += br[m] * wr[q] - bi[m] * wi[q];
                si += br[m] * wi[q] + bi[m] * wr[q];
            }
            bhat[pos] = make_float2((float)(sr / Lc), (float)(si / Lc));
        }
        if ((rc = upload(twc, &gs->d_twc)) || (rc = upload(chirp, &gs->d_chirp)) || (rc = upload(bhat, &gs->d_bhat))) {
            sp.fused = gs;
            g2_destroy(sp);
            return no("table upload failed");
        }
        g.tw_c = gs->d_twc; g.chirp = gs->d_chirp; g.bhat = gs->d_bhat;
    }
    g.rowz_of_k = gs->d_rowz; g.kya_of_jp = gs->d_kya; g.zmult = gs->d_zmult; g.prow_of_row = gs->d_prow;
    gs->dev = g;
    gs->smem1 = p1_smem(g);
    gs->smem2 = p2_smem(g, p->nbins, gs->nshare);
    gs->p1_threads = big_unit ? 1024 : (gs->smem1 > 112 * 1024) ? 512 : 256;   // a CTA that has the SM to itself gets more threads
    if (gs->smem2 > 227 * 1024 || gs->smem1 > 227 * 1024) {
        sp.fused = gs;
        g2_destroy(sp);
        return no("tile does not fit the shared memory");
    }
#ifndef PS21_EMU
    if (cudaFuncSetAttribute(k_g2_p1<256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(k_g2_p1<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(k_g2_p1<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(k_g2_p1<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(k_g2_p1<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
        cudaGetLastError();
        sp.fused = gs;
        g2_destroy(sp);
        return no("device setup failed");
    }
#endif
    {
        const int r0 = R0, r2 = R2, ns = gs->nshare;
        const size_t sm = 227 * 1024;
#define PS21_G2_ATTR(A, B, C) rc = set_attr<A, B, C>(sm)
        PS21_G2_DISPATCH(r0, r2, ns, PS21_G2_ATTR);
#undef PS21_G2_ATTR
        if (rc) {
            sp.fused = gs;
            g2_destroy(sp);
            return no("device setup failed");
        }
    }
    sp.fused = gs;
    for (int c : sp.chunks) p->chunks[c].nslots = g.slots;
    return 0;
}

void g2_destroy(ShapePlan& sp) {
    if (!sp.fused) return;
    auto* gs = static_cast<g2::G2Shape*>(sp.fused);
    if (gs->d_rowz) cudaFree(gs->d_rowz);
    if (gs->d_kya) cudaFree(gs->d_kya);
    if (gs->d_kxpos) cudaFree(gs->d_kxpos);
    if (gs->d_kzrow) cudaFree(gs->d_kzrow);
    if (gs->d_prow) cudaFree(gs->d_prow);
    if (gs->d_zmult) cudaFree(gs->d_zmult);
    if (gs->d_twc) cudaFree(gs->d_twc);
    if (gs->d_chirp) cudaFree(gs->d_chirp);
    if (gs->d_bhat) cudaFree(gs->d_bhat);
    delete gs;
    sp.fused = nullptr;
}

long long g2_scratch_elems(const ShapePlan& sp) {
    const auto* gs = static_cast<const g2::G2Shape*>(sp.fused);
    return (long long)gs->dev.ntiles * gs->dev.n0 * 32;
}

int g2_build_lut(ps21_plan* p, int chunk, const double* d_k0sq, const double* d_k1sq, const double* d_k2sq,
                 const double* d_edges, cudaStream_t st) {
    const ChunkDev& c = p->chunks[chunk];
    const ShapePlan& sp = p->shapes[c.shape];
    const auto* gs = static_cast<const g2::G2Shape*>(sp.fused);
    const int RL = gs->dev.fx.st[gs->dev.fx.nstages - 1].radix;
    PS21_LAUNCH(g2::k_g2_lut, dim3(PS21_EMU_OR(8, 2048)), dim3(256), 0, st, reinterpret_cast<uint8_t*>(p->d_lut[chunk]),
                (unsigned long long*)(p->d_counts + c.out_off), gs->dev, RL, gs->d_kxpos, gs->d_kzrow, d_k0sq, d_k1sq,
                d_k2sq, d_edges, p->nbins, p->ig_kperp, p->ig_kpar, p->ig_k0);
    PS21_CUDA(cudaGetLastError());
    return PS21_OK;
}

int g2_forward(ps21_plan* p, ShapePlan& sp, const float* d_boxes, long long batch_stride, long long item0,
               long long nitems, cudaStream_t st) {
    using namespace g2;
    auto* gs = static_cast<G2Shape*>(sp.fused);
    const G2Dev& g = gs->dev;
    ItemMap im{sp.d_chunks, (int)sp.chunks.size(), item0};
    const int pf_mode = l2_prefetch_mode();
    const int sms = sm_count(p->device);
    int aligned8 = ((reinterpret_cast<uintptr_t>(d_boxes) & 7) == 0 && batch_stride % 2 == 0 && p->n2_total % 2 == 0) ? 1 : 0;
    for (int c : sp.chunks)
        if (p->chunks[c].start % 2) aligned8 = 0;
    if (aligned8 && (reinterpret_cast<uintptr_t>(d_boxes) & 31) == 0 && batch_stride % 8 == 0 && p->n2_total % 8 == 0) {
        aligned8 = 2;   // every row starts on a 32-byte sector
        for (int c : sp.chunks)
            if (p->chunks[c].start % 8) aligned8 = 1;
    }
    {
        ProfScope ps(p, KK_PASS1, st);
        const int minb = gs->p1_threads != 256 ? 1
                         : (gs->smem1 <= 55 * 1024 && !std::getenv("PS21_G2_P1_MINB3")) ? 4
                         : gs->smem1 <= 74 * 1024 ? 3 : 2;
        im.pf = pf_mode >= 0 ? pf_mode : (minb == 1 ? 1 : 0);
        if (const char* e = std::getenv("PS21_PF_WHERE")) im.pf_where = std::atoi(e);
        im.pf_ahead = prefetch_ahead(minb * sms);
        if (gs->p1_threads == 1024)
            PS21_LAUNCH((k_g2_p1<1024, 1>), dim3(g.n0 / g.slabs, (unsigned)nitems), dim3(1024), gs->smem1, st, d_boxes, batch_stride,
                        p->n2_total, im, p->d_chunks, p->nchunks, p->d_mean_parts, p->d_scratch, g, aligned8);
        else if (gs->p1_threads == 256 && gs->smem1 <= 55 * 1024 && !std::getenv("PS21_G2_P1_MINB3"))
            // four CTAs per SM at 64 registers (a few spills) beat three at 80: 4.43 vs 4.67 ms on 200^3
            PS21_LAUNCH((k_g2_p1<256, 4>), dim3(g.n0 / g.slabs, (unsigned)nitems), dim3(256), gs->smem1, st, d_boxes, batch_stride,
                        p->n2_total, im, p->d_chunks, p->nchunks, p->d_mean_parts, p->d_scratch, g, aligned8);
        else if (gs->p1_threads == 256 && gs->smem1 <= 74 * 1024)   // three CTAs per SM: 80 registers
            PS21_LAUNCH((k_g2_p1<256, 3>), dim3(g.n0 / g.slabs, (unsigned)nitems), dim3(256), gs->smem1, st, d_boxes, batch_stride,
                        p->n2_total, im, p->d_chunks, p->nchunks, p->d_mean_parts, p->d_scratch, g, aligned8);
        else if (gs->p1_threads == 256)                          // two CTAs per SM: 128 registers
            PS21_LAUNCH((k_g2_p1<256, 2>), dim3(g.n0 / g.slabs, (unsigned)nitems), dim3(256), gs->smem1, st, d_boxes, batch_stride,
                        p->n2_total, im, p->d_chunks, p->nchunks, p->d_mean_parts, p->d_scratch, g, aligned8);
        else
            PS21_LAUNCH((k_g2_p1<512, 1>), dim3(g.n0 / g.slabs, (unsigned)nitems), dim3(512), gs->smem1, st, d_boxes, batch_stride,
                        p->n2_total, im, p->d_chunks, p->nchunks, p->d_mean_parts, p->d_scratch, g, aligned8);
        PS21_CUDA(cudaGetLastError());
    }
    {
        ProfScope ps(p, KK_PASS_BIN, st);
        const int r0 = gs->R0, r2 = g.R2, ns = gs->nshare;
        {   // resident pass-2 CTAs per SM: shared memory, or 128 registers per thread
            const int by_smem = (int)((227 * 1024) / (gs->smem2 + 1024)), by_regs = 65536 / (128 * g.p2_threads);
            im.pf = std::max(0, pf_mode);
            im.pf_ahead = prefetch_ahead(std::max(1, std::min(by_smem, by_regs)) * sms);
        }
#define PS21_G2_LAUNCH(A, B, C)                                                                                      \
    PS21_LAUNCH((k_g2_p2<A, B, C>), dim3(g.slots, (unsigned)nitems), dim3(g.p2_threads), gs->smem2, st, p->d_scratch, im, \
                p->d_chunks, (const uint16_t* const*)p->d_lut_table, p->d_partials, p->out_per_box, p->slots_max,   \
                p->d_dc, p->nchunks, g, p->nbins)
        PS21_G2_DISPATCH(r0, r2, ns, PS21_G2_LAUNCH);
#undef PS21_G2_LAUNCH
        PS21_CUDA(cudaGetLastError());
    }
    p->launches_last += 2;
    return PS21_OK;
}

}  // namespace ps21
