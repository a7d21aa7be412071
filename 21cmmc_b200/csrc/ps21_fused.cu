// Fused two-pass path for the benchmark shape: 256 x 256 sky plane, 256-slice chunks (BASELINE config 3/5).
//
// HBM traffic = 4N (read lightcone) + 8N_h (write half spectrum) + 8N_h (read it), the B_alg of SURVEY 8d.
//
//   pass 1  k_fused_p1   CTA = one x slab, 4 units (y_lo = 0..3) of 64 z-rows (y = y_lo + 4j) each.
//                        rows land in shared memory by bulk async copies (TMA, cp.async.bulk + mbarrier);
//                        z: 256 real -> 128-point complex FFT on the (even,odd) packing, radix 16 x 8,
//                           in place inside the row, only __syncwarp between stages;
//                        untangle (real-FFT post-processing) fused into the first y stage;
//                        y: 64-point FFT over j (radix 8 x 8, in place), times w_256^(y_lo ky_a);
//                        results go to HBM in tile order so that pass 2 reads one contiguous 64 KB
//                        block per tile: tile = (ky_a, 8 consecutive k_z), inner [x][y_lo][k_z].
//   pass 2  k_fused_p2   CTA loops over tiles: one 64 KB bulk copy per tile (+ its 8 KB of bin indices),
//                        4-point FFT over y_lo and 256-point FFT over x in TWO stages of 32 points per
//                        thread (4x8, then 32; in place, XOR-swizzled 64-byte chunks so both stages are
//                        bank-conflict free) -- the kernel is bound by shared-memory bandwidth, so every
//                        stage saved counts; |delta_k|^2 straight from registers in rising |k_x| order, bin
//                        index from the byte LUT staged next to the tile, lane-private (<= 40 bins) or
//                        4-lane-group (<= 254 bins) fp32 bins in shared memory, flushed every 4 tiles into
//                        fp64 bins in a fixed order (bitwise reproducible), one partial-sum record per CTA
//                        (no global atomics).  The copy of the next tile is requested inside S2, as soon as
//                        every thread holds its inputs and bin indices in registers (bar.arrive/bar.sync).
//
// All transforms are decimation-in-frequency and in place: a thread writes exactly the locations it
// read, so no ping-pong buffer and no barrier between the read and the write of a stage.
#include <cmath>

#include "ps21_internal.cuh"
#include "ptx_sm100.cuh"
#include "../../include/ps21.h"

namespace ps21 {
namespace fz {

constexpr int NX = 256, NY = 256, NZ = 256;   // the headline shape; Geo<128> covers 128 x 128 x 256
constexpr int NHC = NZ / 2;            // 128 complex per packed row
constexpr int R1 = 64, R2 = 4;         // NY = R1 * R2: pass 1 transforms R1 rows, pass 2 the remaining R2
constexpr int T = 8;                   // k_z columns per tile (64-byte runs)
constexpr int ROWLD = 136;             // complex slots per row in pass-1 shared memory (1088 B: bank offset 64 B)
constexpr int TILE = NX * R2 * T;      // 8192 complex = 64 KB (both geometries)
constexpr int NT_MAIN = R1 * (NHC / T);  // 1024 tiles with k_z < 128
constexpr int NTILES = NT_MAIN + R1 / T; // + 8 tiles holding the Nyquist plane (k_z = 128): 1032
constexpr int SLOTS = 86;              // CTAs per spectrum in pass 2 (12 tiles each)
constexpr int P1_THREADS = 256;
constexpr int P2_THREADS = 256;
constexpr int MAX_BINS = 40;           // lane-private bin mode (fastest): at most 40 bins (|k|, or k_perp when cylindrical)
constexpr int MAX_BINS_SPH = 254;      // most spherical bins the byte LUT can address (255 = discarded mode)
constexpr int FLUSH_EVERY = 4;         // lane-private mode: tiles between two fp32 -> fp64 flushes of the lane bins
// pass-2 bin accumulation modes
constexpr int BM_LANE = 0;             // spherical, nbins <= 40: lane-private fp32 bins (conflict free, no shuffles)
constexpr int BM_CYL = 2;              // cylindrical, k_perp bins <= 40: lane-private bins, [bins][8 k_z] per CTA
constexpr int BM_GROUP = 3;            // spherical, 41..254 bins: one fp32 bin column per group of 4 lanes that share
                                       // (warp, k_z) and differ only in m' & 3 -> nearly the same |k|; the 4 lanes update
                                       // the column one after the other (fixed order, no shuffles, no atomics)
constexpr int MAX_BINS_GROUP = 254;    // 254 bins x 64 groups x 4 B = 65 KB (one CTA per SM above 144 bins)
constexpr int NHK = NZ / 2 + 1;        // 129 k_z values of the half spectrum
// cylindrical mode: 64 CTAs take the main tiles with tl % 64 == slot (16 tiles each: always the same 8 k_z),
// 8 CTAs take one Nyquist tile each
constexpr int CYL_SLOTS = 72;
static_assert(NTILES % SLOTS == 0, "tiles must divide evenly over the slots");
static_assert((long long)NTILES * TILE == (long long)NX * NY * (NZ / 2 + 1), "tile order covers the half spectrum exactly");

// w = exp(-2 pi i q / 256) as (w.x, w.y, -w.y, w.x); read through the constant cache where the index is
// warp-uniform (the y twiddles of pass 1), which keeps those reads off the shared-memory pipe
#ifndef PS21_EMU
__constant__ float4 c_tw256[256];
#else
static float4 c_tw256[256];
#endif

struct FusedShape {
    int S = 256;                // sky-plane size: 256 or 128
    double* cylpart = nullptr;  // cylindrical mode: compact partial sums [item][CYL_SLOTS][nbins][8]
    long long cylpart_items = 0;
    float4* tw256 = nullptr;   // w = exp(-2 pi i q / 256) stored as (w.x, w.y, -w.y, w.x) for cmul_tw
};

// ------------------------------------------------------------------------------------------------
// layout of the half-spectrum scratch ("tile order"), per sky-plane size S (S x S x 256 chunks)
// ------------------------------------------------------------------------------------------------
template <int S>
struct Geo;

// Pass 2 (S = 256) bins the 32 outputs r of a thread (k_x = m' + 8 r) in the order 0, 31, 1, 30, ...: |k_x| then
// rises monotonically (m', 8 - m', m' + 8, 16 - m', ...), so a thread meets every bin in one run instead of two (one on
// the way up through the positive, one on the way down through the negative frequencies).  The LUT is stored in that
// visiting order.
__host__ __device__ constexpr int visit256(int i) { return (i & 1) ? 31 - (i >> 1) : (i >> 1); }
__host__ __device__ constexpr int visit8(int j) { return (j & 1) ? 7 - (j >> 1) : (j >> 1); }   // S = 128, see there

// S = 256: y = y_lo + 4 j, tile = (ky_a, 8 consecutive k_z), inner [x 256][y_lo 4][k_z 8]
template <>
struct Geo<256> {
    static constexpr int NXS = 256, R2S = 4, TS = 8, NT_MAIN_S = 1024, NTILES_S = 1032, SLOTS_S = 86, TWMUL = 1;
    static constexpr int P1_SLABS = 1;   // x slabs per pass-1 CTA (two measured 1.4 % slower here)
    // 64-byte chunk (x, y_lo) inside a tile; the low bit of y_lo is flipped by bit0(x)^bit5(x) so that both x
    // stages of pass 2 (lanes over consecutive x, and over x in steps of 32) touch both halves of every
    // 128-byte bank line (conflict free).
    __device__ static __forceinline__ int tile_idx(int x, int yl, int t) {
        return (4 * x + (yl ^ ((x ^ (x >> 5)) & 1))) * 8 + t;
    }
    // where pass 1 stores the coefficient (x, y_lo, ky_a, kz)
    __device__ static __forceinline__ long long scratch_addr(int x, int yl, int kya, int kz) {
        int tile, t;
        if (kz < NHC) {
            tile = kya * 16 + (kz >> 3);
            t = kz & 7;
        } else {
            tile = NT_MAIN_S + (kya >> 3);
            t = kya & 7;
        }
        return (long long)tile * TILE + tile_idx(x, yl, t);
    }
    // which Fourier mode the pass-2 output (tile, task, r) is: r = 0..31 -> kx = m' + 8 r.  The 32 lanes of a
    // warp hold (t, m' & 3): nearly the same |k| for every r, so their bin sequences change at the same r.
    __device__ static __forceinline__ void mode_of(int tile, int task, int r, int& kx, int& ky, int& kz) {
        const int t = task & 7, m1 = ((task >> 3) & 3) | ((task >> 7) << 2), kyb = (task >> 5) & 3;
        kx = m1 + 8 * visit256(r);   // r = position in the visiting order
        int kya;
        if (tile < NT_MAIN_S) {
            kya = tile >> 4;
            kz = 8 * (tile & 15) + t;
        } else {
            kya = 8 * (tile - NT_MAIN_S) + t;
            kz = NHC;
        }
        ky = kya + R1 * kyb;
    }
};

// S = 128: y = y_lo + 2 j, tile = (ky_a, 32 consecutive k_z), inner [x 128][y_lo 2][k_z 32]; every access of
// pass 2 has its lanes along the 32 k_z of a 256-byte chunk, so no swizzle is needed.
template <>
struct Geo<128> {
    static constexpr int NXS = 128, R2S = 2, TS = 32, NT_MAIN_S = 256, NTILES_S = 258, SLOTS_S = 43, TWMUL = 2;
    static constexpr int P1_SLABS = 2;   // x slabs per pass-1 CTA: a slab has only two units, half of them cold starts
    __device__ static __forceinline__ int tile_idx(int x, int yl, int t) { return (2 * x + yl) * 32 + t; }
    __device__ static __forceinline__ long long scratch_addr(int x, int yl, int kya, int kz) {
        int tile, t;
        if (kz < NHC) {
            tile = kya * 4 + (kz >> 5);
            t = kz & 31;
        } else {
            tile = NT_MAIN_S + (kya >> 5);
            t = kya & 31;
        }
        return (long long)tile * TILE + tile_idx(x, yl, t);
    }
    // task = (t, ky_b, m' group), r = 8 i + q: butterfly i of the thread (m' = 4 group + i), output q -> kx = m' + 16 q
    __device__ static __forceinline__ void mode_of(int tile, int task, int r, int& kx, int& ky, int& kz) {
        const int t = task & 31, g = task >> 5;
        const int kyb = g & 1, m1 = 4 * (g >> 1) + (r & 3);   // r = position in the visiting order of pass 2
        kx = m1 + 16 * visit8(r >> 2);
        int kya;
        if (tile < NT_MAIN_S) {
            kya = tile >> 2;
            kz = 32 * (tile & 3) + t;
        } else {
            kya = 32 * (tile - NT_MAIN_S) + t;
            kz = NHC;
        }
        ky = kya + R1 * kyb;
    }
};

// ------------------------------------------------------------------------------------------------
// LUT (one byte per mode, in pass-2 register order) + counts
// ------------------------------------------------------------------------------------------------
template <int S>
__global__ void k_fused_lut(uint8_t* __restrict__ lut, unsigned long long* __restrict__ counts,
                            const double* __restrict__ k0sq, const double* __restrict__ k1sq,
                            const double* __restrict__ k2sq, const double* __restrict__ edges, int nbins,
                            int ig_kperp, int ig_kpar, int ig_k0, int cyl) {
    const long long total = (long long)Geo<S>::NTILES_S * TILE;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int tile = (int)(i / TILE), e = (int)(i % TILE);
        int kx, ky, kz;
        Geo<S>::mode_of(tile, e >> 5, e & 31, kx, ky, kz);
        int w = (kz == 0 || kz == NHC) ? 1 : 2;
        const bool perp0 = (kx == 0 && ky == 0);
        if (ig_kperp && perp0) w = 0;
        if (ig_kpar && kz == 0) w = 0;
        if (ig_k0 && perp0 && kz == 0) w = 0;
        // spherical: |k| bin.  cylindrical: k_perp bin (the k_par index is implied by the tile position)
        const int b = digitize_mode(k0sq[kx], k1sq[ky], cyl ? 0.0 : k2sq[kz], edges, nbins);
        uint8_t v = 255;
        if (b >= 0 && w > 0) {
            v = (uint8_t)b;
            atomicAdd(&counts[cyl ? b * NHK + kz : b], (unsigned long long)w);
        }
        lut[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// pass 1
// ------------------------------------------------------------------------------------------------
// 128-point DIF FFT of one packed row, executed by the 8 threads (t = 0..7) that own the row.
__device__ __forceinline__ void zfft_row(float2* __restrict__ row, int t, float mean, const float2* __restrict__ twz) {
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        v[m] = csub(row[t + 8 * m], make_float2(mean, mean));
    }
    small_dft<16>(v);
#pragma unroll
    for (int m = 1; m < 16; ++m) v[m] = cmul_w2(v[m], twz[8 * m + t]);  // w_128^(t m'), lane-contiguous table
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 16; ++m) row[8 * m + (t ^ (m & 7))] = v[m];  // swizzled inside each group of 8
    __syncwarp();
    float2 a[8], b[8];
    const int m0 = t, m1 = t + 8;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        a[u] = row[8 * m0 + (u ^ (m0 & 7))];
        b[u] = row[8 * m1 + (u ^ (m1 & 7))];
    }
    small_dft<8>(a);
    small_dft<8>(b);
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 8; ++u) {  // C[m' + 16 u'], natural order
        row[m0 + 16 * u] = a[u];
        row[m1 + 16 * u] = b[u];
    }
}

template <int S>
__global__ void __launch_bounds__(P1_THREADS, 3)
k_fused_p1(const float* __restrict__ boxes, long long batch_stride, int n2_total, ItemMap im,
           const ChunkDev* __restrict__ chunks, int nchunks, const double* __restrict__ mean_parts,
           float2* __restrict__ scratch, const float4* __restrict__ tw_g) {
    PS21_DYN_SMEM(float2, buf);                                          // [64][ROWLD]
    float2* twz = buf + R1 * ROWLD;                                      // [16][8] w_128^(t m)
    float4* twu = reinterpret_cast<float4*>(twz + 128);                  // [65]    w_256^p * (-i/2)
    uint64_t* bar = reinterpret_cast<uint64_t*>(twu + 66);               // [8]: one per group of 8 rows
    const int tid = threadIdx.x;
    long long box;
    int cid;
    item_to_box_chunk(im, blockIdx.y, box, cid);
    const ChunkDev ch = chunks[cid];
    const float mean = chunk_offset(mean_parts, box * nchunks + cid, S, ch.nz);
    constexpr int R2S = Geo<S>::R2S;  // units per x slab (y = y_lo + R2S j)
    constexpr int P1_SLABS = Geo<S>::P1_SLABS;
    // A CTA takes P1_SLABS consecutive x slabs of one spectrum; the row prefetch runs across them, so only the first
    // unit of the first slab waits for its loads with nothing to overlap.  (Fully persistent CTAs, 3 per SM looping
    // over everything, were measured 29 % slower: started together they march in phase.)
    int x = blockIdx.x * P1_SLABS;
    const int x_end = x + P1_SLABS;
    const float* spec_src = boxes + box * batch_stride + ch.start;
    const float* src = spec_src + (long long)x * S * n2_total;
    float2* dst = scratch + (long long)blockIdx.y * ((long long)Geo<S>::NTILES_S * TILE);
    if (tid < 128) {
        const float4 w = tw_g[(2 * (tid & 7) * (tid >> 3)) & 255];
        twz[tid] = make_float2(w.x, w.y);
    }
    if (tid >= 192 && tid < 192 + 65) {
        const float4 w = tw_g[tid - 192];  // (wx, wy, -wy, wx);  w * (-i/2) = (wy/2, -wx/2)
        twu[tid - 192] = make_float4(0.5f * w.y, -0.5f * w.x, 0.5f * w.x, 0.5f * w.y);
    }
    if (tid == 0) {
        twu[64] = make_float4(0.f, 0.f, 0.f, 0.f);  // column 64 is handled separately
        const float4 w = tw_g[64];
        twu[64] = make_float4(0.5f * w.y, -0.5f * w.x, 0.5f * w.x, 0.5f * w.y);
        for (int g = 0; g < 8; ++g) mbar_init(bar + g, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint64_t pol = l2_policy_evict_first();
    // rows 8g .. 8g+7 of unit yl, 1 KB each; called by one whole warp
    auto issue_group = [&](const float* src, int yl, int g) {
        const int lane = tid & 31;
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar + g, 8 * NZ * 4);
        }
        __syncwarp();
        if (lane < 8) {
            const int j = 8 * g + lane;
            bulk_g2s_hint(buf + j * ROWLD, src + (long long)(yl + R2S * j) * n2_total, NZ * 4, bar + g, pol);
        }
    };
    // second y stage (radix 8 over rows 8 m1 + u) of one (kz, m1) straight to HBM, times w_S^(y_lo ky_a)
    auto ystage2 = [&](int yl, int kz, int m1) {
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = buf[(8 * m1 + u) * ROWLD + kz];
        small_dft<8>(v);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int kya = m1 + 8 * u;
            // w_S^(y_lo ky_a), warp-uniform index
            const float2 o = (yl == 0) ? v[u] : cmul_tw(v[u], c_tw256[(Geo<S>::TWMUL * yl * kya) & 255]);
            dst[Geo<S>::scratch_addr(x, yl, kya, kz)] = o;
        }
    };
    if (tid < 32)
        for (int g = 0; g < 8; ++g) issue_group(src, 0, g);
    uint32_t phase = 0;
    for (; x < x_end; ++x, src += (long long)S * n2_total)
    for (int yl = 0; yl < R2S; ++yl) {
        // ---- L2 prefetch of the rows whose bulk copies are issued at the end of this unit (the next unit of this
        //      CTA, or the first unit of the CTA that will take its place): by then they come from L2, not DRAM
        if (im.pf && tid >= 192) {   // 64 threads, one 1 KB row each
            const float* nsrc = nullptr;
            int nyl = 0;
            if (yl + 1 < R2S) {
                nsrc = src;
                nyl = yl + 1;
            } else if (x + 1 < x_end) {
                nsrc = src + (long long)S * n2_total;
            } else if (im.pf & 2) {
                const long long n = (long long)blockIdx.y * gridDim.x + blockIdx.x + im.pf_ahead;
                const long long it2 = n / gridDim.x;
                if (it2 < (long long)gridDim.y) {
                    long long b2;
                    int c2;
                    item_to_box_chunk(im, it2, b2, c2);
                    nsrc = boxes + b2 * batch_stride + chunks[c2].start +
                           (long long)((int)(n % gridDim.x) * P1_SLABS) * S * n2_total;
                }
            }
            if (nsrc) bulk_prefetch_l2(nsrc + (long long)(nyl + R2S * (tid - 192)) * n2_total, NZ * 4);
        }
        // ---- z: one 128-point FFT per row, 8 threads per row; rows 0..31 as soon as they have landed
        // thread tid owns rows tid/8 (groups 0..3) and 32 + tid/8 (groups 4..7)
        mbar_wait(bar + (tid >> 6), phase);
        zfft_row(buf + (tid >> 3) * ROWLD, tid & 7, mean, twz);
        mbar_wait(bar + 4 + (tid >> 6), phase);
        zfft_row(buf + (R1 / 2 + (tid >> 3)) * ROWLD, tid & 7, mean, twz);
        phase ^= 1;
        __syncthreads();
        // ---- untangle + first y stage (radix 8 over rows u + 8m), in place: 64 column pairs (p, 128 - p) x 8
        //      butterfly groups u = exactly two tasks per thread; the thread of the cheap pair (0, 128) also
        //      takes the unpaired column 64
        for (int task = tid; task < 64 * 8; task += P1_THREADS) {
            const int p = task & 63, u = task >> 6;
            float2 c1[8], c2[8];
            const float4 w = twu[p];  // w_256^p * (-i/2)
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const float2* r = buf + (u + 8 * m) * ROWLD;
                const float2 ck = r[p];
                if (p == 0) {
                    c1[m] = make_float2(ck.x + ck.y, 0.f);
                    c2[m] = make_float2(ck.x - ck.y, 0.f);
                } else {
                    untangle(ck, r[NHC - p], w, c1[m], c2[m]);
                }
            }
            small_dft<8>(c1);
            small_dft<8>(c2);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                float2* r = buf + (u + 8 * m) * ROWLD;
                const float4 f = c_tw256[(4 * u * m) & 255];  // w_64^(u m'), warp-uniform index
                r[p] = (m == 0) ? c1[m] : cmul_tw(c1[m], f);
                r[NHC - p] = (m == 0) ? c2[m] : cmul_tw(c2[m], f);
            }
            if (p == 0) {  // column 64: A_64 = conj(C_64)
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const float2 ck = buf[(u + 8 * m) * ROWLD + 64];
                    c1[m] = make_float2(ck.x, -ck.y);
                }
                small_dft<8>(c1);
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const float4 f = c_tw256[(4 * u * m) & 255];
                    buf[(u + 8 * m) * ROWLD + 64] = (m == 0) ? c1[m] : cmul_tw(c1[m], f);
                }
            }
        }
        __syncthreads();
        // ---- second y stage straight to HBM.  The Nyquist column (8 tasks) goes first.  Each half of the
        //      CTA (128 threads = the 128 k_z of one butterfly group) walks its own groups m1 = half, half + 2,
        //      ...: a group is 8 rows, and as soon as the half has consumed two of them (128-thread named barrier)
        //      its first warp starts the bulk copies of the next unit into those rows.
        {
            const int half = tid >> 7;
            // the Nyquist column of group m1 = 2q + half is taken by thread q of the half that owns (and later
            // releases) that group, so the half's own barrier orders it before the reload
            if ((tid & 127) < 4) ystage2(yl, NHC, 2 * (tid & 127) + half);
            // two groups per half-CTA barrier (measured: one group per barrier 7.49 ms, two 7.43 ms, all four 7.55 ms)
#pragma unroll 1
            for (int q = 0; q < 4; q += 2) {
                const int m1 = 2 * q + half, m2 = m1 + 2;
                ystage2(yl, tid & (NHC - 1), m1);
                ystage2(yl, tid & (NHC - 1), m2);
                if (half == 0) named_bar_sync(1, 128);
                else named_bar_sync(2, 128);
                if ((tid & 127) < 32) {
                    if (yl + 1 < R2S) {
                        issue_group(src, yl + 1, m1);
                        issue_group(src, yl + 1, m2);
                    } else if (x + 1 < x_end) {   // first unit of the next slab
                        issue_group(src + (long long)S * n2_total, 0, m1);
                        issue_group(src + (long long)S * n2_total, 0, m2);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pass 2
// ------------------------------------------------------------------------------------------------
// BM_LANE / BM_GROUP: spherical bins, partial sums [box][slot][bins] (slot = blockIdx.x of SLOTS).
// BM_CYL    : (k_perp bin, k_z) bins; every thread has one k_z per tile (t = tid & 7), so the lane-private
//              bins are k_perp bins and the per-tile flush keeps the 8 k_z of the tile apart; partial sums go
//              to a compact [item][CYL_SLOTS][bins][8] record that k_fused_cyl_reduce expands.
template <int BM, bool EARLY>
__global__ void __launch_bounds__(P2_THREADS, 2)
k_fused_p2(const float2* __restrict__ scratch, ItemMap im, const ChunkDev* __restrict__ chunks,
           const uint16_t* const* __restrict__ luts, double* __restrict__ partials, long long out_per_box,
           int slots_max, float2* __restrict__ dc, int nchunks, const float4* __restrict__ tw_g, int nbins) {
    // EARLY: the next tile is requested as soon as the tile and bin-index buffers are dead (inside S2, below);
    // otherwise after the binning (a compile-time switch: the barrier branch costs the vote-heavy group mode)
    constexpr int early = EARLY ? 1 : 0;
    PS21_DYN_SMEM(float2, tile);               // [TILE]
    uint8_t* lutb = reinterpret_cast<uint8_t*>(tile + TILE);           // [TILE] bin index of every output of the tile
    float4* tw = reinterpret_cast<float4*>(lutb + TILE);               // [256]
    uint64_t* bar = reinterpret_cast<uint64_t*>(tw + 256);
    constexpr bool CYL = BM == BM_CYL, GROUPB = BM == BM_GROUP;
    constexpr int NCOL = GROUPB ? P2_THREADS / 4 : P2_THREADS;  // private bin columns per CTA
    constexpr int CB = GROUPB ? MAX_BINS_GROUP : MAX_BINS;      // CTA fp64 bins
    // BM_LANE: CTA fp64 bins [MAX_BINS] + lane fp32 bins [nbins][256];  BM_CYL: [MAX_BINS][8] + lane bins;
    double* cbins = reinterpret_cast<double*>(bar + 2);
    float* lbins = reinterpret_cast<float*>(cbins + CB * (CYL ? 8 : 1));  // [nbins][NCOL]
    const int tid = threadIdx.x;
    long long box;
    int cid;
    item_to_box_chunk(im, blockIdx.y, box, cid);
    const ChunkDev ch = chunks[cid];
    const uint8_t* __restrict__ lut = reinterpret_cast<const uint8_t*>(luts[cid]);
    const float2* src = scratch + (long long)blockIdx.y * ((long long)NTILES * TILE);
    tw[tid] = tw_g[tid];
    for (int i = tid; i < nbins * NCOL; i += P2_THREADS) lbins[i] = 0.f;
    for (int i = tid; i < CB * (CYL ? 8 : 1); i += P2_THREADS) cbins[i] = 0.0;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    // tile sequence of this CTA: first, first + stride, ... (count tiles)
    int tl, stride, count;
    if (!CYL) {
        tl = blockIdx.x;
        stride = SLOTS;
        count = NTILES / SLOTS;
    } else if (blockIdx.x < 64) {
        tl = blockIdx.x;
        stride = 64;
        count = NT_MAIN / 64;
    } else {
        tl = NT_MAIN + (blockIdx.x - 64);
        stride = 0;
        count = 1;
    }
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    const int t_end = tl + stride * count;   // first tile past this CTA's sequence
    auto issue_tile = [&](int t_next) {
        mbar_expect_tx(bar, TILE * 8 + TILE);
        bulk_g2s_hint(tile, src + (long long)t_next * TILE, TILE * 8, bar, pol_stream);
        bulk_g2s_hint(lutb, lut + (long long)t_next * TILE, TILE, bar, pol_keep);
        if (im.pf) {   // the tile after it -> L2 (or the first tile of the CTA that will take this one's place)
            if (stride > 0 && t_next + stride < t_end) {
                bulk_prefetch_l2(src + (long long)(t_next + stride) * TILE, TILE * 8);
            } else if ((im.pf & 2) && !CYL) {
                const long long n = (long long)blockIdx.y * gridDim.x + blockIdx.x + im.pf_ahead;
                const long long it2 = n / gridDim.x;
                if (it2 < (long long)gridDim.y)
                    bulk_prefetch_l2(scratch + it2 * ((long long)NTILES * TILE) + (n % gridDim.x) * TILE, TILE * 8);
            }
        }
    };
    if (tid == 0) issue_tile(tl);
    uint32_t phase = 0;
    for (int it = 0; it < count; ++it, tl += stride) {
        mbar_wait(bar, phase);
        phase ^= 1;
        // ---- S1: 4-point FFT over y_lo (-> ky_b) x radix 8 over x = xa + 32 m (-> m'), twiddle w_256^(xa m');
        //      one task of 32 points per thread, in place
        {
            const int t = tid & 7, xa = tid >> 3;
            float2 v[8][4];  // [m][yl]
#pragma unroll
            for (int m = 0; m < 8; ++m)
#pragma unroll
                for (int yl = 0; yl < 4; ++yl) v[m][yl] = tile[Geo<256>::tile_idx(xa + 32 * m, yl, t)];
#pragma unroll
            for (int m = 0; m < 8; ++m) small_dft<4>(v[m]);
#pragma unroll
            for (int yl = 0; yl < 4; ++yl) {
                float2 c[8];
#pragma unroll
                for (int m = 0; m < 8; ++m) c[m] = v[m][yl];
                small_dft<8>(c);
#pragma unroll
                for (int m = 0; m < 8; ++m) v[m][yl] = c[m];
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const float4 f = tw[(xa * m) & 255];
#pragma unroll
                for (int yl = 0; yl < 4; ++yl)
                    tile[Geo<256>::tile_idx(xa + 32 * m, yl, t)] = (m == 0) ? v[m][yl] : cmul_tw(v[m][yl], f);
            }
        }
        __syncthreads();
        // ---- S2: radix 32 over xa inside block m' (-> r, kx = m' + 8 r), |delta_k|^2, bins.
        //      Along r the bin index rises (kx >= 0) and falls again (kx < 0) and the lanes of a warp see
        //      nearly the same |k|: powers are summed in a register while the bin stays the same and the
        //      lane-private shared-memory bin is touched only where some lane changes bin (warp vote).
        {
            const int t = tid & 7, m1 = ((tid >> 3) & 3) | ((tid >> 7) << 2), kyb = (tid >> 5) & 3;
            float2 v[32];
#pragma unroll
            for (int xa = 0; xa < 32; ++xa) v[xa] = tile[Geo<256>::tile_idx(32 * m1 + xa, kyb, t)];
            small_dft<32>(v);
            // Every thread has its 32 inputs and its 32 bin indices in registers: tile and table buffers are dead.
            // Warps 1..7 signal that without waiting (bar.arrive); warp 0 waits for them (bar.sync) and starts the
            // copy of the next tile, which then overlaps the binning and the flush of this tile.  (Issuing before the
            // 32-point transforms, or from the last warp to finish via a counter, measured no better.)
            const uint4 lv2[2] = {reinterpret_cast<const uint4*>(lutb + tid * 32)[0],
                                  reinterpret_cast<const uint4*>(lutb + tid * 32)[1]};
            if (early && it + 1 < count) {
                if (tid >= 32) {
                    named_bar_arrive(3, P2_THREADS);
                } else {
                    named_bar_sync(3, P2_THREADS);
                    if (tid == 0) {
                        fence_proxy_async();
                        issue_tile(tl + stride);
                    }
                }
            }
            if (tl == 0 && tid == 0) dc[box * nchunks + cid] = v[0];
            const float mult = (tl >= NT_MAIN || ((tl & 15) == 0 && t == 0)) ? 1.f : 2.f;
            uint32_t cur = 255u;
            float acc = 0.f;
            // private bin column of this lane (BM_LANE / BM_CYL) or of its 4-lane group (BM_GROUP: lanes t + 8 j)
            const int col = GROUPB ? ((tid & 7) | ((tid >> 5) << 3)) : tid;
            const int member = (tid >> 3) & 3;
            auto lane_update = [&](uint32_t b, float val, bool on) {
                if (!GROUPB) {
                    if (on && b != 255u) lbins[b * NCOL + col] += val;
                } else {  // the four lanes of a group take turns (warp-uniform call site)
#pragma unroll
                    for (int mm = 0; mm < 4; ++mm) {
                        if (on && b != 255u && member == mm) lbins[b * NCOL + col] += val;
                        __syncwarp();
                    }
                }
            };
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint4 lv = lv2[h];
                const uint32_t w4[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
                for (int g = 0; g < 4; ++g) {  // four outputs per 32-bit word of bin indices
                    const int r0 = 16 * h + 4 * g;
                    float pw[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 o = v[visit256(r0 + i)];
                        pw[i] = o.x * o.x + o.y * o.y;
                    }
                    if (__all_sync(0xffffffffu, w4[g] == cur * 0x01010101u)) {
                        acc += (pw[0] + pw[1]) + (pw[2] + pw[3]);  // whole word stays in the current bin, all lanes
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t b = (w4[g] >> (8 * i)) & 255u;
                            const bool chg = b != cur;
                            if (!GROUPB || __any_sync(0xffffffffu, chg)) lane_update(cur, mult * acc, chg);
                            if (chg) {
                                cur = b;
                                acc = 0.f;
                            }
                            acc += pw[i];
                        }
                    }
                }
            }
            lane_update(cur, mult * acc, true);
        }
        // ---- every FLUSH_EVERY tiles (and after the last one): flush the lane-private fp32 bins into the CTA's
        //      fp64 bins in a fixed order; a lane-private bin then holds at most 32 * FLUSH_EVERY fp32 terms.
        //      (Between flushes no barrier is needed here: a thread only touches its own bin column, the tile
        //      buffer was handed over above, and the table buffers alternate.)
        if (!early) {
            __syncthreads();
            if (tid == 0 && it + 1 < count) {
                fence_proxy_async();
                issue_tile(tl + stride);
            }
        }
        if ((it % FLUSH_EVERY) == FLUSH_EVERY - 1 || it + 1 == count) {
            if (early) __syncthreads();
            const int warp = tid >> 5, lane = tid & 31;
            for (int b = warp; b < nbins; b += P2_THREADS / 32) {
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < NCOL / 32; ++i) {
                    float* q = lbins + b * NCOL + lane + 32 * i;
                    acc += (double)*q;
                    *q = 0.f;
                }
                if (!CYL) {
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
                    if (lane == 0) cbins[b] += acc;
                } else {  // lanes with equal (lane & 7) share the k_z of the tile
                    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                    if (lane < 8) cbins[b * 8 + lane] += acc;
                }
            }
        }
        // (the next iteration's barriers order the zeroed lane bins before they are used again)
    }
    __syncthreads();
    if (!CYL) {
        double* out = partials + (box * slots_max + blockIdx.x) * out_per_box + ch.out_off;
        for (int i = tid; i < nbins; i += P2_THREADS) out[i] = cbins[i];
    } else {
        double* out = partials + ((long long)blockIdx.y * CYL_SLOTS + blockIdx.x) * nbins * 8;
        for (int i = tid; i < nbins * 8; i += P2_THREADS) out[i] = cbins[i];
    }
}

// pass 2 for the 128 x 128 sky plane (spherical bins): tile = [x 128][y_lo 2][k_z 32].  Lanes run along the 32
// k_z of a chunk in every access.  S1: 2-point FFT over y_lo x radix 16 over x = xa + 8 m (32 points per
// thread); S2: four radix-8 butterflies over xa per thread (kx = m' + 16 q), |delta_k|^2 and bins from registers.
template <int BM, bool EARLY>
__global__ void __launch_bounds__(P2_THREADS, 2)
k_fused_p2_s128(const float2* __restrict__ scratch, ItemMap im, const ChunkDev* __restrict__ chunks,
                const uint16_t* const* __restrict__ luts, double* __restrict__ partials, long long out_per_box,
                int slots_max, float2* __restrict__ dc, int nchunks, const float4* __restrict__ tw_g, int nbins) {
    using G = Geo<128>;
    PS21_DYN_SMEM(float2, tile);                                       // [TILE]
    uint8_t* lutb = reinterpret_cast<uint8_t*>(tile + TILE);           // [TILE]
    float4* tw = reinterpret_cast<float4*>(lutb + TILE);               // [256]
    uint64_t* bar = reinterpret_cast<uint64_t*>(tw + 256);
    constexpr bool GROUPB = BM == BM_GROUP;
    // BM_GROUP: the four lanes 4c .. 4c+3 (adjacent k_z, nearly the same |k|) share one fp32 bin column
    constexpr int NCOL = GROUPB ? P2_THREADS / 4 : P2_THREADS;
    constexpr int CB = GROUPB ? MAX_BINS_GROUP : MAX_BINS;
    double* cbins = reinterpret_cast<double*>(bar + 2);                // [CB]
    float* lbins = reinterpret_cast<float*>(cbins + CB);               // [nbins][NCOL] (lane / group mode)
    const int tid = threadIdx.x;
    long long box;
    int cid;
    item_to_box_chunk(im, blockIdx.y, box, cid);
    const ChunkDev ch = chunks[cid];
    const uint8_t* __restrict__ lut = reinterpret_cast<const uint8_t*>(luts[cid]);
    const float2* src = scratch + (long long)blockIdx.y * ((long long)G::NTILES_S * TILE);
    tw[tid] = tw_g[tid];
    for (int i = tid; i < nbins * NCOL; i += P2_THREADS) lbins[i] = 0.f;
    if (tid < CB) cbins[tid] = 0.0;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    auto issue_tile = [&](int t_next) {
        mbar_expect_tx(bar, TILE * 8 + TILE);
        bulk_g2s_hint(tile, src + (long long)t_next * TILE, TILE * 8, bar, pol_stream);
        bulk_g2s_hint(lutb, lut + (long long)t_next * TILE, TILE, bar, pol_keep);
    };
    int tl = blockIdx.x;
    constexpr int stride = G::SLOTS_S, count = G::NTILES_S / G::SLOTS_S;
    if (tid == 0) issue_tile(tl);
    uint32_t phase = 0;
    for (int it = 0; it < count; ++it, tl += stride) {
        mbar_wait(bar, phase);
        phase ^= 1;
        {   // ---- S1
            const int t = tid & 31, xa = tid >> 5;
            float2 v[2][16];  // [yl][m]
#pragma unroll
            for (int m = 0; m < 16; ++m)
#pragma unroll
                for (int yl = 0; yl < 2; ++yl) v[yl][m] = tile[G::tile_idx(xa + 8 * m, yl, t)];
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const float2 a = v[0][m], b = v[1][m];
                v[0][m] = cadd(a, b);
                v[1][m] = csub(a, b);
            }
            small_dft<16>(v[0]);
            small_dft<16>(v[1]);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const float4 f = tw[(2 * xa * m) & 255];  // w_128^(xa m')
#pragma unroll
                for (int yl = 0; yl < 2; ++yl)
                    tile[G::tile_idx(xa + 8 * m, yl, t)] = (m == 0 || xa == 0) ? v[yl][m] : cmul_tw(v[yl][m], f);
            }
        }
        __syncthreads();
        {   // ---- S2 + bins
            const int t = tid & 31, g = tid >> 5, kyb = g & 1, mgrp = g >> 1;
            const float mult = (tl >= G::NT_MAIN_S || ((tl & 3) == 0 && t == 0)) ? 1.f : 2.f;
            const uint4* lp = reinterpret_cast<const uint4*>(lutb + tid * 32);
            const uint4 l0 = lp[0], l1 = lp[1];
            const uint32_t w8[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            uint32_t cur = 255u;
            float acc = 0.f;
            const int col = GROUPB ? (tid >> 2) : tid;
            auto lane_update = [&](uint32_t b, float val, bool on) {
                if (!GROUPB) {
                    if (on && b != 255u) lbins[b * NCOL + col] += val;
                } else {  // the four lanes of a group take turns (warp-uniform call site)
#pragma unroll
                    for (int mm = 0; mm < 4; ++mm) {
                        if (on && b != 255u && (tid & 3) == mm) lbins[b * NCOL + col] += val;
                        __syncwarp();
                    }
                }
            };
            float2 vv[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int xa = 0; xa < 8; ++xa) vv[i][xa] = tile[G::tile_idx(xa + 8 * (4 * mgrp + i), kyb, t)];
            // inputs and bin indices are in registers: tile and table buffers are dead (see k_fused_p2)
            if (EARLY && it + 1 < count) {
                if (tid >= 32) {
                    named_bar_arrive(3, P2_THREADS);
                } else {
                    named_bar_sync(3, P2_THREADS);
                    if (tid == 0) {
                        fence_proxy_async();
                        issue_tile(tl + stride);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) small_dft<8>(vv[i]);
            if (tl == 0 && tid == 0) dc[box * nchunks + cid] = vv[0][0];
            // the 32 outputs (k_x = 4 mgrp + i + 16 q) are binned in rising |k_x| order: position p = 4 j + i with
            // q = visit8(j) = 0, 7, 1, 6, ... -> every bin is met in one run; the LUT is stored in that order
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                const float2 o = vv[p & 3][visit8(p >> 2)];
                const uint32_t b = (w8[p >> 2] >> (8 * (p & 3))) & 255u;
                const float pw = o.x * o.x + o.y * o.y;
                if (__any_sync(0xffffffffu, b != cur)) {
                    const bool chg = b != cur;
                    lane_update(cur, mult * acc, chg);
                    if (chg) {
                        cur = b;
                        acc = 0.f;
                    }
                }
                acc += pw;
            }
            lane_update(cur, mult * acc, true);
        }
        if (!EARLY) {
            __syncthreads();
            if (tid == 0 && it + 1 < count) {
                fence_proxy_async();
                issue_tile(tl + stride);
            }
        }
        if ((it % FLUSH_EVERY) == FLUSH_EVERY - 1 || it + 1 == count) {  // fixed-order fp32 -> fp64 flush
            if (EARLY) __syncthreads();
            const int warp = tid >> 5, lane = tid & 31;
            for (int b = warp; b < nbins; b += P2_THREADS / 32) {
                double a = 0.0;
#pragma unroll
                for (int i = 0; i < NCOL / 32; ++i) {
                    float* q = lbins + b * NCOL + lane + 32 * i;
                    a += (double)*q;
                    *q = 0.f;
                }
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) a += __shfl_xor_sync(0xffffffffu, a, sft);
                if (lane == 0) cbins[b] += a;
            }
        }
    }
    __syncthreads();
    double* out = partials + (box * slots_max + blockIdx.x) * out_per_box + ch.out_off;
    for (int i = tid; i < nbins; i += P2_THREADS) out[i] = cbins[i];
}

// cylindrical mode: expand the compact per-CTA records into the standard partial-sum layout (one slot)
__global__ void k_fused_cyl_reduce(const double* __restrict__ cylpart, ItemMap im, const ChunkDev* __restrict__ chunks,
                                   double* __restrict__ partials, long long out_per_box, int slots_max, int nbins) {
    long long box;
    int cid;
    item_to_box_chunk(im, blockIdx.x, box, cid);
    const ChunkDev ch = chunks[cid];
    const double* rec = cylpart + (long long)blockIdx.x * CYL_SLOTS * nbins * 8;
    double* out = partials + (box * slots_max) * out_per_box + ch.out_off;
    for (int i = threadIdx.x; i < nbins * NHK; i += blockDim.x) {
        const int b = i / NHK, kz = i - b * NHK;
        double s = 0.0;
        if (kz < NHC) {
            for (int q = 0; q < 4; ++q) s += rec[(((kz >> 3) + 16 * q) * nbins + b) * 8 + (kz & 7)];
        } else {
            for (int sl = 64; sl < CYL_SLOTS; ++sl)
                for (int t = 0; t < 8; ++t) s += rec[(sl * nbins + b) * 8 + t];
        }
        out[i] = s;
    }
}

static size_t p1_smem() { return (size_t)R1 * ROWLD * 8 + 128 * 8 + 66 * 16 + 64; }
static size_t p2_smem(int nbins, int bm) {
    const size_t bins = bm == BM_GROUP ? (size_t)MAX_BINS_GROUP * 8 + (size_t)nbins * (P2_THREADS / 4) * 4
                                       : (size_t)MAX_BINS * 8 * (bm == BM_CYL ? 8 : 1) + (size_t)nbins * P2_THREADS * 4;
    return (size_t)TILE * 8 + TILE + 256 * 16 + 16 + bins;
}
static int bin_mode(const ps21_plan* p) {
    if (p->mode == PS21_MODE_CYLINDRICAL) return BM_CYL;
    // lane-private columns up to 40 bins; on the 128^2 plane only while two CTAs still fit an SM (<= 36 bins; above,
    // the group mode measured 7 % faster there, 4 % slower on the 256^2 plane)
    if (p->nbins <= MAX_BINS && (p->n0 == 256 || p2_smem(p->nbins, BM_LANE) <= (228 * 1024) / 2 - 1024)) return BM_LANE;
    return BM_GROUP;   // up to MAX_BINS_GROUP = 254 bins (the byte LUT's range); one CTA per SM above 144 bins
}

}  // namespace fz

int fused_try_setup(ps21_plan* p, ShapePlan& sp) {
    using namespace fz;
    const int S = p->n0;
    if (p->n0 != p->n1 || (S != 256 && S != 128) || sp.nz != NZ) {
        set_error("fused path supports 256 x 256 x 256 and 128 x 128 x 256 chunks");
        return -1;
    }
    if (S == 128 && p->mode != PS21_MODE_SPHERICAL) {
        set_error("fused cylindrical bins need the 256 x 256 sky plane");
        return -1;
    }
    if (p->nbins > (p->mode == PS21_MODE_CYLINDRICAL ? MAX_BINS : MAX_BINS_SPH)) {
        set_error("fused path needs n_psbins <= 254 (spherical) / 40 (cylindrical)");
        return -1;
    }
    if (p->n2_total % 4 != 0) {
        set_error("fused path needs n2_total % 4 == 0 (16-byte aligned rows)");
        return -1;
    }
    for (int c : sp.chunks)
        if (p->chunks[c].start % 4 != 0) {
            set_error("fused path needs chunk starts that are multiples of 4 slices");
            return -1;
        }
    auto* fs = new FusedShape();
    fs->S = S;
    std::vector<float4> tw(256);
    for (int q = 0; q < 256; ++q) {
        const double a = -2.0 * 3.14159265358979323846 * q / 256.0;
        const float c = (float)cos(a), sn = (float)sin(a);
        tw[q] = make_float4(c, sn, -sn, c);
    }
#ifndef PS21_EMU
    if (cudaMemcpyToSymbol(c_tw256, tw.data(), sizeof(float4) * 256) != cudaSuccess) {
        cudaGetLastError();
        delete fs;
        set_error("fused path: constant table upload failed");
        return -1;
    }
#else
    std::memcpy(c_tw256, tw.data(), sizeof(float4) * 256);
#endif
    if (cudaMalloc((void**)&fs->tw256, sizeof(float4) * 256) != cudaSuccess ||
        cudaMemcpy(fs->tw256, tw.data(), sizeof(float4) * 256, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaFuncSetAttribute(k_fused_p1<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p1_smem()) != cudaSuccess ||
        cudaFuncSetAttribute(k_fused_p1<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p1_smem()) != cudaSuccess ||
        cudaFuncSetAttribute(k_fused_p2_s128<BM_LANE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)p2_smem(MAX_BINS, BM_LANE)) != cudaSuccess ||
        cudaFuncSetAttribute(k_fused_p2_s128<BM_GROUP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)p2_smem(MAX_BINS_GROUP, BM_GROUP)) != cudaSuccess ||
        cudaFuncSetAttribute(k_fused_p2<BM_LANE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)p2_smem(MAX_BINS, BM_LANE)) != cudaSuccess ||
        cudaFuncSetAttribute(k_fused_p2<BM_LANE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)p2_smem(MAX_BINS, BM_LANE)) != cudaSuccess ||
        cudaFuncSetAttribute(k_fused_p2<BM_GROUP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)p2_smem(MAX_BINS_GROUP, BM_GROUP)) != cudaSuccess ||
        cudaFuncSetAttribute(k_fused_p2<BM_GROUP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)p2_smem(MAX_BINS_GROUP, BM_GROUP)) != cudaSuccess ||
        cudaFuncSetAttribute(k_fused_p2<BM_CYL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)p2_smem(MAX_BINS, BM_CYL)) != cudaSuccess ||
        cudaFuncSetAttribute(k_fused_p2<BM_CYL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)p2_smem(MAX_BINS, BM_CYL)) != cudaSuccess) {
        cudaGetLastError();
        if (fs->tw256) cudaFree(fs->tw256);
        delete fs;
        set_error("fused path: device setup failed");
        return -1;
    }
    sp.fused = fs;
    for (int c : sp.chunks)
        p->chunks[c].nslots = (p->mode == PS21_MODE_CYLINDRICAL) ? 1 : (S == 256 ? SLOTS : Geo<128>::SLOTS_S);
    return 0;
}

void fused_destroy(ShapePlan& sp) {
    if (!sp.fused) return;
    auto* fs = static_cast<fz::FusedShape*>(sp.fused);
    if (fs->tw256) cudaFree(fs->tw256);
    if (fs->cylpart) cudaFree(fs->cylpart);
    delete fs;
    sp.fused = nullptr;
}

long long fused_scratch_elems(const ShapePlan&, int n0, int n1) { return (long long)n0 * n1 * fz::NHK; }

int fused_build_lut(ps21_plan* p, int chunk, const double* d_k0sq, const double* d_k1sq, const double* d_k2sq,
                    const double* d_edges, cudaStream_t st) {
    const ChunkDev& c = p->chunks[chunk];
    if (p->n0 == 256)
        PS21_LAUNCH(fz::k_fused_lut<256>, dim3(PS21_EMU_OR(8, 2048)), dim3(256), 0, st,
                    reinterpret_cast<uint8_t*>(p->d_lut[chunk]), (unsigned long long*)(p->d_counts + c.out_off), d_k0sq,
                    d_k1sq, d_k2sq, d_edges, p->nbins, p->ig_kperp, p->ig_kpar, p->ig_k0,
                    p->mode == PS21_MODE_CYLINDRICAL ? 1 : 0);
    else
        PS21_LAUNCH(fz::k_fused_lut<128>, dim3(PS21_EMU_OR(8, 1024)), dim3(256), 0, st,
                    reinterpret_cast<uint8_t*>(p->d_lut[chunk]), (unsigned long long*)(p->d_counts + c.out_off), d_k0sq,
                    d_k1sq, d_k2sq, d_edges, p->nbins, p->ig_kperp, p->ig_kpar, p->ig_k0, 0);
    PS21_CUDA(cudaGetLastError());
    return PS21_OK;
}

int fused_reserve(ps21_plan* p, ShapePlan& sp, long long nitems) {
    using namespace fz;
    auto* fs = static_cast<FusedShape*>(sp.fused);
    if (p->mode != PS21_MODE_CYLINDRICAL || fs->cylpart_items >= nitems) return PS21_OK;
    drop_graphs(p);
    if (fs->cylpart) cudaFree(fs->cylpart);
    fs->cylpart = nullptr;
    fs->cylpart_items = 0;
    PS21_CUDA(cudaMalloc((void**)&fs->cylpart, sizeof(double) * nitems * CYL_SLOTS * p->nbins * 8));
    fs->cylpart_items = nitems;
    return PS21_OK;
}

int fused_forward(ps21_plan* p, ShapePlan& sp, const float* d_boxes, long long batch_stride, long long item0,
                  long long nitems, cudaStream_t st) {
    using namespace fz;
    auto* fs = static_cast<FusedShape*>(sp.fused);
    if ((reinterpret_cast<uintptr_t>(d_boxes) & 15) != 0 || batch_stride % 4 != 0) {
        set_error("fused path needs 16-byte aligned boxes and batch_stride % 4 == 0");
        return PS21_ERR_INVALID;
    }
    ItemMap im{sp.d_chunks, (int)sp.chunks.size(), item0};
    im.pf = std::max(0, l2_prefetch_mode());
    const int sms = sm_count(p->device);
    {
        ProfScope ps(p, KK_PASS1, st);
        im.pf_ahead = prefetch_ahead(3 * sms);
        if (fs->S == 256)
            PS21_LAUNCH(k_fused_p1<256>, dim3(256 / Geo<256>::P1_SLABS, (unsigned)nitems), dim3(P1_THREADS), p1_smem(), st, d_boxes,
                        batch_stride, p->n2_total, im, p->d_chunks, p->nchunks, p->d_mean_parts, p->d_scratch, fs->tw256);
        else
            PS21_LAUNCH(k_fused_p1<128>, dim3(128 / Geo<128>::P1_SLABS, (unsigned)nitems), dim3(P1_THREADS), p1_smem(), st, d_boxes,
                        batch_stride, p->n2_total, im, p->d_chunks, p->nchunks, p->d_mean_parts, p->d_scratch, fs->tw256);
        PS21_CUDA(cudaGetLastError());
    }
    {
        ProfScope ps(p, KK_PASS_BIN, st);
        const int bm = bin_mode(p);
        // request the next tile inside S2: -8 % in the lane and cylindrical modes, +1.5 % in the vote-heavy group and
        // warp modes (measured), which therefore keep the request after the binning
        const int early = bm == BM_LANE || bm == BM_CYL;
        im.pf_ahead = prefetch_ahead(2 * sms);
#define PS21_P2_LAUNCH(BM, NSLOTS, ...)                                                                           \
    do {                                                                                                          \
        if (early)                                                                                                \
            PS21_LAUNCH((k_fused_p2<BM, true>), dim3(NSLOTS, (unsigned)nitems), dim3(P2_THREADS),                 \
                        p2_smem(p->nbins, BM), st, __VA_ARGS__);                                                  \
        else                                                                                                      \
            PS21_LAUNCH((k_fused_p2<BM, false>), dim3(NSLOTS, (unsigned)nitems), dim3(P2_THREADS),                \
                        p2_smem(p->nbins, BM), st, __VA_ARGS__);                                                  \
    } while (0)
#define PS21_P2_ARGS                                                                                                  \
    p->d_scratch, im, p->d_chunks, (const uint16_t* const*)p->d_lut_table, p->d_partials, p->out_per_box, p->slots_max, \
        p->d_dc, p->nchunks, fs->tw256, p->nbins
        if (fs->S == 128) {
            if (bm == BM_LANE)
                PS21_LAUNCH((k_fused_p2_s128<BM_LANE, true>), dim3(Geo<128>::SLOTS_S, (unsigned)nitems),
                            dim3(P2_THREADS), p2_smem(p->nbins, bm), st, PS21_P2_ARGS);
            else
                PS21_LAUNCH((k_fused_p2_s128<BM_GROUP, false>), dim3(Geo<128>::SLOTS_S, (unsigned)nitems),
                            dim3(P2_THREADS), p2_smem(p->nbins, bm), st, PS21_P2_ARGS);
            PS21_CUDA(cudaGetLastError());
        } else if (bm == BM_LANE) {
            PS21_P2_LAUNCH(BM_LANE, SLOTS, PS21_P2_ARGS);
            PS21_CUDA(cudaGetLastError());
        } else if (bm == BM_GROUP) {
            PS21_P2_LAUNCH(BM_GROUP, SLOTS, PS21_P2_ARGS);
            PS21_CUDA(cudaGetLastError());
        } else {
            {
                const int rrc = fused_reserve(p, sp, nitems);
                if (rrc) return rrc;
            }
            PS21_P2_LAUNCH(BM_CYL, CYL_SLOTS, p->d_scratch, im, p->d_chunks, (const uint16_t* const*)p->d_lut_table,
                           fs->cylpart, p->out_per_box, p->slots_max, p->d_dc, p->nchunks, fs->tw256, p->nbins);
            PS21_CUDA(cudaGetLastError());
            PS21_LAUNCH(k_fused_cyl_reduce, dim3((unsigned)nitems), dim3(256), 0, st, fs->cylpart, im, p->d_chunks,
                        p->d_partials, p->out_per_box, p->slots_max, p->nbins);
            PS21_CUDA(cudaGetLastError());
            p->launches_last += 1;
        }
    }
    p->launches_last += 2;
    return PS21_OK;
}

}  // namespace ps21
