// Generic two-pass path ("g2"): the layout of the fused kernels (ps21_fused.cu) for any box whose axes factor into
// the radices of the in-place engine below -- 200^3 and 512^3 coeval cubes (BASELINE configs 2 and 4), lightcone chunks
// of any even length with factors up to 13 (250, 300, 182, ...), sky planes 64 ... 512.
//
// HBM traffic = 4N (read the box) + 8N_h (write the half spectrum) + 8N_h (read it) = the B_alg of SURVEY 8d; the
// three-pass general path (ps21_general.cu) moves 4N + 32N_h.
//
//   y = y_lo + R2 j,  k_y = ky_a + R1 ky_b   (n1 = R1 R2, R2 = 4 or 8, T = 32 / R2)
//
//   pass 1  k_g2_p1   CTA = one x slab, R2 units (y_lo) of R1 z-rows each.  Rows land in shared memory as packed
//                     (even, odd) complex pairs [z/2][j]; z: complex FFT of length nz/2 over the rows + untangle
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
    FastDiv div_q;
};
struct IpFft {
    int n, nstages;
    IpStage st[kMaxStages];
};

struct G2Dev {
    int n0, n1, nz, M, nh;      // M = nz / 2 complex points per packed row, nh = M + 1 half-spectrum rows
    int R1, R2, T, lgT;         // n1 = R1 R2, T = 32 / R2 k_z rows per tile
    int ld;                     // pass 1: complex slots per shared-memory row (odd -> conflict free both ways)
    int nzt, ntiles;            // z tiles per ky_a position, tiles per spectrum
    int slots;                  // pass-2 CTAs per spectrum
    int p2_threads;
    IpFft fz, fy, fx;
    FastDiv div_M, div_R1, div_nh;
    const float2 *tw_z, *tw_y, *tw_x;   // exp(-2 pi i q / n) for n = nz, n1, n0
    const uint16_t* rowz_of_k;          // [M]   row that holds frequency k after the in-place z transform
    const uint16_t* kya_of_jp;          // [R1]  ky_a at position jp after the in-place y transform
    const uint8_t* zmult;               // [nzt T] Hermitian multiplicity of the k_z in row r (0: padding row)
};

struct G2Shape {
    G2Dev dev{};
    uint16_t *d_rowz = nullptr, *d_kya = nullptr, *d_kxpos = nullptr, *d_kzrow = nullptr;
    uint8_t* d_zmult = nullptr;
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
template <int R, class Addr>
__device__ __forceinline__ void ip_stage(float2* __restrict__ buf, const IpStage st, int nbf, int ncols, FastDiv div_cols,
                                         const Addr addr, const float2* __restrict__ tw) {
    const int ntasks = nbf * ncols;
    for (int task = threadIdx.x; task < ntasks; task += blockDim.x) {
        const int bf = fast_div(task, div_cols), c = task - bf * ncols;
        const int b = fast_div(bf, st.div_q), o = bf - b * st.q;
        const int i0 = b * st.q * R + o;
        float2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = buf[addr(i0 + m * st.q, c)];
        small_dft<R>(v);
        if (st.q > 1) {
            const int step = o * st.twmul;
            int idx = step;
#pragma unroll
            for (int m = 1; m < R; ++m) {
                v[m] = cmul_w2(v[m], __ldg(tw + idx));
                idx += step;
            }
        }
#pragma unroll
        for (int m = 0; m < R; ++m) buf[addr(i0 + m * st.q, c)] = v[m];
    }
}

template <bool BIG, class Addr>
__device__ __forceinline__ void ip_stage_any(float2* buf, const IpStage st, int n, int ncols, FastDiv div_cols,
                                             const Addr addr, const float2* tw) {
    const int nbf = n / st.radix;
    switch (st.radix) {
        case 2: ip_stage<2>(buf, st, nbf, ncols, div_cols, addr, tw); break;
        case 3: ip_stage<3>(buf, st, nbf, ncols, div_cols, addr, tw); break;
        case 4: ip_stage<4>(buf, st, nbf, ncols, div_cols, addr, tw); break;
        case 5: ip_stage<5>(buf, st, nbf, ncols, div_cols, addr, tw); break;
        case 7: ip_stage<7>(buf, st, nbf, ncols, div_cols, addr, tw); break;
        case 8: ip_stage<8>(buf, st, nbf, ncols, div_cols, addr, tw); break;
        case 11: ip_stage<11>(buf, st, nbf, ncols, div_cols, addr, tw); break;
        case 13: ip_stage<13>(buf, st, nbf, ncols, div_cols, addr, tw); break;
        case 16: ip_stage<16>(buf, st, nbf, ncols, div_cols, addr, tw); break;
        case 25: if (BIG) ip_stage<25>(buf, st, nbf, ncols, div_cols, addr, tw); break;
        default: if (BIG) ip_stage<32>(buf, st, nbf, ncols, div_cols, addr, tw); break;
    }
}

struct AddrStrided {   // pass 1: index stride si, column stride sc
    int si, sc;
    __device__ __forceinline__ int operator()(int i, int c) const { return i * si + c * sc; }
};
struct AddrTile {      // pass 2: row x holds 32 columns (ky_b, t); the ky_b slot is XOR-ed with the low bits of x
    int swmask, lgT;   // swmask = R2 - 1
    __device__ __forceinline__ int operator()(int i, int c) const { return i * 32 + (c ^ ((i & swmask) << lgT)); }
};

// ------------------------------------------------------------------------------------------------
// LUT (one byte per mode, in the order pass 2 visits its outputs) + counts
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int visit_order(int v, int R) { return (v & 1) ? R - 1 - (v >> 1) : (v >> 1); }

__global__ void k_g2_lut(uint8_t* __restrict__ lut, unsigned long long* __restrict__ counts, G2Dev g, int RL,
                         const uint16_t* __restrict__ kx_of_pos, const uint16_t* __restrict__ kz_of_row,
                         const double* __restrict__ k0sq, const double* __restrict__ k1sq,
                         const double* __restrict__ k2sq, const double* __restrict__ edges, int nbins, int ig_kperp,
                         int ig_kpar, int ig_k0) {
    const long long tile_elems = (long long)g.n0 * 32;
    const long long total = (long long)g.ntiles * tile_elems;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int tile = (int)(i / tile_elems), e = (int)(i % tile_elems);
        // e = (bf * 32 + c) * RL + v : butterfly bf of the last x stage, column c = (ky_b, t), output in visiting order
        const int v = e % RL, task = e / RL, c = task & 31, bf = task >> 5;
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
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_g2_p1(const float* __restrict__ boxes, long long batch_stride, int n2_total, ItemMap im,
                        const ChunkDev* __restrict__ chunks, int nchunks, const double* __restrict__ mean_parts,
                        float2* __restrict__ scratch, G2Dev g, int aligned8) {
    PS21_DYN_SMEM(float2, buf);   // [nh][ld]
    const int tid = threadIdx.x, nthr = blockDim.x;
    long long box;
    int cid;
    item_to_box_chunk(im, blockIdx.y, box, cid);
    const ChunkDev ch = chunks[cid];
    const float mean = chunk_offset(mean_parts, box * nchunks + cid, g.n1, ch.nz);
    const int x = blockIdx.x;
    const int M = g.M, R1 = g.R1, ld = g.ld, nh = g.nh;
    const float* slab = boxes + box * batch_stride + ch.start + (long long)x * g.n1 * n2_total;
    float2* dst = scratch + (long long)blockIdx.y * ((long long)g.ntiles * g.n0 * 32);
    const AddrStrided az{ld, 1};    // z transform: index = row (z pair / k_z), column = j
    const AddrStrided ay{1, ld};    // y transform: index = j, column = row
    for (int yl = 0; yl < g.R2; ++yl) {
        // ---- rows y = yl + R2 j as packed (even, odd) pairs: buf[zc][j]; lanes along z -> coalesced, conflict free
        for (int e = tid; e < R1 * M; e += nthr) {
            const int j = fast_div(e, g.div_M), zc = e - j * M;
            const float* row = slab + (long long)(yl + g.R2 * j) * n2_total;
            float2 v;
            if (aligned8) {
                v = *reinterpret_cast<const float2*>(row + 2 * zc);
            } else {
                v.x = row[2 * zc];
                v.y = row[2 * zc + 1];
            }
            buf[zc * ld + j] = make_float2(v.x - mean, v.y - mean);
        }
        __syncthreads();
        // ---- z: complex FFT of length M over the row index, in place
        for (int s = 0; s < g.fz.nstages; ++s) {
            ip_stage_any<false>(buf, g.fz.st[s], M, R1, g.div_R1, az, g.tw_z);
            __syncthreads();
        }
        // ---- untangle: pairs (p, M - p) -> A_p, A_{M-p}; p = 0 also yields the Nyquist row M; p = M/2 is its own pair
        {
            const int npair = M / 2 + 1;
            for (int e = tid; e < npair * R1; e += nthr) {
                const int p = fast_div(e, g.div_R1), j = e - p * R1;
                if (p == 0) {
                    const float2 c0 = buf[j];   // frequency 0 stays in row 0
                    buf[j] = make_float2(c0.x + c0.y, 0.f);
                    buf[M * ld + j] = make_float2(c0.x - c0.y, 0.f);
                } else if (2 * p == M) {
                    float2* q = buf + g.rowz_of_k[p] * ld + j;
                    const float2 c0 = *q;
                    *q = make_float2(c0.x, -c0.y);
                } else if (2 * p < M) {
                    float2* q1 = buf + g.rowz_of_k[p] * ld + j;
                    float2* q2 = buf + g.rowz_of_k[M - p] * ld + j;
                    const float2 w = __ldg(g.tw_z + p);   // w_nz^p; times -i/2 -> (wy/2, -wx/2)
                    float2 a1, a2;
                    untangle(*q1, *q2, make_float4(0.5f * w.y, -0.5f * w.x, 0.5f * w.x, 0.5f * w.y), a1, a2);
                    *q1 = a1;
                    *q2 = a2;
                }
            }
        }
        __syncthreads();
        // ---- y: FFT of length R1 over j for every one of the nh rows, in place
        for (int s = 0; s < g.fy.nstages; ++s) {
            ip_stage_any<false>(buf, g.fy.st[s], R1, nh, g.div_nh, ay, g.tw_y);
            __syncthreads();
        }
        // ---- times w_n1^(yl ky_a), to HBM in tile order: runs of T consecutive rows
        {
            const int slot = (yl ^ (x & (g.R2 - 1)));
            const long long inner = ((long long)x * g.R2 + slot) << g.lgT;
            for (int e = tid; e < R1 * nh; e += nthr) {
                const int jp = fast_div(e, g.div_nh), r = e - jp * nh;
                float2 v = buf[r * ld + jp];
                if (yl) v = cmul_w2(v, __ldg(g.tw_y + (yl * g.kya_of_jp[jp]) % g.n1));
                const long long tile = (long long)jp * g.nzt + (r >> g.lgT);
                dst[tile * ((long long)g.n0 * 32) + inner + (r & (g.T - 1))] = v;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// pass 2
// ------------------------------------------------------------------------------------------------
// first x stage (radix R0, sub-block = the whole axis) fused with the R2-point transform over y_lo
template <int R0, int R2>
__device__ __forceinline__ void p2_first_stage(float2* __restrict__ tile, const G2Dev& g, const float2* __restrict__ tw) {
    constexpr int T = 32 / R2;
    constexpr int lgT = (T == 16) ? 4 : (T == 8) ? 3 : (T == 4) ? 2 : 1;
    const int q0 = g.n0 / R0;
    const int ntasks = q0 * T;
    for (int task = threadIdx.x; task < ntasks; task += blockDim.x) {
        const int t = task & (T - 1), o = task >> lgT;
        float2 v[R0][R2];
#pragma unroll
        for (int m = 0; m < R0; ++m) {
            const int x = o + m * q0;
#pragma unroll
            for (int yl = 0; yl < R2; ++yl) v[m][yl] = tile[x * 32 + (((yl ^ (x & (R2 - 1))) << lgT) | t)];
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
            const int x = o + m * q0;
            const float2 w = __ldg(tw + idx);
            idx += step;
#pragma unroll
            for (int kyb = 0; kyb < R2; ++kyb)
                tile[x * 32 + (((kyb ^ (x & (R2 - 1))) << lgT) | t)] = (m == 0) ? v[m][kyb] : cmul_w2(v[m][kyb], w);
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

// last x stage (radix RL, contiguous sub-blocks) -> |delta_k|^2 -> bins, straight from registers
template <int RL, int NSHARE>
__device__ __forceinline__ void p2_last_stage(const float2* __restrict__ tile, const uint8_t* __restrict__ lutb,
                                              const G2Dev& g, const AddrTile addr, BinWalk<NSHARE>& bw,
                                              float2* __restrict__ dc_out) {
    const int ntasks = (g.n0 / RL) * 32;
    for (int task = threadIdx.x; task < ntasks; task += blockDim.x) {
        const int c = task & 31, bf = task >> 5;
        float2 v[RL];
#pragma unroll
        for (int m = 0; m < RL; ++m) v[m] = tile[addr(bf * RL + m, c)];
        small_dft<RL>(v);
        if (dc_out && task == 0) *dc_out = v[0];
        const uint8_t* lp = lutb + task * RL;
        if (RL % 4 == 0) {
            const uint32_t* lw = reinterpret_cast<const uint32_t*>(lp);
#pragma unroll
            for (int w4 = 0; w4 < RL / 4; ++w4) {
                const uint32_t word = lw[w4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 o = v[visit_order(4 * w4 + i, RL)];
                    bw.add((word >> (8 * i)) & 255u, o.x * o.x + o.y * o.y);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < RL; ++i) {
                const float2 o = v[visit_order(i, RL)];
                bw.add(lp[i], o.x * o.x + o.y * o.y);
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
    PS21_DYN_SMEM(float2, tile);                                            // [n0][32]
    uint8_t* lutb = reinterpret_cast<uint8_t*>(tile + tile_elems);           // [n0 * 32]
    uint64_t* bar = reinterpret_cast<uint64_t*>(lutb + tile_elems);
    double* cbins = reinterpret_cast<double*>(bar + 2);                      // [nbins] fp64 bins of the CTA
    float* lbins = reinterpret_cast<float*>(cbins + MAX_BINS);               // [nbins][ncol]
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
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    auto issue_tile = [&](int tl) {
        mbar_expect_tx(bar, (uint32_t)tile_elems * 9u);
        bulk_g2s_hint(tile, src + (long long)tl * tile_elems, (uint32_t)tile_elems * 8u, bar, pol_stream);
        bulk_g2s_hint(lutb, lut + (long long)tl * tile_elems, (uint32_t)tile_elems, bar, pol_keep);
    };
    const AddrTile addr{R2 - 1, lgT};
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
    for (; tl < g.ntiles; tl += g.slots, ++it) {
        mbar_wait(bar, phase);
        phase ^= 1;
        p2_first_stage<R0, R2>(tile, g, g.tw_x);
        __syncthreads();
        for (int s = 1; s + 1 < g.fx.nstages; ++s) {
            ip_stage_any<true>(tile, g.fx.st[s], g.n0, 32, div32, addr, g.tw_x);
            __syncthreads();
        }
        {
            const int zt = tl % g.nzt;
            bw.mult = (float)g.zmult[zt * T + (tid & (T - 1))];
            float2* dc_out = (tl == 0) ? dc + (box * nchunks + cid) : nullptr;
            switch (g.fx.st[g.fx.nstages - 1].radix) {
                case 32: p2_last_stage<32, NSHARE>(tile, lutb, g, addr, bw, dc_out); break;
                case 25: p2_last_stage<25, NSHARE>(tile, lutb, g, addr, bw, dc_out); break;
                case 16: p2_last_stage<16, NSHARE>(tile, lutb, g, addr, bw, dc_out); break;
                case 8: p2_last_stage<8, NSHARE>(tile, lutb, g, addr, bw, dc_out); break;
                case 5: p2_last_stage<5, NSHARE>(tile, lutb, g, addr, bw, dc_out); break;
                default: p2_last_stage<4, NSHARE>(tile, lutb, g, addr, bw, dc_out); break;
            }
        }
        __syncthreads();   // tile and table buffers are dead, lane bins of this tile complete
        if (tid == 0 && tl + g.slots < g.ntiles) {
            fence_proxy_async();
            issue_tile(tl + g.slots);
        }
        if ((it % FLUSH_EVERY) == FLUSH_EVERY - 1 || tl + g.slots >= g.ntiles) {
            // fixed-order fp32 -> fp64 flush; a lane bin then holds at most RL * tasks * FLUSH_EVERY fp32 terms
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
static bool split_radices(int n, std::vector<int>& out, bool big = true) {
    // radices of the in-place engine: 25s, then the power of two in pieces of at most 5 bits (as even as possible),
    // then 5, 3, 7, 11, 13; false when another prime remains.  big = false: no 25- and 32-point butterflies (pass 1
    // keeps its registers for occupancy: 16-point pieces, 25 = 5 * 5)
    out.clear();
    int m = n, a = 0;
    while (m % 2 == 0) { m /= 2; ++a; }
    while (big && m % 25 == 0) { out.push_back(25); m /= 25; }
    if (a > 0) {
        const int mb = big ? 5 : 4;
        const int pieces = (a + mb - 1) / mb, base = a / pieces, extra = a % pieces;
        for (int i = 0; i < pieces; ++i) out.push_back(1 << (base + (i < extra ? 1 : 0)));
    }
    for (int p : {5, 3, 7, 11, 13})
        while (m % p == 0) { out.push_back(p); m /= p; }
    if (m != 1 || (int)out.size() > kMaxStages) return false;
    if (out.empty()) out.push_back(1);
    std::sort(out.begin(), out.end(), [](int x, int y) { return x > y; });
    return true;
}

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

static size_t p1_smem(const G2Dev& g) { return (size_t)g.nh * g.ld * sizeof(float2); }
static size_t p2_smem(const G2Dev& g, int nbins, int nshare) {
    return (size_t)g.n0 * 32 * 9 + 16 + (size_t)MAX_BINS * 8 + (size_t)nbins * (g.p2_threads / nshare) * 4;
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
        else { WHAT(2, 8, 4); }                                                                 \
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
    if (nz % 2 || nz < 4) return no("even chunk length >= 4 needed");
    if (n0 < 8 || n1 < 8 || n0 > 2048 || n1 > 2048 || nz > 4096) return no("axis length out of range");
    G2Dev g{};
    g.n0 = n0; g.n1 = n1; g.nz = nz; g.M = nz / 2; g.nh = g.M + 1;
    std::vector<int> rz, ry, rx;
    if (!split_radices(g.M, rz, false)) return no("chunk length / 2 has a prime factor above 13");
    // y split: R2 = 4 (64-byte runs) while the pass-1 buffer leaves two CTAs per SM, else 8
    int R2 = 0;
    const size_t budget2 = 110 * 1024, budget1 = 220 * 1024;
    int force_r2 = 0;
    if (const char* e = std::getenv("PS21_G2_R2")) force_r2 = std::atoi(e);
    for (int cand : {4, 8}) {
        if (force_r2 && cand != force_r2) continue;
        if (n1 % cand) continue;
        const int R1 = n1 / cand;
        std::vector<int> tmp;
        if (R1 < 2 || !split_radices(R1, tmp, false)) continue;
        const size_t need = (size_t)g.nh * (R1 | 1) * sizeof(float2);
        if (need <= (cand == 4 && !force_r2 ? budget2 : budget1)) { R2 = cand; break; }
    }
    if (!R2)
        for (int cand : {4, 8}) {   // second chance: R2 = 4 with one CTA per SM
            if (force_r2 && cand != force_r2) continue;
            if (n1 % cand) continue;
            std::vector<int> tmp;
            if (n1 / cand >= 2 && split_radices(n1 / cand, tmp, false) &&
                (size_t)g.nh * ((n1 / cand) | 1) * sizeof(float2) <= budget1) { R2 = cand; break; }
        }
    if (!R2) return no("no y split fits the shared memory");
    g.R2 = R2; g.R1 = n1 / R2; g.T = 32 / R2; g.lgT = (g.T == 8) ? 3 : 2;
    g.ld = g.R1 | 1;
    split_radices(g.R1, ry, false);
    // x: first stage R0 with R0 * R2 <= 32 values per thread, the largest remaining radix last (it feeds the bins)
    int R0 = 0;
    for (int cand : (R2 == 4 ? std::vector<int>{8, 4, 2} : std::vector<int>{4, 2})) {
        if (n0 % cand) continue;
        std::vector<int> rest;
        if (n0 / cand < 2 || !split_radices(n0 / cand, rest)) continue;
        const int last = rest.front();   // sorted descending
        if (last != 32 && last != 25 && last != 16 && last != 8 && last != 5 && last != 4) continue;
        rx.clear();
        rx.push_back(cand);
        for (size_t i = 1; i < rest.size(); ++i) rx.push_back(rest[i]);
        rx.push_back(last);
        R0 = cand;
        break;
    }
    if (!R0) return no("sky-plane size does not split into the x radices");
    if ((long long)g.R1 * g.M >= 65536 || (long long)g.R1 * g.nh >= 65536 || (long long)(g.M / 2 + 1) * g.R1 >= 65536)
        return no("unit too large for the index arithmetic");
    make_ipfft(g.M, rz, 2, g.fz);          // w_M^q = w_nz^(2q)
    make_ipfft(g.R1, ry, R2, g.fy);        // w_R1^q = w_n1^(R2 q)
    make_ipfft(n0, rx, 1, g.fx);
    g.div_M = make_fastdiv(g.M); g.div_R1 = make_fastdiv(g.R1); g.div_nh = make_fastdiv(g.nh);
    g.nzt = (g.nh + g.T - 1) / g.T;
    g.ntiles = g.R1 * g.nzt;
    {   // pass-2 CTAs per spectrum: about 12 tiles each
        int per = 12;
        g.slots = (g.ntiles + per - 1) / per;
        if (g.slots < 1) g.slots = 1;
    }
    g.p2_threads = (n0 * 32 >= 12288) ? 512 : 256;
    auto* gs = new G2Shape();
    gs->R0 = R0;
    gs->nshare = p->nbins <= MAX_BINS_LANE ? 1 : 4;
    g.tw_z = sp.tw_z; g.tw_y = sp.tw_y; g.tw_x = sp.tw_x;
    // tables
    const std::vector<int> fz_pos = freq_of_pos(g.M, g.fz), fy_pos = freq_of_pos(g.R1, g.fy), fx_pos = freq_of_pos(n0, g.fx);
    std::vector<uint16_t> rowz(g.M), kya(g.R1), kxpos(n0), kzrow((size_t)g.nzt * g.T, 0xFFFF);
    std::vector<uint8_t> zmult((size_t)g.nzt * g.T, 0);
    for (int r = 0; r < g.M; ++r) {
        rowz[fz_pos[r]] = (uint16_t)r;
        kzrow[r] = (uint16_t)fz_pos[r];
    }
    kzrow[g.M] = (uint16_t)g.M;   // the Nyquist row
    for (int r = 0; r < g.nh; ++r) zmult[r] = (kzrow[r] == 0 || 2 * kzrow[r] == nz) ? 1 : 2;
    for (int j = 0; j < g.R1; ++j) kya[j] = (uint16_t)fy_pos[j];
    for (int x = 0; x < n0; ++x) kxpos[x] = (uint16_t)fx_pos[x];
    int rc = PS21_OK;
    if ((rc = upload(rowz, &gs->d_rowz)) || (rc = upload(kya, &gs->d_kya)) || (rc = upload(kxpos, &gs->d_kxpos)) ||
        (rc = upload(kzrow, &gs->d_kzrow)) || (rc = upload(zmult, &gs->d_zmult))) {
        sp.fused = gs;
        g2_destroy(sp);
        return no("table upload failed");
    }
    g.rowz_of_k = gs->d_rowz; g.kya_of_jp = gs->d_kya; g.zmult = gs->d_zmult;
    gs->dev = g;
    gs->smem1 = p1_smem(g);
    gs->smem2 = p2_smem(g, p->nbins, gs->nshare);
    gs->p1_threads = (gs->smem1 > 72 * 1024) ? 512 : 256;   // one or two big CTAs per SM get 512 threads
    if (gs->smem2 > 227 * 1024 || gs->smem1 > 227 * 1024) {
        sp.fused = gs;
        g2_destroy(sp);
        return no("tile does not fit the shared memory");
    }
#ifndef PS21_EMU
    if (cudaFuncSetAttribute(k_g2_p1<256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
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
    if (gs->d_zmult) cudaFree(gs->d_zmult);
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
    int aligned8 = ((reinterpret_cast<uintptr_t>(d_boxes) & 7) == 0 && batch_stride % 2 == 0 && p->n2_total % 2 == 0) ? 1 : 0;
    for (int c : sp.chunks)
        if (p->chunks[c].start % 2) aligned8 = 0;
    {
        ProfScope ps(p, KK_PASS1, st);
        if (gs->p1_threads == 256)
            PS21_LAUNCH((k_g2_p1<256, 3>), dim3(g.n0, (unsigned)nitems), dim3(256), gs->smem1, st, d_boxes, batch_stride,
                        p->n2_total, im, p->d_chunks, p->nchunks, p->d_mean_parts, p->d_scratch, g, aligned8);
        else
            PS21_LAUNCH((k_g2_p1<512, 1>), dim3(g.n0, (unsigned)nitems), dim3(512), gs->smem1, st, d_boxes, batch_stride,
                        p->n2_total, im, p->d_chunks, p->nchunks, p->d_mean_parts, p->d_scratch, g, aligned8);
        PS21_CUDA(cudaGetLastError());
    }
    {
        ProfScope ps(p, KK_PASS_BIN, st);
        const int r0 = gs->R0, r2 = g.R2, ns = gs->nshare;
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
