// General three-pass path: any box size (radices 2/3/4/5/8/16, 25- and 32-point stages for the y and x axes,
// generic prime radix for everything else).
//   g1: z pencils, real-to-complex (two real rows per complex transform), offset subtracted on load
//   g2: y transform in place on the half spectrum; columns = flattened (x, k_z) pairs; a CTA pipelines 8 tiles
//       through three rotating shared-memory buffers (cp.async prefetch of tile i+1 while tile i is transformed)
//   g3: x transform fused with |delta_k|^2 and the bin histogram (spectrum is never written back); same pipeline,
//       bin indices from a table kept in this pass's tile order
// HBM traffic 4N + 32 N_h bytes per spectrum.  The fused two-pass kernels (ps21_fused.cu) replace it
// for the power-of-two shapes of the benchmark configurations.
#include "ps21_internal.cuh"
#include "../../include/ps21.h"

namespace ps21 {

static int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

// ---------------------------------------------------------------------------------------------
// LUT: bin index of every half-spectrum mode in the order the x pass consumes it: the (k_y,k_z) plane is cut
// into tiles of `tile` consecutive columns m = y*nh + kz; tile tl holds [x][t] (m = tl*tile + t) in a block of
// lut_blk entries (n0*tile rounded up to 16 bytes) -> one contiguous asynchronous copy per tile.  The table is
// preset to kDiscard, which also covers the padding.
// ---------------------------------------------------------------------------------------------
__global__ void k_general_lut(uint16_t* __restrict__ lut, unsigned long long* __restrict__ counts, int n0, int n1,
                              int nz, int nh, const double* __restrict__ k0sq, const double* __restrict__ k1sq,
                              const double* __restrict__ k2sq, const double* __restrict__ edges, int nbins,
                              int mode, int ig_kperp, int ig_kpar, int ig_k0, int y_zero, int lc,
                              long long lut_blk) {
    const long long total = (long long)n0 * n1 * nh;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int kz = (int)(i % nh);
        const int y = (int)((i / nh) % n1);
        const int x = (int)(i / ((long long)nh * n1));
        int w = (kz == 0 || (2 * kz == nz)) ? 1 : 2;
        const bool perp0 = (x == 0 && y == y_zero);
        if (ig_kperp && perp0) w = 0;
        if (ig_kpar && kz == 0) w = 0;
        if (ig_k0 && perp0 && kz == 0) w = 0;
        int b;
        if (mode == PS21_MODE_SPHERICAL) {
            b = digitize_mode(k0sq[x], k1sq[y], k2sq[kz], edges, nbins);
        } else {
            b = digitize_mode(k0sq[x], k1sq[y], 0.0, edges, nbins);
            if (b >= 0) b = b * nh + kz;
        }
        uint16_t v = kDiscard;
        if (b >= 0 && w > 0) {
            v = (uint16_t)b;
            atomicAdd(&counts[b], (unsigned long long)w);
        }
        const long long m = (long long)y * nh + kz;
        lut[(m >> lc) * lut_blk + ((long long)x << lc) + (m & ((1 << lc) - 1))] = v;
    }
}

long long general_lut_block(int n0, int tile) { return (((long long)n0 * tile + 7) / 8) * 8; }

long long general_lut_elems(int n0, int n1, int nh, int tile) {
    const long long plane = (long long)n1 * nh;
    return ((plane + tile - 1) / tile) * general_lut_block(n0, tile);
}

int general_build_lut(ps21_plan* p, int chunk, const double* d_k0sq, const double* d_k1sq, const double* d_k2sq,
                      const double* d_edges, cudaStream_t st) {
    const ChunkDev& c = p->chunks[chunk];
    const long long total = (long long)p->n0 * p->n1 * c.nh;
    int grid = (int)((total + 255) / 256);
    if (grid > 4096) grid = 4096;
    // y index of the "k_perp = 0" line exactly as the reference writes it: shifted [n0//2, n0//2, :]
    int y_zero = ((p->n0 / 2 - p->n1 / 2) % p->n1 + p->n1) % p->n1;
    const ShapePlan& sp = p->shapes[c.shape];
    PS21_CUDA(cudaMemsetAsync(p->d_lut[chunk], 0xFF,
                              sizeof(uint16_t) * general_lut_elems(p->n0, p->n1, c.nh, sp.g3_tile), st));
    PS21_LAUNCH(k_general_lut, dim3(grid), dim3(256), 0, st, p->d_lut[chunk],
                (unsigned long long*)(p->d_counts + c.out_off), p->n0, p->n1, c.nz, c.nh, d_k0sq, d_k1sq, d_k2sq,
                d_edges, p->nbins, p->mode, p->ig_kperp, p->ig_kpar, p->ig_k0, y_zero, ilog2(sp.g3_tile),
                general_lut_block(p->n0, sp.g3_tile));
    PS21_CUDA(cudaGetLastError());
    return PS21_OK;
}

// ---------------------------------------------------------------------------------------------
// Offset estimate: mean of the x = 0 slab of every chunk (fixed-order reduction -> deterministic).
// Any constant close to the mean works: subtracting it changes only the k = 0 coefficient, which is
// restored in fp64 by the finalize kernel when that mode is binned at all.
// ---------------------------------------------------------------------------------------------

__global__ void k_slab_mean_part(const float* __restrict__ boxes, long long batch_stride, int n1, int n2_total,
                                 const ChunkDev* __restrict__ chunks, int nchunks, double* __restrict__ parts) {
    __shared__ double red[256];
    const int box = blockIdx.x / nchunks;
    const int c = blockIdx.x % nchunks;
    const ChunkDev ch = chunks[c];
    const float* src = boxes + (long long)box * batch_stride + ch.start;
    // rows y = part, part + kMeanParts, ... of the x = 0 slab; one warp per row, lanes along z
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    double acc = 0.0;
    for (int y = blockIdx.y + kMeanParts * warp; y < n1; y += kMeanParts * nwarps) {
        const float* r = src + (long long)y * n2_total;
        for (int z = lane; z < ch.nz; z += 32) acc += (double)r[z];
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) parts[((long long)box * nchunks + c) * kMeanParts + blockIdx.y] = red[0];
}

int launch_slab_mean(ps21_plan* p, const float* d_boxes, long long batch_stride, long long batch, cudaStream_t st) {
    const long long nitems = batch * p->nchunks;
    // the consumers (first FFT pass, finalize) add the kMeanParts partial sums themselves in a fixed order
    // (chunk_offset), so no second launch is needed
    PS21_LAUNCH(k_slab_mean_part, dim3((unsigned)nitems, kMeanParts), dim3(256), 0, st, d_boxes, batch_stride, p->n1,
                p->n2_total, p->d_chunks, p->nchunks, p->d_mean_parts);
    PS21_CUDA(cudaGetLastError());
    p->launches_last += 1;
    return PS21_OK;
}

// ---------------------------------------------------------------------------------------------
// g1: z pass.  CTA = `cols` complex columns = 2*cols consecutive (x,y) rows of one spectrum.
// ---------------------------------------------------------------------------------------------
// launch bounds: 64 registers (four CTAs per SM for short rows) and a known block size; measured 2.68 -> 2.38 ms
// on 64 spectra of 256^3 against the unconstrained build with the same register count
template <bool R32>
__global__ void __launch_bounds__(256, R32 ? 2 : 4) k_general_zpass(const float* __restrict__ boxes, long long batch_stride, int n2_total, int nrows,
                                ItemMap im, const ChunkDev* __restrict__ chunks, int nchunks,
                                const double* __restrict__ mean_parts, int n1, float2* __restrict__ scratch,
                                long long spec_elems, FftFactors fz, const float2* __restrict__ tw, int lc) {
    PS21_DYN_SMEM(float2, smem);
    const int nz = fz.n;
    const int nh = nz / 2 + 1;
    const int cols = 1 << lc;
    const int ld = cols | 1;
    float2* A = smem;
    float2* B = smem + (size_t)nz * ld;
    long long box;
    int cid;
    item_to_box_chunk(im, blockIdx.y, box, cid);
    const ChunkDev ch = chunks[cid];
    const float mean = chunk_offset(mean_parts, box * nchunks + cid, n1, ch.nz);
    const int row0 = blockIdx.x * 2 * cols;
    const float* src = boxes + box * batch_stride + ch.start;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // one pair of real rows (= one complex column) per warp iteration, lanes along z: two coalesced loads and one
    // 8-byte shared-memory store per element (row stride ld is odd -> conflict free)
    for (int pr = warp; pr < cols; pr += nwarps) {
        const int ra = row0 + 2 * pr, rb = ra + 1;
        const float* pa = src + (long long)ra * n2_total;
        const float* pb = src + (long long)rb * n2_total;
        for (int z = lane; z < nz; z += 32) {
            float2 v = make_float2(0.f, 0.f);
            if (ra < nrows) v.x = pa[z] - mean;
            if (rb < nrows) v.y = pb[z] - mean;
            A[z * ld + pr] = v;
        }
    }
    __syncthreads();
    const float2* R = column_fft<R32>(A, B, fz, ld, lc, tw);
    float2* dst = scratch + (long long)blockIdx.y * spec_elems;
    for (int pr = warp; pr < cols; pr += nwarps) {
        const int ra = row0 + 2 * pr, rb = ra + 1;
        if (ra >= nrows) continue;
        float2* oa = dst + (long long)ra * nh;
        float2* ob = dst + (long long)rb * nh;
        for (int k = lane; k < nh; k += 32) {
            const float2 Z = R[k * ld + pr];
            const float2 Zc = R[((k == 0) ? 0 : nz - k) * ld + pr];
            // even row: (Z + conj(Z*))/2, odd row: (Z - conj(Z*)) / (2i)
            oa[k] = make_float2(0.5f * (Z.x + Zc.x), 0.5f * (Z.y - Zc.y));
            if (rb < nrows) ob[k] = make_float2(0.5f * (Z.y + Zc.y), -0.5f * (Z.x - Zc.x));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// g2: y pass, in place.  CTA = fixed x, `tile` consecutive k_z.
// ---------------------------------------------------------------------------------------------
// Columns of the y transform are the (x, k_z) pairs, flattened to q = x*nh + k_z and cut into tiles of `tile`
// consecutive q (no ragged last tile per x: nh = nz/2+1 is odd for even nz).  A CTA owns kYPassItems consecutive
// tiles and pipelines them through three rotating buffers: the rows of tile i+1 stream in with cp.async while tile i
// is transformed and stored.  blockDim is a multiple of the tile width, so a thread always serves the same column.
constexpr int kYPassItems = 8;

template <bool R32>
__global__ void k_general_ypass(float2* __restrict__ scratch, long long spec_elems, int n0, int n1, int nh,
                                FftFactors fy, const float2* __restrict__ tw, int lc) {
    PS21_DYN_SMEM(float2, smem);
    const int tile = 1 << lc;
    const int ld = tile | 1;
    const size_t bufsz = (size_t)n1 * ld;
    const long long ncol = (long long)n0 * nh;
    const int nwork = (int)((ncol + tile - 1) >> lc);
    const int w0 = blockIdx.x * kYPassItems;
    const int w1 = min(nwork, w0 + kYPassItems);
    float2* spec = scratch + (long long)blockIdx.y * spec_elems;
    const int t = threadIdx.x & (tile - 1);
    const int ystep = blockDim.x >> lc, yfirst = threadIdx.x >> lc;
    // address of (y = 0) of this thread's column in tile w, or nullptr past the last column
    auto column = [&](int w) -> float2* {
        const long long q = ((long long)w << lc) + t;
        if (q >= ncol) return nullptr;
        const int x = (int)(q / nh);
        return spec + (long long)x * n1 * nh + (q - (long long)x * nh);
    };
    auto fetch = [&](int w, float2* X) {
        const float2* col = column(w);
        const float2* g = col + (long long)yfirst * nh;
        const long long gs = (long long)ystep * nh;
        float2* sp = X + yfirst * ld + t;
        const int ss = ystep * ld;
        for (int y = yfirst; y < n1; y += ystep) {
            if (col) cp_async8(sp, g);
            else *sp = make_float2(0.f, 0.f);
            g += gs;
            sp += ss;
        }
        cp_async_commit();
    };
    int c = 0;
    if (w0 < w1) fetch(w0, smem);
    for (int w = w0; w < w1; ++w) {
        float2* X0 = smem + c * bufsz;
        float2* X1 = smem + ((c + 1) % 3) * bufsz;
        float2* X2 = smem + ((c + 2) % 3) * bufsz;
        cp_async_wait<0>();
        __syncthreads();   // tile w has landed in X0; the previous tile's result has been stored by everybody
        if (w + 1 < w1) fetch(w + 1, X2);
        const float2* R = column_fft<R32>(X0, X1, fy, ld, lc, tw);
        float2* col = column(w);
        if (col) {
            float2* g = col + (long long)yfirst * nh;
            const long long gs = (long long)ystep * nh;
            const float2* sp = R + yfirst * ld + t;
            const int ss = ystep * ld;
            for (int y = yfirst; y < n1; y += ystep) {
                *g = *sp;
                g += gs;
                sp += ss;
            }
        }
        c = (c + 2) % 3;
    }
}

// ---------------------------------------------------------------------------------------------
// g3: x pass + |delta_k|^2 + histogram.  CTA (slot s of spectrum) loops over tiles of `tile`
// consecutive (k_y,k_z) columns; fp64 bins privatised in shared memory; one flush per CTA into its
// own partial-sum slot (no global atomics -> the cross-CTA reduction order is fixed).
// ---------------------------------------------------------------------------------------------
// Binning walk of the x pass: each thread walks a contiguous k_x segment of one column.  The bin index changes
// rarely along k_x, so powers are summed in a register (at most 32 fp32 terms) and the CTA's fp64 bins in shared
// memory are updated once per run with an atomic add.  (Two warp-cooperative variants without atomics -- a shuffle
// flush per run and lanes-along-k_x warp sums -- were measured 20-130 % slower: the dependent shuffle chains cost
// more than the few contended atomics.)
//
// Software pipeline: three FFT buffers rotate.  While tile i is transformed (ping-pong between two of them) and
// binned, the strided rows of tile i+1 stream into the third with cp.async; the bin indices of tile i arrive in
// their own commit group during the FFT.  The CTA therefore always has a tile of loads in flight.
template <bool R32>
__global__ void k_general_xpass_bin(const float2* __restrict__ scratch, long long spec_elems, int n0, int n1,
                                    int nh, int nz, FftFactors fx, const float2* __restrict__ tw, int lc,
                                    ItemMap im, const ChunkDev* __restrict__ chunks,
                                    const uint16_t* const* __restrict__ luts, long long lut_blk,
                                    double* __restrict__ partials, long long out_per_box, int slots_max,
                                    float2* __restrict__ dc, int nchunks, int ncopy) {
    PS21_DYN_SMEM(float2, smem_raw);
    long long box;
    int cid;
    item_to_box_chunk(im, blockIdx.y, box, cid);
    const ChunkDev ch = chunks[cid];
    const uint16_t* __restrict__ lut = luts[cid];
    const int tile = 1 << lc;
    const int ld = tile | 1;
    const size_t bufsz = (size_t)n0 * ld;
    // layout: bin indices first (16-byte aligned for the 16-byte async copies; lut_blk*2 is a multiple of 16),
    // then the three FFT buffers, then the fp64 bins
    uint16_t* lutS = reinterpret_cast<uint16_t*>(smem_raw);            // [n0][tile] bin index of the tile's modes
    float2* smem = reinterpret_cast<float2*>(lutS + lut_blk);
    double* bins = reinterpret_cast<double*>(smem + 3 * bufsz);
    // `ncopy` (power of two) copies of the fp64 bins, chosen by the low thread-index bits (= the k_z column of the
    // tile): threads that flush at the same step mostly hit different copies -> short CAS loops in the atomics
    const int nbin_store = ncopy * ch.nbins_out;
    for (int i = threadIdx.x; i < nbin_store; i += blockDim.x) bins[i] = 0.0;
    double* mybins = bins + (threadIdx.x & (ncopy - 1)) * ch.nbins_out;
    const long long plane = (long long)n1 * nh;
    const long long ntiles = (plane + tile - 1) / tile;
    const float2* base = scratch + (long long)blockIdx.y * spec_elems;
    // blockDim is a multiple of the tile width: a thread always serves column t and walks x in steps of xs
    const int tcol = threadIdx.x & (tile - 1), xfirst = threadIdx.x >> lc, xs = blockDim.x >> lc;
    auto fetch_tile = [&](long long tl, float2* X) {
        const long long m0 = tl * tile;
        const bool in = m0 + tcol < plane;
        const float2* g = base + (long long)xfirst * plane + m0 + tcol;
        const long long gs = (long long)xs * plane;
        float2* sp = X + xfirst * ld + tcol;
        const int ss = xs * ld;
        for (int x = xfirst; x < n0; x += xs) {
            if (in) cp_async8(sp, g);
            else *sp = make_float2(0.f, 0.f);
            g += gs;
            sp += ss;
        }
        cp_async_commit();
    };
    int c = 0;
    if ((long long)blockIdx.x < ntiles) fetch_tile(blockIdx.x, smem);
    for (long long tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const long long m0 = tl * tile;
        const int cols = (int)min((long long)tile, plane - m0);
        float2* X0 = smem + c * bufsz;
        float2* X1 = smem + ((c + 1) % 3) * bufsz;
        float2* X2 = smem + ((c + 2) % 3) * bufsz;
        cp_async_wait<0>();
        __syncthreads();   // tile tl has landed in X0; everybody is done with the previous tile's buffers and lutS
        {
            const uint16_t* lsrc = lut + tl * lut_blk;
            for (int i = threadIdx.x; i < (int)(lut_blk >> 3); i += blockDim.x) cp_async16(lutS + 8 * i, lsrc + 8 * i);
            cp_async_commit();
        }
        if (tl + gridDim.x < ntiles) fetch_tile(tl + gridDim.x, X2);
        else cp_async_commit();   // keep the group count uniform
        const float2* R = column_fft<R32>(X0, X1, fx, ld, lc, tw);
        cp_async_wait<1>();        // bin indices of this tile (the next tile's rows may still be in flight)
        __syncthreads();
        {
            const int nseg = blockDim.x >> lc;                   // segments per column
            const int t = threadIdx.x & (tile - 1), seg = threadIdx.x >> lc;
            const int xlen = (n0 + nseg - 1) / nseg;
            const int x0 = seg * xlen, x1 = min(n0, x0 + xlen);
            const long long m = m0 + t;
            const int kz = (int)(m % nh);
            const float w = (kz == 0 || 2 * kz == nz) ? 1.f : 2.f;
            if (t < cols) {
                unsigned cur = kDiscard;
                float acc = 0.f;
                int run = 0;
                auto step = [&](float2 v, unsigned b) {
                    if (b != cur || run == 32) {
                        if (cur != kDiscard) atomicAdd(&mybins[cur], (double)(w * acc));
                        cur = b;
                        acc = 0.f;
                        run = 0;
                    }
                    acc += v.x * v.x + v.y * v.y;
                    ++run;
                };
                if (x0 == 0 && m == 0 && x1 > 0) dc[box * nchunks + cid] = R[t];
                int x = x0;
                for (; x + 4 <= x1; x += 4) {   // four independent shared-memory loads in flight per step
                    float2 v[4];
                    unsigned b[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v[j] = R[(x + j) * ld + t];
                        b[j] = lutS[((x + j) << lc) + t];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) step(v[j], b[j]);
                }
                for (; x < x1; ++x) step(R[x * ld + t], lutS[(x << lc) + t]);
                if (cur != kDiscard) atomicAdd(&mybins[cur], (double)(w * acc));
            }
        }
        c = (c + 2) % 3;
    }
    cp_async_wait<0>();
    __syncthreads();
    double* out = partials + (box * slots_max + blockIdx.x) * out_per_box + ch.out_off;
    for (int i = threadIdx.x; i < ch.nbins_out; i += blockDim.x) {
        double a = 0.0;
        for (int q = 0; q < ncopy; ++q) a += bins[q * ch.nbins_out + i];
        out[i] = a;
    }
}

// shared memory of the x pass besides the two ping-pong FFT buffers
int general_bin_copies(int nb_out) {
    int c = 16;
    while (c > 1 && (size_t)c * nb_out * sizeof(double) > 16 * 1024) c >>= 1;
    return c;
}

size_t general_xpass_extra(int nb_out, int n0, int tile) {
    return (size_t)n0 * (tile | 1) * sizeof(float2)   // third FFT buffer (prefetch target)
           + (size_t)general_bin_copies(nb_out) * nb_out * sizeof(double) + (size_t)general_lut_block(n0, tile) * sizeof(uint16_t);
}

int general_init() {
#ifndef PS21_EMU
    const int big = 220 * 1024;
    PS21_CUDA(cudaFuncSetAttribute(k_general_zpass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    PS21_CUDA(cudaFuncSetAttribute(k_general_ypass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    PS21_CUDA(cudaFuncSetAttribute(k_general_xpass_bin<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    PS21_CUDA(cudaFuncSetAttribute(k_general_zpass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    PS21_CUDA(cudaFuncSetAttribute(k_general_ypass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    PS21_CUDA(cudaFuncSetAttribute(k_general_xpass_bin<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
#endif
    return PS21_OK;
}

static size_t fft_smem_bytes(int n, int tile) { return 2 * (size_t)n * (tile | 1) * sizeof(float2); }

static bool has_radix32(const FftFactors& f) {
    for (int s = 0; s < f.nstages; ++s)
        if (f.radix[s] == 32 || f.radix[s] == 25) return true;
    return false;
}

int general_forward(ps21_plan* p, ShapePlan& sp, const float* d_boxes, long long batch_stride, long long item0,
                    long long nitems, cudaStream_t st) {
    const int n0 = p->n0, n1 = p->n1;
    const long long spec_elems = (long long)n0 * n1 * sp.nh;
    ItemMap im{sp.d_chunks, (int)sp.chunks.size(), item0};
    const int nrows = n0 * n1;
    {
        ProfScope ps(p, KK_PASS1, st);
        dim3 grid((nrows + 2 * sp.g1_cols - 1) / (2 * sp.g1_cols), (unsigned)nitems);
#define PS21_ZPASS_ARGS                                                                                            \
    d_boxes, batch_stride, p->n2_total, nrows, im, p->d_chunks, p->nchunks, p->d_mean_parts, p->n1, p->d_scratch, spec_elems,  \
        sp.fz, sp.tw_z, ilog2(sp.g1_cols)
        if (has_radix32(sp.fz))
            PS21_LAUNCH(k_general_zpass<true>, grid, dim3(256), fft_smem_bytes(sp.nz, sp.g1_cols), st, PS21_ZPASS_ARGS);
        else
            PS21_LAUNCH(k_general_zpass<false>, grid, dim3(256), fft_smem_bytes(sp.nz, sp.g1_cols), st, PS21_ZPASS_ARGS);
        PS21_CUDA(cudaGetLastError());
    }
    {
        ProfScope ps(p, KK_PASS_MID, st);
        const long long ntile = ((long long)n0 * sp.nh + sp.g2_tile - 1) / sp.g2_tile;
        dim3 grid((unsigned)((ntile + kYPassItems - 1) / kYPassItems), (unsigned)nitems);
#define PS21_YPASS_ARGS p->d_scratch, spec_elems, n0, n1, sp.nh, sp.fy, sp.tw_y, ilog2(sp.g2_tile)
        if (has_radix32(sp.fy))
            PS21_LAUNCH(k_general_ypass<true>, grid, dim3(256), fft_smem_bytes(n1, sp.g2_tile) * 3 / 2, st,
                        PS21_YPASS_ARGS);
        else
            PS21_LAUNCH(k_general_ypass<false>, grid, dim3(256), fft_smem_bytes(n1, sp.g2_tile) * 3 / 2, st,
                        PS21_YPASS_ARGS);
        PS21_CUDA(cudaGetLastError());
    }
    {
        ProfScope ps(p, KK_PASS_BIN, st);
        const int nb_out = p->chunks[sp.chunks[0]].nbins_out;
        dim3 grid(sp.g3_slots, (unsigned)nitems);
        size_t smem = fft_smem_bytes(n0, sp.g3_tile) + general_xpass_extra(nb_out, n0, sp.g3_tile);
#define PS21_XPASS_ARGS                                                                                          \
    p->d_scratch, spec_elems, n0, n1, sp.nh, sp.nz, sp.fx, sp.tw_x, ilog2(sp.g3_tile), im, p->d_chunks,           \
        (const uint16_t* const*)p->d_lut_table, general_lut_block(n0, sp.g3_tile), p->d_partials, p->out_per_box, \
        p->slots_max, p->d_dc, p->nchunks, general_bin_copies(nb_out)
        if (has_radix32(sp.fx))
            PS21_LAUNCH(k_general_xpass_bin<true>, grid, dim3(256), smem, st, PS21_XPASS_ARGS);
        else
            PS21_LAUNCH(k_general_xpass_bin<false>, grid, dim3(256), smem, st, PS21_XPASS_ARGS);
        PS21_CUDA(cudaGetLastError());
    }
    p->launches_last += 3;
    return PS21_OK;
}

}  // namespace ps21
