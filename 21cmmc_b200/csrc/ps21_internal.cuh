// Internal declarations shared by the translation units of libps21.so.
#pragma once

#ifdef PS21_EMU
#include "cuda_emu.h"  // tests/emu: host emulation of the CUDA subset used by the general path (tests only)
#else
#include <cuda_runtime.h>
#endif

#include <cstdint>
#include <string>
#include <vector>

#include "fft_smem.cuh"

#ifndef PS21_EMU
// dynamic shared memory of a kernel, typed view
#define PS21_DYN_SMEM(type, name)                                         \
    extern __shared__ __align__(128) unsigned char ps21_dyn_smem_raw[];   \
    type* name = reinterpret_cast<type*>(ps21_dyn_smem_raw)
#endif

#ifdef PS21_EMU
#define PS21_EMU_OR(emu, real) (emu)
#else
#define PS21_EMU_OR(emu, real) (real)
#endif

#ifdef PS21_EMU
#define PS21_LAUNCH(kernel, grid, block, smem, stream, ...) \
    ps21emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#else
#define PS21_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

namespace ps21 {

constexpr int kMaxBinsSmem = 8192;     // fp64 bins held in shared memory by the binning kernels
constexpr uint16_t kDiscard = 0xFFFF;  // LUT value: mode carries zero weight or lies outside the edges

// One chunk (= one spectrum) of a box.
struct ChunkDev {
    int start;            // first slice on axis 2
    int nz;               // slices
    int nh;               // nz/2 + 1
    int shape;            // index into the plan's shape table
    int nbins_out;        // nbins (spherical) or nbins*nh (cylindrical)
    int dc_bin;           // output bin that holds the k=0 mode, or -1
    long long out_off;    // offset of this chunk in a box's power record
    int nslots;           // partial-sum slots written per spectrum by the binning kernel of its path
    double norm;          // V / N^2
    double ncells;        // N = n0*n1*nz
};

// Per distinct chunk length.
struct ShapePlan {
    int nz = 0, nh = 0;
    int path = 0;  // PS21_PATH_GENERAL / PS21_PATH_FUSED
    FftFactors fz{}, fy{}, fx{};
    float2* tw_z = nullptr;  // device twiddle tables exp(-2 pi i q / n)
    float2* tw_y = nullptr;
    float2* tw_x = nullptr;
    std::vector<int> chunks;      // chunk ids with this length
    int* d_chunks = nullptr;      // same on device
    // general path tiling
    int g1_cols = 0;   // complex columns (= row pairs) per CTA in the z pass
    int g2_tile = 0;   // k_z columns per CTA in the y pass
    int g3_tile = 0;   // (k_y,k_z) columns per CTA in the x pass
    int g3_slots = 0;  // CTAs per spectrum in the x pass (= partial-sum slots)
    // fused path (ps21_fused.cu)
    void* fused = nullptr;   // FusedShape (path FUSED) or G2Shape (path TWOPASS)
};

struct DataTables {
    int nsets = 0;
    long long cap = 0;           // doubles allocated per table
    long long gen = 0;           // bumped by every ps21_set_data (captured graphs hold the old values)
    double* d_data = nullptr;    // [nsets][out_per_box]
    double* d_noise2 = nullptr;  // [nsets][out_per_box] noise^2
    uint8_t* d_mask = nullptr;   // [nsets][out_per_box]
    double model_uncertainty = 0.15;
    int error_on_model = 1;
};

}  // namespace ps21

struct ps21_plan {
    int n0 = 0, n1 = 0, n2_total = 0, nchunks = 0, nbins = 0, mode = 0, device = 0, path_req = 0;
    int ig_kperp = 0, ig_kpar = 0, ig_k0 = 0;
    long long out_per_box = 0;
    long long alg_bytes_per_box = 0;
    long long workspace_cap = 0;
    int slots_max = 0;
    std::vector<ps21::ChunkDev> chunks;
    std::vector<ps21::ShapePlan> shapes;
    ps21::ChunkDev* d_chunks = nullptr;
    // per chunk LUT: bin index of every half-spectrum mode, layout defined by the path of its shape
    std::vector<uint16_t*> d_lut;  // [nchunks]
    std::vector<char> lut_owned;   // chunks with identical geometry share one table
    uint16_t** d_lut_table = nullptr;  // the same pointers in device memory
    std::vector<long long> h_counts;
    long long* d_counts = nullptr;   // [out_per_box]
    double* d_kfac = nullptr;        // [out_per_box] k^3 / (2 pi^2)  (spherical only)
    // workspace
    float2* d_scratch = nullptr;
    long long scratch_elems = 0;
    long long spec_max = 0;
    double* d_partials = nullptr;    // [batch][slots_max][out_per_box]
    double* d_mean_parts = nullptr;  // [batch][nchunks][kMeanParts] partial sums of the offset estimate
    float2* d_dc = nullptr;          // [batch][nchunks] DC Fourier coefficient of the mean-subtracted chunk
    long long batch_cap = 0;
    ps21::DataTables data;
    long long launches_last = 0;
    // optional per-kernel timing with CUDA events on the launching stream (ps21_plan_profile)
    bool profiling = false;
    struct ProfRec { int kind; void* e0; void* e1; };
    std::vector<ProfRec> prof;
    // host staging (ps21_*_host)
    float* d_stage[2] = {nullptr, nullptr};
    long long stage_boxes = 0;
    void* stage_stream[2] = {nullptr, nullptr};
    void* stage_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // copy done [2], compute done [2]
    double* d_res[3] = {nullptr, nullptr, nullptr};             // power, Delta^2, lnL of a *_host call
    // pageable host inputs: page-locked ring the staging threads copy through (2 slots per thread) + "slot free" events
    static constexpr int kStageThreads = 16;   // at most; PS21_STAGE_THREADS or half of the hardware threads are used
    static constexpr long long kStageSlotBytes = 16LL << 20;
    char* h_ring = nullptr;
    void* ring_ev[2 * kStageThreads] = {};
    long long res_boxes = 0;
    // CUDA graphs of whole calls (ps21_plan_set_graphs)
    struct GraphKey {
        const float* boxes; long long batch, stride; double *power, *delta, *lnl; long long box0, data_gen;
    };
    struct GraphEntry { GraphKey key; void* exec; long long launches; unsigned long long last_use; };
    std::vector<GraphEntry> graphs;
    int graph_mode = 0;
    int graph_miss_streak = 0;
    unsigned long long graph_clock = 0;
    long long graph_captures = 0, graph_replays = 0;
    void* graph_stream = nullptr;
    void* graph_ev[2] = {nullptr, nullptr};
};

namespace ps21 {
enum KernelKind { KK_MEAN = 0, KK_PASS1 = 1, KK_PASS_MID = 2, KK_PASS_BIN = 3, KK_FINALIZE = 4, KK_COUNT = 5 };
// RAII: records an event pair around the launches issued in its scope when profiling is on
struct ProfScope {
    ps21_plan* p; cudaStream_t st; int idx = -1;
    ProfScope(ps21_plan* plan, int kind, cudaStream_t s);
    ~ProfScope();
};
void set_error(const std::string& msg);
void drop_graphs(ps21_plan* p);   // captured launches hold raw pointers: dropped whenever a buffer they name is reallocated
int cuda_fail(cudaError_t e, const char* what);
#define PS21_CUDA(call)                                        \
    do {                                                       \
        cudaError_t _e = (call);                               \
        if (_e != cudaSuccess) return ps21::cuda_fail(_e, #call); \
    } while (0)

// general path (ps21_general.cu)
int general_build_lut(ps21_plan* p, int chunk, const double* d_k0sq, const double* d_k1sq, const double* d_k2sq,
                      const double* d_edges, cudaStream_t st);
// cp.async (LDGSTS): 8- and 16-byte asynchronous global -> shared copies, commit groups
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
#ifdef PS21_EMU
    *reinterpret_cast<float2*>(smem_dst) = *reinterpret_cast<const float2*>(gsrc);
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc)
                 : "memory");
#endif
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
#ifdef PS21_EMU
    *reinterpret_cast<float*>(smem_dst) = *reinterpret_cast<const float*>(gsrc);
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc)
                 : "memory");
#endif
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
#ifdef PS21_EMU
    *reinterpret_cast<float4*>(smem_dst) = *reinterpret_cast<const float4*>(gsrc);
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc)
                 : "memory");
#endif
}
// the same with a 32-bit shared-window address (ptx_sm100.cuh: smem_u32)
__device__ __forceinline__ void cp_async8_s32(unsigned smem_addr, const void* gsrc) {
#ifdef PS21_EMU
    *reinterpret_cast<float2*>(ps21emu::g_dyn_smem + smem_addr) = *reinterpret_cast<const float2*>(gsrc);
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async4_s32(unsigned smem_addr, const void* gsrc) {
#ifdef PS21_EMU
    *reinterpret_cast<float*>(ps21emu::g_dyn_smem + smem_addr) = *reinterpret_cast<const float*>(gsrc);
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef PS21_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#ifndef PS21_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

int general_init();
size_t general_xpass_extra(int nb_out, int n0, int tile);
long long general_lut_block(int n0, int tile);
long long general_lut_elems(int n0, int n1, int nh, int tile);
int launch_slab_mean(ps21_plan* p, const float* d_boxes, long long batch_stride, long long batch, cudaStream_t st);
int general_forward(ps21_plan* p, ShapePlan& sp, const float* d_boxes, long long batch_stride, long long item0,
                    long long nitems, cudaStream_t st);

// fused path (ps21_fused.cu); returns 0 and sets sp.fused when the shape is supported
int fused_try_setup(ps21_plan* p, ShapePlan& sp);
void fused_destroy(ShapePlan& sp);
int fused_build_lut(ps21_plan* p, int chunk, const double* d_k0sq, const double* d_k1sq, const double* d_k2sq,
                    const double* d_edges, cudaStream_t st);
int fused_forward(ps21_plan* p, ShapePlan& sp, const float* d_boxes, long long batch_stride, long long item0,
                  long long nitems, cudaStream_t st);
long long fused_scratch_elems(const ShapePlan& sp, int n0, int n1);
// allocations fused_forward would need for `nitems` spectra per launch (none may happen while a graph is captured)
int fused_reserve(ps21_plan* p, ShapePlan& sp, long long nitems);

// generic two-pass path (ps21_g2.cu); returns 0 and sets sp.fused when the shape is supported
int g2_try_setup(ps21_plan* p, ShapePlan& sp);
void g2_destroy(ShapePlan& sp);
int g2_build_lut(ps21_plan* p, int chunk, const double* d_k0sq, const double* d_k1sq, const double* d_k2sq,
                 const double* d_edges, cudaStream_t st);
int g2_forward(ps21_plan* p, ShapePlan& sp, const float* d_boxes, long long batch_stride, long long item0,
               long long nitems, cudaStream_t st);
long long g2_scratch_elems(const ShapePlan& sp);

// exact n / d for 0 <= n < 65536, 1 <= d < 65536: q = (n * mul) >> shift with mul = ceil(2^(16+L) / d), L = ceil(log2 d)
inline FastDiv make_fastdiv(int d) {
    int L = 0;
    while ((1 << L) < d) ++L;
    FastDiv f;
    f.shift = 16 + L;
    f.mul = (unsigned)(((1ull << f.shift) + d - 1) / d);
    return f;
}

// item -> (box, chunk) for a shape whose chunk-id list is `ids[0..nper)`
struct ItemMap {
    const int* ids;
    int nper;
    long long item0;
    // L2 prefetch (two-pass kernels): bit 0 = the CTA's own next unit / tile but one, bit 1 = the first unit / tile of
    // the CTA that will take this CTA's place (`pf_ahead` CTAs later in launch order).  Set per launch by the host.
    int pf = 0;
    int pf_ahead = 0;
    int pf_where = 0;   // generic pass 1: 0 = at the top of a unit, 1 = after the z stages, 2 = before the last y stage
};
int l2_prefetch_mode();              // PS21_L2_PREFETCH (default in ps21_plan.cu)
int prefetch_ahead(int resident);    // launch-order distance to the CTA that takes a finishing CTA's place
int sm_count(int device);
__device__ __forceinline__ void item_to_box_chunk(const ItemMap& m, long long local, long long& box, int& chunk) {
    const long long it = m.item0 + local;
    box = it / m.nper;
    chunk = m.ids[it % m.nper];
}

// Offset estimate of a chunk (mean of its x = 0 slab): k_slab_mean_part leaves kMeanParts partial sums per chunk;
// every consumer adds them in the same fixed order, so all CTAs of a spectrum subtract the same constant.
constexpr int kMeanParts = 8;
__device__ __forceinline__ float chunk_offset(const double* __restrict__ parts, long long item, int n1, int nz) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < kMeanParts; ++q) s += parts[item * kMeanParts + q];
    return (float)(s / ((double)n1 * (double)nz));
}

// exact fp64 digitisation shared by both LUT builders: np.digitize(sqrt((a+b)+c), edges) - 1, or -1
__device__ __forceinline__ int digitize_mode(double a, double b, double c, const double* __restrict__ edges,
                                             int nbins) {
    const double kmag = __dsqrt_rn(__dadd_rn(__dadd_rn(a, b), c));
    // number of edges <= kmag (upper bound), edges ascending
    int lo = 0, hi = nbins + 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (edges[mid] <= kmag) lo = mid + 1; else hi = mid;
    }
    return (lo >= 1 && lo <= nbins) ? lo - 1 : -1;
}
}  // namespace ps21
