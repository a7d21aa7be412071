// Plan management, finalize / likelihood kernels and the extern "C" entry points of include/ps21.h.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <thread>

#include "ps21_internal.cuh"
#include "../../include/ps21.h"

namespace ps21 {

// L2 prefetch in the two-pass kernels (see ItemMap): PS21_L2_PREFETCH = 0 off, 1 own next unit / tile, 3 also the first
// unit / tile of the succeeding CTA; unset (-1) = the measured default of each kernel: on (1) only in generic pass-1
// kernels whose units have the SM to themselves (512^3: 18.2 -> 16.6 ms) -- with several resident CTAs the kernels are
// bound by DRAM throughput, not latency, and the extra requests cost 1.5 % (256^3 fused), 3 % (200^3) of pass 1;
// pass 2 does not react at all (profiles/r02_g2_experiments.md).  PS21_PF_AHEAD overrides the launch-order distance to
// the succeeding CTA.
int l2_prefetch_mode() {
    static const int mode = [] {
        const char* e = std::getenv("PS21_L2_PREFETCH");
        return e ? std::atoi(e) : -1;
    }();
    return mode;
}
int prefetch_ahead(int resident) {
    static const int forced = [] {
        const char* e = std::getenv("PS21_PF_AHEAD");
        return e ? std::atoi(e) : 0;
    }();
    return forced > 0 ? forced : resident;
}
int sm_count(int device) {
#ifdef PS21_EMU
    (void)device;
    return 148;
#else
    static int cached[64] = {0};
    if (device < 0 || device >= 64) return 148;
    if (!cached[device]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        cached[device] = v;
    }
    return cached[device];
#endif
}

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return PS21_ERR_CUDA;
}
ProfScope::ProfScope(ps21_plan* plan, int kind, cudaStream_t s) : p(plan), st(s) {
#ifndef PS21_EMU
    if (!p->profiling) return;
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    cudaEventRecord(a, st);
    idx = (int)p->prof.size();
    p->prof.push_back({kind, (void*)a, (void*)b});
#endif
}
ProfScope::~ProfScope() {
#ifndef PS21_EMU
    if (idx >= 0) cudaEventRecord((cudaEvent_t)p->prof[idx].e1, st);
#endif
}

static int invalid(const std::string& msg) {
    g_err = msg;
    return PS21_ERR_INVALID;
}

static FftFactors factorize(int n, bool allow32) {
    FftFactors f{};
    f.n = n;
    int m = n, ns = 0;
    auto take = [&](int r) {
        while (m % r == 0 && ns < 14) {
            f.radix[ns++] = r;
            m /= r;
        }
    };
    // powers of two: as few stages as radix 16 allows, bits spread evenly (512 = 8*8*8, not 16*16*2: the same
    // number of stages and barriers, but every stage keeps all threads of a small tile busy)
    int k = 0;
    while (m % 2 == 0) {
        m /= 2;
        ++k;
    }
    // 512 = 32*16, 1024 = 32*32: two stages instead of three.  Measured on 512^3: y pass 17.0 -> 14.9 ms, x pass
    // 19.0 -> 16.2 ms, but the z pass 11.7 -> 13.9 ms (its untangle epilogue pays for the registers) -> y and x only
    if (allow32 && (k == 9 || k == 10)) {
        f.radix[ns++] = 32;
        f.radix[ns++] = 1 << (k - 5);
    } else if (k > 0) {
        const int st = (k + 3) / 4, base = k / st, extra = k % st;
        for (int i = 0; i < st && ns < 14; ++i) f.radix[ns++] = 1 << (base + (i < extra ? 1 : 0));
    }
    // 200 = 8*25, 100 = 4*25, 250 = 2*25*5: one 25-point stage (y, x; in the z pass it was measured slower like
    // the 32-point stage: 200^3 4.6 -> 6.9 ms)
    if (allow32 && m % 25 == 0 && ns < 14) {
        f.radix[ns++] = 25;
        m /= 25;
    }
    take(3);
    take(5);
    for (int pr = 7; m > 1 && ns < 14; pr += 2) take(pr);
    f.nstages = ns;
    if (m != 1 || n >= 65536) f.nstages = -1;
    int Ns = 1;
    for (int s = 0; s < ns; ++s) {
        f.div_ns[s] = make_fastdiv(Ns);
        f.div_nb[s] = make_fastdiv(n / f.radix[s]);
        Ns *= f.radix[s];
    }
    return f;
}

static int upload_twiddles(int n, float2** out) {
    std::vector<float2> tw(n);
    for (int q = 0; q < n; ++q) {
        const double a = -2.0 * M_PI * (double)q / (double)n;
        tw[q] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    PS21_CUDA(cudaMalloc((void**)out, sizeof(float2) * n));
    PS21_CUDA(cudaMemcpy(*out, tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    return PS21_OK;
}

template <class Extra>
static int pick_tile(int n, Extra extra) {
    const int cand[5] = {16, 8, 4, 2, 1};   // 32-column tiles measured slower (z pass of 200^3: 4.5 -> 6.8 ms)
    // first choice: the widest tile that leaves two CTAs per SM ((228 KB - 2 x 1 KB reserved) / 2)
    for (size_t budget : {(size_t)115712, (size_t)220 * 1024})
        for (int t : cand)
            if (2 * (size_t)n * (t | 1) * sizeof(float2) + extra(t) <= budget) return t;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// finalize: partial sums -> power -> Delta^2 -> chi^2   (likelihood.py:863/:652 and :584-604)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double chi2_term(double delta, double data, double noise2, double mu, int err_on_model) {
    const double mod = mu * (err_on_model ? delta : data);
    const double d = delta - data;
    return d * d / (mod * mod + noise2);
}

constexpr int kFinalizeThreads = 512;   // 16 warps: one output per warp and round (80 outputs per config-3 box)
__global__ void __launch_bounds__(kFinalizeThreads)
k_finalize(const double* __restrict__ partials, int slots_max, long long out_per_box,
           const ChunkDev* __restrict__ chunks, int nchunks, const long long* __restrict__ counts,
           const double* __restrict__ kfac, const double* __restrict__ mean_parts, int n1,
           const float2* __restrict__ dc, double* __restrict__ power, double* __restrict__ delta,
           double* __restrict__ lnl, const double* __restrict__ data, const double* __restrict__ noise2,
           const uint8_t* __restrict__ mask, int nsets, double mu, int err_on_model, long long box0) {
    __shared__ double red[kFinalizeThreads];
    const long long box = blockIdx.x;
    const long long set = (box0 + box) % nsets;  // box0: index of this call's first box in the caller's batch
    double chi = 0.0;
    // Many partial-sum slots (fused path: 43-129 per spectrum): one warp per output, lanes over the slots, fixed-order
    // shuffle tree -- a thread walking 86 slots alone made this kernel 27 us of pure load latency.  Few slots
    // (cylindrical records have one): one thread per output.  gridDim.y > 1 only without lnL.
    const bool by_warp = slots_max >= 8;
    const int lane = threadIdx.x & 31;
    const long long first = by_warp ? (long long)blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5)
                                    : (long long)blockIdx.y * blockDim.x + threadIdx.x;
    const long long stride = by_warp ? (long long)gridDim.y * (blockDim.x >> 5) : (long long)gridDim.y * blockDim.x;
    for (long long i = first; i < out_per_box; i += stride) {
        int c = 0;
        while (c + 1 < nchunks && chunks[c + 1].out_off <= i) ++c;
        const ChunkDev ch = chunks[c];
        double s = 0.0;
        if (by_warp) {
            for (int sl = lane; sl < ch.nslots; sl += 32) s += partials[(box * slots_max + sl) * out_per_box + i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane != 0) continue;
        } else {
            for (int sl = 0; sl < ch.nslots; ++sl) s += partials[(box * slots_max + sl) * out_per_box + i];
        }
        if (ch.dc_bin >= 0 && ch.out_off + ch.dc_bin == i) {
            // restore the k=0 power removed by the offset subtraction: |F'(0) + N c|^2 - |F'(0)|^2
            const float2 f0 = dc[box * nchunks + c];
            const double shift = ch.ncells * (double)chunk_offset(mean_parts, box * nchunks + c, n1, ch.nz);
            const double full = (double)f0.x + shift;
            s += full * full - (double)f0.x * (double)f0.x;
        }
        const double pw = s / (double)counts[i] * ch.norm;  // 0/0 -> NaN for empty bins, as the reference
        if (power) power[box * out_per_box + i] = pw;
        if (kfac) {
            const double d = pw * kfac[i];
            if (delta) delta[box * out_per_box + i] = d;
            if (lnl) {   // likelihood.py:601-604; a NaN bin inside the k mask propagates, as in the reference
                const long long j = set * out_per_box + i;
                if (mask[j]) chi += chi2_term(d, data[j], noise2[j], mu, err_on_model);
            }
        }
    }
    if (lnl) {   // fixed-order tree sum over the block (gridDim.y == 1): bitwise reproducible
        red[threadIdx.x] = chi;
        __syncthreads();
        for (int s = kFinalizeThreads / 2; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) lnl[box] = -0.5 * red[0];
    }
}

__global__ void k_lnl_from_delta(const double* __restrict__ delta, long long out_per_box,
                                 const ChunkDev* __restrict__ chunks, int nchunks, double* __restrict__ lnl,
                                 const double* __restrict__ data, const double* __restrict__ noise2,
                                 const uint8_t* __restrict__ mask, int nsets, double mu, int err_on_model,
                                 long long batch) {
    const long long box = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (box >= batch) return;
    const long long set = box % nsets;
    double total = 0.0;
    for (int c = 0; c < nchunks; ++c) {
        const ChunkDev ch = chunks[c];
        double cs = 0.0;
        for (int b = 0; b < ch.nbins_out; ++b) {
            const long long i = ch.out_off + b;
            const long long j = set * out_per_box + i;
            if (mask[j]) cs += chi2_term(delta[box * out_per_box + i], data[j], noise2[j], mu, err_on_model);
        }
        total += -0.5 * cs;
    }
    lnl[box] = total;
}

__global__ void k_chi2_rows(const double* __restrict__ delta, long long n, long long nrec,
                            const double* __restrict__ data, const double* __restrict__ noise,
                            const uint8_t* __restrict__ mask, long long nseg, const long long* __restrict__ seg_len,
                            double mu, int err_on_model, double* __restrict__ lnl) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrec) return;
    double total = 0.0;
    long long i = 0;
    for (long long s = 0; s < nseg; ++s) {
        double cs = 0.0;
        for (long long e = i + seg_len[s]; i < e; ++i)
            if (mask[i]) {
                const double nz = noise ? noise[i] : 0.0;
                cs += chi2_term(delta[r * n + i], data[i], nz * nz, mu, err_on_model);
            }
        total += -0.5 * cs;
    }
    lnl[r] = total;
}

// one CTA per model, one thread per window-function row k (tens of rows of tens of values each); the per-row terms are
// summed by a fixed-order tree, so lnL is bitwise reproducible
// gauss = 0: one-sided erf likelihood of upper limits (likelihood.py:2324-2328);
// gauss = 1: Gaussian -0.5 (after - value)^2 / err^2 of measured band powers (likelihood.py:569-583)
__global__ void __launch_bounds__(128)
k_upper_limit(const double* __restrict__ model, long long nrec, int nk, int nkwf,
              const double* __restrict__ wf, const double* __restrict__ limit,
              const double* __restrict__ var, const double* __restrict__ model_err, double frac,
              int gauss, double* __restrict__ lnl) {
    __shared__ double red[128];
    const long long r = blockIdx.x;
    const double* m = model + r * nkwf;
    double total = 0.0;
    for (int k = threadIdx.x; k < nk; k += blockDim.x) {
        double after = 0.0, up = 0.0, lo = 0.0;
        for (int j = 0; j < nkwf; ++j) {
            const double w = wf[(long long)k * nkwf + j];
            after += w * m[j];
            if (model_err) {
                up += w * (m[j] + model_err[j]);
                lo += w * (m[j] - model_err[j]);
            }
        }
        double e2 = var[k] + (frac * after) * (frac * after);
        if (model_err) {
            const double me = 0.5 * ((up - after) + (after - lo));
            e2 += me * me;
        }
        if (gauss) {
            const double d = after - limit[k];
            total += -0.5 * (d * d) / e2;   // np.sum: NaN propagates
            continue;
        }
        double like = 0.5 + 0.5 * erf((limit[k] - after) / (sqrt(2.0) * sqrt(e2)));
        if (like <= 0.0) like = 1e-50;
        const double l = log(like);
        if (l == l) total += l;  // nansum
    }
    red[threadIdx.x] = total;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) lnl[r] += red[0];
}

static void free_dev(void* p) {
    if (p) cudaFree(p);
}

// temporary device buffer of an entry point: released on every return path
template <class T>
struct DevTmp {
    T* p = nullptr;
    ~DevTmp() { free_dev(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc((void**)&p, sizeof(T) * (n ? n : 1)); }
};

void drop_graphs(ps21_plan* p) {
#ifndef PS21_EMU
    for (auto& g : p->graphs)
        if (g.exec) cudaGraphExecDestroy((cudaGraphExec_t)g.exec);
#endif
    p->graphs.clear();
}

static int ensure_batch(ps21_plan* p, long long batch) {
    if (batch <= p->batch_cap) return PS21_OK;
    drop_graphs(p);   // captured launches hold the old pointers
    free_dev(p->d_partials);
    free_dev(p->d_mean_parts);
    free_dev(p->d_dc);
    p->d_partials = nullptr;
    p->d_mean_parts = nullptr;
    p->d_dc = nullptr;
    p->batch_cap = 0;
    PS21_CUDA(cudaMalloc((void**)&p->d_partials, sizeof(double) * batch * p->slots_max * p->out_per_box));
    PS21_CUDA(cudaMalloc((void**)&p->d_mean_parts, sizeof(double) * batch * p->nchunks * kMeanParts));
    PS21_CUDA(cudaMalloc((void**)&p->d_dc, sizeof(float2) * batch * p->nchunks));
    PS21_CUDA(cudaMemset(p->d_dc, 0, sizeof(float2) * batch * p->nchunks));
    p->batch_cap = batch;
    return PS21_OK;
}

static int ensure_scratch(ps21_plan* p, long long want_elems) {
    if (p->scratch_elems >= want_elems) return PS21_OK;
    drop_graphs(p);
    free_dev(p->d_scratch);
    p->d_scratch = nullptr;
    p->scratch_elems = 0;
    if (cudaMalloc((void**)&p->d_scratch, sizeof(float2) * want_elems) != cudaSuccess) {
        cudaGetLastError();
        set_error("cannot allocate half-spectrum scratch");
        return PS21_ERR_NOMEM;
    }
    p->scratch_elems = want_elems;
    return PS21_OK;
}

static long long spectrum_elems(const ps21_plan* p, const ShapePlan& sp) {
    if (sp.path == PS21_PATH_FUSED) return fused_scratch_elems(sp, p->n0, p->n1);
    if (sp.path == PS21_PATH_TWOPASS) return g2_scratch_elems(sp);
    return (long long)p->n0 * p->n1 * sp.nh;
}

static int forward_any(ps21_plan* p, ShapePlan& sp, const float* d_boxes, long long batch_stride, long long i0,
                       long long n, cudaStream_t st) {
    if (sp.path == PS21_PATH_FUSED) return fused_forward(p, sp, d_boxes, batch_stride, i0, n, st);
    if (sp.path == PS21_PATH_TWOPASS) return g2_forward(p, sp, d_boxes, batch_stride, i0, n, st);
    return general_forward(p, sp, d_boxes, batch_stride, i0, n, st);
}

// spectra of shape `sp` per scratch slice for a batch of `batch` boxes (equal slices: 256 spectra with room for
// 121 run as 86 + 85 + 85, not 121 + 121 + 14)
static long long slice_items(const ps21_plan* p, const ShapePlan& sp, long long batch, long long* spec_out) {
    const long long total = batch * (long long)sp.chunks.size();
    const long long spec = spectrum_elems(p, sp);
    long long per = p->workspace_cap / (spec * (long long)sizeof(float2));
    if (per < 1) per = 1;
    if (per > 32768) per = 32768;
    if (per > total) per = total;
    const long long nslices = (total + per - 1) / per;
    per = (total + nslices - 1) / nslices;
    if (spec_out) *spec_out = spec;
    return per;
}

// Every kernel of one call, in order, on `st` (directly, or while `st` is being captured into a CUDA graph).
static int issue_launches(ps21_plan* p, const float* d_boxes, long long batch, long long batch_stride,
                          double* d_power, double* d_delta, double* d_lnl, cudaStream_t st, long long box0) {
    int rc;
    {
        ProfScope ps(p, KK_MEAN, st);
        rc = launch_slab_mean(p, d_boxes, batch_stride, batch, st);
    }
    if (rc) return rc;
    for (auto& sp : p->shapes) {
        const long long total = batch * (long long)sp.chunks.size();
        const long long per = slice_items(p, sp, batch, nullptr);
        for (long long i0 = 0; i0 < total; i0 += per) {
            const long long n = (total - i0 < per) ? (total - i0) : per;
            rc = forward_any(p, sp, d_boxes, batch_stride, i0, n, st);
            if (rc) return rc;
        }
    }
    const bool want_lnl = d_lnl != nullptr;
    ProfScope psf(p, KK_FINALIZE, st);
    unsigned fparts = want_lnl ? 1u : (unsigned)((p->out_per_box + 1023) / 1024);
    if (fparts > 64u) fparts = 64u;
    PS21_LAUNCH(k_finalize, dim3((unsigned)batch, fparts), dim3(kFinalizeThreads), 0, st, p->d_partials, p->slots_max,
                p->out_per_box, p->d_chunks, p->nchunks, p->d_counts,
                p->mode == PS21_MODE_SPHERICAL ? p->d_kfac : nullptr, p->d_mean_parts, p->n1, p->d_dc, d_power,
                d_delta, d_lnl, p->data.d_data, p->data.d_noise2, p->data.d_mask,
                p->data.nsets > 0 ? p->data.nsets : 1, p->data.model_uncertainty, p->data.error_on_model, box0);
    PS21_CUDA(cudaGetLastError());
    p->launches_last++;
    return PS21_OK;
}

// All kernels of one call, asynchronous on `st`.  With graphs enabled (ps21_plan_set_graphs) the launch sequence
// of a call is captured once per distinct argument set and replayed with one cudaGraphLaunch afterwards: an MCMC
// step calls with the same device buffers every iteration, and at 8 walkers per GPU (config 3 on 8 GPUs) the whole
// step is ~1.4 ms, where launch gaps count.
static int run_device(ps21_plan* p, const float* d_boxes, long long batch, long long batch_stride, double* d_power,
                      double* d_delta, double* d_lnl, cudaStream_t st, long long box0 = 0, bool allow_graph = true) {
    if (batch <= 0) return PS21_OK;
    if (batch_stride < (long long)p->n0 * p->n1 * p->n2_total)
        return invalid("batch_stride smaller than one box");
    if (d_lnl && p->data.nsets == 0) {
        set_error("ps21_set_data must be called before a likelihood evaluation");
        return PS21_ERR_STATE;
    }
    if ((d_lnl || d_delta) && p->mode != PS21_MODE_SPHERICAL)
        return invalid("Delta^2 / lnL are defined for the spherical mode only");
    int rc = ensure_batch(p, batch);
    if (rc) return rc;
    {   // the scratch serves the shapes one after the other: size it for the largest slice up front
        long long want = 0;
        for (auto& sp : p->shapes) {
            long long spec = 0;
            const long long per = slice_items(p, sp, batch, &spec);
            if (per * spec > want) want = per * spec;
            if (sp.path == PS21_PATH_FUSED && (rc = fused_reserve(p, sp, per))) return rc;
        }
        rc = ensure_scratch(p, want);
        if (rc) return rc;
    }
#ifndef PS21_EMU
    if (allow_graph && p->graph_mode && !p->profiling) {
        const ps21_plan::GraphKey key{d_boxes, batch, batch_stride, d_power, d_delta, d_lnl, box0, p->data.gen};
        ps21_plan::GraphEntry* hit = nullptr;
        for (auto& g : p->graphs)
            if (std::memcmp(&g.key, &key, sizeof(key)) == 0) hit = &g;
        if (!p->graph_stream) {
            cudaStream_t s;
            PS21_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
            p->graph_stream = (void*)s;
            cudaEvent_t e0, e1;
            PS21_CUDA(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
            PS21_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            p->graph_ev[0] = (void*)e0;
            p->graph_ev[1] = (void*)e1;
        }
        cudaStream_t gs = (cudaStream_t)p->graph_stream;
        if (!hit && p->graph_miss_streak < 16) {
            // capture on the plan's own stream (the caller's may be the legacy default stream, which cannot be captured)
            if (p->graphs.size() >= 16) {   // evict the least recently used entry
                size_t lru = 0;
                for (size_t i = 1; i < p->graphs.size(); ++i)
                    if (p->graphs[i].last_use < p->graphs[lru].last_use) lru = i;
                if (p->graphs[lru].exec) cudaGraphExecDestroy((cudaGraphExec_t)p->graphs[lru].exec);
                p->graphs.erase(p->graphs.begin() + lru);
            }
            p->launches_last = 0;
            PS21_CUDA(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
            rc = issue_launches(p, d_boxes, batch, batch_stride, d_power, d_delta, d_lnl, gs, box0);
            cudaGraph_t graph = nullptr;
            cudaError_t ce = cudaStreamEndCapture(gs, &graph);
            if (rc) {
                if (graph) cudaGraphDestroy(graph);
                return rc;
            }
            if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture");
            cudaGraphExec_t exec = nullptr;
            ce = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return cuda_fail(ce, "cudaGraphInstantiate");
            p->graphs.push_back({key, (void*)exec, p->launches_last, 0});
            hit = &p->graphs.back();
            p->graph_captures++;
            p->graph_miss_streak++;
        } else if (hit) {
            p->graph_miss_streak = 0;
        }
        if (hit) {
            hit->last_use = ++p->graph_clock;
            p->launches_last = hit->launches;
            // fork from / join into the caller's stream
            PS21_CUDA(cudaEventRecord((cudaEvent_t)p->graph_ev[0], st));
            PS21_CUDA(cudaStreamWaitEvent(gs, (cudaEvent_t)p->graph_ev[0], 0));
            PS21_CUDA(cudaGraphLaunch((cudaGraphExec_t)hit->exec, gs));
            PS21_CUDA(cudaEventRecord((cudaEvent_t)p->graph_ev[1], gs));
            PS21_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)p->graph_ev[1], 0));
            p->graph_replays++;
            return PS21_OK;
        }
        // the caller changes buffers on every call: graphs would be re-captured each time -> direct launches
    }
#endif
    p->launches_last = 0;
    return issue_launches(p, d_boxes, batch, batch_stride, d_power, d_delta, d_lnl, st, box0);
}

// everything that regrows on demand: half-spectrum scratch, partial sums, staging and result buffers, graphs
static void release_workspace(ps21_plan* p) {
    drop_graphs(p);
    free_dev(p->d_scratch);
    p->d_scratch = nullptr;
    p->scratch_elems = 0;
    free_dev(p->d_partials);
    free_dev(p->d_mean_parts);
    free_dev(p->d_dc);
    p->d_partials = nullptr;
    p->d_mean_parts = nullptr;
    p->d_dc = nullptr;
    p->batch_cap = 0;
    for (int i = 0; i < 2; ++i) {
        free_dev(p->d_stage[i]);
        p->d_stage[i] = nullptr;
    }
    p->stage_boxes = 0;
    for (int i = 0; i < 3; ++i) {
        free_dev(p->d_res[i]);
        p->d_res[i] = nullptr;
    }
    p->res_boxes = 0;
#ifndef PS21_EMU
    if (p->h_ring) cudaFreeHost(p->h_ring);
    p->h_ring = nullptr;
    for (auto& e : p->ring_ev) {
        if (e) cudaEventDestroy((cudaEvent_t)e);
        e = nullptr;
    }
#endif
}

}  // namespace ps21

using namespace ps21;

extern "C" {

const char* ps21_last_error(void) { return g_err.c_str(); }
int ps21_version(void) { return 100; }

int ps21_plan_destroy(ps21_plan* p) {
    if (!p) return PS21_OK;
    cudaSetDevice(p->device);
    for (auto& sp : p->shapes) {
        free_dev(sp.tw_z);
        free_dev(sp.tw_y);
        free_dev(sp.tw_x);
        free_dev(sp.d_chunks);
        if (sp.path == PS21_PATH_TWOPASS) g2_destroy(sp);
        else fused_destroy(sp);
    }
    for (size_t i = 0; i < p->d_lut.size(); ++i)
        if (i < p->lut_owned.size() && p->lut_owned[i]) free_dev(p->d_lut[i]);
    free_dev(p->d_lut_table);
    free_dev(p->d_chunks);
    free_dev(p->d_counts);
    free_dev(p->d_kfac);
    release_workspace(p);
    free_dev(p->data.d_data);
    free_dev(p->data.d_noise2);
    free_dev(p->data.d_mask);
    for (int i = 0; i < 2; ++i)
        if (p->stage_stream[i]) cudaStreamDestroy((cudaStream_t)p->stage_stream[i]);
    for (int i = 0; i < 4; ++i)
        if (p->stage_ev[i]) cudaEventDestroy((cudaEvent_t)p->stage_ev[i]);
#ifndef PS21_EMU
    if (p->graph_stream) cudaStreamDestroy((cudaStream_t)p->graph_stream);
    for (int i = 0; i < 2; ++i)
        if (p->graph_ev[i]) cudaEventDestroy((cudaEvent_t)p->graph_ev[i]);
#endif
    delete p;
    return PS21_OK;
}

int ps21_plan_release_workspace(ps21_plan* p) {
    if (!p) return invalid("null argument");
    PS21_CUDA(cudaSetDevice(p->device));
    PS21_CUDA(cudaDeviceSynchronize());
    release_workspace(p);
    return PS21_OK;
}

int ps21_plan_set_graphs(ps21_plan* p, int32_t on) {
    if (!p) return invalid("null argument");
    p->graph_mode = on ? 1 : 0;
    p->graph_miss_streak = 0;
    if (!on) drop_graphs(p);
    return PS21_OK;
}

int ps21_plan_graph_stats(const ps21_plan* p, int64_t* captures, int64_t* replays) {
    if (!p || !captures || !replays) return invalid("null argument");
    *captures = p->graph_captures;
    *replays = p->graph_replays;
    return PS21_OK;
}

int ps21_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return invalid("null argument / bytes <= 0");
#ifdef PS21_EMU
    *ptr = std::malloc((size_t)bytes);
    return *ptr ? PS21_OK : PS21_ERR_NOMEM;
#else
    PS21_CUDA(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocPortable));
    return PS21_OK;
#endif
}

int ps21_host_free(void* ptr) {
#ifdef PS21_EMU
    std::free(ptr);
#else
    if (ptr) PS21_CUDA(cudaFreeHost(ptr));
#endif
    return PS21_OK;
}

int ps21_host_register(void* ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return invalid("null argument / bytes <= 0");
#ifndef PS21_EMU
    PS21_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
#endif
    return PS21_OK;
}

int ps21_host_unregister(void* ptr) {
#ifndef PS21_EMU
    if (ptr) PS21_CUDA(cudaHostUnregister(ptr));
#endif
    return PS21_OK;
}

int ps21_plan_create(const ps21_desc* d, ps21_plan** out) {
    if (!d || !out) return invalid("null argument");
    *out = nullptr;
    if (d->n0 < 1 || d->n1 < 1 || d->n2_total < 1 || d->nchunks < 1 || d->nbins < 1)
        return invalid("n0, n1, n2_total, nchunks and nbins must be positive");
    if (!d->chunk_start || !d->chunk_end || !d->volume || !d->k0sq || !d->k1sq || !d->k2sq || !d->edges ||
        !d->kcentres)
        return invalid("null array in descriptor");
    if (d->mode != PS21_MODE_SPHERICAL && d->mode != PS21_MODE_CYLINDRICAL) return invalid("unknown mode");
    if ((d->ignore_kperp_zero || d->ignore_k_zero) && d->n0 != d->n1)
        return invalid("ignore_kperp_zero / ignore_k_zero need n0 == n1 (the reference indexes both sky axes "
                       "with n0//2, likelihood.py:775-780)");
    PS21_CUDA(cudaSetDevice(d->device));
    {
        int irc = general_init();
        if (irc) return irc;
    }
    ps21_plan* p = new ps21_plan();
    struct Guard {  // destroys the half-built plan on every early return
        ps21_plan*& plan;
        ps21_plan** out;
        ~Guard() {
            if (*out == nullptr && plan) ps21_plan_destroy(plan);
        }
    } guard{p, out};
    p->n0 = d->n0; p->n1 = d->n1; p->n2_total = d->n2_total; p->nchunks = d->nchunks; p->nbins = d->nbins;
    p->mode = d->mode; p->device = d->device; p->path_req = d->path;
    p->ig_kperp = d->ignore_kperp_zero; p->ig_kpar = d->ignore_kpar_zero; p->ig_k0 = d->ignore_k_zero;
    if (d->workspace_bytes > 0) {
        p->workspace_cap = d->workspace_bytes;
    } else {  // default: a sixth of the device memory, at most 24 GiB (a 64-walker config-3 step needs 17.3 GB)
        p->workspace_cap = 8LL << 30;
#ifndef PS21_EMU
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b > 0) {
            long long cap = (long long)(total_b / 6);
            if (cap > (24LL << 30)) cap = 24LL << 30;
            if (cap < (1LL << 30)) cap = 1LL << 30;
            p->workspace_cap = cap;
        } else {
            cudaGetLastError();
        }
#endif
    }
    int rc = PS21_OK;
    auto fail = [&](int code) { return code; };  // the guard releases the plan
    // chunks and shapes
    std::map<int, int> shape_of_nz;
    long long off = 0, k2off = 0;
    std::vector<long long> k2offs;
    for (int c = 0; c < d->nchunks; ++c) {
        const int s = d->chunk_start[c], e = d->chunk_end[c];
        if (s < 0 || e <= s || e > d->n2_total) return fail(invalid("bad chunk bounds"));
        ChunkDev ch{};
        ch.start = s; ch.nz = e - s; ch.nh = ch.nz / 2 + 1;
        ch.nbins_out = d->mode == PS21_MODE_SPHERICAL ? d->nbins : d->nbins * ch.nh;
        if (ch.nbins_out >= (int)kDiscard || ch.nbins_out > kMaxBinsSmem)
            return fail(invalid("too many output bins per spectrum (limit 8192)"));
        ch.out_off = off;
        off += ch.nbins_out;
        ch.ncells = (double)d->n0 * d->n1 * ch.nz;
        ch.norm = d->volume[c] / (ch.ncells * ch.ncells);
        auto it = shape_of_nz.find(ch.nz);
        if (it == shape_of_nz.end()) {
            ShapePlan sp;
            sp.nz = ch.nz; sp.nh = ch.nh;
            shape_of_nz[ch.nz] = (int)p->shapes.size();
            ch.shape = (int)p->shapes.size();
            p->shapes.push_back(sp);
        } else ch.shape = it->second;
        p->shapes[ch.shape].chunks.push_back(c);
        // k = 0 mode: binned only when it has weight and lies inside the edges (linear bins start at 0)
        ch.dc_bin = -1;
        if (!d->ignore_kperp_zero && !d->ignore_k_zero && !d->ignore_kpar_zero) {
            const double* ed = d->edges + (long long)c * (d->nbins + 1);
            int cnt = 0;
            for (int j = 0; j <= d->nbins; ++j) cnt += (ed[j] <= 0.0);
            if (cnt >= 1 && cnt <= d->nbins) ch.dc_bin = (d->mode == PS21_MODE_SPHERICAL) ? cnt - 1 : (cnt - 1) * ch.nh;
        }
        p->chunks.push_back(ch);
        k2offs.push_back(k2off);
        k2off += ch.nz;
        p->alg_bytes_per_box += 4LL * d->n0 * d->n1 * ch.nz + 16LL * d->n0 * d->n1 * ch.nh;
    }
    p->out_per_box = off;
    // shapes: factorisations, twiddles, tiling, path
    long long spec_max = 0;
    p->slots_max = 1;
    for (auto& sp : p->shapes) {
        sp.fz = factorize(sp.nz, false);
        sp.fy = factorize(d->n1, true);
        sp.fx = factorize(d->n0, true);
        if (sp.fz.nstages < 0 || sp.fy.nstages < 0 || sp.fx.nstages < 0)
            return fail(invalid("axis length has too many prime factors"));
        if ((rc = upload_twiddles(sp.nz, &sp.tw_z))) return fail(rc);
        if ((rc = upload_twiddles(d->n1, &sp.tw_y))) return fail(rc);
        if ((rc = upload_twiddles(d->n0, &sp.tw_x))) return fail(rc);
        PS21_CUDA(cudaMalloc((void**)&sp.d_chunks, sizeof(int) * sp.chunks.size()));
        PS21_CUDA(cudaMemcpy(sp.d_chunks, sp.chunks.data(), sizeof(int) * sp.chunks.size(), cudaMemcpyHostToDevice));
        // fastest first: the hand-tuned fused kernels (256^2 / 128^2 planes, 256-slice chunks), then the generic
        // two-pass kernels, then the three-pass general path (any size)
        sp.path = PS21_PATH_GENERAL;
        if ((d->path == PS21_PATH_AUTO || d->path == PS21_PATH_FUSED) && fused_try_setup(p, sp) == 0)
            sp.path = PS21_PATH_FUSED;
        if (d->path == PS21_PATH_FUSED && sp.path != PS21_PATH_FUSED)
            return fail(invalid("fused path requested but this shape/alignment is not supported by it: " + g_err));
        if (sp.path == PS21_PATH_GENERAL && (d->path == PS21_PATH_AUTO || d->path == PS21_PATH_TWOPASS) &&
            !(d->path == PS21_PATH_AUTO && std::getenv("PS21_NO_TWOPASS")) && g2_try_setup(p, sp) == 0)
            sp.path = PS21_PATH_TWOPASS;
        if (d->path == PS21_PATH_TWOPASS && sp.path != PS21_PATH_TWOPASS)
            return fail(invalid("two-pass path requested but this shape is not supported by it: " + g_err));
        const int nb_out = p->chunks[sp.chunks[0]].nbins_out;
        auto none = [](int) { return (size_t)0; };
        sp.g1_cols = pick_tile(sp.nz, none);
        sp.g2_tile = pick_tile(d->n1, [&](int t) { return (size_t)d->n1 * (t | 1) * sizeof(float2); });  // third buffer
        sp.g3_tile = pick_tile(d->n0, [&](int t) { return general_xpass_extra(nb_out, d->n0, t); });
        // tuning experiments only: another power-of-two tile (<= 16) than the automatic choice, if it fits
        auto tile_override = [](const char* name, int& tile, int n, auto extra) {
            const char* e = std::getenv(name);
            if (!e) return;
            const int t = std::atoi(e);
            if (t >= 1 && t <= 16 && (t & (t - 1)) == 0 &&
                2 * (size_t)n * (t | 1) * sizeof(float2) + extra(t) <= (size_t)220 * 1024)
                tile = t;
        };
        tile_override("PS21_G1_TILE", sp.g1_cols, sp.nz, none);
        tile_override("PS21_G2_TILE", sp.g2_tile, d->n1,
                      [&](int t) { return (size_t)d->n1 * (t | 1) * sizeof(float2); });
        tile_override("PS21_G3_TILE", sp.g3_tile, d->n0, [&](int t) { return general_xpass_extra(nb_out, d->n0, t); });
        if (!sp.g1_cols || !sp.g2_tile || !sp.g3_tile) return fail(invalid("axis too long for the shared-memory FFT"));
        const long long ntiles = ((long long)d->n1 * sp.nh + sp.g3_tile - 1) / sp.g3_tile;
        const long long cap = d->mode == PS21_MODE_SPHERICAL ? 64 : 16;
        sp.g3_slots = (int)(ntiles < cap ? ntiles : cap);
        long long spec = (long long)d->n0 * d->n1 * sp.nh;
        int nslots = sp.g3_slots;
        if (sp.path != PS21_PATH_GENERAL) {
            spec = spectrum_elems(p, sp);
            nslots = p->chunks[sp.chunks[0]].nslots;  // set by fused_try_setup / g2_try_setup
        }
        for (int c : sp.chunks) p->chunks[c].nslots = nslots;
        if (nslots > p->slots_max) p->slots_max = nslots;
        if (spec > spec_max) spec_max = spec;
    }
    // scratch grows on demand (ensure_scratch) up to the cap; one spectrum of the largest shape to start with
    p->spec_max = spec_max;
    PS21_CUDA(cudaMalloc((void**)&p->d_chunks, sizeof(ChunkDev) * p->nchunks));
    PS21_CUDA(cudaMemcpy(p->d_chunks, p->chunks.data(), sizeof(ChunkDev) * p->nchunks, cudaMemcpyHostToDevice));
    // k^3/(2 pi^2) at the bin centres (likelihood.py:863, :652)
    if (d->mode == PS21_MODE_SPHERICAL) {
        std::vector<double> kf(p->out_per_box);
        for (int c = 0; c < d->nchunks; ++c)
            for (int b = 0; b < d->nbins; ++b) {
                const double k = d->kcentres[(long long)c * d->nbins + b];
                kf[p->chunks[c].out_off + b] = k * k * k / (2.0 * M_PI * M_PI);
            }
        PS21_CUDA(cudaMalloc((void**)&p->d_kfac, sizeof(double) * p->out_per_box));
        PS21_CUDA(cudaMemcpy(p->d_kfac, kf.data(), sizeof(double) * p->out_per_box, cudaMemcpyHostToDevice));
    }
    // LUTs + counts
    PS21_CUDA(cudaMalloc((void**)&p->d_counts, sizeof(long long) * p->out_per_box));
    PS21_CUDA(cudaMemset(p->d_counts, 0, sizeof(long long) * p->out_per_box));
    {
        double *dk0 = nullptr, *dk1 = nullptr, *dk2 = nullptr, *ded = nullptr;
        PS21_CUDA(cudaMalloc((void**)&dk0, sizeof(double) * d->n0));
        PS21_CUDA(cudaMalloc((void**)&dk1, sizeof(double) * d->n1));
        PS21_CUDA(cudaMalloc((void**)&dk2, sizeof(double) * k2off));
        PS21_CUDA(cudaMalloc((void**)&ded, sizeof(double) * d->nchunks * (d->nbins + 1)));
        PS21_CUDA(cudaMemcpy(dk0, d->k0sq, sizeof(double) * d->n0, cudaMemcpyHostToDevice));
        PS21_CUDA(cudaMemcpy(dk1, d->k1sq, sizeof(double) * d->n1, cudaMemcpyHostToDevice));
        PS21_CUDA(cudaMemcpy(dk2, d->k2sq, sizeof(double) * k2off, cudaMemcpyHostToDevice));
        PS21_CUDA(cudaMemcpy(ded, d->edges, sizeof(double) * d->nchunks * (d->nbins + 1), cudaMemcpyHostToDevice));
        p->d_lut.assign(d->nchunks, nullptr);
        p->lut_owned.assign(d->nchunks, 0);
        std::vector<int> twin(d->nchunks, -1);
        for (int c = 0; c < d->nchunks && rc == PS21_OK; ++c) {
            const ShapePlan& sp = p->shapes[p->chunks[c].shape];
            // chunks of equal length, edges and k_z share one table (the usual case: equal-length chunks)
            for (int o = 0; o < c && twin[c] < 0; ++o)
                if (p->chunks[o].shape == p->chunks[c].shape && twin[o] < 0 &&
                    std::memcmp(d->edges + (long long)o * (d->nbins + 1), d->edges + (long long)c * (d->nbins + 1),
                                sizeof(double) * (d->nbins + 1)) == 0 &&
                    std::memcmp(d->k2sq + k2offs[o], d->k2sq + k2offs[c], sizeof(double) * p->chunks[c].nz) == 0)
                    twin[c] = o;
            if (twin[c] >= 0) {
                p->d_lut[c] = p->d_lut[twin[c]];
                continue;
            }
            p->lut_owned[c] = 1;
            const long long spec = (sp.path != PS21_PATH_GENERAL) ? spectrum_elems(p, sp)
                                                                   : general_lut_elems(d->n0, d->n1, sp.nh, sp.g3_tile);
            if (cudaMalloc((void**)&p->d_lut[c], sizeof(uint16_t) * spec) != cudaSuccess) {
                cudaGetLastError();
                set_error("cannot allocate bin-index table");
                rc = PS21_ERR_NOMEM;
                break;
            }
            const double* e = ded + (long long)c * (d->nbins + 1);
            rc = (sp.path == PS21_PATH_FUSED)     ? fused_build_lut(p, c, dk0, dk1, dk2 + k2offs[c], e, 0)
                 : (sp.path == PS21_PATH_TWOPASS) ? g2_build_lut(p, c, dk0, dk1, dk2 + k2offs[c], e, 0)
                                                  : general_build_lut(p, c, dk0, dk1, dk2 + k2offs[c], e, 0);
        }
        cudaError_t se = cudaDeviceSynchronize();
        for (int c = 0; c < d->nchunks && se == cudaSuccess; ++c)
            if (twin[c] >= 0)
                se = cudaMemcpy(p->d_counts + p->chunks[c].out_off, p->d_counts + p->chunks[twin[c]].out_off,
                                sizeof(long long) * p->chunks[c].nbins_out, cudaMemcpyDeviceToDevice);
        free_dev(dk0); free_dev(dk1); free_dev(dk2); free_dev(ded);
        if (rc) return fail(rc);
        if (se != cudaSuccess) return fail(cuda_fail(se, "LUT build"));
    }
    PS21_CUDA(cudaMalloc((void**)&p->d_lut_table, sizeof(uint16_t*) * d->nchunks));
    PS21_CUDA(cudaMemcpy(p->d_lut_table, p->d_lut.data(), sizeof(uint16_t*) * d->nchunks, cudaMemcpyHostToDevice));
    p->h_counts.resize(p->out_per_box);
    PS21_CUDA(cudaMemcpy(p->h_counts.data(), p->d_counts, sizeof(long long) * p->out_per_box, cudaMemcpyDeviceToHost));
    *out = p;
    return PS21_OK;
}

int ps21_plan_info(const ps21_plan* p, ps21_info* info) {
    if (!p || !info) return invalid("null argument");
    info->nshapes = (int)p->shapes.size();
    info->path_of_chunk0 = p->shapes[p->chunks[0].shape].path;
    info->out_per_box = p->out_per_box;
    info->workspace_bytes = p->scratch_elems * (long long)sizeof(float2);
    info->alg_bytes_per_box = p->alg_bytes_per_box;
    info->launches_last_call = p->launches_last;
    info->dominant_ms_last = 0.0;
    return PS21_OK;
}

int ps21_plan_set_profiling(ps21_plan* p, int32_t on) {
    if (!p) return invalid("null argument");
    p->profiling = on != 0;
    return PS21_OK;
}

int ps21_plan_profile(ps21_plan* p, double* ms_by_kind, int64_t* launches_by_kind) {
    if (!p || !ms_by_kind || !launches_by_kind) return invalid("null argument");
    for (int k = 0; k < KK_COUNT; ++k) { ms_by_kind[k] = 0.0; launches_by_kind[k] = 0; }
#ifndef PS21_EMU
    PS21_CUDA(cudaSetDevice(p->device));
    for (auto& r : p->prof) {
        cudaEvent_t a = (cudaEvent_t)r.e0, b = (cudaEvent_t)r.e1;
        float ms = 0.f;
        if (cudaEventSynchronize(b) == cudaSuccess && cudaEventElapsedTime(&ms, a, b) == cudaSuccess) {
            ms_by_kind[r.kind] += ms;
            launches_by_kind[r.kind] += 1;
        }
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
#endif
    p->prof.clear();
    return PS21_OK;
}

int ps21_plan_layout(const ps21_plan* p, int64_t* offsets) {
    if (!p || !offsets) return invalid("null argument");
    for (int c = 0; c < p->nchunks; ++c) offsets[c] = p->chunks[c].out_off;
    offsets[p->nchunks] = p->out_per_box;
    return PS21_OK;
}

int ps21_plan_counts(const ps21_plan* p, int64_t* counts) {
    if (!p || !counts) return invalid("null argument");
    for (long long i = 0; i < p->out_per_box; ++i) counts[i] = p->h_counts[i];
    return PS21_OK;
}

int ps21_set_data(ps21_plan* p, int32_t nsets, const double* data, const double* noise, const uint8_t* kmask,
                  double model_uncertainty, int32_t error_on_model) {
    if (!p || !data || !kmask || nsets < 1) return invalid("null argument / nsets < 1");
    if (p->mode != PS21_MODE_SPHERICAL) return invalid("likelihood data need the spherical mode");
    PS21_CUDA(cudaSetDevice(p->device));
    const long long n = (long long)nsets * p->out_per_box;
    if (n > p->data.cap) {   // tables are reused while they are large enough
        drop_graphs(p);
        free_dev(p->data.d_data); free_dev(p->data.d_noise2); free_dev(p->data.d_mask);
        p->data.d_data = p->data.d_noise2 = nullptr;
        p->data.d_mask = nullptr;
        p->data.cap = 0;
        p->data.nsets = 0;
        PS21_CUDA(cudaMalloc((void**)&p->data.d_data, sizeof(double) * n));
        PS21_CUDA(cudaMalloc((void**)&p->data.d_noise2, sizeof(double) * n));
        PS21_CUDA(cudaMalloc((void**)&p->data.d_mask, n));
        p->data.cap = n;
    }
    std::vector<double> n2(n, 0.0);
    if (noise) for (long long i = 0; i < n; ++i) n2[i] = noise[i] * noise[i];
    PS21_CUDA(cudaMemcpy(p->data.d_data, data, sizeof(double) * n, cudaMemcpyHostToDevice));
    PS21_CUDA(cudaMemcpy(p->data.d_noise2, n2.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    PS21_CUDA(cudaMemcpy(p->data.d_mask, kmask, n, cudaMemcpyHostToDevice));
    p->data.nsets = nsets;
    p->data.model_uncertainty = model_uncertainty;
    p->data.error_on_model = error_on_model ? 1 : 0;
    p->data.gen++;   // nsets / model_uncertainty / error_on_model are baked into captured launches
    return PS21_OK;
}

int ps21_power_batched(ps21_plan* p, const float* d_boxes, int64_t batch, int64_t batch_stride, double* d_power,
                       void* stream) {
    if (!p || !d_boxes || !d_power) return invalid("null argument");
    PS21_CUDA(cudaSetDevice(p->device));
    p->launches_last = 0;
    return run_device(p, d_boxes, batch, batch_stride, d_power, nullptr, nullptr, (cudaStream_t)stream);
}

int ps21_lnl_batched(ps21_plan* p, const float* d_boxes, int64_t batch, int64_t batch_stride, double* d_lnl,
                     double* d_delta, double* d_power, void* stream) {
    if (!p || !d_boxes || !d_lnl) return invalid("null argument");
    PS21_CUDA(cudaSetDevice(p->device));
    p->launches_last = 0;
    return run_device(p, d_boxes, batch, batch_stride, d_power, d_delta, d_lnl, (cudaStream_t)stream);
}

#ifndef PS21_EMU
// H2D of `nboxes` host boxes into `d_dst` on stream `st`.  Page-locked sources (cudaHostAlloc / cudaHostRegister:
// ps21_host_alloc, ps21_host_register, torch pinned memory) are handed to the DMA engine as they are.  PAGEABLE
// sources -- what an unmodified simulator returns -- would make cudaMemcpyAsync stage them synchronously through the
// driver's bounce buffer on the calling thread (measured 10.8 GB/s); instead a few threads (3/4 of the hardware
// threads, at most 12; PS21_STAGE_THREADS) copy 16 MiB pieces into a page-locked ring (two slots per thread) and issue the DMA of
// each piece themselves, so the host memcpy of one piece overlaps the DMA of the others (measured 36 GB/s with 12).
static int stage_boxes_h2d(ps21_plan* p, float* d_dst, const float* const* h_ptrs, const float* h_boxes,
                           long long first, long long batch_stride, long long nboxes, long long box_elems,
                           cudaStream_t st) {
    const long long box_bytes = box_elems * (long long)sizeof(float);
    auto src_of = [&](long long i) { return h_ptrs ? h_ptrs[first + i] : h_boxes + (first + i) * batch_stride; };
    bool pageable = false;
    for (long long i = 0; i < nboxes && !pageable; ++i) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, src_of(i)) != cudaSuccess) {
            cudaGetLastError();
            pageable = true;
        } else if (at.type == cudaMemoryTypeUnregistered) {
            pageable = true;
        }
    }
    if (!pageable || std::getenv("PS21_NO_STAGING_THREADS")) {
        if (!h_ptrs && batch_stride == box_elems) {
            PS21_CUDA(cudaMemcpyAsync(d_dst, h_boxes + first * batch_stride, (size_t)(nboxes * box_bytes),
                                      cudaMemcpyHostToDevice, st));
        } else {
            for (long long i = 0; i < nboxes; ++i)
                PS21_CUDA(cudaMemcpyAsync(d_dst + i * box_elems, src_of(i), (size_t)box_bytes, cudaMemcpyHostToDevice, st));
        }
        return PS21_OK;
    }
    constexpr int NTMAX = ps21_plan::kStageThreads;
    int NT = (int)std::thread::hardware_concurrency() * 3 / 4;   // measured on a 16-thread host: 4-8 threads 28 GB/s, 12-16 36 GB/s
    if (NT > 12) NT = 12;
    if (const char* e = std::getenv("PS21_STAGE_THREADS")) NT = std::atoi(e);
    if (NT < 1) NT = 1;
    if (NT > NTMAX) NT = NTMAX;
    const long long SLOT = ps21_plan::kStageSlotBytes;
    if (!p->h_ring) {
        PS21_CUDA(cudaHostAlloc((void**)&p->h_ring, (size_t)(2 * NTMAX * SLOT), cudaHostAllocPortable));
        for (auto& e : p->ring_ev) {
            cudaEvent_t ev;
            PS21_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            e = (void*)ev;
        }
    }
    const long long per_box = (box_bytes + SLOT - 1) / SLOT, npieces = nboxes * per_box;
    cudaError_t errs[NTMAX];
    std::thread th[NTMAX];
    for (int t = 0; t < NT; ++t) {
        errs[t] = cudaSuccess;
        th[t] = std::thread([&, t]() {
            cudaError_t e = cudaSetDevice(p->device);
            int turn = 0;
            for (long long pc = t; pc < npieces && e == cudaSuccess; pc += NT, ++turn) {
                const int slot = 2 * t + (turn & 1);
                const long long i = pc / per_box, off = (pc % per_box) * SLOT;
                const long long n = (box_bytes - off < SLOT) ? box_bytes - off : SLOT;
                char* ring = p->h_ring + (long long)slot * SLOT;
                cudaEvent_t ev = (cudaEvent_t)p->ring_ev[slot];
                e = cudaEventSynchronize(ev);   // the DMA that last read this slot (this call or an earlier one) is done
                if (e != cudaSuccess) break;
                std::memcpy(ring, reinterpret_cast<const char*>(src_of(i)) + off, (size_t)n);
                e = cudaMemcpyAsync(reinterpret_cast<char*>(d_dst + i * box_elems) + off, ring, (size_t)n,
                                    cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) e = cudaEventRecord(ev, st);
            }
            errs[t] = e;
        });
    }
    for (int t = 0; t < NT; ++t) th[t].join();
    for (int t = 0; t < NT; ++t)
        if (errs[t] != cudaSuccess) return cuda_fail(errs[t], "host staging thread");
    return PS21_OK;
}
#endif

// Host-buffer variants: boxes are staged through two device buffers; the copy of slice i+1 overlaps the kernels
// of slice i.  Box b is read from h_ptrs[b] when a pointer array is given (walker-owned buffers, no gather on the
// host), else from h_boxes + b * batch_stride.  Staging buffers, result buffers, streams and events belong to
// the plan and are reused from call to call.
static int run_host(ps21_plan* p, const float* const* h_ptrs, const float* h_boxes, long long batch,
                    long long batch_stride, double* h_lnl, double* h_delta, double* h_power) {
    PS21_CUDA(cudaSetDevice(p->device));
    p->launches_last = 0;
    if (batch <= 0) return PS21_OK;
    const long long box_elems = (long long)p->n0 * p->n1 * p->n2_total;
    if (!h_ptrs && batch_stride < box_elems) return invalid("batch_stride smaller than one box");
    if (h_ptrs)
        for (long long i = 0; i < batch; ++i)
            if (!h_ptrs[i]) return invalid("null box pointer");
    // slice size: at most 1 GiB per staging buffer (PS21_STAGE_BYTES overrides), at least one box
    long long stage_bytes = 1LL << 30;
    if (const char* e = getenv("PS21_STAGE_BYTES")) {
        const long long v = atoll(e);
        if (v > 0) stage_bytes = v;
    }
    long long per = stage_bytes / (box_elems * 4);
    if (per < 1) per = 1;
    if (per > batch) per = batch;
    if (per > p->stage_boxes) {
        for (int i = 0; i < 2; ++i) {
            free_dev(p->d_stage[i]);
            p->d_stage[i] = nullptr;
        }
        p->stage_boxes = 0;
        for (int i = 0; i < 2; ++i) PS21_CUDA(cudaMalloc((void**)&p->d_stage[i], sizeof(float) * per * box_elems));
        p->stage_boxes = per;
    }
    for (int i = 0; i < 2; ++i) {
        if (!p->stage_stream[i]) {
            cudaStream_t s;
            PS21_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
            p->stage_stream[i] = (void*)s;
        }
    }
    for (int i = 0; i < 4; ++i) {
        if (!p->stage_ev[i]) {
            cudaEvent_t e;
            PS21_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            p->stage_ev[i] = (void*)e;
        }
    }
    if (batch > p->res_boxes) {
        for (int i = 0; i < 3; ++i) {
            free_dev(p->d_res[i]);
            p->d_res[i] = nullptr;
        }
        p->res_boxes = 0;
        PS21_CUDA(cudaMalloc((void**)&p->d_res[0], sizeof(double) * batch * p->out_per_box));
        PS21_CUDA(cudaMalloc((void**)&p->d_res[1], sizeof(double) * batch * p->out_per_box));
        PS21_CUDA(cudaMalloc((void**)&p->d_res[2], sizeof(double) * batch));
        p->res_boxes = batch;
    }
    cudaStream_t s_copy = (cudaStream_t)p->stage_stream[0], s_comp = (cudaStream_t)p->stage_stream[1];
    int rc = ensure_batch(p, per);
    if (rc) return rc;
    double *d_res_p = p->d_res[0], *d_res_d = h_delta ? p->d_res[1] : nullptr, *d_res_l = h_lnl ? p->d_res[2] : nullptr;
    // "copy into buffer b done" is recorded per slice on s_copy and compute waits on it; the copy of slice i+2
    // waits for the compute of slice i (recorded on s_comp) before it overwrites the buffer
    cudaEvent_t ev_copy[2] = {(cudaEvent_t)p->stage_ev[0], (cudaEvent_t)p->stage_ev[1]};
    cudaEvent_t ev_comp[2] = {(cudaEvent_t)p->stage_ev[2], (cudaEvent_t)p->stage_ev[3]};
    long long done = 0, launches = 0;
    int sl = 0;
    while (done < batch && rc == PS21_OK) {
        const long long n = (batch - done < per) ? (batch - done) : per;
        const int b = sl & 1;
        if (sl >= 2) cudaStreamWaitEvent(s_copy, ev_comp[b], 0);
#ifndef PS21_EMU
        rc = stage_boxes_h2d(p, p->d_stage[b], h_ptrs, h_boxes, done, batch_stride, n, box_elems, s_copy);
        if (rc) break;
#else
        for (long long i = 0; i < n; ++i)
            cudaMemcpyAsync(p->d_stage[b] + i * box_elems,
                            h_ptrs ? h_ptrs[done + i] : h_boxes + (done + i) * batch_stride,
                            sizeof(float) * box_elems, cudaMemcpyHostToDevice, s_copy);
#endif
        cudaEventRecord(ev_copy[b], s_copy);
        cudaStreamWaitEvent(s_comp, ev_copy[b], 0);
        rc = run_device(p, p->d_stage[b], n, box_elems, d_res_p + done * p->out_per_box,
                        d_res_d ? d_res_d + done * p->out_per_box : nullptr, d_res_l ? d_res_l + done : nullptr,
                        s_comp, done, /*allow_graph=*/false);
        launches += p->launches_last;
        cudaEventRecord(ev_comp[b], s_comp);
        done += n;
        ++sl;
    }
    p->launches_last = launches;
    if (rc == PS21_OK) {
        if (h_power)
            cudaMemcpyAsync(h_power, d_res_p, sizeof(double) * batch * p->out_per_box, cudaMemcpyDeviceToHost, s_comp);
        if (h_delta)
            cudaMemcpyAsync(h_delta, d_res_d, sizeof(double) * batch * p->out_per_box, cudaMemcpyDeviceToHost, s_comp);
        if (h_lnl) cudaMemcpyAsync(h_lnl, d_res_l, sizeof(double) * batch, cudaMemcpyDeviceToHost, s_comp);
    }
    cudaError_t e1 = cudaStreamSynchronize(s_copy);
    cudaError_t e2 = cudaStreamSynchronize(s_comp);
    if (rc) return rc;
    if (e1 != cudaSuccess) return cuda_fail(e1, "host staging copy stream");
    if (e2 != cudaSuccess) return cuda_fail(e2, "host staging compute stream");
    return PS21_OK;
}

int ps21_power_host(ps21_plan* p, const float* h_boxes, int64_t batch, int64_t batch_stride, double* h_power) {
    if (!p || !h_boxes || !h_power) return invalid("null argument");
    return run_host(p, nullptr, h_boxes, batch, batch_stride, nullptr, nullptr, h_power);
}

int ps21_lnl_host(ps21_plan* p, const float* h_boxes, int64_t batch, int64_t batch_stride, double* h_lnl,
                  double* h_delta, double* h_power) {
    if (!p || !h_boxes || !h_lnl) return invalid("null argument");
    if (p->data.nsets == 0) {
        set_error("ps21_set_data must be called before a likelihood evaluation");
        return PS21_ERR_STATE;
    }
    return run_host(p, nullptr, h_boxes, batch, batch_stride, h_lnl, h_delta, h_power);
}

int ps21_power_host_ptrs(ps21_plan* p, const float* const* h_box_ptrs, int64_t batch, double* h_power) {
    if (!p || !h_box_ptrs || !h_power) return invalid("null argument");
    return run_host(p, h_box_ptrs, nullptr, batch, 0, nullptr, nullptr, h_power);
}

int ps21_lnl_host_ptrs(ps21_plan* p, const float* const* h_box_ptrs, int64_t batch, double* h_lnl, double* h_delta,
                       double* h_power) {
    if (!p || !h_box_ptrs || !h_lnl) return invalid("null argument");
    if (p->data.nsets == 0) {
        set_error("ps21_set_data must be called before a likelihood evaluation");
        return PS21_ERR_STATE;
    }
    return run_host(p, h_box_ptrs, nullptr, batch, 0, h_lnl, h_delta, h_power);
}

int ps21_lnl_from_delta(ps21_plan* p, const double* h_delta, int64_t batch, double* h_lnl) {
    if (!p || !h_delta || !h_lnl) return invalid("null argument");
    if (p->data.nsets == 0) {
        set_error("ps21_set_data must be called before a likelihood evaluation");
        return PS21_ERR_STATE;
    }
    PS21_CUDA(cudaSetDevice(p->device));
    if (batch <= 0) return PS21_OK;
    DevTmp<double> tdd, tdl;
    PS21_CUDA(tdd.alloc((size_t)batch * p->out_per_box));
    PS21_CUDA(tdl.alloc((size_t)batch));
    double *dd = tdd.p, *dl = tdl.p;
    PS21_CUDA(cudaMemcpy(dd, h_delta, sizeof(double) * batch * p->out_per_box, cudaMemcpyHostToDevice));
    PS21_LAUNCH(k_lnl_from_delta, dim3((unsigned)((batch + 63) / 64)), dim3(64), 0, (cudaStream_t)0, dd,
                p->out_per_box, p->d_chunks, p->nchunks, dl, p->data.d_data, p->data.d_noise2, p->data.d_mask,
                p->data.nsets, p->data.model_uncertainty, p->data.error_on_model, (long long)batch);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(h_lnl, dl, sizeof(double) * batch, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(e, "lnl_from_delta");
    p->launches_last = 1;
    return PS21_OK;
}

int ps21_chi2_host(int32_t device, int64_t nrec, int64_t n, const double* delta, const double* data,
                   const double* noise, const uint8_t* mask, int64_t nseg, const int64_t* seg_len,
                   double model_uncertainty, int32_t error_on_model, double* lnl) {
    if (!delta || !data || !mask || !seg_len || !lnl || nrec < 0 || n < 0) return invalid("null argument");
    long long tot = 0;
    for (int64_t s = 0; s < nseg; ++s) tot += seg_len[s];
    if (tot != n) return invalid("segment lengths do not add up to n");
    if (nrec == 0) return PS21_OK;
    PS21_CUDA(cudaSetDevice(device));
    double *dd = nullptr, *da = nullptr, *dn = nullptr, *dl = nullptr;
    uint8_t* dm = nullptr;
    long long* ds = nullptr;
    cudaError_t e = cudaSuccess;
    auto up = [&](void** dst, const void* src, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(dst, bytes ? bytes : 1);
        if (e == cudaSuccess && bytes) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    up((void**)&dd, delta, sizeof(double) * nrec * n);
    up((void**)&da, data, sizeof(double) * n);
    if (noise) up((void**)&dn, noise, sizeof(double) * n);
    up((void**)&dm, mask, (size_t)n);
    up((void**)&ds, seg_len, sizeof(long long) * nseg);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dl, sizeof(double) * nrec);
    if (e == cudaSuccess) {
        PS21_LAUNCH(k_chi2_rows, dim3((unsigned)((nrec + 63) / 64)), dim3(64), 0, (cudaStream_t)0, dd, (long long)n,
                    (long long)nrec, da, dn, dm, (long long)nseg, ds, model_uncertainty, error_on_model ? 1 : 0, dl);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(lnl, dl, sizeof(double) * nrec, cudaMemcpyDeviceToHost);
    free_dev(dd); free_dev(da); free_dev(dn); free_dev(dm); free_dev(ds); free_dev(dl);
    if (e != cudaSuccess) return cuda_fail(e, "ps21_chi2_host");
    return PS21_OK;
}

// Grow-only device scratch of the plan-free entry points (one per host thread): ps21_upper_limit_host is called once
// per (band, field) block -- ten times per evaluation with the HERA file -- and used to pay six cudaMalloc / cudaFree
// pairs and six copies each time.
struct ScratchArena {
    void* d = nullptr;
    size_t cap = 0;
    int device = -1;
    std::vector<double> h;
    int ensure(int dev, size_t bytes) {
        if (dev == device && bytes <= cap) return PS21_OK;
        if (d) {
            cudaSetDevice(device);
            cudaFree(d);
            d = nullptr;
            cap = 0;
        }
        PS21_CUDA(cudaSetDevice(dev));
        const size_t want = bytes < (1u << 20) ? (1u << 20) : bytes * 2;
        PS21_CUDA(cudaMalloc(&d, want));
        cap = want;
        device = dev;
        return PS21_OK;
    }
};
static thread_local ScratchArena g_arena;

static int windowed_host(int32_t device, int64_t nrec, int64_t nk, int64_t nkwf, const double* model,
                         const double* wf, const double* limit, const double* var, const double* model_err,
                         double frac, int gauss, double* lnl) {
    if (!model || !wf || !limit || !var || !lnl || nrec < 0 || nk < 0 || nkwf < 0) return invalid("null argument");
    if (nrec == 0 || nk == 0) return PS21_OK;
    PS21_CUDA(cudaSetDevice(device));
    // one packed buffer: [model nrec*nkwf][wf nk*nkwf][limit nk][var nk][model_err nkwf | -][lnl nrec]
    const size_t n_model = (size_t)(nrec * nkwf), n_wf = (size_t)(nk * nkwf), n_err = model_err ? (size_t)nkwf : 0;
    const size_t o_wf = n_model, o_lim = o_wf + n_wf, o_var = o_lim + (size_t)nk, o_err = o_var + (size_t)nk,
                 o_out = o_err + n_err, total = o_out + (size_t)nrec;
    ScratchArena& A = g_arena;
    int rc = A.ensure(device, total * sizeof(double));
    if (rc) return rc;
    A.h.resize(total);
    std::memcpy(A.h.data(), model, n_model * sizeof(double));
    std::memcpy(A.h.data() + o_wf, wf, n_wf * sizeof(double));
    std::memcpy(A.h.data() + o_lim, limit, (size_t)nk * sizeof(double));
    std::memcpy(A.h.data() + o_var, var, (size_t)nk * sizeof(double));
    if (model_err) std::memcpy(A.h.data() + o_err, model_err, n_err * sizeof(double));
    std::memcpy(A.h.data() + o_out, lnl, (size_t)nrec * sizeof(double));
    double* d = static_cast<double*>(A.d);
    PS21_CUDA(cudaMemcpy(d, A.h.data(), total * sizeof(double), cudaMemcpyHostToDevice));
    PS21_LAUNCH(k_upper_limit, dim3((unsigned)nrec), dim3(128), 0, (cudaStream_t)0, d, (long long)nrec, (int)nk, (int)nkwf,
                d + o_wf, d + o_lim, d + o_var, model_err ? d + o_err : nullptr, frac, gauss, d + o_out);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(lnl, d + o_out, sizeof(double) * nrec, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(e, gauss ? "ps21_windowed_gauss_host" : "ps21_upper_limit_host");
    return PS21_OK;
}

int ps21_upper_limit_host(int32_t device, int64_t nrec, int64_t nk, int64_t nkwf, const double* model,
                          const double* wf, const double* limit, const double* var, const double* model_err,
                          double frac, double* lnl) {
    return windowed_host(device, nrec, nk, nkwf, model, wf, limit, var, model_err, frac, 0, lnl);
}

int ps21_windowed_gauss_host(int32_t device, int64_t nrec, int64_t nk, int64_t nkwf, const double* model,
                             const double* wf, const double* value, const double* var, const double* model_err,
                             double frac, double* lnl) {
    return windowed_host(device, nrec, nk, nkwf, model, wf, value, var, model_err, frac, 1, lnl);
}

}  // extern "C"
