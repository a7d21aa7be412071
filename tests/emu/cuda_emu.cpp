// TEST INFRASTRUCTURE ONLY: fiber scheduler of the host CUDA emulation (see cuda_emu.h).
#include "cuda_emu.h"

namespace ps21emu {
uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);
char* g_dyn_smem = nullptr;

namespace {
constexpr size_t kStack = 128 * 1024;
enum State { RUN, WAIT_BLOCK, WAIT_WARP, DONE };
struct Fiber {
    ucontext_t ctx;
    State st = RUN;
};
ucontext_t g_sched;
std::vector<Fiber> g_fibers;
std::vector<char> g_stacks;
std::vector<unsigned long long> g_shfl;
int g_cur = -1;
const std::function<void()>* g_body = nullptr;

void trampoline() {
    (*g_body)();
    g_fibers[g_cur].st = DONE;
    swapcontext(&g_fibers[g_cur].ctx, &g_sched);
}
void yield_as(State s) {
    g_fibers[g_cur].st = s;
    swapcontext(&g_fibers[g_cur].ctx, &g_sched);
}
}  // namespace

void syncthreads() { yield_as(WAIT_BLOCK); }
void syncwarp() { yield_as(WAIT_WARP); }
void spin_yield() { yield_as(RUN); }
unsigned long long* shfl_slot(int tid) { return &g_shfl[tid]; }
int linear_tid() { return g_cur; }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const int nt = (int)(block.x * block.y * block.z);
    std::vector<char> dyn(smem + 256);
    g_dyn_smem = dyn.data() + (128 - (reinterpret_cast<uintptr_t>(dyn.data()) & 127)) % 128;
    g_body = &body;
    blockDim = block;
    gridDim = grid;
    if (g_stacks.size() < (size_t)nt * kStack) g_stacks.resize((size_t)nt * kStack);
    g_shfl.assign(nt, 0);
    g_fibers.assign(nt, Fiber());
    const int nwarps = (nt + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx = uint3{bx, by, bz};
                for (int t = 0; t < nt; ++t) {
                    Fiber& f = g_fibers[t];
                    f.st = RUN;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = g_stacks.data() + (size_t)t * kStack;
                    f.ctx.uc_stack.ss_size = kStack;
                    f.ctx.uc_link = &g_sched;
                    makecontext(&f.ctx, trampoline, 0);
                }
                for (;;) {
                    // run every runnable fiber up to its next barrier
                    bool ran = false;
                    for (int t = 0; t < nt; ++t) {
                        if (g_fibers[t].st != RUN) continue;
                        g_cur = t;
                        threadIdx = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y),
                                          (unsigned)(t / (block.x * block.y))};
                        swapcontext(&g_sched, &g_fibers[t].ctx);
                        ran = true;
                    }
                    // release warps whose live lanes all wait at a warp barrier
                    bool released = false;
                    for (int w = 0; w < nwarps; ++w) {
                        bool all = true, any = false;
                        for (int t = 32 * w; t < nt && t < 32 * w + 32; ++t) {
                            if (g_fibers[t].st == WAIT_WARP) any = true;
                            else if (g_fibers[t].st != DONE) all = false;
                        }
                        if (any && all) {
                            for (int t = 32 * w; t < nt && t < 32 * w + 32; ++t)
                                if (g_fibers[t].st == WAIT_WARP) g_fibers[t].st = RUN;
                            released = true;
                        }
                    }
                    if (released) continue;
                    // block barrier: every live fiber waits at it
                    bool all = true, any = false;
                    for (int t = 0; t < nt; ++t) {
                        if (g_fibers[t].st == WAIT_BLOCK) any = true;
                        else if (g_fibers[t].st != DONE) all = false;
                    }
                    if (any && all) {
                        for (int t = 0; t < nt; ++t)
                            if (g_fibers[t].st == WAIT_BLOCK) g_fibers[t].st = RUN;
                        continue;
                    }
                    if (!any && all) break;  // everything done
                    if (!ran) std::abort();  // deadlock: mismatched barriers
                }
            }
    g_dyn_smem = nullptr;
}
}  // namespace ps21emu
