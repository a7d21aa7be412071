// TEST INFRASTRUCTURE ONLY: fiber scheduler of the host CUDA emulation (see cuda_emu.h).
#include "cuda_emu.h"

namespace ps21emu {
uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);
char* g_dyn_smem = nullptr;

namespace {
constexpr size_t kStack = 96 * 1024;
struct Fiber {
    ucontext_t ctx;
    bool done = false;
};
ucontext_t g_sched;
std::vector<Fiber> g_fibers;
std::vector<char> g_stacks;
int g_cur = -1;
const std::function<void()>* g_body = nullptr;

void trampoline() {
    (*g_body)();
    g_fibers[g_cur].done = true;
    swapcontext(&g_fibers[g_cur].ctx, &g_sched);
}
}  // namespace

void syncthreads() { swapcontext(&g_fibers[g_cur].ctx, &g_sched); }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const int nt = (int)(block.x * block.y * block.z);
    std::vector<char> dyn(smem + 64);
    g_dyn_smem = dyn.data();
    g_body = &body;
    blockDim = block;
    gridDim = grid;
    if (g_stacks.size() < (size_t)nt * kStack) g_stacks.resize((size_t)nt * kStack);
    g_fibers.assign(nt, Fiber());
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx = uint3{bx, by, bz};
                for (int t = 0; t < nt; ++t) {
                    Fiber& f = g_fibers[t];
                    f.done = false;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = g_stacks.data() + (size_t)t * kStack;
                    f.ctx.uc_stack.ss_size = kStack;
                    f.ctx.uc_link = &g_sched;
                    makecontext(&f.ctx, trampoline, 0);
                }
                bool alive = true;
                while (alive) {  // one sweep = all threads advance to their next barrier
                    alive = false;
                    for (int t = 0; t < nt; ++t) {
                        if (g_fibers[t].done) continue;
                        g_cur = t;
                        threadIdx = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y),
                                          (unsigned)(t / (block.x * block.y))};
                        swapcontext(&g_sched, &g_fibers[t].ctx);
                        if (!g_fibers[t].done) alive = true;
                    }
                }
            }
    g_dyn_smem = nullptr;
}
}  // namespace ps21emu
