// cusim.cpp - fiber scheduler of the CPU execution-model simulator (TEST TOOL ONLY, see cusim.h).
#include "cusim.h"

namespace cusim {

Block* g_blk = nullptr;
unsigned long g_progress = 0;
uint3_ g_blockIdx;
dim3 g_blockDim, g_gridDim;
static const size_t kStack = 256 * 1024;
static std::vector<unsigned char> g_dyn;

unsigned char* dyn_smem() { return g_blk->dyn_smem; }

static void trampoline() {
    Block* b = g_blk;
    b->body();
    b->fibers[b->cur].done = true;
    swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body) {
    int nthreads = block.x * block.y * block.z;
    Block blk;
    blk.nthreads = nthreads;
    blk.body = body;
    int nwarps = (nthreads + 31) / 32;
    blk.wbar_count.assign(nwarps, 0);
    blk.wbar_gen.assign(nwarps, 0);
    blk.wslot.assign(nwarps * 32, 0);
    blk.fibers.resize(nthreads);
    g_dyn.assign(smem + 64, 0);
    blk.dyn_smem = (unsigned char*)(((uintptr_t)g_dyn.data() + 15) & ~(uintptr_t)15);
    std::vector<char*> stacks(nthreads);
    for (int t = 0; t < nthreads; ++t) stacks[t] = (char*)malloc(kStack);
    g_blockDim = block;
    g_gridDim = grid;
    Block* saved = g_blk;
    g_blk = &blk;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = uint3_{bx, by, bz};
                blk.bar_count = 0;
                std::fill(blk.wbar_count.begin(), blk.wbar_count.end(), 0);
                for (int t = 0; t < nthreads; ++t) {
                    Fiber& f = blk.fibers[t];
                    f.lin = t;
                    f.tid.x = t % block.x;
                    f.tid.y = (t / block.x) % block.y;
                    f.tid.z = t / (block.x * block.y);
                    f.done = false;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = stacks[t];
                    f.ctx.uc_stack.ss_size = kStack;
                    f.ctx.uc_link = &blk.sched;
                    makecontext(&f.ctx, (void (*)())trampoline, 0);
                }
                int alive = nthreads;
                while (alive > 0) {
                    int progressed = 0;
                    unsigned long p0 = g_progress;
                    for (int t = 0; t < nthreads; ++t) {
                        if (blk.fibers[t].done) continue;
                        blk.cur = t;
                        swapcontext(&blk.sched, &blk.fibers[t].ctx);
                        if (blk.fibers[t].done) { --alive; ++progressed; }
                    }
                    if (!progressed && g_progress == p0) {
                        fprintf(stderr, "cusim: deadlock (barrier divergence?) in block %u,%u,%u\n", bx, by, bz);
                        abort();
                    }
                }
            }
    g_blk = saved;
    for (int t = 0; t < nthreads; ++t) free(stacks[t]);
}

}  // namespace cusim
