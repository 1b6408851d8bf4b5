/* TEST INFRASTRUCTURE — block / fiber scheduler of the SIMT emulator (see simt_emu.h). */
#include "simt_emu.h"
#include <sys/mman.h>

namespace simt {
Block g_blk;
unsigned char* g_dyn_smem = nullptr;
dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
static long g_launches = 0;
static std::vector<char*> g_stacks;

long launches() { return g_launches; }

static void next_or_main() {
    Block& b = g_blk;
    int n = b.cur;
    for (int k = 0; k < b.nthreads; k++) {
        n = (n + 1 == b.nthreads) ? 0 : n + 1;
        if (!b.fibers[n].done) { b.cur = n; g_threadIdx.x = (unsigned)n; setcontext(&b.fibers[n].ctx); }
    }
    setcontext(&b.main_ctx);
}

static void trampoline() {
    Block& b = g_blk;
    (*b.body)();
    b.fibers[b.cur].done = true;      // a returned thread no longer takes part in barriers (waiters re-evaluate)
    b.live--;
    b.spins = 0;
    next_or_main();
}

void launch(int grid, int block, size_t smem, const std::function<void()>& body) {
    g_launches++;
    if (std::getenv("SIMT_TRACE")) std::fprintf(stderr, "simt launch %ld: grid %d block %d smem %zu\n", g_launches, grid, block, smem);
    if (block < 1 || block > 1024 || grid < 0) { std::fprintf(stderr, "simt_emu: bad launch configuration\n"); std::abort(); }
    Block& b = g_blk;
    while ((int)g_stacks.size() < block) {
        void* p = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { std::perror("simt_emu: mmap"); std::abort(); }
        g_stacks.push_back((char*)p);
    }
    std::vector<unsigned char> dyn(smem + 64, 0xCB);      // uninitialised shared memory reads show up as garbage
    g_dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    g_gridDim = dim3((unsigned)grid); g_blockDim = dim3((unsigned)block);
    b.body = &body;
    for (int blk = 0; blk < grid; blk++) {
        g_blockIdx = dim3((unsigned)blk);
        b.nthreads = block; b.live = block; b.cur = 0; b.bar_count = 0; b.bar_gen = 0; b.spins = 0;
        b.fibers.assign((size_t)block, Fiber());
        b.warps.assign((size_t)(block + 31) / 32, Warp());
        for (int t = 0; t < block; t++) {
            Fiber& f = b.fibers[t];
            f.stack = g_stacks[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK_BYTES; f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, trampoline, 0);
        }
        g_threadIdx = dim3(0);
        swapcontext(&b.main_ctx, &b.fibers[0].ctx);
    }
    b.body = nullptr;
    g_dyn_smem = nullptr;
}
}  // namespace simt
