// cuda_shim.cpp -- TEST INFRASTRUCTURE: the kernel launcher of the host emulation (cuda_shim.h)
//
// Lock-stepped lanes (WARP / BLOCK mode) are cooperative FIBERS on the calling OS thread
// (x86-64: a six-register stack switch; a barrier / shuffle / ballot that has to wait hands
// the core to the next lane), so a __syncthreads costs a few hundred nanoseconds instead of
// a round of kernel scheduling.  Elsewhere (or with B200SPH_EMUL_THREADS=1) every lane is a
// real OS thread, which is what this file did before and is kept as the cross-check: lanes
// then really run concurrently, atomics included.
#include "cuda_shim.h"

#include <cstdlib>
#include <memory>

namespace emu {
thread_local idx3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
thread_local WarpCtx *t_warp = nullptr;
thread_local BlockCtx *t_block = nullptr;
thread_local int t_lane = 0;

static void set_ids(long long grid, long long block, long long b, long long t)
{
    t_gridDim = idx3{(unsigned)grid, 1, 1};
    t_blockDim = idx3{(unsigned)block, 1, 1};
    t_blockIdx = idx3{(unsigned)b, 0, 0};
    t_threadIdx = idx3{(unsigned)t, 0, 0};
    t_lane = (int)(t & 31);
}

#if defined(__x86_64__)
#define EMU_FIBERS 1
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

struct Fiber {
    void *sp = nullptr;
    char *stack = nullptr;
    bool done = true;
    long long tid = 0;
    WarpCtx *warp = nullptr;
};
static const size_t FIBER_STACK = 512 * 1024;
struct FiberSet {
    std::vector<Fiber> f;
    void *main_sp = nullptr;
    int cur = -1;
    const std::function<void()> *body = nullptr;
    ~FiberSet() { for (auto &x : f) free(x.stack); }
};
static thread_local FiberSet *t_set = nullptr;

static void fiber_entry()
{
    FiberSet *S = t_set;
    Fiber &me = S->f[(size_t)S->cur];
    (*S->body)();
    me.done = true;
    emu_switch(&me.sp, S->main_sp);
    abort();   // a finished fiber is never resumed
}

static void run_fibers(long long grid, long long block, long long b, long long first, int n, BlockCtx *bc,
                       WarpCtx *wc, const std::function<void()> &body)
{
    static thread_local std::unique_ptr<FiberSet> pool;
    if (!pool) pool.reset(new FiberSet);
    FiberSet *S = pool.get();
    if ((int)S->f.size() < n) S->f.resize((size_t)n);
    S->body = &body;
    FiberSet *outer = t_set;
    t_set = S;
    for (int i = 0; i < n; i++) {
        Fiber &f = S->f[(size_t)i];
        if (!f.stack) f.stack = (char *)malloc(FIBER_STACK);
        uintptr_t top = ((uintptr_t)f.stack + FIBER_STACK) & ~(uintptr_t)15;
        void **sp = (void **)top;
        *--sp = nullptr;                 // keeps rsp = 8 (mod 16) at fiber_entry, as after a call
        *--sp = (void *)&fiber_entry;    // popped by emu_switch's ret
        for (int r = 0; r < 6; r++) *--sp = nullptr;
        f.sp = sp;
        f.done = false;
        f.tid = first + i;
        f.warp = bc ? &bc->warps[(size_t)(f.tid >> 5)] : wc;
    }
    int left = n;
    while (left > 0) {
        for (int i = 0; i < n; i++) {
            Fiber &f = S->f[(size_t)i];
            if (f.done) continue;
            S->cur = i;
            set_ids(grid, block, b, f.tid);
            t_block = bc;
            t_warp = f.warp;
            emu_switch(&S->main_sp, f.sp);
            if (f.done) left--;
        }
    }
    S->cur = -1;
    t_block = nullptr;
    t_warp = nullptr;
    t_set = outer;
}

void yield()
{
    FiberSet *S = t_set;
    if (S && S->cur >= 0) emu_switch(&S->f[(size_t)S->cur].sp, S->main_sp);
    else std::this_thread::yield();
}
#else
void yield() { std::this_thread::yield(); }
#endif

void launch(long long grid, long long block, int mode, const std::function<void()> &body)
{
#ifdef EMU_FIBERS
    static const bool use_threads = getenv("B200SPH_EMUL_THREADS") && atoi(getenv("B200SPH_EMUL_THREADS")) != 0;
#else
    static const bool use_threads = true;
#endif
    for (long long b = 0; b < grid; b++) {
        if (mode == SEQ) {
            for (long long t = 0; t < block; t++) {
                set_ids(grid, block, b, t);
                body();
            }
        } else if (mode == WARP) {   // warps are independent: one warp at a time, 32 lock-stepped lanes
            for (long long w = 0; w < (block + 31) / 32; w++) {
                WarpCtx wc;
#ifdef EMU_FIBERS
                if (!use_threads) {
                    run_fibers(grid, block, b, w * 32, 32, nullptr, &wc, body);   // block sizes are multiples of 32 in this library
                    continue;
                }
#endif
                std::vector<std::thread> lanes;
                for (int l = 0; l < 32; l++)
                    lanes.emplace_back([&, l] {
                        set_ids(grid, block, b, w * 32 + l);
                        t_warp = &wc;
                        if (w * 32 + l < block) body();
                        t_warp = nullptr;
                    });
                for (auto &t : lanes) t.join();
            }
        } else {                     // BLOCK: every thread of the block runs in lock step
            BlockCtx bc((int)block);
#ifdef EMU_FIBERS
            if (!use_threads) {
                run_fibers(grid, block, b, 0, (int)block, &bc, nullptr, body);
                continue;
            }
#endif
            std::vector<std::thread> ths;
            for (long long t = 0; t < block; t++)
                ths.emplace_back([&, t] {
                    set_ids(grid, block, b, t);
                    t_block = &bc;
                    t_warp = &bc.warps[(size_t)(t >> 5)];
                    body();
                    t_block = nullptr;
                    t_warp = nullptr;
                });
            for (auto &t : ths) t.join();
        }
    }
}
}  // namespace emu
