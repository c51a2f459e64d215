// cuda_shim.cpp -- TEST INFRASTRUCTURE: the kernel launcher of the host emulation (cuda_shim.h)
#include "cuda_shim.h"

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

void launch(long long grid, long long block, int mode, const std::function<void()> &body)
{
    for (long long b = 0; b < grid; b++) {
        if (mode == SEQ) {
            for (long long t = 0; t < block; t++) {
                set_ids(grid, block, b, t);
                body();
            }
        } else if (mode == WARP) {   // warps are independent: one warp at a time, 32 real threads
            for (long long w = 0; w < (block + 31) / 32; w++) {
                WarpCtx wc;
                std::vector<std::thread> lanes;
                for (int l = 0; l < 32; l++)
                    lanes.emplace_back([&, l] {
                        set_ids(grid, block, b, w * 32 + l);
                        t_warp = &wc;
                        if (w * 32 + l < block) body();
                        else for (;;) break;   // block sizes are multiples of 32 in this library
                        t_warp = nullptr;
                    });
                for (auto &t : lanes) t.join();
            }
        } else {                     // BLOCK: every thread of the block is a real thread
            BlockCtx bc((int)block);
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
