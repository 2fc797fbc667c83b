// cuemu — DEVELOPMENT TOOL (see include/cuda_runtime.h): the block scheduler.
#include "include/cuda_runtime.h"

#include <thread>
#include <ucontext.h>
#include <vector>

namespace cuemu
{
namespace
{
constexpr size_t kStack = 256 * 1024;
struct Block
{
    ucontext_t                     sched;
    std::vector<ucontext_t>        ctx;
    std::vector<std::vector<char>> stack;
    std::vector<char>              done;
    int                            current = -1;
    const std::function<void()>*   body = nullptr;
};
thread_local Block    t_block;
thread_local Builtins t_builtins;

void thread_entry()
{
    Block&    b = t_block;
    const int i = b.current;
    (*b.body)();
    b.done[size_t(i)] = 1; // returning activates uc_link == the scheduler
}

void run_block(dim3 grid, dim3 block, uint3 bidx, const std::function<void()>& body)
{
    Block&    b = t_block;
    const int n = int(block.x * block.y * block.z);
    if (int(b.ctx.size()) < n) b.ctx.resize(size_t(n)), b.stack.resize(size_t(n));
    b.done.assign(size_t(n), 0);
    b.body = &body;
    for (int i = 0; i < n; ++i)
    {
        if (b.stack[size_t(i)].empty()) b.stack[size_t(i)].resize(kStack);
        getcontext(&b.ctx[size_t(i)]);
        b.ctx[size_t(i)].uc_stack.ss_sp   = b.stack[size_t(i)].data();
        b.ctx[size_t(i)].uc_stack.ss_size = kStack;
        b.ctx[size_t(i)].uc_link          = &b.sched;
        makecontext(&b.ctx[size_t(i)], thread_entry, 0);
    }
    for (bool running = true; running;)
    {
        running = false;
        for (int i = 0; i < n; ++i)
        {
            if (b.done[size_t(i)]) continue;
            b.current  = i;
            t_builtins = Builtins{uint3{unsigned(i) % block.x, (unsigned(i) / block.x) % block.y, unsigned(i) / (block.x * block.y)}, bidx, block, grid};
            swapcontext(&b.sched, &b.ctx[size_t(i)]); // runs thread i up to its next __syncthreads(), or to its end
            running = running || !b.done[size_t(i)];
        }
    }
    b.current = -1;
}
} // namespace

Builtins& builtins() { return t_builtins; }

void sync_threads()
{
    Block& b = t_block;
    if (b.current < 0) return;
    swapcontext(&b.ctx[size_t(b.current)], &b.sched); // every thread that has not finished reaches the same barrier before anyone continues
}

void launch(dim3 grid, dim3 block, const std::function<void()>& body)
{
    const unsigned total   = grid.x * grid.y * grid.z;
    const char*    env     = std::getenv("CUEMU_THREADS");
    const unsigned workers = std::max(1u, std::min(total, env ? unsigned(std::atoi(env)) : std::max(1u, std::thread::hardware_concurrency())));
    auto           work    = [&](unsigned w) {
        for (unsigned i = w; i < total; i += workers) run_block(grid, block, uint3{i % grid.x, (i / grid.x) % grid.y, i / (grid.x * grid.y)}, body);
    };
    if (workers == 1) return work(0);
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < workers; ++w) pool.emplace_back(work, w);
    for (auto& t : pool) t.join();
}
} // namespace cuemu
