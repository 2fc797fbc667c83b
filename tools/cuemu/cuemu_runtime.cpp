// cuemu — DEVELOPMENT TOOL (see include/cuda_runtime.h): the block scheduler.
#include "include/cuda_runtime.h"

#include <atomic>
#include <fcntl.h>
#include <map>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

// A context switch without the two sigprocmask system calls swapcontext() makes: callee-saved registers on the old stack,
// switch stack pointers, restore from the new stack (System V x86-64).
extern "C" void cuemu_switch(void** save_sp, void* load_sp);
asm(R"(
    .text
    .globl cuemu_switch
    .type cuemu_switch, @function
cuemu_switch:
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
    .size cuemu_switch, .-cuemu_switch
)");
#if !defined(__x86_64__)
#    error "cuemu's context switch is written for x86-64"
#endif

namespace cuemu
{
namespace
{
constexpr size_t kStack = 256 * 1024;
struct Block
{
    void*                          sched_sp = nullptr;
    std::vector<void*>             sp;
    std::vector<std::vector<char>> stack;
    std::vector<char>              done;
    int                            current = -1;
    int                            alive = 0, arrived = 0; // __syncthreads(): threads still running / threads waiting at the barrier
    unsigned                       generation = 0;         // bumped when a barrier releases
    const std::function<void()>*   body = nullptr;
};
thread_local Block    t_block;
thread_local Builtins t_builtins;

void thread_entry()
{
    Block&    b = t_block;
    const int i = b.current;
    (*b.body)();
    b.done[size_t(i)] = 1;
    if (--b.alive > 0 && b.arrived >= b.alive) b.arrived = 0, ++b.generation; // a thread that has exited no longer holds a barrier up
    void* dead;
    cuemu_switch(&dead, b.sched_sp); // back to the scheduler for good
    __builtin_trap();
}

void* fresh_stack(std::vector<char>& mem)
{
    if (mem.empty()) mem.resize(kStack);
    // top of stack: [r15 r14 r13 r12 rbx rbp][return address = thread_entry]; the return-address slot is 16-byte aligned so
    // that thread_entry starts with the stack the ABI expects after a call
    uintptr_t top = (reinterpret_cast<uintptr_t>(mem.data()) + kStack) & ~uintptr_t(15);
    void**    p   = reinterpret_cast<void**>(top - 16);
    p[0]          = reinterpret_cast<void*>(&thread_entry);
    for (int k = 1; k <= 6; ++k) p[-k] = nullptr;
    return p - 6;
}

void run_block(dim3 grid, dim3 block, uint3 bidx, const std::function<void()>& body)
{
    Block&    b = t_block;
    const int n = int(block.x * block.y * block.z);
    if (int(b.sp.size()) < n) b.sp.resize(size_t(n)), b.stack.resize(size_t(n));
    b.done.assign(size_t(n), 0);
    b.body = &body;
    b.alive = n, b.arrived = 0;
    for (int i = 0; i < n; ++i) b.sp[size_t(i)] = fresh_stack(b.stack[size_t(i)]);
    // CUEMU_ORDER=reverse | shuffle: the order in which the threads of a block run between two barriers. Results must not
    // depend on it; if they do, a __syncthreads() is missing (or shared memory is read before it is written).
    static const int order = [] {
        const char* e = std::getenv("CUEMU_ORDER");
        return !e ? 0 : (e[0] == 'r' ? 1 : 2);
    }();
    unsigned         lcg = 12345u + bidx.x * 7919u + bidx.y * 104729u;
    std::vector<int> perm(static_cast<size_t>(n), 0);
    for (int k = 0; k < n; ++k) perm[size_t(k)] = order == 1 ? n - 1 - k : k;
    for (bool running = true; running;)
    {
        running = false;
        if (order == 2) // a fresh permutation every round (Fisher-Yates)
            for (int k = n - 1; k > 0; --k)
            {
                lcg = lcg * 1664525u + 1013904223u;
                std::swap(perm[size_t(k)], perm[size_t((lcg >> 8) % unsigned(k + 1))]);
            }
        for (int k = 0; k < n; ++k)
        {
            const int i = perm[size_t(k)];
            if (b.done[size_t(i)]) continue;
            b.current  = i;
            t_builtins = Builtins{uint3{unsigned(i) % block.x, (unsigned(i) / block.x) % block.y, unsigned(i) / (block.x * block.y)}, bidx, block, grid};
            cuemu_switch(&b.sched_sp, b.sp[size_t(i)]); // runs thread i up to its next __syncthreads(), or to its end
            running = running || !b.done[size_t(i)];
        }
    }
    b.current = -1;
}
} // namespace

Builtins& builtins() { return t_builtins; }

void yield()
{
    Block& b = t_block;
    if (b.current < 0) return;
    cuemu_switch(&b.sp[size_t(b.current)], b.sched_sp); // the scheduler resumes this thread in its next round
}

// Counting barrier (threads may also yield elsewhere, e.g. while polling an mbarrier, so "everybody has yielded" is not enough)
void sync_threads()
{
    Block& b = t_block;
    if (b.current < 0) return;
    const unsigned g = b.generation;
    if (++b.arrived >= b.alive)
    {
        b.arrived = 0, ++b.generation;
        return;
    }
    while (b.generation == g) yield();
}

// Persistent workers: the per-thread coroutine stacks (64 MB per worker) are allocated once, not per launch.
namespace
{
struct Pool
{
    std::mutex               m;
    std::condition_variable  cv_work, cv_done;
    std::vector<std::thread> threads;
    const std::function<void(unsigned)>* job = nullptr;
    unsigned long long       generation = 0;
    unsigned                 pending = 0, active = 0;

    explicit Pool(unsigned n)
    {
        for (unsigned w = 0; w < n; ++w)
            threads.emplace_back([this, w] {
                unsigned long long seen = 0;
                for (;;)
                {
                    const std::function<void(unsigned)>* j;
                    {
                        std::unique_lock<std::mutex> lk(m);
                        cv_work.wait(lk, [&] { return generation != seen; });
                        seen = generation;
                        j    = job;
                    }
                    if (w < active) (*j)(w);
                    {
                        std::lock_guard<std::mutex> lk(m);
                        if (--pending == 0) cv_done.notify_one();
                    }
                }
            });
        for (auto& t : threads) t.detach();
    }
    void run(unsigned workers, const std::function<void(unsigned)>& fn)
    {
        std::unique_lock<std::mutex> lk(m);
        job = &fn, active = workers, pending = unsigned(threads.size()), ++generation;
        cv_work.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};
Pool& pool()
{
    const char*  env = std::getenv("CUEMU_THREADS");
    static Pool* p   = new Pool(std::max(1u, env ? unsigned(std::atoi(env)) : std::max(1u, std::thread::hardware_concurrency())));
    return *p;
}
std::mutex g_launch_mutex; // one grid at a time (the C-ABI may be called from several host threads)
} // namespace

void launch(dim3 grid, dim3 block, const std::function<void()>& body)
{
    std::lock_guard<std::mutex> one(g_launch_mutex);
    const unsigned total   = grid.x * grid.y * grid.z;
    Pool&          p       = pool();
    const unsigned workers = std::max(1u, std::min(total, unsigned(p.threads.size())));
    std::atomic<unsigned> next{0};
    p.run(workers, [&](unsigned) {
        for (unsigned i = next.fetch_add(1); i < total; i = next.fetch_add(1))
            run_block(grid, block, uint3{i % grid.x, (i / grid.x) % grid.y, i / (grid.x * grid.y)}, body);
    });
}
} // namespace cuemu

// ---- device memory: heap, or (CUEMU_IPC=1) named shared-memory segments that another process can map ----
namespace
{
struct Segment
{
    std::string name;
    size_t      bytes;
    bool        owner;
};
std::mutex                g_seg_mutex;
std::map<void*, Segment>  g_segments;
std::atomic<unsigned>     g_seg_counter{0};
bool ipc_mode() { return std::getenv("CUEMU_IPC") != nullptr; }
} // namespace

cudaError_t cudaMalloc(void** p, size_t bytes)
{
    *p = nullptr;
    if (!ipc_mode()) return posix_memalign(p, 256, bytes ? bytes : 256) == 0 ? cudaSuccess : cudaErrorMemoryAllocation;
    const std::string name = "/cuemu-" + std::to_string(getpid()) + "-" + std::to_string(g_seg_counter++);
    const int         fd   = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return cudaErrorMemoryAllocation;
    const size_t n = std::max<size_t>(bytes, 256);
    if (ftruncate(fd, off_t(n)) != 0) return close(fd), shm_unlink(name.c_str()), cudaErrorMemoryAllocation;
    void* m = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return shm_unlink(name.c_str()), cudaErrorMemoryAllocation;
    std::lock_guard<std::mutex> lk(g_seg_mutex);
    g_segments[m] = Segment{name, n, true};
    return *p = m, cudaSuccess;
}
cudaError_t cudaFree(void* p)
{
    if (!p) return cudaSuccess;
    std::lock_guard<std::mutex> lk(g_seg_mutex);
    auto                        it = g_segments.find(p);
    if (it == g_segments.end()) return free(p), cudaSuccess;
    munmap(p, it->second.bytes);
    if (it->second.owner) shm_unlink(it->second.name.c_str());
    g_segments.erase(it);
    return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p)
{
    std::lock_guard<std::mutex> lk(g_seg_mutex);
    auto                        it = g_segments.find(p);
    if (it == g_segments.end() || it->second.name.size() >= sizeof(h->reserved)) return cudaErrorNotSupported;
    std::memset(h->reserved, 0, sizeof(h->reserved));
    std::memcpy(h->reserved, it->second.name.c_str(), it->second.name.size());
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned)
{
    h.reserved[sizeof(h.reserved) - 1] = 0;
    const int fd = shm_open(h.reserved, O_RDWR, 0600);
    if (fd < 0) return cudaErrorNotSupported;
    struct stat st;
    if (fstat(fd, &st) != 0) return close(fd), cudaErrorNotSupported;
    void* m = mmap(nullptr, size_t(st.st_size), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return cudaErrorNotSupported;
    std::lock_guard<std::mutex> lk(g_seg_mutex);
    g_segments[m] = Segment{h.reserved, size_t(st.st_size), false};
    return *p = m, cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) { return cudaFree(p); }
