// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h). Pinned against the reference's own shaders (oracle/refshader, tests/test_reference_shaders.py).
//
// Texture model: what the HLSL relies on from the fixed-function units:
//   * Texture.Load out of bounds returns 0 (D3D)                                            -> Tex::load
//   * point / bilinear SampleLevel with clamp or border(0) addressing                      -> sample_*()
//   * SampleLevel(point-mip sampler, fractional LOD) picks the nearest mip (round half up)  -> nearest_mip()
//   * render-target storage: fp32 planes (the north-star layout); no UNORM/half/R11G11B10 quantisation.
// Bilinear SampleLevel positions are snapped to 1/256 texel (8-bit sub-texel precision of the fixed-function sampler);
// the weights derived from the snapped position are then exact in fp32. Shader-side bilinear arithmetic
// (GetBilinearSamplingInfoUC) is NOT snapped — it is ordinary fp32 shader math.
#pragma once
#include "oracle_math.h"
#include <vector>
#include <thread>
#include <functional>
#include <mutex>
#include <condition_variable>
#include <atomic>
#include <algorithm>
#include <memory>

namespace orc
{

template <class T>
struct Tex
{
    int            w = 0, h = 0;
    std::vector<T> d;

    Tex() = default;
    Tex(int w_, int h_, T fill = T()) : w(w_), h(h_), d(size_t(w_) * size_t(h_), fill) {}
    void resize(int w_, int h_, T fill = T())
    {
        w = w_;
        h = h_;
        d.assign(size_t(w_) * size_t(h_), fill);
    }
    void     fill(T v) { std::fill(d.begin(), d.end(), v); }
    T&       at(int x, int y) { return d[size_t(y) * w + x]; }
    const T& at(int x, int y) const { return d[size_t(y) * w + x]; }
    // Texture.Load: out-of-bounds -> 0
    T load(int x, int y) const
    {
        if (x < 0 || y < 0 || x >= w || y >= h) return T();
        return d[size_t(y) * w + x];
    }
    T load(int2 p) const { return load(p.x, p.y); }
    T load_clamped(int x, int y) const { return d[size_t(clampi(y, 0, h - 1)) * w + clampi(x, 0, w - 1)]; }
};

template <class T>
struct MipTex
{
    std::vector<Tex<T>> mip;
    void create(int w, int h, int levels, T fill = T())
    {
        mip.clear();
        for (int i = 0; i < levels; ++i) mip.emplace_back(std::max(w >> i, 1), std::max(h >> i, 1), fill);
    }
    int levels() const { return int(mip.size()); }
};

// ComputeMipLevelsCount (DiligentCore GraphicsAccessories): number of mips down to 1x1 of the larger dimension.
inline int compute_mip_levels_count(int w, int h)
{
    int m = std::max(w, h), n = 0;
    while (m > 0)
    {
        ++n;
        m >>= 1;
    }
    return n;
}

enum class Address
{
    Clamp,
    Border
};

// Point sampling at normalised uv: texel = floor(uv * size), clamp addressing.
template <class T>
inline T sample_point_clamp(const Tex<T>& t, float2 uv)
{
    int x = int(std::floor(uv.x * float(t.w)));
    int y = int(std::floor(uv.y * float(t.h)));
    return t.load_clamped(x, y);
}

// Bilinear sampling at normalised uv.
template <class T>
inline T sample_linear(const Tex<T>& t, float2 uv, Address addr)
{
    float px = uv.x * float(t.w) - 0.5f;
    float py = uv.y * float(t.h) - 0.5f;
    // fixed-function samplers resolve the sample position to 8 fractional bits (D3D11 functional spec, texture
    // coordinate / filter-weight precision): snap to the nearest 1/256 texel. Taps aimed at texel centres or texel
    // corners therefore get exact weights (1, or 1/2-1/2), as they do on the reference's GPUs.
    px = std::floor(px * 256.0f + 0.5f) * (1.0f / 256.0f);
    py = std::floor(py * 256.0f + 0.5f) * (1.0f / 256.0f);
    float fx0 = std::floor(px), fy0 = std::floor(py);
    int   x0 = int(fx0), y0 = int(fy0);
    float fx = px - fx0, fy = py - fy0;
    T     t00, t10, t01, t11;
    if (addr == Address::Clamp)
    {
        t00 = t.load_clamped(x0, y0);
        t10 = t.load_clamped(x0 + 1, y0);
        t01 = t.load_clamped(x0, y0 + 1);
        t11 = t.load_clamped(x0 + 1, y0 + 1);
    }
    else
    {
        t00 = t.load(x0, y0);
        t10 = t.load(x0 + 1, y0);
        t01 = t.load(x0, y0 + 1);
        t11 = t.load(x0 + 1, y0 + 1);
    }
    float w00 = (1.0f - fx) * (1.0f - fy), w10 = fx * (1.0f - fy), w01 = (1.0f - fx) * fy, w11 = fx * fy;
    return t00 * w00 + t10 * w10 + t01 * w01 + t11 * w11;
}

// Nearest-mip selection for a point-mip sampler given a fractional LOD.
inline int nearest_mip(float lod, int levels) { return clampi(int(std::floor(lod + 0.5f)), 0, levels - 1); }

// Row-parallel helper: the oracle's "all host cores" mode for the cpu_baseline leg. threads <= 1 -> serial.
// A persistent pool (no thread creation per pass) deals rows in small dynamic chunks: per-row cost varies by an order of magnitude
// (sky rows vs reflective rows of the SSR march), and the small pyramid levels are too short for one static chunk per thread.
// Rows are independent in every pass, so the results do not depend on the schedule.
class RowPool
{
public:
    static RowPool& get()
    {
        static RowPool pool;
        return pool;
    }
    void run(int y0, int y1, int threads, const std::function<void(int, int)>& fn)
    {
        const int rows = y1 - y0;
        if (threads <= 1 || rows < 2 || in_job())
        {
            fn(y0, y1);
            return;
        }
        std::lock_guard<std::mutex> serial(run_mutex_); // one job at a time (two oracles in one process take turns)
        const int helpers = std::min(threads - 1, rows - 1); // never more threads than rows
        ensure_workers(helpers);
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_      = &fn;
            end_     = y1;
            chunk_   = std::max(1, rows / ((helpers + 1) * 8));
            next_.store(y0, std::memory_order_relaxed);
            pending_ = helpers;
            ++generation_;
            for (int i = 0; i < helpers; ++i) workers_[i]->wanted = generation_;
        }
        for (int i = 0; i < helpers; ++i) workers_[i]->cv.notify_one(); // only the workers this job uses wake up
        in_job() = true;
        drain();
        in_job() = false;
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }
    ~RowPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        for (auto& w : workers_) w->cv.notify_one();
        for (auto& w : workers_) w->th.join();
    }

private:
    struct Worker
    {
        std::thread             th;
        std::condition_variable cv;
        unsigned                wanted = 0, seen = 0; // generation this worker is asked to join / has joined (guarded by m_)
    };
    static bool& in_job()
    {
        static thread_local bool flag = false;
        return flag;
    }
    void drain()
    {
        for (;;)
        {
            const int a = next_.fetch_add(chunk_, std::memory_order_relaxed);
            if (a >= end_) break;
            (*fn_)(a, std::min(a + chunk_, end_));
        }
    }
    void ensure_workers(int n)
    {
        std::lock_guard<std::mutex> lk(m_);
        while (int(workers_.size()) < n)
        {
            workers_.emplace_back(new Worker);
            Worker* w = workers_.back().get();
            w->th     = std::thread([this, w] {
                in_job() = true;
                std::unique_lock<std::mutex> lk2(m_);
                for (;;)
                {
                    w->cv.wait(lk2, [&] { return stop_ || w->wanted != w->seen; });
                    if (stop_) return;
                    w->seen = w->wanted;
                    lk2.unlock();
                    drain();
                    lk2.lock();
                    if (--pending_ == 0) cv_done_.notify_one();
                }
            });
        }
    }
    std::mutex                           run_mutex_, m_;
    std::condition_variable              cv_done_;
    std::vector<std::unique_ptr<Worker>> workers_;
    const std::function<void(int, int)>* fn_ = nullptr;
    std::atomic<int>                     next_{0};
    int                                  end_ = 0, chunk_ = 1, pending_ = 0;
    unsigned                             generation_ = 0;
    bool                                 stop_ = false;
};
inline void parallel_rows(int y0, int y1, int threads, const std::function<void(int, int)>& fn) { RowPool::get().run(y0, y1, threads, fn); }

} // namespace orc
