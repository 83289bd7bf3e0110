// staging.hpp -- host <-> device copies for the host-pointer entry points.
//
// barretenberg allocates every polynomial and scalar array with aligned_alloc (bb/common/mem.hpp:26-44): pageable
// memory.  cudaMemcpyAsync from/to pageable memory is staged by the driver on ONE thread (measured here: 10.6 GB/s
// H2D, 19 GB/s D2H, against ~52 GB/s for pinned memory), which makes the copies -- not the kernels -- the cost of a
// drop-in FFT call (2^18 elements: 1.39 ms wall, 0.07 ms of kernels).  For large pageable buffers this file stages
// through two pinned buffers with a small pool of host threads doing the memcpy, so the DMA of chunk k overlaps
// the host copy of chunk k+1.  Pinned / registered / managed memory and small copies go straight to cudaMemcpyAsync.
#pragma once
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "ctx.cuh"

namespace bbg {

class CopyPool {
  public:
    explicit CopyPool(unsigned nthreads) : n_(nthreads < 1 ? 1 : nthreads)
    {
        for (unsigned i = 1; i < n_; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++gen_;
        }
        cv_start_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // dst[0, bytes) = src[0, bytes), split over the pool; returns when every slice is copied
    void copy(void* dst, const void* src, size_t bytes)
    {
        if (n_ == 1 || bytes < (256u << 10)) {
            memcpy(dst, src, bytes);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = (char*)dst;
            src_ = (const char*)src;
            bytes_ = bytes;
            remaining_ = n_ - 1;
            ++gen_;
        }
        cv_start_.notify_all();
        slice(0);
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [this] { return remaining_ == 0; });
    }

  private:
    void slice(unsigned i)
    {
        const size_t per = ((bytes_ + n_ - 1) / n_ + 4095) & ~(size_t)4095;
        const size_t lo = (size_t)i * per;
        if (lo >= bytes_) return;
        const size_t len = bytes_ - lo < per ? bytes_ - lo : per;
        memcpy(dst_ + lo, src_ + lo, len);
    }
    void loop(unsigned i)
    {
        uint64_t seen = 0;
        while (true) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_start_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            slice(i);
            {
                std::lock_guard<std::mutex> lk(mu_);
                --remaining_;
            }
            cv_done_.notify_one();
        }
    }
    unsigned n_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_start_, cv_done_;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0;
    unsigned remaining_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

struct Staging {
    static constexpr size_t CHUNK = 4u << 20;     // bytes per pinned buffer
    // Measured on the B200 box (ifft through bbg_ntt, pageable numpy buffers, wall ms, driver path -> staged path):
    // 2 MB 0.45 -> 0.68, 8 MB 1.39 -> 1.81, 32 MB 5.08 -> 4.78, 128 MB 19.8 -> 14.9.  The pool's wake-up latency per
    // 4 MB chunk costs more than it saves on small buffers, so only large copies are staged.
    static constexpr size_t MIN_BYTES = 32u << 20;
    void* buf[2] = { nullptr, nullptr };
    cudaEvent_t ev[2] = { nullptr, nullptr };
    CopyPool* pool = nullptr;
    bool disabled = false;

    int ensure()
    {
        if (buf[0] != nullptr || disabled) return BBG_OK;
        const char* v = getenv("BBG_STAGING_THREADS"); // 0 disables the staged path
        unsigned hw = std::thread::hardware_concurrency();
        // measured on a 16-vCPU B200 box (MSM 2^20 end to end from pageable scalars, ms): driver path 5.05; pool of 2: 4.49, 4: 4.54,
        // 8: 4.67, 12: 7.63, 16: 7.65 (pinned scalars: 3.69) -- a few copy threads beat the driver, many fight the CUDA threads
        unsigned n = v && *v ? (unsigned)atoi(v) : (hw >= 8 ? 4u : (hw >= 4 ? 2u : 1u));
        if (v && *v && n == 0) {
            disabled = true;
            return BBG_OK;
        }
        for (int i = 0; i < 2; ++i) {
            BBG_CUDA(cudaHostAlloc(&buf[i], CHUNK, cudaHostAllocDefault));
            BBG_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        pool = new CopyPool(n);
        return BBG_OK;
    }
    void release()
    {
        delete pool;
        pool = nullptr;
        for (int i = 0; i < 2; ++i) {
            if (buf[i]) cudaFreeHost(buf[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
            buf[i] = nullptr;
            ev[i] = nullptr;
        }
    }
    static bool pageable(const void* p)
    {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
            cudaGetLastError();
            return true;
        }
        return a.type == cudaMemoryTypeUnregistered;
    }

    // asynchronous with respect to the device (ordered on `st`); the host buffer may be reused on return
    int h2d(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st)
    {
        if (bytes == 0) return BBG_OK;
        int rc = ensure();
        if (rc) return rc;
        if (disabled || bytes < MIN_BYTES || !pageable(h_src)) {
            BBG_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
            return BBG_OK;
        }
        size_t k = 0;
        for (size_t off = 0; off < bytes; off += CHUNK, ++k) {
            const int b = (int)(k & 1);
            const size_t len = bytes - off < CHUNK ? bytes - off : CHUNK;
            if (k >= 2) BBG_CUDA(cudaEventSynchronize(ev[b])); // the DMA that last read this buffer is done
            pool->copy(buf[b], (const char*)h_src + off, len);
            BBG_CUDA(cudaMemcpyAsync((char*)d_dst + off, buf[b], len, cudaMemcpyHostToDevice, st));
            BBG_CUDA(cudaEventRecord(ev[b], st));
        }
        return BBG_OK;
    }
    // synchronous: on return h_dst holds the data (everything queued on `st` before the call has completed)
    int d2h(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st)
    {
        if (bytes == 0) return BBG_OK;
        int rc = ensure();
        if (rc) return rc;
        if (disabled || bytes < MIN_BYTES || !pageable(h_dst)) {
            BBG_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
            BBG_CUDA(cudaStreamSynchronize(st));
            return BBG_OK;
        }
        const size_t chunks = (bytes + CHUNK - 1) / CHUNK;
        auto issue = [&](size_t k) -> int {
            const size_t off = k * CHUNK;
            const size_t len = bytes - off < CHUNK ? bytes - off : CHUNK;
            BBG_CUDA(cudaMemcpyAsync(buf[k & 1], (const char*)d_src + off, len, cudaMemcpyDeviceToHost, st));
            BBG_CUDA(cudaEventRecord(ev[k & 1], st));
            return BBG_OK;
        };
        auto drain = [&](size_t k) -> int {
            const size_t off = k * CHUNK;
            const size_t len = bytes - off < CHUNK ? bytes - off : CHUNK;
            BBG_CUDA(cudaEventSynchronize(ev[k & 1]));
            pool->copy((char*)h_dst + off, buf[k & 1], len);
            return BBG_OK;
        };
        if ((rc = issue(0))) return rc;
        for (size_t k = 1; k < chunks; ++k) {
            if ((rc = issue(k))) return rc; // buffer k & 1 held chunk k - 2, drained in the previous iteration
            if ((rc = drain(k - 1))) return rc;
        }
        return drain(chunks - 1);
    }
};

} // namespace bbg
