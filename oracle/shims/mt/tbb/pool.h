// Oracle shim (timed CPU baseline flavour): a std::thread pool standing in for oneTBB's arena.
// oneTBB is not installed in this image and cannot be fetched (no network); this pool gives the
// reference's loops the same fan-out over all host cores with static contiguous chunking.
// Worker w always takes chunk w, so anything ordered by worker id is ordered by input index.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
namespace tbb {
namespace shim {
inline int &worker_id() { static thread_local int id = 0; return id; }
inline bool &in_parallel() { static thread_local bool f = false; return f; }
class Pool {
public:
    static Pool &get() { static Pool p; return p; }
    int size() const { return n_; }
    // run body(chunk_index, begin, end) over [0, total) split into size() contiguous chunks
    void run(std::size_t total, const std::function<void(int, std::size_t, std::size_t)> &body) {
        if (total == 0) return;
        if (n_ == 1 || in_parallel() || total < 2) { body(worker_id(), 0, total); return; }
        {
            std::unique_lock<std::mutex> lk(m_);
            body_ = &body; total_ = total; pending_ = n_ - 1; ++epoch_;
        }
        cv_.notify_all();
        exec(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
        body_ = nullptr;
    }
private:
    Pool() {
        const char *e = std::getenv("LIMU_REF_THREADS");
        n_ = e ? std::atoi(e) : static_cast<int>(std::thread::hardware_concurrency());
        if (n_ < 1) n_ = 1;
        for (int w = 1; w < n_; ++w) threads_.emplace_back([this, w] { loop(w); });
    }
    ~Pool() {
        { std::unique_lock<std::mutex> lk(m_); stop_ = true; ++epoch_; }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    void exec(int w) {
        const std::size_t per = (total_ + n_ - 1) / n_;
        const std::size_t b = per * w, e = b + per < total_ ? b + per : total_;
        const int saved = worker_id();
        worker_id() = w; in_parallel() = true;
        if (b < e) (*body_)(w, b, e);
        in_parallel() = false; worker_id() = saved;
    }
    void loop(int w) {
        unsigned long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
            }
            exec(w);
            std::unique_lock<std::mutex> lk(m_);
            if (--pending_ == 0) done_.notify_one();
        }
    }
    int n_ = 1;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int, std::size_t, std::size_t)> *body_ = nullptr;
    std::size_t total_ = 0;
    int pending_ = 0;
    unsigned long epoch_ = 0;
    bool stop_ = false;
};
}  // namespace shim
}  // namespace tbb
