// thread_pool.h -- minimal blocking fork/join pool for the host-side scene preparation.
// (No OpenMP in the product: libgomp's default spin-waiting costs tens of milliseconds per
// parallel region inside CPU-quota'd containers, which would dominate small-scene build time.)
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace b200rt {

class ThreadPool {
public:
    explicit ThreadPool(int threads) {
        if (threads < 1) threads = 1;
        n_ = threads;
        for (int i = 1; i < n_; ++i) workers_.emplace_back([this] { worker(); });
    }
    ~ThreadPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return n_; }

    // Runs fn(i) for i in [0, count) on the pool (the caller participates); returns when all done.
    void parallel_for(int count, const std::function<void(int)> &fn) {
        if (count <= 0) return;
        if (n_ == 1 || count == 1) {
            for (int i = 0; i < count; ++i) fn(i);
            return;
        }
        auto job = std::make_shared<Job>();
        job->fn = &fn;
        job->count = count;
        job->pending.store(count, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> g(m_);
            job_ = job;
            ++generation_;
        }
        cv_.notify_all();
        run_items(*job);
        std::unique_lock<std::mutex> lk(m_);
        done_cv_.wait(lk, [&] { return job->pending.load(std::memory_order_acquire) == 0; });
        job_.reset();
    }

    static int hardware_threads() {
        unsigned n = std::thread::hardware_concurrency();
        return n ? (int)n : 1;
    }

private:
    struct Job {
        const std::function<void(int)> *fn = nullptr;
        int count = 0;
        std::atomic<int> next{0};
        std::atomic<int> pending{0};
    };
    void run_items(Job &job) {
        while (true) {
            const int i = job.next.fetch_add(1, std::memory_order_relaxed);
            if (i >= job.count) break;
            (*job.fn)(i);
            if (job.pending.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                std::lock_guard<std::mutex> g(m_);
                done_cv_.notify_all();
            }
        }
    }
    void worker() {
        unsigned long seen = 0;
        while (true) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                job = job_;
            }
            if (job) run_items(*job);
        }
    }

    int n_ = 1;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    std::shared_ptr<Job> job_;
    unsigned long generation_ = 0;
    bool stop_ = false;
};

}  // namespace b200rt
