// lbk_team.hpp -- the launcher threads of a single-process multi-device context (lbk_init_multi).
//
// The reference is one process, one thread (lambda-blas.cabal:36-48; call surface Single.hs:84-85); a drop-in
// call such as `sdot n u 1 v 1` on vectors sharded over the GPUs of a box therefore has to fan out from ONE
// host thread.  Launching N kernels one after the other costs N x (2-3 us) before the last GPU starts, which
// is as long as the kernels themselves once n/N elements stream in a few tens of microseconds.  A Team keeps
// one launcher thread per rank >= 1 (rank 0 is served by the calling thread): a call publishes one job, every
// member runs it for its own rank, and the caller returns when all have.  Members spin for a short while
// after a job (calls usually come in bursts) and then sleep on a condition variable, so an idle context
// costs no CPU.
//
// Pure host C++ (no CUDA types): the CPU tests drive it through lbk_team_selftest.
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace lbk_team {

class Team {
  public:
    // `on_start(rank)` runs once in each member thread before its first job (the library binds the thread to its device)
    Team(int nranks, std::function<void(int)> on_start) : n_(nranks), status_(nranks > 0 ? nranks : 1, 0)
    {
        for (int r = 1; r < n_; ++r) threads_.emplace_back([this, r, on_start] { member(r, on_start); });
    }
    ~Team()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            seq_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    Team(const Team &) = delete;
    Team &operator=(const Team &) = delete;

    int size() const { return n_; }

    // Runs job(rank) for EVERY rank of the team -- rank 0 on the calling thread, the others on their own
    // threads -- and returns the first non-zero status in rank order (0 if none).  One call at a time.  (Every
    // member takes part in every job, so no member can still be looking at job k when job k+1 is published.)
    int run(const std::function<int(int)> &job)
    {
        if (n_ <= 1) return n_ == 1 ? job(0) : 0;
        job_ = &job;
        pending_.store(n_ - 1, std::memory_order_relaxed);
        {
            // the lock orders the sequence bump against a member that is about to sleep (no lost wake-up)
            std::lock_guard<std::mutex> lk(m_);
            seq_.fetch_add(1, std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        status_[0] = job(0);
        while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();
        for (int r = 0; r < n_; ++r)
            if (status_[r]) return status_[r];
        return 0;
    }

  private:
    static void cpu_relax()
    {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }

    void member(int rank, const std::function<void(int)> &on_start)
    {
        if (on_start) on_start(rank);
        unsigned long long seen = 0;
        for (;;) {
            // wait for the next job: spin (bursts of calls), then sleep
            const auto t0 = std::chrono::steady_clock::now();
            unsigned spins = 0;
            while (seq_.load(std::memory_order_acquire) == seen) {
                cpu_relax();
                if ((++spins & 1023u) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) {
                    std::unique_lock<std::mutex> lk(m_);
                    sleepers_.fetch_add(1, std::memory_order_release);
                    cv_.wait(lk, [&] { return seq_.load(std::memory_order_acquire) != seen; });
                    sleepers_.fetch_sub(1, std::memory_order_release);
                }
            }
            seen = seq_.load(std::memory_order_acquire);
            if (stop_) return;
            status_[rank] = (*job_)(rank);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }

    const int n_;
    std::vector<std::thread> threads_;
    std::vector<int> status_;
    const std::function<int(int)> *job_ = nullptr;
    std::atomic<unsigned long long> seq_{0};
    std::atomic<int> pending_{0};
    std::atomic<int> sleepers_{0};
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
};

}  // namespace lbk_team
