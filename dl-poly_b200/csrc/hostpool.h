// hostpool.h -- the worker pool behind the packed record transfers of hostio.cu: plain C++ (no CUDA), so that tests/test_hostpool.py
// can build it with a host compiler and a thread sanitizer.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace dlp_hostpool {

constexpr int MAXCH = 64;   // chunks per transfer (one event each on the way down)

// A fixed set of workers that run fn(0), fn(1), ... fn(n-1), handing the chunks out in increasing order.
class HostPool {
 public:
  explicit HostPool(int nthreads) : nthreads_(nthreads) {
    for (int c = 0; c < MAXCH; ++c) done_[c].store(0, std::memory_order_relaxed);
    for (int t = 0; t < nthreads_; ++t) th_.emplace_back([this] { work(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int threads() const { return nthreads_; }
  void start(int n, std::function<void(int)> fn) {   // the previous job must have been waited for (wait_all)
    if (nthreads_ == 0) {   // no workers: the calling thread does the chunks itself
      for (int c = 0; c < n; ++c) { fn(c); done_[c].store(1, std::memory_order_relaxed); }
      return;
    }
    for (int c = 0; c < n; ++c) done_[c].store(0, std::memory_order_relaxed);
    next_.store(0, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = std::move(fn); n_ = n; left_ = nthreads_; ++gen_;
    }
    cv_.notify_all();
  }
  void wait_chunk(int c) {
    unsigned spins = 0;
    while (!done_[c].load(std::memory_order_acquire)) pause(++spins);
  }
  void wait_all() {
    if (nthreads_ == 0) return;
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [this] { return left_ == 0; });
  }
  static void pause(unsigned spins) {
    if (spins & 0x3f) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    } else {
      std::this_thread::yield();
    }
  }

 private:
  void work() {
    unsigned long long seen = 0;
    for (;;) {
      std::function<void(int)> fn;
      int n;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_; fn = fn_; n = n_;
      }
      for (;;) {
        const int c = next_.fetch_add(1, std::memory_order_relaxed);
        if (c >= n) break;
        fn(c);
        done_[c].store(1, std::memory_order_release);
      }
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--left_ == 0) cv_done_.notify_all();
      }
    }
  }
  int nthreads_;
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, cv_done_;
  std::function<void(int)> fn_;
  int n_ = 0, left_ = 0;
  unsigned long long gen_ = 0;
  bool stop_ = false;
  std::atomic<int> next_{0};
  std::atomic<int> done_[MAXCH];
};

}  // namespace dlp_hostpool
