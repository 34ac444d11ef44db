// Persistent host worker threads shared by the batch packer (batch_wire.cu) and the many-structure graph builder
// (graph_builder.cu): parked on a condition variable between jobs, never joined.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>

#include <unistd.h>

namespace chg {

// ---- persistent workers ------------------------------------------------------------------------------------------
class WorkerPool {
 public:
  explicit WorkerPool(int n_workers) : pid_(getpid()) {
    for (int i = 0; i < n_workers; ++i) std::thread(&WorkerPool::loop, this, i + 1).detach();
    n_workers_ = n_workers;
  }
  pid_t pid() const { return pid_; }
  int size() const { return n_workers_ + 1; }
  // fn(k) for k in [0, n): k = 0 on the calling thread, the rest on the workers; returns when all are done
  void run(int n, const std::function<void(int)>& fn) {
    std::lock_guard<std::mutex> one_job(run_m_);  // callers on different threads take turns
    n = std::max(1, std::min(n, size()));
    if (n > 1) {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &fn;
      job_n_ = n;
      pending_.store(n - 1, std::memory_order_relaxed);
      ++generation_;
    }
    if (n > 1) start_.notify_all();
    fn(0);
    if (n > 1) {
      std::unique_lock<std::mutex> lk(m_);
      done_.wait(lk, [&] { return pending_.load(std::memory_order_acquire) == 0; });
      job_ = nullptr;
    }
  }

 private:
  void loop(int index) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* job = nullptr;
      {
        std::unique_lock<std::mutex> lk(m_);
        start_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (index < job_n_) job = job_;
      }
      if (job != nullptr) {
        (*job)(index);
        if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
          std::lock_guard<std::mutex> lk(m_);
          done_.notify_one();
        }
      }
    }
  }
  std::mutex m_, run_m_;
  std::condition_variable start_, done_;
  const std::function<void(int)>* job_ = nullptr;
  int job_n_ = 0, n_workers_ = 0;
  uint64_t generation_ = 0;
  std::atomic<int> pending_{0};
  pid_t pid_;
};

// never destroyed (detached workers may be parked in it at exit); re-created in a forked child, whose threads are gone
inline WorkerPool& pool() {
  static std::mutex m;
  static WorkerPool* p = nullptr;
  std::lock_guard<std::mutex> lk(m);
  if (p == nullptr || p->pid() != getpid()) {
    const unsigned hc = std::thread::hardware_concurrency();
    int n = (int)std::min<unsigned>(hc == 0 ? 4 : hc, 16);
    // one process per GPU (torchrun sets LOCAL_WORLD_SIZE): share the physical cores (2 hardware threads each) between the ranks
    if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) {
      const int ranks = std::atoi(e);
      if (ranks > 1 && hc > 0) n = std::max(2, std::min(n, (int)(hc / 2) / ranks));
    }
    if (const char* e = std::getenv("CHG_PACK_THREADS")) n = std::max(1, std::min(std::atoi(e), 64));
    p = new WorkerPool(n - 1);
  }
  return *p;
}

}  // namespace chg
