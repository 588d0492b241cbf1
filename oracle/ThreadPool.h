// ORACLE / TEST INFRASTRUCTURE ONLY. Minimal persistent thread pool with a chunked parallel_for; stands in
// for the dispenso pool the reference's BackendFast owns (reference baspacho/baspacho/MatOpsFast.cpp:32-46).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace oracle {

class ThreadPool {
 public:
  explicit ThreadPool(int numThreads) : n_(numThreads < 1 ? 1 : numThreads) {
    for (int i = 1; i < n_; i++) workers_.emplace_back([this, i] { workerLoop(i); });
  }
  ~ThreadPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int numThreads() const { return n_; }

  // fn(chunkBegin, chunkEnd, threadSlot) over [begin,end) in chunks; the caller participates as slot 0
  void parallelFor(int64_t begin, int64_t end, int64_t chunk, const std::function<void(int64_t, int64_t, int)>& fn) {
    if (end <= begin) return;
    if (n_ == 1 || end - begin <= chunk) {
      fn(begin, end, 0);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn;
      next_.store(begin);
      end_ = end;
      chunk_ = chunk;
      pending_ = n_ - 1;
      gen_++;
    }
    cv_.notify_all();
    run(0);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void run(int slot) {
    for (;;) {
      int64_t b = next_.fetch_add(chunk_);
      if (b >= end_) break;
      (*fn_)(b, std::min(b + chunk_, end_), slot);
    }
  }
  void workerLoop(int slot) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      run(slot);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }

  int n_;
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int64_t, int64_t, int)>* fn_ = nullptr;
  std::atomic<int64_t> next_{0};
  int64_t end_ = 0, chunk_ = 1;
  int pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

}  // namespace oracle
