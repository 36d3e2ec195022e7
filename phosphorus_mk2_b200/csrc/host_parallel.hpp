// host_parallel.hpp — the two fork-join helpers the host-side build and re-pack use (std::thread only).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace phos {

// host threads to use: hardware_concurrency, PHOS_THREADS overrides
inline int worker_count() {
  int t = (int)std::thread::hardware_concurrency();
  if (const char* e = std::getenv("PHOS_THREADS")) t = std::atoi(e);
  return std::max(1, std::min(t, 64));
}

// fn(chunk_begin, chunk_end, chunk_index) over [0, n) in chunks of `grain`; returns when all are done
template <class F>
void parallel_chunks(size_t n, size_t grain, int threads, F&& fn) {
  const size_t chunks = (n + grain - 1) / grain;
  if (threads <= 1 || chunks <= 1) {
    for (size_t c = 0; c < chunks; ++c) fn(c * grain, std::min(n, (c + 1) * grain), c);
    return;
  }
  std::atomic<size_t> next{0};
  auto body = [&]() {
    for (size_t c; (c = next.fetch_add(1)) < chunks;) fn(c * grain, std::min(n, (c + 1) * grain), c);
  };
  std::vector<std::thread> pool;
  const int extra = (int)std::min<size_t>((size_t)threads, chunks) - 1;
  for (int t = 0; t < extra; ++t) pool.emplace_back(body);
  body();
  for (auto& t : pool) t.join();
}

// A bag of tasks that may add more tasks; run() returns when the bag is empty and every worker idle.
class TaskBag {
 public:
  void add(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> g(m_);
      q_.push_back(std::move(f));
    }
    cv_.notify_one();
  }
  void run(int threads) {
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back([this] { work(); });
    work();
    for (auto& t : pool) t.join();
  }

 private:
  void work() {
    std::unique_lock<std::mutex> g(m_);
    for (;;) {
      if (!q_.empty()) {
        std::function<void()> f = std::move(q_.front());
        q_.pop_front();
        ++busy_;
        g.unlock();
        f();
        g.lock();
        --busy_;
        if (q_.empty() && busy_ == 0) cv_.notify_all();
        continue;
      }
      if (busy_ == 0) return;  // nothing queued, nobody who could queue more
      cv_.wait(g);
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> q_;
  int busy_ = 0;
};

}  // namespace phos
