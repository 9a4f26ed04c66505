// Which compute lane a host thread gets.  Free of CUDA types so that tests/hostcheck can stress it on
// the CPU.  A thread that finds every lane taken waits for ANY lane to be released (the first
// version waited for lane 0 only, which left other lanes idle while the extra host thread of the
// end-to-end arm stood in line).
#pragma once
#include <condition_variable>
#include <mutex>

namespace zkb {

template <int N>
class LanePool {
 public:
  // blocks until one of the first `active` lanes is free, marks it taken and returns its index
  int acquire(int active) {
    if (active < 1) active = 1;
    if (active > N) active = N;
    std::unique_lock<std::mutex> lk(m_);
    int got = -1;
    cv_.wait(lk, [&] {
      for (int i = 0; i < active; i++)
        if (!busy_[i]) { got = i; return true; }
      return false;
    });
    busy_[got] = true;
    return got;
  }
  void release(int lane) {
    {
      std::lock_guard<std::mutex> lk(m_);
      busy_[lane] = false;
    }
    cv_.notify_one();
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  bool busy_[N] = {};
};

}  // namespace zkb
