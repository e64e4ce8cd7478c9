// Host-side parallel loop of the CPU operators.
//
// at::parallel_for only fans out inside translation units compiled with -fopenmp; linking a second OpenMP runtime
// next to the one bundled with the torch wheel is what this avoids.  The loop is split into at most
// at::get_num_threads() contiguous chunks of at least `grain` items; chunk 0 runs on the calling thread, the others
// on short-lived std::threads (no persistent pool: nothing to re-create after fork(), which the sampling producers
// and channel tests rely on).  The first exception of any chunk is re-thrown on the caller.
#pragma once
#include <ATen/Parallel.h>

#include <algorithm>
#include <cstdint>
#include <exception>
#include <mutex>
#include <system_error>
#include <thread>
#include <vector>

namespace glt {

template <class F>
inline void parallel_for(int64_t begin, int64_t end, int64_t grain, const F& f) {
  const int64_t n = end - begin;
  if (n <= 0) return;
  grain = std::max<int64_t>(grain, 1);
  int64_t nt = std::min<int64_t>(at::get_num_threads(), (n + grain - 1) / grain);
  if (nt <= 1 || at::in_parallel_region()) {
    f(begin, end);
    return;
  }
  const int64_t chunk = (n + nt - 1) / nt;
  nt = (n + chunk - 1) / chunk;
  std::exception_ptr err;
  std::mutex err_mu;
  auto run = [&](int64_t lo, int64_t hi) {
    try {
      f(lo, hi);
    } catch (...) {
      std::lock_guard<std::mutex> g(err_mu);
      if (!err) err = std::current_exception();
    }
  };
  std::vector<std::thread> workers;
  workers.reserve(nt - 1);
  int64_t inline_from = nt;          // chunks [inline_from, nt) run on the caller when no more threads can be started
  for (int64_t t = 1; t < nt; ++t) {
    const int64_t lo = begin + t * chunk, hi = std::min(end, lo + chunk);
    try {
      workers.emplace_back(run, lo, hi);
    } catch (const std::system_error&) {   // thread limit reached: finish the rest here instead of terminating
      inline_from = t;
      break;
    }
  }
  run(begin, std::min(end, begin + chunk));
  for (int64_t t = inline_from; t < nt; ++t) run(begin + t * chunk, std::min(end, begin + (t + 1) * chunk));
  for (auto& w : workers) w.join();
  if (err) std::rethrow_exception(err);
}

}  // namespace glt
