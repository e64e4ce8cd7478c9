// Native unit test of the shared-memory ring (no Python, no torch).
//   threads : 3 producers x 2 consumers inside one process -- every message exactly once
//   fork    : 2 sender + 2 receiver child processes attached by name (reference idea:
//             test/cpp/test_shm_queue.cu:72-145 forks 4 children)
//   timeout : Dequeue on an empty ring throws QueueTimeoutError; Close() wakes a blocked waiter
// Build + run: scripts/run_cpp_ut.sh  (also builds a -fsanitize=thread variant of the threads test)
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "cpu/shm_queue.h"

using glt::ShmQueue;

#define CHECK(cond)                                                              \
  do {                                                                           \
    if (!(cond)) {                                                               \
      std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
      std::exit(1);                                                              \
    }                                                                            \
  } while (0)

struct Msg {
  uint32_t producer, seq, len, pad;
};

static void produce(const std::shared_ptr<ShmQueue>& q, uint32_t id, uint32_t n) {
  std::mt19937 rng(1234 + id);
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t len = 16 + rng() % 2000;
    q->Enqueue(sizeof(Msg) + len, [&](void* p) {
      Msg m{id, i, len, 0};
      std::memcpy(p, &m, sizeof(m));
      std::memset(static_cast<char*>(p) + sizeof(m), static_cast<int>((id * 131 + i) & 0xff), len);
    });
  }
}

// returns number of messages consumed; verifies payloads and per-producer ordering of what it saw
static uint64_t consume(const std::shared_ptr<ShmQueue>& q, std::vector<std::atomic<uint32_t>>* seen, int producers,
                        uint32_t per_producer) {
  uint64_t got = 0;
  std::vector<int64_t> last(producers, -1);
  for (;;) {
    std::shared_ptr<glt::ShmBlock> b;
    try {
      b = q->Dequeue(2000);
    } catch (const glt::QueueTimeoutError&) {
      break;  // producers are done
    } catch (const glt::QueueClosedError&) {
      break;
    }
    Msg m;
    std::memcpy(&m, b->data(), sizeof(m));
    CHECK(m.producer < static_cast<uint32_t>(producers) && m.seq < per_producer);
    CHECK(b->size() >= sizeof(Msg) + m.len);
    const unsigned char* pay = static_cast<const unsigned char*>(b->data()) + sizeof(Msg);
    const unsigned char want = static_cast<unsigned char>((m.producer * 131 + m.seq) & 0xff);
    CHECK(pay[0] == want && pay[m.len - 1] == want && pay[m.len / 2] == want);
    CHECK(static_cast<int64_t>(m.seq) > last[m.producer]);  // FIFO per producer within one consumer
    last[m.producer] = m.seq;
    if (seen) CHECK((*seen)[m.producer * per_producer + m.seq].fetch_add(1) == 0);  // exactly once
    ++got;
  }
  return got;
}

static void test_threads() {
  const int P = 3, C = 2;
  const uint32_t N = 4000;
  auto q = ShmQueue::Create(64, 1 << 20);
  std::vector<std::atomic<uint32_t>> seen(P * N);
  for (auto& s : seen) s.store(0);
  std::vector<std::thread> th;
  std::atomic<uint64_t> total{0};
  for (int c = 0; c < C; ++c) th.emplace_back([&] { total += consume(q, &seen, P, N); });
  for (int p = 0; p < P; ++p) th.emplace_back([&, p] { produce(q, p, N); });
  for (auto& t : th) t.join();
  CHECK(total.load() == static_cast<uint64_t>(P) * N);
  for (auto& s : seen) CHECK(s.load() == 1);
  std::printf("threads: %d producers x %d consumers, %llu messages exactly once\n", P, C,
              static_cast<unsigned long long>(total.load()));
}

static void test_fork() {
  const int P = 2, C = 2;
  const uint32_t N = 3000;
  auto q = ShmQueue::Create(32, 1 << 19);
  const std::string name = q->name();
  int pipes[C][2];
  std::vector<pid_t> kids;
  for (int c = 0; c < C; ++c) {
    CHECK(pipe(pipes[c]) == 0);
    pid_t pid = fork();
    CHECK(pid >= 0);
    if (pid == 0) {
      auto qq = ShmQueue::Attach(name);
      const uint64_t got = consume(qq, nullptr, P, N);
      CHECK(write(pipes[c][1], &got, sizeof(got)) == sizeof(got));
      _exit(0);
    }
    kids.push_back(pid);
  }
  for (int p = 0; p < P; ++p) {
    pid_t pid = fork();
    CHECK(pid >= 0);
    if (pid == 0) {
      auto qq = ShmQueue::Attach(name);
      produce(qq, p, N);
      _exit(0);
    }
    kids.push_back(pid);
  }
  uint64_t total = 0;
  for (int c = 0; c < C; ++c) {
    uint64_t got = 0;
    CHECK(read(pipes[c][0], &got, sizeof(got)) == sizeof(got));
    total += got;
  }
  for (pid_t k : kids) {
    int st = 0;
    CHECK(waitpid(k, &st, 0) == k);
    CHECK(WIFEXITED(st) && WEXITSTATUS(st) == 0);
  }
  CHECK(total == static_cast<uint64_t>(P) * N);
  std::printf("fork: %d sender + %d receiver processes, %llu messages, all children exited 0\n", P, C,
              static_cast<unsigned long long>(total));
}

static void test_timeout_and_close() {
  auto q = ShmQueue::Create(4, 1 << 12);
  bool timed_out = false;
  const auto t0 = std::chrono::steady_clock::now();
  try {
    q->Dequeue(100);
  } catch (const glt::QueueTimeoutError&) {
    timed_out = true;
  }
  const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
  CHECK(timed_out && ms >= 90 && ms < 2000);
  std::atomic<bool> woke{false};
  std::thread waiter([&] {
    try {
      q->Dequeue(0);
    } catch (const glt::QueueClosedError&) {
      woke = true;
    }
  });
  std::this_thread::sleep_for(std::chrono::milliseconds(100));
  q->Close();
  waiter.join();
  CHECK(woke.load() && q->closed());
  std::printf("timeout after %lld ms, Close() woke the blocked consumer\n", static_cast<long long>(ms));
}

int main(int argc, char** argv) {
  const bool threads_only = argc > 1 && std::strcmp(argv[1], "--threads-only") == 0;
  test_threads();
  test_timeout_and_close();
  if (!threads_only) test_fork();
  std::printf("SHM_QUEUE_OK\n");
  return 0;
}
