#include "shm_queue.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <cstring>
#include <ctime>
#include <random>

namespace glt {

namespace {
constexpr uint64_t kMagic = 0x474c5442323030ULL;  // "GLTB200"
constexpr uint64_t kAlign = 64;
enum BlockState : uint32_t { kWriting = 0, kReady = 1, kReading = 2, kConsumed = 3 };
inline uint64_t align_up(uint64_t x) { return (x + kAlign - 1) & ~(kAlign - 1); }
}  // namespace

struct ShmQueue::BlockHdr {
  uint64_t payload;  // payload bytes
  uint64_t total;    // header + payload, 64-B aligned
  uint32_t state;
  uint32_t wrap;  // 1: filler covering the tail of the ring
  uint8_t pad[kAlign - 24];
};


struct ShmQueue::Header {
  uint64_t magic;
  uint64_t ring_bytes;
  uint64_t max_msgs;
  uint64_t head;     // oldest unreleased block
  uint64_t tail;     // next allocation
  uint64_t cursor;   // next block to hand to a consumer
  uint64_t used;     // bytes between head and tail (incl. fillers)
  uint64_t live;     // allocated, not yet released (excl. fillers)
  uint64_t pending;  // allocated, not yet dequeued (excl. fillers)
  uint32_t closed;
  uint32_t _pad;
  pthread_mutex_t mu;
  pthread_cond_t space;
  pthread_cond_t data;
};

ShmBlock::~ShmBlock() {
  if (q_) q_->Release(off_);
}

void ShmQueue::Lock() const {
  int rc = pthread_mutex_lock(&hdr_->mu);
  if (rc == EOWNERDEAD) pthread_mutex_consistent(&hdr_->mu);  // a peer died holding the lock
}
void ShmQueue::Unlock() const { pthread_mutex_unlock(&hdr_->mu); }

std::shared_ptr<ShmQueue> ShmQueue::Create(size_t max_msgs, size_t buf_bytes) {
  std::shared_ptr<ShmQueue> q(new ShmQueue());
  std::random_device rd;
  static std::atomic<uint64_t> counter{0};
  q->name_ = "/glt_b200_" + std::to_string(getpid()) + "_" + std::to_string(counter++) + "_" +
             std::to_string(rd() & 0xffffff);
  q->owner_ = true;
  uint64_t ring = align_up(buf_bytes + (max_msgs + 1) * sizeof(BlockHdr));
  uint64_t hdr_bytes = align_up(sizeof(Header));
  q->map_bytes_ = hdr_bytes + ring;
  int fd = shm_open(q->name_.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) throw std::runtime_error(std::string("shm_open failed: ") + strerror(errno));
  if (ftruncate(fd, q->map_bytes_) != 0) {
    close(fd);
    shm_unlink(q->name_.c_str());
    throw std::runtime_error(std::string("ftruncate failed: ") + strerror(errno));
  }
  q->base_ = mmap(nullptr, q->map_bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (q->base_ == MAP_FAILED) {
    shm_unlink(q->name_.c_str());
    throw std::runtime_error(std::string("mmap failed: ") + strerror(errno));
  }
  q->hdr_ = reinterpret_cast<Header*>(q->base_);
  q->ring_ = reinterpret_cast<uint8_t*>(q->base_) + hdr_bytes;
  std::memset(q->hdr_, 0, sizeof(Header));
  q->hdr_->ring_bytes = ring;
  q->hdr_->max_msgs = max_msgs;
  pthread_mutexattr_t ma;
  pthread_mutexattr_init(&ma);
  pthread_mutexattr_setpshared(&ma, PTHREAD_PROCESS_SHARED);
  pthread_mutexattr_setrobust(&ma, PTHREAD_MUTEX_ROBUST);
  pthread_mutex_init(&q->hdr_->mu, &ma);
  pthread_mutexattr_destroy(&ma);
  pthread_condattr_t ca;
  pthread_condattr_init(&ca);
  pthread_condattr_setpshared(&ca, PTHREAD_PROCESS_SHARED);
  pthread_condattr_setclock(&ca, CLOCK_MONOTONIC);
  pthread_cond_init(&q->hdr_->space, &ca);
  pthread_cond_init(&q->hdr_->data, &ca);
  pthread_condattr_destroy(&ca);
  q->hdr_->magic = kMagic;
  return q;
}

std::shared_ptr<ShmQueue> ShmQueue::Attach(const std::string& name) {
  std::shared_ptr<ShmQueue> q(new ShmQueue());
  q->name_ = name;
  int fd = shm_open(name.c_str(), O_RDWR, 0600);
  if (fd < 0) throw std::runtime_error("shm_open(attach) failed for " + name + ": " + strerror(errno));
  struct stat st;
  fstat(fd, &st);
  q->map_bytes_ = st.st_size;
  q->base_ = mmap(nullptr, q->map_bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (q->base_ == MAP_FAILED) throw std::runtime_error("mmap(attach) failed");
  q->hdr_ = reinterpret_cast<Header*>(q->base_);
  if (q->hdr_->magic != kMagic) throw std::runtime_error("not a glt_b200 shm queue: " + name);
  q->ring_ = reinterpret_cast<uint8_t*>(q->base_) + align_up(sizeof(Header));
  return q;
}

ShmQueue::~ShmQueue() {
  if (base_ && base_ != MAP_FAILED) munmap(base_, map_bytes_);
  if (owner_) shm_unlink(name_.c_str());
}

size_t ShmQueue::capacity_bytes() const { return hdr_->ring_bytes; }
size_t ShmQueue::max_msgs() const { return hdr_->max_msgs; }
size_t ShmQueue::size() const {
  Lock();
  size_t n = hdr_->pending;
  Unlock();
  return n;
}
bool ShmQueue::closed() const { return hdr_->closed != 0; }

void ShmQueue::Close() {
  Lock();
  hdr_->closed = 1;
  pthread_cond_broadcast(&hdr_->space);
  pthread_cond_broadcast(&hdr_->data);
  Unlock();
}

void ShmQueue::Enqueue(size_t bytes, const std::function<void(void*)>& writer) {
  const uint64_t total = align_up(sizeof(BlockHdr) + bytes);
  Header* h = hdr_;
  if (total + sizeof(BlockHdr) > h->ring_bytes)
    throw std::runtime_error("message of " + std::to_string(bytes) + " B exceeds shm ring of " +
                             std::to_string(h->ring_bytes) + " B");
  Lock();
  uint64_t off = 0;
  while (true) {
    if (h->closed) { Unlock(); throw QueueClosedError("shm queue closed"); }
    bool ok = false;
    if (h->live < h->max_msgs) {
      if (h->used == 0) { h->head = h->tail = h->cursor = 0; }
      if (h->used == 0 || h->tail > h->head) {
        uint64_t end_space = h->ring_bytes - h->tail;
        if (end_space >= total) {
          off = h->tail; ok = true;
        } else if (h->head >= total && (h->used > 0)) {
          // filler over the ring tail, then wrap to offset 0
          if (end_space > 0) {
            BlockHdr* f = reinterpret_cast<BlockHdr*>(ring_ + h->tail);
            f->payload = 0; f->total = end_space; f->wrap = 1; f->state = kReady;
            h->used += end_space;
          }
          h->tail = 0;
          off = 0; ok = true;
        } else if (h->used == 0) {
          off = 0; ok = true;  // empty ring always fits (checked above)
        }
      } else if (h->tail < h->head) {
        if (h->head - h->tail >= total) { off = h->tail; ok = true; }
      }
    }
    if (ok) break;
    pthread_cond_wait(&h->space, &h->mu);
  }
  BlockHdr* b = reinterpret_cast<BlockHdr*>(ring_ + off);
  b->payload = bytes; b->total = total; b->wrap = 0; b->state = kWriting;
  h->tail = off + total;
  if (h->tail == h->ring_bytes) h->tail = 0;
  h->used += total;
  h->live++;
  h->pending++;
  Unlock();
  try {
    writer(reinterpret_cast<uint8_t*>(b) + sizeof(BlockHdr));
  } catch (...) {
    Lock(); b->payload = 0; b->state = kReady; pthread_cond_broadcast(&h->data); Unlock();
    throw;
  }
  Lock();
  b->state = kReady;
  pthread_cond_broadcast(&h->data);
  Unlock();
}

std::shared_ptr<ShmBlock> ShmQueue::Dequeue(int64_t timeout_ms) {
  Header* h = hdr_;
  timespec deadline;
  if (timeout_ms > 0) {
    clock_gettime(CLOCK_MONOTONIC, &deadline);
    deadline.tv_sec += timeout_ms / 1000;
    deadline.tv_nsec += (timeout_ms % 1000) * 1000000L;
    if (deadline.tv_nsec >= 1000000000L) { deadline.tv_sec++; deadline.tv_nsec -= 1000000000L; }
  }
  Lock();
  while (true) {
    if (h->pending > 0) {
      BlockHdr* b = reinterpret_cast<BlockHdr*>(ring_ + h->cursor);
      if (b->wrap) {  // step over the filler
        b->state = kConsumed;
        h->cursor = 0;
        b = reinterpret_cast<BlockHdr*>(ring_);
      }
      if (b->state == kReady) {
        uint64_t off = h->cursor;
        b->state = kReading;
        h->cursor = off + b->total;
        if (h->cursor == h->ring_bytes) h->cursor = 0;
        h->pending--;
        Unlock();
        return std::make_shared<ShmBlock>(shared_from_this(), off,
                                          reinterpret_cast<uint8_t*>(b) + sizeof(BlockHdr),
                                          b->payload);
      }
    }
    if (h->closed) { Unlock(); throw QueueClosedError("shm queue closed"); }
    if (timeout_ms > 0) {
      int rc = pthread_cond_timedwait(&h->data, &h->mu, &deadline);
      if (rc == ETIMEDOUT) { Unlock(); throw QueueTimeoutError("shm queue dequeue timed out"); }
      if (rc == EOWNERDEAD) pthread_mutex_consistent(&h->mu);
    } else {
      pthread_cond_wait(&h->data, &h->mu);
    }
  }
}

void ShmQueue::Release(uint64_t off) {
  Header* h = hdr_;
  Lock();
  reinterpret_cast<BlockHdr*>(ring_ + off)->state = kConsumed;
  h->live--;
  // in-order reclamation from the head
  while (h->used > 0) {
    BlockHdr* b = reinterpret_cast<BlockHdr*>(ring_ + h->head);
    if (b->state != kConsumed) break;
    h->used -= b->total;
    h->head += b->total;
    if (h->head == h->ring_bytes) h->head = 0;
  }
  pthread_cond_broadcast(&h->space);
  Unlock();
}

}  // namespace glt
