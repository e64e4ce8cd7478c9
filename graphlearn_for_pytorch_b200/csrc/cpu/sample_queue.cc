#include "sample_queue.h"

#include <cuda_runtime.h>

#include <cstring>

namespace glt {

namespace {
constexpr size_t kDataAlign = 64;
inline size_t align_up(size_t x) { return (x + kDataAlign - 1) & ~(kDataAlign - 1); }

// record: u32 key_len | key | i32 dtype | u32 ndim | i64 shape[ndim] | u64 nbytes | pad | data
size_t record_header_bytes(const std::string& key, const torch::Tensor& t) {
  return 4 + key.size() + 4 + 4 + 8 * t.dim() + 8;
}
}  // namespace

size_t tensor_map_bytes(const TensorMap& m) {
  size_t n = 8;
  for (const auto& kv : m) {
    n = align_up(n + record_header_bytes(kv.first, kv.second));
    n = align_up(n + kv.second.numel() * kv.second.element_size());
  }
  return n;
}

void tensor_map_write(const TensorMap& m, void* dst) {
  uint8_t* base = reinterpret_cast<uint8_t*>(dst);
  size_t off = 0;
  uint64_t count = m.size();
  std::memcpy(base, &count, 8);
  off = 8;
  bool any_cuda = false;
  for (const auto& kv : m) {
    const torch::Tensor t = kv.second.contiguous();
    uint32_t klen = kv.first.size();
    std::memcpy(base + off, &klen, 4); off += 4;
    std::memcpy(base + off, kv.first.data(), klen); off += klen;
    int32_t dt = static_cast<int32_t>(t.scalar_type());
    std::memcpy(base + off, &dt, 4); off += 4;
    uint32_t nd = t.dim();
    std::memcpy(base + off, &nd, 4); off += 4;
    for (uint32_t i = 0; i < nd; ++i) { int64_t s = t.size(i); std::memcpy(base + off, &s, 8); off += 8; }
    uint64_t nbytes = t.numel() * t.element_size();
    std::memcpy(base + off, &nbytes, 8); off += 8;
    off = align_up(off);
    if (nbytes) {
      if (t.is_cuda()) {
        // device -> shm block directly (fast when the ring is pinned)
        cudaMemcpyAsync(base + off, t.data_ptr(), nbytes, cudaMemcpyDeviceToHost, 0);
        any_cuda = true;
      } else {
        std::memcpy(base + off, t.data_ptr(), nbytes);
      }
    }
    off = align_up(off + nbytes);
  }
  if (any_cuda) cudaStreamSynchronize(0);
}

TensorMap tensor_map_read(std::shared_ptr<ShmBlock> block) {
  TensorMap out;
  uint8_t* base = reinterpret_cast<uint8_t*>(block->data());
  if (block->size() < 8) return out;
  uint64_t count;
  std::memcpy(&count, base, 8);
  size_t off = 8;
  for (uint64_t r = 0; r < count; ++r) {
    uint32_t klen; std::memcpy(&klen, base + off, 4); off += 4;
    std::string key(reinterpret_cast<char*>(base + off), klen); off += klen;
    int32_t dt; std::memcpy(&dt, base + off, 4); off += 4;
    uint32_t nd; std::memcpy(&nd, base + off, 4); off += 4;
    std::vector<int64_t> shape(nd);
    for (uint32_t i = 0; i < nd; ++i) { std::memcpy(&shape[i], base + off, 8); off += 8; }
    uint64_t nbytes; std::memcpy(&nbytes, base + off, 8); off += 8;
    off = align_up(off);
    auto opts = torch::TensorOptions().dtype(static_cast<torch::ScalarType>(dt)).device(torch::kCPU);
    // zero-copy view; every tensor keeps the block (and thus the ring space) alive
    if (nbytes == 0) {
      // no storage to alias: a from_blob deleter would never fire and pin the ring block
      out[key] = torch::empty(shape, opts);
    } else {
      std::shared_ptr<ShmBlock> keep = block;
      out[key] = torch::from_blob(base + off, shape, [keep](void*) mutable { keep.reset(); }, opts);
    }
    off = align_up(off + nbytes);
  }
  return out;
}

void SampleQueue::send(const TensorMap& m) {
  const size_t bytes = tensor_map_bytes(m);
  q_->Enqueue(bytes, [&](void* dst) { tensor_map_write(m, dst); });
}

TensorMap SampleQueue::recv(int64_t timeout_ms) { return tensor_map_read(q_->Dequeue(timeout_ms)); }

void SampleQueue::pin_memory() {
  if (q_->pinned()) return;
  cudaError_t e = cudaHostRegister(q_->base(), q_->mapped_bytes(), cudaHostRegisterPortable);
  if (e == cudaSuccess) q_->set_pinned(true);
  else cudaGetLastError();  // not fatal: the channel still works unpinned
}

}  // namespace glt
