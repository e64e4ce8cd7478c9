// TensorMap <-> shared-memory block serialisation and the SampleQueue built on it.
// Capability parity with the reference's TensorMapSerializer / SampleQueue
// (include/tensor_map.h:24-52, csrc/tensor_map.cc:71-169, include/sample_queue.h:26-50):
// CUDA tensors are copied D2H straight into the ring block, and loads are
// zero-copy `from_blob` views whose deleter recycles the block.
#pragma once
#include <torch/extension.h>

#include <map>
#include <memory>
#include <string>

#include "shm_queue.h"

namespace glt {

using TensorMap = std::map<std::string, torch::Tensor>;

size_t tensor_map_bytes(const TensorMap& m);
void tensor_map_write(const TensorMap& m, void* dst);
TensorMap tensor_map_read(std::shared_ptr<ShmBlock> block);

class SampleQueue {
 public:
  SampleQueue(size_t max_msgs, size_t buf_bytes) : q_(ShmQueue::Create(max_msgs, buf_bytes)) {}
  explicit SampleQueue(const std::string& name) : q_(ShmQueue::Attach(name)) {}
  const std::string& name() const { return q_->name(); }
  void send(const TensorMap& m);
  TensorMap recv(int64_t timeout_ms);
  bool empty() const { return q_->empty(); }
  size_t size() const { return q_->size(); }
  void close() { q_->Close(); }
  void pin_memory();
  std::shared_ptr<ShmQueue> raw() const { return q_; }

 private:
  std::shared_ptr<ShmQueue> q_;
};

}  // namespace glt
