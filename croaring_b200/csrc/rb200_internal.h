// rb200_internal.h — hooks shared between the translation units of libroaring_b200.so
// (rb200_host.cu owns the context; rb200_shard.cu owns NCCL and the host-side blob algebra).
// Internal: nothing here crosses the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <string>

struct rb200_comm;

namespace rb200 {

// record the message rb200_last_error() returns (calling thread's view)
void set_error(const std::string &msg);
// run fn(i), i < n, on the library's host worker pool (blocking)
void parallel_for(size_t n, const std::function<void(size_t)> &fn);
// in-place ncclAllReduce(sum) of count u32 / u64 words on stream s; no-op for a 1-rank communicator
bool comm_allreduce_sum(rb200_comm *c, void *buf, size_t count, bool u64, cudaStream_t s, std::string &err);

}  // namespace rb200
