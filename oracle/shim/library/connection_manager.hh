// oracle/shim: in-process stand-in for the ibverbs wrapper (rdma-library/library/
// {context,queue_pair,memory_region,connection_manager,configuration}.hh).
// TEST INFRASTRUCTURE ONLY.  It lets the reference's own hot-path headers
// (src/hnsw/*.hh, src/rdma/*.hh, src/compute_thread.hh, src/shared_context.hh) compile
// UNMODIFIED without libibverbs: a "memory node" is a plain buffer in this process, a
// one-sided READ/WRITE is a memcpy, CAS/FAA are local atomics, and every signaled
// request yields exactly one completion carrying its wr_id.  Only the surface listed in
// SURVEY.md App. E is provided; semantics follow rdma-library/library/queue_pair.hh:50-114
// and context.hh:53-71.
#pragma once
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

#include <library/types.hh>
#include <library/utils.hh>

enum ibv_wr_opcode { IBV_WR_RDMA_WRITE = 0, IBV_WR_SEND = 2, IBV_WR_RDMA_READ = 4 };
struct ibv_wc {
  u64 wr_id{};
  int status{};
};
// completion queue = list of wr_ids of finished signaled requests
struct ibv_cq {
  std::mutex mu;
  std::vector<u64> done;
  bool shared{false};  // more than one polling thread
};

namespace configuration {
class Configuration {
public:
  i32 max_poll_cqes{16};
  i32 max_send_queue_wr{1024};
  i32 max_recv_queue_wr{1024};
};
}  // namespace configuration

struct MemoryRegionToken {
  u64 address;
  u32 lkey;
  u32 rkey;
};
using MRT = u_ptr<MemoryRegionToken>;
using MemoryRegionTokens = vec<MRT>;

class Context {
public:
  using Configuration = configuration::Configuration;
  explicit Context(Configuration& config) : config_(config) {}
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;

  Configuration& get_config() const { return config_; }
  ibv_cq* get_send_cq() { return &send_cq_; }
  ibv_cq* get_receive_cq() { return &recv_cq_; }
  u16 get_lid() const { return 0; }

  static i32 poll_send_cq(ibv_wc*, i32 max_cqes, ibv_cq* cq, const func<void(u64)>& id_handler) {
    i32 n = 0;
    if (cq->shared) {
      std::lock_guard<std::mutex> g(cq->mu);
      n = drain(cq, max_cqes, id_handler);
    } else {
      n = drain(cq, max_cqes, id_handler);
    }
    return n;
  }

private:
  static i32 drain(ibv_cq* cq, i32 max_cqes, const func<void(u64)>& id_handler) {
    i32 n = 0;
    // completions are delivered in posting order, as on a reliable-connected QP
    size_t take = std::min<size_t>(cq->done.size(), static_cast<size_t>(max_cqes));
    for (size_t i = 0; i < take; ++i, ++n) id_handler(cq->done[i]);
    cq->done.erase(cq->done.begin(), cq->done.begin() + take);
    return n;
  }

  Configuration& config_;
  ibv_cq send_cq_;
  ibv_cq recv_cq_;
};

class QueuePair {
public:
  QueuePair(Context*, ibv_cq* send_cq, ibv_cq*) : cq_(send_cq) {}

  // READ: remote -> local; WRITE: local -> remote (queue_pair.hh:80-89)
  void post_send(u64 laddr, u32 size, u32, ibv_wr_opcode opcode, bool signaled, bool, MemoryRegionToken* token,
                 u64 remote_offset, u64 local_offset, u64 wr_id) {
    byte_t* remote = reinterpret_cast<byte_t*>(token->address + remote_offset);
    byte_t* local = reinterpret_cast<byte_t*>(laddr + local_offset);
    if (opcode == IBV_WR_RDMA_READ) {
      std::memcpy(local, remote, size);
    } else {
      std::memcpy(remote, local, size);
    }
    complete(signaled, wr_id);
  }

  // inlined WRITE (queue_pair.hh:50-56)
  void post_send_inlined(const void* address, u32 size, ibv_wr_opcode opcode, bool signaled = true,
                         MemoryRegionToken* token = nullptr, u64 remote_offset = 0, u64 wr_id = 0) {
    lib_assert(opcode == IBV_WR_RDMA_WRITE && token != nullptr, "oracle shim: only one-sided WRITEs are inlined");
    byte_t* remote = reinterpret_cast<byte_t*>(token->address + remote_offset);
    if (size == 1) {  // unlock bytes race with CAS on the same word
      __atomic_store_n(remote, *static_cast<const byte_t*>(address), __ATOMIC_RELEASE);
    } else {
      std::memcpy(remote, address, size);
    }
    complete(signaled, wr_id);
  }

  // old value lands in *laddr (queue_pair.hh:99-114)
  void post_CAS(u64 laddr, u32, MemoryRegionToken* token, u64 remote_offset, u64 compare, u64 swap,
                bool signaled = true, u64 wr_id = 0) {
    u64* remote = reinterpret_cast<u64*>(token->address + remote_offset);
    u64 expected = compare;
    __atomic_compare_exchange_n(remote, &expected, swap, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
    *reinterpret_cast<u64*>(laddr) = expected;  // == compare on success, current value on failure
    complete(signaled, wr_id);
  }

  void post_FAA(u64 laddr, u32, MemoryRegionToken* token, u64 remote_offset, u64 to_add, bool signaled = true,
                u64 wr_id = 0) {
    u64* remote = reinterpret_cast<u64*>(token->address + remote_offset);
    *reinterpret_cast<u64*>(laddr) = __atomic_fetch_add(remote, to_add, __ATOMIC_ACQ_REL);
    complete(signaled, wr_id);
  }

private:
  void complete(bool signaled, u64 wr_id) {
    if (!signaled) return;
    if (cq_->shared) {
      std::lock_guard<std::mutex> g(cq_->mu);
      cq_->done.push_back(wr_id);
    } else {
      cq_->done.push_back(wr_id);
    }
  }
  ibv_cq* cq_;
};

using QP = u_ptr<QueuePair>;
using QPs = vec<QP>;

class LocalMemoryRegion {
public:
  LocalMemoryRegion(Context&, void*, size_t) {}
  u32 get_lkey() const { return 0; }
};

// one entry in server_qps per memory node; nothing is actually connected
class ClientConnectionManager {
public:
  explicit ClientConnectionManager(u32 num_memory_nodes) : server_qps(num_memory_nodes) {}
  const bool is_initiator{true};
  u32 client_id{0};
  u32 num_total_clients{1};
  QPs server_qps;
};
