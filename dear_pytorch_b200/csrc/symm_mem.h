// symm_mem.h — symmetric (peer-mapped) memory arenas for the DeAR runtime.
//
// An arena is an allocation of identical size on every rank whose peers'
// copies are mapped into the local address space, so a kernel can load from /
// store to any rank's copy over NVLink.  The first kSignalPadBytes of every
// arena are the rank's signal pad (flags written by peers); the rest is data.
//
// Providers:
//   HOST_SHM  POSIX shared memory; CPU-only boxes / gloo plumbing tests.
//   CUDA_IPC  cudaMalloc + cudaIpcGetMemHandle (P2P over NVLink, no multicast).
//   CUDA_VMM  cuMemCreate + POSIX-fd handles passed over AF_UNIX (SCM_RIGHTS),
//             optionally bound to an NVLS multicast object (multimem.* path).
//   EXTERNAL  pointers supplied by the caller (e.g. torch symmetric memory).
//
// Replaces the reference's MPI + ncclCommInitRank bootstrap
// (common/comm_core/src/communicator.cpp:43-66): rendezvous goes through the
// c10d Store that torchrun already provides, no MPI.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include <torch/csrc/distributed/c10d/Store.hpp>
#include "dear_common.h"

namespace dear {

enum class Provider : int { HOST_SHM = 0, CUDA_IPC = 1, CUDA_VMM = 2, EXTERNAL = 3 };

struct ArenaOptions {
  Provider provider = Provider::CUDA_IPC;
  bool want_multicast = false;
  int device = -1;           // CUDA device index; -1 for host
  double timeout_s = 120.0;  // rendezvous timeout
};

class SymmArena {
 public:
  // Collective: every rank must call with the same `data_bytes` and `key`.
  static std::shared_ptr<SymmArena> create(size_t data_bytes, int rank, int world,
                                           const c10::intrusive_ptr<c10d::Store>& store,
                                           const std::string& key, const ArenaOptions& opt);
  // Wrap memory somebody else made symmetric. `bases[r]` is rank r's arena base
  // (signal pad first).  `mc_base` may be 0.
  static std::shared_ptr<SymmArena> from_external(const std::vector<uint64_t>& bases,
                                                  uint64_t mc_base, size_t data_bytes, int rank,
                                                  int world, int device);
  ~SymmArena();

  int rank() const { return rank_; }
  int world() const { return world_; }
  int device() const { return device_; }
  bool is_cuda() const { return provider_ != Provider::HOST_SHM && device_ >= 0; }
  Provider provider() const { return provider_; }
  size_t data_bytes() const { return data_bytes_; }
  bool has_multicast() const { return mc_base_ != nullptr; }

  char* data(int r) const { return bases_[r] + kSignalPadBytes; }
  char* local_data() const { return data(rank_); }
  char* mc_data() const { return mc_base_ ? mc_base_ + kSignalPadBytes : nullptr; }
  void* sig(int r) const { return bases_[r]; }
  uint32_t* ctrl() const { return ctrl_; }

  PeerTable data_table(size_t byte_off) const;
  PeerTable sig_table() const;

  // Cross-rank barrier through the store (host side).
  void store_barrier(const std::string& tag);

 private:
  SymmArena() = default;
  void init_host_shm(const std::string& key);
  void init_cuda_ipc(const std::string& key);
  void init_cuda_vmm(const std::string& key, bool want_mc);
  void alloc_ctrl();

  Provider provider_ = Provider::HOST_SHM;
  int rank_ = 0, world_ = 1, device_ = -1;
  size_t data_bytes_ = 0, total_bytes_ = 0, mapped_bytes_ = 0;
  std::vector<char*> bases_;
  char* mc_base_ = nullptr;
  uint32_t* ctrl_ = nullptr;
  c10::intrusive_ptr<c10d::Store> store_;
  std::string key_;
  double timeout_s_ = 120.0;
  int barrier_seq_ = 0;
  // provider-private state
  std::vector<std::string> shm_names_;
  std::vector<uint64_t> vmm_handles_;   // CUmemGenericAllocationHandle per rank
  uint64_t mc_handle_ = 0;
  bool owns_ = true;
};

// Host-mapped status word (one per process): kernels flag spin-wait timeouts here.
uint32_t* status_word_host();
uint32_t* status_word_device();   // device alias of the same word (nullptr on CPU-only boxes)
bool cuda_runtime_usable();

}  // namespace dear
