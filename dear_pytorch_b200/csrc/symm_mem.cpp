// symm_mem.cpp — see symm_mem.h.
#include "symm_mem.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>
#include <poll.h>

#include <chrono>
#include <cstring>
#include "dear_msg.h"
#include <stdexcept>
#include <thread>

namespace dear {

#define DEAR_CHECK(cond, msg)                                                        \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      dear::Msg _oss;                                                                     \
      _oss << "dear: " << msg << " (" << __FILE__ << ":" << __LINE__ << ")";         \
      throw std::runtime_error(_oss.str());                                          \
    }                                                                                \
  } while (0)

#define DEAR_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      dear::Msg _oss;                                                                     \
      _oss << "dear: CUDA error '" << cudaGetErrorString(_e) << "' in " #expr " ("   \
           << __FILE__ << ":" << __LINE__ << ")";                                    \
      throw std::runtime_error(_oss.str());                                          \
    }                                                                                \
  } while (0)

// ---------------------------------------------------------------------------
// driver API through cudaGetDriverEntryPoint (the extension never links libcuda,
// so it imports on CPU-only boxes)
// ---------------------------------------------------------------------------
namespace {

template <typename Fn>
Fn driver_fn(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    throw std::runtime_error(std::string("dear: driver entry point not available: ") + name);
  }
  return reinterpret_cast<Fn>(fn);
}

#define DEAR_CU(call)                                                                \
  do {                                                                               \
    CUresult _r = (call);                                                            \
    if (_r != CUDA_SUCCESS) {                                                        \
      dear::Msg _oss;                                                                     \
      _oss << "dear: driver error " << int(_r) << " in " #call " (" << __FILE__      \
           << ":" << __LINE__ << ")";                                                \
      throw std::runtime_error(_oss.str());                                          \
    }                                                                                \
  } while (0)

struct DriverApi {
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
  CUresult (*DeviceGet)(CUdevice*, int);

  static const DriverApi& get() {
    static DriverApi api = [] {
      DriverApi a;
      a.MemGetAllocationGranularity = driver_fn<decltype(a.MemGetAllocationGranularity)>("cuMemGetAllocationGranularity");
      a.MemCreate = driver_fn<decltype(a.MemCreate)>("cuMemCreate");
      a.MemRelease = driver_fn<decltype(a.MemRelease)>("cuMemRelease");
      a.MemExportToShareableHandle = driver_fn<decltype(a.MemExportToShareableHandle)>("cuMemExportToShareableHandle");
      a.MemImportFromShareableHandle = driver_fn<decltype(a.MemImportFromShareableHandle)>("cuMemImportFromShareableHandle");
      a.MemAddressReserve = driver_fn<decltype(a.MemAddressReserve)>("cuMemAddressReserve");
      a.MemAddressFree = driver_fn<decltype(a.MemAddressFree)>("cuMemAddressFree");
      a.MemMap = driver_fn<decltype(a.MemMap)>("cuMemMap");
      a.MemUnmap = driver_fn<decltype(a.MemUnmap)>("cuMemUnmap");
      a.MemSetAccess = driver_fn<decltype(a.MemSetAccess)>("cuMemSetAccess");
      a.MulticastCreate = driver_fn<decltype(a.MulticastCreate)>("cuMulticastCreate");
      a.MulticastAddDevice = driver_fn<decltype(a.MulticastAddDevice)>("cuMulticastAddDevice");
      a.MulticastBindMem = driver_fn<decltype(a.MulticastBindMem)>("cuMulticastBindMem");
      a.MulticastGetGranularity = driver_fn<decltype(a.MulticastGetGranularity)>("cuMulticastGetGranularity");
      a.DeviceGetAttribute = driver_fn<decltype(a.DeviceGetAttribute)>("cuDeviceGetAttribute");
      a.DeviceGet = driver_fn<decltype(a.DeviceGet)>("cuDeviceGet");
      return a;
    }();
    return api;
  }
};

std::vector<uint8_t> to_bytes(const void* p, size_t n) {
  const uint8_t* b = reinterpret_cast<const uint8_t*>(p);
  return std::vector<uint8_t>(b, b + n);
}

std::string sanitize(const std::string& s) {
  std::string o;
  for (char c : s) o.push_back((isalnum(static_cast<unsigned char>(c)) || c == '_') ? c : '_');
  return o;
}

size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- fd passing over AF_UNIX datagram sockets (abstract namespace) ----------
struct FdChannel {
  int sock = -1;
  std::string prefix;
  int rank;

  static sockaddr_un addr_of(const std::string& prefix, int r, socklen_t* len) {
    sockaddr_un a;
    std::memset(&a, 0, sizeof(a));
    a.sun_family = AF_UNIX;
    std::string name = prefix + "_" + std::to_string(r);
    DEAR_CHECK(name.size() + 1 < sizeof(a.sun_path), "socket name too long");
    a.sun_path[0] = '\0';  // abstract namespace: no filesystem entry to clean up
    std::memcpy(a.sun_path + 1, name.data(), name.size());
    *len = static_cast<socklen_t>(offsetof(sockaddr_un, sun_path) + 1 + name.size());
    return a;
  }

  FdChannel(const std::string& prefix_, int rank_) : prefix(prefix_), rank(rank_) {
    sock = ::socket(AF_UNIX, SOCK_DGRAM, 0);
    DEAR_CHECK(sock >= 0, "socket() failed: " << strerror(errno));
    socklen_t len;
    sockaddr_un a = addr_of(prefix, rank, &len);
    DEAR_CHECK(::bind(sock, reinterpret_cast<sockaddr*>(&a), len) == 0,
               "bind() failed: " << strerror(errno));
  }
  ~FdChannel() {
    if (sock >= 0) ::close(sock);
  }

  void send_fd(int dst_rank, int fd, int32_t tag) {
    socklen_t len;
    sockaddr_un a = addr_of(prefix, dst_rank, &len);
    int32_t payload[2] = {rank, tag};
    iovec iov{payload, sizeof(payload)};
    char cbuf[CMSG_SPACE(sizeof(int))];
    std::memset(cbuf, 0, sizeof(cbuf));
    msghdr msg;
    std::memset(&msg, 0, sizeof(msg));
    msg.msg_name = &a;
    msg.msg_namelen = len;
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = cbuf;
    msg.msg_controllen = sizeof(cbuf);
    cmsghdr* c = CMSG_FIRSTHDR(&msg);
    c->cmsg_level = SOL_SOCKET;
    c->cmsg_type = SCM_RIGHTS;
    c->cmsg_len = CMSG_LEN(sizeof(int));
    std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
    for (int tries = 0;; ++tries) {
      ssize_t n = ::sendmsg(sock, &msg, 0);
      if (n >= 0) return;
      if ((errno == ECONNREFUSED || errno == ENOENT || errno == EAGAIN || errno == ENOBUFS) && tries < 2000) {
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
        continue;
      }
      DEAR_CHECK(false, "sendmsg(SCM_RIGHTS) failed: " << strerror(errno));
    }
  }

  // Receives one fd; returns (src_rank, tag, fd).
  void recv_fd(int* src_rank, int32_t* tag, int* fd, double timeout_s) {
    pollfd pfd{sock, POLLIN, 0};
    int pr = ::poll(&pfd, 1, static_cast<int>(timeout_s * 1000));
    DEAR_CHECK(pr > 0, "timed out waiting for a peer's memory handle");
    int32_t payload[2] = {-1, -1};
    iovec iov{payload, sizeof(payload)};
    char cbuf[CMSG_SPACE(sizeof(int))];
    msghdr msg;
    std::memset(&msg, 0, sizeof(msg));
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = cbuf;
    msg.msg_controllen = sizeof(cbuf);
    ssize_t n = ::recvmsg(sock, &msg, 0);
    DEAR_CHECK(n == sizeof(payload), "recvmsg failed: " << strerror(errno));
    cmsghdr* c = CMSG_FIRSTHDR(&msg);
    DEAR_CHECK(c != nullptr && c->cmsg_type == SCM_RIGHTS, "no fd in message");
    std::memcpy(fd, CMSG_DATA(c), sizeof(int));
    *src_rank = payload[0];
    *tag = payload[1];
  }
};

}  // namespace

// ---------------------------------------------------------------------------
// status word
// ---------------------------------------------------------------------------
namespace {
uint32_t g_status_fallback = 0;
uint32_t* g_status_host = nullptr;
uint32_t* g_status_dev = nullptr;
bool g_status_init = false;

void init_status() {
  if (g_status_init) return;
  g_status_init = true;
  if (cuda_runtime_usable()) {
    void* h = nullptr;
    if (cudaHostAlloc(&h, sizeof(uint32_t) * 16, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
      std::memset(h, 0, sizeof(uint32_t) * 16);
      void* d = nullptr;
      if (cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
        g_status_host = reinterpret_cast<uint32_t*>(h);
        g_status_dev = reinterpret_cast<uint32_t*>(d);
        return;
      }
    }
    cudaGetLastError();
  }
  g_status_host = &g_status_fallback;
  g_status_dev = nullptr;
}
}  // namespace

bool cuda_runtime_usable() {
  static int usable = -1;
  if (usable < 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) cudaGetLastError();
    usable = (e == cudaSuccess && n > 0) ? 1 : 0;
  }
  return usable == 1;
}

uint32_t* status_word_host() {
  init_status();
  return g_status_host;
}
uint32_t* status_word_device() {
  init_status();
  return g_status_dev;
}

// ---------------------------------------------------------------------------
// SymmArena
// ---------------------------------------------------------------------------
void SymmArena::store_barrier(const std::string& tag) {
  if (world_ == 1) return;
  const std::string base = key_ + "/bar/" + tag + "/" + std::to_string(barrier_seq_++) + "/";
  store_->set(base + std::to_string(rank_), std::vector<uint8_t>{1});
  std::vector<std::string> keys;
  for (int r = 0; r < world_; ++r) keys.push_back(base + std::to_string(r));
  store_->wait(keys, std::chrono::milliseconds(static_cast<int64_t>(timeout_s_ * 1000)));
}

PeerTable SymmArena::data_table(size_t byte_off) const {
  PeerTable t;
  for (int r = 0; r < kMaxRanks; ++r) t.ptr[r] = r < world_ ? data(r) + byte_off : nullptr;
  return t;
}

PeerTable SymmArena::sig_table() const {
  PeerTable t;
  for (int r = 0; r < kMaxRanks; ++r) t.ptr[r] = r < world_ ? sig(r) : nullptr;
  return t;
}

void SymmArena::alloc_ctrl() {
  const size_t n = sizeof(uint32_t) * 3 * kNumChannels;   // epochs | arrival counters | per-bucket update counts
  if (is_cuda()) {
    void* p = nullptr;
    DEAR_CUDA(cudaMalloc(&p, n));
    DEAR_CUDA(cudaMemset(p, 0, n));
    ctrl_ = reinterpret_cast<uint32_t*>(p);
  } else {
    ctrl_ = reinterpret_cast<uint32_t*>(std::calloc(1, n));
  }
}

std::shared_ptr<SymmArena> SymmArena::create(size_t data_bytes, int rank, int world,
                                             const c10::intrusive_ptr<c10d::Store>& store,
                                             const std::string& key, const ArenaOptions& opt) {
  DEAR_CHECK(world >= 1 && world <= kMaxRanks, "world size " << world << " unsupported (max " << kMaxRanks << ")");
  DEAR_CHECK(rank >= 0 && rank < world, "bad rank");
  DEAR_CHECK(world == 1 || store, "a c10d store is required for world > 1");
  std::shared_ptr<SymmArena> a(new SymmArena());
  a->provider_ = opt.provider;
  a->rank_ = rank;
  a->world_ = world;
  a->device_ = opt.provider == Provider::HOST_SHM ? -1 : opt.device;
  a->data_bytes_ = round_up(data_bytes, 256);
  a->total_bytes_ = a->data_bytes_ + kSignalPadBytes;
  a->store_ = store;
  a->key_ = key;
  a->timeout_s_ = opt.timeout_s;
  a->bases_.assign(world, nullptr);
  switch (opt.provider) {
    case Provider::HOST_SHM: a->init_host_shm(key); break;
    case Provider::CUDA_IPC:
      DEAR_CUDA(cudaSetDevice(opt.device));
      a->init_cuda_ipc(key);
      break;
    case Provider::CUDA_VMM:
      DEAR_CUDA(cudaSetDevice(opt.device));
      a->init_cuda_vmm(key, opt.want_multicast);
      break;
    default: DEAR_CHECK(false, "use from_external() for Provider::EXTERNAL");
  }
  a->alloc_ctrl();
  a->store_barrier("created");
  return a;
}

std::shared_ptr<SymmArena> SymmArena::from_external(const std::vector<uint64_t>& bases, uint64_t mc_base,
                                                    size_t data_bytes, int rank, int world, int device) {
  DEAR_CHECK(static_cast<int>(bases.size()) == world, "need one base pointer per rank");
  std::shared_ptr<SymmArena> a(new SymmArena());
  a->provider_ = Provider::EXTERNAL;
  a->rank_ = rank;
  a->world_ = world;
  a->device_ = device;
  a->data_bytes_ = data_bytes;
  a->total_bytes_ = data_bytes + kSignalPadBytes;
  a->owns_ = false;
  for (uint64_t b : bases) a->bases_.push_back(reinterpret_cast<char*>(b));
  a->mc_base_ = reinterpret_cast<char*>(mc_base);
  a->alloc_ctrl();
  return a;
}

// ---- HOST_SHM -------------------------------------------------------------
void SymmArena::init_host_shm(const std::string& key) {
  const std::string base = "/dear_" + sanitize(key) + "_";
  shm_names_.resize(world_);
  for (int r = 0; r < world_; ++r) shm_names_[r] = base + std::to_string(r);
  {
    ::shm_unlink(shm_names_[rank_].c_str());
    int fd = ::shm_open(shm_names_[rank_].c_str(), O_CREAT | O_RDWR | O_EXCL, 0600);
    DEAR_CHECK(fd >= 0, "shm_open(" << shm_names_[rank_] << ") failed: " << strerror(errno));
    DEAR_CHECK(::ftruncate(fd, static_cast<off_t>(total_bytes_)) == 0, "ftruncate failed: " << strerror(errno));
    void* p = ::mmap(nullptr, total_bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    ::close(fd);
    DEAR_CHECK(p != MAP_FAILED, "mmap failed: " << strerror(errno));
    std::memset(p, 0, total_bytes_);
    bases_[rank_] = reinterpret_cast<char*>(p);
  }
  store_barrier("shm_created");
  for (int r = 0; r < world_; ++r) {
    if (r == rank_) continue;
    int fd = ::shm_open(shm_names_[r].c_str(), O_RDWR, 0600);
    DEAR_CHECK(fd >= 0, "shm_open(peer " << r << ") failed: " << strerror(errno));
    void* p = ::mmap(nullptr, total_bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    ::close(fd);
    DEAR_CHECK(p != MAP_FAILED, "mmap(peer) failed: " << strerror(errno));
    bases_[r] = reinterpret_cast<char*>(p);
  }
  store_barrier("shm_mapped");
  ::shm_unlink(shm_names_[rank_].c_str());   // mappings stay valid; name is gone
}

// ---- CUDA_IPC -------------------------------------------------------------
void SymmArena::init_cuda_ipc(const std::string& key) {
  void* p = nullptr;
  DEAR_CUDA(cudaMalloc(&p, total_bytes_));
  DEAR_CUDA(cudaMemset(p, 0, total_bytes_));
  DEAR_CUDA(cudaDeviceSynchronize());
  bases_[rank_] = reinterpret_cast<char*>(p);
  if (world_ == 1) return;
  cudaIpcMemHandle_t h;
  DEAR_CUDA(cudaIpcGetMemHandle(&h, p));
  store_->set(key + "/ipc/" + std::to_string(rank_), to_bytes(&h, sizeof(h)));
  for (int r = 0; r < world_; ++r) {
    if (r == rank_) continue;
    std::vector<uint8_t> b = store_->get(key + "/ipc/" + std::to_string(r));
    DEAR_CHECK(b.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle from rank " << r);
    cudaIpcMemHandle_t ph;
    std::memcpy(&ph, b.data(), sizeof(ph));
    void* q = nullptr;
    DEAR_CUDA(cudaIpcOpenMemHandle(&q, ph, cudaIpcMemLazyEnablePeerAccess));
    bases_[r] = reinterpret_cast<char*>(q);
  }
}

// ---- CUDA_VMM (+ NVLS multicast) ---------------------------------------------
void SymmArena::init_cuda_vmm(const std::string& key, bool want_mc) {
  const DriverApi& cu = DriverApi::get();
  DEAR_CUDA(cudaFree(nullptr));   // make sure the primary context exists
  CUdevice dev;
  DEAR_CU(cu.DeviceGet(&dev, device_));

  CUmemAllocationProp prop;
  std::memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device_;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  DEAR_CU(cu.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));

  // Is multicast possible at all?  (all ranks must agree; decided below.)
  int mc_supported = 0;
  if (want_mc && world_ > 1) {
    if (cu.DeviceGetAttribute(&mc_supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS)
      mc_supported = 0;
  }
  CUmulticastObjectProp mcprop;
  std::memset(&mcprop, 0, sizeof(mcprop));
  size_t mc_gran = 0;
  if (mc_supported) {
    mcprop.numDevices = static_cast<unsigned>(world_);
    mcprop.size = total_bytes_;
    mcprop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    if (cu.MulticastGetGranularity(&mc_gran, &mcprop, CU_MULTICAST_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS)
      mc_supported = 0;
  }
  if (mc_gran > gran) gran = mc_gran;
  mapped_bytes_ = round_up(total_bytes_, gran);

  CUmemGenericAllocationHandle local;
  DEAR_CU(cu.MemCreate(&local, mapped_bytes_, &prop, 0));
  vmm_handles_.assign(world_, 0);
  vmm_handles_[rank_] = local;

  auto map_handle = [&](CUmemGenericAllocationHandle h) -> char* {
    CUdeviceptr va = 0;
    DEAR_CU(cu.MemAddressReserve(&va, mapped_bytes_, gran, 0, 0));
    DEAR_CU(cu.MemMap(va, mapped_bytes_, 0, h, 0));
    CUmemAccessDesc acc;
    std::memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device_;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    DEAR_CU(cu.MemSetAccess(va, mapped_bytes_, &acc, 1));
    return reinterpret_cast<char*>(va);
  };

  bases_[rank_] = map_handle(local);
  DEAR_CUDA(cudaMemset(bases_[rank_], 0, mapped_bytes_));
  DEAR_CUDA(cudaDeviceSynchronize());
  if (world_ == 1) return;

  // exchange POSIX fds over AF_UNIX datagram sockets
  FdChannel chan("dear_" + sanitize(key), rank_);
  store_barrier("vmm_sock");
  int my_fd = -1;
  DEAR_CU(cu.MemExportToShareableHandle(&my_fd, local, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  for (int k = 1; k < world_; ++k) chan.send_fd((rank_ + k) % world_, my_fd, /*tag=*/0);
  for (int k = 1; k < world_; ++k) {
    int src = -1, fd = -1;
    int32_t tag = -1;
    chan.recv_fd(&src, &tag, &fd, timeout_s_);
    DEAR_CHECK(tag == 0 && src >= 0 && src < world_ && src != rank_, "unexpected handle message");
    CUmemGenericAllocationHandle ph;
    DEAR_CU(cu.MemImportFromShareableHandle(&ph, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                            CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    ::close(fd);
    vmm_handles_[src] = ph;
    bases_[src] = map_handle(ph);
  }
  ::close(my_fd);
  store_barrier("vmm_mapped");

  // NVLS multicast: all-or-nothing across ranks.
  std::string votes_key = key + "/mcvote/";
  store_->set(votes_key + std::to_string(rank_), std::vector<uint8_t>{static_cast<uint8_t>(mc_supported ? 1 : 0)});
  bool all_mc = mc_supported != 0;
  for (int r = 0; r < world_; ++r) {
    std::vector<uint8_t> v = store_->get(votes_key + std::to_string(r));
    if (v.empty() || v[0] == 0) all_mc = false;
  }
  if (!all_mc) return;

  bool ok = true;
  CUmemGenericAllocationHandle mch = 0;
  try {
    mcprop.size = mapped_bytes_;
    if (rank_ == 0) {
      DEAR_CU(cu.MulticastCreate(&mch, &mcprop));
      int mfd = -1;
      DEAR_CU(cu.MemExportToShareableHandle(&mfd, mch, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
      for (int r = 1; r < world_; ++r) chan.send_fd(r, mfd, /*tag=*/1);
      ::close(mfd);
    } else {
      int src = -1, fd = -1;
      int32_t tag = -1;
      chan.recv_fd(&src, &tag, &fd, timeout_s_);
      DEAR_CHECK(tag == 1 && src == 0, "unexpected multicast handle message");
      DEAR_CU(cu.MemImportFromShareableHandle(&mch, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                              CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
      ::close(fd);
    }
    DEAR_CU(cu.MulticastAddDevice(mch, dev));
  } catch (const std::exception&) {
    ok = false;
  }
  // every device must be added before anyone binds
  store_->set(key + "/mcadd/" + std::to_string(rank_), std::vector<uint8_t>{static_cast<uint8_t>(ok ? 1 : 0)});
  for (int r = 0; r < world_; ++r) {
    std::vector<uint8_t> v = store_->get(key + "/mcadd/" + std::to_string(r));
    if (v.empty() || v[0] == 0) ok = false;
  }
  if (ok) {
    try {
      DEAR_CU(cu.MulticastBindMem(mch, 0, local, 0, mapped_bytes_, 0));
      CUdeviceptr va = 0;
      DEAR_CU(cu.MemAddressReserve(&va, mapped_bytes_, gran, 0, 0));
      DEAR_CU(cu.MemMap(va, mapped_bytes_, 0, mch, 0));
      CUmemAccessDesc acc;
      std::memset(&acc, 0, sizeof(acc));
      acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      acc.location.id = device_;
      acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      DEAR_CU(cu.MemSetAccess(va, mapped_bytes_, &acc, 1));
      mc_base_ = reinterpret_cast<char*>(va);
      mc_handle_ = mch;
    } catch (const std::exception&) {
      ok = false;
    }
  }
  store_->set(key + "/mcbind/" + std::to_string(rank_), std::vector<uint8_t>{static_cast<uint8_t>(ok ? 1 : 0)});
  for (int r = 0; r < world_; ++r) {
    std::vector<uint8_t> v = store_->get(key + "/mcbind/" + std::to_string(r));
    if (v.empty() || v[0] == 0) ok = false;
  }
  if (!ok) mc_base_ = nullptr;   // P2P path stays fully functional
}

SymmArena::~SymmArena() {
  // Best effort; never throw from a destructor.
  try {
    if (ctrl_) {
      if (is_cuda()) cudaFree(ctrl_); else std::free(ctrl_);
    }
    if (!owns_) return;
    if (provider_ == Provider::HOST_SHM) {
      for (char* b : bases_) if (b) ::munmap(b, total_bytes_);
    } else if (provider_ == Provider::CUDA_IPC) {
      for (int r = 0; r < world_; ++r) {
        if (!bases_[r]) continue;
        if (r == rank_) cudaFree(bases_[r]); else cudaIpcCloseMemHandle(bases_[r]);
      }
    } else if (provider_ == Provider::CUDA_VMM) {
      const DriverApi& cu = DriverApi::get();
      if (mc_base_) {
        cu.MemUnmap(reinterpret_cast<CUdeviceptr>(mc_base_), mapped_bytes_);
        cu.MemAddressFree(reinterpret_cast<CUdeviceptr>(mc_base_), mapped_bytes_);
      }
      if (mc_handle_) cu.MemRelease(mc_handle_);
      for (int r = 0; r < world_; ++r) {
        if (!bases_[r]) continue;
        cu.MemUnmap(reinterpret_cast<CUdeviceptr>(bases_[r]), mapped_bytes_);
        cu.MemAddressFree(reinterpret_cast<CUdeviceptr>(bases_[r]), mapped_bytes_);
        if (vmm_handles_[r]) cu.MemRelease(vmm_handles_[r]);
      }
    }
  } catch (...) {
  }
}

}  // namespace dear
