// dear_common.h — shared declarations for the B200-native DeAR runtime.
//
// Everything in this header is usable from host C++ and from sm_100a device
// code: the host-emulation backend (emu.cpp, used for CPU/gloo plumbing
// tests) and the CUDA kernels (kernels.cu) share the SAME argument structs
// and the SAME per-element math, so a CPU test of the emulation backend
// validates the indexing/SGD logic that the kernels execute.
//
// Parity notes (reference = lzhangbv/dear_pytorch):
//   * reduce-scatter call site  : dear/tensorfusion.py:475-476 -> comm_core reduceScatter
//   * all-gather call site      : dear/tensorfusion.py:477-478 -> comm_core allGather
//   * per-parameter SGD update  : dear/dear_dopt.py:310-336
// Here both collectives are fused with their adjacent elementwise work and
// run over peer-mapped (NVLink / NVSwitch) memory instead of NCCL.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cmath>

#if defined(__CUDACC__)
#define DEAR_HD __host__ __device__ __forceinline__
#else
#define DEAR_HD inline
#endif

namespace dear {

constexpr int kMaxRanks = 16;            // one NVSwitch domain (8 on HGX B200)
constexpr int kNumChannels = 4096;       // signal-pad channels per arena
constexpr int kChannelsPerBucket = 4;    // RS_READY, RS_DONE, AG_ARRIVE, AG_PUSHED
constexpr int kGeneralChannels = 64;     // channels [0,64) are for general ops
constexpr uint32_t kPackTileBytes = 65536;
constexpr size_t kSignalPadBytes = size_t(kNumChannels) * kMaxRanks * sizeof(uint32_t);

// Channel ids for the general-purpose ops (all-reduce, broadcast, ...).
enum GeneralChannel : uint32_t {
  CH_BARRIER = 0,
  CH_AR_READY = 2,
  CH_AR_DONE = 3,
  CH_BCAST_READY = 4,
  CH_BCAST_DONE = 5,
  CH_REDUCE_READY = 6,
  CH_REDUCE_DONE = 7,
  CH_SENDRECV_READY = 8,
  CH_SENDRECV_DONE = 9,
  CH_AG_READY = 10,
  CH_AG_DONE = 11,
};

enum BucketChannel : uint32_t { RS_READY = 0, RS_DONE = 1, AG_ARRIVE = 2, AG_PUSHED = 3 };

DEAR_HD uint32_t bucket_channel(uint32_t bucket, uint32_t which) {
  return kGeneralChannels + bucket * kChannelsPerBucket + which;
}

// Error codes written to the (host-mapped) status word when a spin-wait
// times out.  A kernel never hangs the GPU: it gives up, flags, and exits.
enum Status : uint32_t {
  ST_OK = 0,
  ST_TIMEOUT_RS_READY = 1,
  ST_TIMEOUT_RS_DONE = 2,
  ST_TIMEOUT_AG_ARRIVE = 3,
  ST_TIMEOUT_AG_PUSHED = 4,
  ST_TIMEOUT_GENERAL = 5,
};

enum DType : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };

DEAR_HD size_t dtype_size(int dt) { return dt == DT_F32 ? 4 : 2; }

struct PeerTable {
  void* ptr[kMaxRanks];
};

// One gradient segment of a bucket (== one parameter's gradient).
struct PackSeg {
  const void* src;        // local gradient storage; nullptr => nothing to copy
  uint64_t dst_off;       // byte offset inside the bucket
  uint64_t nbytes;        // bytes to copy
  uint32_t tile_begin;    // exclusive prefix sum of 64 KiB tiles
  uint32_t flags;         // bit0: zero-fill the destination (gradient absent)
};
constexpr uint32_t SEG_ZERO_FILL = 1u;

// Contiguous element range of a bucket sharing one set of optimizer hyper-parameters.
enum OptKind : uint32_t { OPT_SGD = 0, OPT_ADAM = 1 /* L2 in the gradient */, OPT_ADAMW = 2 /* decoupled decay */ };
struct HyperSeg {
  uint64_t end;           // exclusive end (element offset within the bucket)
  float lr;
  float weight_decay;
  float momentum;         // SGD momentum, or Adam beta1
  float dampening;
  uint32_t nesterov;      // bit0: Nesterov momentum; bit1 (HYPER_SKIP): this rank saw no gradient for the range in
                          // this step — where the REDUCED gradient is exactly zero too (absent on every rank), leave
                          // parameter and state untouched, like torch.optim skips ``p.grad is None``
  uint32_t opt;           // OptKind
  float beta2;            // Adam only
  float eps;              // Adam only
};

constexpr uint32_t HYPER_NESTEROV = 1u;
constexpr uint32_t HYPER_SKIP = 2u;

// ---- per-element math shared by the CUDA kernels and the host emulation ----

// torch-1.8 SGD semantics (reference dear/dear_dopt.py:310-336):
//   g <- g + wd*p ; buf <- (first ? g : m*buf + (1-damp)*g) ; g <- nesterov ? g + m*buf : buf
//   p <- p - lr*g
// `g` is already averaged (the 1/P scale is fused into the reduce-scatter).
DEAR_HD float sgd_update(float p, float g, float& mom, const HyperSeg& h, bool first_step,
                         bool has_mom_buf) {
  if ((h.nesterov & HYPER_SKIP) && g == 0.f) return p;
  if (h.weight_decay != 0.f) g = g + h.weight_decay * p;
  if (h.momentum > 0.f && has_mom_buf) {
    float buf = first_step ? g : (h.momentum * mom + (1.f - h.dampening) * g);
    mom = buf;
    g = (h.nesterov & HYPER_NESTEROV) ? (g + h.momentum * buf) : buf;
  }
  return p - h.lr * g;
}

// torch.optim.Adam / AdamW semantics (non-amsgrad): exp_avg `m`, exp_avg_sq `v`, bias corrections
// bc1 = 1 - beta1^t, bc2 = 1 - beta2^t.  This extends the reference, whose DeAR path is SGD-only
// (dear/dear_dopt.py:310-336; its BERT driver had to drop AdamW, dear/bert_benchmark.py:118-122).
DEAR_HD float adam_update(float p, float g, float& m, float& v, const HyperSeg& h, float bc1, float sqrt_bc2) {
  if ((h.nesterov & HYPER_SKIP) && g == 0.f) return p;
  if (h.opt == OPT_ADAM && h.weight_decay != 0.f) g = g + h.weight_decay * p;
  m = h.momentum * m + (1.f - h.momentum) * g;
  v = h.beta2 * v + (1.f - h.beta2) * g * g;
#if defined(__CUDA_ARCH__)
  const float denom = sqrtf(v) / sqrt_bc2 + h.eps;
#else
  const float denom = std::sqrt(v) / sqrt_bc2 + h.eps;
#endif
  if (h.opt == OPT_ADAMW) p = p * (1.f - h.lr * h.weight_decay);
  return p - (h.lr / bc1) * (m / denom);
}

// Index of the hyper segment containing element `e` (segments sorted by end).
DEAR_HD uint32_t find_hyper(const HyperSeg* segs, uint32_t n, uint64_t e) {
  uint32_t lo = 0, hi = n - 1;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (e < segs[mid].end) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// Index of the pack segment owning tile `t` (tile_begin is a prefix sum).
DEAR_HD uint32_t find_pack_seg(const PackSeg* segs, uint32_t n, uint32_t t) {
  uint32_t lo = 0, hi = n;          // invariant: segs[lo].tile_begin <= t
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (segs[mid].tile_begin <= t) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- kernel parameter blocks ------------------------------------------------

// Kernel A: fused [pack local grads -> symmetric bucket] + cross-GPU ready
// barrier + pull-reduce of this rank's shard from every peer + 1/P scale.
struct RSParams {
  PeerTable grad;          // every rank's bucket base (grad[rank] is local)
  void* mc_grad;           // NVLS multicast alias of the bucket (or nullptr)
  float* out;              // local fp32 reduced shard [shard_elems]
  uint64_t shard_elems;    // elements per shard (bucket = world * shard_elems)
  float scale;             // 1/world
  const PackSeg* segs;     // device table (nullptr => nothing to pack)
  uint32_t nseg;
  uint32_t ntiles;
  PeerTable sig;           // every rank's signal pad (sig[rank] is local)
  uint32_t* ctrl;          // local control block: epoch[kNumChannels], counter[kNumChannels]
  uint32_t bucket;         // bucket id (selects channels)
  uint32_t direct_out;     // world==1 && fp32: pack straight into `out`, skip the pull phase
  int rank;
  int world;
  int dtype;               // DType of the gradient bucket
  uint32_t* status;        // host-mapped status word
  uint64_t timeout_ns;
  // stripe-pipelined variant (rs_pipe.cu): the shard is cut into `nstripes` stripes of `stripe_bytes`
  // (a multiple of kPipePackPiece); stripe k of EVERY shard is packed and published before stripe k+1
  uint32_t nstripes;
  uint32_t reserved;
  uint64_t stripe_bytes;
  // pack work list of the pipelined variant: `pieces` (device) holds <= kPipePackPiece-byte copies in stripe-major
  // order, stripe k owning entries [piece_first[k], piece_first[k+1])
  const PackSeg* pieces;
  uint32_t piece_first[17];
};

// RS_READY flag encoding shared by both reduce-scatter kernels: (epoch << 8) | stripes_published; the one-shot
// kernel publishes kAllStripes.  Monotonic, so the wrap-safe ">=" wait works for either producer.
constexpr uint32_t kAllStripes = 255u;
constexpr uint32_t kMaxStripes = 16u;
constexpr uint32_t kPipeChunk = 16384;        // bytes per TMA bulk copy of the pull ring
constexpr uint32_t kPipePackPiece = 32768;    // bytes per pack work item

enum RsAlgo : int {
  RS_ALGO_ONESHOT = 0,    // pack, one cross-GPU flag round, pull with 128-bit loads (small buckets: fewest sync rounds)
  RS_ALGO_PIPE = 1,       // stripe-pipelined: pack warps + TMA (cp.async.bulk) pull ring + shared-memory reduce
  RS_ALGO_NVLS = 2,       // pack, then multimem.ld_reduce (the NVSwitch reduces); needs a multicast-bound arena
};

// Kernel B: fused [sharded SGD/momentum update] + push of the updated
// parameter shard into every peer's parameter bucket + completion barrier.
struct AGParams {
  PeerTable param;         // every rank's parameter bucket base
  void* mc_param;          // NVLS multicast alias (or nullptr)
  const float* grad_shard; // local fp32 averaged gradient shard
  float* mom_shard;        // local fp32 momentum / Adam exp_avg shard (nullptr if unused)
  float* var_shard;        // local fp32 Adam exp_avg_sq shard (nullptr for SGD)
  uint32_t* step_ctr;      // device-resident count of applied updates (Adam bias correction; graph-safe)
  uint32_t adam;           // 1 => every hyper segment is Adam/AdamW
  float* master_shard;     // local fp32 master shard (nullptr => param bucket is fp32 master)
  void* zero_grad;         // local gradient bucket to zero after use (nullptr => skip)
  uint64_t zero_bytes;
  uint64_t shard_elems;
  const HyperSeg* hyper;   // device table
  uint32_t nhyper;
  uint32_t first_step;     // momentum buffers are uninitialised
  uint32_t entry_barrier;  // wait for every peer to reach this kernel before pushing
  uint32_t do_update;      // 0 => pure all-gather of the shard (no SGD)
  PeerTable sig;
  uint32_t* ctrl;
  uint32_t bucket;
  int rank;
  int world;
  int dtype;               // DType of the parameter bucket
  uint32_t* status;
  uint64_t timeout_ns;
};

// General ops on a symmetric staging buffer.
enum GenOp : int {
  GEN_ALLREDUCE = 0,   // sum over ranks of staging[0:n] -> dst (every rank)
  GEN_BCAST = 1,       // root's staging -> dst on every rank
  GEN_REDUCE = 2,      // sum over ranks -> dst on root only
  GEN_SENDRECV = 3,    // dst <- peer's staging
  GEN_ALLGATHER = 4,   // dst[r*n:(r+1)*n] <- rank r's staging
  GEN_BARRIER = 5,
  GEN_REDUCE_SCATTER = 6,  // dst[0:n/P] <- sum over ranks of staging[rank*n/P : ...]
};

struct GenParams {
  PeerTable stage;       // every rank's staging buffer
  void* mc_stage;
  const void* src;       // local input (copied into local staging first); may be nullptr
  void* dst;             // local output
  uint64_t nelems;       // elements of the op (per-rank input size)
  int op;
  int root_or_peer;
  float scale;           // applied to reductions
  uint32_t ready_chan;
  uint32_t done_chan;
  PeerTable sig;
  uint32_t* ctrl;
  int rank;
  int world;
  int dtype;             // DT_F32 / DT_BF16 / DT_F16 ; int64 is moved as raw bytes (no reduce)
  uint32_t elem_bytes;   // for raw moves
  uint64_t dst_stride_bytes;  // GEN_ALLGATHER: byte distance between ranks' slots in dst (0 => contiguous)
  uint32_t* status;
  uint64_t timeout_ns;
};

}  // namespace dear
