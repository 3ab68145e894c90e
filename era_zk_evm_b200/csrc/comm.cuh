// Multi-GPU stream exchange behind the C ABI (SURVEY §8e, §8b `zkb_gather`): NCCL driven from C++ so that a host in any
// language reaches it (round 1 had it only in Python over torch.distributed).  One process per GPU; the host hands a
// 128-byte NCCL unique id from rank 0 to every rank through its own channel and calls zkb_comm_create.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2"): libzkb.so has no link-time dependency on it, a single-GPU user
// never loads it, and inside a process that already carries a libnccl (PyTorch's) the same copy is picked up.
//
//   zkb_gather_streams   concatenation of whole streams on ONE rank, in rank order (variable length): the sink can rotate
//                        from step to step (dst_rank = step % world), so no GPU takes every step's ingress
//   zkb_exchange_logs    the balanced form for the query log: every LOG record goes to rank  slot_hash(shard, address,
//                        key) % world  -- the partition a global per-slot sort / dedup wants anyway (zkb_sort_log_queries
//                        runs on each rank's share afterwards).  Per-GPU ingress is then ~1/world of every rank's log =
//                        its own log's size, whatever world is.  The pack is fused into the partition: one kernel reads
//                        the records straight from the per-VM slabs and drops them into the per-destination send regions
//                        (deterministic positions from a count + scan pass; no atomics), NCCL moves the regions all-to-all.
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <time.h>

namespace zkb {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};

static NcclApi* nccl_api(std::string* why) {
  static NcclApi api;
  static bool tried = false, ok = false;
  static std::string err;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      err = std::string("dlopen libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
    } else {
      bool all = true;
      auto sym = [&](const char* name) {
        void* p = dlsym(api.handle, name);
        if (!p) {
          all = false;
          err = std::string("libnccl: missing symbol ") + name;
        }
        return p;
      };
      api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
      api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
      api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
      api.Send = (decltype(api.Send))sym("ncclSend");
      api.Recv = (decltype(api.Recv))sym("ncclRecv");
      api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
      ok = all;
    }
  }
  if (!ok && why) *why = err;
  return ok ? &api : nullptr;
}

// ---- kernels of the hash-partitioned exchange ----------------------------------------------------------------------
// pass 1: per VM and destination, how many of the VM's LOG records go there.  One warp per VM; lane l takes records
// l, l + 32, ...
// (128 threads x <= 32 registers: fits next to the persistent interpreter CTA on an SM, which leaves 4 K registers free)
__global__ void __launch_bounds__(128, 16) zkb_bucket_count_kernel(const DevBatch B, uint32_t world, uint32_t* __restrict__ counts /* [world][n_vms] */) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t vm = blockIdx.x * 4 + warp; vm < B.n_vms; vm += gridDim.x * 4) {
    const uint32_t n = B.hot[vm].x[X_COUNT0 + ZKB_STREAM_LOG];
    const uint32_t* recs = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_LOG] + (size_t)vm * B.cap[ZKB_STREAM_LOG] * ZKB_LOG_BYTES);
    uint32_t mine[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t r = lane; r < n; r += 32) {
      const uint32_t* p = recs + (size_t)r * 32;
      const uint32_t dst = (uint32_t)((slot_hash64(p[1] >> 24, p + 2, p + 8) >> 20) % world);
#pragma unroll
      for (uint32_t d = 0; d < 8; d++) mine[d] += dst == d ? 1u : 0u;
    }
#pragma unroll
    for (uint32_t d = 0; d < 8; d++) {
      uint32_t v = mine[d];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && d < world) counts[(size_t)d * B.n_vms + vm] = v;
    }
  }
}

// exclusive scan of counts[d][0..n) per destination d (one 128-thread block each -- small enough to sit next to the
// persistent interpreter CTA); totals[d] = records this rank sends to d
__global__ void __launch_bounds__(128, 16) zkb_bucket_scan_kernel(uint32_t* __restrict__ counts, uint32_t n, uint64_t* __restrict__ totals) {
  __shared__ uint32_t s_warp[4];
  uint32_t* c = counts + (size_t)blockIdx.x * n;
  const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
  uint64_t carry = 0;
  for (uint32_t base = 0; base < n; base += 128) {
    const uint32_t i = base + t;
    const uint32_t v = i < n ? c[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < warp; w++) before += s_warp[w];
    const uint32_t tile_total = s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
    if (i < n) c[i] = (uint32_t)carry + before + incl - v;
    carry += tile_total;
    __syncthreads();
  }
  if (t == 0) totals[blockIdx.x] = carry;
}

// pass 2 (the fused pack): records straight from the per-VM slabs into the per-destination regions of the send buffer,
// at  region_base[d] + offs[d][vm] + (rank of the record among the VM's records for d).  One warp per VM; a record is
// moved by the whole warp (32 lanes x 4 bytes), in record order, so positions are deterministic.
__global__ void __launch_bounds__(128, 16) zkb_bucket_pack_kernel(const DevBatch B, uint32_t world, const uint32_t* __restrict__ offs /* [world][n_vms] */,
                                                              const uint64_t* __restrict__ totals, uint32_t* __restrict__ send) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint64_t base[8];
  uint64_t run = 0;
#pragma unroll
  for (uint32_t d = 0; d < 8; d++) {
    base[d] = run;
    run += d < world ? totals[d] : 0ull;
  }
  for (uint32_t vm = blockIdx.x * 4 + warp; vm < B.n_vms; vm += gridDim.x * 4) {
    const uint32_t n = B.hot[vm].x[X_COUNT0 + ZKB_STREAM_LOG];
    const uint32_t* recs = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_LOG] + (size_t)vm * B.cap[ZKB_STREAM_LOG] * ZKB_LOG_BYTES);
    uint32_t next[8];
#pragma unroll
    for (uint32_t d = 0; d < 8; d++) next[d] = d < world ? offs[(size_t)d * B.n_vms + vm] : 0u;
    for (uint32_t r0 = 0; r0 < n; r0 += 32) {
      // lane l classifies record r0 + l; then the warp moves the 32 records one after the other
      uint32_t my_dst = 0xFFu;
      if (r0 + lane < n) {
        const uint32_t* p = recs + (size_t)(r0 + lane) * 32;
        my_dst = (uint32_t)((slot_hash64(p[1] >> 24, p + 2, p + 8) >> 20) % world);
      }
      const uint32_t m = min(32u, n - r0);
      for (uint32_t k = 0; k < m; k++) {
        const uint32_t d = __shfl_sync(0xffffffffu, my_dst, k);
        uint32_t pos = 0;
#pragma unroll
        for (uint32_t q = 0; q < 8; q++)
          if (q == d) pos = next[q]++;
        send[(base[d] + pos) * 32 + lane] = recs[(size_t)(r0 + k) * 32 + lane];
      }
    }
  }
}


// ---- one-sided variants: the same partition, written STRAIGHT into the destination GPU's memory over NVLink ----------
// (peer memory mapped through CUDA IPC).  No NCCL on the data path, so no SM has to be kept free for it: every kernel
// here is 128 threads x <= 32 registers and co-resides with the persistent interpreter CTA, i.e. the exchange of pass k
// runs underneath the interpreter launch of pass k + 1 without any host synchronisation (counts are read on the device).
// Receive layout on every rank: `world` fixed-capacity regions, region s = what source rank s sent; a mailbox row per
// source carries its record / byte counts and the step tag that marks the row complete.
struct PushPeers {
  uint8_t* rbuf[8];      // receive buffers of all ranks (own = local pointer)
  uint64_t* mbox[8];     // mailboxes of all ranks: [world][16] u64
};

__global__ void __launch_bounds__(128, 16) zkb_bucket_push_kernel(const DevBatch B, uint32_t world, uint32_t my_rank, const uint32_t* __restrict__ offs,
                                                                 PushPeers P, uint64_t region_bytes) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t cap_records = (uint32_t)(region_bytes / ZKB_LOG_BYTES);   // beyond it nothing is written; the mailbox row reports the overflow
  const uint32_t sub = lane >> 3, l8 = lane & 7u;   // a record is moved by 8 lanes x 16 bytes: four records per warp step
  for (uint32_t vm = blockIdx.x * 4 + warp; vm < B.n_vms; vm += gridDim.x * 4) {
    const uint32_t n = B.hot[vm].x[X_COUNT0 + ZKB_STREAM_LOG];
    const uint32_t* recs = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_LOG] + (size_t)vm * B.cap[ZKB_STREAM_LOG] * ZKB_LOG_BYTES);
    uint32_t next[8];
#pragma unroll
    for (uint32_t d = 0; d < 8; d++) next[d] = d < world ? offs[(size_t)d * B.n_vms + vm] : 0u;
    for (uint32_t r0 = 0; r0 < n; r0 += 32) {
      // lane l classifies record r0 + l ...
      uint32_t my_dst = 0u;
      if (r0 + lane < n) {
        const uint32_t* p = recs + (size_t)(r0 + lane) * 32;
        my_dst = (uint32_t)((slot_hash64(p[1] >> 24, p + 2, p + 8) >> 20) % world);
      }
      const uint32_t m = min(32u, n - r0);
      // ... then the warp moves them four at a time (positions are handed out in record order: deterministic layout)
      for (uint32_t k = 0; k < m; k += 4) {
        uint32_t d_mine = 0, pos_mine = 0;
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) {
          const uint32_t d = __shfl_sync(0xffffffffu, my_dst, (k + j) & 31u);
          uint32_t pos = 0;
#pragma unroll
          for (uint32_t q = 0; q < 8; q++)
            if (q == d && k + j < m) pos = next[q]++;
          if (sub == j) {
            d_mine = d;
            pos_mine = pos;
          }
        }
        if (k + sub < m && pos_mine < cap_records) {
          const uint4 v = __ldcs(reinterpret_cast<const uint4*>(recs + (size_t)(r0 + k + sub) * 32) + l8);
          uint8_t* base = P.rbuf[0];
#pragma unroll
          for (uint32_t q = 1; q < 8; q++)
            if (q == d_mine) base = P.rbuf[q];
          reinterpret_cast<uint4*>(base + (size_t)my_rank * region_bytes + (size_t)pos_mine * ZKB_LOG_BYTES)[l8] = v;
        }
      }
    }
  }
}

// per-VM byte counts of one stream kind (input of the scan that places a VM's records in the sink's region)
__global__ void __launch_bounds__(128, 16) zkb_kind_count_kernel(const DevBatch B, uint32_t kind, uint32_t rec_bytes_k, uint32_t* __restrict__ counts) {
  for (uint32_t vm = blockIdx.x * blockDim.x + threadIdx.x; vm < B.n_vms; vm += gridDim.x * blockDim.x) counts[vm] = B.hot[vm].x[X_COUNT0 + kind] * rec_bytes_k;
}

// concat of one stream kind into the sink's region of this rank: one warp per VM, 8-byte copies (RefundRec is 8 bytes)
__global__ void __launch_bounds__(128, 16) zkb_kind_push_kernel(const DevBatch B, uint32_t kind, uint32_t rec_bytes_k, const uint32_t* __restrict__ offs,
                                                               uint8_t* __restrict__ dst_region, uint64_t region_bytes) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t vm = blockIdx.x * 4 + warp; vm < B.n_vms; vm += gridDim.x * 4) {
    const uint32_t bytes = B.hot[vm].x[X_COUNT0 + kind] * rec_bytes_k;
    const uint8_t* src = B.streams[kind] + (size_t)vm * B.cap[kind] * rec_bytes_k;
    if ((uint64_t)offs[vm] + bytes > region_bytes) continue;   // overflow: reported through the mailbox row
    uint8_t* dst = dst_region + offs[vm];
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | bytes) & 15u) == 0) {
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      uint4* d4 = reinterpret_cast<uint4*>(dst);
      const uint32_t n16 = bytes / 16;
      uint32_t i = lane;
      for (; i + 96 < n16; i += 128) {   // four independent transfers in flight per lane
        const uint4 a = __ldcs(s4 + i), b = __ldcs(s4 + i + 32), c = __ldcs(s4 + i + 64), d = __ldcs(s4 + i + 96);
        d4[i] = a;
        d4[i + 32] = b;
        d4[i + 64] = c;
        d4[i + 96] = d;
      }
      for (; i < n16; i += 32) d4[i] = __ldcs(s4 + i);
    } else {
      const uint2* s2 = reinterpret_cast<const uint2*>(src);
      uint2* d2 = reinterpret_cast<uint2*>(dst);
      for (uint32_t i = lane; i < bytes / 8; i += 32) d2[i] = s2[i];
    }
  }
}

// the mailbox rows: what this rank sent to every destination (records of the exchange), what it pushed to the sink
// (bytes per gathered kind), and the step tag LAST (fenced) -- a row whose tag equals the step is complete
__global__ void zkb_mailbox_kernel(PushPeers P, uint32_t world, uint32_t my_rank, const uint64_t* __restrict__ sent /* [8] records per destination */,
                                   uint32_t sink, uint32_t gather_mask, const uint64_t* __restrict__ kind_bytes /* [6] */, uint64_t step,
                                   uint64_t x_region, const uint64_t* __restrict__ g_kind) {
  const uint32_t d = threadIdx.x;
  if (d >= world) return;
  volatile uint64_t* row = P.mbox[d] + (size_t)my_rank * 16;
  row[0] = sent[d];
  uint64_t overflow = sent[d] * ZKB_LOG_BYTES > x_region ? 1ull : 0ull;
  for (uint32_t k = 0; k < ZKB_N_STREAMS; k++)
    if (d == sink && ((gather_mask >> k) & 1u) && kind_bytes[k] > g_kind[k]) overflow = 1ull;
  row[2] = overflow;
  for (uint32_t k = 0; k < ZKB_N_STREAMS; k++) row[8 + k] = d == sink ? kind_bytes[k] : 0ull;
  row[1] = gather_mask;
  row[14] = sink;
  __threadfence_system();
  row[15] = step;
}

}  // namespace zkb

struct ZkbComm {
  int device = 0, rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  zkb::NcclApi* api = nullptr;
  uint8_t* recv = nullptr;      // grow-only receive buffer of the exchange
  uint64_t recv_capacity = 0;
  uint8_t* concat = nullptr;    // grow-only concat buffer of the gather (its own: a gather must not clobber the last exchange's share)
  uint64_t concat_capacity = 0;
  uint8_t* send = nullptr;      // grow-only send buffer of the exchange
  uint64_t send_capacity = 0;
  uint64_t* d_sizes = nullptr;  // [world][16] sizes matrix on the device: [0..8) records per exchange destination, [8..14) packed stream bytes
  uint64_t* h_sizes = nullptr;  // pinned mirror
  uint32_t* d_counts = nullptr; // [world][n_vms] of the exchange
  uint64_t counts_capacity = 0;
  cudaEvent_t ev = nullptr;
  cudaEvent_t ev_packed = nullptr;   // recorded once the last collective has read everything it needs from the batch
  // one-sided path (zkb_push_*): IPC-mapped receive buffers + mailboxes of every rank
  bool push_ready = false;
  uint8_t* rbuf = nullptr;           // [world] exchange regions, then [world] gather regions
  uint64_t* mbox = nullptr;          // [world][16]
  uint64_t x_region = 0, g_region = 0;   // bytes per source rank: exchange region, gather region (sum of g_kind)
  uint64_t g_kind[ZKB_N_STREAMS] = {};     // bytes per source rank and gathered kind
  uint32_t g_mask = 0;
  zkb::PushPeers peers{};
  void* peer_opened[16] = {};
  uint32_t* d_kind_offs = nullptr;   // [6][n_vms] per-VM byte offsets of the gathered kinds
  uint64_t* d_kind_bytes = nullptr;  // [8]
  uint64_t* d_sent = nullptr;        // [8]
  uint64_t kind_offs_capacity = 0;
  uint64_t push_step = 0;
  uint64_t* h_mbox = nullptr;        // pinned mirror of the mailbox
};

#define NCCL_OK(c, expr)                                                                                          \
  do {                                                                                                            \
    ncclResult_t _r = (expr);                                                                                     \
    if (_r != ncclSuccess) return set_err(ZKB_ERR_CUDA, std::string(#expr) + ": " + (c)->api->GetErrorString(_r)); \
  } while (0)

static int32_t comm_ensure(uint8_t** buf, uint64_t* cap, uint64_t need) {
  if (need <= *cap) return ZKB_OK;
  CUDA_OK(cudaDeviceSynchronize());
  if (*buf) CUDA_OK(cudaFree(*buf));
  *buf = nullptr;
  *cap = 0;
  const uint64_t c = need + need / 8 + 4096;
  cudaError_t e = cudaMalloc(buf, c);
  if (e != cudaSuccess) return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("comm buffer cudaMalloc: ") + cudaGetErrorString(e));
  *cap = c;
  return ZKB_OK;
}

extern "C" {

int32_t zkb_comm_unique_id(uint8_t id_out[128]) {
  if (!id_out) return ZKB_ERR_INVALID_ARGUMENT;
  std::string why;
  zkb::NcclApi* api = zkb::nccl_api(&why);
  if (!api) return set_err(ZKB_ERR_CUDA, why);
  static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id size");
  ncclUniqueId id;
  ncclResult_t r = api->GetUniqueId(&id);
  if (r != ncclSuccess) return set_err(ZKB_ERR_CUDA, std::string("ncclGetUniqueId: ") + api->GetErrorString(r));
  memcpy(id_out, &id, 128);
  return ZKB_OK;
}

int32_t zkb_comm_create(int32_t device, int32_t rank, int32_t world, const uint8_t id[128], ZkbComm** out) {
  if (!id || !out || world < 1 || world > 8 || rank < 0 || rank >= world) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_comm_create: rank / world (1..8)");
  std::string why;
  zkb::NcclApi* api = zkb::nccl_api(&why);
  if (!api) return set_err(ZKB_ERR_CUDA, why);
  CUDA_OK(cudaSetDevice(device));
  ZkbComm* c = new ZkbComm();
  c->device = device;
  c->rank = rank;
  c->world = world;
  c->api = api;
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  ncclResult_t r = api->CommInitRank(&c->comm, world, uid, rank);
  if (r != ncclSuccess) {
    std::string msg = std::string("ncclCommInitRank: ") + api->GetErrorString(r);
    delete c;
    return set_err(ZKB_ERR_CUDA, msg);
  }
  cudaError_t e = cudaMalloc(&c->d_sizes, (size_t)world * 16 * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaHostAlloc(&c->h_sizes, (size_t)world * 16 * sizeof(uint64_t), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming);
  if (e != cudaSuccess) return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("zkb_comm_create: ") + cudaGetErrorString(e));
  *out = c;
  return ZKB_OK;
}

int32_t zkb_comm_destroy(ZkbComm* c) {
  if (!c) return ZKB_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (c->comm) c->api->CommDestroy(c->comm);
  if (c->recv) cudaFree(c->recv);
  if (c->concat) cudaFree(c->concat);
  if (c->send) cudaFree(c->send);
  if (c->d_sizes) cudaFree(c->d_sizes);
  if (c->h_sizes) cudaFreeHost(c->h_sizes);
  if (c->d_counts) cudaFree(c->d_counts);
  if (c->ev) cudaEventDestroy(c->ev);
  if (c->ev_packed) cudaEventDestroy(c->ev_packed);
  for (void* p : c->peer_opened)
    if (p) cudaIpcCloseMemHandle(p);
  if (c->rbuf) cudaFree(c->rbuf);
  if (c->mbox) cudaFree(c->mbox);
  if (c->d_kind_offs) cudaFree(c->d_kind_offs);
  if (c->d_kind_bytes) cudaFree(c->d_kind_bytes);
  if (c->d_sent) cudaFree(c->d_sent);
  if (c->h_mbox) cudaFreeHost(c->h_mbox);
  delete c;
  return ZKB_OK;
}

// One step of the multi-GPU path: (optionally) the balanced LOG exchange and (optionally) the concat of whole streams on
// dst_rank, with ONE size all-gather and ONE host synchronisation in front of all the transfers, which are then left in
// flight on `st`:
//   local prep   exchange: bucket count + scan;  gather: the pack kernels          (reads of the batch end here: ev_packed)
//   sizes        all-gather of 16 words per rank, D2H, stream sync
//   transfers    one grouped batch of ncclSend / ncclRecv for both
static int32_t comm_step(ZkbBatch* b, ZkbComm* c, bool do_exchange, uint32_t kinds_mask, int32_t dst_rank, void** share_out, uint64_t* n_share_out,
                         uint64_t* src_offsets_out, void** concat_out, uint64_t* concat_offsets_out, cudaStream_t st) {
  CUDA_OK(cudaSetDevice(c->device));
  const uint32_t* cnt = nullptr;
  int32_t rc = summary(b, &cnt);   // waits for THIS batch's run
  if (rc != ZKB_OK) return rc;
  const uint32_t n = b->cfg.n_vms, world = (uint32_t)c->world;
  uint64_t* h_mine = c->h_sizes + (size_t)c->rank * 16;
  uint64_t* d_mine = c->d_sizes + (size_t)c->rank * 16;
  memset(h_mine, 0, 128);
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device);
  const int grid = (int)std::max<uint32_t>(1u, std::min<uint32_t>((n + 3) / 4, (uint32_t)n_sm * 4));
  // ---- local prep ----
  uint8_t* packed[ZKB_N_STREAMS] = {};
  for (uint32_t k = 0; k < ZKB_N_STREAMS; k++) {
    if (!((kinds_mask >> k) & 1u)) continue;
    rc = pack_async(b, k, st, &packed[k], &h_mine[8 + k]);
    if (rc != ZKB_OK) return rc;
  }
  CUDA_OK(cudaMemcpyAsync(d_mine + 8, h_mine + 8, 64, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemsetAsync(d_mine, 0, 64, st));
  if (do_exchange) {
    if ((uint64_t)world * n > c->counts_capacity) {
      if (c->d_counts) CUDA_OK(cudaFree(c->d_counts));
      c->d_counts = nullptr;
      CUDA_OK(cudaMalloc(&c->d_counts, (size_t)world * n * 4));
      c->counts_capacity = (uint64_t)world * n;
    }
    zkb::zkb_bucket_count_kernel<<<grid, 128, 0, st>>>(b->d, world, c->d_counts);
    zkb::zkb_bucket_scan_kernel<<<world, 128, 0, st>>>(c->d_counts, n, d_mine);
    CUDA_OK(cudaGetLastError());
  }
  // ---- sizes ----
  NCCL_OK(c, c->api->AllGather(d_mine, c->d_sizes, 16, ncclUint64, c->comm, st));
  CUDA_OK(cudaMemcpyAsync(c->h_sizes, c->d_sizes, (size_t)world * 128, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  auto H = [&](uint32_t r, uint32_t i) { return c->h_sizes[(size_t)r * 16 + i]; };
  uint64_t send_total = 0, recv_total = 0;
  if (do_exchange) {
    for (uint32_t d = 0; d < world; d++) send_total += H(c->rank, d);
    for (uint32_t s = 0; s < world; s++) recv_total += H(s, c->rank);
    rc = comm_ensure(&c->send, &c->send_capacity, send_total * ZKB_LOG_BYTES);
    if (rc == ZKB_OK) rc = comm_ensure(&c->recv, &c->recv_capacity, recv_total * ZKB_LOG_BYTES);
    if (rc != ZKB_OK) return rc;
    if (send_total) zkb::zkb_bucket_pack_kernel<<<grid, 128, 0, st>>>(b->d, world, c->d_counts, d_mine, (uint32_t*)c->send);
    CUDA_OK(cudaGetLastError());
  }
  CUDA_OK(cudaEventRecord(c->ev_packed, st));   // nothing below reads the batch's stream slabs
  // concat layout on dst: for each kind in the mask (ascending), the ranks' shares in rank order; kinds start 256-byte aligned
  uint64_t kind_base[ZKB_N_STREAMS] = {}, at = 0;
  for (uint32_t k = 0; k < ZKB_N_STREAMS; k++) {
    if (!((kinds_mask >> k) & 1u)) continue;
    kind_base[k] = at;
    uint64_t run = 0;
    for (uint32_t r = 0; r < world; r++) {
      if (concat_offsets_out) concat_offsets_out[(size_t)k * (world + 1) + r] = at + run;
      run += H(r, 8 + k);
    }
    if (concat_offsets_out) concat_offsets_out[(size_t)k * (world + 1) + world] = at + run;
    at = (at + run + 255) / 256 * 256;
  }
  if (kinds_mask && c->rank == dst_rank) {
    rc = comm_ensure(&c->concat, &c->concat_capacity, at);
    if (rc != ZKB_OK) return rc;
  }
  // ---- transfers ----
  NCCL_OK(c, c->api->GroupStart());
  if (do_exchange) {
    uint64_t s_at = 0, r_at = 0;
    for (uint32_t p = 0; p < world; p++) {
      const uint64_t s_n = H(c->rank, p) * ZKB_LOG_BYTES, r_n = H(p, c->rank) * ZKB_LOG_BYTES;
      if (src_offsets_out) src_offsets_out[p] = r_at / ZKB_LOG_BYTES;
      if ((int)p == c->rank) {
        if (s_n) CUDA_OK(cudaMemcpyAsync(c->recv + r_at, c->send + s_at, s_n, cudaMemcpyDeviceToDevice, st));
      } else {
        if (s_n) NCCL_OK(c, c->api->Send(c->send + s_at, s_n, ncclUint8, (int)p, c->comm, st));
        if (r_n) NCCL_OK(c, c->api->Recv(c->recv + r_at, r_n, ncclUint8, (int)p, c->comm, st));
      }
      s_at += s_n;
      r_at += r_n;
    }
    if (src_offsets_out) src_offsets_out[world] = recv_total;
  }
  for (uint32_t k = 0; k < ZKB_N_STREAMS; k++) {
    if (!((kinds_mask >> k) & 1u)) continue;
    if (c->rank == dst_rank) {
      uint64_t run = kind_base[k];
      for (uint32_t r = 0; r < world; r++) {
        const uint64_t sz = H(r, 8 + k);
        if ((int)r == c->rank) {
          if (sz) CUDA_OK(cudaMemcpyAsync(c->concat + run, packed[k], sz, cudaMemcpyDeviceToDevice, st));
        } else if (sz) {
          NCCL_OK(c, c->api->Recv(c->concat + run, sz, ncclUint8, (int)r, c->comm, st));
        }
        run += sz;
      }
    } else if (H(c->rank, 8 + k)) {
      NCCL_OK(c, c->api->Send(packed[k], H(c->rank, 8 + k), ncclUint8, dst_rank, c->comm, st));
    }
  }
  NCCL_OK(c, c->api->GroupEnd());
  if (share_out) *share_out = do_exchange ? c->recv : nullptr;
  if (n_share_out) *n_share_out = recv_total;
  if (concat_out) *concat_out = (kinds_mask && c->rank == dst_rank) ? c->concat : nullptr;
  return ZKB_OK;
}

int32_t zkb_gather_streams(ZkbBatch* b, ZkbComm* c, uint32_t kinds_mask, int32_t dst_rank, void** dptr_out, uint64_t* offsets_out, void* cuda_stream) {
  if (!b || !c || !b->cfg.witness_mode || dst_rank < 0 || dst_rank >= c->world || !(kinds_mask & 63u)) return ZKB_ERR_INVALID_ARGUMENT;
  return comm_step(b, c, false, kinds_mask & 63u, dst_rank, nullptr, nullptr, nullptr, dptr_out, offsets_out, (cudaStream_t)cuda_stream);
}

int32_t zkb_comm_wait_packed(ZkbComm* c, void* cuda_stream) {
  if (!c) return ZKB_ERR_INVALID_ARGUMENT;
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaStreamWaitEvent((cudaStream_t)cuda_stream, c->ev_packed, 0));
  return ZKB_OK;
}

int32_t zkb_exchange_logs(ZkbBatch* b, ZkbComm* c, void** dptr_out, uint64_t* n_records_out, uint64_t* src_offsets_out, void* cuda_stream) {
  if (!b || !c || !b->cfg.witness_mode) return ZKB_ERR_INVALID_ARGUMENT;
  return comm_step(b, c, true, 0u, 0, dptr_out, n_records_out, src_offsets_out, nullptr, nullptr, (cudaStream_t)cuda_stream);
}

int32_t zkb_exchange_step(ZkbBatch* b, ZkbComm* c, uint32_t gather_kinds_mask, int32_t dst_rank, void** share_out, uint64_t* n_share_out,
                          uint64_t* src_offsets_out, void** concat_out, uint64_t* concat_offsets_out, void* cuda_stream) {
  if (!b || !c || !b->cfg.witness_mode || dst_rank < 0 || dst_rank >= c->world) return ZKB_ERR_INVALID_ARGUMENT;
  return comm_step(b, c, true, gather_kinds_mask & 63u, dst_rank, share_out, n_share_out, src_offsets_out, concat_out, concat_offsets_out,
                   (cudaStream_t)cuda_stream);
}

// ---- one-sided exchange over peer memory ---------------------------------------------------------------------------
// collective, once: receive regions sized for `b`'s capacities, IPC handles all-gathered through NCCL, peers mapped
static int32_t push_setup(ZkbBatch* b, ZkbComm* c, uint32_t gather_mask, cudaStream_t st) {
  const uint32_t world = (uint32_t)c->world;
  const uint64_t n = b->cfg.n_vms;
  // regions are sized from what this batch ACTUALLY emitted (x 1.5: a destination cannot receive more than a source's whole
  // log; per-VM slab capacities are several times that), the maximum over the ranks; a later, larger batch is caught by
  // the overflow flag of the mailbox row
  const uint32_t* cnt = nullptr;
  int32_t rc0 = summary(b, &cnt);
  if (rc0 != ZKB_OK) return rc0;
  uint64_t actual[ZKB_N_STREAMS] = {};
  for (uint64_t v = 0; v < n; v++)
    for (uint32_t k = 0; k < ZKB_N_STREAMS; k++) actual[k] += (uint64_t)cnt[v * 8 + k] * REC_BYTES[k];
  uint64_t* h = c->h_sizes;
  uint64_t* mine = h + (size_t)c->rank * 16;
  memset(mine, 0, 128);
  mine[0] = std::min<uint64_t>(n * b->cfg.cap_records[ZKB_STREAM_LOG] * ZKB_LOG_BYTES, actual[ZKB_STREAM_LOG] + actual[ZKB_STREAM_LOG] / 2 + (1u << 20));
  for (uint32_t k = 0; k < ZKB_N_STREAMS; k++)
    if ((gather_mask >> k) & 1u) mine[8 + k] = std::min<uint64_t>(n * b->cfg.cap_records[k] * REC_BYTES[k], actual[k] + actual[k] / 4 + (1u << 20));
  CUDA_OK(cudaMemcpyAsync(c->d_sizes + (size_t)c->rank * 16, mine, 128, cudaMemcpyHostToDevice, st));
  NCCL_OK(c, c->api->AllGather(c->d_sizes + (size_t)c->rank * 16, c->d_sizes, 16, ncclUint64, c->comm, st));
  CUDA_OK(cudaMemcpyAsync(h, c->d_sizes, (size_t)world * 128, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  c->x_region = 0;
  c->g_region = 0;
  c->g_mask = gather_mask;
  for (uint32_t k = 0; k < ZKB_N_STREAMS; k++) c->g_kind[k] = 0;
  for (uint32_t r = 0; r < world; r++) {
    c->x_region = std::max(c->x_region, h[(size_t)r * 16 + 0]);
    for (uint32_t k = 0; k < ZKB_N_STREAMS; k++) c->g_kind[k] = std::max(c->g_kind[k], h[(size_t)r * 16 + 8 + k]);
  }
  c->x_region = (c->x_region + 255) / 256 * 256;
  for (uint32_t k = 0; k < ZKB_N_STREAMS; k++) {
    c->g_kind[k] = (c->g_kind[k] + 255) / 256 * 256;
    c->g_region += c->g_kind[k];
  }
  cudaError_t e = cudaMalloc(&c->rbuf, (size_t)world * (c->x_region + c->g_region));
  if (e == cudaSuccess) e = cudaMalloc(&c->mbox, (size_t)world * 16 * 8);
  if (e == cudaSuccess) e = cudaMemset(c->mbox, 0, (size_t)world * 16 * 8);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_kind_bytes, 128);   // [0..8) bytes pushed per kind, [8..14) the per-kind region sizes
  if (e == cudaSuccess) e = cudaMalloc(&c->d_sent, 64);
  if (e == cudaSuccess) e = cudaHostAlloc(&c->h_mbox, (size_t)world * 16 * 8, cudaHostAllocDefault);
  if (e != cudaSuccess) return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("zkb push buffers: ") + cudaGetErrorString(e));
  // IPC handles of (rbuf, mbox) to everybody: 2 x 64 bytes per rank through the same all-gather
  struct Handles {
    cudaIpcMemHandle_t rbuf, mbox;
  };
  static_assert(sizeof(Handles) == 128, "two CUDA IPC handles");
  Handles my_handles;
  CUDA_OK(cudaIpcGetMemHandle(&my_handles.rbuf, c->rbuf));
  CUDA_OK(cudaIpcGetMemHandle(&my_handles.mbox, c->mbox));
  memcpy(h + (size_t)c->rank * 16, &my_handles, 128);
  CUDA_OK(cudaMemcpyAsync(c->d_sizes + (size_t)c->rank * 16, h + (size_t)c->rank * 16, 128, cudaMemcpyHostToDevice, st));
  NCCL_OK(c, c->api->AllGather(c->d_sizes + (size_t)c->rank * 16, c->d_sizes, 16, ncclUint64, c->comm, st));
  CUDA_OK(cudaMemcpyAsync(h, c->d_sizes, (size_t)world * 128, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  for (uint32_t r = 0; r < world; r++) {
    if ((int)r == c->rank) {
      c->peers.rbuf[r] = c->rbuf;
      c->peers.mbox[r] = c->mbox;
      continue;
    }
    Handles hr;
    memcpy(&hr, h + (size_t)r * 16, 128);
    void *p0 = nullptr, *p1 = nullptr;
    cudaError_t e0 = cudaIpcOpenMemHandle(&p0, hr.rbuf, cudaIpcMemLazyEnablePeerAccess);
    cudaError_t e1 = e0 == cudaSuccess ? cudaIpcOpenMemHandle(&p1, hr.mbox, cudaIpcMemLazyEnablePeerAccess) : e0;
    if (e1 != cudaSuccess) return set_err(ZKB_ERR_CUDA, std::string("cudaIpcOpenMemHandle (peer ") + std::to_string(r) + "): " + cudaGetErrorString(e1));
    c->peer_opened[2 * r] = p0;
    c->peer_opened[2 * r + 1] = p1;
    c->peers.rbuf[r] = (uint8_t*)p0;
    c->peers.mbox[r] = (uint64_t*)p1;
  }
  CUDA_OK(cudaMemcpy(c->d_kind_bytes + 8, c->g_kind, sizeof(c->g_kind), cudaMemcpyHostToDevice));
  c->push_ready = true;
  return ZKB_OK;
}

int32_t zkb_push_step(ZkbBatch* b, ZkbComm* c, uint32_t gather_kinds_mask, int32_t dst_rank, uint64_t* step_out, void* cuda_stream) {
  if (!b || !c || !b->cfg.witness_mode || dst_rank < 0 || dst_rank >= c->world) return ZKB_ERR_INVALID_ARGUMENT;
  const uint32_t allowed = 1u << ZKB_STREAM_DECOMMIT | 1u << ZKB_STREAM_FRAME | 1u << ZKB_STREAM_REFUND;
  if (gather_kinds_mask & ~allowed) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_push_step gathers DECOMMIT / FRAME / REFUND (whole-witness concat: zkb_gather_streams)");
  CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const uint32_t n = b->cfg.n_vms, world = (uint32_t)c->world;
  if (!c->push_ready) {
    int32_t rc = push_setup(b, c, gather_kinds_mask, st);
    if (rc != ZKB_OK) return rc;
  }
  if (gather_kinds_mask & ~c->g_mask) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_push_step: stream kinds outside the mask of the first call (the receive regions are sized then)");
  if ((uint64_t)world * n > c->counts_capacity) {
    if (c->d_counts) CUDA_OK(cudaFree(c->d_counts));
    c->d_counts = nullptr;
    CUDA_OK(cudaMalloc(&c->d_counts, (size_t)world * n * 4));
    c->counts_capacity = (uint64_t)world * n;
  }
  if ((uint64_t)ZKB_N_STREAMS * n > c->kind_offs_capacity) {
    if (c->d_kind_offs) CUDA_OK(cudaFree(c->d_kind_offs));
    c->d_kind_offs = nullptr;
    CUDA_OK(cudaMalloc(&c->d_kind_offs, (size_t)ZKB_N_STREAMS * n * 4));
    c->kind_offs_capacity = (uint64_t)ZKB_N_STREAMS * n;
  }
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device);
  const int grid = (int)std::max<uint32_t>(1u, std::min<uint32_t>((n + 3) / 4, (uint32_t)n_sm * 2));
  // NOTHING below waits on the host: record counts are read on the device, so the caller only has to order `st` behind
  // the batch's launch (an event) and may queue the next pass at once
  const uint64_t step = ++c->push_step;
  CUDA_OK(cudaMemsetAsync(c->d_sent, 0, 64, st));
  CUDA_OK(cudaMemsetAsync(c->d_kind_bytes, 0, 64, st));
  zkb::zkb_bucket_count_kernel<<<grid, 128, 0, st>>>(b->d, world, c->d_counts);
  zkb::zkb_bucket_scan_kernel<<<world, 128, 0, st>>>(c->d_counts, n, c->d_sent);
  zkb::zkb_bucket_push_kernel<<<grid, 128, 0, st>>>(b->d, world, (uint32_t)c->rank, c->d_counts, c->peers, c->x_region);
  uint64_t at = 0;
  for (uint32_t k = 0; k < ZKB_N_STREAMS; k++) {
    if (!((gather_kinds_mask >> k) & 1u)) continue;
    uint32_t* offs = c->d_kind_offs + (size_t)k * n;
    zkb::zkb_kind_count_kernel<<<std::max(1u, std::min<uint32_t>((n + 127) / 128, (uint32_t)n_sm)), 128, 0, st>>>(b->d, k, REC_BYTES[k], offs);
    zkb::zkb_bucket_scan_kernel<<<1, 128, 0, st>>>(offs, n, c->d_kind_bytes + k);
    uint8_t* region = c->peers.rbuf[dst_rank] + (size_t)world * c->x_region + (size_t)c->rank * c->g_region + at;
    zkb::zkb_kind_push_kernel<<<grid, 128, 0, st>>>(b->d, k, REC_BYTES[k], offs, region, c->g_kind[k]);
    at += c->g_kind[k];
  }
  zkb::zkb_mailbox_kernel<<<1, 32, 0, st>>>(c->peers, world, (uint32_t)c->rank, c->d_sent, (uint32_t)dst_rank, gather_kinds_mask, c->d_kind_bytes, step, c->x_region, c->d_kind_bytes + 8);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(c->ev_packed, st));
  if (step_out) *step_out = step;
  return ZKB_OK;
}

// Waits (host) until every source rank's mailbox row carries `step`, i.e. all of that step's pushes into THIS rank's
// regions have landed.  share_ptrs_out[s] / share_records_out[s]: source s's LOG records for this rank; on that step's sink
// also concat_ptrs_out[s * 6 + k] / concat_bytes_out[s * 6 + k] for the gathered kinds.  timeout_ms = 0: a single check
// (ZKB_ERR_CUDA "not complete" when rows are missing).
int32_t zkb_push_result(ZkbBatch* b, ZkbComm* c, uint64_t step, uint32_t timeout_ms, void** share_ptrs_out, uint64_t* share_records_out,
                        void** concat_ptrs_out, uint64_t* concat_bytes_out) {
  if (!b || !c || !c->push_ready) return ZKB_ERR_INVALID_ARGUMENT;
  CUDA_OK(cudaSetDevice(c->device));
  const uint32_t world = (uint32_t)c->world;
  const uint64_t n = b->cfg.n_vms;
  for (uint32_t waited = 0;; waited++) {
    CUDA_OK(cudaMemcpy(c->h_mbox, c->mbox, (size_t)world * 16 * 8, cudaMemcpyDeviceToHost));
    bool all = true;
    for (uint32_t s = 0; s < world; s++) all = all && c->h_mbox[(size_t)s * 16 + 15] >= step;
    if (all) break;
    if (waited >= timeout_ms * 10u) return set_err(ZKB_ERR_CUDA, "zkb_push_result: not complete");
    struct timespec ts = {0, 100000};
    nanosleep(&ts, nullptr);
  }
  for (uint32_t s = 0; s < world; s++) {
    const uint64_t* row = c->h_mbox + (size_t)s * 16;
    if (row[15] != step) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_push_result: a later step has already overwritten this one");
    if (row[2]) return set_err(ZKB_ERR_OUT_OF_MEMORY, "zkb_push_result: a receive region overflowed (the batch emitted more than the one the regions were sized from)");
    if (share_ptrs_out) share_ptrs_out[s] = c->rbuf + (size_t)s * c->x_region;
    if (share_records_out) share_records_out[s] = row[0];
    uint64_t at = 0;
    const bool to_me = row[14] == (uint64_t)c->rank;
    for (uint32_t k = 0; k < ZKB_N_STREAMS; k++) {
      const bool in_mask = ((row[1] >> k) & 1u) != 0;
      if (concat_ptrs_out) concat_ptrs_out[(size_t)s * 6 + k] = (to_me && in_mask) ? c->rbuf + (size_t)world * c->x_region + (size_t)s * c->g_region + at : nullptr;
      if (concat_bytes_out) concat_bytes_out[(size_t)s * 6 + k] = (to_me && in_mask) ? row[8 + k] : 0;
      if (in_mask) at += c->g_kind[k];
    }
  }
  return ZKB_OK;
}

}  // extern "C"
