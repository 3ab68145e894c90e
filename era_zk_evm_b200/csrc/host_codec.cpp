// Host half of the transport codec (include/zkb_codec.h): the decoder entry points of the C ABI.  Plain C++ (no CUDA):
// a consumer that only receives blobs links or dlopens the same library and needs no GPU for this part.
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/zkb.h"
#include "../../include/zkb_codec.h"

extern "C" {

int32_t zkb_decode_stream(const void* blob, uint64_t blob_bytes, uint32_t vm, uint32_t kind, void* dst, uint64_t max_bytes, uint64_t* n_bytes) {
  zkb_codec::EncodedView v;
  if (!v.open(blob, blob_bytes)) return ZKB_ERR_INVALID_ARGUMENT;
  const uint64_t n = v.decode(vm, kind, dst, max_bytes);
  if (n == UINT64_MAX) return ZKB_ERR_INVALID_ARGUMENT;
  if (n_bytes) *n_bytes = n;
  return ZKB_OK;
}

int32_t zkb_decode_counts(const void* blob, uint64_t blob_bytes, uint32_t vm, uint32_t counts_out[8]) {
  zkb_codec::EncodedView v;
  if (!v.open(blob, blob_bytes) || vm >= v.n_vms() || !counts_out) return ZKB_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < 8; i++) counts_out[i] = v.counts(vm)[i];
  return ZKB_OK;
}

// bulk decode: stream `kind` of EVERY VM as canonical records, VM-major and back to back -- the layout
// zkb_fetch_stream_packed produces -- on n_threads host threads (0 = hardware concurrency).  offsets_out[n_vms + 1]
// receives each VM's byte offset in dst; dst == NULL: only the offsets (and through offsets_out[n_vms] the total).
int32_t zkb_decode_all(const void* blob, uint64_t blob_bytes, uint32_t kind, void* dst, uint64_t capacity, uint64_t* offsets_out,
                       uint32_t n_threads) {
  zkb_codec::EncodedView v;
  if (!v.open(blob, blob_bytes) || kind >= ZKB_N_STREAMS || !offsets_out) return ZKB_ERR_INVALID_ARGUMENT;
  const uint32_t n = v.n_vms();
  const uint64_t rec = (uint64_t)ZKB_CODEC_REC_WORDS[kind] * 4;
  offsets_out[0] = 0;
  for (uint32_t vm = 0; vm < n; vm++) offsets_out[vm + 1] = offsets_out[vm] + (uint64_t)v.counts(vm)[kind] * rec;
  if (!dst) return ZKB_OK;
  if (offsets_out[n] > capacity) return ZKB_ERR_INVALID_ARGUMENT;
  uint32_t t = n_threads ? n_threads : std::max(1u, std::thread::hardware_concurrency());
  t = std::max(1u, std::min(t, n));
  std::vector<int> bad(t, 0);
  auto work = [&](uint32_t id) {
    // contiguous VM ranges of roughly equal BYTES per thread
    const uint64_t lo_b = offsets_out[n] / t * id, hi_b = id + 1 == t ? offsets_out[n] + 1 : offsets_out[n] / t * (id + 1);
    for (uint32_t vm = 0; vm < n; vm++) {
      if (offsets_out[vm] < lo_b || offsets_out[vm] >= hi_b) continue;
      const uint64_t len = offsets_out[vm + 1] - offsets_out[vm];
      if (v.decode(vm, kind, (uint8_t*)dst + offsets_out[vm], len) != len) bad[id] = 1;
    }
  };
  std::vector<std::thread> th;
  for (uint32_t i = 1; i < t; i++) th.emplace_back(work, i);
  work(0);
  for (auto& x : th) x.join();
  for (int b : bad)
    if (b) return ZKB_ERR_INVALID_ARGUMENT;
  return ZKB_OK;
}

}  // extern "C"
