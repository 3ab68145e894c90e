// ASan fuzz harness for the host decoders of include/zkb_codec.h: reads a blob, flips bytes, decodes everything
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include "../include/zkb_codec.h"
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> blob(n); if (fread(blob.data(), 1, n, f) != (size_t)n) return 2; fclose(f);
  const ZkbEncodedHeader* h0 = (const ZkbEncodedHeader*)blob.data();
  std::mt19937_64 rng(42);
  long ok = 0, bad = 0;
  const int trials = argc > 2 ? atoi(argv[2]) : 2000;
  for (int t = 0; t < trials; t++) {
    std::vector<uint8_t> b = blob;
    // exact-size heap copy so that ASan sees any overread of the blob
    int flips = 1 + rng() % 6;
    uint64_t lo = (t % 4 == 0) ? sizeof(ZkbEncodedHeader) : h0->payload_offset[0];   // every 4th trial also hits the tables
    for (int i = 0; i < flips; i++) b[lo + rng() % (n - lo)] = (uint8_t)rng();
    zkb_codec::EncodedView v;
    if (!v.open(b.data(), b.size())) { bad++; continue; }
    for (uint32_t vm = 0; vm < v.n_vms(); vm++)
      for (uint32_t k = 0; k < ZKB_N_STREAMS; k++)
        for (int ref = 0; ref < 2; ref++) {
          uint64_t need = v.decode(vm, k, nullptr, 0);
          if (need == UINT64_MAX || need > (64u << 20)) { bad++; continue; }
          std::vector<uint8_t> out(need);
          uint64_t got = v.decode(vm, k, out.data(), need, ref != 0);
          (got == UINT64_MAX ? bad : ok)++;
        }
  }
  printf("decodes ok %ld rejected %ld\n", ok, bad);
  return 0;
}
