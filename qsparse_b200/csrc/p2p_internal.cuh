#pragma once
#include "qsb_common.cuh"

namespace qsb {

constexpr int kMaxRanks = 16;

// one rank's statistics row as 8-byte {data, flag} packets: 3 words per channel (low / high
// half of the fp64 sum, bits of max|x|), see step_epilogue.cuh
inline int64_t p2p_row_bytes(int64_t channels) { return (channels * 24 + 255) / 256 * 256; }

// what the kernel needs (passed by value)
struct P2PDev {
  int rank, world;
  int64_t row_bytes;
  unsigned char *bufs[kMaxRanks];  // bufs[rank] is the local buffer
  int *error;                      // set to 1 when a peer's packets did not arrive in time
  unsigned long long timeout_ns;   // how long the kernel waits for a peer before it poisons the step
};

__host__ __device__ inline int64_t p2p_slot_offset(const P2PDev &d, int parity, int r) {
  return ((int64_t)parity * d.world + r) * d.row_bytes;
}

}  // namespace qsb

struct qsb_p2p_group {
  qsb::P2PDev dev;
  int64_t channels;
};
