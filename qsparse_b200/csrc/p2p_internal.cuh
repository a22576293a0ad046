#pragma once
#include "qsb_common.cuh"

namespace qsb {

constexpr int kMaxRanks = 16;

inline int64_t p2p_row_bytes(int64_t channels) { return (channels * 12 + 255) / 256 * 256; }

// what the kernel needs (passed by value)
struct P2PDev {
  int rank, world;
  int64_t row_bytes;
  unsigned char *bufs[kMaxRanks];  // bufs[rank] is the local buffer
  int *error;                      // set to 1 if a peer's stamp never arrived
};

__host__ __device__ inline int64_t p2p_slot_offset(const P2PDev &d, int parity, int r) {
  return ((int64_t)parity * d.world + r) * d.row_bytes;
}
__host__ __device__ inline int64_t p2p_flag_offset(const P2PDev &d, int parity, int r) {
  return 2 * (int64_t)d.world * d.row_bytes +
         ((int64_t)parity * kMaxRanks + r) * (int64_t)sizeof(unsigned long long);
}

}  // namespace qsb

struct qsb_p2p_group {
  qsb::P2PDev dev;
  int64_t channels;
};
