// Peer-memory plumbing for the fused statistics exchange.
//
// Every rank owns one small exchange buffer in its own HBM:
//     slots [2 parities][world][row_bytes]   row = 3 * C packets of 8 bytes {data word, stamp}
// and maps every peer's buffer through CUDA IPC (same node, NVLink / NVSwitch).
// The parameter step (step_epilogue.cuh: the tail of the fused reduction, or the stand-alone
// kernel) WRITES its row into slot [parity][rank] of every peer with 64-bit stores over
// NVLink — data and "ready" stamp in the same store, no fence — polls its own buffer until
// every peer's packets carry the current stamp and combines the rows in rank order: the
// collective is part of the kernel, there is no NCCL launch.
#include <string.h>

#include "p2p_internal.cuh"

using namespace qsb;

extern "C" int64_t qsb_p2p_group_bytes(int world, int64_t channels) {
  if (world < 1 || world > kMaxRanks || channels < 1) return 0;
  const int64_t row = p2p_row_bytes(channels);
  return 2 * world * row + 256;
}

// cudaMalloc (not the torch caching allocator: IPC handles need a whole allocation),
// zeroed, plus its 64-byte IPC handle to send to the peers.
extern "C" int qsb_p2p_alloc(int64_t bytes, void **dev_ptr, unsigned char *handle64) {
  if (bytes <= 0 || !dev_ptr || !handle64) return QSB_E_BADARG;
  QSB_CUDA_TRY(cudaMalloc(dev_ptr, (size_t)bytes));
  QSB_CUDA_TRY(cudaMemset(*dev_ptr, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  QSB_CUDA_TRY(cudaIpcGetMemHandle(&h, *dev_ptr));
  static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}

extern "C" int qsb_p2p_open(const unsigned char *handle64, void **peer_ptr) {
  if (!handle64 || !peer_ptr) return QSB_E_BADARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  QSB_CUDA_TRY(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int qsb_p2p_close(void *peer_ptr) {
  if (!peer_ptr) return 0;
  QSB_CUDA_TRY(cudaIpcCloseMemHandle(peer_ptr));
  return 0;
}

extern "C" int qsb_p2p_free(void *dev_ptr) {
  if (!dev_ptr) return 0;
  QSB_CUDA_TRY(cudaFree(dev_ptr));
  return 0;
}

extern "C" int qsb_p2p_group_create(qsb_p2p_group **out, int rank, int world,
                                    int64_t channels, void *const *bufs) {
  if (!out || world < 1 || world > kMaxRanks || rank < 0 || rank >= world ||
      channels < 1 || !bufs)
    return QSB_E_BADARG;
  qsb_p2p_group *g = new qsb_p2p_group();
  g->dev.rank = rank;
  g->dev.world = world;
  g->dev.row_bytes = p2p_row_bytes(channels);
  g->dev.timeout_ns = 30ull * 1000000000ull;  // qsb_p2p_group_set_timeout_ms
  g->channels = channels;
  for (int r = 0; r < world; ++r) {
    if (!bufs[r]) {
      delete g;
      return QSB_E_BADARG;
    }
    g->dev.bufs[r] = static_cast<unsigned char *>(bufs[r]);
  }
  if (cudaMalloc(&g->dev.error, sizeof(int)) != cudaSuccess ||
      cudaMemset(g->dev.error, 0, sizeof(int)) != cudaSuccess) {
    delete g;
    return (int)cudaGetLastError();
  }
  *out = g;
  return 0;
}

extern "C" int qsb_p2p_group_error(qsb_p2p_group *g, int *error_out) {
  if (!g || !error_out) return QSB_E_BADARG;
  QSB_CUDA_TRY(cudaMemcpy(error_out, g->dev.error, sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int qsb_p2p_group_set_timeout_ms(qsb_p2p_group *g, int64_t timeout_ms) {
  if (!g || timeout_ms <= 0) return QSB_E_BADARG;
  g->dev.timeout_ns = (unsigned long long)timeout_ms * 1000000ull;
  return 0;
}

// asynchronous form of qsb_p2p_group_error: the copy is ordered after everything queued on
// `stream` so far; *error_out_pinned is valid once the stream (or an event after this call)
// has completed.  Lets a training loop poll the flag without a synchronisation per step.
extern "C" int qsb_p2p_group_error_async(qsb_p2p_group *g, int *error_out_pinned, void *stream) {
  if (!g || !error_out_pinned) return QSB_E_BADARG;
  QSB_CUDA_TRY(cudaMemcpyAsync(error_out_pinned, g->dev.error, sizeof(int), cudaMemcpyDeviceToHost,
                               (cudaStream_t)stream));
  return 0;
}

extern "C" int qsb_p2p_group_destroy(qsb_p2p_group *g) {
  if (!g) return 0;
  cudaFree(g->dev.error);
  delete g;
  return 0;
}
