// Streaming elementwise "map" kernel shared by every fake-quant / STE / mask op.
//
// One persistent grid (SMs x resident CTAs), each thread moves U vectors of V
// floats per iteration: all loads are issued first (U x 32 B in flight per
// thread), then the arithmetic, then the stores.  HBM-bound: 4 B read + 4 B
// written per element and per stream, nothing is staged in shared memory
// because no element is touched twice.
//
// Per-channel parameters are derived on the fly from the raw device-side
// parameter arrays (decimal / scale / lines / channel mask) — no parameter
// prep launch; the arrays are tiny and L1/L2 resident.
#pragma once
#include "qsb_common.cuh"

namespace qsb {

struct MapIO {
  const float *in0;
  const float *in1;
  const uint8_t *inb;
  float *out0;
  float *out1;
  uint8_t *outb;
};

// Op requirements:
//   struct P;                                  per-channel derived constants
//   static constexpr bool kIn1, kInB, kOut0, kOut1, kOutB, kCanSkip;
//   __device__ P params(int32_t c) const;
//   __device__ bool skip(const P&) const;      (kCanSkip) output independent of in0
//   __device__ void apply(float a, float b, uint8_t mb, const P&,
//                         float &o0, float &o1, uint8_t &ob) const;

template <class Op, int V, Hint LH, Hint SH>
struct MapVec {
  VecF<V> a, b;
  VecB<V> m;
  bool loaded;
};

template <class Op, int V, int U, bool CHAN, Hint LH, Hint SH>
__global__ void __launch_bounds__(QSB_THREADS)
    map_kernel(Op op, MapIO io, int64_t n, Layout L, ChanStep step_u,
               ChanStep step_iter) {
  using P = typename Op::P;
  constexpr int64_t kTile = (int64_t)QSB_THREADS * V * U;
  const int64_t n_main = (n / V) * V;
  const int64_t stride_u = (int64_t)QSB_THREADS * V;

  P p_tensor;
  if constexpr (!CHAN) p_tensor = op.params(0);

  int64_t e_base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * V;
  ChanPos pos_base;
  if constexpr (CHAN) pos_base = chan_pos_of(e_base < n ? e_base : 0, L);

  for (; e_base < n_main; e_base += (int64_t)gridDim.x * kTile) {
    VecF<V> a[U], b[U];
    VecB<V> mb[U];
    P p0[U], p1[U];
    bool mixed[U], skipv[U];
    ChanPos pos = pos_base;
    // ---- phase 1: parameters + loads -----------------------------------
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = e_base + u * stride_u;
      mixed[u] = false;
      skipv[u] = false;
      if constexpr (CHAN) {
        if (e < n_main) {
          if (L.inner >= V) {
            p0[u] = op.params(pos.c);
            mixed[u] = (pos.col + V > L.inner);
            if (mixed[u]) {
              int32_t c1 = pos.c + 1;
              if (c1 >= (int32_t)L.channels) c1 = 0;
              p1[u] = op.params(c1);
            } else {
              p1[u] = p0[u];
            }
            if constexpr (Op::kCanSkip)
              skipv[u] = op.skip(p0[u]) && op.skip(p1[u]);
          }
        }
        advance(pos, step_u, L);
      }
      if (e < n_main) {
        if (!skipv[u]) a[u] = ld_vec<V, LH>(io.in0 + e);
        if constexpr (Op::kIn1) b[u] = ld_vec<V, LH>(io.in1 + e);
        if constexpr (Op::kInB) mb[u] = ld_bytes<V>(io.inb + e);
      }
    }
    // ---- phase 2: arithmetic + stores ----------------------------------
    pos = pos_base;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = e_base + u * stride_u;
      if (e < n_main) {
        VecF<V> o0, o1;
        VecB<V> ob;
        if (CHAN && L.inner < V) {
          // rows shorter than one vector: walk the channels element by element
          ChanPos q = pos;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            P pj = op.params(q.c);
            op.apply(a[u].v[j], Op::kIn1 ? b[u].v[j] : 0.f,
                     Op::kInB ? mb[u].b[j] : (uint8_t)1, pj, o0.v[j], o1.v[j],
                     ob.b[j]);
            q.col += 1;
            if (q.col >= L.inner) {
              q.col = 0;
              q.c += 1;
              if (q.c >= (int32_t)L.channels) q.c = 0;
            }
          }
        } else {
          const int64_t left = CHAN ? (L.inner - pos.col) : (int64_t)V;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const P &pj = CHAN ? ((mixed[u] && j >= left) ? p1[u] : p0[u])
                               : p_tensor;
            float av = a[u].v[j];
            if (Op::kCanSkip && CHAN && skipv[u]) av = 0.f;
            op.apply(av, Op::kIn1 ? b[u].v[j] : 0.f,
                     Op::kInB ? mb[u].b[j] : (uint8_t)1, pj, o0.v[j], o1.v[j],
                     ob.b[j]);
          }
        }
        if constexpr (Op::kOut0) st_vec<V, SH>(io.out0 + e, o0);
        if constexpr (Op::kOut1) st_vec<V, SH>(io.out1 + e, o1);
        if constexpr (Op::kOutB) st_bytes<V>(io.outb + e, ob);
      }
      if constexpr (CHAN) advance(pos, step_u, L);
    }
    if constexpr (CHAN) advance(pos_base, step_iter, L);
  }

  // ---- tail: the last n % V elements, scalar, by block 0 -----------------
  if (blockIdx.x == 0) {
    const int64_t e = n_main + threadIdx.x;
    if (e < n) {
      P pj;
      if constexpr (CHAN)
        pj = op.params(chan_pos_of(e, L).c);
      else
        pj = p_tensor;
      float o0, o1;
      uint8_t ob;
      op.apply(io.in0[e], Op::kIn1 ? io.in1[e] : 0.f,
               Op::kInB ? io.inb[e] : (uint8_t)1, pj, o0, o1, ob);
      if constexpr (Op::kOut0) io.out0[e] = o0;
      if constexpr (Op::kOut1) io.out1[e] = o1;
      if constexpr (Op::kOutB) io.outb[e] = ob;
    }
  }
}

// Tuning knobs (benchmark use; defaults are what the product uses).
struct MapTuning {
  int ctas_per_sm;  // 0: occupancy-derived persistent grid; -1: one CTA per tile
};
MapTuning &map_tuning();

template <class Op, int V, int U, bool CHAN, Hint LH, Hint SH>
int launch_map_variant(const Op &op, const MapIO &io, int64_t n,
                       const Layout &L, cudaStream_t stream) {
  if (n <= 0) return 0;
  auto kern = map_kernel<Op, V, U, CHAN, LH, SH>;
  constexpr int64_t kTile = (int64_t)QSB_THREADS * V * U;
  static int occ = 0;  // per instantiation
  if (occ == 0) {
    int o = 0;
    QSB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern,
                                                               QSB_THREADS, 0));
    occ = o > 0 ? o : 1;
  }
  const int64_t tiles = (n + kTile - 1) / kTile;
  int per_sm = map_tuning().ctas_per_sm;
  int64_t grid;
  if (per_sm < 0)
    grid = tiles;
  else
    grid = (int64_t)device_props().sm_count * (per_sm > 0 ? per_sm : occ);
  if (grid > tiles) grid = tiles;
  if (grid < 1) grid = 1;
  ChanStep su = make_chan_step((int64_t)QSB_THREADS * V, L);
  ChanStep si = make_chan_step(grid * kTile, L);
  kern<<<(unsigned)grid, QSB_THREADS, 0, stream>>>(op, io, n, L, su, si);
  QSB_LAUNCH_CHECK();
  return 0;
}

// Picks the widest vector the pointers allow (32 B -> V=8, 16 B -> V=4, else
// scalar) and per-tensor vs per-channel addressing.
template <class Op, Hint LH = Hint::KEEP, Hint SH = Hint::STREAM>
int launch_map(const Op &op, const MapIO &io, const Layout &L,
               cudaStream_t stream) {
  const int64_t n = L.numel();
  if (n <= 0) return 0;
  uintptr_t bits = 0;
  auto acc = [&](const void *p, int scale) {
    // byte streams advance 1 B per element: a V-element vector needs V-byte
    // alignment, i.e. the same element alignment as 4V bytes of floats.
    if (p) bits |= reinterpret_cast<uintptr_t>(p) * scale;
  };
  acc(io.in0, 1);
  if (Op::kIn1) acc(io.in1, 1);
  if (Op::kOut0) acc(io.out0, 1);
  if (Op::kOut1) acc(io.out1, 1);
  if (Op::kInB) acc(io.inb, 4);
  if (Op::kOutB) acc(io.outb, 4);
  if (bits & 3) return QSB_E_ALIGN;
  const bool chan = L.channels > 1;
  if ((bits & 31) == 0) {
    return chan ? launch_map_variant<Op, 8, 2, true, LH, SH>(op, io, n, L, stream)
                : launch_map_variant<Op, 8, 2, false, LH, SH>(op, io, n, L, stream);
  } else if ((bits & 15) == 0) {
    return chan ? launch_map_variant<Op, 4, 4, true, LH, SH>(op, io, n, L, stream)
                : launch_map_variant<Op, 4, 4, false, LH, SH>(op, io, n, L, stream);
  }
  return chan ? launch_map_variant<Op, 1, 4, true, LH, SH>(op, io, n, L, stream)
              : launch_map_variant<Op, 1, 4, false, LH, SH>(op, io, n, L, stream);
}

}  // namespace qsb
