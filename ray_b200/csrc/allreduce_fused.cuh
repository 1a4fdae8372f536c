// allreduce_fused.cuh — staged NVLS all-reduce with the HBM staging hidden behind the NVLink phase.
//
// The phase-by-phase kernel runs stage-in (HBM-bound), the NVSwitch reduction (link-bound) and
// stage-out (HBM-bound) one after the other.  Because every CTA follows the same schedule the
// phases line up across the whole GPU and the HBM pipes idle while NVLink works and vice versa.
// Here each thread interleaves the three streams row by row: in iteration k it has in flight
//
//     the staging loads of row k+1          (user tensor -> registers)
//     the multimem.ld_reduce of row k       (its own 16-byte unit of the row)
//     the staging loads of row k-1          (reduced slot -> registers)
//
// and then issues the three matching stores.  The switch reduction is throughput-bound (a
// ld_reduce takes tens of microseconds under load), so the HBM traffic completes in its shadow.
//
// Synchronisation is per CTA and per row, with two monotonically increasing counters per
// (CTA, peer) in the signal pads: rows-in[b][p] ("rank p's CTA b staged that many rows") and
// rows-out[b][p] ("... published that many").  Iteration k waits for its peers' iteration k-1,
// which in turn waited for this rank's iteration k-2: no cycle.  Counters continue across
// launches (cumulative count in LocalState::fused_rows), so nothing is ever reset.
#pragma once
#include "allreduce_core.cuh"

namespace b200 {

// threads [0,n) poll one flag array, threads [n,2n) the other; `need_*` disables a side.
__device__ __forceinline__ bool fused_wait(const DevComm &c, bool need_in, uint32_t want_in, bool need_out,
                                           uint32_t want_out) {
  const int n = c.world, t = threadIdx.x;
  int ok = 1;
  if (t < n) {
    if (need_in) ok = wait_flag_ge(c, c.sig[c.rank] + kSigRowsIn + size_t(blockIdx.x) * kMaxRanks + t, want_in);
  } else if (t < 2 * n) {
    if (need_out)
      ok = wait_flag_ge(c, c.sig[c.rank] + kSigRowsOut + size_t(blockIdx.x) * kMaxRanks + (t - n), want_out);
  }
  return __syncthreads_and(ok) != 0;
}

// after __syncthreads: threads [0,n) publish rows-in, threads [n,2n) rows-out (one release each)
__device__ __forceinline__ void fused_signal(const DevComm &c, bool sig_in, uint32_t val_in, bool sig_out,
                                             uint32_t val_out) {
  const int n = c.world, t = threadIdx.x;
  __syncthreads();
  if (t < n) {
    if (sig_in) st_release_sys(c.sig[t] + kSigRowsIn + size_t(blockIdx.x) * kMaxRanks + c.rank, val_in);
  } else if (t < 2 * n) {
    if (sig_out) st_release_sys(c.sig[t - n] + kSigRowsOut + size_t(blockIdx.x) * kMaxRanks + c.rank, val_out);
  }
}

template <typename T, int OP, typename LoadFn, typename StoreFn>
__device__ __forceinline__ void allreduce_fused_nvls(const DevComm &c, size_t off, const RowGeom &g, LoadFn load,
                                                     StoreFn store) {
  using Tr = Traits<T>;
  const int n = c.world, r = c.rank, t = threadIdx.x;
  const size_t G = gridDim.x, b = blockIdx.x;
  const size_t K = b < g.R ? (g.R - b + G - 1) / G : 0;  // rows b, b+G, ... of this CTA
  const uint32_t base = c.st->fused_rows[b];
  char *mine = c.data[r] + off;
  char *mc = c.mc_data + off;
  if (K == 0) return;

  // prologue: stage row 0
  {
    uint4 v[kMaxRanks];
    const size_t u0 = b * g.row_units + t;
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q) {
      const size_t u = u0 + size_t(q) * kThreads;
      if (q < n && u < g.U) v[q] = load(u);
    }
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q) {
      const size_t u = u0 + size_t(q) * kThreads;
      if (q < n && u < g.U) st_vec(mine + (u << 4), v[q]);
    }
  }
  fused_signal(c, true, base + 1, false, 0);

  for (size_t k = 0; k < K; ++k) {
    const size_t row = b + k * G;
    const bool has_next = k + 1 < K, has_prev = k > 0;
    // loads of the next row's staging (independent of any flag)
    uint4 vin[kMaxRanks];
    const size_t un0 = (row + G) * g.row_units + t;
    if (has_next) {
#pragma unroll
      for (int q = 0; q < kMaxRanks; ++q) {
        const size_t u = un0 + size_t(q) * kThreads;
        if (q < n && u < g.U) vin[q] = load(u);
      }
    }
    // every rank staged row k; every rank published row k-1
    if (!fused_wait(c, true, base + uint32_t(k) + 1, has_prev, base + uint32_t(k))) return;
    const size_t ur = row * g.row_units + size_t(r) * kThreads + t;
    uint4 vr;
    if (ur < g.U) vr = Multimem<T>::ld_reduce_sum(mc + (ur << 4));
    uint4 vout[kMaxRanks];
    const size_t up0 = (row - G) * g.row_units + t;
    if (has_prev) {
#pragma unroll
      for (int q = 0; q < kMaxRanks; ++q) {
        const size_t u = up0 + size_t(q) * kThreads;
        if (q < n && u < g.U) vout[q] = ld_peer(mine + (u << 4));
      }
    }
    // the three matching stores
    if (has_next) {
#pragma unroll
      for (int q = 0; q < kMaxRanks; ++q) {
        const size_t u = un0 + size_t(q) * kThreads;
        if (q < n && u < g.U) st_vec(mine + (u << 4), vin[q]);
      }
    }
    if (ur < g.U) {
      if (OP == B200_AVG) {
        typename Tr::Acc acc = Tr::unpack(vr);
        Tr::average(acc, n);
        vr = Tr::pack(acc);
      }
      multimem_st(mc + (ur << 4), vr);
    }
    if (has_prev) {
#pragma unroll
      for (int q = 0; q < kMaxRanks; ++q) {
        const size_t u = up0 + size_t(q) * kThreads;
        if (q < n && u < g.U) store(u, vout[q]);
      }
    }
    fused_signal(c, has_next, base + uint32_t(k) + 2, true, base + uint32_t(k) + 1);
  }

  // epilogue: every rank published the last row; copy it out
  if (!fused_wait(c, false, 0, true, base + uint32_t(K))) return;
  {
    uint4 v[kMaxRanks];
    const size_t u0 = (b + (K - 1) * G) * g.row_units + t;
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q) {
      const size_t u = u0 + size_t(q) * kThreads;
      if (q < n && u < g.U) v[q] = ld_peer(mine + (u << 4));
    }
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q) {
      const size_t u = u0 + size_t(q) * kThreads;
      if (q < n && u < g.U) store(u, v[q]);
    }
  }
  __syncthreads();
  if (t == 0) c.st->fused_rows[b] = base + uint32_t(K);
}

}  // namespace b200
