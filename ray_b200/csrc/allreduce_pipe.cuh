// allreduce_pipe.cuh — role-specialised, flag-pipelined all-reduce for large staged messages.
//
// The three phases of the staged all-reduce (stage-in to the symmetric slot, reduce+publish over
// NVLink / NVSwitch, stage-out to the user tensor) run CONCURRENTLY on disjoint sets of CTAs
// of one launch instead of back to back separated by grid-wide barriers:
//
//   CTAs [0, g_in)              stagers-in : copy tile t of the user tensor into the slot, then
//                                            release-store flag in[t][me] into every rank's pad
//   CTAs [g_in, g_in+g_red)     reducers   : wait in[t][p] for all p, reduce the slice this rank
//                                            owns of tile t (peer loads or multimem.ld_reduce),
//                                            publish it to every rank (peer stores / multimem.st),
//                                            release-store flag out[t][me] into every rank's pad
//   CTAs [g_in+g_red, G)        stagers-out: wait out[t][p] for all p, copy tile t to the user
//
// A tile is one row of the row geometry (n*512 16-byte units; rank r owns units [r*512,(r+1)*512)).
// Flags hold the launch epoch (monotonic), so they are never reset.  HBM staging traffic thus
// overlaps the NVLink phase; the only serial parts left are the first tile in and the last out.
//
// Scheduling order = data-flow order (stagers-in have the lowest CTA indices), so a resident
// consumer CTA always has its producers resident or finished: no deadlock even if the grid
// is not fully co-resident.
#pragma once
#include "allreduce_core.cuh"

namespace b200 {

struct PipeSplit {
  int g_in, g_red, g_out;
};

// all threads call; threads with `mine` poll `flag` until >= epoch.  Returns false on abort/timeout.
__device__ __forceinline__ bool cta_wait_flags(const DevComm &c, bool mine, const uint32_t *flag, uint32_t epoch) {
  int ok = 1;
  if (mine) ok = wait_flag_ge(c, flag, epoch) ? 1 : 0;
  return __syncthreads_and(ok) != 0;
}

template <typename T, int OP, bool NVLS, typename LoadFn, typename StoreFn>
__device__ __forceinline__ void allreduce_pipelined(const DevComm &c, uint32_t epoch, size_t off,
                                                    const RowGeom &g, PipeSplit sp, LoadFn load, StoreFn store) {
  using Tr = Traits<T>;
  const int n = c.world, r = c.rank, t = threadIdx.x, b = blockIdx.x;
  const size_t T_tiles = g.R;
  uint32_t *my_in = c.sig[r] + kSigTileIn;
  uint32_t *my_out = c.sig[r] + kSigTileOut;

  if (b < sp.g_in) {
    // ---------------- stager-in: up to 8 loads in flight per thread, tiles signalled one by one
    char *mine = c.data[r] + off;
    const int tiles_per_iter = n >= 8 ? 1 : 8 / n;
    for (size_t t0 = size_t(b) * tiles_per_iter; t0 < T_tiles; t0 += size_t(sp.g_in) * tiles_per_iter) {
      uint4 v[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int j = s / n, k = s - j * n;
        const size_t u = (t0 + j) * g.row_units + size_t(k) * kThreads + t;
        if (j < tiles_per_iter && t0 + j < T_tiles && u < g.U) v[s] = load(u);
      }
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int j = s / n, k = s - j * n;
        const size_t u = (t0 + j) * g.row_units + size_t(k) * kThreads + t;
        if (j < tiles_per_iter && t0 + j < T_tiles && u < g.U) st_vec(mine + (u << 4), v[s]);
      }
      __syncthreads();
      if (t < n * tiles_per_iter) {
        const int j = t / n, p = t - j * n;
        if (t0 + j < T_tiles) st_release_sys(c.sig[p] + kSigTileIn + (t0 + j) * kMaxRanks + r, epoch);
      }
    }
  } else if (b < sp.g_in + sp.g_red) {
    // ---------------- reducer
    constexpr int UNR = NVLS ? 4 : 2;
    const int i = b - sp.g_in;
    char *mc = NVLS ? c.mc_data + off : nullptr;
    for (size_t t0 = i; t0 < T_tiles; t0 += size_t(sp.g_red) * UNR) {
      // wait until every rank staged the tiles of this batch
      {
        const int j = t / n, p = t - j * n;
        const size_t tile = t0 + size_t(j) * sp.g_red;
        const bool mine = t < n * UNR && tile < T_tiles;
        if (!cta_wait_flags(c, mine, my_in + tile * kMaxRanks + p, epoch)) return;
      }
      if (NVLS) {
        uint4 v[UNR];
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
          const size_t tile = t0 + size_t(j) * sp.g_red;
          const size_t u = tile * g.row_units + size_t(r) * kThreads + t;
          if (tile < T_tiles && u < g.U) v[j] = Multimem<T>::ld_reduce_sum(mc + (u << 4));
        }
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
          const size_t tile = t0 + size_t(j) * sp.g_red;
          const size_t u = tile * g.row_units + size_t(r) * kThreads + t;
          if (tile < T_tiles && u < g.U) {
            if (OP == B200_AVG) {
              typename Tr::Acc acc = Tr::unpack(v[j]);
              Tr::average(acc, n);
              v[j] = Tr::pack(acc);
            }
            multimem_st(mc + (u << 4), v[j]);
          }
        }
      } else {
        uint4 v[UNR][kMaxRanks];
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
          const size_t tile = t0 + size_t(j) * sp.g_red;
          const size_t u = tile * g.row_units + size_t(r) * kThreads + t;
          if (tile < T_tiles && u < g.U) {
#pragma unroll
            for (int p = 0; p < kMaxRanks; ++p)
              if (p < n) v[j][p] = ld_peer(c.data[p] + off + (u << 4));
          }
        }
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
          const size_t tile = t0 + size_t(j) * sp.g_red;
          const size_t u = tile * g.row_units + size_t(r) * kThreads + t;
          if (tile < T_tiles && u < g.U) {
            typename Tr::Acc acc = Tr::unpack(v[j][0]);
#pragma unroll
            for (int p = 1; p < kMaxRanks; ++p)
              if (p < n) Tr::template reduce<OP>(acc, Tr::unpack(v[j][p]));
            if (OP == B200_AVG) Tr::average(acc, n);
            const uint4 res = Tr::pack(acc);
#pragma unroll
            for (int q = 0; q < kMaxRanks; ++q) {
              if (q < n) {
                int p = r + q;
                if (p >= n) p -= n;
                st_vec(c.data[p] + off + (u << 4), res);
              }
            }
          }
        }
      }
      __syncthreads();
      if (t < n * UNR) {
        const int j = t / n, p = t - j * n;
        const size_t tile = t0 + size_t(j) * sp.g_red;
        if (tile < T_tiles) st_release_sys(c.sig[p] + kSigTileOut + tile * kMaxRanks + r, epoch);
      }
    }
  } else {
    // ---------------- stager-out
    const int i = b - sp.g_in - sp.g_red;
    const char *mine = c.data[r] + off;
    const int tiles_per_iter = n >= 8 ? 1 : 8 / n;
    for (size_t t0 = size_t(i) * tiles_per_iter; t0 < T_tiles; t0 += size_t(sp.g_out) * tiles_per_iter) {
      {
        const int j = t / n, p = t - j * n;
        const bool mine_flag = t < n * tiles_per_iter && t0 + j < T_tiles;
        if (!cta_wait_flags(c, mine_flag, my_out + (t0 + j) * kMaxRanks + p, epoch)) return;
      }
      uint4 v[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int j = s / n, k = s - j * n;
        const size_t u = (t0 + j) * g.row_units + size_t(k) * kThreads + t;
        if (j < tiles_per_iter && t0 + j < T_tiles && u < g.U) v[s] = ld_peer(mine + (u << 4));
      }
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int j = s / n, k = s - j * n;
        const size_t u = (t0 + j) * g.row_units + size_t(k) * kThreads + t;
        if (j < tiles_per_iter && t0 + j < T_tiles && u < g.U) store(u, v[s]);
      }
    }
  }
}

// Role split for a grid of G CTAs: explicit B200_PARAM_PIPE_CTAS_IN/OUT, else env
// B200_PIPE_SPLIT="in,out", else one sixth of the grid on each staging side.
inline bool pick_split(const b200_comm *c, int G, PipeSplit *sp) {
  static int env_in = -1, env_out = -1;
  static bool parsed = false;
  if (!parsed) {
    parsed = true;
    const char *s = getenv("B200_PIPE_SPLIT");
    if (s) sscanf(s, "%d,%d", &env_in, &env_out);
  }
  int gin = int(c->params[B200_PARAM_PIPE_CTAS_IN]), gout = int(c->params[B200_PARAM_PIPE_CTAS_OUT]);
  if (gin <= 0) gin = env_in;
  if (gout <= 0) gout = env_out;
  if (gin <= 0) gin = G / 6;
  if (gout <= 0) gout = G / 6;
  if (gin < 1 || gout < 1 || gin + gout + 1 > G) return false;
  *sp = PipeSplit{gin, G - gin - gout, gout};
  return true;
}

inline size_t pipe_min_bytes(const b200_comm *c) {
  static long long env = [] {
    const char *s = getenv("B200_PIPE_MIN_BYTES");
    return s ? atoll(s) : -1ll;
  }();
  long long v = c->params[B200_PARAM_PIPE_MIN_BYTES];
  if (v < 0) v = env;
  if (v < 0) v = (long long)(8 << 20);
  return size_t(v);
}

}  // namespace b200
