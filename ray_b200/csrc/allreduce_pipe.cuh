// allreduce_pipe.cuh — warp-specialised, flag-pipelined NVLS all-reduce for large staged messages.
//
// The phase-by-phase kernels run stage-in, the NVSwitch reduction and stage-out back to back,
// separated by cross-GPU barriers: the HBM copies (2 x 2S bytes) sit idle while the NVLink phase
// runs and vice versa.  Here every CTA has three 256-thread roles that run CONCURRENTLY on the
// same SM and are coupled only through per-tile flags in the signal pads:
//
//   role 0  stager-in : copy tile t of the user tensor into the symmetric slot, then
//                       release-store flag in[t][me] into every rank's pad
//   role 1  reducer   : wait in[t][p] for all p; multimem.ld_reduce the slice of tile t this rank
//                       owns; multimem.st it to every rank; release-store out[t][me] everywhere
//   role 2  stager-out: wait out[t][p] for all p; copy tile t from the slot to the user tensor
//
// While the reducer warps of an SM wait on NVLink latency, its stager warps keep HBM busy, so
// the staging traffic hides behind the NVLink phase.  A tile is n*256 16-byte units; rank r
// owns units [r*256,(r+1)*256) of it.  Flags carry the launch epoch (monotonic): never reset.
// CTA b's roles all work on tiles b, b+G, ...; the stager-in role never waits, so the data flow
// cannot deadlock however the CTAs are scheduled.
#pragma once
#include "allreduce_core.cuh"

namespace b200 {

constexpr int kPipeRole = 256;             // threads per role
constexpr int kPipeThreads = 3 * kPipeRole;  // CTA size of the pipelined kernels

__device__ __forceinline__ void role_bar(int role) {
  asm volatile("bar.sync %0, %1;" ::"r"(role + 1), "n"(kPipeRole) : "memory");
}

// All threads of `role` call.  Threads with `mine` poll `flag` until >= epoch.  `ok_smem` is the
// role's shared word.  Returns false if any thread of the role gave up (abort / watchdog).
__device__ __forceinline__ bool role_wait_flags(const DevComm &c, int role, int *ok_smem, bool mine,
                                                const uint32_t *flag, uint32_t epoch) {
  if (mine && !wait_flag_ge(c, flag, epoch)) *ok_smem = 0;
  role_bar(role);
  return *ok_smem != 0;
}

template <typename T, int OP, typename LoadFn, typename StoreFn>
__device__ __forceinline__ void allreduce_pipelined_nvls(const DevComm &c, uint32_t epoch, size_t off,
                                                         size_t U, LoadFn load, StoreFn store) {
  using Tr = Traits<T>;
  __shared__ int ok[3];
  const int n = c.world, r = c.rank;
  const int role = threadIdx.x / kPipeRole, t = threadIdx.x % kPipeRole;
  const size_t G = gridDim.x, b = blockIdx.x;
  const size_t tile_units = size_t(n) * kPipeRole;
  const size_t T_tiles = (U + tile_units - 1) / tile_units;
  if (threadIdx.x < 3) ok[threadIdx.x] = 1;
  __syncthreads();

  if (role == 0) {
    // ---------------- stager-in: 8 loads in flight per thread.  Software pipelined: the loads of
    // batch k+1 are issued before the release fence of batch k, so the fence drains behind them.
    char *mine = c.data[r] + off;
    const int tpi = n >= 8 ? 1 : 8 / n;  // tiles per batch
    uint4 v[8];
    auto issue_loads = [&](size_t first) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int j = s / n, q = s - j * n;
        const size_t tile = first + size_t(j) * G;
        const size_t u = tile * tile_units + size_t(q) * kPipeRole + t;
        if (j < tpi && tile < T_tiles && u < U) v[s] = load(u);
      }
    };
    size_t first = b;
    if (first < T_tiles) issue_loads(first);
    while (first < T_tiles) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int j = s / n, q = s - j * n;
        const size_t tile = first + size_t(j) * G;
        const size_t u = tile * tile_units + size_t(q) * kPipeRole + t;
        if (j < tpi && tile < T_tiles && u < U) st_vec(mine + (u << 4), v[s]);
      }
      const size_t next = first + size_t(tpi) * G;
      if (next < T_tiles) issue_loads(next);
      role_bar(0);
      if (t < n * tpi) {
        const int j = t / n, p = t - j * n;
        const size_t tile = first + size_t(j) * G;
        if (tile < T_tiles) st_release_sys(c.sig[p] + kSigTileIn + tile * kMaxRanks + r, epoch);
      }
      first = next;
    }
  } else if (role == 1) {
    // ---------------- reducer: 8 multimem.ld_reduce in flight per thread, software pipelined the
    // same way (the ld_reduce of batch k+1 are in flight while the multimem.st of batch k drain).
    constexpr int UNR = 8;
    char *mc = c.mc_data + off;
    const uint32_t *my_in = c.sig[r] + kSigTileIn;
    uint4 v[UNR];
    auto wait_and_load = [&](size_t first) -> bool {
      const int j = t / n, p = t - j * n;
      const size_t ftile = first + size_t(j) * G;
      const bool mine_flag = t < n * UNR && ftile < T_tiles;
      if (!role_wait_flags(c, 1, &ok[1], mine_flag, my_in + ftile * kMaxRanks + p, epoch)) return false;
#pragma unroll
      for (int jj = 0; jj < UNR; ++jj) {
        const size_t tile = first + size_t(jj) * G;
        const size_t u = tile * tile_units + size_t(r) * kPipeRole + t;
        if (tile < T_tiles && u < U) v[jj] = Multimem<T>::ld_reduce_sum(mc + (u << 4));
      }
      return true;
    };
    size_t first = b;
    if (first < T_tiles && !wait_and_load(first)) return;
    while (first < T_tiles) {
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        const size_t tile = first + size_t(j) * G;
        const size_t u = tile * tile_units + size_t(r) * kPipeRole + t;
        if (tile < T_tiles && u < U) {
          if (OP == B200_AVG) {
            typename Tr::Acc acc = Tr::unpack(v[j]);
            Tr::average(acc, n);
            v[j] = Tr::pack(acc);
          }
          multimem_st(mc + (u << 4), v[j]);
        }
      }
      const size_t next = first + size_t(UNR) * G;
      if (next < T_tiles) {
        if (!wait_and_load(next)) return;  // contains the role barrier that orders the stores above
      } else {
        role_bar(1);
      }
      if (t < n * UNR) {
        const int j = t / n, p = t - j * n;
        const size_t tile = first + size_t(j) * G;
        if (tile < T_tiles) st_release_sys(c.sig[p] + kSigTileOut + tile * kMaxRanks + r, epoch);
      }
      first = next;
    }
  } else {
    // ---------------- stager-out
    const char *mine = c.data[r] + off;
    const uint32_t *my_out = c.sig[r] + kSigTileOut;
    const int tpi = n >= 8 ? 1 : 8 / n;
    for (size_t k = 0;; ++k) {
      const size_t first = (k * tpi) * G + b;
      if (first >= T_tiles) break;
      {
        const int j = t / n, p = t - j * n;
        const size_t tile = first + size_t(j) * G;
        const bool mine_flag = t < n * tpi && tile < T_tiles;
        if (!role_wait_flags(c, 2, &ok[2], mine_flag, my_out + tile * kMaxRanks + p, epoch)) return;
      }
      uint4 v[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int j = s / n, q = s - j * n;
        const size_t tile = first + size_t(j) * G;
        const size_t u = tile * tile_units + size_t(q) * kPipeRole + t;
        if (j < tpi && tile < T_tiles && u < U) v[s] = ld_peer(mine + (u << 4));
      }
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int j = s / n, q = s - j * n;
        const size_t tile = first + size_t(j) * G;
        const size_t u = tile * tile_units + size_t(q) * kPipeRole + t;
        if (j < tpi && tile < T_tiles && u < U) store(u, v[s]);
      }
    }
  }
}

// Same completion protocol as finish_launch(), for CTAs whose roles return independently.
__device__ __forceinline__ void finish_launch_pipe(const DevComm &c) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    uint32_t t = atomicAdd(&c.st->done_ctr, 1u);
    if (t == gridDim.x - 1) {
      c.st->done_ctr = 0;
      __threadfence();
      atomicAdd(&c.st->launch_ctr, 1u);
    }
  }
}

inline size_t pipe_tiles(size_t U, int world) {
  const size_t tile_units = size_t(world) * kPipeRole;
  return (U + tile_units - 1) / tile_units;
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
