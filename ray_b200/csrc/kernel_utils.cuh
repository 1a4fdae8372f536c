// kernel_utils.cuh — copy/stage helpers and launch plumbing shared by the collective kernels.
#pragma once
#include "comm.h"

namespace b200 {

// A message of `nbytes` is processed in 16-byte units.  `full` units are complete,
// a trailing partial unit (tail bytes) is staged zero-padded.
struct Units {
  size_t full;   // number of complete 16-byte units
  int tail;      // bytes in the trailing partial unit (0..15)
  __host__ __device__ size_t total() const { return full + (tail ? 1 : 0); }
};
__host__ __device__ inline Units make_units(size_t nbytes) {
  Units u;
  u.full = nbytes >> 4;
  u.tail = int(nbytes & 15);
  return u;
}

// Load unit `u` of a user tensor (arbitrary alignment handled by the slow path).
__device__ __forceinline__ uint4 load_user_unit(const char *src, size_t u, const Units &un, bool aligned) {
  if (u < un.full) {
    if (aligned) return ld_stream(src + (u << 4));
    uint4 v;
    unsigned char *b = reinterpret_cast<unsigned char *>(&v);
#pragma unroll
    for (int i = 0; i < 16; ++i) b[i] = reinterpret_cast<const unsigned char *>(src)[(u << 4) + i];
    return v;
  }
  uint4 v = make_uint4(0, 0, 0, 0);
  unsigned char *b = reinterpret_cast<unsigned char *>(&v);
  for (int i = 0; i < un.tail; ++i) b[i] = reinterpret_cast<const unsigned char *>(src)[(u << 4) + i];
  return v;
}

__device__ __forceinline__ void store_user_unit(char *dst, size_t u, const Units &un, bool aligned, uint4 v) {
  if (u < un.full) {
    if (aligned) {
      st_vec(dst + (u << 4), v);
      return;
    }
    const unsigned char *b = reinterpret_cast<const unsigned char *>(&v);
#pragma unroll
    for (int i = 0; i < 16; ++i) reinterpret_cast<unsigned char *>(dst)[(u << 4) + i] = b[i];
    return;
  }
  const unsigned char *b = reinterpret_cast<const unsigned char *>(&v);
  for (int i = 0; i < un.tail; ++i) reinterpret_cast<unsigned char *>(dst)[(u << 4) + i] = b[i];
}

__host__ __device__ inline bool is_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Offset of the staging slot used by the current launch (two slots alternate so a
// rank may start staging the next message while a slow peer still reads the previous
// one; see DESIGN.md "slot rotation").
__device__ __forceinline__ size_t staging_slot_offset(uint32_t launch, size_t staging_bytes) {
  return (launch & 1u) ? staging_bytes : 0;
}

inline int pick_blocks(const b200_comm *c, size_t work_items, int cap) {
  if (c->forced_blocks > 0) cap = c->forced_blocks;
  size_t want = work_items < 1 ? 1 : work_items;
  int g = int(want < size_t(cap) ? want : size_t(cap));
  if (g > kMaxBlocks) g = kMaxBlocks;
  return g < 1 ? 1 : g;
}

#define B200_LAUNCH_CHECK(c)                                                     \
  do {                                                                           \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) {                                                     \
      b200::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                      __FILE__, __LINE__);                                       \
      return B200_ERR_CUDA;                                                      \
    }                                                                            \
    (c)->launches.fetch_add(1);                                                  \
  } while (0)

// Dispatch helpers ---------------------------------------------------------------
#define B200_DISPATCH_DTYPE(dtype, T, ...)                                   \
  switch (dtype) {                                                           \
    case B200_U8: { using T = uint8_t; __VA_ARGS__; break; }                 \
    case B200_I8: { using T = int8_t; __VA_ARGS__; break; }                  \
    case B200_I32: { using T = int32_t; __VA_ARGS__; break; }                \
    case B200_U32: { using T = uint32_t; __VA_ARGS__; break; }               \
    case B200_I64: { using T = int64_t; __VA_ARGS__; break; }                \
    case B200_U64: { using T = uint64_t; __VA_ARGS__; break; }               \
    case B200_F16: { using T = __half; __VA_ARGS__; break; }                 \
    case B200_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }         \
    case B200_F32: { using T = float; __VA_ARGS__; break; }                  \
    case B200_F64: { using T = double; __VA_ARGS__; break; }                 \
    default: b200::set_error("unsupported dtype %d", int(dtype)); return B200_ERR_UNSUPPORTED; \
  }

#define B200_DISPATCH_OP(op, OP, ...)                                        \
  switch (op) {                                                              \
    case B200_SUM: { constexpr int OP = B200_SUM; __VA_ARGS__; break; }      \
    case B200_PROD: { constexpr int OP = B200_PROD; __VA_ARGS__; break; }    \
    case B200_MIN: { constexpr int OP = B200_MIN; __VA_ARGS__; break; }      \
    case B200_MAX: { constexpr int OP = B200_MAX; __VA_ARGS__; break; }      \
    case B200_AVG: { constexpr int OP = B200_AVG; __VA_ARGS__; break; }      \
    default: b200::set_error("unsupported reduce op %d", int(op)); return B200_ERR_UNSUPPORTED; \
  }

}  // namespace b200
