// fsb200 — host-side helpers shared by every translation unit of libfsb200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fsb200.h"

namespace fsb {

// Thread-local error string returned by fsb_last_error().
void set_error(const char* fmt, ...);

#define FSB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::fsb::set_error(__VA_ARGS__);      \
      return FSB_ERR_INVALID;             \
    }                                     \
  } while (0)

#define FSB_CUDA_LAUNCH_CHECK()                                                       \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      ::fsb::set_error("%s:%d CUDA launch error: %s", __FILE__, __LINE__,             \
                       cudaGetErrorString(e__));                                      \
      return FSB_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

int num_sms();

// 2-D / 3-D bf16 tensor map with 128B swizzle. dims/box innermost-first; strides (bytes) for dims 1.. .
// Returns 0 on success (error string set otherwise).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace fsb
