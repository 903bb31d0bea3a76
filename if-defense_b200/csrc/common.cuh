// common.cuh -- error plumbing, launch accounting and small device helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "../../include/ifd_b200.h"

namespace ifd {

std::string& last_error_ref();
long long& launch_counter_ref();

inline int fail(int code, const std::string& msg) {
  last_error_ref() = msg;
  return code;
}
inline int cuda_fail(cudaError_t e, const char* what) {
  last_error_ref() = std::string(what) + ": " + cudaGetErrorString(e);
  return IFD_ERR_CUDA;
}
inline void count_launch(int n = 1) { launch_counter_ref() += n; }

#define IFD_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return ::ifd::cuda_fail(_e, #expr); \
  } while (0)

// call after every kernel launch
#define IFD_LAUNCH_CHECK(name)                                      \
  do {                                                              \
    ::ifd::count_launch();                                          \
    cudaError_t _e = cudaGetLastError();                            \
    if (_e != cudaSuccess) return ::ifd::cuda_fail(_e, "launch " name); \
  } while (0)

#define IFD_REQUIRE(cond, msg)                                        \
  do {                                                                \
    if (!(cond)) return ::ifd::fail(IFD_ERR_INVALID, std::string(msg)); \
  } while (0)

// optional per-kernel event timing (api.cu); `kind` as documented at ifd_profile_read
struct ProfileScope {
  ProfileScope(int kind, cudaStream_t st);
  ~ProfileScope();
  int kind_;
  cudaStream_t st_;
  cudaEvent_t e0_ = nullptr;
};

// Raise a kernel's dynamic shared-memory limit once per (kernel, device): the setting is sticky, and the runtime call
// takes the context lock -- made before every one of the ~400 launches of a restoration it made the host the
// bottleneck of the end-to-end path (worse with NCCL's threads contending for the same lock).
inline cudaError_t set_max_dyn_smem(const void* kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> g(mu);
  size_t& cur = done[std::make_pair(kernel, dev)];
  if (bytes <= cur) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

// Every device pointer of one restoration loop.  The loop's kernels can take their pointers from such a record in device
// memory instead of from their launch arguments, so that ONE instantiated CUDA graph of the ~400 launches serves any set of
// buffers: a launch of the loop is then "write the record, launch the graph" (restore.cu).
struct LoopJob {
  const float* planes;   // [3][B][R][R][C] channels-last
  const float* W;        // packed decoder blob
  const float* Wimg;     // UMMA weight images (workspace)
  float* xyz;            // [B][K][3] in / out
  float* g_occ;          // [B][K][3] (workspace)
  float* m;              // Adam state (workspace)
  float* v;
  int32_t* nbr;          // [B][K][8] warm-start neighbour lists (workspace)
  float* jac;            // [B][K][3][32] d c / d xyz of the decode kernel's forward gather (workspace), or nullptr
};
bool profile_on();       // api.cu: per-kernel event timing is enabled on this thread

inline cudaStream_t as_stream(ifd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace ifd
