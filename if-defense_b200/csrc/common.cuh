// common.cuh -- error plumbing, launch accounting and small device helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

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

inline cudaStream_t as_stream(ifd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace ifd
