// api.cu -- library-level plumbing: error string, launch counter, device probe, and the host-buffer
// convenience call that is the end-to-end seam (H2D -> layout conversion -> loop -> D2H).
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace ifd {
std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}
long long& launch_counter_ref() {
  static thread_local long long n = 0;
  return n;
}

struct ProfRec { int kind; cudaEvent_t e0, e1; };
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec>* g_prof = nullptr;

bool profile_on() { return g_prof_on; }

ProfileScope::ProfileScope(int kind, cudaStream_t st) : kind_(kind), st_(st) {
  if (!g_prof_on) return;
  if (cudaEventCreate(&e0_) != cudaSuccess) { e0_ = nullptr; return; }
  cudaEventRecord(e0_, st_);
}
ProfileScope::~ProfileScope() {
  if (!e0_) return;
  cudaEvent_t e1;
  if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0_); return; }
  cudaEventRecord(e1, st_);
  if (!g_prof) g_prof = new std::vector<ProfRec>();
  g_prof->push_back({kind_, e0_, e1});
}

struct HostCache {
  void* dev = nullptr;
  size_t bytes = 0;
  float* pinned = nullptr;
  size_t pinned_bytes = 0;
  cudaStream_t stream = nullptr;
  int device = -1;          // the device `dev` and `stream` belong to
};
static thread_local HostCache g_cache;

static int current_device(int* dev) {
  IFD_CUDA_TRY(cudaGetDevice(dev));
  return IFD_OK;
}

static int ensure_cache(size_t dev_bytes, size_t pinned_bytes) {
  int dev = 0;
  if (int rc = current_device(&dev)) return rc;
  if (g_cache.device != dev && (g_cache.stream || g_cache.dev)) {   // the caller switched devices: start over on this one
    if (g_cache.dev) cudaFree(g_cache.dev);
    if (g_cache.stream) cudaStreamDestroy(g_cache.stream);
    g_cache.dev = nullptr;
    g_cache.bytes = 0;
    g_cache.stream = nullptr;
  }
  g_cache.device = dev;
  if (!g_cache.stream) IFD_CUDA_TRY(cudaStreamCreateWithFlags(&g_cache.stream, cudaStreamNonBlocking));
  if (g_cache.bytes < dev_bytes) {
    if (g_cache.dev) cudaFree(g_cache.dev);
    g_cache.dev = nullptr;
    g_cache.bytes = 0;
    IFD_CUDA_TRY(cudaMalloc(&g_cache.dev, dev_bytes));
    g_cache.bytes = dev_bytes;
  }
  if (g_cache.pinned_bytes < pinned_bytes) {
    if (g_cache.pinned) cudaFreeHost(g_cache.pinned);
    g_cache.pinned = nullptr;
    g_cache.pinned_bytes = 0;
    IFD_CUDA_TRY(cudaMallocHost((void**)&g_cache.pinned, pinned_bytes));
    g_cache.pinned_bytes = pinned_bytes;
  }
  return IFD_OK;
}
}  // namespace ifd

using namespace ifd;

extern "C" const char* ifd_last_error(void) { return last_error_ref().c_str(); }
extern "C" int ifd_abi_version(void) { return IFD_ABI_VERSION; }
extern "C" long long ifd_launch_count(int reset) {
  const long long n = launch_counter_ref();
  if (reset) launch_counter_ref() = 0;
  return n;
}
extern "C" int ifd_device_cc(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    cudaGetLastError();
    return fail(IFD_ERR_CUDA, "no CUDA device");
  }
  return prop.major * 10 + prop.minor;
}

extern "C" void ifd_profile_enable(int on) { g_prof_on = on != 0; }
extern "C" int ifd_profile_read(double* ms_out, long long* launches_out) {
  for (int k = 0; k < IFD_PROFILE_KINDS; ++k) {
    if (ms_out) ms_out[k] = 0.0;
    if (launches_out) launches_out[k] = 0;
  }
  if (!g_prof) return IFD_PROFILE_KINDS;
  cudaDeviceSynchronize();
  for (const ProfRec& r : *g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess && r.kind >= 0 && r.kind < IFD_PROFILE_KINDS) {
      if (ms_out) ms_out[r.kind] += ms;
      if (launches_out) launches_out[r.kind] += 1;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof->clear();
  return IFD_PROFILE_KINDS;
}

namespace ifd { void release_pipe(); void release_graphs(); }
extern "C" void ifd_release_cache(void) {
  release_pipe();
  release_graphs();
  if (g_cache.dev) cudaFree(g_cache.dev);
  if (g_cache.pinned) cudaFreeHost(g_cache.pinned);
  if (g_cache.stream) cudaStreamDestroy(g_cache.stream);
  g_cache = HostCache();
}

extern "C" int ifd_convonet_opt_host(const float* planes_nchw_host, const float* dec_weights_host, float* xyz_host,
                                     int B, int K, int R, int C, int H, int n_blocks, const ifd_opt_params* P,
                                     double* stats_out_host) {
  IFD_REQUIRE(planes_nchw_host && dec_weights_host && xyz_host && P && B > 0 && K > 0 && R > 0 && C > 0,
              "ifd_convonet_opt_host: bad arguments");
  const size_t nw = ifd_convonet_decoder_nfloats(C, H, n_blocks);
  if (nw == 0) return fail(IFD_ERR_UNSUPPORTED, "ifd_convonet_opt_host: unsupported decoder shape");
  const size_t plane_bytes = (size_t)3 * B * C * R * R * sizeof(float);
  const size_t xyz_bytes = (size_t)B * K * 3 * sizeof(float);
  const size_t w_bytes = align_up(nw * sizeof(float), 256);
  const int n_stat = P->want_stats && stats_out_host ? (P->n_steps > 0 ? (P->n_steps - 1) / 100 + 1 : 0) : 0;
  const size_t stat_bytes = align_up((size_t)n_stat * 4 * sizeof(double) + 8, 256);
  const size_t ws_bytes = ifd_convonet_opt_workspace_bytes(B, K);
  const size_t total = 2 * align_up(plane_bytes, 256) + w_bytes + align_up(xyz_bytes, 256) + stat_bytes + ws_bytes;
  int rc = ensure_cache(total, 0);
  if (rc) return rc;
  char* base = (char*)g_cache.dev;
  float* d_nchw = (float*)base; base += align_up(plane_bytes, 256);
  float* d_cl = (float*)base; base += align_up(plane_bytes, 256);
  float* d_w = (float*)base; base += w_bytes;
  float* d_xyz = (float*)base; base += align_up(xyz_bytes, 256);
  double* d_stat = (double*)base; base += stat_bytes;
  void* d_ws = base;
  cudaStream_t st = g_cache.stream;
  IFD_CUDA_TRY(cudaMemcpyAsync(d_nchw, planes_nchw_host, plane_bytes, cudaMemcpyHostToDevice, st));
  IFD_CUDA_TRY(cudaMemcpyAsync(d_w, dec_weights_host, nw * sizeof(float), cudaMemcpyHostToDevice, st));
  IFD_CUDA_TRY(cudaMemcpyAsync(d_xyz, xyz_host, xyz_bytes, cudaMemcpyHostToDevice, st));
  if ((rc = ifd_planes_nchw_to_cl(d_nchw, d_cl, 3 * B, C, R, st))) return rc;
  if ((rc = ifd_convonet_opt(d_cl, d_w, d_xyz, nullptr, nullptr, B, K, R, C, H, n_blocks, P, n_stat ? d_stat : nullptr, d_ws,
                             ws_bytes, st)))
    return rc;
  IFD_CUDA_TRY(cudaMemcpyAsync(xyz_host, d_xyz, xyz_bytes, cudaMemcpyDeviceToHost, st));
  if (n_stat) IFD_CUDA_TRY(cudaMemcpyAsync(stats_out_host, d_stat, (size_t)n_stat * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
  IFD_CUDA_TRY(cudaStreamSynchronize(st));
  return IFD_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Many batches, pipelined: what defend_point_cloud (ConvONet/opt_defense.py:272-312) does batch after batch, with the
// host<->device copies overlapped with the loops, and the loops of up to four consecutive batches running side by side (one
// run stream per loop, one buffer slot more than loops in flight, copy streams for H2D and D2H).
// Every batch still pays its own H2D of planes + init points and D2H of the restored cloud.
namespace ifd {
int opt_lanes();                    // restore.cu: loops side by side (ifd_test_hook(2, n), default 4)
void opt_side_by_side(int on);      // restore.cu: the loops enqueued next run next to others (selects the one-CTA tail)
constexpr int kPipeRun = 4;         // run streams (>= opt_lanes())
constexpr int kPipeSlots = kPipeRun + 1;     // buffer slots: one more than loops in flight, for the upload / download of the next
struct PipeCache {
  void* dev = nullptr;
  size_t bytes = 0;
  cudaStream_t h2d = nullptr, run[kPipeRun] = {}, d2h = nullptr;
  cudaEvent_t ready[kPipeSlots] = {}, done[kPipeSlots] = {}, freed[kPipeSlots] = {};
};
static thread_local PipeCache g_pipe;
static thread_local int g_pipe_device = -1;
void release_pipe();
static int ensure_pipe(size_t dev_bytes) {
  int dev = 0;
  if (int rc = current_device(&dev)) return rc;
  if (g_pipe_device != dev && (g_pipe.h2d || g_pipe.dev)) release_pipe();      // streams / buffers of another device
  g_pipe_device = dev;
  if (!g_pipe.h2d) {
    IFD_CUDA_TRY(cudaStreamCreateWithFlags(&g_pipe.h2d, cudaStreamNonBlocking));
    for (int r = 0; r < kPipeRun; ++r) IFD_CUDA_TRY(cudaStreamCreateWithFlags(&g_pipe.run[r], cudaStreamNonBlocking));
    IFD_CUDA_TRY(cudaStreamCreateWithFlags(&g_pipe.d2h, cudaStreamNonBlocking));
    for (int s = 0; s < kPipeSlots; ++s) {
      IFD_CUDA_TRY(cudaEventCreateWithFlags(&g_pipe.ready[s], cudaEventDisableTiming));
      IFD_CUDA_TRY(cudaEventCreateWithFlags(&g_pipe.done[s], cudaEventDisableTiming));
      IFD_CUDA_TRY(cudaEventCreateWithFlags(&g_pipe.freed[s], cudaEventDisableTiming));
    }
  }
  if (g_pipe.bytes < dev_bytes) {
    if (g_pipe.dev) cudaFree(g_pipe.dev);
    g_pipe.dev = nullptr;
    g_pipe.bytes = 0;
    IFD_CUDA_TRY(cudaMalloc(&g_pipe.dev, dev_bytes));
    g_pipe.bytes = dev_bytes;
  }
  return IFD_OK;
}
}  // namespace ifd

extern "C" int ifd_convonet_opt_host_batches(int n_batches, const float* const* planes_nchw_host, const float* dec_weights_host,
                                             float* const* xyz_host, int B, int K, int R, int C, int H, int n_blocks,
                                             const ifd_opt_params* P) {
  IFD_REQUIRE(n_batches >= 0 && planes_nchw_host && dec_weights_host && xyz_host && P && B > 0 && K > 0 && R > 0 && C > 0,
              "ifd_convonet_opt_host_batches: bad arguments");
  if (n_batches == 0) return IFD_OK;
  for (int j = 0; j < n_batches; ++j)
    IFD_REQUIRE(planes_nchw_host[j] && xyz_host[j], "ifd_convonet_opt_host_batches: null batch pointer");
  const size_t nw = ifd_convonet_decoder_nfloats(C, H, n_blocks);
  if (nw == 0) return fail(IFD_ERR_UNSUPPORTED, "ifd_convonet_opt_host_batches: unsupported decoder shape");
  const size_t plane_bytes = align_up((size_t)3 * B * C * R * R * sizeof(float), 256);
  const size_t xyz_bytes = (size_t)B * K * 3 * sizeof(float);
  const size_t xyz_al = align_up(xyz_bytes, 256);
  const size_t w_bytes = align_up(nw * sizeof(float), 256);
  const size_t ws_bytes = align_up(ifd_convonet_opt_workspace_bytes(B, K), 256);
  const int lanes = opt_lanes() < 1 ? 1 : (opt_lanes() > kPipeRun ? kPipeRun : opt_lanes());
  const int slots = lanes + 1;
  const size_t total = w_bytes + (size_t)lanes * ws_bytes + (size_t)slots * (2 * plane_bytes + xyz_al);
  int rc = ensure_pipe(total);
  if (rc) return rc;
  char* base = (char*)g_pipe.dev;
  float* d_w = (float*)base; base += w_bytes;
  void* d_ws[kPipeRun];                     // one workspace per run stream: the loops of `lanes` batches run side by side
  for (int r = 0; r < lanes; ++r) {
    d_ws[r] = base;
    base += ws_bytes;
  }
  float *d_nchw[kPipeSlots], *d_cl[kPipeSlots], *d_xyz[kPipeSlots];    // `lanes` batches run while one more is uploaded / downloaded
  for (int s = 0; s < slots; ++s) {
    d_nchw[s] = (float*)base; base += plane_bytes;
    d_cl[s] = (float*)base; base += plane_bytes;
    d_xyz[s] = (float*)base; base += xyz_al;
  }
  IFD_CUDA_TRY(cudaMemcpyAsync(d_w, dec_weights_host, nw * sizeof(float), cudaMemcpyHostToDevice, g_pipe.h2d));
  opt_side_by_side(lanes > 1 && n_batches > 1);
  for (int j = 0; j < n_batches && rc == IFD_OK; ++j) {
    const int s = j % slots, r = j % lanes;   // buffer slot, run stream
    cudaError_t e = cudaSuccess;
    if (j >= slots) e = cudaStreamWaitEvent(g_pipe.h2d, g_pipe.freed[s], 0);       // slot s drained (loop + D2H of j - slots)
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d_nchw[s], planes_nchw_host[j], (size_t)3 * B * C * R * R * sizeof(float), cudaMemcpyHostToDevice, g_pipe.h2d);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_xyz[s], xyz_host[j], xyz_bytes, cudaMemcpyHostToDevice, g_pipe.h2d);
    if (e != cudaSuccess) { rc = cuda_fail(e, "ifd_convonet_opt_host_batches: upload"); break; }
    if ((rc = ifd_planes_nchw_to_cl(d_nchw[s], d_cl[s], 3 * B, C, R, g_pipe.h2d))) break;
    e = cudaEventRecord(g_pipe.ready[s], g_pipe.h2d);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g_pipe.run[r], g_pipe.ready[s], 0);
    if (e != cudaSuccess) { rc = cuda_fail(e, "ifd_convonet_opt_host_batches: hand-over"); break; }
    if ((rc = ifd_convonet_opt(d_cl[s], d_w, d_xyz[s], nullptr, nullptr, B, K, R, C, H, n_blocks, P, nullptr, d_ws[r], ws_bytes,
                               g_pipe.run[r])))
      break;
    e = cudaEventRecord(g_pipe.done[s], g_pipe.run[r]);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g_pipe.d2h, g_pipe.done[s], 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(xyz_host[j], d_xyz[s], xyz_bytes, cudaMemcpyDeviceToHost, g_pipe.d2h);
    if (e == cudaSuccess) e = cudaEventRecord(g_pipe.freed[s], g_pipe.d2h);
    if (e != cudaSuccess) rc = cuda_fail(e, "ifd_convonet_opt_host_batches: download");
  }
  opt_side_by_side(0);
  // drain on every exit path: the buffers are cached and the next call reuses them
  cudaError_t es = cudaStreamSynchronize(g_pipe.d2h);
  for (int r = 0; r < lanes; ++r) {
    const cudaError_t er = cudaStreamSynchronize(g_pipe.run[r]);
    if (es == cudaSuccess) es = er;
  }
  const cudaError_t eh = cudaStreamSynchronize(g_pipe.h2d);
  if (es == cudaSuccess) es = eh;
  if (rc == IFD_OK && es != cudaSuccess) rc = cuda_fail(es, "ifd_convonet_opt_host_batches: synchronize");
  return rc;
}

namespace ifd {
void release_pipe() {
  if (g_pipe.dev) cudaFree(g_pipe.dev);
  for (int s = 0; s < kPipeSlots; ++s) {
    if (g_pipe.ready[s]) cudaEventDestroy(g_pipe.ready[s]);
    if (g_pipe.done[s]) cudaEventDestroy(g_pipe.done[s]);
    if (g_pipe.freed[s]) cudaEventDestroy(g_pipe.freed[s]);
  }
  if (g_pipe.h2d) cudaStreamDestroy(g_pipe.h2d);
  for (int r = 0; r < kPipeRun; ++r)
    if (g_pipe.run[r]) cudaStreamDestroy(g_pipe.run[r]);
  if (g_pipe.d2h) cudaStreamDestroy(g_pipe.d2h);
  g_pipe = PipeCache();
}
}  // namespace ifd
