// encoder.cu -- the scatter / gather operators of ConvONet's LocalPoolPointnet (run once per batch, before the loop).
//
// Reference: ConvONet/src/encoder/pointnet.py:68-86 (generate_plane_features: scatter_mean into R^2 bins),
// :104-122 (pool_local: scatter_max into bins, gather back to the points, summed over the planes),
// ConvONet/src/common.py:235-258 (normalize_coordinate), :300-315 (coordinate2index).  The reference calls
// torch_scatter (CUDA atomics: the mean's summation order changes run to run); here every bin's members are walked
// in ascending point order, so the results are bitwise reproducible and equal to a sequential scatter_add.
//
// One CTA per cloud.  Bins are few-membered (T = 600 points into 4096 bins), the op runs 4 + 3 times per batch:
// it is latency-trivial next to the loop, what matters is exactness.
#include "common.cuh"
#include "ifd_math.cuh"

namespace ifd {

constexpr int kEncThreads = 256;

// bins[pl][b][t] = floor(u0 * R) + R * floor(u1 * R) for the three planes xz, xy, yz  (u = normalize_coordinate(p))
__global__ void plane_bins_kernel(const float* __restrict__ p, int n, int R, float denom, int32_t* __restrict__ bins) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float x = p[(size_t)e * 3 + 0], y = p[(size_t)e * 3 + 1], z = p[(size_t)e * 3 + 2];
  const float fR = (float)R;
  // (x * reso).long(): truncation toward zero of the fp32 product; u is in [0, 1) after the clamps
  const int ix = (int)mul_rn(plane_coord(x, denom).u, fR);
  const int iy = (int)mul_rn(plane_coord(y, denom).u, fR);
  const int iz = (int)mul_rn(plane_coord(z, denom).u, fR);
  bins[(size_t)0 * n + e] = ix + R * iz;   // 'xz': (x, z)
  bins[(size_t)1 * n + e] = ix + R * iy;   // 'xy': (x, y)
  bins[(size_t)2 * n + e] = iy + R * iz;   // 'yz': (y, z)
}

// Per-bin member lists in ascending point order: head[bin] = first member (or -1), next[t] = next member (or -1).
// Built by one thread walking t downwards -- T is a few hundred, and it makes the order canonical.
__device__ __forceinline__ int safe_bin(int b, int nbins) { return b < 0 ? 0 : (b >= nbins ? nbins - 1 : b); }

__device__ __forceinline__ void build_lists(const int32_t* __restrict__ bin, int T, int nbins, int* head, int* next) {
  for (int i = threadIdx.x; i < nbins; i += kEncThreads) head[i] = -1;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = T - 1; t >= 0; --t) {
      const int b = safe_bin(bin[t], nbins);
      next[t] = head[b];
      head[b] = t;
    }
  }
  __syncthreads();
}

// out[b][t][c] = sum over planes (in plane order, starting from 0) of max_{t' : bin_pl[t'] == bin_pl[t]} src[b][t'][c]
__global__ void __launch_bounds__(kEncThreads) scatter_max_gather_kernel(const float* __restrict__ src,
                                                                         const int32_t* __restrict__ bins, int P, int B, int T,
                                                                         int C, int nbins, float* __restrict__ out) {
  extern __shared__ int enc_smem[];
  int* head = enc_smem;            // [nbins]
  int* next = enc_smem + nbins;    // [T]
  const int b = blockIdx.x;
  const float* s = src + (size_t)b * T * C;
  float* o = out + (size_t)b * T * C;
  for (int pl = 0; pl < P; ++pl) {
    const int32_t* bin = bins + ((size_t)pl * B + b) * T;
    build_lists(bin, T, nbins, head, next);
    for (int e = threadIdx.x; e < T * C; e += kEncThreads) {
      const int t = e / C, c = e - t * C;
      float m = -INFINITY;
      for (int u = head[safe_bin(bin[t], nbins)]; u >= 0; u = next[u]) m = fmaxf(m, s[(size_t)u * C + c]);
      o[e] = pl == 0 ? add_rn(0.0f, m) : add_rn(o[e], m);          // c_out = 0; c_out += fea   (pointnet.py:108-121)
    }
    __syncthreads();
  }
}

// plane[b][bin][c] = (sum in ascending t of src[b][t][c] over the members of the bin) / max(count, 1); channels-last
__global__ void __launch_bounds__(kEncThreads) scatter_mean_cl_kernel(const float* __restrict__ src,
                                                                      const int32_t* __restrict__ bins, int T, int C, int nbins,
                                                                      float* __restrict__ plane) {
  extern __shared__ int enc_smem[];
  int* head = enc_smem;
  int* next = enc_smem + nbins;
  const int b = blockIdx.x;
  const float* s = src + (size_t)b * T * C;
  const int32_t* bin = bins + (size_t)b * T;
  float* o = plane + (size_t)b * nbins * C;
  build_lists(bin, T, nbins, head, next);
  for (int e = threadIdx.x; e < nbins * C; e += kEncThreads) {
    const int g = e / C, c = e - g * C;
    float sum = 0.0f;
    int cnt = 0;
    for (int u = head[g]; u >= 0; u = next[u]) {
      sum = add_rn(sum, s[(size_t)u * C + c]);
      ++cnt;
    }
    o[e] = div_rn(sum, (float)(cnt > 0 ? cnt : 1));
  }
}

}  // namespace ifd

using namespace ifd;

extern "C" int ifd_plane_bins(const float* xyz, int B, int T, int R, double padding, int32_t* bins_out, ifd_stream_t stream) {
  IFD_REQUIRE(xyz && bins_out && B > 0 && T > 0 && R > 0 && R <= 1024, "ifd_plane_bins: bad arguments");
  const int n = B * T;
  plane_bins_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(xyz, n, R, (float)(1.0 + padding + 10e-6), bins_out);
  IFD_LAUNCH_CHECK("plane_bins_kernel");
  return IFD_OK;
}

static int enc_smem_bytes(int T, int nbins, size_t* out) {
  const size_t bytes = ((size_t)nbins + T) * sizeof(int);
  if (bytes > 200 * 1024) return fail(IFD_ERR_UNSUPPORTED, "encoder scatter: nbins + T must be <= 51200");
  *out = bytes;
  return IFD_OK;
}

extern "C" int ifd_scatter_max_gather(const float* src, const int32_t* bins, int P, int B, int T, int C, int nbins, float* out,
                                      ifd_stream_t stream) {
  IFD_REQUIRE(src && bins && out && src != out && P > 0 && B > 0 && T > 0 && C > 0 && nbins > 0, "ifd_scatter_max_gather: bad arguments");
  size_t smem;
  int rc = enc_smem_bytes(T, nbins, &smem);
  if (rc) return rc;
  IFD_CUDA_TRY(set_max_dyn_smem((const void*)scatter_max_gather_kernel, smem));
  scatter_max_gather_kernel<<<B, kEncThreads, smem, as_stream(stream)>>>(src, bins, P, B, T, C, nbins, out);
  IFD_LAUNCH_CHECK("scatter_max_gather_kernel");
  return IFD_OK;
}

extern "C" int ifd_scatter_mean_cl(const float* src, const int32_t* bins, int B, int T, int C, int nbins, float* plane_out,
                                   ifd_stream_t stream) {
  IFD_REQUIRE(src && bins && plane_out && B > 0 && T > 0 && C > 0 && nbins > 0, "ifd_scatter_mean_cl: bad arguments");
  size_t smem;
  int rc = enc_smem_bytes(T, nbins, &smem);
  if (rc) return rc;
  IFD_CUDA_TRY(set_max_dyn_smem((const void*)scatter_mean_cl_kernel, smem));
  scatter_mean_cl_kernel<<<B, kEncThreads, smem, as_stream(stream)>>>(src, bins, T, C, nbins, plane_out);
  IFD_LAUNCH_CHECK("scatter_mean_cl_kernel");
  return IFD_OK;
}
