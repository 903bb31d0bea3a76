// geometry.cu -- index-producing geometry kernels: kNN (3-D and feature space), FPS, ball query, SOR.
//
// All of them are FP32/FP64-ALU + shared-memory bound and HBM-trivial (12 B/point in); the candidate set
// of one cloud is staged in shared memory once per CTA and scanned by every thread as a warp-broadcast.
// Index results must equal the reference bit for bit, so distances are evaluated in the reference's own
// association (ifd_math.cuh) and ties resolve to the lowest index (topk.cuh).
#include "common.cuh"
#include "ifd_math.cuh"
#include "topk.cuh"

namespace ifd {

// ------------------------------------------------------------------------------------------------
// kNN, C == 3  (ConvONet/defense/pn_utils.py:64-83; baselines/model/dgcnn.py:7-13 first layer)
// ------------------------------------------------------------------------------------------------
constexpr int kKnnThreads = 128;
constexpr int kKnnTile = 2048;  // candidates per shared-memory tile (32 KB of float4)

template <int KK>
__global__ void __launch_bounds__(kKnnThreads) knn3_kernel(const float* __restrict__ xyz, int N, int k, int drop,
                                                           int32_t* __restrict__ idx_out,
                                                           float* __restrict__ key_out) {
  __shared__ float4 cand[kKnnTile];
  const int b = blockIdx.y;
  const int q = blockIdx.x * kKnnThreads + threadIdx.x;
  const float* cloud = xyz + (size_t)b * N * 3;
  float qx = 0.f, qy = 0.f, qz = 0.f, qxx = 0.f;
  if (q < N) {
    qx = cloud[q * 3 + 0];
    qy = cloud[q * 3 + 1];
    qz = cloud[q * 3 + 2];
    qxx = sqnorm3(qx, qy, qz);
  }
  TopK<KK> top;
  top.init(INFINITY);
  for (int t0 = 0; t0 < N; t0 += kKnnTile) {
    const int tn = min(kKnnTile, N - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < tn; j += kKnnThreads) {
      const float x = cloud[(t0 + j) * 3 + 0], y = cloud[(t0 + j) * 3 + 1], z = cloud[(t0 + j) * 3 + 2];
      cand[j] = make_float4(x, y, z, sqnorm3(x, y, z));
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < tn; ++j) {
      const float4 c = cand[j];
      const float d = knn_key(qxx, c.w, dot3_chain(qx, qy, qz, c.x, c.y, c.z));
      top.offer(d, t0 + j);
    }
  }
  if (q < N) {
    int32_t* o = idx_out + ((size_t)b * N + q) * k;
#pragma unroll
    for (int s = 0; s < KK; ++s)
      if (s >= drop && s - drop < k) {
        o[s - drop] = top.id[s];
        if (key_out) key_out[((size_t)b * N + q) * k + (s - drop)] = top.key[s];
      }
  }
}

// ------------------------------------------------------------------------------------------------
// kNN in feature space, C % 4 == 0  (baselines/model/dgcnn.py:7-13 layers 2-4: C = 64, 64, 128)
// Query features live in shared memory as [C/4][threads] float4 (conflict-free), candidates as
// [tile][C] (warp-broadcast).  dot = FMA chain over c, |x|^2 = left-to-right sum of rounded squares.
// ------------------------------------------------------------------------------------------------
constexpr int kKnnCTile = 32;

template <int KK>
__global__ void __launch_bounds__(kKnnThreads) knnc_kernel(const float* __restrict__ x, int N, int C, int k, int drop,
                                                           int32_t* __restrict__ idx_out,
                                                           float* __restrict__ key_out) {
  extern __shared__ float4 smem4[];
  const int C4 = C / 4;
  float4* qf = smem4;                                 // [C4][kKnnThreads]
  float4* cf = smem4 + (size_t)C4 * kKnnThreads;      // [kKnnCTile][C4]
  float* cxx = reinterpret_cast<float*>(cf + (size_t)kKnnCTile * C4);  // [kKnnCTile]
  const int b = blockIdx.y;
  const int q = blockIdx.x * kKnnThreads + threadIdx.x;
  const float* cloud = x + (size_t)b * N * C;
  float qxx = 0.f;
  {
    const int qq = min(q, N - 1);
    const float4* src = reinterpret_cast<const float4*>(cloud + (size_t)qq * C);
    for (int c4 = 0; c4 < C4; ++c4) {
      const float4 v = src[c4];
      qf[c4 * kKnnThreads + threadIdx.x] = v;
      if (c4 == 0) qxx = mul_rn(v.x, v.x); else qxx = add_rn(qxx, mul_rn(v.x, v.x));
      qxx = add_rn(qxx, mul_rn(v.y, v.y));
      qxx = add_rn(qxx, mul_rn(v.z, v.z));
      qxx = add_rn(qxx, mul_rn(v.w, v.w));
    }
  }
  TopK<KK> top;
  top.init(INFINITY);
  for (int t0 = 0; t0 < N; t0 += kKnnCTile) {
    const int tn = min(kKnnCTile, N - t0);
    __syncthreads();
    for (int e = threadIdx.x; e < tn * C4; e += kKnnThreads)
      cf[e] = reinterpret_cast<const float4*>(cloud + (size_t)t0 * C)[e];
    __syncthreads();
    if (threadIdx.x < tn) {
      float s = 0.f;
      for (int c4 = 0; c4 < C4; ++c4) {
        const float4 v = cf[threadIdx.x * C4 + c4];
        if (c4 == 0) s = mul_rn(v.x, v.x); else s = add_rn(s, mul_rn(v.x, v.x));
        s = add_rn(s, mul_rn(v.y, v.y));
        s = add_rn(s, mul_rn(v.z, v.z));
        s = add_rn(s, mul_rn(v.w, v.w));
      }
      cxx[threadIdx.x] = s;
    }
    __syncthreads();
    for (int j0 = 0; j0 < tn; j0 += 4) {
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
      const int j1 = min(j0 + 1, tn - 1), j2 = min(j0 + 2, tn - 1), j3 = min(j0 + 3, tn - 1);
      for (int c4 = 0; c4 < C4; ++c4) {
        const float4 qv = qf[c4 * kKnnThreads + threadIdx.x];
        const float4 a = cf[j0 * C4 + c4], bb = cf[j1 * C4 + c4], cc = cf[j2 * C4 + c4], dd = cf[j3 * C4 + c4];
        if (c4 == 0) {
          d0 = mul_rn(qv.x, a.x); d1 = mul_rn(qv.x, bb.x); d2 = mul_rn(qv.x, cc.x); d3 = mul_rn(qv.x, dd.x);
        } else {
          d0 = fma_rn(qv.x, a.x, d0); d1 = fma_rn(qv.x, bb.x, d1); d2 = fma_rn(qv.x, cc.x, d2); d3 = fma_rn(qv.x, dd.x, d3);
        }
        d0 = fma_rn(qv.y, a.y, d0); d1 = fma_rn(qv.y, bb.y, d1); d2 = fma_rn(qv.y, cc.y, d2); d3 = fma_rn(qv.y, dd.y, d3);
        d0 = fma_rn(qv.z, a.z, d0); d1 = fma_rn(qv.z, bb.z, d1); d2 = fma_rn(qv.z, cc.z, d2); d3 = fma_rn(qv.z, dd.z, d3);
        d0 = fma_rn(qv.w, a.w, d0); d1 = fma_rn(qv.w, bb.w, d1); d2 = fma_rn(qv.w, cc.w, d2); d3 = fma_rn(qv.w, dd.w, d3);
      }
      top.offer(knn_key(qxx, cxx[j0], d0), t0 + j0);
      if (j0 + 1 < tn) top.offer(knn_key(qxx, cxx[j0 + 1], d1), t0 + j0 + 1);
      if (j0 + 2 < tn) top.offer(knn_key(qxx, cxx[j0 + 2], d2), t0 + j0 + 2);
      if (j0 + 3 < tn) top.offer(knn_key(qxx, cxx[j0 + 3], d3), t0 + j0 + 3);
    }
  }
  if (q < N) {
    int32_t* o = idx_out + ((size_t)b * N + q) * k;
#pragma unroll
    for (int s = 0; s < KK; ++s)
      if (s >= drop && s - drop < k) {
        o[s - drop] = top.id[s];
        if (key_out) key_out[((size_t)b * N + q) * k + (s - drop)] = top.key[s];
      }
  }
}

template <int KK>
static int launch_knn(const float* x, int B, int N, int C, int k, int drop, int32_t* idx_out, float* key_out,
                      cudaStream_t st) {
  dim3 grid((N + kKnnThreads - 1) / kKnnThreads, B);
  if (C == 3) {
    knn3_kernel<KK><<<grid, kKnnThreads, 0, st>>>(x, N, k, drop, idx_out, key_out);
    IFD_LAUNCH_CHECK("knn3_kernel");
  } else {
    const size_t smem = ((size_t)(C / 4) * kKnnThreads + (size_t)kKnnCTile * (C / 4)) * sizeof(float4) +
                        kKnnCTile * sizeof(float);
    IFD_CUDA_TRY(set_max_dyn_smem((const void*)knnc_kernel<KK>, smem));
    knnc_kernel<KK><<<grid, kKnnThreads, smem, st>>>(x, N, C, k, drop, idx_out, key_out);
    IFD_LAUNCH_CHECK("knnc_kernel");
  }
  return IFD_OK;
}

// ------------------------------------------------------------------------------------------------
// FPS  (baselines/model/pointnet2.py:53-74; ConvONet/defense/pn_utils.py:26-48)
// One CTA per cloud; each thread keeps its points and running min-distance in registers.
// ------------------------------------------------------------------------------------------------
constexpr int kFpsThreads = 512;
constexpr int kFpsPerThread = 16;  // N <= 8192

__global__ void __launch_bounds__(kFpsThreads) fps_kernel(const float* __restrict__ xyz, int N, int npoint,
                                                          const int32_t* __restrict__ start_idx,
                                                          int32_t* __restrict__ idx_out) {
  __shared__ float s_val[kFpsThreads / 32];
  __shared__ int s_idx[kFpsThreads / 32];
  __shared__ int s_far;
  const int b = blockIdx.x;
  const float* cloud = xyz + (size_t)b * N * 3;
  float px[kFpsPerThread], py[kFpsPerThread], pz[kFpsPerThread], dist[kFpsPerThread];
#pragma unroll
  for (int e = 0; e < kFpsPerThread; ++e) {
    const int i = threadIdx.x + e * kFpsThreads;  // strided ownership: index grows with e for a fixed thread
    if (i < N) {
      px[e] = cloud[i * 3 + 0];
      py[e] = cloud[i * 3 + 1];
      pz[e] = cloud[i * 3 + 2];
      dist[e] = 1e10f;
    } else {
      px[e] = py[e] = pz[e] = 0.f;
      dist[e] = -1.f;  // never selected
    }
  }
  int far = start_idx[b];
  for (int it = 0; it < npoint; ++it) {
    if (threadIdx.x == 0) idx_out[(size_t)b * npoint + it] = far;
    const float cx = cloud[far * 3 + 0], cy = cloud[far * 3 + 1], cz = cloud[far * 3 + 2];
    float best = -2.f;
    int besti = 0x7fffffff;
#pragma unroll
    for (int e = 0; e < kFpsPerThread; ++e) {
      const int i = threadIdx.x + e * kFpsThreads;
      const float dx = sub_rn(px[e], cx), dy = sub_rn(py[e], cy), dz = sub_rn(pz[e], cz);
      const float d = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));
      if (i < N && d < dist[e]) dist[e] = d;
      if (dist[e] > best) {  // strict: the lowest index among equal maxima wins (e ascending = index ascending)
        best = dist[e];
        besti = i;
      }
    }
    // argmax with first-occurrence semantics (torch.max(distance, -1)[1] on CPU)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) {
        best = ov;
        besti = oi;
      }
    }
    if ((threadIdx.x & 31) == 0) {
      s_val[threadIdx.x >> 5] = best;
      s_idx[threadIdx.x >> 5] = besti;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = threadIdx.x < kFpsThreads / 32 ? s_val[threadIdx.x] : -3.f;
      int vi = threadIdx.x < kFpsThreads / 32 ? s_idx[threadIdx.x] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
        if (ov > v || (ov == v && oi < vi)) {
          v = ov;
          vi = oi;
        }
      }
      if (threadIdx.x == 0) s_far = vi;
    }
    __syncthreads();
    far = s_far;
  }
}

// ------------------------------------------------------------------------------------------------
// Ball query  (baselines/model/pointnet2.py:77-98 with square_distance :9-30)
// d2(s, n) = ((-2 * dot(new_xyz_s, xyz_n)) + |new_xyz_s|^2) + |xyz_n|^2 ; keep d2 <= r2 in ascending n.
// ------------------------------------------------------------------------------------------------
constexpr int kBallThreads = 128;
constexpr int kBallTile = 2048;

__global__ void __launch_bounds__(kBallThreads) ball_query_kernel(const float* __restrict__ xyz,
                                                                  const float* __restrict__ new_xyz, int N, int S,
                                                                  float r2, int nsample,
                                                                  int32_t* __restrict__ idx_out) {
  __shared__ float4 cand[kBallTile];
  const int b = blockIdx.y;
  const int s = blockIdx.x * kBallThreads + threadIdx.x;
  const float* cloud = xyz + (size_t)b * N * 3;
  float qx = 0.f, qy = 0.f, qz = 0.f, qxx = 0.f;
  if (s < S) {
    const float* c = new_xyz + ((size_t)b * S + s) * 3;
    qx = c[0]; qy = c[1]; qz = c[2];
    qxx = sqnorm3(qx, qy, qz);
  }
  int32_t* o = idx_out + ((size_t)b * S + (s < S ? s : 0)) * nsample;
  int cnt = 0, first = N;
  for (int t0 = 0; t0 < N; t0 += kBallTile) {
    const int tn = min(kBallTile, N - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < tn; j += kBallThreads) {
      const float x = cloud[(t0 + j) * 3 + 0], y = cloud[(t0 + j) * 3 + 1], z = cloud[(t0 + j) * 3 + 2];
      cand[j] = make_float4(x, y, z, sqnorm3(x, y, z));
    }
    __syncthreads();
    if (s < S) {
      for (int j = 0; j < tn && cnt < nsample; ++j) {
        const float4 c = cand[j];
        const float d = add_rn(add_rn(mul_rn(-2.0f, dot3_chain(qx, qy, qz, c.x, c.y, c.z)), qxx), c.w);
        if (!(d > r2)) {
          if (cnt == 0) first = t0 + j;
          o[cnt++] = t0 + j;
        }
      }
    }
  }
  if (s < S)
    for (; cnt < nsample; ++cnt) o[cnt] = first;
}

// ------------------------------------------------------------------------------------------------
// SOR  (ConvONet/defense/SOR.py:22-49) -- float64 throughout, one CTA per cloud.
// ------------------------------------------------------------------------------------------------
constexpr int kSorThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kSorThreads / 32; ++w) t += red[w];
  return t;
}

template <int KK>
__global__ void __launch_bounds__(kSorThreads) sor_kernel(const float* __restrict__ xyz, int K, int k, double alpha,
                                                          uint8_t* __restrict__ keep_out,
                                                          double* __restrict__ value_out) {
  extern __shared__ double sm[];  // [K][4] x y z xx, then [K] value
  __shared__ double red[kSorThreads / 32];
  double* cand = sm;
  double* val = sm + (size_t)K * 4;
  const int b = blockIdx.x;
  const float* cloud = xyz + (size_t)b * K * 3;
  for (int j = threadIdx.x; j < K; j += kSorThreads) {
    const double x = cloud[j * 3 + 0], y = cloud[j * 3 + 1], z = cloud[j * 3 + 2];
    cand[j * 4 + 0] = x;
    cand[j * 4 + 1] = y;
    cand[j * 4 + 2] = z;
    cand[j * 4 + 3] = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
  }
  __syncthreads();
  double lsum = 0.0;
  for (int q = threadIdx.x; q < K; q += kSorThreads) {
    const double qx = cand[q * 4 + 0], qy = cand[q * 4 + 1], qz = cand[q * 4 + 2], qxx = cand[q * 4 + 3];
    TopK<KK, double> top;
    top.init((double)INFINITY);
    for (int j = 0; j < K; ++j) {
      const double dot = __fma_rn(qz, cand[j * 4 + 2], __fma_rn(qy, cand[j * 4 + 1], __dmul_rn(qx, cand[j * 4 + 0])));
      const double d = __dadd_rn(__dadd_rn(cand[j * 4 + 3], __dmul_rn(-2.0, dot)), qxx);
      top.offer(d, j);
    }
    double s = 0.0;
#pragma unroll
    for (int t = 1; t < KK; ++t)
      if (t <= k) s += top.key[t];
    const double v = s / (double)k;
    val[q] = v;
    lsum += v;
  }
  const double mean = block_sum(lsum, red) / (double)K;
  double lsq = 0.0;
  for (int q = threadIdx.x; q < K; q += kSorThreads) {
    const double dv = val[q] - mean;
    lsq += dv * dv;
  }
  const double var = block_sum(lsq, red) / (double)(K - 1);
  const double thr = mean + alpha * sqrt(var);
  for (int q = threadIdx.x; q < K; q += kSorThreads) {
    keep_out[(size_t)b * K + q] = val[q] <= thr ? 1 : 0;
    if (value_out) value_out[(size_t)b * K + q] = val[q];
  }
}

}  // namespace ifd

using namespace ifd;

extern "C" int ifd_knn(const float* x, int B, int N, int C, int k, int drop_first, int32_t* idx_out, float* key_out,
                       ifd_stream_t stream) {
  IFD_REQUIRE(x && idx_out, "ifd_knn: null pointer");
  IFD_REQUIRE(B > 0 && N > 0 && k > 0 && drop_first >= 0, "ifd_knn: bad sizes");
  IFD_REQUIRE(k + drop_first <= N, "ifd_knn: k + drop_first exceeds N");
  if (!(C == 3 || (C % 4 == 0 && C >= 4 && C <= 256)))
    return fail(IFD_ERR_UNSUPPORTED, "ifd_knn: C must be 3 or a multiple of 4 up to 256");
  const int kk = k + drop_first;
  cudaStream_t st = as_stream(stream);
  if (kk <= 4) return launch_knn<4>(x, B, N, C, k, drop_first, idx_out, key_out, st);
  if (kk <= 8) return launch_knn<8>(x, B, N, C, k, drop_first, idx_out, key_out, st);
  if (kk <= 16) return launch_knn<16>(x, B, N, C, k, drop_first, idx_out, key_out, st);
  if (kk <= 24) return launch_knn<24>(x, B, N, C, k, drop_first, idx_out, key_out, st);
  if (kk <= 32) return launch_knn<32>(x, B, N, C, k, drop_first, idx_out, key_out, st);
  return fail(IFD_ERR_UNSUPPORTED, "ifd_knn: k + drop_first must be <= 32");
}

extern "C" int ifd_fps(const float* xyz, int B, int N, int npoint, const int32_t* start_idx, int32_t* idx_out,
                       ifd_stream_t stream) {
  IFD_REQUIRE(xyz && start_idx && idx_out, "ifd_fps: null pointer");
  IFD_REQUIRE(B > 0 && N > 0 && npoint > 0, "ifd_fps: bad sizes");
  if (N > kFpsThreads * kFpsPerThread) return fail(IFD_ERR_UNSUPPORTED, "ifd_fps: N must be <= 8192");
  fps_kernel<<<B, kFpsThreads, 0, as_stream(stream)>>>(xyz, N, npoint, start_idx, idx_out);
  IFD_LAUNCH_CHECK("fps_kernel");
  return IFD_OK;
}

extern "C" int ifd_ball_query(const float* xyz, const float* new_xyz, int B, int N, int S, float radius_sq,
                              int nsample, int32_t* idx_out, ifd_stream_t stream) {
  IFD_REQUIRE(xyz && new_xyz && idx_out, "ifd_ball_query: null pointer");
  IFD_REQUIRE(B > 0 && N > 0 && S > 0 && nsample > 0, "ifd_ball_query: bad sizes");
  dim3 grid((S + kBallThreads - 1) / kBallThreads, B);
  ball_query_kernel<<<grid, kBallThreads, 0, as_stream(stream)>>>(xyz, new_xyz, N, S, radius_sq, nsample, idx_out);
  IFD_LAUNCH_CHECK("ball_query_kernel");
  return IFD_OK;
}

extern "C" int ifd_sor(const float* xyz, int B, int K, int k, double alpha, uint8_t* keep_out, double* value_out,
                       ifd_stream_t stream) {
  IFD_REQUIRE(xyz && keep_out, "ifd_sor: null pointer");
  IFD_REQUIRE(B > 0 && K > 1 && k > 0 && k + 1 <= K, "ifd_sor: bad sizes");
  if (k > 7) return fail(IFD_ERR_UNSUPPORTED, "ifd_sor: k must be <= 7");
  const size_t smem = (size_t)K * 5 * sizeof(double);
  if (smem > 200 * 1024) return fail(IFD_ERR_UNSUPPORTED, "ifd_sor: K must be <= 5120");
  cudaStream_t st = as_stream(stream);
  if (k <= 3) {
    IFD_CUDA_TRY(set_max_dyn_smem((const void*)sor_kernel<4>, smem));
    sor_kernel<4><<<B, kSorThreads, smem, st>>>(xyz, K, k, alpha, keep_out, value_out);
  } else {
    IFD_CUDA_TRY(set_max_dyn_smem((const void*)sor_kernel<8>, smem));
    sor_kernel<8><<<B, kSorThreads, smem, st>>>(xyz, K, k, alpha, keep_out, value_out);
  }
  IFD_LAUNCH_CHECK("sor_kernel");
  return IFD_OK;
}

// ------------------------------------------------------------------------------------------------
// preprocess_pc for a batch (ConvONet/opt_defense.py:114-130, numpy part), fused with the ragged selection that
// follows SOR (:86-111): one CTA per cloud.  The kept points are compacted in input order, centred on their float32
// mean (summed in point order, as np.mean(axis=0) does on a [K,3] array), divided by the largest bounding-box extent and
// multiplied by padding_scale -- every operation rounded to float32 as numpy rounds it.
// ------------------------------------------------------------------------------------------------
namespace ifd {
constexpr int kPreThreads = 256;

__global__ void __launch_bounds__(kPreThreads) preprocess_pc_kernel(const float* __restrict__ xyz, const uint8_t* __restrict__ keep,
                                                                    int K, float padding_scale, float* __restrict__ out,
                                                                    int32_t* __restrict__ counts) {
  __shared__ int warp_cnt[kPreThreads / 32];
  __shared__ int base_s;
  __shared__ float mean_s[3], ext_lo[3][kPreThreads / 32], ext_hi[3][kPreThreads / 32], scale_s;
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  const float* src = xyz + (size_t)b * K * 3;
  float* dst = out + (size_t)b * K * 3;
  if (t == 0) base_s = 0;
  __syncthreads();
  for (int c0 = 0; c0 < K; c0 += kPreThreads) {                      // stable compaction of the kept points
    const int i = c0 + t;
    const bool on = i < K && (keep == nullptr || keep[(size_t)b * K + i] != 0);
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    if (lane == 0) warp_cnt[w] = __popc(bal);
    __syncthreads();
    int before = base_s;
    for (int q = 0; q < w; ++q) before += warp_cnt[q];
    if (on) {
      const int at = before + __popc(bal & ((1u << lane) - 1));
      dst[at * 3 + 0] = src[i * 3 + 0];
      dst[at * 3 + 1] = src[i * 3 + 1];
      dst[at * 3 + 2] = src[i * 3 + 2];
    }
    __syncthreads();
    if (t == 0) {
      int tot = 0;
      for (int q = 0; q < kPreThreads / 32; ++q) tot += warp_cnt[q];
      base_s += tot;
    }
    __syncthreads();
  }
  const int n = base_s;
  if (t < 3) {                                                       // np.mean(pc, axis=0): sequential float32 sum / n
    float s = 0.0f;
    for (int i = 0; i < n; ++i) s = __fadd_rn(s, dst[i * 3 + t]);
    mean_s[t] = __fdiv_rn(s, (float)n);
  }
  __syncthreads();
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = t; i < n; i += kPreThreads)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float c = __fsub_rn(dst[i * 3 + d], mean_s[d]);
      dst[i * 3 + d] = c;
      lo[d] = fminf(lo[d], c);
      hi[d] = fmaxf(hi[d], c);
    }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    if (lane == 0) {
      ext_lo[d][w] = lo[d];
      ext_hi[d][w] = hi[d];
    }
  }
  __syncthreads();
  if (t == 0) {
    float scale = -INFINITY;
    for (int d = 0; d < 3; ++d) {
      float l = INFINITY, h = -INFINITY;
      for (int q = 0; q < kPreThreads / 32; ++q) {
        l = fminf(l, ext_lo[d][q]);
        h = fmaxf(h, ext_hi[d][q]);
      }
      scale = fmaxf(scale, __fsub_rn(h, l));                         // (max_dim - min_dim).max()
    }
    scale_s = scale;
    counts[b] = n;
  }
  __syncthreads();
  const float scale = scale_s;
  for (int e = t; e < K * 3; e += kPreThreads)
    dst[e] = e < n * 3 ? __fmul_rn(__fdiv_rn(dst[e], scale), padding_scale) : 0.0f;   // centered / scale * padding_scale
}
}  // namespace ifd

extern "C" int ifd_preprocess_pc(const float* xyz, const uint8_t* keep, int B, int K, float padding_scale, float* out,
                                 int32_t* counts, ifd_stream_t stream) {
  IFD_REQUIRE(xyz && out && counts && xyz != out && B > 0 && K > 0, "ifd_preprocess_pc: bad arguments");
  preprocess_pc_kernel<<<B, kPreThreads, 0, as_stream(stream)>>>(xyz, keep, K, padding_scale, out, counts);
  IFD_LAUNCH_CHECK("preprocess_pc_kernel");
  return IFD_OK;
}
