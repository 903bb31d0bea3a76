// onet.cu -- ONet-Opt: DecoderCBatchNorm forward + dgrad on the tensor cores, and the restoration loop.
//
// Reference: ONet/im2mesh/onet/models/decoder.py:115-133 (DecoderCBatchNorm.forward), ONet/im2mesh/layers.py:98-107
// (CResnetBlockConv1d), :226-242 (CBatchNorm1d), ONet/opt_defense.py:182-239 (optimize_points).
//
// The model is in eval() (ONet/opt_defense.py:71), so each of the 11 conditional batch norms is a per-cloud,
// per-channel affine  y = s_b[ch] * x + t_b[ch]  that is constant over the 201 iterations: it is folded once per
// batch (onet_fold_cbn_kernel) and applied as the prologue of the GEMM that consumes it.
//
// The ten 256x256 layers per direction are real GEMMs (M = B*K points, N = K = 256): one tcgen05 kernel,
//   Y[M][256] = prologue(A)[M][256] . W^T  (+ bias, + residual)                       forward
//   Y[M][256] = (G[M][256] . W) * s_b * [s_b * Xsaved + t_b > 0]  (+ residual)        dgrad through CBN+ReLU
// The ten layers of a direction run as ONE launch of the shared warp-specialised GEMM engine (tc::gemm_chain_kernel: a row tile
// of layer l + 1 needs only the same row tile of layer l, which the same CTA produced); every layer (tc_gemm.cuh: A producers with the CBN + ReLU prologue, weight
// chunks [256 x 32] hi/lo by TMA bulk copies multicast across a 2-CTA cluster, twelve tcgen05.mma per chunk into one of two
// TMEM accumulator stages, epilogue warps with bias / residual / dgrad mask).  The forward pass keeps three activation tensors
// ([M][256] fp32, warp-transposed layout, see act_off4: net ping, h, net pong) and, for the dgrad, one sign bit per CBN
// pre-activation (10 layers x 32 B per row = 21 MB at B=64, K=1024, instead of the 740 MB of the activations themselves).
#include "common.cuh"
#include "ifd_math.cuh"
#include "tc_gemm.cuh"
#include "umma.cuh"

namespace ifd {

constexpr int kOH = 256;          // hidden size
constexpr int kOC = 512;          // c_dim
constexpr int kOnetCbn = 11;      // block{0..4}.bn_{0,1}, bn
constexpr int kChunk = 32;                              // K columns per pipeline step
constexpr int kChunkImgFloats = 2 * kOH * kChunk;       // hi + lo of a [256 x 32] chunk
constexpr int kLayerImgFloats = (kOH / kChunk) * kChunkImgFloats;   // 131072 floats = 512 KB

// Activation tensors of the chain ([M][256] logically: net_i, h_i, gradients) live in a WARP-TRANSPOSED layout: rows in blocks
// of 32, and inside a block the 64 float4 column groups one after the other, each holding its 32 rows contiguously:
//     float offset of (row m, column k) = (((m / 32) * 64 + k / 4) * 32 + m % 32) * 4 + k % 4.
// The kernels that touch them are thread-per-row (a TMEM lane is a row), and in this layout the 32 lanes of a warp reading
// "their row's float4 number q" touch 512 contiguous bytes -- 4 L1 wavefronts instead of the 32 a row-major tensor costs
// (ncu on the row-major version: LSU wavefronts 71 % of peak, tensor pipe 31 %: the layer was bound by its own addressing).
__host__ __device__ __forceinline__ size_t act_off4(int m, int k4) {          // in float4 units
  return ((size_t)(m >> 5) * 64 + (size_t)k4) * 32 + (size_t)(m & 31);
}
__host__ __device__ inline size_t act_floats(int M) { return (size_t)((M + 31) / 32) * 32 * kOH; }

// ---------------------------------------------------------------------------------------------- packing
// One layer's weight W[n][k] (row-major [256][256]) -> chunked K-major UMMA images, hi then lo per chunk.
// transpose = 1 packs W^T (the dgrad operand).
__global__ void onet_pack_layer_kernel(const float* __restrict__ W, int transpose, float* __restrict__ img) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= kOH * kOH) return;
  const int n = e / kOH, k = e % kOH;                     // element (n, k) of the B operand [N][K]
  const float w = transpose ? W[k * kOH + n] : W[n * kOH + k];
  const int kc = k / kChunk, kl = k % kChunk;
  float* chunk = img + (size_t)kc * kChunkImgFloats;
  const uint32_t off = umma::img_offset(n, kl, kOH) / 4;
  chunk[off] = __uint_as_float(umma::tf32_hi(w));
  chunk[kOH * kChunk + off] = __uint_as_float(umma::tf32_lo(w));
}

// Fold the 11 eval-mode CBNs: s[cbn][b][ch] = gamma / sqrt(var + 1e-5), t = beta - mean * s,
// gamma = Wg c_b + bg, beta = Wb c_b + bb  (layers.py:226-242).  One warp per (cbn, b, ch).
struct CbnParams {
  const float* Wg; const float* bg; const float* Wb; const float* bb; const float* mean; const float* var;
};
struct CbnTable { CbnParams p[kOnetCbn]; };
__global__ void onet_fold_cbn_kernel(const CbnTable tab, const float* __restrict__ c, int B, float* __restrict__ s_out,
                                     float* __restrict__ t_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= kOnetCbn * B * kOH) return;
  const int ch = warp % kOH, b = (warp / kOH) % B, j = warp / (kOH * B);
  const CbnParams p = tab.p[j];
  const float* cb = c + (size_t)b * kOC;
  float g = 0.f, be = 0.f;
  for (int k = lane; k < kOC; k += 32) {
    g = fmaf(p.Wg[(size_t)ch * kOC + k], cb[k], g);
    be = fmaf(p.Wb[(size_t)ch * kOC + k], cb[k], be);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    g += __shfl_xor_sync(0xffffffffu, g, o);
    be += __shfl_xor_sync(0xffffffffu, be, o);
  }
  if (lane == 0) {
    g += p.bg[ch];
    be += p.bb[ch];
    const float inv = 1.0f / sqrtf(p.var[ch] + 1e-5f);
    const float s = g * inv;
    s_out[((size_t)j * B + b) * kOH + ch] = s;
    t_out[((size_t)j * B + b) * kOH + ch] = be - p.mean[ch] * s;
  }
}

// ---------------------------------------------------------------------------------------------- the GEMM
struct GemmArgs {
  const float* A;        // [M][256] input rows
  const float* pro_s;    // [B][256] prologue scale (nullptr: identity prologue, no ReLU)
  const float* pro_t;    // [B][256]
  const float* img;      // packed weight images of this layer / direction
  const float* bias;     // [256] or nullptr
  const float* resid;    // [M][256] or nullptr
  uint32_t* mask_out;    // forward prologue: sign bits of the CBN pre-activations it forms, one word per (row, 32-column chunk):
                         // word of (row m, chunk c) at ((m / 32) * 8 + c) * 32 + m % 32  (nullptr: not recorded)
  const uint32_t* mask_bits;   // dgrad epilogue: those bits of the layer being differentiated (nullptr: plain epilogue)
  const float* mask_s;   // [B][256] the CBN scale that goes with them
  float* out;            // [M][256]
  int M, K;              // rows, points per cloud
};

// The same layer on the shared GEMM engine (tc_gemm.cuh): warp-specialised, A chunks produced into shared memory while the
// previous chunk's MMAs run, weights by TMA bulk copies, epilogue of tile i under the main loop of tile i + 1.  The weight
// images of onet_pack_layer_kernel ARE the engine's format (one N tile of 256, eight chunks, hi | lo per chunk).
struct OnetLayerParams {
  int M, n_chunks, n_tiles_n;
  const float* wimg;
  GemmArgs g;
};
struct OnetLayerPolicy {
  using Params = OnetLayerParams;
  struct Row { int row, b; };
  static __device__ __forceinline__ Row row_begin(const Params& P, int row) {
    const int rc = min(row, P.M - 1);
    return Row{row, rc / P.g.K};
  }
  static __device__ __forceinline__ void load(const Params& P, const Row& r, int kc, float (&x)[32]) {
    if (r.row >= P.M) {
#pragma unroll
      for (int k = 0; k < 32; ++k) x[k] = 0.0f;
      return;
    }
    const float4* p4 = reinterpret_cast<const float4*>(P.g.A) + act_off4(r.row, kc * 8);      // + 32 per float4 column group
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = __ldcg(p4 + q * 32);       // coherent: in a chain launch these rows were written by this kernel
      x[4 * q + 0] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
    if (P.g.pro_s) {                    // CBN (eval mode, folded) + ReLU on read
      const float4* s4 = reinterpret_cast<const float4*>(P.g.pro_s + (size_t)r.b * kOH + kc * kChunk);
      const float4* t4 = reinterpret_cast<const float4*>(P.g.pro_t + (size_t)r.b * kOH + kc * kChunk);
      uint32_t m = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 s = __ldg(s4 + q), t = __ldg(t4 + q);
        const float p0 = fmaf(s.x, x[4 * q + 0], t.x), p1 = fmaf(s.y, x[4 * q + 1], t.y);
        const float p2 = fmaf(s.z, x[4 * q + 2], t.z), p3 = fmaf(s.w, x[4 * q + 3], t.w);
        m |= (p0 > 0.0f ? 1u : 0u) << (4 * q) | (p1 > 0.0f ? 1u : 0u) << (4 * q + 1) | (p2 > 0.0f ? 1u : 0u) << (4 * q + 2) |
             (p3 > 0.0f ? 1u : 0u) << (4 * q + 3);
        x[4 * q + 0] = fmaxf(p0, 0.0f);
        x[4 * q + 1] = fmaxf(p1, 0.0f);
        x[4 * q + 2] = fmaxf(p2, 0.0f);
        x[4 * q + 3] = fmaxf(p3, 0.0f);
      }
      // the dgrad needs nothing of this activation but these signs: 32 bytes per row and layer instead of the 1 KB row itself
      if (P.g.mask_out) P.g.mask_out[((size_t)(r.row >> 5) * 8 + kc) * 32 + (r.row & 31)] = m;
    }
  }
  static __device__ __forceinline__ void store(const Params& P, const Row& r, int col0, const float (&y_in)[32]) {
    if (r.row >= P.M) return;
    const GemmArgs& a = P.g;
    const size_t off = act_off4(r.row, col0 >> 2);
    float y[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) y[k] = y_in[k];
    float4 rv[8];
    if (a.resid) {                      // coherent loads (the dgrad residual buffer is updated in place; in a chain launch an earlier
      const float4* r4 = reinterpret_cast<const float4*>(a.resid) + off;      // layer of this kernel wrote it), issued before any store
#pragma unroll
      for (int q = 0; q < 8; ++q) rv[q] = __ldcg(r4 + q * 32);
    }
    if (a.mask_bits) {                  // dgrad through relu(s * x + t): scale by s where the pre-activation was positive
      const uint32_t m = __ldg(a.mask_bits + ((size_t)(r.row >> 5) * 8 + (col0 >> 5)) * 32 + (r.row & 31));
      const float4* s4 = reinterpret_cast<const float4*>(a.mask_s + (size_t)r.b * kOH + col0);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 s = __ldg(s4 + q);
        y[4 * q + 0] = ((m >> (4 * q + 0)) & 1u) ? y[4 * q + 0] * s.x : 0.0f;
        y[4 * q + 1] = ((m >> (4 * q + 1)) & 1u) ? y[4 * q + 1] * s.y : 0.0f;
        y[4 * q + 2] = ((m >> (4 * q + 2)) & 1u) ? y[4 * q + 2] * s.z : 0.0f;
        y[4 * q + 3] = ((m >> (4 * q + 3)) & 1u) ? y[4 * q + 3] * s.w : 0.0f;
      }
    }
    if (a.bias) {
      const float4* b4 = reinterpret_cast<const float4*>(a.bias + col0);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 bv = __ldg(b4 + q);
        y[4 * q + 0] += bv.x; y[4 * q + 1] += bv.y; y[4 * q + 2] += bv.z; y[4 * q + 3] += bv.w;
      }
    }
    if (a.resid) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        y[4 * q + 0] += rv[q].x; y[4 * q + 1] += rv[q].y; y[4 * q + 2] += rv[q].z; y[4 * q + 3] += rv[q].w;
      }
    }
    float4* o4 = reinterpret_cast<float4*>(a.out) + off;
#pragma unroll
    for (int q = 0; q < 8; ++q) o4[q * 32] = make_float4(y[4 * q + 0], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
  }
};

// ---------------------------------------------------------------------------------------------- thin layers
// All thread-per-row on the warp-transposed activation layout (a warp = one block of 32 rows: every access is 512 contiguous bytes).

// net0[m][n] = Wp[n][:] . p_m + bp[n]        (fc_p: Conv1d(3, 256, 1))
__global__ void onet_fcp_kernel(const float* __restrict__ xyz, const float* __restrict__ Wp, const float* __restrict__ bp,
                                int M, float* __restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float px = xyz[(size_t)m * 3 + 0], py = xyz[(size_t)m * 3 + 1], pz = xyz[(size_t)m * 3 + 2];
  float4* o4 = reinterpret_cast<float4*>(out) + act_off4(m, 0);
#pragma unroll 4
  for (int q = 0; q < kOH / 4; ++q) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = 4 * q + j;
      float t = __ldg(bp + n);
      t = fmaf(__ldg(Wp + n * 3 + 0), px, t);
      t = fmaf(__ldg(Wp + n * 3 + 1), py, t);
      t = fmaf(__ldg(Wp + n * 3 + 2), pz, t);
      v[j] = t;
    }
    o4[q * 32] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// logit = w_out . relu(s_f * net + t_f) + b_out ; then either store the logit (forward seam), or turn it into the gradient
// of the loss w.r.t. net:  g_net = glogit * w_out * [pre > 0] * s_f  with glogit = grad_logits[m] (seam) or
// (sigmoid(logit) - target) * ginv (loop).  The dot product runs over n = 0 .. 255 in order (one thread, one chain).
__global__ void onet_head_kernel(const float* __restrict__ net, const float* __restrict__ s, const float* __restrict__ t,
                                 const float* __restrict__ wout, const float* __restrict__ bout, int M, int K,
                                 float* __restrict__ logits_out, const float* __restrict__ grad_logits, int bce, float target,
                                 float ginv, float* __restrict__ gnet_out, double* __restrict__ stat_part) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp_in_block = threadIdx.x >> 5;
  double s0 = 0.0, s1 = 0.0;
  if (m < M) {
    const int b = m / K;
    const float4* row4 = reinterpret_cast<const float4*>(net) + act_off4(m, 0);
    const float4* sb4 = reinterpret_cast<const float4*>(s + (size_t)b * kOH);
    const float4* tb4 = reinterpret_cast<const float4*>(t + (size_t)b * kOH);
    const float4* w4 = reinterpret_cast<const float4*>(wout);
    float acc = 0.f;
#pragma unroll 4
    for (int q = 0; q < kOH / 4; ++q) {
      const float4 x = __ldg(row4 + q * 32), sv = __ldg(sb4 + q), tv = __ldg(tb4 + q), wv = __ldg(w4 + q);
      acc = fmaf(wv.x, fmaxf(fmaf(sv.x, x.x, tv.x), 0.0f), acc);
      acc = fmaf(wv.y, fmaxf(fmaf(sv.y, x.y, tv.y), 0.0f), acc);
      acc = fmaf(wv.z, fmaxf(fmaf(sv.z, x.z, tv.z), 0.0f), acc);
      acc = fmaf(wv.w, fmaxf(fmaf(sv.w, x.w, tv.w), 0.0f), acc);
    }
    const float logit = acc + bout[0];
    if (logits_out) logits_out[m] = logit;
    if (gnet_out) {
      const float sg = sigmoidf_(logit);
      const float gl = bce ? (sg - target) * ginv : grad_logits[m];
      float4* g4 = reinterpret_cast<float4*>(gnet_out) + act_off4(m, 0);
#pragma unroll 4
      for (int q = 0; q < kOH / 4; ++q) {
        const float4 x = __ldg(row4 + q * 32), sv = __ldg(sb4 + q), tv = __ldg(tb4 + q), wv = __ldg(w4 + q);
        float4 o;
        o.x = fmaf(sv.x, x.x, tv.x) > 0.0f ? gl * wv.x * sv.x : 0.0f;
        o.y = fmaf(sv.y, x.y, tv.y) > 0.0f ? gl * wv.y * sv.y : 0.0f;
        o.z = fmaf(sv.z, x.z, tv.z) > 0.0f ? gl * wv.z * sv.z : 0.0f;
        o.w = fmaf(sv.w, x.w, tv.w) > 0.0f ? gl * wv.w * sv.w : 0.0f;
        g4[q * 32] = o;
      }
      if (stat_part) {
        s0 = (double)bce_with_logits(logit, target);
        s1 = (double)sg;
      }
    }
  }
  if (stat_part) {                       // block-uniform: deterministic block sums
    __shared__ double red[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) {
      red[0][warp_in_block] = s0;
      red[1][warp_in_block] = s1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t0 = 0.0, t1 = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        t0 += red[0][w];
        t1 += red[1][w];
      }
      stat_part[blockIdx.x * 2 + 0] = t0;
      stat_part[blockIdx.x * 2 + 1] = t1;
    }
  }
}

// g_p[m][d] = sum_n Wp[n][d] * g_net0[m][n]
__global__ void onet_fcp_bwd_kernel(const float* __restrict__ gnet, const float* __restrict__ Wp, int M, float* __restrict__ gp) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float4* row4 = reinterpret_cast<const float4*>(gnet) + act_off4(m, 0);
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll 4
  for (int q = 0; q < kOH / 4; ++q) {
    const float4 g = __ldg(row4 + q * 32);
    const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = 4 * q + j;
      g0 = fmaf(__ldg(Wp + n * 3 + 0), gv[j], g0);
      g1 = fmaf(__ldg(Wp + n * 3 + 1), gv[j], g1);
      g2 = fmaf(__ldg(Wp + n * 3 + 2), gv[j], g2);
    }
  }
  gp[(size_t)m * 3 + 0] = g0;
  gp[(size_t)m * 3 + 1] = g1;
  gp[(size_t)m * 3 + 2] = g2;
}

}  // namespace ifd

using namespace ifd;

// ------------------------------------------------------------------------------------------------ host side
namespace {
struct OnetWs {
  float* img;      // [2 dirs][10 layers][kLayerImgFloats]
  float* s;        // [11][B][256]
  float* t;
  float* act;      // [3][M][256] (warp-transposed): net ping, h, net pong -- the dgrad needs only sign bits of them
  uint32_t* mask;  // [10 layers][M / 32][8 chunks][32 rows]: sign bits of the CBN pre-activation each layer's prologue forms
  float* g0;       // [M][256] gradient ping
  float* g1;       // [M][256] gradient pong
  double* stat;    // [M/8 blocks][2]
  OnetLayerParams* chain;   // [3][10]: the layer records of the forward chain with masks, the forward chain without, the dgrad chain
  size_t bytes;
};
OnetWs carve_onet(void* base, int B, int K) {
  OnetWs w;
  size_t off = 0;
  const size_t M = (size_t)B * K;
  auto take = [&](size_t bytes) {
    void* p = base ? (char*)base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  w.img = (float*)take((size_t)2 * 10 * kLayerImgFloats * 4);
  w.s = (float*)take((size_t)kOnetCbn * B * kOH * 4);
  w.t = (float*)take((size_t)kOnetCbn * B * kOH * 4);
  const size_t MH = act_floats((int)M);       // rows padded to whole 32-row blocks of the warp-transposed layout
  w.act = (float*)take((size_t)3 * MH * 4);
  w.mask = (uint32_t*)take((size_t)10 * (MH / kOH) * 8 * 4);
  w.g0 = (float*)take(MH * 4);
  w.g1 = (float*)take(MH * 4);
  w.stat = (double*)take(((M + 255) / 256) * 2 * sizeof(double));
  w.chain = (OnetLayerParams*)take((size_t)3 * 10 * sizeof(OnetLayerParams));
  w.bytes = off;
  return w;
}
}  // namespace

// Packed ONet decoder parameters (float32), in this order (names = reference state_dict keys under `decoder.`):
//   fc_p.weight [256][3], fc_p.bias [256],
//   11 x CBN { conv_gamma.weight [256][512], conv_gamma.bias [256], conv_beta.weight [256][512], conv_beta.bias [256],
//              bn.running_mean [256], bn.running_var [256] }   in the order block0.bn_0, block0.bn_1, ..., block4.bn_1, bn
//   10 x { weight [256][256], bias [256] }                     in the order block0.fc_0, block0.fc_1, ..., block4.fc_1
//   fc_out.weight [256], fc_out.bias [1]
namespace {
constexpr size_t kCbnFloats = 2 * ((size_t)kOH * kOC + kOH) + 2 * kOH;
constexpr size_t kFcFloats = (size_t)kOH * kOH + kOH;
constexpr size_t kOffCbn = (size_t)kOH * 3 + kOH;
constexpr size_t kOffFc = kOffCbn + kOnetCbn * kCbnFloats;
constexpr size_t kOffOut = kOffFc + 10 * kFcFloats;
constexpr size_t kOnetFloats = kOffOut + kOH + 1;
}  // namespace

extern "C" size_t ifd_onet_decoder_nfloats(void) { return kOnetFloats; }
extern "C" size_t ifd_onet_workspace_bytes(int B, int K) {
  if (B <= 0 || K <= 0) return 0;
  return carve_onet(nullptr, B, K).bytes + ifd_convonet_opt_workspace_bytes(B, K);
}

// Once per batch: pack the weight images and fold the CBNs for the given latent codes c [B][512].
extern "C" int ifd_onet_prepare(const float* dec_weights, const float* c, int B, int K, void* workspace, size_t workspace_bytes,
                                ifd_stream_t stream) {
  IFD_REQUIRE(dec_weights && c && workspace && B > 0 && K > 0, "ifd_onet_prepare: bad arguments");
  if (workspace_bytes < ifd_onet_workspace_bytes(B, K)) return fail(IFD_ERR_WORKSPACE, "ifd_onet_prepare: workspace too small");
  cudaStream_t st = as_stream(stream);
  OnetWs w = carve_onet(workspace, B, K);
  for (int l = 0; l < 10; ++l)
    for (int dir = 0; dir < 2; ++dir) {
      onet_pack_layer_kernel<<<(kOH * kOH + 255) / 256, 256, 0, st>>>(dec_weights + kOffFc + l * kFcFloats, dir,
                                                                     w.img + ((size_t)dir * 10 + l) * kLayerImgFloats);
      IFD_LAUNCH_CHECK("onet_pack_layer_kernel");
    }
  CbnTable tab;
  for (int j = 0; j < kOnetCbn; ++j) {
    const float* p = dec_weights + kOffCbn + j * kCbnFloats;
    tab.p[j].Wg = p;
    tab.p[j].bg = p + (size_t)kOH * kOC;
    tab.p[j].Wb = p + (size_t)kOH * kOC + kOH;
    tab.p[j].bb = p + 2 * (size_t)kOH * kOC + kOH;
    tab.p[j].mean = p + 2 * ((size_t)kOH * kOC + kOH);
    tab.p[j].var = tab.p[j].mean + kOH;
  }
  const size_t warps = (size_t)kOnetCbn * B * kOH;
  onet_fold_cbn_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(tab, c, B, w.s, w.t);
  IFD_LAUNCH_CHECK("onet_fold_cbn_kernel");
  return IFD_OK;
}

namespace {
inline float* onet_net5(const OnetWs& w, int M) { return w.act + (size_t)2 * act_floats(M); }
inline uint32_t* onet_mask(const OnetWs& w, int M, int layer) { return w.mask + (size_t)layer * (act_floats(M) / kOH) * 8; }
int g_onet_chain = 1;      // ifd_test_hook(8, 0 / 1): the ten layers of a direction as ONE launch of the GEMM engine (tc::gemm_chain_kernel)

OnetLayerParams as_layer(const GemmArgs& a) {
  OnetLayerParams P{};
  P.M = a.M; P.n_chunks = kOH / kChunk; P.n_tiles_n = 1; P.wimg = a.img; P.g = a;
  return P;
}
// the ten forward layers in order (fc_0, fc_1 of block 0, ...), as onet_forward describes them
void forward_layers(const float* W, const OnetWs& w, int B, int K, bool masks, OnetLayerParams (&out)[10]) {
  const int M = B * K;
  const size_t MH = act_floats(M);
  for (int blk = 0; blk < 5; ++blk) {
    float* net = w.act + (size_t)((blk & 1) ? 2 : 0) * MH;
    float* h = w.act + MH;
    float* net_next = w.act + (size_t)((blk & 1) ? 0 : 2) * MH;
    GemmArgs a{};
    a.M = M; a.K = K;
    a.A = net; a.pro_s = w.s + (size_t)(2 * blk) * B * kOH; a.pro_t = w.t + (size_t)(2 * blk) * B * kOH;
    a.img = w.img + (size_t)(2 * blk) * kLayerImgFloats;
    a.bias = W + kOffFc + (2 * blk) * kFcFloats + (size_t)kOH * kOH;
    a.mask_out = masks ? onet_mask(w, M, 2 * blk) : nullptr;
    a.out = h;
    GemmArgs c{};
    c.M = M; c.K = K;
    c.A = h; c.pro_s = w.s + (size_t)(2 * blk + 1) * B * kOH; c.pro_t = w.t + (size_t)(2 * blk + 1) * B * kOH;
    c.img = w.img + (size_t)(2 * blk + 1) * kLayerImgFloats;
    c.bias = W + kOffFc + (2 * blk + 1) * kFcFloats + (size_t)kOH * kOH;
    c.mask_out = masks ? onet_mask(w, M, 2 * blk + 1) : nullptr;
    c.resid = net;
    c.out = net_next;
    out[2 * blk] = as_layer(a);
    out[2 * blk + 1] = as_layer(c);
  }
}
// the ten dgrad layers in order (fc_1, fc_0 of block 4, ...): g_net5 in w.g0 -> g_net0 in w.g0
void backward_layers(const OnetWs& w, int B, int K, OnetLayerParams (&out)[10]) {
  const int M = B * K;
  float* gnet = w.g0;
  float* gtmp = w.g1;
  int n = 0;
  for (int blk = 4; blk >= 0; --blk) {
    GemmArgs a{};                                   // g_h = (g_net . W1) * s1 * [cbn1(h) > 0]
    a.M = M; a.K = K; a.A = gnet;
    a.img = w.img + ((size_t)10 + 2 * blk + 1) * kLayerImgFloats;
    a.mask_bits = onet_mask(w, M, 2 * blk + 1); a.mask_s = w.s + (size_t)(2 * blk + 1) * B * kOH;
    a.out = gtmp;
    GemmArgs c{};                                   // g_net = g_net + (g_h . W0) * s0 * [cbn0(net) > 0]   (in place on gnet)
    c.M = M; c.K = K; c.A = gtmp;
    c.img = w.img + ((size_t)10 + 2 * blk) * kLayerImgFloats;
    c.mask_bits = onet_mask(w, M, 2 * blk); c.mask_s = w.s + (size_t)(2 * blk) * B * kOH;
    c.resid = gnet;
    c.out = gnet;                                   // each thread reads and writes only its own row chunk
    out[n++] = as_layer(a);
    out[n++] = as_layer(c);
  }
}
// ten layers: one chain launch (the records go to w.chain[slot] first), or ten launches
int run_layers(const OnetLayerParams (&L)[10], const OnetWs& w, int slot, bool upload, cudaStream_t st) {
  if (g_onet_chain) {
    OnetLayerParams* dev = w.chain + (size_t)slot * 10;
    if (upload) IFD_CUDA_TRY(cudaMemcpyAsync(dev, L, sizeof(L), cudaMemcpyHostToDevice, st));
    const int rc = tc::launch_chain<OnetLayerPolicy, 256>(dev, 10, L[0], st);
    if (rc != 1) return rc;                         // 1: the shape does not fit the chain form
  }
  for (int l = 0; l < 10; ++l) {
    const int rc = tc::launch<OnetLayerPolicy>(L[l], 256, st);
    if (rc) return rc;
  }
  return IFD_OK;
}

// forward through the decoder.  net ping-pongs between act slots 0 and 2 (net_5 ends in slot 2, see onet_net5), h lives in slot 1;
// the sign bits of every CBN pre-activation go to w.mask[layer] when `masks` (the dgrad needs nothing else of the activations).
// `upload`: the layer records of this (workspace, weights) pair are not in w.chain yet (the loop uploads them once).
int onet_forward(const float* W, const OnetWs& w, const float* xyz, int B, int K, bool masks, cudaStream_t st, bool upload = true) {
  const int M = B * K;
  onet_fcp_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(xyz, W, W + kOH * 3, M, w.act);
  IFD_LAUNCH_CHECK("onet_fcp_kernel");
  OnetLayerParams L[10];
  forward_layers(W, w, B, K, masks, L);
  return run_layers(L, w, masks ? 0 : 1, upload, st);
}
// dgrad from g_net5 (in w.g0) down to grad_xyz
int onet_backward(const float* W, const OnetWs& w, int B, int K, float* grad_xyz, cudaStream_t st, bool upload = true) {
  const int M = B * K;
  OnetLayerParams L[10];
  backward_layers(w, B, K, L);
  int rc = run_layers(L, w, 2, upload, st);
  if (rc) return rc;
  onet_fcp_bwd_kernel<<<(M + 255) / 256, 256, 0, st>>>(w.g0, W, M, grad_xyz);
  IFD_LAUNCH_CHECK("onet_fcp_bwd_kernel");
  return IFD_OK;
}
}  // namespace

// OccupancyNetwork.decode(p, z, c).logits with z_dim == 0 (ONet/im2mesh/onet/models/__init__.py, decoder.py:115-133).
// Requires ifd_onet_prepare on the same workspace.  logits_out [B][K].
extern "C" int ifd_onet_decode_fwd(const float* dec_weights, const float* xyz, int B, int K, float* logits_out, void* workspace,
                                   size_t workspace_bytes, ifd_stream_t stream) {
  IFD_REQUIRE(dec_weights && xyz && logits_out && workspace && B > 0 && K > 0, "ifd_onet_decode_fwd: bad arguments");
  if (workspace_bytes < ifd_onet_workspace_bytes(B, K)) return fail(IFD_ERR_WORKSPACE, "ifd_onet_decode_fwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  OnetWs w = carve_onet(workspace, B, K);
  int rc = onet_forward(dec_weights, w, xyz, B, K, false, st);
  if (rc) return rc;
  const int M = B * K;
  onet_head_kernel<<<(M + 255) / 256, 256, 0, st>>>(onet_net5(w, M), w.s + (size_t)10 * B * kOH, w.t + (size_t)10 * B * kOH,
                                               dec_weights + kOffOut, dec_weights + kOffOut + kOH, M, K, logits_out, nullptr, 0, 0.f,
                                               0.f, nullptr, nullptr);
  IFD_LAUNCH_CHECK("onet_head_kernel");
  return IFD_OK;
}

// Forward + backward w.r.t. xyz for given grad_logits [B][K] (the autograd seam).
extern "C" int ifd_onet_decode_bwd(const float* dec_weights, const float* xyz, const float* grad_logits, int B, int K,
                                   float* grad_xyz_out, void* workspace, size_t workspace_bytes, ifd_stream_t stream) {
  IFD_REQUIRE(dec_weights && xyz && grad_logits && grad_xyz_out && workspace && B > 0 && K > 0, "ifd_onet_decode_bwd: bad arguments");
  if (workspace_bytes < ifd_onet_workspace_bytes(B, K)) return fail(IFD_ERR_WORKSPACE, "ifd_onet_decode_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  OnetWs w = carve_onet(workspace, B, K);
  int rc = onet_forward(dec_weights, w, xyz, B, K, true, st);
  if (rc) return rc;
  const int M = B * K;
  onet_head_kernel<<<(M + 255) / 256, 256, 0, st>>>(onet_net5(w, M), w.s + (size_t)10 * B * kOH, w.t + (size_t)10 * B * kOH,
                                               dec_weights + kOffOut, dec_weights + kOffOut + kOH, M, K, nullptr, grad_logits, 0, 0.f,
                                               0.f, w.g0, nullptr);
  IFD_LAUNCH_CHECK("onet_head_kernel");
  return onet_backward(dec_weights, w, B, K, grad_xyz_out, st);
}

// Loop pieces shared with the ConvONet path (restore.cu)
namespace ifd {
void onet_set_chain(int on) { g_onet_chain = on ? 1 : 0; }
int opt_step_tail(float* xyz, float* m, float* v, const float* g_occ, int B, int K, const ifd_opt_params* P, int i, void* conv_ws,
                  bool stat, const double* dec_part, int n_dec, double* stats_out, bool warm_ok, cudaStream_t st,
                  const LoopJob* job, bool fresh);
int opt_begin(float* m, float* v, bool zero_state, int B, int K, const ifd_opt_params* P, void* conv_ws, cudaStream_t st);
int opt_finish(float* xyz, int B, int K, int normalize, cudaStream_t st, const LoopJob* job);
float* opt_ws_gocc(void* conv_ws, int B, int K);
float* opt_ws_m(void* conv_ws, int B, int K);
float* opt_ws_v(void* conv_ws, int B, int K);
}  // namespace ifd

// optimize_points for ONet (ONet/opt_defense.py:182-239): same loop, decode(p, z, c) with the CBN decoder.
extern "C" int ifd_onet_opt(const float* dec_weights, const float* c, float* xyz, float* adam_m, float* adam_v, int B, int K,
                            const ifd_opt_params* P, double* stats_out, void* workspace, size_t workspace_bytes,
                            ifd_stream_t stream) {
  IFD_REQUIRE(dec_weights && c && xyz && P && workspace && B > 0 && K > 0, "ifd_onet_opt: bad arguments");
  IFD_REQUIRE(P->n_steps >= 0 && P->step0 >= 0 && P->B_ref > 0, "ifd_onet_opt: bad step counts / B_ref");
  IFD_REQUIRE((adam_m == nullptr) == (adam_v == nullptr), "ifd_onet_opt: pass both adam_m and adam_v or neither");
  IFD_REQUIRE(!(P->step0 > 0 && !adam_m), "ifd_onet_opt: resuming (step0 > 0) needs adam_m / adam_v");
  if (workspace_bytes < ifd_onet_workspace_bytes(B, K)) return fail(IFD_ERR_WORKSPACE, "ifd_onet_opt: workspace too small");
  cudaStream_t st = as_stream(stream);
  OnetWs w = carve_onet(workspace, B, K);
  void* conv_ws = (char*)workspace + w.bytes;
  int rc = ifd_onet_prepare(dec_weights, c, B, K, workspace, workspace_bytes, stream);
  if (rc) return rc;
  float* m = adam_m ? adam_m : opt_ws_m(conv_ws, B, K);
  float* v = adam_v ? adam_v : opt_ws_v(conv_ws, B, K);
  const bool fresh = !adam_m || P->step0 == 0;
  if ((rc = opt_begin(m, v, fresh, B, K, P, conv_ws, st))) return rc;
  float* g_occ = opt_ws_gocc(conv_ws, B, K);
  const int M = B * K;
  const float ginv = (float)K / (float)((long long)P->B_ref * K);
  const int n_dec = (M + 255) / 256;
  for (int i = 0; i < P->n_steps; ++i) {
    const bool stat = P->want_stats && stats_out && (i % 100 == 0);
    {
      ProfileScope ps(0, st);
      if ((rc = onet_forward(dec_weights, w, xyz, B, K, true, st, i == 0))) return rc;
      onet_head_kernel<<<n_dec, 256, 0, st>>>(onet_net5(w, M), w.s + (size_t)10 * B * kOH, w.t + (size_t)10 * B * kOH,
                                              dec_weights + kOffOut, dec_weights + kOffOut + kOH, M, K, nullptr, nullptr, 1,
                                              (float)P->occ_target, ginv, w.g0, stat ? w.stat : nullptr);
      IFD_LAUNCH_CHECK("onet_head_kernel");
      if ((rc = onet_backward(dec_weights, w, B, K, g_occ, st, i == 0))) return rc;
    }
    if ((rc = opt_step_tail(xyz, m, v, g_occ, B, K, P, i, conv_ws, stat, w.stat, n_dec, stats_out, true, st, nullptr, fresh))) return rc;
  }
  return opt_finish(xyz, B, K, P->normalize_out, st, nullptr);
}

