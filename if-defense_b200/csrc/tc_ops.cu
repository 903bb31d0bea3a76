// tc_ops.cu -- the dense layers of the two encoders on the tcgen05 GEMM engine (tc_gemm.cuh), fp32 semantics (3xTF32):
//   * Linear layers with fused pre-activation ReLU, bias, residual and K-concatenated inputs: ResnetBlockFC
//     (ConvONet/src/layers.py:39-48, ONet/im2mesh/layers.py:8-64) = two launches, the shortcut rides in the second one;
//     LocalPoolPointnet.fc_pos / fc_c (ConvONet/src/encoder/pointnet.py:41-45), ResnetPointnet (ONet/im2mesh/encoder/pointnet.py:85-113);
//   * the U-Net (ConvONet/src/encoder/unet.py:117-239) on channels-last tensors: conv3x3 (padding 1) as an implicit GEMM whose
//     A rows are gathered tap by tap -- optionally through the 2x2 max-pool of the level above and / or from the two halves of
//     the skip concatenation, so neither the pooled nor the concatenated tensor is ever written -- ConvTranspose2d(2, stride 2)
//     as a GEMM with a pixel-shuffle epilogue, the final 1x1 convolution as a Linear layer.
// Weight images are packed once per model (ifd_tc_pack).
#include "common.cuh"
#include "tc_gemm.cuh"

namespace ifd {
namespace tc {

__device__ __forceinline__ void load32(const float* __restrict__ p, float (&x)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p) + q);
    x[4 * q + 0] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
  }
}
__device__ __forceinline__ void zero32(float (&x)[32]) {
#pragma unroll
  for (int k = 0; k < 32; ++k) x[k] = 0.0f;
}

// ------------------------------------------------------------------------------------------------ Linear
struct LinearParams {
  int M, n_chunks, n_tiles_n;
  const float* wimg;
  const float* a_ptr[3];     // up to three K segments, concatenated in this order
  int a_ld[3], a_w[3], a_relu[3], a_chunk0[3];      // row stride, width, ReLU on read, first chunk of the segment
  int a_blocked[3];          // 1: the segment is stored in the warp-transposed layout (tc_gemm.cuh, blk_off4), width % 4 == 0
  int out_blocked;           // 1: `out` (N % 4 == 0; with the pixel-shuffle epilogue: [B * 2H * 2W][cout]) is warp-transposed
  int a_group[3];            // > 0: the segment has one row per GROUP of a_group consecutive rows (a per-cloud vector that the
                             // reference expands over the cloud's points: ONet/im2mesh/encoder/pointnet.py:103-104)
  int n_seg;
  const float* bias;         // [N] or null
  const float* resid;        // [M][ld_resid] or null
  int ld_resid, relu_out;
  float* out;                // [M][ld_out]
  int ld_out, N;
  int shuffle_cout, shuffle_H, shuffle_W;   // > 0: ConvTranspose2d(2, 2) epilogue, out is [B][2H][2W][cout] and N = 4 cout
};
struct LinearPolicy {
  using Params = LinearParams;
  struct Row { int row; };
  static __device__ __forceinline__ Row row_begin(const Params&, int row) { return Row{row}; }
  static __device__ __forceinline__ void load(const Params& P, const Row& r, int kc, float (&x)[32]) {
    if (r.row >= P.M) return zero32(x);
    int seg = 0;
    if (P.n_seg > 1 && kc >= P.a_chunk0[1]) seg = 1;
    if (P.n_seg > 2 && kc >= P.a_chunk0[2]) seg = 2;
    const int k0 = (kc - P.a_chunk0[seg]) * kChunk;
    if (P.a_blocked[seg]) {
      const int C4 = P.a_w[seg] >> 2;
      const float4* p4 = reinterpret_cast<const float4*>(P.a_ptr[seg]) + blk_off4((size_t)r.row, k0 >> 2, C4);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = (k0 >> 2) + q < C4 ? __ldg(p4 + q * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
        x[4 * q + 0] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
      }
      if (P.a_relu[seg]) {
#pragma unroll
        for (int k = 0; k < 32; ++k) x[k] = fmaxf(x[k], 0.0f);
      }
      return;
    }
    const int arow = P.a_group[seg] > 0 ? r.row / P.a_group[seg] : r.row;
    const float* p = P.a_ptr[seg] + (size_t)arow * P.a_ld[seg] + k0;
    const int w = P.a_w[seg] - k0;
    if (w >= 32 && (P.a_ld[seg] & 3) == 0) {
      load32(p, x);
    } else {
#pragma unroll
      for (int k = 0; k < 32; ++k) x[k] = k < w ? __ldg(p + k) : 0.0f;
    }
    if (P.a_relu[seg]) {
#pragma unroll
      for (int k = 0; k < 32; ++k) x[k] = fmaxf(x[k], 0.0f);
    }
  }
  static __device__ __forceinline__ void store(const Params& P, const Row& r, int col0, const float (&y_in)[32]) {
    if (r.row >= P.M || col0 >= P.N) return;
    float y[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) y[k] = y_in[k];
    if (P.shuffle_cout > 0) {                // n = (i * 2 + j) * cout + co  ->  out[b][2 y + i][2 x + j][co] (+ bias[co])
      const int q = col0 / P.shuffle_cout, co = col0 % P.shuffle_cout;
      const int hw = P.shuffle_H * P.shuffle_W;
      const int b = r.row / hw, yy = (r.row % hw) / P.shuffle_W, xx = r.row % P.shuffle_W;
      const size_t opix = (((size_t)b * 2 * P.shuffle_H + 2 * yy + (q >> 1)) * 2 * P.shuffle_W) + 2 * xx + (q & 1);
      float4* o4 = P.out_blocked ? reinterpret_cast<float4*>(P.out) + blk_off4(opix, co >> 2, P.shuffle_cout >> 2)
                                 : reinterpret_cast<float4*>(P.out + opix * P.shuffle_cout + co);
      const int ostep = P.out_blocked ? 32 : 1;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 bv = P.bias ? __ldg(reinterpret_cast<const float4*>(P.bias + co) + g) : make_float4(0.f, 0.f, 0.f, 0.f);
        o4[g * ostep] = make_float4(y[4 * g] + bv.x, y[4 * g + 1] + bv.y, y[4 * g + 2] + bv.z, y[4 * g + 3] + bv.w);
      }
      return;
    }
    const int n_here = min(32, P.N - col0);
    if (P.bias) {
#pragma unroll
      for (int k = 0; k < 32; ++k) y[k] += k < n_here ? __ldg(P.bias + col0 + k) : 0.0f;
    }
    if (P.resid) {
      const float* rr = P.resid + (size_t)r.row * P.ld_resid + col0;
#pragma unroll
      for (int k = 0; k < 32; ++k) y[k] += k < n_here ? rr[k] : 0.0f;
    }
    if (P.relu_out) {
#pragma unroll
      for (int k = 0; k < 32; ++k) y[k] = fmaxf(y[k], 0.0f);
    }
    if (P.out_blocked) {
      float4* o4 = reinterpret_cast<float4*>(P.out) + blk_off4((size_t)r.row, col0 >> 2, P.N >> 2);
#pragma unroll
      for (int g = 0; g < 8; ++g)
        if (4 * g < n_here) o4[g * 32] = make_float4(y[4 * g], y[4 * g + 1], y[4 * g + 2], y[4 * g + 3]);
      return;
    }
    float* o = P.out + (size_t)r.row * P.ld_out + col0;
    if (n_here == 32 && (P.ld_out & 3) == 0) {
#pragma unroll
      for (int g = 0; g < 8; ++g) reinterpret_cast<float4*>(o)[g] = make_float4(y[4 * g], y[4 * g + 1], y[4 * g + 2], y[4 * g + 3]);
    } else {
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < n_here) o[k] = y[k];
    }
  }
};

// ------------------------------------------------------------------------------------------------ conv3x3 (padding 1), channels-last
struct ConvParams {
  int M, n_chunks, n_tiles_n;
  const float* wimg;
  const float* src0;       // [B][H][W][C0], or [B][2H][2W][C0] when pool
  const float* src1;       // [B][H][W][C1] second half of the channel concatenation, or null
  int C0, C1, H, W, pool, cpt;      // cpt = (C0 + C1) / 32 chunks per tap
  int blk0, blk1, blk_out;          // 1: src0 / src1 / out are in the warp-transposed layout (tc_gemm.cuh, blk_off4)
  const float* bias;
  int relu_out, N;
  float* out;              // [B][H][W][N]
};
struct ConvPolicy {
  using Params = ConvParams;
  struct Row { int row, b, y, x; };
  static __device__ __forceinline__ Row row_begin(const Params& P, int row) {
    const int hw = P.H * P.W;
    return Row{row, row / hw, (row % hw) / P.W, row % P.W};
  }
  // 32 channels c .. c + 31 of pixel `pix` of a [n_pix][C] tensor in either layout
  static __device__ __forceinline__ void load_px(const float* __restrict__ src, size_t pix, int C, int c, int blocked, float (&x)[32]) {
    const float4* p4 = blocked ? reinterpret_cast<const float4*>(src) + blk_off4(pix, c >> 2, C >> 2)
                               : reinterpret_cast<const float4*>(src + pix * C + c);
    const int step = blocked ? 32 : 1;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = __ldg(p4 + q * step);
      x[4 * q + 0] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
  }
  static __device__ __forceinline__ void load(const Params& P, const Row& r, int kc, float (&x)[32]) {
    const int tap = kc / P.cpt, c = (kc % P.cpt) * kChunk;
    const int yy = r.y + tap / 3 - 1, xx = r.x + tap % 3 - 1;
    if (r.row >= P.M || yy < 0 || yy >= P.H || xx < 0 || xx >= P.W) return zero32(x);
    if (P.pool) {            // F.max_pool2d(x, 2, 2) of the level above, taken on read
      const int W2 = 2 * P.W;
      const size_t p00 = ((size_t)r.b * 2 * P.H + 2 * yy) * W2 + 2 * xx;
      float t[32];
      load_px(P.src0, p00, P.C0, c, P.blk0, x);
      load_px(P.src0, p00 + 1, P.C0, c, P.blk0, t);
#pragma unroll
      for (int k = 0; k < 32; ++k) x[k] = fmaxf(x[k], t[k]);
      load_px(P.src0, p00 + W2, P.C0, c, P.blk0, t);
#pragma unroll
      for (int k = 0; k < 32; ++k) x[k] = fmaxf(x[k], t[k]);
      load_px(P.src0, p00 + W2 + 1, P.C0, c, P.blk0, t);
#pragma unroll
      for (int k = 0; k < 32; ++k) x[k] = fmaxf(x[k], t[k]);
      return;
    }
    const size_t pix = ((size_t)r.b * P.H + yy) * P.W + xx;
    if (c < P.C0) load_px(P.src0, pix, P.C0, c, P.blk0, x);
    else load_px(P.src1, pix, P.C1, c - P.C0, P.blk1, x);
  }
  static __device__ __forceinline__ void store(const Params& P, const Row& r, int col0, const float (&y)[32]) {
    if (r.row >= P.M || col0 >= P.N) return;
    float4* o4 = P.blk_out ? reinterpret_cast<float4*>(P.out) + blk_off4((size_t)r.row, col0 >> 2, P.N >> 2)
                           : reinterpret_cast<float4*>(P.out + (size_t)r.row * P.N + col0);
    const int step = P.blk_out ? 32 : 1;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(P.bias + col0) + g);
      float4 v = make_float4(y[4 * g] + bv.x, y[4 * g + 1] + bv.y, y[4 * g + 2] + bv.z, y[4 * g + 3] + bv.w);
      if (P.relu_out) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
      o4[g * step] = v;
    }
  }
};

// out[g][c] = max over the T rows of group g of x[g * T + t][c]   (net.max(dim=1), ONet/im2mesh/encoder/pointnet.py:103,110)
__global__ void group_max_kernel(const float* __restrict__ x, int T, int C, float* __restrict__ out) {
  const int g = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* p = x + (size_t)g * T * C + c;
    float m = -INFINITY;
    for (int t = 0; t < T; ++t) m = fmaxf(m, p[(size_t)t * C]);
    out[(size_t)g * C + c] = m;
  }
}

}  // namespace tc
}  // namespace ifd

using namespace ifd;

static int seg_chunks(int w) { return (w + tc::kChunk - 1) / tc::kChunk; }

extern "C" size_t ifd_tc_blocked_floats(long long M, int C) {
  if (M <= 0 || C <= 0 || (C & 3)) return 0;
  return (size_t)((M + 31) / 32) * 32 * (size_t)C;
}

extern "C" size_t ifd_tc_packed_floats(int N, const int* seg_widths, int n_seg) {
  if (N <= 0 || !seg_widths || n_seg < 1 || n_seg > 3) return 0;
  const int NT = tc::pick_nt(N);
  size_t chunks = 0;
  for (int s = 0; s < n_seg; ++s) {
    if (seg_widths[s] <= 0) return 0;
    chunks += seg_chunks(seg_widths[s]);
  }
  return (size_t)((N + NT - 1) / NT) * chunks * 2 * NT * tc::kChunk;
}

// W: row-major [N][sum of seg_widths] (row stride ldw); each K segment is padded to a multiple of 32 columns in the image.
extern "C" int ifd_tc_pack(const float* W, int N, const int* seg_widths, int n_seg, int ldw, float* out, ifd_stream_t stream) {
  IFD_REQUIRE(W && out && N > 0 && seg_widths && n_seg >= 1 && n_seg <= 3 && ldw > 0, "ifd_tc_pack: bad arguments");
  const int NT = tc::pick_nt(N), n_tiles = (N + NT - 1) / NT;
  int total_chunks = 0;
  for (int s = 0; s < n_seg; ++s) total_chunks += seg_chunks(seg_widths[s]);
  // one launch per segment: the segment's chunks land at their offset inside every N tile's run of chunks
  int k0 = 0, c0 = 0;
  for (int s = 0; s < n_seg; ++s) {
    const int nc = seg_chunks(seg_widths[s]);
    for (int nt = 0; nt < n_tiles; ++nt) {
      const long long total = (long long)nc * NT * tc::kChunk;
      const int n_here = N - nt * NT < NT ? N - nt * NT : NT;
      tc::pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
          W + (size_t)nt * NT * ldw + k0, n_here, seg_widths[s], ldw, 0, NT, 1, nc,
          out + ((size_t)nt * total_chunks + c0) * (2 * NT * tc::kChunk));
      IFD_LAUNCH_CHECK("tc::pack_weights_kernel");
    }
    k0 += seg_widths[s];
    c0 += nc;
  }
  return IFD_OK;
}

extern "C" int ifd_tc_linear(const ifd_tc_linear_args* a, ifd_stream_t stream) {
  IFD_REQUIRE(a && a->wimg && a->out && a->M >= 0 && a->N > 0 && a->n_seg >= 1 && a->n_seg <= 3, "ifd_tc_linear: bad arguments");
  if (a->M == 0) return IFD_OK;
  tc::LinearParams P{};
  const int NT = tc::pick_nt(a->N);
  P.M = a->M; P.N = a->N; P.n_tiles_n = (a->N + NT - 1) / NT; P.wimg = a->wimg;
  int c0 = 0;
  for (int s = 0; s < a->n_seg; ++s) {
    IFD_REQUIRE(a->a_ptr[s] && a->a_width[s] > 0 && a->a_ld[s] >= a->a_width[s], "ifd_tc_linear: bad A segment");
    IFD_REQUIRE(((uintptr_t)a->a_ptr[s] & 15) == 0 || (a->a_ld[s] & 3) != 0, "ifd_tc_linear: A segments must be 16-byte aligned");
    P.a_ptr[s] = a->a_ptr[s]; P.a_ld[s] = a->a_ld[s]; P.a_w[s] = a->a_width[s]; P.a_relu[s] = a->a_relu[s]; P.a_chunk0[s] = c0;
    P.a_group[s] = a->a_group[s] > 0 ? a->a_group[s] : 0;
    P.a_blocked[s] = a->a_blocked[s] ? 1 : 0;
    IFD_REQUIRE(!P.a_blocked[s] || ((a->a_width[s] & 3) == 0 && !P.a_group[s]), "ifd_tc_linear: a blocked segment needs width % 4 == 0 and no grouping");
    c0 += seg_chunks(a->a_width[s]);
  }
  P.n_seg = a->n_seg; P.n_chunks = c0;
  P.bias = a->bias; P.resid = a->resid; P.ld_resid = a->ld_resid; P.relu_out = a->relu_out;
  P.out = a->out; P.ld_out = a->ld_out; P.out_blocked = a->out_blocked ? 1 : 0;
  IFD_REQUIRE(!P.out_blocked || ((a->N & 3) == 0 && !a->resid), "ifd_tc_linear: a blocked output needs N % 4 == 0 and no residual");
  IFD_REQUIRE(a->ld_out >= a->N || a->shuffle_cout > 0 || a->out_blocked, "ifd_tc_linear: ld_out < N");
  IFD_REQUIRE(((uintptr_t)a->wimg & 15) == 0 && ((uintptr_t)a->out & 15) == 0, "ifd_tc_linear: wimg / out must be 16-byte aligned");
  if (a->shuffle_cout > 0) {
    IFD_REQUIRE(a->shuffle_cout % 32 == 0 && a->N == 4 * a->shuffle_cout && a->shuffle_H > 0 && a->shuffle_W > 0 &&
                    a->M % (a->shuffle_H * a->shuffle_W) == 0 && !a->resid && !a->relu_out,
                "ifd_tc_linear: bad pixel-shuffle epilogue");
    P.shuffle_cout = a->shuffle_cout; P.shuffle_H = a->shuffle_H; P.shuffle_W = a->shuffle_W;
  }
  return tc::launch<tc::LinearPolicy>(P, NT, as_stream(stream));
}

extern "C" int ifd_tc_conv3x3(const float* src0, int C0, const float* src1, int C1, int B, int H, int W, int pool, const float* wimg,
                              const float* bias, int relu_out, int Cout, float* out, int layout, ifd_stream_t stream) {
  IFD_REQUIRE(src0 && wimg && bias && out && B > 0 && H > 0 && W > 0 && Cout > 0, "ifd_tc_conv3x3: bad arguments");
  IFD_REQUIRE(C0 > 0 && C0 % 32 == 0 && C1 >= 0 && C1 % 32 == 0 && Cout % 32 == 0, "ifd_tc_conv3x3: channel counts must be multiples of 32");
  IFD_REQUIRE((C1 == 0) == (src1 == nullptr) && !(pool && C1), "ifd_tc_conv3x3: bad source combination");
  IFD_REQUIRE((((uintptr_t)src0 | (uintptr_t)src1 | (uintptr_t)wimg | (uintptr_t)bias | (uintptr_t)out) & 15) == 0, "ifd_tc_conv3x3: pointers must be 16-byte aligned");
  IFD_REQUIRE((long long)B * H * W < (1ll << 31), "ifd_tc_conv3x3: too many pixels");
  tc::ConvParams P{};
  const int NT = tc::pick_nt(Cout);
  P.M = B * H * W; P.N = Cout; P.n_tiles_n = (Cout + NT - 1) / NT; P.wimg = wimg;
  P.src0 = src0; P.src1 = src1; P.C0 = C0; P.C1 = C1; P.H = H; P.W = W; P.pool = pool ? 1 : 0;
  P.blk0 = layout & 1; P.blk1 = (layout >> 1) & 1; P.blk_out = (layout >> 2) & 1;
  P.cpt = (C0 + C1) / 32; P.n_chunks = 9 * P.cpt;
  P.bias = bias; P.relu_out = relu_out; P.out = out;
  return tc::launch<tc::ConvPolicy>(P, NT, as_stream(stream));
}

extern "C" int ifd_group_max(const float* x, int groups, int T, int C, float* out, ifd_stream_t stream) {
  IFD_REQUIRE(x && out && groups > 0 && T > 0 && C > 0, "ifd_group_max: bad arguments");
  tc::group_max_kernel<<<groups, 256, 0, as_stream(stream)>>>(x, T, C, out);
  IFD_LAUNCH_CHECK("tc::group_max_kernel");
  return IFD_OK;
}

namespace ifd {
void tc_set_cluster(int n) { tc::g_cluster_size = n >= 4 ? 4 : (n >= 2 ? 2 : 1); }
}  // namespace ifd
