// decode_v2.cuh -- the hot decode kernel of the restoration loop (BCE mode), second generation.
//
// What ncu said about v1 (profiles/r01_v1_*): FMA pipe 28 % busy, issue 36 %, stalls dominated by
// short_scoreboard (weight LDS), no_instruction (120 KB of unrolled layer bodies thrash the I-cache) and
// long_scoreboard (thread-per-point gather: every LDG.128 touches 32 different lines).  v2 changes:
//   * one generic forward-layer body and one generic backward-layer body (~20 KB of SASS each), executed 15
//     times each with register-role glue in between, instead of 30 unrolled bodies;
//   * two points per thread, so every weight LDS.128 feeds 8 FMAs instead of 4 (the LSU is the co-limiter);
//   * packed fma.rn.f32x2 (FFMA2): half the issue slots for the same FP32 work, leaving slots for LDS;
//   * warp-cooperative gather: 8 lanes x float4 read one 128-byte texel, 4 points per instruction, i.e. one
//     L1 wavefront per tap instead of 32, both in the forward (feature) and backward (dot with g_c) pass.
// Arithmetic per point is the same as convonet_point.cuh (same accumulation order in the MLP), so v1 stays
// the readable definition and the step-level seam; tests compare the two.
#pragma once
#include "convonet_point.cuh"

namespace ifd {

constexpr int kV2Threads = 256;
constexpr int kV2Pts = 512;             // points per CTA: thread t owns points t and t + 256 of the tile
constexpr int kV2Stride = kV2Pts + 1;   // feature rows [k][pt] padded: conflict-free for both access patterns

struct DecodeV2Smem {
  // weights (wtotal4 float4) | feat [32][kV2Stride] | gpart [kV2Pts][4]
  static __host__ __device__ size_t bytes(int wtotal4) {
    return (size_t)wtotal4 * 16 + (size_t)32 * kV2Stride * 4 + (size_t)kV2Pts * 16;
  }
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// T[p][o] += sum_k W[k][o] * X[p][k]   for two points; T as 16 float2 pairs over o.
__device__ __forceinline__ void layer_fwd(const float4* __restrict__ Wl4, const float (&x)[2][32], float2 (&t)[2][16]) {
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float2 xa = make_float2(x[0][k], x[0][k]);
    const float2 xb = make_float2(x[1][k], x[1][k]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 w = Wl4[k * 8 + j];
      const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
      t[0][2 * j] = ffma2(w01, xa, t[0][2 * j]);
      t[0][2 * j + 1] = ffma2(w23, xa, t[0][2 * j + 1]);
      t[1][2 * j] = ffma2(w01, xb, t[1][2 * j]);
      t[1][2 * j + 1] = ffma2(w23, xb, t[1][2 * j + 1]);
    }
  }
}

// T[p][k] = sum_o W[k][o] * G[p][o]   (dgrad), G as 16 float2 pairs over o; same 4 partial sums as v1.
__device__ __forceinline__ void layer_bwd(const float4* __restrict__ Wl4, const float2 (&g)[2][16], float (&t)[2][32]) {
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    float2 a0 = make_float2(0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 w = Wl4[k * 8 + j];
      const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
      a0 = ffma2(w01, g[0][2 * j], a0);
      a1 = ffma2(w23, g[0][2 * j + 1], a1);
      b0 = ffma2(w01, g[1][2 * j], b0);
      b1 = ffma2(w23, g[1][2 * j + 1], b1);
    }
    t[0][k] = (a0.x + a0.y) + (a1.x + a1.y);
    t[1][k] = (b0.x + b0.y) + (b1.x + b1.y);
  }
}

struct TapSet {          // bilinear geometry of one point, all three planes
  int off[3][4];         // float offsets of the 4 taps inside a plane
  float w[3][4];
  Axis ax[3];
};
__device__ __forceinline__ void tapset(float px, float py, float pz, int R, float denom, TapSet& ts) {
  ts.ax[0] = axis_setup(plane_coord(px, denom), R, denom);
  ts.ax[1] = axis_setup(plane_coord(py, denom), R, denom);
  ts.ax[2] = axis_setup(plane_coord(pz, denom), R, denom);
#pragma unroll
  for (int pl = 0; pl < 3; ++pl) {
    const Axis& aw = ts.ax[plane_axis_w(pl)];
    const Axis& ah = ts.ax[plane_axis_h(pl)];
    const int w0 = aw.i0, w1 = aw.has1 ? aw.i0 + 1 : aw.i0;
    const int h0 = ah.i0, h1 = ah.has1 ? ah.i0 + 1 : ah.i0;
    ts.off[pl][0] = (h0 * R + w0) * 32; ts.off[pl][1] = (h0 * R + w1) * 32;
    ts.off[pl][2] = (h1 * R + w0) * 32; ts.off[pl][3] = (h1 * R + w1) * 32;
    const float fw = aw.has1 ? aw.f : 0.0f, fh = ah.has1 ? ah.f : 0.0f;
    ts.w[pl][0] = ah.near_w * aw.near_w; ts.w[pl][1] = ah.near_w * fw;
    ts.w[pl][2] = fh * aw.near_w;        ts.w[pl][3] = fh * fw;
  }
}

struct DecodeV2Args {
  const float* planes;   // [3][B][R][R][32]
  const float* W;
  const float* xyz;      // [B][K][3]
  float* grad_out;       // [B][K][3]
  double* stat_part;     // optional [gridDim.x][2]
  int n, K, B, R, n_blocks, wtotal4;
  float denom, target, ginv;
};

__global__ void __launch_bounds__(kV2Threads, 1) convonet_decode_v2_kernel(const DecodeV2Args a) {
  extern __shared__ float4 smem4[];
  float* Wb = reinterpret_cast<float*>(smem4);
  float* feat = Wb + (size_t)a.wtotal4 * 4;                  // [32][kV2Stride]: c on the way in, g_c on the way out
  float4* gpart = reinterpret_cast<float4*>(feat + 32 * kV2Stride);   // [kV2Pts]: fc_p part of the gradient
  using L = ConvDecLayout<32>;
  {
    const float4* src = reinterpret_cast<const float4*>(a.W);
    for (int i = threadIdx.x; i < a.wtotal4; i += kV2Threads) smem4[i] = src[i];
  }
  const int tile0 = blockIdx.x * kV2Pts;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane >> 3, j4 = lane & 7;                  // 8 lanes x float4 = one 128-byte texel
  const size_t plane_sz = (size_t)a.R * a.R * 32;

  // ---------------- forward gather: warp w serves its own threads' points (tile slots 32w.. and 256+32w..)
  for (int it = 0; it < 16; ++it) {
    const int slot = (it < 8 ? 0 : kV2Pts / 2) + warp * 32 + (it & 7) * 4 + grp;
    const int pi = min(tile0 + slot, a.n - 1);
    const int b = pi / a.K;
    const float px = a.xyz[(size_t)pi * 3 + 0], py = a.xyz[(size_t)pi * 3 + 1], pz = a.xyz[(size_t)pi * 3 + 2];
    TapSet ts;
    tapset(px, py, pz, a.R, a.denom, ts);
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      const float* base = a.planes + ((size_t)pl * a.B + b) * plane_sz + j4 * 4;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(base + ts.off[pl][t]));
        s.x = fmaf(v.x, ts.w[pl][t], s.x);
        s.y = fmaf(v.y, ts.w[pl][t], s.y);
        s.z = fmaf(v.z, ts.w[pl][t], s.z);
        s.w = fmaf(v.w, ts.w[pl][t], s.w);
      }
      c.x += s.x; c.y += s.y; c.z += s.z; c.w += s.w;
    }
    feat[(j4 * 4 + 0) * kV2Stride + slot] = c.x;
    feat[(j4 * 4 + 1) * kV2Stride + slot] = c.y;
    feat[(j4 * 4 + 2) * kV2Stride + slot] = c.z;
    feat[(j4 * 4 + 3) * kV2Stride + slot] = c.w;
  }
  __syncthreads();   // weights staged by all threads; features are warp-local but one barrier covers both

  // ---------------- MLP forward, two points per thread
  const int slot0 = threadIdx.x, slot1 = threadIdx.x + kV2Pts / 2;
  const int pi0 = min(tile0 + slot0, a.n - 1), pi1 = min(tile0 + slot1, a.n - 1);
  float p[2][3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    p[0][d] = a.xyz[(size_t)pi0 * 3 + d];
    p[1][d] = a.xyz[(size_t)pi1 * 3 + d];
  }
  float net[2][32];
  float x[2][32];
  float2 t[2][16];
  uint32_t mask_a[2][kMaxBlocks], mask_h[2][kMaxBlocks];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int o = 0; o < 32; ++o) {
      float v = Wb[L::kFcpB + o];
      v = fmaf(Wb[L::kFcpW + 0 * 32 + o], p[q][0], v);
      v = fmaf(Wb[L::kFcpW + 1 * 32 + o], p[q][1], v);
      v = fmaf(Wb[L::kFcpW + 2 * 32 + o], p[q][2], v);
      net[q][o] = v;
    }
  const int n_layers = 3 * a.n_blocks;
#pragma unroll 1
  for (int l = 0; l < n_layers; ++l) {
    const int type = l % 3, blk = l / 3;
    const float* Wl = Wb + L::kBlk0 + l * L::kLayer;
    if (type == 0) {                       // fc_c[i](c): input is the sampled feature
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        x[0][k] = feat[k * kV2Stride + slot0];
        x[1][k] = feat[k * kV2Stride + slot1];
      }
    } else if (type == 1) {                // fc_0(relu(net))
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t m = 0;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          m |= (net[q][k] > 0.0f ? 1u : 0u) << k;
          x[q][k] = fmaxf(net[q][k], 0.0f);
        }
        mask_a[q][blk] = m;
      }
    } else {                               // fc_1(relu(h)), h is the previous layer's output
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t m = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          m |= (t[q][k].x > 0.0f ? 1u : 0u) << (2 * k);
          m |= (t[q][k].y > 0.0f ? 1u : 0u) << (2 * k + 1);
          x[q][2 * k] = fmaxf(t[q][k].x, 0.0f);
          x[q][2 * k + 1] = fmaxf(t[q][k].y, 0.0f);
        }
        mask_h[q][blk] = m;
      }
    }
    if (type == 1) {                       // h = bias + W x
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 16; ++k) t[q][k] = make_float2(Wl[1024 + 2 * k], Wl[1024 + 2 * k + 1]);
    } else if (type == 0) {                // tmp = bias + W c ; net += tmp   (v1 order)
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 16; ++k) t[q][k] = make_float2(Wl[1024 + 2 * k], Wl[1024 + 2 * k + 1]);
    } else {                               // net = (net + bias) + W relu(h)   (v1 order: accumulate into net)
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 16; ++k)
          t[q][k] = make_float2(net[q][2 * k] + Wl[1024 + 2 * k], net[q][2 * k + 1] + Wl[1024 + 2 * k + 1]);
    }
    layer_fwd(reinterpret_cast<const float4*>(Wl), x, t);
    if (type == 0) {
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          net[q][2 * k] += t[q][k].x;
          net[q][2 * k + 1] += t[q][k].y;
        }
    } else if (type == 2) {
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          net[q][2 * k] = t[q][k].x;
          net[q][2 * k + 1] = t[q][k].y;
        }
    }
  }
  // ---------------- fc_out, BCE gradient, optional diagnostics
  const float* wo = Wb + L::out_w(a.n_blocks);
  float glogit[2], logit[2];
  uint32_t mask_f[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float lg = Wb[L::out_b(a.n_blocks)];
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      m |= (net[q][k] > 0.0f ? 1u : 0u) << k;
      lg = fmaf(wo[k], fmaxf(net[q][k], 0.0f), lg);
    }
    mask_f[q] = m;
    logit[q] = lg;
    glogit[q] = (sigmoidf_(lg) - a.target) * a.ginv;
  }
  if (a.stat_part) {
    __shared__ double red[2][kV2Threads / 32];
    const bool l0 = tile0 + slot0 < a.n, l1 = tile0 + slot1 < a.n;
    double s0 = (l0 ? (double)bce_with_logits(logit[0], a.target) : 0.0) + (l1 ? (double)bce_with_logits(logit[1], a.target) : 0.0);
    double s1 = (l0 ? (double)sigmoidf_(logit[0]) : 0.0) + (l1 ? (double)sigmoidf_(logit[1]) : 0.0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) {
      red[0][warp] = s0;
      red[1][warp] = s1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t0 = 0.0, t1 = 0.0;
      for (int w = 0; w < kV2Threads / 32; ++w) {
        t0 += red[0][w];
        t1 += red[1][w];
      }
      a.stat_part[blockIdx.x * 2 + 0] = t0;
      a.stat_part[blockIdx.x * 2 + 1] = t1;
    }
  }

  // ---------------- MLP backward (dgrad): roles  gnet := net,  g2 := t,  tt := x
  float (&gnet)[2][32] = net;
  float2 (&g2)[2][16] = t;
  float (&tt)[2][32] = x;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int k = 0; k < 32; ++k) gnet[q][k] = ((mask_f[q] >> k) & 1u) ? glogit[q] * wo[k] : 0.0f;
#pragma unroll 1
  for (int l = n_layers - 1; l >= 0; --l) {
    const int type = l % 3, blk = l / 3;
    const float* Wl = Wb + L::kBlk0 + l * L::kLayer;
    if (type != 1) {                       // through fc_1^T or fc_c^T: the incoming gradient is gnet
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 16; ++k) g2[q][k] = make_float2(gnet[q][2 * k], gnet[q][2 * k + 1]);
    } else {                               // through fc_0^T: incoming is gh = (fc_1^T gnet) masked by relu(h)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t m = mask_h[q][blk];
#pragma unroll
        for (int k = 0; k < 16; ++k)
          g2[q][k] = make_float2(((m >> (2 * k)) & 1u) ? tt[q][2 * k] : 0.0f, ((m >> (2 * k + 1)) & 1u) ? tt[q][2 * k + 1] : 0.0f);
      }
    }
    layer_bwd(reinterpret_cast<const float4*>(Wl), g2, tt);
    if (type == 1) {                       // gnet += (fc_0^T gh) masked by relu(net)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t m = mask_a[q][blk];
#pragma unroll
        for (int k = 0; k < 32; ++k) gnet[q][k] += ((m >> k) & 1u) ? tt[q][k] : 0.0f;
      }
    } else if (type == 0) {                // g_c += fc_c^T gnet, kept in the feature buffer (c is dead by now)
      const bool first = blk == a.n_blocks - 1;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float* f0 = feat + k * kV2Stride + slot0;
        float* f1 = feat + k * kV2Stride + slot1;
        *f0 = first ? tt[0][k] : *f0 + tt[0][k];
        *f1 = first ? tt[1][k] : *f1 + tt[1][k];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {           // through fc_p
    float g[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float s = 0.0f;
#pragma unroll
      for (int o = 0; o < 32; ++o) s = fmaf(Wb[L::kFcpW + d * 32 + o], gnet[q][o], s);
      g[d] = s;
    }
    gpart[q == 0 ? slot0 : slot1] = make_float4(g[0], g[1], g[2], 0.f);
  }
  __syncwarp();      // g_c and gpart of a warp's points were written by that warp's own threads

  // ---------------- backward gather: q_t = <g_c, texel_t> for the 12 taps, reduced over the 8 lanes of a group
  for (int it = 0; it < 16; ++it) {
    const int slot = (it < 8 ? 0 : kV2Pts / 2) + warp * 32 + (it & 7) * 4 + grp;
    const int pi_raw = tile0 + slot;
    const int pi = min(pi_raw, a.n - 1);
    const int b = pi / a.K;
    const float px = a.xyz[(size_t)pi * 3 + 0], py = a.xyz[(size_t)pi * 3 + 1], pz = a.xyz[(size_t)pi * 3 + 2];
    TapSet ts;
    tapset(px, py, pz, a.R, a.denom, ts);
    const float4 gc = make_float4(feat[(j4 * 4 + 0) * kV2Stride + slot], feat[(j4 * 4 + 1) * kV2Stride + slot],
                                  feat[(j4 * 4 + 2) * kV2Stride + slot], feat[(j4 * 4 + 3) * kV2Stride + slot]);
    float gi[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      const float* base = a.planes + ((size_t)pl * a.B + b) * plane_sz + j4 * 4;
      float qv[4];
#pragma unroll
      for (int t4 = 0; t4 < 4; ++t4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(base + ts.off[pl][t4]));
        float s = (v.x * gc.x + v.y * gc.y) + (v.z * gc.z + v.w * gc.w);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        qv[t4] = s;
      }
      const Axis& aw = ts.ax[plane_axis_w(pl)];
      const Axis& ah = ts.ax[plane_axis_h(pl)];
      const float q_ne = aw.has1 ? qv[1] : 0.0f;
      const float q_sw = ah.has1 ? qv[2] : 0.0f;
      const float q_se = (aw.has1 && ah.has1) ? qv[3] : 0.0f;
      gi[plane_axis_w(pl)] += (q_ne - qv[0]) * (1.0f - ah.f) + (q_se - q_sw) * ah.f;
      gi[plane_axis_h(pl)] += (q_sw - qv[0]) * (1.0f - aw.f) + (q_se - q_ne) * aw.f;
    }
    if (j4 == 0 && pi_raw < a.n) {
      const float4 gp = gpart[slot];
      a.grad_out[(size_t)pi * 3 + 0] = gp.x + gi[0] * ts.ax[0].dscale;
      a.grad_out[(size_t)pi * 3 + 1] = gp.y + gi[1] * ts.ax[1].dscale;
      a.grad_out[(size_t)pi * 3 + 2] = gp.z + gi[2] * ts.ax[2].dscale;
    }
  }
}

}  // namespace ifd
