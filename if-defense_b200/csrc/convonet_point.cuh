// convonet_point.cuh -- one query point through ConvONet's LocalDecoder, forward and d/dxyz.
//
// Restates (own arithmetic, reference semantics):
//   LocalDecoder.forward            ConvONet/src/conv_onet/models/decoder.py:69-95
//   sample_plane_feature            decoder.py:50-57   (F.grid_sample bilinear / border / align_corners=True)
//   normalize_coordinate            ConvONet/src/common.py:235-258
//   ResnetBlockFC.forward           ConvONet/src/layers.py:39-48
// The model is frozen (opt_defense.py:72-73): backward is dgrad only and needs nothing from the forward
// but the ReLU sign masks (2 per block + 1) and the bilinear tap geometry, so nothing is spilled to memory.
//
// __host__ __device__: inlined into the kernels; the host instantiation exists only for tests/_mathcheck.
#pragma once
#include "ifd_math.cuh"

#if !defined(__CUDACC__)
struct alignas(16) float4 { float x, y, z, w; };
#endif

namespace ifd {

// Packed-parameter layout (see include/ifd_b200.h, ifd_convonet_decoder_nfloats); H == C.
template <int H>
struct ConvDecLayout {
  static constexpr int kFcpW = 0;              // [3][H]
  static constexpr int kFcpB = 3 * H;          // [H]
  static constexpr int kBlk0 = 4 * H;
  static constexpr int kLayer = H * H + H;     // W^T [H][H] then b [H]
  static constexpr int kBlk = 3 * kLayer;      // fc_c, fc_0, fc_1
  IFD_HD static int fc_c(int i) { return kBlk0 + i * kBlk; }
  IFD_HD static int fc_0(int i) { return kBlk0 + i * kBlk + kLayer; }
  IFD_HD static int fc_1(int i) { return kBlk0 + i * kBlk + 2 * kLayer; }
  IFD_HD static int out_w(int nb) { return kBlk0 + nb * kBlk; }
  IFD_HD static int out_b(int nb) { return kBlk0 + nb * kBlk + H; }
  IFD_HD static int total(int nb) { return kBlk0 + nb * kBlk + H + 1; }
};

constexpr int kMaxBlocks = 8;

// acc[o] += sum_k x[k] * WT[k][o]      (forward Linear, weights stored transposed [in][out])
template <int H, bool RELU_IN>
IFD_HD void matvec_acc(const float* __restrict__ WT, const float (&x)[H], float (&acc)[H]) {
#pragma unroll
  for (int k = 0; k < H; ++k) {
    float xk = x[k];
    if (RELU_IN) xk = xk > 0.0f ? xk : 0.0f;
#pragma unroll
    for (int o = 0; o < H; o += 4) {
      const float4 w = *reinterpret_cast<const float4*>(WT + k * H + o);
      acc[o + 0] = fmaf(w.x, xk, acc[o + 0]);
      acc[o + 1] = fmaf(w.y, xk, acc[o + 1]);
      acc[o + 2] = fmaf(w.z, xk, acc[o + 2]);
      acc[o + 3] = fmaf(w.w, xk, acc[o + 3]);
    }
  }
}

// out[k] (+)= [mask bit k] * sum_o WT[k][o] * g[o]     (dgrad through Linear then through the ReLU feeding it)
template <int H, bool ACCUM>
IFD_HD void matvecT_masked(const float* __restrict__ WT, const float (&g)[H], uint32_t mask, float (&out)[H]) {
#pragma unroll
  for (int k = 0; k < H; ++k) {
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
    for (int o = 0; o < H; o += 4) {
      const float4 w = *reinterpret_cast<const float4*>(WT + k * H + o);
      s0 = fmaf(w.x, g[o + 0], s0);
      s1 = fmaf(w.y, g[o + 1], s1);
      s2 = fmaf(w.z, g[o + 2], s2);
      s3 = fmaf(w.w, g[o + 3], s3);
    }
    float s = (s0 + s1) + (s2 + s3);
    s = ((mask >> k) & 1u) ? s : 0.0f;
    out[k] = ACCUM ? out[k] + s : s;
  }
}

template <int H>
IFD_HD uint32_t sign_mask(const float (&x)[H]) {
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < H; ++k) m |= (x[k] > 0.0f ? 1u : 0u) << k;
  return m;
}

// Which coordinate feeds the W (fast) and H (slow) axis of each plane: xz, xy, yz (common.py:243-248;
// grid_sample's grid[...,0] indexes W, grid[...,1] indexes H).
IFD_HD int plane_axis_w(int pl) { return pl == 2 ? 1 : 0; }
IFD_HD int plane_axis_h(int pl) { return pl == 1 ? 1 : 2; }

template <int H>
struct ConvPoint {
  static_assert(H == 32, "sign masks are one 32-bit word per layer");
  using L = ConvDecLayout<H>;
  Axis ax[3];                      // bilinear geometry of x, y, z (shared by the planes that use them)
  uint32_t mask_a[kMaxBlocks];     // net > 0 before fc_0 of block i
  uint32_t mask_h[kMaxBlocks];     // h > 0 before fc_1 of block i
  uint32_t mask_f;                 // net > 0 before fc_out

  // One plane's 4 taps: offsets (in floats, channels-last) and weights nw, ne, sw, se.
  IFD_HD void taps(int pl, int R, int (&off)[4], float (&w)[4]) const {
    const Axis& aw = ax[plane_axis_w(pl)];
    const Axis& ah = ax[plane_axis_h(pl)];
    const int w0 = aw.i0, w1 = aw.has1 ? aw.i0 + 1 : aw.i0;
    const int h0 = ah.i0, h1 = ah.has1 ? ah.i0 + 1 : ah.i0;
    off[0] = (h0 * R + w0) * H; off[1] = (h0 * R + w1) * H;
    off[2] = (h1 * R + w0) * H; off[3] = (h1 * R + w1) * H;
    const float fw = aw.has1 ? aw.f : 0.0f, fh = ah.has1 ? ah.f : 0.0f;
    w[0] = ah.near_w * aw.near_w;  // nw = s * e
    w[1] = ah.near_w * fw;         // ne = s * w
    w[2] = fh * aw.near_w;         // sw = n * e
    w[3] = fh * fw;                // se = n * w
  }

  // planes[pl] points at this cloud's [R][R][H] channels-last plane.  Returns the logit.
  IFD_HD float forward(const float* __restrict__ Wb, const float* const (&planes)[3], float px, float py, float pz,
                       int R, float denom, int n_blocks) {
    ax[0] = axis_setup(plane_coord(px, denom), R, denom);
    ax[1] = axis_setup(plane_coord(py, denom), R, denom);
    ax[2] = axis_setup(plane_coord(pz, denom), R, denom);

    float c[H];
#pragma unroll
    for (int k = 0; k < H; ++k) c[k] = 0.0f;
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      int off[4];
      float tw[4];
      taps(pl, R, off, tw);
      float s[H];
#pragma unroll
      for (int k = 0; k < H; ++k) s[k] = 0.0f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4* v = reinterpret_cast<const float4*>(planes[pl] + off[t]);
#pragma unroll
        for (int q = 0; q < H / 4; ++q) {
          const float4 x = v[q];
          s[4 * q + 0] = fmaf(x.x, tw[t], s[4 * q + 0]);
          s[4 * q + 1] = fmaf(x.y, tw[t], s[4 * q + 1]);
          s[4 * q + 2] = fmaf(x.z, tw[t], s[4 * q + 2]);
          s[4 * q + 3] = fmaf(x.w, tw[t], s[4 * q + 3]);
        }
      }
#pragma unroll
      for (int k = 0; k < H; ++k) c[k] += s[k];                  // c += sample_plane_feature(...)
    }

    float net[H], h[H];
#pragma unroll
    for (int o = 0; o < H; ++o) {                                 // net = fc_p(p)
      float a = Wb[L::kFcpB + o];
      a = fmaf(Wb[L::kFcpW + 0 * H + o], px, a);
      a = fmaf(Wb[L::kFcpW + 1 * H + o], py, a);
      a = fmaf(Wb[L::kFcpW + 2 * H + o], pz, a);
      net[o] = a;
    }
#pragma unroll 1
    for (int i = 0; i < n_blocks; ++i) {
      const float* wc = Wb + L::fc_c(i);
      const float* w0 = Wb + L::fc_0(i);
      const float* w1 = Wb + L::fc_1(i);
#pragma unroll
      for (int o = 0; o < H; ++o) h[o] = wc[H * H + o];
      matvec_acc<H, false>(wc, c, h);                             // fc_c[i](c)
#pragma unroll
      for (int o = 0; o < H; ++o) net[o] += h[o];                 // net = net + fc_c[i](c)
      mask_a[i] = sign_mask<H>(net);
#pragma unroll
      for (int o = 0; o < H; ++o) h[o] = w0[H * H + o];
      matvec_acc<H, true>(w0, net, h);                            // h = fc_0(relu(net))
      mask_h[i] = sign_mask<H>(h);
#pragma unroll
      for (int o = 0; o < H; ++o) net[o] += w1[H * H + o];
      matvec_acc<H, true>(w1, h, net);                            // net = net + fc_1(relu(h))
    }
    mask_f = sign_mask<H>(net);
    float logit = Wb[L::out_b(n_blocks)];
    const float* wo = Wb + L::out_w(n_blocks);
#pragma unroll
    for (int k = 0; k < H; ++k) logit = fmaf(wo[k], net[k] > 0.0f ? net[k] : 0.0f, logit);
    return logit;
  }

  // d(glogit * logit)/d(px,py,pz).  Must follow forward() on the same object.
  IFD_HD void backward(const float* __restrict__ Wb, const float* const (&planes)[3], float glogit, int R,
                       int n_blocks, float (&gp)[3]) const {
    float gnet[H], gh[H], gc[H];
    const float* wo = Wb + L::out_w(n_blocks);
#pragma unroll
    for (int k = 0; k < H; ++k) {
      gnet[k] = ((mask_f >> k) & 1u) ? glogit * wo[k] : 0.0f;
      gc[k] = 0.0f;
    }
#pragma unroll 1
    for (int i = n_blocks - 1; i >= 0; --i) {
      matvecT_masked<H, false>(Wb + L::fc_1(i), gnet, mask_h[i], gh);     // through fc_1 and relu(h)
      matvecT_masked<H, true>(Wb + L::fc_0(i), gh, mask_a[i], gnet);      // through fc_0 and relu(net), + residual
      matvecT_masked<H, true>(Wb + L::fc_c(i), gnet, 0xffffffffu, gc);    // into the sampled feature
    }
    float g[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {                                          // through fc_p
      float s = 0.0f;
#pragma unroll
      for (int o = 0; o < H; ++o) s = fmaf(Wb[L::kFcpW + a * H + o], gnet[o], s);
      g[a] = s;
    }
    float gi[3] = {0.0f, 0.0f, 0.0f};                                      // d/d(ix) per coordinate axis
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      int off[4];
      float tw[4];
      taps(pl, R, off, tw);
      float q[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4* v = reinterpret_cast<const float4*>(planes[pl] + off[t]);
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
          const float4 x = v[k4];
          s0 = fmaf(x.x, gc[4 * k4 + 0], s0);
          s1 = fmaf(x.y, gc[4 * k4 + 1], s1);
          s2 = fmaf(x.z, gc[4 * k4 + 2], s2);
          s3 = fmaf(x.w, gc[4 * k4 + 3], s3);
        }
        q[t] = (s0 + s1) + (s2 + s3);
      }
      const Axis& aw = ax[plane_axis_w(pl)];
      const Axis& ah = ax[plane_axis_h(pl)];
      // a missing far corner contributes value 0 (ATen masks out-of-bounds corners)
      const float q_ne = aw.has1 ? q[1] : 0.0f;
      const float q_sw = ah.has1 ? q[2] : 0.0f;
      const float q_se = (aw.has1 && ah.has1) ? q[3] : 0.0f;
      const float fh = ah.f, fw = aw.f;
      gi[plane_axis_w(pl)] += (q_ne - q[0]) * (1.0f - fh) + (q_se - q_sw) * fh;
      gi[plane_axis_h(pl)] += (q_sw - q[0]) * (1.0f - fw) + (q_se - q_ne) * fw;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) gp[a] = g[a] + gi[a] * ax[a].dscale;
  }
};

}  // namespace ifd
