// grid_point.cuh -- one query point through ConvONet's LocalDecoder in its 'grid' variant: the conditioning feature is a
// trilinear sample of a feature VOLUME instead of the sum of three plane samples.
//
// Restates (own arithmetic, reference semantics):
//   LocalDecoder.sample_grid_feature   ConvONet/src/conv_onet/models/decoder.py:59-67  (F.grid_sample on [B,C,D,H,W],
//                                      bilinear = trilinear, border, align_corners=True; grid[...,0] -> W = x, 1 -> H = y, 2 -> D = z)
//   normalize_3d_coordinate            ConvONet/src/common.py:260-276  (divisor 1 + padding + 10e-4, clamp to 1 - 10e-4)
//   the ResNet-MLP after it            decoder.py:83-93 (the same as the plane variant: ConvPoint's layer helpers)
// The volume is channels-last [R z][R y][R x][H]: one corner = one 128-byte line; coordinate2index('3d') = x + R (y + R z)
// (common.py:313), so the encoder's scatter_mean fills it in this order.
//
// __host__ __device__: inlined into the kernels; the host instantiation exists only for tests/_mathcheck.
#pragma once
#include "convonet_point.cuh"

namespace ifd {

IFD_HD PlaneCoord grid_coord(float p, float denom /* fp32(1 + padding + 10e-4) */) {
  PlaneCoord r;
  float u = add_rn(div_rn(p, denom), 0.5f);
  r.live = 1.0f;
  if (u >= 1.0f) { u = 0.999f; r.live = 0.0f; }     // p_nor[p_nor >= 1] = 1 - 10e-4
  if (u < 0.0f)  { u = 0.0f;   r.live = 0.0f; }     // p_nor[p_nor < 0] = 0.0
  r.u = u;
  return r;
}

template <int H>
struct GridPoint {
  static_assert(H == 32, "sign masks are one 32-bit word per layer");
  using L = ConvDecLayout<H>;
  Axis ax[3];                      // x (W), y (H), z (D)
  uint32_t mask_a[kMaxBlocks], mask_h[kMaxBlocks], mask_f;

  // corner t = 4 dz + 2 dy + dx (ATen's order tnw, tne, tsw, tse, bnw, bne, bsw, bse): offset in floats and weight
  // (x-term * y-term) * z-term with near = (floor + 1) - i, far = i - floor  (ATen grid_sampler_3d_cpu_impl)
  IFD_HD void taps(int R, int (&off)[8], float (&w)[8]) const {
    int i0[3], i1[3];
    float wn[3], wf[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      i0[a] = ax[a].i0;
      i1[a] = ax[a].has1 ? ax[a].i0 + 1 : ax[a].i0;
      wn[a] = ax[a].near_w;
      wf[a] = ax[a].has1 ? ax[a].f : 0.0f;          // an out-of-bounds far corner contributes 0
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int dx = t & 1, dy = (t >> 1) & 1, dz = t >> 2;
      off[t] = (((dz ? i1[2] : i0[2]) * R + (dy ? i1[1] : i0[1])) * R + (dx ? i1[0] : i0[0])) * H;
      w[t] = ((dx ? wf[0] : wn[0]) * (dy ? wf[1] : wn[1])) * (dz ? wf[2] : wn[2]);
    }
  }

  // vol points at this cloud's [R][R][R][H] channels-last volume.  Returns the logit.
  IFD_HD float forward(const float* __restrict__ Wb, const float* __restrict__ vol, float px, float py, float pz, int R,
                       float denom, int n_blocks) {
    ax[0] = axis_setup(grid_coord(px, denom), R, denom);
    ax[1] = axis_setup(grid_coord(py, denom), R, denom);
    ax[2] = axis_setup(grid_coord(pz, denom), R, denom);
    int off[8];
    float tw[8];
    taps(R, off, tw);
    float c[H];
#pragma unroll
    for (int k = 0; k < H; ++k) c[k] = 0.0f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float4* v = reinterpret_cast<const float4*>(vol + off[t]);
#pragma unroll
      for (int q = 0; q < H / 4; ++q) {
        const float4 x = v[q];
        c[4 * q + 0] = fmaf(x.x, tw[t], c[4 * q + 0]);
        c[4 * q + 1] = fmaf(x.y, tw[t], c[4 * q + 1]);
        c[4 * q + 2] = fmaf(x.z, tw[t], c[4 * q + 2]);
        c[4 * q + 3] = fmaf(x.w, tw[t], c[4 * q + 3]);
      }
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
      for (int o = 0; o < H; ++o) net[o] += h[o];
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
  IFD_HD void backward(const float* __restrict__ Wb, const float* __restrict__ vol, float glogit, int R, int n_blocks,
                       float (&gp)[3]) const {
    float gnet[H], gh[H], gc[H];
    const float* wo = Wb + L::out_w(n_blocks);
#pragma unroll
    for (int k = 0; k < H; ++k) {
      gnet[k] = ((mask_f >> k) & 1u) ? glogit * wo[k] : 0.0f;
      gc[k] = 0.0f;
    }
#pragma unroll 1
    for (int i = n_blocks - 1; i >= 0; --i) {
      matvecT_masked<H, false>(Wb + L::fc_1(i), gnet, mask_h[i], gh);
      matvecT_masked<H, true>(Wb + L::fc_0(i), gh, mask_a[i], gnet);
      matvecT_masked<H, true>(Wb + L::fc_c(i), gnet, 0xffffffffu, gc);
    }
    float g[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float s = 0.0f;
#pragma unroll
      for (int o = 0; o < H; ++o) s = fmaf(Wb[L::kFcpW + a * H + o], gnet[o], s);
      g[a] = s;
    }
    // q[t] = <corner t, gc>; d c / d ix is linear in them: corners differing in dx, weighted by the other two axes' weights
    int off[8];
    float tw[8];
    taps(R, off, tw);
    float q[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float4* v = reinterpret_cast<const float4*>(vol + off[t]);
      float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
      for (int k4 = 0; k4 < H / 4; ++k4) {
        const float4 x = v[k4];
        s0 = fmaf(x.x, gc[4 * k4 + 0], s0);
        s1 = fmaf(x.y, gc[4 * k4 + 1], s1);
        s2 = fmaf(x.z, gc[4 * k4 + 2], s2);
        s3 = fmaf(x.w, gc[4 * k4 + 3], s3);
      }
      const int dx = t & 1, dy = (t >> 1) & 1, dz = t >> 2;
      const bool exists = (!dx || ax[0].has1) && (!dy || ax[1].has1) && (!dz || ax[2].has1);
      q[t] = exists ? (s0 + s1) + (s2 + s3) : 0.0f;               // ATen masks out-of-bounds corners
    }
    float gi[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int b = (a + 1) % 3, c2 = (a + 2) % 3;                // the two other axes
      const float wb[2] = {ax[b].near_w, ax[b].f}, wc[2] = {ax[c2].near_w, ax[c2].f};
      float s = 0.0f;
#pragma unroll
      for (int eb = 0; eb < 2; ++eb)
#pragma unroll
        for (int ec = 0; ec < 2; ++ec) {
          const int lo = (eb << b) | (ec << c2), hi = lo | (1 << a);
          s += (q[hi] - q[lo]) * (wb[eb] * wc[ec]);
        }
      gi[a] = s;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) gp[a] = g[a] + gi[a] * ax[a].dscale;
  }
};

}  // namespace ifd
