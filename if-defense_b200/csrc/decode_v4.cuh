// decode_v4.cuh -- decode v3 (decode_v3.cuh: gather + ResNet-MLP on tcgen05, 3xTF32) re-cut so that TWO CTAs share an SM.
//
// ncu on v3 (profiles/r01_final_*): one 512-thread CTA per SM, issue slots 40 % busy -- the kernel is a latency chain
// (two L2-bound gather phases around a 30-layer tcgen05 round-trip chain), and since the loops of two batches run
// side by side (ifd_convonet_opt_batches) SM time is what limits throughput.  v4 = 256 threads, 256 points, 2 tiles,
// 192 TMEM columns, 128 registers, and 88 KB of shared memory because the weight images are streamed per ResNet block
// (one TMA bulk copy per stage) instead of held for the whole direction: two CTAs (from the same or from different launches) are resident per SM and
// fill each other's stalls.  Arithmetic, layouts and helpers are v3's (same bits per point).
#pragma once
#include "decode_v3.cuh"

namespace ifd {

constexpr int kV4Threads = 256;
constexpr int kV4Pts = 256;
constexpr int kV4Stride = kV4Pts + 1;
constexpr int kV4StageFloats = 3 * kV3ImgFloats;   // one ResNet block, one direction: 3 layers x (hi + lo) = 24 KB
constexpr uint32_t kV4TmemCols = 256;

struct DecodeV4Smem {
  static __host__ __device__ size_t bytes(int n_blocks) {
    return (size_t)2 * kV4StageFloats * 4 + (size_t)32 * kV4Stride * 4 + (size_t)kV4Pts * 16 + 256 +
           (size_t)(3 * n_blocks + 6) * 32 * 4;
  }
};

__global__ void __launch_bounds__(kV4Threads, 2) convonet_decode_v4_kernel(const DecodeV3Args a) {
  extern __shared__ float4 smem4[];
  // graph replay: the buffers of THIS launch are in the job record (uniform branch, five 8-byte loads per thread)
  const float* __restrict__ g_planes = a.job ? a.job->planes : a.planes;
  const float* __restrict__ g_W = a.job ? a.job->W : a.W;
  const float* __restrict__ g_Wimg = a.job ? a.job->Wimg : a.Wimg;
  const float* __restrict__ g_xyz = a.job ? a.job->xyz : a.xyz;
  float* __restrict__ g_grad = a.job ? a.job->g_occ : a.grad_out;
  using L = ConvDecLayout<32>;
  const int n_layers = 3 * a.n_blocks;
  float* wimg = reinterpret_cast<float*>(smem4);                           // [2 stage buffers][3 layers][2048]
  float* feat = wimg + (size_t)2 * kV4StageFloats;                          // [32][kV4Stride]
  float4* gpart = reinterpret_cast<float4*>(feat + 32 * kV4Stride);         // [kV4Pts]
  uint64_t* bars = reinterpret_cast<uint64_t*>(gpart + kV4Pts);             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* vec = reinterpret_cast<float*>(reinterpret_cast<char*>(bars) + 256);   // [n_layers][32] biases | fc_p 4x32 | fc_out 2x32
  const float* Wb = g_W;

  const int tile0 = blockIdx.x * kV4Pts;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int group = warp >> 2;                                              // tile of this thread
  const int grp = lane >> 3, j4 = lane & 7;
  const float4* __restrict__ planes4 = reinterpret_cast<const float4*>(g_planes);
  const uint32_t plane4 = (uint32_t)(a.R * a.R * 8);              // one plane of one cloud, in float4 units

  if (warp == 0) umma::tmem_alloc(tmem_slot, kV4TmemCols);
  if (threadIdx.x == 32) {
    for (int g = 0; g < kV4Threads / 128; ++g) umma::mbar_init(&bars[g], 1);
    umma::fence_mbar_init();
  }
  // Weight images are streamed one ResNet block (3 layers, 24 KB) at a time through two stage buffers: stage t < n_blocks
  // is forward block t, stage t >= n_blocks is backward (transposed images) block 2 n_blocks - 1 - t.  One thread fetches
  // the next stage with ONE TMA bulk copy (cp.async.bulk -> mbarrier complete_tx) while the current one computes, so a CTA
  // needs 48 KB of images instead of 120 KB and two CTAs share an SM -- one's gathers (L2-bound) run under the other's MLP
  // chain (latency-bound).
  uint64_t* wbar = bars + 2;                                                 // [2]: bytes of stage buffer t & 1 have landed
  const int n_stages = 2 * a.n_blocks;
  auto prefetch_stage = [&](int t) {                                         // thread 0 only
    const int blk = t < a.n_blocks ? t : n_stages - 1 - t;
    const float* src = g_Wimg + ((size_t)(t < a.n_blocks ? 0 : n_layers) + 3 * blk) * kV3ImgFloats;
    umma::mbar_arrive_expect_tx(&wbar[t & 1], kV4StageFloats * 4);
    umma::bulk_g2s(wimg + (size_t)(t & 1) * kV4StageFloats, src, kV4StageFloats * 4, &wbar[t & 1]);
  };
  // every thread: everybody is done with stage t - 1 (whose buffer stage t + 1 takes), the bytes of stage t have landed
  auto begin_stage = [&](int t) {
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    umma::mbar_wait(&wbar[t & 1], (uint32_t)((t >> 1) & 1));
    if (threadIdx.x == 0 && t + 1 < n_stages) prefetch_stage(t + 1);
  };
  if (threadIdx.x == 0) {
    umma::mbar_init(&wbar[0], 1);
    umma::mbar_init(&wbar[1], 1);
    umma::fence_mbar_init();
    prefetch_stage(0);
  }
  for (int i = threadIdx.x; i < (n_layers + 6) * 32; i += kV4Threads) {
    const int row = i >> 5, c = i & 31;
    float v;
    if (row < n_layers) v = Wb[L::kBlk0 + row * L::kLayer + 1024 + c];
    else if (row < n_layers + 4) v = Wb[(row - n_layers) * 32 + c];          // fc_p W^T rows 0..2, then fc_p.b
    else if (row == n_layers + 4) v = Wb[L::out_w(a.n_blocks) + c];
    else v = c == 0 ? Wb[L::out_b(a.n_blocks)] : 0.0f;
    vec[i] = v;
  }
  // ---------------- own point (slot = thread) and its geometry
  const int slot = threadIdx.x;
  const int pi = min(tile0 + slot, a.n - 1);
  const float p0 = g_xyz[(size_t)pi * 3 + 0], p1 = g_xyz[(size_t)pi * 3 + 1], p2 = g_xyz[(size_t)pi * 3 + 2];
  const V3Geom geo = v3_geom(p0, p1, p2, a.R, a.denom, pi / a.K);
  // ---------------- forward gather: warp w serves tile slots 32w .. 32w+31 (its own threads' points), four points
  //                  per pass, 8 lanes x float4 = one 128-byte texel
#pragma unroll 2
  for (int it = 0; it < 8; ++it) {
    const int src = it * 4 + grp, gslot = warp * 32 + src;
    int pk[3];
    float fr[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      pk[ax] = __shfl_sync(0xffffffffu, geo.pk[ax], src);
      fr[ax] = __shfl_sync(0xffffffffu, geo.f[ax], src);
    }
    const int b = __shfl_sync(0xffffffffu, geo.b, src);
    V3Taps ts;
    v3_taps(pk, fr, a.R, (uint32_t)b * plane4 + (uint32_t)j4, (uint32_t)a.B * plane4, ts);
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4 v = __ldg(planes4 + ts.off[pl][t]);
        s.x = fmaf(v.x, ts.w[pl][t], s.x);
        s.y = fmaf(v.y, ts.w[pl][t], s.y);
        s.z = fmaf(v.z, ts.w[pl][t], s.z);
        s.w = fmaf(v.w, ts.w[pl][t], s.w);
      }
      c.x += s.x; c.y += s.y; c.z += s.z; c.w += s.w;
    }
    feat[(j4 * 4 + 0) * kV4Stride + gslot] = c.x;
    feat[(j4 * 4 + 1) * kV4Stride + gslot] = c.y;
    feat[(j4 * 4 + 2) * kV4Stride + gslot] = c.z;
    feat[(j4 * 4 + 3) * kV4Stride + gslot] = c.w;
  }
  begin_stage(0);                      // (also publishes feat, vec, the TMEM slot and the mbarriers)
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tile_taddr = tmem_base + group * kV3TileCols;                         // lane 0 of the tile
  const uint32_t lane_taddr = tile_taddr + ((uint32_t)((warp & 3) * 32) << 16);        // this warp's lane quarter
  const uint32_t wimg_saddr = umma::smem_u32(wimg);
  uint64_t* bar = &bars[group];
  uint32_t parity = 0;
  const bool leader = (threadIdx.x & 127) == 0;

  // ---------------- MLP forward: thread = point
  float net[32], x[32];
  uint32_t d[32];
  uint32_t mask_a[kMaxBlocks], mask_h[kMaxBlocks];
  const float* fcp = vec + n_layers * 32;
#pragma unroll
  for (int o = 0; o < 32; ++o) {
    float v = fcp[3 * 32 + o];
    v = fmaf(fcp[0 * 32 + o], p0, v);
    v = fmaf(fcp[1 * 32 + o], p1, v);
    v = fmaf(fcp[2 * 32 + o], p2, v);
    net[o] = v;
  }
#pragma unroll 1
  for (int blk = 0; blk < a.n_blocks; ++blk) {
    if (blk > 0) begin_stage(blk);
    const uint32_t img = wimg_saddr + (uint32_t)(blk & 1) * (kV4StageFloats * 4);
    const float* bc = vec + (3 * blk + 0) * 32;
    const float* b0 = vec + (3 * blk + 1) * 32;
    const float* b1 = vec + (3 * blk + 2) * 32;
#pragma unroll
    for (int k = 0; k < 32; ++k) x[k] = feat[k * kV4Stride + slot];
    v3_layer(x, d, tile_taddr, lane_taddr, img + 0 * kV3ImgFloats * 4, bar, parity, group, leader);
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      net[k] += __uint_as_float(d[k]) + bc[k];               // net = net + fc_c(c)
      m |= (net[k] > 0.0f ? 1u : 0u) << k;
      x[k] = fmaxf(net[k], 0.0f);
    }
    mask_a[blk] = m;
    v3_layer(x, d, tile_taddr, lane_taddr, img + 1 * kV3ImgFloats * 4, bar, parity, group, leader);
    m = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float h = __uint_as_float(d[k]) + b0[k];         // h = fc_0(relu(net))
      m |= (h > 0.0f ? 1u : 0u) << k;
      x[k] = fmaxf(h, 0.0f);
    }
    mask_h[blk] = m;
    v3_layer(x, d, tile_taddr, lane_taddr, img + 2 * kV3ImgFloats * 4, bar, parity, group, leader);
#pragma unroll
    for (int k = 0; k < 32; ++k) net[k] += __uint_as_float(d[k]) + b1[k];   // net = net + fc_1(relu(h))
  }
  const float* wo = vec + (n_layers + 4) * 32;
  float logit = vec[(n_layers + 5) * 32];
  uint32_t mask_f = 0;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    mask_f |= (net[k] > 0.0f ? 1u : 0u) << k;
    logit = fmaf(wo[k], fmaxf(net[k], 0.0f), logit);
  }
  const float sg = sigmoidf_(logit);
  const float glogit = (sg - a.target) * a.ginv;
  if (a.stat_part) {
    __shared__ double red[2][kV4Threads / 32];
    const bool live = tile0 + slot < a.n;
    double s0 = live ? (double)bce_with_logits(logit, a.target) : 0.0;
    double s1 = live ? (double)sg : 0.0;
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
      for (int w = 0; w < kV4Threads / 32; ++w) {
        t0 += red[0][w];
        t1 += red[1][w];
      }
      a.stat_part[blockIdx.x * 2 + 0] = t0;
      a.stat_part[blockIdx.x * 2 + 1] = t1;
    }
  }

  // ---------------- MLP backward (dgrad)
  float (&gnet)[32] = net;
#pragma unroll
  for (int k = 0; k < 32; ++k) gnet[k] = ((mask_f >> k) & 1u) ? glogit * wo[k] : 0.0f;
#pragma unroll 1
  for (int blk = a.n_blocks - 1; blk >= 0; --blk) {
    const int stage = n_stages - 1 - blk;
    begin_stage(stage);
    const uint32_t img = wimg_saddr + (uint32_t)(stage & 1) * (kV4StageFloats * 4);
    v3_layer(gnet, d, tile_taddr, lane_taddr, img + 2 * kV3ImgFloats * 4, bar, parity, group, leader);
    const uint32_t mh = mask_h[blk], ma = mask_a[blk];
#pragma unroll
    for (int k = 0; k < 32; ++k) x[k] = ((mh >> k) & 1u) ? __uint_as_float(d[k]) : 0.0f;      // gh
    v3_layer(x, d, tile_taddr, lane_taddr, img + 1 * kV3ImgFloats * 4, bar, parity, group, leader);
#pragma unroll
    for (int k = 0; k < 32; ++k) gnet[k] += ((ma >> k) & 1u) ? __uint_as_float(d[k]) : 0.0f;
    v3_layer(gnet, d, tile_taddr, lane_taddr, img + 0 * kV3ImgFloats * 4, bar, parity, group, leader);
    const bool first = blk == a.n_blocks - 1;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      float* f = feat + k * kV4Stride + slot;
      *f = first ? __uint_as_float(d[k]) : *f + __uint_as_float(d[k]);
    }
  }
  {
    float g[3];
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
      float s = 0.0f;
#pragma unroll
      for (int o = 0; o < 32; ++o) s = fmaf(fcp[dd * 32 + o], gnet[o], s);
      g[dd] = s;
    }
    gpart[slot] = make_float4(g[0], g[1], g[2], 0.f);
  }
  __syncwarp();

  // ---------------- backward gather (warp-local: slots 32w .. 32w+31); geometry recomputed once per point (its
  //                  registers were handed to the MLP)
  const V3Geom geo2 = v3_geom(p0, p1, p2, a.R, a.denom, pi / a.K);
  const float dsc = ((float)(a.R - 1) * 0.5f) * 2.0f / a.denom;        // Axis::dscale of a live, unclipped axis
#pragma unroll 2
  for (int it = 0; it < 8; ++it) {
    const int src = it * 4 + grp, gslot = warp * 32 + src;
    const int pi_raw = tile0 + gslot;
    int pk[3];
    float fr[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      pk[ax] = __shfl_sync(0xffffffffu, geo2.pk[ax], src);
      fr[ax] = __shfl_sync(0xffffffffu, geo2.f[ax], src);
    }
    const int b = __shfl_sync(0xffffffffu, geo2.b, src);
    V3Taps ts;
    v3_taps(pk, fr, a.R, (uint32_t)b * plane4 + (uint32_t)j4, (uint32_t)a.B * plane4, ts);
    const float4 gc = make_float4(feat[(j4 * 4 + 0) * kV4Stride + gslot], feat[(j4 * 4 + 1) * kV4Stride + gslot],
                                  feat[(j4 * 4 + 2) * kV4Stride + gslot], feat[(j4 * 4 + 3) * kV4Stride + gslot]);
    // d c / d (ix, iy) is linear in the four texel . g_c dot products, so each lane combines its 4-channel partial
    // dots into partial axis gradients first and only THREE values cross the 8 lanes (instead of twelve)
    float gi[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      float qv[4];
#pragma unroll
      for (int t4 = 0; t4 < 4; ++t4) {
        const float4 v = __ldg(planes4 + ts.off[pl][t4]);
        qv[t4] = (v.x * gc.x + v.y * gc.y) + (v.z * gc.z + v.w * gc.w);
      }
      const int aw = plane_axis_w(pl), ah = plane_axis_h(pl);
      const float q_ne = ts.has1[aw] ? qv[1] : 0.0f;
      const float q_sw = ts.has1[ah] ? qv[2] : 0.0f;
      const float q_se = (ts.has1[aw] && ts.has1[ah]) ? qv[3] : 0.0f;
      gi[aw] += (q_ne - qv[0]) * (1.0f - ts.f[ah]) + (q_se - q_sw) * ts.f[ah];
      gi[ah] += (q_sw - qv[0]) * (1.0f - ts.f[aw]) + (q_se - q_ne) * ts.f[aw];
    }
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      gi[ax] += __shfl_xor_sync(0xffffffffu, gi[ax], 1);
      gi[ax] += __shfl_xor_sync(0xffffffffu, gi[ax], 2);
      gi[ax] += __shfl_xor_sync(0xffffffffu, gi[ax], 4);
    }
    if (j4 == 0 && pi_raw < a.n) {
      const float4 gp = gpart[gslot];
      const size_t o = (size_t)pi_raw * 3;
      g_grad[o + 0] = gp.x + gi[0] * (((pk[0] >> 17) & 1) ? dsc : 0.0f);
      g_grad[o + 1] = gp.y + gi[1] * (((pk[1] >> 17) & 1) ? dsc : 0.0f);
      g_grad[o + 2] = gp.z + gi[2] * (((pk[2] >> 17) & 1) ? dsc : 0.0f);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, kV4TmemCols);
}


}  // namespace ifd
