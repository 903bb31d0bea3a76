// decode_v5.cuh -- the production ConvONet decode: gather + ResNet-MLP on tcgen05 (building blocks: decode_v3.cuh) in 256-point
// CTAs, two per SM, weight images streamed per stage by TMA bulk copies -- and, new against the retired v4 kernel, the fc_c
// layers taken OFF the dependent chain: 22 tensor-core round trips per point and Adam step instead of 30.
//
// LocalDecoder.forward (ConvONet/src/conv_onet/models/decoder.py:82-98) is, per ResNet block i,
//     net = net + fc_c[i](c);  h = fc_0[i](relu(net));  net = net + fc_1[i](relu(h))
// and decode v3 / v4 ran the three Linear layers as three dependent tcgen05 round trips (st A -> barrier -> 12 MMA -> commit
// -> wait -> ld D).  ncu on v4 (profiles/r01_final_*): every pipe under 45 %, the kernel time is the LATENCY of that chain.
// But fc_c[i](c) does not depend on the chain -- c is the gathered plane feature, the same operand for all blocks -- so
//   * forward: the products of fc_c[i + 1] are issued together with those of fc_1[i], into the SAME accumulator
//       D = relu(h) . W1[i]^T + c . Wc[i + 1]^T        (24 MMAs, one commit, one round trip)
//     c_hi stays in TMEM for the whole forward pass (32 columns), c_lo is a K-major image in shared memory (it takes the place
//     of the gather's staging buffer), so c is split and stored ONCE instead of five times;
//   * backward: d loss / d c = sum_i g_i . Wc[i] where g_i is the gradient on the residual stream below block i -- exactly the
//     A operand the chain stores for fc_1[i - 1]^T.  Its products go to a second accumulator (32 TMEM columns, the ones c_hi
//     occupied), which is read and summed in fp32 in the shadow of the next round trip.  (Accumulating over the blocks inside
//     TMEM was measured: every MMA truncates the accumulator, 60 of them at full magnitude raised the median gradient error from
//     2.0e-7 to 2.6e-7 of the largest component and moved coordinates whose gradient is of the order of Adam's eps.)
// Round trips: forward 1 + 2 n_blocks, backward 2 n_blocks + 1.  Everything else (gathers, 3xTF32 split, masks, BCE gradient,
// fc_p, weight stages by TMA bulk copy, two CTAs per SM) is v4's.  The association of the residual sum differs from v4
// ((net + dx) + fc_c  ->  net + (dx + fc_c), formed in the fp32 accumulator), so v5 and v4 agree to rounding, not bitwise.
#pragma once
#include "decode_v3.cuh"

namespace ifd {

constexpr int kV5Threads = 256;                    // two tiles of 128 points
constexpr int kV5Pts = 256;
constexpr int kV5Stride = kV5Pts + 1;
constexpr int kV5StageFloats = 3 * kV3ImgFloats;   // one stage buffer: up to 3 images x (hi + lo) = 24 KB
constexpr uint32_t kV5TmemCols = 256;
constexpr int kV5TileCols = 128;                   // D | A_hi | A_lo | c_hi (forward) / D_gc (backward)
constexpr int kV5ColA = 32, kV5ColX = 96;

// Stage table of the weight images (one stage = what the rounds between two CTA-wide syncs read; <= 3 images of 8 KB):
//   s = 0                      : Wc[0]
//   s = 1 + b, b < nb          : W0[b], W1[b], Wc[b + 1] (b + 1 < nb)
//   s = nb + 1 + j, j < nb     : blk = nb - 1 - j:  W1[blk]^T, W0[blk]^T, Wc[blk + 1]^T (blk + 1 < nb)
//   s = 2 nb + 1               : Wc[0]^T
// Images lie in the workspace in exactly this order, so a stage is one contiguous bulk copy.
__host__ __device__ inline int v5_stage_first(int s, int nb) {
  if (s == 0) return 0;
  if (s <= nb) return 1 + 3 * (s - 1);
  const int j = s - nb - 1;
  if (j == 0) return 3 * nb;
  if (j < nb) return 3 * nb + 2 + 3 * (j - 1);
  return 6 * nb - 1;
}
__host__ __device__ inline int v5_stage_images(int s, int nb) {
  if (s == 0 || s == 2 * nb + 1) return 1;
  if (s == nb || s == nb + 1) return 2;
  return 3;
}
// position (image index) of layer l = 3 blk + {0: fc_c, 1: fc_0, 2: fc_1} in the forward / backward part
__host__ __device__ inline int v5_fwd_pos(int l, int nb) {
  const int blk = l / 3, t = l % 3;
  if (t == 0) return blk == 0 ? 0 : v5_stage_first(blk, nb) + 2;       // Wc[blk] rides with block blk - 1
  return v5_stage_first(1 + blk, nb) + (t - 1);
}
__host__ __device__ inline int v5_bwd_pos(int l, int nb) {
  const int blk = l / 3, t = l % 3;
  if (t == 0) return blk == 0 ? 6 * nb - 1 : v5_stage_first(nb + 1 + (nb - blk), nb) + 2;   // Wc[blk]^T rides with block blk - 1
  return v5_stage_first(nb + 1 + (nb - 1 - blk), nb) + (t == 2 ? 0 : 1);
}

// blob (kernel layout, W^T [in][out] per layer) -> UMMA images hi | lo in stage order.
__global__ void convonet_pack_umma_v5_kernel(const float* __restrict__ Wb_arg, int n_blocks, float* __restrict__ out_arg,
                                             const LoopJob* __restrict__ job) {
  using L = ConvDecLayout<32>;
  const float* __restrict__ Wb = job ? job->W : Wb_arg;
  float* __restrict__ out = job ? const_cast<float*>(job->Wimg) : out_arg;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * n_blocks * 1024) return;
  const int l = e >> 10, i = (e >> 5) & 31, o = e & 31;        // WT[i][o] = W[o][i]
  const float w = Wb[L::kBlk0 + l * L::kLayer + i * 32 + o];
  const float hi = __uint_as_float(umma::tf32_hi(w)), lo = __uint_as_float(umma::tf32_lo(w));
  float* f = out + (size_t)v5_fwd_pos(l, n_blocks) * kV3ImgFloats;      // forward: B[n = o][k = i]
  f[umma::img_offset(o, i, 32) / 4] = hi;
  f[1024 + umma::img_offset(o, i, 32) / 4] = lo;
  float* b = out + (size_t)v5_bwd_pos(l, n_blocks) * kV3ImgFloats;      // backward (dgrad): B[n = i][k = o]
  b[umma::img_offset(i, o, 32) / 4] = hi;
  b[1024 + umma::img_offset(i, o, 32) / 4] = lo;
}

// ---- the pieces of one round trip -----------------------------------------------------------------------------------------
// A := split(x): hi -> TMEM columns 32..63, lo (exact remainder) -> 64..95 of this thread's lane.  `d` is staging.
__device__ __forceinline__ void v5_put_a(const float (&x)[32], uint32_t (&d)[32], uint32_t lane_taddr) {
#pragma unroll
  for (int k = 0; k < 32; ++k) d[k] = umma::tf32_hi_cvt(x[k]);
  umma::tmem_st32(lane_taddr + kV5ColA, d);
#pragma unroll
  for (int k = 0; k < 32; ++k) d[k] = __float_as_uint(x[k] - __uint_as_float(d[k]));
  umma::tmem_st32(lane_taddr + kV5ColA + 32, d);
  umma::tmem_wait_st();
}
// Leader only.  The tensor core truncates the accumulator at every MMA, so what costs accuracy is the number of accumulations
// at FULL magnitude: the two small cross terms of every product (2^-11 of the magnitude) are issued first, while the
// accumulator is still small, and only the hi.hi MMAs (four per product) land on the full value -- as many as in v4.
//   D[d_col] (+)= A_lo . B_hi + A_hi . B_lo          A in TMEM columns 32..95
__device__ __forceinline__ void v5_chain_small(uint32_t tile_taddr, uint32_t img_saddr, uint32_t d_col, uint32_t acc0) {
  constexpr uint32_t idesc = umma::idesc_tf32(128, 32);
  const uint32_t d_t = tile_taddr + d_col, a_hi = tile_taddr + kV5ColA, a_lo = a_hi + 32;
#pragma unroll
  for (int part = 0; part < 2; ++part) {
    const uint32_t a_t = part == 0 ? a_lo : a_hi;
    const uint32_t b_s = img_saddr + (part == 1 ? 4096u : 0u);
#pragma unroll
    for (int s = 0; s < 4; ++s)
      umma::mma_tf32_ts(d_t, a_t + s * 8, umma::smem_desc_kmajor(b_s + s * 1024, 512, 128), idesc, (part | s) ? 1u : acc0);
  }
}
//   D[d_col] += A_hi . B_hi
__device__ __forceinline__ void v5_chain_big(uint32_t tile_taddr, uint32_t img_saddr, uint32_t d_col) {
  constexpr uint32_t idesc = umma::idesc_tf32(128, 32);
  const uint32_t d_t = tile_taddr + d_col, a_hi = tile_taddr + kV5ColA;
#pragma unroll
  for (int s = 0; s < 4; ++s)
    umma::mma_tf32_ts(d_t, a_hi + s * 8, umma::smem_desc_kmajor(img_saddr + s * 1024, 512, 128), idesc, 1u);
}
//   D[0] (+)= c_lo . B_hi + c_hi . B_lo      c_hi in TMEM columns 96..127, c_lo a [128 x 32] K-major image in shared memory
__device__ __forceinline__ void v5_c_small(uint32_t tile_taddr, uint32_t clo_saddr, uint32_t img_saddr, uint32_t acc0) {
  constexpr uint32_t idesc = umma::idesc_tf32(128, 32);
  const uint32_t c_hi = tile_taddr + kV5ColX;
#pragma unroll
  for (int s = 0; s < 4; ++s)
    umma::mma_tf32_ss(tile_taddr, umma::smem_desc_kmajor(clo_saddr + s * 4096, 2048, 128),
                      umma::smem_desc_kmajor(img_saddr + s * 1024, 512, 128), idesc, s ? 1u : acc0);
#pragma unroll
  for (int s = 0; s < 4; ++s)
    umma::mma_tf32_ts(tile_taddr, c_hi + s * 8, umma::smem_desc_kmajor(img_saddr + 4096u + s * 1024, 512, 128), idesc, 1u);
}
//   D[0] += c_hi . B_hi
__device__ __forceinline__ void v5_c_big(uint32_t tile_taddr, uint32_t img_saddr) {
  constexpr uint32_t idesc = umma::idesc_tf32(128, 32);
  const uint32_t c_hi = tile_taddr + kV5ColX;
#pragma unroll
  for (int s = 0; s < 4; ++s)
    umma::mma_tf32_ts(tile_taddr, c_hi + s * 8, umma::smem_desc_kmajor(img_saddr + s * 1024, 512, 128), idesc, 1u);
}
// Sign bits of 32 values as a mask (bit 31 - k = sign of v[k]): four independent funnel-shift chains of eight, then one merge.
struct V5Signs {
  uint32_t q[4];
  __device__ __forceinline__ void clear() { q[0] = q[1] = q[2] = q[3] = 0u; }
  __device__ __forceinline__ void push(int k, float v) { q[k >> 3] = __funnelshift_l(__float_as_uint(v), q[k >> 3], 1); }
  __device__ __forceinline__ uint32_t mask() const { return (q[0] << 24) | (q[1] << 16) | (q[2] << 8) | q[3]; }
};
// every thread of the 128-thread group, around the leader's MMAs
__device__ __forceinline__ void v5_round_begin(int group) {
  umma::fence_before_sync();
  asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
}
__device__ __forceinline__ void v5_round_end(uint32_t (&d)[32], uint32_t taddr, uint64_t* bar, uint32_t& parity) {
  __syncwarp();
  umma::mbar_wait(bar, parity);
  parity ^= 1;
  umma::fence_after_sync();
  umma::tmem_ld32(taddr, d);
}

struct DecodeV5Smem {
  static __host__ __device__ size_t bytes(int n_blocks) {      // stage buffers | feat / c_lo | gpart | barriers | bias / fc_p / fc_out table
    return (size_t)2 * kV5StageFloats * 4 + (size_t)32 * kV5Stride * 4 + (size_t)kV5Pts * 16 + 256 +
           (size_t)(3 * n_blocks + 6) * 32 * 4;
  }
};

__global__ void __launch_bounds__(kV5Threads, 2) convonet_decode_v5_kernel(const DecodeV3Args a) {
  extern __shared__ float4 smem4[];
  const float* __restrict__ g_planes = a.job ? a.job->planes : a.planes;
  const float* __restrict__ g_W = a.job ? a.job->W : a.W;
  const float* __restrict__ g_Wimg = a.job ? a.job->Wimg : a.Wimg;
  const float* __restrict__ g_xyz = a.job ? a.job->xyz : a.xyz;
  float* __restrict__ g_grad = a.job ? a.job->g_occ : a.grad_out;
  float4* __restrict__ g_jac4 = reinterpret_cast<float4*>(a.job ? a.job->jac : a.jac);
  using L = ConvDecLayout<32>;
  const int nb = a.n_blocks;
  const int n_layers = 3 * nb;
  float* wimg = reinterpret_cast<float*>(smem4);                           // [2 stage buffers][3 images][2048]
  float* feat = wimg + (size_t)2 * kV5StageFloats;                          // [32][kV5Stride]; forward MLP: c_lo images [2 tiles][128 x 32]
  float4* gpart = reinterpret_cast<float4*>(feat + 32 * kV5Stride);         // [kV5Pts]
  uint64_t* bars = reinterpret_cast<uint64_t*>(gpart + kV5Pts);             // [2] tiles, [2] weight stages
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* vec = reinterpret_cast<float*>(reinterpret_cast<char*>(bars) + 256);   // [n_layers][32] biases | fc_p 4x32 | fc_out 2x32
  const float* Wb = g_W;

  const int tile0 = blockIdx.x * kV5Pts;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int group = warp >> 2;                                              // tile of this thread
  const int grp = lane >> 3, j4 = lane & 7;
  const float4* __restrict__ planes4 = reinterpret_cast<const float4*>(g_planes);
  const uint32_t plane4 = (uint32_t)(a.R * a.R * 8);              // one plane of one cloud, in float4 units

  if (warp == 0) umma::tmem_alloc(tmem_slot, kV5TmemCols);
  if (threadIdx.x == 32) {
    for (int g = 0; g < kV5Threads / 128; ++g) umma::mbar_init(&bars[g], 1);
    umma::fence_mbar_init();
  }
  uint64_t* wbar = bars + 2;                                                 // [2]: bytes of stage buffer t & 1 have landed
  const int n_stages = 2 * nb + 2;
  auto prefetch_stage = [&](int t) {                                         // thread 0 only: ONE TMA bulk copy per stage
    const uint32_t bytes = (uint32_t)v5_stage_images(t, nb) * (kV3ImgFloats * 4);
    umma::mbar_arrive_expect_tx(&wbar[t & 1], bytes);
    umma::bulk_g2s(wimg + (size_t)(t & 1) * kV5StageFloats, g_Wimg + (size_t)v5_stage_first(t, nb) * kV3ImgFloats, bytes, &wbar[t & 1]);
  };
  // Weight stages without CTA-wide barriers: the two tiles of a CTA run their chains independently (one in its ALU epilogue
  // while the other waits for the tensor core).  Only the lane that issues a tile's MMAs touches the weights: it waits for the
  // bytes of stage t before its first MMA (wait_weights), and after the tile's last round of stage t it releases buffer t & 1
  // (release_stage) -- the SECOND tile to release it starts the copy of stage t + 2 into it.  A tile can therefore be at most
  // one stage ahead of the other, which is also what keeps the shared staging buffer (feat / c_lo) safe.
  uint32_t* rel = tmem_slot + 1;                                             // [2] tiles that released stage buffer t & 1
  auto wait_weights = [&](int t) { umma::mbar_wait(&wbar[t & 1], (uint32_t)((t >> 1) & 1)); };
  auto release_stage = [&](int t) {                                          // issuing lane of a tile, after the tile's last round of stage t
    if (t + 2 >= n_stages) return;
    __threadfence_block();
    if (atomicAdd(&rel[t & 1], 1u) == (uint32_t)(kV5Threads / 128 - 1)) {
      atomicExch(&rel[t & 1], 0u);
      __threadfence_block();
      prefetch_stage(t + 2);
    }
  };
  if (threadIdx.x == 0) {
    umma::mbar_init(&wbar[0], 1);
    umma::mbar_init(&wbar[1], 1);
    umma::fence_mbar_init();
    rel[0] = rel[1] = 0u;
    prefetch_stage(0);
    prefetch_stage(1);
  }
  for (int i = threadIdx.x; i < (n_layers + 6) * 32; i += kV5Threads) {
    const int row = i >> 5, c = i & 31;
    float v;
    if (row < n_layers) v = Wb[L::kBlk0 + row * L::kLayer + 1024 + c];
    else if (row < n_layers + 4) v = Wb[(row - n_layers) * 32 + c];          // fc_p W^T rows 0..2, then fc_p.b
    else if (row == n_layers + 4) v = Wb[L::out_w(nb) + c];
    else v = c == 0 ? Wb[L::out_b(nb)] : 0.0f;
    if (row < n_layers && v == 0.0f) v = -0.0f;     // a zero bias is kept as -0: see the sign-bit masks of the forward pass
    vec[i] = v;
  }
  // ---------------- own point (slot = thread) and its geometry
  const int slot = threadIdx.x;
  const int pi = min(tile0 + slot, a.n - 1);
  const float p0 = g_xyz[(size_t)pi * 3 + 0], p1 = g_xyz[(size_t)pi * 3 + 1], p2 = g_xyz[(size_t)pi * 3 + 2];
  const V3Geom geo = v3_geom(p0, p1, p2, a.R, a.denom, pi / a.K);
  // ---------------- forward gather  warp w serves tile slots 32w .. 32w+31, four points per pass
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
    // d c / d (grid coordinate of axis ax), this lane's four channels: the bilinear form differentiated tap by tap (the
    // expressions of the backward gather below, applied to the texel vectors instead of to their dot products with g_c).
    // The SAME thread reads it back after the MLP: 384 B per point through L2 instead of a second 1536 B texel gather.
    float4 jx[3] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      float4 v[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) v[t] = __ldg(planes4 + ts.off[pl][t]);
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        s.x = fmaf(v[t].x, ts.w[pl][t], s.x);
        s.y = fmaf(v[t].y, ts.w[pl][t], s.y);
        s.z = fmaf(v[t].z, ts.w[pl][t], s.z);
        s.w = fmaf(v[t].w, ts.w[pl][t], s.w);
      }
      c.x += s.x; c.y += s.y; c.z += s.z; c.w += s.w;
      if (g_jac4) {
        // d/dw and d/dh of the bilinear form as tap coefficients (a clamped neighbour contributes nothing: has1 == 0)
        const int aw = plane_axis_w(pl), ah = plane_axis_h(pl);
        const float mw = ts.has1[aw] ? 1.0f : 0.0f, mh = ts.has1[ah] ? 1.0f : 0.0f;
        const float fw = ts.f[aw], fh = ts.f[ah], nfw = 1.0f - fw, nfh = 1.0f - fh;
        const float cw[4] = {-nfh, mw * nfh, -(mh * fh), mw * mh * fh};       // nw, ne, sw, se
        const float ch[4] = {-nfw, -(mw * fw), mh * nfw, mw * mh * fw};
        float4& jw = jx[aw];
        float4& jh = jx[ah];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          jw.x = fmaf(v[t].x, cw[t], jw.x); jw.y = fmaf(v[t].y, cw[t], jw.y); jw.z = fmaf(v[t].z, cw[t], jw.z); jw.w = fmaf(v[t].w, cw[t], jw.w);
          jh.x = fmaf(v[t].x, ch[t], jh.x); jh.y = fmaf(v[t].y, ch[t], jh.y); jh.z = fmaf(v[t].z, ch[t], jh.z); jh.w = fmaf(v[t].w, ch[t], jh.w);
        }
      }
    }
    if (g_jac4 && tile0 + gslot < a.n) {
      float4* dst = g_jac4 + ((size_t)(tile0 + gslot) * 3) * 8 + j4;
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        const bool on = (pk[ax] >> 17) & 1;                      // clamped / dead axis: no gradient
        __stcg(dst + ax * 8, on ? jx[ax] : make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
    feat[(j4 * 4 + 0) * kV5Stride + gslot] = c.x;
    feat[(j4 * 4 + 1) * kV5Stride + gslot] = c.y;
    feat[(j4 * 4 + 2) * kV5Stride + gslot] = c.z;
    feat[(j4 * 4 + 3) * kV5Stride + gslot] = c.w;
  }
  umma::fence_before_sync();
  __syncthreads();                     // publishes feat, vec, the TMEM slot, the mbarriers and the release counters
  umma::fence_after_sync();
  // warp-uniform copies of everything the MMA issue reads (umma::elect_one: the operands must sit in uniform registers)
  const int warp_u = (int)umma::warp_bcast((uint32_t)warp), group_u = warp_u >> 2;
  const uint32_t tmem_base = umma::warp_bcast(*tmem_slot);
  const uint32_t tile_taddr = tmem_base + group_u * kV5TileCols;                       // lane 0 of the tile
  const uint32_t lane_taddr = tile_taddr + ((uint32_t)((warp & 3) * 32) << 16);        // this warp's lane quarter
  const uint32_t wimg_saddr = umma::smem_u32(wimg);
  unsigned char* clo = reinterpret_cast<unsigned char*>(feat) + (size_t)group_u * (128 * 32 * 4);   // c_lo image of this tile
  const uint32_t clo_saddr = umma::smem_u32(clo);
  uint64_t* bar = &bars[group_u];
  uint32_t parity = 0;
  const bool lead_warp = (warp_u & 3) == 0;                                            // first warp of the tile issues its MMAs
  auto stage_img = [&](int t, int i) { return wimg_saddr + (uint32_t)(t & 1) * (kV5StageFloats * 4) + (uint32_t)i * (kV3ImgFloats * 4); };

  // ---------------- MLP forward: thread = point
  float net[32], x[32];
  uint32_t d[32];
  uint32_t mask_a[kMaxBlocks], mask_h[kMaxBlocks];
  const float* fcp = vec + n_layers * 32;
  // c: hi -> TMEM (kept for the whole forward pass), lo -> K-major image over the gather's staging buffer
#pragma unroll
  for (int k = 0; k < 32; ++k) x[k] = feat[k * kV5Stride + slot];
  __syncthreads();                     // every thread holds its c row: the staging buffer may be overwritten
#pragma unroll
  for (int k = 0; k < 32; ++k) d[k] = umma::tf32_hi_cvt(x[k]);
  umma::tmem_st32(lane_taddr + kV5ColX, d);
  {
    const int r = slot & 127;
    unsigned char* row = clo + (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      uint4 l;
      l.x = __float_as_uint(x[4 * g + 0] - __uint_as_float(d[4 * g + 0])); l.y = __float_as_uint(x[4 * g + 1] - __uint_as_float(d[4 * g + 1]));
      l.z = __float_as_uint(x[4 * g + 2] - __uint_as_float(d[4 * g + 2])); l.w = __float_as_uint(x[4 * g + 3] - __uint_as_float(d[4 * g + 3]));
      *reinterpret_cast<uint4*>(row + g * 2048) = l;
    }
  }
  umma::fence_proxy_async();           // generic-proxy stores -> visible to the tensor core's async proxy
  umma::tmem_wait_st();
  v5_round_begin(group);
  if (lead_warp && umma::elect_one()) {
    wait_weights(0);
    umma::fence_after_sync();
    v5_c_small(tile_taddr, clo_saddr, stage_img(0, 0), 0u);
    v5_c_big(tile_taddr, stage_img(0, 0));
    umma::commit(bar);
  }
  v5_round_end(d, lane_taddr, bar, parity);
  if (lead_warp && umma::elect_one()) release_stage(0);
  // The residual stream and the hidden pre-activations are carried NEGATED (nnet = -net, exact: rounding is symmetric), because
  // then "pre-activation > 0" is the sign bit of the stored value and a 32-bit ReLU mask costs one funnel shift per element
  // instead of compare + select + or (ncu: 470 instructions per thread and round trip, 80 of them for the masks).  Zeros: every
  // sum below is formed as (-a) + (-b) with zero biases stored as -0, so a true zero comes out as +0 (sign clear = "not > 0",
  // the reference's relu'(0) = 0), never as -0.  Bit 31 - k of a mask belongs to element k.
  {
    const float* bc = vec;
#pragma unroll
    for (int o = 0; o < 32; ++o) {
      float v = fcp[3 * 32 + o];
      v = fmaf(fcp[0 * 32 + o], p0, v);
      v = fmaf(fcp[1 * 32 + o], p1, v);
      v = fmaf(fcp[2 * 32 + o], p2, v);
      net[o] = __fadd_rn(-v, __fadd_rn(-__uint_as_float(d[o]), -bc[o]));            // -(fc_p(p) + fc_c[0](c))
    }
  }
#pragma unroll 1
  for (int blk = 0; blk < nb; ++blk) {
    const int st = 1 + blk;
    const float* b0 = vec + (3 * blk + 1) * 32;
    const float* b1 = vec + (3 * blk + 2) * 32;
    V5Signs sg4;
    sg4.clear();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      sg4.push(k, net[k]);
      x[k] = fmaxf(-net[k], 0.0f);
    }
    mask_a[blk] = sg4.mask();
    v5_put_a(x, d, lane_taddr);
    v5_round_begin(group);
    if (lead_warp && umma::elect_one()) {
      wait_weights(st);
      umma::fence_after_sync();
      v5_chain_small(tile_taddr, stage_img(st, 0), 0u, 0u);
      v5_chain_big(tile_taddr, stage_img(st, 0), 0u);
      umma::commit(bar);
    }
    v5_round_end(d, lane_taddr, bar, parity);
    sg4.clear();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float nh = __fadd_rn(-__uint_as_float(d[k]), -b0[k]);         // -h,  h = fc_0(relu(net))
      sg4.push(k, nh);
      x[k] = fmaxf(-nh, 0.0f);
    }
    mask_h[blk] = sg4.mask();
    v5_put_a(x, d, lane_taddr);
    v5_round_begin(group);
    const bool more = blk + 1 < nb;
    if (lead_warp && umma::elect_one()) {
      umma::fence_after_sync();
      v5_chain_small(tile_taddr, stage_img(st, 1), 0u, 0u);
      if (more) {
        v5_c_small(tile_taddr, clo_saddr, stage_img(st, 2), 1u);
        v5_c_big(tile_taddr, stage_img(st, 2));
      }
      v5_chain_big(tile_taddr, stage_img(st, 1), 0u);
      umma::commit(bar);
    }
    v5_round_end(d, lane_taddr, bar, parity);
    if (lead_warp && umma::elect_one()) release_stage(st);
    if (more) {
      const float* bc = vec + (3 * blk + 3) * 32;
#pragma unroll
      for (int k = 0; k < 32; ++k)      // net + fc_1(relu(h)) + fc_c[blk + 1](c), negated
        net[k] = __fadd_rn(__fadd_rn(net[k], __fadd_rn(-__uint_as_float(d[k]), -b1[k])), -bc[k]);
    } else {
#pragma unroll
      for (int k = 0; k < 32; ++k) net[k] = __fadd_rn(net[k], __fadd_rn(-__uint_as_float(d[k]), -b1[k]));
    }
  }
  const float* wo = vec + (n_layers + 4) * 32;
  float logit = vec[(n_layers + 5) * 32];
  V5Signs sgf;
  sgf.clear();
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    sgf.push(k, net[k]);
    logit = fmaf(wo[k], fmaxf(-net[k], 0.0f), logit);
  }
  const uint32_t mask_f = sgf.mask();
  const float sg = sigmoidf_(logit);
  const float glogit = (sg - a.target) * a.ginv;
  if (a.stat_part) {
    __shared__ double red[2][kV5Threads / 32];
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
      for (int w = 0; w < kV5Threads / 32; ++w) {
        t0 += red[0][w];
        t1 += red[1][w];
      }
      a.stat_part[blockIdx.x * 2 + 0] = t0;
      a.stat_part[blockIdx.x * 2 + 1] = t1;
    }
  }

  // ---------------- MLP backward (dgrad).  d loss / d c = sum_blk g_blk . Wc[blk]: the product of block blk + 1 is issued with
  //                  fc_1[blk]^T (same A operand) into columns 96..127 and added to `feat` in the shadow of the next round trip.
  float (&gnet)[32] = net;
#pragma unroll
  for (int k = 0; k < 32; ++k) gnet[k] = ((mask_f >> (31 - k)) & 1u) ? glogit * wo[k] : 0.0f;
#pragma unroll 1
  for (int blk = nb - 1; blk >= 0; --blk) {
    const int st = nb + 1 + (nb - 1 - blk);
    const bool with_c = blk + 1 < nb;
    v5_put_a(gnet, d, lane_taddr);
    v5_round_begin(group);
    if (lead_warp && umma::elect_one()) {
      wait_weights(st);
      umma::fence_after_sync();
      v5_chain_small(tile_taddr, stage_img(st, 0), 0u, 0u);                                      // . W1[blk]
      v5_chain_big(tile_taddr, stage_img(st, 0), 0u);
      if (with_c) {                                                                               // . Wc[blk + 1]
        v5_chain_small(tile_taddr, stage_img(st, 2), kV5ColX, 0u);
        v5_chain_big(tile_taddr, stage_img(st, 2), kV5ColX);
      }
      umma::commit(bar);
    }
    v5_round_end(d, lane_taddr, bar, parity);
    const uint32_t mh = mask_h[blk], ma = mask_a[blk];
#pragma unroll
    for (int k = 0; k < 32; ++k) x[k] = ((mh >> (31 - k)) & 1u) ? __uint_as_float(d[k]) : 0.0f;      // gh
    v5_put_a(x, d, lane_taddr);
    v5_round_begin(group);
    if (lead_warp && umma::elect_one()) {
      umma::fence_after_sync();
      v5_chain_small(tile_taddr, stage_img(st, 1), 0u, 0u);                                      // . W0[blk]
      v5_chain_big(tile_taddr, stage_img(st, 1), 0u);
      umma::commit(bar);
    }
    if (with_c) {                      // while those MMAs run: this block's share of d loss / d c
      umma::tmem_ld32(lane_taddr + kV5ColX, d);
      const bool first = blk + 2 == nb;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float* f = feat + k * kV5Stride + slot;
        *f = first ? __uint_as_float(d[k]) : *f + __uint_as_float(d[k]);
      }
    }
    v5_round_end(d, lane_taddr, bar, parity);
    if (lead_warp && umma::elect_one()) release_stage(st);
#pragma unroll
    for (int k = 0; k < 32; ++k) gnet[k] += ((ma >> (31 - k)) & 1u) ? __uint_as_float(d[k]) : 0.0f;
  }
  {
    const int st = 2 * nb + 1;
    v5_put_a(gnet, d, lane_taddr);
    v5_round_begin(group);
    if (lead_warp && umma::elect_one()) {
      wait_weights(st);
      umma::fence_after_sync();
      v5_chain_small(tile_taddr, stage_img(st, 0), kV5ColX, 0u);                                 // . Wc[0]
      v5_chain_big(tile_taddr, stage_img(st, 0), kV5ColX);
      umma::commit(bar);
    }
    v5_round_end(d, lane_taddr + kV5ColX, bar, parity);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      float* f = feat + k * kV5Stride + slot;
      *f = nb == 1 ? __uint_as_float(d[k]) : *f + __uint_as_float(d[k]);
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

  const float dsc = ((float)(a.R - 1) * 0.5f) * 2.0f / a.denom;        // Axis::dscale of a live, unclipped axis
  if (g_jac4) {
    // ---------------- backward "gather" from the Jacobian this thread stored in the forward pass (warp-local slots).  Fully
    //                  unrolled: the MLP's registers are free here, so all 24 loads of a thread are in flight at once -- one L2
    //                  round trip instead of four (this phase was 10 % of the kernel's warp time at unroll 2)
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int src = it * 4 + grp, gslot = warp * 32 + src;
      const int pi_raw = tile0 + gslot;
      const bool in = pi_raw < a.n;
      const float4 gc = make_float4(feat[(j4 * 4 + 0) * kV5Stride + gslot], feat[(j4 * 4 + 1) * kV5Stride + gslot],
                                    feat[(j4 * 4 + 2) * kV5Stride + gslot], feat[(j4 * 4 + 3) * kV5Stride + gslot]);
      const float4* srcj = g_jac4 + ((size_t)(in ? pi_raw : 0) * 3) * 8 + j4;
      float gi[3];
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        const float4 j = in ? __ldcg(srcj + ax * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
        gi[ax] = (j.x * gc.x + j.y * gc.y) + (j.z * gc.z + j.w * gc.w);
      }
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        gi[ax] += __shfl_xor_sync(0xffffffffu, gi[ax], 1);
        gi[ax] += __shfl_xor_sync(0xffffffffu, gi[ax], 2);
        gi[ax] += __shfl_xor_sync(0xffffffffu, gi[ax], 4);
      }
      if (j4 == 0 && in) {
        const float4 gp = gpart[gslot];
        const size_t o = (size_t)pi_raw * 3;
        g_grad[o + 0] = gp.x + gi[0] * dsc;
        g_grad[o + 1] = gp.y + gi[1] * dsc;
        g_grad[o + 2] = gp.z + gi[2] * dsc;
      }
    }
  } else {
  // ---------------- backward gather (warp-local: slots 32w .. 32w+31)
  const V3Geom geo2 = v3_geom(p0, p1, p2, a.R, a.denom, pi / a.K);
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
    const float4 gc = make_float4(feat[(j4 * 4 + 0) * kV5Stride + gslot], feat[(j4 * 4 + 1) * kV5Stride + gslot],
                                  feat[(j4 * 4 + 2) * kV5Stride + gslot], feat[(j4 * 4 + 3) * kV5Stride + gslot]);
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
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, kV5TmemCols);
}

}  // namespace ifd
