// decode_v3.cuh -- ConvONet decode with the ResNet-MLP on the 5th-generation tensor cores.
//
// Why: ncu on decode v2 (profiles/r01_v2_*) shows the SIMT formulation pinned between the FMA pipe (45 %) and
// the shared-memory pipe (41 % wavefronts for the warp-broadcast weight loads), two warps per scheduler and
// 255 registers -- the structure of a small-N GEMM chain on CUDA cores.  The 30 layers per point are
// [128 points x 32] . [32 x 32] products: exactly one tcgen05.mma tile (M=128, N=32, K=32 in four K=8 steps).
//
// Mapping (one CTA = 512 points = 4 tiles of 128; 16 warps; 1 CTA/SM):
//   * tile g is owned by warps 4g..4g+3: thread r of the group is point r, TMEM lane r;
//   * A operand (activations) lives in TMEM: every layer the owning threads write hi/lo TF32 halves of their
//     own row with tcgen05.st -- no shared-memory staging, no swizzle, no proxy fence for A;
//   * B operand (weights) is a K-major no-swizzle image in shared memory, hi and lo halves, packed once per
//     call by convonet_pack_umma_kernel (forward images W[o][k], backward images W^T[k][o]);
//   * D (32 fp32 columns per tile) in TMEM; tcgen05.commit -> mbarrier -> tcgen05.ld.32x32b gives each thread
//     its own row back, so bias / residual / ReLU / sign masks / BCE stay thread-private in registers;
//   * 3xTF32: D = A_lo.B_hi + A_hi.B_lo + A_hi.B_hi, fp32 accumulation -- fp32-class accuracy (error per term
//     ~2^-21, the same order as a 32-term fp32 FMA chain), checked against the oracle at the same tolerance.
//   * gather forward / backward: the warp-cooperative scheme of decode v2 (8 lanes x float4 = one texel).
#pragma once
#include "convonet_point.cuh"
#include "decode_v2.cuh"
#include "umma.cuh"

namespace ifd {

constexpr int kV3Threads = 512;
constexpr int kV3Pts = 512;
constexpr int kV3Stride = kV3Pts + 1;
constexpr int kV3ImgFloats = 2048;                 // one layer, one direction: hi (1024) + lo (1024)
constexpr int kV3TileCols = 96;                    // D | A_hi | A_lo
constexpr uint32_t kV3TmemCols = 512;

// blob (kernel layout, W^T [in][out] per layer) -> UMMA images.  out: [2 dirs][n_layers][hi|lo][1024] floats.
__global__ void convonet_pack_umma_kernel(const float* __restrict__ Wb_arg, int n_layers, float* __restrict__ out_arg,
                                          const LoopJob* __restrict__ job) {
  using L = ConvDecLayout<32>;
  const float* __restrict__ Wb = job ? job->W : Wb_arg;
  float* __restrict__ out = job ? const_cast<float*>(job->Wimg) : out_arg;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_layers * 1024) return;
  const int l = e >> 10, i = (e >> 5) & 31, o = e & 31;        // WT[i][o] = W[o][i]
  const float w = Wb[L::kBlk0 + l * L::kLayer + i * 32 + o];
  const float hi = __uint_as_float(umma::tf32_hi(w)), lo = __uint_as_float(umma::tf32_lo(w));
  // forward: B[n = o][k = i]
  float* f = out + ((size_t)0 * n_layers + l) * kV3ImgFloats;
  f[umma::img_offset(o, i, 32) / 4] = hi;
  f[1024 + umma::img_offset(o, i, 32) / 4] = lo;
  // backward (dgrad): B[n = i][k = o]
  float* b = out + ((size_t)1 * n_layers + l) * kV3ImgFloats;
  b[umma::img_offset(i, o, 32) / 4] = hi;
  b[1024 + umma::img_offset(i, o, 32) / 4] = lo;
}

// Bilinear geometry of one point, packed for warp shuffles: the lane that OWNS a point (slot = thread) computes the
// three axes once (normalize_coordinate's division and clamps, grid_sample's unnormalise / clip / floor), and the
// 8-lane group that gathers the point's texels receives 7 words by shuffle instead of redoing that work 8 times.
struct V3Geom {
  int pk[3];     // i0 | has1 << 16 | grad_on << 17   per axis
  float f[3];    // weight of the far corner
  int b;         // cloud of the point
};
__device__ __forceinline__ V3Geom v3_geom(float px, float py, float pz, int R, float denom, int b) {
  V3Geom g;
  const float p[3] = {px, py, pz};
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const Axis a = axis_setup(plane_coord(p[ax], denom), R, denom);
    g.pk[ax] = a.i0 | (a.has1 << 16) | ((a.dscale != 0.0f ? 1 : 0) << 17);
    g.f[ax] = a.f;
  }
  g.b = b;
  return g;
}
// the tap offsets / weights of tapset() (decode_v2.cuh) rebuilt from the packed axes -- same expressions, same bits
struct V3Taps {
  uint32_t off[3][4];
  float w[3][4];
  float f[3], near_w[3];
  int has1[3];
};
// off[pl][t] is a 32-bit index in float4 units into the whole [3][B][R][R][32] plane array (cloud, plane, texel and the
// lane's 4-channel slice folded in), so every tap load is one base + 16 * index address computation.
__device__ __forceinline__ void v3_taps(const int (&pk)[3], const float (&f)[3], int R, uint32_t cloud4, uint32_t plane4,
                                        V3Taps& t) {
  int i0[3];
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    i0[ax] = pk[ax] & 0xffff;
    t.has1[ax] = (pk[ax] >> 16) & 1;
    t.f[ax] = f[ax];
    t.near_w[ax] = sub_rn(1.0f, f[ax]);
  }
#pragma unroll
  for (int pl = 0; pl < 3; ++pl) {
    const int aw = plane_axis_w(pl), ah = plane_axis_h(pl);
    const int w0 = i0[aw], w1 = t.has1[aw] ? i0[aw] + 1 : i0[aw];
    const int h0 = i0[ah], h1 = t.has1[ah] ? i0[ah] + 1 : i0[ah];
    const uint32_t o = cloud4 + (uint32_t)pl * plane4;
    t.off[pl][0] = o + (uint32_t)(h0 * R + w0) * 8u; t.off[pl][1] = o + (uint32_t)(h0 * R + w1) * 8u;
    t.off[pl][2] = o + (uint32_t)(h1 * R + w0) * 8u; t.off[pl][3] = o + (uint32_t)(h1 * R + w1) * 8u;
    const float fw = t.has1[aw] ? t.f[aw] : 0.0f, fh = t.has1[ah] ? t.f[ah] : 0.0f;
    t.w[pl][0] = t.near_w[ah] * t.near_w[aw]; t.w[pl][1] = t.near_w[ah] * fw;
    t.w[pl][2] = fh * t.near_w[aw];           t.w[pl][3] = fh * fw;
  }
}

struct DecodeV3Args {
  const float* planes;
  const float* W;        // plain blob (biases, fc_p, fc_out)
  const float* Wimg;     // packed UMMA images
  const float* xyz;
  float* grad_out;
  double* stat_part;
  int n, K, B, R, n_blocks;
  float denom, target, ginv;
  const LoopJob* job;    // non-null: planes / W / Wimg / xyz / grad_out come from this record (graph replay)
};

struct DecodeV3Smem {
  static __host__ __device__ size_t bytes(int n_blocks) {
    return (size_t)3 * n_blocks * kV3ImgFloats * 4 + (size_t)32 * kV3Stride * 4 + (size_t)kV3Pts * 16 + 256 +
           (size_t)(3 * n_blocks + 6) * 32 * 4;       // + per-layer biases, fc_p (W^T 3x32, b), fc_out (w, b)
  }
};

// One layer on the tensor core for this thread's tile: A := split(x) ; D = A.B ; returns D row in `d`.
// Every thread of the 128-thread group must call it (named barrier id = 1 + group).
__device__ __forceinline__ void v3_layer(const float (&x)[32], uint32_t (&d)[32], uint32_t tile_taddr, uint32_t lane_taddr,
                                         uint32_t img_saddr, uint64_t* bar, uint32_t& parity, int group, bool leader) {
  // `d` doubles as the staging registers of the two tcgen05.st (it is dead until the final tcgen05.ld)
#pragma unroll
  for (int k = 0; k < 32; ++k) d[k] = umma::tf32_hi_fast(x[k]);
  umma::tmem_st32(lane_taddr + 32, d);
  // lo = x - hi exactly (13 significant bits at most); the tensor core reads its top 11 -- no explicit rounding.
  // Error per product <= 2^-21 |a||b| for this term (2^-22 with a rounded lo), same order as the dropped lo.lo term.
#pragma unroll
  for (int k = 0; k < 32; ++k) d[k] = __float_as_uint(x[k] - __uint_as_float(d[k]));
  umma::tmem_st32(lane_taddr + 64, d);
  umma::tmem_wait_st();
  umma::fence_before_sync();
  asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
  if (leader) {
    umma::fence_after_sync();
    constexpr uint32_t idesc = umma::idesc_tf32(128, 32);
    const uint32_t d_t = tile_taddr, a_hi = tile_taddr + 32, a_lo = tile_taddr + 64;
#pragma unroll
    for (int part = 0; part < 3; ++part) {          // lo.hi, hi.lo, hi.hi
      const uint32_t a_t = part == 0 ? a_lo : a_hi;
      const uint32_t b_s = img_saddr + (part == 1 ? 4096u : 0u);
#pragma unroll
      for (int s = 0; s < 4; ++s)
        umma::mma_tf32_ts(d_t, a_t + s * 8, umma::smem_desc_kmajor(b_s + s * 1024, 512, 128), idesc, (part | s) ? 1u : 0u);
    }
    umma::commit(bar);
  }
  __syncwarp();
  umma::mbar_wait(bar, parity);
  parity ^= 1;
  umma::fence_after_sync();
  umma::tmem_ld32(lane_taddr, d);
}

__global__ void __launch_bounds__(kV3Threads, 1) convonet_decode_v3_kernel(const DecodeV3Args a) {
  extern __shared__ float4 smem4[];
  using L = ConvDecLayout<32>;
  const int n_layers = 3 * a.n_blocks;
  float* wimg = reinterpret_cast<float*>(smem4);                           // [n_layers][2048]
  float* feat = wimg + (size_t)n_layers * kV3ImgFloats;                     // [32][kV3Stride]
  float4* gpart = reinterpret_cast<float4*>(feat + 32 * kV3Stride);         // [kV3Pts]
  uint64_t* bars = reinterpret_cast<uint64_t*>(gpart + kV3Pts);             // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* vec = reinterpret_cast<float*>(reinterpret_cast<char*>(bars) + 256);   // [n_layers][32] biases | fc_p 4x32 | fc_out 2x32
  const float* Wb = a.W;

  const int tile0 = blockIdx.x * kV3Pts;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int group = warp >> 2;                                              // tile of this thread
  const int grp = lane >> 3, j4 = lane & 7;
  const float4* __restrict__ planes4 = reinterpret_cast<const float4*>(a.planes);
  const uint32_t plane4 = (uint32_t)(a.R * a.R * 8);              // one plane of one cloud, in float4 units

  if (warp == 0) umma::tmem_alloc(tmem_slot, kV3TmemCols);
  if (threadIdx.x == 32) {
    for (int g = 0; g < 4; ++g) umma::mbar_init(&bars[g], 1);
    umma::fence_mbar_init();
  }
  {  // forward weight images -> smem with cp.async (no register staging): the copy runs under the forward gather
    const float4* src = reinterpret_cast<const float4*>(a.Wimg);
    const uint32_t dst = umma::smem_u32(wimg);
    for (int i = threadIdx.x; i < n_layers * kV3ImgFloats / 4; i += kV3Threads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (uint32_t)i), "l"(src + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (n_layers + 6) * 32; i += kV3Threads) {
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
  const float p0 = a.xyz[(size_t)pi * 3 + 0], p1 = a.xyz[(size_t)pi * 3 + 1], p2 = a.xyz[(size_t)pi * 3 + 2];
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
    feat[(j4 * 4 + 0) * kV3Stride + gslot] = c.x;
    feat[(j4 * 4 + 1) * kV3Stride + gslot] = c.y;
    feat[(j4 * 4 + 2) * kV3Stride + gslot] = c.z;
    feat[(j4 * 4 + 3) * kV3Stride + gslot] = c.w;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  umma::fence_proxy_async();          // weight images were written through the generic proxy
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
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
    const float* bc = vec + (3 * blk + 0) * 32;
    const float* b0 = vec + (3 * blk + 1) * 32;
    const float* b1 = vec + (3 * blk + 2) * 32;
#pragma unroll
    for (int k = 0; k < 32; ++k) x[k] = feat[k * kV3Stride + slot];
    v3_layer(x, d, tile_taddr, lane_taddr, wimg_saddr + (3 * blk + 0) * kV3ImgFloats * 4, bar, parity, group, leader);
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      net[k] += __uint_as_float(d[k]) + bc[k];               // net = net + fc_c(c)
      m |= (net[k] > 0.0f ? 1u : 0u) << k;
      x[k] = fmaxf(net[k], 0.0f);
    }
    mask_a[blk] = m;
    v3_layer(x, d, tile_taddr, lane_taddr, wimg_saddr + (3 * blk + 1) * kV3ImgFloats * 4, bar, parity, group, leader);
    m = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float h = __uint_as_float(d[k]) + b0[k];         // h = fc_0(relu(net))
      m |= (h > 0.0f ? 1u : 0u) << k;
      x[k] = fmaxf(h, 0.0f);
    }
    mask_h[blk] = m;
    v3_layer(x, d, tile_taddr, lane_taddr, wimg_saddr + (3 * blk + 2) * kV3ImgFloats * 4, bar, parity, group, leader);
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
    __shared__ double red[2][kV3Threads / 32];
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
      for (int w = 0; w < kV3Threads / 32; ++w) {
        t0 += red[0][w];
        t1 += red[1][w];
      }
      a.stat_part[blockIdx.x * 2 + 0] = t0;
      a.stat_part[blockIdx.x * 2 + 1] = t1;
    }
  }

  // ---------------- swap in the backward (transposed) weight images
  umma::fence_before_sync();
  __syncthreads();                    // every tile's forward MMAs have completed (each group waited on its own)
  {
    const float4* src = reinterpret_cast<const float4*>(a.Wimg + (size_t)n_layers * kV3ImgFloats);
    float4* dst = reinterpret_cast<float4*>(wimg);
    for (int i = threadIdx.x; i < n_layers * kV3ImgFloats / 4; i += kV3Threads) dst[i] = src[i];
  }
  umma::fence_proxy_async();
  __syncthreads();
  umma::fence_after_sync();

  // ---------------- MLP backward (dgrad)
  float (&gnet)[32] = net;
#pragma unroll
  for (int k = 0; k < 32; ++k) gnet[k] = ((mask_f >> k) & 1u) ? glogit * wo[k] : 0.0f;
#pragma unroll 1
  for (int blk = a.n_blocks - 1; blk >= 0; --blk) {
    v3_layer(gnet, d, tile_taddr, lane_taddr, wimg_saddr + (3 * blk + 2) * kV3ImgFloats * 4, bar, parity, group, leader);
    const uint32_t mh = mask_h[blk], ma = mask_a[blk];
#pragma unroll
    for (int k = 0; k < 32; ++k) x[k] = ((mh >> k) & 1u) ? __uint_as_float(d[k]) : 0.0f;      // gh
    v3_layer(x, d, tile_taddr, lane_taddr, wimg_saddr + (3 * blk + 1) * kV3ImgFloats * 4, bar, parity, group, leader);
#pragma unroll
    for (int k = 0; k < 32; ++k) gnet[k] += ((ma >> k) & 1u) ? __uint_as_float(d[k]) : 0.0f;
    v3_layer(gnet, d, tile_taddr, lane_taddr, wimg_saddr + (3 * blk + 0) * kV3ImgFloats * 4, bar, parity, group, leader);
    const bool first = blk == a.n_blocks - 1;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      float* f = feat + k * kV3Stride + slot;
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
    const float4 gc = make_float4(feat[(j4 * 4 + 0) * kV3Stride + gslot], feat[(j4 * 4 + 1) * kV3Stride + gslot],
                                  feat[(j4 * 4 + 2) * kV3Stride + gslot], feat[(j4 * 4 + 3) * kV3Stride + gslot]);
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
      a.grad_out[o + 0] = gp.x + gi[0] * (((pk[0] >> 17) & 1) ? dsc : 0.0f);
      a.grad_out[o + 1] = gp.y + gi[1] * (((pk[1] >> 17) & 1) ? dsc : 0.0f);
      a.grad_out[o + 2] = gp.z + gi[2] * (((pk[2] >> 17) & 1) ? dsc : 0.0f);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, kV3TmemCols);
}

// Tensor-core self test: D[128][32] = A[128][32] . Bm[32][32]^T through the exact code path of v3_layer.
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                               float* __restrict__ D) {
  __shared__ __align__(128) float img[kV3ImgFloats];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 128);
  if (threadIdx.x == 32) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  for (int e = threadIdx.x; e < 1024; e += 128) {
    const int n = e >> 5, k = e & 31;
    const float w = Bm[n * 32 + k];
    img[umma::img_offset(n, k, 32) / 4] = __uint_as_float(umma::tf32_hi(w));
    img[1024 + umma::img_offset(n, k, 32) / 4] = __uint_as_float(umma::tf32_lo(w));
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tile_taddr = tmem_slot;
  const uint32_t lane_taddr = tile_taddr + ((uint32_t)(warp * 32) << 16);
  float x[32];
  uint32_t d[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) x[k] = A[threadIdx.x * 32 + k];
  uint32_t parity = 0;
  for (int rep = 0; rep < 2; ++rep)      // twice: exercises the mbarrier phase flip and D overwrite
    v3_layer(x, d, tile_taddr, lane_taddr, umma::smem_u32(img), &bar, parity, 0, threadIdx.x == 0);
#pragma unroll
  for (int k = 0; k < 32; ++k) D[threadIdx.x * 32 + k] = __uint_as_float(d[k]);
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tile_taddr, 128);
}

}  // namespace ifd
