// decode_v3.cuh -- ConvONet decode with the ResNet-MLP on the 5th-generation tensor cores: the building blocks (weight images,
// bilinear geometry packed for shuffles, one layer as a tcgen05 round trip, self test).  The kernels that use them are
// decode_v5.cuh; the earlier kernels of this generation (v3: 512-point CTAs, all images resident, one CTA per SM; v4: 256-point
// CTAs, two per SM, 30 dependent round trips) were retired in round 2.
//
// Why: ncu on decode v2 (profiles/r01_v2_*) shows the SIMT formulation pinned between the FMA pipe (45 %) and
// the shared-memory pipe (41 % wavefronts for the warp-broadcast weight loads), two warps per scheduler and
// 255 registers -- the structure of a small-N GEMM chain on CUDA cores.  The 30 layers per point are
// [128 points x 32] . [32 x 32] products: exactly one tcgen05.mma tile (M=128, N=32, K=32 in four K=8 steps).
//
// Mapping (a CTA = tiles of 128 points, 4 warps per tile):
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

constexpr int kV3ImgFloats = 2048;                 // one layer, one direction: hi (1024) + lo (1024)
constexpr int kV3TileCols = 96;                    // D | A_hi | A_lo

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
  const LoopJob* job;    // non-null: planes / W / Wimg / xyz / grad_out / jac come from this record (graph replay)
  float* jac;            // [n][3][32] workspace: the forward gather leaves d c / d xyz here, the backward pass reads it back
                         // instead of gathering the texels again (nullptr: gather twice)
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
