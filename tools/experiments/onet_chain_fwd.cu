// EXPERIMENT, NOT PART OF THE BUILD (round-2 head start for DESIGN.md 8b item 1; never run on hardware yet).
//
// The ONet decoder (DecoderCBatchNorm, ONet/im2mesh/onet/models/decoder.py:115-133) as ONE forward kernel per 128-row tile:
// the tile lives in TMEM for the whole 10-layer chain.
//   TMEM columns   0..255  `net`   (fc_p output; every block's fc_1 ACCUMULATES onto it: the residual is free)
//   TMEM columns 256..511  `h`     (fc_0 output of the current block)
// TMEM is full, so the A operand comes from shared memory (SS form of tcgen05.mma): per K chunk of 32 the row-owning thread
// reads its 32 columns of the source region (tcgen05.ld), adds the pending bias, applies the folded CBN + ReLU, splits hi / lo
// and stores the K-major image [128 x 32] (2 x 16 KB, double-buffered); the weight chunk [256 x 32] hi + lo (64 KB,
// double-buffered, the product's packed images: csrc/onet.cu onet_pack_layer_kernel) streams in with cp.async; one thread
// issues the 12 MMAs (3xTF32 x four K = 8 steps, M 128, N 256).  A-preparation of chunk g overlaps the MMAs of chunk g - 1;
// only at a layer boundary does the tile wait for the tensor pipe (the next layer reads what the last one wrote).
// Biases are not added in TMEM: `net` carries a pending vector bacc = sum of the fc_1 biases so far, `h` carries b0; both are
// added when the region is read.
// Numerics: same products and K order as the layered kernels; the residual is accumulated by the tensor core instead of
// added in the epilogue, so results agree to fp32 rounding (not bitwise).
//
// Shared memory: 2 x 64 KB (B) + 2 x 32 KB (A) + 2 KB = 194 KB -> one CTA per SM, 128 threads.
// Needs K % 128 == 0 or B == 1 (a tile must not straddle two clouds: the CBN scale / shift are per cloud).
#include "common.cuh"
#include "umma.cuh"

namespace ifd {

constexpr int kCH = 256;                       // hidden size
constexpr int kCChunk = 32;
constexpr int kCBFloats = 2 * kCH * kCChunk;   // one weight chunk, hi + lo: 16384 floats = 64 KB (= onet.cu kChunkImgFloats)
constexpr int kCAFloats = 2 * 128 * kCChunk;   // one A chunk, hi + lo: 8192 floats = 32 KB
constexpr int kChainThreads = 128;

// D[tmem] (+)= A[smem] . B[smem]^T, both operands through shared-memory descriptors
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

struct ChainArgs {
  const float* xyz;     // [M][3]
  const float* Wp;      // fc_p.weight [256][3]
  const float* bp;      // fc_p.bias [256]
  const float* img;     // forward images of the 10 layers: [10][8 chunks][hi, lo][256 x 32]  (workspace of ifd_onet_prepare)
  const float* fc;      // 10 x { weight [256][256], bias [256] }: only the biases are read here
  const float* s;       // [11][B][256] folded CBN scale
  const float* t;       // [11][B][256] folded CBN shift
  const float* wout;    // fc_out.weight [256]
  const float* bout;    // fc_out.bias [1]
  float* logits;        // [M]
  int M, K, B;
};

__global__ void __launch_bounds__(kChainThreads, 1) onet_chain_fwd_kernel(const ChainArgs a) {
  extern __shared__ float4 smem4[];
  float* bslot = reinterpret_cast<float*>(smem4);                   // [2][kCBFloats]
  float* aslot = bslot + 2 * kCBFloats;                             // [2][kCAFloats]
  float* rbias = aslot + 2 * kCAFloats;                             // [256] bias pending on the region being read
  float* bacc = rbias + kCH;                                        // [256] sum of the fc_1 biases so far (pending on `net`)
  uint64_t* bars = reinterpret_cast<uint64_t*>(bacc + kCH);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  if (tid == 32) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], 1);
    umma::fence_mbar_init();
  }
  // weight chunk g = layer * 8 + kc: 64 KB, contiguous in the packed images
  auto load_b = [&](int g, int buf) {
    const float4* src = reinterpret_cast<const float4*>(a.img + (size_t)g * kCBFloats);
    const uint32_t dst = umma::smem_u32(bslot + (size_t)buf * kCBFloats);
    for (int i = tid; i < kCBFloats / 4; i += kChainThreads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (uint32_t)i), "l"(src + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_b(0, 0);
  for (int n = tid; n < kCH; n += kChainThreads) {
    bacc[n] = 0.0f;
    rbias[n] = 0.0f;
  }
  __syncthreads();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_t = tmem + ((uint32_t)(warp * 32) << 16);
  const int row = blockIdx.x * kChainThreads + tid;
  const int rowc = min(row, a.M - 1);
  const int b = rowc / a.K;
  const bool leader = tid == 0;
  uint32_t parity[2] = {0u, 0u};
  constexpr uint32_t idesc = umma::idesc_tf32(128, kCH);

  // net0 = fc_p(p) -> TMEM columns 0..255 of this thread's lane
  {
    const float px = a.xyz[(size_t)rowc * 3 + 0], py = a.xyz[(size_t)rowc * 3 + 1], pz = a.xyz[(size_t)rowc * 3 + 2];
#pragma unroll 1
    for (int cg = 0; cg < kCH / 32; ++cg) {
      uint32_t u[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int n = cg * 32 + k;
        float v = __ldg(a.bp + n);
        v = fmaf(__ldg(a.Wp + n * 3 + 0), px, v);
        v = fmaf(__ldg(a.Wp + n * 3 + 1), py, v);
        v = fmaf(__ldg(a.Wp + n * 3 + 2), pz, v);
        u[k] = __float_as_uint(v);
      }
      umma::tmem_st32(lane_t + cg * 32, u);
    }
    umma::tmem_wait_st();
  }

  const int n_chunks = 10 * (kCH / kCChunk);
#pragma unroll 1
  for (int g = 0; g < n_chunks; ++g) {
    const int l = g >> 3, kc = g & 7, buf = g & 1;
    const uint32_t src_col = (l & 1) ? 256u : 0u;                    // even layers (fc_0) read `net`, odd (fc_1) read `h`
    const uint32_t dst_col = (l & 1) ? 0u : 256u;
    if (kc == 0 && g > 0) {
      // layer boundary: this layer reads what the previous one wrote -> all of its MMAs must have completed.  The last
      // commit (chunk g - 1, slot buf ^ 1) covers every earlier MMA.
      umma::mbar_wait(&bars[buf ^ 1], parity[buf ^ 1]);
      parity[buf ^ 1] ^= 1;
      umma::fence_after_sync();
      // pending bias of the region that is read now: after fc_0 the region `h` lacks b0 of layer l - 1; after fc_1 the
      // region `net` lacks every b1 so far
      const float* bprev = a.fc + (size_t)(l - 1) * ((size_t)kCH * kCH + kCH) + (size_t)kCH * kCH;
      __syncthreads();                                              // nobody still reads rbias of the previous layer
      for (int n = tid; n < kCH; n += kChainThreads) {
        if (l & 1) rbias[n] = __ldg(bprev + n);                     // reading h: + b0
        else {
          bacc[n] += __ldg(bprev + n);                              // reading net: + sum of b1
          rbias[n] = bacc[n];
        }
      }
      __syncthreads();
    }
    // A chunk: columns kc*32 .. +32 of the source region, + pending bias, CBN l, ReLU
    uint32_t d[32];
    umma::tmem_ld32(lane_t + src_col + kc * 32, d);
    float x[32];
    {
      const float4* s4 = reinterpret_cast<const float4*>(a.s + ((size_t)l * a.B + b) * kCH + kc * 32);
      const float4* t4 = reinterpret_cast<const float4*>(a.t + ((size_t)l * a.B + b) * kCH + kc * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 s = __ldg(s4 + q), t = __ldg(t4 + q);
        const float4 rb = *reinterpret_cast<const float4*>(rbias + kc * 32 + 4 * q);
        x[4 * q + 0] = fmaxf(fmaf(s.x, __uint_as_float(d[4 * q + 0]) + rb.x, t.x), 0.0f);
        x[4 * q + 1] = fmaxf(fmaf(s.y, __uint_as_float(d[4 * q + 1]) + rb.y, t.y), 0.0f);
        x[4 * q + 2] = fmaxf(fmaf(s.z, __uint_as_float(d[4 * q + 2]) + rb.z, t.z), 0.0f);
        x[4 * q + 3] = fmaxf(fmaf(s.w, __uint_as_float(d[4 * q + 3]) + rb.w, t.w), 0.0f);
      }
    }
    if (g + 1 < n_chunks) {                  // prefetch the next weight chunk into the other slot
      if (g >= 1 && kc != 0) {               // MMA g-1 read that slot and A[buf ^ 1]; at kc == 0 it was waited for above
        umma::mbar_wait(&bars[buf ^ 1], parity[buf ^ 1]);
        parity[buf ^ 1] ^= 1;
        umma::fence_after_sync();
      }
      load_b(g + 1, buf ^ 1);
    }
    // K-major image of this thread's row: k group q (4 floats) of row r at ((q * 16 + (r >> 3)) * 128 + (r & 7) * 16) bytes
    {
      float* ahi = aslot + (size_t)buf * kCAFloats;
      float* alo = ahi + 128 * kCChunk;
      const int r = tid;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t off = umma::img_offset(r, 4 * q, 128) / 4;
        float4 hi, lo;
        hi.x = __uint_as_float(umma::tf32_hi_fast(x[4 * q + 0]));
        hi.y = __uint_as_float(umma::tf32_hi_fast(x[4 * q + 1]));
        hi.z = __uint_as_float(umma::tf32_hi_fast(x[4 * q + 2]));
        hi.w = __uint_as_float(umma::tf32_hi_fast(x[4 * q + 3]));
        lo.x = x[4 * q + 0] - hi.x;                                   // exact residual, top 11 bits used
        lo.y = x[4 * q + 1] - hi.y;
        lo.z = x[4 * q + 2] - hi.z;
        lo.w = x[4 * q + 3] - hi.w;
        *reinterpret_cast<float4*>(ahi + off) = hi;
        *reinterpret_cast<float4*>(alo + off) = lo;
      }
    }
    if (g + 1 < n_chunks) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    umma::fence_proxy_async();               // generic-proxy writes (A image, cp.async data) -> visible to the tensor core
    umma::fence_before_sync();
    __syncthreads();
    if (leader) {
      umma::fence_after_sync();
      const uint32_t sa = umma::smem_u32(aslot + (size_t)buf * kCAFloats);
      const uint32_t sb = umma::smem_u32(bslot + (size_t)buf * kCBFloats);
      // fc_0 starts a fresh accumulator in `h`; fc_1 accumulates onto `net`
      const bool fresh = (l & 1) == 0;
#pragma unroll
      for (int part = 0; part < 3; ++part) {        // lo.hi, hi.lo, hi.hi
        const uint32_t as = sa + (part == 0 ? (uint32_t)(128 * kCChunk * 4) : 0u);
        const uint32_t bs = sb + (part == 1 ? (uint32_t)(kCH * kCChunk * 4) : 0u);
#pragma unroll
        for (int s = 0; s < 4; ++s)
          mma_tf32_ss(tmem + dst_col, umma::smem_desc_kmajor(as + s * 2 * (128 / 8) * 128, (128 / 8) * 128, 128),
                      umma::smem_desc_kmajor(bs + s * 2 * (kCH / 8) * 128, (kCH / 8) * 128, 128), idesc,
                      (fresh && kc == 0 && part == 0 && s == 0) ? 0u : 1u);
      }
      umma::commit(&bars[buf]);
    }
    __syncwarp();
  }
  // drain: the last commit (chunk 79, slot 1) covers everything
  umma::mbar_wait(&bars[(n_chunks - 1) & 1], parity[(n_chunks - 1) & 1]);
  umma::fence_after_sync();
  __syncthreads();
  {
    const float* b9 = a.fc + (size_t)9 * ((size_t)kCH * kCH + kCH) + (size_t)kCH * kCH;
    for (int n = tid; n < kCH; n += kChainThreads) rbias[n] = bacc[n] + __ldg(b9 + n);
  }
  __syncthreads();
  // logit = w_out . relu(s_10 * net + t_10) + b_out
  float acc = 0.0f;
#pragma unroll 1
  for (int cg = 0; cg < kCH / 32; ++cg) {
    uint32_t d[32];
    umma::tmem_ld32(lane_t + cg * 32, d);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int n = cg * 32 + k;
      const float pre = fmaf(__ldg(a.s + ((size_t)10 * a.B + b) * kCH + n), __uint_as_float(d[k]) + rbias[n],
                             __ldg(a.t + ((size_t)10 * a.B + b) * kCH + n));
      acc = fmaf(__ldg(a.wout + n), fmaxf(pre, 0.0f), acc);
    }
  }
  if (row < a.M) a.logits[row] = acc + __ldg(a.bout);
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace ifd

// Device pointers.  dec_weights: the packed ONet decoder blob (include/ifd_b200.h); prepared_ws: a workspace on which
// ifd_onet_prepare(dec_weights, c, B, K, ...) has run (weight images at its head, then s, then t).
extern "C" int exp_onet_chain_fwd(const float* dec_weights, const float* xyz, int B, int K, const float* img, const float* s,
                                  const float* t, float* logits, void* stream) {
  using namespace ifd;
  if (!(B == 1 || K % 128 == 0)) return -1;
  const size_t kOffCbnX = (size_t)kCH * 3 + kCH;
  const size_t kCbnFloatsX = 2 * ((size_t)kCH * 512 + kCH) + 2 * kCH;
  const size_t kOffFcX = kOffCbnX + 11 * kCbnFloatsX;
  const size_t kOffOutX = kOffFcX + 10 * ((size_t)kCH * kCH + kCH);
  ChainArgs a{xyz, dec_weights, dec_weights + kCH * 3, img, dec_weights + kOffFcX, s, t, dec_weights + kOffOutX,
              dec_weights + kOffOutX + kCH, logits, B * K, K, B};
  const size_t smem = (size_t)(2 * kCBFloats + 2 * kCAFloats + 2 * kCH) * sizeof(float) + 64;
  if (cudaFuncSetAttribute((const void*)onet_chain_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
  onet_chain_fwd_kernel<<<(a.M + kChainThreads - 1) / kChainThreads, kChainThreads, smem, (cudaStream_t)stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
