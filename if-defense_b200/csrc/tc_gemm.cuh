// tc_gemm.cuh -- one warp-specialised, persistent tcgen05 GEMM engine for every dense contraction of the path that is not
// the fused ConvONet decode: the Linear stacks of both encoders, the U-Net's convolutions (implicit GEMM, channels-last),
// the transposed convolutions, and the 256-wide ONet layers.
//
//   Out[m][n] = epilogue( sum_k A(m, k) * W[n][k] ),   fp32 semantics through 3xTF32 (A_lo.B_hi + A_hi.B_lo + A_hi.B_hi,
//   fp32 accumulation in TMEM).  The tensor core truncates the accumulator at every MMA, so what limits the accuracy of a
//   long K is the NUMBER of accumulations at full magnitude: for NT <= 128 the two small cross terms go to a second
//   accumulator (2^-11 of the magnitude, so its truncations do not count) and only the four hi.hi MMAs per chunk touch the
//   main one; the epilogue adds the two.  Measured: 3.3e-5 -> ~1e-5 absolute at K = 1152, |result| ~ 1.
//
// A CTA walks tiles of 128 rows x NT columns (tile index = blockIdx.x + i * gridDim.x).  Roles:
//   warps 0-3  A producers: thread r owns row r of the tile; per K chunk of 32 it asks the Policy for the 32 values of
//              its row (a global row of a matrix, a 3x3 tap of a channels-last image, a pooled pixel, ...), splits them into
//              TF32 hi / lo and writes the two K-major no-swizzle UMMA images of the chunk into a pipeline stage;
//   warp 9     B loader: one thread, ONE TMA bulk copy (cp.async.bulk) per chunk of the pre-packed weight image
//              [NT x 32] hi | lo, completion counted in bytes on the stage's mbarrier; in a cluster every CTA fetches 1 / CL
//              of the chunk and multicasts it to all CTAs (L2 -> SM weight traffic / CL);
//   warp 8     MMA issuer: one thread, 12 tcgen05.mma (M = 128, N = NT, K = 8) per chunk, both operands from shared memory,
//              tcgen05.commit releases the stage; the accumulator of a tile lives in one of TWO TMEM stages;
//   warps 4-7  epilogue: tcgen05.ld the finished accumulator (thread = row), hand 32 columns at a time to the Policy
//              (bias, ReLU, residual, pixel shuffle, ...), while the MMA warp is already on the next tile.
// Stage hand-off is mbarrier-only (no __syncthreads in the main loop).
//
// Measured and kept out (round 2, ncu on the 256-wide ONet layer: 93 us, LSU wavefronts 71 % of peak, tensor pipe 31 %): the
// thread-per-row float4 accesses of producers and epilogue cost 32 L1 wavefronts per instruction.  A coalesced mapping (eight
// lanes per 128-byte row piece, bank-rotated image stride, epilogue transposed through an XOR-swizzled scratch) cut the LSU
// load to 13 % and was still SLOWER (ONet step 2.1 -> 3.0 ms, U-Net 12.5 -> 18.4 ms): eight address computations per chunk
// and lane instead of one, and a latency chain per 32 columns in the epilogue.  The way out is TMA tensor loads of the raw A
// tile and TMA stores of the output tile (no LSU traffic at all), not another lane mapping.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace ifd {
namespace tc {

constexpr int kThreads = 320;
constexpr int kChunk = 32;                       // K per pipeline stage
constexpr int kAStageBytes = 2 * 128 * kChunk * 4;   // hi + lo images of the A chunk: 32 KB

__host__ __device__ constexpr int b_stage_bytes(int NT) { return 2 * NT * kChunk * 4; }
__host__ __device__ constexpr int stages_for(int NT) { return NT >= 256 ? 2 : (NT >= 128 ? 3 : 4); }
__host__ __device__ constexpr size_t smem_bytes(int NT) {
  return (size_t)stages_for(NT) * (kAStageBytes + b_stage_bytes(NT)) + 256;      // stages | barriers (<= 24 x 8 B) + TMEM slot
}

using umma::mma_tf32_ss;       // D[tmem] (+)= A[smem] . B[smem]^T
// Warp-transposed ("blocked") row layout for tensors the engine reads and writes thread-per-row: rows in blocks of 32, inside a
// block the C / 4 float4 column groups one after the other, each holding its 32 rows contiguously.  The 32 lanes of a warp that
// access "their row's float4 number q" then touch 512 contiguous bytes (4 L1 wavefronts) instead of 32 separate lines.
// Buffers hold ceil(M / 32) * 32 rows.  Offset in float4 units:
__host__ __device__ __forceinline__ size_t blk_off4(size_t m, int c4, int C4) { return ((m >> 5) * (size_t)C4 + (size_t)c4) * 32 + (m & 31); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

// Weight pre-pack: W (row-major [N][K], row stride ldw; or its transpose when `transposed`: element (n, k) = W[k * ldw + n])
// -> [N_pad / NT tiles][K_pad / 32 chunks][hi | lo][NT x 32] K-major UMMA images, zero padded.
static __global__ void pack_weights_kernel(const float* __restrict__ W, int N, int K, int ldw, int transposed, int NT, int n_tiles,
                                    int n_chunks, float* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_tiles * n_chunks * NT * kChunk;
  if (e >= total) return;
  const int kl = (int)(e % kChunk);
  const int nl = (int)((e / kChunk) % NT);
  const int kc = (int)((e / ((long long)kChunk * NT)) % n_chunks);
  const int nt = (int)(e / ((long long)kChunk * NT * n_chunks));
  const int n = nt * NT + nl, k = kc * kChunk + kl;
  float w = 0.0f;
  if (n < N && k < K) w = transposed ? W[(size_t)k * ldw + n] : W[(size_t)n * ldw + k];
  float* img = out + ((size_t)nt * n_chunks + kc) * (2 * NT * kChunk);
  const uint32_t off = umma::img_offset(nl, kl, NT) / 4;
  img[off] = __uint_as_float(umma::tf32_hi(w));
  img[NT * kChunk + off] = __uint_as_float(umma::tf32_lo(w));
}

// A Policy provides
//   struct Params (trivially copyable; must contain  int M, n_chunks, n_tiles_n;  const float* wimg;)
//   struct Row;                                                            per-row state of the producer / epilogue thread
//   static __device__ Row row_begin(const Params&, int row);               row < M is NOT guaranteed (check yourself)
//   static __device__ void load(const Params&, const Row&, int kc, float (&x)[32]);        the 32 A values of chunk kc
//   static __device__ void store(const Params&, const Row&, int col0, const float (&y)[32]); columns col0 .. col0 + 31
// ---- thread-block clusters: the CTAs of a cluster work on different row tiles of the SAME column tile in lock step, so a
// weight chunk is fetched from L2 once per cluster: every CTA loads 1 / CL of it and TMA-multicasts that piece into the
// same stage of all CL shared memories (each CTA's mbarrier counts the bytes of all pieces); a stage is reused when the
// MMAs of ALL CTAs have read it (tcgen05.commit multicast onto every CTA's `empty` barrier).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          umma::smem_u32(dst_smem)),
      "l"(src), "r"(bytes), "r"(umma::smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   umma::smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// A CHAIN of layers in one launch (n_layers > 1): the layers share M, n_chunks and n_tiles_n, layer l + 1 reads rows that layer l
// wrote, and nothing else couples them -- a row tile depends only on the same row tile of the layer before.  Tile t of every
// layer runs on the same CTA, so the only hand-over needed is inside the CTA: the epilogue warps fence their stores and arrive
// on tile_done[local tile], the producers wait for it before they read that tile's rows for the next layer.  No launch
// boundary, no pipeline drain and refill between layers (14 us of 58 per 256-wide ONet layer, 20 layers per Adam step).
// Policies used in a chain must read activations with coherent loads (__ldcg), not through the read-only path.
constexpr int kMaxLocalTiles = 8;

template <class Policy, int NT, int CL>
__device__ __forceinline__ void gemm_body(const typename Policy::Params* __restrict__ layers, int n_layers) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int S = stages_for(NT);
  constexpr int kBStage = b_stage_bytes(NT);
  constexpr bool kSplit = NT <= 128;                        // second accumulator for the small cross terms (see the header)
  constexpr uint32_t kAcc = kSplit ? 2 * NT : NT;           // TMEM columns per accumulator stage (main | small)
  constexpr uint32_t kTmemCols = 2 * kAcc;                  // 64 .. 512, a power of two
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1);
  unsigned char* a_st = smem_raw;                           // [S][32 KB]
  unsigned char* b_st = smem_raw + (size_t)S * kAStageBytes;    // [S][kBStage]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_st + (size_t)S * kBStage);
  uint64_t* full_a = bars;                 // [S] 128 producer arrivals
  uint64_t* full_b = bars + S;             // [S] 1 arrival + tx bytes
  uint64_t* empty = bars + 2 * S;          // [S] tcgen05.commit of every CTA of the cluster
  uint64_t* acc_full = bars + 3 * S;       // [2] tcgen05.commit after the last chunk of a tile
  uint64_t* acc_empty = bars + 3 * S + 2;  // [2] 128 epilogue arrivals
  uint64_t* tile_done = bars + 3 * S + 4;  // [kMaxLocalTiles] 128 epilogue arrivals per layer (chains only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4 + kMaxLocalTiles);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      umma::mbar_init(&full_a[s], 128);
      umma::mbar_init(&full_b[s], 1);
      umma::mbar_init(&empty[s], CL);
    }
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(&acc_full[s], 1);
      umma::mbar_init(&acc_empty[s], 128);
    }
    for (int s = 0; s < kMaxLocalTiles; ++s) umma::mbar_init(&tile_done[s], 128);
    umma::fence_mbar_init();
  }
  if (warp == 8) umma::tmem_alloc(tmem_slot, kTmemCols);
  umma::fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // every CTA's barriers exist before a peer multicasts into them
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  // Tiles: cluster tile ct -> column tile ct % n_tiles_n, row tiles (ct / n_tiles_n) * CL + rank.  Every CTA of a cluster runs
  // the same number of rounds; a row tile beyond M is a dummy (the policies load zeros and store nothing for rows >= M).
  const typename Policy::Params& P0 = layers[0];
  const int m_tiles = (P0.M + 127) / 128;
  const int n_tiles_n = P0.n_tiles_n;
  const int n_ct = ((m_tiles + CL - 1) / CL) * n_tiles_n;
  const int n_clusters = gridDim.x / CL, cid = blockIdx.x / CL;
  const int n_chunks = P0.n_chunks;
  auto row_tile = [&](int ct) { return (ct / n_tiles_n) * CL + rank; };

  if (warp < 4) {
    // ------------------------------------------------------------------ A producers
    const int r = threadIdx.x;                      // row of the tile
    uint32_t it = 0;                                // global chunk counter -> stage / phase
    // The values of chunk kc + 1 (of the next tile after the last chunk) are fetched BEFORE chunk kc is split and stored: one
    // producer warp per scheduler cannot hide a global-load latency any other way.
    float xn[32];
    typename Policy::Row row = Policy::row_begin(P0, row_tile(cid) * 128 + r);
    bool have = false;                              // xn holds the values of the chunk that comes next
    if (cid < n_ct) {
      Policy::load(P0, row, 0, xn);
      have = true;
    }
    for (int l = 0; l < n_layers; ++l) {
      const typename Policy::Params& P = layers[l];
      int tloc = 0;
      for (int ct = cid; ct < n_ct; ct += n_clusters, ++tloc) {
        for (int kc = 0; kc < n_chunks; ++kc, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          if (!have) {                              // first chunk of a layer whose rows this CTA's epilogue may still be writing
            umma::mbar_wait(&tile_done[tloc], (uint32_t)((l - 1) & 1));
            row = Policy::row_begin(P, row_tile(ct) * 128 + r);
            Policy::load(P, row, kc, xn);
          }
          float x[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) x[k] = xn[k];
          have = true;
          if (kc + 1 < n_chunks) {
            Policy::load(P, row, kc + 1, xn);
          } else if (ct + n_clusters < n_ct) {      // next tile of this layer (its rows of layer l - 1 were finished a round ago)
            if (l > 0) umma::mbar_wait(&tile_done[tloc + 1], (uint32_t)((l - 1) & 1));
            row = Policy::row_begin(P, row_tile(ct + n_clusters) * 128 + r);
            Policy::load(P, row, 0, xn);
          } else if (l + 1 < n_layers && tloc > 0) {   // first tile of the next layer: finished while later tiles of this layer ran
            umma::mbar_wait(&tile_done[0], (uint32_t)(l & 1));
            row = Policy::row_begin(layers[l + 1], row_tile(cid) * 128 + r);
            Policy::load(layers[l + 1], row, 0, xn);
          } else {
            have = false;                           // (a CTA with ONE tile: its epilogue still needs the chunk stored below)
          }
          umma::mbar_wait(&empty[s], ph ^ 1);
          unsigned char* hi = a_st + (size_t)s * kAStageBytes + (r >> 3) * 128 + (r & 7) * 16;
          unsigned char* lo = hi + 128 * kChunk * 4;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            uint4 h, l4;
            h.x = umma::tf32_hi_cvt(x[4 * g + 0]); h.y = umma::tf32_hi_cvt(x[4 * g + 1]);
            h.z = umma::tf32_hi_cvt(x[4 * g + 2]); h.w = umma::tf32_hi_cvt(x[4 * g + 3]);
            l4.x = __float_as_uint(x[4 * g + 0] - __uint_as_float(h.x)); l4.y = __float_as_uint(x[4 * g + 1] - __uint_as_float(h.y));
            l4.z = __float_as_uint(x[4 * g + 2] - __uint_as_float(h.z)); l4.w = __float_as_uint(x[4 * g + 3] - __uint_as_float(h.w));
            *reinterpret_cast<uint4*>(hi + g * 2048) = h;      // k group g: 16 core-matrix rows of 128 B per 8 tile rows
            *reinterpret_cast<uint4*>(lo + g * 2048) = l4;
          }
          umma::fence_proxy_async();                  // generic-proxy stores -> visible to the tensor core's async proxy
          mbar_arrive(&full_a[s]);
        }
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ epilogue (warp - 4 = TMEM lane quarter)
    const int r = threadIdx.x - 128;
    uint32_t tl = 0;
    for (int l = 0; l < n_layers; ++l) {
      const typename Policy::Params& P = layers[l];
      int tloc = 0;
      for (int ct = cid; ct < n_ct; ct += n_clusters, ++tl, ++tloc) {
        const int nt = ct % n_tiles_n;
        const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
        const typename Policy::Row row = Policy::row_begin(P, row_tile(ct) * 128 + r);
        umma::mbar_wait(&acc_full[as], aph);
        umma::fence_after_sync();
        const uint32_t taddr = tmem + as * kAcc + ((uint32_t)((warp - 4) * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < NT / 32; ++c) {
          uint32_t d[32];
          umma::tmem_ld32(taddr + c * 32, d);
          float y[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) y[k] = __uint_as_float(d[k]);
          if (kSplit) {
            umma::tmem_ld32(taddr + NT + c * 32, d);
#pragma unroll
            for (int k = 0; k < 32; ++k) y[k] += __uint_as_float(d[k]);
          }
          Policy::store(P, row, nt * NT + c * 32, y);
        }
        umma::fence_before_sync();
        mbar_arrive(&acc_empty[as]);
        if (n_layers > 1) {                           // this tile's rows of layer l are in memory: the next layer may read them
          __threadfence();
          mbar_arrive(&tile_done[tloc]);
        }
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the loop and waits on the barriers; one ELECTED lane issues (umma::elect_one: with warp-uniform
    // operands the twelve MMAs of a chunk are twelve SASS instructions, not twelve ELECT / R2UR / BRA.U.ANY loops).
    {
      constexpr uint32_t idesc = umma::idesc_tf32(128, NT);
      const uint32_t tmem_u = umma::warp_bcast(tmem);
      const uint32_t cid_u = umma::warp_bcast((uint32_t)cid), ncl_u = umma::warp_bcast((uint32_t)n_clusters);
      uint32_t it = 0, tl = 0;
      for (int l = 0; l < n_layers; ++l) {
        for (uint32_t ct = cid_u; ct < (uint32_t)n_ct; ct += ncl_u, ++tl) {
          const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
          umma::mbar_wait(&acc_empty[as], aph ^ 1);   // the epilogue drained this accumulator stage (two tiles ago)
          umma::fence_after_sync();
          const uint32_t d_t = tmem_u + as * kAcc;
          for (int kc = 0; kc < n_chunks; ++kc, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            umma::mbar_wait(&full_a[s], ph);
            umma::mbar_wait(&full_b[s], ph);
            umma::fence_after_sync();
            const uint32_t sa = umma::smem_u32(a_st + (size_t)s * kAStageBytes);
            const uint32_t sb = umma::smem_u32(b_st + (size_t)s * kBStage);
            if (umma::elect_one()) {
#pragma unroll
              for (int part = 0; part < 3; ++part) {    // lo.hi, hi.lo, hi.hi (small terms first)
                const uint32_t a_s = sa + (part == 0 ? 128 * kChunk * 4 : 0);
                const uint32_t b_s = sb + (part == 1 ? NT * kChunk * 4 : 0);
                const uint32_t d_p = (kSplit && part < 2) ? d_t + NT : d_t;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint32_t later = kSplit ? (part == 1 ? 1u : (uint32_t)q) : (uint32_t)(part | q);    // 0 only for the first MMA into an accumulator
                  mma_tf32_ss(d_p, umma::smem_desc_kmajor(a_s + q * 4096, 2048, 128),
                              umma::smem_desc_kmajor(b_s + q * (NT * 32), NT * 16, 128), idesc, (kc | later) ? 1u : 0u);
                }
              }
              // the stage is free once these MMAs have read it -- in EVERY CTA of the cluster (their loaders write into it)
              if (CL > 1) commit_multicast(&empty[s], kMask);
              else umma::commit(&empty[s]);
              if (kc + 1 == n_chunks) umma::commit(&acc_full[as]);    // accumulator complete -> epilogue
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ B loader (TMA bulk copies)
    if (lane == 0) {
      constexpr uint32_t kPiece = kBStage / CL;
      uint32_t it = 0;
      for (int l = 0; l < n_layers; ++l) {
        const float* wimg = layers[l].wimg;
        for (int ct = cid; ct < n_ct; ct += n_clusters) {
          const int nt = ct % n_tiles_n;
          const unsigned char* src = reinterpret_cast<const unsigned char*>(wimg + (size_t)nt * n_chunks * (2 * NT * kChunk));
          for (int kc = 0; kc < n_chunks; ++kc, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            umma::mbar_wait(&empty[s], ph ^ 1);
            umma::mbar_arrive_expect_tx(&full_b[s], kBStage);       // all CL pieces land here, one of them is ours
            unsigned char* dst = b_st + (size_t)s * kBStage + (size_t)rank * kPiece;
            const unsigned char* from = src + (size_t)kc * kBStage + (size_t)rank * kPiece;
            if (CL > 1) bulk_g2s_multicast(dst, from, kPiece, &full_b[s], kMask);
            else umma::bulk_g2s(dst, from, kPiece, &full_b[s]);
          }
        }
      }
    }
    __syncwarp();
  }
  umma::fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // no CTA leaves while a peer may still multicast into its shared memory
  if (warp == 8) umma::tmem_dealloc(tmem, kTmemCols);
}

template <class Policy, int NT, int CL>
__global__ void __launch_bounds__(kThreads, 1) gemm_kernel(const typename Policy::Params P) {
  gemm_body<Policy, NT, CL>(&P, 1);
}
template <class Policy, int NT, int CL>
__global__ void __launch_bounds__(kThreads, 1) gemm_chain_kernel(const typename Policy::Params* layers, int n_layers) {
  gemm_body<Policy, NT, CL>(layers, n_layers);
}

// ------------------------------------------------------------------------------------------------ host side
// Tile width of the encoder operators: at most 128 columns, so that every tile keeps the small cross terms in a second
// accumulator (the long K of a convolution, up to 9 x 256, needs it to stay at fp32-class accuracy).  256-wide tiles are
// used by the ONet decoder layers only (K = 256, one A pass per row matters more there).
inline int pick_nt(int N) { return N <= 32 ? 32 : (N <= 64 ? 64 : 128); }

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

inline int g_cluster_size = 2;       // ifd_test_hook(6, n): CTAs per cluster of the GEMM engine (1, 2 or 4)

template <class Policy, int NT, int CL>
int launch_one(const typename Policy::Params& P, int clusters_wanted, cudaStream_t st) {
  const size_t smem = smem_bytes(NT);
  const void* fn = (const void*)gemm_kernel<Policy, NT, CL>;
  IFD_CUDA_TRY(set_max_dyn_smem(fn, smem));
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  int max_clusters = sm_count() / CL;
  if (CL > 1) {                        // how many clusters of this shape the chip holds at once (GPC boundaries), once per kernel
    static int cached = -1;
    if (cached < 0) {
      cfg.gridDim = dim3(sm_count() / CL * CL);
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) == cudaSuccess && n > 0) cached = n;
      else { cudaGetLastError(); cached = max_clusters; }
    }
    max_clusters = cached < max_clusters ? cached : max_clusters;
  }
  const int n = clusters_wanted < max_clusters ? clusters_wanted : max_clusters;
  cfg.gridDim = dim3(n * CL);
  IFD_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_kernel<Policy, NT, CL>, P));
  IFD_LAUNCH_CHECK("tc::gemm_kernel");
  return IFD_OK;
}

template <class Policy, int NT>
int launch_nt(const typename Policy::Params& P, cudaStream_t st) {
  const int m_tiles = (P.M + 127) / 128;
  if (m_tiles <= 0) return IFD_OK;
  int cl = g_cluster_size;
  while (cl > 1 && m_tiles < 2 * cl) cl >>= 1;           // tiny problems: no point in padding row tiles
  const int wanted = ((m_tiles + cl - 1) / cl) * P.n_tiles_n;
  if (cl >= 4) return launch_one<Policy, NT, 4>(P, wanted, st);
  if (cl == 2) return launch_one<Policy, NT, 2>(P, wanted, st);
  return launch_one<Policy, NT, 1>(P, wanted, st);
}

// A chain of n_layers layers (device array `dev_layers`; `first` = a host copy of its first element, for the shapes) in one
// launch.  -> IFD_OK, or 1 when the shape does not fit the chain form (more than kMaxLocalTiles row tiles per CTA): the caller
// then launches the layers one by one.
template <class Policy, int NT, int CL>
int launch_chain_one(const typename Policy::Params* dev_layers, int n_layers, const typename Policy::Params& first, cudaStream_t st) {
  const size_t smem = smem_bytes(NT);
  const void* fn = (const void*)gemm_chain_kernel<Policy, NT, CL>;
  IFD_CUDA_TRY(set_max_dyn_smem(fn, smem));
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  int max_clusters = sm_count() / CL;
  if (CL > 1) {
    static int cached = -1;
    if (cached < 0) {
      cfg.gridDim = dim3(sm_count() / CL * CL);
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) == cudaSuccess && n > 0) cached = n;
      else { cudaGetLastError(); cached = max_clusters; }
    }
    max_clusters = cached < max_clusters ? cached : max_clusters;
  }
  const int m_tiles = (first.M + 127) / 128;
  const int wanted = ((m_tiles + CL - 1) / CL) * first.n_tiles_n;
  const int n = wanted < max_clusters ? wanted : max_clusters;
  if ((wanted + n - 1) / n > kMaxLocalTiles) return 1;
  cfg.gridDim = dim3(n * CL);
  IFD_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_chain_kernel<Policy, NT, CL>, dev_layers, n_layers));
  IFD_LAUNCH_CHECK("tc::gemm_chain_kernel");
  return IFD_OK;
}
template <class Policy, int NT>
int launch_chain(const typename Policy::Params* dev_layers, int n_layers, const typename Policy::Params& first, cudaStream_t st) {
  const int m_tiles = (first.M + 127) / 128;
  if (m_tiles <= 0 || n_layers <= 0) return IFD_OK;
  int cl = g_cluster_size;
  while (cl > 1 && m_tiles < 2 * cl) cl >>= 1;
  if (cl >= 4) return launch_chain_one<Policy, NT, 4>(dev_layers, n_layers, first, st);
  if (cl == 2) return launch_chain_one<Policy, NT, 2>(dev_layers, n_layers, first, st);
  return launch_chain_one<Policy, NT, 1>(dev_layers, n_layers, first, st);
}

template <class Policy>
int launch(const typename Policy::Params& P, int NT, cudaStream_t st) {
  switch (NT) {
    case 32: return launch_nt<Policy, 32>(P, st);
    case 64: return launch_nt<Policy, 64>(P, st);
    case 128: return launch_nt<Policy, 128>(P, st);
    default: return launch_nt<Policy, 256>(P, st);
  }
}

}  // namespace tc
}  // namespace ifd
