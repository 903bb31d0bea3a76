// EXPERIMENT, NOT PART OF THE BUILD (round-2 head start for SURVEY.md §8 f1; never run on hardware yet).
//
// The ConvONet encoder's U-Net (ConvONet/src/encoder/unet.py:117-239) as implicit-GEMM 3x3 convolutions on tcgen05,
// obtained by generalising the validated ONet GEMM kernel (csrc/onet.cu: A chunk of the row-owning thread in TMEM, weight
// chunk streamed with cp.async through a double-buffered shared-memory slot, 3xTF32, 12 MMAs per chunk):
//   * a GEMM row is an output PIXEL of a channels-last image [B][H][W][C]; the K loop runs over 9 taps x (Cin / 32)
//     channel chunks; the A chunk of (tap, chunk) is the 128-byte channel slice of the shifted pixel, or zeros outside
//     the image (padding = 1);
//   * two inputs (skip connection + upsampled tensor) are read as one concatenated channel range, so torch.cat goes away;
//   * N = output channels handled by one CTA (32 / 64 / 128), blockIdx.y selects the channel block;
//   * epilogue: bias, optional ReLU, channels-last store;
//   * the same kernel with one tap is the 1x1 convolution and, with N = 4 Cout and a pixel-shuffle store, the 2x2
//     transposed convolution.
// Compile check:  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I if-defense_b200/csrc -I include \
//                      -c tools/experiments/unet_conv3x3.cu -o /tmp/unet_conv3x3.o
// To do when it moves into csrc/: parity test against F.conv2d (fp32, TF32 off) at all five U-Net shapes, maxpool and the
// 2x2 transposed convolution (a plain GEMM with N = 4 Cout and a pixel-shuffle store), bench against cuDNN.
#include "common.cuh"
#include "umma.cuh"

namespace ifd {

constexpr int kConvThreads = 128;
constexpr int kConvChunk = 32;     // channels per K chunk

// W [Cout][Cin][3][3] (torch Conv2d layout) -> per (tap, channel chunk) a K-major image of the [Cout x 32] operand, hi then lo.
// Chunk order: kc = tap * (Cin / 32) + cc.
__global__ void conv3x3_pack_kernel(const float* __restrict__ W, int Cout, int Cin, float* __restrict__ img) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Cout * Cin * 9) return;
  const int tap = e % 9, ci = (e / 9) % Cin, n = e / (9 * Cin);
  const float w = W[e];                                   // e = (n * Cin + ci) * 9 + tap
  const int kc = tap * (Cin / kConvChunk) + ci / kConvChunk, kl = ci % kConvChunk;
  float* chunk = img + (size_t)kc * 2 * Cout * kConvChunk;
  const uint32_t off = umma::img_offset(n, kl, Cout) / 4;
  chunk[off] = __uint_as_float(umma::tf32_hi(w));
  chunk[Cout * kConvChunk + off] = __uint_as_float(umma::tf32_lo(w));
}

// Generic per-pixel GEMM operand: Bm[n][k] given by a functor of the torch tensor -> chunked K-major images (one tap).
// mode 0: Conv2d 1x1 weight [N][K][1][1];  mode 1: ConvTranspose2d weight [K][Cout][2][2] with n = (dy * 2 + dx) * Cout + co.
__global__ void gemm_pack_kernel(const float* __restrict__ W, int N, int K, int mode, int Cout, float* __restrict__ img) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * K) return;
  const int n = e / K, k = e % K;
  float w;
  if (mode == 0) w = W[(size_t)n * K + k];
  else {
    const int q = n / Cout, co = n % Cout;
    w = W[(((size_t)k * Cout + co) * 2 + q / 2) * 2 + q % 2];
  }
  const int kc = k / kConvChunk, kl = k % kConvChunk;
  float* chunk = img + (size_t)kc * 2 * N * kConvChunk;
  const uint32_t off = umma::img_offset(n, kl, N) / 4;
  chunk[off] = __uint_as_float(umma::tf32_hi(w));
  chunk[N * kConvChunk + off] = __uint_as_float(umma::tf32_lo(w));
}

struct ConvArgs {
  const float* in0;   // [B][H][W][C0] channels-last
  const float* in1;   // [B][H][W][C1] or nullptr (C1 = 0): channels C0 .. C0 + C1 - 1 of the concatenation
  const float* img;   // packed weights (conv3x3_pack_kernel), Cin = C0 + C1
  const float* bias;  // [Cout]
  float* out;         // [B][H][W][Cout]
  int B, H, W, C0, C1, Cout, relu;
  int taps;           // 9: 3x3 convolution with padding 1; 1: 1x1 convolution / per-pixel GEMM
  int shuffle;        // 0: out[pixel][Cout]; otherwise Cout = 4 * shuffle and GEMM column q * shuffle + co goes to pixel
                      // (2y + q / 2, 2x + q % 2), channel co of an image [B][2H][2W][shuffle]  (ConvTranspose2d k = 2, s = 2)
};

template <int N>
__global__ void __launch_bounds__(kConvThreads, 2) conv3x3_kernel(const ConvArgs a) {
  constexpr int kSlotFloats = 2 * N * kConvChunk;                   // hi + lo of a [N x 32] block of the chunk
  extern __shared__ float4 smem4[];
  float* bbuf = reinterpret_cast<float*>(smem4);                    // [2][kSlotFloats]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bbuf + 2 * kSlotFloats);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;
  const int nb = blockIdx.y;                                        // output-channel block
  const int Cin = a.C0 + a.C1;
  const int n_chunks = a.taps * (Cin / kConvChunk);
  if (warp == 0) umma::tmem_alloc(tmem_slot, 256);                  // D: columns 0..N-1, A chunks: 128..255
  if (threadIdx.x == 32) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], 1);
    umma::fence_mbar_init();
  }
  // rows [N nb, N nb + N) of chunk kc: in the packed image every 4-wide k group holds its Cout rows contiguously
  // (Cout x 16 bytes); the block is 16 pieces of N x 16 bytes (8 k groups x {hi, lo}), stored back to back.
  auto load_b = [&](int kc, int buf) {
    const float4* src = reinterpret_cast<const float4*>(a.img + (size_t)kc * 2 * a.Cout * kConvChunk);
    const uint32_t dst = umma::smem_u32(bbuf + (size_t)buf * kSlotFloats);
    for (int i = threadIdx.x; i < kSlotFloats / 4; i += kConvThreads) {
      const int piece = i / N, within = i % N;                      // N float4 per piece
      const int sidx = (piece < 8 ? piece * a.Cout : 8 * a.Cout + (piece - 8) * a.Cout) + nb * N + within;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (uint32_t)i), "l"(src + sidx) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_b(0, 0);
  __syncthreads();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_t = tmem + ((uint32_t)(warp * 32) << 16);
  const int M = a.B * a.H * a.W;
  const int row = blockIdx.x * kConvThreads + threadIdx.x;
  const int rowc = min(row, M - 1);
  const int px = rowc % a.W, py = (rowc / a.W) % a.H, pb = rowc / (a.W * a.H);
  const bool leader = threadIdx.x == 0;
  uint32_t parity[2] = {0u, 0u};
  constexpr uint32_t idesc = umma::idesc_tf32(128, N);

  // the 128-byte channel slice of (tap, chunk) for this thread's pixel, or nullptr outside the image
  auto slice = [&](int kc) -> const float4* {
    const int per_tap = Cin / kConvChunk;
    const int tap = kc / per_tap, cc = kc % per_tap;
    const int y = a.taps == 9 ? py + tap / 3 - 1 : py, x = a.taps == 9 ? px + tap % 3 - 1 : px;
    if (y < 0 || y >= a.H || x < 0 || x >= a.W) return nullptr;
    const size_t pix = ((size_t)pb * a.H + y) * a.W + x;
    const int c = cc * kConvChunk;
    const float* p = c < a.C0 ? a.in0 + pix * a.C0 + c : a.in1 + pix * a.C1 + (c - a.C0);
    return reinterpret_cast<const float4*>(p);
  };
  float4 nx[8];
  {
    const float4* p4 = slice(0);
#pragma unroll
    for (int q = 0; q < 8; ++q) nx[q] = p4 ? __ldg(p4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }

#pragma unroll 1
  for (int kc = 0; kc < n_chunks; ++kc) {
    const int buf = kc & 1;
    float x[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      x[4 * q + 0] = nx[q].x; x[4 * q + 1] = nx[q].y; x[4 * q + 2] = nx[q].z; x[4 * q + 3] = nx[q].w;
    }
    if (kc + 1 < n_chunks) {
      const float4* p4 = slice(kc + 1);
#pragma unroll
      for (int q = 0; q < 8; ++q) nx[q] = p4 ? __ldg(p4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (kc >= 1) {                        // MMA kc-1 read the other slot (and its A columns): it must have completed
        umma::mbar_wait(&bars[buf ^ 1], parity[buf ^ 1]);
        parity[buf ^ 1] ^= 1;
        umma::fence_after_sync();
      }
      load_b(kc + 1, buf ^ 1);
    }
    uint32_t u[32];
    const uint32_t a_hi = lane_t + 128 + buf * 64, a_lo = a_hi + 32;
#pragma unroll
    for (int k = 0; k < 32; ++k) u[k] = umma::tf32_hi_fast(x[k]);
    umma::tmem_st32(a_hi, u);
#pragma unroll
    for (int k = 0; k < 32; ++k) u[k] = __float_as_uint(x[k] - __uint_as_float(u[k]));
    umma::tmem_st32(a_lo, u);
    umma::tmem_wait_st();
    if (kc + 1 < n_chunks) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    if (leader) {
      umma::fence_after_sync();
      const uint32_t ta_hi = tmem + 128 + buf * 64, ta_lo = ta_hi + 32;
      const uint32_t sb = umma::smem_u32(bbuf + (size_t)buf * kSlotFloats);
#pragma unroll
      for (int part = 0; part < 3; ++part) {        // lo.hi, hi.lo, hi.hi
        const uint32_t ta = part == 0 ? ta_lo : ta_hi;
        const uint32_t bs = sb + (part == 1 ? (uint32_t)(N * kConvChunk * 4) : 0u);
#pragma unroll
        for (int s = 0; s < 4; ++s)
          umma::mma_tf32_ts(tmem, ta + s * 8, umma::smem_desc_kmajor(bs + s * 2 * (N / 8) * 128, (N / 8) * 128, 128), idesc,
                            (kc | part | s) ? 1u : 0u);
      }
      umma::commit(&bars[buf]);
    }
    __syncwarp();
  }
  // the last commit covers every earlier MMA; chunk counts here are odd (9 x ...) or even: wait on the last one used
  {
    const int last = (n_chunks - 1) & 1;
    umma::mbar_wait(&bars[last], parity[last]);
    umma::fence_after_sync();
  }
  float* orow = a.out + (size_t)rowc * a.Cout + nb * N;
  if (a.shuffle) {                                   // the whole column block lies inside one (dy, dx) quadrant (N <= shuffle)
    const int q = (nb * N) / a.shuffle, co0 = (nb * N) % a.shuffle;
    const size_t opix = ((size_t)pb * 2 * a.H + 2 * py + q / 2) * 2 * a.W + 2 * px + q % 2;
    orow = a.out + opix * a.shuffle + co0;
  }
  const bool live = row < M;
#pragma unroll 1
  for (int cl = 0; cl < N / 32; ++cl) {
    uint32_t d[32];
    umma::tmem_ld32(lane_t + cl * 32, d);
    float y[32];
    const float4* b4 = reinterpret_cast<const float4*>(a.bias + (a.shuffle ? (nb * N) % a.shuffle : nb * N) + cl * 32);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 bv = __ldg(b4 + q);
      y[4 * q + 0] = __uint_as_float(d[4 * q + 0]) + bv.x;
      y[4 * q + 1] = __uint_as_float(d[4 * q + 1]) + bv.y;
      y[4 * q + 2] = __uint_as_float(d[4 * q + 2]) + bv.z;
      y[4 * q + 3] = __uint_as_float(d[4 * q + 3]) + bv.w;
    }
    if (a.relu) {
#pragma unroll
      for (int k = 0; k < 32; ++k) y[k] = fmaxf(y[k], 0.0f);
    }
    if (live) {
      float4* o4 = reinterpret_cast<float4*>(orow + cl * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) o4[q] = make_float4(y[4 * q + 0], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

// 2x2 max pooling, channels-last (unet.py:77: MaxPool2d(kernel_size=2, stride=2))
__global__ void maxpool2_cl_kernel(const float* __restrict__ in, int B, int H, int W, int C, float* __restrict__ out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // one float4 of an output pixel
  const int C4 = C / 4, Ho = H / 2, Wo = W / 2;
  if (e >= (size_t)B * Ho * Wo * C4) return;
  const int c4 = (int)(e % C4);
  const size_t pix = e / C4;
  const int x = (int)(pix % Wo), y = (int)((pix / Wo) % Ho), b = (int)(pix / ((size_t)Wo * Ho));
  const float4* src = reinterpret_cast<const float4*>(in);
  auto at = [&](int yy, int xx) { return __ldg(src + (((size_t)b * H + yy) * W + xx) * C4 + c4); };
  const float4 p = at(2 * y, 2 * x), q = at(2 * y, 2 * x + 1), r = at(2 * y + 1, 2 * x), s = at(2 * y + 1, 2 * x + 1);
  reinterpret_cast<float4*>(out)[e] = make_float4(fmaxf(fmaxf(p.x, q.x), fmaxf(r.x, s.x)), fmaxf(fmaxf(p.y, q.y), fmaxf(r.y, s.y)),
                                                  fmaxf(fmaxf(p.z, q.z), fmaxf(r.z, s.z)), fmaxf(fmaxf(p.w, q.w), fmaxf(r.w, s.w)));
}

template <int N>
static int launch_conv(const ConvArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)2 * 2 * N * kConvChunk * sizeof(float) + 64;
  IFD_CUDA_TRY(set_max_dyn_smem((const void*)conv3x3_kernel<N>, smem));
  const int M = a.B * a.H * a.W;
  conv3x3_kernel<N><<<dim3((M + kConvThreads - 1) / kConvThreads, a.Cout / N), kConvThreads, smem, st>>>(a);
  IFD_LAUNCH_CHECK("conv3x3_kernel");
  return IFD_OK;
}

// in0 / in1 channels-last, weights packed by conv3x3_pack_kernel; C0, C1, Cout multiples of 32
int conv3x3_cl(const float* in0, const float* in1, const float* img, const float* bias, float* out, int B, int H, int W, int C0,
               int C1, int Cout, int relu, cudaStream_t st) {
  IFD_REQUIRE(in0 && img && bias && out && B > 0 && H > 0 && W > 0, "conv3x3_cl: bad arguments");
  IFD_REQUIRE(C0 % 32 == 0 && C1 % 32 == 0 && C0 > 0 && (C1 == 0 || in1) && Cout % 32 == 0 && Cout > 0, "conv3x3_cl: channels must be multiples of 32");
  ConvArgs a{in0, in1, img, bias, out, B, H, W, C0, C1, Cout, relu, 9, 0};
  if (Cout % 128 == 0) return launch_conv<128>(a, st);
  if (Cout % 64 == 0) return launch_conv<64>(a, st);
  return launch_conv<32>(a, st);
}

// 1x1 convolution (unet.py conv_final): weights packed by gemm_pack_kernel with N = Cout
int conv1x1_cl(const float* in, const float* img, const float* bias, float* out, int B, int H, int W, int Cin, int Cout, cudaStream_t st) {
  IFD_REQUIRE(in && img && bias && out && Cin % 32 == 0 && Cout % 32 == 0 && Cin > 0 && Cout > 0, "conv1x1_cl: bad arguments");
  ConvArgs a{in, nullptr, img, bias, out, B, H, W, Cin, 0, Cout, 0, 1, 0};
  return Cout % 128 == 0 ? launch_conv<128>(a, st) : Cout % 64 == 0 ? launch_conv<64>(a, st) : launch_conv<32>(a, st);
}

// ConvTranspose2d(Cin, Cout, kernel_size = 2, stride = 2) (unet.py upconv2x2): one GEMM with N = 4 Cout and a pixel-shuffle
// store; out [B][2H][2W][Cout].  Weights packed by upconv_pack_kernel.
int upconv2x2_cl(const float* in, const float* img, const float* bias, float* out, int B, int H, int W, int Cin, int Cout, cudaStream_t st) {
  IFD_REQUIRE(in && img && bias && out && Cin % 32 == 0 && Cout % 32 == 0 && Cin > 0 && Cout > 0, "upconv2x2_cl: bad arguments");
  ConvArgs a{in, nullptr, img, bias, out, B, H, W, Cin, 0, 4 * Cout, 0, 1, Cout};
  return Cout % 128 == 0 ? launch_conv<128>(a, st) : Cout % 64 == 0 ? launch_conv<64>(a, st) : launch_conv<32>(a, st);
}

}  // namespace ifd

// C entry points for tools/experiments/test_unet_conv3x3.py (device pointers)
extern "C" int exp_conv3x3_pack(const float* W, int Cout, int Cin, float* img, void* stream) {
  const int n = Cout * Cin * 9;
  ifd::conv3x3_pack_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, Cout, Cin, img);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
extern "C" int exp_conv3x3(const float* in0, const float* in1, const float* img, const float* bias, float* out, int B, int H, int W,
                           int C0, int C1, int Cout, int relu, void* stream) {
  return ifd::conv3x3_cl(in0, in1, img, bias, out, B, H, W, C0, C1, Cout, relu, (cudaStream_t)stream);
}
extern "C" int exp_gemm_pack(const float* W, int N, int K, int mode, int Cout, float* img, void* stream) {
  ifd::gemm_pack_kernel<<<(N * K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, N, K, mode, Cout, img);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
extern "C" int exp_conv1x1(const float* in, const float* img, const float* bias, float* out, int B, int H, int W, int Cin, int Cout,
                           void* stream) {
  return ifd::conv1x1_cl(in, img, bias, out, B, H, W, Cin, Cout, (cudaStream_t)stream);
}
extern "C" int exp_upconv2x2(const float* in, const float* img, const float* bias, float* out, int B, int H, int W, int Cin, int Cout,
                             void* stream) {
  return ifd::upconv2x2_cl(in, img, bias, out, B, H, W, Cin, Cout, (cudaStream_t)stream);
}
extern "C" int exp_maxpool2(const float* in, int B, int H, int W, int C, float* out, void* stream) {
  const size_t n = (size_t)B * (H / 2) * (W / 2) * (C / 4);
  ifd::maxpool2_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, B, H, W, C, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
