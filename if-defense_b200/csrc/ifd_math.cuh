// ifd_math.cuh -- the scalar arithmetic of the restoration path, written once.
//
// Everything here is __host__ __device__ so that the exact same source is (a) inlined into the sm_100a
// kernels and (b) compiled by g++ into tests/_mathcheck (test-only) where it is compared against the
// oracle on the CPU before any GPU time is spent.  The product never runs the host instantiation.
//
// Rounding matters on this path (kNN indices must be bit-exact), so every operation whose
// association the reference fixes is spelled with an explicitly rounded primitive that the compiler may
// not contract or reassociate.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define IFD_HD __host__ __device__ __forceinline__
#else
#define IFD_HD inline
#endif

namespace ifd {

// ---- explicitly rounded fp32 primitives (never contracted) ------------------------------------
IFD_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
IFD_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
IFD_HD float sub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}
IFD_HD float fma_rn(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
IFD_HD float div_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
IFD_HD float sqrt_rn(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}

// ---- kNN ranking key: pn_utils.py:76-78 / SURVEY.md appendix A4 --------------------------------
// |x|^2 of a 3-vector: squares rounded separately, summed left to right (torch.sum over a size-3 dim).
IFD_HD float sqnorm3(float x, float y, float z) {
  return add_rn(add_rn(mul_rn(x, x), mul_rn(y, y)), mul_rn(z, z));
}
// torch.matmul with inner dimension 3 on CPU == FMA chain in x, y, z order.
IFD_HD float dot3_chain(float ax, float ay, float az, float bx, float by, float bz) {
  return fma_rn(az, bz, fma_rn(ay, by, mul_rn(ax, bx)));
}
// dist[i][j] = (xx[j] + (-2 * dot_ij)) + xx[i]
IFD_HD float knn_key(float xx_i, float xx_j, float dot_ij) {
  return add_rn(add_rn(xx_j, mul_rn(-2.0f, dot_ij)), xx_i);
}

// ---- repulsion pair: repulsion_loss.py:43-53 / appendix A3 ------------------------------------
struct RepPair {
  float loss;   // (radius - d) * w
  float gcoef;  // d loss / d (x_j) = gcoef * (x_j - x_i);  d loss / d (x_i) = -gcoef * (x_j - x_i)
};
IFD_HD RepPair repulsion_pair(float dx, float dy, float dz, float radius, float h, float eps) {
  const float d2 = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));
  const float d2c = d2 > eps ? d2 : eps;          // torch.max(dist2, eps)
  const float d = sqrt_rn(d2c);
  const float q = div_rn(d, h);
  const float w = expf(-mul_rn(q, q));
  const float rd = sub_rn(radius, d);
  RepPair r;
  r.loss = mul_rn(rd, w);
  // autograd chain: dl/dd = -w + rd * (-w) * 2q / h ;  dd/dd2c = 1/(2d) ;  dd2c/dd2 = [d2 > eps] (1/2 on a tie)
  const float dl_dd = -w - rd * w * (2.0f * q) / h;
  float pass = d2 > eps ? 1.0f : (d2 == eps ? 0.5f : 0.0f);
  r.gcoef = pass * dl_dd / (2.0f * d) * 2.0f;      // times d(d2)/d(diff) = 2 * diff
  return r;
}

// ---- plane coordinate: common.py:235-258 + decoder.py:52-56 + grid_sample(align_corners, border) ------
struct PlaneCoord {
  float u;     // normalised coordinate after the in-place clamps
  float live;  // 1 if the gradient flows to p, 0 if a clamp overwrote the value
};
IFD_HD PlaneCoord plane_coord(float p, float denom /* fp32(1 + padding + 10e-6) */) {
  PlaneCoord r;
  float u = add_rn(div_rn(p, denom), 0.5f);
  r.live = 1.0f;
  if (u >= 1.0f) { u = 0.99999f; r.live = 0.0f; }   // xy_new[xy_new >= 1] = 1 - 10e-6
  if (u < 0.0f)  { u = 0.0f;     r.live = 0.0f; }   // xy_new[xy_new < 0] = 0.0
  r.u = u;
  return r;
}
struct Axis {   // one axis of a bilinear tap set
  int i0;       // floor(ix) clipped into [0, R-1]
  int has1;     // 1 if i0 + 1 <= R-1 (the far corner exists), else its weight is dropped
  float f;      // w = ix - floor(ix): weight of the far corner
  float near_w; // e = 1 - w: weight of the near corner (ATen GridSamplerKernel compute_interp_params)
  float dscale; // d ix / d p including unnormalise (R-1)/2, vgrid 2x, 1/denom, clip and clamp masks
};
IFD_HD Axis axis_setup(PlaneCoord pc, int R, float denom) {
  Axis a;
  const float g = sub_rn(mul_rn(2.0f, pc.u), 1.0f);                       // vgrid = 2.0 * xy - 1.0
  float ix = mul_rn(div_rn(add_rn(g, 1.0f), 2.0f), (float)(R - 1));       // ((g + 1) / 2) * (size - 1)
  float clip = 1.0f;
  if (!(ix > 0.0f)) { ix = 0.0f; clip = 0.0f; }                           // clip_coordinates (border)
  if (!(ix < (float)(R - 1))) { ix = (float)(R - 1); clip = 0.0f; }
  const float fl = floorf(ix);
  a.i0 = (int)fl;
  a.has1 = (a.i0 + 1 <= R - 1) ? 1 : 0;
  a.f = sub_rn(ix, fl);
  a.near_w = sub_rn(1.0f, a.f);
  a.dscale = pc.live * clip * ((float)(R - 1) * 0.5f) * 2.0f / denom;
  return a;
}

// ---- Adam (torch 2.11 _single_tensor_adam, appendix A5) ---------------------------------------
struct AdamStepConst {
  float neg_step_size;   // -(lr / (1 - beta1^t))
  float bc2_sqrt;        // sqrt(1 - beta2^t)
};
IFD_HD void adam_update(float& p, float& m, float& v, float g, float one_minus_b1, float b2, float one_minus_b2,
                        float eps, AdamStepConst c) {
  m = fma_rn(sub_rn(g, m), one_minus_b1, m);                      // exp_avg.lerp_(grad, 1 - beta1)
  v = add_rn(mul_rn(v, b2), mul_rn(mul_rn(one_minus_b2, g), g));  // mul_(beta2).addcmul_(g, g, 1 - beta2)
  const float denom = add_rn(div_rn(sqrt_rn(v), c.bc2_sqrt), eps);
  p = add_rn(p, div_rn(mul_rn(c.neg_step_size, m), denom));       // addcdiv_(exp_avg, denom, -step_size)
}

// ---- BCE-with-logits pieces (opt_defense.py:213-216) ------------------------------------------
IFD_HD float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
IFD_HD float bce_with_logits(float l, float t) {   // ATen: (1 - t) * l + max(-l, 0) + log1p(exp(-|l|))
  const float mx = l < 0.0f ? -l : 0.0f;
  return (1.0f - t) * l + mx + log1pf(expf(-fabsf(l)));
}

// ---- exact, order-independent accumulation of repulsion gradients ----------------------------------
// The scatter-add of pair gradients into neighbours must be (a) bitwise reproducible and (b) precise at
// every magnitude, because Adam is scale-invariant per coordinate: a total gradient of 1e-20 moves a point
// as far as one of 1e-2.  Floating-point atomics fail (a), a single fixed-point word fails (b).  So each
// fp32 term is added *exactly* into a long accumulator of kFxLimbs 64-bit integer limbs at radix 2^32
// (limb l counts units of 2^(-160 + 32 l); a 24-bit mantissa shifted by < 32 bits fits one limb with room
// for 2^8 terms).  Integer addition is associative, so any atomic order gives the same limbs, and the
// limbs are then summed high to low in fp64.
constexpr int kFxLimbs = 6;     // covers 2^-160 .. 2^32
struct FxTerm {
  int limb;        // -1: nothing to add (zero / non-finite)
  long long val;
};
IFD_HD FxTerm fx_term(float g) {
  union { float f; uint32_t u; } cv;
  cv.f = g;
  const uint32_t ex = (cv.u >> 23) & 0xffu;
  uint32_t man = cv.u & 0x7fffffu;
  int e;
  if (ex == 0) { e = -149; } else { man |= 0x800000u; e = (int)ex - 150; }
  FxTerm t;
  const int p = e + 160;
  if (man == 0 || ex == 255 || (p >> 5) >= kFxLimbs) { t.limb = -1; t.val = 0; return t; }
  t.limb = p >> 5;
  const long long v = (long long)((unsigned long long)man << (p & 31));
  t.val = (cv.u >> 31) ? -v : v;
  return t;
}
IFD_HD float fx_value(const long long* limbs) {
  double s = 0.0, w = 1.0;                 // limb l weighs 2^(-160 + 32 l): 2^0 for the top limb
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int l = kFxLimbs - 1; l >= 0; --l) {
    s += (double)limbs[l] * w;
    w *= (1.0 / 4294967296.0);
  }
  return (float)s;
}

}  // namespace ifd
