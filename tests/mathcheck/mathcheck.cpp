// TEST INFRASTRUCTURE ONLY.  Host instantiation of the __host__ __device__ arithmetic in
// if-defense_b200/csrc/{ifd_math,convonet_point}.cuh, driven serially exactly as the kernels in restore.cu
// drive it.  It lets the CPU test-suite compare the product's arithmetic with the oracle before any GPU time
// is spent.  The product never loads this library.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../if-defense_b200/csrc/convonet_point.cuh"
#include "../../if-defense_b200/csrc/grid_point.cuh"

using namespace ifd;

static void fx_add(long long* limbs, float g) {
  const FxTerm t = fx_term(g);
  if (t.limb >= 0) limbs[t.limb] += t.val;
}

extern "C" {

// mode 0: forward only; 1: backward with grad_logits; 2: BCE gradient (target, ginv)
void mc_convonet_decode(const float* W, const float* planes_cl, const float* xyz, int B, int K, int R, int n_blocks,
                        double padding, int mode, const float* grad_logits, float target, float ginv,
                        float* logits_out, float* grad_out) {
  const float denom = (float)(1.0 + padding + 10e-6);
  const size_t plane_sz = (size_t)R * R * 32;
#pragma omp parallel for schedule(static)
  for (int pi = 0; pi < B * K; ++pi) {
    const int b = pi / K;
    const float* const planes[3] = {planes_cl + ((size_t)0 * B + b) * plane_sz, planes_cl + ((size_t)1 * B + b) * plane_sz,
                                    planes_cl + ((size_t)2 * B + b) * plane_sz};
    ConvPoint<32> pt;
    const float logit = pt.forward(W, planes, xyz[pi * 3], xyz[pi * 3 + 1], xyz[pi * 3 + 2], R, denom, n_blocks);
    if (logits_out) logits_out[pi] = logit;
    if (mode == 0) continue;
    const float gl = mode == 1 ? grad_logits[pi] : (sigmoidf_(logit) - target) * ginv;
    float gp[3];
    pt.backward(W, planes, gl, R, n_blocks, gp);
    grad_out[pi * 3] = gp[0];
    grad_out[pi * 3 + 1] = gp[1];
    grad_out[pi * 3 + 2] = gp[2];
  }
}

// The 'grid' variant (grid_point.cuh): vol_cl [B][R][R][R][32] channels-last.  mode as above.
void mc_convonet_grid_decode(const float* W, const float* vol_cl, const float* xyz, int B, int K, int R, int n_blocks,
                             double padding, int mode, const float* grad_logits, float target, float ginv,
                             float* logits_out, float* grad_out) {
  const float denom = (float)(1.0 + padding + 10e-4);
  const size_t vol_sz = (size_t)R * R * R * 32;
#pragma omp parallel for schedule(static)
  for (int pi = 0; pi < B * K; ++pi) {
    const float* vol = vol_cl + (size_t)(pi / K) * vol_sz;
    GridPoint<32> pt;
    const float logit = pt.forward(W, vol, xyz[pi * 3], xyz[pi * 3 + 1], xyz[pi * 3 + 2], R, denom, n_blocks);
    if (logits_out) logits_out[pi] = logit;
    if (mode == 0) continue;
    const float gl = mode == 1 ? grad_logits[pi] : (sigmoidf_(logit) - target) * ginv;
    float gp[3];
    pt.backward(W, vol, gl, R, n_blocks, gp);
    grad_out[pi * 3] = gp[0];
    grad_out[pi * 3 + 1] = gp[1];
    grad_out[pi * 3 + 2] = gp[2];
  }
}

// knn_repulsion_kernel + repulsion_finalize_kernel, serially.  acc must hold B*K*3 zeros on entry when
// `finalize` is 0 (loop use); with finalize = 1 it is private scratch.
void mc_knn_repulsion(const float* xyz, int B, int K, int k, float radius, float h, float eps, const float* grad_loss,
                      int32_t* idx_out, float* loss_out, float* grad_out, long long* acc_io) {
  std::vector<long long> own;
  long long* acc = acc_io;
  if (!acc) {
    own.assign((size_t)B * K * 3 * kFxLimbs, 0);
    acc = own.data();
  }
  for (int b = 0; b < B; ++b) {
    const float* c = xyz + (size_t)b * K * 3;
    std::vector<float> xx(K);
    for (int j = 0; j < K; ++j) xx[j] = sqnorm3(c[j * 3], c[j * 3 + 1], c[j * 3 + 2]);
    std::vector<float> lq(K, 0.f);
    std::vector<int> nbr((size_t)K * (k + 1));
#pragma omp parallel for schedule(static)
    for (int q = 0; q < K; ++q) {
      std::vector<std::pair<float, int>> top(k + 1, {INFINITY, -1});
      for (int j = 0; j < K; ++j) {
        const float d = knn_key(xx[q], xx[j], dot3_chain(c[q * 3], c[q * 3 + 1], c[q * 3 + 2], c[j * 3], c[j * 3 + 1], c[j * 3 + 2]));
        if (d < top[k].first) {
          top[k] = {d, j};
          for (int s = k; s > 0 && top[s].first < top[s - 1].first; --s) std::swap(top[s], top[s - 1]);
        }
      }
      for (int s = 0; s <= k; ++s) nbr[(size_t)q * (k + 1) + s] = top[s].second;
    }
    long long* ab = acc + (size_t)b * K * 3 * kFxLimbs;
    for (int q = 0; q < K; ++q) {
      float sx = 0.f, sy = 0.f, sz = 0.f, ls = 0.f;
      for (int s = 1; s <= k; ++s) {
        const int j = nbr[(size_t)q * (k + 1) + s];
        if (idx_out) idx_out[((size_t)b * K + q) * k + (s - 1)] = j;
        const float dx = sub_rn(c[j * 3], c[q * 3]), dy = sub_rn(c[j * 3 + 1], c[q * 3 + 1]), dz = sub_rn(c[j * 3 + 2], c[q * 3 + 2]);
        const RepPair p = repulsion_pair(dx, dy, dz, radius, h, eps);
        ls += p.loss;
        const float gx = p.gcoef * dx, gy = p.gcoef * dy, gz = p.gcoef * dz;
        fx_add(ab + ((size_t)j * 3 + 0) * kFxLimbs, gx);
        fx_add(ab + ((size_t)j * 3 + 1) * kFxLimbs, gy);
        fx_add(ab + ((size_t)j * 3 + 2) * kFxLimbs, gz);
        sx -= gx; sy -= gy; sz -= gz;
      }
      fx_add(ab + ((size_t)q * 3 + 0) * kFxLimbs, sx);
      fx_add(ab + ((size_t)q * 3 + 1) * kFxLimbs, sy);
      fx_add(ab + ((size_t)q * 3 + 2) * kFxLimbs, sz);
      lq[q] = ls;
    }
    if (loss_out) {
      // same reduction tree as the kernel: chunks of 128 queries, xor-shuffle within a warp, warps left to right
      float total = 0.f;
      for (int c0 = 0; c0 < K; c0 += 128) {
        float part = 0.f;
        for (int w = 0; w < 4; ++w) {
          float v[32];
          for (int l = 0; l < 32; ++l) { const int q = c0 + w * 32 + l; v[l] = q < K ? lq[q] : 0.f; }
          for (int o = 16; o > 0; o >>= 1) { float t[32]; for (int l = 0; l < 32; ++l) t[l] = v[l] + v[l ^ o]; memcpy(v, t, sizeof(v)); }
          part += v[0];
        }
        total += part;
      }
      loss_out[b] = total / (float)(K * k);
    }
    if (grad_out) {
      const float go = grad_loss ? grad_loss[b] : 1.0f;
      for (int e = 0; e < K * 3; ++e) grad_out[(size_t)b * K * 3 + e] = fx_value(ab + (size_t)e * kFxLimbs) * (go / (float)(K * k));
    }
  }
}

void mc_adam(float* xyz, float* m, float* v, const float* g, int n, double lr, double beta1, double beta2, double eps, int t) {
  AdamStepConst sc;
  sc.neg_step_size = (float)(-(lr / (1.0 - pow(beta1, (double)t))));
  sc.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)t));
  for (int e = 0; e < n; ++e) adam_update(xyz[e], m[e], v[e], g[e], (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, sc);
}

void mc_normalize(float* xyz, int B, int K) {
  for (int b = 0; b < B; ++b) {
    float* c = xyz + (size_t)b * K * 3;
    double s[3] = {0, 0, 0};
    for (int i = 0; i < K; ++i) for (int a = 0; a < 3; ++a) s[a] += c[i * 3 + a];
    float ctr[3] = {(float)(s[0] / K), (float)(s[1] / K), (float)(s[2] / K)};
    float mx = 0.f;
    for (int i = 0; i < K; ++i) {
      for (int a = 0; a < 3; ++a) c[i * 3 + a] = sub_rn(c[i * 3 + a], ctr[a]);
      mx = fmaxf(mx, sqrt_rn(add_rn(add_rn(mul_rn(c[i * 3], c[i * 3]), mul_rn(c[i * 3 + 1], c[i * 3 + 1])), mul_rn(c[i * 3 + 2], c[i * 3 + 2]))));
    }
    for (int e = 0; e < K * 3; ++e) c[e] = div_rn(c[e], mx);
  }
}

// ifd_convonet_opt, serially (same order of operations as restore.cu).  trace_xyz (optional): [n_trace][B*K*3]
// snapshots after the steps listed in trace_steps.
void mc_convonet_opt(const float* W, const float* planes_cl, float* xyz, int B, int K, int R, int n_blocks, int n_steps,
                     int B_ref, int knn_k, double lr, double beta1, double beta2, double adam_eps, double target,
                     double rep_weight, double radius, double h, double eps, double padding, int normalize,
                     const int* trace_steps, int n_trace, float* trace_xyz) {
  const size_t n = (size_t)B * K * 3;
  std::vector<float> g(n), m(n, 0.f), v(n, 0.f);
  std::vector<long long> acc(n * kFxLimbs, 0);
  const float ginv = (float)K / (float)((long long)B_ref * K);
  const float rep_coef = ((float)rep_weight / (float)B_ref) / (float)(K * knn_k);
  for (int i = 0; i < n_steps; ++i) {
    mc_convonet_decode(W, planes_cl, xyz, B, K, R, n_blocks, padding, 2, nullptr, (float)target, ginv, nullptr, g.data());
    if (rep_weight > 0.0) {
      mc_knn_repulsion(xyz, B, K, knn_k, (float)radius, (float)h, (float)eps, nullptr, nullptr, nullptr, nullptr, acc.data());
      for (size_t e = 0; e < n; ++e) {
        g[e] = g[e] + fx_value(acc.data() + e * kFxLimbs) * rep_coef;
        for (int l = 0; l < kFxLimbs; ++l) acc[e * kFxLimbs + l] = 0;
      }
    }
    mc_adam(xyz, m.data(), v.data(), g.data(), (int)n, lr, beta1, beta2, adam_eps, i + 1);
    for (int t = 0; t < n_trace; ++t)
      if (trace_steps[t] == i) memcpy(trace_xyz + (size_t)t * n, xyz, n * sizeof(float));
  }
  if (normalize) mc_normalize(xyz, B, K);
}

}  // extern "C"
