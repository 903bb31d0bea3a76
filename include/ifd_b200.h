/*
 * ifd_b200.h -- C ABI of libifd_b200.so: the B200 (sm_100a) implementation of IF-Defense's
 * optimisation-based restoration hot path.
 *
 * The reference (Wuziyi616/IF-Defense) has no FFI/plugin layer: its seams are Python call sites.
 * Each entry point below names the reference call site it replaces (file:line under the reference
 * tree).  INTEGRATION.md shows the ctypes binding a reference maintainer would add at that seam.
 *
 * Conventions
 *   - every pointer is caller-owned DEVICE memory unless the parameter name ends in `_host`;
 *   - no allocation inside the library except in the *_host convenience calls; scratch comes from
 *     the caller-provided workspace (query the size with the matching *_workspace_bytes);
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*); no internal host
 *     synchronisation except in the *_host calls;
 *   - return value: 0 on success, <0 on error (IFD_ERR_*); ifd_last_error() returns a thread-local
 *     message.  The Python shim raises RuntimeError, the reference's only error convention
 *     (ConvONet/defense/repulsion_loss.py:37, ConvONet/defense/SOR.py:64);
 *   - indices are int32 (the reference uses int64 tensors; values are identical);
 *   - results are bitwise reproducible run to run (no floating-point atomics anywhere).
 */
#ifndef IFD_B200_H_
#define IFD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IFD_OK 0
#define IFD_ERR_INVALID (-1)     /* bad argument */
#define IFD_ERR_CUDA (-2)        /* CUDA runtime error (message has the cudaError string) */
#define IFD_ERR_UNSUPPORTED (-3) /* shape/config outside what the kernels are built for */
#define IFD_ERR_WORKSPACE (-4)   /* workspace too small */

#define IFD_ABI_VERSION 1

typedef void* ifd_stream_t; /* cudaStream_t */

const char* ifd_last_error(void);
int ifd_abi_version(void);
/* Compute capability of the current device as major*10+minor, or <0.  The kernels are sm_100a-only. */
int ifd_device_cc(void);

/* ------------------------------------------------------------------------------------------------
 * Geometry kernels
 * ---------------------------------------------------------------------------------------------- */

/* knn_point (ConvONet/defense/pn_utils.py:64-83), DGCNN knn (baselines/model/dgcnn.py:7-13),
 * PointConv knn_point (baselines/model/pointconv.py:104-116).
 * x: [B][N][C] row-major float32 (C = 3 for point clouds; feature-space kNN up to C = 128).
 * Ranks key(i,j) = (|x_j|^2 + (-2 x_i.x_j)) + |x_i|^2 evaluated in the reference's association
 * (dot product = FMA chain over c, squares rounded separately and summed left to right), ascending;
 * exact ties resolve to the lowest index.  The first `drop_first` entries of each sorted row are dropped
 * (knn_point: drop_first = 1 "assuming it is self"; DGCNN: 0).  idx_out: [B][N][k] int32;
 * key_out (optional): [B][N][k] float32 ranked keys.  k + drop_first <= 32. */
int ifd_knn(const float* x, int B, int N, int C, int k, int drop_first, int32_t* idx_out, float* key_out,
            ifd_stream_t stream);

/* RepulsionLoss.get_repulsion_loss fwd + bwd (ConvONet/defense/repulsion_loss.py:18-54) fused with
 * its knn_point and index_points (pn_utils.py:6-23,64-83).
 *   loss_out[b]       = mean_{K,k} (radius - d) * exp(-(d/h)^2),  d = sqrt(max(|x_j - x_i|^2, eps))
 *   grad_xyz_out      = d( sum_b grad_loss[b] * loss[b] ) / d xyz   (grad_loss == NULL means all ones)
 * idx_out (optional): [B][K][k].  Workspace: ifd_knn_repulsion_workspace_bytes(B, K). */
size_t ifd_knn_repulsion_workspace_bytes(int B, int K);
int ifd_knn_repulsion(const float* xyz, int B, int K, int k, double radius, double h, double eps,
                      const float* grad_loss, int32_t* idx_out, float* loss_out, float* grad_xyz_out,
                      void* workspace, size_t workspace_bytes, ifd_stream_t stream);

/* farthest_point_sample (baselines/model/pointnet2.py:53-74, ConvONet/defense/pn_utils.py:26-48).
 * The reference draws the first index with torch.randint; here it is data: start_idx[B].
 * idx_out: [B][npoint]. */
int ifd_fps(const float* xyz, int B, int N, int npoint, const int32_t* start_idx, int32_t* idx_out,
            ifd_stream_t stream);

/* query_ball_point (baselines/model/pointnet2.py:77-98): for each of S centres the first `nsample`
 * indices (ascending) with expanded-form d^2 <= radius_sq, padded with the first hit.  radius_sq is
 * fp32(radius ** 2) exactly as torch casts the Python scalar (the caller squares in double, then rounds).
 * xyz [B][N][3], new_xyz [B][S][3], idx_out [B][S][nsample].  A centre with no hit yields N everywhere
 * (the reference's behaviour: it would then fail in index_points). */
int ifd_ball_query(const float* xyz, const float* new_xyz, int B, int N, int S, float radius_sq, int nsample,
                   int32_t* idx_out, ifd_stream_t stream);

/* SORDefense.outlier_removal (ConvONet/defense/SOR.py:22-49): float64 expanded-form kNN-k mean squared
 * distance per point, threshold mean + alpha * std (unbiased).  keep_out [B][K] uint8 (1 = kept),
 * value_out (optional) [B][K] float64. */
int ifd_sor(const float* xyz, int B, int K, int k, double alpha, uint8_t* keep_out, double* value_out,
            ifd_stream_t stream);

/* preprocess_pc (ConvONet/opt_defense.py:114-130, the numpy part; ONet/remesh_defense.py:104-110) for a batch, fused with the
 * ragged selection that follows SOR (opt_defense.py:86-111).  keep [B][K] uint8 from ifd_sor, or NULL (all points).
 * out [B][K][3]: the counts_out[b] kept points of cloud b in input order, minus their float32 mean (summed in point order like
 * np.mean(axis=0)), divided by the largest bounding-box extent, times padding_scale; the remaining rows are zero.  Bitwise equal
 * to the numpy statements.  The random subset / init draws that follow stay with the caller. */
int ifd_preprocess_pc(const float* xyz, const uint8_t* keep, int B, int K, float padding_scale, float* out,
                      int32_t* counts_out, ifd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * ConvONet decoder (LocalDecoder, 3 feature planes)
 * ---------------------------------------------------------------------------------------------- */

/* Packed decoder parameters in kernel layout (float32, all Linear weights TRANSPOSED to [in][out]):
 *   fc_p.W^T [3][H], fc_p.b [H],
 *   n_blocks x { fc_c[i].W^T [C][H], fc_c[i].b [H], blocks[i].fc_0.W^T [H][H], .b [H], blocks[i].fc_1.W^T [H][H], .b [H] },
 *   fc_out.w [H], fc_out.b [1]
 * Names are the reference state_dict keys under `decoder.`. 16001 floats for the
 * shipped config (C = H = 32, n_blocks = 5). */
size_t ifd_convonet_decoder_nfloats(int C, int H, int n_blocks);

/* encode_inputs() output -> kernel layout.  The reference hands the decoder a dict of
 * [B][C][R][R] NCHW planes (ConvONet/src/encoder/pointnet.py:68-86); the kernels read channels-last
 * [B][R][R][C] so that one bilinear tap is one 128-byte line. */
int ifd_planes_nchw_to_cl(const float* nchw, float* cl, int B, int C, int R, ifd_stream_t stream);

/* ConvolutionalOccupancyNetwork.decode(p, c).logits (ConvONet/src/conv_onet/models/__init__.py:67-77 ->
 * LocalDecoder.forward, models/decoder.py:69-95, sample_plane_feature :50-57, normalize_coordinate
 * src/common.py:235-258, ResnetBlockFC src/layers.py:39-48).
 * planes_cl: [3][B][R][R][C] in the order xz, xy, yz.  xyz [B][K][3].  logits_out [B][K]. */
int ifd_convonet_decode_fwd(const float* planes_cl, const float* dec_weights, const float* xyz, int B, int K,
                            int R, int C, int H, int n_blocks, double padding, float* logits_out,
                            ifd_stream_t stream);

/* Backward of the above w.r.t. xyz only (the model is frozen: ConvONet/opt_defense.py:72-73).
 * grad_logits [B][K] -> grad_xyz_out [B][K][3].  Recomputes the forward (nothing is saved). */
int ifd_convonet_decode_bwd(const float* planes_cl, const float* dec_weights, const float* xyz,
                            const float* grad_logits, int B, int K, int R, int C, int H, int n_blocks,
                            double padding, float* grad_xyz_out, ifd_stream_t stream);

/* The fused "decode + BCE-with-logits gradient" step of the loop as a seam of its own (opt_defense.py:212-216
 * followed by backward to xyz): grad_xyz_out = d( K * mean_{B_ref,K} BCEWithLogits(logit, occ_target) ) / d xyz,
 * computed by the decode kernel selected with decode_kernel (see ifd_opt_params).  logits are not returned.
 * Workspace: ifd_convonet_opt_workspace_bytes(B, K). */
int ifd_convonet_decode_bce_grad(const float* planes_cl, const float* dec_weights, const float* xyz, int B, int K,
                                 int R, int C, int H, int n_blocks, double padding, double occ_target, int B_ref,
                                 int decode_kernel, float* grad_xyz_out, void* workspace, size_t workspace_bytes,
                                 ifd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * ConvONet encoder operators (LocalPoolPointnet; run once per batch, before the loop)
 * ---------------------------------------------------------------------------------------------- */

/* coordinate2index(normalize_coordinate(p, plane), R) for the three planes (ConvONet/src/common.py:235-258,300-315;
 * called at ConvONet/src/encoder/pointnet.py:131-143): bins_out [3][B][T] int32 in the order xz, xy, yz,
 * bin = floor(u0 * R) + R * floor(u1 * R). */
int ifd_plane_bins(const float* xyz, int B, int T, int R, double padding, int32_t* bins_out, ifd_stream_t stream);

/* pool_local (ConvONet/src/encoder/pointnet.py:104-122): for each of the P planes torch_scatter.scatter_max into the
 * bins, gather back to the points, summed over the planes in order.  src, out: [B][T][C] point-major (the reference
 * permutes to [B,C,T] around the call); bins [P][B][T]. */
int ifd_scatter_max_gather(const float* src, const int32_t* bins, int P, int B, int T, int C, int nbins, float* out,
                           ifd_stream_t stream);

/* generate_plane_features up to the U-Net (ConvONet/src/encoder/pointnet.py:68-80): torch_scatter.scatter_mean of the
 * point features into the R^2 bins of one plane, empty bins 0.  src [B][T][C]; bins [B][T]; plane_out [B][nbins][C]
 * channels-last (the memory of a [B,C,R,R] torch tensor in channels_last format).  Members of a bin are summed in
 * ascending point order: reproducible, and equal to a sequential scatter_add (the reference's CUDA atomics are not). */
int ifd_scatter_mean_cl(const float* src, const int32_t* bins, int B, int T, int C, int nbins, float* plane_out,
                        ifd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Dense layers of the encoders on the tensor cores (tcgen05, 3xTF32 = fp32 semantics; csrc/tc_gemm.cuh)
 *
 * One warp-specialised GEMM engine  Out[m][n] = epilogue(sum_k A(m,k) W[n][k])  serves nn.Linear with fused ReLU / bias /
 * residual (ResnetBlockFC: ConvONet/src/layers.py:39-48; fc_pos, fc_c: ConvONet/src/encoder/pointnet.py:41-45,
 * ONet/im2mesh/encoder/pointnet.py:85-113), nn.Conv2d(3x3, padding 1) on channels-last tensors, nn.ConvTranspose2d(2, stride 2)
 * and the 1x1 convolution of the U-Net (ConvONet/src/encoder/unet.py:117-239).  Weights are packed once per model into
 * the engine's image format with ifd_tc_pack.
 * ---------------------------------------------------------------------------------------------- */

/* Size in floats of the packed image of W [N][sum(seg_widths)] (every K segment is padded to a multiple of 32). */
size_t ifd_tc_packed_floats(int N, const int* seg_widths, int n_seg);
/* W: device, row-major [N][sum(seg_widths)] with row stride ldw -> out (device, ifd_tc_packed_floats floats, 16-byte aligned). */
int ifd_tc_pack(const float* W, int N, const int* seg_widths, int n_seg, int ldw, float* out, ifd_stream_t stream);

typedef struct ifd_tc_linear_args {
  int32_t M, N;              /* rows, output features */
  int32_t n_seg;             /* 1..3 input segments, concatenated along K in this order (torch.cat(..., dim = -1) on read) */
  const float* a_ptr[3];     /* [M][a_ld] row-major device pointers */
  int32_t a_ld[3];           /* row strides in floats */
  int32_t a_width[3];        /* columns used of each segment (= the widths given to ifd_tc_pack) */
  int32_t a_relu[3];         /* 1: the segment passes through ReLU on read (fc(relu(x))) */
  int32_t a_blocked[3];      /* 1: the segment is in the warp-transposed layout (see ifd_tc_blocked_floats), width % 4 == 0 */
  int32_t out_blocked;       /* 1: out is in the warp-transposed layout (N % 4 == 0, no resid) */
  int32_t a_group[3];        /* > 0: the segment holds one row per group of a_group consecutive rows (a per-cloud vector
                                expanded over the cloud's points on read); 0: one row per output row */
  const float* wimg;         /* packed weights */
  const float* bias;         /* [N] or NULL */
  const float* resid;        /* [M][ld_resid] added to the result, or NULL */
  int32_t ld_resid;
  int32_t relu_out;          /* 1: ReLU on the result */
  float* out;                /* [M][ld_out] */
  int32_t ld_out;
  /* ConvTranspose2d(kernel 2, stride 2) epilogue: rows are the B*H*W input pixels (channels-last), N = 4 * cout with
   * n = (i * 2 + j) * cout + co, out is [B][2H][2W][cout] channels-last, bias is [cout].  0 = plain Linear. */
  int32_t shuffle_cout, shuffle_H, shuffle_W;
} ifd_tc_linear_args;
int ifd_tc_linear(const ifd_tc_linear_args* args, ifd_stream_t stream);

/* The engine reads and writes tensors thread-per-row; intermediate tensors that only the engine touches (the U-Net's feature
 * maps) can therefore live in a WARP-TRANSPOSED layout in which those accesses are coalesced: rows (pixels) in blocks of 32,
 * inside a block the C / 4 float4 column groups one after the other, each holding its 32 rows contiguously.  A tensor of M
 * rows x C columns (C % 4 == 0) occupies ifd_tc_blocked_floats(M, C) floats; element (m, c) sits at float offset
 * (((m / 32) * (C / 4) + c / 4) * 32 + m % 32) * 4 + c % 4. */
size_t ifd_tc_blocked_floats(long long M, int C);

/* out[g][c] = max_t x[g * T + t][c]: the global max-pool of ResnetPointnet (ONet/im2mesh/encoder/pointnet.py:103,110). */
int ifd_group_max(const float* x, int groups, int T, int C, float* out, ifd_stream_t stream);

/* nn.Conv2d(C0 + C1 -> Cout, 3, padding = 1) (+ ReLU) on channels-last device tensors.  Input = the channel concatenation
 * [src0 | src1] (torch.cat((from_up, from_down), 1), unet.py:187; src1 = NULL, C1 = 0 for a single input) of
 * [B][H][W][C] tensors; with pool != 0 the input is F.max_pool2d(src0, 2, 2) of a [B][2H][2W][C0] tensor, taken on read
 * (unet.py:151).  wimg = ifd_tc_pack of the weight as [Cout][9 * (C0 + C1)] with k = (ky * 3 + kx) * (C0 + C1) + ci.
 * Channel counts are multiples of 32.  out [B][H][W][Cout].  layout: bit 0 / 1 / 2 = src0 / src1 / out are in the
 * warp-transposed layout (rows = pixels in [b][y][x] order) instead of plain channels-last. */
int ifd_tc_conv3x3(const float* src0, int C0, const float* src1, int C1, int B, int H, int W, int pool, const float* wimg,
                   const float* bias, int relu_out, int Cout, float* out, int layout, ifd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * The restoration loop
 * ---------------------------------------------------------------------------------------------- */

/* Everything optimize_points reads from argparse / module constants
 * (ConvONet/opt_defense.py:33-46,59,207; defense/repulsion_loss.py:9-10). */
typedef struct ifd_opt_params {
  int32_t n_steps;       /* iterations + 1 (opt_defense.py:210) -> 201 */
  int32_t step0;         /* Adam steps already taken (0 for a fresh run; >0 resumes with adam_m/adam_v) */
  int32_t B_ref;         /* batch size of the reference call this cloud belongs to: the losses are batch
                            means, so every gradient carries 1/B_ref */
  int32_t knn_k;         /* 5 */
  int32_t normalize_out; /* 1: apply normalize_batch_pc (opt_defense.py:76-83) after the last step */
  int32_t want_stats;    /* 1: fill stats_out every 100 steps as the reference prints (:229-236) */
  /* Python floats travel as doubles and are rounded to fp32 where torch rounds them. */
  double lr;             /* 1e-3 */
  double beta1, beta2;   /* 0.9, 0.999 */
  double adam_eps;       /* 1e-8 */
  double occ_target;     /* cfg['test']['threshold'] = 0.2 */
  double rep_weight;     /* 500 */
  double rep_radius, rep_h, rep_eps; /* 0.07, 0.03, 1e-12 */
  double padding;        /* cfg['data']['padding'] = 0.1 */
  int32_t decode_kernel; /* which decode kernel the loop launches -- 0: the production default (= 5);
                            1: v1, thread-per-point fp32 SIMT (the step-level seam; from-scratch kNN every step);
                            2: v2, fp32 SIMT, cooperative gather, 2 points/thread, FFMA2, generic layer bodies;
                            3, 4: retired (earlier tensor-core kernels: 512- / 256-point CTAs with 30 dependent tensor-core
                               round trips per step);
                            5: v5, ResNet-MLP on tcgen05 tensor cores (3xTF32, A in TMEM, fp32-class accuracy), 256-point
                               CTAs, two per SM, weight images streamed per stage by TMA bulk copies, the fc_c layers off
                               the dependent chain (their products ride in the accumulator of the preceding fc_1 / in a
                               second accumulator of the dgrad): 22 tensor-core round trips per step */
  int32_t tail_kernel;   /* what follows the decode in a step -- 0: the production default, one fused kernel per cloud
                            (grid kNN + repulsion gather + Adam, K <= 1024, knn_k <= 7); 1: first-generation kernels
                            (brute-force kNN + exact long accumulator, separate Adam), any K <= 12000 */
} ifd_opt_params;

void ifd_opt_params_default(ifd_opt_params* p);

/* optimize_points (ConvONet/opt_defense.py:182-239): n_steps x { decode -> K*mean BCEWithLogits(logit,
 * target) + rep_weight*mean repulsion -> d/dxyz -> Adam } then centre + unit-sphere normalise.
 * xyz [B][K][3] is updated in place (the reference also mutates opt_points in place).
 * adam_m / adam_v: optional [B][K][3] state in/out (NULL: zero-initialised scratch in the workspace).
 * stats_out: optional [(n_steps-1)/100 + 1][4] float64 = {loss, occ_loss, rep_loss, mean sigmoid(logit)}
 * over the B clouds passed (batch means use B_ref). */
size_t ifd_convonet_opt_workspace_bytes(int B, int K);
int ifd_convonet_opt(const float* planes_cl, const float* dec_weights, float* xyz, float* adam_m, float* adam_v,
                     int B, int K, int R, int C, int H, int n_blocks, const ifd_opt_params* params,
                     double* stats_out, void* workspace, size_t workspace_bytes, ifd_stream_t stream);

/* The tail of one iteration as a seam of its own (opt_defense.py:219-228: rep_loss = repulsion_loss(p).mean() *
 * rep_weight, its backward into p, then opt.step()): one fused kernel per cloud -- grid-accelerated knn_point
 * (identical indices to the brute-force scan), repulsion pair terms gathered per point in a canonical order, Adam.
 * xyz / adam_m / adam_v [B][K][3] are updated in place with the gradient g_occ + d(rep_loss)/dp; Adam's bias
 * corrections use t = step_index + 1.  nbr [B][K][8] int32 receives the k+1 ranked neighbours per point (column 0
 * is the column knn_point drops); with warm != 0 it must hold the lists of the previous call on the same points
 * moved by a small step (the search is bounded by their current distances).  loss_sum_out (optional) [B]: sum of
 * the pair losses per cloud; rep_grad_out (optional) [B][K][3]: summed pair-loss gradients before the
 * rep_weight / (B_ref K k) factor.  K <= 1024, knn_k <= 7. */
int ifd_opt_tail_step(float* xyz, float* adam_m, float* adam_v, const float* g_occ, int32_t* nbr, int warm, int B,
                      int K, const ifd_opt_params* params, int step_index, float* loss_sum_out, float* rep_grad_out,
                      ifd_stream_t stream);

/* The same for a sequence of batches on DEVICE buffers (the loop of defend_point_cloud over batches, opt_defense.py:
 * 272-312, with the encoder outputs already in kernel layout): batch j uses planes_cl[j] / xyz[j] (in place), all of B
 * clouds.  The loops of up to four consecutive batches run side by side on internal streams forked from and joined back
 * into `stream` (joined on error paths too) -- every launch of a loop depends on the one before it and none fills the
 * machine; in this mode the per-cloud tail runs as ONE CTA per cloud (half the SMs per launch).  workspace:
 * ifd_convonet_opt_batches_workspace_bytes(B, K) (one 256-byte-rounded ifd_convonet_opt_workspace_bytes per loop in
 * flight).  Results are bit-identical to n_batches calls of ifd_convonet_opt. */
size_t ifd_convonet_opt_batches_workspace_bytes(int B, int K);   /* = lanes x the 256-byte-rounded single-loop workspace */
int ifd_convonet_opt_batches(int n_batches, const float* const* planes_cl, const float* dec_weights, float* const* xyz,
                             int B, int K, int R, int C, int H, int n_blocks, const ifd_opt_params* params,
                             void* workspace, size_t workspace_bytes, ifd_stream_t stream);


/* ------------------------------------------------------------------------------------------------
 * The 'grid' variant of ConvONet: ONE feature volume sampled trilinearly instead of three planes sampled bilinearly
 * (LocalDecoder.sample_grid_feature, ConvONet/src/conv_onet/models/decoder.py:59-67,72-73; normalize_3d_coordinate,
 * ConvONet/src/common.py:260-276; LocalPoolPointnet.generate_grid_features, ConvONet/src/encoder/pointnet.py:88-99).
 * No shipped config selects it (configs/ has no grid_resolution); it is what BASELINE.json's north_star describes.
 * volume_cl: [B][R][R][R][C] float32 channels-last, indexed [z][y][x] (coordinate2index '3d' = x + R (y + R z)); C = H = 32.
 * Same arguments and semantics as the plane entry points above; thread-per-point fp32 kernels.
 * ---------------------------------------------------------------------------------------------- */
int ifd_grid_bins(const float* xyz, int B, int T, int R, double padding, int32_t* bins_out /* [B][T] */, ifd_stream_t stream);
int ifd_convonet_grid_decode_fwd(const float* volume_cl, const float* dec_weights, const float* xyz, int B, int K, int R, int C,
                                 int H, int n_blocks, double padding, float* logits_out, ifd_stream_t stream);
int ifd_convonet_grid_decode_bwd(const float* volume_cl, const float* dec_weights, const float* xyz, const float* grad_logits,
                                 int B, int K, int R, int C, int H, int n_blocks, double padding, float* grad_xyz_out,
                                 ifd_stream_t stream);
int ifd_convonet_grid_opt(const float* volume_cl, const float* dec_weights, float* xyz, float* adam_m, float* adam_v, int B, int K,
                          int R, int C, int H, int n_blocks, const ifd_opt_params* params, double* stats_out, void* workspace,
                          size_t workspace_bytes, ifd_stream_t stream);

/* Host-buffer convenience call (the end-to-end seam): planes in the reference's NCHW layout
 * [3][B][C][R][R], weights, xyz are HOST pointers; does H2D, layout conversion, the loop, D2H of xyz and
 * synchronises.  Device scratch is cached per thread between calls and released by ifd_release_cache(). */
int ifd_convonet_opt_host(const float* planes_nchw_host, const float* dec_weights_host, float* xyz_host,
                          int B, int K, int R, int C, int H, int n_blocks, const ifd_opt_params* params,
                          double* stats_out_host);
/* The same for a sequence of batches -- the loop of defend_point_cloud over batches (ConvONet/opt_defense.py:272-312):
 * planes_nchw_host[j] / xyz_host[j] are the HOST buffers of batch j (all batches B clouds; pinned memory makes the
 * copies asynchronous).  Three device buffer slots and two run streams: uploads, layout conversions and downloads
 * overlap the loops, and the loops of two consecutive batches run side by side (one launch fills 128 of the 148 SMs);
 * every batch still pays its own H2D and D2H.  Synchronises before
 * returning.  Results are bit-identical to n_batches calls of ifd_convonet_opt_host. */
int ifd_convonet_opt_host_batches(int n_batches, const float* const* planes_nchw_host, const float* dec_weights_host,
                                  float* const* xyz_host, int B, int K, int R, int C, int H, int n_blocks,
                                  const ifd_opt_params* params);
void ifd_release_cache(void);

/* ------------------------------------------------------------------------------------------------
 * ONet decoder (DecoderCBatchNorm, c_dim 512, hidden 256, z_dim 0) and its restoration loop
 * ---------------------------------------------------------------------------------------------- */

/* Packed ONet decoder parameters (float32; names are the reference state_dict keys under `decoder.`):
 *   fc_p.weight [256][3], fc_p.bias [256],
 *   11 x { conv_gamma.weight [256][512], conv_gamma.bias [256], conv_beta.weight [256][512], conv_beta.bias [256],
 *          bn.running_mean [256], bn.running_var [256] }  in the order block0.bn_0, block0.bn_1, ..., block4.bn_1, bn
 *   10 x { weight [256][256], bias [256] }                in the order block0.fc_0, block0.fc_1, ..., block4.fc_1
 *   fc_out.weight [256], fc_out.bias [1] */
size_t ifd_onet_decoder_nfloats(void);
size_t ifd_onet_workspace_bytes(int B, int K);

/* Once per batch (the reference recomputes gamma(c), beta(c) in every CBN call, ONet/im2mesh/layers.py:226-242;
 * in eval mode they are constants of the batch): packs the tensor-core weight images and folds the 11 CBNs for
 * the latent codes c [B][512] = encode_inputs() (ONet/im2mesh/encoder/pointnet.py:85-113) into the workspace. */
int ifd_onet_prepare(const float* dec_weights, const float* c, int B, int K, void* workspace, size_t workspace_bytes,
                     ifd_stream_t stream);

/* OccupancyNetwork.decode(p, z, c).logits, z_dim = 0 (ONet/opt_defense.py:212; ONet/im2mesh/onet/models/decoder.py:
 * 115-133; CResnetBlockConv1d ONet/im2mesh/layers.py:98-107).  Needs ifd_onet_prepare on the same workspace with the same
 * B; the prepared state sits at the head of the workspace and does not depend on K, so K may differ from the prepare call
 * as long as workspace_bytes >= ifd_onet_workspace_bytes(B, K) (chunked evaluation of one shape). */
int ifd_onet_decode_fwd(const float* dec_weights, const float* xyz, int B, int K, float* logits_out, void* workspace,
                        size_t workspace_bytes, ifd_stream_t stream);
/* Forward + backward w.r.t. xyz for grad_logits [B][K] -> grad_xyz_out [B][K][3]. */
int ifd_onet_decode_bwd(const float* dec_weights, const float* xyz, const float* grad_logits, int B, int K,
                        float* grad_xyz_out, void* workspace, size_t workspace_bytes, ifd_stream_t stream);

/* optimize_points of ONet/opt_defense.py:182-239 (same loop as ConvONet's; decode(p, z, c) with the CBN decoder).
 * c [B][512]; xyz [B][K][3] in place; calls ifd_onet_prepare itself. */
int ifd_onet_opt(const float* dec_weights, const float* c, float* xyz, float* adam_m, float* adam_v, int B, int K,
                 const ifd_opt_params* params, double* stats_out, void* workspace, size_t workspace_bytes,
                 ifd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * ONet-Mesh tail: marching cubes over the occupancy lattice, surface sampling
 * ---------------------------------------------------------------------------------------------- */

/* libmcubes.marching_cubes(volume, isovalue) (ONet/im2mesh/utils/libmcubes/mcubes.pyx:20-25 -> pywrapper.cpp:90-127 ->
 * marchingcubes.h:23-189) on the device, in two calls because the output sizes are data dependent.
 * volume: device array [nx][ny][nz], dtype 0 = float32, 1 = float64 (the reference hands float64 in; float32 values
 * are widened exactly).  pad = 1 adds the one-voxel border of pad_value that Generator3D.extract_mesh adds with np.pad
 * (ONet/im2mesh/onet/generation.py:172-173: -1e6) without materialising it.  Vertices [V][3] float64 and faces [F][3]
 * int64 come out in the reference's order with the reference's bits.
 * ifd_mc_count fills the workspace (ifd_mc_workspace_bytes) and synchronises `stream` to return the counts to the
 * host; ifd_mc_emit then needs the same arguments and the untouched workspace.  to_box != 0 applies
 * generation.py:177-183 to the vertices (-0.5, -1, / (n - 1) of the unpadded lattice, box_size * (v - 0.5)). */
size_t ifd_mc_workspace_bytes(int nx, int ny, int nz, int pad);
int ifd_mc_count(const void* volume, int dtype, int nx, int ny, int nz, int pad, double pad_value, double isovalue,
                 void* workspace, size_t workspace_bytes, long long* n_verts_host, long long* n_faces_host,
                 ifd_stream_t stream);
int ifd_mc_emit(const void* volume, int dtype, int nx, int ny, int nz, int pad, double pad_value, double isovalue,
                int to_box, double box_size, const void* workspace, size_t workspace_bytes, double* verts_out,
                long long* faces_out, ifd_stream_t stream);

/* trimesh.sample.sample_surface(mesh, count) (un-vendored trimesh 3.7.7; call site ONet/remesh_defense.py:157) with
 * the random numbers handed in (the library never draws any): uniforms [count][3] in [0, 1): column 0 picks the face
 * by cumulative area (searchsorted, left), columns 1-2 are the two edge lengths, reflected into the triangle when their
 * sum exceeds 1.  verts [V][3] float64, faces [F][3] int64 (device); xyz_out [count][3] float64; face_out [count] or
 * null.  A mesh without faces is IFD_ERR_INVALID (the reference's IndexError, remesh_defense.py:160). */
size_t ifd_sample_surface_workspace_bytes(long long n_faces);
/* numpy.cumsum over a device array of non-negative float64 values, in place and bit for bit (the cumulative-area prefix that
 * trimesh.sample.sample_surface searches: one left-to-right chain of rounded additions, reproduced by an integer prefix sum
 * per binade of the running sum).  total_out (device, optional): the last element. */
int ifd_cumsum_f64(double* data, long long n, double* total_out, ifd_stream_t stream);
int ifd_sample_surface(const double* verts, long long n_verts, const long long* faces, long long n_faces,
                       const double* uniforms, int count, double* xyz_out, long long* face_out, void* workspace,
                       size_t workspace_bytes, ifd_stream_t stream);

/* MISE, the octree controller of Generator3D.generate_from_latent (ONet/im2mesh/utils/libmise/mise.pyx:33-369; driven by
 * ONet/im2mesh/onet/generation.py:113-130), with its state in a caller-provided device workspace:
 *   ifd_mise_init       MISE(resolution_0, depth, threshold).__cinit__: the (resolution_0 + 1)^3 coarse points exist
 *   ifd_mise_query      .query(): the lattice points (x, y, z int64, on the (resolution_0 << depth) + 1 lattice) that exist
 *                       and have no value yet.  points_out == NULL only counts; synchronises `stream` for the count
 *   ifd_mise_update     .update(points, values): sets the values, then subdivides every leaf voxel whose known points hold
 *                       a value >= threshold and a value <= threshold.  A point that is not in the grid is
 *                       IFD_ERR_INVALID "Point not in grid!" (the reference's ValueError); synchronises for that check
 *   ifd_mise_to_dense   .to_dense(): [n][n][n] float64, unevaluated points completed along x, then y, then z
 * The result equals the reference's for the same evaluation function (set semantics; the order of the queried points
 * is unspecified). */
size_t ifd_mise_workspace_bytes(int resolution0, int depth);
int ifd_mise_init(int resolution0, int depth, void* workspace, size_t workspace_bytes, ifd_stream_t stream);
int ifd_mise_query(int resolution0, int depth, void* workspace, size_t workspace_bytes, long long* points_out,
                   long long capacity, long long* count_host, ifd_stream_t stream);
int ifd_mise_update(int resolution0, int depth, double threshold, void* workspace, size_t workspace_bytes,
                    const long long* points, const double* values, long long count, ifd_stream_t stream);
int ifd_mise_to_dense(int resolution0, int depth, void* workspace, size_t workspace_bytes, double* dense_out,
                      ifd_stream_t stream);

/* Number of kernel launches issued by this library on the calling thread since the last reset. */
long long ifd_launch_count(int reset);

/* Per-kernel device timing for the roofline report (bench.py).  While enabled, every launch inside the
 * restoration loop is bracketed by CUDA events on the launching stream.  ifd_profile_read synchronises the
 * device, sums the elapsed time per kernel kind, clears the record and returns the number of kinds written:
 * 0 = decode (gather + MLP fwd/bwd), 1 = kNN + repulsion, 2 = Adam, 3 = everything else. */
/* Tensor-core self test: D[128][32] = A[128][32] . Bm[32][32]^T (Bm is [n][k]) through the tcgen05 / TMEM /
 * 3xTF32 path of the decode kernels (one layer as a tensor-core round trip).  Device pointers. */
int ifd_selftest_umma(const float* A, const float* Bm, float* D, ifd_stream_t stream);

/* Test instrumentation.  key 1: capacity (0..16, default 16) of the per-point inbox of non-mutual in-edges in the
 * fused tail kernel; lowering it forces the ordered-scan fallback that hubs take.  Results do not depend on it.
 * key 2: number of loops ifd_convonet_opt_batches and the host pipeline run side by side (1..4, default 4; size the workspace
 * with ifd_convonet_opt_batches_workspace_bytes AFTER setting it).
 * key 3: 0 = enqueue the loop of ifd_convonet_opt as direct launches, 1 (default) = replay the cached CUDA graph (same bits).
 * key 4: cluster barrier in front of cloud_step_kernel's first remote store: 0 none, 1 release / acquire, 2 (default) relaxed.
 * key 5: form of the fused tail -- 0 (default) by context: a 2-CTA cluster per cloud for a loop of at most 74 clouds that runs
 *        alone, one CTA per cloud for larger batches and for loops that run side by side; 1 / 2 force the one-CTA / cluster form.
 *        Same bits either way.
 * key 6: CTAs per thread-block cluster of the GEMM engine (1, 2 (default) or 4): weight chunks are TMA-multicast across it.
 * key 7: 1 (default) = decode v5 keeps d c / d xyz of its forward gather and reads it back, 0 = it gathers the texels twice.
 * key 8: 1 (default) = the ten ONet decoder layers of a direction run as one chain launch, 0 = ten launches (same bits). */
void ifd_test_hook(int key, int value);

#define IFD_PROFILE_KINDS 4
void ifd_profile_enable(int on);
int ifd_profile_read(double* ms_out, long long* launches_out);

#ifdef __cplusplus
}
#endif
#endif /* IFD_B200_H_ */
