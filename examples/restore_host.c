/* Minimal C caller of the library through the host-buffer seam: restores B clouds of K points with ConvONet-Opt.
 *
 *   gcc -std=c99 -Iinclude examples/restore_host.c -Lif-defense_b200 -lifd_b200 -Wl,-rpath,$PWD/if-defense_b200 -o restore_host
 *   ./restore_host planes.f32 weights.f32 init_xyz.f32 B K out_xyz.f32
 *
 * planes.f32   [3][B][32][64][64] float32, the encoder's feature planes in the reference's NCHW layout
 *              (generator.model.encode_inputs, ConvONet/opt_defense.py:300; planes in the order xz, xy, yz)
 * weights.f32  ifd_convonet_decoder_nfloats(32, 32, 5) float32, packed as include/ifd_b200.h describes
 * init_xyz.f32 [B][K][3] float32, the output of init_points (opt_defense.py:149-179; the library never draws random numbers)
 * Needs a B200 (sm_100a): without one the call fails with an error message, there is no CPU fallback. */
#include <stdio.h>
#include <stdlib.h>

#include "ifd_b200.h"

static float* read_f32(const char* path, size_t n) {
  float* p = (float*)malloc(n * sizeof(float));
  FILE* f = fopen(path, "rb");
  if (!p || !f || fread(p, sizeof(float), n, f) != n) {
    fprintf(stderr, "cannot read %zu floats from %s\n", n, path);
    exit(2);
  }
  fclose(f);
  return p;
}

int main(int argc, char** argv) {
  if (argc != 7) {
    fprintf(stderr, "usage: %s planes.f32 weights.f32 init_xyz.f32 B K out_xyz.f32\n", argv[0]);
    return 2;
  }
  const int B = atoi(argv[4]), K = atoi(argv[5]);
  const int R = 64, C = 32, H = 32, n_blocks = 5;
  if (B <= 0 || K <= 0) return 2;
  float* planes = read_f32(argv[1], (size_t)3 * B * C * R * R);
  float* weights = read_f32(argv[2], ifd_convonet_decoder_nfloats(C, H, n_blocks));
  float* xyz = read_f32(argv[3], (size_t)B * K * 3);

  ifd_opt_params p;
  ifd_opt_params_default(&p); /* 201 Adam steps, lr 1e-3, rep_weight 500, threshold 0.2: the shipped defaults */
  p.B_ref = B;                /* the reference batch these clouds belong to (the 1/B of the batch-mean losses) */
  const int rc = ifd_convonet_opt_host(planes, weights, xyz, B, K, R, C, H, n_blocks, &p, NULL);
  if (rc != IFD_OK) {
    fprintf(stderr, "ifd_convonet_opt_host failed (%d): %s\n", rc, ifd_last_error());
    return 1;
  }
  FILE* out = fopen(argv[6], "wb");
  if (!out || fwrite(xyz, sizeof(float), (size_t)B * K * 3, out) != (size_t)B * K * 3) return 1;
  fclose(out);
  ifd_release_cache();
  free(planes);
  free(weights);
  free(xyz);
  return 0;
}
