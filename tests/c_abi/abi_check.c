/* Plain C99 consumer of include/ifd_b200.h: proves the boundary is a C ABI (no C++ or torch types in the signatures),
 * that every entry point links, and that argument validation answers before any CUDA call (so it runs without a GPU).
 * Built and run by tests/test_capi_symbols.py. */
#include <stdio.h>
#include <string.h>

#include "ifd_b200.h"

int main(void) {
  int bad = 0;
  ifd_opt_params p;
  ifd_opt_params_default(&p);
  if (p.n_steps != 201 || p.knn_k != 5) { printf("defaults\n"); ++bad; }
  if (ifd_abi_version() != 1) { printf("abi version\n"); ++bad; }
  if (ifd_knn(NULL, 1, 8, 3, 2, 1, NULL, NULL, NULL) != IFD_ERR_INVALID) { printf("knn validation\n"); ++bad; }
  if (strstr(ifd_last_error(), "null") == NULL) { printf("error text\n"); ++bad; }
  if (ifd_mc_workspace_bytes(129, 129, 129, 1) == 0) { printf("mc workspace\n"); ++bad; }
  if (ifd_mise_workspace_bytes(32, 2) == 0 || ifd_mise_workspace_bytes(32, 9) != 0) { printf("mise workspace\n"); ++bad; }
  {
    long long nv = -1, nf = -1;
    if (ifd_mc_count(NULL, 1, 4, 4, 4, 0, 0.0, 0.0, NULL, 0, &nv, &nf, NULL) != IFD_ERR_INVALID) { printf("mc validation\n"); ++bad; }
  }
  if (ifd_sample_surface(NULL, 0, NULL, 0, NULL, 0, NULL, NULL, NULL, 0, NULL) != IFD_ERR_INVALID) { printf("sample validation\n"); ++bad; }
  if (ifd_preprocess_pc(NULL, NULL, 1, 8, 0.9f, NULL, NULL, NULL) != IFD_ERR_INVALID) { printf("preprocess validation\n"); ++bad; }
  if (ifd_convonet_opt_workspace_bytes(64, 1024) == 0 || ifd_onet_workspace_bytes(1, 1024) == 0) { printf("workspaces\n"); ++bad; }
  printf(bad ? "FAILED %d\n" : "abi ok\n", bad);
  return bad;
}
