// topk.cuh -- per-thread sorted top-KK selection kept entirely in registers.
//
// Order is lexicographic on (key, index): candidates are offered in ascending index order and a new
// candidate only moves past strictly larger keys, so among equal keys the lowest index stays first.  That
// is the tie policy of this library (torch.topk's CPU tie order is implementation-defined).
#pragma once
#include <stdint.h>

namespace ifd {

template <int KK, typename KeyT = float>
struct TopK {
  KeyT key[KK];
  int id[KK];

  __device__ __forceinline__ void init(KeyT inf) {
#pragma unroll
    for (int s = 0; s < KK; ++s) {
      key[s] = inf;
      id[s] = -1;
    }
  }
  __device__ __forceinline__ KeyT worst() const { return key[KK - 1]; }

  // caller guarantees d < worst() (strict) -- keeps the common path to one compare
  __device__ __forceinline__ void insert(KeyT d, int j) {
    key[KK - 1] = d;
    id[KK - 1] = j;
#pragma unroll
    for (int s = KK - 1; s > 0; --s) {
      const bool sw = key[s] < key[s - 1];
      const KeyT ka = key[s - 1], kb = key[s];
      const int ia = id[s - 1], ib = id[s];
      key[s - 1] = sw ? kb : ka;
      key[s] = sw ? ka : kb;
      id[s - 1] = sw ? ib : ia;
      id[s] = sw ? ia : ib;
    }
  }
  __device__ __forceinline__ void offer(KeyT d, int j) {
    if (d < key[KK - 1]) insert(d, j);
  }

  // Same (key, index) order for candidates offered in ARBITRARY index order (grid traversal): the index is part
  // of the comparison instead of being implied by the visiting order.
  __device__ __forceinline__ void offer_lex(KeyT d, int j) {
    if (!(d < key[KK - 1] || (d == key[KK - 1] && j < id[KK - 1]))) return;
    key[KK - 1] = d;
    id[KK - 1] = j;
#pragma unroll
    for (int s = KK - 1; s > 0; --s) {
      const bool sw = key[s] < key[s - 1] || (key[s] == key[s - 1] && id[s] < id[s - 1]);
      const KeyT ka = key[s - 1], kb = key[s];
      const int ia = id[s - 1], ib = id[s];
      key[s - 1] = sw ? kb : ka;
      key[s] = sw ? ka : kb;
      id[s - 1] = sw ? ib : ia;
      id[s] = sw ? ia : ib;
    }
  }
};

}  // namespace ifd
