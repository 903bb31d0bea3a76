// mesh.cu -- the ONet-Mesh tail on the device: marching cubes over the occupancy lattice and area-weighted surface
// sampling.
//
// Reference: ONet/im2mesh/onet/generation.py:165-186 (extract_mesh: pad with -1e6, libmcubes.marching_cubes, undo the
// half-cell shift and the padding, scale to the box), ONet/im2mesh/utils/libmcubes/marchingcubes.h:23-189 (the
// sequential algorithm and its vertex sharing), marchingcubes.cpp:288-326 (vertex placement),
// ONet/remesh_defense.py:151-158 (trimesh.sample.sample_surface of the mesh).
//
// The reference walks the cells in (i, j, k) order, appends a vertex the first time an edge is met and remembers the
// indices of the three edges at a cell's far corner for its neighbours.  The same numbering falls out of a prefix sum:
// a cell "creates" the vertices of its far-corner edges 6, 5, 10 (and, on the i/j/k == 0 faces, of the edges nobody
// before it could have created -- duplicates included, as in the reference), so
//     index(vertex) = (# vertices created by earlier cells) + (position inside the cell's creation order)
// and a neighbour's index is a lookup of that neighbour's prefix.  Output order and every bit of the coordinates
// (fp64, the reference's association, no FMA contraction) equal the sequential code; HBM-bound streaming passes.
#include "common.cuh"

namespace ifd {

#include "mc_tables.inc"
__device__ const unsigned long long kMcTri[256] = {IFD_MC_TRI_WORDS};
__device__ const unsigned short kMcEdges[256] = {IFD_MC_EDGE_MASKS};

constexpr int kMcBlock = 1024;  // cells per scan block

struct McGrid {
  const void* vol;
  int is_f64;
  int nx, ny, nz;       // dimensions of the array that is handed in
  int pad;              // 1: a one-voxel border of pad_value is added around it (np.pad(..., 1, 'constant'))
  double pad_value, iso;
  int px, py, pz;       // dimensions after padding
  int cx, cy, cz;       // cells
  long long ncells;
};

__device__ __forceinline__ double mc_value(const McGrid& g, int x, int y, int z) {
  x -= g.pad;
  y -= g.pad;
  z -= g.pad;
  if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz) return g.pad_value;
  const size_t o = ((size_t)x * g.ny + y) * g.nz + z;
  return g.is_f64 ? static_cast<const double*>(g.vol)[o] : (double)static_cast<const float*>(g.vol)[o];
}

__device__ __forceinline__ void mc_corners(const McGrid& g, int i, int j, int k, double v[8]) {
  v[0] = mc_value(g, i, j, k);
  v[1] = mc_value(g, i + 1, j, k);
  v[2] = mc_value(g, i + 1, j + 1, k);
  v[3] = mc_value(g, i, j + 1, k);
  v[4] = mc_value(g, i, j, k + 1);
  v[5] = mc_value(g, i + 1, j, k + 1);
  v[6] = mc_value(g, i + 1, j + 1, k + 1);
  v[7] = mc_value(g, i, j + 1, k + 1);
}

// edges whose vertex THIS cell appends (marchingcubes.h:69-178): always 6, 5, 10; the others only on the low faces
__device__ __forceinline__ unsigned mc_created_mask(int i, int j, int k) {
  unsigned m = 0x040u | 0x020u | 0x400u;
  if (j == 0 || k == 0) m |= 0x001u;
  if (k == 0) m |= 0x002u | 0x004u;
  if (i == 0 || k == 0) m |= 0x008u;
  if (j == 0) m |= 0x010u | 0x200u;
  if (i == 0) m |= 0x080u | 0x800u;
  if (i == 0 || j == 0) m |= 0x100u;
  return m;
}

__device__ __forceinline__ int mc_tri_count(unsigned long long w) {
  int n = 0;
#pragma unroll
  for (int s = 0; s < 15; ++s) n += ((w >> (4 * s)) & 0xF) != 0xF;      // entries are contiguous from nibble 0
  return n;
}

// pass 1: case index and counts per cell, sums per block of kMcBlock cells (low 32 bits vertices, high 32 indices)
__global__ void __launch_bounds__(kMcBlock) mc_classify_kernel(McGrid g, uint8_t* __restrict__ cases,
                                                               uint8_t* __restrict__ counts,
                                                               unsigned long long* __restrict__ block_sums) {
  __shared__ unsigned long long warp_sums[kMcBlock / 32];
  const long long c = (long long)blockIdx.x * kMcBlock + threadIdx.x;
  unsigned long long mine = 0;
  if (c < g.ncells) {
    const int k = (int)(c % g.cz), j = (int)((c / g.cz) % g.cy), i = (int)(c / ((long long)g.cz * g.cy));
    double v[8];
    mc_corners(g, i, j, k, v);
    unsigned ci = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m) ci |= (v[m] <= g.iso ? 1u : 0u) << m;
    const int nv = __popc((unsigned)kMcEdges[ci] & mc_created_mask(i, j, k));
    const int ni = mc_tri_count(kMcTri[ci]);
    cases[c] = (uint8_t)ci;
    counts[c] = (uint8_t)(nv | (ni << 4));
    mine = (unsigned long long)nv | ((unsigned long long)ni << 32);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned long long s = warp_sums[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
  }
}

// pass 2: exclusive scan of the block sums in place (one CTA; a few thousand entries), totals to totals[0]
__global__ void __launch_bounds__(1024) mc_scan_blocks_kernel(unsigned long long* __restrict__ block_sums, int n,
                                                              unsigned long long* __restrict__ totals) {
  __shared__ unsigned long long warp_sums[32];
  __shared__ unsigned long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int e = base + threadIdx.x;
    const unsigned long long x = e < n ? block_sums[e] : 0;
    unsigned long long inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned long long w = warp_sums[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_sums[threadIdx.x] = w;     // inclusive over warps
    }
    __syncthreads();
    const unsigned long long before = carry_s + (threadIdx.x >= 32 ? warp_sums[(threadIdx.x >> 5) - 1] : 0);
    if (e < n) block_sums[e] = before + inc - x;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = before + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[0] = carry_s;
}

// pass 3: per-cell bases (vertices created / indices emitted by all earlier cells)
__global__ void __launch_bounds__(kMcBlock) mc_bases_kernel(long long ncells, const uint8_t* __restrict__ counts,
                                                            const unsigned long long* __restrict__ block_sums,
                                                            uint32_t* __restrict__ vbase, uint32_t* __restrict__ ibase) {
  __shared__ unsigned long long warp_sums[kMcBlock / 32];
  const long long c = (long long)blockIdx.x * kMcBlock + threadIdx.x;
  const unsigned cnt = c < ncells ? counts[c] : 0;
  const unsigned long long x = (unsigned long long)(cnt & 15) | ((unsigned long long)(cnt >> 4) << 32);
  unsigned long long inc = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
    if ((threadIdx.x & 31) >= o) inc += y;
  }
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = inc;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned long long w = warp_sums[threadIdx.x];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, w, o);
      if (threadIdx.x >= o) w += y;
    }
    warp_sums[threadIdx.x] = w;
  }
  __syncthreads();
  const unsigned long long ex = block_sums[blockIdx.x] + (threadIdx.x >= 32 ? warp_sums[(threadIdx.x >> 5) - 1] : 0) + inc - x;
  if (c < ncells) {
    vbase[c] = (uint32_t)(ex & 0xffffffffull);
    ibase[c] = (uint32_t)(ex >> 32);
  }
}

// marchingcubes.cpp:288-295: (x2 - x1) * (iso - f1) / (f2 - f1) + x1, midpoint when f1 == f2; every operation rounded
__device__ __forceinline__ double mc_interp(double iso, double f1, double f2, double x1, double x2) {
  if (f2 == f1) return __ddiv_rn(__dadd_rn(x2, x1), 2.0);
  return __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(x2, x1), __dsub_rn(iso, f1)), __dsub_rn(f2, f1)), x1);
}

struct McXform {   // generation.py:177-183: v -= 0.5; v -= 1; v /= (n - 1); v = box * (v - 0.5)   (numpy fp64, one op each)
  int on;
  double dx, dy, dz, box;
};

__device__ __forceinline__ double mc_xform1(double v, double d, double box) {
  v = __dsub_rn(v, 0.5);
  v = __dsub_rn(v, 1.0);
  v = __ddiv_rn(v, d);
  return __dmul_rn(box, __dsub_rn(v, 0.5));
}

// pass 4: vertices and triangle indices
__global__ void __launch_bounds__(256) mc_emit_kernel(McGrid g, McXform xf, const uint8_t* __restrict__ cases,
                                                      const uint32_t* __restrict__ vbase, const uint32_t* __restrict__ ibase,
                                                      double* __restrict__ verts, long long* __restrict__ indices) {
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  if (c >= g.ncells) return;
  const unsigned ci = cases[c];
  const unsigned edges = kMcEdges[ci];
  if (edges == 0) return;
  const int k = (int)(c % g.cz), j = (int)((c / g.cz) % g.cy), i = (int)(c / ((long long)g.cz * g.cy));
  double v[8];
  mc_corners(g, i, j, k, v);
  // marchingcubes.h:40-55 with lower = 0, upper = n - 1: dx = 1, x = 0 + 1 * i + 1 / 2
  const double x = (double)i + 0.5, xd = (double)(i + 1) + 0.5;
  const double y = (double)j + 0.5, yd = (double)(j + 1) + 0.5;
  const double z = (double)k + 0.5, zd = (double)(k + 1) + 0.5;
  const unsigned created = edges & mc_created_mask(i, j, k);
  long long idx[12];
  uint32_t next = vbase[c];
  const long long sz = 1, sy = g.cz, sx = (long long)g.cz * g.cy;   // cell strides
  // index of slot s (0: edge 6, 1: edge 5, 2: edge 10) of the cell at linear index n
  auto shared_slot = [&](long long n, int s) -> long long {
    const unsigned en = kMcEdges[cases[n]];
    const unsigned before = s == 0 ? 0u : (s == 1 ? (en & 0x040u) : (en & 0x060u));
    return (long long)vbase[n] + __popc(before);
  };
  auto put = [&](int e, double px, double py, double pz) {
    idx[e] = next;
    double* o = verts + (size_t)next * 3;
    if (xf.on) {
      px = mc_xform1(px, xf.dx, xf.box);
      py = mc_xform1(py, xf.dy, xf.box);
      pz = mc_xform1(pz, xf.dz, xf.box);
    }
    o[0] = px;
    o[1] = py;
    o[2] = pz;
    ++next;
  };
  // creation order of marchingcubes.h: 6, 5, 10, then 0, 1, 2, 3, 4, 7, 8, 9, 11
  if (edges & 0x040u) put(6, mc_interp(g.iso, v[6], v[7], xd, x), yd, zd);
  if (edges & 0x020u) put(5, xd, mc_interp(g.iso, v[5], v[6], y, yd), zd);
  if (edges & 0x400u) put(10, xd, __dadd_rn(y, 1.0), mc_interp(g.iso, v[2], v[6], z, zd));     // "y + dx" (:89)
  if (edges & 0x001u) {
    if (created & 0x001u) put(0, mc_interp(g.iso, v[0], v[1], x, xd), y, z);
    else idx[0] = shared_slot(c - sy - sz, 0);
  }
  if (edges & 0x002u) {
    if (created & 0x002u) put(1, xd, mc_interp(g.iso, v[1], v[2], y, yd), z);
    else idx[1] = shared_slot(c - sz, 1);
  }
  if (edges & 0x004u) {
    if (created & 0x004u) put(2, mc_interp(g.iso, v[2], v[3], xd, x), yd, z);
    else idx[2] = shared_slot(c - sz, 0);
  }
  if (edges & 0x008u) {
    if (created & 0x008u) put(3, x, mc_interp(g.iso, v[3], v[0], yd, y), z);
    else idx[3] = shared_slot(c - sx - sz, 1);
  }
  if (edges & 0x010u) {
    if (created & 0x010u) put(4, mc_interp(g.iso, v[4], v[5], x, xd), y, zd);
    else idx[4] = shared_slot(c - sy, 0);
  }
  if (edges & 0x080u) {
    if (created & 0x080u) put(7, x, mc_interp(g.iso, v[7], v[4], yd, y), zd);
    else idx[7] = shared_slot(c - sx, 1);
  }
  if (edges & 0x100u) {
    if (created & 0x100u) put(8, x, y, mc_interp(g.iso, v[0], v[4], z, zd));
    else idx[8] = shared_slot(c - sx - sy, 2);
  }
  if (edges & 0x200u) {
    if (created & 0x200u) put(9, xd, y, mc_interp(g.iso, v[1], v[5], z, zd));
    else idx[9] = shared_slot(c - sy, 2);
  }
  if (edges & 0x800u) {
    if (created & 0x800u) put(11, x, yd, mc_interp(g.iso, v[3], v[7], z, zd));
    else idx[11] = shared_slot(c - sx, 2);
  }
  const unsigned long long w = kMcTri[ci];
  long long* o = indices + ibase[c];
#pragma unroll
  for (int s = 0; s < 15; ++s) {
    const int e = (int)((w >> (4 * s)) & 0xF);
    if (e == 0xF) break;
    long long val = 0;
#pragma unroll
    for (int q = 0; q < 12; ++q)
      if (q == e) val = idx[q];       // keeps idx[] in registers
    o[s] = val;
  }
}

// ------------------------------------------------------------------------------------------------
// area-weighted surface sampling (trimesh 3.7.7 sample.sample_surface; the uniform numbers are the caller's)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tri_corner(const double* __restrict__ verts, long long v, double p[3]) {
  p[0] = verts[v * 3 + 0];
  p[1] = verts[v * 3 + 1];
  p[2] = verts[v * 3 + 2];
}

// area = |cross(b - a, c - a)| / 2 per face
constexpr int kAreaBlock = 256;
__global__ void __launch_bounds__(kAreaBlock) face_area_kernel(const double* __restrict__ verts, const long long* __restrict__ faces,
                                                               long long nf, double* __restrict__ area) {
  const long long f = (long long)blockIdx.x * kAreaBlock + threadIdx.x;
  if (f >= nf) return;
  double p0[3], p1[3], p2[3];
  tri_corner(verts, faces[f * 3 + 0], p0);
  tri_corner(verts, faces[f * 3 + 1], p1);
  tri_corner(verts, faces[f * 3 + 2], p2);
  const double ux = p1[0] - p0[0], uy = p1[1] - p0[1], uz = p1[2] - p0[2];
  const double vx = p2[0] - p0[0], vy = p2[1] - p0[1], vz = p2[2] - p0[2];
  const double cx = __dsub_rn(__dmul_rn(uy, vz), __dmul_rn(uz, vy));
  const double cy = __dsub_rn(__dmul_rn(uz, vx), __dmul_rn(ux, vz));
  const double cz = __dsub_rn(__dmul_rn(ux, vy), __dmul_rn(uy, vx));
  area[f] = __dmul_rn(sqrt(__dadd_rn(__dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy)), __dmul_rn(cz, cz))), 0.5);
}

// np.cumsum(a) in place, bit for bit, for a >= 0.  numpy's add.accumulate is ONE left-to-right chain of rounded fp64
// additions, s_i = fl(s_{i-1} + a_i), and a prefix rounded where numpy rounds it cannot come out of a tree of fp64 adds.  But
// while the running sum stays in one binade [2^e, 2^(e+1)) its ulp u is fixed, s = M u with an integer M, and
//     fl(s + a) = (M + k + c) u,   k = floor(a / u),   c = [a mod u > u / 2]   (a tie a mod u == u / 2 depends on M's parity),
// i.e. the chain IS an integer prefix sum of increments that depend on the binade only -- a parallel scan.  One CTA walks the
// array in pieces: every thread forms the increment of its element against the binade of the current sum, a block scan gives
// M + prefix, and the piece is accepted up to the first element where the sum would leave the binade (M + prefix >= 2^53), a
// tie, or an addend that is not smaller than the binade (or negative / non-finite); that one element is added by a real
// __dadd_rn and the walk resumes behind it.  116 k face areas: ~150 block steps instead of 116 k dependent DADDs (2.2 ms as
// a one-warp serial chain, ncu).
constexpr int kCumThreads = 1024;
__global__ void __launch_bounds__(kCumThreads) cumsum_exact_kernel(double* __restrict__ a, long long n, double* __restrict__ total) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long s_sum;          // bits of the running sum
  __shared__ int s_stop;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr unsigned long long kMant = (1ull << 52) - 1, kOne = 1ull << 52;
  if (tid == 0) s_sum = 0ull;                   // +0.0
  long long base = 0;
  while (base < n) {
    if (tid == 0) s_stop = kCumThreads;
    __syncthreads();                            // s_sum, s_stop visible
    const long long i = base + tid;
    const unsigned long long ab = i < n ? (unsigned long long)__double_as_longlong(a[i]) : 0ull;    // beyond n: + 0.0
    const unsigned long long sb = s_sum;
    const int Es = (int)((sb >> 52) & 0x7ff);
    const unsigned long long M = (sb & kMant) | kOne;
    bool bad = (Es == 0) || (Es == 0x7ff) || (sb >> 63);       // sum zero / subnormal / non-finite / negative: real additions
    unsigned long long inc = 0;
    if (!bad && ab != 0ull) {
      const int Ea = (int)((ab >> 52) & 0x7ff);
      if ((ab >> 63) || Ea == 0x7ff) {
        bad = true;
      } else {
        const unsigned long long A = Ea ? ((ab & kMant) | kOne) : (ab & kMant);
        const int d = Es - (Ea ? Ea : 1);       // a = A 2^(ea - 1075), s = M 2^(Es - 1075)
        if (d < 0) {
          bad = true;
        } else if (d == 0) {
          inc = A;
        } else if (d < 64) {
          const unsigned long long f = A & ((1ull << d) - 1), half = 1ull << (d - 1);
          inc = A >> d;
          if (f > half) ++inc;
          else if (f == half) bad = true;       // tie: the rounding depends on the parity of the sum
        }                                       // d >= 64: a < u / 2, contributes nothing
      }
    }
    // inclusive block scan of the increments
    unsigned long long pre = inc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    if (lane == 31) warp_tot[warp] = pre;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = warp_tot[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;                  // exclusive
    }
    __syncthreads();
    pre += warp_tot[warp];
    // the scan saturates nowhere: 1024 increments below 2^53 each stay below 2^63
    const unsigned long long Mi = M + pre;
    const bool stop_here = bad || Mi >= (1ull << 53);
    const unsigned int ball = __ballot_sync(0xffffffffu, stop_here);
    if (lane == 0 && ball) atomicMin(&s_stop, warp * 32 + (__ffs(ball) - 1));
    __syncthreads();
    const int stop = s_stop;
    const unsigned long long out = ((unsigned long long)Es << 52) | (Mi & kMant);
    if (tid < stop && i < n) a[i] = __longlong_as_double((long long)out);
    __syncthreads();                            // everybody has read s_sum / s_stop
    if (stop == 0) {                            // the first element of the piece needs a real addition
      if (tid == 0) {
        const double r = __dadd_rn(__longlong_as_double((long long)sb), a[base]);
        a[base] = r;
        s_sum = (unsigned long long)__double_as_longlong(r);
      }
      base += 1;
    } else {
      if (tid == stop - 1) s_sum = out;
      base += stop;
    }
  }
  __syncthreads();
  if (tid == 0 && total) *total = __longlong_as_double((long long)s_sum);
}

// one sample per thread: face = searchsorted(cum, u0 * total) (left), point = origin + l0 * e0 + l1 * e1 with the pair
// (l0, l1) folded back into the triangle when l0 + l1 > 1
__global__ void sample_surface_kernel(const double* __restrict__ verts, const long long* __restrict__ faces, long long nf,
                                      const double* __restrict__ cum, const double* __restrict__ total,
                                      const double* __restrict__ u, int count, double* __restrict__ out,
                                      long long* __restrict__ face_out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= count) return;
  const double pick = __dmul_rn(u[(size_t)s * 3 + 0], *total);
  long long lo = 0, hi = nf;               // first f with cum[f] >= pick
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (cum[mid] < pick) lo = mid + 1;
    else hi = mid;
  }
  const long long f = lo < nf ? lo : nf - 1;
  double p0[3], p1[3], p2[3];
  tri_corner(verts, faces[f * 3 + 0], p0);
  tri_corner(verts, faces[f * 3 + 1], p1);
  tri_corner(verts, faces[f * 3 + 2], p2);
  double l0 = u[(size_t)s * 3 + 1], l1 = u[(size_t)s * 3 + 2];
  if (__dadd_rn(l0, l1) > 1.0) {
    l0 = __dsub_rn(l0, 1.0);
    l1 = __dsub_rn(l1, 1.0);
  }
  l0 = fabs(l0);
  l1 = fabs(l1);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double e0 = __dsub_rn(p1[d], p0[d]), e1 = __dsub_rn(p2[d], p0[d]);
    out[(size_t)s * 3 + d] = __dadd_rn(__dadd_rn(__dmul_rn(e0, l0), __dmul_rn(e1, l1)), p0[d]);
  }
  if (face_out) face_out[s] = f;
}

static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

struct McWorkspace {
  uint8_t *cases, *counts;
  uint32_t *vbase, *ibase;
  unsigned long long *block_sums, *totals;
  int nblocks;
  size_t bytes;
};

static McWorkspace mc_carve(void* base, long long ncells) {
  McWorkspace w;
  w.nblocks = (int)((ncells + kMcBlock - 1) / kMcBlock);
  size_t off = 0;
  auto take = [&](size_t b) {
    char* p = base ? static_cast<char*>(base) + off : nullptr;
    off += align256(b);
    return p;
  };
  w.cases = reinterpret_cast<uint8_t*>(take((size_t)ncells));
  w.counts = reinterpret_cast<uint8_t*>(take((size_t)ncells));
  w.vbase = reinterpret_cast<uint32_t*>(take((size_t)ncells * 4));
  w.ibase = reinterpret_cast<uint32_t*>(take((size_t)ncells * 4));
  w.block_sums = reinterpret_cast<unsigned long long*>(take((size_t)w.nblocks * 8));
  w.totals = reinterpret_cast<unsigned long long*>(take(8));
  w.bytes = off;
  return w;
}

static int mc_grid(McGrid* g, const void* vol, int dtype, int nx, int ny, int nz, int pad, double pad_value, double iso) {
  IFD_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1 && (pad == 0 || pad == 1) && (dtype == 0 || dtype == 1), "marching cubes: bad arguments");
  g->vol = vol;
  g->is_f64 = dtype;
  g->nx = nx;
  g->ny = ny;
  g->nz = nz;
  g->pad = pad;
  g->pad_value = pad_value;
  g->iso = iso;
  g->px = nx + 2 * pad;
  g->py = ny + 2 * pad;
  g->pz = nz + 2 * pad;
  g->cx = g->px - 1;
  g->cy = g->py - 1;
  g->cz = g->pz - 1;
  g->ncells = (long long)g->cx * g->cy * g->cz;
  // vertex / index prefixes are 32-bit: at most 12 vertices and 15 indices per cell
  IFD_REQUIRE(g->ncells <= (1ll << 28), "marching cubes: lattice too large (more than 2^28 cells)");
  return IFD_OK;
}

}  // namespace ifd

using namespace ifd;

extern "C" size_t ifd_mc_workspace_bytes(int nx, int ny, int nz, int pad) {
  McGrid g;
  if (mc_grid(&g, nullptr, 0, nx, ny, nz, pad, 0.0, 0.0) != IFD_OK) return 0;
  return mc_carve(nullptr, g.ncells > 0 ? g.ncells : 1).bytes;
}

extern "C" int ifd_mc_count(const void* volume, int dtype, int nx, int ny, int nz, int pad, double pad_value, double isovalue,
                            void* workspace, size_t workspace_bytes, long long* n_verts, long long* n_faces, ifd_stream_t stream) {
  IFD_REQUIRE(volume && workspace && n_verts && n_faces, "ifd_mc_count: null argument");
  McGrid g;
  int rc = mc_grid(&g, volume, dtype, nx, ny, nz, pad, pad_value, isovalue);
  if (rc) return rc;
  *n_verts = *n_faces = 0;
  if (g.ncells <= 0) return IFD_OK;
  McWorkspace w = mc_carve(workspace, g.ncells);
  IFD_REQUIRE(workspace_bytes >= w.bytes, "ifd_mc_count: workspace too small (ifd_mc_workspace_bytes)");
  cudaStream_t s = as_stream(stream);
  mc_classify_kernel<<<w.nblocks, kMcBlock, 0, s>>>(g, w.cases, w.counts, w.block_sums);
  IFD_LAUNCH_CHECK("mc_classify_kernel");
  mc_scan_blocks_kernel<<<1, 1024, 0, s>>>(w.block_sums, w.nblocks, w.totals);
  IFD_LAUNCH_CHECK("mc_scan_blocks_kernel");
  mc_bases_kernel<<<w.nblocks, kMcBlock, 0, s>>>(g.ncells, w.counts, w.block_sums, w.vbase, w.ibase);
  IFD_LAUNCH_CHECK("mc_bases_kernel");
  unsigned long long tot = 0;
  IFD_CUDA_TRY(cudaMemcpyAsync(&tot, w.totals, 8, cudaMemcpyDeviceToHost, s));
  IFD_CUDA_TRY(cudaStreamSynchronize(s));
  *n_verts = (long long)(tot & 0xffffffffull);
  *n_faces = (long long)(tot >> 32) / 3;
  return IFD_OK;
}

extern "C" int ifd_mc_emit(const void* volume, int dtype, int nx, int ny, int nz, int pad, double pad_value, double isovalue,
                           int to_box, double box_size, const void* workspace, size_t workspace_bytes, double* verts_out,
                           long long* faces_out, ifd_stream_t stream) {
  IFD_REQUIRE(volume && workspace && verts_out && faces_out, "ifd_mc_emit: null argument");
  McGrid g;
  int rc = mc_grid(&g, volume, dtype, nx, ny, nz, pad, pad_value, isovalue);
  if (rc) return rc;
  if (g.ncells <= 0) return IFD_OK;
  McWorkspace w = mc_carve(const_cast<void*>(workspace), g.ncells);
  IFD_REQUIRE(workspace_bytes >= w.bytes, "ifd_mc_emit: workspace too small (ifd_mc_workspace_bytes)");
  McXform xf;
  xf.on = to_box ? 1 : 0;
  // generation.py:181: divide by (n - 1) of the UNPADDED lattice
  xf.dx = (double)(g.px - 2 - 1);
  xf.dy = (double)(g.py - 2 - 1);
  xf.dz = (double)(g.pz - 2 - 1);
  xf.box = box_size;
  IFD_REQUIRE(!to_box || (xf.dx > 0 && xf.dy > 0 && xf.dz > 0), "ifd_mc_emit: to_box needs an unpadded lattice of >= 2 points per axis");
  const int blocks = (int)((g.ncells + 255) / 256);
  mc_emit_kernel<<<blocks, 256, 0, as_stream(stream)>>>(g, xf, w.cases, w.vbase, w.ibase, verts_out, faces_out);
  IFD_LAUNCH_CHECK("mc_emit_kernel");
  return IFD_OK;
}

// np.cumsum over a device array of non-negative float64 values, in place and bit for bit (the prefix ifd_sample_surface
// searches); total_out (device, optional) receives the last element.
extern "C" int ifd_cumsum_f64(double* data, long long n, double* total_out, ifd_stream_t stream) {
  IFD_REQUIRE(data && n > 0, "ifd_cumsum_f64: bad arguments");
  cumsum_exact_kernel<<<1, kCumThreads, 0, as_stream(stream)>>>(data, n, total_out);
  IFD_LAUNCH_CHECK("cumsum_exact_kernel");
  return IFD_OK;
}

extern "C" size_t ifd_sample_surface_workspace_bytes(long long n_faces) {
  const long long nb = (n_faces + kAreaBlock - 1) / kAreaBlock;
  return align256((size_t)(n_faces > 0 ? n_faces : 1) * 8) + align256((size_t)(nb > 0 ? nb : 1) * 8) + 256;
}

extern "C" int ifd_sample_surface(const double* verts, long long n_verts, const long long* faces, long long n_faces,
                                  const double* uniforms, int count, double* xyz_out, long long* face_out, void* workspace,
                                  size_t workspace_bytes, ifd_stream_t stream) {
  IFD_REQUIRE(verts && faces && uniforms && xyz_out && workspace && count > 0 && n_verts > 0, "ifd_sample_surface: bad arguments");
  // the reference's sample_surface raises IndexError on a mesh without faces (remesh_defense.py:160-171 catches it)
  if (n_faces <= 0) return fail(IFD_ERR_INVALID, "ifd_sample_surface: the mesh has no faces");
  IFD_REQUIRE(workspace_bytes >= ifd_sample_surface_workspace_bytes(n_faces), "ifd_sample_surface: workspace too small");
  const int nb = (int)((n_faces + kAreaBlock - 1) / kAreaBlock);
  char* base = static_cast<char*>(workspace);
  double* cum = reinterpret_cast<double*>(base);
  double* bsum = reinterpret_cast<double*>(base + align256((size_t)n_faces * 8));
  double* total = reinterpret_cast<double*>(base + align256((size_t)n_faces * 8) + align256((size_t)nb * 8));
  cudaStream_t s = as_stream(stream);
  face_area_kernel<<<nb, kAreaBlock, 0, s>>>(verts, faces, n_faces, cum);
  IFD_LAUNCH_CHECK("face_area_kernel");
  cumsum_exact_kernel<<<1, kCumThreads, 0, s>>>(cum, n_faces, total);
  IFD_LAUNCH_CHECK("cumsum_exact_kernel");
  sample_surface_kernel<<<(count + 127) / 128, 128, 0, s>>>(verts, faces, n_faces, cum, total, uniforms, count, xyz_out, face_out);
  IFD_LAUNCH_CHECK("sample_surface_kernel");
  return IFD_OK;
}

// ------------------------------------------------------------------------------------------------
// MISE -- multiresolution iso-surface extraction (ONet/im2mesh/utils/libmise/mise.pyx:33-369): the octree controller
// that decides which lattice points generate_from_latent evaluates (generation.py:113-130).
//
// The reference keeps vectors of voxels and grid points and a hash map; every decision it takes is a set decision, so
// the state here is dense and the passes are data parallel:
//   exists[n^3], known[n^3]   lattice points that were created / whose value was set        (n = (resolution_0 << depth) + 1)
//   sub[L][(r0 << L)^3]       voxel of level L was subdivided; a voxel exists when L == 0 or its parent was subdivided
// A leaf voxel below the last level is subdivided when the known points on its closure (corners, and points that finer
// neighbours put on its faces and edges -- subdivide_voxels visits the 8 unit cells around every known point,
// mise.pyx:195-219) hold both a value >= threshold and a value <= threshold; subdividing creates the 27 points of the
// half-size lattice (:247-282).  Levels are processed finest first so that children created in this round are not
// examined in it (the reference fixes the flags of the existing voxels before it subdivides, :221-237).
// ------------------------------------------------------------------------------------------------
namespace ifd {

struct MiseState {
  int r0, depth, res, n;
  long long npts;
  double* val;
  uint8_t *exists, *known;
  uint8_t* sub[8];
  unsigned long long* counter;   // [0] query cursor, [1] error flag
  size_t bytes;
};

static int mise_carve(MiseState* m, void* base, int r0, int depth) {
  IFD_REQUIRE(r0 >= 1 && depth >= 0 && depth <= 7 && ((long long)r0 << depth) <= 1024, "MISE: resolution_0 << depth must be <= 1024, depth <= 7");
  m->r0 = r0;
  m->depth = depth;
  m->res = r0 << depth;
  m->n = m->res + 1;
  m->npts = (long long)m->n * m->n * m->n;
  size_t off = 0;
  auto take = [&](size_t b) {
    char* p = base ? static_cast<char*>(base) + off : nullptr;
    off += align256(b);
    return p;
  };
  m->val = reinterpret_cast<double*>(take((size_t)m->npts * 8));
  m->exists = reinterpret_cast<uint8_t*>(take((size_t)m->npts));
  m->known = reinterpret_cast<uint8_t*>(take((size_t)m->npts));
  for (int L = 0; L < 8; ++L) m->sub[L] = nullptr;
  for (int L = 0; L < depth; ++L) {
    const size_t r = (size_t)r0 << L;
    m->sub[L] = reinterpret_cast<uint8_t*>(take(r * r * r));
  }
  m->counter = reinterpret_cast<unsigned long long*>(take(16));
  m->bytes = off;
  return IFD_OK;
}

__global__ void mise_init_kernel(MiseState m) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= m.npts) return;
  const int z = (int)(p % m.n), y = (int)((p / m.n) % m.n), x = (int)(p / ((long long)m.n * m.n));
  const int s = 1 << m.depth;
  m.exists[p] = (x % s == 0 && y % s == 0 && z % s == 0) ? 1 : 0;
  m.known[p] = 0;
  m.val[p] = __longlong_as_double(0x7ff8000000000000ll);       // np.nan (to_dense, mise.pyx:130)
}

// unknown points that exist: count only (out == nullptr) or append (x, y, z)
__global__ void mise_query_kernel(MiseState m, long long* __restrict__ out, long long cap) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool want = p < m.npts && m.exists[p] && !m.known[p];
  const unsigned ballot = __ballot_sync(0xffffffffu, want);
  if (ballot == 0) return;
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(m.counter, (unsigned long long)__popc(ballot));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (want && out) {
    const long long at = (long long)base + __popc(ballot & ((1u << lane) - 1));
    if (at < cap) {
      out[at * 3 + 0] = p / ((long long)m.n * m.n);
      out[at * 3 + 1] = (p / m.n) % m.n;
      out[at * 3 + 2] = p % m.n;
    }
  }
}

__global__ void mise_update_kernel(MiseState m, const long long* __restrict__ pts, const double* __restrict__ values, long long count) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count) return;
  const long long x = pts[e * 3 + 0], y = pts[e * 3 + 1], z = pts[e * 3 + 2];
  if (x < 0 || y < 0 || z < 0 || x >= m.n || y >= m.n || z >= m.n || !m.exists[(x * m.n + y) * m.n + z]) {
    m.counter[1] = 1;                                           // ValueError('Point not in grid!'), mise.pyx:101-102
    return;
  }
  const long long p = (x * m.n + y) * m.n + z;
  m.val[p] = values[e];
  m.known[p] = 1;
}

// one thread per voxel of level L
__global__ void mise_subdivide_kernel(MiseState m, int L, double thr) {
  const int r = m.r0 << L;
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= (long long)r * r * r) return;
  const int k = (int)(v % r), j = (int)((v / r) % r), i = (int)(v / ((long long)r * r));
  if (L > 0) {
    const int rp = r >> 1;
    if (!m.sub[L - 1][((long long)(i >> 1) * rp + (j >> 1)) * rp + (k >> 1)]) return;      // the voxel does not exist
  }
  if (m.sub[L][v]) return;                                                               // not a leaf
  const int s = 1 << (m.depth - L);                                                      // voxel size on the fine lattice
  const int x0 = i * s, y0 = j * s, z0 = k * s;
  bool pos = false, neg = false;
  for (int a = 0; a <= s; ++a)
    for (int b = 0; b <= s; ++b)
      for (int c = 0; c <= s; ++c) {
        const long long p = ((long long)(x0 + a) * m.n + (y0 + b)) * m.n + (z0 + c);
        if (!m.known[p]) continue;
        const double f = m.val[p];
        pos = pos || f >= thr;
        neg = neg || f <= thr;
      }
  if (!(pos && neg)) return;
  m.sub[L][v] = 1;
  const int h = s >> 1;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      for (int c = 0; c < 3; ++c) m.exists[((long long)(x0 + a * h) * m.n + (y0 + b * h)) * m.n + (z0 + c * h)] = 1;
}

// to_dense (mise.pyx:128-163): NaN where nothing was evaluated, then completed along x, then y, then z
__global__ void mise_fill_kernel(const MiseState m, double* __restrict__ out, int axis) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m.n * m.n) return;
  const int a = t / m.n, b = t % m.n;
  const long long n = m.n;
  long long start, stride;
  if (axis == 0) { start = a * n + b; stride = n * n; }              // (j, k) lines along i
  else if (axis == 1) { start = a * n * n + b; stride = n; }         // (i, k) lines along j
  else { start = (a * n + b) * n; stride = 1; }                      // (i, j) lines along k
  double prev = out[start];
  for (int q = 1; q < m.n; ++q) {
    const long long p = start + q * stride;
    const double f = out[p];
    if (f != f) out[p] = prev;
    else prev = f;
  }
}

}  // namespace ifd

extern "C" size_t ifd_mise_workspace_bytes(int resolution0, int depth) {
  MiseState m;
  if (mise_carve(&m, nullptr, resolution0, depth) != IFD_OK) return 0;
  return m.bytes;
}

#define IFD_MISE_STATE(m)                                                                           \
  MiseState m;                                                                                      \
  {                                                                                                 \
    IFD_REQUIRE(workspace, "MISE: null workspace");                                                 \
    int rc_ = mise_carve(&m, workspace, resolution0, depth);                                        \
    if (rc_) return rc_;                                                                            \
    IFD_REQUIRE(workspace_bytes >= m.bytes, "MISE: workspace too small (ifd_mise_workspace_bytes)"); \
  }

extern "C" int ifd_mise_init(int resolution0, int depth, void* workspace, size_t workspace_bytes, ifd_stream_t stream) {
  IFD_MISE_STATE(m);
  cudaStream_t s = as_stream(stream);
  for (int L = 0; L < depth; ++L) {
    const size_t r = (size_t)resolution0 << L;
    IFD_CUDA_TRY(cudaMemsetAsync(m.sub[L], 0, r * r * r, s));
  }
  IFD_CUDA_TRY(cudaMemsetAsync(m.counter, 0, 16, s));
  mise_init_kernel<<<(unsigned)((m.npts + 255) / 256), 256, 0, s>>>(m);
  IFD_LAUNCH_CHECK("mise_init_kernel");
  return IFD_OK;
}

extern "C" int ifd_mise_query(int resolution0, int depth, void* workspace, size_t workspace_bytes, long long* points_out,
                              long long capacity, long long* count_host, ifd_stream_t stream) {
  IFD_MISE_STATE(m);
  IFD_REQUIRE(count_host, "ifd_mise_query: null count");
  cudaStream_t s = as_stream(stream);
  IFD_CUDA_TRY(cudaMemsetAsync(m.counter, 0, 8, s));
  mise_query_kernel<<<(unsigned)((m.npts + 255) / 256), 256, 0, s>>>(m, points_out, points_out ? capacity : 0);
  IFD_LAUNCH_CHECK("mise_query_kernel");
  unsigned long long cnt = 0;
  IFD_CUDA_TRY(cudaMemcpyAsync(&cnt, m.counter, 8, cudaMemcpyDeviceToHost, s));
  IFD_CUDA_TRY(cudaStreamSynchronize(s));
  *count_host = (long long)cnt;
  IFD_REQUIRE(!points_out || (long long)cnt <= capacity, "ifd_mise_query: capacity smaller than the number of unknown points");
  return IFD_OK;
}

extern "C" int ifd_mise_update(int resolution0, int depth, double threshold, void* workspace, size_t workspace_bytes,
                               const long long* points, const double* values, long long count, ifd_stream_t stream) {
  IFD_MISE_STATE(m);
  IFD_REQUIRE(count == 0 || (points && values), "ifd_mise_update: null argument");
  cudaStream_t s = as_stream(stream);
  if (count > 0) {
    mise_update_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(m, points, values, count);
    IFD_LAUNCH_CHECK("mise_update_kernel");
    unsigned long long bad = 0;
    IFD_CUDA_TRY(cudaMemcpyAsync(&bad, m.counter + 1, 8, cudaMemcpyDeviceToHost, s));
    IFD_CUDA_TRY(cudaStreamSynchronize(s));
    if (bad) {
      IFD_CUDA_TRY(cudaMemsetAsync(m.counter + 1, 0, 8, s));
      return fail(IFD_ERR_INVALID, "Point not in grid!");
    }
  }
  for (int L = depth - 1; L >= 0; --L) {
    const long long r = (long long)resolution0 << L;
    mise_subdivide_kernel<<<(unsigned)((r * r * r + 127) / 128), 128, 0, s>>>(m, L, threshold);
    IFD_LAUNCH_CHECK("mise_subdivide_kernel");
  }
  return IFD_OK;
}

extern "C" int ifd_mise_to_dense(int resolution0, int depth, void* workspace, size_t workspace_bytes, double* dense_out,
                                 ifd_stream_t stream) {
  IFD_MISE_STATE(m);
  IFD_REQUIRE(dense_out, "ifd_mise_to_dense: null output");
  cudaStream_t s = as_stream(stream);
  IFD_CUDA_TRY(cudaMemcpyAsync(dense_out, m.val, (size_t)m.npts * 8, cudaMemcpyDeviceToDevice, s));
  for (int axis = 0; axis < 3; ++axis) {
    mise_fill_kernel<<<(m.n * m.n + 127) / 128, 128, 0, s>>>(m, dense_out, axis);
    IFD_LAUNCH_CHECK("mise_fill_kernel");
  }
  return IFD_OK;
}
