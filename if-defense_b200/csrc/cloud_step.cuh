// cloud_step.cuh -- everything of one Adam step that is NOT the decoder, for one cloud, in one CTA:
//   knn_point (defense/pn_utils.py:64-83) -> repulsion pairs + their scatter-add backward
//   (defense/repulsion_loss.py:43-53, index_points backward) -> g = g_occ + coef * g_rep -> Adam
//   (opt_defense.py:226-228, torch/optim/adam.py single-tensor path).
//
// Why a redesign: ncu on the previous generation (profiles/r01_v3_*) shows the brute-force K^2 scan at 50 us and
// the 6-limb global accumulator traffic of the separate Adam kernel at 15 us per step -- 45 % of a step.  Here
//   * the warm-start bound tau_i (largest current key among last step's k+1 neighbours, an upper bound on the
//     (k+1)-th smallest key) turns the scan into a range query on a uniform 16^3 grid rebuilt in shared memory
//     every step (counting sort of the K indices): ~20 candidates per query instead of K.  The candidate set is a
//     superset of {j : key(i,j) <= tau_i} (radius inflated by the rounding error bound of the expanded-form key),
//     keys are evaluated in the reference's association and ranked by (key, index): bit-identical indices to the
//     brute-force scan, ties included;
//   * the scatter-add becomes a gather.  A mutual edge pair (i->j, j->i) contributes to i the bit-identical
//     value twice (d^2 is symmetric and x_j - x_i = -(x_i - x_j) exactly in floating point), so only NON-mutual
//     edges are sent to the target's inbox (shared-memory slots, one integer atomic per such edge); a hub whose
//     inbox overflows recovers its in-edges by an ordered scan of the neighbour lists;
//   * every point sums its terms in one canonical order (mutual terms by rank, own sum, in-edges by ascending
//     source) in fp64, so the result does not depend on atomic arrival order: bitwise reproducible;
//   * Adam runs in the same thread; xyz, m, v make one round trip to HBM/L2 per step.
//
// Two forms of the same body (template on the cluster size, same bits):
//   * cloud_step_kernel: two CTAs (a thread-block cluster) per cloud.  Both build the whole grid, the cloud is split at the
//     cell boundary nearest the median (a deterministic function of the cell histogram), each CTA ranks / sums / updates the
//     points on its side (the idle half of its threads walks every other row of a query's box), neighbour lists and
//     non-mutual in-edges cross through distributed shared memory.  Shortest launch (42 us at B = 64): a loop running alone.
//   * cloud_step_solo_kernel: one CTA per cloud, one thread per query, no cluster traffic.  A longer launch (53 us) on HALF
//     the SMs -- and a tail CTA takes a whole SM's registers, so nothing shares an SM with it: when the loops of several
//     batches run side by side this form leaves the other loops' decode launches twice the room (restore.cu picks it then).
// The range walk reads candidates from a copy of the positions in CELL order (one load per candidate, no index hop), takes
// two candidates per step, and fetches the next row's bounds while the current row is evaluated.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "ifd_math.cuh"
#include "topk.cuh"

namespace ifd {

constexpr int kCsThreads = 1024;
constexpr int kCsMaxK = 1024;               // one thread per point
constexpr int kCsG = 16;                    // grid cells per axis
constexpr int kCsCells = kCsG * kCsG * kCsG;
constexpr int kCsInbox = 16;                // non-mutual in-edges kept per target before the ordered fallback
constexpr int kCsKK = 8;                    // neighbour-list width (k + 1 <= 8)
constexpr int kCsMaxHub = 128;              // hubs served by the warp-cooperative pass (beyond: serial fallback)

struct CloudStepArgs {
  float* xyz;            // [B][K][3] in/out
  float* m;              // [B][K][3] Adam exp_avg in/out
  float* v;              // [B][K][3] Adam exp_avg_sq in/out
  const float* g_occ;    // [B][K][3] d occ_loss / d xyz of this step (decode kernel)
  int32_t* nbr;          // [B][K][8] neighbour lists: in = previous step's (warm), out = this step's
  float* loss_part;      // [B] sum of the pair losses of the cloud (optional)
  float* rep_grad_out;   // [B][K][3] sum of the pair-loss gradients per point, before rep_coef (optional, diagnostics)
  int K, k, warm;
  int inbox_cap;         // <= kCsInbox (tests lower it to force the hub fallback)
  float radius, h, eps, rep_coef;
  float omb1, b2, omb2, adam_eps;
  AdamStepConst sc;
  int bar_mode;          // cluster barrier before the first remote store -- 0: none, 1: arrive.release / wait.acquire,
                         // 2: arrive.relaxed / wait (execution barrier only; the default)
  const LoopJob* job;    // non-null: xyz / m / v / g_occ / nbr come from this record (graph replay)
  int zero_mv;           // first step of a fresh run: Adam state is zero, m and v are not read
};

struct CloudStepSmem {
  float4 pos[kCsMaxK];                       // x y z |x|^2
  uint32_t cell[kCsCells + 4];               // counting sort: cell[c] = first slot of cell c, cell[c + 1] = end
  uint16_t nbr[kCsMaxK][kCsKK];              // last step's lists (warm start), filled locally from global memory
  uint16_t nbrn[kCsMaxK][kCsKK];             // this step's lists (column 0 = the dropped "self" column): own rows written here
                                             // and into the peer CTA's copy -- a separate array, so the peer's warm-start
                                             // initialisation of `nbr` can never overwrite a list that arrived early
  uint16_t inbox[kCsMaxK][kCsInbox];         // sources of non-mutual in-edges
  uint32_t inbox_cnt[kCsMaxK];
  uint16_t sorted[kCsMaxK];                  // point indices in cell order
  float4 spos[kCsMaxK];                      // pos in cell order: the range walk reads candidate t with ONE load (no index hop)
  float gmv[3][3 * kCsMaxK];                 // g_occ, m, v of the cloud, flat [K][3] (coalesced in, coalesced out)
  uint16_t cid[kCsMaxK];                     // cell of every point (ownership: cell < split -> CTA 0)
  uint16_t hcnt[kCsMaxK];                    // candidates found by the helper thread of each query slot
  double hubsum[kCsMaxHub][3];               // in-edge sums of hubs (points whose inbox overflowed)
  uint16_t hub[kCsMaxHub];
  float red[8][32];
  uint32_t scan[32];
  float bbox[8];
  uint32_t nhub;
  int split_cell, split_rank;                // first cell / first sorted rank owned by CTA 1
  float peer_loss;
};

__device__ __forceinline__ int cs_cell(float v, float lo, float inv_h) {
  float f = floorf((v - lo) * inv_h);                     // monotone in v: range queries stay supersets
  f = fminf(fmaxf(f, 0.0f), (float)(kCsG - 1));           // (NaN -> 0)
  return (int)f;
}

// Linear index of cell (cx, cy, cz), x fastest: for fixed (cy, cz) an x-range of cells is one contiguous range of
// the sorted point array.
__device__ __forceinline__ int cs_index(int cx, int cy, int cz) { return (cz * kCsG + cy) * kCsG + cx; }

// gradient on the TARGET t of the pair loss of edge (source s -> target t):  gcoef * (x_t - x_s), and the loss
__device__ __forceinline__ void cs_pair(const float4& src, const float4& tgt, float radius, float h, float eps, float& gx,
                                        float& gy, float& gz, float& loss) {
  const float dx = sub_rn(tgt.x, src.x), dy = sub_rn(tgt.y, src.y), dz = sub_rn(tgt.z, src.z);
  const RepPair p = repulsion_pair(dx, dy, dz, radius, h, eps);
  gx = p.gcoef * dx;
  gy = p.gcoef * dy;
  gz = p.gcoef * dz;
  loss = p.loss;
}

__device__ __forceinline__ bool cs_row_has(const uint16_t* row, int k, int who) {
  const uint4 r = *reinterpret_cast<const uint4*>(row);   // 8 x uint16
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  bool hit = false;
#pragma unroll
  for (int s = 1; s < kCsKK; ++s) {
    const uint32_t e = (s & 1) ? (w[s >> 1] >> 16) : (w[s >> 1] & 0xffffu);
    hit |= (s <= k) && (e == (uint32_t)who);
  }
  return hit;
}

// CL = 2: the cluster form described above.  CL = 1: one CTA owns the whole cloud (no helpers at K = 1024, no DSMEM, the
// cluster barriers become __syncthreads): a longer walk per thread, but half the SMs per launch -- what matters when the
// loops of two batches share the GPU (ifd_test_hook(5, 1)).
template <int CL>
__device__ __forceinline__ void cloud_step_body(const CloudStepArgs& a) {
  extern __shared__ __align__(16) unsigned char cs_raw[];
  // graph replay: the buffers of THIS launch are in the job record.  (Plain locals on purpose: copying the argument struct
  // and overwriting its pointer members made nvcc 12.9 drop the assignment of `xyz` -- the kernel then kept using the
  // capture-time pointer, which only shows when a replay runs on other buffers than the capture did.)
  float* const g_xyz = a.job ? a.job->xyz : a.xyz;
  float* const g_m = a.job ? a.job->m : a.m;
  float* const g_v = a.job ? a.job->v : a.v;
  const float* const g_gocc = a.job ? a.job->g_occ : a.g_occ;
  int32_t* const g_nbr = a.job ? a.job->nbr : a.nbr;
  CloudStepSmem& S = *reinterpret_cast<CloudStepSmem*>(cs_raw);
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
  const int half = CL == 2 ? (int)cluster.block_rank() : 0;    // which side of the split this CTA owns
  CloudStepSmem& P = CL == 2 ? *cluster.map_shared_rank(&S, half ^ 1) : S;   // the peer CTA's shared memory
  auto cluster_sync = [&]() {
    if (CL == 2) cluster.sync();
    else __syncthreads();
  };
  const int b = CL == 2 ? blockIdx.x >> 1 : blockIdx.x, i = threadIdx.x, K = a.K, k = a.k;
  const int lane = i & 31, warp = i >> 5;
  bool live = i < K;

  // ---- all global reads happen here, coalesced over the flat [K][3] arrays (both CTAs read the whole cloud)
  const size_t cloud3 = (size_t)b * K * 3;
  float* xs = reinterpret_cast<float*>(&S.inbox[0][0]);        // xyz staging (the inbox is not in use yet)
  // (every load is issued before the first shared-memory store: the kernel starts with a cold L1, so what this phase costs is
  // L2 round trips -- one, not one per array and iteration; ncu had 22 % of the kernel's warp time on this loop)
  {
    static_assert(3 * kCsMaxK == 3 * kCsThreads, "three elements per thread and array");
    float rx[3], rg[3], rm[3], rv[3];
    int4 p0 = make_int4(0, 0, 0, 0), p1 = p0;
    const bool have_mv = !a.zero_mv;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int e = i + u * kCsThreads;
      const bool in = e < 3 * K;
      rx[u] = in ? __ldcg(g_xyz + cloud3 + e) : 0.0f;
      rg[u] = in ? __ldcg(g_gocc + cloud3 + e) : 0.0f;
      rm[u] = (in && have_mv) ? __ldcg(g_m + cloud3 + e) : 0.0f;
      rv[u] = (in && have_mv) ? __ldcg(g_v + cloud3 + e) : 0.0f;
    }
    if (live && a.warm) {
      const int4* pv = reinterpret_cast<const int4*>(g_nbr + ((size_t)b * K + i) * kCsKK);
      p0 = __ldcg(pv);
      p1 = __ldcg(pv + 1);
    }
    for (int c = i; c < kCsCells + 4; c += kCsThreads) S.cell[c] = 0;      // (independent work under the load latency)
    S.inbox_cnt[i] = 0;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int e = i + u * kCsThreads;
      if (e < 3 * K) {
        xs[e] = rx[u];
        S.gmv[0][e] = rg[u];
        S.gmv[1][e] = rm[u];
        S.gmv[2][e] = rv[u];
      }
    }
    if (live && a.warm) {                                      // previous lists -> 16-bit rows
      const int prev[kCsKK] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      uint32_t w[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int e0 = (unsigned)prev[2 * s] < (unsigned)K ? prev[2 * s] : i;
        const int e1 = (unsigned)prev[2 * s + 1] < (unsigned)K ? prev[2 * s + 1] : i;
        w[s] = (uint32_t)e0 | ((uint32_t)e1 << 16);
      }
      *reinterpret_cast<uint4*>(&S.nbr[i][0]) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  if (i == 0) {
    S.nhub = 0;
    S.split_cell = kCsCells;
    S.split_rank = K;
    S.peer_loss = 0.0f;
  }
  __syncthreads();
  // Distributed shared memory may only be touched once the peer CTA is known to be running: arrive here, wait just before
  // the first remote store -- the grid build and the kNN walk hide the barrier latency.  No memory ordering is needed from
  // this pair (the rows written remotely, `nbrn`, are not initialised by their owner; the inbox counters are only pushed
  // to after the full cluster.sync below), so the arrive is relaxed.
  if (CL == 2) {
    if (a.bar_mode == 1) cluster.barrier_arrive();
    else if (a.bar_mode == 2) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  }
  float4 me0 = make_float4(0.f, 0.f, 0.f, 0.f);        // point i (load order); re-assigned in cell order below
  if (live) {
    const float x = xs[3 * i + 0], y = xs[3 * i + 1], z = xs[3 * i + 2];
    me0 = make_float4(x, y, z, sqnorm3(x, y, z));
    S.pos[i] = me0;
  }

  // ---- bounding box of the cloud
  {
    float lo[3] = {live ? me0.x : INFINITY, live ? me0.y : INFINITY, live ? me0.z : INFINITY};
    float hi[3] = {live ? me0.x : -INFINITY, live ? me0.y : -INFINITY, live ? me0.z : -INFINITY};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        lo[ax] = fminf(lo[ax], __shfl_xor_sync(0xffffffffu, lo[ax], o));
        hi[ax] = fmaxf(hi[ax], __shfl_xor_sync(0xffffffffu, hi[ax], o));
      }
    if (lane == 0) {
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        S.red[ax][warp] = lo[ax];
        S.red[3 + ax][warp] = hi[ax];
      }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        float l = S.red[ax][lane], h2 = S.red[3 + ax][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
          h2 = fmaxf(h2, __shfl_xor_sync(0xffffffffu, h2, o));
        }
        if (lane == 0) {
          S.bbox[ax] = l;
          S.bbox[3 + ax] = h2;
        }
      }
    }
    __syncthreads();
  }
  const float lox = S.bbox[0], loy = S.bbox[1], loz = S.bbox[2];
  const float ihx = (float)kCsG / fmaxf(S.bbox[3] - lox, 1e-30f);
  const float ihy = (float)kCsG / fmaxf(S.bbox[4] - loy, 1e-30f);
  const float ihz = (float)kCsG / fmaxf(S.bbox[5] - loz, 1e-30f);
  // rounding-error bound of the expanded-form key against the true squared distance (~1.5e-6 * max|x|^2)
  float m2 = 0.0f;
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const float t = fmaxf(fabsf(S.bbox[ax]), fabsf(S.bbox[3 + ax]));
    m2 = fmaf(t, t, m2);
  }
  const float key_margin = 4e-6f * m2 + 1e-30f;

  // ---- counting sort of the point indices by cell
  int cid = 0;
  if (live) {
    cid = cs_index(cs_cell(me0.x, lox, ihx), cs_cell(me0.y, loy, ihy), cs_cell(me0.z, loz, ihz));
    S.cid[i] = (uint16_t)cid;
    atomicAdd(&S.cell[cid + 1], 1u);
  }
  __syncthreads();
  {
    static_assert(kCsCells == 4 * kCsThreads, "each thread scans 4 cells");
    uint32_t* mine = &S.cell[1 + 4 * i];
    const uint32_t c0 = mine[0], c1 = mine[1], c2 = mine[2], c3 = mine[3];
    const uint32_t tot = c0 + c1 + c2 + c3;
    uint32_t inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) S.scan[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const uint32_t w = S.scan[lane];
      uint32_t winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      S.scan[lane] = winc - w;
    }
    __syncthreads();
    const uint32_t ex = S.scan[warp] + inc - tot;
    mine[0] = ex;
    mine[1] = ex + c0;
    mine[2] = ex + c0 + c1;
    mine[3] = ex + c0 + c1 + c2;
    // split: the first cell whose exclusive start reaches K/2 (unique; none -> CTA 0 owns everything)
    const uint32_t e4[5] = {ex, ex + c0, ex + c0 + c1, ex + c0 + c1 + c2, ex + tot};
    const uint32_t halfK = (uint32_t)((K + 1) / 2);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (e4[u] < halfK && e4[u + 1] >= halfK && 4 * i + u + 1 < kCsCells) {
        S.split_cell = 4 * i + u + 1;
        S.split_rank = (int)e4[u + 1];
      }
  }
  __syncthreads();
  if (live) {
    const uint32_t slot = atomicAdd(&S.cell[cid + 1], 1u);
    S.sorted[slot] = (uint16_t)i;
    S.spos[slot] = me0;
  }
  __syncthreads();                      // now cell[c] .. cell[c + 1] delimit cell c (cell[0] == 0)

  // ---- from here on thread t owns the t-th point in CELL order: the lanes of a warp are spatial neighbours, so
  //      their range queries walk (nearly) the same rows and candidates -- coherent loops, broadcast loads
  const int rank0 = (CL == 1 || half == 0) ? 0 : S.split_rank, rank1 = (CL == 2 && half == 0) ? S.split_rank : K;
  const int n_own = rank1 - rank0;
  live = i < n_own;
  // Two threads per query wherever the CTA has threads to spare (the split is at the median, so about half of them
  // are): thread n_own + q walks the odd rows of query q's box, thread q the even ones -- the idle half of the CTA
  // halves the length of the divergent candidate walk.
  const int n_help = min(kCsThreads - n_own, n_own);           // queries [0, n_help) have a helper
  const bool helper = i >= n_own && i - n_own < n_help;
  const int qi = helper ? i - n_own : i;                       // query slot this thread works for
  const bool helped = qi < n_help;                             // this thread's query is shared by two threads
  const bool scan = live || helper;
  const int p = scan ? (int)S.sorted[rank0 + qi] : (int)S.sorted[0];
  const float4 me = S.pos[p];

  // ---- kNN
  TopK<kCsKK> top;
  top.init(INFINITY);
  {
    // the scan threshold and the scan itself use one expression: -2 * dot(a, b) == dot(-2 a, b) bit for bit (scaling
    // by a power of two commutes with rounding), folded into the query once; the ranking keys are knn_key proper
    const float mx = -2.0f * me.x, my = -2.0f * me.y, mz = -2.0f * me.z;
    float tau = scan ? INFINITY : -INFINITY;                   // idle lanes accept nothing
    if (scan && a.warm) {
      const uint4 pr = *reinterpret_cast<const uint4*>(&S.nbr[p][0]);
      const uint32_t pw[4] = {pr.x, pr.y, pr.z, pr.w};
      tau = -INFINITY;
#pragma unroll
      for (int s = 0; s < kCsKK; ++s)
        if (s <= k) {
          const int j = (int)((s & 1) ? (pw[s >> 1] >> 16) : (pw[s >> 1] & 0xffffu));
          const float4 c = S.pos[j];
          tau = fmaxf(tau, add_rn(add_rn(c.w, dot3_chain(mx, my, mz, c.x, c.y, c.z)), me.w));
        }
      if (!(tau < INFINITY)) tau = INFINITY;                   // NaN -> unbounded
    }
    // Cell box of this lane's query sphere.  Each lane walks its own rows and candidates with a FLAT iterator -- one
    // action per warp iteration, either "advance to the next row" or "evaluate one candidate" -- so the warp runs
    // for max-over-lanes (rows + candidates) iterations instead of the product of the per-row maxima that nested
    // loops cost (measured: 2400 instructions per warp for ~300 useful per lane).
    int bx0 = 0, bx1 = -1, by0 = 0, by1 = -1, bz0 = 0, bz1 = -1;
    if (scan) {
      if (tau < INFINITY) {
        const float r = sqrtf(fmaxf(tau, 0.0f) + key_margin) * 1.0001f + 1e-7f;
        bx0 = cs_cell(me.x - r, lox, ihx); bx1 = cs_cell(me.x + r, lox, ihx);
        by0 = cs_cell(me.y - r, loy, ihy); by1 = cs_cell(me.y + r, loy, ihy);
        bz0 = cs_cell(me.z - r, loz, ihz); bz1 = cs_cell(me.z + r, loz, ihz);
      } else {
        bx1 = by1 = bz1 = kCsG - 1;
      }
    }
    // pass 1: collect the candidates with key <= tau (typically k+1 .. k+4 of them) in a per-thread column of
    // shared memory (the inbox array, not yet in use); pass 2 ranks them.  The insertion network stays out of the
    // scan loop: some lane of the warp would trigger it on almost every iteration.
    uint16_t* col = &S.inbox[0][0] + i;                        // element c at col[c * kCsThreads]
    int cnt = 0;
    {
      const int ny = by1 - by0 + 1, stride = (scan && helped) ? 2 : 1;
      int cz = bz0, cy = by0 + (helper ? 1 : 0) - stride;      // rows in (cz, cy) order, every stride-th one
      // The bounds of the NEXT row are fetched while the current row's candidates are evaluated: a step is "take the
      // prefetched row if the current one is used up, evaluate up to two candidates", so an exhausted row costs no
      // step of its own and no shared-memory round trip on the critical path.
      int nt = 0, nt1 = 0;
      auto fetch_row = [&]() -> bool {
        cy += stride;
        while (cy > by1) {
          cy -= ny;
          ++cz;
        }
        if (cz > bz1) return false;
        const int row = cs_index(0, cy, cz);
        nt = (int)S.cell[row + bx0];
        nt1 = (int)S.cell[row + bx1 + 1];
        return true;
      };
      bool more = scan && ny > 0;
      bool have_next = more && fetch_row();
      int t = 0, t1 = 0;
      more = have_next;
      while (__any_sync(0xffffffffu, more)) {
        if (more) {
          if (t >= t1) {                                       // current row used up: the prefetched one becomes current
            if (have_next) {
              t = nt;
              t1 = nt1;
              have_next = fetch_row();
            } else {
              more = false;
            }
          }
          if (more && t < t1) {                                // two candidates per step (independent loads and key chains)
            const bool two = t + 1 < t1;
            const float4 c0 = S.spos[t], c1 = S.spos[two ? t + 1 : t];
            const float d0 = add_rn(add_rn(c0.w, dot3_chain(mx, my, mz, c0.x, c0.y, c0.z)), me.w);
            const float d1 = add_rn(add_rn(c1.w, dot3_chain(mx, my, mz, c1.x, c1.y, c1.z)), me.w);
            if (d0 <= tau) {
              if (cnt < kCsInbox) col[cnt * kCsThreads] = S.sorted[t];
              ++cnt;
            }
            if (two && d1 <= tau) {
              if (cnt < kCsInbox) col[cnt * kCsThreads] = S.sorted[t + 1];
              ++cnt;
            }
            t += two ? 2 : 1;
          }
        }
      }
    }
    int cnt2 = 0;
    if (helper) S.hcnt[qi] = (uint16_t)min(cnt, 0xffff);       // hand the helper's column over
    __syncthreads();
    if (live && helped) cnt2 = S.hcnt[i];
    if (live && cnt2 > kCsInbox) cnt = kCsInbox + 1;           // either column overflowed: rescan below, nothing offered yet
    if (live && cnt <= kCsInbox) {
      const uint16_t* col2 = col + n_own;                      // the helper's column
      for (int c0 = 0; c0 < cnt2; ++c0) {
        const int j = col2[c0 * kCsThreads];
        const float4 c = S.pos[j];
        top.offer_lex(knn_key(me.w, c.w, dot3_chain(me.x, me.y, me.z, c.x, c.y, c.z)), j);
      }
      for (int c0 = 0; c0 < cnt; ++c0) {
        const int j = col[c0 * kCsThreads];
        const float4 c = S.pos[j];
        top.offer_lex(knn_key(me.w, c.w, dot3_chain(me.x, me.y, me.z, c.x, c.y, c.z)), j);
      }
    }
    if (live && cnt > kCsInbox) {        // loose bound (first step, or a neighbour moved far away): rank while scanning
      for (int cz = bz0; cz <= bz1; ++cz)
        for (int cy = by0; cy <= by1; ++cy) {
          const int row = cs_index(0, cy, cz);
          const int t1 = (int)S.cell[row + bx1 + 1];
          for (int t = (int)S.cell[row + bx0]; t < t1; ++t) {
            const int j = S.sorted[t];
            const float4 c = S.pos[j];
            if (add_rn(add_rn(c.w, dot3_chain(mx, my, mz, c.x, c.y, c.z)), me.w) <= tau)
              top.offer_lex(knn_key(me.w, c.w, dot3_chain(me.x, me.y, me.z, c.x, c.y, c.z)), j);
          }
        }
    }
  }
  if (CL == 2) {
    if (a.bar_mode == 1) cluster.barrier_wait();          // the peer has started (see the arrive above)
    else if (a.bar_mode == 2) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  }
  if (live) {
    uint32_t w[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int e0 = (unsigned)top.id[2 * s] < (unsigned)K ? top.id[2 * s] : p;
      const int e1 = (unsigned)top.id[2 * s + 1] < (unsigned)K ? top.id[2 * s + 1] : p;
      w[s] = (uint32_t)e0 | ((uint32_t)e1 << 16);
    }
    const uint4 row16 = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(&S.nbrn[p][0]) = row16;           // this CTA's copy
    if (CL == 2) *reinterpret_cast<uint4*>(&P.nbrn[p][0]) = row16;           // the peer's copy (distributed shared memory)
    int4* pv = reinterpret_cast<int4*>(g_nbr + ((size_t)b * K + p) * kCsKK);   // warm start of the next step
    pv[0] = make_int4(top.id[0], top.id[1], top.id[2], top.id[3]);
    pv[1] = make_int4(top.id[4], top.id[5], top.id[6], top.id[7]);
  }
  cluster_sync();                       // both CTAs now hold all K lists

  // ---- own edges: pair terms, mutual test, inbox pushes
  double ax = 0.0, ay = 0.0, az = 0.0;
  float lsum = 0.0f;
  if (live) {
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
#pragma unroll
    for (int s = 1; s < kCsKK; ++s)
      if (s <= k) {
        const int j = (unsigned)top.id[s] < (unsigned)K ? top.id[s] : p;
        float gx, gy, gz, l;
        cs_pair(me, S.pos[j], a.radius, a.h, a.eps, gx, gy, gz, l);       // gradient on j; on i it is the negative
        lsum += l;
        sx -= gx;
        sy -= gy;
        sz -= gz;
        if (cs_row_has(S.nbrn[j], k, p)) {     // j -> p exists too: its gradient on p is bitwise -g
          ax += (double)(-gx);
          ay += (double)(-gy);
          az += (double)(-gz);
        } else {
          CloudStepSmem& O = (CL == 1 || ((int)S.cid[j] >= S.split_cell) == (half == 1)) ? S : P;     // the CTA that owns j
          const uint32_t slot = atomicAdd(&O.inbox_cnt[j], 1u);
          if (slot < (uint32_t)a.inbox_cap) O.inbox[j][slot] = (uint16_t)p;
        }
      }
    ax += (double)sx;
    ay += (double)sy;
    az += (double)sz;
  }
  cluster_sync();                       // all pushes (local and remote) have landed
  float* ls = reinterpret_cast<float*>(&S.cell[0]);             // the grid is no longer needed: per-point loss staging
  ls[i] = 0.0f;

  // ---- hubs: a point with more non-mutual in-edges than inbox slots (an outlier-rich or very uneven cloud) gets
  //      them from a warp-cooperative scan of all K lists -- one thread doing that scan alone would hold the whole
  //      CTA back (measured: one such point tripled the kernel time)
  const int in_cnt = live ? (int)S.inbox_cnt[p] : 0;
  if (in_cnt > a.inbox_cap) {
    const uint32_t h = atomicAdd(&S.nhub, 1u);
    if (h < (uint32_t)kCsMaxHub) S.hub[h] = (uint16_t)p;
    S.inbox[p][0] = (uint16_t)(h < (uint32_t)kCsMaxHub ? h : 0xffffu);
  }
  __syncthreads();
  {
    const int nh = min((int)S.nhub, kCsMaxHub);
    for (int h = warp; h < nh; h += kCsThreads / 32) {
      const int j = S.hub[h];
      const float4 tj = S.pos[j];
      double hx = 0.0, hy = 0.0, hz = 0.0;
      for (int q = lane; q < K; q += 32) {
        if (!cs_row_has(S.nbrn[q], k, j) || cs_row_has(S.nbrn[j], k, q)) continue;
        float gx, gy, gz, l;
        cs_pair(S.pos[q], tj, a.radius, a.h, a.eps, gx, gy, gz, l);
        hx += (double)gx;
        hy += (double)gy;
        hz += (double)gz;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        hx += __shfl_xor_sync(0xffffffffu, hx, o);
        hy += __shfl_xor_sync(0xffffffffu, hy, o);
        hz += __shfl_xor_sync(0xffffffffu, hz, o);
      }
      if (lane == 0) {
        S.hubsum[h][0] = hx;
        S.hubsum[h][1] = hy;
        S.hubsum[h][2] = hz;
      }
    }
    if (nh > 0) __syncthreads();           // block-uniform
  }

  // ---- in-edges from non-mutual sources, ascending source index
  if (live) {
    if (in_cnt <= a.inbox_cap) {
      int last = -1;
      for (int t = 0; t < in_cnt; ++t) {
        int q = 0x7fffffff;
        for (int u = 0; u < in_cnt; ++u) {
          const int e = S.inbox[p][u];
          if (e > last && e < q) q = e;
        }
        last = q;
        float gx, gy, gz, l;
        cs_pair(S.pos[q], me, a.radius, a.h, a.eps, gx, gy, gz, l);
        ax += (double)gx;
        ay += (double)gy;
        az += (double)gz;
      }
    } else if (S.inbox[p][0] != 0xffffu) {     // hub served above
      const int h = S.inbox[p][0];
      ax += S.hubsum[h][0];
      ay += S.hubsum[h][1];
      az += S.hubsum[h][2];
    } else {                                   // more than kCsMaxHub hubs: ordered scan of every list
      for (int q = 0; q < K; ++q) {
        if (!cs_row_has(S.nbrn[q], k, p)) continue;
        bool mutual = false;
#pragma unroll
        for (int s = 1; s < kCsKK; ++s) mutual |= (s <= k) && (top.id[s] == q);
        if (mutual) continue;
        float gx, gy, gz, l;
        cs_pair(S.pos[q], me, a.radius, a.h, a.eps, gx, gy, gz, l);
        ax += (double)gx;
        ay += (double)gy;
        az += (double)gz;
      }
    }
    // ---- Adam (torch 2.11 single-tensor formula, ifd_math.cuh)
    const size_t o3 = cloud3 + (size_t)3 * p;
    float p3[3] = {me.x, me.y, me.z};
    const float r3[3] = {(float)ax, (float)ay, (float)az};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float g = S.gmv[0][3 * p + c] + r3[c] * a.rep_coef;
      float mm = S.gmv[1][3 * p + c], vv = S.gmv[2][3 * p + c];
      adam_update(p3[c], mm, vv, g, a.omb1, a.b2, a.omb2, a.adam_eps, a.sc);
      // results go back through the staging arrays and leave coalesced below (a thread's point is an arbitrary index: written
      // from here, every store instruction of a warp would touch 32 separate sectors)
      S.gmv[0][3 * p + c] = p3[c];
      S.gmv[1][3 * p + c] = mm;
      S.gmv[2][3 * p + c] = vv;
      if (a.rep_grad_out) a.rep_grad_out[o3 + c] = r3[c];
    }
    ls[p] = lsum;                                               // pair-loss sum of point p
  }
  __syncthreads();
  for (int e = i; e < 3 * K; e += kCsThreads) {                 // each CTA writes the points it owns
    if (CL == 1 || ((int)S.cid[e / 3] >= S.split_cell) == (half == 1)) {
      g_xyz[cloud3 + e] = S.gmv[0][e];
      g_m[cloud3 + e] = S.gmv[1][e];
      g_v[cloud3 + e] = S.gmv[2][e];
    }
  }
  // ---- sum of the pair losses in POINT order (the thread <-> point map depends on atomic arrival order), CTA 1's
  //      part crosses to CTA 0
  if (a.loss_part) {
    __syncthreads();
    float l = ls[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if (lane == 0) S.red[6][warp] = l;
    __syncthreads();
    float t = 0.0f;
    if (i == 0) {
      for (int w = 0; w < kCsThreads / 32; ++w) t += S.red[6][w];
      if (half == 1) P.peer_loss = t;
    }
    cluster_sync();
    if (i == 0 && half == 0) a.loss_part[b] = t + S.peer_loss;
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kCsThreads, 1) cloud_step_kernel(const CloudStepArgs a) {
  cloud_step_body<2>(a);
}
__global__ void __launch_bounds__(kCsThreads, 1) cloud_step_solo_kernel(const CloudStepArgs a) { cloud_step_body<1>(a); }

}  // namespace ifd
