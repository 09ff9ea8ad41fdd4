// octree_algo.inl -- deterministic block-parallel DistributeOctTree.
//
// Reproduces OrbExtractor::DistributeOctTree (orb_extractor.cc:542-742) including its node
// list ORDER, without the std::list: the node list is kept as an array in list order and every
// pass rebuilds it with prefix sums.
//
//   reference                                    here
//   ---------                                    ----
//   push_front(child) in creation order          new position = total_children-1-creation_idx
//   untouched nodes keep their relative order    new position = total_children + rank among them
//   coarse pass: split every node with >1 pts    processing order = list order
//   fine phase: stable_sort (count, UL.x) asc,   processing order = (count desc, UL.x desc,
//     walk from the back, stop at size >= N        position asc); prefix sum of (children-1)
//                                                  finds the stopping node
//   best point per node, first wins ties         max response, then min reference-order key
//
// The file is included twice: by octree.cu (one CTA per (frame, level) problem) and by the CPU
// harness tests/host_emul.cc (-DORBX_HOST_EMUL), where OT_FOR degenerates to a serial loop.  The
// harness checks this very source against the oracle on thousands of point sets, so only the
// CUDA plumbing (barriers, atomics, scans) is left to verify on the GPU.
#pragma once

#include "orbx_math.cuh"

#if defined(ORBX_HOST_EMUL)
#define OT_DEV inline
#define OT_HD inline
#define OT_FOR(i, n) for (int i = 0; i < (n); ++i)
#define OT_SYNC() ((void)0)
#define OT_TID0 true
static inline int ot_atomic_add(int* p, int v) { int o = *p; *p += v; return o; }
static inline void ot_atomic_max(int* p, int v) { if (v > *p) *p = v; }
static inline void ot_atomic_min_u(unsigned* p, unsigned v) { if (v < *p) *p = v; }
static inline void ot_atomic_max64(unsigned long long* p, unsigned long long v) { if (v > *p) *p = v; }
#else
#define OT_DEV __device__ __forceinline__
#define OT_HD __host__ __device__ __forceinline__
#define OT_FOR(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)
#define OT_SYNC() __syncthreads()
#define OT_TID0 (threadIdx.x == 0)
__device__ __forceinline__ int ot_atomic_add(int* p, int v) { return atomicAdd(p, v); }
__device__ __forceinline__ void ot_atomic_max(int* p, int v) { atomicMax(p, v); }
__device__ __forceinline__ void ot_atomic_min_u(unsigned* p, unsigned v) { atomicMin(p, v); }
__device__ __forceinline__ void ot_atomic_max64(unsigned long long* p, unsigned long long v) { atomicMax(p, v); }
#endif

namespace orbx {

// Working set of one quadtree problem.  All arrays have `cap` entries (cc/cpos: 4*cap) and live
// in shared memory on the device.
struct OtWork {
  int cap;
  int* tab;     // node tables, double buffered, in list order: x0, y0, x1, y1, cnt of buffer b at tab + (5 * b + k) * cap
                // (addressed by arithmetic, not through an array of pointers: a run-time index into such an array puts
                // the struct in local memory and turns every table access into a generic 64-bit load)
  int* cc;      // [cap][4] child point counts of nodes being split
  int* cpos;    // [cap][4] new list position of each non-empty child
  int* newpos;  // [cap+1]  new list position of nodes that are kept
  int* rank;    // [cap]    processing rank of a node in this pass, -1 = not split
  int* seq;     // [cap]    inverse of rank
  int* scan;    // [cap+1]  scratch for prefix sums
  unsigned long long* best;  // [cap] (response << 32 | ~order key) of the best point of each node
  int* sv;      // small scalar block (>= 16 ints), shared between threads
  int* wsum;    // [2][32] warp totals of the block scan (device only; part of the same shared object as everything else)
  int* misc;    // [2] scalars of the calling kernel
};

OT_DEV int* ot_tab(const OtWork& w, int b, int k) { return w.tab + (5 * b + k) * w.cap; }
enum { OT_X0 = 0, OT_Y0, OT_X1, OT_Y1, OT_CNT };

enum { SV_N = 0, SV_F, SV_M, SV_MEFF, SV_TOTALC, SV_DONE, SV_PHASE, SV_TOEXP, SV_CUR, SV_ERR };

OT_HD int ot_work_ints(int cap) { return cap * (10 + 4 + 4 + 1 + 1 + 2) + 2 * (cap + 1) + 16 + 2 + 64 + 2; }

OT_DEV void ot_carve(OtWork& w, int* mem, int cap) {
  w.cap = cap;
  int* p = mem;
  w.tab = p; p += 10 * cap;
  w.cc = p; p += 4 * cap;
  w.cpos = p; p += 4 * cap;
  w.newpos = p; p += cap + 1;
  w.rank = p; p += cap;
  w.seq = p; p += cap;
  w.scan = p; p += cap + 1;
  p += (p - mem) & 1;  // 8-byte alignment (the block itself is at least 8-byte aligned); index arithmetic keeps the pointer's address space known
  w.best = (unsigned long long*)p; p += 2 * cap;
  w.sv = p; p += 16;
  w.wsum = p; p += 64;
  w.misc = p;
}

// exclusive prefix sum of a[0..n) in place; a[n] receives the total.  Block-wide.
OT_DEV void ot_excl_scan(int* a, int n, int* wsum) {
#if defined(ORBX_HOST_EMUL)
  (void)wsum;
  int s = 0;
  for (int i = 0; i < n; i++) { int v = a[i]; a[i] = s; s += v; }
  a[n] = s;
#else
  // warp-shuffle scan; the running total lives in a register of every thread (read back from the
  // last warp's sum), so a chunk of blockDim.x elements costs two barriers
  int(*warp_sums)[32] = reinterpret_cast<int(*)[32]>(wsum);  // in the kernel's one shared object: no second shared-window base
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  int carry = 0, buf = 0;
  for (int base = 0; base < n; base += nt, buf ^= 1) {
    const int i = base + tid;
    const int v = i < n ? a[i] : 0;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) warp_sums[buf][wid] = s;
    __syncthreads();
    // every warp scans the (<= 16) warp totals itself: no second barrier, no broadcast through shared memory
    int ws = lane < nw ? warp_sums[buf][lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, ws, o);
      if (lane >= o) ws += t;
    }
    const int before_warp = __shfl_sync(0xffffffffu, ws, wid > 0 ? wid - 1 : 0);
    const int total = __shfl_sync(0xffffffffu, ws, nw - 1);
    if (i < n) a[i] = carry + (wid ? before_warp : 0) + s - v;
    carry += total;
    // warp_sums is double buffered, so the next chunk may overwrite the other buffer right away
  }
  if (tid == 0) a[n] = carry;
  __syncthreads();
#endif
}

// child quadrant of point (px,py) in node with bounds (x0,y0,x1,y1): ExtractorNode::DivideNode,
// orb_extractor.cc:476-511.  halfX = ceil((UR.x-UL.x)/2.f).
OT_DEV int ot_quadrant(int px, int py, int x0, int y0, int x1, int y1) {
  const int mx = x0 + ((x1 - x0 + 1) >> 1), my = y0 + ((y1 - y0 + 1) >> 1);
  return (px < mx ? 0 : 1) + (py < my ? 0 : 2);  // 0:n1 1:n2 2:n3 3:n4
}

// Reference-order key of a candidate (x,y relative to (16,16)): position in to_dist_kps, i.e.
// cell-row-major, then row-major inside the cell's detection domain (orb_extractor.cc:767-823).
// v / d for 0 <= v, v * d < 2^32, rcp = ceil(2^32 / d): one multiply-high, exact
OT_DEV int ot_div(int v, int d, unsigned rcp) {
#if defined(ORBX_HOST_EMUL)
  (void)rcp;
  return v / d;
#else
  (void)d;
  return (int)__umulhi((unsigned)v, rcp);
#endif
}
OT_HD unsigned ot_rcp(int d) { return (unsigned)((0x100000000ull + (unsigned long long)d - 1) / (unsigned long long)d); }

OT_DEV unsigned ot_order_key(int x, int y, int wcell, int hcell, int ncols, unsigned wrcp, unsigned hrcp) {
  const int ci = ot_div(y - 3, hcell, hrcp), cj = ot_div(x - 3, wcell, wrcp);
  const int yr = (y - 3) - ci * hcell, xr = (x - 3) - cj * wcell;
  return (unsigned)(((ci * ncols + cj) * hcell + yr) * wcell + xr);
}

// One quadtree problem.  Points: xy[p] = (y << 16) | x relative to (16,16), sc[p] = response.
// node_of[p]: scratch (global memory on the device).  Output: sel_xy/sel_sc in list order with
// LEVEL coordinates (the +16 of orb_extractor.cc:838-839 applied), *n_sel nodes.
// Returns (via *n_sel) -1 if the node table capacity was exceeded.
OT_DEV void ot_select(const uint32_t* xy, const uint8_t* sc, int P, int* node_of, OtWork& w,
                      int width, int height /* max_x-min_x, max_y-min_y */, int n_roots,
                      float root_hx, int quota, int wcell, int hcell, int ncols, unsigned wrcp, unsigned hrcp,
                      uint32_t* sel_xy, uint8_t* sel_sc, int* n_sel) {
  int* sv = w.sv;
  int cur = 0;
  // ---- roots (:548-585) ----
  OT_FOR(i, w.cap) {
    if (i < n_roots) {
      ot_tab(w, 0, OT_X0)[i] = (int)f_mul(root_hx, (float)i);
      ot_tab(w, 0, OT_X1)[i] = (int)f_mul(root_hx, (float)(i + 1));
      ot_tab(w, 0, OT_Y0)[i] = 0;
      ot_tab(w, 0, OT_Y1)[i] = height;
    }
    ot_tab(w, 0, OT_CNT)[i] = 0;
  }
  if (OT_TID0) { sv[SV_DONE] = 0; sv[SV_PHASE] = 0; sv[SV_F] = 0; sv[SV_ERR] = (n_roots > w.cap); }
  OT_SYNC();
  if (P <= 0 || sv[SV_ERR]) {
    if (OT_TID0) *n_sel = sv[SV_ERR] ? -1 : 0;
    return;
  }
  OT_FOR(p, P) {
    int r = (int)f_div((float)(xy[p] & 0xFFFFu), root_hx);
    if (r >= n_roots) r = n_roots - 1;
    node_of[p] = r;
    ot_atomic_add(&ot_tab(w, 0, OT_CNT)[r], 1);
  }
  OT_SYNC();
  // erase empty roots, keep order
  OT_FOR(i, n_roots) w.scan[i] = ot_tab(w, 0, OT_CNT)[i] > 0;
  OT_SYNC();
  ot_excl_scan(w.scan, n_roots, w.wsum);
  OT_FOR(i, n_roots) {
    if (ot_tab(w, 0, OT_CNT)[i] > 0) {
      const int j = w.scan[i];
      for (int k = OT_X0; k <= OT_CNT; k++) ot_tab(w, 1, k)[j] = ot_tab(w, 0, k)[i];
    }
  }
  OT_SYNC();
  OT_FOR(p, P) node_of[p] = w.scan[node_of[p]];
  if (OT_TID0) sv[SV_N] = w.scan[n_roots];
  OT_SYNC();
  cur = 1;

  // ---- subdivision passes ----
  for (;;) {
    const int n = sv[SV_N];
    const int F = sv[SV_F];
    const int fine = sv[SV_PHASE];
    int *X0 = ot_tab(w, cur, OT_X0), *Y0 = ot_tab(w, cur, OT_Y0), *X1 = ot_tab(w, cur, OT_X1), *Y1 = ot_tab(w, cur, OT_Y1), *CNT = ot_tab(w, cur, OT_CNT);
    int *NX0 = ot_tab(w, cur ^ 1, OT_X0), *NY0 = ot_tab(w, cur ^ 1, OT_Y0), *NX1 = ot_tab(w, cur ^ 1, OT_X1), *NY1 = ot_tab(w, cur ^ 1, OT_Y1);
    int* NCNT = ot_tab(w, cur ^ 1, OT_CNT);

    // processing rank of the nodes split in this pass
    if (!fine) {
      OT_FOR(i, n) w.scan[i] = CNT[i] > 1;  // !no_more_ (:597-603)
      OT_SYNC();
      ot_excl_scan(w.scan, n, w.wsum);
      OT_FOR(i, n) {
        const int r = CNT[i] > 1 ? w.scan[i] : -1;
        w.rank[i] = r;
        if (r >= 0) w.seq[r] = i;
      }
      if (OT_TID0) sv[SV_M] = w.scan[n];
      OT_SYNC();
    } else {
      // candidates: children created by the previous pass that hold > 1 point (kps_size_and_nd,
      // :616-650); processed from the back of stable_sort((count, UL.x) ascending) (:671-673)
      // key = (count, UL.x) packed; 0 marks a node that is not a candidate (candidates have count >= 2)
      unsigned long long* key = w.best;  // free until the best-point pass at the end
      OT_FOR(i, n) key[i] = (i < F && CNT[i] > 1) ? (((unsigned long long)CNT[i] << 32) | (unsigned)X0[i]) : 0ull;
      OT_SYNC();
      OT_FOR(i, n) {
        int r = -1;
        const unsigned long long ki = key[i];
        if (ki) {
          r = 0;
#pragma unroll 8
          for (int j = 0; j < F; j++) {
            const unsigned long long kj = key[j];
            r += (int)(kj > ki) + (int)((kj == ki) & (j < i));  // processed before node i
          }
        }
        w.rank[i] = r;
        w.scan[i] = (r >= 0);
      }
      OT_SYNC();
      OT_FOR(i, n) if (w.rank[i] >= 0) w.seq[w.rank[i]] = i;
      ot_excl_scan(w.scan, n, w.wsum);
      if (OT_TID0) sv[SV_M] = w.scan[n];
      OT_SYNC();
    }
    const int m = sv[SV_M];
    if (m == 0) break;  // nothing left to split: size == size_prev (:660 / :716)

    // child point counts of every node that may be split
    OT_FOR(i, 4 * n) w.cc[i] = 0;
    OT_SYNC();
    OT_FOR(p, P) {
      const int i = node_of[p];
      if (w.rank[i] >= 0) {
        const uint32_t v = xy[p];
        ot_atomic_add(&w.cc[4 * i + ot_quadrant((int)(v & 0xFFFFu), (int)(v >> 16), X0[i], Y0[i], X1[i], Y1[i])], 1);
      }
    }
    OT_SYNC();

    // how many nodes are actually processed: the fine phase stops at the first node after which
    // the list holds >= quota nodes (:713)
    OT_FOR(r, m) {
      const int* c = &w.cc[4 * w.seq[r]];
      w.scan[r] = (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0);
    }
    OT_SYNC();
    ot_excl_scan(w.scan, m, w.wsum);  // scan[r] = children created before node r, scan[m] = total
    if (OT_TID0) sv[SV_MEFF] = m;
    OT_SYNC();
    if (fine) {
      // list size after processing nodes 0..r:  n + (scan[r+1] - (r+1))
      OT_FOR(r, m) {
        const int after = n + w.scan[r + 1] - (r + 1);
        const int before = n + w.scan[r] - r;
        if (after >= quota && before < quota) sv[SV_MEFF] = r + 1;  // unique r: sizes are monotone
      }
      OT_SYNC();
    }
    const int meff = sv[SV_MEFF];
    const int totalc = w.scan[meff];
    if (totalc + (n - meff) > w.cap) {  // cannot happen for cap >= quota+4; defensive
      if (OT_TID0) sv[SV_ERR] = 1;
      OT_SYNC();
      break;
    }

    // new list positions
    OT_FOR(r, meff) {
      const int i = w.seq[r];
      int c = w.scan[r];
      for (int k = 0; k < 4; k++) {
        if (w.cc[4 * i + k] > 0) { w.cpos[4 * i + k] = totalc - 1 - c; c++; }
        else w.cpos[4 * i + k] = -1;
      }
    }
    OT_SYNC();
    // kept nodes: rank among the kept ones in list order (scan buffer is reused, so stash totals)
    OT_FOR(i, n) w.newpos[i] = !(w.rank[i] >= 0 && w.rank[i] < meff);
    OT_SYNC();
    ot_excl_scan(w.newpos, n, w.wsum);
    OT_FOR(i, n) {
      const bool split = w.rank[i] >= 0 && w.rank[i] < meff;
      if (split) {
        const int mx = X0[i] + ((X1[i] - X0[i] + 1) >> 1), my = Y0[i] + ((Y1[i] - Y0[i] + 1) >> 1);
        for (int k = 0; k < 4; k++) {
          const int d = w.cpos[4 * i + k];
          if (d < 0) continue;
          NX0[d] = (k & 1) ? mx : X0[i];
          NX1[d] = (k & 1) ? X1[i] : mx;
          NY0[d] = (k & 2) ? my : Y0[i];
          NY1[d] = (k & 2) ? Y1[i] : my;
          NCNT[d] = w.cc[4 * i + k];
        }
      } else {
        const int d = totalc + w.newpos[i];
        NX0[d] = X0[i]; NY0[d] = Y0[i]; NX1[d] = X1[i]; NY1[d] = Y1[i]; NCNT[d] = CNT[i];
      }
    }
    OT_FOR(p, P) {
      const int i = node_of[p];
      if (w.rank[i] >= 0 && w.rank[i] < meff) {
        const uint32_t v = xy[p];
        node_of[p] = w.cpos[4 * i + ot_quadrant((int)(v & 0xFFFFu), (int)(v >> 16), X0[i], Y0[i], X1[i], Y1[i])];
      } else {
        node_of[p] = totalc + w.newpos[i];
      }
    }
    if (OT_TID0) sv[SV_TOEXP] = 0;
    OT_SYNC();
    const int n_new = totalc + (n - meff);
    // nodes to expand next: fresh children with more than one point (:612-650).  Only their number is
    // needed (the order comes from the ranks of the next pass): a count, not a scan.
    OT_FOR(i, totalc) if (NCNT[i] > 1) ot_atomic_add(&sv[SV_TOEXP], 1);
    OT_SYNC();
    if (OT_TID0) {
      const int to_expand = sv[SV_TOEXP];
      sv[SV_N] = n_new;
      sv[SV_F] = totalc;
      if (n_new >= quota || n_new == n) sv[SV_DONE] = 1;                 // :660 / :716
      else if (!fine && n_new + 3 * to_expand > quota) sv[SV_PHASE] = 1;  // :662
    }
    OT_SYNC();
    cur ^= 1;
    if (sv[SV_DONE]) break;
  }

  // ---- best point per node (:723-739) ----
  const int n = sv[SV_N];
  if (sv[SV_ERR]) {
    if (OT_TID0) *n_sel = -1;
    return;
  }
  // max response, then the earliest point in the reference's candidate order: one 64-bit key per point
  OT_FOR(i, n) w.best[i] = 0ull;
  OT_SYNC();
  OT_FOR(p, P) {
    const uint32_t v = xy[p];
    const unsigned okey = ot_order_key((int)(v & 0xFFFFu), (int)(v >> 16), wcell, hcell, ncols, wrcp, hrcp);
    ot_atomic_max64(&w.best[node_of[p]], ((unsigned long long)(sc[p] + 1u) << 32) | (unsigned long long)(0xFFFFFFFFu - okey));
  }
  OT_SYNC();
  OT_FOR(p, P) {
    const uint32_t v = xy[p];
    const int x = (int)(v & 0xFFFFu), y = (int)(v >> 16);
    const unsigned okey = ot_order_key(x, y, wcell, hcell, ncols, wrcp, hrcp);
    const int i = node_of[p];
    if (w.best[i] == (((unsigned long long)(sc[p] + 1u) << 32) | (unsigned long long)(0xFFFFFFFFu - okey))) {
      sel_xy[i] = ((uint32_t)(y + kFastBorder) << 16) | (uint32_t)(x + kFastBorder);
      sel_sc[i] = sc[p];
    }
  }
  if (OT_TID0) *n_sel = n;
  OT_SYNC();
}

}  // namespace orbx
