// octree.cu -- DistributeOctTree on the device: one CTA per (frame, level) problem.
//
// Replaces OrbExtractor::DistributeOctTree / ExtractorNode::DivideNode
// (orb_extractor.cc:476-742).  The algorithm itself lives in octree_algo.inl, which is also
// compiled for the CPU by tests/host_emul; this file only supplies the launch plumbing.
#include <stdlib.h>

#include "octree_algo.inl"
#include "orbx_kernels.cuh"

namespace orbx {

__device__ __forceinline__ void octree_problem(const FrameGeom& g, const int lev, const int f, const uint32_t* __restrict__ raw_xy,
                                               const uint8_t* __restrict__ raw_sc, const int32_t* __restrict__ cell_strong,
                                               uint32_t* __restrict__ cand_xy, uint8_t* __restrict__ cand_sc,
                                               int32_t* __restrict__ node_of, int32_t* __restrict__ n_cand,
                                               uint32_t* __restrict__ sel_xy, uint8_t* __restrict__ sel_sc,
                                               int32_t* __restrict__ n_sel) {
  extern __shared__ __align__(16) int ot_mem[];
  const LevelGeom& L = g.lv[lev];
  OtWork w;
  ot_carve(w, ot_mem, g.node_cap);
  const size_t cbase = (size_t)f * g.cand_frame_cap + L.cand_off;
  const size_t sbase = (size_t)f * g.sel_frame_cap + L.sel_off;
  int P = n_cand[f * ORBX_MAX_LEVELS + lev];
  if (P > L.cand_cap) P = L.cand_cap;
  int& nsel_sh = w.misc[0];  // (scalars in the same shared object as the tree: one shared-window base for the kernel)
  int& n_keep = w.misc[1];
  // The iniThFAST -> minThFAST retry of orb_extractor.cc:783-801: k_fast emitted every FAST(min, nms)
  // survivor with its cell (in node_of); a cell that holds a survivor with response >= ini keeps only
  // those, any other cell keeps all.  The filtered list is what the reference hands to the quadtree.
  if (threadIdx.x == 0) n_keep = 0;
  __syncthreads();
  {
    const bool filter = g.ini_th > g.min_th;
    const int32_t* strong = cell_strong + (size_t)f * g.total_cells + L.cell_base;
    for (int p0 = 0; p0 < P; p0 += blockDim.x) {
      const int p = p0 + threadIdx.x;
      bool keep = false;
      uint32_t v = 0;
      uint8_t sc = 0;
      if (p < P) {
        v = raw_xy[cbase + p];
        sc = raw_sc[cbase + p];
        keep = !filter || (int)sc >= g.ini_th || !strong[node_of[cbase + p]];
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (m) {
        int wb = 0;
        const int lane = threadIdx.x & 31;
        if (lane == 0) wb = atomicAdd(&n_keep, __popc(m));
        wb = __shfl_sync(0xffffffffu, wb, 0);
        if (keep) {
          const int pos = wb + __popc(m & ((1u << lane) - 1));
          cand_xy[cbase + pos] = v;
          cand_sc[cbase + pos] = sc;
        }
      }
    }
  }
  __syncthreads();
  P = n_keep;
  if (threadIdx.x == 0) n_cand[f * ORBX_MAX_LEVELS + lev] = P;  // the count ORBX_STAGE_CAND reports
  ot_select(cand_xy + cbase, cand_sc + cbase, P, node_of + cbase, w, L.w - 2 * kFastBorder, L.h - 2 * kFastBorder,
            L.n_roots, L.root_hx, L.quota, L.wcell, L.hcell, L.ncols, L.wcell_rcp, L.hcell_rcp, sel_xy + sbase, sel_sc + sbase, &nsel_sh);
  __syncthreads();
  if (threadIdx.x == 0) n_sel[f * ORBX_MAX_LEVELS + lev] = nsel_sh;
}

#ifndef ORBX_OCTREE_MINB
#define ORBX_OCTREE_MINB 5  // 48 registers: 0.193 -> 0.187 ms per 512 frames (4: 0.200, 6: 0.201)
#endif
__global__ void __launch_bounds__(256, ORBX_OCTREE_MINB) k_octree(const __grid_constant__ FrameGeom g, const uint32_t* __restrict__ raw_xy,
                                                const uint8_t* __restrict__ raw_sc, const int32_t* __restrict__ cell_strong,
                                                uint32_t* __restrict__ cand_xy, uint8_t* __restrict__ cand_sc,
                                                int32_t* __restrict__ node_of, int32_t* __restrict__ n_cand,
                                                uint32_t* __restrict__ sel_xy, uint8_t* __restrict__ sel_sc,
                                                int32_t* __restrict__ n_sel) {
  // level-major launch order (level = blockIdx.y, the slow index): the long level-0 problems of all frames start first and
  // the short top-level ones fill the tail of the launch
  octree_problem(g, blockIdx.y, blockIdx.x, raw_xy, raw_sc, cell_strong, cand_xy, cand_sc, node_of, n_cand, sel_xy, sel_sc, n_sel);
}

// the same problem for the levels lev, lev + 1, ... (grid = frames x levels): the single-frame pipeline runs levels as parallel branches
__global__ void __launch_bounds__(1024) k_octree_level(const __grid_constant__ FrameGeom g, const uint32_t* __restrict__ raw_xy,
                                                      const uint8_t* __restrict__ raw_sc, const int32_t* __restrict__ cell_strong,
                                                      uint32_t* __restrict__ cand_xy, uint8_t* __restrict__ cand_sc,
                                                      int32_t* __restrict__ node_of, int32_t* __restrict__ n_cand,
                                                      uint32_t* __restrict__ sel_xy, uint8_t* __restrict__ sel_sc,
                                                      int32_t* __restrict__ n_sel, int lev) {
  octree_problem(g, lev + blockIdx.y, blockIdx.x, raw_xy, raw_sc, cell_strong, cand_xy, cand_sc, node_of, n_cand, sel_xy, sel_sc, n_sel);
}

size_t octree_smem_bytes(int node_cap) { return sizeof(int) * (size_t)ot_work_ints(node_cap); }

cudaError_t octree_configure(int node_cap) {
  cudaError_t e = cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)octree_smem_bytes(node_cap));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_octree_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)octree_smem_bytes(node_cap));
  return e;
}

int launch_octree_levels(const FrameGeom& g, const BatchBuffers& b, int frames, int lev, int lev_end, cudaStream_t st) {
  static const int forced = [] { const char* e = getenv("ORBX_OCTREE_LEVEL_THREADS"); const int t = e ? atoi(e) : 0; return t >= 32 && t <= 1024 && t % 32 == 0 ? t : 0; }();
  // a single frame has one CTA per level: the passes' loops over points and nodes shorten with the thread count until the
  // barriers take over (p50 of the blocking call: 32 threads 0.157 ms, 64: 0.121, 128: 0.108, 256: 0.098, 512: 0.096, 1024: 0.098)
  const int threads = forced ? forced : 512;
  k_octree_level<<<dim3(frames, lev_end - lev), threads, octree_smem_bytes(g.node_cap), st>>>(g, b.cand_raw_xy, b.cand_raw_sc, b.cell_strong, b.cand_xy, b.cand_sc,
                                                                    b.node_of, b.n_cand, b.sel_xy, b.sel_sc, b.n_sel, lev);
  return 1;
}

int launch_octree(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st) {
  dim3 grid(frames, g.nlev);
  // The passes are short and barrier-bound.  Batches: 128-thread CTAs waste fewer idle warps (0.31 vs 0.45 ms per
  // 512 frames); a single frame has only 8 CTAs in flight and wants the shorter point loops of 256 threads
  // (54 vs 65 us).  ORBX_OCTREE_THREADS overrides (A/B runs).
  static const int forced = [] { const char* e = getenv("ORBX_OCTREE_THREADS"); const int t = e ? atoi(e) : 0; return t >= 64 && t <= 256 && t % 32 == 0 ? t : 0; }();
  const int threads = forced ? forced : (frames >= 16 ? 128 : 256);
  k_octree<<<grid, threads, octree_smem_bytes(g.node_cap), st>>>(g, b.cand_raw_xy, b.cand_raw_sc, b.cell_strong, b.cand_xy,
                                                            b.cand_sc, b.node_of, b.n_cand, b.sel_xy, b.sel_sc, b.n_sel);
  return 1;
}

}  // namespace orbx
