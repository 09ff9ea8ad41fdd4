// octree.cu -- DistributeOctTree on the device: one CTA per (frame, level) problem.
//
// Replaces OrbExtractor::DistributeOctTree / ExtractorNode::DivideNode
// (orb_extractor.cc:476-742).  The algorithm itself lives in octree_algo.inl, which is also
// compiled for the CPU by tests/host_emul; this file only supplies the launch plumbing.
#include "octree_algo.inl"
#include "orbx_kernels.cuh"

namespace orbx {

__global__ void __launch_bounds__(256) k_octree(const __grid_constant__ FrameGeom g, const uint32_t* __restrict__ cand_xy,
                                                const uint8_t* __restrict__ cand_sc, int32_t* __restrict__ node_of,
                                                const int32_t* __restrict__ n_cand, uint32_t* __restrict__ sel_xy,
                                                uint8_t* __restrict__ sel_sc, int32_t* __restrict__ n_sel) {
  extern __shared__ int ot_mem[];
  const int lev = blockIdx.x, f = blockIdx.y;
  const LevelGeom& L = g.lv[lev];
  OtWork w;
  ot_carve(w, ot_mem, g.node_cap);
  const size_t cbase = (size_t)f * g.cand_frame_cap + L.cand_off;
  const size_t sbase = (size_t)f * g.sel_frame_cap + L.sel_off;
  int P = n_cand[f * ORBX_MAX_LEVELS + lev];
  if (P > L.cand_cap) P = L.cand_cap;
  __shared__ int nsel_sh;
  ot_select(cand_xy + cbase, cand_sc + cbase, P, node_of + cbase, w, L.w - 2 * kFastBorder, L.h - 2 * kFastBorder,
            L.n_roots, L.root_hx, L.quota, L.wcell, L.hcell, L.ncols, sel_xy + sbase, sel_sc + sbase, &nsel_sh);
  __syncthreads();
  if (threadIdx.x == 0) n_sel[f * ORBX_MAX_LEVELS + lev] = nsel_sh;
}

size_t octree_smem_bytes(int node_cap) { return sizeof(int) * (size_t)ot_work_ints(node_cap); }

cudaError_t octree_configure(int node_cap) {
  return cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)octree_smem_bytes(node_cap));
}

int launch_octree(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st) {
  dim3 grid(g.nlev, frames);
  k_octree<<<grid, 256, octree_smem_bytes(g.node_cap), st>>>(g, b.cand_xy, b.cand_sc, b.node_of, b.n_cand, b.sel_xy,
                                                            b.sel_sc, b.n_sel);
  return 1;
}

}  // namespace orbx
