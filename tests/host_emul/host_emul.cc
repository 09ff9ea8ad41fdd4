// tests/host_emul/host_emul.cc -- CPU harness over the PRODUCT's shared kernel sources.
//
// There is no GPU in the development container, so the arithmetic that the CUDA kernels share
// (orb_slam_fusion_b200/csrc/orbx_math.cuh) and the block-parallel quadtree
// (octree_algo.inl, here with OT_FOR degenerating to a serial loop) are compiled with g++ and
// compared against the oracle by tests/test_host_emul.py.  This is a TEST of product source;
// nothing in the product ever calls it.
#define ORBX_HOST_EMUL 1
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../orb_slam_fusion_b200/csrc/orbx_math.cuh"
#include "../../orb_slam_fusion_b200/csrc/octree_algo.inl"

using namespace orbx;

extern "C" {

float emul_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }

int emul_fast9_score(const uint8_t* p, long stride, int t) {
  static const int dx[16] = ORBX_RING_DX, dy[16] = ORBX_RING_DY;
  int r[16];
  for (int k = 0; k < 16; k++) r[k] = p[dy[k] * stride + dx[k]];
  return fast9_score(p[0], r, t);
}

int emul_has_run9(unsigned m) { return has_run9(m); }

void emul_rbrief_offset(float a, float b, int px, int py, int* row, int* col) { rbrief_offset(a, b, px, py, *row, *col); }

int emul_resize_vcombine(int h0, int h1, int b0, int b1) { return resize_vcombine(h0, h1, b0, b1); }

uint64_t emul_splitmix64(uint64_t x) { return splitmix64(x); }

// quadtree on (x, y, response) int triples relative to (16,16) of a w x h level; writes the
// selected (x+16, y+16, response) in list order; returns the count (or -1).
int emul_octree(const int* xyr, int n, int w, int h, int quota, int wcell, int hcell, int ncols,
                int* out_xyr, int cap_out) {
  std::vector<uint32_t> xy(n ? n : 1);
  std::vector<uint8_t> sc(n ? n : 1);
  for (int i = 0; i < n; i++) {
    xy[i] = ((uint32_t)xyr[3 * i + 1] << 16) | (uint32_t)xyr[3 * i];
    sc[i] = (uint8_t)xyr[3 * i + 2];
  }
  const int width = (w - 16) - 16, height = (h - 16) - 16;
  const int n_roots = (int)roundf((float)width / (float)height);
  if (n_roots < 1) return -1;
  const float hx = (float)width / n_roots;
  int cap = quota + 4;
  if (4 * n_roots > cap) cap = 4 * n_roots;
  cap += 1;
  std::vector<int> mem(ot_work_ints(cap));
  OtWork wk;
  ot_carve(wk, mem.data(), cap);
  std::vector<int> node_of(n ? n : 1);
  std::vector<uint32_t> sxy(cap);
  std::vector<uint8_t> ssc(cap);
  int nsel = 0;
  ot_select(xy.data(), sc.data(), n, node_of.data(), wk, width, height, n_roots, hx, quota, wcell, hcell,
            ncols, ot_rcp(wcell), ot_rcp(hcell), sxy.data(), ssc.data(), &nsel);
  for (int i = 0; i < nsel && i < cap_out; i++) {
    out_xyr[3 * i] = (int)(sxy[i] & 0xFFFF);
    out_xyr[3 * i + 1] = (int)(sxy[i] >> 16);
    out_xyr[3 * i + 2] = ssc[i];
  }
  return nsel;
}

}  // extern "C"
