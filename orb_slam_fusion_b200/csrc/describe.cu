// describe.cu -- orientation, steered rBRIEF descriptors and output assembly.
//
// Replaces computeOrientation / IC_Angle (orb_extractor.cc:76-100, 467-474),
// ComputeDescriptors / ComputeOrbDescriptor (:102-146, 1002-1009) and the output loop of
// operator() (:1030-1090: scale to level-0 pixels, mono block from the front, lapping block
// from the back).  Arithmetic: SURVEY.md A.5, A.7, A.8.
//
//   k_plan     one CTA per frame: output slot of every selected keypoint (prefix sum of the
//              lapping flags in the reference's traversal order), n and n_mono.
//   k_describe one warp per keypoint: the intensity centroid over the patch's rows read as aligned 64-bit units
//              (DP4A with signed per-byte weights), then lane L builds descriptor byte L
//              (16 gathers from the blurred level) and the warp stores the 32-byte row at once.
#include "orbx_kernels.cuh"
#include "orbx_math.cuh"

namespace orbx {

__constant__ int8_t c_pattern[1024] = {
#include "../../include/orb_pattern31.inc"
};
// umax_ of orb_extractor.cc:452-464 for kHalfPatchSize = 15
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// IC_Angle as DP4A over aligned 64-bit units: the 31 patch rows are read as 5 aligned 8-byte units each (40 bytes cover
// the 31 columns for any alignment a = 0..7 of the patch's first column); item i = 5 * row + unit.  Entry [a][i] holds
// the per-byte SIGNED weights of that unit inside the radius-15 disc -- (x, y): u per byte of the low / high word,
// (z, w): v per byte (0 outside the disc) -- so that m10 and m01 are accumulated by DP4A itself (dp4a.u32.s32 with the
// running sum as addend): four instructions per 8 pixels and nothing else.  Filled once per device by k_pattern_init;
// 8 x 155 x 16 B = 19.8 KB, read by every warp with one coalesced 128-bit load per item (the pixel loads do not depend
// on it: row and unit come from the item index).
constexpr int kOriItems = 31 * 5;
__device__ uint4 g_ori_w[8 * kOriItems];
__device__ __forceinline__ int dp4a_us(uint32_t pix, uint32_t w, int acc) {  // unsigned pixels x signed weights
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pix), "r"(w), "r"(acc));
  return d;
}

// the same pattern as floats, transposed: entry [k][L] = (x0, y0, x1, y1) of bit k of descriptor byte L, so that a
// warp reads one coalesced 512-byte line per bit (filled once per device by k_pattern_init)
__device__ float4 g_pattern_f[8 * 32];
__global__ void k_pattern_init() {
  const int t = threadIdx.x;  // 256 threads = 32 bytes x 8 bits
  const int L = t >> 3, k = t & 7;
  const int8_t* p = &c_pattern[(L * 8 + k) * 4];
  g_pattern_f[k * 32 + L] = make_float4((float)p[0], (float)p[1], (float)p[2], (float)p[3]);
  for (int e = t; e < 8 * kOriItems; e += blockDim.x) {
    const int a = e / kOriItems, i = e - a * kOriItems, r = i / 5, c = i - r * 5;
    const int v = r - kHalfPatch, d = c_umax[v < 0 ? -v : v];
    const int u0 = 8 * c - a - kHalfPatch;  // u of byte 0 of this unit
    uint32_t wu[2] = {0, 0}, wv[2] = {0, 0};
    for (int j = 0; j < 8; j++) {
      const int u = u0 + j;
      if (u >= -d && u <= d) {
        wu[j >> 2] |= (uint32_t)(uint8_t)(int8_t)u << (8 * (j & 3));
        wv[j >> 2] |= (uint32_t)(uint8_t)(int8_t)v << (8 * (j & 3));
      }
    }
    g_ori_w[e] = make_uint4(wu[0], wu[1], wv[0], wv[1]);
  }
}

constexpr int kPlanThreads = 256;
constexpr int kWorkLevShift = 24;  // work[] entry = output slot | level << 24 (slots < 2^24), or -1

__global__ void __launch_bounds__(kPlanThreads) k_plan(const __grid_constant__ FrameGeom g, const uint32_t* __restrict__ sel_xy,
                                                       const int32_t* __restrict__ n_sel, int32_t* __restrict__ work,
                                                       int cap, int32_t* __restrict__ n_out, int32_t* __restrict__ n_mono_out,
                                                       int out_frame0) {
  __shared__ int lev_start[ORBX_MAX_LEVELS + 1];
  __shared__ int warp_sums[kPlanThreads / 32];
  __shared__ int carry, bad;
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {
    int s = 0, b = 0;
    for (int l = 0; l < g.nlev; l++) {
      int n = n_sel[f * ORBX_MAX_LEVELS + l];
      if (n < 0) { b = 1; n = 0; }
      lev_start[l] = s;
      s += n;
    }
    lev_start[g.nlev] = s;
    carry = 0;
    bad = b;
  }
  __syncthreads();
  const int N = lev_start[g.nlev];
  int32_t* wk = work + (size_t)f * g.sel_frame_cap;
  for (int i = tid; i < g.sel_frame_cap; i += kPlanThreads) wk[i] = -1;  // entries no keypoint owns
  __syncthreads();
  const uint32_t* sxy = sel_xy + (size_t)f * g.sel_frame_cap;
  const bool fits = N <= cap && !bad;
  // No keypoint lies left of column 19 (level coordinates, scaled up by >= 1), so a lapping area that ends before it --
  // {0, 0}, every camera that is not a fisheye pair -- flags nothing: slots are the traversal order itself, no scan
  const bool no_laps = g.lap1 < kEdge;
  if (no_laps) {
    for (int i = tid; i < N; i += kPlanThreads) {
      int lev = 0;
      while (i >= lev_start[lev + 1]) lev++;
      wk[g.lv[lev].sel_off + (i - lev_start[lev])] = fits ? (i | (lev << kWorkLevShift)) : -1;
    }
  }
  for (int base = 0; base < (no_laps ? 0 : N); base += kPlanThreads) {
    const int i = base + tid;
    int flag = 0, lev = 0, idx = 0;
    if (i < N) {
      while (i >= lev_start[lev + 1]) lev++;
      idx = i - lev_start[lev];
      float x = (float)(sxy[g.lv[lev].sel_off + idx] & 0xFFFFu);
      if (lev != 0) x = f_mul(x, g.lv[lev].scale);                 // :1071-1073
      flag = (x >= (float)g.lap0 && x <= (float)g.lap1) ? 1 : 0;    // :1075-1076
    }
    // block-wide exclusive scan of the flags
    int s = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) warp_sums[wid] = s;
    __syncthreads();
    int before = carry;
    for (int w = 0; w < wid; w++) before += warp_sums[w];
    const int laps_before = before + s - flag;
    if (i < N) {
      // lapping keypoints fill the tail backwards (stereoIndex--), the rest the head (monoIndex++)
      const int slot = flag ? (N - 1 - laps_before) : (i - laps_before);
      wk[g.lv[lev].sel_off + idx] = fits ? (slot | (lev << kWorkLevShift)) : -1;  // k_describe reads slot and level from here
    }
    __syncthreads();
    if (tid == kPlanThreads - 1) carry = before + s;
    __syncthreads();
  }
  if (tid == 0) {
    n_out[out_frame0 + f] = bad ? INT32_MIN : (fits ? N : -N);
    n_mono_out[out_frame0 + f] = fits ? N - carry : 0;
  }
}

#ifndef ORBX_DESC_WARPS
#define ORBX_DESC_WARPS 8
#endif
constexpr int kDescWarps = ORBX_DESC_WARPS;
constexpr int kPatchRows = 37;    // rotated pattern offsets stay within +-18 px (A.7)
constexpr int kPatchPitch = 80;   // bytes per staged row: 4 x 16-byte chunks from a 16-byte aligned column (<= 15 + 37 bytes used);
                                  // 20 words of pitch spread vertical neighbours over the banks

#ifndef ORBX_DESC_MINB
#define ORBX_DESC_MINB 8
#endif
__global__ void __launch_bounds__(32 * kDescWarps, ORBX_DESC_MINB) k_describe(const __grid_constant__ FrameGeom g, const uint8_t* __restrict__ pyr,
                                                              const uint8_t* __restrict__ blur, const uint32_t* __restrict__ sel_xy,
                                                              const uint8_t* __restrict__ sel_sc, const int32_t* __restrict__ n_sel,
                                                              const int32_t* __restrict__ work, orbx_kp* __restrict__ kps,
                                                              uint8_t* __restrict__ desc, int cap, int out_frame0) {
  // per warp: the 37 x 37 blurred patch around the keypoint, rows of kPatchPitch bytes
  __shared__ __align__(16) uint8_t patch[kDescWarps][kPatchRows * kPatchPitch];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int f = blockIdx.y;
  const int s = blockIdx.x * kDescWarps + wid;  // position in the frame's selected list
  if (s >= g.sel_frame_cap) return;
  const size_t so = (size_t)f * g.sel_frame_cap + s;
  const int wv = work[so];  // slot | level << 24 from k_plan; -1: no keypoint here (or it does not fit)
  const uint32_t xy = sel_xy[so];  // (issued together with the load above)
  if (wv < 0) return;
  const int slot = wv & ((1 << kWorkLevShift) - 1), lev = wv >> kWorkLevShift;
  const LevelGeom& L = g.lv[lev];
  const int pitch = L.pitch;  // level fields live in the parameter bank behind a run-time index: read once
  const int cx = (int)(xy & 0xFFFFu), cy = (int)(xy >> 16);
  const size_t fo = (size_t)f * g.pyr_frame_bytes;

  // Blurred 37 x 37 patch for the descriptor: issue its coalesced word loads first so that they are in
  // flight together with the orientation loads (the 512 samples of a descriptor are single-byte
  // gathers over 37 rows: straight from global memory each would cost one L1 wavefront per lane).
  const int xb = (cx - 18) & ~15;  // >= 0: keypoints are >= 19 px inside; interior rows are 16-byte aligned
  uint8_t* pw = patch[wid];
  {
    // 16-byte LDGSTS: global -> shared without staging registers; completes while the orientation is computed
    const uint8_t* bsrc = blur + fo + px_off(L, xb, cy - 18);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(pw);
    // (staging three chunks per row when they cover the 37 columns -- 3 in 4 keypoints -- moves a quarter less data but was
    // measured slower, 0.4375 vs 0.4323 ms per 512 frames: the kernel is bound by instruction issue and L2 latency, not bytes)
#pragma unroll
    for (int t = 0; t < (4 * kPatchRows + 31) / 32; t++) {
      const int i = lane + 32 * t;
      const int r = i >> 2, c = i & 3;
      if (i < 4 * kPatchRows)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sbase + (unsigned)(r * kPatchPitch + 16 * c)),
                     "l"(bsrc + r * pitch + 16 * c));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  }

  // ---- IC_Angle (:76-100): m10 = sum u*I, m01 = sum v*I over the radius-15 disc.  The 31 rows are read as
  // aligned 64-bit units (5 per row cover any alignment): item = (row v, unit c), 155 items over the 32
  // lanes = 5 loads per lane instead of 31 byte loads; DP4A with per-byte signed weights (u, v inside the disc, else 0).
  int m10 = 0, m01 = 0;
  {
    const int xa = (cx - kHalfPatch) & ~7, a = (cx - kHalfPatch) - xa;  // aligned start, 0..7 bytes before the patch
    // 8-byte aligned: rows start on 16-byte boundaries (plane rows, and the caller's rows when level 0 is read in place)
    const bool ext = lev == 0 && g.ext0 != nullptr;
    const uint8_t* base = ext ? g.ext0 + (size_t)f * g.ext0_frame + (size_t)(cy - kHalfPatch) * g.ext0_pitch + xa
                              : pyr + fo + px_off(L, xa, cy - kHalfPatch);
    const int rpitch = ext ? (int)g.ext0_pitch : pitch;
    const uint4* wt = g_ori_w + a * kOriItems;
#pragma unroll
    for (int t = 0; t < (kOriItems + 31) / 32; t++) {
      const int i = lane + 32 * t;
      if (i < kOriItems) {
        const int r = (i * 205) >> 10, c = i - r * 5;  // i / 5 for i < 1024
        const uint2 w = __ldg(reinterpret_cast<const uint2*>(base + r * rpitch) + c);
        const uint4 e = __ldg(wt + i);
        m10 = dp4a_us(w.y, e.y, dp4a_us(w.x, e.x, m10));
        m01 = dp4a_us(w.y, e.w, dp4a_us(w.x, e.z, m01));
      }
    }
  }
  m10 = __reduce_add_sync(0xffffffffu, m10);  // REDUX.SUM: one instruction instead of a 5-step shuffle tree
  m01 = __reduce_add_sync(0xffffffffu, m01);
  const float angle = fast_atan2_deg((float)m01, (float)m10);

  // ---- steered rBRIEF (:102-146).  cos/sin are the correctly rounded float values
  // (float of the double-precision result); see SURVEY.md A.7 for the libm tolerance class.
  const float factor_pi = (float)(3.1415926535897932384626433832795 / 180.0);
  const float rad = f_mul(angle, factor_pi);
  double sd, cd;
  sincos((double)rad, &sd, &cd);
  const float a = (float)cd, b = (float)sd;
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncwarp();
  // cvRound by the magic-number add (orbx_math.cuh f_round): the integer sits in the low mantissa bits above the bias
  // 0x4B400000.  row * pitch + col is formed from the BIASED integers with one IMAD and the bias (constant) is folded
  // into the base pointer, two subtractions less per sample point.
  constexpr int kBias = 0x4B400000;
  // (32-bit shared-memory addresses, arithmetic modulo 2^32)
  const uint32_t bc = (uint32_t)__cvta_generic_to_shared(pw) + (uint32_t)(18 * kPatchPitch + (cx - xb)) -
                      ((uint32_t)kBias * (uint32_t)kPatchPitch + (uint32_t)kBias);
  uint32_t byte = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const float4 p = __ldg(&g_pattern_f[k * 32 + lane]);
    // orb_extractor.cc:108-113: row = cvRound(x*b + y*a), col = cvRound(x*a - y*b), float32, no FMA
    const unsigned r0 = (unsigned)__float_as_int(__fadd_rn(f_add(f_mul(p.x, b), f_mul(p.y, a)), 12582912.0f));
    const unsigned c0 = (unsigned)__float_as_int(__fadd_rn(f_sub(f_mul(p.x, a), f_mul(p.y, b)), 12582912.0f));
    const unsigned r1 = (unsigned)__float_as_int(__fadd_rn(f_add(f_mul(p.z, b), f_mul(p.w, a)), 12582912.0f));
    const unsigned c1 = (unsigned)__float_as_int(__fadd_rn(f_sub(f_mul(p.z, a), f_mul(p.w, b)), 12582912.0f));
    uint32_t t0, t1;
    asm("ld.shared.u8 %0, [%1];\n" : "=r"(t0) : "r"(bc + r0 * (uint32_t)kPatchPitch + c0));
    asm("ld.shared.u8 %0, [%1];\n" : "=r"(t1) : "r"(bc + r1 * (uint32_t)kPatchPitch + c1));
    byte |= (uint32_t)(t0 < t1) << k;
  }
  const size_t o = (size_t)(out_frame0 + f) * cap + slot;
  desc[o * 32 + lane] = (uint8_t)byte;

  // ---- keypoint record (:834-843, 1071-1073): lane L stores float L of the 28-byte record; the value is picked by a
  // chain of selects (no per-lane branches)
  {
    float fx = (float)cx, fy = (float)cy;
    if (lev != 0) { fx = f_mul(fx, L.scale); fy = f_mul(fy, L.scale); }
    float v = __int_as_float(-1);                       // class_id
    v = lane == 5 ? __int_as_float(lev) : v;            // octave
    v = lane == 4 ? (float)sel_sc[so] : v;              // response
    v = lane == 3 ? angle : v;
    v = lane == 2 ? (float)L.scaled_patch : v;          // size
    v = lane == 1 ? fy : v;
    v = lane == 0 ? fx : v;
    if (lane < 7) reinterpret_cast<float*>(kps + o)[lane] = v;
  }
}

int launch_pattern_init(cudaStream_t st) {
  k_pattern_init<<<1, 256, 0, st>>>();
  return 1;
}

int launch_describe(const FrameGeom& g, const BatchBuffers& b, int frames, orbx_kp* kps, uint8_t* desc, int cap,
                    int32_t* n, int32_t* n_mono, int out_frame0, cudaStream_t st) {
  k_plan<<<frames, kPlanThreads, 0, st>>>(g, b.sel_xy, b.n_sel, b.work, cap, n, n_mono, out_frame0);
  dim3 grid((g.sel_frame_cap + kDescWarps - 1) / kDescWarps, frames);
  k_describe<<<grid, 32 * kDescWarps, 0, st>>>(g, b.pyr, b.blur, b.sel_xy, b.sel_sc, b.n_sel, b.work, kps, desc, cap,
                                              out_frame0);
  return 2;
}

}  // namespace orbx
