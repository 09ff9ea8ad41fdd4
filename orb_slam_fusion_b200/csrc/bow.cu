// bow.cu -- bag-of-words transform of descriptor batches on the device.
//
// Replaces, for whole batches of frames, Frame::ComputeBoW (src/map/frame.cc:761-766) =
// TemplatedVocabulary<FORB>::transform(features, BowVector, FeatureVector, levelsup)
// (3rdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1056-1118) with the per-feature tree descent of :1139-1179,
// FORB::distance (FORB.cpp:71-88), BowVector::addWeight / addIfNotExist / normalize (BowVector.cpp:30-70) and
// FeatureVector::addFeature (FeatureVector.cpp:28-38).
//
//   k_bow_descend  one thread per descriptor: the query stays in 8 registers, the <= k children of the
//                  current node are contiguous 32-byte rows of the slot-ordered node table (one 32-byte
//                  sector each, L2-resident: 35 MB for k=10, L=6), 8 POPC per child, strict "<" keeps the
//                  first minimum like the reference's loop.
//   k_bow_frame    one CTA per frame: the std::map insertions of the reference are a sort of
//                  (word id, feature index) keys in shared memory (bitonic, 64-bit keys); every word's value
//                  is accumulated in feature order and the L1 / L2 norm is summed in increasing word id by
//                  one lane, so every double is produced by the same sequence of IEEE operations as on the
//                  CPU (the library is built with --fmad=false); the FeatureVector is the same sort on
//                  (node id, feature index).
#include "orbx_kernels.cuh"

namespace orbx {

__device__ __forceinline__ int ham256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
         __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void __launch_bounds__(128) k_bow_descend(const VocabDev v, const uint8_t* __restrict__ desc, int cap,
                                                     const int32_t* __restrict__ n_per_frame, int n_frames, int levelsup,
                                                     uint32_t* __restrict__ word_id, double* __restrict__ weight,
                                                     uint32_t* __restrict__ node_id) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
  if (i >= cap || f >= n_frames) return;
  const int n = n_per_frame ? n_per_frame[f] : cap;
  if (i >= n) return;
  const size_t o = (size_t)f * cap + i;
  const uint4* q = reinterpret_cast<const uint4*>(desc + 32 * o);
  const uint4 q0 = __ldg(q), q1 = __ldg(q + 1);
  const int nid_level = v.L - levelsup;
  int beg = v.root_beg, cnt = v.root_cnt, level = 0, best_slot = -1;
  uint32_t nid = 0;  // the root when nid_level <= 0 (:1151)
  bool nid_set = nid_level <= 0;
  while (cnt > 0) {  // do { } while (!isLeaf()) of :1156-1174; an empty vocabulary never gets here with cnt > 0
    ++level;
    int best_d = 1 << 30;
    // the children's rows are fetched five at a time, side by side, before any of them is compared (a loop of load - compare
    // keeps one row in flight per thread, and the three deepest levels miss L1)
    constexpr int kG = 5;
    for (int c0 = 0; c0 < cnt; c0 += kG) {
      uint4 r0[kG], r1[kG];
#pragma unroll
      for (int u = 0; u < kG; u++) {
        const uint4* r = v.sdesc + 2 * (size_t)(beg + min(c0 + u, cnt - 1));
        r0[u] = __ldg(r);
        r1[u] = __ldg(r + 1);
      }
#pragma unroll
      for (int u = 0; u < kG; u++) {
        const int d = ham256(q0, q1, r0[u], r1[u]);
        if (c0 + u < cnt && d < best_d) { best_d = d; best_slot = beg + c0 + u; }  // strict <: the first minimum wins (:1166)
      }
    }
    if (level == nid_level) { nid = (uint32_t)__ldg(v.snode + best_slot); nid_set = true; }
    const int2 ch = __ldg(v.schild + best_slot);
    beg = ch.x;
    cnt = ch.y;
  }
  if (best_slot < 0) {  // empty vocabulary
    word_id[o] = 0; weight[o] = 0.0; node_id[o] = 0;
    return;
  }
  if (!nid_set) nid = (uint32_t)__ldg(v.snode + best_slot);  // leaf above nid_level: unset in the reference
  word_id[o] = (uint32_t)__ldg(v.sword + best_slot);
  weight[o] = __ldg(v.sweight + best_slot);
  node_id[o] = nid;
}

constexpr int kBowThreads = 256;
constexpr unsigned long long kBowPad = ~0ull;

// bitonic sort of P (power of two) 64-bit keys in shared memory, ascending
__device__ void bow_sort(unsigned long long* a, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += kBowThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
        const unsigned long long x = a[i], y = a[l];
        const bool up = (i & k) == 0;
        if ((x > y) == up) { a[i] = y; a[l] = x; }
      }
      __syncthreads();
    }
  }
}

// exclusive prefix over the block of one int per thread; returns the prefix, *total = block sum
__device__ int bow_block_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int s = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, s, o);
    if (lane >= o) s += t;
  }
  if (lane == 31) warp_sums[wid] = s;
  __syncthreads();
  int before = 0, all = 0;
  for (int w = 0; w < kBowThreads / 32; w++) {
    const int x = warp_sums[w];
    if (w < wid) before += x;
    all += x;
  }
  __syncthreads();
  *total = all;
  return before + s - v;
}

__global__ void __launch_bounds__(kBowThreads) k_bow_frame(int P, int cap, const int32_t* __restrict__ n_per_frame, int tf_weighting,
                                                           int must_normalize, int l2_norm, const uint32_t* __restrict__ word_id,
                                                           const double* __restrict__ weight, const uint32_t* __restrict__ node_id,
                                                           uint32_t* __restrict__ bow_ids, double* __restrict__ bow_vals,
                                                           int32_t* __restrict__ bow_n, uint32_t* __restrict__ fv_nodes,
                                                           int32_t* __restrict__ fv_begin, int32_t* __restrict__ fv_n,
                                                           uint32_t* __restrict__ fv_feats, int32_t* __restrict__ fv_total) {
  extern __shared__ unsigned long long bow_keys[];  // P keys: (word or node id) << 32 | feature index
  __shared__ int warp_sums[kBowThreads / 32];
  __shared__ int m_valid;
  __shared__ double norm_sh;
  const int f = blockIdx.x, tid = threadIdx.x;
  int n = n_per_frame ? n_per_frame[f] : cap;
  n = n < 0 ? 0 : (n > cap ? cap : n);
  const size_t fo = (size_t)f * cap;
  const uint32_t* wid = word_id + fo;
  const double* w = weight + fo;
  const uint32_t* nid = node_id + fo;
  if (tid == 0) m_valid = 0;
  __syncthreads();

  // ---- BowVector: keys of the features that are not stopped (weight > 0, :1084)
  int mine = 0;
  for (int i = tid; i < P; i += kBowThreads) {
    unsigned long long key = kBowPad;
    if (i < n && w[i] > 0) { key = ((unsigned long long)wid[i] << 32) | (unsigned)i; mine++; }
    bow_keys[i] = key;
  }
  if (mine) atomicAdd(&m_valid, mine);
  __syncthreads();
  const int m = m_valid;
  bow_sort(bow_keys, P);
  // heads of equal-word runs -> dense output slots (chunk of consecutive keys per thread)
  const int per = (P + kBowThreads - 1) / kBowThreads, c_beg = tid * per, c_end = min(m, c_beg + per);
  int heads = 0;
  for (int p = c_beg; p < c_end; p++) heads += p == 0 || (bow_keys[p] >> 32) != (bow_keys[p - 1] >> 32);
  int nb;
  int out = bow_block_scan(heads, warp_sums, &nb);
  for (int p = c_beg; p < c_end; p++) {
    const uint32_t word = (uint32_t)(bow_keys[p] >> 32);
    if (p == 0 || word != (uint32_t)(bow_keys[p - 1] >> 32)) {
      double val = w[(uint32_t)bow_keys[p]];  // first insertion (:36); later ones add in feature order (:33)
      if (tf_weighting)
        for (int r = p + 1; r < m && (uint32_t)(bow_keys[r] >> 32) == word; r++) val = __dadd_rn(val, w[(uint32_t)bow_keys[r]]);
      bow_ids[fo + out] = word;
      bow_vals[fo + out] = val;
      out++;
    }
  }
  __syncthreads();  // this CTA's bow_vals are visible to all its threads
  if (tf_weighting && !must_normalize && nb > 0) {  // :1091-1096
    const double nd = (double)nb;
    for (int i = tid; i < nb; i += kBowThreads) bow_vals[fo + i] = __ddiv_rn(bow_vals[fo + i], nd);
    __syncthreads();
  }
  if (must_normalize) {  // BowVector::normalize (BowVector.cpp:52-66): the sum runs in increasing word id
    if (tid < 32) {
      double norm = 0.0;
      for (int b = 0; b < nb; b += 32) {
        const double x = b + tid < nb ? bow_vals[fo + b + tid] : 0.0;
        const double t = l2_norm ? __dmul_rn(x, x) : fabs(x);
        const int cnt = min(32, nb - b);
        for (int l = 0; l < cnt; l++) norm = __dadd_rn(norm, __shfl_sync(0xffffffffu, t, l));
      }
      if (l2_norm) norm = sqrt(norm);
      if (tid == 0) norm_sh = norm;
    }
    __syncthreads();
    const double norm = norm_sh;
    if (norm > 0.0)
      for (int i = tid; i < nb; i += kBowThreads) bow_vals[fo + i] = __ddiv_rn(bow_vals[fo + i], norm);
  }
  if (tid == 0) { bow_n[f] = nb; fv_total[f] = m; }
  __syncthreads();

  // ---- FeatureVector: the same features keyed by the node at level L - levelsup
  for (int i = tid; i < P; i += kBowThreads) {
    unsigned long long key = kBowPad;
    if (i < n && w[i] > 0) key = ((unsigned long long)nid[i] << 32) | (unsigned)i;
    bow_keys[i] = key;
  }
  __syncthreads();
  bow_sort(bow_keys, P);
  heads = 0;
  for (int p = c_beg; p < c_end; p++) heads += p == 0 || (bow_keys[p] >> 32) != (bow_keys[p - 1] >> 32);
  int nf;
  out = bow_block_scan(heads, warp_sums, &nf);
  for (int p = c_beg; p < c_end; p++) {
    const uint32_t node = (uint32_t)(bow_keys[p] >> 32);
    if (p == 0 || node != (uint32_t)(bow_keys[p - 1] >> 32)) {
      fv_nodes[fo + out] = node;
      fv_begin[fo + out] = p;
      out++;
    }
    fv_feats[fo + p] = (uint32_t)bow_keys[p];
  }
  if (tid == 0) fv_n[f] = nf;
}

int bow_max_features() { return 16384; }  // 128 KB of keys per CTA

cudaError_t bow_configure() {
  return cudaFuncSetAttribute(k_bow_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, bow_max_features() * 8);
}

int launch_bow_descend(const VocabDev& v, const uint8_t* desc, int cap, const int32_t* n_per_frame, int n_frames, int levelsup,
                       uint32_t* word_id, double* weight, uint32_t* node_id, cudaStream_t st) {
  if (cap <= 0 || n_frames <= 0) return 0;
  dim3 grid((cap + 127) / 128, n_frames);
  k_bow_descend<<<grid, 128, 0, st>>>(v, desc, cap, n_per_frame, n_frames, levelsup, word_id, weight, node_id);
  return 1;
}

int launch_bow_frame(int cap, const int32_t* n_per_frame, int n_frames, int tf_weighting, int must_normalize, int l2_norm,
                     const uint32_t* word_id, const double* weight, const uint32_t* node_id, uint32_t* bow_ids, double* bow_vals,
                     int32_t* bow_n, uint32_t* fv_nodes, int32_t* fv_begin, int32_t* fv_n, uint32_t* fv_feats, int32_t* fv_total,
                     cudaStream_t st) {
  if (n_frames <= 0) return 0;
  int P = 2;
  while (P < cap) P <<= 1;
  k_bow_frame<<<n_frames, kBowThreads, (size_t)P * 8, st>>>(P, cap, n_per_frame, tf_weighting, must_normalize, l2_norm, word_id,
                                                          weight, node_id, bow_ids, bow_vals, bow_n, fv_nodes, fv_begin, fv_n,
                                                          fv_feats, fv_total);
  return 1;
}

}  // namespace orbx
