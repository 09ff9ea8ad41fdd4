// orbv_api.cu -- C ABI of the vocabulary half of liborbx_b200.so (include/orbx.h, orbv_*).
// Host side: the text loader (TemplatedVocabulary::loadFromTextFile, TemplatedVocabulary.h:1246-1330),
// the slot layout of the tree for the device, staging of host buffers; kernels in bow.cu.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "orbx_kernels.cuh"

using namespace orbx;

struct orbv_vocab {
  int device = 0;
  int k = 0, L = 0, scoring = 0, weighting = 0;
  int32_t n_nodes = 0, n_words = 0;
  cudaStream_t stream = nullptr;
  VocabDev dev{};
  std::vector<void*> allocs;
  uint8_t* arena = nullptr;  // staging of host-memory calls + per-feature scratch
  size_t arena_bytes = 0, arena_used = 0;
  // the arena is shared by every call: a call on another stream than the previous one waits for the event the
  // previous call recorded behind its last use of it
  cudaEvent_t busy_ev = nullptr;
  cudaStream_t busy_stream = nullptr;
  bool busy = false;
  long long launches = 0;
  char err[256] = "";
};

namespace {

int fail(orbv_t* v, int code, const char* fmt, ...) {
  if (v) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(v->err, sizeof(v->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CU(v, call)                                                                              \
  do {                                                                                           \
    const cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) return fail(v, ORBX_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)
#define TRY(x)           \
  do {                   \
    const int rc_ = (x); \
    if (rc_) return rc_; \
  } while (0)

size_t pad256(size_t b) { return (b + 255) / 256 * 256; }

int arena_reserve(orbv_t* v, size_t bytes) {
  v->arena_used = 0;
  if (bytes <= v->arena_bytes) return ORBX_OK;
  if (v->busy) CU(v, cudaEventSynchronize(v->busy_ev));
  if (v->arena) cudaFree(v->arena);
  v->arena = nullptr;
  v->arena_bytes = 0;
  CU(v, cudaMalloc((void**)&v->arena, bytes));
  v->arena_bytes = bytes;
  return ORBX_OK;
}

template <class T>
T* arena_take(orbv_t* v, size_t count) {
  T* p = reinterpret_cast<T*>(v->arena + v->arena_used);
  v->arena_used += pad256(count * sizeof(T));
  return p;
}

template <class T>
int upload(orbv_t* v, const std::vector<T>& h, const T** out) {
  void* d = nullptr;
  CU(v, cudaMalloc(&d, (h.size() ? h.size() : 1) * sizeof(T)));
  v->allocs.push_back(d);
  if (!h.empty()) CU(v, cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = (const T*)d;
  return ORBX_OK;
}

// Slot layout: breadth-first from the root, children in increasing node id (the push_back order of
// loadFromTextFile, :1300), so that the children of a node are consecutive slots.
int build_device_tree(orbv_t* v, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc, const double* weight) {
  const int n = v->n_nodes;
  std::vector<int32_t> cbeg((size_t)n + 1, 0), child((size_t)(n > 0 ? n : 1)), fill((size_t)n, 0), word((size_t)n, 0);
  int words = 0;
  for (int i = 1; i < n; i++) {
    cbeg[parent[i] + 1]++;
    if (is_leaf[i]) word[i] = words++;  // :1319-1324
  }
  v->n_words = words;
  for (int i = 0; i < n; i++) cbeg[i + 1] += cbeg[i];
  for (int i = 1; i < n; i++) child[cbeg[parent[i]] + fill[parent[i]]++] = i;
  // breadth-first over the part of the tree that hangs from the root
  std::vector<int32_t> order;  // order[slot] = node id
  order.reserve((size_t)n);
  std::vector<int32_t> slot_children_beg, slot_children_cnt;
  size_t head = 0;
  const int root_cnt = cbeg[1] - cbeg[0];
  for (int c = cbeg[0]; c < cbeg[1]; c++) order.push_back(child[c]);
  while (head < order.size()) {
    const int node = order[head++];
    const int b = cbeg[node], e = cbeg[node + 1];
    slot_children_beg.push_back((int32_t)order.size());
    slot_children_cnt.push_back(e - b);
    for (int c = b; c < e; c++) order.push_back(child[c]);
    if (order.size() > (size_t)n) return fail(v, ORBX_E_ARG, "the parent array does not describe a tree");
  }
  const size_t ns = order.size();
  std::vector<uint4> sdesc(2 * ns);
  std::vector<int2> schild(ns);
  std::vector<int32_t> snode(ns), sword(ns);
  std::vector<double> sweight(ns);
  for (size_t s = 0; s < ns; s++) {
    const int node = order[s];
    memcpy(&sdesc[2 * s], desc + 32 * (size_t)node, 32);
    schild[s] = make_int2(slot_children_beg[s], slot_children_cnt[s]);
    snode[s] = node;
    sword[s] = word[node];
    sweight[s] = weight[node];
  }
  TRY(upload(v, sdesc, &v->dev.sdesc));
  TRY(upload(v, schild, &v->dev.schild));
  TRY(upload(v, snode, &v->dev.snode));
  TRY(upload(v, sword, &v->dev.sword));
  TRY(upload(v, sweight, &v->dev.sweight));
  v->dev.root_beg = 0;
  v->dev.root_cnt = root_cnt;
  v->dev.L = v->L;
  return ORBX_OK;
}

int begin(orbv_t* v, int mem, void* stream, cudaStream_t* st) {
  if (!v) return ORBX_E_ARG;
  if (mem != ORBX_MEM_HOST && mem != ORBX_MEM_DEVICE) return fail(v, ORBX_E_ARG, "bad mem kind");
  CU(v, cudaSetDevice(v->device));
  *st = (mem == ORBX_MEM_DEVICE && stream) ? (cudaStream_t)stream : v->stream;
  if (!v->busy_ev) CU(v, cudaEventCreateWithFlags(&v->busy_ev, cudaEventDisableTiming));
  if (v->busy && v->busy_stream != *st) CU(v, cudaStreamWaitEvent(*st, v->busy_ev, 0));
  return ORBX_OK;
}

int end(orbv_t* v, int mem, cudaStream_t st) {
  CU(v, cudaEventRecord(v->busy_ev, st));
  v->busy_stream = st;
  v->busy = true;
  if (mem == ORBX_MEM_HOST) CU(v, cudaStreamSynchronize(st));
  CU(v, cudaGetLastError());
  return ORBX_OK;
}

template <class T>
int stage_in(orbv_t* v, int mem, const T* src, size_t count, const T** out, cudaStream_t st) {
  if (mem == ORBX_MEM_DEVICE || !src) { *out = src; return ORBX_OK; }
  T* d = arena_take<T>(v, count);
  if (count) CU(v, cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
  *out = d;
  return ORBX_OK;
}
template <class T>
T* stage_out(orbv_t* v, int mem, T* dst, size_t count) {
  return mem == ORBX_MEM_DEVICE ? dst : arena_take<T>(v, count);
}
template <class T>
int finish_out(orbv_t* v, int mem, T* dst, const T* dev, size_t count, cudaStream_t st) {
  if (mem == ORBX_MEM_DEVICE || !count) return ORBX_OK;
  CU(v, cudaMemcpyAsync(dst, dev, count * sizeof(T), cudaMemcpyDeviceToHost, st));
  return ORBX_OK;
}

}  // namespace

extern "C" {

int orbv_create(int device, int k, int L, int scoring, int weighting, int32_t n_nodes, const int32_t* parent,
                const uint8_t* is_leaf, const uint8_t* desc, const double* weight, orbv_t** out) {
  if (!out) return ORBX_E_ARG;
  *out = nullptr;
  if (n_nodes < 1 || (n_nodes > 1 && (!parent || !is_leaf || !desc || !weight)) || scoring < 0 || scoring > 5 || weighting < 0 ||
      weighting > 3)
    return ORBX_E_ARG;
  for (int i = 1; i < n_nodes; i++)
    if (parent[i] < 0 || parent[i] >= n_nodes || parent[i] == i) return ORBX_E_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return ORBX_E_CUDA;  // no CPU fallback
  orbv_t* v = new (std::nothrow) orbv_vocab();
  if (!v) return ORBX_E_NOMEM;
  v->device = device;
  v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->n_nodes = n_nodes;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking) != cudaSuccess ||
      bow_configure() != cudaSuccess) {
    delete v;
    return ORBX_E_CUDA;
  }
  const int rc = build_device_tree(v, parent, is_leaf, desc, weight);
  if (rc != ORBX_OK) {
    orbv_destroy(v);
    return rc;
  }
  *out = v;
  return ORBX_OK;
}

int orbv_load_text(int device, const char* path, orbv_t** out) {
  if (!out || !path) return ORBX_E_ARG;
  *out = nullptr;
  FILE* f = fopen(path, "r");
  if (!f) return ORBX_E_ARG;
  int k, L, n1, n2;
  if (fscanf(f, "%d %d %d %d", &k, &L, &n1, &n2) != 4 || k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 ||
      n2 > 3) {  // :1267-1272
    fclose(f);
    return ORBX_E_ARG;
  }
  std::vector<int32_t> parent(1, 0);
  std::vector<uint8_t> leaf(1, 0), desc(32, 0);
  std::vector<double> weight(1, 0.0);
  // one node per line: parent, leaf flag, 32 descriptor bytes as decimals (FORB::fromString, FORB.cpp:105-116), weight
  for (;;) {
    int pid, il, e[32];
    double w;
    if (fscanf(f, "%d %d", &pid, &il) != 2) break;
    bool ok = true;
    for (int i = 0; i < 32 && ok; i++) ok = fscanf(f, "%d", &e[i]) == 1;
    if (!ok || fscanf(f, "%lf", &w) != 1) break;
    parent.push_back(pid);
    leaf.push_back(il > 0);  // :1318
    for (int i = 0; i < 32; i++) desc.push_back((uint8_t)e[i]);
    weight.push_back(w);
  }
  fclose(f);
  return orbv_create(device, k, L, n1, n2, (int32_t)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data(), out);
}

void orbv_destroy(orbv_t* v) {
  if (!v) return;
  cudaSetDevice(v->device);
  if (v->stream) cudaStreamSynchronize(v->stream);
  if (v->busy) cudaEventSynchronize(v->busy_ev);
  if (v->busy_ev) cudaEventDestroy(v->busy_ev);
  for (void* p : v->allocs) cudaFree(p);
  if (v->arena) cudaFree(v->arena);
  if (v->stream) cudaStreamDestroy(v->stream);
  delete v;
}

const char* orbv_last_error(const orbv_t* v) { return v ? v->err : "null handle"; }

int orbv_info(const orbv_t* v, int* k, int* L, int* scoring, int* weighting, int32_t* n_nodes, int32_t* n_words) {
  if (!v) return ORBX_E_ARG;
  if (k) *k = v->k;
  if (L) *L = v->L;
  if (scoring) *scoring = v->scoring;
  if (weighting) *weighting = v->weighting;
  if (n_nodes) *n_nodes = v->n_nodes;
  if (n_words) *n_words = v->n_words;
  return ORBX_OK;
}

int orbv_sync(orbv_t* v) {
  if (!v) return ORBX_E_ARG;
  CU(v, cudaSetDevice(v->device));
  CU(v, cudaStreamSynchronize(v->stream));
  return ORBX_OK;
}

long long orbv_launch_count(const orbv_t* v) { return v ? v->launches : 0; }

int orbv_max_features(void) { return bow_max_features(); }

int orbv_features(orbv_t* v, const uint8_t* desc, int n, int levelsup, uint32_t* word_id, double* weight, uint32_t* node_id,
                  int mem, void* stream) {
  cudaStream_t st;
  TRY(begin(v, mem, stream, &st));
  if (n < 0 || (n > 0 && (!desc || !word_id || !weight || !node_id))) return fail(v, ORBX_E_ARG, "null buffer");
  if (n == 0) return ORBX_OK;
  if (mem == ORBX_MEM_HOST) TRY(arena_reserve(v, pad256((size_t)n * 32) + pad256((size_t)n * 4) * 2 + pad256((size_t)n * 8)));
  const uint8_t* d_desc;
  TRY(stage_in(v, mem, desc, (size_t)n * 32, &d_desc, st));
  uint32_t* d_w = stage_out(v, mem, word_id, (size_t)n);
  double* d_wt = stage_out(v, mem, weight, (size_t)n);
  uint32_t* d_n = stage_out(v, mem, node_id, (size_t)n);
  v->launches += launch_bow_descend(v->dev, d_desc, n, nullptr, 1, levelsup, d_w, d_wt, d_n, st);
  TRY(finish_out(v, mem, word_id, d_w, (size_t)n, st));
  TRY(finish_out(v, mem, weight, d_wt, (size_t)n, st));
  TRY(finish_out(v, mem, node_id, d_n, (size_t)n, st));
  return end(v, mem, st);
}

int orbv_transform(orbv_t* v, const uint8_t* desc, int cap, const int32_t* n_per_frame, int n_frames, int levelsup,
                   uint32_t* bow_ids, double* bow_vals, int32_t* bow_n, uint32_t* fv_nodes, int32_t* fv_begin, int32_t* fv_n,
                   uint32_t* fv_feats, int32_t* fv_total, int mem, void* stream) {
  cudaStream_t st;
  TRY(begin(v, mem, stream, &st));
  if (cap < 0 || n_frames < 0) return fail(v, ORBX_E_ARG, "negative size");
  if (cap > bow_max_features()) return fail(v, ORBX_E_UNSUPPORTED, "more than %d features per frame", bow_max_features());
  if (n_frames == 0) return ORBX_OK;
  if (!bow_n || !fv_n || !fv_total || (cap > 0 && (!desc || !bow_ids || !bow_vals || !fv_nodes || !fv_begin || !fv_feats)))
    return fail(v, ORBX_E_ARG, "null buffer");
  const size_t tot = (size_t)n_frames * (size_t)cap, nf = (size_t)n_frames;
  // scratch: word / weight / node per feature; host calls also stage every input and output
  size_t need = pad256(tot * 4) * 2 + pad256(tot * 8);
  if (mem == ORBX_MEM_HOST) need += pad256(tot * 32) + pad256(nf * 4) * 4 + pad256(tot * 4) * 4 + pad256(tot * 8);
  if (need > v->arena_bytes || mem == ORBX_MEM_HOST) TRY(arena_reserve(v, need));
  else v->arena_used = 0;
  uint32_t* s_word = arena_take<uint32_t>(v, tot);
  double* s_weight = arena_take<double>(v, tot);
  uint32_t* s_node = arena_take<uint32_t>(v, tot);
  const uint8_t* d_desc;
  const int32_t* d_npf;
  TRY(stage_in(v, mem, desc, tot * 32, &d_desc, st));
  TRY(stage_in(v, mem, n_per_frame, nf, &d_npf, st));
  uint32_t* d_bi = stage_out(v, mem, bow_ids, tot);
  double* d_bv = stage_out(v, mem, bow_vals, tot);
  int32_t* d_bn = stage_out(v, mem, bow_n, nf);
  uint32_t* d_fn = stage_out(v, mem, fv_nodes, tot);
  int32_t* d_fb = stage_out(v, mem, fv_begin, tot);
  int32_t* d_fc = stage_out(v, mem, fv_n, nf);
  uint32_t* d_ff = stage_out(v, mem, fv_feats, tot);
  int32_t* d_ft = stage_out(v, mem, fv_total, nf);
  if (v->n_words == 0 || cap == 0) {  // empty(): transform() clears both vectors and returns (:1063)
    CU(v, cudaMemsetAsync(d_bn, 0, nf * 4, st));
    CU(v, cudaMemsetAsync(d_fc, 0, nf * 4, st));
    CU(v, cudaMemsetAsync(d_ft, 0, nf * 4, st));
  } else {
    v->launches += launch_bow_descend(v->dev, d_desc, cap, d_npf, n_frames, levelsup, s_word, s_weight, s_node, st);
    const int tf = v->weighting == 0 || v->weighting == 1;  // TF_IDF, TF: addWeight; IDF, BINARY: addIfNotExist
    const int must = v->scoring != 5, l2 = v->scoring == 1;  // DBoW2/ScoringObject.h:76-91
    v->launches += launch_bow_frame(cap, d_npf, n_frames, tf, must, l2, s_word, s_weight, s_node, d_bi, d_bv, d_bn, d_fn, d_fb, d_fc,
                                    d_ff, d_ft, st);
  }
  TRY(finish_out(v, mem, bow_ids, d_bi, tot, st));
  TRY(finish_out(v, mem, bow_vals, d_bv, tot, st));
  TRY(finish_out(v, mem, bow_n, d_bn, nf, st));
  TRY(finish_out(v, mem, fv_nodes, d_fn, tot, st));
  TRY(finish_out(v, mem, fv_begin, d_fb, tot, st));
  TRY(finish_out(v, mem, fv_n, d_fc, nf, st));
  TRY(finish_out(v, mem, fv_feats, d_ff, tot, st));
  TRY(finish_out(v, mem, fv_total, d_ft, nf, st));
  return end(v, mem, st);
}

}  // extern "C"
