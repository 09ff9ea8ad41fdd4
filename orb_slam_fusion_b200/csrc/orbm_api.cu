// orbm_api.cu -- C ABI of the matcher half of liborbx_b200.so (include/orbx.h, orbm_*).
// Host buffers are staged through a per-handle device arena; device buffers are used in place.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "nccl_dl.cuh"
#include "orbx_kernels.cuh"

using namespace orbx;

struct orbm_matcher {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint8_t* arena = nullptr;  // staging for host-memory calls
  size_t arena_bytes = 0, arena_used = 0;
  void* partials = nullptr;  // knn2 per-chunk top-2
  size_t partial_bytes = 0;
  unsigned long long* keys = nullptr;  // sharded search: this rank's packed top-2, then the gathered keys of every rank
  size_t key_bytes = 0;
  // Handle-owned scratch (arena, partials, keys) is shared by every call: a call that runs on another stream than the
  // previous one first waits for the event the previous call recorded behind its last use of the scratch.
  cudaEvent_t busy_ev = nullptr;
  cudaStream_t busy_stream = nullptr;
  bool busy = false;
  int claim_sequential = 0;  // ORBM_OPT_CLAIM_SEQUENTIAL
  long long launches = 0;
  char err[256] = "";
};

namespace {

int fail(orbm_t* m, int code, const char* fmt, ...) {
  if (m) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(m->err, sizeof(m->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CU(m, call)                                                                              \
  do {                                                                                           \
    const cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) return fail(m, ORBX_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

int arena_reserve(orbm_t* m, size_t bytes) {
  m->arena_used = 0;
  if (bytes <= m->arena_bytes) return ORBX_OK;
  if (m->busy) CU(m, cudaEventSynchronize(m->busy_ev));  // the previous call, whatever stream it ran on, is done with the arena
  if (m->arena) cudaFree(m->arena);
  m->arena = nullptr;
  m->arena_bytes = 0;
  CU(m, cudaMalloc((void**)&m->arena, bytes));
  m->arena_bytes = bytes;
  return ORBX_OK;
}

size_t pad256(size_t b) { return (b + 255) / 256 * 256; }

template <class T>
T* arena_take(orbm_t* m, size_t count) {
  T* p = reinterpret_cast<T*>(m->arena + m->arena_used);
  m->arena_used += pad256(count * sizeof(T));
  return p;
}

// device copy of a host input (or the pointer itself for device memory)
template <class T>
int stage_in(orbm_t* m, int mem, const T* src, size_t count, const T** out, cudaStream_t st) {
  if (mem == ORBX_MEM_DEVICE || !src) { *out = src; return ORBX_OK; }
  T* d = arena_take<T>(m, count);
  if (count) CU(m, cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
  *out = d;
  return ORBX_OK;
}

template <class T>
T* stage_out(orbm_t* m, int mem, T* dst, size_t count) {
  return mem == ORBX_MEM_DEVICE ? dst : arena_take<T>(m, count);
}

template <class T>
int finish_out(orbm_t* m, int mem, T* dst, const T* dev, size_t count, cudaStream_t st) {
  if (mem == ORBX_MEM_DEVICE || !count) return ORBX_OK;
  CU(m, cudaMemcpyAsync(dst, dev, count * sizeof(T), cudaMemcpyDeviceToHost, st));
  return ORBX_OK;
}

int begin(orbm_t* m, int mem, void* stream, cudaStream_t* st) {
  if (!m) return ORBX_E_ARG;
  if (mem != ORBX_MEM_HOST && mem != ORBX_MEM_DEVICE) return fail(m, ORBX_E_ARG, "bad mem kind");
  CU(m, cudaSetDevice(m->device));
  *st = (mem == ORBX_MEM_DEVICE && stream) ? (cudaStream_t)stream : m->stream;
  if (m->busy && m->busy_stream != *st) CU(m, cudaStreamWaitEvent(*st, m->busy_ev, 0));
  return ORBX_OK;
}

int end(orbm_t* m, int mem, cudaStream_t st) {
  CU(m, cudaEventRecord(m->busy_ev, st));
  m->busy_stream = st;
  m->busy = true;
  if (mem == ORBX_MEM_HOST) CU(m, cudaStreamSynchronize(st));
  CU(m, cudaGetLastError());
  return ORBX_OK;
}

#define TRY(x)                \
  do {                        \
    const int rc_ = (x);      \
    if (rc_) return rc_;      \
  } while (0)

// grow a handle-owned device buffer; the previous call (on whatever stream) must be done with the old one first
int buffer_reserve(orbm_t* m, void** buf, size_t* have, size_t want) {
  if (want <= *have) return ORBX_OK;
  if (m->busy) CU(m, cudaEventSynchronize(m->busy_ev));
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *have = 0;
  CU(m, cudaMalloc(buf, want));
  *have = want;
  return ORBX_OK;
}
int partials_reserve(orbm_t* m, size_t bytes) { return buffer_reserve(m, &m->partials, &m->partial_bytes, bytes); }

}  // namespace

extern "C" {

int orbm_create(int device, orbm_t** out) {
  if (!out) return ORBX_E_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return ORBX_E_CUDA;
  orbm_t* m = new (std::nothrow) orbm_matcher();
  if (!m) return ORBX_E_NOMEM;
  m->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->busy_ev, cudaEventDisableTiming) != cudaSuccess || projection_configure() != cudaSuccess) {
    if (m->busy_ev) cudaEventDestroy(m->busy_ev);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return ORBX_E_CUDA;
  }
  *out = m;
  return ORBX_OK;
}

void orbm_destroy(orbm_t* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaStreamSynchronize(m->stream);
  if (m->busy) cudaEventSynchronize(m->busy_ev);
  if (m->arena) cudaFree(m->arena);
  if (m->partials) cudaFree(m->partials);
  if (m->keys) cudaFree(m->keys);
  cudaEventDestroy(m->busy_ev);
  cudaStreamDestroy(m->stream);
  delete m;
}

const char* orbm_last_error(const orbm_t* m) { return m ? m->err : "null handle"; }

int orbm_sync(orbm_t* m) {
  if (!m) return ORBX_E_ARG;
  CU(m, cudaSetDevice(m->device));
  CU(m, cudaStreamSynchronize(m->stream));
  return ORBX_OK;
}

long long orbm_launch_count(const orbm_t* m) { return m ? m->launches : 0; }

int orbm_set_option(orbm_t* m, int option, int value) {
  if (!m) return ORBX_E_ARG;
  if (option == ORBM_OPT_CLAIM_SEQUENTIAL) { m->claim_sequential = value != 0; return ORBX_OK; }
  return fail(m, ORBX_E_ARG, "unknown option %d", option);
}

int orbm_hamming_pairs(orbm_t* m, const uint8_t* a, const uint8_t* b, int64_t n, int32_t* out, int mem, void* stream) {
  cudaStream_t st;
  TRY(begin(m, mem, stream, &st));
  if (n < 0 || (n > 0 && (!a || !b || !out))) return fail(m, ORBX_E_ARG, "bad argument");
  if (n == 0) return ORBX_OK;
  if (mem == ORBX_MEM_HOST) TRY(arena_reserve(m, 2 * pad256((size_t)n * 32) + pad256((size_t)n * 4)));
  const uint8_t *da, *db;
  TRY(stage_in(m, mem, a, (size_t)n * 32, &da, st));
  TRY(stage_in(m, mem, b, (size_t)n * 32, &db, st));
  int32_t* dout = stage_out(m, mem, out, (size_t)n);
  m->launches += launch_hamming_pairs(da, db, n, dout, st);
  TRY(finish_out(m, mem, out, dout, (size_t)n, st));
  return end(m, mem, st);
}

int orbm_knn2(orbm_t* m, const uint8_t* q, int nq, const uint8_t* db, int64_t nd, int64_t db_index_base, int64_t* idx,
              int32_t* dist, int mem, void* stream) {
  cudaStream_t st;
  TRY(begin(m, mem, stream, &st));
  if (nq < 0 || nd < 0 || (nq > 0 && (!q || !idx || !dist)) || (nd > 0 && !db)) return fail(m, ORBX_E_ARG, "bad argument");
  if (nq == 0) return ORBX_OK;
  if (((uintptr_t)db & 15) != 0 && mem == ORBX_MEM_DEVICE) return fail(m, ORBX_E_ARG, "device database must be 16-byte aligned");
  TRY(partials_reserve(m, knn2_partial_bytes(nq, nd)));
  if (mem == ORBX_MEM_HOST)
    TRY(arena_reserve(m, pad256((size_t)nq * 32) + pad256((size_t)nd * 32) + pad256((size_t)nq * 16) + pad256((size_t)nq * 8)));
  const uint8_t *dq, *ddb;
  TRY(stage_in(m, mem, q, (size_t)nq * 32, &dq, st));
  TRY(stage_in(m, mem, db, (size_t)nd * 32, &ddb, st));
  int64_t* didx = stage_out(m, mem, idx, (size_t)nq * 2);
  int32_t* ddist = stage_out(m, mem, dist, (size_t)nq * 2);
  m->launches += launch_knn2(dq, nq, ddb, nd, db_index_base, m->partials, didx, ddist, st);
  TRY(finish_out(m, mem, idx, didx, (size_t)nq * 2, st));
  TRY(finish_out(m, mem, dist, ddist, (size_t)nq * 2, st));
  return end(m, mem, st);
}

int orbm_top2_merge(orbm_t* m, const int64_t* idx_parts, const int32_t* dist_parts, int n_parts, int nq, int64_t* idx,
                    int32_t* dist, int mem, void* stream) {
  cudaStream_t st;
  TRY(begin(m, mem, stream, &st));
  if (nq < 0 || n_parts < 0 || (nq > 0 && (!idx || !dist)) || (nq > 0 && n_parts > 0 && (!idx_parts || !dist_parts)))
    return fail(m, ORBX_E_ARG, "bad argument");
  if (nq == 0) return ORBX_OK;
  const size_t np = (size_t)n_parts * nq * 2;
  if (mem == ORBX_MEM_HOST) TRY(arena_reserve(m, pad256(np * 8) + pad256(np * 4) + pad256((size_t)nq * 16) + pad256((size_t)nq * 8)));
  const int64_t* di;
  const int32_t* dd;
  TRY(stage_in(m, mem, idx_parts, np, &di, st));
  TRY(stage_in(m, mem, dist_parts, np, &dd, st));
  int64_t* didx = stage_out(m, mem, idx, (size_t)nq * 2);
  int32_t* ddist = stage_out(m, mem, dist, (size_t)nq * 2);
  m->launches += launch_top2_merge(di, dd, n_parts, nq, didx, ddist, st);
  TRY(finish_out(m, mem, idx, didx, (size_t)nq * 2, st));
  TRY(finish_out(m, mem, dist, ddist, (size_t)nq * 2, st));
  return end(m, mem, st);
}

int orbm_ratio_test(orbm_t* m, const int64_t* idx, const int32_t* dist, int nq, double ratio, uint8_t* accept, int mem,
                    void* stream) {
  cudaStream_t st;
  TRY(begin(m, mem, stream, &st));
  if (nq < 0 || (nq > 0 && (!idx || !dist || !accept))) return fail(m, ORBX_E_ARG, "bad argument");
  if (nq == 0) return ORBX_OK;
  if (mem == ORBX_MEM_HOST) TRY(arena_reserve(m, pad256((size_t)nq * 16) + pad256((size_t)nq * 8) + pad256((size_t)nq)));
  const int64_t* di;
  const int32_t* dd;
  TRY(stage_in(m, mem, idx, (size_t)nq * 2, &di, st));
  TRY(stage_in(m, mem, dist, (size_t)nq * 2, &dd, st));
  uint8_t* dacc = stage_out(m, mem, accept, (size_t)nq);
  m->launches += launch_ratio_test(di, dd, nq, ratio, dacc, st);
  TRY(finish_out(m, mem, accept, dacc, (size_t)nq, st));
  return end(m, mem, st);
}

/* ---- database sharded over the GPUs of one node ---- */

int orbm_nccl_unique_id(uint8_t id[ORBM_NCCL_ID_BYTES]) {
  const NcclApi* nc = nccl_api(nullptr);
  if (!nc || !id) return nc ? ORBX_E_ARG : ORBX_E_UNSUPPORTED;
  static_assert(sizeof(NcclUniqueId) == ORBM_NCCL_ID_BYTES, "ncclUniqueId is 128 bytes");
  NcclUniqueId u;
  if (nc->GetUniqueId(&u) != 0) return ORBX_E_CUDA;
  memcpy(id, &u, sizeof(u));
  return ORBX_OK;
}

int orbm_nccl_comm_create(const uint8_t id[ORBM_NCCL_ID_BYTES], int n_ranks, int rank, int device, ncclComm_t* comm) {
  const NcclApi* nc = nccl_api(nullptr);
  if (!nc) return ORBX_E_UNSUPPORTED;
  if (!id || !comm || n_ranks < 1 || rank < 0 || rank >= n_ranks) return ORBX_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return ORBX_E_CUDA;
  NcclUniqueId u;
  memcpy(&u, id, sizeof(u));
  void* c = nullptr;
  if (nc->CommInitRank(&c, n_ranks, u, rank) != 0) return ORBX_E_CUDA;
  *comm = (ncclComm_t)c;
  return ORBX_OK;
}

int orbm_nccl_comm_destroy(ncclComm_t comm) {
  const NcclApi* nc = nccl_api(nullptr);
  if (!nc) return ORBX_E_UNSUPPORTED;
  return comm && nc->CommDestroy((void*)comm) != 0 ? ORBX_E_CUDA : ORBX_OK;
}

int orbm_nccl_version(void) {
  const NcclApi* nc = nccl_api(nullptr);
  int v = 0;
  return nc && nc->GetVersion(&v) == 0 ? v : 0;
}

int orbm_knn2_sharded(orbm_t* m, ncclComm_t comm, const uint8_t* q, int nq, const uint8_t* db_local, int64_t nd_local,
                      int64_t db_index_base, double ratio, int64_t* idx, int32_t* dist, uint8_t* accept, int mem, void* stream) {
  cudaStream_t st;
  TRY(begin(m, mem, stream, &st));
  if (nq < 0 || nd_local < 0 || db_index_base < 0 || db_index_base + nd_local >= ((int64_t)1 << 40) ||
      (nq > 0 && (!q || !idx || !dist)) || (nd_local > 0 && !db_local))
    return fail(m, ORBX_E_ARG, "bad argument");
  if (((uintptr_t)db_local & 15) != 0 && mem == ORBX_MEM_DEVICE) return fail(m, ORBX_E_ARG, "device database must be 16-byte aligned");
  int n_ranks = 1;
  const NcclApi* nc = nullptr;
  if (comm) {
    const char* why = "";
    nc = nccl_api(&why);
    if (!nc) return fail(m, ORBX_E_UNSUPPORTED, "%s", why);
    const int rc = nc->CommCount((void*)comm, &n_ranks);
    if (rc != 0 || n_ranks < 1) return fail(m, ORBX_E_ARG, "ncclCommCount: %s", nc->GetErrorString(rc));
  }
  if (nq == 0) return ORBX_OK;  // every rank passes the same nq, so no rank waits in the collective
  TRY(partials_reserve(m, knn2_partial_bytes(nq, nd_local)));
  const size_t key_count = (size_t)nq * 2;
  TRY(buffer_reserve(m, (void**)&m->keys, &m->key_bytes, (size_t)(n_ranks + 1) * key_count * sizeof(unsigned long long)));
  if (mem == ORBX_MEM_HOST)
    TRY(arena_reserve(m, pad256((size_t)nq * 32) + pad256((size_t)nd_local * 32) + pad256((size_t)nq * 16) + pad256((size_t)nq * 8) +
                             pad256((size_t)nq)));
  const uint8_t *dq, *ddb;
  TRY(stage_in(m, mem, q, (size_t)nq * 32, &dq, st));
  TRY(stage_in(m, mem, db_local, (size_t)nd_local * 32, &ddb, st));
  int64_t* didx = stage_out(m, mem, idx, (size_t)nq * 2);
  int32_t* ddist = stage_out(m, mem, dist, (size_t)nq * 2);
  uint8_t* dacc = accept ? stage_out(m, mem, accept, (size_t)nq) : nullptr;
  unsigned long long* mine = m->keys;                // [nq][2]
  unsigned long long* all = m->keys + key_count;     // [n_ranks][nq][2]
  m->launches += launch_knn2_keys(dq, nq, ddb, nd_local, db_index_base, m->partials, n_ranks > 1 ? mine : all, st);
  if (n_ranks > 1) {
    // the one exchange step of the path: 16 bytes per query and rank, on the stream the kernels run on
    const int rc = nc->AllGather(mine, all, key_count, kNcclUint64, (void*)comm, st);
    if (rc != 0) return fail(m, ORBX_E_CUDA, "ncclAllGather: %s", nc->GetErrorString(rc));
  }
  m->launches += launch_top2_keys_ratio(all, n_ranks, nq, ratio, didx, ddist, dacc, st);
  TRY(finish_out(m, mem, idx, didx, (size_t)nq * 2, st));
  TRY(finish_out(m, mem, dist, ddist, (size_t)nq * 2, st));
  if (accept) TRY(finish_out(m, mem, accept, dacc, (size_t)nq, st));
  return end(m, mem, st);
}

int orbm_stereo_rowband(orbm_t* m, const orbx_kp* kl, const uint8_t* dl, int nl, const orbx_kp* kr, const uint8_t* dr,
                        int nr, const float* scale_factors, int n_levels, int n_rows, float min_d, float max_d,
                        int32_t* best_idx, int32_t* best_dist, int mem, void* stream) {
  cudaStream_t st;
  TRY(begin(m, mem, stream, &st));
  if (nl < 0 || nr < 0 || n_levels < 1 || !scale_factors || (nl > 0 && (!kl || !dl || !best_idx || !best_dist)) ||
      (nr > 0 && (!kr || !dr)))
    return fail(m, ORBX_E_ARG, "bad argument");
  if (nl == 0) return ORBX_OK;
  if (mem == ORBX_MEM_HOST)
    TRY(arena_reserve(m, pad256((size_t)nl * 28) + pad256((size_t)nl * 32) + pad256((size_t)nr * 28) + pad256((size_t)nr * 32) +
                             pad256((size_t)n_levels * 4) + 2 * pad256((size_t)nl * 4)));
  const orbx_kp *dkl, *dkr;
  const uint8_t *ddl, *ddr;
  const float* dsf;
  TRY(stage_in(m, mem, kl, (size_t)nl, &dkl, st));
  TRY(stage_in(m, mem, dl, (size_t)nl * 32, &ddl, st));
  TRY(stage_in(m, mem, kr, (size_t)nr, &dkr, st));
  TRY(stage_in(m, mem, dr, (size_t)nr * 32, &ddr, st));
  TRY(stage_in(m, mem, scale_factors, (size_t)n_levels, &dsf, st));
  int32_t* dbi = stage_out(m, mem, best_idx, (size_t)nl);
  int32_t* dbd = stage_out(m, mem, best_dist, (size_t)nl);
  m->launches += launch_stereo_rowband(dkl, ddl, nl, dkr, ddr, nr, dsf, n_levels, n_rows, min_d, max_d, dbi, dbd, st);
  TRY(finish_out(m, mem, best_idx, dbi, (size_t)nl, st));
  TRY(finish_out(m, mem, best_dist, dbd, (size_t)nl, st));
  return end(m, mem, st);
}

int orbm_stereo_refine(orbm_t* m, const orbx_t* left, const orbx_t* right, const orbx_kp* kl, int nl, const orbx_kp* kr,
                       int nr, const int32_t* best_idx, const int32_t* best_dist, int th_orb_dist, float min_d, float max_d,
                       float bf, float* u_right, float* depth, int32_t* sad, int mem, void* stream) {
  cudaStream_t st;
  TRY(begin(m, mem, stream, &st));
  if (nl < 0 || nr < 0 || (nl > 0 && (!kl || !best_idx || !best_dist || !u_right || !depth || !sad)) || (nr > 0 && !kr))
    return fail(m, ORBX_E_ARG, "bad argument");
  FrameGeom gl, gr;
  const uint8_t *pl, *pr;
  const float *sf, *isf, *sf_r, *isf_r;
  int dl, dr;
  if (!orbx_peek_pyramid(left, &gl, &pl, &sf, &isf, &dl) || !orbx_peek_pyramid(right, &gr, &pr, &sf_r, &isf_r, &dr))
    return fail(m, ORBX_E_ARG, "both extractors must hold the pyramid of their last frame");
  if (dl != m->device || dr != m->device || gl.nlev != gr.nlev) return fail(m, ORBX_E_ARG, "extractors on another device / level count");
  if (nl == 0) return ORBX_OK;
  if (mem == ORBX_MEM_HOST)
    TRY(arena_reserve(m, pad256((size_t)nl * 28) + pad256((size_t)nr * 28) + 2 * pad256((size_t)nl * 4) + 3 * pad256((size_t)nl * 4)));
  const orbx_kp *dkl, *dkr;
  const int32_t *dbi, *dbd;
  TRY(stage_in(m, mem, kl, (size_t)nl, &dkl, st));
  TRY(stage_in(m, mem, kr, (size_t)nr, &dkr, st));
  TRY(stage_in(m, mem, best_idx, (size_t)nl, &dbi, st));
  TRY(stage_in(m, mem, best_dist, (size_t)nl, &dbd, st));
  float* dur = stage_out(m, mem, u_right, (size_t)nl);
  float* ddp = stage_out(m, mem, depth, (size_t)nl);
  int32_t* dsad = stage_out(m, mem, sad, (size_t)nl);
  // the pyramids belong to the extractors: read them after the extractors' last kernels, and let the extractors' next
  // frame wait for this read
  CU(m, orbx_pyramid_acquire(left, st));
  CU(m, orbx_pyramid_acquire(right, st));
  m->launches += launch_stereo_refine(gl, pl, gr, pr, sf, isf, dkl, nl, dkr, nr, dbi, dbd, th_orb_dist, min_d, max_d, bf, dur, ddp,
                                      dsad, st);
  CU(m, orbx_pyramid_release(left, st));
  if (right != left) CU(m, orbx_pyramid_release(right, st));
  TRY(finish_out(m, mem, u_right, dur, (size_t)nl, st));
  TRY(finish_out(m, mem, depth, ddp, (size_t)nl, st));
  TRY(finish_out(m, mem, sad, dsad, (size_t)nl, st));
  return end(m, mem, st);
}

int orbm_stereo_matches_last(orbm_t* m, const orbx_t* left, const orbx_t* right, float bf, float mb, float* u_right, float* depth,
                             int cap, int* nl_out) {
  cudaStream_t st;
  TRY(begin(m, ORBX_MEM_HOST, nullptr, &st));
  if (!left || !right || !u_right || !depth || !nl_out || !(mb > 0.f)) return fail(m, ORBX_E_ARG, "bad argument");
  FrameGeom gl, gr;
  const uint8_t *pl, *pr, *ddl, *ddr;
  const orbx_kp *dkl, *dkr;
  const float *sf, *isf, *sf_r, *isf_r;
  int dl, dr, nl, nr;
  if (!orbx_peek_pyramid(left, &gl, &pl, &sf, &isf, &dl) || !orbx_peek_pyramid(right, &gr, &pr, &sf_r, &isf_r, &dr) ||
      !orbx_peek_single(left, &dkl, &ddl, &nl) || !orbx_peek_single(right, &dkr, &ddr, &nr))
    return fail(m, ORBX_E_ARG, "both extractors must hold the result of a single-frame call");
  if (dl != m->device || dr != m->device || gl.nlev != gr.nlev) return fail(m, ORBX_E_ARG, "extractors on another device / level count");
  *nl_out = nl;
  if (nl > cap) return fail(m, ORBX_E_CAP, "%d left keypoints, capacity %d", nl, cap);
  if (nl == 0) return ORBX_OK;
  const float min_d = 0.f, max_d = bf / mb;  // frame.cc:853-856
  TRY(arena_reserve(m, pad256((size_t)gl.nlev * 4) + 5 * pad256((size_t)nl * 4)));
  float* dsf = arena_take<float>(m, (size_t)gl.nlev);
  int32_t* dbi = arena_take<int32_t>(m, (size_t)nl);
  int32_t* dbd = arena_take<int32_t>(m, (size_t)nl);
  float* dur = arena_take<float>(m, (size_t)nl);   // dur and ddp are adjacent: one copy brings both back
  float* ddp = arena_take<float>(m, (size_t)nl);
  int32_t* dsad = arena_take<int32_t>(m, (size_t)nl);
  CU(m, cudaMemcpyAsync(dsf, sf, (size_t)gl.nlev * 4, cudaMemcpyHostToDevice, st));
  CU(m, orbx_pyramid_acquire(left, st));   // keypoints, descriptors and pyramids were written on the extractors' streams
  CU(m, orbx_pyramid_acquire(right, st));
  m->launches += launch_stereo_rowband(dkl, ddl, nl, dkr, ddr, nr, dsf, gl.nlev, gl.lv[0].h, min_d, max_d, dbi, dbd, st);
  m->launches += launch_stereo_refine(gl, pl, gr, pr, sf, isf, dkl, nl, dkr, nr, dbi, dbd, 75 /* (TH_HIGH + TH_LOW) / 2, frame.cc:832 */,
                                      min_d, max_d, bf, dur, ddp, dsad, st);
  CU(m, orbx_pyramid_release(left, st));
  if (right != left) CU(m, orbx_pyramid_release(right, st));
  CU(m, cudaMemcpyAsync(u_right, dur, (size_t)nl * 4, cudaMemcpyDeviceToHost, st));
  CU(m, cudaMemcpyAsync(depth, ddp, (size_t)nl * 4, cudaMemcpyDeviceToHost, st));
  return end(m, ORBX_MEM_HOST, st);
}

int orbm_distinctive(orbm_t* m, const uint8_t* desc, const int32_t* offsets, int n_points, int max_rows, int32_t* best_idx,
                     int32_t* best_median, int mem, void* stream) {
  cudaStream_t st;
  TRY(begin(m, mem, stream, &st));
  if (n_points < 0 || max_rows < 0 || max_rows > 6000 || (n_points > 0 && (!offsets || !best_idx || !best_median)))
    return fail(m, ORBX_E_ARG, "bad argument");
  if (n_points == 0) return ORBX_OK;
  size_t total = 0;
  if (mem == ORBX_MEM_HOST) {
    total = (size_t)offsets[n_points];
    for (int p = 0; p < n_points; p++)
      if (offsets[p + 1] - offsets[p] > max_rows) return fail(m, ORBX_E_ARG, "a point owns more rows than max_rows");
    TRY(arena_reserve(m, pad256(total * 32) + pad256((size_t)(n_points + 1) * 4) + 2 * pad256((size_t)n_points * 4)));
  }
  const uint8_t* dd;
  const int32_t* doff;
  TRY(stage_in(m, mem, desc, total * 32, &dd, st));
  TRY(stage_in(m, mem, offsets, (size_t)n_points + 1, &doff, st));
  int32_t* dbi = stage_out(m, mem, best_idx, (size_t)n_points);
  int32_t* dbm = stage_out(m, mem, best_median, (size_t)n_points);
  m->launches += launch_distinctive(dd, doff, n_points, max_rows, dbi, dbm, st);
  TRY(finish_out(m, mem, best_idx, dbi, (size_t)n_points, st));
  TRY(finish_out(m, mem, best_median, dbm, (size_t)n_points, st));
  return end(m, mem, st);
}

static int window_search(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int n, const orbm_grid_geom* geom,
                         const orbm_window_query* queries, const uint8_t* qdesc, int nq, const uint8_t* skip,
                         const float* kp_u_right, const float* q_u_right, const float* q_max_err, const float* inv_sigma2,
                         int n_levels, orbm_window_result* out, int mem, void* stream) {
  cudaStream_t st;
  TRY(begin(m, mem, stream, &st));
  if (n < 0 || nq < 0 || !geom || geom->cols < 1 || geom->rows < 1 || (int64_t)geom->cols * geom->rows >= (1 << 20) ||
      n >= (1 << 24) || (n > 0 && (!kps || !desc)) || (nq > 0 && (!queries || !qdesc || !out)) ||
      (kp_u_right && nq > 0 && (!q_u_right || (!q_max_err && !inv_sigma2))) ||
      (inv_sigma2 && (n_levels < 1 || n_levels > ORBX_MAX_LEVELS)))
    return fail(m, ORBX_E_ARG, "bad argument");
  if (nq == 0) return ORBX_OK;
  if (mem == ORBX_MEM_HOST)
    TRY(arena_reserve(m, pad256((size_t)n * 28) + pad256((size_t)n * 32) + pad256((size_t)nq * sizeof(orbm_window_query)) +
                             pad256((size_t)nq * 32) + pad256((size_t)n) + pad256((size_t)n * 4) + 2 * pad256((size_t)nq * 4) +
                             pad256((size_t)nq * sizeof(orbm_window_result)) + 256));
  const orbx_kp* dk;
  const uint8_t *dd, *dqd, *dskip;
  const orbm_window_query* dq;
  const float *dur = nullptr, *dqr = nullptr, *dqe = nullptr;
  TRY(stage_in(m, mem, kps, (size_t)n, &dk, st));
  TRY(stage_in(m, mem, desc, (size_t)n * 32, &dd, st));
  TRY(stage_in(m, mem, queries, (size_t)nq, &dq, st));
  TRY(stage_in(m, mem, qdesc, (size_t)nq * 32, &dqd, st));
  TRY(stage_in(m, mem, skip, (size_t)n, &dskip, st));
  if (kp_u_right) {
    TRY(stage_in(m, mem, kp_u_right, (size_t)n, &dur, st));
    TRY(stage_in(m, mem, q_u_right, (size_t)nq, &dqr, st));
    if (q_max_err) TRY(stage_in(m, mem, q_max_err, (size_t)nq, &dqe, st));
  }
  const float* dis = nullptr;
  if (inv_sigma2) TRY(stage_in(m, mem, inv_sigma2, (size_t)n_levels, &dis, st));
  orbm_window_result* dout = stage_out(m, mem, out, (size_t)nq);
  m->launches += launch_window_search(dk, dd, n, *geom, dq, dqd, nq, dskip, dur, dqr, dqe, dout, st, dis, n_levels);
  TRY(finish_out(m, mem, out, dout, (size_t)nq, st));
  return end(m, mem, st);
}

int orbm_window_search_stereo(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int n, const orbm_grid_geom* geom,
                              const orbm_window_query* queries, const uint8_t* qdesc, int nq, const uint8_t* skip,
                              const float* kp_u_right, const float* q_u_right, const float* q_max_err,
                              orbm_window_result* out, int mem, void* stream) {
  return window_search(m, kps, desc, n, geom, queries, qdesc, nq, skip, kp_u_right, q_u_right, q_max_err, nullptr, 0, out, mem, stream);
}

int orbm_window_search_fuse(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int n, const orbm_grid_geom* geom,
                            const orbm_window_query* queries, const uint8_t* qdesc, int nq, const float* kp_u_right,
                            const float* q_u_right, const float* inv_level_sigma2, int n_levels, orbm_window_result* out,
                            int mem, void* stream) {
  if (!inv_level_sigma2) return m ? fail(m, ORBX_E_ARG, "inv_level_sigma2 is required") : ORBX_E_ARG;
  return window_search(m, kps, desc, n, geom, queries, qdesc, nq, nullptr, kp_u_right, q_u_right, nullptr, inv_level_sigma2, n_levels,
                       out, mem, stream);
}

int orbm_window_search(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int n, const orbm_grid_geom* geom,
                       const orbm_window_query* queries, const uint8_t* qdesc, int nq, const uint8_t* skip,
                       orbm_window_result* out, int mem, void* stream) {
  return orbm_window_search_stereo(m, kps, desc, n, geom, queries, qdesc, nq, skip, nullptr, nullptr, nullptr, out, mem, stream);
}

static int search_by_projection(bool last_frame, const float* q_angle, int check_orientation, orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int n, const orbm_grid_geom* geom,
                              const orbm_window_query* queries, const uint8_t* qdesc, int nq, const uint8_t* skip,
                              const float* kp_u_right, const float* q_u_right, const float* q_max_err, int th_high,
                              float nnratio, int32_t* assigned, int32_t* n_matches, int mem, void* stream) {
  cudaStream_t st;
  TRY(begin(m, mem, stream, &st));
  if (n < 0 || nq < 0 || !geom || geom->cols < 1 || geom->rows < 1 || (int64_t)geom->cols * geom->rows >= (1 << 20) ||
      n >= (1 << 24) || (n > 0 && (!kps || !desc || !assigned)) || (nq > 0 && (!queries || !qdesc)) || !n_matches ||
      (kp_u_right && nq > 0 && (!q_u_right || !q_max_err)))
    return fail(m, ORBX_E_ARG, "bad argument");
  if ((size_t)((n + 31) / 32 + 1) * 4 > 200 * 1024) return fail(m, ORBX_E_UNSUPPORTED, "more than 1.6 M keypoints in one frame");
  if (last_frame && (size_t)n * 4 > 180 * 1024) return fail(m, ORBX_E_UNSUPPORTED, "more than 46 k keypoints in one frame");
  if (last_frame && check_orientation && nq > 0 && !q_angle) return fail(m, ORBX_E_ARG, "orientation check without angles");
  const size_t scratch = pad256(projection_scratch_bytes(nq));
  if (mem == ORBX_MEM_HOST) {
    TRY(arena_reserve(m, pad256((size_t)n * 28) + pad256((size_t)n * 32) + pad256((size_t)nq * sizeof(orbm_window_query)) +
                             pad256((size_t)nq * 32) + pad256((size_t)n) + pad256((size_t)n * 4) + 2 * pad256((size_t)nq * 4) +
                             pad256((size_t)n * 4) + pad256((size_t)nq * 4) + 256 + scratch));
  } else {
    TRY(arena_reserve(m, scratch));  // device-memory calls still need the per-query scratch
  }
  const orbx_kp* dk;
  const uint8_t *dd, *dqd, *dskip;
  const orbm_window_query* dq;
  const float *dur = nullptr, *dqr = nullptr, *dqe = nullptr;
  TRY(stage_in(m, mem, kps, (size_t)n, &dk, st));
  TRY(stage_in(m, mem, desc, (size_t)n * 32, &dd, st));
  TRY(stage_in(m, mem, queries, (size_t)nq, &dq, st));
  TRY(stage_in(m, mem, qdesc, (size_t)nq * 32, &dqd, st));
  TRY(stage_in(m, mem, skip, (size_t)n, &dskip, st));
  if (kp_u_right) {
    TRY(stage_in(m, mem, kp_u_right, (size_t)n, &dur, st));
    TRY(stage_in(m, mem, q_u_right, (size_t)nq, &dqr, st));
    TRY(stage_in(m, mem, q_max_err, (size_t)nq, &dqe, st));
  }
  const float* dqa = nullptr;
  if (q_angle) TRY(stage_in(m, mem, q_angle, (size_t)nq, &dqa, st));
  int32_t* dassigned = stage_out(m, mem, assigned, (size_t)n);
  int32_t* dnm = stage_out(m, mem, n_matches, (size_t)1);
  void* dscratch = arena_take<uint8_t>(m, projection_scratch_bytes(nq));
  m->launches += launch_search_by_projection(dk, dd, n, *geom, dq, dqd, nq, dskip, dur, dqr, dqe, th_high, nnratio, dqa,
                                             check_orientation, last_frame, m->claim_sequential != 0, dscratch, dassigned, dnm, st);
  TRY(finish_out(m, mem, assigned, dassigned, (size_t)n, st));
  TRY(finish_out(m, mem, n_matches, dnm, (size_t)1, st));
  return end(m, mem, st);
}

int orbm_search_by_projection(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int n, const orbm_grid_geom* geom,
                              const orbm_window_query* queries, const uint8_t* qdesc, int nq, const uint8_t* skip,
                              const float* kp_u_right, const float* q_u_right, const float* q_max_err, int th_high,
                              float nnratio, int32_t* assigned, int32_t* n_matches, int mem, void* stream) {
  return search_by_projection(false, nullptr, 0, m, kps, desc, n, geom, queries, qdesc, nq, skip, kp_u_right, q_u_right, q_max_err,
                              th_high, nnratio, assigned, n_matches, mem, stream);
}

int orbm_search_by_projection_last(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int n, const orbm_grid_geom* geom,
                                   const orbm_window_query* queries, const uint8_t* qdesc, const float* q_angle, int nq,
                                   const uint8_t* skip, const float* kp_u_right, const float* q_u_right, const float* q_max_err,
                                   int th_high, int check_orientation, int32_t* assigned, int32_t* n_matches, int mem,
                                   void* stream) {
  return search_by_projection(true, q_angle, check_orientation, m, kps, desc, n, geom, queries, qdesc, nq, skip, kp_u_right, q_u_right,
                              q_max_err, th_high, 0.f, assigned, n_matches, mem, stream);
}

static int search_by_bow(bool keyframes, orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int cap, int n_frames, const int32_t* n_per_frame,
                       const uint32_t* fv_nodes, const int32_t* fv_begin, const int32_t* fv_n, const uint32_t* fv_feats,
                       const int32_t* fv_total, const uint8_t* has_point, const int32_t* pair_kf, const int32_t* pair_f,
                       int n_pairs, float nnratio, int check_orientation, int32_t* match, int32_t* n_matches, int mem,
                       void* stream) {
  cudaStream_t st;
  TRY(begin(m, mem, stream, &st));
  if (cap < 1 || cap > 2048 || n_frames < 1 || n_pairs < 0 || !kps || !desc || !fv_nodes || !fv_begin || !fv_n || !fv_feats ||
      !fv_total || (n_pairs > 0 && (!pair_kf || !pair_f || !match || !n_matches)))
    return fail(m, ORBX_E_ARG, "bad argument");
  if (n_pairs == 0) return ORBX_OK;
  const size_t fc = (size_t)n_frames * cap;
  if (mem == ORBX_MEM_HOST) {
    for (int p = 0; p < n_pairs; p++)
      if (pair_kf[p] < 0 || pair_kf[p] >= n_frames || pair_f[p] < 0 || pair_f[p] >= n_frames)
        return fail(m, ORBX_E_ARG, "pair %d refers to a frame outside the pool", p);
    TRY(arena_reserve(m, pad256(fc * sizeof(orbx_kp)) + pad256(fc * 32) + 3 * pad256(fc * 4) + pad256(fc) + 3 * pad256((size_t)n_frames * 4) +
                             3 * pad256((size_t)n_pairs * 4) + pad256((size_t)n_pairs * cap * 4)));
  }
  const orbx_kp* dk;
  const uint8_t *dd, *dhp;
  const uint32_t *dnodes, *dfeats;
  const int32_t *dbegin, *dfn, *dft, *dnpf, *dpk, *dpf;
  TRY(stage_in(m, mem, kps, fc, &dk, st));
  TRY(stage_in(m, mem, desc, fc * 32, &dd, st));
  TRY(stage_in(m, mem, fv_nodes, fc, &dnodes, st));
  TRY(stage_in(m, mem, fv_begin, fc, &dbegin, st));
  TRY(stage_in(m, mem, fv_feats, fc, &dfeats, st));
  TRY(stage_in(m, mem, has_point, fc, &dhp, st));
  TRY(stage_in(m, mem, fv_n, (size_t)n_frames, &dfn, st));
  TRY(stage_in(m, mem, fv_total, (size_t)n_frames, &dft, st));
  TRY(stage_in(m, mem, n_per_frame, (size_t)n_frames, &dnpf, st));
  TRY(stage_in(m, mem, pair_kf, (size_t)n_pairs, &dpk, st));
  TRY(stage_in(m, mem, pair_f, (size_t)n_pairs, &dpf, st));
  int32_t* dmatch = stage_out(m, mem, match, (size_t)n_pairs * cap);
  int32_t* dnm = stage_out(m, mem, n_matches, (size_t)n_pairs);
  m->launches += launch_search_by_bow(dk, dd, cap, n_frames, dnodes, dbegin, dfn, dfeats, dft, dnpf, dhp, dpk, dpf, n_pairs, nnratio,
                                      check_orientation, keyframes, dmatch, dnm, st);
  TRY(finish_out(m, mem, match, dmatch, (size_t)n_pairs * cap, st));
  TRY(finish_out(m, mem, n_matches, dnm, (size_t)n_pairs, st));
  return end(m, mem, st);
}

int orbm_search_by_bow(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int cap, int n_frames, const int32_t* n_per_frame,
                       const uint32_t* fv_nodes, const int32_t* fv_begin, const int32_t* fv_n, const uint32_t* fv_feats,
                       const int32_t* fv_total, const uint8_t* has_point, const int32_t* pair_kf, const int32_t* pair_f,
                       int n_pairs, float nnratio, int check_orientation, int32_t* match, int32_t* n_matches, int mem,
                       void* stream) {
  return search_by_bow(false, m, kps, desc, cap, n_frames, n_per_frame, fv_nodes, fv_begin, fv_n, fv_feats, fv_total, has_point,
                       pair_kf, pair_f, n_pairs, nnratio, check_orientation, match, n_matches, mem, stream);
}

int orbm_search_by_bow_kf(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int cap, int n_frames, const int32_t* n_per_frame,
                          const uint32_t* fv_nodes, const int32_t* fv_begin, const int32_t* fv_n, const uint32_t* fv_feats,
                          const int32_t* fv_total, const uint8_t* has_point, const int32_t* pair_1, const int32_t* pair_2,
                          int n_pairs, float nnratio, int check_orientation, int32_t* match, int32_t* n_matches, int mem,
                          void* stream) {
  return search_by_bow(true, m, kps, desc, cap, n_frames, n_per_frame, fv_nodes, fv_begin, fv_n, fv_feats, fv_total, has_point,
                       pair_1, pair_2, n_pairs, nnratio, check_orientation, match, n_matches, mem, stream);
}

int orbm_search_for_triangulation(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int cap, int n_frames,
                                  const int32_t* n_per_frame, const uint32_t* fv_nodes, const int32_t* fv_begin, const int32_t* fv_n,
                                  const uint32_t* fv_feats, const int32_t* fv_total, const uint8_t* has_point, const float* u_right,
                                  const int32_t* pair_1, const int32_t* pair_2, int n_pairs, const float* pair_f12,
                                  const float* pair_ep, const float* scale_factors, const float* level_sigma2, int n_levels,
                                  int only_stereo, int coarse, int check_orientation, int32_t* match, int32_t* n_matches, int mem,
                                  void* stream) {
  cudaStream_t st;
  TRY(begin(m, mem, stream, &st));
  if (cap < 1 || cap > 2048 || n_frames < 1 || n_pairs < 0 || n_levels < 1 || n_levels > ORBX_MAX_LEVELS || !kps || !desc || !fv_nodes ||
      !fv_begin || !fv_n || !fv_feats || !fv_total || !has_point || !u_right || !scale_factors || !level_sigma2 ||
      (n_pairs > 0 && (!pair_1 || !pair_2 || !pair_f12 || !pair_ep || !match || !n_matches)))
    return fail(m, ORBX_E_ARG, "bad argument");
  if (n_pairs == 0) return ORBX_OK;
  const size_t fc = (size_t)n_frames * cap;
  if (mem == ORBX_MEM_HOST) {
    for (int p = 0; p < n_pairs; p++)
      if (pair_1[p] < 0 || pair_1[p] >= n_frames || pair_2[p] < 0 || pair_2[p] >= n_frames)
        return fail(m, ORBX_E_ARG, "pair %d refers to a frame outside the pool", p);
    TRY(arena_reserve(m, pad256(fc * sizeof(orbx_kp)) + pad256(fc * 32) + 4 * pad256(fc * 4) + pad256(fc) + 3 * pad256((size_t)n_frames * 4) +
                             3 * pad256((size_t)n_pairs * 4) + pad256((size_t)n_pairs * 36) + pad256((size_t)n_pairs * 8) + 2 * 256 +
                             pad256((size_t)n_pairs * cap * 4)));
  }
  const orbx_kp* dk;
  const uint8_t *dd, *dhp;
  const uint32_t *dnodes, *dfeats;
  const int32_t *dbegin, *dfn, *dft, *dnpf, *dp1, *dp2;
  const float *dur, *df12, *dep, *dsf, *ds2;
  TRY(stage_in(m, mem, kps, fc, &dk, st));
  TRY(stage_in(m, mem, desc, fc * 32, &dd, st));
  TRY(stage_in(m, mem, fv_nodes, fc, &dnodes, st));
  TRY(stage_in(m, mem, fv_begin, fc, &dbegin, st));
  TRY(stage_in(m, mem, fv_feats, fc, &dfeats, st));
  TRY(stage_in(m, mem, has_point, fc, &dhp, st));
  TRY(stage_in(m, mem, u_right, fc, &dur, st));
  TRY(stage_in(m, mem, fv_n, (size_t)n_frames, &dfn, st));
  TRY(stage_in(m, mem, fv_total, (size_t)n_frames, &dft, st));
  TRY(stage_in(m, mem, n_per_frame, (size_t)n_frames, &dnpf, st));
  TRY(stage_in(m, mem, pair_1, (size_t)n_pairs, &dp1, st));
  TRY(stage_in(m, mem, pair_2, (size_t)n_pairs, &dp2, st));
  TRY(stage_in(m, mem, pair_f12, (size_t)n_pairs * 9, &df12, st));
  TRY(stage_in(m, mem, pair_ep, (size_t)n_pairs * 2, &dep, st));
  TRY(stage_in(m, mem, scale_factors, (size_t)n_levels, &dsf, st));
  TRY(stage_in(m, mem, level_sigma2, (size_t)n_levels, &ds2, st));
  int32_t* dmatch = stage_out(m, mem, match, (size_t)n_pairs * cap);
  int32_t* dnm = stage_out(m, mem, n_matches, (size_t)n_pairs);
  m->launches += launch_search_for_triangulation(dk, dd, cap, n_frames, dnodes, dbegin, dfn, dfeats, dft, dnpf, dhp, dur, dp1, dp2, n_pairs, df12, dep,
                                                 dsf, ds2, n_levels, only_stereo, coarse, check_orientation, dmatch, dnm, st);
  TRY(finish_out(m, mem, match, dmatch, (size_t)n_pairs * cap, st));
  TRY(finish_out(m, mem, n_matches, dnm, (size_t)n_pairs, st));
  return end(m, mem, st);
}

int orbm_synth_descriptors(int device, uint8_t* dst, int64_t first, int64_t n, uint64_t seed, void* stream) {
  if (!dst || n < 0) return ORBX_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return ORBX_E_CUDA;
  launch_synth_descriptors(dst, first, n, seed, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_popc_peak(int device, int mode, double* per_s) {
  if (!per_s || mode < 0 || mode > 2) return ORBX_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return ORBX_E_CUDA;
  return popc_bench(mode, per_s) == 0 ? ORBX_OK : ORBX_E_CUDA;
}

}  // extern "C"
