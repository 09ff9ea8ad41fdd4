/*
 * oracle/orb_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * See orb_oracle.h for scope and pinning.  Compile with -ffp-contract=off.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  Nothing here is used by the CUDA product path.
 */
#include "orb_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "cvprim.h"

#define EDGE 19    /* kEdgeThreshold, orb_extractor.cc:74 */
#define HALF_PATCH 15 /* kHalfPatchSize, :73 */
#define PATCH 31   /* kPatchSize, :72 */

static const int8_t kPattern[1024] = {
#include "../include/orb_pattern31.inc"
};

typedef struct {
  uint8_t* buf; /* bordered buffer */
  uint8_t* px;  /* pixel (0,0) */
  int w, h;
  size_t stride;
  uint8_t* blur; /* w*h continuous, or NULL */
  int blur_valid;
  int* cand; /* xyr triples */
  int ncand;
  orc_kp* sel;
  int nsel;
} orc_level_t;

struct orc_extractor {
  orc_params p;
  double scale_factor_d; /* the reference keeps scale_factor_ as double (orb_extractor.h:91) */
  float scale[ORC_MAX_LEVELS], inv_scale[ORC_MAX_LEVELS], sigma2[ORC_MAX_LEVELS],
      inv_sigma2[ORC_MAX_LEVELS];
  int quota[ORC_MAX_LEVELS];
  int umax[HALF_PATCH + 1];
  int trig;
  orc_level_t lv[ORC_MAX_LEVELS];
};

/* ---- orb_extractor.cc:407-465 ---- */
orc_extractor* orc_create(const orc_params* p) {
  if (!p || p->num_levs < 1 || p->num_levs > ORC_MAX_LEVELS) return NULL;
  orc_extractor* e = (orc_extractor*)calloc(1, sizeof(*e));
  e->p = *p;
  e->scale_factor_d = (double)p->scale_factor;
  const int L = p->num_levs;
  e->scale[0] = 1.0f;
  e->sigma2[0] = 1.0f;
  for (int i = 1; i < L; i++) {
    e->scale[i] = (float)((double)e->scale[i - 1] * e->scale_factor_d); /* :419 */
    e->sigma2[i] = e->scale[i] * e->scale[i];                           /* :420 */
  }
  for (int i = 0; i < L; i++) {
    e->inv_scale[i] = 1.0f / e->scale[i];
    e->inv_sigma2[i] = 1.0f / e->sigma2[i];
  }
  float factor = (float)(1.0 / e->scale_factor_d); /* 1.0f / double -> double, narrowed (:433) */
  /* num_feats_ * (1 - factor) is int*float -> float; divided by float -> float (:434-436) */
  float per_scale = ((float)p->num_feats * (1 - factor)) / (1 - (float)pow((double)factor, (double)L));
  int sum = 0;
  for (int lev = 0; lev < L - 1; lev++) {
    e->quota[lev] = cvp_round_f(per_scale);
    sum += e->quota[lev];
    per_scale *= factor;
  }
  e->quota[L - 1] = p->num_feats - sum > 0 ? p->num_feats - sum : 0;

  /* :452-464 */
  const int vmax = cvp_floor_f(HALF_PATCH * sqrtf(2.f) / 2 + 1);
  const int vmin = cvp_ceil_f(HALF_PATCH * sqrtf(2.f) / 2);
  const double hp2 = HALF_PATCH * HALF_PATCH;
  for (int v = 0; v <= vmax; ++v) e->umax[v] = cvp_round_d(sqrt(hp2 - v * v));
  for (int v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
    while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
    e->umax[v] = v0;
    ++v0;
  }
  e->trig = ORC_TRIG_LIBM;
  return e;
}

static void level_free(orc_level_t* l) {
  free(l->buf); free(l->blur); free(l->cand); free(l->sel);
  memset(l, 0, sizeof(*l));
}

void orc_destroy(orc_extractor* e) {
  if (!e) return;
  for (int i = 0; i < ORC_MAX_LEVELS; i++) level_free(&e->lv[i]);
  free(e);
}

void orc_set_trig(orc_extractor* e, int mode) { e->trig = mode; }

void orc_tables(const orc_extractor* e, float* scale, float* inv_scale, float* sigma2,
                float* inv_sigma2, int* quota, int* umax) {
  const int L = e->p.num_levs;
  if (scale) memcpy(scale, e->scale, sizeof(float) * L);
  if (inv_scale) memcpy(inv_scale, e->inv_scale, sizeof(float) * L);
  if (sigma2) memcpy(sigma2, e->sigma2, sizeof(float) * L);
  if (inv_sigma2) memcpy(inv_sigma2, e->inv_sigma2, sizeof(float) * L);
  if (quota) memcpy(quota, e->quota, sizeof(int) * L);
  if (umax) memcpy(umax, e->umax, sizeof(int) * (HALF_PATCH + 1));
}

/* ---- orb_extractor.cc:1093-1117 ---- */
int orc_compute_pyramid(orc_extractor* e, const uint8_t* img, int w, int h, size_t stride) {
  if (!img || w <= 0 || h <= 0) return -1;
  for (int lev = 0; lev < e->p.num_levs; lev++) {
    orc_level_t* l = &e->lv[lev];
    level_free(l);
    const float s = e->inv_scale[lev];
    l->w = cvp_round_f((float)w * s); /* :1096 */
    l->h = cvp_round_f((float)h * s);
    if (l->w < 1 || l->h < 1) return -3;
    l->stride = (size_t)l->w + 2 * EDGE;
    l->buf = (uint8_t*)malloc(l->stride * ((size_t)l->h + 2 * EDGE));
    l->px = l->buf + EDGE * l->stride + EDGE;
    if (lev != 0) {
      const orc_level_t* p = &e->lv[lev - 1];
      /* resize into a temporary, then border (copyMakeBorder with BORDER_ISOLATED reads only
       * the level itself; :1106-1111) */
      uint8_t* tmp = (uint8_t*)malloc((size_t)l->w * l->h);
      cvp_resize_linear_u8(p->px, p->w, p->h, p->stride, tmp, l->w, l->h, (size_t)l->w);
      cvp_border_reflect101_u8(tmp, l->w, l->h, (size_t)l->w, l->buf, l->stride, EDGE);
      free(tmp);
    } else {
      cvp_border_reflect101_u8(img, w, h, stride, l->buf, l->stride, EDGE); /* :1113 */
    }
  }
  return 0;
}

const uint8_t* orc_level(const orc_extractor* e, int lev, int* w, int* h, size_t* stride) {
  const orc_level_t* l = &e->lv[lev];
  if (w) *w = l->w;
  if (h) *h = l->h;
  if (stride) *stride = l->stride;
  return l->px;
}

static void ensure_blur(orc_level_t* l) {
  if (l->blur_valid) return;
  /* :1054-1055: clone() makes the level continuous, so the reflect border is taken at the
   * true image edge, and the 8U fixed-point GaussianBlur path is used. */
  uint8_t* clone = (uint8_t*)malloc((size_t)l->w * l->h);
  for (int y = 0; y < l->h; y++) memcpy(clone + (size_t)y * l->w, l->px + (size_t)y * l->stride, l->w);
  if (!l->blur) l->blur = (uint8_t*)malloc((size_t)l->w * l->h);
  cvp_gauss7x7_u8(clone, l->w, l->h, (size_t)l->w, l->blur, (size_t)l->w);
  free(clone);
  l->blur_valid = 1;
}

const uint8_t* orc_blurred_level(orc_extractor* e, int lev, size_t* stride) {
  orc_level_t* l = &e->lv[lev];
  ensure_blur(l);
  if (stride) *stride = (size_t)l->w;
  return l->blur;
}

/* ---- grid FAST, orb_extractor.cc:744-825 ---- */
int orc_fast_grid(const uint8_t* lvl, int w, int h, size_t stride, int ini_th, int min_th,
                  int* xyr, int cap) {
  const float W = 35;
  const int min_bx = EDGE - 3, min_by = min_bx;
  const int max_bx = w - EDGE + 3, max_by = h - EDGE + 3;
  const float width = (float)(max_bx - min_bx), height = (float)(max_by - min_by);
  const int ncols = (int)(width / W), nrows = (int)(height / W);
  if (ncols < 1 || nrows < 1) return 0;
  const int wcell = (int)ceilf(width / ncols), hcell = (int)ceilf(height / nrows);
  int n = 0;
  const int tcap = (wcell + 6) * (hcell + 6);
  int* cell = (int*)malloc(sizeof(int) * 3 * (size_t)tcap);
  for (int i = 0; i < nrows; i++) {
    const float ini_y = (float)(min_by + i * hcell);
    float max_y = ini_y + hcell + 6;
    if (ini_y >= max_by - 3) continue;
    if (max_y > max_by) max_y = (float)max_by;
    for (int j = 0; j < ncols; j++) {
      const float ini_x = (float)(min_bx + j * wcell);
      float max_x = ini_x + wcell + 6;
      if (ini_x >= max_bx - 3) continue; /* fork: -3 (upstream -6), :777-778 */
      if (max_x > max_bx) max_x = (float)max_bx;
      const int x0 = (int)ini_x, y0 = (int)ini_y, cw = (int)max_x - x0, ch = (int)max_y - y0;
      const uint8_t* roi = lvl + (size_t)y0 * stride + x0;
      int c = cvp_fast9_nms_u8(roi, cw, ch, stride, ini_th, cell, tcap);
      if (c == 0) c = cvp_fast9_nms_u8(roi, cw, ch, stride, min_th, cell, tcap); /* :799-801 */
      for (int k = 0; k < c; k++) {
        if (n < cap) {
          xyr[3 * n] = cell[3 * k] + j * wcell; /* :819-820 */
          xyr[3 * n + 1] = cell[3 * k + 1] + i * hcell;
          xyr[3 * n + 2] = cell[3 * k + 2];
        }
        n++;
      }
    }
  }
  free(cell);
  return n;
}

/* ---- DistributeOctTree, orb_extractor.cc:476-742 (list-faithful restatement) ---- */
typedef struct {
  int x0, y0, x1, y1; /* UL=(x0,y0) UR=(x1,y0) BL=(x0,y1) BR=(x1,y1) */
  int* pts;           /* indices into the candidate array, original order preserved */
  int npts;
  int no_more;
  int prev, next; /* doubly linked list */
} ot_node;

typedef struct {
  ot_node* nodes;
  int n_nodes, cap_nodes;
  int head, tail, size;
} ot_list;

static int ot_new_node(ot_list* L) {
  if (L->n_nodes == L->cap_nodes) {
    L->cap_nodes = L->cap_nodes ? L->cap_nodes * 2 : 64;
    L->nodes = (ot_node*)realloc(L->nodes, sizeof(ot_node) * (size_t)L->cap_nodes);
  }
  ot_node* nd = &L->nodes[L->n_nodes];
  memset(nd, 0, sizeof(*nd));
  nd->prev = nd->next = -1;
  return L->n_nodes++;
}
static void ot_push_back(ot_list* L, int id) {
  L->nodes[id].prev = L->tail;
  L->nodes[id].next = -1;
  if (L->tail >= 0) L->nodes[L->tail].next = id; else L->head = id;
  L->tail = id;
  L->size++;
}
static void ot_push_front(ot_list* L, int id) {
  L->nodes[id].next = L->head;
  L->nodes[id].prev = -1;
  if (L->head >= 0) L->nodes[L->head].prev = id; else L->tail = id;
  L->head = id;
  L->size++;
}
static int ot_erase(ot_list* L, int id) { /* returns the following node */
  const int p = L->nodes[id].prev, n = L->nodes[id].next;
  if (p >= 0) L->nodes[p].next = n; else L->head = n;
  if (n >= 0) L->nodes[n].prev = p; else L->tail = p;
  L->size--;
  free(L->nodes[id].pts);
  L->nodes[id].pts = NULL;
  return n;
}

/* ExtractorNode::DivideNode, :476-524.  Children are created as list nodes ids c[0..3]
 * (not yet linked). */
static void ot_divide(ot_list* L, int id, const int* xyr, int c[4]) {
  for (int k = 0; k < 4; k++) c[k] = ot_new_node(L);
  ot_node* nd = &L->nodes[id];
  const int half_x = (int)ceilf((float)(nd->x1 - nd->x0) / 2);
  const int half_y = (int)ceilf((float)(nd->y1 - nd->y0) / 2);
  ot_node* n1 = &L->nodes[c[0]];
  ot_node* n2 = &L->nodes[c[1]];
  ot_node* n3 = &L->nodes[c[2]];
  ot_node* n4 = &L->nodes[c[3]];
  n1->x0 = nd->x0; n1->y0 = nd->y0; n1->x1 = nd->x0 + half_x; n1->y1 = nd->y0 + half_y;
  n2->x0 = nd->x0 + half_x; n2->y0 = nd->y0; n2->x1 = nd->x1; n2->y1 = nd->y0 + half_y;
  n3->x0 = nd->x0; n3->y0 = nd->y0 + half_y; n3->x1 = nd->x0 + half_x; n3->y1 = nd->y1;
  n4->x0 = nd->x0 + half_x; n4->y0 = nd->y0 + half_y; n4->x1 = nd->x1; n4->y1 = nd->y1;
  for (int k = 0; k < 4; k++) L->nodes[c[k]].pts = (int*)malloc(sizeof(int) * (size_t)(nd->npts ? nd->npts : 1));
  for (int i = 0; i < nd->npts; i++) {
    const int pi = nd->pts[i];
    const int px = xyr[3 * pi], py = xyr[3 * pi + 1];
    ot_node* t;
    if (px < n1->x1) t = (py < n1->y1) ? n1 : n3;
    else t = (py < n1->y1) ? n2 : n4;
    t->pts[t->npts++] = pi;
  }
  for (int k = 0; k < 4; k++)
    if (L->nodes[c[k]].npts == 1) L->nodes[c[k]].no_more = 1;
}

typedef struct { int count; int node; } ot_rec;

/* CompareNodes :527-540 + stable_sort :671 -> stable insertion sort ascending (count, UL.x) */
static void ot_stable_sort(ot_rec* a, int n, const ot_list* L) {
  for (int i = 1; i < n; i++) {
    ot_rec v = a[i];
    int j = i - 1;
    while (j >= 0) {
      const int less = (v.count < a[j].count) ||
                       (v.count == a[j].count && L->nodes[v.node].x0 < L->nodes[a[j].node].x0);
      if (!less) break;
      a[j + 1] = a[j];
      j--;
    }
    a[j + 1] = v;
  }
}

int orc_octree(const int* xyr, int n, int min_x, int max_x, int min_y, int max_y, int quota,
               int* out_index, int cap) {
  if (n <= 0) return 0;
  ot_list L;
  memset(&L, 0, sizeof(L));
  L.head = L.tail = -1;
  const int n_ini = (int)roundf((float)(max_x - min_x) / (float)(max_y - min_y)); /* :548 */
  if (n_ini < 1) return -1; /* the reference divides by zero here */
  const float h_x = (float)(max_x - min_x) / n_ini;
  int* ini = (int*)malloc(sizeof(int) * (size_t)n_ini);
  for (int i = 0; i < n_ini; i++) {
    const int id = ot_new_node(&L);
    ot_node* nd = &L.nodes[id];
    nd->x0 = (int)(h_x * (float)i);
    nd->x1 = (int)(h_x * (float)(i + 1));
    nd->y0 = 0;
    nd->y1 = max_y - min_y;
    nd->pts = (int*)malloc(sizeof(int) * (size_t)n);
    ot_push_back(&L, id);
    ini[i] = id;
  }
  for (int i = 0; i < n; i++) {
    int r = (int)((float)xyr[3 * i] / h_x); /* :572 */
    if (r >= n_ini) r = n_ini - 1;
    ot_node* nd = &L.nodes[ini[r]];
    nd->pts[nd->npts++] = i;
  }
  free(ini);
  for (int it = L.head; it >= 0;) { /* :577-585 */
    ot_node* nd = &L.nodes[it];
    if (nd->npts == 1) { nd->no_more = 1; it = nd->next; }
    else if (nd->npts == 0) it = ot_erase(&L, it);
    else it = nd->next;
  }

  int finished = 0;
  ot_rec* rec = NULL; int nrec = 0, caprec = 0;
  ot_rec* prev = NULL; int capprev = 0;
#define REC_PUSH(cnt, nodeid) do { if (nrec == caprec) { caprec = caprec ? caprec * 2 : 64; \
    rec = (ot_rec*)realloc(rec, sizeof(ot_rec) * (size_t)caprec); } rec[nrec].count = (cnt); rec[nrec].node = (nodeid); nrec++; } while (0)

  while (!finished) {
    int size_prev = L.size;
    int to_expand = 0;
    nrec = 0;
    for (int it = L.head; it >= 0;) { /* :594-656 */
      if (L.nodes[it].no_more) { it = L.nodes[it].next; continue; }
      int c[4];
      ot_divide(&L, it, xyr, c);
      for (int k = 0; k < 4; k++) {
        if (L.nodes[c[k]].npts > 0) {
          ot_push_front(&L, c[k]);
          if (L.nodes[c[k]].npts > 1) { to_expand++; REC_PUSH(L.nodes[c[k]].npts, c[k]); }
        } else { free(L.nodes[c[k]].pts); L.nodes[c[k]].pts = NULL; }
      }
      it = ot_erase(&L, it);
    }
    if (L.size >= quota || L.size == size_prev) {
      finished = 1;
    } else if (L.size + to_expand * 3 > quota) { /* :662-719 */
      while (!finished) {
        size_prev = L.size;
        if (nrec > capprev) { capprev = nrec; prev = (ot_rec*)realloc(prev, sizeof(ot_rec) * (size_t)capprev); }
        const int nprev = nrec;
        if (nprev) memcpy(prev, rec, sizeof(ot_rec) * (size_t)nprev);
        nrec = 0;
        ot_stable_sort(prev, nprev, &L);
        for (int j = nprev - 1; j >= 0; j--) {
          int c[4];
          ot_divide(&L, prev[j].node, xyr, c);
          for (int k = 0; k < 4; k++) {
            if (L.nodes[c[k]].npts > 0) {
              ot_push_front(&L, c[k]);
              if (L.nodes[c[k]].npts > 1) REC_PUSH(L.nodes[c[k]].npts, c[k]);
            } else { free(L.nodes[c[k]].pts); L.nodes[c[k]].pts = NULL; }
          }
          ot_erase(&L, prev[j].node);
          if (L.size >= quota) break;
        }
        if (L.size >= quota || L.size == size_prev) finished = 1;
      }
    }
  }
#undef REC_PUSH
  /* :723-739 keep the best point per node; first one wins ties (strict >) */
  int nout = 0;
  for (int it = L.head; it >= 0; it = L.nodes[it].next) {
    const ot_node* nd = &L.nodes[it];
    int best = nd->pts[0];
    for (int k = 1; k < nd->npts; k++)
      if (xyr[3 * nd->pts[k] + 2] > xyr[3 * best + 2]) best = nd->pts[k];
    if (nout < cap) out_index[nout] = best;
    nout++;
  }
  for (int i = 0; i < L.n_nodes; i++) free(L.nodes[i].pts);
  free(L.nodes); free(rec); free(prev);
  return nout;
}

/* ---- IC_Angle, orb_extractor.cc:76-100 ---- */
static float ic_angle_umax(const uint8_t* lvl, size_t stride, int cx, int cy, const int* umax) {
  int m_01 = 0, m_10 = 0;
  const uint8_t* center = lvl + (size_t)cy * stride + cx;
  for (int u = -HALF_PATCH; u <= HALF_PATCH; ++u) m_10 += u * center[u];
  const int step = (int)stride;
  for (int v = 1; v <= HALF_PATCH; ++v) {
    int v_sum = 0;
    const int d = umax[v];
    for (int u = -d; u <= d; ++u) {
      const int val_plus = center[u + v * step], val_minus = center[u - v * step];
      v_sum += (val_plus - val_minus);
      m_10 += u * (val_plus + val_minus);
    }
    m_01 += v * v_sum;
  }
  return cvp_fast_atan2((float)m_01, (float)m_10);
}

static const int kUmax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

float orc_ic_angle(const uint8_t* lvl, size_t stride, int cx, int cy) {
  return ic_angle_umax(lvl, stride, cx, cy, kUmax);
}

/* ---- ComputeOrbDescriptor, orb_extractor.cc:102-146 ---- */
void orc_rbrief(const uint8_t* img, size_t stride, int cx, int cy, float angle_deg, int trig_mode,
                uint8_t desc[32]) {
  const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
  const float angle = angle_deg * factorPI;
  float a, b;
  if (trig_mode == ORC_TRIG_CR) { a = (float)cos((double)angle); b = (float)sin((double)angle); }
  else { a = cosf(angle); b = sinf(angle); }
  const uint8_t* center = img + (size_t)cy * stride + cx;
  const int step = (int)stride;
  const int8_t* pat = kPattern;
#define GETV(idx) center[cvp_round_f(pat[2 * (idx)] * b + pat[2 * (idx) + 1] * a) * step + \
                         cvp_round_f(pat[2 * (idx)] * a - pat[2 * (idx) + 1] * b)]
  for (int i = 0; i < 32; ++i, pat += 32) {
    int val = 0;
    for (int k = 0; k < 8; k++) {
      const int t0 = GETV(2 * k), t1 = GETV(2 * k + 1);
      val |= (t0 < t1) << k;
    }
    desc[i] = (uint8_t)val;
  }
#undef GETV
}

/* ---- operator(), orb_extractor.cc:1011-1091 (+ ComputeKeyPointsOctTree :744-849) ---- */
int orc_extract(orc_extractor* e, const uint8_t* img, int w, int h, size_t stride, int lap0,
                int lap1, orc_kp* kps, uint8_t* desc, int cap, int* n_out, int* n_mono) {
  if (!img || w <= 0 || h <= 0) return -1; /* :1016 */
  int rc = orc_compute_pyramid(e, img, w, h, stride);
  if (rc) return rc;
  const int L = e->p.num_levs;
  int total = 0;
  for (int lev = 0; lev < L; lev++) {
    orc_level_t* l = &e->lv[lev];
    const int ccap = l->w * l->h / 4 + 16;
    l->cand = (int*)malloc(sizeof(int) * 3 * (size_t)ccap);
    l->ncand = orc_fast_grid(l->px, l->w, l->h, l->stride, e->p.ini_th_fast, e->p.min_th_fast,
                             l->cand, ccap);
    const int min_b = EDGE - 3;
    const int max_bx = l->w - EDGE + 3, max_by = l->h - EDGE + 3;
    int* idx = (int*)malloc(sizeof(int) * (size_t)(e->quota[lev] + 8 + l->ncand));
    int ns = orc_octree(l->cand, l->ncand, min_b, max_bx, min_b, max_by, e->quota[lev], idx,
                        e->quota[lev] + 8 + l->ncand);
    if (ns < 0) ns = 0;
    l->sel = (orc_kp*)malloc(sizeof(orc_kp) * (size_t)(ns ? ns : 1));
    l->nsel = ns;
    const int scaled_patch = (int)(PATCH * e->scale[lev]); /* :834 */
    for (int i = 0; i < ns; i++) {
      orc_kp* k = &l->sel[i];
      k->x = (float)l->cand[3 * idx[i]] + (float)min_b;
      k->y = (float)l->cand[3 * idx[i] + 1] + (float)min_b;
      k->size = (float)scaled_patch;
      k->response = (float)l->cand[3 * idx[i] + 2];
      k->octave = lev;
      k->class_id = -1;
      k->angle = ic_angle_umax(l->px, l->stride, cvp_round_f(k->x), cvp_round_f(k->y), e->umax);
    }
    free(idx);
    total += ns;
  }
  if (n_out) *n_out = total;
  if (total > cap) return -2;
  int mono = 0, stereo = total - 1;
  for (int lev = 0; lev < L; lev++) {
    orc_level_t* l = &e->lv[lev];
    if (l->nsel == 0) continue;
    ensure_blur(l);
    const float scale = e->scale[lev];
    for (int i = 0; i < l->nsel; i++) {
      orc_kp k = l->sel[i];
      uint8_t d[32];
      orc_rbrief(l->blur, (size_t)l->w, cvp_round_f(k.x), cvp_round_f(k.y), k.angle, e->trig, d);
      if (lev != 0) { k.x *= scale; k.y *= scale; } /* :1071-1073 */
      int slot;
      if (k.x >= (float)lap0 && k.x <= (float)lap1) slot = stereo--; /* :1075-1084 */
      else slot = mono++;
      kps[slot] = k;
      memcpy(desc + 32 * (size_t)slot, d, 32);
    }
  }
  if (n_mono) *n_mono = mono;
  return 0;
}

int orc_candidates(const orc_extractor* e, int lev, int* xyr, int cap) {
  const orc_level_t* l = &e->lv[lev];
  const int n = l->ncand < cap ? l->ncand : cap;
  if (xyr && n > 0) memcpy(xyr, l->cand, sizeof(int) * 3 * (size_t)n);
  return l->ncand;
}

int orc_selected(const orc_extractor* e, int lev, orc_kp* out, int cap) {
  const orc_level_t* l = &e->lv[lev];
  const int n = l->nsel < cap ? l->nsel : cap;
  if (out && n > 0) memcpy(out, l->sel, sizeof(orc_kp) * (size_t)n);
  return l->nsel;
}

/* ---- matching ---- */
int orc_hamming(const uint8_t* a, const uint8_t* b) { /* orb_matcher.cc:1877-1891 */
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t pa, pb;
    memcpy(&pa, a + 4 * i, 4);
    memcpy(&pb, b + 4 * i, 4);
    uint32_t v = pa ^ pb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (int)((((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24);
  }
  return dist;
}

typedef struct {
  const uint8_t* q; int q0, q1; const uint8_t* d; int64_t nd; int64_t* idx; int* dist;
} knn_job;

static void* knn_worker(void* arg) {
  knn_job* j = (knn_job*)arg;
  for (int qi = j->q0; qi < j->q1; qi++) {
    int b0 = INT_MAX, b1 = INT_MAX;
    int64_t i0 = -1, i1 = -1;
    const uint8_t* qd = j->q + 32 * (size_t)qi;
    for (int64_t r = 0; r < j->nd; r++) {
      const int dd = orc_hamming(qd, j->d + 32 * (size_t)r);
      if (dd < b0) { b1 = b0; i1 = i0; b0 = dd; i0 = r; }
      else if (dd < b1) { b1 = dd; i1 = r; }
    }
    j->idx[2 * qi] = i0; j->idx[2 * qi + 1] = i1;
    j->dist[2 * qi] = b0; j->dist[2 * qi + 1] = b1;
  }
  return NULL;
}

void orc_knn2(const uint8_t* q, int nq, const uint8_t* d, int64_t nd, int64_t* idx, int* dist,
              int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nq) nthreads = nq > 0 ? nq : 1;
  knn_job* jobs = (knn_job*)malloc(sizeof(knn_job) * (size_t)nthreads);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (knn_job){q, (int)((int64_t)nq * t / nthreads), (int)((int64_t)nq * (t + 1) / nthreads), d, nd, idx, dist};
    if (nthreads == 1) knn_worker(&jobs[t]);
    else pthread_create(&th[t], NULL, knn_worker, &jobs[t]);
  }
  if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(jobs); free(th);
}

int orc_ratio_accept(int d0, int d1, int have2, double ratio) {
  /* DMatch.distance is float; (*it)[1].distance * 0.7 is a double product (frame.cc:1162) */
  return have2 && ((float)d0 < (float)d1 * ratio);
}

void orc_stereo_rowband(const orc_kp* kl, const uint8_t* dl, int nl, const orc_kp* kr,
                        const uint8_t* dr, int nr, const float* sf, int n_rows, float min_d,
                        float max_d, int* best_idx, int* best_dist) {
  /* frame.cc:836-851 row table */
  int* cnt = (int*)calloc((size_t)n_rows + 1, sizeof(int));
  int** rows = (int**)calloc((size_t)n_rows, sizeof(int*));
  for (int ir = 0; ir < nr; ir++) {
    const float r = 2.0f * sf[kr[ir].octave];
    const int maxr = (int)ceilf(kr[ir].y + r), minr = (int)floorf(kr[ir].y - r);
    for (int yi = minr; yi <= maxr; yi++) if (yi >= 0 && yi < n_rows) cnt[yi]++;
  }
  for (int y = 0; y < n_rows; y++) { rows[y] = (int*)malloc(sizeof(int) * (size_t)(cnt[y] ? cnt[y] : 1)); cnt[y] = 0; }
  for (int ir = 0; ir < nr; ir++) {
    const float r = 2.0f * sf[kr[ir].octave];
    const int maxr = (int)ceilf(kr[ir].y + r), minr = (int)floorf(kr[ir].y - r);
    for (int yi = minr; yi <= maxr; yi++) if (yi >= 0 && yi < n_rows) rows[yi][cnt[yi]++] = ir;
  }
  for (int il = 0; il < nl; il++) {
    best_idx[il] = -1;
    best_dist[il] = 100; /* ORBmatcher::TH_HIGH */
    const int level_l = kl[il].octave;
    const float vl = kl[il].y, ul = kl[il].x;
    const int row = (int)vl; /* vRowIndices[vL], frame.cc:868 */
    if (row < 0 || row >= n_rows || cnt[row] == 0) continue;
    const float min_u = ul - max_d, max_u = ul - min_d;
    if (max_u < 0) continue;
    for (int ic = 0; ic < cnt[row]; ic++) {
      const int ir = rows[row][ic];
      if (kr[ir].octave < level_l - 1 || kr[ir].octave > level_l + 1) continue;
      const float ur = kr[ir].x;
      if (ur >= min_u && ur <= max_u) {
        const int dist = orc_hamming(dl + 32 * (size_t)il, dr + 32 * (size_t)ir);
        if (dist < best_dist[il]) { best_dist[il] = dist; best_idx[il] = ir; }
      }
    }
  }
  for (int y = 0; y < n_rows; y++) free(rows[y]);
  free(rows); free(cnt);
}

/* ---- mappoint.cc:365-428 ---- */
static int int_cmp(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }
void orc_distinctive(const uint8_t* desc, const int* offsets, int n_points, int* best_idx, int* best_median) {
  for (int p = 0; p < n_points; p++) {
    const int o = offsets[p], N = offsets[p + 1] - o;
    best_idx[p] = -1; best_median[p] = INT_MAX;
    if (N <= 0) continue;
    int* row = (int*)malloc(sizeof(int) * (size_t)N);
    int bm = INT_MAX, bi = 0;
    for (int i = 0; i < N; i++) {
      for (int j = 0; j < N; j++) row[j] = i == j ? 0 : orc_hamming(desc + 32 * (size_t)(o + i), desc + 32 * (size_t)(o + j));
      qsort(row, (size_t)N, sizeof(int), int_cmp);
      const int median = row[(int)(0.5 * (N - 1))]; /* :414 */
      if (median < bm) { bm = median; bi = i; }
    }
    free(row);
    best_idx[p] = bi; best_median[p] = bm;
  }
}

/* ---- frame.cc:903-985 ---- */
typedef struct { int dist; int il; } sad_rec;
static int sad_rec_cmp(const void* a, const void* b) { /* std::sort of pair<int,int> */
  const sad_rec* x = (const sad_rec*)a; const sad_rec* y = (const sad_rec*)b;
  if (x->dist != y->dist) return x->dist < y->dist ? -1 : 1;
  return x->il < y->il ? -1 : (x->il > y->il ? 1 : 0);
}

void orc_stereo_refine(const orc_level_view* left, const orc_level_view* right, int n_levels,
                       const orc_kp* kl, int nl, const orc_kp* kr, const int* best_idx,
                       const int* best_dist, const float* sf, const float* isf, int th_orb_dist,
                       float min_d, float max_d, float bf, float* u_right, float* depth, int* sad) {
  sad_rec* recs = (sad_rec*)malloc(sizeof(sad_rec) * (size_t)(nl ? nl : 1));
  int nrec = 0;
  for (int il = 0; il < nl; il++) {
    u_right[il] = -1.0f; depth[il] = -1.0f; sad[il] = -1;
    if (best_idx[il] < 0 || !(best_dist[il] < th_orb_dist)) continue; /* :903 */
    const orc_kp* kpl = &kl[il];
    const int oct = kpl->octave;
    if (oct < 0 || oct >= n_levels) continue;
    const float ur0 = kr[best_idx[il]].x;
    const float scale = isf[oct];
    const float sul = roundf(kpl->x * scale), svl = roundf(kpl->y * scale), sur0 = roundf(ur0 * scale); /* :907-909 */
    const int w = 5, L = 5;
    const orc_level_view* pl = &left[oct];
    const orc_level_view* pr = &right[oct];
    const float iniu = sur0 + L - w, endu = sur0 + L + w + 1; /* :923-927 */
    if (iniu < 0 || endu >= (float)pr->w) continue;
    int best = INT_MAX, best_inc = 0;
    float dists[11];
    const int y0 = (int)(svl - w), xl0 = (int)(sul - w);
    for (int inc = -L; inc <= L; inc++) { /* :929-942 */
      const int xr0 = (int)(sur0 + inc - w);
      int acc = 0;
      for (int dy = 0; dy < 2 * w + 1; dy++)
        for (int dx = 0; dx < 2 * w + 1; dx++) {
          const int a = pl->px[(ptrdiff_t)(y0 + dy) * (ptrdiff_t)pl->stride + xl0 + dx];
          const int b = pr->px[(ptrdiff_t)(y0 + dy) * (ptrdiff_t)pr->stride + xr0 + dx];
          acc += a > b ? a - b : b - a;
        }
      const float dist = (float)acc; /* cv::norm(..., NORM_L1) */
      if (dist < (float)best) { best = (int)dist; best_inc = inc; }
      dists[L + inc] = dist;
    }
    if (best_inc == -L || best_inc == L) continue; /* :944 */
    const float d1 = dists[L + best_inc - 1], d2 = dists[L + best_inc], d3 = dists[L + best_inc + 1];
    const float delta = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2)); /* :951-952 */
    if (delta < -1 || delta > 1) continue;
    float best_ur = sf[oct] * ((float)sur0 + (float)best_inc + delta); /* :957-958 */
    float disparity = kpl->x - best_ur;
    if (disparity >= min_d && disparity < max_d) { /* :962-970 */
      if (disparity <= 0) { disparity = 0.01f; best_ur = (float)((double)kpl->x - 0.01); }
      depth[il] = bf / disparity;
      u_right[il] = best_ur;
      sad[il] = best;
      recs[nrec].dist = best; recs[nrec].il = il; nrec++;
    }
  }
  if (nrec > 0) { /* :974-985 (the reference indexes an empty vector when nothing matched) */
    qsort(recs, (size_t)nrec, sizeof(sad_rec), sad_rec_cmp);
    const float median = (float)recs[nrec / 2].dist;
    const float th = 1.5f * 1.4f * median;
    for (int i = nrec - 1; i >= 0; i--) {
      if ((float)recs[i].dist < th) break;
      u_right[recs[i].il] = -1.0f; depth[recs[i].il] = -1.0f;
    }
  }
  free(recs);
}

void orc_window_search(const orc_kp* kps, const uint8_t* desc, int n, const orc_grid_geom* g,
                       const orc_window_query* q, const uint8_t* qdesc, int nq,
                       const uint8_t* skip, orc_window_result* out) {
  orc_window_search_stereo(kps, desc, n, g, q, qdesc, nq, skip, NULL, NULL, NULL, out);
}

static void window_search_impl(const orc_kp* kps, const uint8_t* desc, int n, const orc_grid_geom* g,
                               const orc_window_query* q, const uint8_t* qdesc, int nq, const uint8_t* skip,
                               const float* kp_u_right, const float* q_u_right, const float* q_max_err,
                               const float* g_fuse_inv_sigma2 /* non-NULL: the chi-square gate of Fuse replaces the stereo gate */,
                               orc_window_result* out);

void orc_window_search_fuse(const orc_kp* kps, const uint8_t* desc, int n, const orc_grid_geom* g, const orc_window_query* q,
                            const uint8_t* qdesc, int nq, const float* kp_u_right, const float* q_u_right,
                            const float* inv_level_sigma2, orc_window_result* out) {
  window_search_impl(kps, desc, n, g, q, qdesc, nq, NULL, kp_u_right, q_u_right, NULL, inv_level_sigma2, out);
}

void orc_window_search_stereo(const orc_kp* kps, const uint8_t* desc, int n, const orc_grid_geom* g,
                              const orc_window_query* q, const uint8_t* qdesc, int nq, const uint8_t* skip,
                              const float* kp_u_right, const float* q_u_right, const float* q_max_err,
                              orc_window_result* out) {
  window_search_impl(kps, desc, n, g, q, qdesc, nq, skip, kp_u_right, q_u_right, q_max_err, NULL, out);
}

static void window_search_impl(const orc_kp* kps, const uint8_t* desc, int n, const orc_grid_geom* g,
                               const orc_window_query* q, const uint8_t* qdesc, int nq, const uint8_t* skip,
                               const float* kp_u_right, const float* q_u_right, const float* q_max_err,
                               const float* g_fuse_inv_sigma2, orc_window_result* out) {
  const int ncell = g->cols * g->rows;
  int* cnt = (int*)calloc((size_t)ncell + 1, sizeof(int));
  int* cell_of = (int*)malloc(sizeof(int) * (size_t)(n ? n : 1));
  for (int i = 0; i < n; i++) { /* frame.cc:438-465, :748-760 */
    const int px = (int)roundf((kps[i].x - g->min_x) * g->inv_w);
    const int py = (int)roundf((kps[i].y - g->min_y) * g->inv_h);
    if (px < 0 || px >= g->cols || py < 0 || py >= g->rows) { cell_of[i] = -1; continue; }
    cell_of[i] = px * g->rows + py;
    cnt[cell_of[i] + 1]++;
  }
  for (int c = 0; c < ncell; c++) cnt[c + 1] += cnt[c];
  int* fill = (int*)calloc((size_t)ncell, sizeof(int));
  int* items = (int*)malloc(sizeof(int) * (size_t)(n ? n : 1));
  for (int i = 0; i < n; i++) if (cell_of[i] >= 0) items[cnt[cell_of[i]] + fill[cell_of[i]]++] = i;

  for (int qi = 0; qi < nq; qi++) {
    orc_window_result r = {256, -1, -1, 256, -1};
    const float x = q[qi].u, y = q[qi].v, fr = q[qi].r;
    /* frame.cc:679-746 */
    int c0x = (int)floorf((x - g->min_x - fr) * g->inv_w); if (c0x < 0) c0x = 0;
    int c1x = (int)ceilf((x - g->min_x + fr) * g->inv_w); if (c1x > g->cols - 1) c1x = g->cols - 1;
    int c0y = (int)floorf((y - g->min_y - fr) * g->inv_h); if (c0y < 0) c0y = 0;
    int c1y = (int)ceilf((y - g->min_y + fr) * g->inv_h); if (c1y > g->rows - 1) c1y = g->rows - 1;
    const int ok = !(c0x >= g->cols || c1x < 0 || c0y >= g->rows || c1y < 0);
    const int check_levels = (q[qi].min_level >= 0) || (q[qi].max_level >= 0);
    if (ok)
      for (int ix = c0x; ix <= c1x; ix++)
        for (int iy = c0y; iy <= c1y; iy++) {
          const int c = ix * g->rows + iy;
          for (int k = cnt[c]; k < cnt[c + 1]; k++) {
            const int i = items[k];
            if (check_levels) {
              if (kps[i].octave < q[qi].min_level) continue;
              if (q[qi].max_level >= 0 && kps[i].octave > q[qi].max_level) continue;
            }
            const float dx = kps[i].x - x, dy = kps[i].y - y;
            if (!(fabsf(dx) < fr && fabsf(dy) < fr)) continue;
            if (skip && skip[i]) continue; /* orb_matcher.cc:86-87 "already matched" */
            if (g_fuse_inv_sigma2) { /* ORBmatcher::Fuse, orb_matcher.cc:1159-1178 */
              const float ex = x - kps[i].x, ey = y - kps[i].y;
              const float inv = g_fuse_inv_sigma2[kps[i].octave];
              if (kp_u_right && kp_u_right[i] >= 0) {
                const float er = q_u_right[qi] - kp_u_right[i];
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * inv > 7.8) continue;
              } else {
                const float e2 = ex * ex + ey * ey;
                if (e2 * inv > 5.99) continue;
              }
            } else
            if (kp_u_right && kp_u_right[i] > 0) { /* stereo observation: orb_matcher.cc:89-92, 1586-1590 */
              const float er = fabsf(q_u_right[qi] - kp_u_right[i]);
              if (er > q_max_err[qi]) continue;
            }
            const int dist = orc_hamming(qdesc + 32 * (size_t)qi, desc + 32 * (size_t)i);
            if (dist < r.best_dist) { /* orb_matcher.cc:98-112 */
              r.best_dist2 = r.best_dist; r.best_dist = dist;
              r.best_level2 = r.best_level; r.best_level = kps[i].octave; r.best_idx = i;
            } else if (dist < r.best_dist2) {
              r.best_level2 = kps[i].octave; r.best_dist2 = dist;
            }
          }
        }
    out[qi] = r;
  }
  free(cnt); free(cell_of); free(fill); free(items);
}

/* ---- ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches), orb_matcher.cc:215-386, for a frame with
 * Nleft == -1 (monocular / rectified stereo; the two-camera branches of :268-291 and :329-356 are not taken).
 * The two FeatureVectors are given as sorted node ids + group starts + feature indices (the layout
 * orc_vocab_transform / orbv_transform produce).  match_of_f[i] = key-frame feature whose map point frame
 * feature i receives (vpMapPointMatches[i]), -1 = none.  Returns nmatches. */
static void three_maxima(const int* histo, int L, int* ind1, int* ind2, int* ind3) { /* orb_matcher.cc:1841-1873 */
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = histo[i];
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; *ind3 = *ind2; *ind2 = *ind1; *ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; *ind3 = *ind2; *ind2 = i; }
    else if (s > max3) { max3 = s; *ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { *ind2 = -1; *ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { *ind3 = -1; }
}

int orc_search_by_bow(const orc_kp* kps_kf, const uint8_t* desc_kf, const uint8_t* has_point_kf,
                      const uint32_t* nodes_kf, const int* begin_kf, int n_nodes_kf, const uint32_t* feats_kf, int total_kf,
                      const orc_kp* kps_f, const uint8_t* desc_f, int n_f,
                      const uint32_t* nodes_f, const int* begin_f, int n_nodes_f, const uint32_t* feats_f, int total_f,
                      float nnratio, int check_orientation, int* match_of_f) {
  enum { HISTO_LENGTH = 30, TH_LOW = 50 };
  int nmatches = 0;
  int hist[HISTO_LENGTH] = {0};
  int* bin_of = (int*)malloc(sizeof(int) * (size_t)(n_f ? n_f : 1));
  const float factor = HISTO_LENGTH / 360.0f; /* :228 */
  for (int i = 0; i < n_f; i++) { match_of_f[i] = -1; bin_of[i] = -1; }
  int a = 0, b = 0;
  while (a < n_nodes_kf && b < n_nodes_f) { /* :237-375: the intersection of two sorted maps */
    if (nodes_kf[a] == nodes_f[b]) {
      const int k0 = begin_kf[a], k1 = a + 1 < n_nodes_kf ? begin_kf[a + 1] : total_kf;
      const int f0 = begin_f[b], f1 = b + 1 < n_nodes_f ? begin_f[b + 1] : total_f;
      for (int ik = k0; ik < k1; ik++) {
        const int real_kf = (int)feats_kf[ik];
        if (has_point_kf && !has_point_kf[real_kf]) continue; /* :246-250: no map point, or a bad one */
        int best1 = 256, best_idx = -1, best2 = 256;
        for (int jf = f0; jf < f1; jf++) {
          const int real_f = (int)feats_f[jf];
          if (match_of_f[real_f] >= 0) continue; /* :265 */
          const int dist = orc_hamming(desc_kf + 32 * (size_t)real_kf, desc_f + 32 * (size_t)real_f);
          if (dist < best1) { best2 = best1; best1 = dist; best_idx = real_f; }
          else if (dist < best2) { best2 = dist; }
        }
        if (best1 <= TH_LOW && (float)best1 < nnratio * (float)best2) { /* :307-309 */
          match_of_f[best_idx] = real_kf;
          if (check_orientation) { /* :318-328 */
            float rot = kps_kf[real_kf].angle - kps_f[best_idx].angle;
            if (rot < 0.0) rot += 360.0f;
            int bin = (int)roundf(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            bin_of[best_idx] = bin;
            hist[bin]++;
          }
          nmatches++;
        }
      }
      a++; b++;
    } else if (nodes_kf[a] < nodes_f[b]) {
      while (a < n_nodes_kf && nodes_kf[a] < nodes_f[b]) a++; /* lower_bound */
    } else {
      while (b < n_nodes_f && nodes_f[b] < nodes_kf[a]) b++;
    }
  }
  if (check_orientation) { /* :377-391 */
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(hist, HISTO_LENGTH, &ind1, &ind2, &ind3);
    for (int i = 0; i < n_f; i++)
      if (match_of_f[i] >= 0 && bin_of[i] != ind1 && bin_of[i] != ind2 && bin_of[i] != ind3) { match_of_f[i] = -1; nmatches--; }
  }
  free(bin_of);
  return nmatches;
}

/* ---- ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12), orb_matcher.cc:697-815 (NLeft == -1):
 * both sides need a good map point, a side-2 feature is claimed once (vbMatched2), the threshold is a strict
 * < TH_LOW, and the result is indexed by the side-1 feature: match_of_1[i] = side-2 feature or -1. */
int orc_search_by_bow_kf(const orc_kp* kps1, const uint8_t* desc1, const uint8_t* has_point1, int n1,
                         const uint32_t* nodes1, const int* begin1, int n_nodes1, const uint32_t* feats1, int total1,
                         const orc_kp* kps2, const uint8_t* desc2, const uint8_t* has_point2, int n2,
                         const uint32_t* nodes2, const int* begin2, int n_nodes2, const uint32_t* feats2, int total2,
                         float nnratio, int check_orientation, int* match_of_1) {
  enum { HISTO_LENGTH = 30, TH_LOW = 50 };
  int nmatches = 0;
  int hist[HISTO_LENGTH] = {0};
  int* bin_of = (int*)malloc(sizeof(int) * (size_t)(n1 ? n1 : 1));
  uint8_t* matched2 = (uint8_t*)calloc((size_t)(n2 ? n2 : 1), 1);
  const float factor = HISTO_LENGTH / 360.0f; /* :716 */
  for (int i = 0; i < n1; i++) { match_of_1[i] = -1; bin_of[i] = -1; }
  int a = 0, b = 0;
  while (a < n_nodes1 && b < n_nodes2) { /* :725-797 */
    if (nodes1[a] == nodes2[b]) {
      const int k0 = begin1[a], k1 = a + 1 < n_nodes1 ? begin1[a + 1] : total1;
      const int f0 = begin2[b], f1 = b + 1 < n_nodes2 ? begin2[b + 1] : total2;
      for (int i1 = k0; i1 < k1; i1++) {
        const int idx1 = (int)feats1[i1];
        if (has_point1 && !has_point1[idx1]) continue; /* :733-735 */
        int best1 = 256, best_idx2 = -1, best2 = 256;
        for (int i2 = f0; i2 < f1; i2++) {
          const int idx2 = (int)feats2[i2];
          if (matched2[idx2] || (has_point2 && !has_point2[idx2])) continue; /* :752-754 */
          const int dist = orc_hamming(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
          if (dist < best1) { best2 = best1; best1 = dist; best_idx2 = idx2; }
          else if (dist < best2) { best2 = dist; }
        }
        if (best1 < TH_LOW && (float)best1 < nnratio * (float)best2) { /* :769-771 */
          match_of_1[idx1] = best_idx2;
          matched2[best_idx2] = 1;
          if (check_orientation) { /* :775-782 */
            float rot = kps1[idx1].angle - kps2[best_idx2].angle;
            if (rot < 0.0) rot += 360.0f;
            int bin = (int)roundf(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            bin_of[idx1] = bin;
            hist[bin]++;
          }
          nmatches++;
        }
      }
      a++; b++;
    } else if (nodes1[a] < nodes2[b]) {
      while (a < n_nodes1 && nodes1[a] < nodes2[b]) a++;
    } else {
      while (b < n_nodes2 && nodes2[b] < nodes1[a]) b++;
    }
  }
  if (check_orientation) { /* :799-812 */
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(hist, HISTO_LENGTH, &ind1, &ind2, &ind3);
    for (int i = 0; i < n1; i++)
      if (match_of_1[i] >= 0 && bin_of[i] != ind1 && bin_of[i] != ind2 && bin_of[i] != ind3) { match_of_1[i] = -1; nmatches--; }
  }
  free(bin_of); free(matched2);
  return nmatches;
}

/* ---- ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints), orb_matcher.cc:42-134,
 * Nleft == -1: the window search above per map point, in order, with the greedy claim: a frame keypoint that holds a map
 * point with observations is skipped (:86-87), and an accepted match stores the map point (:121).  Queries are the map
 * points that passed :50-57, with the window of :62-70. */
int orc_search_by_projection(const orc_kp* kps, const uint8_t* desc, int n, const orc_grid_geom* g,
                             const orc_window_query* q, const uint8_t* qdesc, int nq, const uint8_t* skip,
                             const float* kp_u_right, const float* q_u_right, const float* q_max_err, int th_high,
                             float nnratio, int* assigned) {
  uint8_t* taken = (uint8_t*)calloc((size_t)(n ? n : 1), 1);
  int nmatches = 0;
  for (int i = 0; i < n; i++) { assigned[i] = -1; taken[i] = skip ? skip[i] : 0; }
  for (int qi = 0; qi < nq; qi++) {
    orc_window_result r;
    orc_window_search_stereo(kps, desc, n, g, q + qi, qdesc + 32 * (size_t)qi, 1, taken, kp_u_right,
                             kp_u_right ? q_u_right + qi : NULL, kp_u_right ? q_max_err + qi : NULL, &r);
    if (r.best_idx < 0) continue;                 /* :75 vIndices.empty() or every candidate skipped: bestDist stays 256 */
    if (r.best_dist <= th_high) {                 /* :117 */
      if (r.best_level == r.best_level2 && (float)r.best_dist > nnratio * (float)r.best_dist2) continue; /* :118-119 */
      assigned[r.best_idx] = qi;                  /* :121 */
      taken[r.best_idx] = 1;
      nmatches++;
    }
  }
  free(taken);
  return nmatches;
}

/* ---- ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono), orb_matcher.cc:1518-1728
 * (Nleft == -1), after the projection: per last-frame map point (in order) the nearest unclaimed keypoint of its window,
 * accepted within TH_HIGH (:1578-1604), then the rotation histogram (:1606-1624, :1706-1725). */
int orc_search_by_projection_last(const orc_kp* kps, const uint8_t* desc, int n, const orc_grid_geom* g,
                                  const orc_window_query* q, const uint8_t* qdesc, const float* q_angle, int nq,
                                  const uint8_t* skip, const float* kp_u_right, const float* q_u_right,
                                  const float* q_max_err, int th_high, int check_orientation, int* assigned) {
  enum { HISTO_LENGTH = 30 };
  uint8_t* taken = (uint8_t*)calloc((size_t)(n ? n : 1), 1);
  int* bin_of = (int*)malloc(sizeof(int) * (size_t)(n ? n : 1));
  int hist[HISTO_LENGTH] = {0};
  const float factor = HISTO_LENGTH / 360.0f; /* :1527 */
  int nmatches = 0;
  for (int i = 0; i < n; i++) { assigned[i] = -1; bin_of[i] = -1; taken[i] = skip ? skip[i] : 0; }
  for (int qi = 0; qi < nq; qi++) {
    orc_window_result r;
    orc_window_search_stereo(kps, desc, n, g, q + qi, qdesc + 32 * (size_t)qi, 1, taken, kp_u_right,
                             kp_u_right ? q_u_right + qi : NULL, kp_u_right ? q_max_err + qi : NULL, &r);
    if (r.best_idx < 0 || r.best_dist > th_high) continue; /* :1604 */
    assigned[r.best_idx] = qi;
    taken[r.best_idx] = 1;
    nmatches++;
    if (check_orientation) { /* :1608-1624 */
      float rot = q_angle[qi] - kps[r.best_idx].angle;
      if (rot < 0.0) rot += 360.0f;
      int bin = (int)roundf(rot * factor);
      if (bin == HISTO_LENGTH) bin = 0;
      bin_of[r.best_idx] = bin;
      hist[bin]++;
    }
  }
  if (check_orientation) { /* :1706-1725 */
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(hist, HISTO_LENGTH, &ind1, &ind2, &ind3);
    for (int i = 0; i < n; i++)
      if (assigned[i] >= 0 && bin_of[i] != ind1 && bin_of[i] != ind2 && bin_of[i] != ind3) { assigned[i] = -1; nmatches--; }
  }
  free(taken); free(bin_of);
  return nmatches;
}

/* ---- ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse), orb_matcher.cc:817-1040, for
 * key frames without a second camera (cam2_ == NULL, NLeft == -1) and pinhole cameras: per shared vocabulary node every
 * feature of key frame 1 WITHOUT a map point takes the nearest (ties: the LAST) feature of key frame 2 in the node that has
 * no map point, is not claimed yet, is within TH_LOW, lies away from the epipole (:917-924, monocular pairs only) and on
 * the epipolar line (Pinhole::EpipolarConstrain, pinhole_model.cc:121-134, with F12 given); then the rotation histogram.
 * match_of_1[i] = feature of key frame 2 (vMatches12) or -1; returns nmatches. */
int orc_search_for_triangulation(const orc_kp* kps1, const uint8_t* desc1, const uint8_t* has_point1, const float* u_right1, int n1,
                                 const uint32_t* nodes1, const int* begin1, int n_nodes1, const uint32_t* feats1, int total1,
                                 const orc_kp* kps2, const uint8_t* desc2, const uint8_t* has_point2, const float* u_right2, int n2,
                                 const uint32_t* nodes2, const int* begin2, int n_nodes2, const uint32_t* feats2, int total2,
                                 const float* f12 /* row-major 3x3 */, const float* ep, const float* scale_factors,
                                 const float* level_sigma2, int only_stereo, int coarse, int check_orientation, int* match_of_1) {
  enum { HISTO_LENGTH = 30, TH_LOW = 50 };
  int nmatches = 0;
  int hist[HISTO_LENGTH] = {0};
  int* bin_of = (int*)malloc(sizeof(int) * (size_t)(n1 ? n1 : 1));
  uint8_t* matched2 = (uint8_t*)calloc((size_t)(n2 ? n2 : 1), 1);
  const float factor = HISTO_LENGTH / 360.0f; /* :868 */
  for (int i = 0; i < n1; i++) { match_of_1[i] = -1; bin_of[i] = -1; }
  int a = 0, b = 0;
  while (a < n_nodes1 && b < n_nodes2) { /* :875-1008 */
    if (nodes1[a] == nodes2[b]) {
      const int k0 = begin1[a], k1 = a + 1 < n_nodes1 ? begin1[a + 1] : total1;
      const int f0 = begin2[b], f1 = b + 1 < n_nodes2 ? begin2[b + 1] : total2;
      for (int i1 = k0; i1 < k1; i1++) {
        const int idx1 = (int)feats1[i1];
        if (has_point1[idx1]) continue;                       /* :883-886 */
        const int stereo1 = u_right1[idx1] >= 0;              /* :888 */
        if (only_stereo && !stereo1) continue;
        int best_dist = TH_LOW, best_idx2 = -1;
        for (int i2 = f0; i2 < f1; i2++) {
          const int idx2 = (int)feats2[i2];
          if (matched2[idx2] || has_point2[idx2]) continue;   /* :911 */
          const int stereo2 = u_right2[idx2] >= 0;
          if (only_stereo && !stereo2) continue;
          const int dist = orc_hamming(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
          if (dist > TH_LOW || dist > best_dist) continue;    /* :922 */
          if (!stereo1 && !stereo2) {                         /* :932-939 */
            const float distex = ep[0] - kps2[idx2].x, distey = ep[1] - kps2[idx2].y;
            if (distex * distex + distey * distey < 100 * scale_factors[kps2[idx2].octave]) continue;
          }
          int ok = coarse;
          if (!ok) { /* pinhole_model.cc:121-134 */
            const float x1 = kps1[idx1].x, y1 = kps1[idx1].y, unc = level_sigma2[kps2[idx2].octave];
            const float la = x1 * f12[0] + y1 * f12[3] + f12[6];
            const float lb = x1 * f12[1] + y1 * f12[4] + f12[7];
            const float lc = x1 * f12[2] + y1 * f12[5] + f12[8];
            const float num = la * kps2[idx2].x + lb * kps2[idx2].y + lc;
            const float den = la * la + lb * lb;
            if (den != 0) {
              const float dsqr = num * num / den;
              ok = dsqr < 3.84 * unc;
            }
          }
          if (ok) { best_idx2 = idx2; best_dist = dist; }     /* :977-984 */
        }
        if (best_idx2 >= 0) {                                 /* :987-1006 */
          match_of_1[idx1] = best_idx2;
          matched2[best_idx2] = 1;
          nmatches++;
          if (check_orientation) {
            float rot = kps1[idx1].angle - kps2[best_idx2].angle;
            if (rot < 0.0) rot += 360.0f;
            int bin = (int)roundf(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            bin_of[idx1] = bin;
            hist[bin]++;
          }
        }
      }
      a++; b++;
    } else if (nodes1[a] < nodes2[b]) {
      while (a < n_nodes1 && nodes1[a] < nodes2[b]) a++;
    } else {
      while (b < n_nodes2 && nodes2[b] < nodes1[a]) b++;
    }
  }
  if (check_orientation) { /* :1010-1027 */
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(hist, HISTO_LENGTH, &ind1, &ind2, &ind3);
    for (int i = 0; i < n1; i++)
      if (match_of_1[i] >= 0 && bin_of[i] != ind1 && bin_of[i] != ind2 && bin_of[i] != ind3) { match_of_1[i] = -1; nmatches--; }
  }
  free(bin_of); free(matched2);
  return nmatches;
}

/* ---- synthetic inputs, SURVEY.md 8(d) ---- */
uint64_t orc_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

void orc_synth_blocks_v1(uint8_t* img, int w, int h, size_t stride, uint64_t seed, uint64_t frame,
                         int shift_x, uint64_t noise_seed) {
  const uint64_t base = orc_splitmix64((seed << 32) ^ frame);
  const uint64_t nbase = orc_splitmix64((noise_seed << 32) ^ frame);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) img[(size_t)y * stride + x] = (uint8_t)(40 + (160 * x) / (w - 1));
  const int R = (w * h) / 900;
  for (int k = 0; k < R; k++) {
    const int x0 = (int)(orc_splitmix64(base ^ (uint64_t)(5 * k + 1)) % (uint64_t)w) - shift_x;
    const int y0 = (int)(orc_splitmix64(base ^ (uint64_t)(5 * k + 2)) % (uint64_t)h);
    const int rw = 8 + (int)(orc_splitmix64(base ^ (uint64_t)(5 * k + 3)) % 82);
    const int rh = 8 + (int)(orc_splitmix64(base ^ (uint64_t)(5 * k + 4)) % 82);
    const uint8_t v = (uint8_t)(orc_splitmix64(base ^ (uint64_t)(5 * k + 5)) % 256);
    for (int y = y0; y < y0 + rh && y < h; y++)
      for (int x = (x0 < 0 ? 0 : x0); x < x0 + rw && x < w; x++) img[(size_t)y * stride + x] = v;
  }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const int nz = (int)(orc_splitmix64(nbase ^ 0xABCDEFull ^ ((uint64_t)y << 20) ^ (uint64_t)x) % 7) - 3;
      int v = img[(size_t)y * stride + x] + nz;
      img[(size_t)y * stride + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

void orc_synth_uniform_v1(uint8_t* img, int w, int h, size_t stride, uint64_t seed, uint64_t frame) {
  const uint64_t base = orc_splitmix64((seed << 32) ^ frame);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      img[(size_t)y * stride + x] = (uint8_t)(orc_splitmix64(base ^ ((uint64_t)y << 20) ^ (uint64_t)x) & 255);
}

void orc_synth_descriptors(uint8_t* rows, int64_t first, int64_t n, uint64_t seed) {
  for (int64_t i = 0; i < n; i++)
    for (int j = 0; j < 4; j++) {
      const uint64_t v = orc_splitmix64(seed ^ (uint64_t)(4 * (first + i) + j));
      memcpy(rows + 32 * (size_t)i + 8 * j, &v, 8); /* little-endian words */
    }
}
