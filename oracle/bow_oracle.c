/* oracle/bow_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the reference's DBoW2 transform.
 * See bow_oracle.h for the reference lines each function follows. */
#include "bow_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct orc_vocab {
  int k, L, scoring, weighting;
  int n_nodes, n_words;
  int32_t* parent;
  uint8_t* is_leaf;
  uint8_t* desc;      /* n_nodes x 32 */
  double* weight;
  int32_t* word_id;   /* per node; 0 for inner nodes like Node() (:322) */
  int32_t* child_beg; /* CSR of the children lists, in push_back order (increasing node id) */
  int32_t* child;
};

static uint64_t sm64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* FORB::distance, FORB.cpp:71-88 (the bit-twiddling popcount over 8 x 32 bits) */
static int forb_distance(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t pa, pb;
    memcpy(&pa, a + 4 * i, 4);
    memcpy(&pb, b + 4 * i, 4);
    uint32_t v = pa ^ pb;
    v = v - ((v >> 1) & 0x55555555u);
    v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
    dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
  }
  return dist;
}

orc_vocab* orc_vocab_create(int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent,
                            const uint8_t* is_leaf, const uint8_t* desc, const double* weight) {
  if (n_nodes < 1) return NULL;
  for (int i = 1; i < n_nodes; i++)
    if (parent[i] < 0 || parent[i] >= n_nodes) return NULL;
  orc_vocab* v = (orc_vocab*)calloc(1, sizeof(orc_vocab));
  v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->n_nodes = n_nodes;
  v->parent = (int32_t*)calloc((size_t)n_nodes, sizeof(int32_t));
  v->is_leaf = (uint8_t*)calloc((size_t)n_nodes, 1);
  v->desc = (uint8_t*)calloc((size_t)n_nodes, 32);
  v->weight = (double*)calloc((size_t)n_nodes, sizeof(double));
  v->word_id = (int32_t*)calloc((size_t)n_nodes, sizeof(int32_t));
  v->child_beg = (int32_t*)calloc((size_t)n_nodes + 1, sizeof(int32_t));
  v->child = (int32_t*)calloc((size_t)n_nodes, sizeof(int32_t));
  for (int i = 1; i < n_nodes; i++) {
    v->parent[i] = parent[i];
    v->is_leaf[i] = is_leaf[i] ? 1 : 0;
    memcpy(v->desc + 32 * (size_t)i, desc + 32 * (size_t)i, 32);
    v->weight[i] = weight[i];
    if (is_leaf[i]) v->word_id[i] = v->n_words++;  /* :1319-1324 */
    v->child_beg[parent[i] + 1]++;
  }
  for (int i = 0; i < n_nodes; i++) v->child_beg[i + 1] += v->child_beg[i];
  int32_t* fill = (int32_t*)calloc((size_t)n_nodes, sizeof(int32_t));
  for (int i = 1; i < n_nodes; i++) v->child[v->child_beg[parent[i]] + fill[parent[i]]++] = i;  /* :1300 push_back in file order */
  free(fill);
  return v;
}

void orc_vocab_destroy(orc_vocab* v) {
  if (!v) return;
  free(v->parent); free(v->is_leaf); free(v->desc); free(v->weight); free(v->word_id); free(v->child_beg); free(v->child);
  free(v);
}

int orc_vocab_nodes(const orc_vocab* v) { return v->n_nodes; }
int orc_vocab_words(const orc_vocab* v) { return v->n_words; }

void orc_vocab_arrays(const orc_vocab* v, int* k, int* L, int* scoring, int* weighting, int32_t* parent, uint8_t* is_leaf,
                      uint8_t* desc, double* weight) {
  if (k) *k = v->k;
  if (L) *L = v->L;
  if (scoring) *scoring = v->scoring;
  if (weighting) *weighting = v->weighting;
  if (parent) memcpy(parent, v->parent, sizeof(int32_t) * (size_t)v->n_nodes);
  if (is_leaf) memcpy(is_leaf, v->is_leaf, (size_t)v->n_nodes);
  if (desc) memcpy(desc, v->desc, 32 * (size_t)v->n_nodes);
  if (weight) memcpy(weight, v->weight, sizeof(double) * (size_t)v->n_nodes);
}

/* loadFromTextFile, TemplatedVocabulary.h:1246-1330: header "k L scoring weighting", then one line per
 * node "parent isLeaf d0 .. d31 weight"; node ids are line numbers.  An empty last line is skipped (the
 * reference would append a phantom node there, see bow_oracle.h). */
orc_vocab* orc_vocab_load_text(const char* path) {
  FILE* f = fopen(path, "r");
  if (!f) return NULL;
  int k, L, n1, n2;
  if (fscanf(f, "%d %d %d %d", &k, &L, &n1, &n2) != 4 || k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 ||
      n2 > 3) {  /* :1267-1272 */
    fclose(f);
    return NULL;
  }
  int cap = 1024, n = 1;
  int32_t* parent = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
  uint8_t* leaf = (uint8_t*)malloc((size_t)cap);
  uint8_t* desc = (uint8_t*)malloc(32 * (size_t)cap);
  double* w = (double*)malloc(sizeof(double) * (size_t)cap);
  parent[0] = 0; leaf[0] = 0; w[0] = 0; memset(desc, 0, 32);
  for (;;) {
    int pid, il;
    if (fscanf(f, "%d %d", &pid, &il) != 2) break;
    if (n == cap) {
      cap *= 2;
      parent = (int32_t*)realloc(parent, sizeof(int32_t) * (size_t)cap);
      leaf = (uint8_t*)realloc(leaf, (size_t)cap);
      desc = (uint8_t*)realloc(desc, 32 * (size_t)cap);
      w = (double*)realloc(w, sizeof(double) * (size_t)cap);
    }
    parent[n] = pid;
    leaf[n] = il > 0;  /* :1318 */
    int ok = 1;
    for (int i = 0; i < 32; i++) {
      int e;
      if (fscanf(f, "%d", &e) != 1) { ok = 0; break; }
      desc[32 * (size_t)n + i] = (uint8_t)e;  /* FORB::fromString, FORB.cpp:105-116 */
    }
    if (!ok || fscanf(f, "%lf", &w[n]) != 1) break;
    n++;
  }
  fclose(f);
  orc_vocab* v = orc_vocab_create(k, L, n1, n2, n, parent, leaf, desc, w);
  free(parent); free(leaf); free(desc); free(w);
  return v;
}

/* TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup), :1139-1179 */
static void transform_one(const orc_vocab* v, const uint8_t* feature, int levelsup, uint32_t* word_id, double* weight,
                          uint32_t* nid) {
  const int nid_level = v->L - levelsup;
  int nid_set = 0;
  if (nid_level <= 0) { *nid = 0; nid_set = 1; }  /* root, :1151 */
  int final_id = 0, current_level = 0;
  if (v->child_beg[1] == v->child_beg[0]) {  /* empty(): no words (:1063) -- callers skip; keep the outputs defined */
    *word_id = 0; *weight = 0; *nid = 0;
    return;
  }
  do {
    ++current_level;
    const int b = v->child_beg[final_id], e = v->child_beg[final_id + 1];
    final_id = v->child[b];
    int best_d = forb_distance(feature, v->desc + 32 * (size_t)final_id);  /* a double in the reference; holds an int */
    for (int c = b + 1; c < e; c++) {
      const int id = v->child[c];
      const int d = forb_distance(feature, v->desc + 32 * (size_t)id);
      if (d < best_d) { best_d = d; final_id = id; }  /* strict <: the first minimum wins, :1166 */
    }
    if (current_level == nid_level) { *nid = (uint32_t)final_id; nid_set = 1; }
  } while (v->child_beg[final_id + 1] != v->child_beg[final_id]);  /* !isLeaf(): children non-empty, :1174 */
  if (!nid_set) *nid = (uint32_t)final_id;  /* unset in the reference (leaf above nid_level) */
  *word_id = (uint32_t)v->word_id[final_id];
  *weight = v->weight[final_id];
}

void orc_bow_features(const orc_vocab* v, const uint8_t* desc, int n, int levelsup, uint32_t* word_id, double* weight,
                      uint32_t* node_id) {
  for (int i = 0; i < n; i++) transform_one(v, desc + 32 * (size_t)i, levelsup, &word_id[i], &weight[i], &node_id[i]);
}

typedef struct { uint32_t key; uint32_t idx; } kv;
static int kv_cmp(const void* a, const void* b) {
  const kv *x = (const kv*)a, *y = (const kv*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

/* TemplatedVocabulary::transform(features, v, fv, levelsup), :1056-1118.  The std::map insertions in feature
 * order are restated as: sort (word, feature index), then accumulate each word's weights in feature
 * order (addWeight: v += w, :30-38) or keep the first (addIfNotExist, :42-48). */
void orc_bow_transform(const orc_vocab* v, const uint8_t* desc, int n, int levelsup, uint32_t* bow_ids, double* bow_vals,
                       int* bow_n, uint32_t* fv_nodes, int32_t* fv_begin, int* fv_n, uint32_t* fv_feats, int* fv_total) {
  *bow_n = 0; *fv_n = 0; *fv_total = 0;
  if (n <= 0 || v->n_words == 0) return;  /* empty(), :1063 */
  uint32_t* wid = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
  uint32_t* nid = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
  double* w = (double*)malloc(sizeof(double) * (size_t)n);
  kv* a = (kv*)malloc(sizeof(kv) * (size_t)n);
  kv* b = (kv*)malloc(sizeof(kv) * (size_t)n);
  int m = 0;
  for (int i = 0; i < n; i++) {
    transform_one(v, desc + 32 * (size_t)i, levelsup, &wid[i], &w[i], &nid[i]);
    if (w[i] > 0) {  /* not stopped, :1084 / :1109 */
      a[m].key = wid[i]; a[m].idx = (uint32_t)i;
      b[m].key = nid[i]; b[m].idx = (uint32_t)i;
      m++;
    }
  }
  *fv_total = m;
  qsort(a, (size_t)m, sizeof(kv), kv_cmp);
  qsort(b, (size_t)m, sizeof(kv), kv_cmp);
  const int tf = v->weighting == 0 || v->weighting == 1;  /* TF_IDF or TF: addWeight; IDF / BINARY: addIfNotExist */
  int nb = 0;
  for (int i = 0; i < m;) {
    int j = i;
    double val = w[a[i].idx];
    for (j = i + 1; j < m && a[j].key == a[i].key; j++)
      if (tf) val += w[a[j].idx];
    bow_ids[nb] = a[i].key;
    bow_vals[nb] = val;
    nb++;
    i = j;
  }
  /* normalisation, :1068-1070, 1091-1096, 1117: L1 / L2 when the scoring needs it, else TF weights / size */
  const int must = v->scoring != 5;            /* ScoringObject.h:76-91: all but DOT_PRODUCT */
  const int l2 = v->scoring == 1;
  if (tf && nb > 0 && !must) {
    const double nd = (double)nb;
    for (int i = 0; i < nb; i++) bow_vals[i] /= nd;
  }
  if (must) {  /* BowVector::normalize, BowVector.cpp:52-66 */
    double norm = 0.0;
    if (!l2) {
      for (int i = 0; i < nb; i++) norm += fabs(bow_vals[i]);
    } else {
      for (int i = 0; i < nb; i++) norm += bow_vals[i] * bow_vals[i];
      norm = sqrt(norm);
    }
    if (norm > 0.0)
      for (int i = 0; i < nb; i++) bow_vals[i] /= norm;
  }
  *bow_n = nb;
  int nf = 0;
  for (int i = 0; i < m; i++) {
    if (i == 0 || b[i].key != b[i - 1].key) {
      fv_nodes[nf] = b[i].key;
      fv_begin[nf] = i;
      nf++;
    }
    fv_feats[i] = b[i].idx;
  }
  *fv_n = nf;
  free(wid); free(nid); free(w); free(a); free(b);
}

int orc_synth_vocab(int k, int L, uint64_t seed, int32_t* parent, uint8_t* is_leaf, uint8_t* desc, double* weight) {
  int total = 1, level_n = 1;
  for (int l = 1; l <= L; l++) { level_n *= k; total += level_n; }
  if (!parent) return total;
  /* breadth-first: level l occupies ids [first_l, first_l + k^l); the children of node p at position j of
   * its level are first_{l+1} + k*j .. + k-1 */
  int first = 1, prev_first = 0, prev_n = 1, word = 0;
  parent[0] = 0; is_leaf[0] = 0; weight[0] = 0; memset(desc, 0, 32);
  for (int l = 1; l <= L; l++) {
    const int n = prev_n * k;
    const int flips = 256 >> l;  /* 128, 64, 32, ... bits away from the parent */
    for (int j = 0; j < n; j++) {
      const int id = first + j, p = prev_first + j / k;
      parent[id] = p;
      is_leaf[id] = l == L;
      uint8_t* d = desc + 32 * (size_t)id;
      if (l == 1) {
        for (int q = 0; q < 4; q++) {
          const uint64_t r = sm64(seed ^ (4 * (uint64_t)id + (uint64_t)q));
          memcpy(d + 8 * q, &r, 8);
        }
      } else {
        memcpy(d, desc + 32 * (size_t)p, 32);
        for (int t = 0; t < flips; t++) {
          const unsigned bit = (unsigned)(sm64(seed ^ 0x5151515151ull ^ ((uint64_t)id << 8) ^ (uint64_t)t) & 255u);
          d[bit >> 3] ^= (uint8_t)(1u << (bit & 7));
        }
      }
      if (l == L) {
        weight[id] = (word % 29 == 28) ? 0.0 : (double)(sm64(seed ^ 0xA5A5ull ^ (uint64_t)id) % 997u + 1u) / 64.0;
        word++;
      } else {
        weight[id] = 0.0;
      }
    }
    prev_first = first; prev_n = n; first += n;
  }
  return total;
}

int orc_vocab_save_text(const char* path, int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent,
                        const uint8_t* is_leaf, const uint8_t* desc, const double* weight) {
  FILE* f = fopen(path, "w");
  if (!f) return -1;
  fprintf(f, "%d %d  %d %d", k, L, scoring, weighting);  /* saveToTextFile, :1339-1341 */
  for (int i = 1; i < n_nodes; i++) {
    fprintf(f, "\n%d %d ", parent[i], is_leaf[i] ? 1 : 0);
    for (int j = 0; j < 32; j++) fprintf(f, "%d ", desc[32 * (size_t)i + j]);  /* FORB::toString: bytes as decimals */
    fprintf(f, "%.17g", weight[i]);
  }
  fclose(f);
  return 0;
}
