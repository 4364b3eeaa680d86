/*
 * oracle/orb_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see orb_oracle.h).
 *
 * CPU restatement of ORBextractor (reference: /root/reference/src/orb_extractor/ORBextractor.cc, twin
 * ORBExtractor.cpp).  Every function cites the reference lines it follows.  Build with
 * -ffp-contract=off (the reference is built without FMA).
 */
#include "orb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "cv_prims.h"

#define PATCH_SIZE 31       /* ORBextractor.cc:70 */
#define HALF_PATCH_SIZE 15  /* :71 */
#define EDGE_THRESHOLD 19   /* :72 */

static const int kPattern[1024] = {
#include "orb_pattern_oracle.inc"
};

struct OrbOracle {
    int nfeatures, nlevels, ini_th, min_th, cell_w;
    double scale_factor; /* double member initialised from the float argument, inc/ORBextractor.h:98 */
    float *sf, *inv_sf, *sigma2, *inv_sigma2;
    int* quota;
    int umax[HALF_PATCH_SIZE + 1];
    /* per-frame state */
    int *lw, *lh;
    uint8_t** plane; /* bordered, pitch = lw + 38 */
    uint8_t** blur;  /* lw x lh or NULL */
    int *ncand, **cx, **cy, **cs;
    int* nkp;
    OrbOracleKeyPoint** kp;
};

static int floor_f(float v) { int i = (int)v; return i - (i > v); }
static int ceil_f(float v) { int i = (int)v; return i + (i < v); }

/* ORBextractor::ORBextractor, ORBextractor.cc:408-475 */
OrbOracle* orb_oracle_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int cell_w) {
    if (nlevels < 1 || nfeatures < 0 || cell_w < 1) return NULL;
    OrbOracle* o = (OrbOracle*)calloc(1, sizeof(OrbOracle));
    o->nfeatures = nfeatures; o->nlevels = nlevels; o->ini_th = ini_th; o->min_th = min_th; o->cell_w = cell_w;
    o->scale_factor = (double)scale_factor;
    const int L = nlevels;
    o->sf = (float*)calloc(L, sizeof(float)); o->inv_sf = (float*)calloc(L, sizeof(float));
    o->sigma2 = (float*)calloc(L, sizeof(float)); o->inv_sigma2 = (float*)calloc(L, sizeof(float));
    o->quota = (int*)calloc(L, sizeof(int));
    o->sf[0] = 1.0f; o->sigma2[0] = 1.0f;
    for (int i = 1; i < L; ++i) {                       /* :423-427 */
        o->sf[i] = (float)(o->sf[i - 1] * o->scale_factor);
        o->sigma2[i] = o->sf[i] * o->sf[i];
    }
    for (int i = 0; i < L; ++i) {                       /* :431-435 */
        o->inv_sf[i] = 1.0f / o->sf[i];
        o->inv_sigma2[i] = 1.0f / o->sigma2[i];
    }
    float factor = (float)(1.0f / o->scale_factor);     /* :440 */
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels)); /* :441 */
    int sum = 0;
    for (int l = 0; l < L - 1; ++l) {                   /* :445-450 */
        o->quota[l] = ocv_round_f(nDesired);
        sum += o->quota[l];
        nDesired *= factor;
    }
    o->quota[L - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0; /* :451 */

    /* umax, :459-474 */
    int v, v0;
    const int vmax = floor_f(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
    const int vmin = ceil_f(HALF_PATCH_SIZE * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) o->umax[v] = ocv_round_d(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (o->umax[v0] == o->umax[v0 + 1]) ++v0;
        o->umax[v] = v0;
        ++v0;
    }

    o->lw = (int*)calloc(L, sizeof(int)); o->lh = (int*)calloc(L, sizeof(int));
    o->plane = (uint8_t**)calloc(L, sizeof(uint8_t*)); o->blur = (uint8_t**)calloc(L, sizeof(uint8_t*));
    o->ncand = (int*)calloc(L, sizeof(int));
    o->cx = (int**)calloc(L, sizeof(int*)); o->cy = (int**)calloc(L, sizeof(int*)); o->cs = (int**)calloc(L, sizeof(int*));
    o->nkp = (int*)calloc(L, sizeof(int));
    o->kp = (OrbOracleKeyPoint**)calloc(L, sizeof(OrbOracleKeyPoint*));
    return o;
}

static void free_frame_state(OrbOracle* o) {
    for (int l = 0; l < o->nlevels; ++l) {
        free(o->plane[l]); o->plane[l] = NULL;
        free(o->blur[l]); o->blur[l] = NULL;
        free(o->cx[l]); free(o->cy[l]); free(o->cs[l]); o->cx[l] = o->cy[l] = o->cs[l] = NULL;
        free(o->kp[l]); o->kp[l] = NULL;
        o->ncand[l] = o->nkp[l] = 0;
    }
}

void orb_oracle_destroy(OrbOracle* o) {
    if (!o) return;
    free_frame_state(o);
    free(o->sf); free(o->inv_sf); free(o->sigma2); free(o->inv_sigma2); free(o->quota);
    free(o->lw); free(o->lh); free(o->plane); free(o->blur); free(o->ncand);
    free(o->cx); free(o->cy); free(o->cs); free(o->nkp); free(o->kp);
    free(o);
}

void orb_oracle_scale_table(const OrbOracle* o, int which, float* out) {
    const float* t = which == 0 ? o->sf : which == 1 ? o->inv_sf : which == 2 ? o->sigma2 : o->inv_sigma2;
    memcpy(out, t, sizeof(float) * (size_t)o->nlevels);
}
void orb_oracle_quota(const OrbOracle* o, int* out) { memcpy(out, o->quota, sizeof(int) * (size_t)o->nlevels); }
void orb_oracle_umax(const OrbOracle* o, int* out16) { memcpy(out16, o->umax, sizeof(int) * 16); }

/* ORBextractor::ComputePyramid, ORBextractor.cc:1164-1219 */
static void compute_pyramid(OrbOracle* o, const uint8_t* img, int w, int h, size_t stride) {
    const int E = EDGE_THRESHOLD;
    for (int l = 0; l < o->nlevels; ++l) {
        const float scale = o->inv_sf[l];
        const int sw = ocv_round_f((float)w * scale), sh = ocv_round_f((float)h * scale); /* :1171 */
        o->lw[l] = sw; o->lh[l] = sh;
        const size_t pitch = (size_t)sw + 2 * E;
        o->plane[l] = (uint8_t*)malloc(pitch * (size_t)(sh + 2 * E));                      /* :1173-1177 */
        uint8_t* roi = o->plane[l] + (size_t)E * pitch + E;
        if (l != 0) {
            const size_t ppitch = (size_t)o->lw[l - 1] + 2 * E;
            const uint8_t* proi = o->plane[l - 1] + (size_t)E * ppitch + E;
            ocv_resize_linear_u8(proi, o->lw[l - 1], o->lh[l - 1], ppitch, roi, sw, sh, pitch);  /* :1183 */
            ocv_copy_make_border_reflect101_u8(roi, sw, sh, pitch, o->plane[l], pitch, E, E, E, E); /* :1193 */
        } else {
            ocv_copy_make_border_reflect101_u8(img, w, h, stride, o->plane[l], pitch, E, E, E, E);  /* :1213 */
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * ExtractorNode::DivideNode + ORBextractor::DistributeOctTree, ORBextractor.cc:486-771.
 * Literal std::list semantics; the pointer tie-break of :689 is replaced by the creation sequence
 * number (== pointer order under a monotonic allocator).
 * ---------------------------------------------------------------------------------------------- */
typedef struct QNode {
    int x0, x1, y0, y1; /* UL.x, UR.x, UL.y, BR.y */
    int* keys;
    int nkeys;
    int no_more;
    long seq;
    struct QNode *prev, *next;
} QNode;

typedef struct { QNode *head, *tail; int size; long next_seq; } QList;

static QNode* qnode_new(int cap) {
    QNode* n = (QNode*)calloc(1, sizeof(QNode));
    n->keys = (int*)malloc(sizeof(int) * (size_t)(cap > 0 ? cap : 1));
    return n;
}
static void qnode_free(QNode* n) { free(n->keys); free(n); }
static void qlist_push_front(QList* L, QNode* n) {
    n->seq = L->next_seq++;
    n->prev = NULL; n->next = L->head;
    if (L->head) L->head->prev = n; else L->tail = n;
    L->head = n; L->size++;
}
static void qlist_push_back(QList* L, QNode* n) {
    n->seq = L->next_seq++;
    n->next = NULL; n->prev = L->tail;
    if (L->tail) L->tail->next = n; else L->head = n;
    L->tail = n; L->size++;
}
static QNode* qlist_erase(QList* L, QNode* n) { /* returns the following node */
    QNode* nx = n->next;
    if (n->prev) n->prev->next = n->next; else L->head = n->next;
    if (n->next) n->next->prev = n->prev; else L->tail = n->prev;
    L->size--;
    qnode_free(n);
    return nx;
}

/* DivideNode, :486-542 */
static void divide_node(const QNode* p, const int* xs, const int* ys, QNode* c[4]) {
    const int halfX = (int)ceilf((float)(p->x1 - p->x0) / 2);
    const int halfY = (int)ceilf((float)(p->y1 - p->y0) / 2);
    for (int q = 0; q < 4; ++q) c[q] = qnode_new(p->nkeys);
    c[0]->x0 = p->x0;         c[0]->x1 = p->x0 + halfX; c[0]->y0 = p->y0;         c[0]->y1 = p->y0 + halfY;
    c[1]->x0 = p->x0 + halfX; c[1]->x1 = p->x1;         c[1]->y0 = p->y0;         c[1]->y1 = p->y0 + halfY;
    c[2]->x0 = p->x0;         c[2]->x1 = p->x0 + halfX; c[2]->y0 = p->y0 + halfY; c[2]->y1 = p->y1;
    c[3]->x0 = p->x0 + halfX; c[3]->x1 = p->x1;         c[3]->y0 = p->y0 + halfY; c[3]->y1 = p->y1;
    for (int i = 0; i < p->nkeys; ++i) {
        const int k = p->keys[i];
        int q;
        if ((float)xs[k] < (float)c[0]->x1) q = ((float)ys[k] < (float)c[0]->y1) ? 0 : 2;
        else q = ((float)ys[k] < (float)c[0]->y1) ? 1 : 3;
        c[q]->keys[c[q]->nkeys++] = k;
    }
    for (int q = 0; q < 4; ++q) if (c[q]->nkeys == 1) c[q]->no_more = 1;
}

typedef struct { int size; QNode* node; } SizeNode;
static int cmp_size_node(const void* a, const void* b) { /* pair<int,ExtractorNode*> operator< */
    const SizeNode* x = (const SizeNode*)a; const SizeNode* y = (const SizeNode*)b;
    if (x->size != y->size) return x->size < y->size ? -1 : 1;
    if (x->node->seq != y->node->seq) return x->node->seq < y->node->seq ? -1 : 1;
    return 0;
}

int orb_oracle_distribute(const int* xs, const int* ys, const int* scores, int n,
                          int minX, int maxX, int minY, int maxY, int N, int* kept_idx, int cap) {
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));       /* :548 */
    if (nIni < 1) return -2;                                                   /* UB in the reference */
    const float hX = (float)(maxX - minX) / nIni;                              /* :550 */
    QList L = {NULL, NULL, 0, 0};
    QNode** ini = (QNode**)malloc(sizeof(QNode*) * (size_t)nIni);
    for (int i = 0; i < nIni; ++i) {                                           /* :557-568 */
        QNode* r = qnode_new(n);
        r->x0 = (int)(hX * (float)i); r->x1 = (int)(hX * (float)(i + 1));
        r->y0 = 0; r->y1 = maxY - minY;
        qlist_push_back(&L, r);
        ini[i] = r;
    }
    for (int k = 0; k < n; ++k) {                                              /* :571-575 */
        QNode* r = ini[(size_t)((float)xs[k] / hX)];
        r->keys[r->nkeys++] = k;
    }
    free(ini);
    for (QNode* it = L.head; it;) {                                            /* :577-590 */
        if (it->nkeys == 1) { it->no_more = 1; it = it->next; }
        else if (it->nkeys == 0) it = qlist_erase(&L, it);
        else it = it->next;
    }

    int finish = 0;
    SizeNode* vs = NULL; int nvs = 0, capvs = 0;
    while (!finish) {                                                          /* :599 */
        int prevSize = L.size;
        int nToExpand = 0;
        nvs = 0;
        for (QNode* it = L.head; it;) {                                        /* :611-670 */
            if (it->no_more) { it = it->next; continue; }
            QNode* c[4];
            divide_node(it, xs, ys, c);
            for (int q = 0; q < 4; ++q) {
                if (c[q]->nkeys > 0) {
                    qlist_push_front(&L, c[q]);
                    if (c[q]->nkeys > 1) {
                        nToExpand++;
                        if (nvs == capvs) { capvs = capvs ? capvs * 2 : 64; vs = (SizeNode*)realloc(vs, sizeof(SizeNode) * (size_t)capvs); }
                        vs[nvs].size = c[q]->nkeys; vs[nvs].node = c[q]; nvs++;
                    }
                } else qnode_free(c[q]);
            }
            it = qlist_erase(&L, it);
        }
        if (L.size >= N || L.size == prevSize) {                               /* :674-677 */
            finish = 1;
        } else if (L.size + nToExpand * 3 > N) {                               /* :678 */
            while (!finish) {
                prevSize = L.size;
                SizeNode* prevv = (SizeNode*)malloc(sizeof(SizeNode) * (size_t)(nvs > 0 ? nvs : 1));
                memcpy(prevv, vs, sizeof(SizeNode) * (size_t)nvs);
                const int nprev = nvs;
                nvs = 0;
                qsort(prevv, (size_t)nprev, sizeof(SizeNode), cmp_size_node);  /* :689 */
                for (int j = nprev - 1; j >= 0; --j) {                         /* :690-737 */
                    QNode* c[4];
                    divide_node(prevv[j].node, xs, ys, c);
                    for (int q = 0; q < 4; ++q) {
                        if (c[q]->nkeys > 0) {
                            qlist_push_front(&L, c[q]);
                            if (c[q]->nkeys > 1) {
                                if (nvs == capvs) { capvs = capvs ? capvs * 2 : 64; vs = (SizeNode*)realloc(vs, sizeof(SizeNode) * (size_t)capvs); }
                                vs[nvs].size = c[q]->nkeys; vs[nvs].node = c[q]; nvs++;
                            }
                        } else qnode_free(c[q]);
                    }
                    qlist_erase(&L, prevv[j].node);
                    if (L.size >= N) break;
                }
                free(prevv);
                if (L.size >= N || L.size == prevSize) finish = 1;             /* :739-740 */
            }
        }
    }
    free(vs);

    int nout = 0;
    for (QNode* it = L.head; it;) {                                            /* :751-767 */
        int best = it->keys[0];
        float maxResponse = (float)scores[best];
        for (int k = 1; k < it->nkeys; ++k)
            if ((float)scores[it->keys[k]] > maxResponse) { best = it->keys[k]; maxResponse = (float)scores[best]; }
        if (nout < cap) kept_idx[nout] = best;
        nout++;
        QNode* nx = it->next;
        qnode_free(it);
        it = nx;
    }
    return nout;
}

/* IC_Angle, ORBextractor.cc:75-102 */
float orb_oracle_ic_angle(const uint8_t* center, size_t step_, const int* umax) {
    int m_01 = 0, m_10 = 0;
    const int step = (int)step_;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return ocv_fast_atan2((float)m_01, (float)m_10);
}

/* computeOrbDescriptor, ORBextractor.cc:105-145 */
void orb_oracle_descriptor(const uint8_t* center, size_t step_, float kp_angle, uint8_t* desc) {
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float angle = kp_angle * factorPI;
    const float a = cosf(angle), b = sinf(angle);
    const int step = (int)step_;
    const int* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) { /* 16 points = 8 tests per byte */
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            const int x0 = pat[4 * k + 0], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
            const int t0 = center[ocv_round_f(x0 * b + y0 * a) * step + ocv_round_f(x0 * a - y0 * b)];
            const int t1 = center[ocv_round_f(x1 * b + y1 * a) * step + ocv_round_f(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ORBextractor::ComputeKeyPointsOctTree, ORBextractor.cc:773-888 */
static int compute_keypoints_octtree(OrbOracle* o) {
    const float W = (float)o->cell_w;                                          /* :777 */
    for (int level = 0; level < o->nlevels; ++level) {
        const int cols = o->lw[level], rows = o->lh[level];
        const size_t pitch = (size_t)cols + 2 * EDGE_THRESHOLD;
        const uint8_t* roi = o->plane[level] + (size_t)EDGE_THRESHOLD * pitch + EDGE_THRESHOLD;
        const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;    /* :781-784 */
        const int maxBorderX = cols - EDGE_THRESHOLD + 3, maxBorderY = rows - EDGE_THRESHOLD + 3;
        const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);         /* :792-793 */
        if (nCols < 1 || nRows < 1) return -2;                                 /* division by zero in the reference */
        const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows); /* :794-795 */

        int cap = 1024, n = 0;
        int *cx = (int*)malloc(sizeof(int) * cap), *cy = (int*)malloc(sizeof(int) * cap), *cs = (int*)malloc(sizeof(int) * cap);
        const int cellcap = (wCell + 7) * (hCell + 7);
        int *fx = (int*)malloc(sizeof(int) * cellcap), *fy = (int*)malloc(sizeof(int) * cellcap), *fs = (int*)malloc(sizeof(int) * cellcap);
        for (int i = 0; i < nRows; ++i) {                                      /* :797-864 */
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; ++j) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
                const uint8_t* cell = roi + (size_t)y0 * pitch + x0;
                int m = ocv_fast9_16_nms(cell, x1 - x0, y1 - y0, pitch, o->ini_th, fx, fy, fs, cellcap);   /* :818 */
                if (m == 0) m = ocv_fast9_16_nms(cell, x1 - x0, y1 - y0, pitch, o->min_th, fx, fy, fs, cellcap); /* :837 */
                for (int k = 0; k < m; ++k) {                                  /* :855-860 */
                    if (n == cap) {
                        cap *= 2;
                        cx = (int*)realloc(cx, sizeof(int) * cap); cy = (int*)realloc(cy, sizeof(int) * cap); cs = (int*)realloc(cs, sizeof(int) * cap);
                    }
                    cx[n] = fx[k] + j * wCell; cy[n] = fy[k] + i * hCell; cs[n] = fs[k]; n++;
                }
            }
        }
        free(fx); free(fy); free(fs);
        o->cx[level] = cx; o->cy[level] = cy; o->cs[level] = cs; o->ncand[level] = n;

        const int N = o->quota[level];
        const int capk = n + 8;
        int* kept = (int*)malloc(sizeof(int) * (size_t)capk);
        const int nk = orb_oracle_distribute(cx, cy, cs, n, minBorderX, maxBorderX, minBorderY, maxBorderY, N, kept, capk); /* :869 */
        if (nk < 0) { free(kept); return nk; }
        const int scaledPatchSize = (int)(PATCH_SIZE * o->sf[level]);          /* :872 */
        o->kp[level] = (OrbOracleKeyPoint*)malloc(sizeof(OrbOracleKeyPoint) * (size_t)(nk > 0 ? nk : 1));
        o->nkp[level] = nk;
        for (int i = 0; i < nk; ++i) {                                         /* :875-882 */
            OrbOracleKeyPoint* kp = &o->kp[level][i];
            kp->x = (float)cx[kept[i]] + minBorderX; kp->y = (float)cy[kept[i]] + minBorderY;
            kp->size = (float)scaledPatchSize; kp->angle = -1.f; kp->response = (float)cs[kept[i]];
            kp->octave = level; kp->class_id = -1;
        }
        free(kept);
    }
    for (int level = 0; level < o->nlevels; ++level) {                         /* :886-887 computeOrientation */
        const size_t pitch = (size_t)o->lw[level] + 2 * EDGE_THRESHOLD;
        const uint8_t* roi = o->plane[level] + (size_t)EDGE_THRESHOLD * pitch + EDGE_THRESHOLD;
        for (int i = 0; i < o->nkp[level]; ++i) {
            OrbOracleKeyPoint* kp = &o->kp[level][i];
            const uint8_t* c = roi + (size_t)ocv_round_f(kp->y) * pitch + ocv_round_f(kp->x);
            kp->angle = orb_oracle_ic_angle(c, pitch, o->umax);
        }
    }
    return 0;
}

/* ORBextractor::operator(), ORBextractor.cc:1078-1162 */
int orb_oracle_extract(OrbOracle* o, const uint8_t* img, int w, int h, size_t stride, int lap0, int lap1,
                       OrbOracleKeyPoint* kps, uint8_t* desc, int cap, int* n_out) {
    if (n_out) *n_out = 0;
    if (!img || w <= 0 || h <= 0) return -1;                                   /* :1083 */
    free_frame_state(o);
    /* Geometry the reference cannot handle (a level without a single 30-px cell: division by zero at :794; levels of
     * zero size: cv::resize asserts) is rejected before any pixel work -- the restatement must not loop on it either. */
    for (int l = 0; l < o->nlevels; ++l) {
        const int sw = ocv_round_f((float)w * o->inv_sf[l]), sh = ocv_round_f((float)h * o->inv_sf[l]);
        const float width = (float)(sw - 2 * (EDGE_THRESHOLD - 3)), height = (float)(sh - 2 * (EDGE_THRESHOLD - 3));
        if (sw < 1 || sh < 1 || (int)(width / (float)o->cell_w) < 1 || (int)(height / (float)o->cell_w) < 1) return -2;
    }
    compute_pyramid(o, img, w, h, stride);                                     /* :1090 */
    int rc = compute_keypoints_octtree(o);                                     /* :1093 */
    if (rc < 0) return rc;
    int nkeypoints = 0;
    for (int l = 0; l < o->nlevels; ++l) nkeypoints += o->nkp[l];              /* :1099-1101 */
    if (n_out) *n_out = nkeypoints;
    int monoIndex = 0, stereoIndex = nkeypoints - 1;                           /* :1116 */
    uint8_t d[32];
    for (int level = 0; level < o->nlevels; ++level) {                         /* :1117 */
        const int nl = o->nkp[level];
        if (nl == 0) continue;                                                 /* :1122 */
        const int cols = o->lw[level], rows = o->lh[level];
        const size_t pitch = (size_t)cols + 2 * EDGE_THRESHOLD;
        const uint8_t* roi = o->plane[level] + (size_t)EDGE_THRESHOLD * pitch + EDGE_THRESHOLD;
        o->blur[level] = (uint8_t*)malloc((size_t)cols * rows);
        ocv_gaussian_blur_7x7_s2_u8(roi, cols, rows, pitch, o->blur[level], (size_t)cols); /* :1126-1127 */
        const float scale = o->sf[level];
        for (int i = 0; i < nl; ++i) {
            OrbOracleKeyPoint kp = o->kp[level][i];
            const uint8_t* c = o->blur[level] + (size_t)ocv_round_f(kp.y) * cols + ocv_round_f(kp.x);
            orb_oracle_descriptor(c, (size_t)cols, kp.angle, d);               /* :1131-1132 */
            if (level != 0) { kp.x *= scale; kp.y *= scale; }                  /* :1143-1145 */
            int idx;
            if (kp.x >= lap0 && kp.x <= lap1) idx = stereoIndex--;            /* :1147-1156 */
            else idx = monoIndex++;
            if (idx < cap) {
                if (kps) kps[idx] = kp;
                if (desc) memcpy(desc + (size_t)idx * 32, d, 32);
            }
        }
    }
    return monoIndex;                                                          /* :1161 */
}

int orb_oracle_level_size(const OrbOracle* o, int level, int* w, int* h) {
    if (level < 0 || level >= o->nlevels) return -1;
    *w = o->lw[level]; *h = o->lh[level];
    return 0;
}
const uint8_t* orb_oracle_level_plane(const OrbOracle* o, int level, size_t* pitch) {
    if (pitch) *pitch = (size_t)o->lw[level] + 2 * EDGE_THRESHOLD;
    return o->plane[level];
}
const uint8_t* orb_oracle_level_blur(const OrbOracle* o, int level) { return o->blur[level]; }
int orb_oracle_level_candidates(const OrbOracle* o, int level, int* xs, int* ys, int* scores, int cap) {
    const int n = o->ncand[level];
    for (int i = 0; i < n && i < cap; ++i) { xs[i] = o->cx[level][i]; ys[i] = o->cy[level][i]; scores[i] = o->cs[level][i]; }
    return n;
}
int orb_oracle_level_keypoints(const OrbOracle* o, int level, OrbOracleKeyPoint* kps, int cap) {
    const int n = o->nkp[level];
    for (int i = 0; i < n && i < cap; ++i) kps[i] = o->kp[level][i];
    return n;
}
