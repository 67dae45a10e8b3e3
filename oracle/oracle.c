/*
 * oracle.c -- plain-C CPU restatement of nimrud's multiscale eigenfeature path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load the library built from this file.
 * Nothing under nimrud_b200/ links or calls it.
 *
 * What it restates (citations relative to /root/reference):
 *   orc_grid_widths        nimrud/utils/geometry.py:37-62   (corners, ceil(log2(span/e)), 64-bit check)
 *   orc_unique_voxels      nimrud/utils/geometry.py:103-116 (floor((p-min)/e) -> packed address),
 *                          :150 (np.unique = ascending sort + dedup), :120-138 ((k*e+min)+e*0.5)
 *   orc_radius_*           nimrud/minimal/multiscale.py:103 (scipy cKDTree.query_ball_tree, p=2, eps=0:
 *                          member iff dx*dx+dy*dy+dz*dz <= r*r in float64, inclusive; lists ascending)
 *   feature columns        nimrud/minimal/features.py:21-57 (population, |q-mean|, eigvalsh(cov)/sum,
 *                          two largest, largest first); undefined -> 0 (multiscale.py:4-5)
 *   orc_process            nimrud/minimal/multiscale.py:27-123
 *   orc_knn                no reference code (PARITY UNPINNED): total order (d^2, index)
 *
 * Differences from the reference that do not change results beyond float64 rounding:
 *   neighbor search walks the voxel lattice through a hash of the unique addresses instead of a
 *   kd-tree (the membership predicate is the same expression); eigenvalues come from cyclic
 *   Jacobi on the two-pass covariance instead of LAPACK syevd.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -pthread -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

/* ------------------------------------------------------------------ tiny pthread parallel-for
 * (this image has no libgomp).  work is handed out in blocks through an atomic cursor. */
static int g_threads = 1;

typedef void (*orc_body)(int64_t begin, int64_t end, void *ctx);
typedef struct { orc_body body; void *ctx; int64_t n, block; int64_t *cursor; } orc_job;

static void *orc_worker(void *arg)
{
    orc_job *job = (orc_job *)arg;
    for (;;) {
        int64_t b = __atomic_fetch_add(job->cursor, job->block, __ATOMIC_RELAXED);
        if (b >= job->n) break;
        int64_t e = b + job->block < job->n ? b + job->block : job->n;
        job->body(b, e, job->ctx);
    }
    return NULL;
}

static void orc_parallel_for(int64_t n, int64_t block, orc_body body, void *ctx)
{
    int64_t cursor = 0;
    orc_job job = {body, ctx, n, block, &cursor};
    int nt = g_threads;
    if (nt > 64) nt = 64;
    if (nt <= 1 || n <= block) { orc_worker(&job); return; }
    pthread_t tid[64];
    int started = 0;
    for (int t = 0; t < nt - 1; ++t)
        if (pthread_create(&tid[started], NULL, orc_worker, &job) == 0) ++started;
    orc_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
}

typedef struct {
    double minc[3];
    double edge;
    int64_t widths[3];
    int64_t shifts[3];     /* shifts[0] = 0 */
} orc_grid;

/* ------------------------------------------------------------------ a1 */
int orc_grid_widths(const double *pts, int64_t n, double edge, double *minc, double *maxc,
                    int64_t *widths)
{
    if (n < 2) return 1;
    double lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = pts[a]; hi[a] = pts[a]; }
    for (int64_t i = 1; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            double v = pts[3 * i + a];
            if (v < lo[a]) lo[a] = v;
            if (v > hi[a]) hi[a] = v;
        }
    double total = 0;
    for (int a = 0; a < 3; ++a) {
        minc[a] = lo[a] - edge / 2;
        maxc[a] = hi[a] + edge / 2;
        double w = ceil(log2((maxc[a] - minc[a]) / edge));
        widths[a] = (int64_t)w;
        total += w;
    }
    return total > 64 ? 2 : 0;
}

static void grid_init(orc_grid *g, const double *minc, double edge, const int64_t *widths)
{
    for (int a = 0; a < 3; ++a) { g->minc[a] = minc[a]; g->widths[a] = widths[a]; }
    g->edge = edge;
    g->shifts[0] = 0;
    g->shifts[1] = widths[0];
    g->shifts[2] = widths[0] + widths[1];
}

static inline int64_t cell_of(const orc_grid *g, double v, int a)
{
    return (int64_t)floor((v - g->minc[a]) / g->edge);
}

static inline double centre_of(const orc_grid *g, int64_t k, int a)
{
    double t = (double)k * g->edge;
    t = t + g->minc[a];
    return t + g->edge * 0.5;
}

static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* ------------------------------------------------------------------ a2-a4 */
int64_t orc_unique_voxels(const double *pts, int64_t n, const double *minc, double edge,
                          const int64_t *widths, int64_t *keys_out, double *centres_out)
{
    orc_grid g;
    grid_init(&g, minc, edge, widths);
    for (int64_t i = 0; i < n; ++i) {
        int64_t kx = cell_of(&g, pts[3 * i + 0], 0);
        int64_t ky = cell_of(&g, pts[3 * i + 1], 1);
        int64_t kz = cell_of(&g, pts[3 * i + 2], 2);
        keys_out[i] = kx + (ky << g.shifts[1]) + (kz << g.shifts[2]);
    }
    qsort(keys_out, (size_t)n, sizeof(int64_t), cmp_i64);
    int64_t nv = 0;
    for (int64_t i = 0; i < n; ++i)
        if (i == 0 || keys_out[i] != keys_out[i - 1]) keys_out[nv++] = keys_out[i];
    if (centres_out)
        for (int64_t i = 0; i < nv; ++i) {
            int64_t k = keys_out[i];
            for (int a = 0; a < 3; ++a) {
                int64_t c = (k >> g.shifts[a]) & ((widths[a] >= 64) ? -1LL : ((1LL << widths[a]) - 1));
                centres_out[3 * i + a] = centre_of(&g, c, a);
            }
        }
    return nv;
}

/* ------------------------------------------------------------------ address hash */
typedef struct {
    int64_t *key;
    int64_t *val;
    uint64_t mask;
} orc_hash;

static inline uint64_t mix(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

static int hash_build(orc_hash *h, const int64_t *keys, int64_t n)
{
    uint64_t cap = 16;
    while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
    h->key = (int64_t *)malloc(cap * sizeof(int64_t));
    h->val = (int64_t *)malloc(cap * sizeof(int64_t));
    if (!h->key || !h->val) return 1;
    h->mask = cap - 1;
    for (uint64_t i = 0; i < cap; ++i) h->key[i] = -1;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t s = mix((uint64_t)keys[i]) & h->mask;
        while (h->key[s] != -1) s = (s + 1) & h->mask;
        h->key[s] = keys[i];
        h->val[s] = i;
    }
    return 0;
}

static inline int64_t hash_find(const orc_hash *h, int64_t key)
{
    uint64_t s = mix((uint64_t)key) & h->mask;
    while (h->key[s] != -1) {
        if (h->key[s] == key) return h->val[s];
        s = (s + 1) & h->mask;
    }
    return -1;
}

static void hash_free(orc_hash *h) { free(h->key); free(h->val); }

/* ------------------------------------------------------------------ eigenvalues */
/* cyclic Jacobi on a symmetric 3x3; w ascending */
static void eig3(const double c[6] /* xx xy xz yy yz zz */, double w[3], double v[9])
{
    double a[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
    double q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off <= 1e-300 || off <= 1e-22 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int r = p + 1; r < 3; ++r) {
                if (a[p][r] == 0.0) continue;
                double theta = (a[r][r] - a[p][p]) / (2.0 * a[p][r]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < 3; ++k) {
                    double akp = a[k][p], akr = a[k][r];
                    a[k][p] = cs * akp - sn * akr;
                    a[k][r] = sn * akp + cs * akr;
                }
                for (int k = 0; k < 3; ++k) {
                    double apk = a[p][k], ark = a[r][k];
                    a[p][k] = cs * apk - sn * ark;
                    a[r][k] = sn * apk + cs * ark;
                }
                for (int k = 0; k < 3; ++k) {
                    double qkp = q[k][p], qkr = q[k][r];
                    q[k][p] = cs * qkp - sn * qkr;
                    q[k][r] = sn * qkp + cs * qkr;
                }
            }
    }
    int order[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2 - i; ++j)
            if (a[order[j]][order[j]] > a[order[j + 1]][order[j + 1]]) {
                int t = order[j]; order[j] = order[j + 1]; order[j + 1] = t;
            }
    for (int i = 0; i < 3; ++i) {
        w[i] = a[order[i]][order[i]];
        if (v) for (int k = 0; k < 3; ++k) v[3 * i + k] = q[k][order[i]];
    }
}

/* the four reference columns from an explicit neighbor coordinate list */
static void row_from_points(const double *q, const double *nb, int64_t n, double *out)
{
    out[0] = (double)n; out[1] = 0; out[2] = 0; out[3] = 0;
    if (n == 0) return;
    double m[3] = {0, 0, 0};
    for (int64_t i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) m[a] += nb[3 * i + a];
    for (int a = 0; a < 3; ++a) m[a] /= (double)n;
    double d0 = q[0] - m[0], d1 = q[1] - m[1], d2 = q[2] - m[2];
    out[1] = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    if (n < 2) return;
    double c[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = 0; i < n; ++i) {
        double x = nb[3 * i] - m[0], y = nb[3 * i + 1] - m[1], z = nb[3 * i + 2] - m[2];
        c[0] += x * x; c[1] += x * y; c[2] += x * z; c[3] += y * y; c[4] += y * z; c[5] += z * z;
    }
    for (int k = 0; k < 6; ++k) c[k] /= (double)(n - 1);
    double w[3];
    eig3(c, w, NULL);
    double total = w[0] + w[1] + w[2];
    if (total != 0) { out[2] = w[2] / total; out[3] = w[1] / total; }
}

/* ------------------------------------------------------------------ a5: one query's ball */
/* visits the lattice window around q and calls back for every member; returns count.
 * idx/nbuf may be NULL.  members are produced in ascending address order (z, y, x). */
static int64_t ball_members(const orc_grid *g, const orc_hash *h, const double *q, double radius,
                            int64_t *idx, double *nbuf, int64_t cap)
{
    const double r2 = radius * radius;
    int64_t lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        int64_t lim = g->widths[a] >= 63 ? INT64_MAX : ((1LL << g->widths[a]) - 1);
        double l = floor((q[a] - radius - g->minc[a]) / g->edge) - 1;
        double u = floor((q[a] + radius - g->minc[a]) / g->edge) + 1;
        if (u < 0 || l > (double)lim) return 0;
        lo[a] = l < 0 ? 0 : (int64_t)l;
        hi[a] = u > (double)lim ? lim : (int64_t)u;
    }
    int64_t count = 0;
    for (int64_t kz = lo[2]; kz <= hi[2]; ++kz) {
        double dz = q[2] - centre_of(g, kz, 2);
        for (int64_t ky = lo[1]; ky <= hi[1]; ++ky) {
            double dy = q[1] - centre_of(g, ky, 1);
            for (int64_t kx = lo[0]; kx <= hi[0]; ++kx) {
                double dx = q[0] - centre_of(g, kx, 0);
                double s = dx * dx;
                s += dy * dy;
                s += dz * dz;
                if (!(s <= r2)) continue;
                int64_t id = hash_find(h, kx + (ky << g->shifts[1]) + (kz << g->shifts[2]));
                if (id < 0) continue;
                if (count < cap) {
                    if (idx) idx[count] = id;
                    if (nbuf) {
                        nbuf[3 * count] = centre_of(g, kx, 0);
                        nbuf[3 * count + 1] = centre_of(g, ky, 1);
                        nbuf[3 * count + 2] = centre_of(g, kz, 2);
                    }
                }
                ++count;
            }
        }
    }
    return count;
}

/* ------------------------------------------------------------------ a5-a10 */
typedef struct {
    const orc_grid *g; const orc_hash *h; const double *query; double radius; double *out; int fail;
} feat_ctx;

static void feat_body(int64_t begin, int64_t end, void *p)
{
    feat_ctx *c = (feat_ctx *)p;
    int64_t cap = 4096;
    double *buf = (double *)malloc(sizeof(double) * 3 * cap);
    for (int64_t i = begin; i < end && buf; ++i) {
        int64_t n = ball_members(c->g, c->h, c->query + 3 * i, c->radius, NULL, buf, cap);
        if (n > cap) {
            cap = n * 2;
            free(buf);
            buf = (double *)malloc(sizeof(double) * 3 * cap);
            if (!buf) break;
            n = ball_members(c->g, c->h, c->query + 3 * i, c->radius, NULL, buf, cap);
        }
        row_from_points(c->query + 3 * i, buf, n, c->out + 4 * i);
    }
    if (!buf) c->fail = 1;
    free(buf);
}

typedef struct {
    const orc_grid *g; const orc_hash *h; const double *query; double radius;
    int64_t *offsets; int64_t *indices;
} sets_ctx;

static void sets_count_body(int64_t begin, int64_t end, void *p)
{
    sets_ctx *c = (sets_ctx *)p;
    for (int64_t i = begin; i < end; ++i)
        c->offsets[i + 1] = ball_members(c->g, c->h, c->query + 3 * i, c->radius, NULL, NULL, 0);
}

static void sets_fill_body(int64_t begin, int64_t end, void *p)
{
    sets_ctx *c = (sets_ctx *)p;
    for (int64_t i = begin; i < end; ++i)
        ball_members(c->g, c->h, c->query + 3 * i, c->radius, c->indices + c->offsets[i], NULL,
                     c->offsets[i + 1] - c->offsets[i]);
}

/* (nq,4) features for one scale given the sorted unique addresses of the search voxels */
int orc_radius_features(const double *query, int64_t nq, const int64_t *ukeys, int64_t nv,
                        const double *minc, double edge, const int64_t *widths, double radius,
                        double *out)
{
    orc_grid g;
    grid_init(&g, minc, edge, widths);
    orc_hash h;
    if (hash_build(&h, ukeys, nv)) return 1;
    feat_ctx ctx = {&g, &h, query, radius, out, 0};
    orc_parallel_for(nq, 256, feat_body, &ctx);
    hash_free(&h);
    return ctx.fail;
}

/* CSR neighbor sets.  call with indices == NULL to get counts (offsets[i+1]-offsets[i]); then again
 * with a buffer of offsets[nq] entries. */
int orc_radius_sets(const double *query, int64_t nq, const int64_t *ukeys, int64_t nv,
                    const double *minc, double edge, const int64_t *widths, double radius,
                    int64_t *offsets, int64_t *indices)
{
    orc_grid g;
    grid_init(&g, minc, edge, widths);
    orc_hash h;
    if (hash_build(&h, ukeys, nv)) return 1;
    sets_ctx ctx = {&g, &h, query, radius, offsets, indices};
    if (!indices) {
        offsets[0] = 0;
        orc_parallel_for(nq, 256, sets_count_body, &ctx);
        for (int64_t i = 0; i < nq; ++i) offsets[i + 1] += offsets[i];
    } else {
        orc_parallel_for(nq, 256, sets_fill_body, &ctx);
    }
    hash_free(&h);
    return 0;
}

/* ------------------------------------------------------------------ a11 */
int orc_process(const double *query, int64_t nq, const double *search, int64_t ns,
                const double *edges, const double *radii, int32_t n_scales, double *out)
{
    int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * (size_t)ns);
    double *block = (double *)malloc(sizeof(double) * 4 * (size_t)(nq > 0 ? nq : 1));
    if (!keys || !block) return 1;
    int rc = 0;
    for (int s = 0; s < n_scales && !rc; ++s) {
        double minc[3], maxc[3];
        int64_t widths[3];
        rc = orc_grid_widths(search, ns, edges[s], minc, maxc, widths);
        if (rc) break;
        int64_t nv = orc_unique_voxels(search, ns, minc, edges[s], widths, keys, NULL);
        rc = orc_radius_features(query, nq, keys, nv, minc, edges[s], widths, radii[s], block);
        for (int64_t i = 0; i < nq; ++i)
            memcpy(out + (size_t)i * 4 * n_scales + 4 * s, block + 4 * i, 4 * sizeof(double));
    }
    free(keys);
    free(block);
    return rc;
}

/* ------------------------------------------------------------------ kNN over an explicit point set */
typedef struct { double d2; int64_t idx; } orc_cand;

static inline int cand_less(const orc_cand *a, const orc_cand *b)
{
    return a->d2 < b->d2 || (a->d2 == b->d2 && a->idx < b->idx);
}

/* brute force, exact total order (d^2, index); idx_out/d2_out are (nq,k), padded with -1/inf */
typedef struct {
    const double *query; const double *pts; int64_t np; int32_t k; int64_t *idx_out; double *d2_out;
} knn_ctx;

static void knn_body(int64_t begin, int64_t end, void *p)
{
    knn_ctx *c = (knn_ctx *)p;
    const double *query = c->query, *pts = c->pts;
    const int64_t np = c->np;
    const int32_t k = c->k;
    int64_t *idx_out = c->idx_out;
    double *d2_out = c->d2_out;
    {
        orc_cand *best = (orc_cand *)malloc(sizeof(orc_cand) * (size_t)(k + 1));
        for (int64_t i = begin; i < end; ++i) {
            int have = 0;
            const double *q = query + 3 * i;
            for (int64_t j = 0; j < np; ++j) {
                double dx = q[0] - pts[3 * j], dy = q[1] - pts[3 * j + 1], dz = q[2] - pts[3 * j + 2];
                double s = dx * dx;
                s += dy * dy;
                s += dz * dz;
                orc_cand c = {s, j};
                if (have == k && !cand_less(&c, &best[k - 1])) continue;
                int pos = have < k ? have : k - 1;
                while (pos > 0 && cand_less(&c, &best[pos - 1])) { best[pos] = best[pos - 1]; --pos; }
                best[pos] = c;
                if (have < k) ++have;
            }
            for (int j = 0; j < k; ++j) {
                idx_out[i * k + j] = j < have ? best[j].idx : -1;
                d2_out[i * k + j] = j < have ? best[j].d2 : INFINITY;
            }
        }
        free(best);
    }
}

int orc_knn(const double *query, int64_t nq, const double *pts, int64_t np, int32_t k,
            int64_t *idx_out, double *d2_out)
{
    knn_ctx ctx = {query, pts, np, k, idx_out, d2_out};
    orc_parallel_for(nq, 64, knn_body, &ctx);
    return 0;
}

/* four reference columns over the first k_s entries of each kNN row, for each k_s */
int orc_knn_features(const double *query, int64_t nq, const double *pts, const int64_t *knn_idx,
                     int32_t kmax, const int32_t *ks, int32_t n_k, double *out)
{
    double *buf = (double *)malloc(sizeof(double) * 3 * (size_t)kmax);
    if (!buf) return 1;
    for (int64_t i = 0; i < nq; ++i)
        for (int s = 0; s < n_k; ++s) {
            int64_t n = 0;
            for (int j = 0; j < ks[s] && j < kmax; ++j) {
                int64_t id = knn_idx[i * kmax + j];
                if (id < 0) break;
                memcpy(buf + 3 * n, pts + 3 * id, 3 * sizeof(double));
                ++n;
            }
            row_from_points(query + 3 * i, buf, n, out + (size_t)i * 4 * n_k + 4 * s);
        }
    free(buf);
    return 0;
}

int orc_num_threads(void) { return g_threads; }

void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
