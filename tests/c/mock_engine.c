/* tests/c/mock_engine.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C stand-in for the subset of the C-ABI (include/qce_b200.h) that the host
 * operator layer (query-compiler-executor_b200/src/ *.c) calls, so that the host logic --
 * parser, arranger, the join-kind state machine of build_relations (src/join.c:152-292),
 * update_mid_results / fix_all_mid_results (:486-628), print_sums -- can be run and
 * fuzzed (ASan/UBSan) on a machine WITHOUT a GPU, against the reference's recorded
 * outputs in tests/golden.  It is linked only into tests/c/build/queries_mock by
 * tests/test_host_mock.py; the product (libqce_b200.so, build/queries) never sees it and
 * has no CPU path.  Semantics follow the comments of include/qce_b200.h and
 * oracle/qce_oracle.py, function by function.
 */
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/qce_b200.h"

struct qce_rowids {
    uint64_t *d;
    uint64_t n;
};
struct qce_tuples {
    uint64_t *k, *p;
    uint64_t n;
};

#define MAX_REL 64
#define MAX_COL 64
static struct {
    uint64_t *d;
    uint64_t n;
} g_col[MAX_REL][MAX_COL];
static __thread char g_err[256];

static int fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return -1;
}
static void *xalloc(uint64_t count, size_t size)
{
    void *p = calloc(count ? count : 1, size);
    if (!p) { fprintf(stderr, "mock engine: out of memory\n"); exit(2); }
    return p;
}
/* handles alive: the host layer must hand every one back (QCE_MOCK_BALANCE=1 reports at shutdown) */
static long g_live_ids, g_live_tuples;
static qce_rowids *new_ids(uint64_t n)
{
    __atomic_add_fetch(&g_live_ids, 1, __ATOMIC_RELAXED);
    qce_rowids *r = xalloc(1, sizeof *r);
    r->d = xalloc(n, sizeof(uint64_t));
    r->n = n;
    return r;
}
static qce_tuples *new_tuples(uint64_t n)
{
    __atomic_add_fetch(&g_live_tuples, 1, __ATOMIC_RELAXED);
    qce_tuples *t = xalloc(1, sizeof *t);
    t->k = xalloc(n, sizeof(uint64_t));
    t->p = xalloc(n, sizeof(uint64_t));
    t->n = n;
    return t;
}
static int column(uint32_t rel, uint32_t col, const uint64_t **d, uint64_t *n)
{
    if (rel >= MAX_REL || col >= MAX_COL || g_col[rel][col].d == NULL)
        return fail("relation %u column %u was never uploaded", rel, col);
    *d = g_col[rel][col].d;
    *n = g_col[rel][col].n;
    return 0;
}
static int cmp(uint64_t v, char op, uint64_t c, int *ok)
{
    switch (op) {
    case '=': *ok = v == c; return 0;
    case '>': *ok = v > c; return 0;
    case '<': *ok = v < c; return 0;
    }
    return fail("Wrong operator '%c'", op);
}

/* ---- contexts, batches, ranks: one thread-safe "device", one rank ---------------------------
 * The mock keeps no per-context state: handles are plain malloc blocks and the column table
 * is only written by uploads, so the scheduler's worker threads may call it concurrently. */
void *qce_ctx_create(void) { return xalloc(1, 8); }
int qce_ctx_bind(void *ctx) { (void)ctx; return 0; }
void qce_ctx_destroy(void *ctx) { free(ctx); }
int qce_ctx_solo(int on) { (void)on; return 0; }
int qce_batch_begin(void) { return 0; }
int qce_batch_end(void) { return 0; }
int qce_comm_fork(uint32_t world) { (void)world; return fail("the mock engine is a single rank"); }
uint32_t qce_comm_rank(void) { return 0; }
uint32_t qce_comm_world(void) { return 1; }
int qce_comm_is_child(void) { return 0; }
int qce_comm_finish(int status) { (void)status; return 0; }
int qce_comm_gatherv(const void *mine, uint64_t bytes, char **out, uint64_t *lens)
{
    if (lens) lens[0] = bytes;
    if (out) { *out = xalloc(bytes ? bytes : 1, 1); memcpy(*out, mine, bytes); }
    return 0;
}
int qce_placement_cap(const uint64_t *col_rows, uint32_t ncols, uint64_t *cap_bytes) { (void)col_rows; (void)ncols; *cap_bytes = ~0ull; return 0; }
int qce_set_replicate_bytes(uint64_t bytes) { (void)bytes; return 0; }
int qce_column_would_be_whole(uint64_t rows) { (void)rows; return 1; }
int qce_column_is_whole(uint32_t rel, uint32_t col)
{
    return (rel < MAX_REL && col < MAX_COL && g_col[rel][col].d != NULL) ? 1 : -1;
}
int qce_column_info(uint32_t rel, uint32_t col, uint64_t *n, uint64_t *max_value)
{
    const uint64_t *d;
    uint64_t rows;
    if (column(rel, col, &d, &rows) != 0) return -1;
    if (n) *n = rows;
    if (max_value) { *max_value = 0; for (uint64_t i = 0; i < rows; i++) if (d[i] > *max_value) *max_value = d[i]; }
    return 0;
}

int qce_init(int device) { (void)device; return 0; }
void qce_shutdown(void)
{
    if (getenv("QCE_MOCK_BALANCE"))
        fprintf(stderr, "mock engine: %ld row-id columns and %ld tuple runs were never freed\n", g_live_ids, g_live_tuples);
}
int qce_sync(void) { return 0; }
const char *qce_last_error(void) { return g_err; }
int qce_abi_version(void) { return 1; }

int qce_upload_column(uint32_t rel, uint32_t col, const uint64_t *host, uint64_t n)
{
    if (rel >= MAX_REL || col >= MAX_COL) return fail("mock engine holds %d x %d columns", MAX_REL, MAX_COL);
    free(g_col[rel][col].d);
    g_col[rel][col].d = xalloc(n, sizeof(uint64_t));
    memcpy(g_col[rel][col].d, host, n * sizeof(uint64_t));
    g_col[rel][col].n = n;
    return 0;
}

/* ---- filters (src/filter.c:3-64) */
int qce_filter_scan(uint32_t rel, uint32_t col, char op, uint64_t c, qce_rowids **out)
{
    const uint64_t *d;
    uint64_t n, m = 0;
    int ok;
    if (column(rel, col, &d, &n) != 0 || cmp(0, op, 0, &ok) != 0) return -1;
    qce_rowids *r = new_ids(n);
    for (uint64_t i = 0; i < n; i++) {
        cmp(d[i], op, c, &ok);
        if (ok) r->d[m++] = i;
    }
    r->n = m;
    *out = r;
    return 0;
}
int qce_filter_refine(qce_rowids *ids, uint32_t rel, uint32_t col, char op, uint64_t c, uint64_t *survivors)
{
    const uint64_t *d;
    uint64_t n, m = 0;
    int ok;
    if (!ids) return fail("null row-id column");
    if (column(rel, col, &d, &n) != 0 || cmp(0, op, 0, &ok) != 0) return -1;
    for (uint64_t i = 0; i < ids->n; i++) {
        if (ids->d[i] >= n) return fail("row id out of range");
        cmp(d[ids->d[i]], op, c, &ok);
        if (ok) ids->d[m++] = ids->d[i];
    }
    ids->n = m;
    if (survivors) *survivors = m;
    return 0;
}

/* ---- tuple runs (src/join.c:96-142) and the sort (src/join.c:5-94), stable here */
int qce_build_tuples_base(uint32_t rel, uint32_t col, qce_tuples **out)
{
    const uint64_t *d;
    uint64_t n;
    if (column(rel, col, &d, &n) != 0) return -1;
    qce_tuples *t = new_tuples(n);
    for (uint64_t i = 0; i < n; i++) { t->k[i] = d[i]; t->p[i] = i; }
    *out = t;
    return 0;
}
int qce_build_tuples_rowids(uint32_t rel, uint32_t col, const qce_rowids *ids, qce_tuples **out)
{
    const uint64_t *d;
    uint64_t n;
    if (!ids) return fail("null row-id column");
    if (column(rel, col, &d, &n) != 0) return -1;
    qce_tuples *t = new_tuples(ids->n);
    for (uint64_t i = 0; i < ids->n; i++) {
        if (ids->d[i] >= n) { qce_tuples_free(t); return fail("row id out of range"); }
        t->k[i] = d[ids->d[i]];
        t->p[i] = ids->d[i];
    }
    *out = t;
    return 0;
}
static void merge_sort(uint64_t *k, uint64_t *p, uint64_t *tk, uint64_t *tp, uint64_t lo, uint64_t hi)
{
    if (hi - lo < 2) return;
    const uint64_t mid = lo + (hi - lo) / 2;
    merge_sort(k, p, tk, tp, lo, mid);
    merge_sort(k, p, tk, tp, mid, hi);
    uint64_t a = lo, b = mid, o = lo;
    while (a < mid && b < hi) {
        if (k[b] < k[a]) { tk[o] = k[b]; tp[o++] = p[b++]; }
        else { tk[o] = k[a]; tp[o++] = p[a++]; }
    }
    while (a < mid) { tk[o] = k[a]; tp[o++] = p[a++]; }
    while (b < hi) { tk[o] = k[b]; tp[o++] = p[b++]; }
    memcpy(k + lo, tk + lo, (hi - lo) * sizeof(uint64_t));
    memcpy(p + lo, tp + lo, (hi - lo) * sizeof(uint64_t));
}
static void sort_kp(uint64_t *k, uint64_t *p, uint64_t n)
{
    uint64_t *tk = xalloc(n, sizeof(uint64_t)), *tp = xalloc(n, sizeof(uint64_t));
    merge_sort(k, p, tk, tp, 0, n);
    free(tk);
    free(tp);
}
int qce_sort_tuples(qce_tuples *t)
{
    if (!t) return fail("null tuple run");
    sort_kp(t->k, t->p, t->n);
    return 0;
}
int qce_tuples_is_sorted(const qce_tuples *t, int *sorted)
{
    if (!t || !sorted) return fail("null argument");
    *sorted = 1;
    for (uint64_t i = 1; i < t->n; i++)
        if (t->k[i - 1] > t->k[i]) { *sorted = 0; break; }
    return 0;
}
uint64_t qce_tuples_count(const qce_tuples *t) { return t ? t->n : 0; }
void qce_tuples_free(qce_tuples *t)
{
    if (!t) return;
    __atomic_sub_fetch(&g_live_tuples, 1, __ATOMIC_RELAXED);
    free(t->k);
    free(t->p);
    free(t);
}

/* ---- join_relations, src/join.c:325-392: the literal pointer walk (defined for unsorted input too) */
static void walk(const qce_tuples *R, const qce_tuples *S, qce_rowids **outR, qce_rowids **outS)
{
    uint64_t cap = 1024, m = 0;
    uint64_t *a = xalloc(cap, sizeof(uint64_t)), *b = xalloc(cap, sizeof(uint64_t));
    uint64_t pr = 0, s_start = 0;
    while (pr < R->n && s_start < S->n) {
        uint64_t ps = s_start;
        int flag = 0;
        while (ps < S->n) {
            if (R->k[pr] < S->k[ps]) break;
            if (R->k[pr] > S->k[ps]) {
                ps++;
                if (!flag) s_start = ps;
            } else {
                if (m == cap) {
                    cap *= 2;
                    a = realloc(a, cap * sizeof(uint64_t));
                    b = realloc(b, cap * sizeof(uint64_t));
                    if (!a || !b) { fprintf(stderr, "mock engine: out of memory\n"); exit(2); }
                }
                a[m] = R->p[pr];
                b[m++] = S->p[ps];
                flag = 1;
                ps++;
            }
        }
        pr++;
    }
    qce_rowids *r = xalloc(1, sizeof *r), *s = xalloc(1, sizeof *s);
    __atomic_add_fetch(&g_live_ids, 2, __ATOMIC_RELAXED);
    r->d = a; r->n = m;
    s->d = b; s->n = m;
    if (outR) *outR = r; else qce_rowids_free(r);
    if (outS) *outS = s; else qce_rowids_free(s);
}
/* ---- bystander re-join elision: positions through the merge, then gathers */
int qce_elision_supported(void) { return getenv("QCE_ELIDE") ? atoi(getenv("QCE_ELIDE")) : 1; }
int qce_build_tuples_positions(uint32_t rel, uint32_t col, const qce_rowids *ids, qce_tuples **out)
{
    if (qce_build_tuples_rowids(rel, col, ids, out) != 0) return -1;
    for (uint64_t i = 0; i < (*out)->n; i++) {
        if ((*out)->k[i] >> 32) { qce_tuples_free(*out); return fail("position-carrying runs need keys below 2^32"); }
        (*out)->p[i] = i;
    }
    return 0;
}
int qce_tuples_attach(qce_tuples *t, uint32_t ncols, const qce_rowids *const *cols)
{
    (void)t; (void)ncols; (void)cols; /* one rank: the positions index the caller's arrays */
    return 0;
}
int qce_merge_join_stats(const qce_tuples *R, const qce_tuples *S, qce_rowids **outR, qce_rowids **outS,
                         uint32_t *min_matches, uint32_t *max_matches)
{
    if (!R || !S || !outR || !outS || !min_matches || !max_matches) return fail("null argument");
    walk(R, S, outR, outS);
    uint32_t lo = 0xffffffffu, hi = 0;
    for (uint64_t i = 0; i < R->n; i++) {
        uint32_t c = 0;
        for (uint64_t j = 0; j < S->n; j++) c += S->k[j] == R->k[i]; /* tiny inputs only */
        if (c < lo) lo = c;
        if (c > hi) hi = c;
    }
    *min_matches = R->n && S->n ? lo : 0;
    *max_matches = R->n && S->n ? hi : 0;
    return 0;
}
int qce_rowids_gather(const qce_rowids *src, const qce_rowids *index, qce_rowids **out)
{
    if (!src || !index || !out) return fail("null argument");
    qce_rowids *g = new_ids(index->n);
    for (uint64_t i = 0; i < index->n; i++) {
        if (index->d[i] >= src->n) { qce_rowids_free(g); return fail("gather index out of range"); }
        g->d[i] = src->d[index->d[i]];
    }
    *out = g;
    return 0;
}

int qce_distinct_pairs(const qce_rowids *pairsR, const qce_rowids *pairsS, qce_rowids **distinctR, qce_rowids **distinctS)
{
    if (!pairsR || !pairsS || !distinctR || !distinctS) return fail("null argument");
    if (pairsR->n != pairsS->n) return fail("pair columns differ in length");
    const uint64_t n = pairsR->n;
    /* (r, s)-ascending: sort by s (stable), then by r (stable) */
    uint64_t *r = xalloc(n, sizeof(uint64_t)), *s = xalloc(n, sizeof(uint64_t));
    memcpy(r, pairsR->d, n * sizeof(uint64_t));
    memcpy(s, pairsS->d, n * sizeof(uint64_t));
    sort_kp(s, r, n);
    sort_kp(r, s, n);
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; i++)
        if (i == 0 || r[i] != r[i - 1] || s[i] != s[i - 1]) { r[m] = r[i]; s[m++] = s[i]; }
    qce_rowids *dr = xalloc(1, sizeof *dr), *ds = xalloc(1, sizeof *ds);
    __atomic_add_fetch(&g_live_ids, 2, __ATOMIC_RELAXED);
    dr->d = r; dr->n = m;
    ds->d = s; ds->n = m;
    *distinctR = dr;
    *distinctS = ds;
    return 0;
}
int qce_merge_join(const qce_tuples *R, const qce_tuples *S, qce_rowids **outR, qce_rowids **outS,
                   qce_rowids **distinctR, qce_rowids **distinctS)
{
    if (!R || !S || !outR || !outS) return fail("null argument");
    walk(R, S, outR, outS);
    if (distinctR && distinctS) return qce_distinct_pairs(*outR, *outS, distinctR, distinctS);
    return 0;
}
int qce_merge_join_walk(const qce_tuples *R, const qce_tuples *S, qce_rowids **outR, qce_rowids **outS)
{
    if (!R || !S || !outR || !outS) return fail("null argument");
    walk(R, S, outR, outS);
    return 0;
}

/* ---- scan_join, src/join.c:395-423 */
static int scan_impl(const uint64_t *cr, uint64_t nr, const qce_rowids *ir, const uint64_t *cs, uint64_t ns,
                     const qce_rowids *is, qce_rowids **outR, qce_rowids **outS)
{
    const uint64_t n_r = ir ? ir->n : nr, n_s = is ? is->n : ns, n = n_r < n_s ? n_r : n_s;
    qce_rowids *a = new_ids(n), *b = new_ids(n);
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t x = ir ? ir->d[i] : i, y = is ? is->d[i] : i;
        if (x >= nr || y >= ns) { qce_rowids_free(a); qce_rowids_free(b); return fail("row id out of range"); }
        if (cr[x] == cs[y]) { a->d[m] = x; b->d[m++] = y; }
    }
    a->n = b->n = m;
    *outR = a;
    *outS = b;
    return 0;
}
int qce_scan_join(uint32_t relR, uint32_t colR, const qce_rowids *idsR, uint32_t relS, uint32_t colS,
                  const qce_rowids *idsS, qce_rowids **outR, qce_rowids **outS)
{
    const uint64_t *cr, *cs;
    uint64_t nr, ns;
    if (!idsR || !idsS || !outR || !outS) return fail("null argument");
    if (column(relR, colR, &cr, &nr) != 0 || column(relS, colS, &cs, &ns) != 0) return -1;
    return scan_impl(cr, nr, idsR, cs, ns, idsS, outR, outS);
}
int qce_scan_join_base(uint32_t relR, uint32_t colR, uint32_t relS, uint32_t colS, qce_rowids **outR, qce_rowids **outS)
{
    const uint64_t *cr, *cs;
    uint64_t nr, ns;
    if (!outR || !outS) return fail("null argument");
    if (column(relR, colR, &cr, &nr) != 0 || column(relS, colS, &cs, &ns) != 0) return -1;
    return scan_impl(cr, nr, NULL, cs, ns, NULL, outR, outS);
}

/* ---- join_payloads, src/join.c:426-484 */
int qce_rejoin(const qce_rowids *driver, const qce_rowids *last, const qce_rowids *edit, qce_rowids **out)
{
    if (!driver || !last || !edit || !out) return fail("null argument");
    if (edit->n < last->n) return fail("bystander column shorter than the joined column (the reference reads past it)");
    qce_tuples R, S;
    R.n = last->n;
    R.k = xalloc(R.n, sizeof(uint64_t));
    R.p = xalloc(R.n, sizeof(uint64_t));
    memcpy(R.k, last->d, R.n * sizeof(uint64_t));
    memcpy(R.p, edit->d, R.n * sizeof(uint64_t));
    S.n = driver->n;
    S.k = xalloc(S.n, sizeof(uint64_t));
    S.p = xalloc(S.n, sizeof(uint64_t));
    memcpy(S.k, driver->d, S.n * sizeof(uint64_t));
    sort_kp(R.k, R.p, R.n);
    sort_kp(S.k, S.p, S.n);
    walk(&R, &S, out, NULL);
    free(R.k); free(R.p); free(S.k); free(S.p);
    return 0;
}

/* ---- print_sums inner loop, src/utilities.c:215-219 */
int qce_checksum(const qce_rowids *ids, uint32_t rel, const uint32_t *cols, uint32_t ncols, uint64_t *sums)
{
    if (!ids || !cols || !sums) return fail("null argument");
    for (uint32_t k = 0; k < ncols; k++) {
        const uint64_t *d;
        uint64_t n, s = 0;
        if (column(rel, cols[k], &d, &n) != 0) return -1;
        for (uint64_t i = 0; i < ids->n; i++) {
            if (ids->d[i] >= n) return fail("row id out of range");
            s += d[ids->d[i]];
        }
        sums[k] = s;
    }
    return 0;
}

uint64_t qce_rowids_count(const qce_rowids *ids) { return ids ? ids->n : 0; }
void qce_rowids_free(qce_rowids *ids)
{
    if (!ids) return;
    __atomic_sub_fetch(&g_live_ids, 1, __ATOMIC_RELAXED);
    free(ids->d);
    free(ids);
}
