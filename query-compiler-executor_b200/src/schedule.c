/* schedule.c -- see schedule.h. */
#define _GNU_SOURCE
#include "schedule.h"

#include <pthread.h>
#include <setjmp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/qce_b200.h"
#include "dbg.h"
#include "join.h"
#include "pred_arrange.h"
#include "structs.h"
#include "utilities.h"

/* loader side of utilities.c */
size_t qce_host_file_count(void);
uint64_t qce_host_file_rows(uint32_t rel);
int qce_host_column_pending(uint32_t rel, uint32_t col);
int qce_host_upload_column(uint32_t rel, uint32_t col);

#define MAX_STREAMS 16

typedef struct qjob {
    query *q;
    char *text; /* stdout bytes of the query (count lines + result line) */
    size_t len;
    int failed, fatal;
    int whole;    /* every relation it touches is held whole by every rank: may run on one rank alone */
    int heavy;    /* too large to share the device with other queries */
    int owner;    /* rank that runs it alone; -1 = all ranks together (sharded) */
    int ran;
    uint64_t cost; /* rows of the relations it binds */
} qjob;

typedef struct colref {
    uint32_t rel, col;
    int state; /* 0 pending, 1 resident, 2 failed */
} colref;

typedef struct batch {
    qjob *jobs;
    size_t njobs;
    DArray *meta;
    colref *cols;
    size_t ncols;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    /* light queries of this rank, pulled by the workers */
    size_t *light;
    size_t nlight, next_light;
} batch;

static __thread jmp_buf *tl_fatal_jmp = NULL;

void qce_fatal(void)
{
    if (tl_fatal_jmp) longjmp(*tl_fatal_jmp, 1);
    fflush(stdout);
    exit(EXIT_FAILURE);
}

static long env_long(const char *name, long dflt)
{
    const char *e = getenv(name);
    return e && *e ? atol(e) : dflt;
}

/* ------------------------------------------------------------------ columns of the batch */
static int find_col(batch *b, uint32_t rel, uint32_t col)
{
    for (size_t i = 0; i < b->ncols; i++)
        if (b->cols[i].rel == rel && b->cols[i].col == col) return (int)i;
    return -1;
}
static int note_col(batch *b, size_t *cap, uint32_t rel, uint32_t col)
{
    if (!qce_host_column_pending(rel, col) || find_col(b, rel, col) >= 0) return 0;
    if (b->ncols == *cap) {
        *cap = *cap ? *cap * 2 : 32;
        colref *g = (colref *)realloc(b->cols, *cap * sizeof(colref));
        if (!g) return -1;
        b->cols = g;
    }
    b->cols[b->ncols].rel = rel;
    b->cols[b->ncols].col = col;
    b->cols[b->ncols].state = 0;
    b->ncols++;
    return 0;
}
/* calls f(rel, col) for every column the query reads */
typedef int (*col_fn)(void *ctx, uint32_t rel, uint32_t col);
static int for_each_col(const query *q, col_fn f, void *ctx)
{
    for (size_t i = 0; i < q->predicates_size; i++) {
        const predicate *p = &q->predicates[i];
        if (p->first.relation < q->relations_size && f(ctx, q->relations[p->first.relation], (uint32_t)p->first.column) != 0) return -1;
        if (p->type == 0) {
            const relation_column *s = (const relation_column *)p->second;
            if (s->relation < q->relations_size && f(ctx, q->relations[s->relation], (uint32_t)s->column) != 0) return -1;
        }
    }
    for (size_t i = 0; i < q->select_size; i++)
        if (q->selects[i].relation < q->relations_size &&
            f(ctx, q->relations[q->selects[i].relation], (uint32_t)q->selects[i].column) != 0)
            return -1;
    return 0;
}
struct note_ctx { batch *b; size_t *cap; };
static int note_cb(void *ctx, uint32_t rel, uint32_t col)
{
    struct note_ctx *n = (struct note_ctx *)ctx;
    return note_col(n->b, n->cap, rel, col);
}
static int wait_cb(void *ctx, uint32_t rel, uint32_t col)
{
    batch *b = (batch *)ctx;
    const int i = find_col(b, rel, col);
    if (i < 0) return 0;
    pthread_mutex_lock(&b->mu);
    while (b->cols[i].state == 0) pthread_cond_wait(&b->cv, &b->mu);
    pthread_mutex_unlock(&b->mu);
    return 0; /* a failed load surfaces as "column was never uploaded" in the operator */
}
static void mark_col(batch *b, size_t i, int state)
{
    pthread_mutex_lock(&b->mu);
    b->cols[i].state = state;
    pthread_cond_broadcast(&b->cv);
    pthread_mutex_unlock(&b->mu);
}

/* ------------------------------------------------------------------ one query */
/* Shape test for elided mode (the parity-defined class): distinct relation ids, filters
 * on at most one binding, and at least two join predicates (otherwise no bystander ever exists). */
static int elidable(const query *q)
{
    size_t joins = 0;
    long fbind = -1;
    for (size_t i = 0; i < q->relations_size; i++)
        for (size_t k = i + 1; k < q->relations_size; k++)
            if (q->relations[i] == q->relations[k]) return 0;
    for (size_t i = 0; i < q->predicates_size; i++) {
        const predicate *p = &q->predicates[i];
        if (p->type == 1) {
            if (fbind >= 0 && (long)p->first.relation != fbind) return 0;
            fbind = (long)p->first.relation;
        } else {
            joins++;
        }
    }
    return joins >= 2;
}

static int g_trace_jobs = -1;
static void run_job(batch *b, qjob *j)
{
    if (g_trace_jobs < 0) g_trace_jobs = getenv("QCE_TRACE_JOBS") != NULL;
    if (g_trace_jobs) fprintf(stderr, "[qce] rank %u thread %lx: query %ld starts (cost %lu, heavy %d, owner %d)\n", qce_comm_rank(),
                              (unsigned long)pthread_self(), (long)(j - b->jobs), (unsigned long)j->cost, j->heavy, j->owner);
    for_each_col(j->q, wait_cb, b);
    /* first with the bystander re-joins elided (join.c) when the query's shape allows it; withdrawn
     * attempts are replayed faithfully, their output discarded */
    for (int attempt = elidable(j->q) && qce_elision_supported() ? 0 : 1; attempt < 2; attempt++) {
        free(j->text);
        j->text = NULL;
        j->len = 0;
        j->failed = j->fatal = 0;
        FILE *mem = open_memstream(&j->text, &j->len);
        if (mem == NULL) { j->failed = 1; return; }
        jmp_buf jb;
        volatile int withdrawn = 0;
        tl_fatal_jmp = &jb;
        qce_join_elide_begin(attempt == 0);
        if (setjmp(jb) == 0) {
            if (execute_query_to(j->q, b->meta, mem) != 0) j->failed = 1;
            withdrawn = qce_join_elide_unsafe();
        } else {
            j->fatal = 1; /* the reference exits here; what the query printed so far stays */
            withdrawn = attempt == 0; /* its partial output must be the faithful path's */
        }
        tl_fatal_jmp = NULL;
        qce_join_elide_begin(0);
        fclose(mem);
        if (!withdrawn) break;
        if (g_trace_jobs) fprintf(stderr, "[qce] query %ld: elided attempt withdrawn, replaying\n", (long)(j - b->jobs));
    }
    j->ran = 1;
    if (g_trace_jobs) fprintf(stderr, "[qce] rank %u thread %lx: query %ld done\n", qce_comm_rank(), (unsigned long)pthread_self(), (long)(j - b->jobs));
}

/* ------------------------------------------------------------------ threads */
static void *g_stream_ctx[MAX_STREAMS];
static void *g_loader_ctx = NULL;

typedef struct worker_arg { batch *b; int slot; } worker_arg;
static void *worker_main(void *p)
{
    worker_arg *w = (worker_arg *)p;
    batch *b = w->b;
    if (qce_ctx_bind(g_stream_ctx[w->slot]) != 0) return NULL;
    for (;;) {
        pthread_mutex_lock(&b->mu);
        const size_t k = b->next_light < b->nlight ? b->next_light++ : (size_t)-1;
        pthread_mutex_unlock(&b->mu);
        if (k == (size_t)-1) break;
        run_job(b, &b->jobs[b->light[k]]);
    }
    qce_ctx_bind(NULL);
    return NULL;
}
static void *loader_main(void *p)
{
    batch *b = (batch *)p;
    if (qce_ctx_bind(g_loader_ctx) != 0) {
        for (size_t i = 0; i < b->ncols; i++)
            if (b->cols[i].state == 0) mark_col(b, i, 2);
        return NULL;
    }
    for (size_t i = 0; i < b->ncols; i++) {
        if (b->cols[i].state != 0) continue;
        mark_col(b, i, qce_host_upload_column(b->cols[i].rel, b->cols[i].col) == 0 ? 1 : 2);
    }
    qce_ctx_bind(NULL);
    return NULL;
}

/* ------------------------------------------------------------------ results between ranks */
/* [u32 index][u8 failed][u8 fatal][u32 len][bytes] per query this rank ran alone */
static char *pack_results(const batch *b, int rank, size_t *bytes)
{
    size_t total = 0;
    for (size_t i = 0; i < b->njobs; i++)
        if (b->jobs[i].owner == rank && b->jobs[i].ran) total += 10 + b->jobs[i].len;
    char *buf = (char *)malloc(total ? total : 1), *at = buf;
    for (size_t i = 0; buf && i < b->njobs; i++) {
        const qjob *j = &b->jobs[i];
        if (j->owner != rank || !j->ran) continue;
        uint32_t idx = (uint32_t)i, len = (uint32_t)j->len;
        memcpy(at, &idx, 4);
        at[4] = (char)j->failed;
        at[5] = (char)j->fatal;
        memcpy(at + 6, &len, 4);
        memcpy(at + 10, j->text, j->len);
        at += 10 + j->len;
    }
    *bytes = total;
    return buf;
}
static void unpack_results(batch *b, const char *buf, size_t bytes)
{
    size_t at = 0;
    while (at + 10 <= bytes) {
        uint32_t idx, len;
        memcpy(&idx, buf + at, 4);
        memcpy(&len, buf + at + 6, 4);
        if (idx >= b->njobs || at + 10 + len > bytes) break;
        qjob *j = &b->jobs[idx];
        if (!j->ran) {
            j->failed = buf[at + 4];
            j->fatal = buf[at + 5];
            j->text = (char *)malloc(len ? len : 1);
            if (j->text) memcpy(j->text, buf + at + 10, len);
            j->len = j->text ? len : 0;
            j->ran = 1;
        }
        at += 10 + len;
    }
}

/* ------------------------------------------------------------------ the batch */
static int whole_cb(void *ctx, uint32_t rel, uint32_t col)
{
    (void)col;
    int *whole = (int *)ctx;
    /* placement follows the relation's size: loaded files by their row count, columns an embedder
     * uploaded by what the engine holds */
    int w;
    if (qce_host_file_rows(rel) && qce_host_column_pending(rel, col)) w = qce_column_would_be_whole(qce_host_file_rows(rel));
    else w = qce_column_is_whole(rel, col);
    if (w == 0) *whole = 0;
    return 0;
}
static uint64_t rows_of(uint32_t rel)
{
    uint64_t n = qce_host_file_rows(rel);
    if (n == 0) qce_column_info(rel, 0, &n, NULL);
    return n;
}

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

int qce_run_queries(DArray *q_list, DArray *metadata_arr, FILE *out, int *failed)
{
    const int trace = getenv("QCE_TRACE") != NULL;
    const double t_enter = now_ms();
    const size_t n = DArray_count(q_list);
    int status = QCE_RUN_OK, bad = 0;
    batch b;
    memset(&b, 0, sizeof b);
    b.meta = metadata_arr;
    b.njobs = n;
    pthread_mutex_init(&b.mu, NULL);
    pthread_cond_init(&b.cv, NULL);

    /* one process per GPU: forked before this process touches CUDA */
    const long want_gpus = env_long("QCE_GPUS", 1);
    int forked_here = 0;
    if (want_gpus > 1 && qce_comm_world() == 1) {
        if (out) fflush(out);
        const int r = qce_comm_fork((uint32_t)want_gpus);
        if (r < 0) log_warn("QCE_GPUS=%ld ignored: %s", want_gpus, qce_last_error());
        else forked_here = 1;
    }
    const int world = (int)qce_comm_world(), rank = (int)qce_comm_rank();

    b.jobs = (qjob *)calloc(n ? n : 1, sizeof(qjob));
    b.light = (size_t *)calloc(n ? n : 1, sizeof(size_t));
    if (!b.jobs || !b.light) { log_err("Out of memory."); status = QCE_RUN_FATAL; goto done; }

    /* arrange (decides the join kinds, src/pred_arrange.c:88-93) and collect the columns to load,
     * in the order the batch first reads them */
    size_t ccap = 0;
    struct note_ctx nc = {&b, &ccap};
    for (size_t i = 0; i < n; i++) {
        b.jobs[i].q = (query *)DArray_get(q_list, i);
        arrange_predicates(b.jobs[i].q);
        for_each_col(b.jobs[i].q, note_cb, &nc);
    }

    /* placement (several ranks): replicate what fits, smallest relations first; shard the rest by rows */
    if (world > 1 && b.ncols) {
        uint64_t *rows = (uint64_t *)calloc(b.ncols, sizeof(uint64_t)), cap = 0;
        for (size_t i = 0; rows && i < b.ncols; i++) rows[i] = qce_host_file_rows(b.cols[i].rel);
        if (rows && qce_placement_cap(rows, (uint32_t)b.ncols, &cap) == 0) qce_set_replicate_bytes(cap);
        free(rows);
    }
    /* row-sharded columns are loaded by all ranks together (their windows are mapped into every
     * peer: a collective), before anything runs; whole columns stream in behind the first queries */
    int later = 0;
    for (size_t i = 0; i < b.ncols; i++) {
        if (world > 1 && !qce_column_would_be_whole(qce_host_file_rows(b.cols[i].rel)))
            b.cols[i].state = qce_host_upload_column(b.cols[i].rel, b.cols[i].col) == 0 ? 1 : 2;
        else
            later++;
    }

    /* who runs what */
    const uint64_t heavy_rows = (uint64_t)env_long("QCE_HEAVY_ROWS", 2000000);
    uint64_t *load = (uint64_t *)calloc((size_t)world, sizeof(uint64_t));
    for (size_t i = 0; i < n; i++) {
        qjob *j = &b.jobs[i];
        j->whole = 1;
        for_each_col(j->q, whole_cb, &j->whole);
        for (size_t k = 0; k < j->q->relations_size; k++) j->cost += rows_of(j->q->relations[k]);
        j->heavy = j->cost > heavy_rows;
        if (world == 1) j->owner = 0;
        else if (!j->whole) j->owner = -1;
    }
    if (world > 1 && load) {
        /* a query over replicated relations may still be too large for one rank's fair share of the
         * batch (a lone 100M-row join, an outlier among small queries): it runs sharded like the
         * queries over row-sharded relations -- every rank takes its row share of the whole columns */
        uint64_t total = 0;
        for (size_t i = 0; i < n; i++)
            if (b.jobs[i].whole) total += b.jobs[i].cost;
        for (size_t i = 0; i < n; i++)
            if (b.jobs[i].whole && b.jobs[i].heavy && b.jobs[i].cost > total / (uint64_t)world) {
                b.jobs[i].whole = 0;
                b.jobs[i].owner = -1;
            }
        /* replicas: longest queries first, each to the least loaded rank (the same on every rank) */
        char *assigned = (char *)calloc(n ? n : 1, 1);
        for (;;) {
            size_t pick = (size_t)-1;
            for (size_t i = 0; assigned && i < n; i++)
                if (b.jobs[i].whole && !assigned[i] && (pick == (size_t)-1 || b.jobs[i].cost > b.jobs[pick].cost)) pick = i;
            if (pick == (size_t)-1) break;
            int best = 0;
            for (int r = 1; r < world; r++)
                if (load[r] < load[best]) best = r;
            b.jobs[pick].owner = best;
            assigned[pick] = 1;
            load[best] += b.jobs[pick].cost + 1;
        }
        free(assigned);
    }
    free(load);
    for (size_t i = 0; i < n; i++)
        if (b.jobs[i].owner == rank && !b.jobs[i].heavy) b.light[b.nlight++] = i;

    /* threads: the loader and the streams for this rank's light queries */
    pthread_t loader, workers[MAX_STREAMS];
    int have_loader = 0, nworkers = 0;
    worker_arg wargs[MAX_STREAMS];
    if (later > 0) {
        if (!g_loader_ctx) g_loader_ctx = qce_ctx_create();
        if (g_loader_ctx && pthread_create(&loader, NULL, loader_main, &b) == 0) have_loader = 1;
        else loader_main(&b);
    }
    long streams = env_long("QCE_STREAMS", 8);
    if (streams > MAX_STREAMS) streams = MAX_STREAMS;
    if (b.nlight < 2) streams = 0; /* nothing to overlap */
    if ((size_t)streams > b.nlight) streams = (long)b.nlight;
    const double t_plan = now_ms();
    qce_batch_begin();
    const double t_begin = now_ms();
    for (long s = 0; s < streams; s++) {
        if (!g_stream_ctx[s]) g_stream_ctx[s] = qce_ctx_create();
        if (!g_stream_ctx[s]) break;
        wargs[nworkers].b = &b;
        wargs[nworkers].slot = (int)s;
        if (pthread_create(&workers[nworkers], NULL, worker_main, &wargs[nworkers]) != 0) break;
        nworkers++;
    }

    /* this thread: first the sharded queries, all ranks in step, then this rank's own heavy ones */
    for (size_t i = 0; i < n; i++)
        if (b.jobs[i].owner == -1) run_job(&b, &b.jobs[i]);
    if (world > 1) qce_ctx_solo(1);
    for (size_t i = 0; i < n; i++)
        if (b.jobs[i].owner == rank && b.jobs[i].heavy) run_job(&b, &b.jobs[i]);
    if (nworkers == 0) /* no stream could be created (or a single light query): run them here */
        for (size_t k = 0; k < b.nlight; k++) run_job(&b, &b.jobs[b.light[k]]);
    if (world > 1) qce_ctx_solo(0);
    for (int w = 0; w < nworkers; w++) pthread_join(workers[w], NULL);
    if (have_loader) pthread_join(loader, NULL);
    const double t_ran = now_ms();
    qce_batch_end();
    if (trace)
        fprintf(stderr, "[qce] batch of %zu: plan %.3f ms, batch_begin %.3f, run %.3f, batch_end %.3f (rank %d, %d streams)\n", n,
                t_plan - t_enter, t_begin - t_plan, t_ran - t_begin, now_ms() - t_ran, rank, nworkers);

    /* the queries the other ranks ran alone */
    if (world > 1) {
        size_t bytes = 0;
        char *mine = pack_results(&b, rank, &bytes), *all = NULL;
        uint64_t lens[MAX_STREAMS];
        if (!mine || qce_comm_gatherv(mine, bytes, &all, lens) != 0) {
            log_err("collecting the ranks' results failed: %s", qce_last_error());
            status = QCE_RUN_FATAL;
        } else if (rank == 0 && all) {
            size_t at = 0;
            for (int r = 0; r < world; r++) {
                if (r != 0) unpack_results(&b, all + at, lens[r]);
                at += lens[r];
            }
        }
        free(mine);
        free(all);
    }

    /* stdout in query order; a reference exit(EXIT_FAILURE) ends it after that query's partial output */
    for (size_t i = 0; i < n; i++) {
        qjob *j = &b.jobs[i];
        /* forked children share the parent's stdout: only rank 0 writes it */
        if (out && j->text && j->len && !qce_comm_is_child()) fwrite(j->text, 1, j->len, out);
        if (j->failed || (!j->ran && rank == 0)) bad++;
        if (j->fatal) { status = QCE_RUN_FATAL; break; }
    }

done:
    for (size_t i = 0; b.jobs && i < n; i++) free(b.jobs[i].text);
    free(b.jobs);
    free(b.light);
    free(b.cols);
    pthread_mutex_destroy(&b.mu);
    pthread_cond_destroy(&b.cv);
    if (failed) *failed = bad;
    if (forked_here) {
        if (out) fflush(out);
        /* children leave here; rank 0 collects them */
        const int lost = qce_comm_finish(status == QCE_RUN_OK ? 0 : 1);
        if (lost) log_err("%d of the forked ranks failed", lost);
    }
    return status;
}
