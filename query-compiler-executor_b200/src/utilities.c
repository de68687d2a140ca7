/* utilities.c -- loader, mid-result lookup, projection and the query loop.
 *
 * Same entry points as /root/reference/src/utilities.c (read_relations :124,
 * relation_exists :164, relation_exists_current :183, execute_queries :289);
 * what they do with the data is different:
 *   - read_relations maps each relation file and uploads its columns to HBM
 *     as they are on disk (column-major uint64).  The reference's fill_data
 *     (:105-121) rewrites every column into a 16-byte AoS tuple array with
 *     payload = row index; here the row id is implicit and `tuples` stays NULL.
 *   - print_sums (:197-224) becomes one gather-and-reduce kernel per projected
 *     binding (qce_checksum sums every selected column of a binding in a single
 *     read of its row-id column); only the formatting is host code.
 */
#define _GNU_SOURCE
#include "utilities.h"

#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/qce_b200.h"
#include "filter.h"
#include "join.h"
#include "pred_arrange.h"
#include "schedule.h"

/* ------------------------------------------------------------------ loader */
/* Relation files stay mapped; a column goes to HBM when a batch first references it
 * (the reference copies every column of every relation, src/utilities.c:105-121).
 * Nothing here touches CUDA, so the host layer may still fork one process per GPU
 * (QCE_GPUS, schedule.c) after the relations have been read. */
typedef struct rel_file {
    uint64_t *map;
    size_t bytes;
    uint64_t rows, cols;
    unsigned char *uploaded; /* per column */
} rel_file;
static rel_file *g_files = NULL;
static size_t g_nfiles = 0, g_files_cap = 0;

static int load_relation_file(const char *path, uint32_t rel_index, metadata *out)
{
    int rc = -1;
    uint64_t *map = MAP_FAILED;
    struct stat sb;
    out->data = NULL;
    int fd = open(path, O_RDONLY);
    check(fd != -1, "open failed");
    check(fstat(fd, &sb) != -1, "fstat failed");
    check((size_t)sb.st_size >= 2 * sizeof(uint64_t), "relation file too short");
    map = (uint64_t *)mmap(NULL, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    check(map != MAP_FAILED, "mmap failed");

    out->tuples = map[0];
    out->columns = map[1];
    /* the header is untrusted: no wrap-around in rows x columns, row ids are 32-bit on the device */
    uint64_t cells = 0;
    check(out->tuples < (1ull << 32) && out->columns <= 65536 && !__builtin_mul_overflow(out->tuples, out->columns, &cells),
          "implausible relation header (%lu rows x %lu columns)", (unsigned long)out->tuples, (unsigned long)out->columns);
    check(cells <= ((uint64_t)sb.st_size - 2 * sizeof(uint64_t)) / sizeof(uint64_t), "relation file truncated");
    out->data = MALLOC(relation *, out->columns ? out->columns : 1);
    check_mem(out->data);
    for (uint64_t c = 0; c < out->columns; c++) {
        relation *col = MALLOC(relation, 1);
        check_mem(col);
        col->num_tuples = out->tuples;
        col->tuples = NULL; /* resident in HBM once referenced; the driver's FREE(tuples) is NULL-safe */
        out->data[c] = col;
    }
    if (rel_index >= g_files_cap) {
        size_t cap = g_files_cap ? g_files_cap * 2 : 16;
        while (cap <= rel_index) cap *= 2;
        rel_file *grown = (rel_file *)realloc(g_files, cap * sizeof(rel_file));
        check_mem(grown);
        memset(grown + g_files_cap, 0, (cap - g_files_cap) * sizeof(rel_file));
        g_files = grown;
        g_files_cap = cap;
    }
    rel_file *f = &g_files[rel_index];
    if (f->map) { munmap(f->map, f->bytes); free(f->uploaded); }
    f->map = map;
    f->bytes = (size_t)sb.st_size;
    f->rows = out->tuples;
    f->cols = out->columns;
    f->uploaded = (unsigned char *)calloc(out->columns ? out->columns : 1, 1);
    check_mem(f->uploaded);
    if (rel_index + 1 > g_nfiles) g_nfiles = rel_index + 1;
    map = MAP_FAILED; /* kept */
    rc = 0;

error:
    if (map != MAP_FAILED) munmap(map, (size_t)sb.st_size);
    if (fd != -1) close(fd);
    return rc;
}

int read_relations_from(FILE *in, DArray *metadata_arr)
{
    char *line = NULL;
    size_t cap = 0;
    int rc = 0;
    while (getline(&line, &cap, in) != -1) {
        if (strncmp(line, "Done\n", 5) == 0 || strncmp(line, "done\n", 5) == 0) break;
        size_t len = strlen(line);
        if (len && line[len - 1] == '\n') line[len - 1] = '\0';
        metadata m;
        if (load_relation_file(line, (uint32_t)DArray_count(metadata_arr), &m) != 0) {
            rc = -1;
            break;
        }
        DArray_push(metadata_arr, &m);
    }
    free(line);
    return rc;
}

int read_relations(DArray *metadata_arr) { return read_relations_from(stdin, metadata_arr); }

/* relation files known to the loader (0 when the relations were uploaded by an embedder) */
size_t qce_host_file_count(void) { return g_nfiles; }
uint64_t qce_host_file_rows(uint32_t rel) { return rel < g_nfiles ? g_files[rel].rows : 0; }
/* 1: (rel, col) exists in a mapped file and is not in HBM yet */
int qce_host_column_pending(uint32_t rel, uint32_t col)
{
    return rel < g_nfiles && g_files[rel].map && col < g_files[rel].cols && !g_files[rel].uploaded[col];
}
int qce_host_upload_column(uint32_t rel, uint32_t col)
{
    if (!qce_host_column_pending(rel, col)) return 0;
    rel_file *f = &g_files[rel];
    if (qce_upload_column(rel, col, f->map + 2 + (uint64_t)col * f->rows, f->rows) != 0) {
        log_err("column upload failed: %s", qce_last_error());
        return -1;
    }
    f->uploaded[col] = 1;
    return 0;
}

/* ------------------------------------------------------------------ lookups */
exists_info relation_exists(DArray *mid_results_array, uint64_t relation, uint64_t predicate_id)
{
    exists_info found = {0, -1};
    for (ssize_t e = (ssize_t)DArray_count(mid_results_array) - 1; e >= 0; e--) {
        DArray *entity = *(DArray **)DArray_get(mid_results_array, e);
        for (ssize_t j = 0; j < (ssize_t)DArray_count(entity); j++) {
            const mid_result *m = (const mid_result *)DArray_get(entity, j);
            if (m->relation == relation && m->predicate_id == predicate_id) {
                found.mid_result = e;
                found.index = j;
                return found;
            }
        }
    }
    return found;
}

ssize_t relation_exists_current(DArray *mid_results, uint64_t relation, uint64_t predicate_id)
{
    ssize_t last_hit = -1;
    for (ssize_t j = 0; j < (ssize_t)DArray_count(mid_results); j++) {
        const mid_result *m = (const mid_result *)DArray_get(mid_results, j);
        if (m->relation == relation && m->predicate_id == predicate_id) last_hit = j;
    }
    return last_hit;
}

/* ------------------------------------------------------------------ projection */
#define MAX_SELECTS 64

/* One checksum token per select, in select order: "<sum> " or "NULL " when the
 * binding's row-id column is empty; newline after the last one. */
static int print_sums(DArray *entities, const query *q, FILE *out)
{
    uint64_t sums[MAX_SELECTS];
    int is_null[MAX_SELECTS], done[MAX_SELECTS];
    mid_result *entry[MAX_SELECTS];
    size_t ns = q->select_size;
    if (ns > MAX_SELECTS) {
        log_err("too many selects");
        return -1;
    }
    /* The reference resolves and prints select by select and exits at the first
     * binding without a mid result (src/utilities.c:203-207), i.e. AFTER the earlier
     * tokens were written: `limit` keeps that observable behaviour. */
    size_t limit = ns;
    for (size_t i = 0; i < ns; i++) {
        const uint64_t binding = q->selects[i].relation;
        exists_info where = relation_exists(entities, q->relations[binding], binding);
        if (where.index == -1) {
            limit = i;
            break;
        }
        DArray *entity = *(DArray **)DArray_get(entities, where.mid_result);
        entry[i] = (mid_result *)DArray_get(entity, where.index);
        is_null[i] = qce_rowids_count(entry[i]->payloads) == 0;
        done[i] = is_null[i];
    }
    const size_t ns_all = ns;
    ns = limit;
    /* every selected column of one binding in a single pass over its row ids */
    for (size_t i = 0; i < ns; i++) {
        if (done[i]) continue;
        uint32_t cols[8];
        size_t slot[8];
        uint32_t k = 0;
        for (size_t j = i; j < ns && k < 8; j++) {
            if (!done[j] && entry[j] == entry[i]) {
                cols[k] = (uint32_t)q->selects[j].column;
                slot[k++] = j;
            }
        }
        uint64_t part[8];
        if (qce_checksum(entry[i]->payloads, (uint32_t)entry[i]->relation, cols, k, part) != 0) {
            log_err("checksum failed: %s", qce_last_error());
            return -1;
        }
        for (uint32_t t = 0; t < k; t++) {
            sums[slot[t]] = part[t];
            done[slot[t]] = 1;
        }
    }
    for (size_t i = 0; i < ns; i++) {
        if (is_null[i]) fputs("NULL ", out);
        else fprintf(out, "%lu ", (unsigned long)sums[i]);
    }
    if (limit < ns_all) {
        log_err("Something went really wrong...");
        fflush(out);
        qce_fatal(); /* src/utilities.c:204-207: exit(EXIT_FAILURE) once the earlier queries' lines are out */
    }
    fputc('\n', out);
    return 0;
}

/* ------------------------------------------------------------------ query loop */
static void destroy_entities(DArray *entities)
{
    for (size_t e = 0; e < DArray_count(entities); e++) {
        DArray *entity = *(DArray **)DArray_get(entities, e);
        for (size_t j = 0; j < DArray_count(entity); j++) {
            mid_result *m = (mid_result *)DArray_get(entity, j);
            qce_rowids_free(m->payloads);
        }
        DArray_destroy(entity);
    }
    DArray_destroy(entities);
}

int execute_query_to(query *q, DArray *metadata_arr, FILE *out)
{
    int rc = -1;
    DArray *entities = DArray_create(sizeof(DArray *), 2);
    if (entities == NULL) return -1;
    qce_set_query_stdout(out ? out : stdout);

    for (size_t i = 0; i < q->predicates_size; i++) {
        predicate *p = &q->predicates[i];
        const int r = p->type == 1 ? execute_filter(p, q->relations, metadata_arr, entities)
                                   : execute_join(p, q->relations, metadata_arr, entities);
        if (qce_join_elide_unsafe()) goto error; /* elided attempt withdrawn: the scheduler replays the query */
        if (p->type == 1) {
            check(r != -1, "Filter failed!");
        } else {
            check(r != -1, "Join failed!");
        }
    }
    check(print_sums(entities, q, out ? out : stdout) == 0, "Projection failed!");
    rc = 0;

error:
    destroy_entities(entities);
    qce_set_query_stdout(NULL);
    return rc;
}

void execute_queries(DArray *q_list, DArray *metadata_arr)
{
    /* the batch scheduler (schedule.c): lazy column loads, one process per GPU when QCE_GPUS > 1,
     * small queries overlapped on several streams, stdout in query order */
    const int status = qce_run_queries(q_list, metadata_arr, stdout, NULL);
    fflush(stdout);
    if (status == QCE_RUN_FATAL) exit(EXIT_FAILURE); /* src/utilities.c:204-207, src/join.c:563-620 */
}
