/* queries_driver.c -- stdin -> stdout driver of the B200 engine.
 *
 * Speaks the reference's protocol (relation paths, `Done`, query lines to EOF;
 * /root/reference/main/queries_main.c:24-68) and calls the same five library
 * entry points.  The reference's own main compiles unchanged against this
 * library too (build/queries_refmain); this driver only adds what a benchmark
 * needs and the reference lacks: timing of the load and of execute_queries on
 * stderr (QCE_TIMING=1), never on stdout.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "../../include/qce_b200.h"
#include "../src/DArray.h"
#include "../src/alloc_free.h"
#include "../src/dbg.h"
#include "../src/parsing.h"
#include "../src/structs.h"
#include "../src/utilities.h"

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(void)
{
    const int timing = getenv("QCE_TIMING") != NULL;
    /* CUDA is touched lazily, by the first column load inside execute_queries: with QCE_GPUS=n
     * the host layer forks one process per GPU before that (schedule.c) */
    DArray *relations = DArray_create(sizeof(metadata), 10);
    double t0 = now_s();
    check(read_relations(relations) != -1, "Something went wrong in reading the relations");
    double t1 = now_s();

    DArray *queries = parser();
    check(queries != NULL, "Parsing failed");
    double t2 = now_s();
    execute_queries(queries, relations);
    qce_sync();
    double t3 = now_s();

    if (timing) {
        uint64_t rows = 0;
        for (size_t i = 0; i < DArray_count(relations); i++) rows += ((metadata *)DArray_get(relations, i))->tuples;
        fprintf(stderr, "{\"load_s\": %.6f, \"parse_s\": %.6f, \"execute_s\": %.6f, \"queries\": %u, \"rows_loaded\": %lu}\n",
                t1 - t0, t2 - t1, t3 - t2, DArray_count(queries), (unsigned long)rows);
    }

    for (size_t i = 0; i < DArray_count(queries); i++) {
        query *q = (query *)DArray_get(queries, i);
        FREE(q->relations);
        for (size_t j = 0; j < q->predicates_size; j++) FREE(q->predicates[j].second);
        FREE(q->predicates);
        FREE(q->selects);
    }
    DArray_destroy(queries);
    for (size_t i = 0; i < DArray_count(relations); i++) {
        metadata *m = (metadata *)DArray_get(relations, i);
        for (uint64_t c = 0; c < m->columns; c++) FREE(m->data[c]);
        FREE(m->data);
    }
    DArray_destroy(relations);
    qce_shutdown();
    return EXIT_SUCCESS;

error:
    return EXIT_FAILURE;
}
