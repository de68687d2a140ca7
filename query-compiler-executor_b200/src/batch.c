/* batch.c -- see batch.h. */
#define _GNU_SOURCE
#include "batch.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "parsing.h"
#include "pred_arrange.h"
#include "schedule.h"
#include "structs.h"
#include "utilities.h"

static void free_queries(DArray *queries)
{
    for (size_t i = 0; i < DArray_count(queries); i++) {
        query *q = (query *)DArray_get(queries, i);
        FREE(q->relations);
        for (size_t j = 0; j < q->predicates_size; j++) FREE(q->predicates[j].second);
        FREE(q->predicates);
        FREE(q->selects);
    }
    DArray_destroy(queries);
}

long qce_host_run_batch(const char *text, char *out, size_t cap, int *failed)
{
    FILE *in = fmemopen((void *)text, strlen(text), "r");
    if (in == NULL) return -1;
    DArray *queries = parser_from(in);
    fclose(in);
    if (queries == NULL) return -1;

    char *buf = NULL;
    size_t len = 0;
    FILE *mem = open_memstream(&buf, &len);
    int bad = 0;
    /* the same scheduler as execute_queries (schedule.c); a reference exit(EXIT_FAILURE) site
     * ends the output after that query's partial line and counts as a failure */
    if (qce_run_queries(queries, NULL, mem, &bad) == QCE_RUN_FATAL) bad++;
    fclose(mem);
    free_queries(queries);
    if (failed) *failed = bad;
    if (out && cap) {
        size_t n = len < cap - 1 ? len : cap - 1;
        memcpy(out, buf, n);
        out[n] = '\0';
    }
    free(buf);
    return (long)len;
}
