/* utilities.h -- loader, mid-result lookup and the query loop (entry points of
 * /root/reference/src/utilities.h:33-39).
 *
 * The reference's CPU radix helpers (build_histogram / build_psum /
 * build_reordered_array, src/utilities.h:21-31) have no host counterpart here:
 * they operate on host AoS tuple arrays that no longer exist; their work is
 * done by the radix kernels behind qce_sort_tuples (include/qce_b200.h). */
#ifndef QCE_UTILITIES_H
#define QCE_UTILITIES_H

#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/types.h>

#include "DArray.h"
#include "alloc_free.h"
#include "dbg.h"
#include "structs.h"

/* where a (relation id, binding) pair was found among the entities */
typedef struct exists_info {
    ssize_t mid_result; /* entity index */
    ssize_t index;      /* entry index inside the entity, -1 = not found */
} exists_info;

/* Reads relation file paths from stdin up to `Done`, uploads every column to
 * the GPU and appends one `metadata` per file.  0, or -1 on failure. */
int read_relations(DArray *metadata_arr);
/* Same protocol from any stream. */
int read_relations_from(FILE *in, DArray *metadata_arr);

/* Newest entity first, first matching entry. */
exists_info relation_exists(DArray *mid_results_array, uint64_t relation, uint64_t predicate_id);
/* Inside one entity, last matching entry (-1 if none). */
ssize_t relation_exists_current(DArray *mid_results, uint64_t relation, uint64_t predicate_id);

/* Arranges and executes every query of the list, printing one result line per
 * query (and the stacked-filter count lines) to stdout. */
void execute_queries(DArray *q_list, DArray *metadata_arr);
/* One query; appends its stdout bytes to `out` instead of printing when `out`
 * is not NULL (used by the batch driver to keep stdout in query order). */
int execute_query_to(query *q, DArray *metadata_arr, FILE *out);

#endif /* QCE_UTILITIES_H */
