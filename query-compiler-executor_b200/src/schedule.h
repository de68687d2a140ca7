/* schedule.h -- the batch scheduler around the reference's query loop
 * (execute_queries, /root/reference/src/utilities.c:289-300).
 *
 * The reference runs the queries of a batch one after another on one thread.
 * Here the same entry point
 *   - loads only the columns the batch references, from the mapped relation files,
 *     through pinned staging, while the first queries already run (8f-1);
 *   - with QCE_GPUS=n forks one process per GPU before CUDA is touched (8e): queries over
 *     row-sharded relations run on all ranks together (the operators exchange tuples over
 *     NVLink, include/qce_b200.h), queries over replicated relations run whole on one rank
 *     each ("replicas"), balanced by their input rows;
 *   - runs a rank's small queries on several engine contexts (streams) at once (8f-3) and
 *     reuses sorted base columns across the queries of the batch (engine side);
 *   - keeps stdout byte-identical: per-query output is buffered and written in query order,
 *     the stacked-filter count lines before their result line, and the reference's
 *     exit(EXIT_FAILURE) sites end the output exactly where the reference's would.
 */
#ifndef QCE_SCHEDULE_H
#define QCE_SCHEDULE_H

#include <stdio.h>

#include "DArray.h"

#define QCE_RUN_OK 0
#define QCE_RUN_FATAL 1 /* a query hit one of the reference's exit(EXIT_FAILURE) sites */

/* Arranges and executes every query; writes their stdout bytes to `out` in query order
 * (rank 0 only when several ranks run).  *failed (may be NULL) = queries that failed or were
 * refused (they print nothing but their count lines, as in the reference). */
int qce_run_queries(DArray *q_list, DArray *metadata_arr, FILE *out, int *failed);

/* The reference's "Something went really wrong" exits: ends the running query; the scheduler
 * prints everything up to and including its partial output and reports QCE_RUN_FATAL. */
void qce_fatal(void) __attribute__((noreturn));

#endif /* QCE_SCHEDULE_H */
