/* batch.h -- in-process entry point of the host layer for embedders (bench.py,
 * tests): run query text against relations that are already resident on the
 * GPU and get back the bytes the reference would have written to stdout.
 * No reference counterpart (its only entry is main(), stdin -> stdout). */
#ifndef QCE_BATCH_H
#define QCE_BATCH_H

#include <stddef.h>

/* Parses `queries` (lines `r r|p&p|s s`, 'F' lines skipped), arranges and
 * executes each through execute_filter / execute_join / print_sums and copies
 * the stdout bytes (NUL-terminated) into out[0..cap).  Returns the number of
 * bytes produced (may exceed cap: output truncated), or -1 on failure.
 * *failed (may be NULL) = queries that were refused or failed. */
long qce_host_run_batch(const char *queries, char *out, size_t cap, int *failed);

#endif /* QCE_BATCH_H */
