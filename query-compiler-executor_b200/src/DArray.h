/* DArray.h -- small host-side vector of individually allocated elements.
 *
 * Only the tiny host lists use it: relation metadata, parsed queries, the
 * per-query entity list and its mid_result entries.  Row-id columns -- the one
 * large thing the reference keeps in a DArray (one calloc per row id,
 * /root/reference/src/DArray.h:51-60) -- are device arrays here
 * (struct qce_rowids, include/qce_b200.h).
 *
 * The struct layout and the inline accessors match the reference header
 * (/root/reference/src/DArray.h:13-19,46,83-90) because the reference's own
 * main/queries_main.c is compiled against that header and linked against this
 * library. */
#ifndef QCE_DARRAY_H
#define QCE_DARRAY_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

#include "alloc_free.h"
#include "dbg.h"

typedef struct DArray {
    int32_t end;         /* one past the last used slot */
    int32_t capacity;    /* slots allocated in `contents` */
    uint32_t count;      /* elements stored (== end) */
    size_t element_size; /* bytes copied per element */
    void **contents;     /* one heap block per element */
} DArray;

DArray *DArray_create(size_t element_size, int32_t initial_capacity);
void DArray_destroy(DArray *array);
void DArray_clear(DArray *array);
int DArray_push(DArray *array, void *element);
int DArray_pop(DArray *array);
int DArray_resize(DArray *array, int32_t newsize);
int DArray_expand(DArray *array);

#define DArray_count(A) ((A)->count)
#define DArray_end(A) ((A)->end)
#define DArray_capacity(A) ((A)->capacity)
#define DArray_first(A) ((A)->contents[0])
#define DArray_last(A) ((A)->contents[(A)->end - 1])

static inline void *DArray_get(DArray *array, ssize_t i)
{
    if (i >= array->capacity) {
        log_err("darray attempt to get past capacity");
        errno = 0;
        return NULL;
    }
    return array->contents[i];
}

/* Replace slot i with a copy of *element (append when i is past the end). */
static inline void DArray_set(DArray *array, ssize_t i, void *element)
{
    if (i >= array->capacity) {
        log_err("darray attempt to set past capacity");
        errno = 0;
        return;
    }
    if (i >= array->end) {
        i = array->end;
        array->end++;
        array->count++;
    }
    void *copy = malloc(array->element_size);
    if (copy == NULL) {
        log_err("Out of memory.");
        return;
    }
    memcpy(copy, element, array->element_size);
    free(array->contents[i]);
    array->contents[i] = copy;
}

#endif /* QCE_DARRAY_H */
