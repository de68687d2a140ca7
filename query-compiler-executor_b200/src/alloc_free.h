/* alloc_free.h -- allocation shorthands used by the host layer and by the
 * reference's main (MALLOC / CALLOC / REALLOC / FREE; same meaning as
 * /root/reference/src/alloc_free.h:4-10).  FREE is NULL-safe and clears the
 * pointer, which is what lets metadata `tuples` stay NULL here: the column data
 * lives in HBM, not in host memory. */
#ifndef QCE_ALLOC_FREE_H
#define QCE_ALLOC_FREE_H

#include <stdlib.h>

#define MALLOC(type, items) ((type *)malloc(sizeof(type) * (size_t)(items)))
#define CALLOC(items, size, type) ((type *)calloc((size_t)(items), (size_t)(size)))
#define REALLOC(pointer, size, type) ((type *)realloc((pointer), sizeof(*(pointer)) * (size_t)(size)))
#define FREE(pointer)           \
    do {                        \
        if ((pointer) != NULL) {\
            free(pointer);      \
            (pointer) = NULL;   \
        }                       \
    } while (0)

#endif /* QCE_ALLOC_FREE_H */
