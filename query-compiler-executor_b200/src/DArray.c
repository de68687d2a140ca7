/* DArray.c -- see DArray.h.  Growth doubles the slot table; elements are
 * separate heap blocks so pointers to them stay valid across pushes. */
#include "DArray.h"

DArray *DArray_create(size_t element_size, int32_t initial_capacity)
{
    if (initial_capacity <= 0) {
        log_err("You must set an initial capacity > 0");
        return NULL;
    }
    DArray *a = (DArray *)malloc(sizeof *a);
    if (a == NULL) return NULL;
    a->contents = (void **)calloc((size_t)initial_capacity, sizeof(void *));
    if (a->contents == NULL) {
        free(a);
        return NULL;
    }
    a->end = 0;
    a->count = 0;
    a->capacity = initial_capacity;
    a->element_size = element_size;
    return a;
}

void DArray_clear(DArray *a)
{
    for (int32_t i = 0; i < a->end; i++) {
        free(a->contents[i]);
        a->contents[i] = NULL;
    }
    a->end = 0;
    a->count = 0;
}

void DArray_destroy(DArray *a)
{
    if (a == NULL) return;
    DArray_clear(a);
    free(a->contents);
    free(a);
}

int DArray_resize(DArray *a, int32_t newsize)
{
    if (newsize <= 0) {
        log_err("The new size must be > 0");
        return -1;
    }
    void **grown = (void **)realloc(a->contents, sizeof(void *) * (size_t)newsize);
    if (grown == NULL) return -1;
    for (int32_t i = a->capacity; i < newsize; i++) grown[i] = NULL;
    a->contents = grown;
    a->capacity = newsize;
    return 0;
}

int DArray_expand(DArray *a) { return DArray_resize(a, a->capacity * 2); }

int DArray_push(DArray *a, void *element)
{
    void *copy = malloc(a->element_size ? a->element_size : 1);
    if (copy == NULL) return -1;
    memcpy(copy, element, a->element_size);
    a->contents[a->end++] = copy;
    a->count++;
    return a->end >= a->capacity ? DArray_expand(a) : 0;
}

int DArray_pop(DArray *a)
{
    if (a->end == 0) {
        log_err("Attempt to pop from empty array");
        return -1;
    }
    a->end--;
    a->count--;
    free(a->contents[a->end]);
    a->contents[a->end] = NULL;
    return 0;
}
