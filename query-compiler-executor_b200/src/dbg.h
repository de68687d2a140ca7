/* dbg.h -- stderr diagnostics in the reference's format
 * ("[ERROR] (file:line: errno: ...) message", /root/reference/src/dbg.h:17-29)
 * and the check()/check_mem() goto-error convention its main relies on.
 * stdout is the result channel and is never written from here. */
#ifndef QCE_DBG_H
#define QCE_DBG_H

#include <errno.h>
#include <stdio.h>
#include <string.h>

#ifdef DEBUG
#define debug(M, ...) fprintf(stderr, "DEBUG %s:%d: " M "\n", __FILE__, __LINE__, ##__VA_ARGS__)
#else
#define debug(M, ...)
#endif

#define clean_errno() (errno == 0 ? "None" : strerror(errno))

#define qce_log_(LEVEL, M, ...) \
    fprintf(stderr, "[" LEVEL "] (%s:%d: errno: %s) " M "\n", __FILE__, __LINE__, clean_errno(), ##__VA_ARGS__)
#define log_err(M, ...) qce_log_("ERROR", M, ##__VA_ARGS__)
#define log_warn(M, ...) qce_log_("WARN", M, ##__VA_ARGS__)
#define log_info(M, ...) fprintf(stderr, "[INFO] (%s:%d) " M "\n", __FILE__, __LINE__, ##__VA_ARGS__)

#define check(A, M, ...)              \
    if (!(A)) {                       \
        log_err(M, ##__VA_ARGS__);    \
        errno = 0;                    \
        goto error;                   \
    }
#define sentinel(M, ...)              \
    {                                 \
        log_err(M, ##__VA_ARGS__);    \
        errno = 0;                    \
        goto error;                   \
    }
#define check_mem(A) check((A), "Out of memory.")
#define check_debug(A, M, ...)        \
    if (!(A)) {                       \
        debug(M, ##__VA_ARGS__);      \
        errno = 0;                    \
        goto error;                   \
    }

#endif /* QCE_DBG_H */
