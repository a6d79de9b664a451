/* util.h -- small helpers of the reference's public API (src/sbmlsolver/util.h:40-70): reading a line of any length,
   joining path components, message printers, checked allocation. */
#ifndef ODES_B200_UTIL_H
#define ODES_B200_UTIL_H
#include <stdio.h>
#include "sbmlsolver/exportdefs.h"
#ifdef __cplusplus
extern "C" {
#endif
void *space(unsigned size);                 /* zeroed block; exits on exhaustion */
void *xrealloc(void *p, unsigned size);
char *get_line(FILE *fp);                   /* next line without its newline (caller frees), NULL at end of file */
char *concat(const char *a, const char *b); /* a + "/" + b (no doubled separator) */
void fatal(FILE *hdl, const char *fmt, ...);
void Warn(FILE *hdl, const char *fmt, ...);
void *xalloc(size_t nmemb, size_t size);
void xfree(void *ptr);
#ifdef __cplusplus
}
#endif
#endif
