/*
 * util.c -- helpers of the reference's public API (src/util.c:60-200, declared in sbmlsolver/util.h): line reader,
 * path join, message printers, checked allocation.
 */
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/util.h"

void *xalloc(size_t nmemb, size_t size)
{
  void *p = calloc(nmemb ? nmemb : 1, size ? size : 1);
  if (!p) fatal(stderr, "xalloc(): memory allocation of %lu x %lu bytes failed", (unsigned long)nmemb, (unsigned long)size);
  return p;
}
void xfree(void *ptr) { free(ptr); }
void *space(unsigned size) { return xalloc(1, size); }
void *xrealloc(void *p, unsigned size)
{
  void *q = realloc(p, size ? size : 1);
  if (!q) fatal(stderr, "xrealloc(): memory allocation of %u bytes failed", size);
  return q;
}
/* a whole line, however long, without the terminating newline; the last line of a file may lack the newline */
char *get_line(FILE *fp)
{
  size_t cap = 256, len = 0; int c;
  char *line;
  if ((c = fgetc(fp)) == EOF) return NULL;
  line = (char *)space((unsigned)cap);
  for (; c != EOF && c != '\n'; c = fgetc(fp)) {
    if (len + 2 > cap) { cap *= 2; line = (char *)xrealloc(line, (unsigned)cap); }
    line[len++] = (char)c;
  }
  line[len] = '\0';
  return line;
}
char *concat(const char *a, const char *b)
{
  size_t alen = strlen(a), blen = strlen(b);
  char *r = (char *)xalloc(alen + blen + 2, sizeof(char));
  memcpy(r, a, alen);
  if (alen == 0 || a[alen - 1] != '/') r[alen++] = '/';
  memcpy(r + alen, b, blen + 1);
  return r;
}
void fatal(FILE *hdl, const char *fmt, ...)
{
  va_list ap;
  if (!hdl) hdl = stderr;
  va_start(ap, fmt);
  if (fmt) { fprintf(hdl, "FATAL ERROR: "); vfprintf(hdl, fmt, ap); fprintf(hdl, "\n"); } else fprintf(hdl, "FATAL ERROR\n");
  va_end(ap);
  fflush(hdl);
  exit(EXIT_FAILURE);
}
void Warn(FILE *hdl, const char *fmt, ...)
{
  va_list ap;
  if (!hdl) hdl = stderr;
  va_start(ap, fmt);
  if (fmt) { fprintf(hdl, "WARNING: "); vfprintf(hdl, fmt, ap); fprintf(hdl, "\n"); } else fprintf(hdl, "WARNING\n");
  va_end(ap);
  fflush(hdl);
}
