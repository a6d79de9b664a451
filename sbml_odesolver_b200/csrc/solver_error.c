/*
 * solver_error.c -- typed message store.  Same calling convention as the
 * reference's src/solverError.c:75-358 (per-type lists, codes, dump helpers);
 * a mutex makes it usable from the batched path.
 */
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/solverError.h"

typedef struct { errorCode_t code; char *msg; } entry_t;
static struct { entry_t *e; int n, cap; } store[NUMBER_OF_ERROR_TYPES];
static pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
/* right-aligned type names and tab-separated fields: the dump format unittest/test_solverError.c:131-147 pins */
static const char *TYPE_NAME[NUMBER_OF_ERROR_TYPES] = { "Fatal Error", "      Error", "    Warning", "    Message" };

void SolverError_error(errorType_t type, errorCode_t code, const char *fmt, ...)
{
  char buf[2048]; va_list ap; entry_t *x;
  va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if ((int)type < 0 || type >= NUMBER_OF_ERROR_TYPES) return;
  pthread_mutex_lock(&lock);
  if (store[type].n == store[type].cap) {
    store[type].cap = store[type].cap ? 2 * store[type].cap : 16;
    store[type].e = (entry_t *)realloc(store[type].e, (size_t)store[type].cap * sizeof(entry_t));
  }
  x = &store[type].e[store[type].n++];
  x->code = code; x->msg = (char *)malloc(strlen(buf) + 1); strcpy(x->msg, buf);
  pthread_mutex_unlock(&lock);
}
int SolverError_getNum(errorType_t t) { return store[t].n; }
char *SolverError_getMessage(errorType_t t, int i) { return (i >= 0 && i < store[t].n) ? store[t].e[i].msg : NULL; }
errorCode_t SolverError_getCode(errorType_t t, int i) { return (i >= 0 && i < store[t].n) ? store[t].e[i].code : SOLVER_ERROR_NO_ERROR; }
errorCode_t SolverError_getLastCode(errorType_t t) { return store[t].n ? store[t].e[store[t].n - 1].code : SOLVER_ERROR_NO_ERROR; }
void SolverError_clear(void)
{
  int t, i;
  pthread_mutex_lock(&lock);
  for (t = 0; t < NUMBER_OF_ERROR_TYPES; t++) {
    for (i = 0; i < store[t].n; i++) free(store[t].e[i].msg);
    free(store[t].e); store[t].e = NULL; store[t].n = store[t].cap = 0;
  }
  pthread_mutex_unlock(&lock);
}
char *SolverError_dumpToString(void)
{
  size_t cap = 256, len = 0; char *s = (char *)malloc(cap); int t, i;
  s[0] = 0;
  for (t = 0; t < NUMBER_OF_ERROR_TYPES; t++)
    for (i = 0; i < store[t].n; i++) {
      char line[2300];
      int n = snprintf(line, sizeof line, "%s\t%d\t%s\n", TYPE_NAME[t], (int)store[t].e[i].code, store[t].e[i].msg);
      if (len + (size_t)n + 1 > cap) { cap = 2 * (len + (size_t)n + 1); s = (char *)realloc(s, cap); }
      memcpy(s + len, line, (size_t)n + 1); len += (size_t)n;
    }
  return s;
}
void SolverError_freeDumpString(char *s) { free(s); }
int SolverError_dumpHelper(char *s) { int n = (int)strlen(s); fputs(s, stderr); return n; }
void SolverError_dump(void) { char *s = SolverError_dumpToString(); fputs(s, stderr); free(s); }
void SolverError_dumpAndClearErrors(void) { SolverError_dump(); SolverError_clear(); }
/* src/solverError.c:196-270 */
static int memoryExhaustion = 0;
void SolverError_haltOnErrors(void) { if (SolverError_getNum(ERROR_ERROR_TYPE) || SolverError_getNum(FATAL_ERROR_TYPE)) exit(EXIT_FAILURE); }
int SolverError_isMemoryExhausted(void) { return memoryExhaustion; }
void *SolverError_calloc(size_t num, size_t size)
{
  void *r = calloc(num, size);
  memoryExhaustion = memoryExhaustion || !r;
  return r;
}
