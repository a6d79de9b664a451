/*
 * check.h -- a minimal stand-in for the `check` unit-test framework (absent from this image), just large enough to
 * compile and run the reference's own unittest/*.c UNMODIFIED against this library (tests/test_reference_unittests.py).
 * Test infrastructure only.  Like `check`, every test runs in a forked child (the reference's tests rely on that: statics of
 * one test must not leak into the next); a failing assertion longjmps out of the test function, a crash counts as a
 * failure; fixtures (checked) run around every test of a case.  The subset: START_TEST / END_TEST, ck_assert, ck_assert_msg, ck_assert_int_eq / _ne,
 * ck_assert_str_eq, ck_abort_msg, fail_unless / fail_if, Suite / TCase / SRunner creation, srunner_run_all,
 * srunner_ntests_failed.
 */
#ifndef ODES_B200_COMPAT_CHECK_H
#define ODES_B200_COMPAT_CHECK_H
#include <setjmp.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef void (*TFun)(int);
typedef void (*SFun)(void);
typedef struct TCase { const char *name; TFun tests[64]; const char *tnames[64]; int ntests; SFun setup[8], teardown[8]; int nfix; } TCase;
typedef struct Suite { const char *name; TCase *cases[256]; int ncases; } Suite;
typedef struct SRunner { Suite *suites[64]; int nsuites; int nrun, nfailed; } SRunner;
enum print_output { CK_SILENT, CK_MINIMAL, CK_NORMAL, CK_VERBOSE, CK_ENV };

/* one copy per program: weak definitions are merged across the translation units that include this header */
__attribute__((weak)) jmp_buf ck_jmp_;
__attribute__((weak)) int ck_failed_;
__attribute__((weak)) const char *ck_current_;
static void ck_fail_(const char *file, int line, const char *fmt, ...)
{
  va_list ap;
  fprintf(stderr, "%s:%d: %s: ", file, line, ck_current_ ? ck_current_ : "?");
  va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
  fprintf(stderr, "\n");
  ck_failed_ = 1;
  longjmp(ck_jmp_, 1);
}
#define START_TEST(name_) static void name_(int _i) { (void)_i;
#define END_TEST }
#define ck_assert(cond) do { if (!(cond)) ck_fail_(__FILE__, __LINE__, "Assertion '%s' failed", #cond); } while (0)
#define ck_assert_msg(cond, ...) do { if (!(cond)) ck_fail_(__FILE__, __LINE__, __VA_ARGS__); } while (0)
#define fail_unless(cond, ...) do { if (!(cond)) ck_fail_(__FILE__, __LINE__, "Assertion '%s' failed", #cond); } while (0)
#define fail_if(cond, ...) do { if (cond) ck_fail_(__FILE__, __LINE__, "Failure '%s' occurred", #cond); } while (0)
#define ck_abort_msg(...) ck_fail_(__FILE__, __LINE__, __VA_ARGS__)
#define ck_abort() ck_fail_(__FILE__, __LINE__, "aborted")
#define ck_assert_int_eq(a, b) do { long a_ = (long)(a), b_ = (long)(b); if (a_ != b_) ck_fail_(__FILE__, __LINE__, "Assertion '%s == %s' failed: %ld != %ld", #a, #b, a_, b_); } while (0)
#define ck_assert_int_ne(a, b) do { long a_ = (long)(a), b_ = (long)(b); if (a_ == b_) ck_fail_(__FILE__, __LINE__, "Assertion '%s != %s' failed: both %ld", #a, #b, a_); } while (0)
#define ck_assert_str_eq(a, b) do { const char *a_ = (a), *b_ = (b); if (!a_ || !b_ || strcmp(a_, b_) != 0) ck_fail_(__FILE__, __LINE__, "Assertion '%s == %s' failed: \"%s\" != \"%s\"", #a, #b, a_ ? a_ : "(null)", b_ ? b_ : "(null)"); } while (0)

static Suite *suite_create(const char *name) { Suite *s = (Suite *)calloc(1, sizeof *s); s->name = name; return s; }
static TCase *tcase_create(const char *name) { TCase *t = (TCase *)calloc(1, sizeof *t); t->name = name; return t; }
static void suite_add_tcase(Suite *s, TCase *t) { if (s->ncases < 256) s->cases[s->ncases++] = t; }
static void tcase_add_test_(TCase *t, TFun f, const char *name) { if (t->ntests < 64) { t->tests[t->ntests] = f; t->tnames[t->ntests++] = name; } }
#define tcase_add_test(tc, f) tcase_add_test_((tc), (f), #f)
static void tcase_add_checked_fixture(TCase *t, SFun setup, SFun teardown) { if (t->nfix < 8) { t->setup[t->nfix] = setup; t->teardown[t->nfix++] = teardown; } }
#define tcase_add_unchecked_fixture tcase_add_checked_fixture
static void tcase_set_timeout(TCase *t, double s) { (void)t; (void)s; }
static SRunner *srunner_create(Suite *s) { SRunner *r = (SRunner *)calloc(1, sizeof *r); if (s) r->suites[r->nsuites++] = s; return r; }
static void srunner_add_suite(SRunner *r, Suite *s) { if (r->nsuites < 64) r->suites[r->nsuites++] = s; }
static void srunner_run_all(SRunner *r, enum print_output mode)
{
  int si, ci, ti, fi;
  (void)mode;
  for (si = 0; si < r->nsuites; si++) for (ci = 0; ci < r->suites[si]->ncases; ci++) {
    TCase *t = r->suites[si]->cases[ci];
    for (ti = 0; ti < t->ntests; ti++) {
      pid_t pid; int status = 0, failed;
      fflush(stdout); fflush(stderr);
      pid = fork();
      if (pid == 0) {
        ck_failed_ = 0; ck_current_ = t->tnames[ti];
        if (setjmp(ck_jmp_) == 0) {
          for (fi = 0; fi < t->nfix; fi++) if (t->setup[fi]) t->setup[fi]();
          t->tests[ti](0);
          for (fi = t->nfix - 1; fi >= 0; fi--) if (t->teardown[fi]) t->teardown[fi]();
        }
        fflush(stdout); fflush(stderr);
        _exit(ck_failed_ ? 1 : 0);
      }
      failed = (pid < 0 || waitpid(pid, &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0);
      if (pid > 0 && WIFSIGNALED(status)) fprintf(stderr, "%s: killed by signal %d\n", t->tnames[ti], WTERMSIG(status));
      r->nrun++;
      if (failed) r->nfailed++;
      printf("%s:%s:%s: %s\n", r->suites[si]->name, t->name, t->tnames[ti], failed ? "FAILED" : "ok");
    }
  }
  printf("%d%%: Checks: %d, Failures: %d\n", r->nrun ? 100 * (r->nrun - r->nfailed) / r->nrun : 100, r->nrun, r->nfailed);
}
static int srunner_ntests_failed(SRunner *r) { return r->nfailed; }
static int srunner_ntests_run(SRunner *r) { return r->nrun; }
static void srunner_free(SRunner *r) { free(r); }
static void srunner_set_fork_status(SRunner *r, int s) { (void)r; (void)s; }
#endif
