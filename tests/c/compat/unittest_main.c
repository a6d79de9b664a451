/* runs ONE suite of the reference's unittest/*.c (SUITE = name behind create_suite_); see tests/test_reference_unittests.py */
#include "unittest.h"
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
int main(void)
{
  SRunner *sr = srunner_create(CAT(create_suite_, SUITE)());
  int failed;
  srunner_run_all(sr, CK_NORMAL);
  failed = srunner_ntests_failed(sr);
  srunner_free(sr);
  return failed == 0 ? EXIT_SUCCESS : EXIT_FAILURE;
}
