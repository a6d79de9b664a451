/* compile-time stand-in for the reference's graph drawing header (out of scope: SURVEY section 8), so that
   the reference's own example sources can be compiled against include/sbmlsolver unchanged (tests/test_examples.py) */
#ifndef ODES_B200_COMPAT_DRAWGRAPH_H
#define ODES_B200_COMPAT_DRAWGRAPH_H
#define drawSensitivity(data, file, format, threshold) ((void)0)
#define drawModel(m, file, format) ((void)0)
#define drawJacoby(data, file, format) ((void)0)
#endif
