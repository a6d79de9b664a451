/* charBuffer.h -- header name of the reference (src/sbmlsolver/charBuffer.h); the charBuffer_t functions are declared with the code generation seam */
#ifndef ODES_B200_FWD_CHARBUFFER_H
#define ODES_B200_FWD_CHARBUFFER_H
#include "sbmlsolver/odeModel.h"
#endif
