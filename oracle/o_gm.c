/* o_gm.c -- ORACLE (test infrastructure): Gent-McWilliams placeholder.
 * hdifft_gm (source/hmix_gm.F90:1102-2219) is not restated yet; the dispatcher in o_ops.c never
 * calls into this file until it is (SURVEY 8a row a13 is open -- see DESIGN.md). */
#include "pop_oracle.h"
