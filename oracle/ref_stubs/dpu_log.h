/* TEST INFRASTRUCTURE ONLY: see dpu.h in this directory. */
#include "dpu.h"
