// Complex typedef (reference cfbasics/complexdefs.h); everything lives in mathdefs.h here.
#ifndef CFB200_COMPLEXDEFS_H
#define CFB200_COMPLEXDEFS_H
#include "cfbasics/mathdefs.h"
#endif
