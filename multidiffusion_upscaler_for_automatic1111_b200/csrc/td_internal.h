// Internal helpers shared by the translation units of libtd_b200.so.
#pragma once
#include <stdint.h>

void td_set_error(const char* fmt, ...);

static inline int td_dtype_size(int dtype) { return dtype == 2 /*TD_F32*/ ? 4 : ((dtype == 0 || dtype == 1) ? 2 : 0); }
