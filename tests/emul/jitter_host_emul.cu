// TEST INFRASTRUCTURE: host execution of the per-element bodies of csrc/td_jitter.cu (index arithmetic and rounding
// sequence of the list-driven scatter / blend and the offset combine), so that they can be checked against torch on a
// machine without a GPU.  Built by tests/test_jitter_emulation.py into tests/_build/; never part of libtd_b200.so.
#define TD_JITTER_HOST_EMULATION 1
#include "td_jitter.cu"
