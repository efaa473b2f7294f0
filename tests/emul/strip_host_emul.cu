// TEST INFRASTRUCTURE: host execution of the strip blend kernel of csrc/td_strip.cu -- every CTA, thread by thread and
// phase by phase (table, copies, consume) over an emulated shared-memory array -- so that its index arithmetic, staging
// layout and rounding sequence can be checked against the reference's fixtures on a machine without a GPU.
// Built by tests/test_strip_emulation.py into tests/_build/; never part of libtd_b200.so.
#define TD_STRIP_HOST_EMULATION 1
#include "td_strip.cu"
