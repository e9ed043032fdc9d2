// cooperative_groups.h -- TEST STUB (tests/emul): a cooperative launch is emulated with ONE
// CTA, so the grid barrier is the CTA barrier.
#pragma once
#include "cuda_runtime.h"
namespace cooperative_groups {
struct grid_group { void sync() const { emu::barrier_cta(); } };
static inline grid_group this_grid() { return grid_group(); }
}
