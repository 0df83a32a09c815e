// cooperative_groups.h — host emulation shim: only this_grid().sync(), and only for single-block launches (the emulator runs
// one block at a time, so a grid-wide barrier of a multi-block launch cannot be emulated).
#pragma once
#include <cuda_runtime.h>
namespace cooperative_groups {
struct grid_group {
  void sync() const {
    if (emu::g_gdim.x * emu::g_gdim.y * emu::g_gdim.z != 1) { fprintf(stderr, "emu: grid.sync() of a multi-block launch\n"); abort(); }
    emu::syncthreads();
  }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
