// cooperative_groups.h — host emulation shim: only this_grid().sync().  Single-block launches: a block barrier; multi-block
// grids must be started with emu::launch_cooperative (all blocks run concurrently as fibers).
#pragma once
#include <cuda_runtime.h>
namespace cooperative_groups {
struct grid_group {
  void sync() const {
    if (emu::g_blocks.size() == 1 && emu::g_gdim.x * emu::g_gdim.y * emu::g_gdim.z != 1) {
      fprintf(stderr, "emu: grid.sync() inside a multi-block launch that was not started with launch_cooperative\n");
      abort();
    }
    if (emu::g_blocks.size() == 1) emu::syncthreads(); else emu::gridsync();
  }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
