// emu.h — a very small CUDA execution model on the CPU (test infrastructure, see cuda_runtime.h next to it).
// One thread block at a time; each CUDA thread is a ucontext fiber scheduled round-robin on ONE OS thread, so plain
// memory accesses need no synchronisation.  Rendezvous points:
//   syncthreads()  all live fibers of the block
//   shfl<T>()      the 32 fibers of a warp (full-warp, convergent shuffles only — what the product's kernels use)
// A fiber that returns from the kernel leaves every later rendezvous (like an exited CUDA thread).
#pragma once
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace emu {

struct Fiber {
  ucontext_t ctx;
  dim3 tidx, bidx;
  int block = 0;       // linear block number inside the running launch (cooperative launches hold several)
  bool done = false;
  char *stack = nullptr;
};
struct BlockState { int live = 0, bar_count = 0, bar_gen = 0; };
struct Warp {
  unsigned long long in[32], out[32];
  int count = 0, gen = 0, live = 0;
};

inline dim3 g_bdim, g_gdim;
inline std::vector<Fiber> g_fibers;
inline std::vector<Warp> g_warps;
inline int g_cur_index = 0, g_live = 0;
inline std::vector<BlockState> g_blocks;          // one per block of the running launch
inline int g_grid_count = 0, g_grid_gen = 0;      // grid-wide barrier of a cooperative launch
inline int g_threads_per_block = 0;
inline ucontext_t g_sched;
inline const std::function<void()> *g_body = nullptr;
inline std::vector<unsigned long long> g_dyn;   // dynamic shared memory of the running launch (8-byte aligned)
inline void *dyn_smem() { return g_dyn.data(); }
constexpr size_t STACK = 256 * 1024;

inline Fiber *cur() { return &g_fibers[g_cur_index]; }
// Stress modes (environment, read once):
//   GHICP_EMU_SCHED=<seed>     every scheduling sweep visits the fibers in a fresh pseudo-random order instead of 0, 1, 2 ...
//                              (a correct kernel only depends on its rendezvous points, not on who runs first)
//   GHICP_EMU_TMA_DELAY=<n>    a bulk copy lands 1..n scheduling ticks AFTER it was issued (emu_mbarrier.h): code that reads
//                              the destination without waiting on the mbarrier sees stale bytes
inline unsigned long long g_tick = 0;                 // advances on every yield
inline void (*g_yield_hook)() = nullptr;              // deferred work that becomes due as ticks pass (emu_mbarrier.h)
inline unsigned long long g_rng = 0;
inline int g_sched_random = -1, g_tma_delay = -1;
inline void read_modes() {
  if (g_sched_random >= 0) return;
  const char *s = getenv("GHICP_EMU_SCHED"), *d = getenv("GHICP_EMU_TMA_DELAY");
  g_sched_random = s ? 1 : 0;
  g_rng = s ? (unsigned long long)atoll(s) * 0x9E3779B97F4A7C15ull + 0x1234567ull : 0;
  g_tma_delay = d ? atoi(d) : 0;
}
inline unsigned rnd() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17; return (unsigned)(g_rng >> 11); }
inline void yield() {
  ++g_tick;
  if (g_yield_hook) g_yield_hook();
  swapcontext(&cur()->ctx, &g_sched);
}

inline void syncthreads() {
  BlockState &b = g_blocks[cur()->block];
  const int my = b.bar_gen;
  if (++b.bar_count == b.live) { b.bar_count = 0; ++b.bar_gen; return; }
  while (b.bar_gen == my) yield();
}
inline void gridsync() {   // cooperative launches only: every live fiber of every block
  const int my = g_grid_gen;
  if (++g_grid_count == g_live) { g_grid_count = 0; ++g_grid_gen; return; }
  while (g_grid_gen == my) yield();
}

template <typename T>
inline T shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shfl: at most 64-bit values");
  Warp &w = g_warps[cur()->block * ((g_threads_per_block + 31) / 32) + (cur()->tidx.x >> 5)];
  const int lane = cur()->tidx.x & 31;
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  w.in[lane] = raw;
  const int my = w.gen;
  if (++w.count == w.live) {
    memcpy(w.out, w.in, sizeof(w.out));
    w.count = 0;
    ++w.gen;
  } else {
    while (w.gen == my) yield();
  }
  T r;
  memcpy(&r, &w.out[src_lane & 31], sizeof(T));
  return r;
}

// all 32 lane values of a convergent warp-wide operation (ballot / redux)
template <typename T>
inline void warp_all(T v, T out[32]) {
  Warp &w = g_warps[cur()->block * ((g_threads_per_block + 31) / 32) + (cur()->tidx.x >> 5)];
  const int lane = cur()->tidx.x & 31;
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  w.in[lane] = raw;
  const int my = w.gen;
  if (++w.count == w.live) {
    memcpy(w.out, w.in, sizeof(w.out));
    w.count = 0;
    ++w.gen;
  } else {
    while (w.gen == my) yield();
  }
  for (int l = 0; l < 32; ++l) memcpy(&out[l], &w.out[l], sizeof(T));
}

inline void fiber_entry() {
  (*g_body)();
  Fiber *f = cur();
  f->done = true;
  --g_live;
  BlockState &b = g_blocks[f->block];
  --b.live;
  Warp &w = g_warps[f->block * ((g_threads_per_block + 31) / 32) + (f->tidx.x >> 5)];
  --w.live;
  // an exiting thread may complete a rendezvous the others are waiting in
  if (b.live > 0 && b.bar_count == b.live) { b.bar_count = 0; ++b.bar_gen; }
  if (g_live > 0 && g_grid_count == g_live) { g_grid_count = 0; ++g_grid_gen; }
  if (w.live > 0 && w.count == w.live) { memcpy(w.out, w.in, sizeof(w.out)); w.count = 0; ++w.gen; }
  swapcontext(&f->ctx, &g_sched);
}

// run the fibers [0, nfib): round-robin until all have returned
inline void run_fibers(int nfib) {
  long spins = 0;
  read_modes();
  std::vector<int> order;
  if (g_sched_random) { order.resize(nfib); for (int t = 0; t < nfib; ++t) order[t] = t; }
  while (g_live > 0) {
    if (g_sched_random) for (int t = nfib - 1; t > 0; --t) std::swap(order[t], order[rnd() % (unsigned)(t + 1)]);
    for (int q = 0; q < nfib; ++q) {
      const int t = g_sched_random ? order[q] : q;
      if (g_fibers[t].done) continue;
      g_cur_index = t;
      swapcontext(&g_sched, &g_fibers[t].ctx);
    }
    if (++spins > 100000000L) { fprintf(stderr, "emu: deadlock (divergent barrier / shuffle?)\n"); abort(); }
  }
  if (g_yield_hook) { g_tick += 1ull << 20; g_yield_hook(); }   // whatever is still in flight lands before the launch "returns"
}
inline void prepare_fiber(int t, dim3 tidx, dim3 bidx, int block) {
  Fiber &f = g_fibers[t];
  f.done = false; f.tidx = tidx; f.bidx = bidx; f.block = block;
  getcontext(&f.ctx);
  f.ctx.uc_stack.ss_sp = f.stack;
  f.ctx.uc_stack.ss_size = STACK;
  f.ctx.uc_link = nullptr;
  makecontext(&f.ctx, fiber_entry, 0);
}
inline void ensure_fibers(int n) {
  if ((int)g_fibers.size() < n) {
    const size_t old = g_fibers.size();
    g_fibers.resize(n);
    for (size_t i = old; i < g_fibers.size(); ++i) g_fibers[i].stack = (char *)malloc(STACK);
  }
}

// run `body` (a call of the kernel function) for every thread of every block of the grid (1-D blocks), one block at a time
inline void launch(dim3 grid, dim3 block, const std::function<void()> &body, size_t dyn_smem_bytes = 0) {
  const int nt = (int)block.x;
  g_dyn.assign(dyn_smem_bytes / 8 + 1, 0ull);
  if (block.y != 1 || block.z != 1) { fprintf(stderr, "emu: 1-D blocks only\n"); abort(); }
  ensure_fibers(nt);
  g_bdim = block; g_gdim = grid; g_body = &body; g_threads_per_block = nt;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_warps.assign((nt + 31) / 32, Warp());
        g_blocks.assign(1, BlockState());
        g_blocks[0].live = nt;
        g_live = nt; g_grid_count = 0; g_grid_gen = 0;
        for (int t = 0; t < nt; ++t) { prepare_fiber(t, dim3(t, 0, 0), dim3(bx, by, bz), 0); g_warps[t >> 5].live++; }
        run_fibers(nt);
      }
}
// cooperative launch: ALL blocks of a (small, 1-D) grid run concurrently, gridsync() is a rendezvous of every fiber.
// `static` shared-memory arrays are one copy for the whole launch: only valid for kernels in which a single block uses them.
inline void launch_cooperative(dim3 grid, dim3 block, const std::function<void()> &body) {
  const int nt = (int)block.x, nb = (int)grid.x;
  if (block.y != 1 || block.z != 1 || grid.y != 1 || grid.z != 1) { fprintf(stderr, "emu: 1-D cooperative launches only\n"); abort(); }
  ensure_fibers(nt * nb);
  g_dyn.assign(1, 0ull);
  g_bdim = block; g_gdim = grid; g_body = &body; g_threads_per_block = nt;
  const int wpb = (nt + 31) / 32;
  g_warps.assign((size_t)wpb * nb, Warp());
  g_blocks.assign(nb, BlockState());
  g_live = nt * nb; g_grid_count = 0; g_grid_gen = 0;
  for (int b = 0; b < nb; ++b) {
    g_blocks[b].live = nt;
    for (int t = 0; t < nt; ++t) { prepare_fiber(b * nt + t, dim3(t, 0, 0), dim3(b, 0, 0), b); g_warps[(size_t)b * wpb + (t >> 5)].live++; }
  }
  run_fibers(nt * nb);
}

}  // namespace emu
