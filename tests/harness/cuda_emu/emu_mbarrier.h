// emu_mbarrier.h — host emulation of an mbarrier and of the TMA bulk copy that completes on it (ghicp_stream.cu,
// ghicp_fdtc.cu under GHICP_EMU_HOST).  The barrier lives inside its own 64-bit shared-memory word: pending transaction bytes
// (signed), pending arrivals, the arrival count it was initialised with and the phase bit.  A bulk copy is a memcpy that
// completes its bytes at once; a waiting fiber yields.  Test infrastructure only.
#pragma once
#include <string.h>
#include <vector>
#include "emu.h"
namespace emu {
struct Bar { int tx; unsigned short pending; unsigned short init_phase; };   // init_phase: bit 15 = phase, low 15 bits = count
static_assert(sizeof(Bar) == 8, "an mbarrier is one 64-bit word");
inline void bar_check(Bar *b) {
  if (b->pending == 0 && b->tx == 0) { b->init_phase ^= 0x8000u; b->pending = (unsigned short)(b->init_phase & 0x7fffu); }
}
inline void bar_init(unsigned long long *bar, int count) {
  Bar *b = reinterpret_cast<Bar *>(bar);
  b->tx = 0; b->pending = (unsigned short)count; b->init_phase = (unsigned short)count;
}
inline void bar_arrive(unsigned long long *bar) { Bar *b = reinterpret_cast<Bar *>(bar); b->pending -= 1; bar_check(b); }
inline void bar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
  Bar *b = reinterpret_cast<Bar *>(bar);
  b->tx += (int)bytes; b->pending -= 1;
  bar_check(b);
}
inline void bar_wait(unsigned long long *bar, unsigned parity) {   // returns once the phase of that parity has completed
  static const bool negctl = getenv("GHICP_EMU_NEGCTL_NOWAIT") != nullptr;   // negative control: a kernel that forgot to wait
  if (negctl) return;
  Bar *b = reinterpret_cast<Bar *>(bar);
  while ((unsigned)(b->init_phase >> 15) == parity) yield();
}
struct PendingCopy { void *dst; const void *src; unsigned bytes; unsigned long long *bar; unsigned long long due; };
inline std::vector<PendingCopy> g_pending;
inline void land(const PendingCopy &c) {
  memcpy(c.dst, c.src, c.bytes);
  Bar *b = reinterpret_cast<Bar *>(c.bar);
  b->tx -= (int)c.bytes;
  bar_check(b);
}
inline void tma_progress() {   // yield hook: copies whose time has come land now
  for (size_t i = 0; i < g_pending.size();) {
    if (g_pending[i].due <= g_tick) { const PendingCopy c = g_pending[i]; g_pending[i] = g_pending.back(); g_pending.pop_back(); land(c); }
    else ++i;
  }
}
inline void bulk_copy(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  read_modes();
  const PendingCopy c{dst, src, bytes, bar, 0};
  if (g_tma_delay <= 0) { land(c); return; }
  // GHICP_EMU_TMA_DELAY: poison the destination, land the bytes a few scheduling ticks later
  memset(dst, 0xA5, bytes);
  PendingCopy d = c;
  d.due = g_tick + 1 + rnd() % (unsigned)g_tma_delay;
  g_pending.push_back(d);
  g_yield_hook = tma_progress;
}
}  // namespace emu
