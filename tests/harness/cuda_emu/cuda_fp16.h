// cuda_fp16.h — host emulation shim: just enough of __half for ghicp_device.cuh's h2d() and ghicp_stream.cu's fp16 weights
#pragma once
#include <stdint.h>
#include <string.h>
struct __half { unsigned short bits; };
inline __half __ushort_as_half(unsigned short b) { __half h; h.bits = b; return h; }
inline float __half2float(__half h) {
  const unsigned s = (h.bits >> 15) & 1u, e = (h.bits >> 10) & 31u, m = h.bits & 1023u;
  unsigned out;
  if (e == 0) {
    if (m == 0) out = s << 31;
    else { int k = 0; unsigned mm = m; while (!(mm & 1024u)) { mm <<= 1; ++k; } out = (s << 31) | ((unsigned)(113 - k) << 23) | ((mm & 1023u) << 13); }
  } else if (e == 31) out = (s << 31) | 0x7f800000u | (m << 13);
  else out = (s << 31) | ((e + 112u) << 23) | (m << 13);
  float f; memcpy(&f, &out, 4); return f;
}
inline unsigned short __half_as_ushort(__half h) { return h.bits; }
inline __half __int2half_rn(int v) {   // exact for |v| <= 2048 (the BSC Hamming distances); round-to-nearest-even beyond
  __half h; h.bits = 0;
  if (v == 0) return h;
  const unsigned s = v < 0 ? 1u : 0u;
  unsigned a = (unsigned)(v < 0 ? -v : v);
  int e = 31 - __builtin_clz(a);                  // a = 1.xxx * 2^e
  unsigned m;
  if (e <= 10) m = (a << (10 - e)) & 1023u;
  else {
    const unsigned sh = (unsigned)(e - 10), rem = a & ((1u << sh) - 1u), half = 1u << (sh - 1);
    unsigned q = a >> sh;
    if (rem > half || (rem == half && (q & 1u))) ++q;
    if (q >> 11) { q >>= 1; ++e; }
    m = q & 1023u;
  }
  h.bits = (unsigned short)((s << 15) | ((unsigned)(e + 15) << 10) | m);
  return h;
}
inline __half __float2half_rn(float f) {   // round to nearest even, through the compiler's IEEE binary16 type
  const _Float16 v = (_Float16)f;
  __half h; memcpy(&h.bits, &v, 2); return h;
}
