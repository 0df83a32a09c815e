// cuda_fp16.h — host emulation shim: just enough of __half for ghicp_device.cuh's h2d()
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
