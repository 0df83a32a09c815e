// km_ref_shim.cpp — C wrapper around the REFERENCE's own Km class (include/km.h:32-61,
// src/km.cpp:13-233), compiled verbatim from /root/reference by oracle/Makefile into
// oracle/_ref/libkm_ref.so.  TEST INFRASTRUCTURE ONLY; contains no reference source, only calls it.
// `private` is opened so the wrapper can read Km::gra.match after kmsolve().
#define private public
#include "km.h"
#undef private
#include <unistd.h>
#include <cstdlib>
#include <string>

extern "C" {

// W: n x n row-major weights.  match[y] = x.  Returns 0.
int kmref_solve(const double *W, int n, double eps, int *match) {
  ghicp::Graph g;
  g.GTable.assign(n, std::vector<double>(n));
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) g.GTable[i][j] = W[(size_t)i * n + j];
  g.n = n; g.sp = n; g.tp = n;
  g.lx.resize(n); g.ly.resize(n); g.match.resize(n); g.slack.resize(n);
  g.visx.resize(n); g.visy.resize(n);
  ghicp::Km km(g, eps, 0.0);
  km.kmsolve();
  for (int i = 0; i < n; ++i) match[i] = km.gra.match[i];
  return 0;
}

// Full findcorrespondenceKM-style call: solve + output + Calenergy.  Km::output writes Corres.txt
// into the CWD (km.cpp:148); run from a scratch directory.
int kmref_solve_output(const double *W, int n, int sp, int tp, double eps, double penalty, int *match,
                       int *SP, int *TP, int *SPout, int *nSPout, int *TPout, int *nTPout,
                       double *energy) {
  ghicp::Graph g;
  g.GTable.assign(n, std::vector<double>(n));
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) g.GTable[i][j] = W[(size_t)i * n + j];
  g.n = n; g.sp = sp; g.tp = tp;
  g.lx.resize(n); g.ly.resize(n); g.match.resize(n); g.slack.resize(n);
  g.visx.resize(n); g.visy.resize(n);
  ghicp::Km km(g, eps, penalty);
  km.kmsolve();
  std::vector<int> sp_v, tp_v, spo, tpo;
  int cor = km.output(sp_v, tp_v, spo, tpo);
  double e = km.Calenergy();
  for (int i = 0; i < n; ++i) match[i] = km.gra.match[i];
  for (int i = 0; i < cor; ++i) { SP[i] = sp_v[i]; TP[i] = tp_v[i]; }
  for (size_t i = 0; i < spo.size(); ++i) SPout[i] = spo[i];
  for (size_t i = 0; i < tpo.size(); ++i) TPout[i] = tpo[i];
  *nSPout = (int)spo.size(); *nTPout = (int)tpo.size();
  *energy = e;
  return cor;
}
}
