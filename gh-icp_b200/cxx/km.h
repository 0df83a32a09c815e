// km.h — ghicp::Graph / ghicp::Km with the reference's interface (include/km.h:15-62), solved on the
// GPU through the C ABI (ghicp_km_solve) instead of the recursive O(n^3) Kuhn–Munkres of src/km.cpp.
#ifndef _INCLUDE_KM_H_
#define _INCLUDE_KM_H_
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ghicp_b200.h"
#include "utility.h"

namespace ghicp {

struct Graph {  // include/km.h:15-30
  std::vector<std::vector<double>> GTable;
  int n = 0, sp = 0, tp = 0;
  std::vector<int> match;
  std::vector<double> lx, ly, slack;
  std::vector<bool> visx, visy;
  double energy = 0;
  std::vector<int> min_match;
  int min_n = 0;
};

class Km {
 public:
  Km(Graph graph, double eps0, double penalty0) : penalty(penalty0), gra(graph), eps(eps0) {}  // km.h:38-43

  // src/km.cpp:40-126.  match[y] = x; pairs the reference would drop (weight == -penalty) get -1.
  void kmsolve() {
    const int n = gra.n;
    std::vector<double> W((size_t)n * n);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) W[(size_t)i * n + j] = gra.GTable[i][j];
    gra.match.assign(n, -1);
    const int sp = gra.sp > 0 ? gra.sp : n, tp = gra.tp > 0 ? gra.tp : n;
    int rc = ghicp_km_solve(device, W.data(), n, sp, tp, eps, penalty, gra.match.data(), &gra.energy, &rounds);
    if (rc < 0) throw std::runtime_error(std::string("ghicp_km_solve: ") + ghicp_last_error(nullptr));
  }
  // src/km.cpp:128-141 (computed by kmsolve)
  double Calenergy() { return gra.energy; }
  // src/km.cpp:144-233 without the Corres.txt side effect.  Returns cor_number.
  int output(std::vector<int> &SP, std::vector<int> &TP, std::vector<int> &SPout, std::vector<int> &TPout) {
    int cor_number = 0, cor_exact_num = 0;
    std::vector<char> s_used(gra.n, 0);
    for (int i = 0; i < gra.n; ++i)
      if (gra.match[i] >= 0) {
        if (gra.match[i] == i) cor_exact_num++;
        SP.push_back(gra.match[i]);
        TP.push_back(i);
        s_used[gra.match[i]] = 1;
        cor_number++;
      } else if (i < gra.tp) {
        TPout.push_back(i);
      }
    for (int i = 0; i < gra.sp; ++i)
      if (!s_used[i]) SPout.push_back(i);
    precision = 1.0 * cor_exact_num / cor_number;
    recall = 1.0 * cor_exact_num / gra.n;
    return cor_number;
  }
  bool findpath(int) { return false; }  // kept for source compatibility; the auction has no DFS

  double penalty;
  double precision = 0, recall = 0;
  int device = 0;
  int rounds = 0;

 private:
  Graph gra;
  double eps;
};
}  // namespace ghicp
#endif
