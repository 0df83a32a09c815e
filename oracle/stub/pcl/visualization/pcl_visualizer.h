// STUB (oracle/stub): the viewer object src/ghicp_reg.cpp:28-29 constructs; it does nothing here.
#pragma once
#include <pcl/point_types.h>
#include <string>
namespace pcl { namespace visualization {
class PCLVisualizer {
 public:
  explicit PCLVisualizer(const std::string & = "") {}
  void setBackgroundColor(double, double, double) {}
};
} }
