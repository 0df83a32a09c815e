// STUB (oracle/stub): stands in for the reference's include/cloud_viewer.hpp (VTK / PCLVisualizer windows), which
// src/ghicp_reg.cpp:10 includes and :26, :46-47, :99-100, :109-110 call only when launch_viewer_ is set.  Does nothing.
#pragma once
namespace ghicp {
template <typename PointT> class CloudViewer {
 public:
  template <typename... Args> void displayRegistration_on_fly(Args &&...) {}
};
}
