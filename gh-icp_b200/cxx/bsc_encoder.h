// bsc_encoder.h — host-side mirror of the reference's ghicp::BSCEncoder (include/binary_feature_extraction.hpp:34-120,
// :603-676) over libghicp_b200.so: same constructor arguments, same extractBinaryFeatures result layout (four vectors of
// StereoBinaryFeature; the ones dof_type does not ask for hold default features), the descriptors computed on the GPU by
// ghicp_bsc_extract.  Differences, deliberate:
//   * clouds are ghicp::Cloud (cloud_io.h) and plain index vectors instead of pcl::PointCloud / pcl::PointIndices;
//   * without build_sample_pattern the constructor reads ./sample_pattern.txt like the reference (:107-116) and, when the
//     file does not exist (the reference would silently read zeros), falls back to the pattern the reference's own
//     constructor generates in a fresh process (ghicp_bsc_default_pattern; 7 x 7 grids only);
//   * no CPU path: without a CUDA device extractBinaryFeatures throws.
#pragma once
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/ghicp_b200.h"
#include "cloud_io.h"
#include "ghicp_types.h"

namespace ghicp {

class BSCEncoder {
 public:
  float extract_radius_;
  unsigned int voxel_side_num_;
  std::vector<std::pair<int, int>> grid_index_pairs_2d_;
  bool pattern_from_file_ = false;

  BSCEncoder(float extract_radius, unsigned int voxel_side_num, bool build_sample_pattern = false)
      : extract_radius_(extract_radius), voxel_side_num_(voxel_side_num) {
    const int cells = (int)(voxel_side_num_ * voxel_side_num_);
    if (build_sample_pattern) {   // :75-103: distinct cells, no pair twice (in either order); the C library's rand()
      while ((int)grid_index_pairs_2d_.size() < cells) {
        const int a = rand() % cells, b = rand() % cells;
        bool used = (a == b);
        for (const auto &p : grid_index_pairs_2d_) used = used || (p.first == a && p.second == b) || (p.first == b && p.second == a);
        if (!used) grid_index_pairs_2d_.push_back(std::make_pair(a, b));
      }
      std::ofstream out("sample_pattern.txt");
      for (const auto &p : grid_index_pairs_2d_) out << p.first << " " << p.second << std::endl;
      return;
    }
    std::ifstream in("sample_pattern.txt");
    if (in) {
      grid_index_pairs_2d_.resize(cells);
      for (auto &p : grid_index_pairs_2d_)
        if (!(in >> p.first >> p.second) || p.first < 0 || p.first >= cells || p.second < 0 || p.second >= cells)
          throw std::runtime_error("sample_pattern.txt: expected " + std::to_string(cells) + " pairs of cell indices below " + std::to_string(cells));
      pattern_from_file_ = true;
    } else {
      std::vector<int> flat(2 * (size_t)cells);
      if (ghicp_bsc_default_pattern((int)voxel_side_num_, flat.data()) < 0)
        throw std::runtime_error("BSCEncoder: no sample_pattern.txt in the working directory and no shipped pattern for this grid size");
      for (int i = 0; i < cells; ++i) grid_index_pairs_2d_.push_back(std::make_pair(flat[2 * i], flat[2 * i + 1]));
    }
  }

  // :603-676.  indices = keypoint indices into cloud.  bscFeatures receives four vectors (variants), like the reference.
  void extractBinaryFeatures(const Cloud &cloud, const std::vector<int> &indices, int dof_type, doubleVectorSBF &bscFeatures) {
    if (indices.empty()) { std::cout << "The input indice is NaN\n"; return; }   // :611-615
    const int nkp = (int)indices.size(), cells = (int)(voxel_side_num_ * voxel_side_num_), bits = 9 * cells, nbytes = (bits + 7) / 8;
    std::vector<int> flat(2 * (size_t)cells);
    for (int i = 0; i < cells; ++i) { flat[2 * i] = grid_index_pairs_2d_[i].first; flat[2 * i + 1] = grid_index_pairs_2d_[i].second; }
    std::vector<unsigned char> out((size_t)4 * nkp * nbytes);
    int V = 0;
    if (ghicp_bsc_extract(0, cloud.xyz.data(), (int)cloud.size(), indices.data(), nkp, extract_radius_, (int)voxel_side_num_,
                          flat.data(), dof_type, out.data(), &V, nullptr, nullptr) < 0)
      throw std::runtime_error(std::string("BSCEncoder: ") + ghicp_last_error(nullptr));
    for (int v = 0; v < 4; ++v) {
      vectorSBF feats(nkp);
      if (v < V)
        for (int i = 0; i < nkp; ++i) {
          feats[i] = StereoBinaryFeature((unsigned)bits);
          for (int b = 0; b < nbytes; ++b) feats[i].feature_[b] = (char)out[((size_t)v * nkp + i) * nbytes + b];
        }
      bscFeatures.push_back(feats);
    }
    std::cout << "Extract BSC feature done." << std::endl;   // :675
  }
};

}  // namespace ghicp
