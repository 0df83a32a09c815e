// STUB (oracle/stub), force-included (-include) when the reference's sources are compiled without PCL / Eigen:
// include/utility.h needs the real libraries for helpers the hot path never touches, so its include guard is pre-defined
// (-D_INCLUDE_UTILITY_H) and only the NAMES the path uses are supplied here — the typedefs of utility.h:23-48 and the two
// enums of utility.h:51-64 (same enumerators, same order).  Test infrastructure only.
#pragma once
#include <list>
#include <vector>
#include <pcl/point_types.h>
typedef pcl::PointCloud<pcl::PointXYZI>::Ptr pcXYZIPtr;
typedef pcl::PointCloud<pcl::PointXYZI> pcXYZI;
typedef pcl::PointCloud<pcl::PointXYZ>::Ptr pcXYZPtr;
typedef pcl::PointCloud<pcl::PointXYZ> pcXYZ;
typedef pcl::PointCloud<pcl::Normal>::Ptr NormalsPtr;
typedef pcl::PointCloud<pcl::Normal> Normals;
typedef pcl::PointCloud<pcl::FPFHSignature33>::Ptr fpfhFeaturePtr;
typedef pcl::PointCloud<pcl::FPFHSignature33> fpfhFeature;
namespace ghicp {
enum FeatureType { BSC, RoPS, FPFH, None };
enum CorrespondenceType { NN, NNR, KM };
// named in signatures of include/filter.hpp (utility.h:66-90, 131-...): declarations only, nothing here is called
struct CenterPoint { double x, y, z; };
struct Bounds { double min_x, min_y, min_z, max_x, max_y, max_z; };
template <typename PointT> class CloudUtility {};
}
