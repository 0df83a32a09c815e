// utility.h — enums of the reference's include/utility.h:51-64, same names and order.
// (The PCL typedefs of the reference's utility.h:21-46 are out of scope; include the reference's own
// utility.h instead when building inside its tree — the include guard is the same on purpose.)
#ifndef _INCLUDE_UTILITY_H
#define _INCLUDE_UTILITY_H
namespace ghicp {
enum FeatureType { BSC, RoPS, FPFH, None };
enum CorrespondenceType { NN, NNR, KM };
}  // namespace ghicp
#endif
