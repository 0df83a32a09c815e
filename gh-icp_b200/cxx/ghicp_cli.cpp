// ghicp_cli.cpp — headless command-line driver with the argument list of the reference's executable
// (test/ghicp_main.cpp:56-79, script/run.sh):
//
//   ghicp_cli target source registered  feature(B|F|R|N) corres(K|N|R)  downsample_resolution neighborhood_radius
//             curvature_non_max_radius weight_adjustment_ratio weight_adjustment_step registration_dof(4|6)
//             appro_overlap_ratio launch_realtime_viewer(ignored)
//
// Pipeline = test/ghicp_main.cpp:81-155 on the GPU through libghicp_b200.so: voxel down-sampling of both clouds,
// curvature keypoints (0.65 / 20 neighbours, :96-97), bounding-box magnitude of the down-sampled source (:91-93),
// GHRegistration, the float32 transform applied to the FULL source cloud (pcl::transformPointCloud with
// Rt_final.cast<float>(), :153) and written to `registered`; the 4x4 also goes to stdout and to `registered`.Rt.txt.
// Differences from the reference, all deliberate:
//   * no viewer windows (the last argument is accepted and ignored; :156-157 cannot run headless);
//   * feature 'N' (register on coordinates only) WORKS — the reference's switch falls into "Wrong feature input" and
//     exits (:135-139) although GHRegistration supports Ft = None (src/ghicp_reg.cpp:66-68);
//   * feature 'B' encodes the BSC descriptors on the GPU like :113-116 (BSCEncoder(curvature_non_max_radius, 7), target
//     dof 0, source registration_dof).  The sampling pattern is ./sample_pattern.txt when present — as in the reference,
//     which ships none — else the pattern the reference's constructor generates in a fresh process;
//   * features 'F' and 'R' need PCL's FPFH / RoPS estimators, which this library does not provide: the driver says so
//     and exits with status 2 (FPFH histograms computed elsewhere go through Keypoints::setFPFHfeature);
//   * .las input is not supported (needs libLAS + an interactive prompt).
// Utility modes (no GPU):  ghicp_cli --convert in.{pcd,ply,txt} out.{pcd,ply,txt}
//                          ghicp_cli --sample-pattern     write ./sample_pattern.txt the way the reference's
//                              BSCEncoder(radius, 7, true) does (binary_feature_extraction.hpp:75-103), read it back, report
// Exit status: 0 ok, 2 usage / unsupported option, 3 runtime error (incl. no CUDA device: there is no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "bsc_encoder.h"
#include "cloud_io.h"
#include "ghicp_reg.h"

using namespace ghicp;

static Cloud gather(const Cloud &c, const std::vector<int> &idx, int m) {
  Cloud o;
  o.xyz.reserve(3 * (size_t)m);
  for (int k = 0; k < m; ++k) o.push(c.xyz[3 * (size_t)idx[k]], c.xyz[3 * (size_t)idx[k] + 1], c.xyz[3 * (size_t)idx[k] + 2]);
  return o;
}
static void check_rc(int rc, const char *what) {
  if (rc < 0) throw std::runtime_error(std::string(what) + ": " + ghicp_last_error(nullptr));
}

int main(int argc, char **argv) {
  try {
    if (argc == 4 && std::string(argv[1]) == "--convert") {
      Cloud c;
      read_cloud(argv[2], c);
      write_cloud(argv[3], c);
      std::cout << "converted " << c.size() << " points" << std::endl;
      return 0;
    }
    if (argc == 2 && std::string(argv[1]) == "--sample-pattern") {
      BSCEncoder made(1.0f, 7, true);
      BSCEncoder back(1.0f, 7, false);
      std::vector<int> shipped(98);
      check_rc(ghicp_bsc_default_pattern(7, shipped.data()), "default pattern");
      bool same_back = back.pattern_from_file_, same_shipped = true;
      for (int i = 0; i < 49; ++i) {
        same_back = same_back && back.grid_index_pairs_2d_[i] == made.grid_index_pairs_2d_[i];
        same_shipped = same_shipped && made.grid_index_pairs_2d_[i].first == shipped[2 * i] && made.grid_index_pairs_2d_[i].second == shipped[2 * i + 1];
      }
      std::cout << "wrote sample_pattern.txt (49 pairs); read back " << (same_back ? "ok" : "MISMATCH") << "; "
                << (same_shipped ? "equals" : "differs from") << " the shipped default pattern" << std::endl;
      return same_back ? 0 : 3;
    }
    if (argc < 13) {
      std::cerr << "usage: " << argv[0] << " target source registered feature(B|F|R|N) corres(K|N|R) downsample_resolution "
                << "neighborhood_radius curvature_non_max_radius weight_adjustment_ratio weight_adjustment_step "
                << "registration_dof appro_overlap_ratio [launch_realtime_viewer]\n       " << argv[0]
                << " --convert in.{pcd,ply,txt} out.{pcd,ply,txt}\n       " << argv[0] << " --sample-pattern" << std::endl;
      return 2;
    }
    const std::string filenameT = argv[1], filenameS = argv[2], filenameR = argv[3];   // test/ghicp_main.cpp:56-58
    FeatureType Ft; CorrespondenceType Ct;
    switch (argv[4][0]) {                                                                // include/utility.h match_feature_type
      case 'B': Ft = BSC; break; case 'F': Ft = FPFH; break; case 'R': Ft = RoPS; break; case 'N': Ft = None; break;
      default: std::cerr << "unknown feature '" << argv[4] << "' (B, F, R, N)" << std::endl; return 2;
    }
    switch (argv[5][0]) {
      case 'K': Ct = KM; break; case 'N': Ct = NN; break; case 'R': Ct = NNR; break;
      default: std::cerr << "unknown correspondence method '" << argv[5] << "' (K, N, R)" << std::endl; return 2;
    }
    const float resolution = (float)atof(argv[6]), neighborhood_radius = (float)atof(argv[7]);
    const float curvature_non_max_radius = (float)atof(argv[8]), weight_adjustment_ratio = (float)atof(argv[9]);
    const float weight_adjustment_step = (float)atof(argv[10]);
    const int reg_dof = atoi(argv[11]);
    const float estimated_IoU = (float)atof(argv[12]);
    if (Ft == FPFH || Ft == RoPS) {
      std::cerr << "feature '" << argv[4] << "': the FPFH / RoPS estimators are PCL's and not part of libghicp_b200 (FPFH histograms "
                << "computed elsewhere go through Keypoints::setFPFHfeature); run with B or N" << std::endl;
      return 2;
    }
    if (!(resolution > 0.f) || !(neighborhood_radius > 0.f) || !(curvature_non_max_radius > 0.f)) {
      std::cerr << "resolution and radii must be positive" << std::endl;
      return 2;
    }

    Cloud cloudT, cloudS;                                                                // :81-85
    read_cloud(filenameT, cloudT);
    read_cloud(filenameS, cloudS);
    std::cout << "Target " << cloudT.size() << " points, source " << cloudS.size() << " points" << std::endl;

    // Down-sampling (:87-90)
    std::vector<int> keepT(cloudT.size() + 1), keepS(cloudS.size() + 1);
    int mT = 0, mS = 0;
    check_rc(ghicp_voxel_downsample(0, cloudT.xyz.data(), (int)cloudT.size(), resolution, keepT.data(), &mT), "voxel filter (target)");
    check_rc(ghicp_voxel_downsample(0, cloudS.xyz.data(), (int)cloudS.size(), resolution, keepS.data(), &mS), "voxel filter (source)");
    const Cloud downT = gather(cloudT, keepT, mT), downS = gather(cloudS, keepS, mS);
    std::cout << "Downsample done (" << mT << " / " << mS << " points)" << std::endl;
    float mn[3] = {downS.xyz[0], downS.xyz[1], downS.xyz[2]}, mx[3] = {mn[0], mn[1], mn[2]};   // getCloudBound, :91-93
    for (size_t i = 0; i < downS.size(); ++i)
      for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], downS.xyz[3 * i + a]); mx[a] = std::max(mx[a], downS.xyz[3 * i + a]); }
    const float bbx_magnitude = mx[0] - mn[0] + mx[1] - mn[1] + mx[2] - mn[2];

    // Keypoints (:95-100)
    const float non_stable_ratio_threshold = 0.65f;
    std::vector<int> kpT(mT), kpS(mS);
    int nkpt = 0, nkps = 0;
    check_rc(ghicp_detect_keypoints(0, downT.xyz.data(), mT, neighborhood_radius, non_stable_ratio_threshold, 20,
                                    curvature_non_max_radius, kpT.data(), &nkpt, nullptr, nullptr, nullptr), "keypoints (target)");
    check_rc(ghicp_detect_keypoints(0, downS.xyz.data(), mS, neighborhood_radius, non_stable_ratio_threshold, 20,
                                    curvature_non_max_radius, kpS.data(), &nkps, nullptr, nullptr, nullptr), "keypoints (source)");
    std::cout << "Keypoint detection done (" << nkpt << " / " << nkps << " keypoints)" << std::endl;
    if (nkps == 0 || nkpt == 0) throw std::runtime_error("no keypoints: check the radii against the cloud's scale");
    MatrixX3d kpSXYZ(nkps, 3), kpTXYZ(nkpt, 3);                                          // savecoordinates, dataio.hpp:609-626
    for (int i = 0; i < nkps; ++i) for (int a = 0; a < 3; ++a) kpSXYZ(i, a) = downS.xyz[3 * (size_t)kpS[i] + a];
    for (int i = 0; i < nkpt; ++i) for (int a = 0; a < 3; ++a) kpTXYZ(i, a) = downT.xyz[3 * (size_t)kpT[i] + a];
    Keypoints Kp;
    Kp.setCoordinate(kpSXYZ, kpTXYZ);
    if (Ft == BSC) {                                                                     // :109-119
      kpT.resize(nkpt); kpS.resize(nkps);
      BSCEncoder bsc(curvature_non_max_radius, 7);
      doubleVectorSBF bscT, bscS;
      bsc.extractBinaryFeatures(downT, kpT, 0, bscT);        // fixed feature (one per keypoint)
      bsc.extractBinaryFeatures(downS, kpS, reg_dof, bscS);  // 2 / 4 features per keypoint
      Kp.setBSCfeature(bscS, bscT);
      std::cout << "BSC pattern: " << (bsc.pattern_from_file_ ? "./sample_pattern.txt" : "shipped default") << std::endl;
    }

    // Registration (:141-151)
    Energyfunction Ef;
    Ef.init(nkps, nkpt, bbx_magnitude);
    Matrix4d Rt_final;
    GHRegistration ghreg(Kp, Ef, Ft, Ct, curvature_non_max_radius, weight_adjustment_ratio, weight_adjustment_step, reg_dof,
                         estimated_IoU);
    ghreg.set_viewer(false);
    ghreg.set_max_iterations(getenv("GHICP_MAX_ITER") ? atoi(getenv("GHICP_MAX_ITER")) : 0);
    ghreg.ghicp_reg(Rt_final);
    std::cout << "Registration done in " << ghreg.cor.size() << " iterations, " << (ghreg.cor.empty() ? 0 : ghreg.cor.back())
              << " correspondences, RMSE " << ghreg.RMS << std::endl;

    // pcl::transformPointCloud(*pointCloudS, *pointCloudS_reg, Rt_final.cast<float>()) (:153)
    float R[3][3], t[3];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[i][j] = (float)Rt_final(i, j); t[i] = (float)Rt_final(i, 3); }
    Cloud reg;
    reg.xyz.resize(cloudS.xyz.size());
    for (size_t i = 0; i < cloudS.size(); ++i) {
      const float x = cloudS.xyz[3 * i], y = cloudS.xyz[3 * i + 1], z = cloudS.xyz[3 * i + 2];
      for (int a = 0; a < 3; ++a) reg.xyz[3 * i + a] = R[a][0] * x + R[a][1] * y + R[a][2] * z + t[a];
    }
    write_cloud(filenameR, reg);
    std::ofstream rt((filenameR + ".Rt.txt").c_str());
    rt << std::setprecision(17);
    std::cout << std::setprecision(10) << "Rt_final (source -> target):" << std::endl;
    for (int i = 0; i < 4; ++i) {
      for (int j = 0; j < 4; ++j) { rt << Rt_final(i, j) << (j == 3 ? "\n" : " "); std::cout << Rt_final(i, j) << (j == 3 ? "\n" : " "); }
    }
    return 0;
  } catch (const std::exception &e) {
    std::cerr << "ghicp_cli: " << e.what() << std::endl;
    return 3;
  }
}
