"""Raw clouds -> transform: the steps of the reference's driver (test/ghicp_main.cpp:86-155) over the C ABI, in Python.

    Rt, info = register_clouds(target_xyz, source_xyz, resolution=0.1, neighborhood_radius=0.5,
                               curvature_non_max_radius=1.0)

Voxel down-sampling of both clouds (CFilter::voxelfilter), curvature keypoints (0.65 / 20 neighbours, :96-97), the
bounding-box magnitude of the down-sampled source (:91-93), GHRegistration on the keypoints.  Descriptors are optional:
`features(target_down, target_kp_idx, source_down, source_kp_idx) -> Keypoints` may attach BSC / FPFH descriptors computed
elsewhere (the encoders are not part of this library); without it the registration runs on coordinates
only (Ft = None), which GHRegistration supports (src/ghicp_reg.cpp:66-68) although the reference's own main() rejects it.
With feature_type = FT_BSC and no `features` callback the BSC descriptors are encoded on the GPU like :113-116
(capi.bsc_extract: radius = curvature_non_max_radius, 7 x 7 grids, dof_type 0 for the target, `dof_type` for the source;
`bsc_pattern` = the sampling pairs, default = the pattern the reference's constructor generates).
"""
import numpy as np

from . import capi
from .registration import Energyfunction, GHRegistration, Keypoints


def register_clouds(target_xyz, source_xyz, resolution, neighborhood_radius, curvature_non_max_radius, corr_type=capi.CT_NN,
                    feature_type=capi.FT_NONE, features=None, weight_adjustment_ratio=1.1, weight_adjustment_step=0.1,
                    dof_type=6, estimated_IoU=0.5, max_iter=0, device=0, bsc_pattern=None, **reg_kw):
    T = np.ascontiguousarray(target_xyz, dtype=np.float32)
    S = np.ascontiguousarray(source_xyz, dtype=np.float32)
    if features is None:
        # device-resident pipeline: each raw cloud is uploaded once, the stages chain on the GPU, the keypoint coordinates and
        # descriptors go device-to-device into the registration context (ghicp_prep_run + ghicp_set_from_prep)
        want_bsc = feature_type == capi.FT_BSC
        radius = curvature_non_max_radius if want_bsc else 0.0
        pt = capi.Prep(T, resolution, neighborhood_radius, curvature_non_max_radius, radius, 0, pairs=bsc_pattern, device=device)       # :89-115
        ps = capi.Prep(S, resolution, neighborhood_radius, curvature_non_max_radius, radius, dof_type, pairs=bsc_pattern, device=device)  # :116
        try:
            for name, pr in (("target", pt), ("source", ps)):
                if pr.n_kp == 0:
                    raise capi.GhicpError(-1, f"no keypoints in the {name} cloud: check the radii")
            bbx = ps.bbx_magnitude                                                       # getCloudBound, :91-93
            Ef = Energyfunction().init(ps.n_kp, pt.n_kp, bbx)
            reg = GHRegistration((ps, pt), Ef, feature_type, corr_type, curvature_non_max_radius, weight_adjustment_ratio,
                                 weight_adjustment_step, dof_type, estimated_IoU, max_iter=max_iter, device=device, **reg_kw)
            Rt, iterations = reg.ghicp_reg()
            info = dict(iterations=iterations, n_target_down=pt.n_down, n_source_down=ps.n_down, n_target_kp=pt.n_kp,
                        n_source_kp=ps.n_kp, bbx_magnitude=bbx, cor=len(reg.pairs()[0]),
                        stage_ms=dict(target=pt.stage_ms, source=ps.stage_ms))
            reg.close()
        finally:
            pt.close(); ps.close()
        return Rt, info
    down, kp_idx = {}, {}
    for name, P in (("T", T), ("S", S)):
        keep = capi.voxel_downsample(P, resolution, device=device)                       # :89-90
        D = np.ascontiguousarray(P[keep])
        kp, _, _, _ = capi.detect_keypoints(D, neighborhood_radius, 0.65, 20, curvature_non_max_radius, device=device)   # :96-100
        if len(kp) == 0:
            raise capi.GhicpError(-1, f"no keypoints in the {'target' if name == 'T' else 'source'} cloud: check the radii")
        down[name], kp_idx[name] = D, kp
    ext = down["S"].max(axis=0) - down["S"].min(axis=0)                                   # getCloudBound, :91-93
    bbx = float(np.float32(ext[0] + ext[1] + ext[2]))
    Kp = features(down["T"], kp_idx["T"], down["S"], kp_idx["S"])
    Ef = Energyfunction().init(Kp.kps_num, Kp.kpt_num, bbx)
    reg = GHRegistration(Kp, Ef, feature_type, corr_type, curvature_non_max_radius, weight_adjustment_ratio,
                         weight_adjustment_step, dof_type, estimated_IoU, max_iter=max_iter, device=device, **reg_kw)
    Rt, iterations = reg.ghicp_reg()
    info = dict(iterations=iterations, n_target_down=len(down["T"]), n_source_down=len(down["S"]), n_target_kp=len(kp_idx["T"]),
                n_source_kp=len(kp_idx["S"]), bbx_magnitude=bbx, cor=len(reg.pairs()[0]))
    reg.close()
    return Rt, info


def transform_cloud(xyz, Rt):
    """pcl::transformPointCloud with the float32 matrix (test/ghicp_main.cpp:153)."""
    P = np.ascontiguousarray(xyz, dtype=np.float32)
    R, t = Rt[:3, :3].astype(np.float32), Rt[:3, 3].astype(np.float32)
    return P @ R.T + t
