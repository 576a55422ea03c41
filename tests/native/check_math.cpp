// Host-side harness: compiles intrinsic3d_b200/csrc/i3d_math.cuh with a plain C++ compiler so the
// hand-derived E_g Jacobian row can be checked against the oracle's forward-mode Jets on CPU
// (tests/test_eg_math.py).  Test infrastructure only — never part of the product library.
#define I3D_HD inline
#include "../../intrinsic3d_b200/csrc/i3d_math.cuh"
#include <cmath>
#include <cstdint>

extern "C" int i3dm_eval_eg(const int32_t coord[3], double voxel_size, double pyr_scale, int w, int h, const float* lum,
                            const double sh[9], const double sdf[10], const double alb[4], const double pose[6],
                            const double intr[4], const double dist[5], double* residual, double* jac29_f64, float* jac29_f32)
{
    i3d::PoseCtx<double> pc;
    i3d::pose_ctx_make(pose, &pc);
    i3d::CamParams<double> cam;
    cam.fx = intr[0] * pyr_scale; cam.fy = intr[1] * pyr_scale; cam.cx = intr[2] * pyr_scale; cam.cy = intr[3] * pyr_scale;
    cam.k1 = dist[0]; cam.k2 = dist[1]; cam.k3 = dist[2]; cam.p1 = dist[3]; cam.p2 = dist[4];
    cam.pyr_scale = pyr_scale; cam.w = w; cam.h = h;
    const int c[3] = {coord[0], coord[1], coord[2]};
    double rowd[29]; float rowf[29];
    for (int k = 0; k < 29; ++k) { rowd[k] = 0.0; rowf[k] = 0.f; }
    *residual = i3d::eg_row<double>(sdf, alb, c, voxel_size, pc, cam, lum, sh, rowd);
    i3d::eg_row<float>(sdf, alb, c, voxel_size, pc, cam, lum, sh, rowf);
    for (int k = 0; k < 29; ++k) { jac29_f64[k] = rowd[k]; jac29_f32[k] = rowf[k]; }
    return 0;
}

// Voxel-owned evaluation path of the round-2 kernels (voxel_geom_make + eg_frame_primal + eg_frame_deriv): what k_eg_rows runs.
extern "C" int i3dm_eval_eg_voxel(const int32_t coord[3], double voxel_size, double pyr_scale, int w, int h, const float* lum,
                                  const double sh[9], const double sdf[10], const double alb[4], const double pose[6],
                                  const double intr[4], const double dist[5], double* residual, float* jac29_f32)
{
    i3d::FramePose fp;
    i3d::frame_pose_make(pose, &fp);
    i3d::CamParams<double> cam;
    cam.fx = intr[0] * pyr_scale; cam.fy = intr[1] * pyr_scale; cam.cx = intr[2] * pyr_scale; cam.cy = intr[3] * pyr_scale;
    cam.k1 = dist[0]; cam.k2 = dist[1]; cam.k3 = dist[2]; cam.p1 = dist[3]; cam.p2 = dist[4];
    cam.pyr_scale = pyr_scale; cam.w = w; cam.h = h;
    i3d::CamParams<float> cf;
    cf.fx = float(cam.fx); cf.fy = float(cam.fy); cf.cx = float(cam.cx); cf.cy = float(cam.cy);
    cf.k1 = float(cam.k1); cf.k2 = float(cam.k2); cf.k3 = float(cam.k3); cf.p1 = float(cam.p1); cf.p2 = float(cam.p2);
    cf.pyr_scale = float(pyr_scale); cf.w = w; cf.h = h;
    const int c[3] = {coord[0], coord[1], coord[2]};
    i3d::VoxelGeom vg; i3d::VoxelDeriv vd;
    i3d::voxel_geom_make<true>(sdf, alb, c, voxel_size, sh, &vg, &vd);
    i3d::PointSave sv[4]; float e[4] = {0, 0, 0, 0};
    const double r = i3d::eg_frame_primal<true>(vg, fp, cam, i3d::LinearImage{lum}, sv, e);
    *residual = r;
    for (int k = 0; k < 29; ++k) jac29_f32[k] = 0.f;
    if (r != 0.0) i3d::eg_frame_deriv(vd, fp, cf, sv, e, jac29_f32);
    // the cost-only instantiation must give the same residual
    i3d::VoxelGeom vg2;
    i3d::voxel_geom_make<false>(sdf, alb, c, voxel_size, sh, &vg2, nullptr);
    const double r2 = i3d::eg_frame_primal<false>(vg2, fp, cam, i3d::LinearImage{lum}, nullptr, nullptr);
    return r2 == r ? 0 : 1;
}


// The same row through the shared-memory VIEWS of the per-voxel state (VoxelGeomView / VoxelDerivView over parked columns): what k_eg_rows
// runs since round 2.  Must be bit-identical to i3dm_eval_eg_voxel (same arithmetic, the state only moves through memory).
extern "C" int i3dm_eval_eg_voxel_views(const int32_t coord[3], double voxel_size, double pyr_scale, int w, int h, const float* lum,
                                        const double sh[9], const double sdf[10], const double alb[4], const double pose[6],
                                        const double intr[4], const double dist[5], double* residual, float* jac29_f32)
{
    i3d::FramePose fp;
    i3d::frame_pose_make(pose, &fp);
    i3d::CamParams<double> cam;
    cam.fx = intr[0] * pyr_scale; cam.fy = intr[1] * pyr_scale; cam.cx = intr[2] * pyr_scale; cam.cy = intr[3] * pyr_scale;
    cam.k1 = dist[0]; cam.k2 = dist[1]; cam.k3 = dist[2]; cam.p1 = dist[3]; cam.p2 = dist[4];
    cam.pyr_scale = pyr_scale; cam.w = w; cam.h = h;
    i3d::CamParams<float> cf;
    cf.fx = float(cam.fx); cf.fy = float(cam.fy); cf.cx = float(cam.cx); cf.cy = float(cam.cy);
    cf.k1 = float(cam.k1); cf.k2 = float(cam.k2); cf.k3 = float(cam.k3); cf.p1 = float(cam.p1); cf.p2 = float(cam.p2);
    cf.pyr_scale = float(pyr_scale); cf.w = w; cf.h = h;
    const int c[3] = {coord[0], coord[1], coord[2]};
    i3d::VoxelGeom vg; i3d::VoxelDeriv vd;
    i3d::voxel_geom_make<true>(sdf, alb, c, voxel_size, sh, &vg, &vd);
    // parked with a stride of 7 "threads", this voxel in column 3
    const int stride = 7, col = 3;
    double pg[i3d::kVoxelGeomWords * 7];
    float pd[i3d::kVoxelDerivWords * 7];
    for (double& x : pg) x = -1e300;
    for (float& x : pd) x = -1e30f;
    i3d::voxel_geom_park(vg, pg + col, stride);
    i3d::voxel_deriv_park(vd, pd + col, stride);
    const i3d::VoxelGeomView vgv{pg + col, stride};
    const i3d::VoxelDerivView vdv{pd + col, stride};
    i3d::PointSave sv[4]; float e[4] = {0, 0, 0, 0};
    const double r = i3d::eg_frame_primal<true>(vgv, fp, cam, i3d::LinearImage{lum}, sv, e);
    *residual = r;
    for (int k = 0; k < 29; ++k) jac29_f32[k] = 0.f;
    if (r != 0.0) i3d::eg_frame_deriv(vdv, fp, cf, sv, e, jac29_f32);
    return 0;
}
