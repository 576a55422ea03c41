// Host-side harness: compiles intrinsic3d_b200/csrc/i3d_math.cuh with a plain C++ compiler so the
// hand-derived E_g Jacobian row can be checked against the oracle's forward-mode Jets on CPU
// (tests/test_eg_math.py).  Test infrastructure only — never part of the product library.
#define I3D_HD inline
#include "../../intrinsic3d_b200/csrc/i3d_math.cuh"
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
