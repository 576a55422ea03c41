// nv/lighting/lighting_svsh.h — LightingSVSH with the reference's API (libintrinsic3d/include/nv/lighting/lighting_svsh.h:47-70),
// computed by the B200 engine (i3d_estimate_lighting, include/i3d_c_api.h) instead of Ceres.
//
//   LightingSVSH lighting(grid, subvolume_size, lambda_reg, thres_shell, weighted);
//   if (!lighting.estimate()) ...                       // src/refinement/intrinsic3d.cpp:255-262
//   lighting.computeVoxelShCoeffs(data.voxel_sh_coeffs); // :264
//
// estimate() runs the subvolume generation, the joint SH solve AND the per-voxel blend on the device in one call and
// keeps the results on the host; computeVoxelShCoeffs() hands out the already computed vectors (empty VecXd for voxels the
// reference skips: invalid or outside the thin shell).
#pragma once
#include <vector>

#include <nv/lighting/subvolumes.h>
#include <nv/mat.h>
#include <nv/sparse_voxel_grid.h>

namespace nv
{
class LightingSVSH
{
public:
    LightingSVSH(const SparseVoxelGrid<VoxelSBR>* grid, float subvolume_size, double lambda_reg, double thres_shell = 0.0, bool weighted = false);
    ~LightingSVSH();

    bool estimate();
    const Subvolumes& subvolumes() const;
    std::vector<VecXd> shCoeffs() const;
    bool interpolate(const Vec3i& v_pos, VecXd& sh_coeffs) const;
    bool computeVoxelShCoeffs(std::vector<VecXd>& voxel_coeffs) const;

    void setDevice(int cuda_device) { device_ = cuda_device; }
    // ceres::Solver::Summary equivalents of the last estimate()
    int iterations() const { return iterations_; }
    double initialCost() const { return cost_initial_; }
    double finalCost() const { return cost_final_; }

protected:
    const SparseVoxelGrid<VoxelSBR>* grid_;
    float subvolume_size_;
    double thres_shell_;
    bool weighted_;
    double lambda_reg_;
    Subvolumes subvolumes_;
    std::vector<VecXd> sh_coeffs_;
    std::vector<double> voxel_sh_;        // [n][9] blend computed on the device
    std::vector<uint8_t> voxel_has_sh_;   // [n]
    int device_ = 0;
    int iterations_ = 0;
    double cost_initial_ = 0.0, cost_final_ = 0.0;
};
} // namespace nv
