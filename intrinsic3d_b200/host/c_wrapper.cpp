// Test hook: runs nv::Optimizer::optimize (the reference-shaped host API) on flat arrays, so the Python parity tests can
// drive the C++ shim.  Not part of the drop-in surface.
#include <algorithm>
#include <cmath>
#include <cstring>

#include <nv/refinement/albedo_regularizer.h>
#include <nv/refinement/optimizer.h>
#include <nv/refinement/surface_stab_regularizer.h>
#include <nv/refinement/volumetric_regularizer.h>

extern "C" int i3dh_run_optimizer(int64_t n, const int32_t* xyz, const double* sdf0, double* sdf_refined, double* albedo, const float* weight, const uint8_t* rgb,
                                  float voxel_size, int32_t F, int32_t W, int32_t H, const float* lum, const float* depth, double* poses, double* intr, double* dist,
                                  const double* sh9n, double thres_shell, float occlusion, int32_t num_obs, int32_t iterations, int32_t lm_steps,
                                  const double* lambdas /* g, r0, r1, s0, s1, a */, int32_t fix_poses, int32_t fix_intr, int32_t fix_dist,
                                  int64_t* counts_out /* [4]: applicable E_g stencils, E_r, E_s, E_a creates via the plugin API on the first 2000 voxels */)
{
    using namespace nv;
    SparseVoxelGrid<VoxelSBR>* grid = SparseVoxelGrid<VoxelSBR>::create(voxel_size);
    grid->reserve(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i)
    {
        VoxelSBR v;
        v.sdf = sdf0[i]; v.sdf_refined = sdf_refined[i]; v.albedo = albedo[i]; v.weight = weight[i];
        v.color = Vec3b{rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
        grid->insert(Vec3i{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, v);
    }
    Optimizer::Config cfg;
    cfg.iterations = iterations; cfg.lm_steps = lm_steps;
    cfg.lambda_g = lambdas[0]; cfg.lambda_r0 = lambdas[1]; cfg.lambda_r1 = lambdas[2]; cfg.lambda_s0 = lambdas[3]; cfg.lambda_s1 = lambdas[4]; cfg.lambda_a = lambdas[5];
    cfg.fix_poses = fix_poses; cfg.fix_intrinsics = fix_intr; cfg.fix_distortion = fix_dist;
    Optimizer::Data data;
    data.grid = grid; data.thres_shell = thres_shell; data.grid_level = 0; data.rgbd_level = 0;
    data.voxel_sh_coeffs.resize(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) data.voxel_sh_coeffs[i].assign(sh9n + 9 * i, sh9n + 9 * i + 9);
    Optimizer::ImageFormationModel im;
    for (int k = 0; k < 4; ++k) im.intrinsics[k] = intr[k];
    for (int k = 0; k < 5; ++k) im.distortion_coeffs[k] = dist[k];
    im.poses.resize(F); im.rgbd_pyr.resize(F);
    for (int f = 0; f < F; ++f)
    {
        for (int k = 0; k < 6; ++k) im.poses[f][k] = poses[6 * f + k];
        im.frame_ids.push_back(f);
        im.rgbd_pyr[f].addLevel(ImageF{H, W, lum + static_cast<size_t>(f) * W * H}, ImageF{H, W, depth + static_cast<size_t>(f) * W * H});
        data.shading_cost_data.emplace_back(0, static_cast<double>(voxel_size), W, H, lum + static_cast<size_t>(f) * W * H);
    }
    // exercise the cost-term plugin signatures (applicability only) on a prefix of the grid
    if (counts_out)
    {
        std::memset(counts_out, 0, 4 * sizeof(int64_t));
        int64_t seen = 0;
        for (auto it = grid->begin(); it != grid->end() && seen < 2000; ++it, ++seen)
        {
            const Vec3i& p = it->first;
            VoxelResidual a = ShadingCost::create(grid, p, im.poses[0], im.intrinsics, im.distortion_coeffs, data.voxel_sh_coeffs[seen], &data.shading_cost_data[0]);
            if (a.cost) { counts_out[0]++; delete a.cost; }
            VoxelResidual b = VolumetricRegularizer::create(grid, p);
            if (b.weight > 0.0) { counts_out[1]++; delete b.cost; }
            VoxelResidual c = SurfaceStabRegularizer::create(grid, p);
            if (c.weight > 0.0) { counts_out[2]++; delete c.cost; }
            VoxelResidual d = AlbedoRegularizer::create(grid, p, Vec3i{p[0] + 1, p[1], p[2]});
            if (d.cost) { counts_out[3]++; delete d.cost; }
        }
    }
    SDFColorization::Config ccfg;
    ccfg.max_occlusion_distance = occlusion; ccfg.max_num_observations = static_cast<size_t>(num_obs);
    SDFColorization colorization(grid);
    colorization.setConfig(ccfg);
    Optimizer opt(cfg);
    const bool ok = opt.optimize(colorization, data, im);
    if (ok)
    {
        int64_t i = 0;
        for (auto it = grid->begin(); it != grid->end(); ++it, ++i) { sdf_refined[i] = it->second.sdf_refined; albedo[i] = it->second.albedo; }
        for (int f = 0; f < F; ++f) for (int k = 0; k < 6; ++k) poses[6 * f + k] = im.poses[f][k];
        for (int k = 0; k < 4; ++k) intr[k] = im.intrinsics[k];
        for (int k = 0; k < 5; ++k) dist[k] = im.distortion_coeffs[k];
    }
    delete grid;
    return ok ? 0 : 1;
}

#include "../../include/i3d_c_api.h"

// Test hook: the NLSSolver::addResidual contract.  The engine always solves the complete problem of the attached grid, so
//   bit 0: nothing recorded            -> buildProblem(true) succeeds
//   bit 1: one E_r residual recorded   -> buildProblem(true) FAILS (a subset cannot be honoured; reference contract nls_solver.cpp:172-187)
//   bit 2: a residual with weight 0 or cost == nullptr is refused by addResidual() like the reference does
//   bit 3: a descriptor whose term type does not match the cost id is refused
extern "C" int i3dh_nls_contract(int64_t n, const int32_t* xyz, const double* sdf0, const double* sdf_refined, const double* albedo, const float* weight, const uint8_t* rgb,
                                 float voxel_size, int32_t F, int32_t W, int32_t H, const float* lum, const float* depth, const double* poses, const double* intr,
                                 const double* dist, const double* sh9n, double thres_shell)
{
    using namespace nv;
    I3DEngine* eng = nullptr;
    if (i3d_engine_create(0, &eng) != 0) return -1;
    int result = -1;
    SparseVoxelGrid<VoxelSBR>* grid = SparseVoxelGrid<VoxelSBR>::create(voxel_size);
    do
    {
        if (i3d_upload_grid(eng, n, xyz, sdf0, sdf_refined, albedo, weight, rgb, voxel_size) != 0) break;
        if (i3d_upload_frames(eng, F, W, H, lum, depth, 1.0) != 0) break;
        if (i3d_set_camera(eng, poses, intr, dist) != 0) break;
        if (i3d_set_sh(eng, sh9n) != 0) break;
        Vec3i ring_voxel{0, 0, 0};
        bool have_ring = false;
        for (int64_t i = 0; i < n; ++i)
        {
            VoxelSBR v;
            v.sdf = sdf0[i]; v.sdf_refined = sdf_refined[i]; v.albedo = albedo[i]; v.weight = weight[i];
            grid->insert(Vec3i{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, v);
        }
        for (auto it = grid->begin(); it != grid->end() && !have_ring; ++it)
        {
            VoxelResidual r = VolumetricRegularizer::create(grid, it->first);
            if (r.cost) { have_ring = true; ring_voxel = it->first; delete r.cost; }
        }
        if (!have_ring) break;
        NLSSolver::Binding b;
        b.engine = eng; b.thres_shell = thres_shell;
        result = 0;
        {
            NLSSolver s; s.reset(4); s.attach(b);
            for (int t = 0; t < 4; ++t) s.setCostWeight(t, t == 0 ? 0.2 : 10.0);
            if (s.buildProblem(true)) result |= 1;
        }
        {
            NLSSolver s; s.reset(4); s.attach(b);
            for (int t = 0; t < 4; ++t) s.setCostWeight(t, t == 0 ? 0.2 : 10.0);
            VoxelResidual r = VolumetricRegularizer::create(grid, ring_voxel);
            const bool added = s.addResidual(1, r);
            if (added && !s.buildProblem(true)) result |= 2;
        }
        {
            NLSSolver s; s.reset(4);
            VoxelResidual none;                                         // cost == nullptr, weight == 0
            VoxelResidual zero_w = VolumetricRegularizer::create(grid, ring_voxel);
            zero_w.weight = 0.0;
            if (!s.addResidual(1, none) && !s.addResidual(1, zero_w)) result |= 4;
            VoxelResidual wrong = VolumetricRegularizer::create(grid, ring_voxel);
            if (!s.addResidual(2, wrong)) result |= 8;               // an E_r descriptor under the E_s cost id
        }
    } while (false);
    delete grid;
    i3d_engine_destroy(eng);
    return result;
}

#include <nv/lighting/lighting_svsh.h>

// Test hook: nv::LightingSVSH (estimate + computeVoxelShCoeffs + interpolate) on flat arrays.
// sub_index3 / sub_sh9 have room for `sub_capacity` subvolumes; voxel_sh9n [n][9] (zeros where the reference leaves an empty
// vector), has_sh [n]; interp_err_out = max |LightingSVSH::interpolate(v) - computeVoxelShCoeffs()[v]| over the voxels that have one.
extern "C" int i3dh_run_lighting(int64_t n, const int32_t* xyz, const double* sdf0, const double* sdf_refined, const double* albedo, const float* weight,
                                 const uint8_t* rgb, float voxel_size, float subvolume_size, double lambda_reg, double thres_shell, int32_t weighted,
                                 int64_t sub_capacity, int64_t* num_subvolumes_out, int32_t* sub_index3, double* sub_sh9, double* voxel_sh9n, uint8_t* has_sh,
                                 double* interp_err_out)
{
    using namespace nv;
    SparseVoxelGrid<VoxelSBR>* grid = SparseVoxelGrid<VoxelSBR>::create(voxel_size);
    grid->reserve(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i)
    {
        VoxelSBR v;
        v.sdf = sdf0[i]; v.sdf_refined = sdf_refined[i]; v.albedo = albedo[i]; v.weight = weight[i];
        v.color = Vec3b{rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
        grid->insert(Vec3i{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, v);
    }
    LightingSVSH lighting(grid, subvolume_size, lambda_reg, thres_shell, weighted != 0);
    int rc = 1;
    if (lighting.estimate())
    {
        std::vector<VecXd> voxel_coeffs;
        const std::vector<VecXd> sh = lighting.shCoeffs();
        const Subvolumes& sub = lighting.subvolumes();
        *num_subvolumes_out = static_cast<int64_t>(sub.count());
        if (lighting.computeVoxelShCoeffs(voxel_coeffs) && static_cast<int64_t>(sub.count()) <= sub_capacity && sh.size() == sub.count())
        {
            for (size_t s = 0; s < sub.count(); ++s)
            {
                const Vec3i idx = sub.index(static_cast<int>(s));
                for (int d = 0; d < 3; ++d) sub_index3[3 * s + d] = idx[d];
                for (int k = 0; k < 9; ++k) sub_sh9[9 * s + k] = sh[s][k];
            }
            double err = 0.0;
            int64_t i = 0;
            for (auto it = grid->begin(); it != grid->end(); ++it, ++i)
            {
                has_sh[i] = voxel_coeffs[i].empty() ? 0 : 1;
                for (int k = 0; k < 9; ++k) voxel_sh9n[9 * i + k] = voxel_coeffs[i].empty() ? 0.0 : voxel_coeffs[i][k];
                if (!voxel_coeffs[i].empty())
                {
                    VecXd c;
                    if (!lighting.interpolate(it->first, c) || c.size() != 9) { err = 1e300; continue; }
                    for (int k = 0; k < 9; ++k) err = std::max(err, std::fabs(c[k] - voxel_coeffs[i][k]));
                }
            }
            *interp_err_out = err;
            rc = 0;
        }
    }
    delete grid;
    return rc;
}

// Test hook: nv::SDFColorization::reset / add (per frame) / compute on flat arrays, the way Intrinsic3D::recomputeColors drives them.
// pose_rt: [F][12] floats (rotation row-major, translation); rgb is updated in place.
extern "C" int i3dh_run_recolor(int64_t n, const int32_t* xyz, const double* sdf0, const double* sdf_refined, const double* albedo, const float* weight,
                                uint8_t* rgb, float voxel_size, int32_t F, int32_t W, int32_t H, const float* depth, const uint8_t* bgr, const float* pose_rt,
                                const double* intr, const double* dist, float occlusion, int32_t num_obs)
{
    using namespace nv;
    SparseVoxelGrid<VoxelSBR>* grid = SparseVoxelGrid<VoxelSBR>::create(voxel_size);
    grid->reserve(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i)
    {
        VoxelSBR v;
        v.sdf = sdf0[i]; v.sdf_refined = sdf_refined[i]; v.albedo = albedo[i]; v.weight = weight[i];
        v.color = Vec3b{rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
        grid->insert(Vec3i{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, v);
    }
    Vec4 K; Vec5 D;
    for (int k = 0; k < 4; ++k) K[k] = intr[k];
    for (int k = 0; k < 5; ++k) D[k] = dist[k];
    SDFColorization::Config cfg;
    cfg.max_occlusion_distance = occlusion; cfg.max_num_observations = static_cast<size_t>(num_obs);
    SDFColorization col(grid);
    col.setConfig(cfg);
    bool ok = col.reset(grid, K, D, W, H);
    const size_t px = static_cast<size_t>(W) * H;
    for (int f = 0; f < F && ok; ++f)
    {
        Mat4f P;
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) P(r, c) = pose_rt[12 * f + 3 * r + c]; P(r, 3) = pose_rt[12 * f + 9 + r]; }
        ok = col.add(f, ImageF{H, W, depth + f * px}, ImageBGR{H, W, bgr + f * px * 3}, P);
    }
    ok = ok && col.compute();
    if (ok)
    {
        int64_t i = 0;
        for (auto it = grid->begin(); it != grid->end(); ++it, ++i) for (int k = 0; k < 3; ++k) rgb[3 * i + k] = it->second.color[k];
    }
    delete grid;
    return ok ? 0 : 1;
}

#include <nv/sdf/algorithms.h>

// Test hook: SDFAlgorithms::clearVoxelsOutsideThinShell (op 0) / upsample (op 1) on flat arrays.  The output arrays have room for
// `capacity` voxels; *n_out receives the new count, *voxel_size_out the new voxel size.
extern "C" int i3dh_run_gridop(int32_t op, int64_t n, const int32_t* xyz, const double* sdf0, const double* sdf_refined, const double* albedo, const float* weight,
                               const uint8_t* rgb, float voxel_size, double thres_shell, int64_t capacity, int64_t* n_out, int32_t* xyz_out, double* sdf0_out,
                               double* sdf_out, double* albedo_out, float* weight_out, uint8_t* rgb_out, float* voxel_size_out)
{
    using namespace nv;
    SparseVoxelGrid<VoxelSBR>* grid = SparseVoxelGrid<VoxelSBR>::create(voxel_size);
    grid->reserve(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i)
    {
        VoxelSBR v;
        v.sdf = sdf0[i]; v.sdf_refined = sdf_refined[i]; v.albedo = albedo[i]; v.weight = weight[i];
        v.color = Vec3b{rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
        grid->insert(Vec3i{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, v);
    }
    SparseVoxelGrid<VoxelSBR>* res = grid;
    if (op == 0) SDFAlgorithms::clearVoxelsOutsideThinShell(grid, thres_shell);
    else res = SDFAlgorithms::upsample(grid);
    int rc = 1;
    if (res && static_cast<int64_t>(res->numVoxels()) <= capacity)
    {
        int64_t i = 0;
        for (auto it = res->begin(); it != res->end(); ++it, ++i)
        {
            for (int d = 0; d < 3; ++d) { xyz_out[3 * i + d] = it->first[d]; rgb_out[3 * i + d] = it->second.color[d]; }
            sdf0_out[i] = it->second.sdf; sdf_out[i] = it->second.sdf_refined; albedo_out[i] = it->second.albedo; weight_out[i] = it->second.weight;
        }
        *n_out = i; *voxel_size_out = res->voxelSize();
        rc = 0;
    }
    if (res != grid) delete res;
    delete grid;
    return rc;
}

#include <nv/camera.h>
#include <nv/math.h>
#include <nv/refinement/intrinsic3d.h>

// Test hook: nv::Intrinsic3D::refine on flat arrays.  Pyramid level l of frame f: lum/depth planes at lum_lvl[l] + f * W[l] * H[l];
// colour (B,G,R) only for level 0.  cfg = {num_grid_levels, num_rgbd_levels, thres_shell_factor, thres_shell_factor_final,
// clear_distant_voxels, occlusion, num_observations, subvolume_size, sh_lambda_reg, iterations, lm_steps, lambda_g, r0, r1, s0, s1, a}.
// callbacks_out counts RefinementCallback invocations.
extern "C" int i3dh_run_refine(int64_t n, const int32_t* xyz, const float* sdf, const float* weight, const uint8_t* rgb, float voxel_size, int32_t F, int32_t L,
                               const int32_t* Wl, const int32_t* Hl, const float* const* lum_lvl, const float* const* depth_lvl, const uint8_t* bgr0, double* poses,
                               double* intr, double* dist, const double* cfg, int64_t capacity, int64_t* n_out, int32_t* xyz_out, double* sdf0_out, double* sdf_out,
                               double* albedo_out, float* weight_out, uint8_t* rgb_out, float* voxel_size_out, int32_t* callbacks_out)
{
    using namespace nv;
    SparseVoxelGrid<Voxel>* grid = SparseVoxelGrid<Voxel>::create(voxel_size);
    grid->reserve(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i)
    {
        Voxel v;
        v.sdf = sdf[i]; v.weight = weight[i]; v.color = Vec3b{rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
        grid->insert(Vec3i{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, v);
    }
    Optimizer::ImageFormationModel im;
    for (int k = 0; k < 4; ++k) im.intrinsics[k] = intr[k];
    im.poses.resize(F); im.rgbd_pyr.resize(F);
    for (int f = 0; f < F; ++f)
    {
        for (int k = 0; k < 6; ++k) im.poses[f][k] = poses[6 * f + k];
        im.frame_ids.push_back(f);
        for (int l = 0; l < L; ++l)
        {
            const size_t px = static_cast<size_t>(Wl[l]) * Hl[l];
            const ImageF li{Hl[l], Wl[l], lum_lvl[l] + f * px}, di{Hl[l], Wl[l], depth_lvl[l] + f * px};
            if (l == 0) im.rgbd_pyr[f].addLevel(li, di, ImageBGR{Hl[0], Wl[0], bgr0 + f * px * 3});
            else im.rgbd_pyr[f].addLevel(li, di);
        }
    }
    Intrinsic3D::Config c;
    c.num_grid_levels = static_cast<int>(cfg[0]); c.num_rgbd_levels = static_cast<int>(cfg[1]); c.thres_shell_factor = cfg[2]; c.thres_shell_factor_final = cfg[3];
    c.clear_distant_voxels = cfg[4] != 0.0; c.occlusions_distance = static_cast<float>(cfg[5]); c.num_observations = static_cast<size_t>(cfg[6]);
    c.subvolume_size_sh = static_cast<float>(cfg[7]); c.sh_est_lambda_reg = cfg[8];
    Optimizer::Config oc;
    oc.iterations = static_cast<int>(cfg[9]); oc.lm_steps = static_cast<int>(cfg[10]); oc.lambda_g = cfg[11]; oc.lambda_r0 = cfg[12]; oc.lambda_r1 = cfg[13];
    oc.lambda_s0 = cfg[14]; oc.lambda_s1 = cfg[15]; oc.lambda_a = cfg[16];
    struct Counter : Intrinsic3D::RefinementCallback
    {
        int calls = 0; size_t last_voxels = 0;
        void onSDFRefined(const Intrinsic3D::RefinementInfo& info) override { ++calls; last_voxels = info.grid ? info.grid->numVoxels() : 0; }
    } counter;
    Intrinsic3D app(c, oc, &im);
    app.addRefinementCallback(&counter);
    const bool ok = app.refine(grid);
    int rc = 1;
    SparseVoxelGrid<VoxelSBR>* res = app.refinedGrid();
    if (ok && res && static_cast<int64_t>(res->numVoxels()) <= capacity)
    {
        int64_t i = 0;
        for (auto it = res->begin(); it != res->end(); ++it, ++i)
        {
            for (int d = 0; d < 3; ++d) { xyz_out[3 * i + d] = it->first[d]; rgb_out[3 * i + d] = it->second.color[d]; }
            sdf0_out[i] = it->second.sdf; sdf_out[i] = it->second.sdf_refined; albedo_out[i] = it->second.albedo; weight_out[i] = it->second.weight;
        }
        *n_out = i; *voxel_size_out = res->voxelSize(); *callbacks_out = counter.calls;
        for (int f = 0; f < F; ++f) for (int k = 0; k < 6; ++k) poses[6 * f + k] = im.poses[f][k];
        for (int k = 0; k < 4; ++k) intr[k] = im.intrinsics[k];
        for (int k = 0; k < 5; ++k) dist[k] = im.distortion_coeffs[k];
        rc = 0;
    }
    delete grid;
    return rc;
}

// ---- file-format test hooks (CPU only) ----
// writes a Voxel grid to `path` with SparseVoxelGrid<Voxel>::save, loads it back with load(), returns 0 when identical; also converts
// (SDFAlgorithms::convert) and round-trips the VoxelSBR grid through `path_sbr`.
extern "C" int i3dh_io_tsdf_roundtrip(const char* path, const char* path_sbr, int64_t n, const int32_t* xyz, const float* sdf, const float* weight, const uint8_t* rgb,
                                      float voxel_size, int64_t* n_valid_out)
{
    using namespace nv;
    SparseVoxelGrid<Voxel>* g = SparseVoxelGrid<Voxel>::create(voxel_size);
    for (int64_t i = 0; i < n; ++i)
    {
        Voxel v; v.sdf = sdf[i]; v.weight = weight[i]; v.color = Vec3b{rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
        g->insert(Vec3i{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, v);
    }
    int rc = 0;
    if (!g->save(path)) rc = 1;
    SparseVoxelGrid<Voxel>* h = SparseVoxelGrid<Voxel>::create(1.0f);
    if (!rc && !h->load(path)) rc = 2;
    if (!rc && (h->numVoxels() != g->numVoxels() || h->voxelSize() != g->voxelSize() || h->truncation() != g->truncation())) rc = 3;
    if (!rc)
    {
        auto a = g->begin(); auto b = h->begin();
        for (; a != g->end(); ++a, ++b)
            if (!(a->first == b->first) || a->second.sdf != b->second.sdf || a->second.weight != b->second.weight || !(a->second.color == b->second.color)) { rc = 4; break; }
    }
    SparseVoxelGrid<VoxelSBR>* s = SDFAlgorithms::convert(g);
    *n_valid_out = s ? static_cast<int64_t>(s->numVoxels()) : -1;
    if (!rc && (!s || !s->save(path_sbr))) rc = 5;
    SparseVoxelGrid<VoxelSBR>* t = SparseVoxelGrid<VoxelSBR>::create(1.0f);
    if (!rc && (!t->load(path_sbr) || t->numVoxels() != s->numVoxels())) rc = 6;
    if (!rc)
    {
        auto a = s->begin(); auto b = t->begin();
        for (; a != s->end(); ++a, ++b)
            if (!(a->first == b->first) || a->second.sdf != b->second.sdf || a->second.sdf_refined != b->second.sdf_refined || a->second.albedo != b->second.albedo ||
                a->second.weight != b->second.weight || !(a->second.color == b->second.color)) { rc = 7; break; }
    }
    delete g; delete h; delete s; delete t;
    return rc;
}

// loads a .tsdf of Voxel records written by someone else (the test writes one with numpy in the reference's byte layout)
extern "C" int i3dh_io_tsdf_load(const char* path, int64_t capacity, int64_t* n_out, float* header3, int32_t* xyz, float* sdf, float* weight, uint8_t* rgb)
{
    using namespace nv;
    SparseVoxelGrid<Voxel>* g = SparseVoxelGrid<Voxel>::create(1.0f);
    int rc = g->load(path) ? 0 : 1;
    if (!rc && static_cast<int64_t>(g->numVoxels()) > capacity) rc = 2;
    if (!rc)
    {
        int64_t i = 0;
        for (auto it = g->begin(); it != g->end(); ++it, ++i)
        {
            for (int d = 0; d < 3; ++d) { xyz[3 * i + d] = it->first[d]; rgb[3 * i + d] = it->second.color[d]; }
            sdf[i] = it->second.sdf; weight[i] = it->second.weight;
        }
        *n_out = i; header3[0] = g->voxelSize(); header3[1] = g->truncation(); header3[2] = 0.0f;
    }
    delete g;
    return rc;
}

// intrinsics file + TUM poses round trip and pose-vector conversions; out[0..3] intrinsics, out[4..8] distortion, out[9..10] size,
// out[11] = max |poseMatToVecAA(poseVecAAToMat(v)) - v| over `poses6`, out[12] = max |R R^T - I| entry, out[13] = max abs pose error after the file round trip
extern "C" int i3dh_io_camera_poses(const char* intr_path, const char* poses_path, int32_t F, const double* poses6, double* out)
{
    using namespace nv;
    Camera cam;
    if (!cam.load(intr_path)) return 1;
    const Vec4 k = cam.intrinsicsVec(); const Vec5 d = cam.distortion();
    for (int i = 0; i < 4; ++i) out[i] = k[i];
    for (int i = 0; i < 5; ++i) out[4 + i] = d[i];
    out[9] = cam.width(); out[10] = cam.height();
    std::string copy = std::string(intr_path) + ".copy";
    if (!cam.save(copy)) return 2;
    Camera cam2;
    if (!cam2.load(copy)) return 3;
    for (int i = 0; i < 4; ++i) if (cam2.intrinsicsVec()[i] != k[i]) return 4;
    double e_aa = 0.0, e_orth = 0.0, e_file = 0.0;
    std::vector<Mat4f> cam_to_world; std::vector<double> ts;
    for (int f = 0; f < F; ++f)
    {
        Vec6 v; for (int i = 0; i < 6; ++i) v[i] = poses6[6 * f + i];
        const Mat4 M = math::poseVecAAToMat(v);
        const Vec6 w = math::poseMatToVecAA(M);
        for (int i = 0; i < 6; ++i) e_aa = std::max(e_aa, std::fabs(w[i] - v[i]));
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
        {
            double s = 0.0; for (int j = 0; j < 3; ++j) s += M(r, j) * M(c, j);
            e_orth = std::max(e_orth, std::fabs(s - (r == c ? 1.0 : 0.0)));
        }
        const Mat4 I = math::invertPose(M);            // camera-to-world, what the trajectory file stores
        Mat4f If; for (int i = 0; i < 16; ++i) If.m[i] = static_cast<float>(I.m[i]);
        cam_to_world.push_back(If); ts.push_back(1000.0 + 0.033 * f);
    }
    if (!savePoses(poses_path, cam_to_world, ts)) return 5;
    std::vector<Mat4f> back; std::vector<double> ts2;
    if (!loadPoses(poses_path, back, ts2, false) || static_cast<int>(back.size()) != F) return 6;
    for (int f = 0; f < F; ++f)
    {
        for (int i = 0; i < 12; ++i) e_file = std::max(e_file, static_cast<double>(std::fabs(back[f].m[i] - cam_to_world[f].m[i])));
        if (std::fabs(ts2[f] - ts[f]) > 1e-6) return 7;
    }
    out[11] = e_aa; out[12] = e_orth; out[13] = e_file;
    return 0;
}
