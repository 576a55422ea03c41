// nv/math.h — pose conversions of the refinement path (libintrinsic3d/src/math.cpp:151-178) and the TUM RGB-D trajectory file
// format the reference reads / writes (src/rgbd/sensor.cpp:236-347): "timestamp tx ty tz qx qy qz qw", camera-to-world.
#pragma once
#include <string>
#include <vector>

#include <nv/mat.h>

namespace nv
{
namespace math
{
// pose vector = (angle-axis rotation, translation); Eigen::AngleAxisd(|w|, w/|w|).matrix()
Mat4 poseVecAAToMat(const Vec6& pose_vec_aa);
// inverse: Eigen::AngleAxisd(R): angle in [0, pi], axis normalised
Vec6 poseMatToVecAA(const Mat4& pose);
// rigid inverse
Mat4 invertPose(const Mat4& pose);
} // namespace math

bool loadPoses(const std::string& filename, std::vector<Mat4f>& poses_cam_to_world, std::vector<double>& timestamps, bool first_pose_is_identity = false);
bool savePoses(const std::string& filename, const std::vector<Mat4f>& poses_cam_to_world, const std::vector<double>& timestamps);
} // namespace nv
