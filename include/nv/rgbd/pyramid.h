// nv/rgbd/pyramid.h — per-keyframe image pyramid as the optimiser sees it (reference: include/nv/rgbd/pyramid.h:47-69).
// Building pyramids (cv::pyrDown etc.) is image preparation and out of scope; the caller attaches float luminance and depth
// images per level.
#pragma once
#include <vector>

#include <nv/image.h>

namespace nv
{
class Pyramid
{
public:
    void addLevel(const ImageF& intensity, const ImageF& depth) { intensity_.push_back(intensity); depth_.push_back(depth); color_.push_back(ImageBGR()); }
    void addLevel(const ImageF& intensity, const ImageF& depth, const ImageBGR& color) { intensity_.push_back(intensity); depth_.push_back(depth); color_.push_back(color); }
    int levels() const { return static_cast<int>(intensity_.size()); }
    ImageF intensity(int lvl) const { return intensity_[static_cast<size_t>(lvl)]; }
    ImageF depth(int lvl) const { return depth_[static_cast<size_t>(lvl)]; }
    ImageBGR color(int lvl) const { return color_[static_cast<size_t>(lvl)]; }

private:
    std::vector<ImageF> intensity_, depth_;
    std::vector<ImageBGR> color_;
};
} // namespace nv
