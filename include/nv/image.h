// nv/image.h — non-owning float image view; stands in for the cv::Mat (CV_32FC1) the reference passes around
// (OpenCV C++ is not available here).  rows/cols/ptr mirror the cv::Mat members the hot path touches.
#pragma once
namespace nv
{
struct ImageF
{
    int rows = 0, cols = 0;
    const float* data = nullptr;
    bool empty() const { return data == nullptr || rows <= 0 || cols <= 0; }
};
} // namespace nv
