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
// 8-bit colour image view, interleaved B,G,R like the reference's cv::Mat (CV_8UC3) from Pyramid::color()
struct ImageBGR
{
    int rows = 0, cols = 0;
    const unsigned char* data = nullptr;
    bool empty() const { return data == nullptr || rows <= 0 || cols <= 0; }
};
} // namespace nv
