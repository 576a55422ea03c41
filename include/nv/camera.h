// nv/camera.h — the file-format slice of the reference's Camera (libintrinsic3d/include/nv/camera.h, src/camera.cpp:202-274):
// image size, pinhole intrinsics, 5 lens-distortion coefficients (k1 k2 k3 p1 p2) and the intrinsics text file
//   width height \n fx 0 cx \n 0 fy cy \n 0 0 1 \n k1 k2 k3 p1 p2
// Projection itself lives in the engine (csrc/i3d_kernels.cuh, csrc/i3d_math.cuh).
#pragma once
#include <string>

#include <nv/mat.h>

namespace nv
{
class Camera
{
public:
    Camera() { setDefault(); }
    bool load(const std::string& filename);          // false (and default intrinsics 525/525/319.5/239.5) if unreadable
    bool save(const std::string& filename) const;
    int width() const { return width_; }
    int height() const { return height_; }
    void setSize(int w, int h) { width_ = w; height_ = h; }
    Vec4 intrinsicsVec() const { return Vec4{fx_, fy_, cx_, cy_}; }            // fx, fy, cx, cy
    void setIntrinsics(const Vec4& k) { fx_ = k[0]; fy_ = k[1]; cx_ = k[2]; cy_ = k[3]; }
    Vec5 distortion() const { return dist_; }
    void setDistortion(const Vec5& d) { dist_ = d; }

private:
    void setDefault() { width_ = 640; height_ = 480; fx_ = 525.0; fy_ = 525.0; cx_ = 319.5; cy_ = 239.5; dist_ = Vec5::Zero(); }
    int width_, height_;
    double fx_, fy_, cx_, cy_;      // the reference stores float; values round-trip through float in load()/save()
    Vec5 dist_;
};
} // namespace nv
