// Minimal stand-in for cv_bridge: the reference only reads image.rows/cols/step/data
// (MatchGPULib.cpp:315-323).
#pragma once
#include <cstddef>
#include <memory>
namespace cv_bridge {
struct Mat {
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;
};
struct CvImage {
    Mat image;
};
typedef std::shared_ptr<CvImage> CvImagePtr;
}  // namespace cv_bridge
