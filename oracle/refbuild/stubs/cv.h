// Minimal stand-in for the OpenCV 1 C header: types the reference declares but only uses in
// commented-out debug code.
#pragma once
#include <cstring>
typedef unsigned char uchar;
struct IplImage {
    int width, height, widthStep;
    char* imageData;
};
static inline void cvReleaseImage(IplImage**) {}
