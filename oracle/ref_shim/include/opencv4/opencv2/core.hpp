// stub OpenCV core for compiling the reference's FastDetector.cc / BriefDescriptor.cc / Image.cc unmodified:
// the container types come from the product's cv shim; the three OpenCV functions the hot path calls
// are declared here and implemented in ../../cv_funcs.cpp on top of the (cv2-pinned) oracle.
#pragma once
#include <climits>
#include <memory>
#include <string>

#include "../../../../../ya_vo_b200/host/cvshim/opencv2/core.hpp"
#include "../../../../../ya_vo_b200/host/cvshim/opencv2/imgproc.hpp"

namespace cv {
void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY = 0);
bool eigen(const Mat &src, Mat &eigenvalues);
void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int borderType,
                    const Scalar &value = Scalar());
}  // namespace cv
