// The three OpenCV functions the reference's hot path calls, for the stub build (TEST INFRASTRUCTURE).
// GaussianBlur and eigen forward to the oracle's restatements (pinned against cv2 4.13.0).
#include <stdexcept>

#include "../yavo_oracle.h"
#include "include/opencv4/opencv2/core.hpp"

namespace cv {

void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY) {
    if (src.type() != CV_8UC1 || ksize.width != 9 || ksize.height != 9 || sigmaX != 2.5 || sigmaY != 2.5)
        throw std::runtime_error("stub GaussianBlur supports only the reference's call: CV_8UC1, 9x9, sigma 2.5");
    Mat in = src.clone();  // continuous
    Mat out(src.rows, src.cols, CV_8UC1);
    yavo_oracle_gaussian_blur(in.data, src.rows, src.cols, out.data);
    dst = out;
}

bool eigen(const Mat &src, Mat &eigenvalues) {
    if (src.type() != CV_32FC1 || src.rows != 2 || src.cols != 2)
        throw std::runtime_error("stub eigen supports only the reference's call: 2x2 CV_32FC1");
    float l1, l2;
    yavo_oracle_eigen2x2(src.at<float>(0, 0), src.at<float>(0, 1), src.at<float>(1, 1), &l1, &l2);
    Mat ev(2, 1, CV_32FC1);
    ev.at<float>(0, 0) = l1;
    ev.at<float>(1, 0) = l2;
    eigenvalues = ev;
    return true;
}

void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int borderType,
                    const Scalar &value) {
    if (borderType != BORDER_CONSTANT) throw std::runtime_error("stub copyMakeBorder: BORDER_CONSTANT only");
    Mat out(src.rows + top + bottom, src.cols + left + right, src.type(), value);
    for (int r = 0; r < src.rows; r++)
        std::memcpy(out.data + (size_t)(r + top) * out.step + (size_t)left * src.elemSize(), src.data + (size_t)r * src.step,
                    (size_t)src.cols * src.elemSize());
    dst = out;
}

}  // namespace cv
