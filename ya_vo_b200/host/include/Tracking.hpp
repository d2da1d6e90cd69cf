// Tracking — drop-in for the one OpenCV call the reference's steady-state loop spends its time in after
// feature extraction (LoopHandler::trackLastFrame, src/LoopHandler.cc:372-375):
//
//     cv::calcOpticalFlowPyrLK(lastFrame->rawImage, currentFrame->rawImage, lastFrameKpt, currFrameKpt,
//                              flowStatus, error, cv::Size(11, 11), 3,
//                              cv::TermCriteria(cv::TermCriteria::COUNT + cv::TermCriteria::EPS, 30, 0.01), 0, 0.001);
//
// yavo::calcOpticalFlowPyrLK keeps OpenCV's argument order, defaults and output conventions (points are
// (x = column, y = row); status 1 = tracked; err = mean absolute patch difference, or the minimum eigenvalue with
// OPTFLOW_LK_GET_MIN_EIGENVALS) and runs the pyramid and the per-point iterations on the device
// (include/yavo_b200.h: yavo_klt_track).  A maintainer swaps `cv::` for `yavo::` at that call site; the Image
// overload also skips the upload when the frame is already resident from getFastFeatures / computeBrief.
// Results are bit-identical to OpenCV 4.13's CPU path (tests/golden/klt_golden.npz).
#ifndef YAVO_HOST_TRACKING_HPP
#define YAVO_HOST_TRACKING_HPP

#include <vector>

#include <opencv2/core.hpp>

#include "Image.hpp"

namespace yavo {

enum { OPTFLOW_USE_INITIAL_FLOW = 4, OPTFLOW_LK_GET_MIN_EIGENVALS = 8 };  // cv::OPTFLOW_* values

void calcOpticalFlowPyrLK(const Image &prevImg, const Image &nextImg, const std::vector<cv::Point2f> &prevPts,
                          std::vector<cv::Point2f> &nextPts, std::vector<uchar> &status, std::vector<float> &err,
                          cv::Size winSize = cv::Size(21, 21), int maxLevel = 3,
                          cv::TermCriteria criteria = cv::TermCriteria(cv::TermCriteria::COUNT + cv::TermCriteria::EPS, 30, 0.01),
                          int flags = 0, double minEigThreshold = 1e-4);

// cv::Mat form of the same call (what the reference passes): the pixels are wrapped and uploaded per call
void calcOpticalFlowPyrLK(const cv::Mat &prevImg, const cv::Mat &nextImg, const std::vector<cv::Point2f> &prevPts,
                          std::vector<cv::Point2f> &nextPts, std::vector<uchar> &status, std::vector<float> &err,
                          cv::Size winSize = cv::Size(21, 21), int maxLevel = 3,
                          cv::TermCriteria criteria = cv::TermCriteria(cv::TermCriteria::COUNT + cv::TermCriteria::EPS, 30, 0.01),
                          int flags = 0, double minEigThreshold = 1e-4);

}  // namespace yavo
#endif
