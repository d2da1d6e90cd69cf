#include "../include/Tracking.hpp"

#include "../include/yavo_device.hpp"

using yavo_host::Device;

namespace yavo {

void calcOpticalFlowPyrLK(const Image &prevImg, const Image &nextImg, const std::vector<cv::Point2f> &prevPts,
                          std::vector<cv::Point2f> &nextPts, std::vector<uchar> &status, std::vector<float> &err,
                          cv::Size winSize, int maxLevel, cv::TermCriteria criteria, int flags, double minEigThreshold) {
    const size_t n = prevPts.size();
    // OpenCV: with OPTFLOW_USE_INITIAL_FLOW nextPts must already hold n points; otherwise it is (re)created
    if (flags & OPTFLOW_USE_INITIAL_FLOW) {
        if (nextPts.size() != n) throw yavo_host::DeviceError("calcOpticalFlowPyrLK: nextPts must have prevPts' size with OPTFLOW_USE_INITIAL_FLOW");
    } else {
        nextPts.assign(n, cv::Point2f(0.f, 0.f));
    }
    status.assign(n, 0);
    err.assign(n, 0.f);
    if (n == 0) return;
    if (prevImg.getH() != nextImg.getH() || prevImg.getW() != nextImg.getW())
        throw yavo_host::DeviceError("calcOpticalFlowPyrLK: frames differ in size");
    Device &dev = Device::instance(prevImg.getH(), prevImg.getW());
    std::lock_guard<std::mutex> lk(dev.mutex());
    const int sp = dev.slotFor(prevImg);
    const int sn = dev.slotFor(nextImg);  // least-recently-used eviction never takes the slot just touched
    static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "cv::Point2f is two packed floats");
    dev.check(yavo_klt_track(dev.ctx(), sp, sn, reinterpret_cast<const float *>(prevPts.data()), (int)n,
                             reinterpret_cast<float *>(nextPts.data()), status.data(), err.data(), winSize.width,
                             winSize.height, maxLevel, criteria.type, criteria.maxCount, criteria.epsilon, flags,
                             minEigThreshold));
}

void calcOpticalFlowPyrLK(const cv::Mat &prevImg, const cv::Mat &nextImg, const std::vector<cv::Point2f> &prevPts,
                          std::vector<cv::Point2f> &nextPts, std::vector<uchar> &status, std::vector<float> &err,
                          cv::Size winSize, int maxLevel, cv::TermCriteria criteria, int flags, double minEigThreshold) {
    Image a(prevImg), b(nextImg);
    calcOpticalFlowPyrLK(a, b, prevPts, nextPts, status, err, winSize, maxLevel, criteria, flags, minEigThreshold);
}

}  // namespace yavo
