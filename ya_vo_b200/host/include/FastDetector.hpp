// FastDetector — drop-in for the reference's include/FastDetector.hpp:17-55 / src/FastDetector.cc.
// getFastFeatures runs on the GPU (fused segment-test + blur kernel, Harris scoring of the passing
// pixels, exact replay of the reference's std::sort order, top 2000) through include/yavo_b200.h.
// The small public helpers the reference exposes (ring points, the run test, Sobel / Harris on a
// cv::Mat) are kept with their semantics for source compatibility; they are host-side scalar code and
// are not used by getFastFeatures.
#ifndef YAVO_HOST_FAST_DETECTOR_HPP
#define YAVO_HOST_FAST_DETECTOR_HPP

#include <cstdint>
#include <iostream>
#include <vector>

#include <opencv2/core.hpp>

#include "Image.hpp"

class FastDetector {
   public:
    FastDetector() {}
    // include/FastDetector.hpp:32-38: the intensity threshold argument is ignored (always 40), so is
    // the run length (the literal 12 in src/FastDetector.cc:147); both quirks are preserved.
    FastDetector(int _minDetectionThresold, uint8_t /*_intensityThreshold*/)
        : minDetectionThreshold(_minDetectionThresold), bresRadius(3), intensityThreshold(40),
          fastCornerNumThreshold(2000), harrisThreshold(2) {}
    ~FastDetector() {}

    std::vector<cv::Point> getFastFeatures(const Image &img);
    std::vector<cv::Point> getBresenhamCirclePoints(const Image &img, int x, int y);
    std::vector<cv::Point> getAllSymPoints(int x, int y);
    bool checkContiguousPixels(uint8_t centPixel, const std::vector<cv::Point> &circlePoints, const Image &img);
    inline bool checkInBetween(uint8_t centPixel, uint8_t condPixel);

    void putPixel(Image &img, cv::Point pt);
    void putPixel(Image &img, cv::Point pt, uint8_t pixVal);
    void putPixelColor(Image &img, cv::Point pt);

    void convolve2d(const Image &img, cv::Mat &kernel, cv::Mat &output);
    void gaussianBlur(const Image &img, int sigma, cv::Mat &outImage);
    void preComputeHarris(const Image &img, cv::Mat &Ix, cv::Mat &Iy);
    float getHarrisCornerResponse(const Image &img, int x, int y, const cv::Mat &Ix, const cv::Mat &Iy);

    // not in the reference: scores of the last getFastFeatures call, in the returned order; console
    // chatter of the reference ("sicr:", timing lines) is off unless verbose is set
    const std::vector<float> &lastScores() const { return lastScores_; }
    int lastCandidateCount() const { return lastCandidates_; }
    bool verbose = false;

   private:
    int minDetectionThreshold = 12;
    int bresRadius = 3;
    uint8_t intensityThreshold = 40;
    int fastCornerNumThreshold = 2000;
    int harrisThreshold = 2;
    std::vector<float> lastScores_;
    int lastCandidates_ = 0;
};

inline bool FastDetector::checkInBetween(uint8_t centPixel, uint8_t condPixel) {
    return (centPixel > condPixel - this->intensityThreshold) && (centPixel < condPixel + this->intensityThreshold);
}
#endif
