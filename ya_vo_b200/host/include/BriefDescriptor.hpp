// KeyPoint / Matches / Brief — drop-in for the reference's include/BriefDescriptor.hpp:11-68 and
// src/BriefDescriptor.cc.  Field layout and method signatures are the reference's; the pixel and
// descriptor work (9x9 Gaussian, 256 intensity-pair tests, brute-force Hamming arg-min) runs in the
// sm_100a kernels behind the C ABI of include/yavo_b200.h.  There is no CPU implementation of
// computeBrief / matchFeatures in this class.
#ifndef YAVO_HOST_BRIEF_DESCRIPTOR_HPP
#define YAVO_HOST_BRIEF_DESCRIPTOR_HPP

#include <algorithm>
#include <random>
#include <vector>

#include <opencv2/core.hpp>

class Image;

// include/BriefDescriptor.hpp:11-24: x = row, y = col, id = index in the point list handed to computeBrief
class KeyPoint {
   public:
    KeyPoint() {}
    KeyPoint(const int _x, const int _y, const int _id) : x(_x), y(_y), id(_id) {}
    int x;
    int y;
    int id;
    bool matched = false;
    uchar featVec[32] = {};
};

// include/BriefDescriptor.hpp:27-39
class Matches {
   public:
    Matches() {}
    Matches(const KeyPoint &_pt1, const KeyPoint &_pt2, int distance) : pt1(_pt1), pt2(_pt2), distance(distance) {}
    KeyPoint pt1;
    KeyPoint pt2;
    int distance;
};

// Drop-in for the reference's `Brief`.  Grouped by role rather than in the reference's declaration order:
//   * the three calls LoopHandler makes on the hot path (GPU),
//   * construction / the offset table,
//   * small scalar helpers that the reference happens to expose publicly (host-side, unchanged semantics).
class Brief {
   public:
    // ---- hot path: device kernels behind include/yavo_b200.h ----------------------------------------------------
    // Appends one KeyPoint per admitted point to img.keypoints (a second call appends again, as in the
    // reference); smoothing = cv::GaussianBlur(9x9, sigma 2.5) exactly as OpenCV 4.x computes it for 8-bit input.
    void computeBrief(const std::vector<cv::Point> &detectedCornerPoints, Image &img);
    // One Matches entry per keypoint of img1: the keypoint of img2 with the smallest Hamming distance, the
    // lowest index among equal minima; pt2 carries only x, y and id, like the reference.
    std::vector<Matches> matchFeatures(Image &img1, Image &img2);
    // keeps distance < max(2 * smallest distance, threshold) and marks kept matches on the input list too
    void removeOutliers(std::vector<Matches> &matches, std::vector<Matches> &newMatches, int threshold);

    // ---- construction ----------------------------------------------------------------------------------------------
    Brief() {}
    // numTests is stored as patchSize and used as the test count exactly as the reference does (256 in practice)
    Brief(int numTests) : patchSize(numTests), offsets(preComputeOffsets()) {}
    ~Brief() {}
    // src/BriefDescriptor.cc:4-20: 256 x 4 offsets in [-8,8] from mt19937(random_device) — new every construction
    std::vector<std::vector<int>> preComputeOffsets();
    // not in the reference: inject a fixed table (tests, benchmarks, reproducible runs)
    void setOffsets(const std::vector<std::vector<int>> &table) { offsets = table; }
    const std::vector<std::vector<int>> &getOffsets() const { return offsets; }
    // keypoints whose tests read past the end of the pixel buffer in the last computeBrief
    // (undefined behaviour in the reference; those reads are defined as 0 here)
    int lastOutOfBufferCount() const { return lastOob; }

    // ---- scalar helpers kept for source compatibility ------------------------------------------------------------
    int hammingDistance(uchar featVec1[32], uchar featVec2[32]);
    int popCount(uchar featVec);
    inline bool checkBoundry(int x, int y, int width, int height);
    cv::Mat drawMatches(Image &img1, Image &img2, std::vector<Matches> &matches);  // debug canvas
    void gaussianBlur(const Image &img, int sigma, cv::Mat &outImage);               // the reference's unused hand-rolled blur
    void convolve2d(const Image &img, cv::Mat &kernel, cv::Mat &output);

   private:
    int patchSize = 256;
    std::vector<std::vector<int>> offsets;
    int lastOob = 0;
};

inline bool Brief::checkBoundry(int x, int y, int width, int height) {
    return !(x - 8 < 0 || x + 8 > width || y - 8 < 0 || y + 8 > height);
}
#endif
