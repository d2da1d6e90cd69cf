// Image — drop-in for the reference's include/Image.hpp:14-28 (src/Image.cc:8-26) on the B200 path.
//
// Same public surface: a deep-copied 8-bit single-channel cv::Mat `rawImage`, the keypoint lists that
// Brief::computeBrief appends to, getW/getH/getPixelVal.  In addition every Image carries a small
// device-residency record: FastDetector / Brief upload the pixels to a slot of the shared
// yavo context the first time they see the frame and skip the upload (and reuse the blurred plane and the
// descriptors produced with it) while the host pixels are unchanged.  rawImage is a public, mutable member in the
// reference (its tests write pixels directly), so "unchanged" is never assumed: the Image's identity (id, a generation
// counter bumped by FastDetector::putPixel* / Image::touch(), the address of the pixel buffer, the frame size) is
// the cheap filter, and a hit is confirmed byte for byte against the pinned copy of the pixels the device call kept
// (one memcmp of a cache-resident frame, ~15 us at KITTI size).  YAVO_TRUST_IMAGE_IDENTITY=1 skips the confirmation
// for callers that never write through rawImage between two device calls on one Image (the reference's own loop,
// src/LoopHandler.cc:468-485, builds the Frame, then detects, then describes).
#ifndef YAVO_HOST_IMAGE_HPP
#define YAVO_HOST_IMAGE_HPP

#include <cstdint>
#include <vector>

#include <opencv2/core.hpp>

#include "BriefDescriptor.hpp"

class KeyPoint;

class Image {
   public:
    Image() : yavo_id_(nextId()) {}
    Image(const cv::Mat &img);
    ~Image();

    cv::Mat rawImage;
    std::vector<KeyPoint> keypoints;
    std::vector<KeyPoint> resetKeypoints;

    int getW() const;
    int getH() const;
    void unDistort();  // declared by the reference, never defined or called there; a no-op here
    uint8_t getPixelVal(int i, int j) const;

    // ---- device residency (not part of the reference interface) ----
    uint64_t yavoId() const { return yavo_id_; }
    uint64_t yavoGeneration() const { return yavo_gen_; }
    void touch() { yavo_gen_++; }   // the pixels were written through rawImage: device copies are stale
    uint64_t yavoChecksum() const;  // content hash of rawImage (rows, cols, pixels): slots filled without a pinned copy

   private:
    static uint64_t nextId();
    uint64_t yavo_id_;
    uint64_t yavo_gen_ = 0;
};
#endif
