// extern "C" driver over the reference's unmodified classes (TEST INFRASTRUCTURE; see README.md).
// `private` is made public for THIS translation unit only, so a fixed BRIEF offset table can be
// injected (the reference draws it from std::random_device, src/BriefDescriptor.cc:4-20); the class
// layout is unchanged and the reference's own translation units are compiled untouched.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>

#define private public
#include "include/BriefDescriptor.hpp"
#include "include/FastDetector.hpp"
#include "include/Image.hpp"
#undef private

namespace {
struct Quiet {  // getFastFeatures prints sizes and timings on every call (src/FastDetector.cc:287-349)
    std::streambuf *old;
    std::ostringstream sink;
    Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(old); }
};
cv::Mat wrap(const uint8_t *img, int H, int W) { return cv::Mat(H, W, CV_8UC1, (void *)img); }
}  // namespace

extern "C" {

int ref_ring(int xc, int yc, int32_t *out_xy) {
    cv::Mat z = cv::Mat::zeros(8, 8, CV_8UC1);
    Image im(z);
    FastDetector fd(12, 50);
    std::vector<cv::Point> p = fd.getBresenhamCirclePoints(im, xc, yc);
    for (size_t k = 0; k < p.size() && k < 16; k++) { out_xy[2 * k] = p[k].x; out_xy[2 * k + 1] = p[k].y; }
    return (int)p.size();
}

int ref_check_contiguous(const uint8_t *img, int H, int W, int xc, int yc) {
    Image im(wrap(img, H, W));
    FastDetector fd(12, 50);
    std::vector<cv::Point> p = fd.getBresenhamCirclePoints(im, xc, yc);
    return fd.checkContiguousPixels(im.getPixelVal(xc, yc), p, im) ? 1 : 0;
}

float ref_harris(const uint8_t *img, int H, int W, int x, int y) {
    Image im(wrap(img, H, W));
    FastDetector fd(12, 50);
    cv::Mat Ix = cv::Mat::zeros(H, W, CV_32FC1), Iy = cv::Mat::zeros(H, W, CV_32FC1);
    fd.preComputeHarris(im, Ix, Iy);
    return fd.getHarrisCornerResponse(im, x, y, Ix, Iy);
}

int ref_fast(const uint8_t *img, int H, int W, int32_t *rows, int32_t *cols, int cap) {
    Quiet q;
    Image im(wrap(img, H, W));
    FastDetector fd(12, 50);
    std::vector<cv::Point> f = fd.getFastFeatures(im);
    for (size_t i = 0; i < f.size() && (int)i < cap; i++) { rows[i] = f[i].x; cols[i] = f[i].y; }
    return (int)f.size();
}

// computeBrief on explicit points with an injected offset table; returns the keypoints appended
int ref_brief(const uint8_t *img, int H, int W, const int32_t *offsets, const int32_t *rows, const int32_t *cols, int n,
              int32_t *out_x, int32_t *out_y, int32_t *out_id, uint8_t *out_desc) {
    Image im(wrap(img, H, W));
    Brief brief(256);
    for (int j = 0; j < 256; j++)
        for (int k = 0; k < 4; k++) brief.offsets[j][k] = offsets[4 * j + k];
    std::vector<cv::Point> pts;
    for (int i = 0; i < n; i++) pts.push_back(cv::Point(rows[i], cols[i]));
    brief.computeBrief(pts, im);
    for (size_t i = 0; i < im.keypoints.size(); i++) {
        out_x[i] = im.keypoints[i].x;
        out_y[i] = im.keypoints[i].y;
        out_id[i] = im.keypoints[i].id;
        std::memcpy(out_desc + 32 * i, im.keypoints[i].featVec, 32);
    }
    return (int)im.keypoints.size();
}

// matchFeatures + removeOutliers on explicit descriptor sets; train keypoint j carries id = j
int ref_match(const uint8_t *d1, int n1, const uint8_t *d2, int n2, int threshold, int32_t *out_idx, int32_t *out_dist,
              uint8_t *out_keep) {
    cv::Mat z = cv::Mat::zeros(8, 8, CV_8UC1);
    Image a(z), b(z);
    for (int i = 0; i < n1; i++) { KeyPoint k(i, 2 * i, i); std::memcpy(k.featVec, d1 + 32 * i, 32); a.keypoints.push_back(k); }
    for (int j = 0; j < n2; j++) { KeyPoint k(j, 3 * j, j); std::memcpy(k.featVec, d2 + 32 * j, 32); b.keypoints.push_back(k); }
    Brief brief(256);
    std::vector<Matches> m = brief.matchFeatures(a, b);
    for (size_t i = 0; i < m.size(); i++) {
        out_dist[i] = m[i].distance;
        out_idx[i] = (n2 > 0) ? m[i].pt2.id : -1;
    }
    if (out_keep && !m.empty()) {
        std::vector<Matches> kept;
        brief.removeOutliers(m, kept, threshold);
        for (size_t i = 0; i < m.size(); i++) out_keep[i] = m[i].pt1.matched ? 1 : 0;
    }
    return (int)m.size();
}

// The benchmark call order of tests/BriefDescriptorTest.cc:13-44 over a run of n consecutive frames:
// getFastFeatures + computeBrief on every frame, matchFeatures(previous, current) + removeOutliers(.., threshold)
// for every frame (the first frame is matched against the last one, so n frames cost n detections and n matches).
// stdout chatter of the reference is left alone: the caller points fd 1 at /dev/null.
// Returns the number of frames processed; per-frame keypoint and kept-match counts are written when asked for.
int ref_sequence(const uint8_t *frames, int n, int H, int W, const int32_t *offsets, int threshold, int32_t *out_nkp,
                 int32_t *out_nkept) {
    FastDetector fd(12, 50);
    Brief brief(256);
    for (int j = 0; j < 256; j++)
        for (int k = 0; k < 4; k++) brief.offsets[j][k] = offsets[4 * j + k];
    std::vector<Image> imgs;
    for (int f = 0; f < n; f++) {
        imgs.emplace_back(wrap(frames + (size_t)f * H * W, H, W));
        Image &im = imgs.back();
        std::vector<cv::Point> pts = fd.getFastFeatures(im);
        brief.computeBrief(pts, im);
        if (out_nkp) out_nkp[f] = (int)im.keypoints.size();
    }
    for (int f = 0; f < n; f++) {
        Image &prev = imgs[(f + n - 1) % n], &cur = imgs[f];
        std::vector<Matches> m = brief.matchFeatures(prev, cur), kept;
        if (!m.empty()) brief.removeOutliers(m, kept, threshold);
        if (out_nkept) out_nkept[f] = (int)kept.size();
    }
    return n;
}

// Steady-state form of the same call order for timing (bench.py --impl reference): one detector, one descriptor
// object and the previous frame are kept; every push runs getFastFeatures + computeBrief on the new frame and
// matchFeatures(previous, new) + removeOutliers, as LoopHandler does per frame (src/LoopHandler.cc:468-485,534-537).
struct RefStream {
    FastDetector fd{12, 50};
    Brief brief{256};
    std::vector<Image> prev;  // 0 or 1 element (Image has no default constructor)
};

void *ref_stream_new(const int32_t *offsets) {
    RefStream *s = new RefStream();
    for (int j = 0; j < 256; j++)
        for (int k = 0; k < 4; k++) s->brief.offsets[j][k] = offsets[4 * j + k];
    return s;
}

int ref_stream_push(void *h, const uint8_t *frame, int H, int W, int threshold, int32_t *out_nkp, int32_t *out_nkept) {
    RefStream *s = (RefStream *)h;
    Image cur(wrap(frame, H, W));
    std::vector<cv::Point> pts = s->fd.getFastFeatures(cur);
    s->brief.computeBrief(pts, cur);
    if (out_nkp) *out_nkp = (int)cur.keypoints.size();
    int kept_n = 0;
    if (!s->prev.empty()) {
        std::vector<Matches> m = s->brief.matchFeatures(s->prev[0], cur), kept;
        if (!m.empty()) s->brief.removeOutliers(m, kept, threshold);
        kept_n = (int)kept.size();
    }
    if (out_nkept) *out_nkept = kept_n;
    s->prev.clear();
    s->prev.push_back(cur);
    return 0;
}

void ref_stream_free(void *h) { delete (RefStream *)h; }

}  // extern "C"
