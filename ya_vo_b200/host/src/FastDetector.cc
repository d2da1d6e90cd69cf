#include "../include/FastDetector.hpp"

#include <chrono>
#include <cmath>

#include "../include/yavo_device.hpp"

using yavo_host::Device;

// ---- the hot path ------------------------------------------------------------------------------
// Reference src/FastDetector.cc:277-369.  One C-ABI call: the frame is uploaded on first sight (or
// found resident), the fused kernel produces the corner mask and the blurred plane BRIEF will need,
// candidates are scored and the reference's std::sort order is replayed on the device.
std::vector<cv::Point> FastDetector::getFastFeatures(const Image &img) {
    if (verbose) {
        std::cout << "sicr: " << img.rawImage.rows << std::endl;
        std::cout << img.rawImage.cols << std::endl;
    }
    const auto t1 = std::chrono::steady_clock::now();
    Device &dev = Device::instance(img.getH(), img.getW());
    std::lock_guard<std::mutex> lk(dev.mutex());
    const int cap = std::min(fastCornerNumThreshold > 0 ? fastCornerNumThreshold : Device::kMaxKeypoints,
                             (int)Device::kMaxKeypoints);
    std::vector<int32_t> rows, cols;
    int n = 0, ncand = 0;
    int slot = dev.findSlot(img);
    if (slot >= 0 && dev.features(slot).valid && dev.features(slot).cap == cap) {
        // the same pixels were detected before (same Image, untouched): the answer is on the host already
        const Device::Features &f = dev.features(slot);
        rows = f.rows;
        cols = f.cols;
        lastScores_ = f.scores;
        n = (int)rows.size();
        ncand = f.n_cand;
    } else if (dev.hasBriefOffsets()) {
        // One graph launch and one synchronisation for the reference's per-frame sequence (src/LoopHandler.cc:468-485:
        // getFastFeatures, then computeBrief on its points): upload, detect + score + blur, exact top-K, BRIEF of the
        // admitted points.  The descriptors stay with the slot record; Brief::computeBrief picks them up when it is
        // handed these very points.
        slot = dev.claimSlot(img);
        Device::Features &f = dev.features(slot);
        f.rows.assign(cap, 0);
        f.cols.assign(cap, 0);
        f.scores.assign(cap, 0.f);
        f.ids.assign(cap, 0);
        f.desc.assign((size_t)cap * 32, 0);
        std::vector<int32_t> brows(cap), bcols(cap);
        int nb = 0;
        const cv::Mat &m = img.rawImage;
        dev.check(yavo_frame_features(dev.ctx(), slot, m.data, m.rows, m.cols, (int)m.step, cap, &n, f.rows.data(), f.cols.data(),
                                      f.scores.data(), &nb, brows.data(), bcols.data(), f.ids.data(), f.desc.data(), &ncand));
        f.rows.resize(n);
        f.cols.resize(n);
        f.scores.resize(n);
        f.ids.resize(nb);
        f.desc.resize((size_t)nb * 32);
        f.cap = cap;
        f.n_cand = ncand;
        f.offsets_epoch = dev.offsetsEpoch();
        f.valid = true;
        dev.markShadow(slot);
        rows = f.rows;
        cols = f.cols;
        lastScores_ = f.scores;
    } else {
        // no BRIEF table on the device yet (no Brief::computeBrief has run): detect only
        slot = dev.slotFor(img);
        rows.assign(cap, 0);
        cols.assign(cap, 0);
        lastScores_.assign(cap, 0.f);
        dev.check(yavo_fast_detect(dev.ctx(), slot, cap, rows.data(), cols.data(), lastScores_.data(), &n, &ncand));
        lastScores_.resize(n);
    }
    lastCandidates_ = ncand;
    std::vector<cv::Point> out;
    out.reserve(n);
    for (int i = 0; i < n; i++) out.push_back(cv::Point(rows[i], cols[i]));  // (x = row, y = col)
    if (verbose) {
        const std::chrono::duration<double> dt = std::chrono::steady_clock::now() - t1;
        std::cout << "Looping cost time: " << dt.count() << " seconds." << std::endl;
        if (ncand == 0) std::cout << "No corners found" << std::endl;
    }
    return out;
}

// ---- small public helpers kept for source compatibility (host-side, scalar) -----------------------

std::vector<cv::Point> FastDetector::getBresenhamCirclePoints(const Image & /*img*/, int xc, int yc) {
    int32_t xy[32];
    yavo_ring_points(xc, yc, xy);  // the constant table the reference's set-based generator yields
    std::vector<cv::Point> pts;
    pts.reserve(16);
    for (int k = 0; k < 16; k++) pts.push_back(cv::Point(xy[2 * k], xy[2 * k + 1]));
    return pts;
}

std::vector<cv::Point> FastDetector::getAllSymPoints(int x, int y) {
    // the eight octant images of (x, y), reference order (src/FastDetector.cc:114-116)
    std::vector<cv::Point> p;
    const int sx[8] = {x, y, y, x, -x, -y, -y, -x}, sy[8] = {y, x, -x, -y, -y, -x, x, y};
    for (int k = 0; k < 8; k++) p.push_back(cv::Point(sx[k], sy[k]));
    return p;
}

bool FastDetector::checkContiguousPixels(uint8_t centPixel, const std::vector<cv::Point> &circlePoints,
                                         const Image &img) {
    // 12 consecutive differing ring pixels scanning indices 0..15 once, no wrap-around (:135-153)
    int run = 0;
    for (int k = 0; k < 16; k++) {
        run = checkInBetween(centPixel, img.getPixelVal(circlePoints[k].x, circlePoints[k].y)) ? 0 : run + 1;
        if (run >= 12) return true;
    }
    return false;
}

void FastDetector::putPixel(Image &img, cv::Point pt) { img.rawImage.at<uint8_t>(pt) = 255; img.touch(); }
void FastDetector::putPixel(Image &img, cv::Point pt, uint8_t pixVal) { img.rawImage.at<uint8_t>(pt) = pixVal; img.touch(); }
void FastDetector::putPixelColor(Image &img, cv::Point pt) { img.rawImage.at<cv::Vec3b>(pt) = cv::Vec3b(255, 255, 0); img.touch(); }

void FastDetector::convolve2d(const Image &img, cv::Mat &kernel, cv::Mat &output) {
    // float correlation with a zero border; writes output(r, c) for r <= rows-1-2h, c <= cols-1-2h only,
    // centred on (r, c) — the reference's loop bounds (:164-200), so the tail rows/cols stay untouched
    const int ks = kernel.rows, h = ks / 2, R = img.rawImage.rows, C = img.rawImage.cols;
    for (int r = 0; r < R - 2 * h; r++)
        for (int c = 0; c < C - 2 * h; c++) {
            float sum = 0;
            for (int k = 0; k < ks; k++)
                for (int l = 0; l < ks; l++) {
                    const int rr = r + k - h, cc = c + l - h;
                    const float v = (rr < 0 || cc < 0 || rr >= R || cc >= C) ? 0.f : (float)img.rawImage.at<uchar>(rr, cc);
                    sum += kernel.at<float>(k, l) * v;
                }
            output.at<float>(r, c) = sum;
        }
}

void FastDetector::preComputeHarris(const Image &img, cv::Mat &Ix, cv::Mat &Iy) {
    cv::Mat sobelx = (cv::Mat_<float>(3, 3) << -1, 0, 1, -2, 0, 2, -1, 0, 1);
    cv::Mat sobely = (cv::Mat_<float>(3, 3) << -1, -2, -1, 0, 0, 0, 1, 2, 1);
    convolve2d(img, sobelx, Ix);
    convolve2d(img, sobely, Iy);
}

void FastDetector::gaussianBlur(const Image &img, int sigma, cv::Mat &outImage) {
    // the reference's hand-rolled blur (unused by its own hot path): one-sided 1-D Gaussian of 3*sigma
    // taps, outer product, float correlation, converted to 8 bits
    const int ks = 3 * sigma;
    cv::Mat gx = cv::Mat::zeros(1, ks, CV_32FC1), gy = cv::Mat::zeros(ks, 1, CV_32FC1);
    for (int i = 0; i < ks; i++) {
        const float g = (float)((1 / (std::sqrt(2 * M_PI) * sigma)) * std::exp(-std::pow(i, 2) / (2 * std::pow(sigma, 2))));
        gx.at<float>(0, i) = g;
        gy.at<float>(i, 0) = g;
    }
    cv::Mat k2 = gx * gy;
    convolve2d(img, k2, outImage);
    outImage.convertTo(outImage, CV_8UC1);
}

namespace {
// cv::eigen on a symmetric 2x2 float matrix as OpenCV's Jacobi solver (no-Eigen build) computes it
inline float hyp(float a, float b) {
    a = std::fabs(a);
    b = std::fabs(b);
    if (a > b) { b /= a; return a * std::sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * std::sqrt(1 + a * a); }
    return 0;
}
}  // namespace

float FastDetector::getHarrisCornerResponse(const Image & /*img*/, int x, int y, const cv::Mat &Ix, const cv::Mat &Iy) {
    float m00 = 0, m01 = 0, m11 = 0;
    for (int i = x - 1; i <= x + 1; i++)
        for (int j = y - 1; j <= y + 1; j++) {
            const float gx = Ix.at<float>(i, j), gy = Iy.at<float>(i, j);
            m00 += gx * gx;
            m01 += gx * gy;
            m11 += gy * gy;
        }
    float w0 = m00, w1 = m11;
    if (!(std::fabs(m01) <= 1.1920928955078125e-07f)) {
        const float yv = (float)((w1 - w0) * 0.5);
        float t = std::fabs(yv) + hyp(m01, yv);
        t = (m01 / t) * m01;
        if (yv < 0) t = -t;
        w0 -= t;
        w1 += t;
    }
    const float l1 = std::max(w0, w1), l2 = std::min(w0, w1);
    const float prod = l1 * l2, sum = l2 + l1;
    return (float)((double)prod - 0.04 * ((double)sum * (double)sum));
}
