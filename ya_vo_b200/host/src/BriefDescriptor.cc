#include "../include/BriefDescriptor.hpp"

#include <climits>
#include <cmath>
#include <cstring>

#include <opencv2/imgproc.hpp>

#include "../include/Image.hpp"
#include "../include/yavo_device.hpp"

using yavo_host::Device;

std::vector<std::vector<int>> Brief::preComputeOffsets() {
    // as the reference: a fresh table per construction, non-deterministic seed (src/BriefDescriptor.cc:4-20)
    std::random_device seeder;
    std::mt19937 generator(seeder());
    std::uniform_int_distribution<int> dist(-8, 8);
    std::vector<std::vector<int>> table(256, std::vector<int>(4));
    for (auto &row : table)
        for (int &v : row) v = dist(generator);
    return table;
}

// ---- hot path ------------------------------------------------------------------------------------

void Brief::computeBrief(const std::vector<cv::Point> &detectedCornerPoints, Image &img) {
    // Reference src/BriefDescriptor.cc:86-124.  The smoothed plane is the one the detect kernel left in
    // the frame's slot when getFastFeatures ran on the same pixels; otherwise it is computed now.
    const int n = (int)detectedCornerPoints.size();
    if (n == 0) return;
    if ((int)offsets.size() < 256) throw yavo_host::DeviceError("Brief: offset table missing (default-constructed Brief)");
    Device &dev = Device::instance(img.getH(), img.getW());
    std::lock_guard<std::mutex> lk(dev.mutex());
    int32_t table[1024];
    for (int j = 0; j < 256; j++)
        for (int k = 0; k < 4; k++) table[4 * j + k] = offsets[j][k];
    dev.setBriefOffsets(table);
    {
        // FastDetector::getFastFeatures ran the fused single-frame graph on these pixels with this table, and these are
        // its points: the descriptors are on the host already
        const int hit = dev.findSlot(img);
        if (hit >= 0) {
            const Device::Features &f = dev.features(hit);
            bool same = f.valid && f.offsets_epoch == dev.offsetsEpoch() && (int)f.rows.size() == n;
            for (int i = 0; same && i < n; i++) same = detectedCornerPoints[i].x == f.rows[i] && detectedCornerPoints[i].y == f.cols[i];
            if (same) {
                const int H = img.getH(), W = img.getW();
                lastOob = 0;
                for (size_t k = 0; k < f.ids.size(); k++) {
                    const int i = f.ids[k];
                    KeyPoint kp(f.rows[i], f.cols[i], i);
                    std::memcpy(kp.featVec, f.desc.data() + k * 32, 32);
                    img.keypoints.push_back(kp);
                    if (f.rows[i] + 9 >= H) {  // border keypoint: does any sample read past the pixel buffer (reference UB)?
                        bool oob = false;
                        for (int j = 0; j < 256 && !oob; j++)
                            for (int e = 0; e < 2 && !oob; e++) {
                                int r = f.rows[i] + table[4 * j + 2 * e], c = f.cols[i] + table[4 * j + 2 * e + 1];
                                if (c >= W) { c -= W; r += 1; }
                                oob = r >= H;
                            }
                        lastOob += oob ? 1 : 0;
                    }
                }
                return;
            }
        }
    }
    const int slot = dev.slotFor(img);
    std::vector<int32_t> rows(n), cols(n);
    for (int i = 0; i < n; i++) {
        rows[i] = detectedCornerPoints[i].x;
        cols[i] = detectedCornerPoints[i].y;
    }
    std::vector<uint8_t> desc((size_t)n * 32), valid(n);
    dev.check(yavo_brief_describe(dev.ctx(), slot, rows.data(), cols.data(), n, desc.data(), valid.data(), &lastOob));
    for (int i = 0; i < n; i++) {
        if (!valid[i]) continue;  // checkBoundry rejected it: the reference appends nothing
        KeyPoint kp(rows[i], cols[i], i);
        std::memcpy(kp.featVec, desc.data() + (size_t)i * 32, 32);
        img.keypoints.push_back(kp);
    }
}

std::vector<Matches> Brief::matchFeatures(Image &img1, Image &img2) {
    // Reference src/BriefDescriptor.cc:163-183: one match per keypoint of img1, first minimum wins
    const int n1 = (int)img1.keypoints.size(), n2 = (int)img2.keypoints.size();
    std::vector<Matches> out;
    if (n1 == 0) return out;
    std::vector<uint8_t> d1((size_t)n1 * 32), d2((size_t)n2 * 32);
    for (int i = 0; i < n1; i++) std::memcpy(d1.data() + (size_t)i * 32, img1.keypoints[i].featVec, 32);
    for (int j = 0; j < n2; j++) std::memcpy(d2.data() + (size_t)j * 32, img2.keypoints[j].featVec, 32);
    std::vector<int32_t> idx(n1), dist(n1);
    Device &dev = Device::instance(std::max(img1.getH(), img2.getH()), std::max(img1.getW(), img2.getW()));
    std::lock_guard<std::mutex> lk(dev.mutex());
    dev.check(yavo_match(dev.ctx(), d1.data(), n1, d2.data(), n2, idx.data(), dist.data(), nullptr, nullptr));
    out.reserve(n1);
    for (int i = 0; i < n1; i++) {
        KeyPoint kp2(0, 0, 0);
        if (idx[i] >= 0) {
            const KeyPoint &t = img2.keypoints[idx[i]];
            kp2.x = t.x;
            kp2.y = t.y;
            kp2.id = t.id;
        }
        out.push_back(Matches(img1.keypoints[i], kp2, dist[i]));
    }
    return out;
}

void Brief::removeOutliers(std::vector<Matches> &matches, std::vector<Matches> &newMatches, int threshold) {
    // Reference src/BriefDescriptor.cc:213-231 (which dereferences end() on an empty list; nothing is kept here)
    const int n = (int)matches.size();
    if (n == 0) return;
    std::vector<int32_t> dist(n);
    for (int i = 0; i < n; i++) dist[i] = matches[i].distance;
    std::vector<uint8_t> keep(n);
    yavo_remove_outliers(dist.data(), n, threshold, keep.data());
    for (int i = 0; i < n; i++)
        if (keep[i]) {
            matches[i].pt1.matched = true;
            matches[i].pt2.matched = true;
            newMatches.push_back(matches[i]);
        }
}

// ---- small public helpers kept for source compatibility (host-side, scalar) -------------------------

int Brief::popCount(uchar v) {
    int c = 0;
    for (; v; v = (uchar)(v & (v - 1))) c++;
    return c;
}

int Brief::hammingDistance(uchar a[32], uchar b[32]) {
    int d = 0;
    for (int i = 0; i < 32; i++) d += popCount((uchar)(a[i] ^ b[i]));
    return d;
}

void Brief::convolve2d(const Image &img, cv::Mat &kernel, cv::Mat &output) {
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

void Brief::gaussianBlur(const Image &img, int sigma, cv::Mat &outImage) {
    // dead code in the reference (its call is commented out at src/BriefDescriptor.cc:88); kept callable.
    // Note the reference fills only element (0,1) of the row kernel (:70-72); reproduced.
    const int ks = 3 * sigma;
    cv::Mat gx = cv::Mat::zeros(ks, 1, CV_32FC1), gy = cv::Mat::zeros(1, ks, CV_32FC1);
    for (int i = 0; i < ks; i++) {
        const float g = (float)((1 / (std::sqrt(2 * M_PI) * sigma)) * std::exp(-std::pow(i, 2) / (2 * std::pow(sigma, 2))));
        gx.at<float>(i, 0) = g;
        if (ks > 1) gy.at<float>(0, 1) = g;
    }
    cv::Mat k2 = gx * gy;
    convolve2d(img, k2, outImage);
    outImage.convertTo(outImage, CV_8UC1);
}

cv::Mat Brief::drawMatches(Image &img1, Image &img2, std::vector<Matches> &matches) {
    // debug drawing (reference src/BriefDescriptor.cc:186-210): side-by-side canvas with match lines
    cv::Mat canvas = cv::Mat::zeros(img1.getH(), img1.getW() + img2.getW(), CV_8UC1);
    for (int i = 0; i < img1.getH(); i++)
        for (int j = 0; j < img1.getW() - 1; j++) canvas.at<uchar>(i, j) = img1.getPixelVal(i, j);
    for (int i = 0; i < img2.getH() - 1 && i < canvas.rows; i++)
        for (int j = 0; j < img2.getW() - 1; j++) canvas.at<uchar>(i, j + img1.getW()) = img2.getPixelVal(i, j);
    cv::cvtColor(canvas, canvas, cv::COLOR_GRAY2RGB);
    for (Matches &m : matches)
        cv::line(canvas, cv::Point(m.pt1.y, m.pt1.x), cv::Point(m.pt2.y + img1.getW(), m.pt2.x), cv::Scalar(255, 255, 255), 1);
    return canvas;
}
