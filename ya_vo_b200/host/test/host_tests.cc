// host_tests — the reference's own tests, restated against the drop-in classes (no gtest in this image).
//
//   host_tests known <bresenham_50x50.bin>
//       tests/FastDetectorTest.cc:6-31  BresenhamCircleCheck
//       tests/FastDetectorTest.cc:38-61 CheckContiguosPixels
//       tests/FastDetectorTest.cc:64-80 CheckDiscontinuous
//       tests/ImageTest.cc:23-37        GetPixelMethod
//       (no device needed: these exercise only the host-side helpers)
//   host_tests pipeline <frameA.bin> <frameB.bin> <H> <W> <offsets.bin> <out.bin>
//       tests/BriefDescriptorTest.cc:9-64 call order: FAST x2, BRIEF x2, match, removeOutliers(…,20),
//       drawMatches; results are dumped for the pytest side to compare with the oracle.  Needs a GPU.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <algorithm>

#include "../include/BriefDescriptor.hpp"
#include "../include/FastDetector.hpp"
#include "../include/FrameStream.hpp"
#include "../include/Tracking.hpp"
#include "../include/Image.hpp"
#include "../include/yavo_device.hpp"

static int g_fail = 0;
#define EXPECT_TRUE(c)                                                       \
    do {                                                                     \
        if (!(c)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); g_fail++; } \
    } while (0)
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))

static std::vector<uint8_t> slurp(const char *p) {
    std::ifstream f(p, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

static int known(const char *bres_path) {
    std::vector<uint8_t> gold = slurp(bres_path);
    EXPECT_EQ(gold.size(), (size_t)2500);
    {   // BresenhamCircleCheck
        cv::Mat testcv(cv::Size(50, 50), CV_8UC1, cv::Scalar(0));
        Image testImage(testcv);
        FastDetector fd(12, 50);
        std::vector<cv::Point2i> bres = fd.getBresenhamCirclePoints(testImage, 25, 25);
        EXPECT_EQ(bres.size(), (size_t)16);
        for (auto &p : bres) fd.putPixel(testImage, p);
        // cv::subtract(painted, golden) must lie in [-1, 1): saturating u8 difference is 0 everywhere
        for (int r = 0; r < 50; r++)
            for (int c = 0; c < 50; c++) {
                int d = (int)testImage.rawImage.at<uchar>(r, c) - (int)gold[r * 50 + c];
                if (d < 0) d = 0;
                EXPECT_TRUE(d < 1);
            }
        cv::Mat g(50, 50, CV_8UC1, gold.data());
        Image goldImg(g);
        for (auto &p : bres) EXPECT_EQ((int)goldImg.getPixelVal(p.x, p.y), 255);  // ImageTest GetPixelMethod
    }
    {   // CheckContiguosPixels
        cv::Mat testcv(cv::Size(50, 50), CV_8UC1, cv::Scalar(0));
        Image testImage(testcv);
        FastDetector fd(12, 50);
        std::vector<cv::Point2i> bres = fd.getBresenhamCirclePoints(testImage, 25, 25);
        for (auto &p : bres) fd.putPixel(testImage, p);
        EXPECT_EQ(fd.checkContiguousPixels(testImage.getPixelVal(25, 25), bres, testImage), true);
        fd.putPixel(testImage, cv::Point(25, 25));
        EXPECT_EQ(fd.checkContiguousPixels(testImage.getPixelVal(25, 25), bres, testImage), false);
    }
    {   // CheckDiscontinuous
        cv::Mat testcv(cv::Size(50, 50), CV_8UC1, cv::Scalar(0));
        Image testImage(testcv);
        FastDetector fd(12, 50);
        std::vector<cv::Point2i> bres = fd.getBresenhamCirclePoints(testImage, 25, 25);
        for (int i = 0; i < 11; i++) fd.putPixel(testImage, cv::Point(bres[i].x, bres[i].y));
        EXPECT_EQ(fd.checkContiguousPixels(testImage.getPixelVal(25, 25), bres, testImage), false);
    }
    {   // value semantics the reference relies on (LoopHandler is copied with its detector and Brief: src/main.cc:11)
        Brief a(256);
        Brief b = a;
        EXPECT_EQ(a.getOffsets().size(), (size_t)256);
        EXPECT_TRUE(a.getOffsets() == b.getOffsets());
        for (auto &row : a.getOffsets())
            for (int v : row) EXPECT_TRUE(v >= -8 && v <= 8);
        uchar x[32] = {0}, y[32] = {0};
        x[0] = 0xff; y[31] = 0x0f;
        EXPECT_EQ(a.hammingDistance(x, y), 12);
        EXPECT_EQ(a.popCount(0x5a), 4);
        EXPECT_TRUE(a.checkBoundry(8, 8, 16, 16));
        EXPECT_TRUE(!a.checkBoundry(7, 8, 100, 100));
        EXPECT_TRUE(!a.checkBoundry(93, 8, 100, 100));
        FastDetector f1(12, 50), f2 = f1;
        (void)f2;
        std::vector<Matches> none, outv;
        a.removeOutliers(none, outv, 20);
        EXPECT_EQ(outv.size(), (size_t)0);
        // Sobel helper: reference loop bounds leave the last two rows/cols untouched
        cv::Mat im(cv::Size(12, 10), CV_8UC1, cv::Scalar(0));
        for (int r = 0; r < 10; r++) for (int c = 0; c < 12; c++) im.at<uchar>(r, c) = (uchar)(r * 13 + c * 7);
        Image I(im);
        cv::Mat Ix = cv::Mat::zeros(10, 12, CV_32FC1), Iy = cv::Mat::zeros(10, 12, CV_32FC1);
        f1.preComputeHarris(I, Ix, Iy);
        EXPECT_EQ(Ix.at<float>(9, 5), 0.f);
        EXPECT_EQ(Ix.at<float>(4, 11), 0.f);
        EXPECT_EQ(Ix.at<float>(4, 5), 8.f * 7.f);
        EXPECT_EQ(Iy.at<float>(4, 5), 8.f * 13.f);
    }
    std::printf("known-answer tests: %s (%d failures)\n", g_fail ? "FAILED" : "ok", g_fail);
    return g_fail ? 1 : 0;
}

template <typename T>
static void put(std::ofstream &f, const T &v) { f.write(reinterpret_cast<const char *>(&v), sizeof(T)); }

static int pipeline(char **a) {
    const int H = std::atoi(a[2]), W = std::atoi(a[3]);
    std::vector<uint8_t> A = slurp(a[0]), B = slurp(a[1]), O = slurp(a[4]);
    if ((int)A.size() != H * W || (int)B.size() != H * W || O.size() != 1024 * 4) {
        std::printf("bad inputs\n");
        return 2;
    }
    std::vector<std::vector<int>> table(256, std::vector<int>(4));
    const int32_t *o = reinterpret_cast<const int32_t *>(O.data());
    for (int j = 0; j < 256; j++) for (int k = 0; k < 4; k++) table[j][k] = o[4 * j + k];

    Brief brief(256);
    brief.setOffsets(table);
    cv::Mat testImage1(H, W, CV_8UC1, A.data()), testImage2(H, W, CV_8UC1, B.data());
    Image testObj1(testImage1), testObj2(testImage2);
    FastDetector fd(12, 50);
    {   // warm-up on a throw-away copy (context creation, first launches), then the reference's timing printouts
        Image warm(testImage2);
        auto w = fd.getFastFeatures(warm);
        brief.computeBrief(w, warm);
    }
    auto start = std::chrono::high_resolution_clock::now();
    auto features1 = fd.getFastFeatures(testObj1);
    auto stop = std::chrono::high_resolution_clock::now();
    std::cout << "Time taken for FAST feature detection: "
              << std::chrono::duration_cast<std::chrono::microseconds>(stop - start).count() << " us" << std::endl;
    std::vector<float> scores1 = fd.lastScores();
    auto features2 = fd.getFastFeatures(testObj2);
    start = std::chrono::high_resolution_clock::now();
    brief.computeBrief(features1, testObj1);
    stop = std::chrono::high_resolution_clock::now();
    std::cout << "Time taken for BRIEF descriptor computation: "
              << std::chrono::duration_cast<std::chrono::microseconds>(stop - start).count() << " us" << std::endl;
    brief.computeBrief(features2, testObj2);
    start = std::chrono::high_resolution_clock::now();
    std::vector<Matches> matches = brief.matchFeatures(testObj1, testObj2);
    stop = std::chrono::high_resolution_clock::now();
    std::cout << "Time taken for matchFeatures: "
              << std::chrono::duration_cast<std::chrono::microseconds>(stop - start).count() << " us" << std::endl;
    std::vector<Matches> filterMatches;
    brief.removeOutliers(matches, filterMatches, 20.0);
    cv::Mat sideBySide = brief.drawMatches(testObj1, testObj2, filterMatches);
    // a second computeBrief appends again (reference behaviour, src/BriefDescriptor.cc:121)
    Image again(testImage1);
    brief.computeBrief(features1, again);
    const size_t once = again.keypoints.size();
    brief.computeBrief(features1, again);
    EXPECT_EQ(again.keypoints.size(), 2 * once);
    // host pixels mutated after a detect: the next call must see the new pixels
    Image mut(testImage1);
    auto fa = fd.getFastFeatures(mut);
    for (int r = 0; r < H; r++) for (int c = 0; c < W; c++) mut.rawImage.at<uchar>(r, c) = 7;
    auto fb = fd.getFastFeatures(mut);
    EXPECT_EQ(fb.size(), (size_t)0);
    EXPECT_EQ(sideBySide.cols, 2 * W);

    std::ofstream f(a[5], std::ios::binary);
    put<int32_t>(f, (int32_t)features1.size());
    for (size_t i = 0; i < features1.size(); i++) { put<int32_t>(f, features1[i].x); put<int32_t>(f, features1[i].y); put<float>(f, scores1[i]); }
    put<int32_t>(f, (int32_t)features2.size());
    for (auto &p : features2) { put<int32_t>(f, p.x); put<int32_t>(f, p.y); }
    for (Image *im : {&testObj1, &testObj2}) {
        put<int32_t>(f, (int32_t)im->keypoints.size());
        for (auto &k : im->keypoints) { put<int32_t>(f, k.x); put<int32_t>(f, k.y); put<int32_t>(f, k.id); f.write((const char *)k.featVec, 32); }
    }
    put<int32_t>(f, (int32_t)matches.size());
    for (auto &m : matches) { put<int32_t>(f, m.pt1.id); put<int32_t>(f, m.pt2.id); put<int32_t>(f, m.pt2.x); put<int32_t>(f, m.pt2.y); put<int32_t>(f, m.distance); put<int32_t>(f, (int32_t)m.pt1.matched); }
    put<int32_t>(f, (int32_t)filterMatches.size());
    for (auto &m : filterMatches) { put<int32_t>(f, m.pt1.id); put<int32_t>(f, m.distance); }
    f.close();
    yavo_host::Device::shutdown();
    std::printf("pipeline: %zu/%zu features, %zu/%zu keypoints, %zu matches, %zu kept (%d failures)\n", features1.size(),
                features2.size(), testObj1.keypoints.size(), testObj2.keypoints.size(), matches.size(), filterMatches.size(), g_fail);
    return g_fail ? 1 : 0;
}

// LoopHandler::trackLastFrame's OpenCV call (src/LoopHandler.cc:372-375) through yavo::calcOpticalFlowPyrLK:
// the FAST keypoints of frame A, swapped to (x = col, y = row) as the reference does at :343-347, tracked into B.
static int track(char **a) {
    const int H = std::atoi(a[2]), W = std::atoi(a[3]);
    std::vector<uint8_t> A = slurp(a[0]), B = slurp(a[1]);
    if ((int)A.size() != H * W || (int)B.size() != H * W) {
        std::printf("bad inputs\n");
        return 2;
    }
    cv::Mat m1(H, W, CV_8UC1, A.data()), m2(H, W, CV_8UC1, B.data());
    Image lastFrame(m1), currentFrame(m2);
    FastDetector fd(12, 50);
    auto features = fd.getFastFeatures(lastFrame);
    std::vector<cv::Point2f> lastFrameKpt, currFrameKpt;
    for (auto &p : features) lastFrameKpt.push_back(cv::Point2f((float)p.y, (float)p.x));
    std::vector<uchar> flowStatus;
    std::vector<float> error;
    yavo::calcOpticalFlowPyrLK(lastFrame, currentFrame, lastFrameKpt, currFrameKpt, flowStatus, error, cv::Size(11, 11), 3,
                               cv::TermCriteria(cv::TermCriteria::COUNT + cv::TermCriteria::EPS, 30, 0.01), 0, 0.001);
    {   // wall time of one call with both frames resident and their pyramids built (the steady state of the VO loop)
        std::vector<cv::Point2f> tmp;
        std::vector<uchar> st;
        std::vector<float> er;
        auto start = std::chrono::high_resolution_clock::now();
        yavo::calcOpticalFlowPyrLK(lastFrame, currentFrame, lastFrameKpt, tmp, st, er, cv::Size(11, 11), 3,
                                   cv::TermCriteria(cv::TermCriteria::COUNT + cv::TermCriteria::EPS, 30, 0.01), 0, 0.001);
        auto stop = std::chrono::high_resolution_clock::now();
        std::cout << "Time taken for calcOpticalFlowPyrLK (" << lastFrameKpt.size() << " points): "
                  << std::chrono::duration_cast<std::chrono::microseconds>(stop - start).count() << " us" << std::endl;
    }
    // the cv::Mat form (what the reference passes) must agree
    std::vector<cv::Point2f> again;
    std::vector<uchar> st2;
    std::vector<float> er2;
    yavo::calcOpticalFlowPyrLK(lastFrame.rawImage, currentFrame.rawImage, lastFrameKpt, again, st2, er2, cv::Size(11, 11), 3,
                               cv::TermCriteria(cv::TermCriteria::COUNT + cv::TermCriteria::EPS, 30, 0.01), 0, 0.001);
    EXPECT_EQ(again.size(), currFrameKpt.size());
    for (size_t i = 0; i < again.size(); i++) {
        EXPECT_TRUE(again[i] == currFrameKpt[i]);
        EXPECT_EQ((int)st2[i], (int)flowStatus[i]);
    }
    std::ofstream f(a[4], std::ios::binary);
    put<int32_t>(f, (int32_t)lastFrameKpt.size());
    for (size_t i = 0; i < lastFrameKpt.size(); i++) {
        put<float>(f, lastFrameKpt[i].x); put<float>(f, lastFrameKpt[i].y);
        put<float>(f, currFrameKpt[i].x); put<float>(f, currFrameKpt[i].y);
        put<float>(f, error[i]); put<int32_t>(f, (int32_t)flowStatus[i]);
    }
    f.close();
    yavo_host::Device::shutdown();
    std::printf("track: %zu points (%d failures)\n", lastFrameKpt.size(), g_fail);
    return g_fail ? 1 : 0;
}

// The call shape the reference really uses, timed: ONE frame at a time through FastDetector::getFastFeatures ->
// Brief::computeBrief (LoopHandler::insertFrameFeatures, src/LoopHandler.cc:468-485) and Brief::matchFeatures on
// consecutive frames (src/LoopHandler.cc:189,534; tests/BriefDescriptorTest.cc:21-44).  Every frame is a fresh Image, as
// getNextFrame builds it (src/LoopHandler.cc:917-927).  Prints one JSON object: microseconds per call, p50 / p99 / mean.
static int latency(char **a) {
    const int n = std::atoi(a[1]), H = std::atoi(a[2]), W = std::atoi(a[3]);
    std::vector<uint8_t> F = slurp(a[0]), O = slurp(a[4]);
    const size_t fb = (size_t)H * W;
    if (n < 2 || F.size() < fb * 2 || O.size() != 1024 * 4) { std::printf("bad inputs\n"); return 2; }
    const int distinct = (int)(F.size() / fb);
    std::vector<std::vector<int>> table(256, std::vector<int>(4));
    const int32_t *o = reinterpret_cast<const int32_t *>(O.data());
    for (int j = 0; j < 256; j++) for (int k = 0; k < 4; k++) table[j][k] = o[4 * j + k];
    Brief brief(256);
    brief.setOffsets(table);
    FastDetector fd(12, 50);
    typedef std::chrono::steady_clock clk;
    auto us = [](clk::time_point t0, clk::time_point t1) { return std::chrono::duration<double, std::micro>(t1 - t0).count(); };
    std::vector<double> t_fast, t_brief, t_match, t_frame, t_image;
    std::unique_ptr<Image> prev;
    size_t kp_total = 0, kept_total = 0;
    for (int i = -8; i < n; i++) {  // 8 untimed frames first: context creation, graph capture, first launches
        cv::Mat m(H, W, CV_8UC1, F.data() + (size_t)((i + 8) % distinct) * fb);
        const auto t0 = clk::now();
        std::unique_ptr<Image> cur(new Image(m));  // deep copy, like Frame(img) in getNextFrame
        const auto t1 = clk::now();
        auto features = fd.getFastFeatures(*cur);
        const auto t2 = clk::now();
        brief.computeBrief(features, *cur);
        const auto t3 = clk::now();
        double tm = 0;
        if (prev) {
            std::vector<Matches> matches = brief.matchFeatures(*prev, *cur);
            const auto t4 = clk::now();
            std::vector<Matches> kept;
            brief.removeOutliers(matches, kept, 20);
            tm = us(t3, t4);
            if (i >= 0) kept_total += kept.size();
        }
        if (i >= 0) {
            t_fast.push_back(us(t1, t2));
            t_brief.push_back(us(t2, t3));
            if (prev) t_match.push_back(tm);
            t_frame.push_back(us(t1, t3));
            t_image.push_back(us(t0, t1));
            kp_total += cur->keypoints.size();
        }
        prev = std::move(cur);
    }
    auto stat = [](std::vector<double> v, const char *name) {
        std::sort(v.begin(), v.end());
        double mean = 0;
        for (double x : v) mean += x;
        mean /= std::max<size_t>(v.size(), 1);
        std::printf("\"%s\": {\"p50\": %.1f, \"p99\": %.1f, \"mean\": %.1f, \"min\": %.1f, \"calls\": %zu}", name, v[v.size() / 2],
                    v[std::min(v.size() - 1, (size_t)(v.size() * 0.99))], mean, v[0], v.size());
    };
    std::printf("{\"unit\": \"us per call\", \"frames\": %d, \"frame\": [%d, %d], \"mean_keypoints\": %.1f, \"mean_kept_matches\": %.1f, ", n, H, W,
                (double)kp_total / n, (double)kept_total / std::max(n - 1, 1));
    stat(t_fast, "getFastFeatures");
    std::printf(", ");
    stat(t_brief, "computeBrief");
    std::printf(", ");
    stat(t_match, "matchFeatures");
    std::printf(", ");
    stat(t_frame, "getFastFeatures+computeBrief");
    std::printf(", ");
    stat(t_image, "Image(cv::Mat) deep copy (host only, src/Image.cc:8-13)");
    std::printf("}\n");
    yavo_host::Device::shutdown();
    return 0;
}

// yavo::FrameStream (include/FrameStream.hpp) over a sequence held in one file: per-frame results dumped for the pytest side
static int stream_mode(char **a) {
    const int n = std::atoi(a[1]), H = std::atoi(a[2]), W = std::atoi(a[3]), batch = std::atoi(a[5]);
    std::vector<uint8_t> F = slurp(a[0]), O = slurp(a[4]);
    const size_t fb = (size_t)H * W;
    if (F.size() != fb * n || O.size() != 1024 * 4) { std::printf("bad inputs\n"); return 2; }
    std::ofstream f(a[6], std::ios::binary);
    int seen = 0, next = 0;
    yavo::FrameStream fs(0, n, H, W, batch, reinterpret_cast<const int32_t *>(O.data()), 2000, 2);
    const int delivered = fs.run(
        [&](int frame, uint8_t *dst) { std::memcpy(dst, F.data() + (size_t)frame * fb, fb); },
        [&](const yavo::FrameResult &r) {
            EXPECT_EQ(r.frame, next);  // in order, each frame once
            next++;
            seen++;
            put<int32_t>(f, r.frame);
            put<int32_t>(f, r.n_kp);
            f.write((const char *)r.rows, 4 * (size_t)r.n_kp);
            f.write((const char *)r.cols, 4 * (size_t)r.n_kp);
            f.write((const char *)r.desc, 32 * (size_t)r.n_kp);
            put<int32_t>(f, r.n_prev);
            if (r.n_prev) {
                f.write((const char *)r.match_idx, 4 * (size_t)r.n_prev);
                f.write((const char *)r.match_dist, 4 * (size_t)r.n_prev);
            }
        });
    EXPECT_EQ(delivered, n);
    EXPECT_EQ(seen, n);
    f.close();
    std::printf("stream: %d frames in batches of %d (%d failures)\n", delivered, batch, g_fail);
    return g_fail ? 1 : 0;
}

int main(int argc, char **argv) {
    try {
        if (argc == 7 && !std::strcmp(argv[1], "latency")) return latency(argv + 2);
        if (argc == 9 && !std::strcmp(argv[1], "stream")) return stream_mode(argv + 2);
        if (argc == 3 && !std::strcmp(argv[1], "known")) return known(argv[2]);
        if (argc == 8 && !std::strcmp(argv[1], "pipeline")) return pipeline(argv + 2);
        if (argc == 7 && !std::strcmp(argv[1], "track")) return track(argv + 2);
    } catch (const std::exception &e) {
        std::printf("exception: %s\n", e.what());
        return 3;
    }
    std::printf("usage: host_tests known <bres.bin> | pipeline <A.bin> <B.bin> <H> <W> <offsets.bin> <out.bin> | track <A.bin> <B.bin> <H> <W> <out.bin>\n"
                "       | latency <frames.bin> <n> <H> <W> <offsets.bin> | stream <frames.bin> <n> <H> <W> <offsets.bin> <batch> <out.bin>\n");
    return 2;
}
