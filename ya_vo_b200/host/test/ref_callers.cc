// ref_callers — source-compatibility proof: the bodies of the reference's own CALLERS of the hot path are compiled
// VERBATIM against ya_vo_b200/host/include and run on the GPU.  Nothing of the reference is copied into this repository:
// the Makefile cuts the line ranges out of the reference checkout at build time into test/_gen/*.inc (git-ignored), and
// this file only supplies the surroundings those lines expect.
//
//   _gen/brief_test_body.inc        = /root/reference/tests/BriefDescriptorTest.cc:10-47
//        Brief brief(256); LoopHandler Lh(config); imread x2; Image x2; FastDetector fd(12, 50); getFastFeatures x2 with
//        the chrono printouts; computeBrief x2; matchFeatures; removeOutliers(…, 20.0); drawMatches
//   _gen/insert_frame_features.inc  = /root/reference/src/LoopHandler.cc:469-484
//        the body of LoopHandler::insertFrameFeatures(Frame::ptr): getFastFeatures(*_frame), computeBrief(features, *_frame)
//
//   ref_callers <frameA.bin> <frameB.bin> <H> <W> <out.bin>
// dumps the (randomly drawn, src/BriefDescriptor.cc:4-20) offset table and every result, so the pytest side can replay
// the run through the oracle with the same table.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include <opencv2/imgproc.hpp>

#include "../include/BriefDescriptor.hpp"
#include "../include/FastDetector.hpp"
#include "../include/Image.hpp"
#include "../include/yavo_device.hpp"

static std::vector<uint8_t> g_frames[2];
static int g_H = 0, g_W = 0;

// ---- what the verbatim lines reach for besides the three drop-in classes --------------------------------------------
// tests/BriefDescriptorTest.cc:12-15 builds a LoopHandler only to list image paths and reads two files with cv::imread
struct LoopHandler {
    std::vector<std::string> leftPathTrain;
    explicit LoopHandler(const std::string &) { leftPathTrain = {"0", "1"}; }
};
namespace cv {
#ifdef YAVO_CVSHIM_CORE_HPP  // the stand-in has no imgcodecs; with real OpenCV the files would be read from disk
inline Mat imread(const std::string &path, int /*flags*/) {
    const int i = std::atoi(path.c_str());
    Mat view(g_H, g_W, CV_8UC1, g_frames[i].data());
    Mat own = Mat::zeros(g_H, g_W, CV_8UC1);
    view.copyTo(own);
    return own;
}
#endif
}  // namespace cv

// src/LoopHandler.cc is written with these in scope (include/LoopHandler.hpp, Utils.hpp)
using namespace std;
struct Frame : public Image {  // include/Frame.hpp:10: class Frame : public Image
    typedef std::shared_ptr<Frame> ptr;
    explicit Frame(const cv::Mat &m) : Image(m) {}
};
struct LoopHandlerSlice {  // the two members insertFrameFeatures uses (include/LoopHandler.hpp:47-48)
    FastDetector fd{12, 50};
    Brief brief{256};
    void insertFrameFeatures(Frame::ptr _frame) {
#include "_gen/insert_frame_features.inc"
    }
};

template <typename T>
static void put(std::ofstream &f, const T &v) { f.write(reinterpret_cast<const char *>(&v), sizeof(T)); }
static void put_table(std::ofstream &f, const Brief &b) {
    for (auto &row : b.getOffsets())
        for (int v : row) put<int32_t>(f, v);
}
static void put_keypoints(std::ofstream &f, const Image &im) {
    put<int32_t>(f, (int32_t)im.keypoints.size());
    for (auto &k : im.keypoints) { put<int32_t>(f, k.x); put<int32_t>(f, k.y); put<int32_t>(f, k.id); f.write((const char *)k.featVec, 32); }
}

int main(int argc, char **argv) {
    if (argc != 6) {
        std::printf("usage: ref_callers <A.bin> <B.bin> <H> <W> <out.bin>\n");
        return 2;
    }
    try {
        g_H = std::atoi(argv[3]);
        g_W = std::atoi(argv[4]);
        for (int i = 0; i < 2; i++) {
            std::ifstream f(argv[1 + i], std::ios::binary);
            g_frames[i].assign((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
            if ((int)g_frames[i].size() != g_H * g_W) { std::printf("bad frame %d\n", i); return 2; }
        }
        std::ofstream out(argv[5], std::ios::binary);
        {
            // ---- tests/BriefDescriptorTest.cc:10-47, verbatim ----
#include "_gen/brief_test_body.inc"
            // ---- end of the verbatim lines: dump what they computed ----
            put_table(out, brief);
            put<int32_t>(out, (int32_t)features1.size());
            for (auto &p : features1) { put<int32_t>(out, p.x); put<int32_t>(out, p.y); }
            put<int32_t>(out, (int32_t)features2.size());
            for (auto &p : features2) { put<int32_t>(out, p.x); put<int32_t>(out, p.y); }
            put_keypoints(out, testObj1);
            put_keypoints(out, testObj2);
            put<int32_t>(out, (int32_t)matches.size());
            for (auto &m : matches) { put<int32_t>(out, m.pt1.id); put<int32_t>(out, m.pt2.id); put<int32_t>(out, m.distance); }
            put<int32_t>(out, (int32_t)filterMatches.size());
            for (auto &m : filterMatches) { put<int32_t>(out, m.pt1.id); put<int32_t>(out, m.distance); }
            put<int32_t>(out, sideBySide.cols);
        }
        {
            // ---- src/LoopHandler.cc:468-485 through the slice above, on both frames ----
            LoopHandlerSlice lh;
            put_table(out, lh.brief);
            for (int i = 0; i < 2; i++) {
                cv::Mat m(g_H, g_W, CV_8UC1, g_frames[i].data());
                Frame::ptr frame(new Frame(m));
                lh.insertFrameFeatures(frame);
                put_keypoints(out, *frame);
            }
        }
        out.close();
        yavo_host::Device::shutdown();
    } catch (const std::exception &e) {
        std::printf("exception: %s\n", e.what());
        return 3;
    }
    std::printf("ref_callers ok\n");
    return 0;
}
