// FrameStream — the streaming caller of the front end in C++: the B200 counterpart of the reference's
// LoopHandler::getNextFrame + insertFrameFeatures pair (src/LoopHandler.cc:917-927, 468-485), which reads ONE frame
// (cv::imread(path, 0)), runs FAST + BRIEF on it and only then reads the next.
//
// Here a decoder thread fills page-locked batches ahead of the GPU while consecutive batches overlap through
// yavo_submit_host_batch / yavo_wait_batch (include/yavo_b200.h): decode, PCIe upload, kernels and result download all
// run concurrently; the caller gets one callback per frame, in sequence order, with the frame's keypoints and
// descriptors and its matches against the previous frame (Brief::matchFeatures(previous, current)).  Batches after the
// first start with the last frame of the previous one (the seam frame), so every consecutive pair is matched exactly
// once.  Same behaviour as ya_vo_b200/stream.py (FrameStream + run_sequence), for a maintainer who lives in C++.
#ifndef YAVO_HOST_FRAME_STREAM_HPP
#define YAVO_HOST_FRAME_STREAM_HPP

#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../../include/yavo_b200.h"

namespace yavo {

struct FrameResult {
    int frame = 0;                      // index in the sequence
    int n_kp = 0;                       // keypoints Brief::computeBrief would have appended for this frame
    const int32_t *rows = nullptr;      // [n_kp]  (x = row, y = col as everywhere in the reference)
    const int32_t *cols = nullptr;      // [n_kp]
    const float *scores = nullptr;      // [n_kp]  Harris responses
    const uint8_t *desc = nullptr;      // [n_kp][32]
    int n_prev = 0;                     // keypoints of frame - 1 (0 for the first frame or with matching off)
    const int32_t *match_idx = nullptr; // [n_prev]: keypoint of THIS frame matched to keypoint i of frame - 1
    const int32_t *match_dist = nullptr;// [n_prev]
};

class FrameStream {
   public:
    // read(frame, dst): decode frame `frame` into dst (rows * cols bytes, 8-bit gray) — the caller's cv::imread
    typedef std::function<void(int, uint8_t *)> Reader;
    typedef std::function<void(const FrameResult &)> OnFrame;

    FrameStream(int device, int n_frames, int rows, int cols, int batch, const int32_t *brief_offsets_1024, int max_kp = 2000,
                int prefetch = 2);
    ~FrameStream();
    FrameStream(const FrameStream &) = delete;
    FrameStream &operator=(const FrameStream &) = delete;

    // pushes the whole sequence through the front end; the pointers of a FrameResult are valid during the callback only.
    // Returns the number of frames delivered; throws std::runtime_error with the C ABI's message on failure.
    int run(const Reader &read, const OnFrame &on_frame, bool do_match = true);

   private:
    struct Out {
        int32_t *n_kp, *rows, *cols, *midx, *mdist;
        float *scores;
        uint8_t *desc;
    };
    void check(int rc) const;
    void deliver(int a, int b, const Out &o, const OnFrame &on_frame, bool do_match);
    yavo_ctx *ctx_ = nullptr;
    int n_, rows_, cols_, batch_, K_, depth_;
    std::vector<uint8_t *> in_;  // pinned input batches
    Out out_[2];
    int delivered_ = 0;
};

}  // namespace yavo
#endif
