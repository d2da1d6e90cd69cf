#include "../include/FrameStream.hpp"

#include <condition_variable>
#include <deque>
#include <exception>
#include <mutex>
#include <stdexcept>
#include <thread>

namespace yavo {

namespace {
template <typename T>
T *pinned(size_t n) {
    void *p = yavo_pinned_alloc(n * sizeof(T));
    if (!p) throw std::runtime_error("yavo_pinned_alloc failed");
    return static_cast<T *>(p);
}
}  // namespace

FrameStream::FrameStream(int device, int n_frames, int rows, int cols, int batch, const int32_t *offsets, int max_kp, int prefetch)
    : n_(n_frames), rows_(rows), cols_(cols), batch_(batch), K_(max_kp), depth_(prefetch + 2) {
    if (batch < 2) throw std::invalid_argument("FrameStream needs batch >= 2 (every batch after the first re-uses one frame as its seam)");
    if (n_frames < 0 || rows < 1 || cols < 1 || max_kp < 1 || prefetch < 0) throw std::invalid_argument("FrameStream: bad sizes");
    const int rc = yavo_create(device, batch, rows, cols, max_kp, 0, &ctx_);
    if (rc != 0) throw std::runtime_error(std::string("yavo_create failed: ") + yavo_last_error(nullptr));
    try {
        check(yavo_set_brief_offsets(ctx_, offsets));
        for (int i = 0; i < depth_; i++) in_.push_back(pinned<uint8_t>((size_t)batch * rows * cols));
        for (Out &o : out_) {
            const size_t BK = (size_t)batch * max_kp;
            o.n_kp = pinned<int32_t>(batch);
            o.rows = pinned<int32_t>(BK);
            o.cols = pinned<int32_t>(BK);
            o.midx = pinned<int32_t>(BK);
            o.mdist = pinned<int32_t>(BK);
            o.scores = pinned<float>(BK);
            o.desc = pinned<uint8_t>(BK * 32);
        }
    } catch (...) {
        yavo_destroy(ctx_);
        ctx_ = nullptr;
        throw;
    }
}

FrameStream::~FrameStream() {
    if (ctx_) {
        yavo_wait(ctx_);
        yavo_destroy(ctx_);
    }
    for (uint8_t *p : in_) yavo_pinned_free(p);
    for (Out &o : out_) {
        yavo_pinned_free(o.n_kp);
        yavo_pinned_free(o.rows);
        yavo_pinned_free(o.cols);
        yavo_pinned_free(o.midx);
        yavo_pinned_free(o.mdist);
        yavo_pinned_free(o.scores);
        yavo_pinned_free(o.desc);
    }
}

void FrameStream::check(int rc) const {
    if (rc < 0) throw std::runtime_error(std::string("yavo: ") + yavo_last_error(ctx_));
}

void FrameStream::deliver(int a, int b, const Out &o, const OnFrame &on_frame, bool do_match) {
    for (int i = 0; i < b - a; i++) {
        if (i == 0 && a > 0) continue;  // the seam frame went out with the previous batch
        FrameResult r;
        r.frame = a + i;
        r.n_kp = o.n_kp[i];
        const size_t off = (size_t)i * K_;
        r.rows = o.rows + off;
        r.cols = o.cols + off;
        r.scores = o.scores + off;
        r.desc = o.desc + off * 32;
        if (do_match && r.frame > 0) {  // row i of a batch holds the match of (frame - 1, frame)
            r.n_prev = o.n_kp[i - 1];
            r.match_idx = o.midx + off;
            r.match_dist = o.mdist + off;
        }
        if (on_frame) on_frame(r);
        delivered_++;
    }
}

int FrameStream::run(const Reader &read, const OnFrame &on_frame, bool do_match) {
    struct Item {
        int a, b, buf;
    };
    std::mutex mu;
    std::condition_variable cv_ready, cv_free;
    std::deque<Item> ready;
    std::deque<int> free_bufs;
    for (int i = 0; i < depth_; i++) free_bufs.push_back(i);
    bool done = false;
    std::exception_ptr err;
    const size_t fbytes = (size_t)rows_ * cols_;

    // decoder thread: the reference's getNextFrame, running ahead of the GPU
    std::thread decoder([&] {
        try {
            int a = 0;
            while (a < n_) {
                const int b = std::min(n_, a + batch_);
                int buf;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv_free.wait(lk, [&] { return !free_bufs.empty() || done; });
                    if (done) return;
                    buf = free_bufs.front();
                    free_bufs.pop_front();
                }
                for (int f = a; f < b; f++) read(f, in_[buf] + (size_t)(f - a) * fbytes);
                {
                    std::lock_guard<std::mutex> lk(mu);
                    ready.push_back({a, b, buf});
                }
                cv_ready.notify_one();
                if (b >= n_) break;
                a = b - 1;  // seam: the next batch starts with this batch's last frame
            }
        } catch (...) {
            std::lock_guard<std::mutex> lk(mu);
            err = std::current_exception();
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            ready.push_back({-1, -1, -1});
        }
        cv_ready.notify_one();
    });

    delivered_ = 0;
    struct Pending {
        Item it;
        int out, ticket;
        bool valid;
    } pend = {{0, 0, 0}, 0, 0, false};
    int turn = 0;
    std::exception_ptr fail;
    try {
        for (;;) {
            Item it;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_ready.wait(lk, [&] { return !ready.empty(); });
                it = ready.front();
                ready.pop_front();
            }
            if (it.buf < 0) break;
            const Out &o = out_[turn];
            const int t = yavo_submit_host_batch(ctx_, in_[it.buf], it.b - it.a, rows_, cols_, do_match ? 1 : 0, o.n_kp, o.rows, o.cols,
                                                 o.scores, o.desc, o.midx, o.mdist);
            check(t);
            if (pend.valid) {  // the previous batch finishes while this one is in flight
                check(yavo_wait_batch(ctx_, pend.ticket));
                {
                    std::lock_guard<std::mutex> lk(mu);
                    free_bufs.push_back(pend.it.buf);
                }
                cv_free.notify_one();
                deliver(pend.it.a, pend.it.b, out_[pend.out], on_frame, do_match);
            }
            pend = {it, turn, t, true};
            turn ^= 1;
        }
        if (pend.valid) {
            check(yavo_wait(ctx_));
            deliver(pend.it.a, pend.it.b, out_[pend.out], on_frame, do_match);
        }
    } catch (...) {
        fail = std::current_exception();
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        done = true;
    }
    cv_free.notify_all();
    decoder.join();
    if (ctx_) yavo_wait(ctx_);
    if (fail) std::rethrow_exception(fail);
    if (err) std::rethrow_exception(err);
    return delivered_;
}

}  // namespace yavo
