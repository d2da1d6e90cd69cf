// Process-wide device context shared by the drop-in classes.  FastDetector and Brief are held by value
// inside the reference's LoopHandler and copied with it (src/main.cc:11), so they keep no device
// handle of their own: they ask this registry, which owns one yavo_ctx (include/yavo_b200.h) per
// process and maps Images to device slots.
#ifndef YAVO_HOST_DEVICE_HPP
#define YAVO_HOST_DEVICE_HPP

#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/yavo_b200.h"

class Image;

namespace yavo_host {

class DeviceError : public std::runtime_error {
   public:
    explicit DeviceError(const std::string &m) : std::runtime_error(m) {}
};

class Device {
   public:
    // the context able to hold rows x cols frames (created or re-created larger on demand)
    static Device &instance(int rows, int cols);
    static void shutdown();

    yavo_ctx *ctx() const { return ctx_; }
    // slot holding img's current pixels; uploads them if no slot matches the Image's identity
    int slotFor(const Image &img);
    // slot with a matching identity, or -1 (never uploads)
    int findSlot(const Image &img) const;
    // slot to overwrite with img (least recently used); its record takes img's identity, its cached features are dropped
    int claimSlot(const Image &img);
    bool hasBriefOffsets() const { return offsets_.size() == 1024; }
    uint64_t offsetsEpoch() const { return offsets_epoch_; }

    // what the fused single-frame call (yavo_frame_features) left for a slot: the detector's output and, for the points
    // checkBoundry admits, their index in that list and descriptor — Brief::computeBrief on the same points needs no
    // further device work
    struct Features {
        bool valid = false;
        int cap = 0;
        uint64_t offsets_epoch = 0;
        int n_cand = 0;
        std::vector<int32_t> rows, cols;  // top-K in the reference's order
        std::vector<float> scores;
        std::vector<int32_t> ids;         // admitted points: index into rows/cols
        std::vector<uint8_t> desc;        // 32 bytes each
    };
    Features &features(int slot) { return slots_[slot].feat; }
    void markShadow(int slot) { slots_[slot].shadow = true; }
    void check(int rc) const;  // throws DeviceError with yavo_last_error on rc < 0
    // passes the 256 x 4 BRIEF table to the device unless it is the one already there
    void setBriefOffsets(const int32_t *table1024);
    std::mutex &mutex() { return mu_; }

    static const int kSlots = 4;
    static const int kMaxKeypoints = 2000;  // include/FastDetector.hpp:36

   private:
    Device(int rows, int cols);
    ~Device();
    struct Slot {
        uint64_t id = 0, gen = 0, checksum = 0, stamp = 0;
        const void *data = nullptr;
        int rows = 0, cols = 0;
        bool used = false;
        bool shadow = false;  // filled by yavo_frame_features: the C ABI holds a pinned copy of the pixels to compare with
        Features feat;
    };
    bool matches(int slot, const Image &img) const;
    yavo_ctx *ctx_ = nullptr;
    int maxRows_ = 0, maxCols_ = 0;
    std::vector<Slot> slots_;
    uint64_t clock_ = 0;
    std::vector<int32_t> offsets_;  // table currently on the device
    uint64_t offsets_epoch_ = 0;    // bumped whenever a different table goes to the device
    bool verify_pixels_ = true;     // YAVO_TRUST_IMAGE_IDENTITY=1 switches the pixel comparison off
    std::mutex mu_;
};

}  // namespace yavo_host
#endif
