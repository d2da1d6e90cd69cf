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
    // slot holding img's current pixels; uploads them if the slot table has no (id, checksum) match
    int slotFor(const Image &img);
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
        uint64_t id = 0, checksum = 0, stamp = 0;
        bool used = false;
    };
    yavo_ctx *ctx_ = nullptr;
    int maxRows_ = 0, maxCols_ = 0;
    std::vector<Slot> slots_;
    uint64_t clock_ = 0;
    std::vector<int32_t> offsets_;  // table currently on the device
    std::mutex mu_;
};

}  // namespace yavo_host
#endif
