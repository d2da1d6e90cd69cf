#include "../include/yavo_device.hpp"

#include <algorithm>

#include "../include/Image.hpp"

namespace yavo_host {

namespace {
Device *g_device = nullptr;
std::mutex g_mu;
}  // namespace

Device::Device(int rows, int cols) : maxRows_(rows), maxCols_(cols), slots_(kSlots) {
    const int rc = yavo_create(0, kSlots, rows, cols, kMaxKeypoints, 0, &ctx_);
    if (rc != 0) throw DeviceError(std::string("yavo_create failed: ") + yavo_last_error(nullptr));
}

Device::~Device() { yavo_destroy(ctx_); }

Device &Device::instance(int rows, int cols) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_device && (g_device->maxRows_ < rows || g_device->maxCols_ < cols)) {
        const int r = std::max(rows, g_device->maxRows_), c = std::max(cols, g_device->maxCols_);
        delete g_device;
        g_device = new Device(r, c);
    }
    if (!g_device) g_device = new Device(std::max(rows, 376), std::max(cols, 1241));
    return *g_device;
}

void Device::shutdown() {
    std::lock_guard<std::mutex> lk(g_mu);
    delete g_device;
    g_device = nullptr;
}

void Device::check(int rc) const {
    if (rc < 0) throw DeviceError(std::string("yavo: ") + yavo_last_error(ctx_));
}

void Device::setBriefOffsets(const int32_t *table1024) {
    if (offsets_.size() == 1024 && std::equal(offsets_.begin(), offsets_.end(), table1024)) return;
    check(yavo_set_brief_offsets(ctx_, table1024));
    offsets_.assign(table1024, table1024 + 1024);
}

int Device::slotFor(const Image &img) {
    const uint64_t id = img.yavoId(), sum = img.yavoChecksum();
    int victim = 0;
    for (int s = 0; s < kSlots; s++) {
        if (slots_[s].used && slots_[s].id == id && slots_[s].checksum == sum) {
            slots_[s].stamp = ++clock_;
            return s;  // pixels (and any blurred plane) already resident
        }
        if (!slots_[s].used) victim = s;
    }
    if (slots_[victim].used)
        for (int s = 0; s < kSlots; s++)
            if (slots_[s].stamp < slots_[victim].stamp) victim = s;
    const cv::Mat &m = img.rawImage;
    if (m.empty() || m.type() != CV_8UC1) throw DeviceError("Image must hold a non-empty CV_8UC1 frame");
    check(yavo_upload(ctx_, victim, m.data, m.rows, m.cols, (int)m.step));
    slots_[victim].used = true;
    slots_[victim].id = id;
    slots_[victim].checksum = sum;
    slots_[victim].stamp = ++clock_;
    return victim;
}

}  // namespace yavo_host
