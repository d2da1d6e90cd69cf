#include "../include/yavo_device.hpp"

#include <algorithm>
#include <cstdlib>

#include "../include/Image.hpp"

namespace yavo_host {

namespace {
Device *g_device = nullptr;
std::mutex g_mu;
}  // namespace

Device::Device(int rows, int cols) : maxRows_(rows), maxCols_(cols), slots_(kSlots) {
    const char *v = std::getenv("YAVO_TRUST_IMAGE_IDENTITY");
    verify_pixels_ = !(v && *v && *v != '0');
    // YAVO_DEVICE selects the GPU of this process's context (default 0); one process per GPU is the multi-GPU shape
    const char *dv = std::getenv("YAVO_DEVICE");
    const int device = dv && *dv ? std::atoi(dv) : 0;
    const int rc = yavo_create(device, kSlots, rows, cols, kMaxKeypoints, 0, &ctx_);
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
    offsets_epoch_++;
}

bool Device::matches(int slot, const Image &img) const {
    const Slot &s = slots_[slot];
    const cv::Mat &m = img.rawImage;
    if (!(s.used && s.id == img.yavoId() && s.gen == img.yavoGeneration() && s.data == (const void *)m.data && s.rows == m.rows &&
          s.cols == m.cols))
        return false;
    if (!verify_pixels_) return true;
    // rawImage is public and mutable: the identity is confirmed against the pixels themselves — byte for byte against the
    // pinned copy the fused call kept (one memcmp of a cache-resident 0.47 MB), a checksum for slots filled by yavo_upload
    if (s.shadow) return yavo_slot_holds(ctx_, slot, m.data, m.rows, m.cols, (int)m.step) == 1;
    return s.checksum == img.yavoChecksum();
}

int Device::findSlot(const Image &img) const {
    for (int s = 0; s < kSlots; s++)
        if (matches(s, img)) return s;
    return -1;
}

int Device::claimSlot(const Image &img) {
    int victim = -1;
    for (int s = 0; s < kSlots && victim < 0; s++)
        if (slots_[s].used && slots_[s].id == img.yavoId()) victim = s;  // the same Image with new pixels: its own slot
    for (int s = 0; s < kSlots && victim < 0; s++)
        if (!slots_[s].used) victim = s;
    if (victim < 0) {
        victim = 0;
        for (int s = 1; s < kSlots; s++)
            if (slots_[s].stamp < slots_[victim].stamp) victim = s;
    }
    const cv::Mat &m = img.rawImage;
    if (m.empty() || m.type() != CV_8UC1) throw DeviceError("Image must hold a non-empty CV_8UC1 frame");
    Slot &sl = slots_[victim];
    sl.used = true;
    sl.id = img.yavoId();
    sl.gen = img.yavoGeneration();
    sl.data = m.data;
    sl.rows = m.rows;
    sl.cols = m.cols;
    sl.checksum = 0;
    sl.shadow = false;
    sl.stamp = ++clock_;
    sl.feat.valid = false;
    return victim;
}

int Device::slotFor(const Image &img) {
    const int hit = findSlot(img);
    if (hit >= 0) {
        slots_[hit].stamp = ++clock_;
        return hit;  // pixels (and any blurred plane) already resident
    }
    const int victim = claimSlot(img);
    const cv::Mat &m = img.rawImage;
    check(yavo_upload(ctx_, victim, m.data, m.rows, m.cols, (int)m.step));
    if (verify_pixels_) slots_[victim].checksum = img.yavoChecksum();
    return victim;
}

}  // namespace yavo_host
