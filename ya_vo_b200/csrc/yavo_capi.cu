// yavo_capi.cu — the C ABI of include/yavo_b200.h on top of the kernels in yavo_kernels.cuh.
// One context = one device, one stream, device-resident frame slots plus the per-slot work
// buffers.  No CPU fallback: every compute entry point launches the CUDA kernels or fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/yavo_b200.h"
#include "yavo_kernels.cuh"
#include "klt_kernels.cuh"
#include "match_tc4.cuh"

using namespace yavo;

namespace {
thread_local std::string g_create_error;
}

struct yavo_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t ls = nullptr;  // stream the feature-kernel launch helpers use: `stream`, or one of `aux` inside the overlapped batch path
    // overlapped feature pipeline (yavo_set_overlap): chunks of a batch rotate over the auxiliary streams so that the
    // issue-bound detect kernel of chunk i+1 runs beside the latency-bound select / BRIEF kernels of chunk i
    enum { MAX_AUX = 4 };
    cudaStream_t aux[MAX_AUX] = {};
    cudaEvent_t ev_fork = nullptr, ev_k1[MAX_AUX] = {}, ev_join[MAX_AUX] = {};
    int ov_chunk = 0, ov_streams = 3;    // frames per chunk (0 = no overlap, the default: measured no gain on B200), streams in rotation
    int n_slots = 0, max_rows = 0, max_cols = 0, max_kp = 0, max_cand = 0;
    int pitch = 0;       // device row pitch in bytes (multiple of 128)
    int rows_alloc = 0;  // rows per slot
    int seg_cols = 0;    // tile columns per row of the segment table (pitch / TW)
    size_t frame_stride = 0;
    std::vector<int> slot_rows, slot_cols;
    std::vector<char> slot_blur_valid;  // blurred plane of the slot is current
    // device buffers (per slot)
    uint8_t *d_frames = nullptr, *d_blur = nullptr;
    CUtensorMap frames_map;  // d_frames as a 3-D tensor (byte in row, row, slot): K1 stages a tile with one tensor copy
    CUtensorMap frames_cmap; // d_frames as a 4-D tensor (byte in 16-byte chunk, row, chunk, slot): the same tile in core-matrix order (blur_umma.cuh)
    bool batch_select_only = false;    // YAVO_SELECT_SINGLE=0: the batch instance of the select kernel on single frames too (A/B runs)
    uint8_t *d_blur_consts = nullptr;  // the blur's constant band matrices in their shared-memory layout (blur_umma.cuh)
    CUtensorMap blur_map;    // d_blur likewise, box = the 32 x 17 byte patch the BRIEF kernel stages per keypoint
    yavo_ent *d_pool = nullptr;  // per slot max_cand scored corners in tile order (written by K1)
    uint32_t *d_seg = nullptr;   // per slot rows_alloc x seg_cols: pool offset << 8 | count of each (row, tile column)
    yavo_ent *d_cand = nullptr;  // per slot max_cand: the select kernel's sort buffer when the list exceeds its shared memory
    int *d_ncand = nullptr;
    uint32_t *d_scratch = nullptr;
    // cluster pre-partition of large candidate lists (select_big_kernel): exchange block, handed-over ranges, their count
    int *d_xchg = nullptr, *d_npre = nullptr, *d_team_done = nullptr;
    SelRange *d_pre = nullptr;
    int big_min = -1;  // candidates above which a frame's list takes the cluster path; 0 = never, -1 = automatic (6144, frames >= 2 Mpx)
    int32_t *d_kp_row = nullptr, *d_kp_col = nullptr;
    float *d_kp_score = nullptr;
    int *d_nkp = nullptr;
    int32_t *d_bk_row = nullptr, *d_bk_col = nullptr, *d_bk_id = nullptr;
    float *d_bk_score = nullptr;
    int *d_nbk = nullptr;
    uint32_t *d_desc = nullptr;  // per slot max_kp x 8 words (compacted order)
    int32_t *d_midx = nullptr, *d_mdist = nullptr;
    int32_t *d_pairs = nullptr;  // per slot max_kp x 8 ints: filtered point pairs of (slot-1, slot)
    int *d_npairs = nullptr, *d_minDist = nullptr;
    int *d_status = nullptr, *d_noob = nullptr;  // d_status: STATUS_WORDS ints; [STATUS_GENERAL] for the synchronous entry points, [t] for ticket t
    int *cur_status = nullptr;                   // the word the kernels of the current call report to
    uint32_t *d_offs = nullptr, *d_spos = nullptr;  // BRIEF tests: packed offsets / positions inside a staged patch
    bool offs_set = false;
    // scratch for the explicit-point / explicit-descriptor entry points
    int32_t *d_pt_row = nullptr, *d_pt_col = nullptr;
    uint32_t *d_pt_desc = nullptr;
    uint8_t *d_pt_valid = nullptr;
    int pt_cap = 0;
    uint32_t *d_mq = nullptr, *d_mt = nullptr;
    int mq_cap = 0, mt_cap = 0;
    int32_t *d_mo_idx = nullptr, *d_mo_dist = nullptr, *d_mo_sec = nullptr;
    int mo_cap = 0;
    uint32_t *d_part_key = nullptr, *d_part_sec = nullptr;
    size_t part_cap = 0;
    // dense device staging for uploads (H2D runs as one contiguous copy, a kernel re-pitches)
    uint8_t *d_raw = nullptr;
    size_t raw_bytes = 0;
    // pinned staging
    uint8_t *h_stage = nullptr;
    size_t h_stage_bytes = 0;
    int *h_small = nullptr;  // 64 ints
    long long launches = 0;
    // copy/compute overlap of yavo_process_host_batch
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_repitched[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_done, ev_fetched;
    std::vector<char> fetched_valid;
    int fetched_C = 0;
    cudaEvent_t ev_ticket[16] = {};
    unsigned long long n_tickets = 0;
    char ticket_pending[16] = {};  // submitted, status word not yet looked at
    char raw_used[2] = {0, 0};
    int raw_C = 0;  // frames per staging half of the last submit (the halves sit at buf * raw_C * frame bytes)
    int pipeline_chunk = 0;  // frames per copy/compute stage of the host-batch path (0 = automatic)
    int matcher = 0;    // 0 = tensor-core matcher on packed 4-bit operands (K5t4), 1 = POPC matcher (K5), 2 = tensor-core matcher on FP8 operands (K5t)
    int n_sms = 148;
    int sub_batch = 0;  // frames per kernel sub-batch of yavo_frontend_batch (0 = the whole batch in one set of launches)
    // tracking step (klt_kernels.cuh): pyramid levels 1..KLT_MAX_LEVELS per slot, allocated on first use
    uint8_t *d_pyr = nullptr;
    size_t pyr_slot_stride = 0;
    size_t pyr_off[KLT_MAX_LEVELS + 1] = {};
    int pyr_pitch[KLT_MAX_LEVELS + 1] = {};
    int pyr_levels_alloc = 0;
    std::vector<int> slot_pyr;  // pyramid levels currently built for the slot (0 = none)
    float2 *d_klt_prev = nullptr, *d_klt_next = nullptr;  // explicit-point scratch
    uint8_t *d_klt_status = nullptr;
    float *d_klt_err = nullptr;
    int klt_cap = 0;
    // tracking inside the streaming path (yavo_stream_tracking): parameters and the host arrays of the next submit
    bool strk_on = false;
    KltParams strk_P = {};
    int strk_max_level = 0;
    float *strk_xy = nullptr;
    uint8_t *strk_status = nullptr;
    float *strk_err = nullptr;
    float2 *d_trk_xy = nullptr;  // batch results, per slot max_kp
    uint8_t *d_trk_status = nullptr;
    float *d_trk_err = nullptr;
    // single-frame path (yavo_frame_features): one CUDA graph per (slot, rows, cols, K) = upload + re-pitch + K1 + K3 + K4 +
    // result copies, between pinned staging buffers whose addresses the graphs hold
    struct FrameGraph {
        int slot, rows, cols, K;
        cudaGraphExec_t exec;
        int launches;  // kernels inside the graph
    };
    std::vector<FrameGraph> frame_graphs;
    // per-slot pinned copy of the pixels yavo_frame_features uploaded (dense rows): the graph's H2D source, and what
    // yavo_slot_holds compares a caller's pixels with
    std::vector<uint8_t *> h_shadow;
    std::vector<size_t> h_shadow_bytes;
    std::vector<int> shadow_rows, shadow_cols;  // 0 = the slot's shadow does not describe its current contents
    std::vector<int> stage_cols;                // width of the frame last staged in the slot's shadow (its row padding is zero for that width)
    uint8_t *h_fout = nullptr;  // pinned: 4 ints + 6 x K ints + 32 x K bytes
    uint32_t *d_fpack = nullptr;  // the same on the device (pack_frame_kernel)
    uint8_t *h_mstage = nullptr;  // pinned staging of yavo_match: descriptors in, (idx, dist, second, rev) out
    size_t h_mstage_bytes = 0;
    // optional per-kernel timing (CUDA events on the context's stream around every launch)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<int> ev_class;  // class of pair i (events 2i, 2i+1)
    std::string err;
};

enum { STATUS_TICKETS = 16, STATUS_GENERAL = 16, STATUS_WORDS = 17 };
enum { KC_REPITCH = 0, KC_DETECT, KC_COMPACT, KC_SELECT, KC_BRIEF, KC_MATCH, KC_MATCH_REDUCE, KC_FILTER, KC_PYR, KC_KLT, KC_EPI, KC_MATCH_TC, KC_SELECT_BIG, KC_COUNT };

namespace {

void prof_events(yavo_ctx *c, int cls, cudaEvent_t *e0, cudaEvent_t *e1) {
    while (c->ev_pool.size() < c->ev_used + 2) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        c->ev_pool.push_back(e);
    }
    *e0 = c->ev_pool[c->ev_used];
    *e1 = c->ev_pool[c->ev_used + 1];
    c->ev_used += 2;
    c->ev_class.push_back(cls);
}

int fail(yavo_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, YAVO_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

// PROF(cls, launch): optional event pair around one kernel launch
#define PROF(cls, ...)                                                    \
    do {                                                                  \
        cudaEvent_t e0_ = nullptr, e1_ = nullptr;                         \
        if (ctx->profiling) prof_events(ctx, cls, &e0_, &e1_);            \
        if (e0_) cudaEventRecord(e0_, ctx->ls);                           \
        __VA_ARGS__;                                                      \
        if (e1_) cudaEventRecord(e1_, ctx->ls);                           \
    } while (0)

#define CK_LAUNCH()                                                                                          \
    do {                                                                                                     \
        ctx->launches++;                                                                                     \
        cudaError_t e_ = cudaGetLastError();                                                                 \
        if (e_ != cudaSuccess)                                                                               \
            return fail(ctx, YAVO_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_),      \
                        __FILE__, __LINE__);                                                                 \
    } while (0)

template <typename T>
cudaError_t dalloc(T **p, size_t n) {
    return cudaMalloc(reinterpret_cast<void **>(p), std::max<size_t>(n, 1) * sizeof(T));
}

// CUtensorMap of the frame slots: u8 [n_slots][rows_alloc][pitch], box = one staged tile (SROW x SH bytes of one slot).
// cuTensorMapEncodeTiled is fetched from the driver at run time, so the library does not link libcuda.
int encode_slot_map(yavo_ctx *ctx, CUtensorMap *map, void *base, int box_w, int box_h) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(ctx, YAVO_ERR_CUDA, "the driver does not export cuTensorMapEncodeTiled");
    const cuuint64_t gdim[3] = {(cuuint64_t)ctx->pitch, (cuuint64_t)ctx->rows_alloc, (cuuint64_t)ctx->n_slots};
    const cuuint64_t gstride[2] = {(cuuint64_t)ctx->pitch, (cuuint64_t)ctx->frame_stride};  // bytes, dims 1 and 2
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = reinterpret_cast<encode_fn>(fn)(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, gdim, gstride, box,
                                                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, YAVO_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

// The frame slots once more as a 4-D tensor (byte in a 16-byte chunk, row, chunk of the row, slot): a box {16, SH, SROW / 16, 1}
// lands in shared memory as [chunk][row][16 bytes], the K-major core-matrix order of the tensor cores' pixel operand.
int encode_chunk_map(yavo_ctx *ctx, CUtensorMap *map, void *base) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(ctx, YAVO_ERR_CUDA, "the driver does not export cuTensorMapEncodeTiled");
    const cuuint64_t gdim[4] = {16, (cuuint64_t)ctx->rows_alloc, (cuuint64_t)(ctx->pitch / 16), (cuuint64_t)ctx->n_slots};
    const cuuint64_t gstride[3] = {(cuuint64_t)ctx->pitch, 16, (cuuint64_t)ctx->frame_stride};  // bytes, dims 1..3
    const cuuint32_t box[4] = {16u, (cuuint32_t)SH, (cuuint32_t)(SROW / 16), 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = reinterpret_cast<encode_fn>(fn)(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, base, gdim, gstride, box, estr,
                                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, YAVO_ERR_CUDA, "cuTensorMapEncodeTiled (chunk map) failed (%d)", (int)r);
    return 0;
}

template <int NT>
size_t select_smem_bytes() { return ((sizeof(SelSharedT<NT>) + 15) & ~size_t(15)) + sizeof(yavo_ent) * SEL_SMEM_ENTS; }

void drop_frame_graphs(yavo_ctx *ctx) {
    for (auto &g : ctx->frame_graphs) cudaGraphExecDestroy(g.exec);
    ctx->frame_graphs.clear();
}

int ensure_stage(yavo_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->h_stage_bytes) return 0;
    drop_frame_graphs(ctx);  // they hold the staging buffer's address
    if (ctx->h_stage) CK(cudaFreeHost(ctx->h_stage));
    ctx->h_stage = nullptr;
    ctx->h_stage_bytes = 0;
    CK(cudaMallocHost(reinterpret_cast<void **>(&ctx->h_stage), bytes));
    ctx->h_stage_bytes = bytes;
    return 0;
}

int ensure_raw(yavo_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->raw_bytes) return 0;
    drop_frame_graphs(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_raw) CK(cudaFree(ctx->d_raw));
    ctx->d_raw = nullptr;
    ctx->raw_bytes = 0;
    CK(dalloc(&ctx->d_raw, bytes));
    ctx->raw_bytes = bytes;
    return 0;
}

bool is_pinned_host(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// dense (n x rows x src_pitch) device pixels -> pitched frame slots
int launch_repitch(yavo_ctx *ctx, const uint8_t *d_src, size_t src_pitch, int slot0, int n, int rows, int cols) {
    dim3 grid((ctx->pitch / 4 + 127) / 128, rows, n);
    PROF(KC_REPITCH, repitch_kernel<<<grid, 128, 0, ctx->stream>>>(d_src, src_pitch, src_pitch * rows,
                                                  ctx->d_frames + ctx->frame_stride * slot0, ctx->pitch,
                                                  ctx->frame_stride, cols));
    CK_LAUNCH();
    for (int i = 0; i < n; i++) {
        ctx->slot_rows[slot0 + i] = rows;
        ctx->slot_cols[slot0 + i] = cols;
        ctx->slot_blur_valid[slot0 + i] = 0;
        ctx->slot_pyr[slot0 + i] = 0;
        if (!ctx->shadow_rows.empty()) ctx->shadow_rows[slot0 + i] = 0;
    }
    return 0;
}

// host pixels (row pitch `stride`) for n frames -> slots
int upload_host(yavo_ctx *ctx, int slot0, int n, const uint8_t *pixels, int rows, int cols, int stride) {
    const size_t fbytes = (size_t)rows * cols, total = fbytes * n;
    if (ctx->s_h2d) CK(cudaStreamSynchronize(ctx->s_h2d));  // the pipelined path shares d_raw
    ctx->raw_used[0] = ctx->raw_used[1] = 0;
    if (int r = ensure_raw(ctx, total)) return r;
    if (stride == cols && is_pinned_host(pixels)) {
        CK(cudaMemcpyAsync(ctx->d_raw, pixels, total, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        // pageable (or strided) source: pack into pinned staging first; the staging buffer is
        // reused, so wait for the copy that last read it
        CK(cudaStreamSynchronize(ctx->stream));
        if (int r = ensure_stage(ctx, total)) return r;
        if (stride == cols) memcpy(ctx->h_stage, pixels, total);
        else
            for (size_t r = 0; r < (size_t)rows * n; r++)
                memcpy(ctx->h_stage + r * cols, pixels + r * stride, cols);
        CK(cudaMemcpyAsync(ctx->d_raw, ctx->h_stage, total, cudaMemcpyHostToDevice, ctx->stream));
    }
    return launch_repitch(ctx, ctx->d_raw, cols, slot0, n, rows, cols);
}

int check_slot(yavo_ctx *ctx, int slot, int n = 1) {
    if (!ctx) return YAVO_ERR_INVALID;
    if (slot < 0 || n < 0 || slot + n > ctx->n_slots)
        return fail(ctx, YAVO_ERR_INVALID, "slot range [%d,%d) outside [0,%d)", slot, slot + n, ctx->n_slots);
    return 0;
}

int check_uploaded(yavo_ctx *ctx, int slot0, int n) {
    for (int s = slot0; s < slot0 + n; s++) {
        if (ctx->slot_rows[s] <= 0) return fail(ctx, YAVO_ERR_STATE, "slot %d holds no frame", s);
        if (ctx->slot_rows[s] != ctx->slot_rows[slot0] || ctx->slot_cols[s] != ctx->slot_cols[slot0])
            return fail(ctx, YAVO_ERR_INVALID, "slots %d and %d hold frames of different sizes", slot0, s);
    }
    return 0;
}

// K1 over slots [slot0, slot0+n): corners scored and left in the slots' pools + segment tables, blurred planes
int launch_detect(yavo_ctx *ctx, int slot0, int n, bool do_fast, bool do_blur) {
    const int H = ctx->slot_rows[slot0], W = ctx->slot_cols[slot0];
    const size_t fs = ctx->frame_stride;
    const uint8_t *frames = ctx->d_frames + fs * slot0;
    uint8_t *blur = ctx->d_blur + fs * slot0;
    yavo_ent *pool = ctx->d_pool + (size_t)slot0 * ctx->max_cand;
    uint32_t *seg = ctx->d_seg + (size_t)slot0 * ctx->rows_alloc * ctx->seg_cols;
    int *ncand = ctx->d_ncand + slot0;
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, n);
    if (do_fast) CK(cudaMemsetAsync(ncand, 0, sizeof(int) * (size_t)n, ctx->ls));
    if (do_fast && do_blur)
        PROF(KC_DETECT, detect_blur_kernel<true, true><<<grid, K1_THREADS, 0, ctx->ls>>>(
            ctx->frames_map, ctx->frames_cmap, slot0, frames, fs, ctx->pitch, H, W, blur, pool, ctx->max_cand, ncand, seg, ctx->seg_cols, ctx->rows_alloc, ctx->d_blur_consts));
    else if (do_fast)
        PROF(KC_DETECT, detect_blur_kernel<true, false><<<grid, K1_THREADS, 0, ctx->ls>>>(
            ctx->frames_map, ctx->frames_cmap, slot0, frames, fs, ctx->pitch, H, W, blur, pool, ctx->max_cand, ncand, seg, ctx->seg_cols, ctx->rows_alloc, ctx->d_blur_consts));
    else
        PROF(KC_DETECT, detect_blur_kernel<false, true><<<grid, K1_THREADS, 0, ctx->ls>>>(
            ctx->frames_map, ctx->frames_cmap, slot0, frames, fs, ctx->pitch, H, W, blur, pool, ctx->max_cand, ncand, seg, ctx->seg_cols, ctx->rows_alloc, ctx->d_blur_consts));
    CK_LAUNCH();
    if (do_blur)
        for (int s = slot0; s < slot0 + n; s++) ctx->slot_blur_valid[s] = 1;
    return 0;
}

// K3 over slots [slot0, slot0+n)
int launch_select(yavo_ctx *ctx, int slot0, int n, int K) {
    const int H = ctx->slot_rows[slot0], W = ctx->slot_cols[slot0];
    const size_t o = (size_t)slot0 * ctx->max_kp;
    // large lists (4K frames: 300 k candidates): the top of the partition tree runs on a thread-block cluster per frame
    const int big_min = ctx->big_min > 0 ? ctx->big_min : (ctx->big_min < 0 && (long long)H * W >= 2000000 ? 6144 : 0);
    const bool big = big_min > 0 && ctx->max_cand > big_min;
    if (big) {
        CK(cudaMemsetAsync(ctx->d_npre + slot0, 0, sizeof(int) * (size_t)n, ctx->ls));
        CK(cudaMemsetAsync(ctx->d_team_done + slot0, 0, sizeof(int) * (size_t)n, ctx->ls));
        PROF(KC_SELECT_BIG, select_big_kernel<<<n * BIG_CL, BIG_THREADS, 0, ctx->ls>>>(
            ctx->d_seg + (size_t)slot0 * ctx->rows_alloc * ctx->seg_cols, ctx->seg_cols, ctx->rows_alloc, (W + TW - 1) / TW,
            ctx->d_pool + (size_t)slot0 * ctx->max_cand, ctx->d_cand + (size_t)slot0 * ctx->max_cand, ctx->max_cand, ctx->d_ncand + slot0,
            ctx->d_scratch + (size_t)slot0 * (2 * (size_t)ctx->max_cand + 8), ctx->d_xchg + (size_t)slot0 * BIG_XCHG, K, H, big_min,
            ctx->d_pre + (size_t)slot0 * BIG_PRE, ctx->d_npre + slot0));
        CK_LAUNCH();
    }
    const int team = big ? BIG_CL : 1;
    // one frame on an otherwise idle GPU: the instance with twice the warps (phase 2 is a work queue served by the warps)
    const bool single = n == 1 && !big && !ctx->batch_select_only;
#define YAVO_SELECT_ARGS \
    ctx->d_seg + (size_t)slot0 * ctx->rows_alloc * ctx->seg_cols, ctx->seg_cols, ctx->rows_alloc, (W + TW - 1) / TW, \
    ctx->d_pool + (size_t)slot0 * ctx->max_cand, ctx->d_cand + (size_t)slot0 * ctx->max_cand, ctx->max_cand, ctx->d_ncand + slot0, \
    ctx->d_scratch + (size_t)slot0 * (2 * (size_t)ctx->max_cand + 8), K, H, W, ctx->max_kp, \
    ctx->d_kp_row + o, \
    ctx->d_kp_col + o, ctx->d_kp_score + o, ctx->d_nkp + slot0, ctx->d_bk_row + o, ctx->d_bk_col + o, \
    ctx->d_bk_score + o, ctx->d_bk_id + o, ctx->d_nbk + slot0, ctx->cur_status, \
    big ? ctx->d_pre + (size_t)slot0 * BIG_PRE : nullptr, big ? ctx->d_npre + slot0 : nullptr, team, \
    big ? ctx->d_team_done + slot0 : nullptr
    if (single)
        PROF(KC_SELECT, select_topk_kernel<SEL_THREADS_SINGLE><<<n * team, SEL_THREADS_SINGLE, select_smem_bytes<SEL_THREADS_SINGLE>(), ctx->ls>>>(YAVO_SELECT_ARGS));
    else
        PROF(KC_SELECT, select_topk_kernel<SEL_THREADS_BATCH><<<n * team, SEL_THREADS_BATCH, select_smem_bytes<SEL_THREADS_BATCH>(), ctx->ls>>>(YAVO_SELECT_ARGS));
#undef YAVO_SELECT_ARGS
    CK_LAUNCH();
    return 0;
}

int status_error(yavo_ctx *ctx, int st) {
    if (st == 2) return fail(ctx, YAVO_ERR_CUDA, "select kernel watchdog fired (work queue did not drain); results are invalid");
    if (st != 0)
        return fail(ctx, YAVO_ERR_CAPACITY,
                    "FAST candidate list overflowed max_cand=%d; create the context with a larger max_cand",
                    ctx->max_cand);
    return 0;
}

// synchronises ctx->stream; reports (and clears) what the kernels of the synchronous entry points flagged
int check_status(yavo_ctx *ctx) {
    CK(cudaMemcpyAsync(ctx->h_small, ctx->d_status + STATUS_GENERAL, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int st = ctx->h_small[0];
    if (st != 0) CK(cudaMemsetAsync(ctx->d_status + STATUS_GENERAL, 0, sizeof(int), ctx->stream));
    return status_error(ctx, st);
}

int choose_chunks(int nq, int nt, int pairs) {
    // enough CTAs to cover the 148 SMs a few times over, at least 256 train descriptors per chunk
    const int qblocks = std::max(1, (nq + MQ - 1) / MQ) * std::max(1, pairs);
    int want = (148 * 4 + qblocks - 1) / qblocks;
    int maxc = std::max(1, nt / 256);
    return std::max(1, std::min(want, maxc));
}

int ensure_partials(yavo_ctx *ctx, size_t n) {
    if (n <= ctx->part_cap) return 0;
    if (ctx->d_part_key) CK(cudaFree(ctx->d_part_key));
    if (ctx->d_part_sec) CK(cudaFree(ctx->d_part_sec));
    ctx->d_part_key = ctx->d_part_sec = nullptr;
    ctx->part_cap = 0;
    CK(dalloc(&ctx->d_part_key, n));
    CK(dalloc(&ctx->d_part_sec, n));
    ctx->part_cap = n;
    return 0;
}

// ---- tracking step helpers ------------------------------------------------------------------------------
// cv::buildOpticalFlowPyramid keeps a level while both of its sides exceed the window's
int klt_levels_for(int H, int W, int ww, int wh, int max_level) {
    int lv = 0;
    while (lv < max_level && lv < KLT_MAX_LEVELS) {
        const int nh = (H + 1) / 2, nw = (W + 1) / 2;
        if (nw <= ww || nh <= wh) break;
        H = nh;
        W = nw;
        lv++;
    }
    return lv;
}

int ensure_pyramid_alloc(yavo_ctx *ctx) {
    if (ctx->d_pyr) return 0;
    size_t off = 0;
    int r = ctx->max_rows, c = ctx->max_cols, lv = 0;
    while (lv < KLT_MAX_LEVELS && (r > 1 || c > 1)) {
        r = (r + 1) / 2;
        c = (c + 1) / 2;
        lv++;
        ctx->pyr_pitch[lv] = (c + 15) & ~15;
        ctx->pyr_off[lv] = off;
        off += (size_t)ctx->pyr_pitch[lv] * r;
    }
    ctx->pyr_levels_alloc = lv;
    ctx->pyr_slot_stride = (off + 127) & ~size_t(127);
    CK(dalloc(&ctx->d_pyr, ctx->pyr_slot_stride * ctx->n_slots));
    return 0;
}

void level_dims(int H, int W, int level, int *h, int *w) {
    for (int l = 0; l < level; l++) {
        H = (H + 1) / 2;
        W = (W + 1) / 2;
    }
    *h = H;
    *w = W;
}

// K7 over slots [slot0, slot0+n): levels 1..levels (all slots hold frames of one size)
int launch_pyramid(yavo_ctx *ctx, int slot0, int n, int levels) {
    if (int r = ensure_pyramid_alloc(ctx)) return r;
    if (levels > ctx->pyr_levels_alloc)
        return fail(ctx, YAVO_ERR_CAPACITY, "pyramid level %d exceeds the %d levels this context can hold", levels, ctx->pyr_levels_alloc);
    // contiguous runs of slots that lack some of the requested levels
    int s = slot0;
    while (s < slot0 + n) {
        if (ctx->slot_pyr[s] >= levels) { s++; continue; }
        int e = s;
        while (e < slot0 + n && ctx->slot_pyr[e] < levels) e++;
        const int H0 = ctx->slot_rows[s], W0 = ctx->slot_cols[s];
        for (int l = 1; l <= levels; l++) {
            int sh, sw, oh, ow;
            level_dims(H0, W0, l - 1, &sh, &sw);
            level_dims(H0, W0, l, &oh, &ow);
            const uint8_t *src = l == 1 ? ctx->d_frames + ctx->frame_stride * s : ctx->d_pyr + ctx->pyr_slot_stride * s + ctx->pyr_off[l - 1];
            const size_t sstride = l == 1 ? ctx->frame_stride : ctx->pyr_slot_stride;
            const int spitch = l == 1 ? ctx->pitch : ctx->pyr_pitch[l - 1];
            dim3 grid((ow + PD_TW - 1) / PD_TW, (oh + PD_TH - 1) / PD_TH, e - s);
            PROF(KC_PYR, pyr_down_kernel<<<grid, 256, 0, ctx->stream>>>(src, sstride, spitch, sh, sw,
                                                                         ctx->d_pyr + ctx->pyr_slot_stride * s + ctx->pyr_off[l],
                                                                         ctx->pyr_slot_stride, ctx->pyr_pitch[l], oh, ow));
            CK_LAUNCH();
        }
        for (int k = s; k < e; k++) ctx->slot_pyr[k] = levels;
        s = e;
    }
    return 0;
}

int klt_check_params(yavo_ctx *ctx, int ww, int wh, int max_level, int flags) {
    // OpenCV asserts winSize.width > 2 && winSize.height > 2 (lkpyramid.cpp, SparsePyrLKOpticalFlowImpl::calc)
    if (ww < 3 || wh < 3 || ww > KLT_MAX_WIN || wh > KLT_MAX_WIN)
        return fail(ctx, YAVO_ERR_INVALID, "window %dx%d outside 3..%d", ww, wh, KLT_MAX_WIN);
    if (max_level < 0) return fail(ctx, YAVO_ERR_INVALID, "max_level %d < 0", max_level);
    if (flags & ~(4 | 8)) return fail(ctx, YAVO_ERR_INVALID, "unsupported flags 0x%x (4 = USE_INITIAL_FLOW, 8 = GET_MIN_EIGENVALS)", flags);
    return 0;
}

// criteria clamping of cv::SparsePyrLKOpticalFlow::calc
KltParams klt_params(int ww, int wh, int crit_type, int max_count, double epsilon, int flags, double min_eig) {
    KltParams P;
    P.ww = ww;
    P.wh = wh;
    P.max_count = (crit_type & 1) ? std::min(std::max(max_count, 0), 100) : 30;
    const double e = (crit_type & 2) ? std::min(std::max(epsilon, 0.), 10.) : 0.01;
    P.eps2 = e * e;
    P.flags = flags;
    P.min_eig = (float)min_eig;
    return P;
}

KltLevels klt_levels_struct(yavo_ctx *ctx, int H, int W, int top) {
    KltLevels L;
    memset(&L, 0, sizeof L);
    L.top = top;
    for (int l = 0; l <= top; l++) {
        level_dims(H, W, l, &L.H[l], &L.W[l]);
        L.img[l] = l == 0 ? ctx->d_frames : ctx->d_pyr + ctx->pyr_off[l];
        L.slot_stride[l] = l == 0 ? ctx->frame_stride : ctx->pyr_slot_stride;
        L.pitch[l] = l == 0 ? ctx->pitch : ctx->pyr_pitch[l];
    }
    return L;
}

int ensure_track_buffers(yavo_ctx *ctx) {
    if (ctx->d_trk_xy) return 0;
    const size_t S = (size_t)ctx->n_slots * ctx->max_kp;
    CK(dalloc(&ctx->d_trk_xy, S));
    CK(dalloc(&ctx->d_trk_status, S));
    CK(dalloc(&ctx->d_trk_err, S));
    return 0;
}

// one launch of K8; the reference's 11 x 11 window runs the specialised instance
int launch_klt(yavo_ctx *ctx, dim3 grid, const KltLevels &L, const KltParams &P, int prev_slot0, int next_slot0,
               const float2 *prev_xy, const int32_t *kp_row, const int32_t *kp_col, const int *n_all, int n_fixed,
               int pts_stride, const float2 *init_xy, float2 *next_xy, uint8_t *status, float *err) {
    if (P.ww == 11 && P.wh == 11) {
        const size_t smem = (size_t)KltFixed<11, 11>::SMEM * KLT_WARPS;
        PROF(KC_KLT, klt_track_fixed_kernel<11, 11><<<grid, KLT_WARPS * 32, smem, ctx->stream>>>(
                         L, P, prev_slot0, next_slot0, prev_xy, kp_row, kp_col, n_all, n_fixed, pts_stride, init_xy,
                         next_xy, status, err));
        CK_LAUNCH();
        return 0;
    }
    const size_t smem = klt_smem_per_warp(P.ww, P.wh) * KLT_WARPS;
    if (smem > 48 * 1024)
        CK(cudaFuncSetAttribute(klt_track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PROF(KC_KLT, klt_track_kernel<<<grid, KLT_WARPS * 32, smem, ctx->stream>>>(
                     L, P, prev_slot0, next_slot0, prev_xy, kp_row, kp_col, n_all, n_fixed, pts_stride, init_xy,
                     next_xy, status, err));
    CK_LAUNCH();
    return 0;
}

}  // namespace

extern "C" {

const char *yavo_last_error(const yavo_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
long long yavo_kernel_launches(const yavo_ctx *ctx) { return ctx ? ctx->launches : 0; }

int yavo_create(int device, int n_slots, int max_rows, int max_cols, int max_kp, int max_cand, yavo_ctx **out) {
    yavo_ctx *ctx = nullptr;  // CK() reports through g_create_error while ctx is null
    if (!out) return fail(nullptr, YAVO_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_slots < 1 || max_rows < 1 || max_cols < 1 || max_kp < 1 || max_rows > 65535 || max_cols > 65535)
        return fail(nullptr, YAVO_ERR_INVALID, "bad sizes: n_slots=%d max_rows=%d max_cols=%d max_kp=%d", n_slots,
                    max_rows, max_cols, max_kp);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, YAVO_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path",
                    cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, YAVO_ERR_INVALID, "device %d of %d", device, ndev);
    CK(cudaSetDevice(device));
    yavo_ctx *c = new yavo_ctx();
    c->device = device;
    c->n_slots = n_slots;
    c->max_rows = max_rows;
    c->max_cols = max_cols;
    c->max_kp = max_kp;
    if (max_cand <= 0)
        max_cand = (int)std::min<long long>(SEG_MAX_OFFSET - 1, std::max(1024LL, ((long long)std::max(max_rows - 8, 1) * std::max(max_cols - 8, 1)) / 4));
    if (max_cand >= SEG_MAX_OFFSET) {
        delete c;
        return fail(nullptr, YAVO_ERR_INVALID, "max_cand %d: the segment table addresses at most %d candidates per frame", max_cand, SEG_MAX_OFFSET - 1);
    }
    c->max_cand = max_cand;
    c->pitch = ((max_cols + TW - 1) / TW) * TW;
    c->rows_alloc = max_rows;
    c->seg_cols = c->pitch / TW;
    c->frame_stride = (size_t)c->pitch * c->rows_alloc;
    c->slot_rows.assign(n_slots, 0);
    c->slot_cols.assign(n_slots, 0);
    c->slot_blur_valid.assign(n_slots, 0);
    c->slot_pyr.assign(n_slots, 0);
    ctx = c;
#define CKC(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            fail(nullptr, YAVO_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));              \
            yavo_destroy(c);                                                                           \
            return YAVO_ERR_CUDA;                                                                      \
        }                                                                                              \
    } while (0)
    CKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->ls = c->stream;
    const size_t S = n_slots;
    CKC(dalloc(&c->d_frames, S * c->frame_stride));
    CKC(dalloc(&c->d_blur, S * c->frame_stride + 16));  // the BRIEF kernel's word loads may touch the 4 bytes after a row
    CKC(dalloc(&c->d_pool, S * c->max_cand));
    CKC(dalloc(&c->d_seg, S * c->rows_alloc * c->seg_cols));
    CKC(dalloc(&c->d_cand, S * c->max_cand));
    CKC(dalloc(&c->d_ncand, S));
    CKC(dalloc(&c->d_scratch, S * (2 * (size_t)c->max_cand + 8)));  // stopper lists: a range [f, l) uses [2f, 2f + n + 4)
    CKC(dalloc(&c->d_xchg, S * BIG_XCHG));
    CKC(dalloc(&c->d_pre, S * BIG_PRE));
    CKC(dalloc(&c->d_npre, S));
    CKC(dalloc(&c->d_team_done, S));
    CKC(cudaMemset(c->d_npre, 0, S * sizeof(int)));
    CKC(dalloc(&c->d_kp_row, S * max_kp));
    CKC(dalloc(&c->d_kp_col, S * max_kp));
    CKC(dalloc(&c->d_kp_score, S * max_kp));
    CKC(dalloc(&c->d_nkp, S));
    CKC(dalloc(&c->d_bk_row, S * max_kp));
    CKC(dalloc(&c->d_bk_col, S * max_kp));
    CKC(dalloc(&c->d_bk_score, S * max_kp));
    CKC(dalloc(&c->d_bk_id, S * max_kp));
    CKC(dalloc(&c->d_nbk, S));
    CKC(dalloc(&c->d_desc, S * max_kp * 8));
    CKC(dalloc(&c->d_midx, S * max_kp));
    CKC(dalloc(&c->d_mdist, S * max_kp));
    CKC(dalloc(&c->d_pairs, S * max_kp * 8));
    CKC(dalloc(&c->d_npairs, S));
    CKC(dalloc(&c->d_minDist, S));
    CKC(cudaMemset(c->d_npairs, 0, S * sizeof(int)));
    CKC(cudaMemset(c->d_minDist, 0, S * sizeof(int)));  // row 0 of a yavo_filter_pairs range is never written
    CKC(dalloc(&c->d_status, STATUS_WORDS));
    c->cur_status = c->d_status + STATUS_GENERAL;
    CKC(dalloc(&c->d_noob, 1));
    CKC(dalloc(&c->d_offs, 256));
    CKC(dalloc(&c->d_spos, 256));
    CKC(cudaMemset(c->d_frames, 0, S * c->frame_stride));
    if (const char *e = getenv("YAVO_SELECT_SINGLE")) c->batch_select_only = atoi(e) == 0;  // A/B runs of the single-frame instance
    {
        std::vector<uint8_t> bc(bu::BU_CONST_BYTES);
        bu::bu_fill_constants(bc.data());
        CKC(dalloc(&c->d_blur_consts, bc.size()));
        CKC(cudaMemcpy(c->d_blur_consts, bc.data(), bc.size(), cudaMemcpyHostToDevice));
    }
    // eight CTAs of the detect kernel per SM need the largest shared-memory carve-out
    CKC(cudaFuncSetAttribute(detect_blur_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CKC(cudaFuncSetAttribute(detect_blur_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    memset(&c->frames_cmap, 0, sizeof c->frames_cmap);  // read only by the hybrid tensor-core blur (build variant -DYAVO_BLUR_UMMA=2)
    if (encode_slot_map(c, &c->frames_map, c->d_frames, SROW, SH) != 0 || encode_slot_map(c, &c->blur_map, c->d_blur, BP_ROWB, BP_ROWS) != 0 ||
        (YAVO_BLUR_UMMA == 2 && encode_chunk_map(c, &c->frames_cmap, c->d_frames) != 0)) {
        g_create_error = c->err;
        yavo_destroy(c);
        return YAVO_ERR_CUDA;
    }
    CKC(cudaMemset(c->d_status, 0, STATUS_WORDS * sizeof(int)));
    CKC(cudaMemset(c->d_nbk, 0, S * sizeof(int)));
    CKC(cudaMemset(c->d_nkp, 0, S * sizeof(int)));
    CKC(cudaMallocHost(reinterpret_cast<void **>(&c->h_small), 64 * sizeof(int)));
    if (BIG_CL > 8) CKC(cudaFuncSetAttribute(select_big_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CKC(cudaFuncSetAttribute(select_topk_kernel<SEL_THREADS_BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)select_smem_bytes<SEL_THREADS_BATCH>()));
    CKC(cudaFuncSetAttribute(select_topk_kernel<SEL_THREADS_SINGLE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)select_smem_bytes<SEL_THREADS_SINGLE>()));
    CKC(cudaFuncSetAttribute(tcm::match_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcm::SMEM_BYTES));
    CKC(cudaFuncSetAttribute(tcm4::match_tc4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcm4::SMEM4_BYTES));
    CKC(cudaFuncSetAttribute(tcm4::match_tc4_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcm4::SMEM4_BYTES));
    CKC(cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, device));
#undef CKC
    *out = c;
    return YAVO_OK;
}

void yavo_destroy(yavo_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    void *bufs[] = {c->d_frames, c->d_blur,    c->d_pool,    c->d_seg,      c->d_cand,     c->d_ncand,  c->d_scratch,
                    c->d_kp_row, c->d_kp_col,  c->d_kp_score, c->d_nkp,     c->d_bk_row,   c->d_bk_col, c->d_bk_score,
                    c->d_bk_id,  c->d_nbk,     c->d_desc,    c->d_midx,     c->d_mdist,    c->d_status, c->d_noob,
                    c->d_offs,   c->d_pt_row,  c->d_pt_col,  c->d_pt_desc,  c->d_pt_valid, c->d_mq,     c->d_mt,
                    c->d_mo_idx, c->d_part_key, c->d_part_sec, c->d_raw,
                    c->d_pairs,  c->d_npairs,  c->d_minDist, c->d_spos,
                    c->d_fpack,  c->d_xchg,    c->d_pre,     c->d_npre,    c->d_team_done,
                    c->d_pyr,    c->d_klt_prev, c->d_klt_next, c->d_klt_status, c->d_klt_err, c->d_trk_xy,
                    c->d_trk_status, c->d_trk_err, c->d_blur_consts};
    for (void *b : bufs)
        if (b) cudaFree(b);
    drop_frame_graphs(c);
    for (uint8_t *p : c->h_shadow)
        if (p) cudaFreeHost(p);
    if (c->h_fout) cudaFreeHost(c->h_fout);
    if (c->h_mstage) cudaFreeHost(c->h_mstage);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_small) cudaFreeHost(c->h_small);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_done) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_fetched) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_ticket)
        if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_repitched[i]) cudaEventDestroy(c->ev_repitched[i]);
    }
    for (int i = 0; i < yavo_ctx::MAX_AUX; i++) {
        if (c->aux[i]) cudaStreamDestroy(c->aux[i]);
        if (c->ev_k1[i]) cudaEventDestroy(c->ev_k1[i]);
        if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void *yavo_get_stream(yavo_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int yavo_set_profiling(yavo_ctx *ctx, int on) {
    if (!ctx) return YAVO_ERR_INVALID;
    ctx->profiling = on != 0;
    ctx->ev_used = 0;
    ctx->ev_class.clear();
    return 0;
}

int yavo_profile_collect(yavo_ctx *ctx, double *ms_per_class, int *launches_per_class, int n_classes) {
    if (!ctx || !ms_per_class || !launches_per_class) return YAVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < n_classes; k++) { ms_per_class[k] = 0.0; launches_per_class[k] = 0; }
    for (size_t i = 0; i < ctx->ev_class.size(); i++) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev_pool[2 * i], ctx->ev_pool[2 * i + 1]));
        const int k = ctx->ev_class[i];
        if (k < n_classes) { ms_per_class[k] += ms; launches_per_class[k]++; }
    }
    ctx->ev_used = 0;
    ctx->ev_class.clear();
    return 0;
}

int yavo_sync(yavo_ctx *ctx) {
    if (!ctx) return YAVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- Image --------------------------------------------------------------------------------------

int yavo_upload(yavo_ctx *ctx, int slot, const uint8_t *pixels, int rows, int cols, int stride) {
    if (int r = check_slot(ctx, slot)) return r;
    if (!pixels || rows < 1 || cols < 1 || rows > ctx->max_rows || cols > ctx->max_cols || stride < cols)
        return fail(ctx, YAVO_ERR_INVALID, "bad frame %dx%d stride %d (context max %dx%d)", rows, cols, stride,
                    ctx->max_rows, ctx->max_cols);
    CK(cudaSetDevice(ctx->device));
    return upload_host(ctx, slot, 1, pixels, rows, cols, stride);
}

int yavo_upload_batch(yavo_ctx *ctx, int slot0, int n, const uint8_t *pixels, int rows, int cols) {
    if (int r = check_slot(ctx, slot0, n)) return r;
    if (n == 0) return 0;
    if (!pixels || rows < 1 || cols < 1 || rows > ctx->max_rows || cols > ctx->max_cols)
        return fail(ctx, YAVO_ERR_INVALID, "bad frame size %dx%d", rows, cols);
    CK(cudaSetDevice(ctx->device));
    return upload_host(ctx, slot0, n, pixels, rows, cols, cols);
}

int yavo_upload_from_device(yavo_ctx *ctx, int slot0, int n, const void *d_pixels, int rows, int cols,
                            size_t pitch) {
    if (int r = check_slot(ctx, slot0, n)) return r;
    if (n == 0) return 0;
    if (!d_pixels || rows < 1 || cols < 1 || rows > ctx->max_rows || cols > ctx->max_cols || pitch < (size_t)cols)
        return fail(ctx, YAVO_ERR_INVALID, "bad device frame %dx%d pitch %zu", rows, cols, pitch);
    CK(cudaSetDevice(ctx->device));
    return launch_repitch(ctx, static_cast<const uint8_t *>(d_pixels), pitch, slot0, n, rows, cols);
}

int yavo_download(yavo_ctx *ctx, int slot, uint8_t *pixels, int rows, int cols) {
    if (int r = check_slot(ctx, slot)) return r;
    if (!pixels || rows != ctx->slot_rows[slot] || cols != ctx->slot_cols[slot])
        return fail(ctx, YAVO_ERR_INVALID, "slot %d holds %dx%d, asked for %dx%d", slot, ctx->slot_rows[slot],
                    ctx->slot_cols[slot], rows, cols);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy2DAsync(pixels, cols, ctx->d_frames + ctx->frame_stride * slot, ctx->pitch, cols, rows,
                         cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- FastDetector ---------------------------------------------------------------------------------

void yavo_ring_points(int xc, int yc, int32_t *out_xy) {
    // net result of the reference's set-based generator for radius 3 (src/FastDetector.cc:50-112)
    static const int8_t ring[16][2] = {{0, -3}, {1, -3}, {2, -2}, {3, -1}, {3, 0},  {3, 1},  {2, 2},   {1, 3},
                                       {0, 3},  {-1, 3}, {-2, 2}, {-3, 1}, {-3, 0}, {-3, -1}, {-2, -2}, {-1, -3}};
    for (int k = 0; k < 16; k++) {
        out_xy[2 * k] = xc + ring[k][0];
        out_xy[2 * k + 1] = yc + ring[k][1];
    }
}

int yavo_fast_candidates(yavo_ctx *ctx, int slot, int cap, int32_t *out_rows, int32_t *out_cols,
                         float *out_scores, int *n_cand) {
    if (int r = check_slot(ctx, slot)) return r;
    if (int r = check_uploaded(ctx, slot, 1)) return r;
    CK(cudaSetDevice(ctx->device));
    if (int r = launch_detect(ctx, slot, 1, true, true)) return r;
    PROF(KC_COMPACT, gather_kernel<<<1, K2_THREADS, 0, ctx->stream>>>(
        ctx->d_seg + (size_t)slot * ctx->rows_alloc * ctx->seg_cols, ctx->seg_cols, ctx->rows_alloc, ctx->slot_rows[slot],
        (ctx->slot_cols[slot] + TW - 1) / TW, ctx->d_pool + (size_t)slot * ctx->max_cand,
        ctx->d_cand + (size_t)slot * ctx->max_cand, ctx->max_cand, ctx->d_ncand + slot));
    CK_LAUNCH();
    CK(cudaMemcpyAsync(ctx->h_small, ctx->d_ncand + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int n = ctx->h_small[0];
    if (n_cand) *n_cand = n;
    if (n > ctx->max_cand)
        return fail(ctx, YAVO_ERR_CAPACITY, "%d FAST candidates exceed max_cand=%d", n, ctx->max_cand);
    const int m = std::min(n, cap);
    if (m > 0) {
        std::vector<yavo_ent> h(m);
        CK(cudaMemcpyAsync(h.data(), ctx->d_cand + (size_t)slot * ctx->max_cand, sizeof(yavo_ent) * m,
                           cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < m; i++) {
            const uint32_t p = (uint32_t)h[i];
            if (out_rows) out_rows[i] = (int32_t)(p >> 16);
            if (out_cols) out_cols[i] = (int32_t)(p & 0xffffu);
            if (out_scores) out_scores[i] = yavo_ent_score(h[i]);
        }
    }
    return 0;
}

int yavo_fast_detect(yavo_ctx *ctx, int slot, int max_kp, int32_t *out_rows, int32_t *out_cols,
                     float *out_scores, int *n_out, int *n_cand) {
    if (int r = check_slot(ctx, slot)) return r;
    if (int r = check_uploaded(ctx, slot, 1)) return r;
    if (max_kp <= 0) max_kp = ctx->max_kp;
    if (max_kp > ctx->max_kp)
        return fail(ctx, YAVO_ERR_CAPACITY, "max_kp %d exceeds the context's %d", max_kp, ctx->max_kp);
    CK(cudaSetDevice(ctx->device));
    if (int r = launch_detect(ctx, slot, 1, true, true)) return r;
    if (int r = launch_select(ctx, slot, 1, max_kp)) return r;
    CK(cudaMemcpyAsync(ctx->h_small + 1, ctx->d_ncand + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_small + 2, ctx->d_nkp + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (int r = check_status(ctx)) return r;  // synchronises
    const int n = ctx->h_small[2];
    if (n_cand) *n_cand = ctx->h_small[1];
    if (n_out) *n_out = n;
    const size_t o = (size_t)slot * ctx->max_kp;
    if (n > 0) {
        if (out_rows) CK(cudaMemcpyAsync(out_rows, ctx->d_kp_row + o, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_cols) CK(cudaMemcpyAsync(out_cols, ctx->d_kp_col + o, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_scores)
            CK(cudaMemcpyAsync(out_scores, ctx->d_kp_score + o, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

// One frame, one call, one synchronisation: the shape of the reference's per-frame use (LoopHandler::insertFrameFeatures,
// src/LoopHandler.cc:468-485: getFastFeatures, then computeBrief on its points).  Upload, re-pitch, K1, K3, K4 and the
// result copies are captured once per (slot, frame size, K) as a CUDA graph between pinned staging buffers; a call
// packs the pixels into the staging buffer, launches the graph and waits once.
int yavo_frame_features(yavo_ctx *ctx, int slot, const uint8_t *pixels, int rows, int cols, int stride, int max_kp,
                        int32_t *n_kp, int32_t *kp_rows, int32_t *kp_cols, float *kp_scores, int32_t *n_desc,
                        int32_t *d_rows, int32_t *d_cols, int32_t *d_ids, uint8_t *desc, int *n_cand) {
    if (int r = check_slot(ctx, slot)) return r;
    if (!pixels || rows < 1 || cols < 1 || rows > ctx->max_rows || cols > ctx->max_cols || stride < cols)
        return fail(ctx, YAVO_ERR_INVALID, "bad frame %dx%d stride %d (context max %dx%d)", rows, cols, stride, ctx->max_rows, ctx->max_cols);
    if (max_kp <= 0) max_kp = ctx->max_kp;
    if (max_kp > ctx->max_kp) return fail(ctx, YAVO_ERR_CAPACITY, "max_kp %d exceeds the context's %d", max_kp, ctx->max_kp);
    if (!ctx->offs_set) return fail(ctx, YAVO_ERR_STATE, "yavo_set_brief_offsets has not been called");
    CK(cudaSetDevice(ctx->device));
    const int K = max_kp;
    // The frame is staged in pinned memory in the slot's own layout (rows at the device pitch, padding zero) and
    // copied straight into the slot: no dense staging buffer on the device, no re-pitch kernel on this path.
    const size_t fbytes = (size_t)rows * ctx->pitch;
    if (ctx->h_shadow.empty()) {
        ctx->h_shadow.assign(ctx->n_slots, nullptr);
        ctx->h_shadow_bytes.assign(ctx->n_slots, 0);
        ctx->shadow_rows.assign(ctx->n_slots, 0);
        ctx->shadow_cols.assign(ctx->n_slots, 0);
        ctx->stage_cols.assign(ctx->n_slots, 0);
    }
    if (fbytes > ctx->h_shadow_bytes[slot]) {
        CK(cudaStreamSynchronize(ctx->stream));
        // the slot's graphs hold the old buffer's address
        for (size_t i = 0; i < ctx->frame_graphs.size();)
            if (ctx->frame_graphs[i].slot == slot) {
                cudaGraphExecDestroy(ctx->frame_graphs[i].exec);
                ctx->frame_graphs.erase(ctx->frame_graphs.begin() + i);
            } else i++;
        if (ctx->h_shadow[slot]) CK(cudaFreeHost(ctx->h_shadow[slot]));
        ctx->h_shadow[slot] = nullptr;
        ctx->h_shadow_bytes[slot] = 0;
        const size_t want = std::max(fbytes, (size_t)ctx->max_rows * ctx->pitch);
        CK(cudaMallocHost(reinterpret_cast<void **>(&ctx->h_shadow[slot]), want));
        ctx->h_shadow_bytes[slot] = want;
        ctx->stage_cols[slot] = 0;
    }
    uint8_t *stage = ctx->h_shadow[slot];
    if (ctx->stage_cols[slot] != cols) {  // a frame of another width was staged here: its pixels must not remain in the row padding
        memset(stage, 0, ctx->h_shadow_bytes[slot]);
        ctx->stage_cols[slot] = cols;
    }
    const size_t out_ints = 4 + 6 * (size_t)ctx->max_kp, out_bytes = out_ints * 4 + 32 * (size_t)ctx->max_kp;
    if (!ctx->h_fout) {
        CK(cudaMallocHost(reinterpret_cast<void **>(&ctx->h_fout), out_bytes));
        CK(dalloc(&ctx->d_fpack, out_bytes / 4));
    }
    int32_t *hi = reinterpret_cast<int32_t *>(ctx->h_fout);
    int32_t *h_kr = hi + 4, *h_kc = h_kr + ctx->max_kp, *h_br = h_kc + ctx->max_kp, *h_bc = h_br + ctx->max_kp, *h_bi = h_bc + ctx->max_kp;
    float *h_ks = reinterpret_cast<float *>(h_bi + ctx->max_kp);
    uint8_t *h_desc = ctx->h_fout + out_ints * 4;
    // the previous call's graph has been waited for: the staging buffers are free
    for (int r = 0; r < rows; r++) memcpy(stage + (size_t)r * ctx->pitch, pixels + (size_t)r * stride, cols);
    // host-side bookkeeping of an upload (what launch_repitch does on the batch paths)
    ctx->slot_rows[slot] = rows;
    ctx->slot_cols[slot] = cols;
    ctx->slot_blur_valid[slot] = 0;
    ctx->slot_pyr[slot] = 0;
    cudaGraphExec_t exec = nullptr;
    int graph_launches = 0;
    for (auto &g : ctx->frame_graphs)
        if (g.slot == slot && g.rows == rows && g.cols == cols && g.K == K) {
            exec = g.exec;
            graph_launches = g.launches;
        }
    const long long launches_before = ctx->launches;
    if (!exec || ctx->profiling) {
        const bool capture = !ctx->profiling;  // per-kernel profiling wants plain launches with event pairs
        cudaGraph_t graph = nullptr;
        if (capture) CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        do {
            const size_t o = (size_t)slot * ctx->max_kp;
            if (cudaMemcpyAsync(ctx->d_frames + ctx->frame_stride * slot, stage, fbytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { rc = YAVO_ERR_CUDA; break; }
            if ((rc = launch_detect(ctx, slot, 1, true, true))) break;
            if ((rc = launch_select(ctx, slot, 1, K))) break;
            const int kp_per_block = (K4_THREADS / 32) * BP_KPW;
            dim3 grid((K + kp_per_block - 1) / kp_per_block, 1);
            PROF(KC_BRIEF, brief_kernel<<<grid, K4_THREADS, 0, ctx->stream>>>(
                               ctx->blur_map, slot, ctx->d_blur + ctx->frame_stride * slot, ctx->frame_stride, ctx->pitch, rows, cols,
                               ctx->d_offs, ctx->d_spos, ctx->d_bk_row + o, ctx->d_bk_col + o, ctx->d_nbk + slot, 0, ctx->max_kp,
                               ctx->d_desc + o * 8, nullptr, nullptr));
            ctx->launches++;
            const int words = 8 * K;
            pack_frame_kernel<<<(words + 255) / 256, 256, 0, ctx->stream>>>(
                ctx->d_nkp + slot, ctx->d_nbk + slot, ctx->d_ncand + slot, ctx->d_status + STATUS_GENERAL, ctx->d_kp_row + o,
                ctx->d_kp_col + o, ctx->d_kp_score + o, ctx->d_bk_row + o, ctx->d_bk_col + o, ctx->d_bk_id + o, ctx->d_desc + o * 8,
                ctx->max_kp, ctx->d_fpack);
            ctx->launches++;
            if (cudaMemcpyAsync(ctx->h_fout, ctx->d_fpack, out_bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = YAVO_ERR_CUDA;
        } while (0);
        if (capture) {
            const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
            if (rc || e != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                if (!rc) return fail(ctx, YAVO_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
                return rc;
            }
            const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ei != cudaSuccess) return fail(ctx, YAVO_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ei));
            ctx->frame_graphs.push_back({slot, rows, cols, K, exec, (int)(ctx->launches - launches_before)});
        } else if (rc) {
            return rc;
        }
    } else {
        ctx->launches += graph_launches;  // the launch counter the captured helpers advanced when the graph was recorded
    }
    if (exec && !ctx->profiling) CK(cudaGraphLaunch(exec, ctx->stream));
    ctx->slot_blur_valid[slot] = 1;
    ctx->shadow_rows[slot] = rows;
    ctx->shadow_cols[slot] = cols;
    CK(cudaStreamSynchronize(ctx->stream));
    if (hi[3] != 0) {
        CK(cudaMemsetAsync(ctx->d_status + STATUS_GENERAL, 0, sizeof(int), ctx->stream));
        return status_error(ctx, hi[3]);
    }
    const int n = hi[0], nb = hi[1];
    if (n_kp) *n_kp = n;
    if (n_desc) *n_desc = nb;
    if (n_cand) *n_cand = hi[2];
    if (kp_rows) memcpy(kp_rows, h_kr, 4 * (size_t)n);
    if (kp_cols) memcpy(kp_cols, h_kc, 4 * (size_t)n);
    if (kp_scores) memcpy(kp_scores, h_ks, 4 * (size_t)n);
    if (d_rows) memcpy(d_rows, h_br, 4 * (size_t)nb);
    if (d_cols) memcpy(d_cols, h_bc, 4 * (size_t)nb);
    if (d_ids) memcpy(d_ids, h_bi, 4 * (size_t)nb);
    if (desc) memcpy(desc, h_desc, 32 * (size_t)nb);
    return 0;
}

int yavo_slot_holds(yavo_ctx *ctx, int slot, const uint8_t *pixels, int rows, int cols, int stride) {
    if (!ctx || slot < 0 || slot >= ctx->n_slots || !pixels || stride < cols) return YAVO_ERR_INVALID;
    if (ctx->shadow_rows.empty() || ctx->shadow_rows[slot] != rows || ctx->shadow_cols[slot] != cols || rows < 1) return 0;
    const uint8_t *sh = ctx->h_shadow[slot];  // rows at the device pitch
    for (int r = 0; r < rows; r++)
        if (memcmp(sh + (size_t)r * ctx->pitch, pixels + (size_t)r * stride, cols) != 0) return 0;
    return 1;
}

// ---- Brief ----------------------------------------------------------------------------------------

int yavo_set_brief_offsets(yavo_ctx *ctx, const int32_t *offsets) {
    if (!ctx || !offsets) return YAVO_ERR_INVALID;
    uint32_t packed[256], spos[256];
    for (int j = 0; j < 256; j++) {
        uint32_t w = 0;
        for (int k = 0; k < 4; k++) {
            const int v = offsets[4 * j + k];
            if (v < -8 || v > 8) return fail(ctx, YAVO_ERR_INVALID, "offset %d of test %d outside [-8,8]", v, j);
            w |= (uint32_t)(uint8_t)(int8_t)v << (8 * k);
        }
        packed[j] = w;
        // the same test as byte positions inside the 17 x 32-byte patch the kernel stages per keypoint
        const int p1 = (offsets[4 * j + 0] + 8) * BP_ROWB + offsets[4 * j + 1] + 8;
        const int p2 = (offsets[4 * j + 2] + 8) * BP_ROWB + offsets[4 * j + 3] + 8;
        spos[j] = (uint32_t)p1 | ((uint32_t)p2 << 16);
    }
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->d_offs, packed, sizeof packed, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_spos, spos, sizeof spos, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->offs_set = true;
    return 0;
}

static int ensure_blur(yavo_ctx *ctx, int slot) {
    if (ctx->slot_blur_valid[slot]) return 0;
    return launch_detect(ctx, slot, 1, false, true);
}

int yavo_blurred(yavo_ctx *ctx, int slot, uint8_t *out, int rows, int cols) {
    if (int r = check_slot(ctx, slot)) return r;
    if (int r = check_uploaded(ctx, slot, 1)) return r;
    if (!out || rows != ctx->slot_rows[slot] || cols != ctx->slot_cols[slot])
        return fail(ctx, YAVO_ERR_INVALID, "slot %d holds %dx%d", slot, ctx->slot_rows[slot], ctx->slot_cols[slot]);
    CK(cudaSetDevice(ctx->device));
    if (int r = ensure_blur(ctx, slot)) return r;
    CK(cudaMemcpy2DAsync(out, cols, ctx->d_blur + ctx->frame_stride * slot, ctx->pitch, cols, rows,
                         cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int yavo_brief_describe(yavo_ctx *ctx, int slot, const int32_t *rows, const int32_t *cols, int n,
                        uint8_t *out_desc, uint8_t *out_valid, int *n_oob) {
    if (int r = check_slot(ctx, slot)) return r;
    if (int r = check_uploaded(ctx, slot, 1)) return r;
    if (n < 0 || (n > 0 && (!rows || !cols || !out_desc))) return fail(ctx, YAVO_ERR_INVALID, "bad point list");
    if (!ctx->offs_set) return fail(ctx, YAVO_ERR_STATE, "yavo_set_brief_offsets has not been called");
    if (n_oob) *n_oob = 0;
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    if (n > ctx->pt_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_pt_row) CK(cudaFree(ctx->d_pt_row));
        if (ctx->d_pt_col) CK(cudaFree(ctx->d_pt_col));
        if (ctx->d_pt_desc) CK(cudaFree(ctx->d_pt_desc));
        if (ctx->d_pt_valid) CK(cudaFree(ctx->d_pt_valid));
        ctx->d_pt_row = ctx->d_pt_col = nullptr;
        ctx->d_pt_desc = nullptr;
        ctx->d_pt_valid = nullptr;
        ctx->pt_cap = 0;
        const int cap = std::max(n, 4096);
        CK(dalloc(&ctx->d_pt_row, cap));
        CK(dalloc(&ctx->d_pt_col, cap));
        CK(dalloc(&ctx->d_pt_desc, (size_t)cap * 8));
        CK(dalloc(&ctx->d_pt_valid, cap));
        ctx->pt_cap = cap;
    }
    if (int r = ensure_blur(ctx, slot)) return r;
    const int H = ctx->slot_rows[slot], W = ctx->slot_cols[slot];
    CK(cudaMemcpyAsync(ctx->d_pt_row, rows, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_pt_col, cols, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_noob, 0, sizeof(int), ctx->stream));
    const int kp_per_block = (K4_THREADS / 32) * BP_KPW;
    dim3 grid((n + kp_per_block - 1) / kp_per_block, 1);
    PROF(KC_BRIEF, brief_kernel<<<grid, K4_THREADS, 0, ctx->stream>>>(ctx->blur_map, slot, ctx->d_blur + ctx->frame_stride * slot, ctx->frame_stride,
                                                       ctx->pitch, H, W, ctx->d_offs, ctx->d_spos, ctx->d_pt_row, ctx->d_pt_col,
                                                       nullptr, n, n, ctx->d_pt_desc, ctx->d_pt_valid, ctx->d_noob));
    CK_LAUNCH();
    CK(cudaMemcpyAsync(out_desc, ctx->d_pt_desc, 32 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_valid) CK(cudaMemcpyAsync(out_valid, ctx->d_pt_valid, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_small, ctx->d_noob, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n_oob) *n_oob = ctx->h_small[0];
    return 0;
}

// K5t launch: pairs x q_tiles work items over one persistent CTA per SM
static int launch_match_tc(yavo_ctx *ctx, const uint32_t *dq_all, const int *nq_all, int nq_fixed, const uint32_t *dt_all,
                           const int *nt_all, int nt_fixed, size_t set_stride_words, int q_off, int t_off, int pairs,
                           int max_q, int out_stride, int32_t *o_idx, int32_t *o_dist, int32_t *o_sec = nullptr) {
    const int q_tiles = (max_q + tcm::TQ - 1) / tcm::TQ;
    const long long items = (long long)pairs * q_tiles;
    if (items <= 0) return 0;
    const int grid = (int)std::min<long long>(items, ctx->n_sms);
    if (o_sec)  // ratio-test extension: the packed-4-bit kernel also keeps the second smallest distance of every query
        PROF(KC_MATCH_TC, tcm4::match_tc4_kernel<false, true><<<grid, tcm4::THREADS4, tcm4::SMEM4_BYTES, ctx->stream>>>(
                              dq_all, nq_all, nq_fixed, dt_all, nt_all, nt_fixed, set_stride_words, q_off, t_off, pairs, q_tiles,
                              out_stride, o_idx, o_dist, nullptr, o_sec));
    else if (ctx->matcher == 2)
        PROF(KC_MATCH_TC, tcm::match_tc_kernel<false><<<grid, tcm::THREADS, tcm::SMEM_BYTES, ctx->stream>>>(
                              dq_all, nq_all, nq_fixed, dt_all, nt_all, nt_fixed, set_stride_words, q_off, t_off, pairs, q_tiles,
                              out_stride, o_idx, o_dist, nullptr));
    else
        PROF(KC_MATCH_TC, tcm4::match_tc4_kernel<false><<<grid, tcm4::THREADS4, tcm4::SMEM4_BYTES, ctx->stream>>>(
                              dq_all, nq_all, nq_fixed, dt_all, nt_all, nt_fixed, set_stride_words, q_off, t_off, pairs, q_tiles,
                              out_stride, o_idx, o_dist, nullptr));
    CK_LAUNCH();
    return 0;
}

static int match_device(yavo_ctx *ctx, const uint32_t *dq, int n1, const uint32_t *dt, int n2, int32_t *o_idx,
                        int32_t *o_dist, int32_t *o_sec) {
    // tensor-core matcher (its SECOND variant when the caller wants the second-best distance, the ratio-test extension);
    // yavo_set_matcher(ctx, 1) selects the integer-pipe kernel
    if (ctx->matcher != 1) return launch_match_tc(ctx, dq, nullptr, n1, dt, nullptr, n2, 0, 0, 0, 1, n1, n1, o_idx, o_dist, o_sec);
    const int chunks = choose_chunks(n1, n2, 1);
    const int chunk = std::max(1, (std::max(n2, 1) + chunks - 1) / chunks);
    if (int r = ensure_partials(ctx, (size_t)n1 * chunks)) return r;
    dim3 grid((n1 + MQ - 1) / MQ, chunks, 1);
    if (o_sec)  // second-best tracking only when the caller asked for it (ratio-test extension)
        PROF(KC_MATCH, match_partial_kernel<true><<<grid, MQ, 0, ctx->stream>>>(dq, nullptr, n1, dt, nullptr, n2, 0, 0, 0, chunk,
                                                                                 chunks, n1, ctx->d_part_key, ctx->d_part_sec, 1u << 22));
    else
        PROF(KC_MATCH, match_partial_kernel<false><<<grid, MQ, 0, ctx->stream>>>(dq, nullptr, n1, dt, nullptr, n2, 0, 0, 0, chunk,
                                                                                  chunks, n1, ctx->d_part_key, ctx->d_part_sec, 1u << 22));
    CK_LAUNCH();
    PROF(KC_MATCH_REDUCE, match_reduce_kernel<<<dim3((n1 + 127) / 128, 1), 128, 0, ctx->stream>>>(
        ctx->d_part_key, ctx->d_part_sec, nullptr, n1, 0, nullptr, n2, 0, chunk, chunks, n1, o_idx, o_dist, o_sec));
    CK_LAUNCH();
    return 0;
}

int yavo_match(yavo_ctx *ctx, const uint8_t *d1, int n1, const uint8_t *d2, int n2, int32_t *out_idx,
               int32_t *out_dist, int32_t *out_second, int32_t *out_rev_idx) {
    if (!ctx) return YAVO_ERR_INVALID;
    if (n1 < 0 || n2 < 0 || (n1 > 0 && (!d1 || !out_idx || !out_dist)) || (n2 > 0 && !d2))
        return fail(ctx, YAVO_ERR_INVALID, "bad descriptor sets");
    if (n1 >= (1 << 22) || n2 >= (1 << 22)) return fail(ctx, YAVO_ERR_CAPACITY, "descriptor sets are limited to 2^22-1");
    CK(cudaSetDevice(ctx->device));
    // device buffers and the pinned staging block are kept between calls and only ever grow (the reference matches one
    // pair of ~2000-keypoint frames per call, src/LoopHandler.cc:189,534): a call is two H2D copies, the kernel(s), one
    // D2H copy and ONE synchronisation
    if (n1 > ctx->mq_cap || n2 > ctx->mt_cap || std::max(n1, n2) > ctx->mo_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (n1 > ctx->mq_cap) {
            if (ctx->d_mq) CK(cudaFree(ctx->d_mq));
            ctx->d_mq = nullptr; ctx->mq_cap = 0;
            const int cap = std::max(n1, 4096);
            CK(dalloc(&ctx->d_mq, (size_t)cap * 8));
            ctx->mq_cap = cap;
        }
        if (n2 > ctx->mt_cap) {
            if (ctx->d_mt) CK(cudaFree(ctx->d_mt));
            ctx->d_mt = nullptr; ctx->mt_cap = 0;
            const int cap = std::max(n2, 4096);
            CK(dalloc(&ctx->d_mt, (size_t)cap * 8));
            ctx->mt_cap = cap;
        }
        const int mo = std::max(n1, n2);
        if (mo > ctx->mo_cap) {
            if (ctx->d_mo_idx) CK(cudaFree(ctx->d_mo_idx));
            ctx->d_mo_idx = ctx->d_mo_dist = ctx->d_mo_sec = nullptr; ctx->mo_cap = 0;
            const int cap = std::max(mo, 4096);
            CK(dalloc(&ctx->d_mo_idx, (size_t)cap * 5));  // idx | dist | second | reverse idx | reverse dist, contiguous
            ctx->d_mo_dist = ctx->d_mo_idx + cap;
            ctx->d_mo_sec = ctx->d_mo_idx + 2 * (size_t)cap;
            ctx->mo_cap = cap;
        }
    }
    const size_t cap = ctx->mo_cap;
    int32_t *d_rev = ctx->d_mo_idx + 3 * cap, *d_rev_dist = ctx->d_mo_idx + 4 * cap;
    const size_t in_bytes = 32 * ((size_t)n1 + n2), out_bytes = 4 * 4 * cap, need = in_bytes + out_bytes;
    if (need > ctx->h_mstage_bytes) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->h_mstage) CK(cudaFreeHost(ctx->h_mstage));
        ctx->h_mstage = nullptr;
        ctx->h_mstage_bytes = 0;
        const size_t want = std::max(need, (size_t)(32 * 2 * 4096 + 4 * 4 * 4096));
        CK(cudaMallocHost(reinterpret_cast<void **>(&ctx->h_mstage), want));
        ctx->h_mstage_bytes = want;
    }
    uint8_t *h_in = ctx->h_mstage;
    int32_t *h_out = reinterpret_cast<int32_t *>(ctx->h_mstage + ctx->h_mstage_bytes - out_bytes);  // the tail of the block
    if (n1 > 0) memcpy(h_in, d1, 32 * (size_t)n1);
    if (n2 > 0) memcpy(h_in + 32 * (size_t)n1, d2, 32 * (size_t)n2);
    if (n1 > 0) CK(cudaMemcpyAsync(ctx->d_mq, h_in, 32 * (size_t)n1, cudaMemcpyHostToDevice, ctx->stream));
    if (n2 > 0) CK(cudaMemcpyAsync(ctx->d_mt, h_in + 32 * (size_t)n1, 32 * (size_t)n2, cudaMemcpyHostToDevice, ctx->stream));
    if (n1 > 0)
        if (int r = match_device(ctx, ctx->d_mq, n1, ctx->d_mt, n2, ctx->d_mo_idx, ctx->d_mo_dist,
                                 out_second ? ctx->d_mo_sec : nullptr))
            return r;
    const bool rev = out_rev_idx && n2 > 0;
    if (rev) {
        // cross-check extension: the same kernel with the roles swapped (a second pass over the N1 x N2 distances on the
        // tensor pipe; the operands are already on the device)
        if (int r = match_device(ctx, ctx->d_mt, n2, ctx->d_mq, n1, d_rev, d_rev_dist, nullptr)) return r;
    }
    // one copy back: [idx | dist | second | reverse idx] x cap
    if (n1 > 0 || rev) CK(cudaMemcpyAsync(h_out, ctx->d_mo_idx, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n1 > 0) {
        memcpy(out_idx, h_out, 4 * (size_t)n1);
        memcpy(out_dist, h_out + cap, 4 * (size_t)n1);
        if (out_second) memcpy(out_second, h_out + 2 * cap, 4 * (size_t)n1);
    }
    if (rev) memcpy(out_rev_idx, h_out + 3 * cap, 4 * (size_t)n2);
    return 0;
}

int yavo_remove_outliers(const int32_t *dist, int n, int threshold, uint8_t *keep) {
    if (n <= 0 || !dist || !keep) return 0;  // the reference dereferences end() on an empty list (UB)
    int mn = INT_MAX;
    for (int i = 0; i < n; i++) mn = std::min(mn, dist[i]);
    // the reference's int arithmetic: 2 * INT_MAX (empty train set) wraps to -2, so nothing is kept — the same
    // answer filter_pairs_kernel gives
    const int lim = std::max((int)(2u * (unsigned)mn), threshold);
    int kept = 0;
    for (int i = 0; i < n; i++) {
        keep[i] = dist[i] < lim ? 1 : 0;
        kept += keep[i];
    }
    return kept;
}

// ---- batch front end -------------------------------------------------------------------------------

// detect -> select -> describe on slots [slot0, slot0+n); match pairs (p, p+1) that END inside the range.
// link_prev also matches (slot0-1, slot0), whose query descriptors an earlier call left in place.
// K1 -> K3 -> K4 on slots [slot0, slot0+n), on the current launch stream
static int features_range(yavo_ctx *ctx, int slot0, int n, cudaEvent_t after_detect = nullptr) {
    const int H = ctx->slot_rows[slot0], W = ctx->slot_cols[slot0];
    if (int r = launch_detect(ctx, slot0, n, true, true)) return r;
    if (after_detect) CK(cudaEventRecord(after_detect, ctx->ls));
    if (int r = launch_select(ctx, slot0, n, ctx->max_kp)) return r;
    const size_t o = (size_t)slot0 * ctx->max_kp;
    const int kp_per_block = (K4_THREADS / 32) * BP_KPW;
    dim3 grid((ctx->max_kp + kp_per_block - 1) / kp_per_block, n);
    PROF(KC_BRIEF, brief_kernel<<<grid, K4_THREADS, 0, ctx->ls>>>(
        ctx->blur_map, slot0, ctx->d_blur + ctx->frame_stride * slot0, ctx->frame_stride, ctx->pitch, H, W, ctx->d_offs, ctx->d_spos,
        ctx->d_bk_row + o, ctx->d_bk_col + o, ctx->d_nbk + slot0, 0, ctx->max_kp, ctx->d_desc + o * 8, nullptr,
        nullptr));
    CK_LAUNCH();
    return 0;
}

static int ensure_aux(yavo_ctx *ctx) {
    if (ctx->ev_fork) return 0;
    for (int i = 0; i < yavo_ctx::MAX_AUX; i++) {
        CK(cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_k1[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    return 0;
}

static int frontend_range(yavo_ctx *ctx, int slot0, int n, bool do_match, bool link_prev) {
    const int C = ctx->ov_chunk, NS = std::min<int>(std::max(ctx->ov_streams, 1), yavo_ctx::MAX_AUX);
    if (C > 0 && NS > 1 && !ctx->profiling && n >= 2 * C) {
        // Overlapped: chunk i runs K1 -> K3 -> K4 on auxiliary stream i % NS; its detect kernel waits for the detect kernel of
        // chunk i-1, which staggers the chunks so that an issue-bound detect kernel always has latency-bound select / BRIEF
        // kernels of earlier chunks beside it.  Results are identical (the chunks touch disjoint slots); per-kernel
        // profiling (yavo_set_profiling) runs the serial path so that its event pairs time one kernel each.
        if (int r = ensure_aux(ctx)) return r;
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        const int nchunks = (n + C - 1) / C;
        for (int i = 0; i < nchunks; i++) {
            const int s0 = slot0 + i * C, nc = std::min(C, slot0 + n - s0), si = i % NS;
            if (i < NS) CK(cudaStreamWaitEvent(ctx->aux[si], ctx->ev_fork, 0));
            if (i > 0) CK(cudaStreamWaitEvent(ctx->aux[si], ctx->ev_k1[(i - 1) % NS], 0));
            ctx->ls = ctx->aux[si];
            const int r = features_range(ctx, s0, nc, ctx->ev_k1[si]);
            ctx->ls = ctx->stream;
            if (r) return r;
        }
        for (int j = 0; j < std::min(NS, nchunks); j++) {
            CK(cudaEventRecord(ctx->ev_join[j], ctx->aux[j]));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[j], 0));
        }
    } else if (int r = features_range(ctx, slot0, n)) {
        return r;
    }
    const int m0 = (link_prev && slot0 > 0) ? slot0 - 1 : slot0;  // first query slot
    const int pairs = slot0 + n - 1 - m0;
    if (do_match && pairs > 0 && ctx->matcher != 1) {
        // pair p: queries = slot m0+p, train = slot m0+p+1; results stored at the train slot
        const size_t mo = (size_t)m0 * ctx->max_kp;
        if (int r = launch_match_tc(ctx, ctx->d_desc + mo * 8, ctx->d_nbk + m0, 0, ctx->d_desc + mo * 8, ctx->d_nbk + m0, 0,
                                    (size_t)ctx->max_kp * 8, 0, 1, pairs, ctx->max_kp, ctx->max_kp,
                                    ctx->d_midx + mo + ctx->max_kp, ctx->d_mdist + mo + ctx->max_kp))
            return r;
    } else if (do_match && pairs > 0) {
        const size_t mo = (size_t)m0 * ctx->max_kp;
        const int chunks = choose_chunks(ctx->max_kp, ctx->max_kp, pairs);
        const int chunk = (ctx->max_kp + chunks - 1) / chunks;
        if (int r = ensure_partials(ctx, (size_t)pairs * ctx->max_kp * chunks)) return r;
        // pair p: queries = slot m0+p, train = slot m0+p+1; results stored at the train slot
        dim3 grid((ctx->max_kp + MQ - 1) / MQ, chunks, pairs);
        PROF(KC_MATCH, match_partial_kernel<false><<<grid, MQ, 0, ctx->stream>>>(
            ctx->d_desc + mo * 8, ctx->d_nbk + m0, 0, ctx->d_desc + mo * 8, ctx->d_nbk + m0, 0,
            (size_t)ctx->max_kp * 8, 0, 1, chunk, chunks, ctx->max_kp, ctx->d_part_key, ctx->d_part_sec, 1u << 22));
        CK_LAUNCH();
        PROF(KC_MATCH_REDUCE, match_reduce_kernel<<<dim3((ctx->max_kp + 127) / 128, pairs), 128, 0, ctx->stream>>>(
            ctx->d_part_key, ctx->d_part_sec, ctx->d_nbk + m0, 0, 0, ctx->d_nbk + m0, 0, 1, chunk, chunks,
            ctx->max_kp, ctx->d_midx + mo + ctx->max_kp, ctx->d_mdist + mo + ctx->max_kp, nullptr));
        CK_LAUNCH();
    }
    return 0;
}

int yavo_frontend_batch(yavo_ctx *ctx, int slot0, int n, int do_match) {
    if (int r = check_slot(ctx, slot0, n)) return r;
    if (n == 0) return 0;
    if (int r = check_uploaded(ctx, slot0, n)) return r;
    if (!ctx->offs_set) return fail(ctx, YAVO_ERR_STATE, "yavo_set_brief_offsets has not been called");
    CK(cudaSetDevice(ctx->device));
    // One set of launches for the whole batch by default: measured on B200 (profiles/), L2-sized sub-batches
    // lose more to the latency-bound select kernel (one CTA per frame, wants hundreds in flight) than the
    // scoring / BRIEF kernels gain from L2-resident planes.  sub_batch > 0 overrides.
    const int sub = ctx->sub_batch > 0 ? std::min(ctx->sub_batch, n) : n;
    for (int s0 = 0; s0 < n; s0 += sub)
        if (int r = frontend_range(ctx, slot0 + s0, std::min(sub, n - s0), do_match != 0, s0 > 0)) return r;
    return 0;
}

int yavo_fetch_batch(yavo_ctx *ctx, int slot0, int n, int32_t *n_kp, int32_t *rows, int32_t *cols, float *scores,
                     uint8_t *desc, int32_t *match_idx, int32_t *match_dist) {
    if (int r = check_slot(ctx, slot0, n)) return r;
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    const size_t o = (size_t)slot0 * ctx->max_kp, cnt = (size_t)n * ctx->max_kp;
    if (n_kp) CK(cudaMemcpyAsync(n_kp, ctx->d_nbk + slot0, 4 * (size_t)n, cudaMemcpyDefault, ctx->stream));
    if (rows) CK(cudaMemcpyAsync(rows, ctx->d_bk_row + o, 4 * cnt, cudaMemcpyDefault, ctx->stream));
    if (cols) CK(cudaMemcpyAsync(cols, ctx->d_bk_col + o, 4 * cnt, cudaMemcpyDefault, ctx->stream));
    if (scores) CK(cudaMemcpyAsync(scores, ctx->d_bk_score + o, 4 * cnt, cudaMemcpyDefault, ctx->stream));
    if (desc) CK(cudaMemcpyAsync(desc, ctx->d_desc + o * 8, 32 * cnt, cudaMemcpyDefault, ctx->stream));
    if (match_idx) CK(cudaMemcpyAsync(match_idx, ctx->d_midx + o, 4 * cnt, cudaMemcpyDefault, ctx->stream));
    if (match_dist) CK(cudaMemcpyAsync(match_dist, ctx->d_mdist + o, 4 * cnt, cudaMemcpyDefault, ctx->stream));
    return check_status(ctx);  // synchronises; reports a candidate-list overflow
}

// D2H of the results of slots [slot0, slot0+n) on `st`, into host arrays whose row 0 is slot `base`
static int fetch_async(yavo_ctx *ctx, cudaStream_t st, int base, int slot0, int n, int32_t *n_kp, int32_t *rows,
                       int32_t *cols, float *scores, uint8_t *desc, int32_t *match_idx, int32_t *match_dist) {
    const size_t o = (size_t)slot0 * ctx->max_kp, cnt = (size_t)n * ctx->max_kp, h = (size_t)(slot0 - base) * ctx->max_kp;
    if (n_kp) CK(cudaMemcpyAsync(n_kp + (slot0 - base), ctx->d_nbk + slot0, 4 * (size_t)n, cudaMemcpyDefault, st));
    if (rows) CK(cudaMemcpyAsync(rows + h, ctx->d_bk_row + o, 4 * cnt, cudaMemcpyDefault, st));
    if (cols) CK(cudaMemcpyAsync(cols + h, ctx->d_bk_col + o, 4 * cnt, cudaMemcpyDefault, st));
    if (scores) CK(cudaMemcpyAsync(scores + h, ctx->d_bk_score + o, 4 * cnt, cudaMemcpyDefault, st));
    if (desc) CK(cudaMemcpyAsync(desc + h * 32, ctx->d_desc + o * 8, 32 * cnt, cudaMemcpyDefault, st));
    if (match_idx) CK(cudaMemcpyAsync(match_idx + h, ctx->d_midx + o, 4 * cnt, cudaMemcpyDefault, st));
    if (match_dist) CK(cudaMemcpyAsync(match_dist + h, ctx->d_mdist + o, 4 * cnt, cudaMemcpyDefault, st));
    return 0;
}

int yavo_submit_host_batch(yavo_ctx *ctx, const uint8_t *pixels, int n, int rows, int cols, int do_match,
                           int32_t *n_kp, int32_t *out_rows, int32_t *out_cols, float *scores, uint8_t *desc,
                           int32_t *match_idx, int32_t *match_dist) {
    if (int r = check_slot(ctx, 0, n)) return r;
    if (n == 0) return 0;
    if (!pixels || rows < 1 || cols < 1 || rows > ctx->max_rows || cols > ctx->max_cols)
        return fail(ctx, YAVO_ERR_INVALID, "bad frame size %dx%d", rows, cols);
    if (!ctx->offs_set) return fail(ctx, YAVO_ERR_STATE, "yavo_set_brief_offsets has not been called");
    if (!is_pinned_host(pixels))
        return fail(ctx, YAVO_ERR_INVALID, "yavo_submit_host_batch needs pinned host frames (cudaHostAlloc/cudaHostRegister)");
    CK(cudaSetDevice(ctx->device));
    // three streams: chunk c's pixels cross PCIe (s_h2d) while chunk c-1 is in the kernels (ctx->stream) and
    // chunk c-2's keypoints / descriptors / matches go back (s_d2h); consecutive submits overlap the same way.
    if (!ctx->s_h2d) {
        CK(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CK(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_repitched[i], cudaEventDisableTiming));
        }
    }
    // automatic stage size: a quarter of the batch, between 16 and 128 frames (measured on B200: 128-frame
    // stages keep the latency-bound select kernel efficient; smaller batches still get 4 stages of overlap)
    const int want = ctx->pipeline_chunk > 0 ? ctx->pipeline_chunk : std::min(128, std::max(16, (n + 3) / 4));
    const int C = std::max(1, std::min(want, n));
    const int nchunks = (n + C - 1) / C;
    while ((int)ctx->ev_done.size() < nchunks) {
        cudaEvent_t e, g;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&g, cudaEventDisableTiming));
        ctx->ev_done.push_back(e);
        ctx->ev_fetched.push_back(g);
        ctx->fetched_valid.push_back(0);
    }
    const size_t fbytes = (size_t)rows * cols;
    if (2 * (size_t)C * fbytes > ctx->raw_bytes) {
        // growing the staging buffer: nothing may be in flight
        if (int r = yavo_wait(ctx)) return r;
        if (int r = ensure_raw(ctx, 2 * (size_t)C * fbytes)) return r;
        ctx->raw_used[0] = ctx->raw_used[1] = 0;
    }
    // results of an earlier submit still travelling out of the slots this one overwrites: same chunking ->
    // wait chunk by chunk, otherwise wait for all of them
    const bool same_shape = ctx->fetched_C == C;
    if (!same_shape)
        for (size_t c = 0; c < ctx->fetched_valid.size(); c++)
            if (ctx->fetched_valid[c]) {
                CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_fetched[c], 0));
                ctx->fetched_valid[c] = 0;
            }
    ctx->fetched_C = C;
    const bool track = ctx->strk_on && ctx->strk_xy != nullptr;
    int top = 0;
    if (track) {
        top = klt_levels_for(rows, cols, ctx->strk_P.ww, ctx->strk_P.wh, ctx->strk_max_level);
        if (int r = ensure_pyramid_alloc(ctx)) return r;
        if (int r = ensure_track_buffers(ctx)) return r;
    }
    if (ctx->raw_C != C || ctx->raw_C * fbytes == 0) {
        // the staging halves move when the stage size changes: a new half can cover BOTH halves of the previous
        // submit, so the first copy waits until both have been re-pitched (not only the one with the same index)
        for (int b = 0; b < 2; b++)
            if (ctx->raw_used[b]) CK(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_repitched[b], 0));
    }
    ctx->raw_C = C;
    // per-ticket status word: what the kernels of THIS batch flag (candidate overflow, select watchdog)
    const int ticket = (int)(ctx->n_tickets++ % 16);
    ctx->cur_status = ctx->d_status + ticket;
    CK(cudaMemsetAsync(ctx->cur_status, 0, sizeof(int), ctx->stream));
    struct StatusScope {
        yavo_ctx *c;
        ~StatusScope() { c->cur_status = c->d_status + STATUS_GENERAL; }
    } status_scope{ctx};
    for (int c = 0; c < nchunks; c++) {
        const int buf = c & 1, s0 = c * C, nc = std::min(C, n - s0);
        uint8_t *raw = ctx->d_raw + (size_t)buf * C * fbytes;
        if (ctx->raw_used[buf]) CK(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_repitched[buf], 0));
        CK(cudaMemcpyAsync(raw, pixels + (size_t)s0 * fbytes, (size_t)nc * fbytes, cudaMemcpyHostToDevice, ctx->s_h2d));
        CK(cudaEventRecord(ctx->ev_h2d[buf], ctx->s_h2d));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[buf], 0));
        if (ctx->fetched_valid[c]) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_fetched[c], 0));
        // the match of (slot s0-1, s0) reads the previous chunk's descriptors: in order on ctx->stream
        if (int r = launch_repitch(ctx, raw, cols, s0, nc, rows, cols)) return r;
        CK(cudaEventRecord(ctx->ev_repitched[buf], ctx->stream));
        ctx->raw_used[buf] = 1;
        if (int r = frontend_range(ctx, s0, nc, do_match != 0, c > 0)) return r;
        // streaming tracking: the keypoints of frame f tracked into f+1 for the pairs that END in this chunk (the
        // pair across the chunk seam uses the previous chunk's keypoints and pyramid, still in place)
        const int p0 = c > 0 ? s0 - 1 : s0, tpairs = track ? s0 + nc - 1 - p0 : 0;
        if (track) {
            if (top > 0) {
                if (int r = launch_pyramid(ctx, s0, nc, top)) return r;
            }
            if (tpairs > 0) {
                dim3 grid((ctx->max_kp + KLT_WARPS - 1) / KLT_WARPS, tpairs);
                const size_t o = (size_t)p0 * ctx->max_kp;
                if (int r = launch_klt(ctx, grid, klt_levels_struct(ctx, rows, cols, top), ctx->strk_P, p0, p0 + 1, nullptr,
                                       ctx->d_kp_row, ctx->d_kp_col, ctx->d_nkp, 0, ctx->max_kp, nullptr, ctx->d_trk_xy + o,
                                       ctx->d_trk_status + o, ctx->d_trk_err + o))
                    return r;
            }
        }
        CK(cudaEventRecord(ctx->ev_done[c], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_done[c], 0));
        if (int r = fetch_async(ctx, ctx->s_d2h, 0, s0, nc, n_kp, out_rows, out_cols, scores, desc, match_idx, match_dist))
            return r;
        if (tpairs > 0) {
            const size_t o = (size_t)p0 * ctx->max_kp, cnt = (size_t)tpairs * ctx->max_kp;
            CK(cudaMemcpyAsync(ctx->strk_xy + 2 * o, ctx->d_trk_xy + o, sizeof(float2) * cnt, cudaMemcpyDeviceToHost, ctx->s_d2h));
            if (ctx->strk_status) CK(cudaMemcpyAsync(ctx->strk_status + o, ctx->d_trk_status + o, cnt, cudaMemcpyDeviceToHost, ctx->s_d2h));
            if (ctx->strk_err) CK(cudaMemcpyAsync(ctx->strk_err + o, ctx->d_trk_err + o, sizeof(float) * cnt, cudaMemcpyDeviceToHost, ctx->s_d2h));
        }
        CK(cudaEventRecord(ctx->ev_fetched[c], ctx->s_d2h));
        ctx->fetched_valid[c] = 1;
    }
    // ticket: an event after the last D2H copy of this batch; the batch's status word travels with it
    CK(cudaMemcpyAsync(ctx->h_small + 32 + ticket, ctx->d_status + ticket, sizeof(int), cudaMemcpyDeviceToHost, ctx->s_d2h));
    ctx->ticket_pending[ticket] = 1;
    if (!ctx->ev_ticket[ticket]) CK(cudaEventCreateWithFlags(&ctx->ev_ticket[ticket], cudaEventDisableTiming | cudaEventBlockingSync));
    CK(cudaEventRecord(ctx->ev_ticket[ticket], ctx->s_d2h));
    return ticket;
}

int yavo_wait_batch(yavo_ctx *ctx, int ticket) {
    if (!ctx || ticket < 0 || ticket >= 16 || !ctx->ev_ticket[ticket]) return YAVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev_ticket[ticket]));
    if (!ctx->ticket_pending[ticket]) return 0;
    ctx->ticket_pending[ticket] = 0;
    return status_error(ctx, ctx->h_small[32 + ticket]);  // overflow / watchdog of THIS batch
}

int yavo_wait(yavo_ctx *ctx) {
    if (!ctx) return YAVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->s_h2d) CK(cudaStreamSynchronize(ctx->s_h2d));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->s_d2h) CK(cudaStreamSynchronize(ctx->s_d2h));
    for (size_t c = 0; c < ctx->fetched_valid.size(); c++) ctx->fetched_valid[c] = 0;
    int st = 0;  // batches nobody called yavo_wait_batch for
    for (int t = 0; t < 16; t++)
        if (ctx->ticket_pending[t]) {
            ctx->ticket_pending[t] = 0;
            if (ctx->h_small[32 + t] != 0 && st == 0) st = ctx->h_small[32 + t];
        }
    if (int r = check_status(ctx)) return r;
    return status_error(ctx, st);
}

int yavo_process_host_batch(yavo_ctx *ctx, const uint8_t *pixels, int n, int rows, int cols, int do_match,
                            int32_t *n_kp, int32_t *out_rows, int32_t *out_cols, float *scores, uint8_t *desc,
                            int32_t *match_idx, int32_t *match_dist) {
    if (int r = check_slot(ctx, 0, n)) return r;
    if (n == 0) return 0;
    if (!pixels || rows < 1 || cols < 1 || rows > ctx->max_rows || cols > ctx->max_cols)
        return fail(ctx, YAVO_ERR_INVALID, "bad frame size %dx%d", rows, cols);
    if (!ctx->offs_set) return fail(ctx, YAVO_ERR_STATE, "yavo_set_brief_offsets has not been called");
    CK(cudaSetDevice(ctx->device));
    if (!is_pinned_host(pixels)) {
        // pageable frames: packed through the pinned staging buffer, no overlap
        if (int r = yavo_wait(ctx)) return r;
        if (int r = upload_host(ctx, 0, n, pixels, rows, cols, cols)) return r;
        if (int r = frontend_range(ctx, 0, n, do_match != 0, false)) return r;
        return yavo_fetch_batch(ctx, 0, n, n_kp, out_rows, out_cols, scores, desc, match_idx, match_dist);
    }
    const int t = yavo_submit_host_batch(ctx, pixels, n, rows, cols, do_match, n_kp, out_rows, out_cols, scores, desc,
                                         match_idx, match_dist);
    if (t < 0) return t;
    return yavo_wait(ctx);
}

int yavo_filter_pairs(yavo_ctx *ctx, int slot0, int n, int threshold, int32_t *n_pairs, int32_t *min_dist,
                      int32_t *pairs) {
    if (int r = check_slot(ctx, slot0, n)) return r;
    if (n < 2) return 0;
    CK(cudaSetDevice(ctx->device));
    const size_t o = (size_t)slot0 * ctx->max_kp;
    PROF(KC_FILTER, filter_pairs_kernel<<<n - 1, K6_THREADS, 0, ctx->stream>>>(
        ctx->d_midx + o, ctx->d_mdist + o, ctx->d_bk_row + o, ctx->d_bk_col + o, ctx->d_nbk + slot0, ctx->max_kp,
        threshold, ctx->d_pairs + o * 8, ctx->d_npairs + slot0, ctx->d_minDist + slot0));
    CK_LAUNCH();
    if (n_pairs) CK(cudaMemcpyAsync(n_pairs, ctx->d_npairs + slot0, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (min_dist) CK(cudaMemcpyAsync(min_dist, ctx->d_minDist + slot0, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (pairs)
        CK(cudaMemcpyAsync(pairs, ctx->d_pairs + o * 8, 32 * (size_t)n * ctx->max_kp, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

#ifdef YAVO_SEL_TIMING
// debug build only: phase timestamps (clock64) of the first 64 select CTAs of the last launch
extern "C" int yavo_debug_select_timing(yavo_ctx *ctx, long long *out /* 64 x 8 */) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(out, ctx->d_scratch, sizeof(long long) * (64 * 8 + 8 * 16 * 4), cudaMemcpyDeviceToHost));
    long long dbg[8], zero[8] = {0};
    CK(cudaMemcpyFromSymbol(dbg, yavo::g_sel_dbg, sizeof dbg));
    CK(cudaMemcpyToSymbol(yavo::g_sel_dbg, zero, sizeof zero));
    fprintf(stderr, "select dbg (CTA 0, all launches since last call): partition cycles %lld count %lld elems %lld | leaf cycles %lld count %lld\n",
            dbg[0], dbg[1], dbg[4], dbg[2], dbg[3]);
    return 0;
}
#endif

int yavo_set_matcher(yavo_ctx *ctx, int kind) {
    if (!ctx || kind < 0 || kind > 2) return YAVO_ERR_INVALID;
    ctx->matcher = kind;
    return 0;
}

int yavo_set_overlap(yavo_ctx *ctx, int chunk_frames, int n_streams) {
    if (!ctx || chunk_frames < 0 || n_streams < 1 || n_streams > yavo_ctx::MAX_AUX) return YAVO_ERR_INVALID;
    ctx->ov_chunk = chunk_frames;
    ctx->ov_streams = n_streams;
    return 0;
}

int yavo_set_big_select(yavo_ctx *ctx, int min_candidates) {
    if (!ctx || min_candidates < -1) return YAVO_ERR_INVALID;
    ctx->big_min = min_candidates;
    drop_frame_graphs(ctx);  // captured launches were made with the old setting
    return 0;
}

int yavo_set_sub_batch(yavo_ctx *ctx, int frames) {
    if (!ctx || frames < 0) return YAVO_ERR_INVALID;
    ctx->sub_batch = frames;
    return 0;
}

/* frames per pipeline stage of yavo_process_host_batch (default 32) */
int yavo_set_pipeline_chunk(yavo_ctx *ctx, int frames) {
    if (!ctx || frames < 0) return YAVO_ERR_INVALID;
    ctx->pipeline_chunk = frames;
    return 0;
}

// ---- tracking step (cv::calcOpticalFlowPyrLK, src/LoopHandler.cc:372-375) ------------------------------------

int yavo_build_pyramid(yavo_ctx *ctx, int slot0, int n, int win_w, int win_h, int max_level, int *levels) {
    if (int r = check_slot(ctx, slot0, n)) return r;
    if (int r = klt_check_params(ctx, win_w, win_h, max_level, 0)) return r;
    if (levels) *levels = 0;
    if (n == 0) return 0;
    if (int r = check_uploaded(ctx, slot0, n)) return r;
    CK(cudaSetDevice(ctx->device));
    const int top = klt_levels_for(ctx->slot_rows[slot0], ctx->slot_cols[slot0], win_w, win_h, max_level);
    if (levels) *levels = top;
    return top > 0 ? launch_pyramid(ctx, slot0, n, top) : 0;
}

int yavo_pyramid_level(yavo_ctx *ctx, int slot, int level, uint8_t *out, size_t out_bytes, int *rows, int *cols) {
    if (int r = check_slot(ctx, slot)) return r;
    if (int r = check_uploaded(ctx, slot, 1)) return r;
    if (level < 1 || level > ctx->slot_pyr[slot])
        return fail(ctx, YAVO_ERR_STATE, "slot %d holds pyramid levels 1..%d, asked for %d", slot, ctx->slot_pyr[slot], level);
    int h, w;
    level_dims(ctx->slot_rows[slot], ctx->slot_cols[slot], level, &h, &w);
    if (rows) *rows = h;
    if (cols) *cols = w;
    if (!out || out_bytes < (size_t)h * w) return fail(ctx, YAVO_ERR_INVALID, "level %d is %dx%d, buffer holds %zu bytes", level, h, w, out_bytes);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy2DAsync(out, w, ctx->d_pyr + ctx->pyr_slot_stride * slot + ctx->pyr_off[level], ctx->pyr_pitch[level], w, h,
                         cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int yavo_klt_track(yavo_ctx *ctx, int slot_prev, int slot_next, const float *prev_xy, int n, float *next_xy,
                   uint8_t *status, float *err, int win_w, int win_h, int max_level, int crit_type, int max_count,
                   double epsilon, int flags, double min_eig_threshold) {
    if (int r = check_slot(ctx, slot_prev)) return r;
    if (int r = check_slot(ctx, slot_next)) return r;
    if (int r = klt_check_params(ctx, win_w, win_h, max_level, flags)) return r;
    if (n < 0 || (n > 0 && (!prev_xy || !next_xy || !status))) return fail(ctx, YAVO_ERR_INVALID, "bad point arrays (n=%d)", n);
    if (n == 0) return 0;
    if (int r = check_uploaded(ctx, slot_prev, 1)) return r;
    if (int r = check_uploaded(ctx, slot_next, 1)) return r;
    if (ctx->slot_rows[slot_prev] != ctx->slot_rows[slot_next] || ctx->slot_cols[slot_prev] != ctx->slot_cols[slot_next])
        return fail(ctx, YAVO_ERR_INVALID, "slots %d and %d hold frames of different sizes", slot_prev, slot_next);
    CK(cudaSetDevice(ctx->device));
    const int H = ctx->slot_rows[slot_prev], W = ctx->slot_cols[slot_prev];
    const int top = klt_levels_for(H, W, win_w, win_h, max_level);
    if (top > 0) {
        if (int r = launch_pyramid(ctx, slot_prev, 1, top)) return r;
        if (int r = launch_pyramid(ctx, slot_next, 1, top)) return r;
    } else if (int r = ensure_pyramid_alloc(ctx)) return r;
    if (n > ctx->klt_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        void *old[] = {ctx->d_klt_prev, ctx->d_klt_next, ctx->d_klt_status, ctx->d_klt_err};
        for (void *b : old)
            if (b) CK(cudaFree(b));
        ctx->d_klt_prev = ctx->d_klt_next = nullptr;
        ctx->d_klt_status = nullptr;
        ctx->d_klt_err = nullptr;
        ctx->klt_cap = 0;
        const int cap = std::max(n, 4096);
        CK(dalloc(&ctx->d_klt_prev, cap));
        CK(dalloc(&ctx->d_klt_next, cap));
        CK(dalloc(&ctx->d_klt_status, cap));
        CK(dalloc(&ctx->d_klt_err, cap));
        ctx->klt_cap = cap;
    }
    CK(cudaMemcpyAsync(ctx->d_klt_prev, prev_xy, sizeof(float2) * n, cudaMemcpyHostToDevice, ctx->stream));
    if (flags & 4) CK(cudaMemcpyAsync(ctx->d_klt_next, next_xy, sizeof(float2) * n, cudaMemcpyHostToDevice, ctx->stream));
    const KltParams P = klt_params(win_w, win_h, crit_type, max_count, epsilon, flags, min_eig_threshold);
    const KltLevels L = klt_levels_struct(ctx, H, W, top);
    dim3 grid((n + KLT_WARPS - 1) / KLT_WARPS, 1);
    if (int r = launch_klt(ctx, grid, L, P, slot_prev, slot_next, ctx->d_klt_prev, nullptr, nullptr, nullptr, n, 0,
                           ctx->d_klt_next, ctx->d_klt_next, ctx->d_klt_status, ctx->d_klt_err))
        return r;
    CK(cudaMemcpyAsync(next_xy, ctx->d_klt_next, sizeof(float2) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(status, ctx->d_klt_status, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (err) CK(cudaMemcpyAsync(err, ctx->d_klt_err, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int yavo_klt_track_batch(yavo_ctx *ctx, int slot0, int n, int win_w, int win_h, int max_level, int crit_type,
                         int max_count, double epsilon, int flags, double min_eig_threshold) {
    if (int r = check_slot(ctx, slot0, n)) return r;
    if (int r = klt_check_params(ctx, win_w, win_h, max_level, flags)) return r;
    if (flags & 4) return fail(ctx, YAVO_ERR_INVALID, "the batch form starts from the keypoints themselves (no OPTFLOW_USE_INITIAL_FLOW)");
    if (n < 2) return 0;
    if (int r = check_uploaded(ctx, slot0, n)) return r;
    CK(cudaSetDevice(ctx->device));
    const int H = ctx->slot_rows[slot0], W = ctx->slot_cols[slot0];
    const int top = klt_levels_for(H, W, win_w, win_h, max_level);
    if (top > 0) {
        if (int r = launch_pyramid(ctx, slot0, n, top)) return r;
    } else if (int r = ensure_pyramid_alloc(ctx)) return r;
    if (int r = ensure_track_buffers(ctx)) return r;
    const KltParams P = klt_params(win_w, win_h, crit_type, max_count, epsilon, flags, min_eig_threshold);
    const KltLevels L = klt_levels_struct(ctx, H, W, top);
    dim3 grid((ctx->max_kp + KLT_WARPS - 1) / KLT_WARPS, n - 1);
    const size_t o = (size_t)slot0 * ctx->max_kp;
    // keypoint arrays and counts are indexed by absolute slot inside the kernel; outputs by pair from `o`
    return launch_klt(ctx, grid, L, P, slot0, slot0 + 1, nullptr, ctx->d_kp_row, ctx->d_kp_col, ctx->d_nkp, 0, ctx->max_kp,
                      nullptr, ctx->d_trk_xy + o, ctx->d_trk_status + o, ctx->d_trk_err + o);
}

int yavo_stream_tracking(yavo_ctx *ctx, int enable, int win_w, int win_h, int max_level, int crit_type, int max_count,
                         double epsilon, int flags, double min_eig_threshold) {
    if (!ctx) return YAVO_ERR_INVALID;
    if (!enable) {
        ctx->strk_on = false;
        return 0;
    }
    if (int r = klt_check_params(ctx, win_w, win_h, max_level, flags)) return r;
    if (flags & 4) return fail(ctx, YAVO_ERR_INVALID, "streaming tracking starts from the keypoints themselves (no OPTFLOW_USE_INITIAL_FLOW)");
    ctx->strk_P = klt_params(win_w, win_h, crit_type, max_count, epsilon, flags, min_eig_threshold);
    ctx->strk_max_level = max_level;
    ctx->strk_on = true;
    return 0;
}

int yavo_stream_track_outputs(yavo_ctx *ctx, float *next_xy, uint8_t *status, float *err) {
    if (!ctx) return YAVO_ERR_INVALID;
    ctx->strk_xy = next_xy;
    ctx->strk_status = status;
    ctx->strk_err = err;
    return 0;
}

int yavo_klt_fetch(yavo_ctx *ctx, int slot0, int n, float *next_xy, uint8_t *status, float *err) {
    if (int r = check_slot(ctx, slot0, n)) return r;
    if (n == 0) return 0;
    if (!ctx->d_trk_xy) return fail(ctx, YAVO_ERR_STATE, "yavo_klt_track_batch has not run");
    CK(cudaSetDevice(ctx->device));
    const size_t o = (size_t)slot0 * ctx->max_kp, cnt = (size_t)n * ctx->max_kp;
    if (next_xy) CK(cudaMemcpyAsync(next_xy, ctx->d_trk_xy + o, sizeof(float2) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
    if (status) CK(cudaMemcpyAsync(status, ctx->d_trk_status + o, cnt, cudaMemcpyDeviceToHost, ctx->stream));
    if (err) CK(cudaMemcpyAsync(err, ctx->d_trk_err + o, sizeof(float) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- inlier count of the F-matrix RANSAC (src/3DHandler.cc:163-190) -----------------------------------------

int yavo_epipolar_inliers(yavo_ctx *ctx, const double *F, int m, const int32_t *x1, const int32_t *y1,
                          const int32_t *x2, const int32_t *y2, int n, double threshold, int32_t *counts,
                          int32_t *best, int32_t *best_count, double *residuals) {
    if (!ctx) return YAVO_ERR_INVALID;
    if (m < 0 || n < 0 || (m > 0 && (!F || !counts)) || (n > 0 && (!x1 || !y1 || !x2 || !y2)))
        return fail(ctx, YAVO_ERR_INVALID, "bad arguments (m=%d, n=%d)", m, n);
    if (best) *best = -1;
    if (best_count) *best_count = INT_MIN;
    if (m == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    double *dF = nullptr, *dres = nullptr;
    int32_t *dpts = nullptr, *dcnt = nullptr;
    // small, call-scoped buffers: the reference calls this once per (re)initialisation (src/LoopHandler.cc:225,567)
    cudaError_t e = cudaSuccess;
    auto cleanup = [&]() {
        if (dF) cudaFree(dF);
        if (dres) cudaFree(dres);
        if (dpts) cudaFree(dpts);
        if (dcnt) cudaFree(dcnt);
    };
#define CKE(call)                                                                                          \
    do {                                                                                                   \
        e = (call);                                                                                        \
        if (e != cudaSuccess) {                                                                            \
            cleanup();                                                                                     \
            return fail(ctx, YAVO_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e));               \
        }                                                                                                  \
    } while (0)
    CKE(dalloc(&dF, (size_t)m * 9));
    CKE(dalloc(&dpts, (size_t)std::max(n, 1) * 4));
    CKE(dalloc(&dcnt, (size_t)m + 2));
    if (residuals && n > 0) CKE(dalloc(&dres, (size_t)m * n));
    CKE(cudaMemcpyAsync(dF, F, sizeof(double) * 9 * m, cudaMemcpyHostToDevice, ctx->stream));
    const int32_t *src[4] = {x1, y1, x2, y2};
    for (int k = 0; k < 4 && n > 0; k++)
        CKE(cudaMemcpyAsync(dpts + (size_t)k * n, src[k], sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    PROF(KC_EPI, epipolar_inliers_kernel<<<m, 256, 0, ctx->stream>>>(dF, dpts, dpts + n, dpts + 2 * (size_t)n, dpts + 3 * (size_t)n, n,
                                                                    threshold, dcnt, dres));
    ctx->launches++;
    first_max_kernel<<<1, 32, 0, ctx->stream>>>(dcnt, m, dcnt + m);
    ctx->launches++;
    CKE(cudaGetLastError());
    std::vector<int32_t> h((size_t)m + 2);
    CKE(cudaMemcpyAsync(h.data(), dcnt, sizeof(int32_t) * (m + 2), cudaMemcpyDeviceToHost, ctx->stream));
    if (dres) CKE(cudaMemcpyAsync(residuals, dres, sizeof(double) * (size_t)m * n, cudaMemcpyDeviceToHost, ctx->stream));
    CKE(cudaStreamSynchronize(ctx->stream));
#undef CKE
    cleanup();
    memcpy(counts, h.data(), sizeof(int32_t) * m);
    if (best) *best = h[m];
    if (best_count) *best_count = h[m + 1];
    return 0;
}

// ---- pinned host memory for callers without a CUDA runtime binding of their own -------------------------------

void *yavo_pinned_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void yavo_pinned_free(void *p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
