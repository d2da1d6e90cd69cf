/*
 * yavo_b200.h — C ABI of the B200-native (sm_100a) ORB-style front end of YA_VO.
 *
 * The reference (kartikmadhira1/YA_VO) has no FFI or plugin layer: its hot path is the public
 * member functions of three concrete C++ classes.  This header is the boundary a binding of
 * that path would use; every entry point names the reference interface it replaces
 * (paths relative to the reference checkout).  The C++ classes in ya_vo_b200/host/ keep the
 * reference signatures and call these functions; ya_vo_b200/capi.py binds them with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross the boundary
 *   - every function returns 0 on success, a negative yavo_status otherwise; the message is
 *     available from yavo_last_error(ctx) (yavo_last_error(NULL) for yavo_create failures)
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails
 *   - coordinates follow the reference: a point is (x = row, y = col)
 *   - a context is bound to one device and one stream; calls on one context must not overlap
 *     (the reference's hot path is only ever entered from one thread, src/main.cc:11);
 *     use one context per GPU / host thread
 *   - host pointers may be pageable; uploads are staged through pinned memory
 */
#ifndef YAVO_B200_H
#define YAVO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct yavo_ctx yavo_ctx;

typedef enum yavo_status {
    YAVO_OK = 0,
    YAVO_ERR_INVALID = -1,    /* bad argument */
    YAVO_ERR_CUDA = -2,       /* CUDA runtime error (message has the detail) */
    YAVO_ERR_CAPACITY = -3,   /* a device buffer sized at yavo_create was too small */
    YAVO_ERR_STATE = -4       /* call sequence error (e.g. slot never uploaded) */
} yavo_status;

/* Reference constants (include/FastDetector.hpp:32-38, src/FastDetector.cc:147, src/LoopHandler.cc:7) */
#define YAVO_FAST_THRESHOLD 40
#define YAVO_FAST_RUN 12
#define YAVO_FAST_MAX_KEYPOINTS 2000
#define YAVO_BRIEF_TESTS 256
#define YAVO_DESC_BYTES 32

/* ---- context ---------------------------------------------------------------------------------- */

/* One context per GPU.  n_slots device-resident frames of at most max_rows x max_cols pixels;
 * max_kp keypoints kept per frame (the reference's fastCornerNumThreshold, include/FastDetector.hpp:36);
 * max_cand = capacity of the per-frame FAST candidate list (0 = (max_rows-8)*(max_cols-8)/4,
 * i.e. every fourth interior pixel; at most 2^24 - 1).  Uniform noise yields 3.7 % of the pixels, real frames under 1 %;
 * periodic textures can reach 50 % (tests/test_gpu_round2.py: DENSE_PATTERN) — a frame with more candidates than max_cand
 * is reported with YAVO_ERR_CAPACITY by the call (or the batch ticket) that saw it, never truncated silently.  */
int yavo_create(int device, int n_slots, int max_rows, int max_cols, int max_kp, int max_cand,
                yavo_ctx **out);
void yavo_destroy(yavo_ctx *ctx);
const char *yavo_last_error(const yavo_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
long long yavo_kernel_launches(const yavo_ctx *ctx);
/* block until everything queued on the context's stream has finished */
int yavo_sync(yavo_ctx *ctx);
/* the context's cudaStream_t (as void*), so callers can record their own CUDA events on it */
void *yavo_get_stream(yavo_ctx *ctx);
/* measurement: when on, every kernel launch is bracketed by a CUDA event pair on the context's stream.
 * yavo_profile_collect synchronises and returns, per kernel class
 * {0 repitch, 1 detect_blur, 2 compact_score, 3 select_topk, 4 brief, 5 match_partial, 6 match_reduce,
 *  7 filter_pairs, 8 pyr_down, 9 klt_track, 10 epipolar_inliers, 11 match_tc},
 * the summed device time in ms and the number of launches since the last collect / set_profiling. */
int yavo_set_profiling(yavo_ctx *ctx, int on);
int yavo_profile_collect(yavo_ctx *ctx, double *ms_per_class, int *launches_per_class, int n_classes);

/* ---- Image (include/Image.hpp:14-28, src/Image.cc:8-17) ---------------------------------------- */

/* Image::Image(const cv::Mat&): deep copy of an 8-bit single-channel frame into device slot `slot`.
 * `stride` is the host row pitch in bytes (cols for a continuous Mat).  Pageable memory is packed
 * into pinned staging before the call returns; a continuous frame in pinned (cudaHostAlloc /
 * cudaHostRegister) memory is copied asynchronously straight from the caller's buffer, which must
 * then stay unchanged until the next yavo_sync / fetch on this context. */
int yavo_upload(yavo_ctx *ctx, int slot, const uint8_t *pixels, int rows, int cols, int stride);
/* n continuous frames (n x rows x cols) into slots [slot0, slot0+n) */
int yavo_upload_batch(yavo_ctx *ctx, int slot0, int n, const uint8_t *pixels, int rows, int cols);
/* same, from pixels already in device memory (row pitch `pitch` bytes, frame pitch rows*pitch) */
int yavo_upload_from_device(yavo_ctx *ctx, int slot0, int n, const void *d_pixels, int rows, int cols,
                            size_t pitch);
/* Image::getPixelVal / rawImage read-back (tests): copies slot pixels to a continuous host buffer */
int yavo_download(yavo_ctx *ctx, int slot, uint8_t *pixels, int rows, int cols);

/* ---- FastDetector (include/FastDetector.hpp:17-55, src/FastDetector.cc) ------------------------ */

/* FastDetector::getBresenhamCirclePoints (src/FastDetector.cc:50-112): the 16 ring points around
 * (xc, yc) in ring order, out_xy = {x0,y0,...,x15,y15}.  Host-side constant table. */
void yavo_ring_points(int xc, int yc, int32_t *out_xy);

/* FastDetector::getFastFeatures (src/FastDetector.cc:277-369): segment test over the interior,
 * Harris response of every passing pixel, reference ordering (std::sort by response, descending,
 * replayed exactly including its treatment of tied responses), first max_kp points.
 *   out_rows/out_cols/out_scores: capacity max_kp each (out_scores may be NULL)
 *   *n_out: points written, *n_cand: pixels that passed the segment test (both may be NULL)
 * max_kp <= 0 selects the context's max_kp.  The blurred plane BRIEF needs is produced by the
 * same kernel and cached in the slot. */
int yavo_fast_detect(yavo_ctx *ctx, int slot, int max_kp, int32_t *out_rows, int32_t *out_cols,
                     float *out_scores, int *n_out, int *n_cand);
/* The reference's per-frame sequence in ONE call (LoopHandler::insertFrameFeatures, src/LoopHandler.cc:468-485:
 * FastDetector::getFastFeatures on a fresh frame, then Brief::computeBrief on the points it returned): uploads the host
 * frame (rows x cols, row pitch `stride`) into `slot`, detects, orders and cuts to max_kp as yavo_fast_detect does, and
 * describes the points checkBoundry admits as yavo_brief_describe would.  The frame is staged in pinned memory at the
 * device row pitch and copied straight into the slot (no re-pitch kernel on this path); the whole sequence — copy in,
 * kernels, copy out — is captured once per (slot, frame size, max_kp) as a CUDA graph between pinned staging buffers, so
 * a call costs one graph launch and one synchronisation.  One frame runs the select kernel's 768-thread instance.  yavo_set_brief_offsets must have been called.
 *   kp_rows / kp_cols / kp_scores [max_kp], *n_kp: the detector's output (kp_scores may be NULL)
 *   d_rows / d_cols / d_ids [max_kp], desc [max_kp x 32], *n_desc: the admitted points in order; d_ids = index of the
 *   point in the kp list (the `id` Brief::computeBrief gives the KeyPoint, src/BriefDescriptor.cc:95)
 *   *n_cand: pixels that passed the segment test.  Any output pointer may be NULL. */
int yavo_frame_features(yavo_ctx *ctx, int slot, const uint8_t *pixels, int rows, int cols, int stride, int max_kp,
                        int32_t *n_kp, int32_t *kp_rows, int32_t *kp_cols, float *kp_scores, int32_t *n_desc,
                        int32_t *d_rows, int32_t *d_cols, int32_t *d_ids, uint8_t *desc, int *n_cand);
/* 1 when `slot` was last filled by yavo_frame_features with exactly these pixels (compared byte for byte with the pinned
 * host copy that call kept), 0 when not or unknown, < 0 on bad arguments.  Lets a caller that owns mutable host pixels
 * (Image::rawImage is public in the reference) decide whether device-side results for the slot are still valid. */
int yavo_slot_holds(yavo_ctx *ctx, int slot, const uint8_t *pixels, int rows, int cols, int stride);
/* the unsorted candidate list in scan (row-major) order, as the reference's retCorners holds it
 * before the sort (src/FastDetector.cc:298-335); for parity tests. capacity `cap` entries. */
int yavo_fast_candidates(yavo_ctx *ctx, int slot, int cap, int32_t *out_rows, int32_t *out_cols,
                         float *out_scores, int *n_cand);

/* ---- Brief (include/BriefDescriptor.hpp:41-68, src/BriefDescriptor.cc) ------------------------- */

/* Brief::preComputeOffsets result (src/BriefDescriptor.cc:4-20): 256 x {drow1,dcol1,drow2,dcol2},
 * each in [-8, 8].  The host wrapper draws the table as the reference does and passes it here. */
int yavo_set_brief_offsets(yavo_ctx *ctx, const int32_t *offsets /* 1024 */);

/* cv::GaussianBlur(rawImage, 9x9, 2.5) as Brief::computeBrief applies it (src/BriefDescriptor.cc:90):
 * read-back of the slot's smoothed plane (computed on demand); for parity tests. */
int yavo_blurred(yavo_ctx *ctx, int slot, uint8_t *out, int rows, int cols);

/* Brief::computeBrief (src/BriefDescriptor.cc:86-124) for n points (x=row, y=col), in input order.
 *   out_desc: n x 32 bytes, bit j of the descriptor in byte j/8, bit j%8 (:108-117)
 *   out_valid: n bytes, 1 where checkBoundry (:128-136) admits the point; descriptors of rejected
 *              points are zero and the caller drops them (the reference does not append them)
 *   *n_oob (may be NULL): admitted points with a test read at linear index >= rows*cols
 *              (undefined behaviour in the reference; such reads are defined as pixel value 0) */
int yavo_brief_describe(yavo_ctx *ctx, int slot, const int32_t *rows, const int32_t *cols, int n,
                        uint8_t *out_desc, uint8_t *out_valid, int *n_oob);

/* Brief::matchFeatures (src/BriefDescriptor.cc:163-183) on two descriptor sets (n x 32 bytes):
 * for every query i the train index with the smallest Hamming distance (:139-160), lowest index
 * among equal minima; n2 == 0 gives idx -1 and dist INT_MAX.
 * Extensions with no reference counterpart (may be NULL): out_second = second smallest distance
 * per query (ratio test), out_rev_idx[j] = best query for train j (cross-check). */
int yavo_match(yavo_ctx *ctx, const uint8_t *d1, int n1, const uint8_t *d2, int n2,
               int32_t *out_idx, int32_t *out_dist, int32_t *out_second, int32_t *out_rev_idx);

/* Which kernel computes the matches: 0 (default) = K5t4, tcgen05 tensor cores (kind::mxf4) on +-1 e2m1 expansions of
 * the descriptor bits; 2 = K5t, the same on FP8 expansions (kind::f8f6f4); 1 = K5, XOR / carry-save / POPC on the integer
 * pipes.  All three give identical results (every partial sum of the tensor-core forms is an integer < 2^16, exact in the
 * FP32 accumulator); out_second of yavo_match always comes from K5. */
int yavo_set_matcher(yavo_ctx *ctx, int kind);

/* Brief::removeOutliers (src/BriefDescriptor.cc:213-231): keep[i] = dist[i] < max(2*min(dist), threshold).
 * O(n) host arithmetic on the distances yavo_match returned; returns the kept count (>= 0). */
int yavo_remove_outliers(const int32_t *dist, int n, int threshold, uint8_t *keep);

/* ---- batch front end (LoopHandler::insertFrameFeatures over many frames, src/LoopHandler.cc:468-485,
 *      plus matchFeatures(frame f-1, frame f), :189,534) -------------------------------------------- */

/* Runs detect -> top-K -> describe on slots [slot0, slot0+n) and, if do_match, matches the kept
 * descriptors of slot f-1 (queries) against slot f (train) for f in (slot0, slot0+n).  Results stay
 * in device memory; nothing is copied to the host.  Frames must all have the same size. */
int yavo_frontend_batch(yavo_ctx *ctx, int slot0, int n, int do_match);

/* Copies the results of yavo_frontend_batch for n slots to caller arrays — host memory, or device memory of the context's
 * GPU (the copies use unified addressing; a multi-GPU caller gathers device buffers over NVLink) — any may be NULL:
 *   n_kp[n]; rows/cols/scores [n x max_kp]; desc [n x max_kp x 32]; only keypoints admitted by
 *   checkBoundry are kept, compacted in order, as Brief::computeBrief appends them;
 *   match_idx/match_dist [n x max_kp]: entry (f, i) is the match of keypoint i of slot f-1 in slot f
 *   (row 0 of the batch is unused). */
int yavo_fetch_batch(yavo_ctx *ctx, int slot0, int n, int32_t *n_kp, int32_t *rows, int32_t *cols,
                     float *scores, uint8_t *desc, int32_t *match_idx, int32_t *match_dist);

/* Brief::removeOutliers (src/BriefDescriptor.cc:213-231) fused with the conversion of the kept matches to
 * point pairs that the callers perform (src/LoopHandler.cc:232-237,251-254), on the device, for the matches
 * yavo_frontend_batch left in slots (slot0, slot0+n).  For every f in [1, n): n_pairs[f] kept matches of the
 * pair (slot0+f-1, slot0+f), min_dist[f] the smallest distance, pairs[f][j] = {q_row, q_col, t_row, t_col,
 * dist, q_index, t_index, 0} in match order (arrays are n / n / n x max_kp x 8; entry 0 unused; any may be NULL). */
int yavo_filter_pairs(yavo_ctx *ctx, int slot0, int n, int threshold, int32_t *n_pairs, int32_t *min_dist,
                      int32_t *pairs);

/* frames per copy/compute pipeline stage of the host-batch entry points (0 = automatic, the default:
 * a quarter of the batch, clamped to 16..128 frames) */
int yavo_set_pipeline_chunk(yavo_ctx *ctx, int frames);
/* overlapped feature pipeline of the batch entry points: chunks of `chunk_frames` frames rotate over `n_streams`
 * (2..4) internal streams so that the detect kernel of one chunk runs beside the select / BRIEF kernels of the previous
 * ones; chunk_frames = 0 or n_streams = 1 runs the kernels of a batch back to back on one stream (the default: on
 * B200 the overlap measured within 1 % of the serial order, profiles/r2_overlap_sweep.md).
 * Results do not depend on the setting. */
int yavo_set_overlap(yavo_ctx *ctx, int chunk_frames, int n_streams);
/* Large candidate lists (a 3840x2160 frame of noise has 300 k): frames with more than `min_candidates` FAST candidates
 * have the top of their std::sort replay — the partitions that scan hundreds of thousands of elements — spread over a
 * thread-block cluster of 8 CTAs before the select kernel (then a team of 8 CTAs per frame) takes over.  -1 (default) = automatic: 6144 (measured best on config 4 with the round-2 kernels: 3072 9.7 k, 4096 9.9 k, 6144 10.06 k, 8192 9.9 k, 12288 9.5 k frames/s), for
 * frames of 2 Mpx and more; 0 = never; results do not depend on the setting. */
int yavo_set_big_select(yavo_ctx *ctx, int min_candidates);
/* frames per set of kernel launches inside yavo_frontend_batch (0 = the whole batch at once, the default) */

int yavo_set_sub_batch(yavo_ctx *ctx, int frames);

/* streaming callers: upload + frontend + fetch of one batch of host frames into slots [0, n), with the
 * copies inside (what bench.py's e2e leg times).  When `pixels` is pinned host memory the batch is cut
 * into stages and the H2D copy of stage c+1, the kernels of stage c and the D2H copy of stage c-1 run on
 * three streams; pageable memory goes through the staging buffer without overlap.  Returns after all
 * results are in the host arrays. */
/* Asynchronous form for a continuous stream of batches (the caller side of LoopHandler::getNextFrame /
 * insertFrameFeatures, src/LoopHandler.cc:468-485,917-927): queues the H2D copies, kernels and D2H copies of one
 * batch of PINNED host frames and returns at once.  Further submits may follow before yavo_wait — the next
 * batch's pixels cross PCIe while this batch is in the kernels — provided each in-flight batch has its own
 * output arrays.  Input and output arrays must stay valid and untouched until the batch has been waited for;
 * no other call on the context is allowed in between.  Returns a ticket (0..15, reused round-robin) or a
 * negative yavo_status. */
int yavo_submit_host_batch(yavo_ctx *ctx, const uint8_t *pixels, int n, int rows, int cols, int do_match,
                           int32_t *n_kp, int32_t *out_rows, int32_t *out_cols, float *scores, uint8_t *desc,
                           int32_t *match_idx, int32_t *match_dist);
/* blocks until the batch with this ticket has its results in its host arrays (later batches keep running).
 * Returns YAVO_ERR_CAPACITY / YAVO_ERR_CUDA when THIS batch overflowed the FAST candidate list or tripped the select
 * kernel's watchdog (every batch carries its own status word; its frames' results are then invalid). */
int yavo_wait_batch(yavo_ctx *ctx, int ticket);
/* blocks until every submitted batch has finished and its results are in the host arrays; reports a FAST
 * candidate-list overflow of any batch that was not waited for by ticket */
int yavo_wait(yavo_ctx *ctx);

int yavo_process_host_batch(yavo_ctx *ctx, const uint8_t *pixels, int n, int rows, int cols, int do_match,
                            int32_t *n_kp, int32_t *out_rows, int32_t *out_cols, float *scores,
                            uint8_t *desc, int32_t *match_idx, int32_t *match_dist);

/* ---- tracking step: cv::calcOpticalFlowPyrLK (src/LoopHandler.cc:372-375) ----------------------
 * The reference's steady-state per-frame cost after feature extraction: LoopHandler::trackLastFrame projects the
 * last frame's map points and tracks them with OpenCV's sparse pyramidal Lucas-Kanade,
 *   calcOpticalFlowPyrLK(last, cur, lastKpt, curKpt, status, error, Size(11,11), 3,
 *                        TermCriteria(COUNT + EPS, 30, 0.01), 0, 0.001).
 * These entry points run that call on device-resident frame slots.  Unlike the rest of this header, points here
 * are OpenCV's: interleaved float (x = column, y = row), exactly what the reference hands to OpenCV
 * (it swaps its (row, col) keypoints at :343-347).  crit_type: 1 = COUNT, 2 = EPS (cv::TermCriteria::Type);
 * flags: 4 = OPTFLOW_USE_INITIAL_FLOW (next_xy is read), 8 = OPTFLOW_LK_GET_MIN_EIGENVALS.  Window sides 3..31 (OpenCV rejects smaller ones).
 * Results are bit-identical to cv2 4.13 (tests/golden/klt_golden.npz); err is defined where status == 1
 * (OpenCV leaves it uninitialised elsewhere; here it is 0 or the last minimum eigenvalue). */
#define YAVO_KLT_MAX_LEVEL 7
#define YAVO_KLT_MAX_WIN 31

/* cv::buildOpticalFlowPyramid (the pyrDown chain inside calcOpticalFlowPyrLK): builds levels 1..L of slots
 * [slot0, slot0+n), L = min(max_level, levels OpenCV would keep for this frame size and window); *levels = L.
 * Called implicitly by the tracking entry points; a slot's pyramid is kept until the slot is uploaded again. */
int yavo_build_pyramid(yavo_ctx *ctx, int slot0, int n, int win_w, int win_h, int max_level, int *levels);
/* copies pyramid level `level` (>= 1) of a slot to the host; out is rows x cols as reported, capacity checked */
int yavo_pyramid_level(yavo_ctx *ctx, int slot, int level, uint8_t *out, size_t out_bytes, int *rows, int *cols);

/* One call of calcOpticalFlowPyrLK: tracks n points from the frame in slot_prev to the frame in slot_next.
 * prev_xy [2n] in; next_xy [2n] out (in/out with flag 4); status [n]; err [n] (may be NULL). */
int yavo_klt_track(yavo_ctx *ctx, int slot_prev, int slot_next, const float *prev_xy, int n, float *next_xy,
                   uint8_t *status, float *err, int win_w, int win_h, int max_level, int crit_type, int max_count,
                   double epsilon, int flags, double min_eig_threshold);

/* Batch form over consecutive slots: for every f in [slot0, slot0+n-1) the keypoints yavo_frontend_batch (or
 * yavo_fast_detect) left in slot f — the reference's features, as (x, y) = (col, row) — are tracked into slot
 * f+1.  Results stay on the device until yavo_klt_fetch.  flags may not contain OPTFLOW_USE_INITIAL_FLOW. */
int yavo_klt_track_batch(yavo_ctx *ctx, int slot0, int n, int win_w, int win_h, int max_level, int crit_type,
                         int max_count, double epsilon, int flags, double min_eig_threshold);
/* rows f of the outputs ([n x max_kp x 2], [n x max_kp], [n x max_kp]; any may be NULL) hold the tracks of slot
 * slot0+f's keypoints into slot slot0+f+1; the last row of the batch is unused. */
int yavo_klt_fetch(yavo_ctx *ctx, int slot0, int n, float *next_xy, uint8_t *status, float *err);

/* Tracking inside the streaming path (the steady state of LoopHandler::takeVOStep: FAST + BRIEF on the new frame, then
 * calcOpticalFlowPyrLK from the last frame, src/LoopHandler.cc:453-464,372-375).  Once enabled, every
 * yavo_submit_host_batch / yavo_process_host_batch on pinned frames also builds the pyramids and tracks the top-K
 * keypoints of frame f into frame f+1 for all f, f+1 of the batch, stage by stage in the same copy / compute pipeline,
 * and copies the tracks into the host arrays registered with yavo_stream_track_outputs BEFORE that submit
 * (rows f of [n x max_kp x 2] / [n x max_kp] / [n x max_kp]; the last row unused; status / err may be NULL; the arrays
 * must stay valid until the batch has been waited for).  enable = 0 switches tracking off (the default). */
int yavo_stream_tracking(yavo_ctx *ctx, int enable, int win_w, int win_h, int max_level, int crit_type, int max_count,
                         double epsilon, int flags, double min_eig_threshold);
int yavo_stream_track_outputs(yavo_ctx *ctx, float *next_xy, uint8_t *status, float *err);

/* ---- inlier count of the F-matrix RANSAC: _3DHandler::getFRANSAC (src/3DHandler.cc:145-195) ------
 * The reference fits a fundamental matrix to 8 random matches per iteration (cv::SVD; stays on the host) and counts,
 * over all matches, those with fabs(p2.t() * F * p1) < threshold, p = (pt.x, pt.y, 1) (:163-186); the matrix with the
 * first maximal count wins (:187-190).  This entry point evaluates m candidate matrices against n matches in one
 * launch: F [m x 9] row-major doubles, x1/y1/x2/y2 [n] the matches' pt1.x, pt1.y, pt2.x, pt2.y (integer keypoint
 * coordinates, as the reference reads them).  counts [m]; *best = index of the first maximum, *best_count its count
 * (m == 0: -1, INT_MIN, the reference's initial value); residuals [m x n] optional (NULL to skip).  Doubles are
 * summed in OpenCV's order: results are bit-identical to cv::gemm (tests/golden/epipolar_golden.npz). */
int yavo_epipolar_inliers(yavo_ctx *ctx, const double *F, int m, const int32_t *x1, const int32_t *y1,
                          const int32_t *x2, const int32_t *y2, int n, double threshold, int32_t *counts,
                          int32_t *best, int32_t *best_count, double *residuals);

/* Page-locked host memory for the frame and result arrays of the asynchronous entry points (yavo_submit_host_batch
 * copies to and from them while the call has already returned; with pageable result arrays the driver stages the
 * device-to-host copies and the submit blocks until they are done).  NULL on failure. */
void *yavo_pinned_alloc(size_t bytes);
void yavo_pinned_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* YAVO_B200_H */
