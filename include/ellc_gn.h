/*
 * ellc_gn.h -- C-ABI of the B200-native photometric Gauss-Newton frame-to-keyframe tracker.
 *
 * Drop-in boundary for ELLC's FLAG_DO_PARALLEL_POSE_ESTIMATION path.  The reference has no FFI of its own; the
 * seam is the C++ call surface of GetImagePoseEstimate / PixelWisePyramid / frame.  Every entry
 * point below names the reference interface it replaces (paths relative to the reference repo root).  Plain
 * pointers and sizes only -- no C++ or torch types cross this boundary.  The C++ shim that keeps the reference's
 * class names on top of this ABI is egomotion_with_local_loop_closures_b200/host/; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - pose: Lie-algebra 6-vector [wx wy wz vx vy vz]; exp(hat(pose)) maps keyframe-camera points to current-camera
 *     points (src/PixelWisePyramid.cpp:153).
 *   - images: u8, row-major, contiguous (stride == width).  depth: f32, 0 = no depth (src/Frame.cpp:298).
 *     variance: f32, -1 = invalid (src/DepthPropagation.cpp:1296).  Level l arrays are (width>>l) x (height>>l).
 *   - all functions return 0 (ELLC_OK) or a negative ellc_status; none throws or aborts.  Degenerate inputs behave
 *     like the reference (singular normal equations => zero step, src/PixelWisePyramid.cpp:451).
 *   - a handle owns one CUDA stream and all device memory; calls on one handle are serialised by the caller,
 *     different handles may be used concurrently from different host threads (main thread + loop-closure thread,
 *     src/GlobalOptimize.cpp:241).
 *   - there is NO CPU fallback: every compute entry point fails with ELLC_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef ELLC_GN_H_
#define ELLC_GN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ELLC_LEVELS 4                 /* util::MAX_PYRAMID_LEVEL, src/ExternVariable.h:40 */
#define ELLC_MAX_TRACE_ITERS 16       /* per level; the reference's largest MAX_ITER is 12 (src/main.cpp:34) */
#define ELLC_MAX_RANKS 8              /* GPUs of one NVSwitch box taking part in a result exchange                  */
#define ELLC_IPC_HANDLE_BYTES 64      /* sizeof(cudaIpcMemHandle_t)                                                  */

typedef enum ellc_status {
    ELLC_OK = 0,
    ELLC_ERR_INVALID = -1,            /* bad argument (null pointer, slot out of range, unsupported size) */
    ELLC_ERR_CUDA = -2,               /* CUDA runtime error or no usable device; see ellc_last_error_string */
    ELLC_ERR_NOT_READY = -3           /* slot used before it was uploaded / prepared */
} ellc_status;

/* arithmetic mode of the per-pixel kernel */
#define ELLC_ARITH_FAST   0           /* fp32 with FMA contraction and reciprocal multiplies (default)            */
#define ELLC_ARITH_STRICT 1           /* the reference's exact fp32/double operation sequence, no contraction      */

/* pair flags */
#define ELLC_PAIR_DEFAULT        0
#define ELLC_PAIR_CONST_WEIGHT   1    /* loop-closure pair: inverse-compositional constant-weight variant
                                         (src/PixelWisePyramid.cpp:561-974, selected at src/ImageFunc.cpp:241-244) */
#define ELLC_PAIR_SAVE_WEIGHTS   2    /* accumulate last-iteration weights into the keyframe's weight pyramid
                                         (saveWeights(true), src/ImageFunc.cpp:280-288)                           */

/* Runtime form of the compile-time configuration surface of src/ExternVariable.h and src/main.cpp:34. */
typedef struct ellc_config {
    int32_t width, height;            /* util::ORIG_COLS / ORIG_ROWS (level-0 working size), :50-51              */
    float   fx, fy, cx, cy;           /* util::ORIG_FX / ORIG_FY / ORIG_CX / ORIG_CY, :53-59                     */
    int32_t max_iter[ELLC_LEVELS];    /* util::MAX_ITER[level], src/main.cpp:34 = {4,7,9,12}                     */
    float   huber_d;                  /* util::HUBER_D, :149                                                      */
    float   camera_pixel_noise_2;     /* util::CAMERA_PIXEL_NOISE_2, :148                                         */
    float   weight[6];                /* util::weight[], :76                                                      */
    float   stop_threshold;           /* 1.0f, src/ImageFunc.cpp:251                                              */
    int32_t arithmetic;               /* ELLC_ARITH_*                                                             */
    int32_t jacobian_at_warped;       /* 0: PixelWisePyramid.cpp Jacobian (keyframe pixel/depth);
                                         1: Pyramid.cpp:99-130 variant (warped pixel, transformed depth, weight of an
                                            out-of-bounds pixel not zeroed :629-651); needs ELLC_ARITH_STRICT          */
    int32_t max_keyframes;            /* keyframe slots resident on the device                                    */
    int32_t max_frames;               /* frame slots resident on the device                                       */
    int32_t ctas_per_pair;            /* thread-block cluster size per pair: 1,2,4,8; 0 = choose from batch size  */
    int32_t device;                   /* CUDA device ordinal                                                      */
    int32_t pairs_per_cta;            /* pairs one CTA tracks in lockstep (their serial solves overlap): 1..4;
                                         0 = choose from batch size.  Only used when ctas_per_pair resolves to 1  */
    /* Levenberg-Marquardt damping of the on-device 6x6 solve (north_star).  The reference takes the plain Gauss-Newton step
       unconditionally (hessianInv = hessian.inv(); updatePose(); src/PixelWisePyramid.cpp:451-453): lm_lambda = 0, the default,
       is exactly that, bit for bit.  lm_lambda > 0: solve (H + lambda diag(H)) delta = -b; a step after which the mean weighted
       squared residual rises is rejected on the device (pose restored, lambda *= lm_up, the step retaken and counted as an
       iteration, result status bit 1), an accepted step multiplies lambda by lm_down.  Forward (Huber-reweighted) pairs only. */
    float   lm_lambda;                /* initial damping per level; 0 = off                                       */
    float   lm_up;                    /* > 1; 0 selects 4                                                         */
    float   lm_down;                  /* in (0, 1]; 0 selects 0.5                                                 */
} ellc_config;

/* One frame-keyframe pair = one call of GetImagePoseEstimate (src/ImageFunc.h:31). */
typedef struct ellc_pair {
    int32_t kf_slot;                  /* prev_frame (keyframe) slot                                               */
    int32_t frame_slot;               /* current_frame slot                                                       */
    float   init_pose[6];             /* initial relative pose, i.e. the result of src/ImageFunc.cpp:97-108       */
    int32_t flags;                    /* ELLC_PAIR_*                                                              */
} ellc_pair;

/* Fixed-size result record (256 B) -- also the unit of the multi-GPU gather. */
typedef struct ellc_result {
    float   pose[6];                  /* returned vector<float> of GetImagePoseEstimate (src/ImageFunc.cpp:311)   */
    float   H[21];                    /* upper triangle (row-major) of the last evaluated hessian                 */
    float   b[6];                     /* last evaluated sd_param                                                  */
    int32_t n_selected[ELLC_LEVELS];  /* prev_frame->no_nonZeroDepthPts per level (src/Frame.cpp:299)             */
    int32_t n_iters[ELLC_LEVELS];     /* executed GN iterations per level                                         */
    float   res_first[ELLC_LEVELS];   /* sum w r^2 over selected pixels at the level's first iteration (pose on entry) */
    float   res_last[ELLC_LEVELS];    /* same at the level's last executed iteration (before its update)         */
    float   weighted_pose[ELLC_LEVELS]; /* PixelWisePyramid::weightedPose after the level's last update          */
    int32_t n_oob[ELLC_LEVELS];       /* selected pixels warped fully out of bounds at the last iteration         */
    int32_t status;                   /* 0 ok; bit0: a singular hessian was met (zero step, as the reference);
                                         bit1: a Levenberg-Marquardt step was rejected (lm_lambda > 0 only)       */
    int32_t reserved[6];
} ellc_result;                        /* 64 x 4 B */

/* Per-iteration trace (parity tests / debugging); layout mirrors the members of class PixelWisePyramid. */
typedef struct ellc_iter_trace {
    float   H[36];                    /* PixelWisePyramid::hessian                                                */
    float   b[6];                     /* sd_param                                                                 */
    float   delta[6];                 /* deltapose                                                                */
    float   weighted_pose;            /* weightedPose                                                             */
    float   pose_after[6];            /* *pose after updatePose()                                                 */
    float   res_sum;                  /* sum w r^2 at the pose before the update                                  */
    float   weight_sum;               /* sum w                                                                    */
    int32_t n_oob;
    int32_t executed;                 /* 1 if this iteration ran                                                  */
    float   lm_lambda;                /* damping used by this iteration's solve (0 = plain Gauss-Newton)          */
    int32_t lm_rejected;              /* 1: the residual rose, the previous linearisation point was restored      */
    int32_t pad[3];
} ellc_iter_trace;                    /* 64 x 4 B */

/* Loop-closure candidate gating (globalOptimize::findMatch, src/GlobalOptimize.cpp:274-452). */
typedef struct ellc_lc_candidate {
    int32_t loop_frame_slot;          /* loopFrameArray[i]: a frame slot whose histogram was computed                */
    int32_t test_frame_slot;          /* currentLoopFrame                                                            */
    float   loop_pose_world[6];       /* loopFrameArray[i].poseWrtWorld                                              */
    float   test_pose_world[6];       /* currentLoopFrame.poseWrtWorld                                               */
} ellc_lc_candidate;
typedef struct ellc_lc_stats {
    double  match_value;              /* compareHist(loop, test, CV_COMP_KL_DIV), :351                                */
    float   rms_error;                /* calculateRotationStats, :426                                                 */
    float   relative_view_angle;      /* degrees, :435-436                                                            */
    int32_t pass;                     /* matchValue <= threshold && angle <= max angle, :364-369                      */
    int32_t reserved;
} ellc_lc_stats;

/* Loop-closure pair-list generation on the device (globalOptimize::findMatchParallel + findMatch, src/GlobalOptimize.cpp:274-420,
 * :455-620): one entry of loopFrameArray, and one test frame whose matches are wanted. */
typedef struct ellc_lc_ring_entry {
    int32_t frame_id;                 /* loopFrameArray[i].frameId                                                    */
    int32_t is_valid;                 /* loopFrameArray[i].isValid                                                    */
    int32_t frame_slot;               /* frame slot holding its image (histogram computed by ellc_frame_histograms)   */
    int32_t kf_slot;                  /* keyframe slot holding this_frame + this_currentDepthMap: the pair's keyframe */
    float   pose_world[6];            /* loopFrameArray[i].poseWrtWorld                                               */
} ellc_lc_ring_entry;                 /* 40 B */
typedef struct ellc_lc_query {
    int32_t current_array_id;         /* currentArrayId: the walk starts one position below, :286                     */
    int32_t match_window_beg;         /* match_window_beg / match_window_end, :312-322                                */
    int32_t match_window_end;
    int32_t frame_id;                 /* testFrame->frameId (MIN_MATCH_DIFFERENCE test, :344)                         */
    int32_t frame_slot;               /* frame slot of the test frame (histogram computed)                            */
    int32_t stray;                    /* strayFlag: no pose, only the id gap gates, :356-378                          */
    float   pose_world[6];            /* currentLoopFrame.poseWrtWorld                                                */
} ellc_lc_query;                      /* 48 B */

typedef struct ellc_handle ellc_handle;

/* ---- lifetime -------------------------------------------------------------------------------------------------- */
void        ellc_default_config(ellc_config* cfg, int32_t width, int32_t height);   /* ExternVariable.h defaults  */
int         ellc_create(const ellc_config* cfg, ellc_handle** out);
int         ellc_destroy(ellc_handle* h);
const char* ellc_last_error_string(const ellc_handle* h);                           /* h may be NULL (create)     */
const char* ellc_version(void);

/* ---- keyframes and frames: upload from HOST memory ---------------------------------------------------------------
 * ellc_upload_frame    replaces frame::frame(VideoCapture) after undistort/resize: constructImagePyramids()
 *                      (src/Frame.cpp:170-182) and per-level calculateGradient() (src/Frame.cpp:185-285).
 * ellc_upload_keyframe additionally takes the depth / variance pyramids the depth module hands over
 *                      (frame::depth_pyramid[l], depthMap::depthvararrptr[l]; src/DepthPropagation.h:65-78) and runs
 *                      calculateNonZeroDepthPts() (src/Frame.cpp:295-301) for every level.
 * Both are asynchronous on the handle's COPY stream (they overlap the kernels of a batch in flight; writes to a slot that a
 * batch still reads are ordered behind it automatically); host buffers must stay valid until ellc_synchronize or until the
 * records of a batch launched after the upload have been fetched (pinned memory recommended).  Every consumer (track,
 * ellc_prepare_*, evaluate, read-back) waits for pending uploads on its own stream. */
int ellc_upload_frame(ellc_handle* h, int32_t frame_slot, const uint8_t* image);
int ellc_upload_keyframe(ellc_handle* h, int32_t kf_slot, const uint8_t* image,
                         const float* const depth[ELLC_LEVELS], const float* const var[ELLC_LEVELS]);

/* ---- device-resident inputs ----------------------------------------------------------------------------------------
 * Raw device pointers into a slot, for callers that already hold the data on the GPU (cudaMemcpyAsync D2D, another
 * kernel, ...).  After writing, call ellc_prepare_* for the touched slots.  level_offsets (ELLC_LEVELS+1 entries,
 * in elements) describe how the depth / var levels are concatenated. */
/* Keyframe from the depth module's hypotheses instead of ready-made pyramids: replaces
 * depthMap::updateDepthImage + buildInvVarDepth + mapDepthArr2Mat (src/DepthPropagation.cpp:1254-1315, :1637-1746).  valid /
 * inv_depth_smoothed / variance_smoothed are the isValid / invDepthSmoothed / varianceSmoothed members of the width x height
 * depthhypothesis array as SoA.  The 3-pixel border is invalidated as the reference does; valid_out (may be NULL) receives the
 * updated flags.  Depth / variance pyramids are built on the device (bit-identical to the reference's fp32 sequence). */
int ellc_upload_keyframe_hypotheses(ellc_handle* h, int32_t slot, const uint8_t* image, const uint8_t* valid,
                                    const float* inv_depth_smoothed, const float* variance_smoothed, uint8_t* valid_out);
/* depthMap::calculate_no_of_Seeds (src/DepthPropagation.cpp:1804-1830) on the flags passed to the last
 * ellc_upload_keyframe_hypotheses of this slot: the depthMapOccupancy column of poses_orig.txt (src/main.cpp:361,373). */
int ellc_read_keyframe_occupancy(ellc_handle* h, int32_t slot, int32_t* n_valid, float* occupancy);
/* depth_pyramid[level] (0 = invalid) and depthvararrptr[level] (-1 = invalid) as resident on the device. */
int ellc_read_keyframe_depth(ellc_handle* h, int32_t slot, int32_t level, float* depth, float* var);

/* Loop-closure candidate gating.  ellc_frame_histograms: calculateImageHistogram (src/GlobalOptimize.cpp:40-100)
 * for n uploaded frame slots; the histograms stay resident (hist, n x 256 floats, may be NULL).  ellc_lc_gate: the statistics and
 * the test findMatch applies to one (loop frame, test frame) candidate (:351-369), for n candidates at once; the id-gap test
 * (MIN_MATCH_DIFFERENCE, :346) and the ring / window walk stay with the caller, who takes the first passing candidate. */
int ellc_frame_histograms(ellc_handle* h, int32_t n, const int32_t* frame_slots, float* hist);
int ellc_lc_gate(ellc_handle* h, int32_t n, const ellc_lc_candidate* candidates, float match_threshold, float max_rel_view_angle,
                 ellc_lc_stats* stats);

/* The ring / window walk of findMatch for n_queries test frames at once, on the device: for every query the ring positions are
 * visited as the reference does (start below currentArrayId, step down with wrap-around, stop at the first position outside the
 * match window or at an invalid entry), candidates need frameId gap > min_match_difference (util::MIN_MATCH_DIFFERENCE = 8), and
 * those passing the histogram / view-angle test (ellc_lc_gate's statistics) become ellc_pair records -- keyframe = the loop frame's
 * kf_slot, frame = the test frame's slot, flags = pair_flags, init_pose = log(exp(test pose) exp(loop pose)^-1) as
 * GetImagePoseEstimate derives it for a loop-closure call (src/GlobalOptimize.cpp:560-568, src/ImageFunc.cpp:97-108) -- compacted
 * query-major, walk order inside a query.  The list stays on the device for ellc_track_generated_pairs; *n_pairs and, if not NULL,
 * pairs / stats / query_of_pair (capacity n_queries * ring_len each; stats[k].reserved = ring position of the matched loop frame)
 * are copied to the host.  ring_len <= 64 (the reference: MAX_LOOP_ARRAY_LENGTH_SCALE_AVG = 43). */
int ellc_lc_generate_pairs(ellc_handle* h, int32_t ring_len, const ellc_lc_ring_entry* ring, int32_t n_queries, const ellc_lc_query* queries,
                           int32_t min_match_difference, float match_threshold, float max_rel_view_angle, int32_t pair_flags,
                           int32_t* n_pairs, ellc_pair* pairs, ellc_lc_stats* stats, int32_t* query_of_pair);
/* Track the list the last ellc_lc_generate_pairs left on the device (n_pairs as it reported) without sending it back: the pair
 * records never leave the GPU.  results: HOST array of n_pairs records. */
int ellc_track_generated_pairs(ellc_handle* h, int32_t n_pairs, ellc_result* results);

int ellc_frame_image_devptr(ellc_handle* h, int32_t frame_slot, uint8_t** image);
int ellc_keyframe_devptrs(ellc_handle* h, int32_t kf_slot, uint8_t** image, float** depth, float** var,
                          int64_t level_offsets[ELLC_LEVELS + 1]);
int ellc_prepare_frames(ellc_handle* h, int32_t n, const int32_t* frame_slots);      /* pyramid + gradients, batched */
int ellc_prepare_keyframes(ellc_handle* h, int32_t n, const int32_t* kf_slots);      /* pyramid + mask/count/selection */
/* (extension) The same two preparations on a separate low-priority stream that waits only for pending uploads and for the last
 * batch that READ these slots -- not for the batch that is tracking now.  A caller that alternates between two sets of slots
 * builds the pyramids / texels / selection lists of batch k+1 while batch k tracks.  The caller must not pass slots that the
 * batch in flight (or any other pending call) uses. */
int ellc_prepare_async(ellc_handle* h, int32_t n_frames, const int32_t* frame_slots, int32_t n_keyframes, const int32_t* kf_slots);

/* ---- the hot path -------------------------------------------------------------------------------------------------
 * ellc_track_batch replaces n calls of GetImagePoseEstimate(prev_frame, current_frame, ...) (src/ImageFunc.cpp:49-315)
 * with FLAG_DO_PARALLEL_POSE_ESTIMATION: coarse-to-fine levels 3..0, per level up to max_iter[level] iterations of
 * calculatePixelWiseParallel() (src/PixelWisePyramid.cpp:416-455) = per-pixel warp / sample / residual / weight /
 * Jacobian / 6x6 accumulate, hessian.inv(), updatePose(); early-out when weightedPose < stop_threshold
 * (src/ImageFunc.cpp:251-252).  No host round trip per iteration.  `pairs` and `results` are HOST arrays; the call
 * returns after the results have been copied back.  `trace` may be NULL; otherwise it receives
 * n * ELLC_LEVELS * ELLC_MAX_TRACE_ITERS records indexed [pair][level][iter]. */
int ellc_track_batch(ellc_handle* h, int32_t n, const ellc_pair* pairs, ellc_result* results, ellc_iter_trace* trace);

/* Same, asynchronous: only enqueues.  Results stay on the device in a ring of four buffers: the returned pointer
 * stays valid until three more track calls have been made on this handle (the buffers grow when a larger batch arrives -- all idle
 * ones together, once per batch size; a buffer whose records have not been fetched yet is left alone).  Consecutive forward batches
 * run on two internal streams and overlap at their tails.  Uploads run on their own copy stream, so the uploads of
 * the next batch overlap this batch's kernels as long as they go to slots this batch does not read (slot reuse is
 * detected and ordered automatically). */
int ellc_track_batch_async(ellc_handle* h, int32_t n, const ellc_pair* pairs, const ellc_result** device_results);
/* Wait for the batch that produced `device_results` and copy its n records to the host (own D2H stream: does not queue
 * behind batches enqueued later). */
int ellc_results_download(ellc_handle* h, const ellc_result* device_results, int32_t n, ellc_result* results);
int ellc_synchronize(ellc_handle* h);
/* Batches run on two internal tracking streams (consecutive batches overlap at their tails).  ellc_fence makes the stream
 * returned by ellc_stream() wait for every batch enqueued so far, so that an event recorded there afterwards covers them. */
int ellc_fence(ellc_handle* h);

/* ---- multi-GPU: sharded batches with an in-kernel result exchange over NVLink peer memory -----------------------
 * Frame-keyframe pairs are independent (each is one call of GetImagePoseEstimate, src/ImageFunc.cpp:49-315; the loop-closure thread
 * issues them per keyframe, src/GlobalOptimize.cpp:480-610), so a pair list is sharded over the GPUs of one box with no data-path
 * collective; the only exchange is the gather of the 256-byte ellc_result records (pose, hessian, counters) the host writes to
 * poses_orig.txt / matchframes*.txt (src/main.cpp:373,382).  Here the gather is part of the tracking kernel: every rank owns a ring
 * of result tables in its device memory, maps the tables of its peers (CUDA IPC between processes, peer access inside one process),
 * and the kernel's epilogue stores each record at its GLOBAL pair index straight into the table of every receiving rank -- no NCCL
 * call, no staging copy; a one-warp kernel behind it bumps the receivers' arrival counters.
 *   ellc_exchange_create        allocate this rank's tables (`capacity` records each) and return the IPC handle of the block
 *   ellc_exchange_attach_ipc    map the blocks of all ranks from their IPC handles (world x ELLC_IPC_HANDLE_BYTES bytes, rank order;
 *                               the handles travel by any host channel: MPI, torch.distributed, a pipe, ...)
 *   ellc_exchange_attach_local  the same for `world` handles living in THIS process (one host thread per GPU, or several handles on
 *                               one GPU): peers[d] is the handle of rank d
 *   ellc_track_batch_exchange   ellc_track_batch_async + exchange: global_index[i] in [0, n_total) is the position of pairs[i] in the
 *                               global pair list; root = -1: every rank receives all records (all-gather), root >= 0: only that rank
 *                               (gather).  Every rank calls it once per global batch with the same n_total and root; *token numbers
 *                               the batch (the same on every rank).
 *   ellc_exchange_wait          wait for this rank's batch and -- on a receiving rank -- for the records of ALL ranks, and copy the
 *                               n_total records (global order) to `results` (HOST; may be NULL).  Every rank calls it for every token,
 *                               in order; at most 3 tokens may be outstanding. */
int ellc_exchange_create(ellc_handle* h, int32_t rank, int32_t world, int32_t capacity, uint8_t ipc_handle[ELLC_IPC_HANDLE_BYTES]);
int ellc_exchange_attach_ipc(ellc_handle* h, const uint8_t* ipc_handles);
int ellc_exchange_attach_local(ellc_handle* h, ellc_handle* const* peers);
int ellc_track_batch_exchange(ellc_handle* h, int32_t n, const ellc_pair* pairs, const int32_t* global_index, int32_t n_total,
                              int32_t root, int64_t* token);
int ellc_exchange_wait(ellc_handle* h, int64_t token, ellc_result* results);
int ellc_exchange_destroy(ellc_handle* h);

/* ---- constant-weight loop-closure variant (src/PixelWisePyramid.cpp:500-974) --------------------------------------------------
 * Reference flow with util::FLAG_DO_CONST_WEIGHT_POSE_ESTIMATION: every sequential track of a frame on its keyframe ends each
 * level with saveWeights(true) (src/ImageFunc.cpp:280-288), which adds the level's last display_weightimg to the keyframe's
 * weight_pyramid; when the keyframe is retired, frame::finaliseWeights (src/Frame.cpp:678-695, src/main.cpp:431-434) averages them;
 * loop-closure pairs on that keyframe then run calculatePixelWiseParallelInvCompositional (src/ImageFunc.cpp:241-244).
 * Here: flag the sequential pair ELLC_PAIR_SAVE_WEIGHTS (its weight images stay in the FRAME slot), call
 * ellc_accumulate_weights with the frame slots in tracking order, ellc_finalise_weights, ellc_prepare_keyframes_lc, and flag the
 * loop-closure pairs ELLC_PAIR_CONST_WEIGHT.  A keyframe upload invalidates the loop-closure records (not the weights). */
int ellc_reset_keyframe_weights(ellc_handle* h, int32_t kf_slot);                /* Mat::zeros / numWeightsAdded = 0, src/Frame.cpp:114-122 */
int ellc_accumulate_weights(ellc_handle* h, int32_t kf_slot, int32_t n, const int32_t* frame_slots);   /* saveWeights(true), :546-548;
                                                                                     call before the keyframe's depth is replaced */
int ellc_finalise_weights(ellc_handle* h, int32_t kf_slot);                      /* frame::finaliseWeights */
/* Direct access to prev_frame->weight_pyramid[level] (cols x rows f32) / numWeightsAdded[level]. */
int ellc_upload_keyframe_weights(ellc_handle* h, int32_t kf_slot, const float* const weight[ELLC_LEVELS], const int32_t counts[ELLC_LEVELS]);
int ellc_read_keyframe_weights(ellc_handle* h, int32_t kf_slot, int32_t level, float* weight, int32_t* count);
/* display_weightimg of the frame's last ELLC_PAIR_SAVE_WEIGHTS track; only pixels selected in its keyframe are written. */
int ellc_read_frame_weights(ellc_handle* h, int32_t frame_slot, int32_t level, float* weight);
/* precomputePixelWiseInvCompositional + hessian (:561-680, :938) for the keyframes' current depth and weights. */
int ellc_prepare_keyframes_lc(ellc_handle* h, int32_t n, const int32_t* kf_slots);
/* The same on the pipelined preparation path of ellc_prepare_async: enqueued on the low-priority preparation stream, behind the weights
 * written so far and behind the last batch that READ these keyframe slots only -- a caller that alternates between two sets of
 * keyframe slots builds the loop-closure records of batch k+1 while batch k tracks.  The slots must be prepared
 * (ellc_prepare_async / ellc_prepare_keyframes) and must not be in use by the batch in flight. */
int ellc_prepare_keyframes_lc_async(ellc_handle* h, int32_t n, const int32_t* kf_slots);

/* One evaluation of the normal equations at a given pose and level WITHOUT updating the pose: the body of
 * calculatePixelWiseParallel() up to src/PixelWisePyramid.cpp:442.  out->delta/pose_after/weighted_pose are left 0.
 * weight_image (host, (width>>level)*(height>>level) f32, display_weightimg :361) may be NULL. */
int ellc_gn_evaluate(ellc_handle* h, int32_t kf_slot, int32_t frame_slot, int32_t level, const float pose[6],
                     ellc_iter_trace* out, float* weight_image);

/* One iteration of any of the reference's three per-level drivers at a given pose, for callers that run the iteration loop
 * themselves as src/ImageFunc.cpp:163-253 does:
 *   ELLC_VARIANT_FORWARD       PixelWisePyramid::calculatePixelWiseParallel()                      (src/PixelWisePyramid.cpp:416-455)
 *   ELLC_VARIANT_CONST_WEIGHT  PixelWisePyramid::calculatePixelWiseParallelInvCompositional(iter)   (:917-974; the precomputation of
 *                              `iter == 0` -- steepest-descent rows, hessian, hessianInv -- is what ellc_prepare_keyframes_lc built)
 *   ELLC_VARIANT_PYRAMID       Pyramid::performIterationSteps()                                      (src/Pyramid.cpp:714-726: Jacobian at
 *                              the warped pixel; always the bit-faithful STRICT arithmetic)
 * update = 0: normal equations only (out->delta / pose_after / weighted_pose stay 0); update = 1: hessian.inv() + updatePose() too.
 * weight_image (display_weightimg) and the display planes (display_warpedimg, display_iterationres, savedWarpedPointsX / Y: 0 resp.
 * -2 where the keyframe has no depth, -1 coordinates where the warp left the image, :207-283) are host arrays of (width >> level) x
 * (height >> level) floats, any of them NULL; asking for display planes evaluates with the STRICT arithmetic. */
#define ELLC_VARIANT_FORWARD      0
#define ELLC_VARIANT_CONST_WEIGHT 1
#define ELLC_VARIANT_PYRAMID      2
typedef struct ellc_display_planes {
    float* warped_image;              /* PixelWisePyramid::display_warpedimg                                        */
    float* iteration_residual;        /* display_iterationres                                                       */
    float* warped_x;                  /* savedWarpedPointsX                                                         */
    float* warped_y;                  /* savedWarpedPointsY                                                         */
} ellc_display_planes;
int ellc_gn_iterate(ellc_handle* h, int32_t kf_slot, int32_t frame_slot, int32_t level, int32_t variant, int32_t update,
                    const float pose[6], ellc_iter_trace* out, float* weight_image, const ellc_display_planes* display);
/* hessianInv = hessian.inv() (src/PixelWisePyramid.cpp:451, :939; cv::Mat::inv, DECOMP_LU) by the device's LU; all zeros and
 * *regular = 0 for a singular matrix, as OpenCV returns. */
int ellc_hessian_inverse(ellc_handle* h, const float H[36], float Hinv[36], int32_t* regular);

/* hessian.inv() + updatePose() (src/PixelWisePyramid.cpp:451-491) executed by the device code path of the handle's flavour
 * (STRICT: Pade exponential / exact logarithm; FAST: closed-form exp / log below 11.5 degrees of rotation). */
int ellc_solve_update(ellc_handle* h, const float H[36], const float b[6], const float pose_in[6],
                      float pose_out[6], float delta[6], float* weighted_pose);
/* The same, also returning rows 0..2 of exp(hat(pose_out)) as K5 hands them to the next iteration (SE3_vec,
 * src/PixelWisePyramid.cpp:153-173); equals ellc_se3_exp(pose_out) bit for bit. */
int ellc_solve_update_rt(ellc_handle* h, const float H[36], const float b[6], const float pose_in[6],
                         float pose_out[6], float delta[6], float* weighted_pose, float rt_out[12]);

/* ---- read-back of intermediate products (parity tests) ------------------------------------------------------------ */
/* image_pyramid[level] (pyrDown chain) and gradientx/gradienty after updationOnPyrChange(level) (src/Frame.cpp:316-327).
 * image is (w_l x h_l) with w_l = ceil-halved width (cv::pyrDown), gradients are (width>>level) x (height>>level).
 * Any output pointer may be NULL. */
int ellc_read_frame_level(ellc_handle* h, int32_t frame_slot, int32_t level, uint8_t* image, float* gradx, float* grady);
/* frame::mask (0/255) and no_nonZeroDepthPts after updationOnPyrChange(level) on the keyframe; image as above. */
int ellc_read_keyframe_level(ellc_handle* h, int32_t kf_slot, int32_t level, uint8_t* image, uint8_t* mask, int32_t* count);
/* pyramid image dims of a level: *w = ceil-halved width, *h likewise */
int ellc_level_dims(const ellc_handle* h, int32_t level, int32_t* pyr_w, int32_t* pyr_h, int32_t* cols, int32_t* rows);

/* ---- host-side pose algebra used by the callers of the tracker (6 floats, no device work) ------------------------- */
/* frame::concatenateRelativePose (src/Frame.cpp:503-530): dest = log(exp(a) exp(b)) */
void ellc_concat_relative(const float a[6], const float b[6], float dest[6]);
/* frame::concatenateOriginPose (src/Frame.cpp:534-562): dest = log(exp(a) exp(b)^-1) */
void ellc_concat_origin(const float a[6], const float b[6], float dest[6]);
/* exp(hat(pose)) as 4x4 row-major (src/Frame.cpp:443-471 calculateRandT) */
void ellc_se3_exp(const float pose[6], float T[16]);

/* The closed-form small-rotation exponential / logarithm the FAST flavour of the on-device pose update uses in place of Eigen's
 * Pade .exp() / Schur .log() (src/Frame.cpp:511-521) when every rotation is below 11.5 degrees; host code,
 * exported so that the CPU tests can pin them against the Pade / double-logarithm path.  ellc_se3_log_closed returns 1, or 0
 * (pose untouched) when the rotation is outside the small-angle range. */
void ellc_se3_exp_closed(const float pose[6], float T[16]);
int  ellc_se3_log_closed(const float T[16], float pose[6]);

/* ---- introspection for bench.py ----------------------------------------------------------------------------------- */
/* kernels launched by this handle since creation (or since the last reset) */
int64_t ellc_launch_count(const ellc_handle* h);
void    ellc_reset_launch_count(ellc_handle* h);
/* cudaStream_t of the handle, as an opaque pointer (for CUDA-event timing on the launching stream) */
void*   ellc_stream(ellc_handle* h);
/* Self-test of the GN kernel's shared-reciprocal division (the hand-run div.rn fast path that yields both X'/Z' and Y'/Z' of
 * src/PixelWisePyramid.cpp:250-251) against __fdiv_rn on n pseudo-random operand triples; mismatches[0]: quotients with a normal
 * result that differ, mismatches[1]: differing quotients below 2^-120. */
int ellc_selftest_division(ellc_handle* h, int64_t n, uint64_t seed, int64_t mismatches[2]);
/* Self-test of the GN kernel's UNZERO (src/ExternVariable.h:232, applied to Z' at src/PixelWisePyramid.cpp:246): the fast
 * flavour's three-instruction form against the macro's two comparisons on 64 special values and n pseudo-random bit patterns;
 * *mismatches = inputs whose results differ (two NaNs count as equal). */
int ellc_selftest_unzero(ellc_handle* h, int64_t n, uint64_t seed, int64_t* mismatches);
/* which: 0 = main compute stream (same as ellc_stream), 1 = H2D upload stream, 2 = D2H result stream, 3 / 4 = the two tracking streams (diagnostics). */
void*   ellc_stream_of(ellc_handle* h, int32_t which);
/* device time of the track kernel(s) of the most recent ellc_track_batch* call, in milliseconds (CUDA events) */
float   ellc_last_track_kernel_ms(ellc_handle* h);
/* the same for an earlier batch: batches_ago = 0 is the most recent ellc_track_batch* call, up to 3.  Measured from the moment the
 * batch's stream reaches the kernel to its completion: when batches are pipelined it INCLUDES the time the kernel's CTAs wait for
 * the previous batch to leave the SMs. */
float   ellc_batch_kernel_ms(ellc_handle* h, int32_t batches_ago);
/* completion-to-completion interval between that batch's tracking kernel and the previous batch's: in a pipelined loop the average
 * time a launch occupies the GPU (an upper bound of its exclusive kernel time) */
float   ellc_batch_interval_ms(ellc_handle* h, int32_t batches_ago);

#ifdef __cplusplus
}
#endif
#endif /* ELLC_GN_H_ */
