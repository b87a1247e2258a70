/*
 * bugcar_b200.h -- C ABI of the B200-native perception hot path
 *                  (camera frame -> ENet -> class argmax -> BEV warp -> occupancy grid).
 *
 * The reference (tranqkhue/bugcar_image_segmentation) is pure Python and exposes no
 * FFI; its boundary for this path is the Python call surface cited per entry point
 * below (paths are relative to the reference root).  The drop-in Python modules in
 * bugcar_image_segmentation_b200/{models,bev}.py bind these symbols with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success and a negative bc_status on failure; it
 *     never throws, aborts or prints.  bc_last_error() gives the message.
 *   - one context per thread and per GPU; a context is not thread-safe.
 *   - "d_" pointers are DEVICE pointers on the context's GPU, "h_" pointers are HOST
 *     pointers.  The caller owns every buffer.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     All work is enqueued asynchronously on it; no hidden host synchronisation,
 *     except in the *_host convenience calls, which return after the result has
 *     landed in the host buffer.
 *   - image layouts follow the reference: frames are uint8 HWC BGR (OpenCV), network
 *     input is NCHW float (models.py:92-94), logits are NCHW fp32 (models.py:52),
 *     label maps are uint8 (B,256,512) (models.py:67), grids are int8 (B,Hc,Wc)
 *     row-major (bev.py:244-246).
 */
#ifndef BUGCAR_B200_H
#define BUGCAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bc_ctx bc_ctx;

enum bc_status {
  BC_OK = 0,
  BC_ERR_ARG = -1,      /* bad argument / shape precondition (reference: AssertionError, bev.py:169) */
  BC_ERR_STATE = -2,    /* weights or calibration not loaded yet */
  BC_ERR_CUDA = -3,     /* CUDA runtime error, message in bc_last_error */
  BC_ERR_FORMAT = -4,   /* malformed weight container / missing tensor (reference: TF import error) */
  BC_ERR_NOMEM = -5
};

/* network geometry fixed by the reference: models.py:19 INPUT_WIDTH, INPUT_HEIGHT = (512, 256) */
#define BC_NET_W 512
#define BC_NET_H 256

/* input kinds for the ENet entry points */
enum bc_input_kind {
  BC_IN_BGR_U8 = 0,     /* uint8 (B,256,512,3) BGR frames; normalisation of models.py:91 fused in */
  BC_IN_NCHW_F32 = 1,   /* float  (B,3,256,512), output of ENET.preprocess cast to fp32 (TF feed) */
  BC_IN_NCHW_F64 = 2    /* double (B,3,256,512), output of ENET.preprocess as is (models.py:95) */
};

/* storage precision of activations / GEMM operands */
enum bc_precision {
  BC_PREC_BF16 = 0,     /* bf16 storage, fp32 accumulate, tcgen05 tiles (8 significand bits) */
  BC_PREC_FP32 = 1,     /* exact mode: fp32 storage and arithmetic on CUDA cores */
  BC_PREC_FP16 = 2      /* production (default): IEEE fp16 storage (11 significand bits, stores saturate at
                           +-65504), fp32 accumulate, the same tcgen05 tiles; >= 99.9 % raw argmax agreement
                           with the fp32 network, which bf16 storage misses (99.7 %) */
};

/* ---- lifetime ------------------------------------------------------------------- */
/* Replaces ENET.__init__'s session creation (models.py:21-22).  max_batch bounds B of
 * every later call. */
int bc_create(bc_ctx** out, int device, int max_batch);
void bc_destroy(bc_ctx* ctx);
/* Message of the last failure on this context (ctx may be NULL: last create error). */
const char* bc_last_error(const bc_ctx* ctx);
/* Library/ABI version, for the binding to check. */
int bc_abi_version(void);

/* ---- model + calibration ---------------------------------------------------------- */
/* Replaces the GraphDef import of models.py:23-31.  `h_blob` is a BCENETW1 flat weight
 * container (bugcar_image_segmentation_b200/weights.py).  Folds batch norm, repacks. */
int bc_load_enet(bc_ctx* ctx, const void* h_blob, size_t n_bytes);
int bc_num_classes(const bc_ctx* ctx);
int bc_set_precision(bc_ctx* ctx, int precision /* enum bc_precision */);
/* Frames of a batch are pushed through the network `frames` at a time (smaller chunks keep
 * the inter-layer activations resident in the 126 MB L2, larger ones amortise launches).
 * 0 restores the default: the whole batch in one pass. */
int bc_set_chunk(bc_ctx* ctx, int frames);
/* 1 (default): the network runs as fused tcgen05 kernels in the 16-bit modes;
 * 0: every layer runs on the CUDA-core kernels (bring-up / A-B comparison). */
int bc_set_tensor_cores(bc_ctx* ctx, int enable);
/* 1 (default): bc_pipeline / bc_pipeline_host replay their kernel chain as one CUDA graph
 * per distinct argument set (what makes batch 1 sub-millisecond); 0: plain launches. */
int bc_set_graphs(bc_ctx* ctx, int enable);
/* 1 (default): bc_pipeline_host splits batches >= 32 into sub-batches whose H2D copy (internal
 * copy stream) overlaps the previous sub-batch's kernels; 0: one copy, one pass. */
int bc_set_host_overlap(bc_ctx* ctx, int enable);
/* Replaces bev_transform_tools.__init__/fromJSON state (bev.py:13-41): src->dst
 * homography `h_M` (row-major 3x3, bev.py:31-32), label-map shape (rows, cols)
 * ("input image size", bev.py:30,169), warped size (ww, wh) ("output image size",
 * bev.py:29) and cm_per_px (bev.py:35).  The label map must be 2..32767 pixels per side (the
 * occupancy-grid kernel samples fixed 2x2 blocks of label pixels); BC_ERR_ARG otherwise. */
int bc_set_bev(bc_ctx* ctx, const double h_M[9], int in_rows, int in_cols,
               int warp_w, int warp_h, double cm_per_px);

/* ---- ENET.preprocess  (models.py:84-95) --------------------------------------------- */
/* cv2.resize(bgr,(512,256)) bit-exact (models.py:87).  d_src (B,h,w,3) -> d_dst (B,256,512,3). */
int bc_resize_bgr(bc_ctx* ctx, const uint8_t* d_src, int h, int w, int B,
                  uint8_t* d_dst, void* stream);
/* Full preprocess: resize, BGR->RGB, (u/256 - mean)/std, HWC->CHW (models.py:87-94).
 * out_f64 = 1 writes double (B,3,256,512) exactly as the reference returns;
 * out_f64 = 0 writes float (what TensorFlow is fed). */
int bc_preprocess(bc_ctx* ctx, const uint8_t* d_bgr, int h, int w, int B,
                  void* d_out, int out_f64, void* stream);

/* ---- ENET.predict / predict_binary  (models.py:42-82) ---------------------------------- */
/* sess.run of the frozen graph (models.py:43-44): logits fp32 NCHW (B,C,256,512). */
int bc_enet_logits(bc_ctx* ctx, const void* d_x, int kind, int B, float* d_logits,
                   void* stream);
/* Parity/debug: activation after block `block` (-1 = initial block, 0.. = the 27 bottlenecks
 * in network order) as fp32 NCHW (B,C,H,W); B <= the chunk size.  No reference counterpart:
 * the frozen graph exposes only its output tensor (models.py:16). */
int bc_enet_block_output(bc_ctx* ctx, const void* d_x, int kind, int B, int block, float* d_out,
                         void* stream);
/* Forward + tf.math.argmax(axis=1) (models.py:55) + class LUT fused in the head.
 * h_lut[class] -> label; the 3-way LUT of models.py:56-58 gives predict, the
 * {0,1}->1 LUT of models.py:79-80 gives predict_binary.  d_labels uint8 (B,256,512). */
int bc_enet_labels(bc_ctx* ctx, const void* d_x, int kind, int B, const uint8_t h_lut[256],
                   uint8_t* d_labels, void* stream);
/* Stand-alone argmax + LUT on caller-provided logits (models.py:55-58,67 / 78-81). */
int bc_argmax_lut(bc_ctx* ctx, const float* d_logits, int B, int C, int H, int W,
                  const uint8_t h_lut[256], uint8_t* d_labels, void* stream);

/* ---- bev_transform_tools.create_occupancy_grid[_binary]  (bev.py:97-246) ---------------- */
/* Grid shape for the given request (bev.py:172-176): returns 0 and fills Hc, Wc. */
int bc_occgrid_shape(bc_ctx* ctx, double w_m, double h_m, double cell_m, int* Hc, int* Wc);
/* d_labels uint8 (B,in_rows,in_cols) -> d_grids int8 (B,Hc,Wc); non-laserscan mode.
 * binary = 0: bev.py:166-246 (occupied = {1,3});  binary = 1: bev.py:97-144,165.
 * ros_layout = 1 additionally applies occgrid_to_ros.py:18-21 (flip + rotate 90 CCW),
 * i.e. writes int8 (B,Wc,Hc) ready for OccupancyGrid.data. */
int bc_occgrid(bc_ctx* ctx, const uint8_t* d_labels, int B, double w_m, double h_m,
               double cell_m, int binary, int ros_layout, int8_t* d_grids, void* stream);

/* Laserscan-like grids (the `is_laserscan` branch, bev.py:145-164 binary / bev.py:216-240 three-way): only
 * the first obstacle along each ray from the camera (grid centre (Wc/2 - 1, Hc)) is kept, behind it
 * the grid is unknown.  The reference runs its two cv2.warpPolar calls without WARP_FILL_OUTLIERS and
 * so reads uninitialised memory wherever a ray leaves the grid; here those pixels are 0, i.e. the
 * result equals the reference's with that flag set.  binary = 1: d_grid_plain (may be NULL) receives
 * the ordinary binary grid and d_grid_laser the laserscan one -- the reference's 2-tuple (bev.py:164);
 * binary = 0: d_grid_plain is ignored.  Grids are int8 (B,Hc,Wc).  Allocates on first use of a grid
 * shape / a larger batch: call once outside stream capture. */
int bc_occgrid_laserscan(bc_ctx* ctx, const uint8_t* d_labels, int B, double w_m, double h_m,
                         double cell_m, int binary, int8_t* d_grid_plain, int8_t* d_grid_laser,
                         void* stream);

/* Host-only (no GPU, no context): the two gather tables bc_occgrid_laserscan builds for a (Wc,Hc) grid,
 * i.e. the coordinate maps of cv2.warpPolar(..., WARP_POLAR_LINEAR) and (..., WARP_INVERSE_MAP) after
 * nearest-neighbour rounding.  h_fwd [pol_h][pol_w] -> flat grid cell or -1, h_inv [Hc][Wc] -> flat polar
 * pixel or -1; pass both NULL to query the polar size.  For tests and for hosts that post-process grids. */
int bc_laser_tables(int Wc, int Hc, int binary, int* pol_w, int* pol_h, int* h_fwd, int* h_inv);

/* ---- contour_noise_removal  (image_processing_utils.py:4-44) ----------------------------- */
/* d_seg uint8 (B,H,W) road masks (non-zero = road, as ENET.predict_binary returns them) ->
 * d_out uint8 (B,H,W) in {0,1}: k x k close with k = int(min(H,W)/50) (:6-9), then every contour
 * (findContours RETR_LIST, :12) whose fillPoly covers more than 40 % of the bottom tenth of the
 * image (:19-39) is kept, and all kept contours are filled together, even-odd (:41-42).  Computed
 * by connected-component labelling, bit-exact with the OpenCV formulation.  Needs
 * 50 <= min(H,W) < 1650.  Scratch grows on demand (a first call must happen outside graph capture). */
int bc_contour_noise_removal(bc_ctx* ctx, const uint8_t* d_seg, int H, int W, int B, uint8_t* d_out,
                             void* stream);
/* 1: bc_pipeline* calls with binary = 1 run predict_binary -> contour_noise_removal -> grid, the
 * order the reference intends (models.py:6 imports the filter next to predict_binary);
 * 0 (default): the mask goes to the grid unfiltered. */
int bc_set_contour_filter(bc_ctx* ctx, int enable);

/* ---- whole path --------------------------------------------------------------------- */
/* frames (B,h,w,3) uint8 BGR -> grids: resize (if needed) -> ENet -> argmax+LUT -> grid.
 * d_labels_out may be NULL (labels then live only in context scratch). */
int bc_pipeline(bc_ctx* ctx, const uint8_t* d_bgr, int h, int w, int B,
                const uint8_t h_lut[256], double w_m, double h_m, double cell_m,
                int binary, int ros_layout, uint8_t* d_labels_out, int8_t* d_grids,
                void* stream);
/* Same with HOST buffers: H2D of the frames, the pipeline, D2H of the grids, then a
 * stream synchronise.  This is the call a per-frame driver loop makes. */
int bc_pipeline_host(bc_ctx* ctx, const uint8_t* h_bgr, int h, int w, int B,
                     const uint8_t h_lut[256], double w_m, double h_m, double cell_m,
                     int binary, int ros_layout, int8_t* h_grids, void* stream);

/* Streaming form for a driver loop that keeps the GPU busy: submit returns once the step is
 * enqueued (H2D and D2H on internal copy streams, kernels on `stream`); two staging slots
 * alternate, so the copy of step i+1 overlaps the kernels of step i.  A third submit blocks until
 * the oldest step has finished.  bc_pipeline_host_wait(ctx, k) returns when at most k (0 or 1)
 * submitted steps are still in flight; the grids of finished steps are in their h_grids buffers,
 * which (like h_bgr) must stay valid and should be pinned until then. */
int bc_pipeline_host_submit(bc_ctx* ctx, const uint8_t* h_bgr, int h, int w, int B,
                            const uint8_t h_lut[256], double w_m, double h_m, double cell_m,
                            int binary, int ros_layout, int8_t* h_grids, void* stream);
int bc_pipeline_host_wait(bc_ctx* ctx, int keep_in_flight);

/* ---- multi-GPU gather (frame-batch sharding; grids gathered to rank 0) ----------------- */
/* After this call bc_occgrid/bc_pipeline write their grids to
 * d_gather_base + rank * B * Hc * Wc instead of d_grids when d_grids == NULL.
 * d_gather_base is rank 0's buffer mapped into this process (CUDA IPC / peer access),
 * so K9's stores travel over NVLink; pass NULL to switch the redirection off. */
int bc_gather_setup(bc_ctx* ctx, void* d_gather_base, int rank, int world);

/* Streaming gather: makes bc_pipeline_host_submit / _wait the multi-GPU entry point.  One process per
 * GPU calls submit with ITS frames; the occupancy-grid kernel of every rank stores into rank 0's
 * gather buffer (as with bc_gather_setup), and the ranks synchronise through flags in peer-mapped
 * device memory -- no collective, no host barrier per step:
 *   - after its kernels rank r writes the step number into d_arrive[slot][r] (rank 0's memory);
 *   - rank 0 waits (on its internal D2H stream) for all `world` arrivals of the slot, copies the whole
 *     (world*B, Hc, Wc) buffer to ITS h_grids, then writes the step number into every rank's
 *     d_release[slot] (that rank's own memory), which that rank's next use of the slot waits for.
 * Two slots alternate, so step i+1 computes while step i's grids leave.  All pointers are device
 * pointers valid in this process (CUDA IPC mappings where the memory lives on another GPU):
 *   d_gather[2]       rank 0's two gather buffers, each world * B * Hc * Wc bytes
 *   d_arrive          rank 0's memory, uint32 [2][world], zero-initialised
 *   d_release_mine    this rank's own memory, uint32 [2], zero-initialised
 *   d_release_peers   rank 0 only: [world] pointers, entry r = rank r's d_release_mine; NULL elsewhere
 * h_grids of submit is used on rank 0 only (size world * B * Hc * Wc) and may be NULL elsewhere.  A rank
 * that waits longer than ~10 s for a flag gives up; bc_pipeline_host_wait then returns BC_ERR_STATE.
 * Pass d_gather == NULL to switch the mode off. */
int bc_gather_stream_setup(bc_ctx* ctx, void* const d_gather[2], uint32_t* d_arrive, uint32_t* d_release_mine,
                           uint32_t* const* d_release_peers, int rank, int world);

/* Page-locked host memory for the frame / grid buffers of the *_host calls (cudaHostAlloc).
 * write_combined = 1 gives write-combined pages: fast for the CPU to fill and for the GPU to
 * read over PCIe, slow for the CPU to read back -- for frame staging buffers only. */
int bc_host_alloc(void** h_ptr, size_t bytes, int write_combined);
int bc_host_free(void* h_ptr);

/* number of kernels this context launched since creation (bench.py "gpu_launches") */
long long bc_launch_count(const bc_ctx* ctx);
/* Per-kernel timing for the roofline report: while enabled every launch is bracketed by a
 * CUDA event pair on its stream (CUDA graphs are bypassed); bc_profile_json synchronises
 * and returns a JSON array [{"kernel", "launches", "ms", "bytes", "flops"}] aggregated by
 * kernel since the last bc_set_profile call ("bytes"/"flops" are the ALGORITHMIC traffic
 * and work of those launches).  The string is owned by the context. */
int bc_set_profile(bc_ctx* ctx, int enable);
const char* bc_profile_json(bc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* BUGCAR_B200_H */
