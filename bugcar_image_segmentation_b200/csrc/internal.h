// Internal declarations shared by the translation units of libbugcar_b200.so.
// Not part of the ABI (see include/bugcar_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>

namespace bc {

typedef __nv_bfloat16 bf16;
typedef __half f16;

// ------------------------------------------------------------------ packed layers
// One convolution (+ folded batch norm + activation slope) for the CUDA-core kernels.
// w is [tap][cin][cout] fp32 (values rounded to bf16 in BC_PREC_BF16 so that both the
// CUDA-core and the tcgen05 kernels see identical operands), bias/alpha are [cout].
struct ConvP {
  float* w = nullptr;
  float* bias = nullptr;
  float* alpha = nullptr;    // PReLU slope per output channel (0 = ReLU, 1 = identity)
  int cin = 0, cout = 0, ntaps = 0;
};

// tcgen05 operand pack of one bottleneck kernel launch (enet_umma.cu): the exact
// shared-memory image of the K-major swizzled bf16 weights + the fp32 bias / slope block
struct UmmaPack {
  uint8_t* wblob = nullptr;
  std::vector<float> hf;       // fp32 bias / slope block, passed by value as kernel parameters
  int C = 0, CI = 0, CN = 0, CRES = 0, ntaps = 0;   // output / internal / next-projection / residual channels
  bool has_exp = false, has_next = false;
};

struct Bottleneck {
  std::string name;
  int kind;                  // 0 down, 1 regular, 2 asymmetric, 3 up
  int cin, cout, ci, dilation;
  ConvP c1, c2, c2b, c3, cm; // proj / mid / mid-second (asym) / expand / main (up)
  float* alpha_out = nullptr;
  std::vector<float> s5;     // Stage5Params image (host) when the block is the 16-channel regular bottleneck
  UmmaPack um_f;             // asymmetric bottleneck fused into one kernel (5x1 + 1x5 + expansion + residual + next projection)
  UmmaPack um_a, um_b;       // um_b: second half (1x5 + expansion) of an asymmetric bottleneck; first half (pool + 2x2 conv) of a down-sampling one
};

// stage-5 bottleneck (16 channels, internal width 4): every parameter by value (simt_stage5.cu)
struct Stage5Params {
  float w1[64], b1[4], a1[4];      // projection [16][4]
  float w2[144], b2[4], a2[4];     // 3x3 conv [9][4][4]
  float w3[64], b3[16], a3[16];    // expansion [4][16]
  float aout[16];
};

struct Taps { int8_t dy[9]; int8_t dx[9]; };

struct Lut256 { uint8_t v[256]; };

// ------------------------------------------------------------------ BEV parameters
struct BevGeom {            // kernel parameter block for K9 (passed by value)
  double Mi[9];             // inverse homography (dst -> src), fp64 (cv::invert)
  int bw0;                  // OpenCV warp block width (coordinate association)
  int in_rows, in_cols;     // label map
  int warp_w, warp_h;
  int occ_w_px, occ_h_px;   // template size
  int wl, wt, gl, gt;       // crop / paste offsets (bev.py:183-189)
  int crop_w, crop_h;
  int Wc, Hc;               // grid cells
  int binary, ros_layout;
  int raw_template;         // 1: write the resized template value (0..3) instead of the int8 map (laserscan branch)
  double ifx, ify;          // nearest-resize source step (cv::resize INTER_NEAREST, fp64)
  const uint2* table;       // device table of k_occ_table for this geometry (uint2 [25][Hc*Wc] + uint32 [Hc*Wc]); host-side cache key has it null
};

// ------------------------------------------------------------------ launchers (enet_simt.cu)
struct Net;   // forward (api.cu)

template <typename T>
void launch_initial(const void* x, int kind, int B, T* out, int pool_kernel, const float* w, const float* g,
                    const float* b, const float* alpha, const float* lut, cudaStream_t s);
template <typename T>
void launch_down_a(const T* x, int B, int H, int W, int cin, int ci, T* pooled, uint8_t* idx,
                   T* e1, const ConvP& c1, cudaStream_t s);
template <typename T>
void launch_conv(const T* in, T* out, const T* res, int res_ch, const ConvP& c,
                 const float* alpha_out, int B, int H, int W, const Taps& taps, cudaStream_t s);
template <typename T>
void launch_up_b(const T* x, const T* e1, const uint8_t* idx, T* out, const Bottleneck& bn,
                 int B, int H, int W, cudaStream_t s);
template <typename T>
void launch_fullconv(const T* x, int B, int C, const float* w, float* logits, uint8_t* labels,
                     const Lut256* lut, cudaStream_t s);

// regular bottleneck at 16 channels / internal width 4 (regular5_1) fused into one kernel (simt_stage5.cu)
template <typename T>
void launch_stage5(const T* x, T* y, const Bottleneck& b, int B, int H, int W, cudaStream_t s);

template <typename T>
void launch_export_nchw(const T* in, float* out, int B, int C, int H, int W, cudaStream_t s);

// ------------------------------------------------------------------ tcgen05 kernels (enet_umma.cu, umma_*.cu)
// Those sources are compiled once per 16-bit activation type (bf16: BC_PREC_BF16, fp16: BC_PREC_FP16;
// see umma_common.cuh); Umma<AT> is the per-type set of entry points, declared in umma_api.inc.
extern thread_local int g_umma_reverse;                    // tile walk direction of the next tcgen05 launch (api.cu)
bool umma_supported(const Bottleneck& bn);    // regular / dilated / asymmetric at C = 64 or 128
void umma_free(UmmaPack& p);
template <typename AT> struct Umma;
template <> struct Umma<bf16> {
  typedef bf16 AT;
#include "umma_api.inc"
};
template <> struct Umma<f16> {
  typedef f16 AT;
#include "umma_api.inc"
};

// ------------------------------------------------------------------ launchers (prepost.cu)
struct ResizeTab {          // device tables for cv2.resize INTER_LINEAR, one (h,w)
  int src_h = 0, src_w = 0;
  int mode = 0;             // 0 identity, 1 area 2x2, 2 linear fixed-point
  int* x0 = nullptr; int* x1 = nullptr; int* a0 = nullptr; int* a1 = nullptr;   // [512]
  int* y0 = nullptr; int* y1 = nullptr; int* b0 = nullptr; int* b1 = nullptr;   // [256]
  void* blob = nullptr;
};
void launch_resize(const uint8_t* src, int h, int w, int B, uint8_t* dst, const ResizeTab& t,
                   cudaStream_t s);
void launch_preprocess(const uint8_t* bgr256, int B, void* out, int out_f64, const double* lut64,
                       cudaStream_t s);
void launch_argmax_lut(const float* logits, int B, int C, int H, int W, const Lut256& lut,
                       uint8_t* labels, cudaStream_t s);
size_t occ_table_bytes(int cells);
void launch_occ_table(const BevGeom& g, uint2* table, cudaStream_t s);     // occ_table_bytes(Hc*Wc) bytes, once per geometry
void launch_occgrid(const uint8_t* labels, int B, const BevGeom& g, int8_t* grids, cudaStream_t s);   // needs g.table
cudaError_t prepare_occgrid();            // dynamic shared-memory opt-in of K9 on the current device


// flag kernels of the streaming gather (prepost.cu): system-scope release store / bounded acquire spin
void launch_flag_store(uint32_t* const* ptrs, int n, uint32_t value, cudaStream_t s);       // *ptrs[i] = value, i < n (ptrs: device array)
void launch_flag_store1(uint32_t* ptr, uint32_t value, cudaStream_t s);
void launch_flag_wait(const uint32_t* flags, int n, uint32_t value, int* d_err, cudaStream_t s);   // until flags[i] >= value for all i < n

// ------------------------------------------------------------------ launchers (contour.cu)
// contour_noise_removal (image_processing_utils.py:4-44): uint8 masks (B,H,W) -> uint8 0/1 masks.
// `scratch` must hold contour_scratch_bytes(B,H,W) bytes; needs 50 <= min(H,W) < 1650.
size_t contour_scratch_bytes(int B, int H, int W);
int contour_launch_count();
void launch_contour_noise_removal(const uint8_t* seg, int H, int W, int B, uint8_t* out, void* scratch, cudaStream_t s);


// ------------------------------------------------------------------ launchers (laser.cu)
// laserscan-like grids (bev.py:145-164, 216-240): polar image size, host-built gather tables, per-batch kernels
void laser_polar_size(int Wc, int Hc, int binary, int* pol_w, int* pol_h);
void laser_build_tables(int Wc, int Hc, int pol_w, int pol_h, std::vector<int>& fwd, std::vector<int>& inv);
void launch_laser(const uint8_t* cells, int B, int Wc, int Hc, int pol_w, int pol_h, int binary, const int* d_fwd,
                  const int* d_inv, int* d_first, int8_t* out, cudaStream_t s);

}  // namespace bc
