// "Laserscan-like" occupancy grids (bev.py:145-164 binary, bev.py:216-240 three-way): keep only the
// first obstacle along every ray from the camera, everything behind it becomes unknown.
//
// The reference does it with two cv2.warpPolar calls (grid -> polar, polar -> grid) around a
// per-row "first obstacle column" search and a radius-1 cv2.circle per hit.  It passes no
// WARP_FILL_OUTLIERS, so polar pixels whose ray leaves the grid and grid cells that map outside the
// polar image are read from uninitialised memory: its output is not a function of its input.  This
// file computes the branch with those pixels defined as 0 (= the reference with that flag OR-ed in;
// oracle/laser_oracle.py, pinned against exactly that).
//
// Both warps are nearest-neighbour gathers whose coordinates depend only on the grid shape, so they
// are built ONCE per shape on the host, in the float/double arithmetic of cv::warpPolar,
// cv::cartToPolar (fastAtan32f polynomial) and cv::remap (cvRound), as two index tables.  Per
// batch: K9 writes the plain grid / the raw template, k_laser_first finds the first obstacle of
// every (frame, angle) with one warp, k_laser_mark gathers the plus-shaped marks back per cell.
#include "internal.h"

#include <cmath>
#include <vector>

namespace bc {
namespace {

// cv::fastAtan32f scalar path (mathfuncs_core.simd.hpp), degrees, fp32 without contraction
float fast_atan2_deg(float y, float x) {
  const float s = (float)(180 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s, p5 = 0.1555786518463281f * s,
              p7 = -0.04432655554792128f * s;
  const float eps = (float)2.2204460492503131e-16;
  volatile float ax = std::fabs(x), ay = std::fabs(y);      // volatile: every step rounds to fp32, no FMA
  volatile float c, c2, a;
  if (ax >= ay) {
    c = ay / (ax + eps);
    c2 = c * c;
    a = p7 * c2; a = a + p5; a = a * c2; a = a + p3; a = a * c2; a = a + p1; a = a * c;
  } else {
    c = ax / (ay + eps);
    c2 = c * c;
    a = p7 * c2; a = a + p5; a = a * c2; a = a + p3; a = a * c2; a = a + p1; a = a * c;
    a = 90.f - a;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

inline long cv_round(float v) { return std::lrint(v); }      // round half to even (default rounding mode)

__global__ void k_laser_first(const uint8_t* __restrict__ cells, const int* __restrict__ fwd, int* __restrict__ first,
                              int ncells, int pol_w, int pol_h, int target) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= pol_h) return;
  const uint8_t* frame = cells + (size_t)blockIdx.y * ncells;
  const int* f = fwd + (size_t)row * pol_w;
  int hit = -1;
  for (int base = 0; base < pol_w; base += 32) {
    int col = base + lane;
    int src = col < pol_w ? f[col] : -1;
    bool is = src >= 0 && frame[src] == target;
    unsigned m = __ballot_sync(0xffffffffu, is);
    if (m) { hit = base + __ffs(m) - 1; break; }
  }
  if (lane == 0) first[(size_t)blockIdx.y * pol_h + row] = hit;
}

// variant 0: three-way (cells = raw template 0..3 -> final int8 map), 1: binary (cells = plain int8 grid)
__global__ void k_laser_mark(const uint8_t* __restrict__ cells, const int* __restrict__ inv, const int* __restrict__ first,
                             int8_t* __restrict__ out, int ncells, int pol_w, int pol_h, int binary) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncells) return;
  const int* f = first + (size_t)blockIdx.y * pol_h;
  const int p = inv[cell];
  bool mark = false;
  if (p >= 0) {
    const int pr = p / pol_w, pc = p - pr * pol_w;
    const int f0 = f[pr];
    mark = (f0 >= 0 && abs(f0 - pc) <= 1) || (pr > 0 && f[pr - 1] == pc) || (pr + 1 < pol_h && f[pr + 1] == pc);
  }
  const int v = cells[(size_t)blockIdx.y * ncells + cell];
  int o;
  if (binary) {
    o = v == 255 ? 255 : (mark ? 100 : 0);                   // bev.py:161-163
  } else {
    int nv = v != 3 ? v : (mark ? 1 : 0);                    // bev.py:236
    o = nv == 0 ? 255 : ((200 - nv * 100) & 255);            // bev.py:244-245
  }
  out[(size_t)blockIdx.y * ncells + cell] = (int8_t)o;
}

}  // namespace

void laser_polar_size(int Wc, int Hc, int binary, int* pol_w, int* pol_h) {
  if (binary) { *pol_w = Wc; *pol_h = Hc; return; }          // bev.py:148: dsize = shape
  double R = (double)std::max(Wc, Hc);                        // bev.py:219: dsize = (-1, -1)
  *pol_w = (int)std::lrint(R);
  *pol_h = (int)std::lrint(R * 3.1415926535897932384626433832795);
}

// host: the two gather tables of one grid shape (fwd: [pol_h][pol_w] -> cell or -1; inv: [Hc][Wc] -> polar pixel or -1)
void laser_build_tables(int Wc, int Hc, int pol_w, int pol_h, std::vector<int>& fwd, std::vector<int>& inv) {
  const double two_pi = 6.283185307179586476925286766559;
  const float cx = (float)(Wc / 2.0 - 1), cy = (float)Hc;     // bev.py:148 / :219 centre, as cv::Point2f
  const double R = (double)std::max(Wc, Hc);
  fwd.assign((size_t)pol_w * pol_h, -1);
  inv.assign((size_t)Wc * Hc, -1);
  {                                                            // cv::warpPolar, forward, linear
    const double k_angle = two_pi / pol_h, k_mag = R / pol_w;
    std::vector<float> rhos(pol_w);
    for (int r = 0; r < pol_w; ++r) rhos[r] = (float)(r * k_mag);
    for (int phi = 0; phi < pol_h; ++phi) {
      const double a = k_angle * phi, cp = std::cos(a), sp = std::sin(a);
      for (int r = 0; r < pol_w; ++r) {
        volatile double xd = rhos[r] * cp, yd = rhos[r] * sp;
        const float mx = (float)(xd + cx), my = (float)(yd + cy);
        const long sx = cv_round(mx), sy = cv_round(my);
        if (sx >= 0 && sx < Wc && sy >= 0 && sy < Hc) fwd[(size_t)phi * pol_w + r] = (int)(sy * Wc + sx);
      }
    }
  }
  {                                                            // cv::warpPolar, WARP_INVERSE_MAP (source wrapped by one row)
    const double k_angle = two_pi / pol_h, k_mag = R / pol_w;
    const float deg2rad = (float)(3.14159265358979323846 / 180);
    for (int y = 0; y < Hc; ++y)
      for (int x = 0; x < Wc; ++x) {
        volatile float bx = (float)x - cx, by = (float)y - cy;
        volatile float xx = bx * bx, yy = by * by, ss = xx + yy;
        const float mag = std::sqrt(ss);
        volatile float ang = fast_atan2_deg(by, bx) * deg2rad;
        const float rho = (float)((double)mag / k_mag);
        volatile float phi = (float)((double)ang / k_angle);
        phi = phi + 1.0f;
        const long sx = cv_round(rho), sy = cv_round(phi);
        if (sx >= 0 && sx < pol_w && sy >= 0 && sy < pol_h + 2)
          inv[(size_t)y * Wc + x] = (int)(((sy - 1 + pol_h) % pol_h) * pol_w + sx);
      }
  }
}

void launch_laser(const uint8_t* cells, int B, int Wc, int Hc, int pol_w, int pol_h, int binary, const int* d_fwd,
                  const int* d_inv, int* d_first, int8_t* out, cudaStream_t s) {
  const int ncells = Wc * Hc;
  k_laser_first<<<dim3((pol_h + 3) / 4, B), 128, 0, s>>>(cells, d_fwd, d_first, ncells, pol_w, pol_h, binary ? 100 : 3);
  k_laser_mark<<<dim3((ncells + 255) / 256, B), 256, 0, s>>>(cells, d_inv, d_first, out, ncells, pol_w, pol_h, binary);
}

}  // namespace bc
