// C ABI of libbugcar_b200.so (include/bugcar_b200.h): context, weight container parsing,
// batch-norm folding, the ENet layer schedule, BEV geometry and the whole-path entry points.
//
// Reference call surface replaced (paths relative to the reference root):
//   ENET.__init__            models.py:21-31     -> bc_create + bc_load_enet
//   ENET.preprocess          models.py:84-95     -> bc_resize_bgr / bc_preprocess
//   ENET.predict[_binary]    models.py:42-82     -> bc_enet_logits / bc_enet_labels / bc_argmax_lut
//   bev_transform_tools      bev.py:13-41        -> bc_set_bev
//   create_occupancy_grid*   bev.py:97-246       -> bc_occgrid
//   the per-frame loop       README.md:18-20     -> bc_pipeline / bc_pipeline_host
#include "internal.h"
#include "../../include/bugcar_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <type_traits>

using namespace bc;

namespace {

// ---------------------------------------------------------------- host-side conv pack
struct HostConv {
  std::vector<float> w, bias, alpha;   // [tap][cin][cout], [cout], [cout]
  int cin = 0, cout = 0, ntaps = 0;
};

struct HostBlock {
  std::string name;
  int kind = 0;                        // 0 down, 1 regular, 2 asymmetric, 3 up
  int cin = 0, cout = 0, ci = 0, dilation = 1;
  HostConv c1, c2, c2b, c3, cm;
  std::vector<float> alpha_out;
};

struct Tensor {
  const float* data = nullptr;
  int ndim = 0;
  int dims[4] = {0, 0, 0, 0};
  size_t count() const {
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) n *= (size_t)dims[i];
    return n;
  }
};

struct GraphKey {
  const void* in; const void* out; const void* labels;
  int h, w, B, binary, ros, pad_;
  double w_m, h_m, cell_m;
  uint8_t lut[256];
  bool operator==(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) == 0; }
};

struct GraphEntry {
  GraphKey key;
  cudaGraphExec_t exec = nullptr;
  long long launches = 0;
};

struct ProfRec {
  const char* name;
  double bytes, flops;      // algorithmic bytes moved / flops of this launch
  cudaEvent_t start, stop;
};

// value of a weight as the 16-bit kernels of `precision` read it (fp32 mode: unchanged)
inline float round_to(int precision, float v) {
  if (precision == BC_PREC_BF16) return __bfloat162float(__float2bfloat16_rn(v));
  if (precision == BC_PREC_FP16) return __half2float(__float2half_rn(std::min(std::max(v, -65504.f), 65504.f)));
  return v;
}
template <typename T> struct is16 { static constexpr bool value = std::is_same<T, bf16>::value || std::is_same<T, f16>::value; };

}  // namespace

// ----------------------------------------------------------------------- the context
struct bc_ctx {
  int device = 0;
  int max_batch = 0;
  int num_sms = 0;
  std::string err;
  long long launches = 0;

  int precision = BC_PREC_FP16;
  int initial_pool = 3;     // max-pool of the initial block: 3 = 3x3 s2 p1, 2 = 2x2 s2 (container "__spec__")
  int chunk = 0;            // 0 = default
  int tensor_cores = 1;
  bool umma_ready = false;
  int use_graphs = 1;

  // ---- network (host copies, folded, fp32)
  bool net_loaded = false;
  int num_classes = 0;
  float bn_eps = 1e-5f;
  std::vector<float> h_init_w, h_init_g, h_init_b, h_init_a;   // [27][13], [16] x3
  std::vector<HostBlock> h_blocks;
  std::vector<float> h_full_w;                                  // [9][16][CP]
  // ---- network (device)
  std::vector<void*> dev_allocs;      // weight allocations
  float *d_init_w = nullptr, *d_init_g = nullptr, *d_init_b = nullptr, *d_init_a = nullptr;
  std::vector<Bottleneck> blocks;
  float* d_full_w = nullptr;
  uint8_t* d_head_umma = nullptr;     // tcgen05 operand image of the head (bf16 mode, C <= 16)
  uint8_t* d_init_umma = nullptr;     // tcgen05 operand images of the initial block (bf16 mode): float inputs,
  uint8_t* d_init_umma_u8 = nullptr;  // uint8 frames
  float init_u8_unscale[16] = {0};    // power-of-two un-scale per output channel of the uint8 operand image
  // normalisation LUTs (models.py:91): [256][3] RGB order
  float* d_lut32 = nullptr;
  double* d_lut64 = nullptr;

  // ---- activation scratch
  int scratch_frames = 0, scratch_esz = 0;
  void *bufX = nullptr, *bufY = nullptr, *bufP = nullptr, *bufE1 = nullptr, *bufE2 = nullptr;
  uint8_t *idx1 = nullptr, *idx2 = nullptr;
  uint8_t* d_labels = nullptr;        // [max_batch][256][512]
  uint8_t* d_resized = nullptr;       // [max_batch][256][512][3]
  uint8_t* d_frames_in = nullptr;     // staging for *_host calls
  size_t frames_in_bytes = 0;
  int8_t* d_grids_out = nullptr;
  size_t grids_out_bytes = 0;
  // staging slots of bc_pipeline_host_submit (two steps in flight)
  struct Slot {
    uint8_t* d_in = nullptr; size_t in_bytes = 0;
    int8_t* d_out = nullptr; size_t out_bytes = 0;
    cudaEvent_t copied = nullptr, computed = nullptr, done = nullptr;
    bool busy = false;
  } slots[2];
  long long submitted = 0;

  // ---- BEV
  bool bev_set = false;
  double M[9], Mi[9];
  int in_rows = 0, in_cols = 0, warp_w = 0, warp_h = 0;
  double cm_per_px = 1.0;

  // ---- resize tables, keyed by source size
  std::map<std::pair<int, int>, ResizeTab> resize_tabs;
  // ---- K9 coordinate tables, keyed by the geometry (cleared by bc_set_bev)
  std::vector<std::pair<BevGeom, uint2*>> occ_tables;

  // ---- contour_noise_removal (image_processing_utils.py:4-44)
  void* cn_scratch = nullptr;
  size_t cn_scratch_bytes = 0;
  uint8_t* d_labels_cn = nullptr;     // filtered road masks of the binary pipeline [max_batch][256][512]
  int contour_filter = 0;

  // ---- laserscan-like grids (bev.py:145-164, 216-240): gather tables per grid shape, scratch
  struct LaserTab { int Wc, Hc, binary, pol_w, pol_h; int* d_fwd; int* d_inv; };
  std::vector<LaserTab> laser_tabs;
  uint8_t* laser_cells = nullptr;     // [B][Hc*Wc] plain grid / raw template
  int* laser_first = nullptr;         // [B][pol_h]
  size_t laser_cells_bytes = 0, laser_first_bytes = 0;

  // ---- multi-GPU gather
  int8_t* gather_base = nullptr;
  int rank = 0, world = 1;

  // ---- streaming gather (bc_gather_stream_setup): flags in peer-mapped device memory
  struct GatherStream {
    bool on = false;
    int8_t* gather[2] = {nullptr, nullptr};
    uint32_t* arrive = nullptr;          // rank 0's [2][world]
    uint32_t* release_mine = nullptr;    // this rank's [2]
    uint32_t** d_release_peers = nullptr;  // rank 0: device array [2][world] of pointers to every rank's release flag of slot s
    int* h_err = nullptr;                // pinned, device-mapped: raised by a flag wait that timed out (read without a CUDA call)
    int* d_err = nullptr;                // its device address
    long long steps = 0;
  } gs;

  // ---- CUDA graphs of bc_pipeline[_host]
  std::vector<GraphEntry> graphs;

  // ---- host entry point: copy stream for H2D / compute overlap
  int host_overlap = 1;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t d2h_stream = nullptr;       // grids of the streaming entry point leave on their own stream
  cudaEvent_t copy_done[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t call_start = nullptr;

  // ---- per-kernel profiling (bc_set_profile): one event pair per launch
  int profiling = 0;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  std::string prof_json;
};

static std::string g_create_err;

namespace bc {
// Set by the scheduler (forward_chunk) before each tcgen05 launch: 1 = walk the tiles from the last to the
// first.  The activations of a 256-frame batch (134-268 MB per tensor) do not fit the 126 MB L2, but the tail
// of what a kernel wrote is still there when the next one starts: a consumer that walks in the opposite
// direction reads those tiles first (measured: 3 % per launch).
thread_local int g_umma_reverse = 0;     // one context per thread (include/bugcar_b200.h): no sharing

bool umma_supported(const Bottleneck& bn) {
  if (bn.kind != 1 && bn.kind != 2) return false;
  return (bn.cin == 128 && bn.ci == 32) || (bn.cin == 64 && bn.ci == 16);
}

void umma_free(UmmaPack& p) {
  if (p.wblob) cudaFree(p.wblob);
  p = UmmaPack();
}
}  // namespace bc

namespace {

// Stream-ordered wait for "*flag >= value" WITHOUT occupying an SM: the driver's stream memory operation
// (cuStreamWaitValue32, polled by the GPU front end).  A spinning wait KERNEL next to the persistent compute kernels
// made the SM it landed on a straggler for every kernel of the step (N = 2 end to end 3.85 instead of 2.8 ms per
// step); it remains the fallback when the driver entry point is missing.
typedef int (*PFN_streamWaitValue32)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
PFN_streamWaitValue32 stream_wait_fn() {
  static PFN_streamWaitValue32 fn = nullptr;
  static bool looked = false;
  if (!looked) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (!getenv("BC_GS_SPIN") && cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_streamWaitValue32)p;
    looked = true;
  }
  return fn;
}
void wait_flags(bc_ctx* c, const uint32_t* flags, int n, uint32_t value, int* d_err, cudaStream_t s) {
  if (PFN_streamWaitValue32 w = stream_wait_fn()) {
    bool ok = true;
    for (int i = 0; ok && i < n; ++i) ok = w(s, (unsigned long long)(uintptr_t)(flags + i), value, 0u /* CU_STREAM_WAIT_VALUE_GEQ */) == 0;
    if (ok) return;
  }
  launch_flag_wait(flags, n, value, d_err, s);
  c->launches++;
}

int fail(bc_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_err = msg;
  return code;
}

#define CU(expr)                                                                     \
  do {                                                                               \
    cudaError_t e_ = (expr);                                                         \
    if (e_ != cudaSuccess)                                                           \
      return fail(c, BC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

int check_launch(bc_ctx* c, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(c, BC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return BC_OK;
}

// Every kernel launch goes through L(): counts it and, in profiling mode, brackets it with
// a CUDA event pair on the launching stream and records its algorithmic bytes / flops.
cudaEvent_t prof_event(bc_ctx* c) {
  cudaEvent_t e;
  if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
  cudaEventCreate(&e);
  return e;
}
template <typename F>
inline void L(bc_ctx* c, const char* name, double bytes, double flops, cudaStream_t s, F&& launch) {
  c->launches++;
  if (!c->profiling) { launch(); return; }
  ProfRec r{name, bytes, flops, prof_event(c), prof_event(c)};
  cudaEventRecord(r.start, s);
  launch();
  cudaEventRecord(r.stop, s);
  c->prof.push_back(r);
}

// ----------------------------------------------------------- weight container parsing
// BCENETW1 layout: see bugcar_image_segmentation_b200/weights.py
struct Container {
  std::map<std::string, Tensor> t;
  int num_classes = 0;
  float bn_eps = 1e-5f;
};

int parse_container(bc_ctx* c, const void* blob, size_t n, Container& out) {
  const uint8_t* p = (const uint8_t*)blob;
  if (n < 24 || memcmp(p, "BCENETW1", 8) != 0) return fail(c, BC_ERR_FORMAT, "not a BCENETW1 weight container");
  uint32_t nt, nc, rsv;
  float eps;
  memcpy(&nt, p + 8, 4); memcpy(&nc, p + 12, 4); memcpy(&eps, p + 16, 4); memcpy(&rsv, p + 20, 4);
  const size_t ent = 96 + 4 + 16 + 8;
  size_t data0 = 24 + (size_t)nt * ent;
  data0 += (64 - data0 % 64) % 64;
  if (data0 > n) return fail(c, BC_ERR_FORMAT, "truncated weight container (table)");
  out.num_classes = (int)nc;
  out.bn_eps = eps;
  for (uint32_t i = 0; i < nt; ++i) {
    const uint8_t* e = p + 24 + (size_t)i * ent;
    char name[97];
    memcpy(name, e, 96); name[96] = 0;
    uint32_t ndim, d[4];
    uint64_t off;
    memcpy(&ndim, e + 96, 4); memcpy(d, e + 100, 16); memcpy(&off, e + 116, 8);
    if (ndim > 4) return fail(c, BC_ERR_FORMAT, std::string(name) + ": ndim > 4");
    Tensor t;
    t.ndim = (int)ndim;
    size_t cnt = 1;
    for (int k = 0; k < 4; ++k) {
      t.dims[k] = (int)d[k];
      if (k >= (int)ndim) continue;
      // no ENet tensor has an axis beyond a few hundred: a cap keeps the product far from overflow
      if (d[k] < 1 || d[k] > 65536) return fail(c, BC_ERR_FORMAT, std::string(name) + ": dimension outside [1, 65536]");
      cnt *= d[k];
      if (cnt > (size_t)1 << 32) return fail(c, BC_ERR_FORMAT, std::string(name) + ": tensor too large");
    }
    const size_t avail = n - data0;
    if (off > avail || cnt * 4 > avail - off) return fail(c, BC_ERR_FORMAT, std::string(name) + ": data out of bounds");
    if ((data0 + off) % 4 != 0) return fail(c, BC_ERR_FORMAT, std::string(name) + ": misaligned data");
    t.data = (const float*)(p + data0 + off);
    out.t[name] = t;
  }
  return BC_OK;
}

const Tensor* find(const Container& ct, const std::string& name) {
  auto it = ct.t.find(name);
  return it == ct.t.end() ? nullptr : &it->second;
}

int need(bc_ctx* c, const Container& ct, const std::string& name, int ndim, const int* dims, const Tensor** out) {
  const Tensor* t = find(ct, name);
  if (!t) return fail(c, BC_ERR_FORMAT, "missing tensor " + name);
  bool ok = t->ndim == ndim;
  for (int i = 0; ok && i < ndim; ++i) ok = t->dims[i] == dims[i];
  if (!ok) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: unexpected shape (%d dims: %d,%d,%d,%d)", name.c_str(), t->ndim,
             t->dims[0], t->dims[1], t->dims[2], t->dims[3]);
    return fail(c, BC_ERR_FORMAT, buf);
  }
  *out = t;
  return BC_OK;
}

// batch-norm scale/shift: g = gamma / sqrt(var + eps), b = beta - mean * g
int bn_fold(bc_ctx* c, const Container& ct, const std::string& p, int ch, std::vector<double>& g, std::vector<double>& b) {
  const Tensor *tw, *tb, *tm, *tv;
  int d[1] = {ch};
  int r;
  if ((r = need(c, ct, p + ".weight", 1, d, &tw))) return r;
  if ((r = need(c, ct, p + ".bias", 1, d, &tb))) return r;
  if ((r = need(c, ct, p + ".running_mean", 1, d, &tm))) return r;
  if ((r = need(c, ct, p + ".running_var", 1, d, &tv))) return r;
  g.resize(ch); b.resize(ch);
  for (int i = 0; i < ch; ++i) {
    // the oracle evaluates batch_norm in fp32: keep the fp32 rsqrt argument
    double inv = 1.0 / std::sqrt((double)(float)(tv->data[i] + ct.bn_eps));
    g[i] = (double)tw->data[i] * inv;
    b[i] = (double)tb->data[i] - (double)tm->data[i] * g[i];
  }
  return BC_OK;
}

// activation slopes: missing tensor = ReLU (slope 0), 1 element = shared PReLU, ch elements = per channel
int act_alpha(bc_ctx* c, const Container& ct, const std::string& name, int ch, std::vector<float>& a) {
  a.assign(ch, 0.f);
  const Tensor* t = find(ct, name + ".weight");
  if (!t) return BC_OK;
  size_t n = t->count();
  if (n == 1) { for (int i = 0; i < ch; ++i) a[i] = t->data[0]; return BC_OK; }
  if ((int)n == ch) { for (int i = 0; i < ch; ++i) a[i] = t->data[i]; return BC_OK; }
  return fail(c, BC_ERR_FORMAT, name + ".weight: PReLU slope count matches neither 1 nor the channel count");
}

// conv (cout,cin,kh,kw) [transposed: (cin,cout,kh,kw)] + BN -> [tap][cin][cout] folded.
// act: "" = identity (slope 1), else activation tensor prefix.
int fold_conv(bc_ctx* c, const Container& ct, const std::string& conv, const std::string& bn,
              const std::string& act, bool has_act, int cin, int cout, int kh, int kw, bool transposed, HostConv& out) {
  const Tensor* w;
  int d[4] = {transposed ? cin : cout, transposed ? cout : cin, kh, kw};
  int r;
  if ((r = need(c, ct, conv + ".weight", 4, d, &w))) return r;
  std::vector<double> g, b;
  if ((r = bn_fold(c, ct, bn, cout, g, b))) return r;
  out.cin = cin; out.cout = cout; out.ntaps = kh * kw;
  out.w.resize((size_t)kh * kw * cin * cout);
  out.bias.resize(cout);
  // an optional convolution bias (checkpoints built with bias=True) goes through the batch norm with the output:
  // BN(conv + cb) = conv * g + (b + cb * g)
  if (const Tensor* cb = find(ct, conv + ".bias")) {
    if (cb->count() != (size_t)cout) return fail(c, BC_ERR_FORMAT, conv + ".bias: expected one value per output channel");
    for (int o = 0; o < cout; ++o) b[o] += (double)cb->data[o] * g[o];
  }
  for (int o = 0; o < cout; ++o) out.bias[o] = (float)b[o];
  for (int ky = 0; ky < kh; ++ky)
    for (int kx = 0; kx < kw; ++kx)
      for (int i = 0; i < cin; ++i)
        for (int o = 0; o < cout; ++o) {
          size_t src = transposed ? (((size_t)i * cout + o) * kh + ky) * kw + kx
                                  : (((size_t)o * cin + i) * kh + ky) * kw + kx;
          out.w[((size_t)(ky * kw + kx) * cin + i) * cout + o] = (float)((double)w->data[src] * g[o]);
        }
  if (has_act) return act_alpha(c, ct, act, cout, out.alpha);
  out.alpha.assign(cout, 1.f);   // identity
  return BC_OK;
}

struct BlockSpec { const char* name; int kind; int a, b; };
// canonical ENet encoder/decoder (weights.py ENET_BLOCKS): used when the container carries no
// "__graph__" tensor
const BlockSpec kBlocks[] = {
    {"downsample1_0", 0, 16, 64},
    {"regular1_1", 1, 64, 1}, {"regular1_2", 1, 64, 1}, {"regular1_3", 1, 64, 1}, {"regular1_4", 1, 64, 1},
    {"downsample2_0", 0, 64, 128},
    {"regular2_1", 1, 128, 1}, {"dilated2_2", 1, 128, 2}, {"asymmetric2_3", 2, 128, 1}, {"dilated2_4", 1, 128, 4},
    {"regular2_5", 1, 128, 1}, {"dilated2_6", 1, 128, 8}, {"asymmetric2_7", 2, 128, 1}, {"dilated2_8", 1, 128, 16},
    {"regular3_0", 1, 128, 1}, {"dilated3_1", 1, 128, 2}, {"asymmetric3_2", 2, 128, 1}, {"dilated3_3", 1, 128, 4},
    {"regular3_4", 1, 128, 1}, {"dilated3_5", 1, 128, 8}, {"asymmetric3_6", 2, 128, 1}, {"dilated3_7", 1, 128, 16},
    {"upsample4_0", 3, 128, 64},
    {"regular4_1", 1, 64, 1}, {"regular4_2", 1, 64, 1},
    {"upsample5_0", 3, 64, 16},
    {"regular5_1", 1, 16, 1},
};

// One block of the graph as the container describes it (weights.py "__graph__": kind, stage, index, cin, cout,
// internal width, dilation) or as kBlocks implies it.
struct GraphRow { std::string name; int kind, cin, cout, ci, dilation; };

// The block list of this network, checked against what the kernels implement: the channel chain 16 -> ... -> 16,
// every down-sampling block closed by an up-sampling one of the mirrored widths, and per block the widths the
// CUDA-core kernels are instantiated for (every tcgen05 kernel covers a subset of those and falls back to them).
int read_graph(bc_ctx* c, const Container& ct, std::vector<GraphRow>& rows) {
  rows.clear();
  const Tensor* g = find(ct, "__graph__");
  if (!g) {
    for (const BlockSpec& s : kBlocks) {
      GraphRow r;
      r.name = s.name; r.kind = s.kind;
      if (s.kind == 0 || s.kind == 3) { r.cin = s.a; r.cout = s.b; r.dilation = 1; }
      else { r.cin = r.cout = s.a; r.dilation = s.b; }
      r.ci = r.cin / 4;
      rows.push_back(r);
    }
    return BC_OK;
  }
  if (g->ndim != 2 || g->dims[1] != 7 || g->dims[0] < 1 || g->dims[0] > 256)
    return fail(c, BC_ERR_FORMAT, "__graph__: expected [n_blocks][7] with 1 <= n_blocks <= 256");
  static const char* base[4] = {"downsample", "regular", "asymmetric", "upsample"};
  for (int i = 0; i < g->dims[0]; ++i) {
    const float* v = g->data + (size_t)i * 7;
    for (int k = 0; k < 7; ++k)
      if (!(v[k] >= 0.f && v[k] <= 4096.f) || v[k] != std::floor(v[k])) return fail(c, BC_ERR_FORMAT, "__graph__: entries must be small non-negative integers");
    GraphRow r;
    r.kind = (int)v[0]; r.cin = (int)v[3]; r.cout = (int)v[4]; r.ci = (int)v[5]; r.dilation = (int)v[6];
    if (r.kind > 3) return fail(c, BC_ERR_FORMAT, "__graph__: unknown block kind");
    char nm[64];
    snprintf(nm, sizeof nm, "%s%d_%d", (r.kind == 1 && r.dilation != 1) ? "dilated" : base[r.kind], (int)v[1], (int)v[2]);
    r.name = nm;
    rows.push_back(r);
  }
  // structural validation
  int ch = 16, level = 0;
  std::vector<int> open_down;          // input width of the open down-sampling blocks
  for (const GraphRow& r : rows) {
    char msg[256];
    auto bad = [&](const char* why) {
      snprintf(msg, sizeof msg, "%s (cin %d, cout %d, internal %d, dilation %d): %s", r.name.c_str(), r.cin, r.cout, r.ci, r.dilation, why);
      return fail(c, BC_ERR_FORMAT, msg);
    };
    if (r.cin != ch) return bad("input width does not match the previous block's output");
    if (r.kind == 0) {
      if (!((r.cin == 16 && r.cout == 64) || (r.cin == 64 && r.cout == 128))) return bad("down-sampling blocks are implemented for 16->64 and 64->128");
      if (r.ci != r.cin / 4) return bad("down-sampling blocks are implemented with internal width cin/4 (the out/4 variant of the paper is not)");
      if (level >= 2) return bad("more than two down-sampling levels");
      open_down.push_back(r.cin); ++level;
    } else if (r.kind == 3) {
      if (open_down.empty()) return bad("up-sampling block without an open down-sampling block");
      if (!((r.cin == 128 && r.cout == 64) || (r.cin == 64 && r.cout == 16)) || r.ci != r.cin / 4) return bad("up-sampling blocks are implemented for 128->64 and 64->16 with internal width cin/4");
      if (r.cout != open_down.back()) return bad("output width must equal the input width of the down-sampling block it mirrors");
      open_down.pop_back(); --level;
    } else {
      if (r.cin != r.cout) return bad("regular / asymmetric blocks keep their width");
      if (r.ci != r.cin / 4) return bad("internal width must be a quarter of the block width");
      if (!(r.cin == 64 || r.cin == 128 || (r.cin == 16 && r.kind == 1 && r.dilation == 1))) return bad("implemented widths: 64, 128, and 16 (plain 3x3 only)");
      if (r.kind == 1 && (r.dilation < 1 || r.dilation > 16)) return bad("dilation must be in [1, 16]");
      if (r.kind == 2 && r.cin != 128) return bad("asymmetric blocks are implemented at width 128");
      if ((r.cin == 64 && level != 1) || (r.cin == 128 && level != 2) || (r.cin == 16 && level != 0)) return bad("width does not match the resolution level");
    }
    ch = r.cout;
  }
  if (!open_down.empty() || ch != 16) return fail(c, BC_ERR_FORMAT, "__graph__: the decoder must return to 16 channels at full block resolution");
  return BC_OK;
}

int build_host_net(bc_ctx* c, const Container& ct) {
  int r;
  c->num_classes = ct.num_classes;
  c->bn_eps = ct.bn_eps;
  if (ct.num_classes < 1 || ct.num_classes > 32) return fail(c, BC_ERR_FORMAT, "num_classes must be in [1, 32]");
  // variant switches (weights.py "__spec__"): initial max-pool kernel, head kernel
  c->initial_pool = 3;
  int head_kernel = 3;
  if (const Tensor* sp = find(ct, "__spec__")) {
    if (sp->count() < 2) return fail(c, BC_ERR_FORMAT, "__spec__: expected at least 2 entries");
    c->initial_pool = (int)sp->data[0];
    head_kernel = (int)sp->data[1];
    if (c->initial_pool != 2 && c->initial_pool != 3) return fail(c, BC_ERR_FORMAT, "__spec__: initial max-pool kernel must be 2 (2x2 s2) or 3 (3x3 s2 p1)");
    if (head_kernel != 2 && head_kernel != 3) return fail(c, BC_ERR_FORMAT, "__spec__: head kernel must be 2 (2x2 s2) or 3 (3x3 s2 p1 op1)");
  }
  if (find(ct, "transposed_conv.bias")) return fail(c, BC_ERR_FORMAT, "transposed_conv.bias: a bias on the class head is not implemented");
  std::vector<GraphRow> graph;
  if ((r = read_graph(c, ct, graph))) return r;
  // initial block: raw conv weights, BN as a per-channel affine on all 16 channels
  {
    const Tensor* w;
    int d[4] = {13, 3, 3, 3};
    if ((r = need(c, ct, "initial_block.main_branch.weight", 4, d, &w))) return r;
    c->h_init_w.resize(27 * 13);
    for (int o = 0; o < 13; ++o)
      for (int ch = 0; ch < 3; ++ch)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx)
            c->h_init_w[((ch * 3 + ky) * 3 + kx) * 13 + o] = w->data[((o * 3 + ch) * 3 + ky) * 3 + kx];
    std::vector<double> g, b;
    if ((r = bn_fold(c, ct, "initial_block.batch_norm", 16, g, b))) return r;
    if (const Tensor* cb = find(ct, "initial_block.main_branch.bias")) {
      if (cb->count() != 13) return fail(c, BC_ERR_FORMAT, "initial_block.main_branch.bias: expected 13 values");
      for (int i = 0; i < 13; ++i) b[i] += (double)cb->data[i] * g[i];
    }
    c->h_init_g.resize(16); c->h_init_b.resize(16);
    for (int i = 0; i < 16; ++i) { c->h_init_g[i] = (float)g[i]; c->h_init_b[i] = (float)b[i]; }
    if ((r = act_alpha(c, ct, "initial_block.out_activation", 16, c->h_init_a))) return r;
  }
  c->h_blocks.clear();
  for (const GraphRow& s : graph) {
    HostBlock hb;
    hb.name = s.name;
    hb.kind = s.kind;
    std::string n = s.name;
    if (s.kind == 0) {          // downsample: cin -> cout, internal cin/4
      hb.cin = s.cin; hb.cout = s.cout; hb.ci = s.ci;
      if ((r = fold_conv(c, ct, n + ".ext_conv1.0", n + ".ext_conv1.1", n + ".ext_conv1.2", true, hb.cin, hb.ci, 2, 2, false, hb.c1))) return r;
      if ((r = fold_conv(c, ct, n + ".ext_conv2.0", n + ".ext_conv2.1", n + ".ext_conv2.2", true, hb.ci, hb.ci, 3, 3, false, hb.c2))) return r;
      if ((r = fold_conv(c, ct, n + ".ext_conv3.0", n + ".ext_conv3.1", n + ".ext_conv3.2", true, hb.ci, hb.cout, 1, 1, false, hb.c3))) return r;
    } else if (s.kind == 1 || s.kind == 2) {
      hb.cin = hb.cout = s.cin; hb.ci = s.ci; hb.dilation = s.dilation;
      if ((r = fold_conv(c, ct, n + ".ext_conv1.0", n + ".ext_conv1.1", n + ".ext_conv1.2", true, hb.cin, hb.ci, 1, 1, false, hb.c1))) return r;
      if (s.kind == 1) {
        if ((r = fold_conv(c, ct, n + ".ext_conv2.0", n + ".ext_conv2.1", n + ".ext_conv2.2", true, hb.ci, hb.ci, 3, 3, false, hb.c2))) return r;
      } else {
        if ((r = fold_conv(c, ct, n + ".ext_conv2.0", n + ".ext_conv2.1", n + ".ext_conv2.2", true, hb.ci, hb.ci, 5, 1, false, hb.c2))) return r;
        if ((r = fold_conv(c, ct, n + ".ext_conv2.3", n + ".ext_conv2.4", n + ".ext_conv2.5", true, hb.ci, hb.ci, 1, 5, false, hb.c2b))) return r;
      }
      if ((r = fold_conv(c, ct, n + ".ext_conv3.0", n + ".ext_conv3.1", n + ".ext_conv3.2", true, hb.ci, hb.cout, 1, 1, false, hb.c3))) return r;
    } else {                    // upsample: cin -> cout, internal cin/4
      hb.cin = s.cin; hb.cout = s.cout; hb.ci = s.ci;
      if ((r = fold_conv(c, ct, n + ".main_conv1.0", n + ".main_conv1.1", "", false, hb.cin, hb.cout, 1, 1, false, hb.cm))) return r;
      if ((r = fold_conv(c, ct, n + ".ext_conv1.0", n + ".ext_conv1.1", n + ".ext_conv1.2", true, hb.cin, hb.ci, 1, 1, false, hb.c1))) return r;
      if ((r = fold_conv(c, ct, n + ".ext_tconv1", n + ".ext_tconv1_bnorm", n + ".ext_tconv1_activation", true, hb.ci, hb.ci, 2, 2, true, hb.c2))) return r;
      if ((r = fold_conv(c, ct, n + ".ext_conv2.0", n + ".ext_conv2.1", "", false, hb.ci, hb.cout, 1, 1, false, hb.c3))) return r;
    }
    if ((r = act_alpha(c, ct, n + ".out_activation", hb.cout, hb.alpha_out))) return r;
    c->h_blocks.push_back(std::move(hb));
  }
  {
    const Tensor* w;
    int C = ct.num_classes;
    int d[4] = {16, C, head_kernel, head_kernel};
    if ((r = need(c, ct, "transposed_conv.weight", 4, d, &w))) return r;
    int cp = C <= 16 ? 16 : 32;
    c->h_full_w.assign((size_t)9 * 16 * cp, 0.f);
    // A 2x2 stride-2 transposed conv is the 3x3 s2 p1 op1 form with five taps zero: output (2i+qy, 2j+qx) takes
    // input (i, j) only, which the 3x3 form reaches through taps ky = 1 + qy, kx = 1 + qx.
    const int K = head_kernel, off = K == 3 ? 0 : 1;
    for (int k = 0; k < 16; ++k)
      for (int o = 0; o < C; ++o)
        for (int ky = 0; ky < K; ++ky)
          for (int kx = 0; kx < K; ++kx)
            c->h_full_w[((size_t)((ky + off) * 3 + kx + off) * 16 + k) * cp + o] = w->data[(((size_t)k * C + o) * K + ky) * K + kx];
  }
  return BC_OK;
}

// ------------------------------------------------------------------- device upload
void free_net(bc_ctx* c) {
  for (Bottleneck& b : c->blocks) { umma_free(b.um_a); umma_free(b.um_b); umma_free(b.um_f); }
  for (void* p : c->dev_allocs) cudaFree(p);
  if (c->d_head_umma) cudaFree(c->d_head_umma);
  if (c->d_init_umma) cudaFree(c->d_init_umma);
  if (c->d_init_umma_u8) cudaFree(c->d_init_umma_u8);
  c->d_head_umma = c->d_init_umma = c->d_init_umma_u8 = nullptr;
  c->dev_allocs.clear();
  c->blocks.clear();
  c->d_init_w = c->d_init_g = c->d_init_b = c->d_init_a = c->d_full_w = nullptr;
}

int upload_vec(bc_ctx* c, const std::vector<float>& v, bool round16, float** out) {
  std::vector<float> tmp;
  const float* src = v.data();
  if (round16 && c->precision != BC_PREC_FP32) {
    tmp.resize(v.size());
    for (size_t i = 0; i < v.size(); ++i) tmp[i] = round_to(c->precision, v[i]);
    src = tmp.data();
  }
  float* d = nullptr;
  CU(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(float)));
  c->dev_allocs.push_back(d);
  CU(cudaMemcpy(d, src, v.size() * sizeof(float), cudaMemcpyHostToDevice));
  *out = d;
  return BC_OK;
}

int upload_conv(bc_ctx* c, const HostConv& h, bool bf, ConvP& p) {
  if (h.ntaps == 0) return BC_OK;
  int r;
  p.cin = h.cin; p.cout = h.cout; p.ntaps = h.ntaps;
  if ((r = upload_vec(c, h.w, bf, &p.w))) return r;
  if ((r = upload_vec(c, h.bias, false, &p.bias))) return r;
  if ((r = upload_vec(c, h.alpha, false, &p.alpha))) return r;
  return BC_OK;
}

void invalidate_graphs(bc_ctx* c) {
  for (auto& g : c->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  c->graphs.clear();
}

// tcgen05 operand packs (16-bit modes): conv (+ expansion + the NEXT block's projection)
template <typename AT>
int build_umma_packs(bc_ctx* c) {
  typedef Umma<AT> U;
  if (U::available()) {
    for (size_t i = 0; i < c->blocks.size(); ++i) {
      Bottleneck& b = c->blocks[i];
      if (b.kind == 3) {     // upsampling bottleneck (+ the next regular block's projection when it is 64 wide)
        const HostBlock& hb = c->h_blocks[i];
        const HostBlock* nx = nullptr;
        if (b.cout == 64 && i + 1 < c->blocks.size() && umma_supported(c->blocks[i + 1]) && c->blocks[i + 1].cin == 64)
          nx = &c->h_blocks[i + 1];
        if (!U::up_build(b.um_a, b.cin, b.ci, b.cout, hb.cm.w.data(), hb.cm.bias.data(), hb.c1.w.data(), hb.c1.bias.data(),
                      hb.c1.alpha.data(), hb.c2.w.data(), hb.c2.bias.data(), hb.c2.alpha.data(), hb.c3.w.data(),
                      hb.c3.bias.data(), hb.alpha_out.data(), nx ? nx->c1.w.data() : nullptr,
                      nx ? nx->c1.bias.data() : nullptr, nx ? nx->c1.alpha.data() : nullptr))
          return fail(c, BC_ERR_CUDA, "building the tcgen05 upsampling operands failed");
        continue;
      }
      if (b.kind == 0) {     // down-sampling bottleneck: 3x3 conv + expansion + pooled residual + next projection
        const HostBlock& hb = c->h_blocks[i];
        if (i + 1 >= c->blocks.size() || !umma_supported(c->blocks[i + 1]) || c->blocks[i + 1].cin != b.cout) continue;
        const HostBlock& nx = c->h_blocks[i + 1];
        const int cip = 16;                         // internal width zero-padded to the K granularity (4 -> 16 at stage 1)
        std::vector<float> cw((size_t)9 * cip * cip, 0.f), cb(cip, 0.f), ca(cip, 1.f), ew((size_t)cip * b.cout, 0.f);
        for (int t = 0; t < 9; ++t)
          for (int k = 0; k < b.ci; ++k)
            for (int o = 0; o < b.ci; ++o) cw[((size_t)t * cip + k) * cip + o] = hb.c2.w[((size_t)t * b.ci + k) * b.ci + o];
        for (int o = 0; o < b.ci; ++o) { cb[o] = hb.c2.bias[o]; ca[o] = hb.c2.alpha[o]; }
        for (int k = 0; k < b.ci; ++k)
          for (int o = 0; o < b.cout; ++o) ew[(size_t)k * b.cout + o] = hb.c3.w[(size_t)k * b.cout + o];
        if (!U::build(b.um_a, b.cout, cip, nx.ci, b.cin, cw.data(), 9, cb.data(), ca.data(), ew.data(), hb.c3.bias.data(),
                        hb.c3.alpha.data(), hb.alpha_out.data(), nx.c1.w.data(), nx.c1.bias.data(), nx.c1.alpha.data()))
          return fail(c, BC_ERR_CUDA, "building the tcgen05 down-sampling operands failed");
        if (!U::down_build(b.um_b, b.cin, b.ci, hb.c1.w.data(), hb.c1.bias.data(), hb.c1.alpha.data()))
          return fail(c, BC_ERR_CUDA, "building the tcgen05 pooling / strided-conv operands failed");
        continue;
      }
      if (!umma_supported(b)) continue;
      const HostBlock& hb = c->h_blocks[i];
      const HostBlock* nx = nullptr;
      if (i + 1 < c->blocks.size() && umma_supported(c->blocks[i + 1]) && c->blocks[i + 1].cin == b.cin)
        nx = &c->h_blocks[i + 1];
      const float* nw = nx ? nx->c1.w.data() : nullptr;
      const float* nb = nx ? nx->c1.bias.data() : nullptr;
      const float* na = nx ? nx->c1.alpha.data() : nullptr;
      bool ok;
      if (b.kind == 1) {
        ok = U::build(b.um_a, b.cin, b.ci, b.ci, b.cin, hb.c2.w.data(), 9, hb.c2.bias.data(), hb.c2.alpha.data(), hb.c3.w.data(),
                        hb.c3.bias.data(), hb.c3.alpha.data(), hb.alpha_out.data(), nw, nb, na);
      } else {
        ok = U::build(b.um_a, b.cin, b.ci, b.ci, b.cin, hb.c2.w.data(), 5, hb.c2.bias.data(), hb.c2.alpha.data(), nullptr, nullptr,
                        nullptr, nullptr, nullptr, nullptr, nullptr) &&
             U::build(b.um_b, b.cin, b.ci, b.ci, b.cin, hb.c2b.w.data(), 5, hb.c2b.bias.data(), hb.c2b.alpha.data(), hb.c3.w.data(),
                        hb.c3.bias.data(), hb.c3.alpha.data(), hb.alpha_out.data(), nw, nb, na) &&
             (b.cin != 128 ||        // the one-kernel form of the block (128 channels)
              U::build_asym(b.um_f, hb.c2.w.data(), hb.c2.bias.data(), hb.c2.alpha.data(), hb.c2b.w.data(), hb.c2b.bias.data(),
                            hb.c2b.alpha.data(), hb.c3.w.data(), hb.c3.bias.data(), hb.c3.alpha.data(), hb.alpha_out.data(), nw, nb, na));
      }
      if (!ok) return fail(c, BC_ERR_CUDA, "building the tcgen05 operand packs failed");
    }
    if (!U::initial_build(&c->d_init_umma, c->h_init_w.data()) || !U::initial_build_u8(&c->d_init_umma_u8, c->h_init_w.data(), c->init_u8_unscale))
      return fail(c, BC_ERR_CUDA, "building the tcgen05 initial-block operands failed");
    if (c->num_classes <= 16) {
      std::vector<float> hw(c->h_full_w.size());
      for (size_t i = 0; i < hw.size(); ++i) hw[i] = round_to(c->precision, c->h_full_w[i]);
      if (!U::head_build(&c->d_head_umma, hw.data(), c->num_classes, 16))
        return fail(c, BC_ERR_CUDA, "building the tcgen05 head operands failed");
    }
    c->umma_ready = true;
  }
  return BC_OK;
}

int upload_net(bc_ctx* c) {
  invalidate_graphs(c);
  free_net(c);
  const bool bf = c->precision != BC_PREC_FP32;     // 16-bit storage: operands rounded to that type
  int r;
  // The initial block and the head read fp32 weights in every mode, except that the 16-bit modes
  // round the head's weights (its input is 16-bit and the oracle's emulation does the same).
  if ((r = upload_vec(c, c->h_init_w, false, &c->d_init_w))) return r;
  if ((r = upload_vec(c, c->h_init_g, false, &c->d_init_g))) return r;
  if ((r = upload_vec(c, c->h_init_b, false, &c->d_init_b))) return r;
  if ((r = upload_vec(c, c->h_init_a, false, &c->d_init_a))) return r;
  if ((r = upload_vec(c, c->h_full_w, bf, &c->d_full_w))) return r;
  for (const HostBlock& hb : c->h_blocks) {
    Bottleneck b;
    b.name = hb.name; b.kind = hb.kind; b.cin = hb.cin; b.cout = hb.cout; b.ci = hb.ci; b.dilation = hb.dilation;
    if ((r = upload_conv(c, hb.c1, bf, b.c1))) return r;
    if ((r = upload_conv(c, hb.c2, bf, b.c2))) return r;
    if ((r = upload_conv(c, hb.c2b, bf, b.c2b))) return r;
    if ((r = upload_conv(c, hb.c3, bf, b.c3))) return r;
    if ((r = upload_conv(c, hb.cm, bf, b.cm))) return r;
    if ((r = upload_vec(c, hb.alpha_out, false, &b.alpha_out))) return r;
    if (hb.kind == 1 && hb.cin == 16 && hb.ci == 4) {      // stage 5: parameters travel by value (simt_stage5.cu)
      Stage5Params sp;
      auto wv = [&](float v) { return round_to(c->precision, v); };   // same operand rounding as upload_conv
      for (int i = 0; i < 64; ++i) { sp.w1[i] = wv(hb.c1.w[i]); sp.w3[i] = wv(hb.c3.w[i]); }
      for (int i = 0; i < 144; ++i) sp.w2[i] = wv(hb.c2.w[i]);
      for (int i = 0; i < 4; ++i) { sp.b1[i] = hb.c1.bias[i]; sp.a1[i] = hb.c1.alpha[i]; sp.b2[i] = hb.c2.bias[i]; sp.a2[i] = hb.c2.alpha[i]; }
      for (int i = 0; i < 16; ++i) { sp.b3[i] = hb.c3.bias[i]; sp.a3[i] = hb.c3.alpha[i]; sp.aout[i] = hb.alpha_out[i]; }
      b.s5.resize(sizeof sp / sizeof(float));
      memcpy(b.s5.data(), &sp, sizeof sp);
    }
    c->blocks.push_back(b);
  }
  c->umma_ready = false;
  if (c->precision == BC_PREC_BF16) return build_umma_packs<bf16>(c);
  if (c->precision == BC_PREC_FP16) return build_umma_packs<f16>(c);
  return BC_OK;
}

// ------------------------------------------------------------------------- scratch
void free_scratch(bc_ctx* c) {
  void** ps[] = {&c->bufX, &c->bufY, &c->bufP, &c->bufE1, &c->bufE2, (void**)&c->idx1, (void**)&c->idx2};
  for (void** p : ps) { if (*p) cudaFree(*p); *p = nullptr; }
  c->scratch_frames = 0;
}

int chunk_frames(const bc_ctx* c) {
  // default: the whole batch in one pass (measured on B200: 22.8k frames/s at chunk 256 vs
  // 19.4k at chunk 32 -- launch granularity beats L2 residency with these kernels)
  int ch = c->chunk > 0 ? c->chunk : c->max_batch;
  return std::min(ch, c->max_batch);
}

int ensure_scratch(bc_ctx* c) {
  int frames = chunk_frames(c);
  int esz = c->precision == BC_PREC_FP32 ? 4 : 2;
  if (c->scratch_frames == frames && c->scratch_esz == esz) return BC_OK;
  invalidate_graphs(c);
  free_scratch(c);
  const size_t big = 524288, small = 131072;   // elements per frame
  CU(cudaMalloc(&c->bufX, big * esz * frames));
  CU(cudaMalloc(&c->bufY, big * esz * frames));
  CU(cudaMalloc(&c->bufP, small * esz * frames));
  CU(cudaMalloc(&c->bufE1, small * esz * frames));
  CU(cudaMalloc(&c->bufE2, small * esz * frames));
  CU(cudaMalloc(&c->idx1, small * frames));
  CU(cudaMalloc(&c->idx2, small * frames));
  c->scratch_frames = frames;
  c->scratch_esz = esz;
  return BC_OK;
}

// ------------------------------------------------------------------ layer schedule
Taps taps_for(int kh, int kw, int dil) {
  Taps t{};
  for (int ky = 0; ky < kh; ++ky)
    for (int kx = 0; kx < kw; ++kx) {
      t.dy[ky * kw + kx] = (int8_t)((ky - kh / 2) * dil);
      t.dx[ky * kw + kx] = (int8_t)((kx - kw / 2) * dil);
    }
  return t;
}

// n frames; x points at this chunk's input; exactly one of logits / labels is non-null.
template <typename T>
int forward_chunk(bc_ctx* c, const void* x, int kind, int n, float* logits, uint8_t* labels,
                  const Lut256* lut, cudaStream_t s, int stop_after = -2, float* dump = nullptr) {
  T* X = (T*)c->bufX;
  T* Y = (T*)c->bufY;
  T* P = (T*)c->bufP;
  T* E1 = (T*)c->bufE1;
  T* E2 = (T*)c->bufE2;
  const double esz = sizeof(T);
  const double in_px_bytes = kind == BC_IN_BGR_U8 ? 3.0 : kind == BC_IN_NCHW_F32 ? 12.0 : 24.0;
  bool init_tc = false;
  if constexpr (is16<T>::value) init_tc = c->tensor_cores && c->umma_ready && c->d_init_umma;
  if (init_tc) {
    if constexpr (is16<T>::value) {
      cudaError_t ce = cudaSuccess;
      L(c, "umma_initial", n * (131072.0 * in_px_bytes + 524288.0 * esz), n * 2.0 * 32768 * 27 * 13, s, [&] {
        ce = Umma<T>::launch_initial(x, kind, n, X, c->initial_pool, c->d_init_umma, c->d_init_umma_u8, c->init_u8_unscale, c->d_lut32, c->h_init_g.data(), c->h_init_b.data(),
                                 c->h_init_a.data(), c->num_sms, s);
      });
      if (ce != cudaSuccess) return fail(c, BC_ERR_CUDA, std::string("tcgen05 initial-block launch: ") + cudaGetErrorString(ce));
    }
  } else {
    L(c, "initial", n * (131072.0 * in_px_bytes + 524288.0 * esz), n * 2.0 * 32768 * 27 * 13, s, [&] {
      launch_initial<T>(x, kind, n, X, c->initial_pool, c->d_init_w, c->d_init_g, c->d_init_b, c->d_init_a, c->d_lut32, s);
    });
  }
  int H = 128, W = 256;   // resolution of X
  if (stop_after == -1) { launch_export_nchw<T>(X, dump, n, 16, H, W, s); return BC_OK; }
  int block_index = 0;
  // direction in which the tensor now in X was written; the next kernel walks the other way.  Every kernel
  // after the initial block can walk backwards; the CUDA-core fall-backs (fp32 mode, tensor cores off) ignore
  // the flag, which only costs the L2 hits
  static const bool snake = getenv("BC_NO_SNAKE") == nullptr, asym_fusion = getenv("BC_NO_ASYM_FUSION") == nullptr;
  int x_dir = 0;
  auto flip_dir = [&]() { x_dir = snake ? !x_dir : 0; g_umma_reverse = x_dir; };
  bool e1_ready = false;    // E1 already holds this block's projection (written by the previous tcgen05 kernel)
  const Taps t1 = taps_for(1, 1, 1);
  // conv launch with its algorithmic traffic: input + output (+ residual) activations
  auto conv = [&](const char* name, const T* in, T* out, const T* res, int res_ch, const ConvP& p,
                  const float* alpha_out, const Taps& taps) {
    double px = (double)n * H * W;
    L(c, name, px * (p.cin + p.cout + (res ? res_ch : 0)) * esz, 2.0 * px * p.cin * p.cout * p.ntaps, s,
      [&] { launch_conv<T>(in, out, res, res_ch, p, alpha_out, n, H, W, taps, s); });
  };
  for (const Bottleneck& b : c->blocks) {
    if (b.kind == 0) {
      H /= 2; W /= 2;
      uint8_t* idx = b.cin == 16 ? c->idx1 : c->idx2;
      double px = (double)n * H * W;
      bool tc = false;
      if constexpr (is16<T>::value) tc = c->tensor_cores && c->umma_ready && b.um_a.wblob != nullptr;
      const int cip = tc ? b.um_a.CI : b.ci;       // e1 width as stored (zero-padded to 16 for tcgen05)
      if (tc) {
        if constexpr (is16<T>::value) {
          cudaError_t ce = cudaSuccess;
          L(c, b.cin == 16 ? "umma_pool_conv16" : "umma_pool_conv64", px * (4.0 * b.cin * esz + b.cin * esz + b.cin + cip * esz),
            2.0 * px * 4 * b.cin * b.ci, s,
            [&] { flip_dir(); ce = Umma<T>::launch_down(b.um_b, X, P, idx, E1, n, H, W, c->num_sms, s); });
          if (ce != cudaSuccess) return fail(c, BC_ERR_CUDA, std::string("tcgen05 pooling launch: ") + cudaGetErrorString(ce));
          L(c, b.cout == 64 ? "umma_down64" : "umma_down128", px * (cip + b.cin + b.cout + b.um_a.CN) * esz,
            2.0 * px * (9.0 * b.ci * b.ci + (double)b.ci * b.cout + (double)b.cout * b.um_a.CN), s,
            [&] { flip_dir();                 // against the pooling kernel above, which wrote P / E1
                  ce = Umma<T>::launch(b.um_a, E1, P, Y, E2, n, H, W, taps_for(3, 3, 1), 0, 1, c->num_sms, s); });
          if (ce != cudaSuccess) return fail(c, BC_ERR_CUDA, std::string("tcgen05 down-sampling launch: ") + cudaGetErrorString(ce));
          std::swap(E1, E2);          // e1' of the next block was written to E2
          e1_ready = true;
        }
      } else {
        e1_ready = false;
        L(c, "down_pool_conv2x2", px * (4.0 * b.cin * esz + b.cin * esz + b.cin + b.ci * esz), 2.0 * px * 4 * b.cin * b.ci, s,
          [&] { launch_down_a<T>(X, n, H, W, b.cin, b.ci, P, idx, E1, b.c1, s); });
        conv("down_conv3x3", E1, E2, nullptr, 0, b.c2, nullptr, taps_for(3, 3, 1));
        conv("down_expand_add", E2, Y, P, b.cin, b.c3, b.alpha_out, t1);
      }
      std::swap(X, Y);
    } else if (b.kind == 1 || b.kind == 2) {
      bool done = false;
      if (c->tensor_cores && c->umma_ready && umma_supported(b)) {
        if constexpr (is16<T>::value) {
          const double px = (double)n * H * W;
          if (!e1_ready) conv("proj1x1", X, E1, nullptr, 0, b.c1, nullptr, t1);
          cudaError_t ce = cudaSuccess;
          const int has_next = (b.kind == 1 ? b.um_a : b.um_b).has_next ? 1 : 0;
          const double io = px * (2.0 * b.cin + b.ci * (1 + has_next)) * esz;
          flip_dir();                   // both launches of an asymmetric block walk the same way
          if (b.kind == 1) {
            L(c, b.cin == 64 ? "umma_bottleneck64" : "umma_bottleneck128", io,
              2.0 * px * (9.0 * b.ci * b.ci + b.ci * b.cin * (1 + has_next)), s,
              [&] { ce = Umma<T>::launch(b.um_a, E1, X, Y, E2, n, H, W, taps_for(3, 3, b.dilation), 0, has_next, c->num_sms, s); });
            std::swap(E1, E2);        // e1' of the next block was written to E2
          } else if (b.um_f.wblob && W == 64 && asym_fusion) {
            // the whole asymmetric block in one launch: the 5x1 result stays in shared memory
            L(c, "umma_asym_fused", io, 2.0 * px * (10.0 * b.ci * b.ci + b.ci * b.cin * (1 + has_next)), s,
              [&] { ce = Umma<T>::launch_asym(b.um_f, E1, X, Y, E2, n, H, W, taps_for(5, 1, 1), has_next, c->num_sms, s); });
            std::swap(E1, E2);        // e1' of the next block was written to E2
          } else {
            L(c, "umma_conv5x1", px * 2.0 * b.ci * esz, 2.0 * px * 5.0 * b.ci * b.ci, s,
              [&] { ce = Umma<T>::launch(b.um_a, E1, nullptr, nullptr, E2, n, H, W, taps_for(5, 1, 1), 1, 0, c->num_sms, s); });
            if (ce == cudaSuccess)
              L(c, "umma_bottleneck128_asym", io, 2.0 * px * (5.0 * b.ci * b.ci + b.ci * b.cin * (1 + has_next)), s,
                [&] { ce = Umma<T>::launch(b.um_b, E2, X, Y, E1, n, H, W, taps_for(1, 5, 1), 0, has_next, c->num_sms, s); });
          }
          if (ce != cudaSuccess) return fail(c, BC_ERR_CUDA, std::string("tcgen05 bottleneck launch: ") + cudaGetErrorString(ce));
          e1_ready = has_next != 0;
          done = true;
        }
      }
      if (!done && b.kind == 1 && b.cin == 16 && b.ci == 4 && b.dilation == 1) {
        // stage 5: K = 4 is too skinny for tcgen05; one fused CUDA-core kernel, x in, y out
        const double px = (double)n * H * W;
        L(c, "stage5_bottleneck", px * 2.0 * b.cin * esz, 2.0 * px * (16.0 * 4 + 9.0 * 16 + 4.0 * 16), s,
          [&] { flip_dir(); launch_stage5<T>(X, Y, b, n, H, W, s); });
        e1_ready = false;
        done = true;
      }
      if (!done) {
        e1_ready = false;
        conv("proj1x1", X, E1, nullptr, 0, b.c1, nullptr, t1);
        if (b.kind == 1) {
          conv(b.dilation > 1 ? "conv3x3_dilated" : "conv3x3", E1, E2, nullptr, 0, b.c2, nullptr, taps_for(3, 3, b.dilation));
          conv("expand1x1_add", E2, Y, X, b.cin, b.c3, b.alpha_out, t1);
        } else {
          conv("conv5x1", E1, E2, nullptr, 0, b.c2, nullptr, taps_for(5, 1, 1));
          conv("conv1x5", E2, E1, nullptr, 0, b.c2b, nullptr, taps_for(1, 5, 1));
          conv("expand1x1_add", E1, Y, X, b.cin, b.c3, b.alpha_out, t1);
        }
      }
      std::swap(X, Y);
    } else {
      const uint8_t* idx = b.cout == 16 ? c->idx1 : c->idx2;   // upsample5_0 pairs with downsample1_0
      double px = (double)n * H * W;
      bool done = false;
      if (c->tensor_cores && c->umma_ready && b.um_a.wblob) {
        if constexpr (is16<T>::value) {
          cudaError_t ce = cudaSuccess;
          const int has_next = b.um_a.has_next ? 1 : 0;
          L(c, b.cout == 64 ? "umma_up4" : "umma_up5", px * (b.cin * esz + b.cout + 4.0 * b.cout * esz + has_next * 4.0 * 16 * esz),
            2.0 * px * ((double)b.cin * (b.cout + b.ci) + 4.0 * b.ci * b.ci + 4.0 * b.ci * b.cout + has_next * 4.0 * 64 * 16), s,
            [&] { flip_dir();
                  ce = Umma<T>::launch_up(b.um_a, b.cin, b.cout, X, idx, Y, E1, n, H, W, has_next, c->num_sms, s); });
          if (ce != cudaSuccess) return fail(c, BC_ERR_CUDA, std::string("tcgen05 upsampling launch: ") + cudaGetErrorString(ce));
          e1_ready = has_next != 0;
          done = true;
        }
      }
      if (!done) {
      e1_ready = false;
      conv("up_proj1x1", X, E1, nullptr, 0, b.c1, nullptr, t1);
      L(c, "up_unpool_tconv_expand", px * ((b.cin + b.ci) * esz + b.cout + 4.0 * b.cout * esz),
        2.0 * px * ((double)b.cin * b.cout + 4.0 * b.ci * b.ci + 4.0 * b.ci * b.cout), s,
        [&] { launch_up_b<T>(X, E1, idx, Y, b, n, H, W, s); });
      }
      H *= 2; W *= 2;
      std::swap(X, Y);
    }
    if (block_index++ == stop_after) { launch_export_nchw<T>(X, dump, n, b.cout, H, W, s); return BC_OK; }
  }
  if (labels && c->tensor_cores && c->umma_ready && c->d_head_umma) {
    if constexpr (is16<T>::value) {
      cudaError_t ce = cudaSuccess;
      L(c, "umma_head_argmax_lut", n * (32768.0 * 16 * esz + 131072.0), n * 2.0 * 32768 * 64 * 64, s,
        [&] { flip_dir(); ce = Umma<T>::launch_head(X, n, c->num_classes, c->d_head_umma, labels, *lut, c->num_sms, s); });
      if (ce != cudaSuccess) return fail(c, BC_ERR_CUDA, std::string("tcgen05 head launch: ") + cudaGetErrorString(ce));
      return BC_OK;
    }
  }
  L(c, labels ? "head_tconv_argmax_lut" : "head_tconv_logits",
    n * (32768.0 * 16 * esz + (labels ? 131072.0 : 131072.0 * 4 * c->num_classes)), n * 2.0 * 32768 * 9 * 16 * c->num_classes, s,
    [&] { launch_fullconv<T>(X, n, c->num_classes, c->d_full_w, logits, labels, lut, s); });
  return BC_OK;
}

size_t input_frame_bytes(int kind) {
  const size_t px = (size_t)BC_NET_H * BC_NET_W * 3;
  return kind == BC_IN_BGR_U8 ? px : kind == BC_IN_NCHW_F32 ? px * 4 : px * 8;
}

int forward(bc_ctx* c, const void* x, int kind, int B, float* logits, uint8_t* labels, const Lut256* lut,
            cudaStream_t s) {
  if (!c->net_loaded) return fail(c, BC_ERR_STATE, "bc_load_enet has not been called");
  if (B < 1 || B > c->max_batch) return fail(c, BC_ERR_ARG, "batch size outside [1, max_batch]");
  if (kind < 0 || kind > 2) return fail(c, BC_ERR_ARG, "unknown input kind");
  if (!x) return fail(c, BC_ERR_ARG, "null input pointer");
  int r;
  if ((r = ensure_scratch(c))) return r;
  const int ch = c->scratch_frames;
  const size_t fb = input_frame_bytes(kind);
  const size_t px = (size_t)BC_NET_H * BC_NET_W;
  for (int f0 = 0; f0 < B; f0 += ch) {
    int n = std::min(ch, B - f0);
    const void* xi = (const uint8_t*)x + fb * f0;
    float* lo = logits ? logits + (size_t)f0 * c->num_classes * px : nullptr;
    uint8_t* la = labels ? labels + (size_t)f0 * px : nullptr;
    if (c->precision == BC_PREC_FP16) r = forward_chunk<f16>(c, xi, kind, n, lo, la, lut, s);
    else if (c->precision == BC_PREC_BF16) r = forward_chunk<bf16>(c, xi, kind, n, lo, la, lut, s);
    else r = forward_chunk<float>(c, xi, kind, n, lo, la, lut, s);
    if (r) return r;
  }
  return check_launch(c, "ENet forward");
}

// ---------------------------------------------------------------------- resize tables
// cv::resize INTER_LINEAR, 8-bit: per-axis tap pair + 11-bit coefficients (resize.cpp;
// oracle/cv_ops.py _linear_axis_tables).  x zeroes the fraction at the ends, y only clamps.
void axis_table(int dst_n, int src_n, bool clamp_frac, int* s0, int* s1, int* c0, int* c1) {
  double scale = (double)src_n / (double)dst_n;
  for (int d = 0; d < dst_n; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= (float)s;
    int a, b;
    if (clamp_frac) {
      if (s < 0) { f = 0.f; s = 0; }
      if (s >= src_n - 1) { f = 0.f; s = src_n - 1; }
      a = s; b = std::min(s + 1, src_n - 1);
    } else {
      a = std::min(std::max(s, 0), src_n - 1);
      b = std::min(std::max(s + 1, 0), src_n - 1);
    }
    s0[d] = a; s1[d] = b;
    c0[d] = (int)std::nearbyint((1.f - f) * 2048.f);
    c1[d] = (int)std::nearbyint(f * 2048.f);
  }
}

int get_resize_tab(bc_ctx* c, int h, int w, const ResizeTab** out) {
  auto key = std::make_pair(h, w);
  auto it = c->resize_tabs.find(key);
  if (it != c->resize_tabs.end()) { *out = &it->second; return BC_OK; }
  ResizeTab t;
  t.src_h = h; t.src_w = w;
  if (h == BC_NET_H && w == BC_NET_W) t.mode = 0;
  else if (h == 2 * BC_NET_H && w == 2 * BC_NET_W) t.mode = 1;
  else {
    t.mode = 2;
    std::vector<int> host(4 * BC_NET_W + 4 * BC_NET_H);
    int* x0 = host.data(); int* x1 = x0 + BC_NET_W; int* a0 = x1 + BC_NET_W; int* a1 = a0 + BC_NET_W;
    int* y0 = a1 + BC_NET_W; int* y1 = y0 + BC_NET_H; int* b0 = y1 + BC_NET_H; int* b1 = b0 + BC_NET_H;
    axis_table(BC_NET_W, w, true, x0, x1, a0, a1);
    axis_table(BC_NET_H, h, false, y0, y1, b0, b1);
    CU(cudaMalloc(&t.blob, host.size() * sizeof(int)));
    CU(cudaMemcpy(t.blob, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice));
    int* d = (int*)t.blob;
    t.x0 = d; t.x1 = d + BC_NET_W; t.a0 = d + 2 * BC_NET_W; t.a1 = d + 3 * BC_NET_W;
    d += 4 * BC_NET_W;
    t.y0 = d; t.y1 = d + BC_NET_H; t.b0 = d + 2 * BC_NET_H; t.b1 = d + 3 * BC_NET_H;
  }
  auto ins = c->resize_tabs.emplace(key, t);
  *out = &ins.first->second;
  return BC_OK;
}

// ------------------------------------------------------------------------ BEV geometry
// cv::invert for 3x3 (closed form, adjugate * 1/det); oracle/cv_ops.py invert3x3
void invert3x3(const double* a, double* t) {
  double d = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
  if (d == 0.0) { for (int i = 0; i < 9; ++i) t[i] = 0.0; return; }
  d = 1.0 / d;
  t[0] = (a[4] * a[8] - a[5] * a[7]) * d;
  t[1] = (a[2] * a[7] - a[1] * a[8]) * d;
  t[2] = (a[1] * a[5] - a[2] * a[4]) * d;
  t[3] = (a[5] * a[6] - a[3] * a[8]) * d;
  t[4] = (a[0] * a[8] - a[2] * a[6]) * d;
  t[5] = (a[2] * a[3] - a[0] * a[5]) * d;
  t[6] = (a[3] * a[7] - a[4] * a[6]) * d;
  t[7] = (a[1] * a[6] - a[0] * a[7]) * d;
  t[8] = (a[0] * a[4] - a[1] * a[3]) * d;
}

// bev.py:172-176, 183-190 (float -> int() truncations; Python float == C double)
void free_occ_tables(bc_ctx* c) {
  for (auto& kv : c->occ_tables) cudaFree(kv.second);
  c->occ_tables.clear();
}

// Geometry of one grid request; with `table` also the (cached) device table of K9's
// frame-independent coordinates.  May allocate and synchronise on first use of a geometry,
// so it runs before any graph capture.
int make_geom(bc_ctx* c, double w_m, double h_m, double cell_m, int binary, int ros, BevGeom& g,
              bool with_table = true) {
  memset(&g, 0, sizeof g);                // the struct doubles as a cache key (padding included)
  if (!c->bev_set) return fail(c, BC_ERR_STATE, "bc_set_bev has not been called");
  if (!(cell_m > 0.0) || !(w_m > 0.0) || !(h_m > 0.0)) return fail(c, BC_ERR_ARG, "grid extents and cell size must be positive");
  double cell_px = cell_m * 100 / c->cm_per_px;
  int Wc = (int)(w_m / cell_m);
  int occ_w_px = (int)(Wc * cell_px);
  int Hc = (int)(h_m / cell_m);
  int occ_h_px = (int)(Hc * cell_px);
  if (Wc < 1 || Hc < 1 || occ_w_px < 1 || occ_h_px < 1) return fail(c, BC_ERR_ARG, "empty occupancy grid");
  int left_x = (int)((c->warp_w - occ_w_px) / 2.0);
  int top_y = c->warp_h - occ_h_px;
  memcpy(g.Mi, c->Mi, sizeof g.Mi);
  int bh0 = std::min(16, c->warp_h);
  g.bw0 = std::min(1024 / bh0, c->warp_w);
  g.in_rows = c->in_rows; g.in_cols = c->in_cols;
  g.warp_w = c->warp_w; g.warp_h = c->warp_h;
  g.occ_w_px = occ_w_px; g.occ_h_px = occ_h_px;
  g.wl = std::max(left_x, 0); g.wt = std::max(top_y, 0);
  g.gl = std::max(-left_x, 0); g.gt = std::max(-top_y, 0);
  g.crop_w = std::max(std::min(g.wl + occ_w_px, c->warp_w) - g.wl, 0);
  g.crop_h = std::max(c->warp_h - g.wt, 0);
  g.Wc = Wc; g.Hc = Hc;
  g.binary = binary ? 1 : 0; g.ros_layout = ros ? 1 : 0;
  // cv::resize INTER_NEAREST: sx = min(floor(dx * (1 / (dst / src))), src - 1) in fp64
  g.ifx = 1.0 / ((double)Wc / (double)occ_w_px);
  g.ify = 1.0 / ((double)Hc / (double)occ_h_px);
  if (!with_table) return BC_OK;
  BevGeom key = g;
  key.binary = key.ros_layout = key.raw_template = 0;        // the coordinates do not depend on them
  for (auto& kv : c->occ_tables)
    if (memcmp(&kv.first, &key, sizeof key) == 0) { g.table = kv.second; return BC_OK; }
  CU(cudaSetDevice(c->device));
  uint2* d = nullptr;
  CU(cudaMalloc(&d, occ_table_bytes(Hc * Wc)));
  launch_occ_table(key, d, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(d); return fail(c, BC_ERR_CUDA, std::string("occupancy table: ") + cudaGetErrorString(e)); }
  if (c->occ_tables.size() >= 8) {
    // cached graphs hold the evicted table's address (GraphKey does not): drop them with it.  The synchronise
    // above covered every kernel that could still read the table.
    invalidate_graphs(c);
    cudaFree(c->occ_tables.front().second);
    c->occ_tables.erase(c->occ_tables.begin());
  }
  c->occ_tables.emplace_back(key, d);
  g.table = d;
  return BC_OK;
}

int8_t* grid_dest(bc_ctx* c, int8_t* d_grids, int B, const BevGeom& g) {
  if (d_grids) return d_grids;
  if (!c->gather_base) return nullptr;
  return c->gather_base + (size_t)c->rank * B * g.Hc * g.Wc;
}

int ensure_contour_scratch(bc_ctx* c, int B, int H, int W) {
  size_t need = contour_scratch_bytes(B, H, W);
  if (need <= c->cn_scratch_bytes) return BC_OK;
  if (c->cn_scratch) {
    // pipeline graphs captured with the contour filter on hold the old scratch address
    cudaDeviceSynchronize();
    invalidate_graphs(c);
    cudaFree(c->cn_scratch);
    c->cn_scratch = nullptr; c->cn_scratch_bytes = 0;
  }
  if (cudaMalloc(&c->cn_scratch, need) != cudaSuccess) {
    cudaGetLastError();
    return fail(c, BC_ERR_NOMEM, "contour_noise_removal scratch: out of device memory");
  }
  c->cn_scratch_bytes = need;
  return BC_OK;
}

// 8 kernels (contour.cu); algorithmic traffic = mask in + mask out
void run_contour(bc_ctx* c, const uint8_t* seg, int H, int W, int B, uint8_t* out, cudaStream_t s) {
  L(c, "contour_noise_removal", 2.0 * B * H * W, 0, s,
    [&] { launch_contour_noise_removal(seg, H, W, B, out, c->cn_scratch, s); });
  c->launches += contour_launch_count() - 1;
}

Lut256 make_lut(const uint8_t* h_lut) {
  Lut256 l;
  memcpy(l.v, h_lut, 256);
  return l;
}

int do_pipeline(bc_ctx* c, const uint8_t* d_bgr, int h, int w, int B, const uint8_t* h_lut, const BevGeom& g,
                uint8_t* d_labels_out, int8_t* d_grids, cudaStream_t s) {
  int r;
  const ResizeTab* rt;
  if ((r = get_resize_tab(c, h, w, &rt))) return r;
  const uint8_t* frames = d_bgr;
  if (rt->mode != 0) {
    L(c, "resize", (double)B * (3.0 * h * w + 393216.0), 0, s, [&] { launch_resize(d_bgr, h, w, B, c->d_resized, *rt, s); });
    frames = c->d_resized;
  }
  uint8_t* labels = d_labels_out ? d_labels_out : c->d_labels;
  Lut256 lut = make_lut(h_lut);
  if ((r = forward(c, frames, BC_IN_BGR_U8, B, nullptr, labels, &lut, s))) return r;
  if (g.binary && c->contour_filter) {           // predict_binary -> contour_noise_removal -> grid
    run_contour(c, labels, BC_NET_H, BC_NET_W, B, c->d_labels_cn, s);
    labels = c->d_labels_cn;
  }
  L(c, "occgrid", (double)B * ((double)g.in_rows * g.in_cols + (double)g.Hc * g.Wc), 0, s,
    [&] { launch_occgrid(labels, B, g, d_grids, s); });
  return check_launch(c, "pipeline");
}

// The device part of the path (resize .. occupancy grid) replayed as ONE CUDA graph per
// distinct argument set: a per-frame driver loop then costs a single launch (batch 1 is
// launch-bound otherwise: ~75 kernels of a few microseconds each).
int run_pipeline(bc_ctx* c, const uint8_t* d_bgr, int h, int w, int B, const uint8_t* h_lut, double w_m, double h_m,
                 double cell_m, const BevGeom& g, uint8_t* d_labels_out, int8_t* d_grids, cudaStream_t s) {
  int r;
  const ResizeTab* rt;
  if ((r = get_resize_tab(c, h, w, &rt))) return r;   // may allocate: keep it out of capture
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (s) cudaStreamIsCapturing(s, &st);
  if (!c->use_graphs || c->profiling || st != cudaStreamCaptureStatusNone)
    return do_pipeline(c, d_bgr, h, w, B, h_lut, g, d_labels_out, d_grids, s);
  GraphKey key;
  memset(&key, 0, sizeof key);
  key.in = d_bgr; key.out = d_grids; key.labels = d_labels_out;
  key.h = h; key.w = w; key.B = B; key.binary = g.binary; key.ros = g.ros_layout;
  key.pad_ = c->contour_filter;
  key.w_m = w_m; key.h_m = h_m; key.cell_m = cell_m;
  memcpy(key.lut, h_lut, 256);
  GraphEntry* ge = nullptr;
  for (auto& e : c->graphs) if (e.key == key) { ge = &e; break; }
  if (!ge) {
    cudaStream_t cs;
    CU(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    long long l0 = c->launches;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      r = do_pipeline(c, d_bgr, h, w, B, h_lut, g, d_labels_out, d_grids, cs);
      e = cudaStreamEndCapture(cs, &graph);
    }
    long long nl = c->launches - l0;
    c->launches = l0;
    cudaStreamDestroy(cs);
    if (r) { if (graph) cudaGraphDestroy(graph); return r; }
    if (e != cudaSuccess) return fail(c, BC_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(c, BC_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
    if (c->graphs.size() >= 16) { cudaGraphExecDestroy(c->graphs.front().exec); c->graphs.erase(c->graphs.begin()); }
    GraphEntry ne;
    ne.key = key; ne.exec = exec; ne.launches = nl;
    c->graphs.push_back(ne);
    ge = &c->graphs.back();
  }
  CU(cudaGraphLaunch(ge->exec, s));
  c->launches += ge->launches;
  return BC_OK;
}

}  // namespace

// =============================================================================== ABI
extern "C" {

int bc_abi_version(void) { return 1; }

const char* bc_last_error(const bc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int bc_create(bc_ctx** out, int device, int max_batch) {
  bc_ctx* c = nullptr;
  if (!out) return fail(c, BC_ERR_ARG, "null output pointer");
  *out = nullptr;
  if (max_batch < 1 || max_batch > 65536) return fail(c, BC_ERR_ARG, "max_batch must be in [1, 65536]");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(c, BC_ERR_CUDA, std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                                    " (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(c, BC_ERR_ARG, "device index out of range");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(c, BC_ERR_CUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(c, BC_ERR_CUDA, std::string("device ") + prop.name + " is not sm_100 (Blackwell B200); this library carries sm_100a code only");
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(c, BC_ERR_CUDA, cudaGetErrorString(e));
  std::unique_ptr<bc_ctx, void (*)(bc_ctx*)> ctx(new bc_ctx(), bc_destroy);   // a failed create releases what it made
  c = ctx.get();
  c->device = device;
  c->max_batch = max_batch;
  c->num_sms = prop.multiProcessorCount;
  // kernels that need more than 48 KB of dynamic shared memory: the opt-in is a per-device function
  // attribute, so every context sets it for its own GPU (one process may hold contexts on several GPUs)
  CU(prepare_occgrid());
  CU(Umma<bf16>::prepare_bottleneck()); CU(Umma<f16>::prepare_bottleneck());
  CU(Umma<bf16>::prepare_initial());    CU(Umma<f16>::prepare_initial());
  CU(Umma<bf16>::prepare_down());       CU(Umma<f16>::prepare_down());
  CU(Umma<bf16>::prepare_up());         CU(Umma<f16>::prepare_up());
  CU(Umma<bf16>::prepare_head());       CU(Umma<f16>::prepare_head());
  // streams and events of the host entry points
  CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  for (auto& e : c->copy_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->call_start, cudaEventDisableTiming));
  for (auto& sl : c->slots) {
    CU(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&sl.computed, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  }
  // normalisation LUT, fp64 exactly as numpy evaluates (rgb / 256.0 - mean) / std (models.py:91)
  const double mean[3] = {0.485, 0.456, 0.406}, sd[3] = {0.229, 0.224, 0.225};   // models.py:17-18
  std::vector<double> l64(768);
  std::vector<float> l32(768);
  for (int u = 0; u < 256; ++u)
    for (int ch = 0; ch < 3; ++ch) {
      l64[u * 3 + ch] = ((double)u / 256.0 - mean[ch]) / sd[ch];
      l32[u * 3 + ch] = (float)l64[u * 3 + ch];   // TensorFlow's fp64 -> fp32 feed cast
    }
  CU(cudaMalloc(&c->d_lut64, 768 * sizeof(double)));
  CU(cudaMalloc(&c->d_lut32, 768 * sizeof(float)));
  CU(cudaMemcpy(c->d_lut64, l64.data(), 768 * sizeof(double), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(c->d_lut32, l32.data(), 768 * sizeof(float), cudaMemcpyHostToDevice));
  const size_t px = (size_t)BC_NET_H * BC_NET_W;
  CU(cudaMalloc(&c->d_labels, px * max_batch));
  CU(cudaMalloc(&c->d_resized, px * 3 * max_batch));
  *out = ctx.release();
  return BC_OK;
}

void bc_destroy(bc_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  invalidate_graphs(c);
  free_net(c);
  free_scratch(c);
  for (auto& kv : c->resize_tabs) if (kv.second.blob) cudaFree(kv.second.blob);
  free_occ_tables(c);
  for (auto& r : c->prof) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  for (auto& sl : c->slots) {
    if (sl.d_in) cudaFree(sl.d_in);
    if (sl.d_out) cudaFree(sl.d_out);
    for (cudaEvent_t e : {sl.copied, sl.computed, sl.done}) if (e) cudaEventDestroy(e);
  }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  for (auto e : c->copy_done) if (e) cudaEventDestroy(e);
  if (c->call_start) cudaEventDestroy(c->call_start);
  void* ps[] = {c->d_lut32, c->d_lut64, c->d_labels, c->d_resized, c->d_frames_in, c->d_grids_out, c->cn_scratch,
                c->d_labels_cn, c->laser_cells, c->laser_first, c->gs.d_release_peers};
  if (c->gs.h_err) cudaFreeHost(c->gs.h_err);
  for (auto& t : c->laser_tabs) { cudaFree(t.d_fwd); cudaFree(t.d_inv); }
  for (void* p : ps) if (p) cudaFree(p);
  delete c;
}

int bc_load_enet(bc_ctx* c, const void* h_blob, size_t n_bytes) {
  if (!c) return BC_ERR_ARG;
  if (!h_blob) return fail(c, BC_ERR_ARG, "null weight blob");
  CU(cudaSetDevice(c->device));
  Container ct;
  int r;
  c->net_loaded = false;
  if ((r = parse_container(c, h_blob, n_bytes, ct))) return r;
  if ((r = build_host_net(c, ct))) return r;
  if ((r = upload_net(c))) return r;
  c->net_loaded = true;
  return BC_OK;
}

int bc_num_classes(const bc_ctx* c) { return (c && c->net_loaded) ? c->num_classes : BC_ERR_STATE; }

int bc_set_precision(bc_ctx* c, int precision) {
  if (!c) return BC_ERR_ARG;
  if (precision != BC_PREC_BF16 && precision != BC_PREC_FP32 && precision != BC_PREC_FP16)
    return fail(c, BC_ERR_ARG, "unknown precision");
  if (precision == c->precision) return BC_OK;
  c->precision = precision;
  if (c->net_loaded) {
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    c->net_loaded = false;          // a failed upload leaves no half-built network behind
    int r = upload_net(c);
    if (r) return r;
    c->net_loaded = true;
  }
  return BC_OK;
}

int bc_set_chunk(bc_ctx* c, int frames) {
  if (!c) return BC_ERR_ARG;
  if (frames < 0) return fail(c, BC_ERR_ARG, "chunk must be >= 0");
  if (frames != c->chunk) { CU(cudaDeviceSynchronize()); invalidate_graphs(c); }
  c->chunk = frames;
  return BC_OK;
}

int bc_set_tensor_cores(bc_ctx* c, int enable) {
  if (!c) return BC_ERR_ARG;
  if ((enable != 0) != (c->tensor_cores != 0)) { CU(cudaDeviceSynchronize()); invalidate_graphs(c); }
  c->tensor_cores = enable ? 1 : 0;
  return BC_OK;
}

int bc_set_graphs(bc_ctx* c, int enable) {
  if (!c) return BC_ERR_ARG;
  c->use_graphs = enable ? 1 : 0;
  return BC_OK;
}

int bc_set_host_overlap(bc_ctx* c, int enable) {
  if (!c) return BC_ERR_ARG;
  c->host_overlap = enable ? 1 : 0;
  return BC_OK;
}

int bc_set_bev(bc_ctx* c, const double h_M[9], int in_rows, int in_cols, int warp_w, int warp_h, double cm_per_px) {
  if (!c) return BC_ERR_ARG;
  if (!h_M) return fail(c, BC_ERR_ARG, "null matrix");
  if (in_rows < 2 || in_cols < 2 || in_rows > 32767 || in_cols > 32767)          // K9 samples fixed 2 x 2 blocks of label pixels
    return fail(c, BC_ERR_ARG, "label-map size out of range (2..32767 per side)");
  if (warp_w < 1 || warp_h < 1 || warp_w > 32767 || warp_h > 32767) return fail(c, BC_ERR_ARG, "warped size out of range");
  if (!(cm_per_px > 0.0)) return fail(c, BC_ERR_ARG, "cm_per_px must be positive");
  CU(cudaSetDevice(c->device));
  CU(cudaDeviceSynchronize());
  invalidate_graphs(c);
  free_occ_tables(c);
  memcpy(c->M, h_M, sizeof c->M);
  invert3x3(c->M, c->Mi);
  c->in_rows = in_rows; c->in_cols = in_cols;
  c->warp_w = warp_w; c->warp_h = warp_h;
  c->cm_per_px = cm_per_px;
  c->bev_set = true;
  return BC_OK;
}

int bc_resize_bgr(bc_ctx* c, const uint8_t* d_src, int h, int w, int B, uint8_t* d_dst, void* stream) {
  if (!c) return BC_ERR_ARG;
  if (!d_src || !d_dst) return fail(c, BC_ERR_ARG, "null pointer");
  if (h < 1 || w < 1 || h > 16384 || w > 16384 || B < 1) return fail(c, BC_ERR_ARG, "bad frame shape");
  CU(cudaSetDevice(c->device));
  const ResizeTab* rt;
  int r;
  if ((r = get_resize_tab(c, h, w, &rt))) return r;
  cudaStream_t s = (cudaStream_t)stream;
  if (rt->mode != 0)
    L(c, "resize", (double)B * (3.0 * h * w + 393216.0), 0, s, [&] { launch_resize(d_src, h, w, B, d_dst, *rt, s); });
  else launch_resize(d_src, h, w, B, d_dst, *rt, s);   // same size: a device-to-device copy, no kernel
  return check_launch(c, "resize");
}

int bc_preprocess(bc_ctx* c, const uint8_t* d_bgr, int h, int w, int B, void* d_out, int out_f64, void* stream) {
  if (!c) return BC_ERR_ARG;
  if (!d_bgr || !d_out) return fail(c, BC_ERR_ARG, "null pointer");
  if (h < 1 || w < 1 || h > 16384 || w > 16384) return fail(c, BC_ERR_ARG, "bad frame shape");
  if (B < 1 || B > c->max_batch) return fail(c, BC_ERR_ARG, "batch size outside [1, max_batch]");
  CU(cudaSetDevice(c->device));
  const ResizeTab* rt;
  int r;
  if ((r = get_resize_tab(c, h, w, &rt))) return r;
  const uint8_t* src = d_bgr;
  if (rt->mode != 0) {
    L(c, "resize", (double)B * (3.0 * h * w + 393216.0), 0, (cudaStream_t)stream,
      [&] { launch_resize(d_bgr, h, w, B, c->d_resized, *rt, (cudaStream_t)stream); });
    src = c->d_resized;
  }
  L(c, "preprocess", (double)B * 393216.0 * (1 + (out_f64 ? 8 : 4)), 0, (cudaStream_t)stream,
    [&] { launch_preprocess(src, B, d_out, out_f64, c->d_lut64, (cudaStream_t)stream); });
  return check_launch(c, "preprocess");
}

int bc_enet_logits(bc_ctx* c, const void* d_x, int kind, int B, float* d_logits, void* stream) {
  if (!c) return BC_ERR_ARG;
  if (!d_logits) return fail(c, BC_ERR_ARG, "null logits pointer");
  CU(cudaSetDevice(c->device));
  return forward(c, d_x, kind, B, d_logits, nullptr, nullptr, (cudaStream_t)stream);
}

int bc_enet_labels(bc_ctx* c, const void* d_x, int kind, int B, const uint8_t h_lut[256], uint8_t* d_labels, void* stream) {
  if (!c) return BC_ERR_ARG;
  if (!d_labels || !h_lut) return fail(c, BC_ERR_ARG, "null pointer");
  CU(cudaSetDevice(c->device));
  Lut256 lut = make_lut(h_lut);
  return forward(c, d_x, kind, B, nullptr, d_labels, &lut, (cudaStream_t)stream);
}

int bc_enet_block_output(bc_ctx* c, const void* d_x, int kind, int B, int block, float* d_out, void* stream) {
  if (!c) return BC_ERR_ARG;
  if (!c->net_loaded) return fail(c, BC_ERR_STATE, "bc_load_enet has not been called");
  if (!d_x || !d_out) return fail(c, BC_ERR_ARG, "null pointer");
  if (block < -1 || block >= (int)c->blocks.size()) return fail(c, BC_ERR_ARG, "block index out of range");
  if (kind < 0 || kind > 2) return fail(c, BC_ERR_ARG, "unknown input kind");
  CU(cudaSetDevice(c->device));
  int r;
  if ((r = ensure_scratch(c))) return r;
  if (B < 1 || B > c->scratch_frames) return fail(c, BC_ERR_ARG, "batch size outside [1, chunk]");
  if (c->precision == BC_PREC_FP16) r = forward_chunk<f16>(c, d_x, kind, B, nullptr, nullptr, nullptr, (cudaStream_t)stream, block, d_out);
  else if (c->precision == BC_PREC_BF16) r = forward_chunk<bf16>(c, d_x, kind, B, nullptr, nullptr, nullptr, (cudaStream_t)stream, block, d_out);
  else r = forward_chunk<float>(c, d_x, kind, B, nullptr, nullptr, nullptr, (cudaStream_t)stream, block, d_out);
  if (r) return r;
  return check_launch(c, "block output");
}

int bc_argmax_lut(bc_ctx* c, const float* d_logits, int B, int C, int H, int W, const uint8_t h_lut[256],
                  uint8_t* d_labels, void* stream) {
  if (!c) return BC_ERR_ARG;
  if (!d_logits || !d_labels || !h_lut) return fail(c, BC_ERR_ARG, "null pointer");
  if (B < 1 || C < 1 || C > 256 || H < 1 || W < 1) return fail(c, BC_ERR_ARG, "bad logits shape");
  if ((long long)B * H * W > 0x7fffffffLL) return fail(c, BC_ERR_ARG, "too many pixels");
  CU(cudaSetDevice(c->device));
  L(c, "argmax_lut", (double)B * H * W * (4.0 * C + 1), 0, (cudaStream_t)stream,
    [&] { launch_argmax_lut(d_logits, B, C, H, W, make_lut(h_lut), d_labels, (cudaStream_t)stream); });
  return check_launch(c, "argmax_lut");
}

int bc_contour_noise_removal(bc_ctx* c, const uint8_t* d_seg, int H, int W, int B, uint8_t* d_out, void* stream) {
  if (!c) return BC_ERR_ARG;
  if (!d_seg || !d_out) return fail(c, BC_ERR_ARG, "null pointer");
  if (std::min(H, W) < 50) return fail(c, BC_ERR_ARG, "contour_noise_removal needs min(h, w) >= 50 (closing kernel int(min/50))");
  if (std::min(H, W) >= 1650) return fail(c, BC_ERR_ARG, "contour_noise_removal supports closing kernels up to 32 x 32 (min(h, w) < 1650)");
  if (B < 1 || B > 65535 || H > 65535 || (long long)B * H * W > 0x7fffffffLL)
    return fail(c, BC_ERR_ARG, "batch size outside [1, 65535] or more than 2^31 pixels");
  CU(cudaSetDevice(c->device));
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (stream) cudaStreamIsCapturing((cudaStream_t)stream, &st);
  if (st != cudaStreamCaptureStatusNone && contour_scratch_bytes(B, H, W) > c->cn_scratch_bytes)
    return fail(c, BC_ERR_STATE, "contour_noise_removal scratch must be sized by one call outside stream capture");
  int r = ensure_contour_scratch(c, B, H, W);
  if (r) return r;
  run_contour(c, d_seg, H, W, B, d_out, (cudaStream_t)stream);
  return check_launch(c, "contour_noise_removal");
}

int bc_set_contour_filter(bc_ctx* c, int enable) {
  if (!c) return BC_ERR_ARG;
  if (enable) {      // size everything for max_batch now: the pipeline may run inside a graph capture later
    CU(cudaSetDevice(c->device));
    int r = ensure_contour_scratch(c, c->max_batch, BC_NET_H, BC_NET_W);
    if (r) return r;
    if (!c->d_labels_cn) CU(cudaMalloc(&c->d_labels_cn, (size_t)c->max_batch * BC_NET_H * BC_NET_W));
  }
  c->contour_filter = enable ? 1 : 0;
  return BC_OK;
}

int bc_occgrid_laserscan(bc_ctx* c, const uint8_t* d_labels, int B, double w_m, double h_m, double cell_m, int binary,
                         int8_t* d_grid_plain, int8_t* d_grid_laser, void* stream) {
  if (!c) return BC_ERR_ARG;
  if (!d_labels || !d_grid_laser) return fail(c, BC_ERR_ARG, "null pointer");
  if (B < 1 || B > 65535) return fail(c, BC_ERR_ARG, "batch size outside [1, 65535]");
  BevGeom g;
  int r = make_geom(c, w_m, h_m, cell_m, binary, 0, g);
  if (r) return r;
  CU(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  // gather tables of this grid shape (host-built once, OpenCV's float arithmetic)
  const bc_ctx::LaserTab* tab = nullptr;
  for (auto& t : c->laser_tabs) if (t.Wc == g.Wc && t.Hc == g.Hc && t.binary == g.binary) tab = &t;
  if (!tab) {
    bc_ctx::LaserTab t{g.Wc, g.Hc, g.binary, 0, 0, nullptr, nullptr};
    laser_polar_size(g.Wc, g.Hc, g.binary, &t.pol_w, &t.pol_h);
    std::vector<int> fwd, inv;
    laser_build_tables(g.Wc, g.Hc, t.pol_w, t.pol_h, fwd, inv);
    CU(cudaMalloc(&t.d_fwd, fwd.size() * sizeof(int)));
    CU(cudaMalloc(&t.d_inv, inv.size() * sizeof(int)));
    CU(cudaMemcpy(t.d_fwd, fwd.data(), fwd.size() * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(t.d_inv, inv.data(), inv.size() * sizeof(int), cudaMemcpyHostToDevice));
    c->laser_tabs.push_back(t);
    tab = &c->laser_tabs.back();
  }
  const size_t cells_bytes = (size_t)B * g.Hc * g.Wc, first_bytes = (size_t)B * tab->pol_h * sizeof(int);
  if (cells_bytes > c->laser_cells_bytes) {
    if (c->laser_cells) cudaFree(c->laser_cells);
    c->laser_cells = nullptr; c->laser_cells_bytes = 0;
    CU(cudaMalloc(&c->laser_cells, cells_bytes));
    c->laser_cells_bytes = cells_bytes;
  }
  if (first_bytes > c->laser_first_bytes) {
    if (c->laser_first) cudaFree(c->laser_first);
    c->laser_first = nullptr; c->laser_first_bytes = 0;
    CU(cudaMalloc(&c->laser_first, first_bytes));
    c->laser_first_bytes = first_bytes;
  }
  // binary: the plain grid is the first element of the reference's tuple (bev.py:164) and the source of the
  // polar search; three-way: the search runs on the resized template (values 0..3, bev.py:209-219)
  uint8_t* cells = c->laser_cells;
  if (binary && d_grid_plain) cells = (uint8_t*)d_grid_plain;
  g.raw_template = binary ? 0 : 1;
  L(c, "occgrid", (double)B * ((double)g.in_rows * g.in_cols + (double)g.Hc * g.Wc), 0, s,
    [&] { launch_occgrid(d_labels, B, g, (int8_t*)cells, s); });
  L(c, "laserscan", (double)B * 2.0 * g.Hc * g.Wc, 0, s,
    [&] { launch_laser(cells, B, g.Wc, g.Hc, tab->pol_w, tab->pol_h, g.binary, tab->d_fwd, tab->d_inv, c->laser_first,
                       d_grid_laser, s); });
  c->launches += 1;
  return check_launch(c, "occgrid_laserscan");
}

int bc_laser_tables(int Wc, int Hc, int binary, int* pol_w, int* pol_h, int* h_fwd, int* h_inv) {
  if (Wc < 1 || Hc < 1 || !pol_w || !pol_h) return BC_ERR_ARG;
  laser_polar_size(Wc, Hc, binary, pol_w, pol_h);
  if (!h_fwd && !h_inv) return BC_OK;                  // size query
  std::vector<int> fwd, inv;
  laser_build_tables(Wc, Hc, *pol_w, *pol_h, fwd, inv);
  if (h_fwd) memcpy(h_fwd, fwd.data(), fwd.size() * sizeof(int));
  if (h_inv) memcpy(h_inv, inv.data(), inv.size() * sizeof(int));
  return BC_OK;
}

int bc_occgrid_shape(bc_ctx* c, double w_m, double h_m, double cell_m, int* Hc, int* Wc) {
  if (!c) return BC_ERR_ARG;
  BevGeom g;
  int r = make_geom(c, w_m, h_m, cell_m, 0, 0, g, false);
  if (r) return r;
  if (Hc) *Hc = g.Hc;
  if (Wc) *Wc = g.Wc;
  return BC_OK;
}

int bc_occgrid(bc_ctx* c, const uint8_t* d_labels, int B, double w_m, double h_m, double cell_m, int binary,
               int ros_layout, int8_t* d_grids, void* stream) {
  if (!c) return BC_ERR_ARG;
  if (!d_labels) return fail(c, BC_ERR_ARG, "null labels pointer");
  if (B < 1 || B > 65535) return fail(c, BC_ERR_ARG, "batch size outside [1, 65535]");
  BevGeom g;
  int r = make_geom(c, w_m, h_m, cell_m, binary, ros_layout, g);
  if (r) return r;
  int8_t* dst = grid_dest(c, d_grids, B, g);
  if (!dst) return fail(c, BC_ERR_ARG, "null grid pointer and no gather buffer set");
  CU(cudaSetDevice(c->device));
  L(c, "occgrid", (double)B * ((double)g.in_rows * g.in_cols + (double)g.Hc * g.Wc), 0, (cudaStream_t)stream,
    [&] { launch_occgrid(d_labels, B, g, dst, (cudaStream_t)stream); });
  return check_launch(c, "occgrid");
}

int bc_pipeline(bc_ctx* c, const uint8_t* d_bgr, int h, int w, int B, const uint8_t h_lut[256], double w_m,
                double h_m, double cell_m, int binary, int ros_layout, uint8_t* d_labels_out, int8_t* d_grids,
                void* stream) {
  if (!c) return BC_ERR_ARG;
  if (!d_bgr || !h_lut) return fail(c, BC_ERR_ARG, "null pointer");
  if (h < 1 || w < 1 || h > 16384 || w > 16384) return fail(c, BC_ERR_ARG, "bad frame shape");
  if (B < 1 || B > c->max_batch) return fail(c, BC_ERR_ARG, "batch size outside [1, max_batch]");
  if (!c->net_loaded) return fail(c, BC_ERR_STATE, "bc_load_enet has not been called");
  BevGeom g;
  int r = make_geom(c, w_m, h_m, cell_m, binary, ros_layout, g);
  if (r) return r;
  if (c->in_rows != BC_NET_H || c->in_cols != BC_NET_W)
    return fail(c, BC_ERR_ARG, "calibration input size must be (256, 512) for the ENet pipeline (bev.py:169)");
  int8_t* dst = grid_dest(c, d_grids, B, g);
  if (!dst) return fail(c, BC_ERR_ARG, "null grid pointer and no gather buffer set");
  CU(cudaSetDevice(c->device));
  int r2;
  if ((r2 = ensure_scratch(c))) return r2;
  return run_pipeline(c, d_bgr, h, w, B, h_lut, w_m, h_m, cell_m, g, d_labels_out, dst, (cudaStream_t)stream);
}

int bc_pipeline_host(bc_ctx* c, const uint8_t* h_bgr, int h, int w, int B, const uint8_t h_lut[256], double w_m,
                     double h_m, double cell_m, int binary, int ros_layout, int8_t* h_grids, void* stream) {
  if (!c) return BC_ERR_ARG;
  if (!h_bgr || !h_grids || !h_lut) return fail(c, BC_ERR_ARG, "null pointer");
  if (h < 1 || w < 1 || h > 16384 || w > 16384) return fail(c, BC_ERR_ARG, "bad frame shape");
  if (B < 1 || B > c->max_batch) return fail(c, BC_ERR_ARG, "batch size outside [1, max_batch]");
  if (!c->net_loaded) return fail(c, BC_ERR_STATE, "bc_load_enet has not been called");
  BevGeom g;
  int r = make_geom(c, w_m, h_m, cell_m, binary, ros_layout, g);
  if (r) return r;
  if (c->in_rows != BC_NET_H || c->in_cols != BC_NET_W)
    return fail(c, BC_ERR_ARG, "calibration input size must be (256, 512) for the ENet pipeline (bev.py:169)");
  CU(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  size_t in_bytes = (size_t)B * h * w * 3;
  size_t out_bytes = (size_t)B * g.Hc * g.Wc;
  if (c->frames_in_bytes < in_bytes) {
    CU(cudaStreamSynchronize(s));
    invalidate_graphs(c);
    if (c->d_frames_in) cudaFree(c->d_frames_in);
    c->d_frames_in = nullptr; c->frames_in_bytes = 0;
    CU(cudaMalloc(&c->d_frames_in, in_bytes));
    c->frames_in_bytes = in_bytes;
  }
  if (c->grids_out_bytes < out_bytes) {
    CU(cudaStreamSynchronize(s));
    invalidate_graphs(c);
    if (c->d_grids_out) cudaFree(c->d_grids_out);
    c->d_grids_out = nullptr; c->grids_out_bytes = 0;
    CU(cudaMalloc(&c->d_grids_out, out_bytes));
    c->grids_out_bytes = out_bytes;
  }
  if ((r = ensure_scratch(c))) return r;
  const ResizeTab* rt;
  if ((r = get_resize_tab(c, h, w, &rt))) return r;   // may allocate: keep it out of capture

  // Large batches go through in sub-batches so that the H2D copy of sub-batch i+1 (copy
  // stream) overlaps the kernels of sub-batch i (caller's stream); grids return as they finish.
  const int nsub = c->host_overlap ? (B >= 128 ? 4 : B >= 32 ? 2 : 1) : 1;
  if (nsub == 1) {
    CU(cudaMemcpyAsync(c->d_frames_in, h_bgr, in_bytes, cudaMemcpyHostToDevice, s));
    if ((r = run_pipeline(c, c->d_frames_in, h, w, B, h_lut, w_m, h_m, cell_m, g, nullptr, c->d_grids_out, s))) return r;
    CU(cudaMemcpyAsync(h_grids, c->d_grids_out, out_bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return BC_OK;
  }
  // the copy stream must not overwrite the staging buffer while earlier work on `s` may still read it
  CU(cudaEventRecord(c->call_start, s));
  CU(cudaStreamWaitEvent(c->copy_stream, c->call_start, 0));
  const size_t frame_bytes = (size_t)h * w * 3, grid_bytes = (size_t)g.Hc * g.Wc;
  int f0 = 0;
  for (int i = 0; i < nsub; ++i) {
    int nb = (B - f0) / (nsub - i);
    CU(cudaMemcpyAsync(c->d_frames_in + f0 * frame_bytes, h_bgr + f0 * frame_bytes, nb * frame_bytes,
                       cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaEventRecord(c->copy_done[i], c->copy_stream));
    f0 += nb;
  }
  f0 = 0;
  for (int i = 0; i < nsub; ++i) {
    int nb = (B - f0) / (nsub - i);
    CU(cudaStreamWaitEvent(s, c->copy_done[i], 0));
    if ((r = run_pipeline(c, c->d_frames_in + f0 * frame_bytes, h, w, nb, h_lut, w_m, h_m, cell_m, g, nullptr,
                          c->d_grids_out + f0 * grid_bytes, s))) return r;
    CU(cudaMemcpyAsync(h_grids + f0 * grid_bytes, c->d_grids_out + f0 * grid_bytes, nb * grid_bytes,
                       cudaMemcpyDeviceToHost, s));
    f0 += nb;
  }
  CU(cudaStreamSynchronize(s));
  return BC_OK;
}

// Streaming form of bc_pipeline_host: returns as soon as the work is enqueued.  Two staging
// slots alternate, so the H2D copy of step i+1 (internal copy stream) overlaps the kernels of
// step i (caller's stream) and the step time is max(copy, compute) instead of their sum.
int bc_pipeline_host_submit(bc_ctx* c, const uint8_t* h_bgr, int h, int w, int B, const uint8_t h_lut[256], double w_m,
                            double h_m, double cell_m, int binary, int ros_layout, int8_t* h_grids, void* stream) {
  if (!c) return BC_ERR_ARG;
  if (!h_bgr || !h_lut) return fail(c, BC_ERR_ARG, "null pointer");
  if (h < 1 || w < 1 || h > 16384 || w > 16384) return fail(c, BC_ERR_ARG, "bad frame shape");
  if (B < 1 || B > c->max_batch) return fail(c, BC_ERR_ARG, "batch size outside [1, max_batch]");
  if (!c->net_loaded) return fail(c, BC_ERR_STATE, "bc_load_enet has not been called");
  BevGeom g;
  int r = make_geom(c, w_m, h_m, cell_m, binary, ros_layout, g);
  if (r) return r;
  if (c->in_rows != BC_NET_H || c->in_cols != BC_NET_W)
    return fail(c, BC_ERR_ARG, "calibration input size must be (256, 512) for the ENet pipeline (bev.py:169)");
  CU(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (c->gs.on && !h_grids && c->rank == 0) return fail(c, BC_ERR_ARG, "rank 0 needs a host buffer for the gathered grids");
  if (!c->gs.on && !h_grids) return fail(c, BC_ERR_ARG, "null pointer");
  bc_ctx::Slot& sl = c->slots[c->submitted & 1];
  const size_t in_bytes = (size_t)B * h * w * 3, out_bytes = c->gs.on ? 0 : (size_t)B * g.Hc * g.Wc;
  if (sl.busy) { CU(cudaEventSynchronize(sl.done)); sl.busy = false; }      // at most two steps in flight
  if (sl.in_bytes < in_bytes || sl.out_bytes < out_bytes) {
    CU(cudaDeviceSynchronize());
    invalidate_graphs(c);
    if (sl.in_bytes < in_bytes) {
      if (sl.d_in) cudaFree(sl.d_in);
      sl.d_in = nullptr; sl.in_bytes = 0;
      CU(cudaMalloc(&sl.d_in, in_bytes));
      sl.in_bytes = in_bytes;
    }
    if (sl.out_bytes < out_bytes) {
      if (sl.d_out) cudaFree(sl.d_out);
      sl.d_out = nullptr; sl.out_bytes = 0;
      CU(cudaMalloc(&sl.d_out, out_bytes));
      sl.out_bytes = out_bytes;
    }
  }
  if ((r = ensure_scratch(c))) return r;
  const ResizeTab* rt;
  if ((r = get_resize_tab(c, h, w, &rt))) return r;   // may allocate: keep it out of capture
  // the slot's previous kernels (two submits ago) were waited for above via `done`, which follows them
  CU(cudaMemcpyAsync(sl.d_in, h_bgr, in_bytes, cudaMemcpyHostToDevice, c->copy_stream));
  CU(cudaEventRecord(sl.copied, c->copy_stream));
  CU(cudaStreamWaitEvent(s, sl.copied, 0));
  if (c->gs.on) {
    // multi-GPU streaming gather: this rank's grids go straight into rank 0's buffer of the slot; flags order
    // the ranks (see bc_gather_stream_setup in the header)
    bc_ctx::GatherStream& gs = c->gs;
    const int j = (int)(gs.steps & 1);
    const uint32_t gen = (uint32_t)(gs.steps / 2 + 1);
    const size_t cells = (size_t)g.Hc * g.Wc;
    if (gen > 1) wait_flags(c, gs.release_mine + j, 1, gen - 1, gs.d_err, s);        // rank 0 has drained the slot's previous step
    if ((r = run_pipeline(c, sl.d_in, h, w, B, h_lut, w_m, h_m, cell_m, g, nullptr, gs.gather[j] + (size_t)c->rank * B * cells, s)))
      return r;
    launch_flag_store1(gs.arrive + (size_t)j * c->world + c->rank, gen, s);
    c->launches += 1;
    CU(cudaEventRecord(sl.computed, s));
    if (c->rank == 0) {
      wait_flags(c, gs.arrive + (size_t)j * c->world, c->world, gen, gs.d_err, c->d2h_stream);
      CU(cudaMemcpyAsync(h_grids, gs.gather[j], (size_t)c->world * B * cells, cudaMemcpyDeviceToHost, c->d2h_stream));
      launch_flag_store(gs.d_release_peers + (size_t)j * c->world, c->world, gen, c->d2h_stream);
      c->launches += 1;
      CU(cudaEventRecord(sl.done, c->d2h_stream));
    } else {
      CU(cudaEventRecord(sl.done, s));
    }
    gs.steps++;
    sl.busy = true;
    c->submitted++;
    return check_launch(c, "streaming gather");
  }
  if ((r = run_pipeline(c, sl.d_in, h, w, B, h_lut, w_m, h_m, cell_m, g, nullptr, sl.d_out, s))) return r;
  // the grids leave on a third stream, so the next step's kernels (already queued on `s`) start at once
  CU(cudaEventRecord(sl.computed, s));
  CU(cudaStreamWaitEvent(c->d2h_stream, sl.computed, 0));
  CU(cudaMemcpyAsync(h_grids, sl.d_out, out_bytes, cudaMemcpyDeviceToHost, c->d2h_stream));
  CU(cudaEventRecord(sl.done, c->d2h_stream));
  sl.busy = true;
  c->submitted++;
  return BC_OK;
}

// Blocks until at most `keep_in_flight` (0 or 1) submitted steps are still running; the grids
// of every completed step are then in their host buffers.
int bc_pipeline_host_wait(bc_ctx* c, int keep_in_flight) {
  if (!c) return BC_ERR_ARG;
  if (keep_in_flight < 0 || keep_in_flight > 1) return fail(c, BC_ERR_ARG, "keep_in_flight must be 0 or 1");
  CU(cudaSetDevice(c->device));
  // slots in submission order: the older one is the slot the NEXT submit would take
  for (int k = 0; k < 2 - keep_in_flight; ++k) {
    bc_ctx::Slot& sl = c->slots[(c->submitted + k) & 1];
    if (sl.busy) { CU(cudaEventSynchronize(sl.done)); sl.busy = false; }
  }
  if (c->gs.on) {
    const int err = *(volatile int*)c->gs.h_err;
    if (err) return fail(c, BC_ERR_STATE, "streaming gather: a rank did not deliver its grids within the time-out");
  }
  return BC_OK;
}

int bc_gather_setup(bc_ctx* c, void* d_gather_base, int rank, int world) {
  if (!c) return BC_ERR_ARG;
  if (d_gather_base && (world < 1 || rank < 0 || rank >= world)) return fail(c, BC_ERR_ARG, "rank outside [0, world)");
  c->gather_base = (int8_t*)d_gather_base;
  c->rank = d_gather_base ? rank : 0;
  c->world = d_gather_base ? world : 1;
  return BC_OK;
}

int bc_gather_stream_setup(bc_ctx* c, void* const d_gather[2], uint32_t* d_arrive, uint32_t* d_release_mine,
                           uint32_t* const* d_release_peers, int rank, int world) {
  if (!c) return BC_ERR_ARG;
  CU(cudaSetDevice(c->device));
  CU(cudaDeviceSynchronize());
  bc_ctx::GatherStream& gs = c->gs;
  if (gs.d_release_peers) { cudaFree(gs.d_release_peers); gs.d_release_peers = nullptr; }
  gs.on = false;
  gs.steps = 0;
  if (!d_gather) { c->rank = 0; c->world = 1; return BC_OK; }
  if (world < 1 || world > 64 || rank < 0 || rank >= world) return fail(c, BC_ERR_ARG, "rank outside [0, world), world <= 64");
  if (!d_gather[0] || !d_gather[1] || !d_arrive || !d_release_mine) return fail(c, BC_ERR_ARG, "null pointer");
  if (rank == 0 && !d_release_peers) return fail(c, BC_ERR_ARG, "rank 0 needs every rank's release flags");
  gs.gather[0] = (int8_t*)d_gather[0]; gs.gather[1] = (int8_t*)d_gather[1];
  gs.arrive = d_arrive;
  gs.release_mine = d_release_mine;
  if (!gs.h_err) {
    CU(cudaHostAlloc((void**)&gs.h_err, sizeof(int), cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer((void**)&gs.d_err, gs.h_err, 0));
  }
  *gs.h_err = 0;
  if (rank == 0) {
    std::vector<uint32_t*> tab((size_t)2 * world);          // [slot][rank] -> &release_r[slot]
    for (int sidx = 0; sidx < 2; ++sidx)
      for (int r2 = 0; r2 < world; ++r2) {
        if (!d_release_peers[r2]) return fail(c, BC_ERR_ARG, "null release-flag pointer");
        tab[(size_t)sidx * world + r2] = d_release_peers[r2] + sidx;
      }
    CU(cudaMalloc(&gs.d_release_peers, tab.size() * sizeof(uint32_t*)));
    CU(cudaMemcpy(gs.d_release_peers, tab.data(), tab.size() * sizeof(uint32_t*), cudaMemcpyHostToDevice));
  }
  c->rank = rank; c->world = world;
  c->submitted = 0;
  for (auto& sl : c->slots) sl.busy = false;
  invalidate_graphs(c);
  gs.on = true;
  return BC_OK;
}

int bc_host_alloc(void** h_ptr, size_t bytes, int write_combined) {
  if (!h_ptr || bytes == 0) return BC_ERR_ARG;
  *h_ptr = nullptr;
  cudaError_t e = cudaHostAlloc(h_ptr, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
  if (e != cudaSuccess) { cudaGetLastError(); g_create_err = std::string("cudaHostAlloc: ") + cudaGetErrorString(e); return BC_ERR_NOMEM; }
  return BC_OK;
}

int bc_host_free(void* h_ptr) {
  if (!h_ptr) return BC_OK;
  return cudaFreeHost(h_ptr) == cudaSuccess ? BC_OK : BC_ERR_CUDA;
}

long long bc_launch_count(const bc_ctx* c) { return c ? c->launches : 0; }

int bc_set_profile(bc_ctx* c, int enable) {
  if (!c) return BC_ERR_ARG;
  CU(cudaSetDevice(c->device));
  CU(cudaDeviceSynchronize());
  for (auto& r : c->prof) { c->ev_pool.push_back(r.start); c->ev_pool.push_back(r.stop); }
  c->prof.clear();
  c->profiling = enable ? 1 : 0;
  return BC_OK;
}

const char* bc_profile_json(bc_ctx* c) {
  if (!c) return "[]";
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  struct Agg { long long n = 0; double ms = 0, bytes = 0, flops = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : c->prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.start, r.stop) != cudaSuccess) continue;
    Agg& a = agg[r.name];
    a.n++; a.ms += ms; a.bytes += r.bytes; a.flops += r.flops;
  }
  std::string out = "[";
  bool first = true;
  for (auto& kv : agg) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s{\"kernel\": \"%s\", \"launches\": %lld, \"ms\": %.6f, \"bytes\": %.0f, \"flops\": %.0f}",
             first ? "" : ", ", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.bytes, kv.second.flops);
    out += buf;
    first = false;
  }
  out += "]";
  c->prof_json = out;
  return c->prof_json.c_str();
}

}  // extern "C"
