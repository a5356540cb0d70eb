// extern "C" entry points of libdeepcut_b200.so (see include/deepcut_b200.h).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/deepcut_b200.h"
#include "conv_igemm.cuh"
#include "hbm_kernels.cuh"
#include "layer_kernels.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
// largest split-K cluster dc_conv_forward may pick for under-filled grids (1 = never split); DC_SPLIT_K overrides the default
std::atomic<int> g_split_k_max{[] { const char* e = getenv("DC_SPLIT_K"); const int v = e ? atoi(e) : 4; return (v == 1 || v == 2 || v == 4) ? v : 4; }()};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define DC_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess) return fail(DC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
// SMs the persistent conv grids leave free (dc_set_reserved_sms): a persistent kernel with one CTA per SM and a static tile schedule
// cannot share its SMs -- a concurrent NCCL kernel that takes even a few of them delays the CTAs that should have run there by a whole
// kernel, i.e. doubles that kernel's time.  While a batch exchange is in flight the forwards therefore run on SMs - reserve.
std::atomic<int> g_reserved_sms{[] { const char* e = getenv("DC_RESERVED_SMS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 0; }()};
int persistent_sms() {
  const int n = g_num_sms - g_reserved_sms.load();
  return n >= 2 ? n : 2;
}
// a layer's K loop is shared by a split-K cluster only from this many 64-channel K-steps on (DC_SPLIT_K_MIN_STEPS)
std::atomic<int> g_split_k_min_steps{[] { const char* e = getenv("DC_SPLIT_K_MIN_STEPS"); const int v = e ? atoi(e) : 16; return v >= 8 ? v : 16; }()};
bool g_inited = false;

int ensure_init() {
  if (g_inited) return DC_OK;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(DC_ERR_NO_DEVICE, "no CUDA device: libdeepcut_b200 has no CPU path");
  return dc_init(dev);
}

uint16_t f2h_bits(float f) {
  // round-to-nearest-even fp32 -> fp16 bit pattern (host side, matches __float2half_rn)
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7FFFFFFFu;
  if (x >= 0x7F800000u) return static_cast<uint16_t>(sign | 0x7C00u | ((x > 0x7F800000u) ? 0x200u : 0));
  if (x >= 0x477FF000u) return static_cast<uint16_t>(sign | 0x7C00u);           // overflow -> inf
  if (x < 0x33000001u) return static_cast<uint16_t>(sign);                        // underflow -> 0
  int exp = static_cast<int>(x >> 23) - 127;
  uint32_t man = (x & 0x7FFFFFu) | 0x800000u;
  int shift;
  uint32_t hexp;
  if (exp < -14) { shift = 13 + (-14 - exp); hexp = 0; } else { shift = 13; hexp = static_cast<uint32_t>(exp + 15); }
  uint32_t half_man = man >> shift;
  const uint32_t rem = man & ((1u << shift) - 1u);
  const uint32_t halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (half_man & 1u))) half_man++;
  uint32_t out;
  if (hexp == 0) out = half_man;                  // subnormal (may carry into exponent 1, which is correct)
  else out = ((hexp - 1) << 10) + half_man;       // half_man includes the implicit bit (0x400)
  return static_cast<uint16_t>(sign | out);
}
float h2f_bits(uint16_t h) {
  const uint32_t sign = (h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, out;
  if (exp == 0) {
    if (man == 0) out = sign;
    else {
      int e = -1;
      do { e++; man <<= 1; } while (!(man & 0x400u));
      out = sign | static_cast<uint32_t>(127 - 15 - e) << 23 | ((man & 0x3FFu) << 13);
    }
  } else if (exp == 31) out = sign | 0x7F800000u | (man << 13);
  else out = sign | ((exp + 112) << 23) | (man << 13);
  float f;
  memcpy(&f, &out, 4);
  return f;
}

// Packs one logical GEMM row (K fp32 values gathered by `get`) as scaled split fp16.
template <class Get>
void pack_row(int K, Get get, uint16_t* hi, uint16_t* lo, float* rowscale) {
  float mx = 0.f;
  for (int k = 0; k < K; ++k) mx = fmaxf(mx, fabsf(get(k)));
  float s = 1.f;
  if (mx > 0.f && std::isfinite(mx)) {
    int e;
    frexpf(mx, &e);              // mx = m * 2^e, m in [0.5, 1)
    s = ldexpf(1.f, 10 - e);     // mx * s in [2^9, 2^10)
  }
  for (int k = 0; k < K; ++k) {
    const float v = get(k) * s;  // exact: power-of-two scaling
    const uint16_t h = f2h_bits(v);
    hi[k] = h;
    lo[k] = f2h_bits(v - h2f_bits(h));
  }
  *rowscale = 1.f / s;
}

int tile_n_for(int cout) { return cout > 64 ? 128 : 64; }

// plane_elems: distance between the hi and lo planes in elements (0 = dense n*h*w*c; larger when the launch covers a
// sub-batch of a bigger tensor)
int encode_act_map(CUtensorMap* m, const void* base, int n, int h, int w, int c, int tw, int th, int stride = 1, long long plane_elems = 0) {
  const cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n, 2};
  const cuuint64_t strides[4] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2,
                                 (cuuint64_t)(plane_elems > 0 ? plane_elems : (long long)n * h * w * c) * 2};
  // traversal stride s: the box spans tw*s x th*s input pixels of which every s-th is loaded (tw x th rows of smem)
  const cuuint32_t box[5] = {(cuuint32_t)dc::kBK, (cuuint32_t)(tw * stride), (cuuint32_t)(th * stride), 1, 1};
  const cuuint32_t es[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1, 1};
  // L2 promotion 128 B = exactly the 64-channel row a K-step needs (256 B measured the same: 522 vs 522 images/s at 16x720p)
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DC_ERR_CUDA, "cuTensorMapEncodeTiled(activations) failed: %d", (int)r);
  return DC_OK;
}
// Output (split NHWC) map for the epilogue's TMA stores: box = 32 channels x the 32-pixel sub-rectangle
// one epilogue warp owns, SWIZZLE_64B to match the staging tile.
int encode_out_map(CUtensorMap* m, const void* base, int n, int h, int w, int c, int tw, long long plane_elems = 0) {
  const cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n, 2};
  const cuuint64_t strides[4] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2,
                                 (cuuint64_t)(plane_elems > 0 ? plane_elems : (long long)n * h * w * c) * 2};
  const int bw = tw < 32 ? tw : 32;
  const cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)(32 / bw), 1, 1};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DC_ERR_CUDA, "cuTensorMapEncodeTiled(output) failed: %d", (int)r);
  return DC_OK;
}
int encode_w_map(CUtensorMap* m, const void* base, int rows, long long K, int bn) {
  const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 2};
  const cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)rows * K * 2};
  const cuuint32_t box[3] = {(cuuint32_t)dc::kBK, (cuuint32_t)bn, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DC_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return DC_OK;
}

// conv_igemm requests its first weight tiles before griddepcontrol.wait (DC_EARLY_WEIGHTS=0: after it, like the activations)
int use_early_weights() {
  static const int on = [] { const char* e = getenv("DC_EARLY_WEIGHTS"); return e ? atoi(e) : 1; }();
  return on;
}

bool use_pdl() {
  static const bool on = [] { const char* e = getenv("DC_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

// ksplit > 1 (SK = 1): a cluster of `ksplit` CTAs per work unit, each taking a share of the K loop (conv_igemm.cuh)
template <int BN, int CG, int EW = 8, int SK = 0, int TALL = 0>
int launch_conv(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const dc::ConvParams& p, cudaStream_t st,
                int ksplit = 1) {
  const int units = ((p.n_tiles_m + CG - 1) / CG) * p.n_tiles_n;
  dc::ConvParams pk = p;                       // the launch's copy, with the tile-decode divisors filled in
  pk.div_ntn = dc::fastdiv_make(p.n_tiles_n);
  pk.div_tx = dc::fastdiv_make(p.tiles_x);
  pk.div_ty = dc::fastdiv_make(p.tiles_y);
  const int sms = persistent_sms();
  int grid = units * CG < sms ? units * CG : sms;
  if (CG == 2) grid &= ~1;
  if (SK) grid = units * ksplit;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(dc::conv_threads(EW));
  cfg.dynamicSmemBytes = TALL ? dc::ConvCfg<BN, CG, EW>::kTallSmemBytes : dc::ConvCfg<BN, CG, EW>::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (use_pdl()) {     // the kernel's prologue overlaps the previous kernel's tail (griddepcontrol.wait inside)
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (CG == 2 || SK) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = SK ? ksplit : 2;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  DC_CUDA(cudaLaunchKernelEx(&cfg, dc::conv_igemm_kernel<BN, CG, EW, SK, TALL>, ta, tb, to, pk));
  g_launches++;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

// TALL-mode geometry (conv_igemm.cuh) for the tile rectangle already chosen in `p`: `groups` column offsets `gdx`, `taps_per_group`
// taps each, `row_step` input rows apart, the first one `top` rows from the tile's first output row.  False when the boxes do not fit
// the shared-memory budget or the tap offsets would break the 1024-byte swizzle phase.
bool setup_tall(dc::ConvParams& p, int groups, int taps_per_group, const int* gdx, int top, int row_step, int* tall_rows_out) {
  const int tall_rows = p.TH + (taps_per_group - 1) * row_step;
  if (tall_rows > 256 || (p.TW * row_step) % 8 != 0 || (p.TW * tall_rows) % 8 != 0 || groups > 3) return false;
  const long long plane = static_cast<long long>(tall_rows) * p.TW * 128;
  const long long avail = dc::ConvCfg<64, 2, 8>::kTallOperandBytes - static_cast<long long>(p.ntaps) * 2 * dc::ConvCfg<64, 2, 8>::kBBytes;
  long long stages = avail / (2 * plane);
  if (stages > 4) stages = 4;
  if (stages < 2) return false;
  p.tall_groups = groups;
  p.tall_taps_per_group = taps_per_group;
  for (int g = 0; g < groups; ++g) p.tall_gdx[g] = gdx[g];
  p.tall_top = top;
  p.tall_plane_bytes = static_cast<int>(plane);
  p.tall_tap_bytes = row_step * p.TW * 128;
  p.tall_stages = static_cast<int>(stages);
  *tall_rows_out = tall_rows;
  return true;
}
// the tile rectangles TALL mode chooses from: at least 8 rows, so that a box of TH + (taps - 1) * row_step rows does not dwarf the tile
void choose_tall_tile(dc::ConvParams& p, int n, int out_h, int out_w) {
  const int cand[2][2] = {{16, 8}, {8, 16}};
  long long best = -1;
  for (int i = 0; i < 2; ++i) {
    const int tw = cand[i][0], th = cand[i][1];
    const long long area = static_cast<long long>((out_w + tw - 1) / tw) * tw * ((out_h + th - 1) / th) * th;
    if (best < 0 || area < best) { best = area; p.TW = tw; p.TH = th; }
  }
  for (p.log2_tw = 0; (1 << p.log2_tw) < p.TW; ++p.log2_tw) {}
  p.tiles_x = (out_w + p.TW - 1) / p.TW;
  p.tiles_y = (out_h + p.TH - 1) / p.TH;
  p.n_tiles_m = n * p.tiles_x * p.tiles_y;
}
bool use_tall() {        // read per launch (graph capture time): one process can sweep it
  const char* e = getenv("DC_CONV_TALL");
  return !(e && e[0] == '0');
}

// How many split-K clusters of `s` CTAs (one CTA per SM: ~224 KB of shared memory each) the device keeps resident at once:
// GPCs whose SM count is not a multiple of `s` strand SMs, so this is less than num_sms / s.  A split grid must fit in
// one wave, otherwise splitting only adds a second wave.
template <int BN>
int max_split_clusters(int s) {
  static std::atomic<int> cache[5] = {{-1}, {-1}, {-1}, {-1}, {-1}};
  if (s < 2 || s > 4) return 0;
  int n = cache[s].load();
  if (n >= 0) return n;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(g_num_sms / s * s);
  cfg.blockDim = dim3(dc::conv_threads(8));
  cfg.dynamicSmemBytes = dc::ConvCfg<BN, 1, 8>::kSmemBytes;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = s;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, dc::conv_igemm_kernel<BN, 1, 8, 1>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  cache[s].store(n);
  return n;
}

// CTA-pair (cta_group::2) kernels for the pixel-major modes; DC_CONV_2CTA=0 falls back to single CTAs.
bool use_2cta() {
  static const bool on = [] { const char* e = getenv("DC_CONV_2CTA"); return !(e && e[0] == '0'); }();
  return on;
}

int ew_grid(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(g_num_sms > 0 ? g_num_sms : 148) * 8;
  return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

extern "C" {

int dc_version(void) { return 100; }
int dc_set_split_k(int max_split) {
  if (max_split != 1 && max_split != 2 && max_split != 4) return fail(DC_ERR_INVALID, "dc_set_split_k: %d is not 1, 2 or 4", max_split);
  g_split_k_max.store(max_split);
  return DC_OK;
}
int dc_get_split_k(void) { return g_split_k_max.load(); }
int dc_set_reserved_sms(int n) {
  if (n < 0 || (g_num_sms > 0 && n > g_num_sms - 2)) return fail(DC_ERR_INVALID, "dc_set_reserved_sms: %d out of range", n);
  g_reserved_sms.store(n);
  return DC_OK;
}
int dc_get_reserved_sms(void) { return g_reserved_sms.load(); }
size_t dc_splitk_workspace_bytes(void) {
  // units * S <= SMs and a slot is at most a 128 x 128 fp32 tile: units * (S - 1) slots < SMs slots
  return static_cast<size_t>(g_num_sms > 0 ? g_num_sms : 148) * 128 * 128 * 4;
}
int dc_set_split_k_min_steps(int min_ksteps) {
  if (min_ksteps < 8) return fail(DC_ERR_INVALID, "dc_set_split_k_min_steps: %d < 8 (every CTA of a 4-way split needs K-steps)", min_ksteps);
  g_split_k_min_steps.store(min_ksteps);
  return DC_OK;
}
int dc_get_split_k_min_steps(void) { return g_split_k_min_steps.load(); }
const char* dc_last_error(void) { return g_err; }
long long dc_launch_count(void) { return g_launches.load(); }

int dc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) ok++;
  }
  return ok;
}

int dc_init(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(DC_ERR_NO_DEVICE, "no CUDA device: libdeepcut_b200 has no CPU path");
  }
  DC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  DC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(DC_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is sm_100a only", device, prop.major, prop.minor);
  g_num_sms = prop.multiProcessorCount;
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    DC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail(DC_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  DC_CUDA(cudaFuncSetAttribute(dc::conv_igemm_kernel<128, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, dc::ConvCfg<128, 1, 8>::kSmemBytes));
  DC_CUDA(cudaFuncSetAttribute(dc::conv_igemm_kernel<64, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, dc::ConvCfg<64, 1, 8>::kSmemBytes));
  DC_CUDA(cudaFuncSetAttribute(dc::conv_igemm_kernel<128, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, dc::ConvCfg<128, 2, 8>::kSmemBytes));
  DC_CUDA(cudaFuncSetAttribute(dc::conv_igemm_kernel<64, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, dc::ConvCfg<64, 2, 8>::kSmemBytes));
  DC_CUDA(cudaFuncSetAttribute(dc::conv_igemm_kernel<128, 2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, dc::ConvCfg<128, 2, 16>::kSmemBytes));
  DC_CUDA(cudaFuncSetAttribute(dc::conv_igemm_kernel<256, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, dc::ConvCfg<256, 2, 8>::kSmemBytes));
  DC_CUDA(cudaFuncSetAttribute(dc::conv_igemm_kernel<64, 2, 8, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dc::ConvCfg<64, 2, 8>::kTallSmemBytes));
  DC_CUDA(cudaFuncSetAttribute(dc::conv_igemm_kernel<128, 1, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dc::ConvCfg<128, 1, 8>::kSmemBytes));
  DC_CUDA(cudaFuncSetAttribute(dc::conv_igemm_kernel<64, 1, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dc::ConvCfg<64, 1, 8>::kSmemBytes));
  DC_CUDA(cudaFuncSetAttribute(dc::conv1_7x7s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dc::kC1SmemFloats * 4));
  g_inited = true;
  return DC_OK;
}

// ------------------------------------------------------------------ memory / streams / events
int dc_malloc(void** ptr, size_t bytes) {
  if (int rc = ensure_init()) return rc;
  if (!ptr) return fail(DC_ERR_INVALID, "dc_malloc: null");
  DC_CUDA(cudaMalloc(ptr, bytes));
  return DC_OK;
}
int dc_free(void* ptr) { if (ptr) DC_CUDA(cudaFree(ptr)); return DC_OK; }
int dc_malloc_host(void** ptr, size_t bytes) {
  if (int rc = ensure_init()) return rc;
  if (!ptr) return fail(DC_ERR_INVALID, "dc_malloc_host: null");
  DC_CUDA(cudaMallocHost(ptr, bytes));
  return DC_OK;
}
int dc_free_host(void* ptr) { if (ptr) DC_CUDA(cudaFreeHost(ptr)); return DC_OK; }
int dc_memcpy_async(void* dst, const void* src, size_t bytes, int kind, void* stream) {
  if (int rc = ensure_init()) return rc;
  cudaMemcpyKind k;
  switch (kind) {
    case DC_H2D: k = cudaMemcpyHostToDevice; break;
    case DC_D2H: k = cudaMemcpyDeviceToHost; break;
    case DC_D2D: k = cudaMemcpyDeviceToDevice; break;
    default: return fail(DC_ERR_INVALID, "dc_memcpy_async: bad kind %d", kind);
  }
  if (bytes == 0) return DC_OK;
  DC_CUDA(cudaMemcpyAsync(dst, src, bytes, k, static_cast<cudaStream_t>(stream)));
  return DC_OK;
}
int dc_memset_async(void* ptr, int value, size_t bytes, void* stream) {
  if (int rc = ensure_init()) return rc;
  if (bytes == 0) return DC_OK;
  DC_CUDA(cudaMemsetAsync(ptr, value, bytes, static_cast<cudaStream_t>(stream)));
  return DC_OK;
}
int dc_stream_create(void** stream) {
  if (int rc = ensure_init()) return rc;
  cudaStream_t s;
  DC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = s;
  return DC_OK;
}
int dc_stream_destroy(void* stream) { if (stream) DC_CUDA(cudaStreamDestroy(static_cast<cudaStream_t>(stream))); return DC_OK; }
int dc_stream_sync(void* stream) { DC_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream))); return DC_OK; }
int dc_device_sync(void) { if (int rc = ensure_init()) return rc; DC_CUDA(cudaDeviceSynchronize()); return DC_OK; }
int dc_mem_info(size_t* free_bytes, size_t* total_bytes) {
  if (int rc = ensure_init()) return rc;
  DC_CUDA(cudaMemGetInfo(free_bytes, total_bytes));
  return DC_OK;
}
int dc_event_create(void** event) {
  if (int rc = ensure_init()) return rc;
  cudaEvent_t e;
  DC_CUDA(cudaEventCreate(&e));
  *event = e;
  return DC_OK;
}
int dc_event_destroy(void* event) { if (event) DC_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(event))); return DC_OK; }
int dc_event_record(void* event, void* stream) {
  DC_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), static_cast<cudaStream_t>(stream)));
  return DC_OK;
}
int dc_event_sync(void* event) {
  DC_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(event)));
  return DC_OK;
}
int dc_event_elapsed_ms(void* start, void* stop, float* ms) {
  DC_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(stop)));
  DC_CUDA(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
  return DC_OK;
}

// ------------------------------------------------------------------ CUDA graphs
namespace {
struct GraphExec { cudaGraphExec_t exec; long long launches; };
thread_local long long g_capture_base = -1;
}  // namespace

int dc_graph_begin(void* stream) {
  if (int rc = ensure_init()) return rc;
  DC_CUDA(cudaStreamBeginCapture(static_cast<cudaStream_t>(stream), cudaStreamCaptureModeThreadLocal));
  g_capture_base = g_launches.load();
  return DC_OK;
}
int dc_graph_end(void* stream, void** graph_exec) {
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(static_cast<cudaStream_t>(stream), &graph);
  const long long captured = g_launches.load() - g_capture_base;
  g_launches -= captured;       // captured launches have not run yet; dc_graph_launch counts them per replay
  g_capture_base = -1;
  if (e != cudaSuccess || graph == nullptr) { cudaGetLastError(); return fail(DC_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e)); }
  GraphExec* g = new GraphExec();
  g->launches = captured;
  e = cudaGraphInstantiate(&g->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { delete g; cudaGetLastError(); return fail(DC_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
  *graph_exec = g;
  return DC_OK;
}
int dc_graph_launch(void* graph_exec, void* stream) {
  GraphExec* g = static_cast<GraphExec*>(graph_exec);
  if (!g) return fail(DC_ERR_INVALID, "dc_graph_launch: null graph");
  DC_CUDA(cudaGraphLaunch(g->exec, static_cast<cudaStream_t>(stream)));
  g_launches += g->launches;
  return DC_OK;
}
int dc_graph_destroy(void* graph_exec) {
  GraphExec* g = static_cast<GraphExec*>(graph_exec);
  if (g) { cudaGraphExecDestroy(g->exec); delete g; }
  return DC_OK;
}

// ------------------------------------------------------------------ host-side transforms
int dc_fold_bn_scale(const float* mean_sum, const float* var_sum, float factor, float eps, const float* gamma,
                     const float* beta, int channels, float* a_out, float* b_out) {
  if (!mean_sum || !var_sum || !a_out || !b_out || channels <= 0) return fail(DC_ERR_INVALID, "dc_fold_bn_scale: bad arguments");
  const float sf = factor == 0.f ? 0.f : 1.f / factor;
  for (int c = 0; c < channels; ++c) {
    const float mean = mean_sum[c] * sf;
    const float var = var_sum[c] * sf;
    const double inv = 1.0 / std::sqrt(static_cast<double>(var) + static_cast<double>(eps));
    const double g = gamma ? gamma[c] : 1.0;
    const double b = beta ? beta[c] : 0.0;
    a_out[c] = static_cast<float>(g * inv);
    b_out[c] = static_cast<float>(b - static_cast<double>(mean) * g * inv);
  }
  return DC_OK;
}

int dc_tile_n(int cout) { return tile_n_for(cout); }
int dc_packed_rows(int cout) {
  const int bn = tile_n_for(cout);
  return (cout + bn - 1) / bn * bn;
}

int dc_pack_conv_weight(const float* w, int cout, int cin, int kh, int kw, uint16_t* packed, float* rowscale) {
  if (!w || !packed || !rowscale || cout <= 0 || cin <= 0 || kh <= 0 || kw <= 0) return fail(DC_ERR_INVALID, "dc_pack_conv_weight: bad arguments");
  const int rows = dc_packed_rows(cout);
  const long long K = static_cast<long long>(kh) * kw * cin;
  uint16_t* hi = packed;
  uint16_t* lo = packed + static_cast<long long>(rows) * K;
  for (int r = 0; r < rows; ++r) {
    if (r >= cout) {
      memset(hi + r * K, 0, K * 2);
      memset(lo + r * K, 0, K * 2);
      rowscale[r] = 1.f;
      continue;
    }
    const float* wr = w + static_cast<long long>(r) * cin * kh * kw;
    const int taps = kh * kw;
    pack_row(static_cast<int>(K), [&](int k) { const int t = k / cin, ci = k % cin; return wr[static_cast<long long>(ci) * taps + t]; },
             hi + r * K, lo + r * K, rowscale + r);
  }
  return DC_OK;
}

int dc_pack_deconv_weight(const float* w, int cin, int cout, int kh, int kw, uint16_t* packed, float* rowscale) {
  if (!w || !packed || !rowscale || cout <= 0 || cin <= 0 || kh <= 0 || kw <= 0) return fail(DC_ERR_INVALID, "dc_pack_deconv_weight: bad arguments");
  const int taps = kh * kw;
  const int nrows = cout * taps;
  const int rows = dc_packed_rows(nrows);
  const long long K = cin;
  uint16_t* hi = packed;
  uint16_t* lo = packed + static_cast<long long>(rows) * K;
  for (int r = 0; r < rows; ++r) {
    if (r >= nrows) {
      memset(hi + r * K, 0, K * 2);
      memset(lo + r * K, 0, K * 2);
      rowscale[r] = 1.f;
      continue;
    }
    // W[ci][co][p][q]; row r = co*taps + t  ->  element offset ci*cout*taps + r
    pack_row(static_cast<int>(K), [&](int ci) { return w[static_cast<long long>(ci) * cout * taps + r]; }, hi + r * K,
             lo + r * K, rowscale + r);
  }
  return DC_OK;
}

int dc_pack_conv1_weight(const float* w, float* packed) {
  if (!w || !packed) return fail(DC_ERR_INVALID, "dc_pack_conv1_weight: bad arguments");
  for (int co = 0; co < 64; ++co)
    for (int k = 0; k < 147; ++k) packed[k * 64 + co] = w[co * 147 + k];
  return DC_OK;
}

int dc_pack_conv1_tc_weight(const float* w, uint16_t* packed, float* rowscale) {
  if (!w || !packed || !rowscale) return fail(DC_ERR_INVALID, "dc_pack_conv1_tc_weight: bad arguments");
  const int K = 256;
  std::vector<float> row(K);
  for (int co = 0; co < 64; ++co) {
    std::fill(row.begin(), row.end(), 0.f);
    for (int ci = 0; ci < 3; ++ci)
      for (int P = 0; P < 7; ++P)
        for (int Q = 0; Q < 7; ++Q) {
          // P - 3 = 2 t + py, t = floor((P-3)/2) in {-2..1}
          const int dP = P - 3, dQ = Q - 3;
          const int ty = (dP >= 0) ? dP / 2 : -((-dP + 1) / 2), py = dP - 2 * ty;
          const int tx = (dQ >= 0) ? dQ / 2 : -((-dQ + 1) / 2), px = dQ - 2 * tx;
          row[(ty + 2) * 64 + (tx + 2) * 16 + (py * 2 + px) * 3 + ci] = w[((co * 3 + ci) * 7 + P) * 7 + Q];
        }
    pack_row(K, [&](int k) { return row[k]; }, packed + static_cast<long long>(co) * K, packed + static_cast<long long>(64 + co) * K, rowscale + co);
  }
  return DC_OK;
}

size_t dc_conv1_tc_workspace_bytes(int n, int h, int w) {
  const long long h2 = (h + 1) / 2, w2 = (w + 1) / 2;
  return static_cast<size_t>(2ll * n * h2 * (w2 + 3) * 16 * 2);
}

int dc_pool_out_size(int size, int kernel, int stride) {
  // PoolingLayer::Reshape, pooling_layer.cpp:90-93 (pad 0): ceil((size - k) / s) + 1
  return static_cast<int>(std::ceil(static_cast<float>(size - kernel) / stride)) + 1;
}

// ------------------------------------------------------------------ fused convolution
int dc_conv_forward(const dc_conv_args* a, void* stream) {
  if (int rc = ensure_init()) return rc;
  if (!a || !a->x || !a->w_packed || !a->scale || !a->shift || !a->out) return fail(DC_ERR_INVALID, "dc_conv_forward: null argument");
  if (a->cin % dc::kBK != 0) return fail(DC_ERR_UNSUPPORTED, "dc_conv_forward: cin=%d is not a multiple of 64", a->cin);
  if (a->kh * a->kw > dc::kMaxTaps) return fail(DC_ERR_UNSUPPORTED, "dc_conv_forward: %dx%d filter exceeds 9 taps", a->kh, a->kw);
  if (!a->out_f32_rows && a->cout % 32 != 0) return fail(DC_ERR_UNSUPPORTED, "dc_conv_forward: cout=%d must be a multiple of 32 for split output", a->cout);
  const int stride = a->stride > 1 ? a->stride : 1;
  if (stride > 1 && !(a->kh == 1 && a->kw == 1 && a->pad == 0 && !a->out_f32_rows))
    return fail(DC_ERR_UNSUPPORTED, "dc_conv_forward: stride %d only for 1x1 pad 0 convolutions with split output", stride);
  if (stride > 8) return fail(DC_ERR_UNSUPPORTED, "dc_conv_forward: stride %d exceeds the TMA traversal stride limit (8)", stride);
  const int ho = (a->h + 2 * a->pad - (a->dilation * (a->kh - 1) + 1)) / stride + 1;      // conv_layer.cpp:17-19
  const int wo = (a->w + 2 * a->pad - (a->dilation * (a->kw - 1) + 1)) / stride + 1;
  if (ho <= 0 || wo <= 0) return fail(DC_ERR_INVALID, "dc_conv_forward: empty output");
  int bn = tile_n_for(a->cout);
  const int rows = dc_packed_rows(a->cout);
  if (a->out_f32_rows == 1 && a->ldc < rows) return fail(DC_ERR_INVALID, "dc_conv_forward: ldc=%d < packed rows %d", a->ldc, rows);
  if (a->out_f32_rows == 2) {
    if (!(a->kh == 1 && a->kw == 1 && a->pad == 0) || bn != 128)
      return fail(DC_ERR_UNSUPPORTED, "dc_conv_forward: channel-major output needs a 1x1 convolution with more than 64 outputs");
    if (a->ldc < a->n * a->h * a->w || a->ldc % 4 != 0) return fail(DC_ERR_INVALID, "dc_conv_forward: ldc=%d must be >= n*h*w and a multiple of 4", a->ldc);
    if (a->relu || a->residual) return fail(DC_ERR_UNSUPPORTED, "dc_conv_forward: channel-major output has no ReLU/residual epilogue");
  }

  dc::ConvParams p;
  memset(&p, 0, sizeof(p));
  int n = a->n, h = a->h, w = a->w;
  const bool pointwise = (a->kh == 1 && a->kw == 1 && a->pad == 0 && stride == 1);
  int out_h = ho, out_w = wo;
  if (pointwise) {   // a 1x1 conv is a plain GEMM over all pixels: flatten so tiles never straddle rows
    w = n * h * w; h = 1; n = 1;
    out_h = 1; out_w = w;
  }
  p.H = h; p.W = w; p.Ho = out_h; p.Wo = out_w; p.Cout = a->cout; p.Cin = a->cin;
  p.ntaps = a->kh * a->kw;
  // visiting order of the taps: row-major like the packed K axis, except for the 64 -> 64 channel 3x3 convs, which go column offset by
  // column offset (ConvParams::tap_kblk: the order TALL mode needs, used by the plain kernel too so that both give the same bits)
  const bool col_major_taps = a->kh == 3 && a->kw == 3 && a->cin == 64 && rows == 64;
  for (int pp = 0; pp < a->kh; ++pp)
    for (int q = 0; q < a->kw; ++q) {
      const int t = col_major_taps ? q * a->kh + pp : pp * a->kw + q;
      p.tap_dy[t] = -a->pad + pp * a->dilation;
      p.tap_dx[t] = -a->pad + q * a->dilation;
      p.tap_kblk[t] = pp * a->kw + q;
    }
  // tile rectangle with the least padded area
  const int cand[5][2] = {{128, 1}, {64, 2}, {32, 4}, {16, 8}, {8, 16}};
  long long best = -1;
  for (int i = 0; i < 5; ++i) {
    const int tw = cand[i][0], th = cand[i][1];
    if (tw * stride > 256 || th * stride > 256) continue;        // TMA box extent limit
    const long long area = static_cast<long long>((out_w + tw - 1) / tw) * tw * ((out_h + th - 1) / th) * th;
    if (best < 0 || area < best) { best = area; p.TW = tw; p.TH = th; }
  }
  p.in_stride = stride;
  for (p.log2_tw = 0; (1 << p.log2_tw) < p.TW; ++p.log2_tw) {}
  p.tiles_x = (out_w + p.TW - 1) / p.TW;
  p.tiles_y = (out_h + p.TH - 1) / p.TH;
  p.n_tiles_m = n * p.tiles_x * p.tiles_y;
  // Latency regime (a single image, the demo's case): when 128-channel tiles cannot give every other SM a unit, halve
  // the channel tile.  Twice the CTAs, each streaming 3/4 of the operand bytes per K-chunk; every output element keeps
  // its own K chain in the same order, so the result is bitwise the same as with 128-channel tiles (the packed rows
  // are a multiple of 128, hence of 64).  DC_SMALL_GRID_BN64=0 disables.
  static const bool small_grid_bn64 = [] { const char* e = getenv("DC_SMALL_GRID_BN64"); return !(e && e[0] == '0'); }();
  if (small_grid_bn64 && bn == 128 && a->out_f32_rows != 2 && static_cast<long long>(p.n_tiles_m) * (rows / 128) * 2 <= g_num_sms) bn = 64;
  p.n_tiles_n = rows / bn;
  p.scale = a->scale; p.shift = a->shift;
  p.res = static_cast<const __half*>(a->residual);
  const long long out_elems = static_cast<long long>(a->n) * ho * wo * a->cout;
  if ((a->x_plane && a->x_plane < static_cast<long long>(a->n) * a->h * a->w * a->cin) || (a->out_plane && a->out_plane < out_elems) ||
      (a->residual_plane && a->residual_plane < out_elems))
    return fail(DC_ERR_INVALID, "dc_conv_forward: a plane stride is smaller than the sub-batch it addresses");
  if ((a->x_plane || a->out_plane) && a->out_f32_rows) return fail(DC_ERR_UNSUPPORTED, "dc_conv_forward: sub-batch plane strides apply to split output only");
  p.res_plane = a->residual_plane > 0 ? a->residual_plane : out_elems;
  p.out = a->out;
  p.out_plane = a->out_plane > 0 ? a->out_plane : out_elems;
  p.ldc = a->ldc;
  p.relu = a->relu;
  p.out_mode = a->out_f32_rows == 2 ? dc::kOutF32RowsT : (a->out_f32_rows ? dc::kOutF32Rows : dc::kOutSplitNHWC);
  p.swap_ab = a->out_f32_rows == 2;
  p.early_weights = use_early_weights();
  { const char* e = getenv("DC_DEBUG_SKIP"); p.debug_skip = e ? (atoi(e) & 3) : 0; }      // microbenchmarks only: results are wrong
  // weights evict_last pays when every CTA re-reads the layer's weight tiles for many pixel tiles; a launch with fewer pixel tiles than
  // SMs (a single image) reads each weight tile once, and 251 MB of evict_last lines per forward would only push the activations out
  p.w_evict_last = (a->weights_evict_last && p.n_tiles_m >= g_num_sms) ? 1 : 0;
  p.sk_ws = static_cast<float*>(a->splitk_workspace);

  // 64 -> 64 channel 3x3 convs of a throughput-size batch (res2's branch2b): TALL mode -- one tall activation box per column offset
  // instead of one box per tap, all nine weight tiles resident in shared memory (conv_igemm.cuh).  These layers were bound by L2 -> SM
  // operand traffic (48 KB per 384 MMA cycles); the boxes are a third of those bytes.  Same K order per output element: bitwise the
  // same result.  DC_CONV_TALL=0 disables.
  if (use_tall() && a->kh == 3 && a->kw == 3 && stride == 1 && a->pad == a->dilation && a->cin == 64 && rows == 64 && !a->out_f32_rows &&
      use_2cta() && g_num_sms >= 2) {
    dc::ConvParams pt = p;
    choose_tall_tile(pt, n, out_h, out_w);
    const int gdx[3] = {-a->pad, -a->pad + a->dilation, -a->pad + 2 * a->dilation};
    int tall_rows = 0;
    if (pt.n_tiles_m >= g_num_sms && setup_tall(pt, 3, 3, gdx, -a->pad, a->dilation, &tall_rows)) {
      pt.n_tiles_n = 1;
      CUtensorMap tta, ttb, tto;
      if (int rc = encode_act_map(&tta, a->x, n, h, w, a->cin, pt.TW, tall_rows, 1, a->x_plane)) return rc;
      if (int rc = encode_out_map(&tto, a->out, n, out_h, out_w, a->cout, pt.TW, a->out_plane)) return rc;
      if (int rc = encode_w_map(&ttb, a->w_packed, rows, static_cast<long long>(pt.ntaps) * a->cin, 32)) return rc;
      return launch_conv<64, 2, 8, 0, 1>(tta, ttb, tto, pt, static_cast<cudaStream_t>(stream));
    }
  }
  CUtensorMap ta, tb, to;
  memset(&to, 0, sizeof(to));
  if (int rc = encode_act_map(&ta, a->x, n, h, w, a->cin, p.TW, p.TH, stride, a->x_plane)) return rc;
  if (!a->out_f32_rows) {
    // output geometry as the kernel indexes it (flattened for 1x1): [n][out_h][out_w][cout]
    if (int rc = encode_out_map(&to, a->out, n, out_h, out_w, a->cout, p.TW, a->out_plane)) return rc;
  }
  // CTA pairs for the 128-channel-tile convs (fewer operand bytes per SM, the lean epilogue); single CTAs with the fused
  // N = 2*BN MMA (conv_igemm.cuh) for the 64-channel tiles, the head GEMMs and single images, where they measure faster
  // (profiles/r1_microbench_wide_mma.txt, r2_pairs_3x3_sweep.jsonl).  DC_CONV_PAIR_ALL=1 forces pairs everywhere.
  // (the A/B switches below are read per launch -- launches happen at graph capture -- so one process can sweep them)
  const bool pair_all = [] { const char* e = getenv("DC_CONV_PAIR_ALL"); return e && e[0] == '1'; }();
  static const bool pair_lean_only = [] { const char* e = getenv("DC_CONV_PAIR_LEAN_ONLY"); return e && e[0] == '1'; }();
  const bool lean_on = [] { const char* e = getenv("DC_LEAN_EPILOGUE"); return !(e && e[0] == '0'); }();
  const bool lean_shape = lean_on && bn == 128 && p.out_mode == dc::kOutSplitNHWC && p.res != nullptr && p.ntaps * p.Cin <= 512 && a->cout >= 256;
  // Still fewer units than a quarter / half of the SMs and a K loop worth sharing (>= 16 K-steps; the exchange costs ~4.7 us,
  // a K-step ~0.4 us per CTA, profiles/r1_microbench_latency.txt): split-K clusters of 4 / 2 CTAs per unit.  The summation order over K changes (S partial chains added in rank order), so results differ
  // from the unsplit kernel by fp32 rounding; it is a pure function of the launch geometry, hence still deterministic.
  // dc_set_split_k(1) / DC_SPLIT_K=1 disables.
  int ksplit = 1;
  if (p.out_mode != dc::kOutF32RowsT && a->splitk_workspace != nullptr) {
    const long long units = static_cast<long long>(p.n_tiles_m) * p.n_tiles_n;
    const int ksteps = p.ntaps * (a->cin / dc::kBK);
    for (int s = 4; s >= 2; s -= 2)
      if (s <= g_split_k_max.load() && ksteps >= g_split_k_min_steps.load() && units <= (bn == 128 ? max_split_clusters<128>(s) : max_split_clusters<64>(s)) &&
          static_cast<size_t>(units) * (s - 1) * bn * 128 * 4 <= a->splitk_workspace_bytes) { ksplit = s; break; }
  }
  // Pairs for the 128-channel-tile 3x3 convs of a throughput-size batch as well (round 2, after MMA issue stopped being the bottleneck of
  // the pair kernels, profiles/r2_issue_lane.md): 48 instead of 64 KB of operands per K-step and SM and a fourth pipeline stage; res4 branch2b
  // 7.15 -> 6.39 ms per step, res3 branch2b 1.61 -> 1.45, res5's dilated 3x3 unchanged (profiles/r2_pairs_3x3_sweep.jsonl).  Bitwise the same
  // output.  DC_CONV_PAIR_3X3=0 restores single CTAs with the fused N = 2*BN MMA; a single image (fewer pixel tiles than SMs) keeps them.
  const bool pair_3x3 = [] { const char* e = getenv("DC_CONV_PAIR_3X3"); return !(e && e[0] == '0'); }();
  const bool pair = ksplit == 1 && use_2cta() && !p.swap_ab && g_num_sms >= 2 &&
                    (pair_all || (p.ntaps == 1 && bn == 128 && (!pair_lean_only || lean_shape)) || (pair_3x3 && p.ntaps > 1 && bn == 128 && p.n_tiles_m >= g_num_sms));
  // 256-channel tiles (conv_igemm<256, 2, 8>: one activation tile against 256 output channels, 2/3 of the operand bytes per MMA,
  // single-buffered accumulators; same per-element K chains: bitwise the same output).  While every MMA of the pair kernels sat in
  // ptxas's issue waterfall the wide tile's half as many, twice as long MMAs were worth 7-10 % per unit of work and it was taken
  // wherever it did not cost a wave (res5 branch2a, the projection shortcuts).  With the elect.sync issue path the 128-channel tiles are
  // faster there too (16 x 720p: res5 branch2a 0.87 -> 0.74 ms, res4 / res5 branch1 0.24 / 0.68 -> 0.18 / 0.57 ms per step,
  // profiles/r2_tile_choice_sweep.jsonl), so it is OFF by default: DC_CONV_BN256=1 = where it costs no wave, 2 = wherever legal.
  const bool bn256_on = [] { const char* e = getenv("DC_CONV_BN256"); return e && (e[0] == '1' || e[0] == '2'); }();
  bool wide256 = bn256_on && pair && !lean_shape && bn == 128 && p.out_mode == dc::kOutSplitNHWC && p.res == nullptr && rows % 256 == 0 &&
                 p.ntaps * (a->cin / dc::kBK) >= 8;
  if (wide256) {
    const long long pairs = persistent_sms() / 2, mp = (p.n_tiles_m + 1) / 2;
    const long long rounds128 = (mp * (rows / 128) + pairs - 1) / pairs, rounds256 = (mp * (rows / 256) + pairs - 1) / pairs;
    const bool force = [] { const char* e = getenv("DC_CONV_BN256"); return e && e[0] == '2'; }();      // 2 = wherever legal (A/B runs)
    wide256 = force ? mp * (rows / 256) >= pairs : (p.ntaps == 1 && rounds256 * 18 < rounds128 * 10 && mp * (rows / 256) >= pairs);
  }
  if (wide256) { bn = 256; p.n_tiles_n = rows / 256; }
  if (int rc = encode_w_map(&tb, a->w_packed, rows, static_cast<long long>(p.ntaps) * a->cin, pair ? bn / 2 : bn)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ksplit > 1) return bn == 128 ? launch_conv<128, 1, 8, 1>(ta, tb, to, p, st, ksplit) : launch_conv<64, 1, 8, 1>(ta, tb, to, p, st, ksplit);
  // epilogue-bound layers (short K, wide output: the 1x1 expand convs) get the 16-warp lean epilogue
  const bool lean = pair && lean_shape;
  if (lean) return launch_conv<128, 2, 16>(ta, tb, to, p, st);
  if (wide256) return launch_conv<256, 2, 8>(ta, tb, to, p, st);
  if (bn == 128) return pair ? launch_conv<128, 2>(ta, tb, to, p, st) : launch_conv<128, 1>(ta, tb, to, p, st);
  return pair ? launch_conv<64, 2>(ta, tb, to, p, st) : launch_conv<64, 1>(ta, tb, to, p, st);
}

// ------------------------------------------------------------------ tensor-core stem
int dc_conv1_tc_forward(const float* x, int n, int h, int w, const void* w_packed, const float* scale,
                        const float* shift, void* workspace, void* out, void* stream) {
  if (int rc = ensure_init()) return rc;
  if (!x || !w_packed || !scale || !shift || !workspace || !out) return fail(DC_ERR_INVALID, "dc_conv1_tc_forward: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int h2 = (h + 1) / 2, w2 = (w + 1) / 2, wp = w2 + 3;
  const long long plane = static_cast<long long>(n) * h2 * wp * 16;
  dc::stem_s2d_kernel<<<ew_grid(static_cast<long long>(n) * h2 * wp), 256, 0, st>>>(x, static_cast<__half*>(workspace), plane, n, h, w, h2, w2);
  g_launches++;
  DC_CUDA(cudaGetLastError());

  dc::ConvParams p;
  memset(&p, 0, sizeof(p));
  p.H = h2; p.W = w2; p.Ho = h2; p.Wo = w2; p.Cout = 64; p.Cin = 64;
  p.ntaps = 4;
  p.in_stride = 1;
  for (int t = 0; t < 4; ++t) { p.tap_dy[t] = t - 2; p.tap_dx[t] = 0; p.tap_kblk[t] = t; }
  const int cand[5][2] = {{128, 1}, {64, 2}, {32, 4}, {16, 8}, {8, 16}};
  long long best = -1;
  for (int i = 0; i < 5; ++i) {
    const int tw = cand[i][0], th = cand[i][1];
    const long long area = static_cast<long long>((w2 + tw - 1) / tw) * tw * ((h2 + th - 1) / th) * th;
    if (best < 0 || area < best) { best = area; p.TW = tw; p.TH = th; }
  }
  for (p.log2_tw = 0; (1 << p.log2_tw) < p.TW; ++p.log2_tw) {}
  p.tiles_x = (w2 + p.TW - 1) / p.TW;
  p.tiles_y = (h2 + p.TH - 1) / p.TH;
  p.n_tiles_m = n * p.tiles_x * p.tiles_y;
  p.n_tiles_n = 1;
  p.scale = scale; p.shift = shift;
  p.out = out;
  p.out_plane = static_cast<long long>(n) * h2 * w2 * 64;
  p.relu = 1;
  p.out_mode = dc::kOutSplitNHWC;
  p.early_weights = use_early_weights();

  // TALL mode (conv_igemm.cuh): the four taps are the same 4-pixel windows one row apart -- one box of TH + 3 rows per tile instead of
  // four boxes of TH rows, the four weight tiles resident in shared memory.  DC_CONV_TALL=0 disables.
  bool tall = false;
  int box_rows = p.TH;
  if (use_tall() && use_2cta() && g_num_sms >= 2) {
    dc::ConvParams pt = p;
    choose_tall_tile(pt, n, h2, w2);
    const int gdx[1] = {0};
    int tall_rows = 0;
    if (pt.n_tiles_m >= g_num_sms && setup_tall(pt, 1, 4, gdx, -2, 1, &tall_rows)) { p = pt; box_rows = tall_rows; tall = true; }
  }
  // A: overlapping 4-pixel windows of the padded space-to-depth image (W stride = one 16-channel pixel)
  CUtensorMap ta, tb, to;
  {
    const cuuint64_t dims[5] = {64, (cuuint64_t)w2, (cuuint64_t)h2, (cuuint64_t)n, 2};
    const cuuint64_t strides[4] = {32, (cuuint64_t)wp * 32, (cuuint64_t)h2 * wp * 32, (cuuint64_t)plane * 2};
    const cuuint32_t box[5] = {64, (cuuint32_t)p.TW, (cuuint32_t)box_rows, 1, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, workspace, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DC_ERR_CUDA, "cuTensorMapEncodeTiled(stem windows) failed: %d", (int)r);
  }
  static const bool pair_all = [] { const char* e = getenv("DC_CONV_PAIR_ALL"); return e && e[0] == '1'; }();
  const bool pair = tall || (use_2cta() && g_num_sms >= 2 && pair_all);       // BN = 64 otherwise: single CTAs + fused wide MMA
  if (int rc = encode_w_map(&tb, w_packed, 64, 256, pair ? 32 : 64)) return rc;
  if (int rc = encode_out_map(&to, out, n, h2, w2, 64, p.TW)) return rc;
  if (tall) return launch_conv<64, 2, 8, 0, 1>(ta, tb, to, p, st);
  return pair ? launch_conv<64, 2>(ta, tb, to, p, st) : launch_conv<64, 1>(ta, tb, to, p, st);
}

// ------------------------------------------------------------------ HBM kernels
int dc_conv1_forward(const float* x, int n, int h, int w, const float* w147x64, const float* scale,
                     const float* shift, void* out, void* stream) {
  if (int rc = ensure_init()) return rc;
  if (!x || !w147x64 || !scale || !shift || !out) return fail(DC_ERR_INVALID, "dc_conv1_forward: null argument");
  const int ho = (h + 6 - 7) / 2 + 1, wo = (w + 6 - 7) / 2 + 1;
  const int tiles = n * ((ho + dc::kC1TileH - 1) / dc::kC1TileH) * ((wo + dc::kC1TileW - 1) / dc::kC1TileW);
  dc::conv1_7x7s2_kernel<<<tiles, 256, dc::kC1SmemFloats * 4, static_cast<cudaStream_t>(stream)>>>(
      x, w147x64, scale, shift, static_cast<__half*>(out), static_cast<long long>(n) * ho * wo * 64, n, h, w, ho, wo);
  g_launches++;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

int dc_maxpool_forward(const void* x, int n, int h, int w, int c, int kernel, int stride, void* out, void* stream) {
  if (int rc = ensure_init()) return rc;
  if (!x || !out || c % 8 != 0) return fail(DC_ERR_INVALID, "dc_maxpool_forward: bad arguments (c must be a multiple of 8)");
  const int ho = dc_pool_out_size(h, kernel, stride), wo = dc_pool_out_size(w, kernel, stride);
  const long long total = static_cast<long long>(n) * ho * wo * (c / 8);
  if (kernel < 1 || stride < 1 || h < 1 || w < 1) return fail(DC_ERR_INVALID, "dc_maxpool_forward: bad geometry");
  const __half* xi = static_cast<const __half*>(x);
  __half* xo = static_cast<__half*>(out);
  const long long ip = static_cast<long long>(n) * h * w * c, op = static_cast<long long>(n) * ho * wo * c;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kernel == 3) dc::maxpool_split_kernel<3><<<ew_grid(total), 256, 0, st>>>(xi, ip, xo, op, n, h, w, c, ho, wo, kernel, stride);
  else dc::maxpool_split_kernel<0><<<ew_grid(total), 256, 0, st>>>(xi, ip, xo, op, n, h, w, c, ho, wo, kernel, stride);
  g_launches++;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

int dc_subsample_forward(const void* x, int n, int h, int w, int c, int stride, void* out, void* stream) {
  if (int rc = ensure_init()) return rc;
  if (!x || !out || c % 8 != 0 || stride < 1) return fail(DC_ERR_INVALID, "dc_subsample_forward: bad arguments");
  const int ho = (h - 1) / stride + 1, wo = (w - 1) / stride + 1;
  const long long total = static_cast<long long>(n) * ho * wo * (c / 8) * 2;
  dc::subsample_split_kernel<<<ew_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), static_cast<long long>(n) * h * w * c, static_cast<__half*>(out),
      static_cast<long long>(n) * ho * wo * c, n, h, w, c, ho, wo, stride);
  g_launches++;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

int dc_head_finish(const float* col, long long ldcol, int col_off, const float* skip, long long ldskip, int skip_off,
                   float* out, int n, int cout, int h, int w, int ho, int wo, int sigmoid, void* stream) {
  if (int rc = ensure_init()) return rc;
  if (!col || !skip || !out) return fail(DC_ERR_INVALID, "dc_head_finish: null argument");
  if (n <= 0 || cout <= 0 || h <= 0 || w <= 0 || ho <= 0 || wo <= 0 || ho > 2 * h + 1 || wo > 2 * w + 1)
    return fail(DC_ERR_INVALID, "dc_head_finish: bad geometry (%dx%d -> %dx%d)", h, w, ho, wo);
  const int cells = ((ho + 1) / 2) * ((wo + 1) / 2);
  const dim3 grid(static_cast<unsigned>(n) * cout, static_cast<unsigned>((cells + 255) / 256));
  // float2 rows: even width, even plane strides, 8-byte-aligned bases
  const bool vec = (wo % 2 == 0) && (ldskip % 2 == 0) && (reinterpret_cast<uintptr_t>(out) % 8 == 0) &&
                   (reinterpret_cast<uintptr_t>(skip) % 8 == 0) && ((static_cast<long long>(ho) * wo) % 2 == 0);
  if (vec)
    dc::head_finish_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(col, ldcol, col_off, skip, ldskip, skip_off, out, n,
                                                                                       cout, h, w, ho, wo, sigmoid);
  else
    dc::head_finish_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(col, ldcol, col_off, skip, ldskip, skip_off, out, n,
                                                                                        cout, h, w, ho, wo, sigmoid);
  g_launches++;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

int dc_images_u8_to_blob(const unsigned char* img, int n, int h, int w, const float* mean3, float* out, void* stream) {
  if (int rc = ensure_init()) return rc;
  if (!img || !out || !mean3 || n <= 0 || h <= 0 || w <= 0) return fail(DC_ERR_INVALID, "dc_images_u8_to_blob: bad arguments");
  const long long pixels = static_cast<long long>(n) * h * w;
  const long long hw = static_cast<long long>(h) * w;
  if (hw > 0x7FFFFFFF) return fail(DC_ERR_UNSUPPORTED, "dc_images_u8_to_blob: image too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = (w % 4 == 0) && (reinterpret_cast<uintptr_t>(img) % 4 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  if (vec) dc::images_u8_to_blob_kernel<<<ew_grid(pixels / 4), 256, 0, st>>>(img, out, pixels / 4, static_cast<int>(hw), mean3[0], mean3[1], mean3[2]);
  else dc::images_u8_to_blob_scalar_kernel<<<ew_grid(pixels), 256, 0, st>>>(img, out, pixels, static_cast<int>(hw), mean3[0], mean3[1], mean3[2]);
  g_launches++;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

int dc_pose_from_maps(const float* prob, const float* loc, int n, int joints, int h, int w, float stride,
                      float locref_scale, float scale, float* out, void* stream) {
  if (int rc = ensure_init()) return rc;
  if (!prob || !loc || !out || n <= 0 || joints <= 0 || h <= 0 || w <= 0 || scale == 0.f) return fail(DC_ERR_INVALID, "dc_pose_from_maps: bad arguments");
  dc::pose_from_maps_kernel<<<n * joints, 256, 0, static_cast<cudaStream_t>(stream)>>>(prob, loc, joints, h, w, stride, locref_scale, scale, out);
  g_launches++;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

// ------------------------------------------------------------------ demo pre-processing (estimate_pose.py:83-105)
namespace {
// Pillow's resampling table for the bilinear filter over a full axis (Resample.c: precompute_coeffs followed by
// normalize_coeffs_8bpc), restated: the triangle filter's support widens by the shrink factor, taps are normalised in
// double precision and rounded to 22-bit fixed point.
struct ResampleTable {
  int ksize = 0;
  std::vector<int> bounds;      // {first, count} per output index
  std::vector<int> kk;          // [out][ksize]
};

ResampleTable build_resample_table(int in_size, int out_size) {
  ResampleTable t;
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  t.ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  t.bounds.assign(2 * static_cast<size_t>(out_size), 0);
  t.kk.assign(static_cast<size_t>(out_size) * t.ksize, 0);
  const double ss = 1.0 / filterscale;
  std::vector<double> w(t.ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0) a = -a;
      w[x] = a < 1.0 ? 1.0 - a : 0.0;
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      const double v = (ww != 0.0 ? w[x] / ww : w[x]) * (1 << 22);
      t.kk[static_cast<size_t>(xx) * t.ksize + x] = static_cast<int>(v < 0 ? -0.5 + v : 0.5 + v);
    }
    t.bounds[2 * xx] = xmin;
    t.bounds[2 * xx + 1] = xmax;
  }
  return t;
}
}  // namespace

struct dc_preprocess_plan {
  int h = 0, w = 0;                 // source image
  int res_h = 0, res_w = 0;         // rescaled padded image: int((h+64)*scale), int((w+64)*scale)
  int out_h = 0, out_w = 0;         // net input: ceil(h*scale/8)*8, ceil(w*scale/8)*8
  int valid_h = 0, valid_w = 0;     // min(out, res)
  int mid_rows = 0;                 // rows of the horizontal pass the vertical pass reads
  bool hpass = false, vpass = false;
  int ksize_x = 0, ksize_y = 0;
  int *bounds_x = nullptr, *kk_x = nullptr, *bounds_y = nullptr, *kk_y = nullptr;   // device
};

int dc_preprocess_plan_create(int h, int w, double scale, dc_preprocess_plan** plan) {
  if (int rc = ensure_init()) return rc;
  if (!plan || h <= 0 || w <= 0 || !(scale > 0.0)) return fail(DC_ERR_INVALID, "dc_preprocess_plan_create: bad arguments");
  const int pad = 64, stride = 8;
  auto p = std::make_unique<dc_preprocess_plan>();
  p->h = h;
  p->w = w;
  p->res_h = static_cast<int>((h + pad) * scale);
  p->res_w = static_cast<int>((w + pad) * scale);
  p->out_h = static_cast<int>(std::ceil(static_cast<double>(h) * scale / stride) * stride);
  p->out_w = static_cast<int>(std::ceil(static_cast<double>(w) * scale / stride) * stride);
  if (p->res_h <= 0 || p->res_w <= 0 || p->out_h <= 0 || p->out_w <= 0) return fail(DC_ERR_INVALID, "dc_preprocess_plan_create: scale %g leaves no pixels", scale);
  p->valid_h = std::min(p->out_h, p->res_h);
  p->valid_w = std::min(p->out_w, p->res_w);
  p->hpass = p->res_w != w + pad;
  p->vpass = p->res_h != h + pad;
  p->mid_rows = p->valid_h;
  auto upload = [](const std::vector<int>& v, int** dev) -> int {
    DC_CUDA(cudaMalloc(dev, v.size() * sizeof(int)));
    DC_CUDA(cudaMemcpy(*dev, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
    return DC_OK;
  };
  int rc = DC_OK;
  if (p->vpass) {
    ResampleTable t = build_resample_table(h + pad, p->res_h);
    p->ksize_y = t.ksize;
    p->mid_rows = t.bounds[2 * (p->valid_h - 1)] + t.bounds[2 * (p->valid_h - 1) + 1];
    if ((rc = upload(t.bounds, &p->bounds_y)) || (rc = upload(t.kk, &p->kk_y))) { dc_preprocess_plan_destroy(p.release()); return rc; }
  }
  if (p->hpass) {
    ResampleTable t = build_resample_table(w + pad, p->res_w);
    p->ksize_x = t.ksize;
    if ((rc = upload(t.bounds, &p->bounds_x)) || (rc = upload(t.kk, &p->kk_x))) { dc_preprocess_plan_destroy(p.release()); return rc; }
  }
  *plan = p.release();
  return DC_OK;
}

int dc_preprocess_plan_info(const dc_preprocess_plan* plan, int* out_h, int* out_w, size_t* workspace_bytes) {
  if (!plan) return fail(DC_ERR_INVALID, "dc_preprocess_plan_info: null plan");
  if (out_h) *out_h = plan->out_h;
  if (out_w) *out_w = plan->out_w;
  if (workspace_bytes) *workspace_bytes = plan->hpass ? static_cast<size_t>(plan->mid_rows) * plan->valid_w * 3 : 0;
  return DC_OK;
}

int dc_preprocess_plan_destroy(dc_preprocess_plan* plan) {
  if (!plan) return DC_OK;
  cudaFree(plan->bounds_x);
  cudaFree(plan->kk_x);
  cudaFree(plan->bounds_y);
  cudaFree(plan->kk_y);
  delete plan;
  return DC_OK;
}

int dc_preprocess_u8_forward(const dc_preprocess_plan* p, const unsigned char* img, const float* mean3, float* out, void* workspace,
                             void* stream) {
  if (int rc = ensure_init()) return rc;
  if (!p || !img || !mean3 || !out) return fail(DC_ERR_INVALID, "dc_preprocess_u8_forward: null argument");
  if (p->hpass && !workspace) return fail(DC_ERR_INVALID, "dc_preprocess_u8_forward: this plan needs a workspace (dc_preprocess_plan_info)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned char* src = img;
  if (p->hpass) {
    dim3 grid((p->valid_w + 255) / 256, p->mid_rows);
    dc::preprocess_hpass_kernel<<<grid, 256, 0, st>>>(img, p->h, p->w, p->mid_rows, p->valid_w, reinterpret_cast<const int2*>(p->bounds_x),
                                                      p->kk_x, p->ksize_x, static_cast<unsigned char*>(workspace));
    g_launches++;
    DC_CUDA(cudaGetLastError());
    src = static_cast<const unsigned char*>(workspace);
  }
  dim3 grid((p->out_w + 255) / 256, p->out_h);
  const int2* by = reinterpret_cast<const int2*>(p->bounds_y);
#define DC_PRE_FINISH(HP, VP)                                                                                                        \
  dc::preprocess_finish_kernel<HP, VP><<<grid, 256, 0, st>>>(src, p->h, p->w, p->valid_w, p->valid_h, p->valid_w, by, p->kk_y, p->ksize_y, \
                                                              mean3[0], mean3[1], mean3[2], out, p->out_h, p->out_w)
  if (p->hpass && p->vpass) DC_PRE_FINISH(true, true);
  else if (p->hpass) DC_PRE_FINISH(true, false);
  else if (p->vpass) DC_PRE_FINISH(false, true);
  else DC_PRE_FINISH(false, false);
#undef DC_PRE_FINISH
  g_launches++;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

int dc_nchw_to_split(const float* x, int n, int c, int h, int w, void* out, void* stream) {
  if (int rc = ensure_init()) return rc;
  if (!x || !out) return fail(DC_ERR_INVALID, "dc_nchw_to_split: null argument");
  const long long total = static_cast<long long>(n) * c * h * w;
  dc::nchw_to_split_kernel<<<ew_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__half*>(out), total, n, c, h, w);
  g_launches++;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

int dc_split_to_nchw(const void* x, int n, int c, int h, int w, float* out, void* stream) {
  if (int rc = ensure_init()) return rc;
  if (!x || !out) return fail(DC_ERR_INVALID, "dc_split_to_nchw: null argument");
  const long long total = static_cast<long long>(n) * c * h * w;
  dc::split_to_nchw_kernel<<<ew_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), total, out, n, c, h, w);
  g_launches++;
  DC_CUDA(cudaGetLastError());
  return DC_OK;
}

// ------------------------------------------------------------------ per-layer NCHW kernels
#define DC_EW_LAUNCH(kernel, total, ...)                                                        \
  do {                                                                                          \
    if (int rc = ensure_init()) return rc;                                                      \
    if ((total) > 0) {                                                                          \
      kernel<<<ew_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(__VA_ARGS__);       \
      g_launches++;                                                                             \
      DC_CUDA(cudaGetLastError());                                                              \
    }                                                                                           \
    return DC_OK;                                                                               \
  } while (0)

int dc_bn_forward_nchw(const float* x, const float* mean, const float* stddev, int n, int c, int hw, float* y, void* stream) {
  const long long total = static_cast<long long>(n) * c * hw;
  DC_EW_LAUNCH(dc::bn_nchw_kernel, total, x, mean, stddev, total, c, hw, y);
}
int dc_scale_forward_nchw(const float* x, const float* gamma, const float* beta, int n, int c, int hw, float* y, void* stream) {
  const long long total = static_cast<long long>(n) * c * hw;
  DC_EW_LAUNCH(dc::scale_nchw_kernel, total, x, gamma, beta, total, c, hw, y);
}
int dc_relu_forward(const float* x, long long count, float negative_slope, float* y, void* stream) {
  DC_EW_LAUNCH(dc::relu_kernel, count, x, count, negative_slope, y);
}
int dc_sigmoid_forward(const float* x, long long count, float* y, void* stream) {
  DC_EW_LAUNCH(dc::sigmoid_kernel, count, x, count, y);
}
int dc_axpby_forward(const float* a, float ca, const float* b, float cb, long long count, float* y, void* stream) {
  DC_EW_LAUNCH(dc::axpby_kernel, count, a, ca, b, cb, count, y);
}
int dc_crop_forward_nchw(const float* x, int n, int c, int h, int w, int off_h, int off_w, int ho, int wo, float* y, void* stream) {
  const long long total = static_cast<long long>(n) * c * ho * wo;
  DC_EW_LAUNCH(dc::crop_nchw_kernel, total, x, h, w, off_h, off_w, ho, wo, total, y);
}
int dc_maxpool_forward_nchw(const float* x, int n, int c, int h, int w, int kh, int kw, int sh, int sw, int ph, int pw,
                            int ho, int wo, float* y, void* stream) {
  const long long total = static_cast<long long>(n) * c * ho * wo;
  DC_EW_LAUNCH(dc::maxpool_nchw_kernel, total, x, h, w, kh, kw, sh, sw, ph, pw, ho, wo, total, y);
}
int dc_conv_direct_nchw(const float* x, const float* w, const float* bias, int n, int cin, int h, int wd, int cout,
                        int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, float* y, void* stream) {
  const int ho = (h + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1, wo = (wd + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
  const long long total = static_cast<long long>(n) * cout * ho * wo;
  DC_EW_LAUNCH(dc::conv_direct_nchw_kernel, total, x, w, bias, cin, h, wd, cout, kh, kw, sh, sw, ph, pw, dh, dw, ho, wo, total, y);
}
int dc_deconv_direct_nchw(const float* x, const float* w, const float* bias, int n, int cin, int h, int wd, int cout,
                          int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, float* y, void* stream) {
  const int ho = sh * (h - 1) + dh * (kh - 1) + 1 - 2 * ph, wo = sw * (wd - 1) + dw * (kw - 1) + 1 - 2 * pw;
  const long long total = static_cast<long long>(n) * cout * ho * wo;
  DC_EW_LAUNCH(dc::deconv_direct_nchw_kernel, total, x, w, bias, cin, h, wd, cout, kh, kw, sh, sw, ph, pw, dh, dw, ho, wo, total, y);
}

}  // extern "C"
