// Host side of the tcgen05 GEMM engine: tensor-map construction, tile-shape selection and the C-ABI entry points
// gc_conv_fprop / gc_conv_dgrad / gc_conv_wgrad / gc_linear_fwd / gc_linear_dgrad / gc_linear_wgrad.
// Reference op sites: tools/model.py:131-164 (conv stack), :89-128 (FC body/head), algo/wdgail.py:26-32 (critic trunk),
// algo/wdgail.py:56-98 (gradient-penalty chains reuse dgrad / fprop-with-mask / wgrad).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

#include "gc_umma_kernel.cuh"
#include "../../include/gail_carla_b200.h"

namespace {
using namespace gcu;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

struct MapSpec {
  const void* ptr;
  uint64_t dim[5];
  uint64_t stride[4];  // bytes, dims 1..4
  uint32_t box[5];
  int tf32;     // operand maps round fp32 -> tf32 (RN) inside TMA; output / mask maps stay fp32
  int swizzle;  // 0 none | 1 SWIZZLE_128B (K-major operands, output staging) | 2 SWIZZLE_128B_ATOM_32B (MN-major tf32 operands)
};

std::unordered_map<std::string, CUtensorMap>& map_cache() {
  static std::unordered_map<std::string, CUtensorMap> c;
  return c;
}
std::mutex& map_mutex() {
  static std::mutex m;
  return m;
}

int make_map(const MapSpec& s, CUtensorMap* out) {
  std::string key((const char*)&s, sizeof(MapSpec));
  {
    std::lock_guard<std::mutex> g(map_mutex());
    auto it = map_cache().find(key);
    if (it != map_cache().end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return gc::fail(-3, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 5; ++i) { dims[i] = s.dim[i]; box[i] = s.box[i]; }
  for (int i = 0; i < 4; ++i) strides[i] = s.stride[i];
  static const int exp_map = getenv("GC_EXP_MAP") ? atoi(getenv("GC_EXP_MAP")) : 0;   // experiment: 1 = plain FLOAT32 operand maps, 2 = no L2 promotion, 4 = 128B promotion
  CUresult r = fn(out, (s.tf32 && !(exp_map & 1)) ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)s.ptr, dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  s.swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                 : (s.swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  (exp_map & 2) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : ((exp_map & 4) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B),
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return gc::fail((int)r,
                    "cuTensorMapEncodeTiled failed (%d): ptr=%p dim=[%llu,%llu,%llu,%llu,%llu] stride=[%llu,%llu,%llu,%llu] "
                    "box=[%u,%u,%u,%u,%u]",
                    (int)r, s.ptr, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                    (unsigned long long)dims[3], (unsigned long long)dims[4], (unsigned long long)strides[0],
                    (unsigned long long)strides[1], (unsigned long long)strides[2], (unsigned long long)strides[3], box[0],
                    box[1], box[2], box[3], box[4]);
  std::lock_guard<std::mutex> g(map_mutex());
  if (map_cache().size() > 4096) map_cache().clear();
  map_cache()[key] = *out;
  return 0;
}

// MapSpec builder: dims given innermost first; missing dims are padded with extent 1.
MapSpec spec(const void* ptr, int rank, const uint64_t* dim, const uint64_t* stride_elems, const uint32_t* box, int tf32,
             int swizzle) {
  MapSpec s;
  memset(&s, 0, sizeof(s));
  s.ptr = ptr;
  uint64_t last = 16;
  for (int i = 0; i < 5; ++i) {
    s.dim[i] = i < rank ? dim[i] : 1;
    s.box[i] = i < rank ? box[i] : 1;
    if (i >= 1) {
      if (i < rank) s.stride[i - 1] = stride_elems[i] * 4;
      else s.stride[i - 1] = last;
      last = s.stride[i - 1] * std::max<uint64_t>(1, s.dim[i]);
    } else {
      last = std::max<uint64_t>(16, s.dim[0] * 4);
    }
  }
  s.tf32 = tf32;
  s.swizzle = swizzle;
  return s;
}

int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

struct Plan {
  GemmParams p;
  dim3 grid;
  bool a_mn, b_mn;
  int b_bytes_override = 0;
  bool pair = false;   // CTA-pair (cta_group::2) launch: 2-CTA clusters, B tile split between the CTAs
  Plan() { memset(&p, 0, sizeof(p)); p.e0 = p.e1 = p.f0 = p.g0 = p.g1 = 1; a_mn = b_mn = false; }
};

constexpr int kSmemBudget = 225 * 1024;
constexpr int kSmemFixed = 5120;   // mbarriers, TMEM slot, descriptor table (256 B), bias (2 KB), column sums (1 KB), 1 KB alignment slack

// CTA-pair mode pays when the B tile is a large share of the per-k-iteration TMA rows (N >= 128) and needs an even
// number of m-tiles (a pair works on m-tiles 2j, 2j+1 of the same n-tile) and the two-group bit-mask / plain epilogues.
// An odd m-tile count is made even by the caller with one phantom tile row past the end of the outermost M dimension
// (its TMA loads are zero-filled, its stores clipped, its bit-mask words skipped).
bool pair_ok(long m_tiles, int bn, int epilogue, bool tma_mask) {
  static const bool off = getenv("GC_NO_CTA2") != nullptr;
  return !off && m_tiles >= 16 && bn >= 128 && bn % 32 == 0 && !(epilogue == EPI_MASK && tma_mask);
}

int finish_and_launch(Plan& pl, cudaStream_t st, const char* what) {
  GemmParams& p = pl.p;
  GC_REQUIRE(p.bn % 16 == 0 && p.bn >= 16 && (p.bn <= 256 || p.ngroups > 0), "%s: tile N %d not a multiple of 16 in [16,256]", what, p.bn);
  GC_REQUIRE(p.bk % 8 == 0 && p.bk >= 8, "%s: bk %d not a multiple of 8", what, p.bk);
  if (!pl.a_mn || !pl.b_mn) GC_REQUIRE(p.bk == 32, "%s: K-major operands need bk == 32 (got %d)", what, p.bk);
  // MN-major A: only the panels that exist are staged (Cout = 32 -> 1 of the 4 panels an M = 128 MMA reads; the other
  // three read whatever follows in shared memory into accumulator rows that are never stored)
  if (p.a_bytes == 0) p.a_bytes = pl.a_mn ? ((p.a_panels * p.bk * 128 + 1023) & ~1023) : 16384;
  if (p.taps == 0) p.taps = 1;
  if (p.row_box[0] == 0) { p.row_box[0] = 128; p.row_box[1] = 1; p.row_box[2] = 1; }
  const int bpan = (p.bn + 31) / 32;
  if (p.b_resident) {
    p.b_bytes = 0;
    p.b_slab_bytes = (p.mma_n ? p.mma_n : p.bn) * 128;
    GC_REQUIRE(!pl.b_mn && pl.grid.y == 1 && p.k_iters <= 16 && p.taps <= 4 && p.bk == 32,
               "%s: resident-B mode needs K-major B, one n-tile, <= 16 k-iterations of <= 4 taps", what);
  } else {
    p.b_slabs = 0; p.b_slab_bytes = 0;
    p.b_bytes = pl.b_mn ? bpan * p.bk * 128 : (pl.pair ? p.bn / 2 : p.bn) * 128;
    if (pl.b_bytes_override) p.b_bytes = pl.b_bytes_override;
    p.b_bytes = (p.b_bytes + 1023) & ~1023;
  }
  const int resident = p.b_slabs * p.b_slab_bytes;
  p.nbuf = (p.epilogue == EPI_MASK && p.bits_in == nullptr) ? (p.b_resident ? 3 : 4) : 2;
  if (p.nbuf == 2) {
    // two staging buffers per epilogue group (the TMA store of panel i drains while panel i+1 is formed) when that still
    // leaves a deep operand ring - in practice the small-tile streaming layers (conv1), whose epilogue is the bottleneck
    const int stage_b = p.a_bytes + p.b_bytes;
    if ((kSmemBudget - (4 * 16384 + kSmemFixed + resident)) / stage_b >= 6) p.nbuf = 4;
  }
  if (p.acc_stages == 0) p.acc_stages = 2;
  p.tmem_cols = pow2_cols(p.acc_stages * bpan * 32);
  GC_REQUIRE(p.tmem_cols <= 512, "%s: accumulators need %d TMEM columns", what, p.tmem_cols);
  p.d_row_bytes = p.bn >= 32 ? 128 : p.bn * 4;
  const int stage = p.a_bytes + p.b_bytes;
  const int fixed = p.nbuf * 16384 + kSmemFixed + resident;  // staging + barriers/bias/column sums/alignment slack + resident weights
  int stages = (kSmemBudget - fixed) / stage;
  stages = std::min(stages, 8);
  GC_REQUIRE(stages >= 2, "%s: tile does not fit in shared memory (stage %d B, fixed %d B)", what, stage, fixed);
  p.stages = stages;
  p.mt = pl.grid.x; p.nt = pl.grid.y; p.zt = pl.grid.z;
  // the kernel walks tiles as mixed-radix digits (m0 m1 m2 | n0 n1 | z): the radices must nest exactly
  GC_REQUIRE(p.mt % (p.e0 * p.e1) == 0 && p.nt % p.f0 == 0, "%s: tile grid %dx%d does not nest in (%d,%d | %d)", what, p.mt, p.nt,
             p.e0, p.e1, p.f0);
  if (p.cols_per_map > 0) {
    GC_REQUIRE((p.cols_per_map & (p.cols_per_map - 1)) == 0, "%s: cols_per_map %d must be a power of two", what, p.cols_per_map);
    p.cpm_shift = 0;
    while ((1 << p.cpm_shift) < p.cols_per_map) ++p.cpm_shift;
  }
  const long total_tiles = (long)p.mt * p.nt * p.zt;
  GC_REQUIRE(total_tiles > 0 && total_tiles < (1L << 31), "%s: bad tile count", what);
  const dim3 grid((unsigned)std::min<long>(total_tiles, gc::kNumSMs), 1, 1);
  const size_t smem = (size_t)stages * stage + fixed;
  if (pl.pair) {
    GC_REQUIRE(!pl.a_mn && !pl.b_mn && p.ngroups == 0 && !p.b_resident && p.taps == 1 && total_tiles % 2 == 0 && p.mt % 2 == 0,
               "%s: CTA-pair mode needs K-major operands, plain tiles and an even number of m-tiles", what);
    p.pair = 1;
  }
  auto launch = [&](auto kern) -> int {
    static thread_local const void* configured[3] = {nullptr, nullptr, nullptr};
    (void)configured;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return gc::fail((int)e, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    static long long* d_stats = nullptr;   // debug only: GC_UMMA_STATS=1 prints where each role of the kernel waits
#ifdef GC_UMMA_STATS_BUILD
    static const bool want_stats = getenv("GC_UMMA_STATS") != nullptr;
#else
    static const bool want_stats = false;   // counters are compiled out of the production kernel (build with GC_UMMA_STATS_BUILD=1)
#endif
    if (want_stats) {
      if (!d_stats) cudaMalloc(&d_stats, gc::kNumSMs * 16 * sizeof(long long));
      cudaMemsetAsync(d_stats, 0, gc::kNumSMs * 16 * sizeof(long long), st);
      p.stats = d_stats;
    }
    if (pl.pair) {   // 2-CTA clusters: CTAs (2c, 2c+1) form the pairs
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = grid; cfg.blockDim = dim3(320, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, kern, p);
      if (e != cudaSuccess) return gc::fail((int)e, "%s: cudaLaunchKernelEx (2-CTA clusters): %s", what, cudaGetErrorString(e));
    } else {
      kern<<<grid, 320, smem, st>>>(p);
    }
    if (want_stats) {
      long long h[gc::kNumSMs * 16];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost);
      double m[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (unsigned c = 0; c < grid.x; ++c) for (int i = 0; i < 16; ++i) m[i] += (double)h[c * 16 + i] / grid.x;
      fprintf(stderr, "[umma-stats] %-26s %sctas=%u tiles/cta=%.1f k_iters=%d bn=%d stages=%d | clocks/cta total=%.0f prod_wait_empty=%.0f "
                      "mma_wait_full=%.0f mma_wait_tmem=%.0f epi0_wait_acc=%.0f epi1_wait_acc=%.0f epi0_total=%.0f | epi0 per panel (%.0f panels): free=%.0f ld=%.0f "
                      "math=%.0f sts+fence=%.0f bar=%.0f store=%.0f nbuf=%d\n",
              what, pl.pair ? "PAIR " : "", grid.x, m[7], p.k_iters, p.bn, p.stages, m[5], m[0], m[1], m[2], m[3], m[4], m[6], m[12],
              m[8] / std::max(1.0, m[12]), m[9] / std::max(1.0, m[12]), m[10] / std::max(1.0, m[12]), m[13] / std::max(1.0, m[12]), m[14] / std::max(1.0, m[12]), m[11] / std::max(1.0, m[12]), p.nbuf);
    }
    return gc::launch_status(what);
  };
  if (p.ngroups > 0) {
    GC_REQUIRE(pl.a_mn && pl.b_mn, "%s: slab mode needs MN-major operands", what);
    return launch(umma_gemm_kernel<true, true, true>);
  }
  if (pl.pair) return launch(umma_gemm_kernel<false, false, false, true>);
  if (!pl.a_mn && !pl.b_mn) return launch(umma_gemm_kernel<false, false, false>);
  if (!pl.a_mn && pl.b_mn) return launch(umma_gemm_kernel<false, true, false>);
  if (pl.a_mn && pl.b_mn) return launch(umma_gemm_kernel<true, true, false>);
  return gc::fail(-2, "%s: unsupported operand majors", what);
}

// ---- pixel-box selection -------------------------------------------------------------------
struct PixBox { int ox, oy, b; };

// Tile of output pixels (ox fastest, then oy, then samples) with at most `max_rows` rows whose row count is a
// multiple of `mult`; maximise useful rows / issued rows.
PixBox choose_box(int OW, int OH, int B, int max_rows, int mult, double small_penalty) {
  PixBox best{1, 1, 1};
  double best_score = -1.0;
  for (int bx = 1; bx <= std::min(OW, max_rows); ++bx) {
    for (int by = 1; by <= OH && bx * by <= max_rows; ++by) {
      const int bb_max = (bx == OW) ? std::min(B, max_rows / (bx * by)) : 1;
      for (int bb = 1; bb <= bb_max; ++bb) {
        const int rows = bx * by * bb;
        if (rows % mult) continue;
        const long tiles = (long)((OW + bx - 1) / bx) * ((OH + by - 1) / by) * ((B + bb - 1) / bb);
        const double issued = (double)tiles * (small_penalty > 0 ? rows : max_rows);
        double score = (double)OW * OH * B / issued;
        if (small_penalty > 0) score -= small_penalty / rows;  // fewer, fatter k-iterations amortise barriers
        if (score > best_score + 1e-9) { best_score = score; best = PixBox{bx, by, bb}; }
      }
    }
  }
  return best;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

int check_geom(const gc_conv_geom* g, const char* what) {
  GC_REQUIRE(g != nullptr, "%s: null geometry", what);
  GC_REQUIRE(g->B > 0 && g->Cin > 0 && g->Cout > 0 && g->KH > 0 && g->KW > 0 && g->S > 0, "%s: bad geometry", what);
  GC_REQUIRE((g->KW * g->Cin) % 32 == 0, "%s: KW*Cin=%d must be a multiple of 32", what, g->KW * g->Cin);
  GC_REQUIRE(g->Cout % 16 == 0, "%s: Cout=%d must be a multiple of 16", what, g->Cout);
  GC_REQUIRE((g->OH - 1) * g->S + g->KH <= g->H && (g->OW - 1) * g->S + g->KW <= g->W, "%s: output does not fit input", what);
  GC_REQUIRE(g->Wp >= g->W && g->Hp >= g->H && g->OWp >= g->OW && g->OHp >= g->OH, "%s: pitches smaller than extents", what);
  GC_REQUIRE(g->in_batch_stride >= (long)g->Hp * g->Wp * g->Cin && g->out_batch_stride >= (long)g->OHp * g->OWp * g->Cout,
             "%s: batch strides too small", what);
  GC_REQUIRE(g->in_batch_stride % 4 == 0 && g->out_batch_stride % 4 == 0, "%s: batch strides must be multiples of 4 floats", what);
  return 0;
}

// A-operand window map over x[B][Hp][Wp][Cin]: dims (kx*c, ox, ky, oy, b) - overlapping strides
MapSpec window_spec(const gc_conv_geom* g, const float* x, const PixBox& bx, int swz = 1) {
  const uint64_t dim[5] = {(uint64_t)g->KW * g->Cin, (uint64_t)g->OW, (uint64_t)g->KH, (uint64_t)g->OH, (uint64_t)g->B};
  const uint64_t str[5] = {1, (uint64_t)g->S * g->Cin, (uint64_t)g->Wp * g->Cin, (uint64_t)g->S * g->Wp * g->Cin,
                           (uint64_t)g->in_batch_stride};
  const uint32_t box[5] = {32, (uint32_t)bx.ox, 1, (uint32_t)bx.oy, (uint32_t)bx.b};
  return spec(x, 5, dim, str, box, 1, swz);
}

// map over y[B][OHp][OWp][Cout]: dims (n, ox, oy, b)
MapSpec out_spec(const gc_conv_geom* g, const float* y, const PixBox& bx, int inner, int tf32, int swz) {
  const uint64_t dim[4] = {(uint64_t)g->Cout, (uint64_t)g->OW, (uint64_t)g->OH, (uint64_t)g->B};
  const uint64_t str[4] = {1, (uint64_t)g->Cout, (uint64_t)g->OWp * g->Cout, (uint64_t)g->out_batch_stride};
  const uint32_t box[4] = {(uint32_t)inner, (uint32_t)bx.ox, (uint32_t)bx.oy, (uint32_t)bx.b};
  return spec(y, 4, dim, str, box, tf32, swz);
}

}  // namespace

// ============================================ C ABI ============================================
extern "C" {

// y[b,oy,ox,n] = epi( sum_{ky,kx,c} x[b, S*oy+ky, S*ox+kx, c] * w[n][ky][kx][c] )
int gc_conv_fprop(const gc_conv_geom* g, const float* x, const float* w, const float* bias, const float* mask_src,
                  unsigned* mask_bits, float* y, int epilogue, float slope, void* stream) {
  if (int e = check_geom(g, "gc_conv_fprop")) return e;
  GC_REQUIRE(x && w && y, "gc_conv_fprop: null pointer");
  GC_REQUIRE(epilogue >= 0 && epilogue <= 3, "gc_conv_fprop: bad epilogue %d", epilogue);
  if (epilogue == EPI_BIAS_LRELU || epilogue == EPI_BIAS) GC_REQUIRE(bias, "gc_conv_fprop: bias epilogue without bias");
  if (epilogue == EPI_MASK) GC_REQUIRE(mask_src || mask_bits, "gc_conv_fprop: mask epilogue without mask source");
  // bit mask of the output tensor y: written by the bias+LeakyReLU epilogue, read by the mask epilogue
  auto set_bits = [&](GemmParams& p, const PixBox& bx) {
    if (!mask_bits) return;
    if (epilogue == EPI_BIAS_LRELU) p.bits_out = mask_bits;
    else if (epilogue == EPI_MASK) p.bits_in = mask_bits;
    p.row_box[0] = bx.ox; p.row_box[1] = bx.oy; p.row_box[2] = bx.b;
    p.row_ext[0][0] = g->OW; p.row_ext[0][1] = g->OH; p.row_ext[0][2] = g->B;
    p.bit_str[0] = g->Cout; p.bit_str[1] = (long)g->OWp * g->Cout; p.bit_str[2] = g->out_batch_stride;
    p.bit_base[0] = 0;
  };
  if (mask_bits) GC_REQUIRE(g->Cout % 32 == 0 && g->out_batch_stride % 32 == 0, "gc_conv_fprop: bit masks need Cout and batch stride multiples of 32");
  // Small-channel stride-2 layers (conv2): "patch mode".  The weight matrix (<= 128 KB) is loaded once per CTA and stays in
  // shared memory; for each of the 4 input parity classes (py,px) ONE (8+1)x(16+1)-pixel patch of the class sub-image is
  // TMA-loaded and serves the class's 4 taps (a,b) as shifted A views (descriptor start (a*9+b)*128 B, 8-row groups
  // 9 rows = 1152 B apart - see profiles/r01_umma_descriptor_experiment.txt).  L2->SM traffic per 128-pixel tile drops
  // from 16 x 16 KB (A) + 128 KB (B) to 4 x 19 KB.
  if (g->S == 2 && g->KH == 4 && g->KW == 4 && g->Cin % 32 == 0 && (long)g->Cout * g->Cin * 64 <= 131072 && g->Wp % 2 == 0 &&
      g->Hp % 2 == 0 && g->Cout <= 256 && getenv("GC_NO_PATCH") == nullptr) {
    Plan pl;
    GemmParams& p = pl.p;
    const int C = g->Cin, nchunk = C / 32;
    const PixBox bx{8, 16, 1};
    p.bn = g->Cout;
    p.bk = 32;
    p.e0 = cdiv(g->OW, 8); p.e1 = cdiv(g->OH, 16);
    p.g0 = nchunk; p.g1 = 2;                      // k -> (chunk, px, py)
    p.k_iters = 4 * nchunk;
    {
      const uint64_t dim[5] = {(uint64_t)2 * C, (uint64_t)g->Wp / 2, 2, (uint64_t)g->Hp / 2, (uint64_t)g->B};
      const uint64_t str[5] = {1, (uint64_t)2 * C, (uint64_t)g->Wp * C, (uint64_t)2 * g->Wp * C, (uint64_t)g->in_batch_stride};
      const uint32_t box[5] = {32, 9, 1, 17, 1};
      if (int e = make_map(spec(x, 5, dim, str, box, 1, 1), &p.mapA)) return e;
    }
    p.a.mul[0][K0] = 32; p.a.mul[0][K1] = C; p.a.mul[2][K2] = 1; p.a.mul[1][M0] = 8; p.a.mul[3][M1] = 16; p.a.mul[4][M2] = 1;
    p.a_panels = 1; p.a_panel_bytes = 9 * 17 * 128; p.a_bytes = 20480;
    p.exp_a_sbo = 9 * 128;
    p.taps = 4;
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) p.tap_off[a * 2 + b] = (a * 9 + b) * 128;
    for (int k = 0; k < p.k_iters; ++k) {
      const int chunk = k % nchunk, px = (k / nchunk) % 2, py = k / (2 * nchunk);
      for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b)
        p.b_tab[k * 4 + a * 2 + b] = (unsigned char)((((2 * a + py) * 4 + (2 * b + px)) * nchunk) + chunk);
    }
    {
      const int K = 16 * C;
      const uint64_t dim[2] = {(uint64_t)K, (uint64_t)g->Cout}, str[2] = {1, (uint64_t)K};
      const uint32_t box[2] = {32, (uint32_t)p.bn};
      if (int e = make_map(spec(w, 2, dim, str, box, 1, 1), &p.mapB)) return e;
    }
    p.b_resident = 1; p.b_slabs = 16 * nchunk;
    const int inner = std::min(32, p.bn);
    if (int e = make_map(out_spec(g, y, bx, inner, 0, p.bn >= 32), &p.mapD[0])) return e;
    p.mapX[0] = p.mapD[0];
    p.d.mul[1][M0] = 8; p.d.mul[2][M1] = 16; p.d.mul[3][M2] = 1; p.d.panel[0] = 32;
    p.d_box_bytes = inner * 4 * 128;
    p.epilogue = epilogue; p.slope = slope; p.bias = bias; p.n_total = g->Cout;
    set_bits(p, bx);
    if (p.bits_in == nullptr && epilogue == EPI_MASK) { if (int e = make_map(out_spec(g, mask_src, bx, inner, 0, p.bn >= 32), &p.mapX[0])) return e; }
    pl.grid = dim3(p.e0 * p.e1 * g->B, 1, 1);
    return finish_and_launch(pl, (cudaStream_t)stream, "gc_conv_fprop(patch)");
  }
  // conv1 in its space-to-depth form (2x2 taps, stride 1, 16 channels): "row-patch mode".  A window row (2 pixels x 16
  // channels) is exactly one 128-byte K slab, so the two vertical taps of a 32x4-pixel tile are two views, 32 rows apart,
  // of ONE 32x5-row patch; the 8 KB weight matrix stays resident in shared memory.  TMA row requests per tile drop from
  // 256 (A) + 64 (B) to 160.
  if (g->S == 1 && g->KH == 2 && g->KW == 2 && g->Cin == 16 && g->Cout <= 256 && g->OW >= 32 && getenv("GC_NO_PATCH") == nullptr) {
    Plan pl;
    GemmParams& p = pl.p;
    // Two 32x4-pixel sub-tiles per tile when Cout == 32 (one 32x9-row patch, two accumulator blocks, two output panels):
    // the per-tile bookkeeping of all three roles - as large as the work itself for a 128x32 tile - is paid half as often.
    const int subs = (g->Cout == 32 && getenv("GC_NO_SUBTILES") == nullptr) ? 2 : 1;
    const PixBox pbx{32, 4, 1};                  // one output panel (= one sub-tile)
    p.bn = subs * g->Cout;
    p.mma_n = g->Cout;
    p.bk = 32;
    p.e0 = cdiv(g->OW, 32); p.e1 = cdiv(g->OH, 4 * subs);
    p.k_iters = 1;
    {
      const uint64_t dim[4] = {32, (uint64_t)g->OW, (uint64_t)g->H, (uint64_t)g->B};
      const uint64_t str[4] = {1, (uint64_t)g->Cin, (uint64_t)g->Wp * g->Cin, (uint64_t)g->in_batch_stride};
      const uint32_t box[4] = {32, 32, (uint32_t)(4 * subs + 1), 1};
      if (int e = make_map(spec(x, 4, dim, str, box, 1, 1), &p.mapA)) return e;
    }
    p.a.mul[1][M0] = 32; p.a.mul[2][M1] = 4 * subs; p.a.mul[3][M2] = 1;
    p.a_panels = 1; p.a_panel_bytes = 32 * (4 * subs + 1) * 128; p.a_bytes = (p.a_panel_bytes + 1023) & ~1023;
    p.taps = 2 * subs;
    for (int sub = 0; sub < subs; ++sub) for (int ky = 0; ky < 2; ++ky) {
      const int t = sub * 2 + ky;
      p.tap_off[t] = (sub * 4 + ky) * 32 * 128;   // rows (4*sub + ky .. +3) of the patch
      p.b_tab[t] = (unsigned char)ky;
      p.tap_acc[t] = sub * g->Cout;
      if (ky == 0) p.tap_fresh |= 1 << t;
    }
    {
      const uint64_t dim[2] = {64, (uint64_t)g->Cout}, str[2] = {1, 64};
      const uint32_t box[2] = {32, (uint32_t)g->Cout};
      if (int e = make_map(spec(w, 2, dim, str, box, 1, 1), &p.mapB)) return e;
    }
    p.b_resident = 1; p.b_slabs = 2;
    const int inner = std::min(32, g->Cout);
    if (int e = make_map(out_spec(g, y, pbx, inner, 0, g->Cout >= 32), &p.mapD[0])) return e;
    if (epilogue == EPI_MASK && !mask_bits) { if (int e = make_map(out_spec(g, mask_src, pbx, inner, 0, g->Cout >= 32), &p.mapX[0])) return e; }
    else p.mapX[0] = p.mapD[0];
    p.d.mul[1][M0] = 32; p.d.mul[2][M1] = 4 * subs; p.d.mul[3][M2] = 1; p.d.panel[0] = 32;
    if (subs > 1) {   // panel q = sub-tile q / sub_panels, channel panel q % sub_panels
      p.sub_panels = (g->Cout + 31) / 32;
      p.d.period = p.sub_panels; p.d.panel2[2] = 4;
    }
    p.d_box_bytes = inner * 4 * 128;
    p.epilogue = epilogue; p.slope = slope; p.bias = bias; p.n_total = g->Cout;
    set_bits(p, pbx);
    pl.grid = dim3(p.e0 * p.e1 * g->B, 1, 1);
    return finish_and_launch(pl, (cudaStream_t)stream, "gc_conv_fprop(row-patch)");
  }
  Plan pl;
  GemmParams& p = pl.p;
  const PixBox bx = choose_box(g->OW, g->OH, g->B, 128, 1, 0.0);
  const int K = g->KH * g->KW * g->Cin;
  p.bn = std::min(256, g->Cout);
  p.bk = 32;
  p.e0 = cdiv(g->OW, bx.ox);
  p.e1 = cdiv(g->OH, bx.oy);
  int e2 = cdiv(g->B, bx.b);
  p.f0 = cdiv(g->Cout, p.bn);
  p.g0 = g->KW * g->Cin / 32;
  p.g1 = g->KH;
  p.k_iters = p.g0 * p.g1;
  // A: window map
  if (int e = make_map(window_spec(g, x, bx), &p.mapA)) return e;
  p.a.mul[0][K0] = 32; p.a.mul[1][M0] = bx.ox; p.a.mul[2][K1] = 1; p.a.mul[3][M1] = bx.oy; p.a.mul[4][M2] = bx.b;
  p.a_panels = 1; p.a_panel_bytes = bx.ox * bx.oy * bx.b * 128;
  pl.pair = pair_ok((long)p.e0 * p.e1 * e2, p.bn, epilogue, mask_bits == nullptr);
  if (pl.pair && ((long)p.e0 * p.e1 * e2) % 2) e2 += 1;   // phantom batch tile
  const int b_rows = pl.pair ? p.bn / 2 : p.bn;   // CTA-pair mode: each CTA stages half of the weight tile
  // B: weights [Cout][K]
  {
    const uint64_t dim[2] = {(uint64_t)K, (uint64_t)g->Cout}, str[2] = {1, (uint64_t)K};
    const uint32_t box[2] = {32, (uint32_t)b_rows};
    if (int e = make_map(spec(w, 2, dim, str, box, 1, 1), &p.mapB)) return e;
  }
  p.b.mul[0][K0] = 32; p.b.mul[0][K1] = 32 * p.g0; p.b.mul[1][N0] = p.bn;
  p.b_panels = 1; p.b_panel_bytes = b_rows * 128;
  p.pair_b_dim = 1; p.pair_b_off = b_rows;
  // D (+ mask source with the same geometry)
  const int inner = std::min(32, p.bn);
  if (int e = make_map(out_spec(g, y, bx, inner, 0, p.bn >= 32), &p.mapD[0])) return e;
  if (epilogue == EPI_MASK && !mask_bits) { if (int e = make_map(out_spec(g, mask_src, bx, inner, 0, p.bn >= 32), &p.mapX[0])) return e; }
  else p.mapX[0] = p.mapD[0];
  p.d.mul[0][N0] = p.bn; p.d.mul[1][M0] = bx.ox; p.d.mul[2][M1] = bx.oy; p.d.mul[3][M2] = bx.b; p.d.panel[0] = 32;
  p.d_box_bytes = inner * 4 * bx.ox * bx.oy * bx.b;
  p.epilogue = epilogue; p.slope = slope; p.bias = bias; p.n_total = g->Cout;
  set_bits(p, bx);
  pl.grid = dim3(p.e0 * p.e1 * e2, p.f0, 1);
  return finish_and_launch(pl, (cudaStream_t)stream, "gc_conv_fprop");
}

// dx[b, S*j+py, S*i+px, c] = mask * sum_{a,b',n} dy[b, j-a, i-b', n] * wd[cls][c][a][b'][n],  cls = py*S+px,
// taps a < KH/S, b' < KW/S.  All S*S parity classes share the same A rows (dy at (j-a, i-b')), so they are ONE GEMM with
// N = S*S*Cin output columns: dy is fetched once instead of S*S times and the MMA runs at N >= 128 instead of Cin.  Each
// 32-column output panel belongs to one class and is stored through that class's strided tensor map.
// mask = LeakyReLU'(sign of mask_src at the output position).
int gc_conv_dgrad(const gc_conv_geom* g, const float* dy, const float* wd, const float* mask_src, const unsigned* mask_bits,
                  float* dx, float slope, float* dbias_in, int dbias_samples, void* stream) {
  if (int e = check_geom(g, "gc_conv_dgrad")) return e;
  if (dbias_in) GC_REQUIRE(mask_bits && g->Cin >= 32 && (g->Cin & (g->Cin - 1)) == 0 && g->Cin <= 256,
                           "gc_conv_dgrad: dbias_in needs the bit-mask epilogue and a power-of-two Cin in [32,256]");
  GC_REQUIRE(dy && wd && dx, "gc_conv_dgrad: null pointer");
  GC_REQUIRE(g->KH % g->S == 0 && g->KW % g->S == 0, "gc_conv_dgrad: taps must be a multiple of the stride");
  GC_REQUIRE(g->Cout % 32 == 0 && g->Cin % 16 == 0 && g->Cin <= 256, "gc_conv_dgrad: Cout%%32, Cin%%16, Cin<=256 required");
  GC_REQUIRE(g->S == 1 || (g->S == 2 && g->Cin % 32 == 0), "gc_conv_dgrad: stride 1, or stride 2 with Cin%%32 == 0");
  const int TA = g->KH / g->S, TB = g->KW / g->S;
  const int Kd = TA * TB * g->Cout;
  const int ncls = g->S * g->S;
  const int Ntot = ncls * g->Cin;
  const int NJ = cdiv(g->H, g->S), NI = cdiv(g->W, g->S);  // class (0,0) has the largest pixel grid
  if (mask_bits) GC_REQUIRE(g->Cin % 32 == 0 && g->in_batch_stride % 32 == 0, "gc_conv_dgrad: bit masks need Cin and batch stride multiples of 32");
  const bool want_mask = mask_src != nullptr || mask_bits != nullptr;
  // bit mask of the activation tensor that dx has the geometry of (one strided class view per output map)
  auto set_bits = [&](GemmParams& p, int box_i, int box_j, int box_b) {
    if (!mask_bits) return;
    p.bits_in = mask_bits;
    p.row_box[0] = box_i; p.row_box[1] = box_j; p.row_box[2] = box_b;
    for (int py = 0; py < g->S; ++py) for (int px = 0; px < g->S; ++px) {
      const int cls = py * g->S + px;
      p.row_ext[cls][0] = cdiv(g->W - px, g->S); p.row_ext[cls][1] = cdiv(g->H - py, g->S); p.row_ext[cls][2] = g->B;
      p.bit_base[cls] = ((long)py * g->Wp + px) * g->Cin;
    }
    p.bit_str[0] = (long)g->S * g->Cin; p.bit_str[1] = (long)g->S * g->Wp * g->Cin; p.bit_str[2] = g->in_batch_stride;
    if (dbias_in) {   // bias gradient of the layer below: column sums of the masked dx over the first dbias_samples samples
      p.colsum_out = dbias_in; p.colsum_mask = g->Cin - 1; p.colsum_dim = 2;
      p.colsum_limit = dbias_samples > 0 ? dbias_samples : g->B;
    }
  };
  // Patch mode for the small layers (conv2, and conv1 in its stride-1 space-to-depth form): resident weights (all classes,
  // <= 128 KB) and one (8+1)x(16+1) patch of dy per 32-channel chunk serving the 4 taps (a,b') as shifted A views.
  if ((g->S == 1 || g->S == 2) && TA == 2 && TB == 2 && (long)Ntot * Kd * 4 <= 131072 && Ntot <= 256 && NI >= 8 && NJ >= 16 &&
      getenv("GC_NO_PATCH") == nullptr) {
    Plan pl;
    GemmParams& p = pl.p;
    const int nchunk = g->Cout / 32, S = g->S;
    p.bn = Ntot;
    p.bk = 32;
    p.e0 = cdiv(NI, 8); p.e1 = cdiv(NJ, 16);
    p.g0 = nchunk; p.g1 = 1;
    p.k_iters = nchunk;
    {
      const uint64_t dim[4] = {(uint64_t)g->Cout, (uint64_t)g->OW, (uint64_t)g->OH, (uint64_t)g->B};
      const uint64_t str[4] = {1, (uint64_t)g->Cout, (uint64_t)g->OWp * g->Cout, (uint64_t)g->out_batch_stride};
      const uint32_t box[4] = {32, 9, 17, 1};
      if (int e = make_map(spec(dy, 4, dim, str, box, 1, 1), &p.mapA)) return e;
    }
    p.a.mul[0][K0] = 32; p.a.mul[1][M0] = 8; p.a.off[1] = -1; p.a.mul[2][M1] = 16; p.a.off[2] = -1; p.a.mul[3][M2] = 1;
    p.a_panels = 1; p.a_panel_bytes = 9 * 17 * 128; p.a_bytes = 20480;
    p.exp_a_sbo = 9 * 128;
    p.taps = 4;
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) {
      p.tap_off[a * 2 + b] = ((1 - a) * 9 + (1 - b)) * 128;   // dy[j-a, i-b'] inside the patch whose origin is (j0-1, i0-1)
      for (int k = 0; k < nchunk; ++k) p.b_tab[k * 4 + a * 2 + b] = (unsigned char)((a * 2 + b) * nchunk + k);
    }
    {
      const uint64_t dim[2] = {(uint64_t)Kd, (uint64_t)Ntot}, str[2] = {1, (uint64_t)Kd};
      const uint32_t box[2] = {32, (uint32_t)p.bn};
      if (int e = make_map(spec(wd, 2, dim, str, box, 1, 1), &p.mapB)) return e;
    }
    p.b_resident = 1; p.b_slabs = 4 * nchunk;
    const int inner = std::min(32, g->Cin);
    for (int py = 0; py < S; ++py) for (int px = 0; px < S; ++px) {
      const int cls = py * S + px;
      const long base = ((long)py * g->Wp + px) * g->Cin;
      const uint64_t dim[4] = {(uint64_t)g->Cin, (uint64_t)cdiv(g->W - px, S), (uint64_t)cdiv(g->H - py, S), (uint64_t)g->B};
      const uint64_t str[4] = {1, (uint64_t)S * g->Cin, (uint64_t)S * g->Wp * g->Cin, (uint64_t)g->in_batch_stride};
      const uint32_t box[4] = {(uint32_t)inner, 8, 16, 1};
      if (int e = make_map(spec(dx + base, 4, dim, str, box, 0, g->Cin >= 32), &p.mapD[cls])) return e;
      if (mask_src && !mask_bits) { if (int e = make_map(spec(mask_src + base, 4, dim, str, box, 0, g->Cin >= 32), &p.mapX[cls])) return e; }
      else p.mapX[cls] = p.mapD[cls];
    }
    p.cols_per_map = ncls > 1 ? g->Cin : 0;
    p.d.mul[1][M0] = 8; p.d.mul[2][M1] = 16; p.d.mul[3][M2] = 1; p.d.panel[0] = 32;
    p.d_box_bytes = inner * 4 * 128;
    p.epilogue = want_mask ? EPI_MASK : EPI_STORE; p.slope = slope; p.n_total = Ntot;
    set_bits(p, 8, 16, 1);
    pl.grid = dim3(p.e0 * p.e1 * g->B, 1, 1);
    return finish_and_launch(pl, (cudaStream_t)stream, "gc_conv_dgrad(patch)");
  }
  Plan pl;
  GemmParams& p = pl.p;
  const PixBox bx = choose_box(NI, NJ, g->B, 128, 1, 0.0);
  p.bn = std::min(256, Ntot);
  GC_REQUIRE(Ntot % p.bn == 0, "gc_conv_dgrad: S*S*Cin=%d not a multiple of the tile N %d", Ntot, p.bn);
  p.bk = 32;
  p.e0 = cdiv(NI, bx.ox);
  p.e1 = cdiv(NJ, bx.oy);
  int e2 = cdiv(g->B, bx.b);
  p.f0 = Ntot / p.bn;
  p.g0 = g->Cout / 32;
  p.g1 = TB;
  p.k_iters = p.g0 * TB * TA;
  // A: dy[B][OHp][OWp][Cout] read at (j - a, i - b'); out-of-range taps are zero-filled by TMA
  if (int e = make_map(out_spec(g, dy, bx, 32, 1, 1), &p.mapA)) return e;
  p.a.mul[0][K0] = 32; p.a.mul[1][M0] = bx.ox; p.a.mul[1][K1] = -1; p.a.mul[2][M1] = bx.oy; p.a.mul[2][K2] = -1;
  p.a.mul[3][M2] = bx.b;
  p.a_panels = 1; p.a_panel_bytes = bx.ox * bx.oy * bx.b * 128;
  pl.pair = pair_ok((long)p.e0 * p.e1 * e2, p.bn, want_mask ? EPI_MASK : EPI_STORE, mask_bits == nullptr);
  if (pl.pair && ((long)p.e0 * p.e1 * e2) % 2) e2 += 1;   // phantom batch tile
  const int b_rows = pl.pair ? p.bn / 2 : p.bn;   // CTA-pair mode: each CTA stages half of the weight tile
  // B: wd viewed as [ncls*Cin][Kd], k = (a*TB + b')*Cout + n
  {
    const uint64_t dim[2] = {(uint64_t)Kd, (uint64_t)Ntot};
    const uint64_t str[2] = {1, (uint64_t)Kd};
    const uint32_t box[2] = {32, (uint32_t)b_rows};
    if (int e = make_map(spec(wd, 2, dim, str, box, 1, 1), &p.mapB)) return e;
  }
  p.b.mul[0][K0] = 32; p.b.mul[0][K1] = g->Cout; p.b.mul[0][K2] = TB * g->Cout; p.b.mul[1][N0] = p.bn;
  p.b_panels = 1; p.b_panel_bytes = b_rows * 128;
  p.pair_b_dim = 1; p.pair_b_off = b_rows;
  // D / mask: one strided map per parity class over dx[B][Hp][Wp][Cin]
  const int inner = std::min(32, g->Cin);
  for (int py = 0; py < g->S; ++py) {
    for (int px = 0; px < g->S; ++px) {
      const int cls = py * g->S + px;
      const long base = ((long)py * g->Wp + px) * g->Cin;
      const uint64_t dim[4] = {(uint64_t)g->Cin, (uint64_t)cdiv(g->W - px, g->S), (uint64_t)cdiv(g->H - py, g->S), (uint64_t)g->B};
      const uint64_t str[4] = {1, (uint64_t)g->S * g->Cin, (uint64_t)g->S * g->Wp * g->Cin, (uint64_t)g->in_batch_stride};
      const uint32_t box[4] = {(uint32_t)inner, (uint32_t)bx.ox, (uint32_t)bx.oy, (uint32_t)bx.b};
      if (int e = make_map(spec(dx + base, 4, dim, str, box, 0, g->Cin >= 32), &p.mapD[cls])) return e;
      if (mask_src && !mask_bits) { if (int e = make_map(spec(mask_src + base, 4, dim, str, box, 0, g->Cin >= 32), &p.mapX[cls])) return e; }
      else p.mapX[cls] = p.mapD[cls];
    }
  }
  p.cols_per_map = ncls > 1 ? g->Cin : 0;
  p.d.mul[1][M0] = bx.ox; p.d.mul[2][M1] = bx.oy; p.d.mul[3][M2] = bx.b; p.d.panel[0] = 32;
  p.d_box_bytes = inner * 4 * bx.ox * bx.oy * bx.b;
  p.epilogue = want_mask ? EPI_MASK : EPI_STORE; p.slope = slope; p.n_total = Ntot;
  set_bits(p, bx.ox, bx.oy, bx.b);
  pl.grid = dim3(p.e0 * p.e1 * e2, p.f0, 1);
  return finish_and_launch(pl, (cudaStream_t)stream, "gc_conv_dgrad");
}

// K rows (pixels) per k-iteration for wgrad: a multiple of 8 that leaves room for >= 3 pipeline stages
static int wgrad_max_rows(int bn) {
  const int per_row = 128 * (4 + bn / 32);
  int r = ((kSmemBudget - 2 * 16384 - 2048) / 3) / per_row;
  r = (r / 8) * 8;
  return std::max(8, std::min(64, r));
}

// wgrad tile N: whole (kx,c) rows of as many ky taps as fit in 256 columns (so dy is re-read as rarely as possible)
static void wgrad_tile_n(const gc_conv_geom* g, int& bn, int& ky_per, int& n_tiles) {
  const int KC = g->KW * g->Cin;
  if (KC >= 256) { bn = 256; ky_per = 1; n_tiles = (KC / 256) * g->KH; }
  else {
    ky_per = std::min(g->KH, 256 / KC);
    while (g->KH % ky_per) --ky_per;
    bn = KC * ky_per;
    n_tiles = g->KH / ky_per;
  }
}

// Slab-mode wgrad for the stride-2 4x4 convolutions.  With both dY and the four input parity sub-images flattened to
// rows of pitch P (> OW, zero-filled by TMA beyond the extents), tap (ky,kx) = (2a+py, 2b+px) is a 1-D correlation:
//   dW[n][ky][kx][c] = sum_r dY[r][n] * Xsub_{py,px}[r + a*P + b][c]
// so ONE patch per parity class serves its 4 taps as MN-major B views shifted by a*P+b rows (the b=0/b=1 panels are the
// same rows one apart: LBO = 128 B), and dY is fetched once for all 16 taps of a 32-channel chunk.  The wrap-around
// products of the shift hit the zero columns of dY.  (MN-major shifted views: profiles/r01_umma_descriptor_experiment.txt.)
struct SlabGeom { int P, by, R, rows_patch; bool ok; };
static SlabGeom slab_geom(const gc_conv_geom* g) {
  SlabGeom s{0, 0, 0, 0, false};
  if (!(g->S == 2 && g->KH == 4 && g->KW == 4 && g->Cin % 32 == 0 && g->Cout % 32 == 0 && g->Wp % 2 == 0 && g->Hp % 2 == 0)) return s;
  if (getenv("GC_NO_SLAB") != nullptr) return s;
  if (g->OW * g->OH < 400) return s;   // tiny images (conv4: 10x10) lose more to row padding than the patches save
  s.P = ((g->OW + 1 + 3) / 4) * 4;
  if (s.P > 256 || s.P > g->Wp / 2 + 8) return s;
  for (int by = 1; by <= 8; ++by) {
    if ((s.P * by) % 8) continue;
    const int a_bytes = std::min(4, g->Cout / 32) * s.P * by * 128, b_bytes = 4 * (by + 2) * s.P * 128;
    if (2 * (a_bytes + b_bytes) + 2 * 16384 + 4096 > kSmemBudget) break;   // keep >= 2 pipeline stages
    s.by = by;
    if (s.P * by >= 48) break;
  }
  if (s.by == 0) return s;
  s.R = s.P * s.by; s.rows_patch = (s.by + 2) * s.P; s.ok = true;
  return s;
}

int gc_conv_wgrad_splits(const gc_conv_geom* g) {
  if (check_geom(g, "gc_conv_wgrad_splits")) return -1;
  if (slab_geom(g).ok) {
    const int tiles = cdiv(g->Cout, 128) * (g->Cin / 32);
    return std::max(1, std::min(g->B, (2 * gc::kNumSMs) / std::max(1, tiles)));
  }
  int bn, ky_per, n_tiles;
  wgrad_tile_n(g, bn, ky_per, n_tiles);
  const PixBox bx = choose_box(g->OW, g->OH, g->B, wgrad_max_rows(bn), 8, 4.0);
  const int btiles = cdiv(g->B, bx.b);
  const int tiles = cdiv(g->Cout, 128) * n_tiles;
  int z = std::max(1, std::min(btiles, (2 * gc::kNumSMs) / std::max(1, tiles)));
  return z;
}

// dw_partial[z][n][ky][kx*c] = sum over the z-th slice of samples of dy[pix][n] * x[window(pix)][ky][kx*c]
int gc_conv_wgrad(const gc_conv_geom* g, const float* dy, const float* x, float* dw_partial, int splits, void* stream) {
  if (int e = check_geom(g, "gc_conv_wgrad")) return e;
  GC_REQUIRE(dy && x && dw_partial, "gc_conv_wgrad: null pointer");
  GC_REQUIRE(g->Cout % 32 == 0, "gc_conv_wgrad: Cout %% 32 required");
  GC_REQUIRE(splits >= 1, "gc_conv_wgrad: splits=%d", splits);
  const SlabGeom sg = slab_geom(g);
  if (sg.ok) {
    Plan pl;
    GemmParams& p = pl.p;
    pl.a_mn = pl.b_mn = true;
    const int C = g->Cin, KC = 4 * C;
    p.bn = 512; p.mma_n = 64; p.acc_stages = 1; p.ngroups = 8;
    if (g->Cout <= 64 && getenv("GC_NO_M64") == nullptr) p.mma_m = 64;
    p.bk = sg.R;
    p.e0 = cdiv(g->Cout, 128); p.e1 = 1;
    p.f0 = C / 32;                              // n-tile = 32-channel chunk
    p.g0 = cdiv(g->OH, sg.by);                  // k -> (row tile, sample within the split)
    const int per_split = cdiv(g->B, splits);
    p.g1 = per_split;
    p.k_iters = p.g0 * per_split;
    // A: dY rows (ox padded to P, zero-filled), panels of 32 output channels
    {
      const uint64_t dim[4] = {(uint64_t)g->Cout, (uint64_t)g->OW, (uint64_t)g->OH, (uint64_t)g->B};
      const uint64_t str[4] = {1, (uint64_t)g->Cout, (uint64_t)g->OWp * g->Cout, (uint64_t)g->out_batch_stride};
      const uint32_t box[4] = {32, (uint32_t)sg.P, (uint32_t)sg.by, 1};
      if (int e = make_map(spec(dy, 4, dim, str, box, 1, 2), &p.mapA)) return e;
    }
    p.a.mul[0][M0] = 128; p.a.mul[2][K0] = sg.by; p.a.mul[3][K1] = 1; p.a.mul[3][Z] = per_split; p.a.panel[0] = 32;
    p.a_panels = std::min(4, g->Cout / 32); p.a_panel_bytes = sg.R * 128; p.a_bytes = p.a_panels * sg.R * 128;
    p.a_bytes = (p.a_bytes + 1023) & ~1023;
    // B: one patch per parity class (py,px) of the input, (by+2) rows of pitch P
    {
      const uint64_t dim[5] = {(uint64_t)2 * C, (uint64_t)g->Wp / 2, 2, (uint64_t)g->Hp / 2, (uint64_t)g->B};
      const uint64_t str[5] = {1, (uint64_t)2 * C, (uint64_t)g->Wp * C, (uint64_t)2 * g->Wp * C, (uint64_t)g->in_batch_stride};
      const uint32_t box[5] = {32, (uint32_t)sg.P, 1, (uint32_t)sg.by + 2, 1};
      if (int e = make_map(spec(x, 5, dim, str, box, 1, 2), &p.mapB)) return e;
    }
    p.b.mul[0][N0] = 32; p.b.mul[3][K0] = sg.by; p.b.mul[4][K1] = 1; p.b.mul[4][Z] = per_split;
    p.b.panel[0] = C; p.b.panel2[2] = 1; p.b.period = 2;       // class cls = py*2+px: px -> +C in dim 0, py -> +1 in dim 2
    p.b_panels = 4; p.b_panel_bytes = sg.rows_patch * 128;
    pl.b_bytes_override = 4 * sg.rows_patch * 128;
    p.exp_b_lbo = 128;                                          // panels b=0 / b=1 of a group: same rows, one apart
    for (int cls = 0; cls < 4; ++cls) for (int a = 0; a < 2; ++a) {
      const int gi = cls * 2 + a;
      p.grp_b_off[gi] = cls * p.b_panel_bytes + a * sg.P * 128;
      p.grp_acc[gi] = gi * 64;
      for (int b = 0; b < 2; ++b) {
        const int py = cls / 2, px = cls % 2;
        p.panel_tab0[gi * 2 + b] = (2 * b + px) * C;            // kx * C (+ 32*chunk from the tile)
        p.panel_tab1[gi * 2 + b] = 2 * a + py;                  // ky
      }
    }
    // D: partial[z][Cout][KH][KW*C]
    {
      const uint64_t dim[4] = {(uint64_t)KC, (uint64_t)g->KH, (uint64_t)g->Cout, (uint64_t)splits};
      const uint64_t str[4] = {1, (uint64_t)KC, (uint64_t)g->KH * KC, (uint64_t)g->Cout * g->KH * KC};
      const int rows = std::min(128, g->Cout);
      const uint32_t box[4] = {32, 1, (uint32_t)rows, 1};
      if (int e = make_map(spec(dw_partial, 4, dim, str, box, 0, 1), &p.mapD[0])) return e;
      p.d_box_bytes = rows * 128;
    }
    p.mapX[0] = p.mapD[0];
    p.d.mul[0][N0] = 32; p.d.mul[2][M0] = 128; p.d.mul[3][Z] = 1;
    p.epilogue = EPI_STORE; p.n_total = 16 * C;
    pl.grid = dim3(p.e0, p.f0, splits);
    return finish_and_launch(pl, (cudaStream_t)stream, "gc_conv_wgrad(slab)");
  }
  Plan pl;
  GemmParams& p = pl.p;
  pl.a_mn = pl.b_mn = true;
  const int KC = g->KW * g->Cin;
  int ky_per, n_tiles;
  wgrad_tile_n(g, p.bn, ky_per, n_tiles);
  GC_REQUIRE(KC % 32 == 0 && (KC >= 256 ? KC % 256 == 0 : true), "gc_conv_wgrad: unsupported KW*Cin=%d", KC);
  if (g->Cout <= 64 && getenv("GC_NO_M64") == nullptr) p.mma_m = 64;
  const PixBox bx = choose_box(g->OW, g->OH, g->B, wgrad_max_rows(p.bn), 8, 4.0);
  p.bk = bx.ox * bx.oy * bx.b;
  p.e0 = cdiv(g->Cout, 128);
  p.e1 = 1;
  p.f0 = KC >= 256 ? KC / 256 : 1;   // n-tile -> (n0 = 256-column block inside a ky row, n1 = ky group)
  p.g0 = cdiv(g->OW, bx.ox);
  p.g1 = cdiv(g->OH, bx.oy);
  const int btiles = cdiv(g->B, bx.b);
  const int g2 = cdiv(btiles, splits);
  p.k_iters = p.g0 * p.g1 * g2;
  const int per_ky = std::min(KC, 256) / 32;  // 32-column panels per ky row inside one tile
  // A: dy MN-major, panels of 32 output channels
  if (int e = make_map(out_spec(g, dy, bx, 32, 1, 2), &p.mapA)) return e;
  p.a.mul[0][M0] = 128; p.a.mul[1][K0] = bx.ox; p.a.mul[2][K1] = bx.oy; p.a.mul[3][K2] = bx.b; p.a.mul[3][Z] = g2 * bx.b;
  p.a.panel[0] = 32;
  p.a_panels = std::min(4, g->Cout / 32); p.a_panel_bytes = p.bk * 128;
  // B: x windows MN-major, panels of 32 (kx,c) columns; panels walk (kx,c) first, then ky
  if (int e = make_map(window_spec(g, x, bx, 2), &p.mapB)) return e;
  p.b.mul[0][N0] = 256; p.b.mul[1][K0] = bx.ox; p.b.mul[2][N1] = ky_per; p.b.mul[3][K1] = bx.oy; p.b.mul[4][K2] = bx.b;
  p.b.mul[4][Z] = g2 * bx.b; p.b.panel[0] = 32; p.b.panel2[2] = 1; p.b.period = per_ky;
  p.b_panels = p.bn / 32; p.b_panel_bytes = p.bk * 128;
  // D: partial[z][Cout][KH][KC]
  {
    const uint64_t dim[4] = {(uint64_t)KC, (uint64_t)g->KH, (uint64_t)g->Cout, (uint64_t)splits};
    const uint64_t str[4] = {1, (uint64_t)KC, (uint64_t)g->KH * KC, (uint64_t)g->Cout * g->KH * KC};
    const int rows = std::min(128, g->Cout);
    const uint32_t box[4] = {32, 1, (uint32_t)rows, 1};
    if (int e = make_map(spec(dw_partial, 4, dim, str, box, 0, 1), &p.mapD[0])) return e;
    p.d_box_bytes = rows * 128;
  }
  p.mapX[0] = p.mapD[0];
  p.d.mul[0][N0] = 256; p.d.mul[1][N1] = ky_per; p.d.mul[2][M0] = 128; p.d.mul[3][Z] = 1;
  p.d.panel[0] = 32; p.d.panel2[1] = 1; p.d.period = per_ky;
  p.epilogue = EPI_STORE; p.n_total = KC * g->KH;
  pl.grid = dim3(p.e0, n_tiles, splits);
  return finish_and_launch(pl, (cudaStream_t)stream, "gc_conv_wgrad");
}

// y[z][m][n] = epi( sum_{k in split z} x[m][k] * w[n][k] ); x row pitch ldx, w row pitch ldw (floats, multiples of 4)
int gc_linear_fwd(const float* x, long ldx, const float* w, long ldw, const float* bias, float* y, long ldy, int M, int N,
                  int K, int epilogue, float slope, int splits, void* stream) {
  GC_REQUIRE(x && w && y && M > 0 && N > 0 && K > 0, "gc_linear_fwd: bad arguments");
  GC_REQUIRE(ldx % 4 == 0 && ldw % 4 == 0 && ldy % 4 == 0, "gc_linear_fwd: row pitches must be multiples of 4 floats");
  GC_REQUIRE(epilogue == EPI_STORE || epilogue == EPI_BIAS_LRELU || epilogue == EPI_BIAS, "gc_linear_fwd: bad epilogue");
  GC_REQUIRE(splits >= 1 && (splits == 1 || epilogue == EPI_STORE), "gc_linear_fwd: split-K needs the plain-store epilogue");
  if (epilogue != EPI_STORE) GC_REQUIRE(bias, "gc_linear_fwd: bias epilogue without bias");
  Plan pl;
  GemmParams& p = pl.p;
  p.bn = std::min(256, ((N + 15) / 16) * 16);
  p.bk = 32;
  p.e0 = cdiv(M, 128);
  p.f0 = cdiv(N, p.bn);
  const int kit_total = cdiv(K, 32);
  p.k_iters = cdiv(kit_total, splits);
  p.kz_stride = p.k_iters;
  p.g0 = 1 << 30;
  {
    const uint64_t dim[2] = {(uint64_t)K, (uint64_t)M}, str[2] = {1, (uint64_t)ldx};
    const uint32_t box[2] = {32, 128};
    if (int e = make_map(spec(x, 2, dim, str, box, 1, 1), &p.mapA)) return e;
  }
  p.a.mul[0][K0] = 32; p.a.mul[1][M0] = 128; p.a_panels = 1; p.a_panel_bytes = 16384;
  if (const char* ex = getenv("GC_EXP")) {
    // Bring-up experiment (tests/gpu_probe_desc.py): can a K-major SWIZZLE_128B A descriptor start at a row that is
    // not a multiple of 8 (mode 1), and can its 8-row groups be 9 rows apart (mode 2: SBO = 1152 B)?
    int mode = 0, r = 0, bo = 0;
    if (sscanf(ex, "%d,%d,%d", &mode, &r, &bo) == 3) {
      if (mode == 1) {
        p.a_panels = 2; p.a.panel[1] = 128; p.a_bytes = 32768; p.exp_a_off = r * 128; p.exp_a_baseoff = bo;
      } else if (mode == 2) {
        const uint64_t dim[2] = {(uint64_t)K, (uint64_t)M}, str[2] = {1, (uint64_t)ldx};
        const uint32_t box[2] = {32, 9};
        if (int e = make_map(spec(x, 2, dim, str, box, 1, 1), &p.mapA)) return e;
        p.a_panels = 16; p.a_panel_bytes = 1152; p.a.panel[1] = 9; p.a_bytes = 20480;
        p.exp_a_off = r * 128; p.exp_a_sbo = 1152; p.exp_a_baseoff = bo;
      }
    }
  }
  pl.pair = getenv("GC_EXP") == nullptr && pair_ok(p.e0, p.bn, epilogue, false);
  if (pl.pair && p.e0 % 2) p.e0 += 1;   // phantom row tile
  const int b_rows = pl.pair ? p.bn / 2 : p.bn;   // CTA-pair mode: each CTA stages half of the weight tile
  {
    const uint64_t dim[2] = {(uint64_t)K, (uint64_t)N}, str[2] = {1, (uint64_t)ldw};
    const uint32_t box[2] = {32, (uint32_t)b_rows};
    if (int e = make_map(spec(w, 2, dim, str, box, 1, 1), &p.mapB)) return e;
  }
  p.b.mul[0][K0] = 32; p.b.mul[1][N0] = p.bn; p.b_panels = 1; p.b_panel_bytes = b_rows * 128;
  p.pair_b_dim = 1; p.pair_b_off = b_rows;
  {
    const int inner = std::min(32, p.bn);
    const uint64_t dim[3] = {(uint64_t)N, (uint64_t)M, (uint64_t)splits}, str[3] = {1, (uint64_t)ldy, (uint64_t)ldy * M};
    const uint32_t box[3] = {(uint32_t)inner, 128, 1};
    if (int e = make_map(spec(y, 3, dim, str, box, 0, p.bn >= 32), &p.mapD[0])) return e;
    p.d_box_bytes = inner * 4 * 128;
  }
  p.mapX[0] = p.mapD[0];
  p.d.mul[0][N0] = p.bn; p.d.mul[1][M0] = 128; p.d.mul[2][Z] = 1; p.d.panel[0] = 32;
  p.epilogue = epilogue; p.slope = slope; p.bias = bias; p.n_total = N;
  pl.grid = dim3(p.e0, p.f0, splits);
  return finish_and_launch(pl, (cudaStream_t)stream, "gc_linear_fwd");
}

// dx[m][n] = mask * sum_k dy[m][k] * w[k][n]   (w is the forward weight [out=K][in=N], read MN-major - no transpose)
int gc_linear_dgrad(const float* dy, long lddy, const float* w, long ldw, const float* mask_src, const unsigned* mask_bits,
                    long ldm, float* dx, long lddx, int M, int N, int K, float slope, float* colsum, int colsum_mod,
                    int colsum_rows, void* stream) {
  GC_REQUIRE(dy && w && dx && M > 0 && N > 0 && K > 0, "gc_linear_dgrad: bad arguments");
  if (colsum) GC_REQUIRE(mask_bits && colsum_mod >= 32 && colsum_mod <= 256 && (colsum_mod & (colsum_mod - 1)) == 0 && N % 32 == 0,
                         "gc_linear_dgrad: colsum needs the bit-mask epilogue, N %% 32 == 0 and a power-of-two modulus in [32,256]");
  GC_REQUIRE(lddy % 4 == 0 && ldw % 4 == 0 && lddx % 4 == 0 && ldm % 4 == 0, "gc_linear_dgrad: pitches must be multiples of 4");
  Plan pl;
  GemmParams& p = pl.p;
  pl.b_mn = true;
  p.bn = std::min(256, ((N + 31) / 32) * 32);
  p.bk = 32;
  p.e0 = cdiv(M, 128);
  p.f0 = cdiv(N, p.bn);
  p.k_iters = cdiv(K, 32);
  p.g0 = 1 << 30;
  {
    const uint64_t dim[2] = {(uint64_t)K, (uint64_t)M}, str[2] = {1, (uint64_t)lddy};
    const uint32_t box[2] = {32, 128};
    if (int e = make_map(spec(dy, 2, dim, str, box, 1, 1), &p.mapA)) return e;
  }
  p.a.mul[0][K0] = 32; p.a.mul[1][M0] = 128; p.a_panels = 1; p.a_panel_bytes = 16384;
  {
    const uint64_t dim[2] = {(uint64_t)N, (uint64_t)K}, str[2] = {1, (uint64_t)ldw};
    const uint32_t box[2] = {32, 32};
    if (int e = make_map(spec(w, 2, dim, str, box, 1, 2), &p.mapB)) return e;
  }
  p.b.mul[0][N0] = p.bn; p.b.mul[1][K0] = 32; p.b.panel[0] = 32; p.b_panels = p.bn / 32; p.b_panel_bytes = 32 * 128;
  {
    const uint64_t dim[2] = {(uint64_t)N, (uint64_t)M}, str[2] = {1, (uint64_t)lddx};
    const uint32_t box[2] = {32, 128};
    if (int e = make_map(spec(dx, 2, dim, str, box, 0, 1), &p.mapD[0])) return e;
    if (mask_src && !mask_bits) {
      const uint64_t strm[2] = {1, (uint64_t)ldm};
      if (int e = make_map(spec(mask_src, 2, dim, strm, box, 0, 1), &p.mapX[0])) return e;
    } else p.mapX[0] = p.mapD[0];
    p.d_box_bytes = 128 * 128;
  }
  p.d.mul[0][N0] = p.bn; p.d.mul[1][M0] = 128; p.d.panel[0] = 32;
  p.epilogue = (mask_src || mask_bits) ? EPI_MASK : EPI_STORE; p.slope = slope; p.n_total = N;
  if (mask_bits) {
    GC_REQUIRE(ldm % 32 == 0, "gc_linear_dgrad: bit masks need a row pitch that is a multiple of 32");
    p.bits_in = mask_bits;
    p.row_box[0] = 128; p.row_box[1] = 1; p.row_box[2] = 1;
    p.row_ext[0][0] = M; p.row_ext[0][1] = 1; p.row_ext[0][2] = 1;
    p.bit_str[0] = ldm; p.bit_str[1] = 0; p.bit_str[2] = 0; p.bit_base[0] = 0;
    if (colsum) {
      p.colsum_out = colsum; p.colsum_mask = colsum_mod - 1; p.colsum_dim = 0;
      p.colsum_limit = colsum_rows > 0 ? colsum_rows : M;
    }
  }
  pl.grid = dim3(p.e0, p.f0, 1);
  return finish_and_launch(pl, (cudaStream_t)stream, "gc_linear_dgrad");
}

// dw[z][m][n] = sum_{rows in split z} dy[row][m] * x[row][n]   (both operands MN-major - no transposes)
int gc_linear_wgrad(const float* dy, long lddy, const float* x, long ldx, float* dw, long lddw, int M, int N, int K, int splits,
                    void* stream) {
  GC_REQUIRE(dy && x && dw && M > 0 && N > 0 && K > 0 && splits >= 1, "gc_linear_wgrad: bad arguments");
  GC_REQUIRE(lddy % 4 == 0 && ldx % 4 == 0 && lddw % 4 == 0, "gc_linear_wgrad: pitches must be multiples of 4");
  Plan pl;
  GemmParams& p = pl.p;
  pl.a_mn = pl.b_mn = true;
  p.bn = std::min(256, ((N + 31) / 32) * 32);
  p.bk = 32;
  p.e0 = cdiv(M, 128);
  p.f0 = cdiv(N, p.bn);
  const int kit_total = cdiv(K, p.bk);
  p.k_iters = cdiv(kit_total, splits);
  p.kz_stride = p.k_iters;
  p.g0 = 1 << 30;
  {
    const uint64_t dim[2] = {(uint64_t)M, (uint64_t)K}, str[2] = {1, (uint64_t)lddy};
    const uint32_t box[2] = {32, (uint32_t)p.bk};
    if (int e = make_map(spec(dy, 2, dim, str, box, 1, 2), &p.mapA)) return e;
  }
  p.a.mul[0][M0] = 128; p.a.mul[1][K0] = p.bk; p.a.panel[0] = 32;
  p.a_panels = std::min(4, cdiv(M, 32)); p.a_panel_bytes = p.bk * 128;
  {
    const uint64_t dim[2] = {(uint64_t)N, (uint64_t)K}, str[2] = {1, (uint64_t)ldx};
    const uint32_t box[2] = {32, (uint32_t)p.bk};
    if (int e = make_map(spec(x, 2, dim, str, box, 1, 2), &p.mapB)) return e;
  }
  p.b.mul[0][N0] = p.bn; p.b.mul[1][K0] = p.bk; p.b.panel[0] = 32; p.b_panels = p.bn / 32; p.b_panel_bytes = p.bk * 128;
  if (const char* ex = getenv("GC_EXP")) {
    // Bring-up experiment (tests/gpu_probe_desc.py): MN-major (SWIZZLE_128B_BASE32B) B views that start r rows into the
    // loaded panel (mode 3), and two 32-column panels that are the SAME rows one row apart (LBO = 128 B, mode 4).
    int mode = 0, r = 0, bo = 0;
    if (sscanf(ex, "%d,%d,%d", &mode, &r, &bo) == 3 && (mode == 3 || mode == 4)) {
      const uint64_t dim[2] = {(uint64_t)N, (uint64_t)K}, str[2] = {1, (uint64_t)ldx};
      const uint32_t box[2] = {32, (uint32_t)p.bk + 8};
      if (int e = make_map(spec(x, 2, dim, str, box, 1, 2), &p.mapB)) return e;
      p.b_panel_bytes = (p.bk + 8) * 128;
      p.exp_b_off = r * 128;
      if (mode == 4) { p.b_panels = 1; p.exp_b_lbo = 128; }
      pl.b_bytes_override = (p.bn / 32) * (p.bk + 8) * 128;
    }
  }
  {
    const uint64_t dim[3] = {(uint64_t)N, (uint64_t)M, (uint64_t)splits}, str[3] = {1, (uint64_t)lddw, (uint64_t)lddw * M};
    const uint32_t box[3] = {32, 128, 1};
    if (int e = make_map(spec(dw, 3, dim, str, box, 0, 1), &p.mapD[0])) return e;
    p.d_box_bytes = 128 * 128;
  }
  p.mapX[0] = p.mapD[0];
  p.d.mul[0][N0] = p.bn; p.d.mul[1][M0] = 128; p.d.mul[2][Z] = 1; p.d.panel[0] = 32;
  p.epilogue = EPI_STORE; p.n_total = N;
  pl.grid = dim3(p.e0, p.f0, splits);
  return finish_and_launch(pl, (cudaStream_t)stream, "gc_linear_wgrad");
}

}  // extern "C"
