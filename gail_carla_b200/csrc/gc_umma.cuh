// tcgen05 / TMEM / TMA "affine-tile" TF32 GEMM engine for sm_100a.
//
// One kernel serves every dense contraction of the policy / discriminator trunks
// (tools/model.py:89-164, algo/wdgail.py:26-32,56-98): conv fprop / dgrad / wgrad as implicit GEMMs and the
// fully-connected fwd / dgrad / wgrad.  D[M,N] = sum_k A[M,k] * B[N,k] with
//   * operands fetched by 5-D TMA boxes whose coordinates are affine functions of the tile index, the
//     k-iteration index and blockIdx.z (so im2col windows, parity-class dgrad taps, split-K and batched
//     pixel boxes are all just coefficient tables filled in on the host - no im2col buffer, no transposes),
//   * either operand K-major (rows of 32 tf32 = 128 B, SWIZZLE_128B) or MN-major (panels of 32 tf32 along
//     M/N x bk rows along K), chosen per operand through the UMMA instruction descriptor,
//   * tcgen05.mma.kind::tf32 (M=128, N=16..256, K=8) issued by one thread, fp32 accumulators in TMEM,
//   * a 4-warp epilogue (tcgen05.ld -> bias / LeakyReLU / LeakyReLU'-mask -> swizzled smem -> TMA store);
//     TMA clips partial tiles on store and zero-fills them on load, so the kernel has no bounds logic.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace gcu {

constexpr int kSrc = 9;  // coordinate sources: m0 m1 m2 | n0 n1 | k0 k1 k2 | z
enum Src { M0 = 0, M1, M2, N0, N1, K0, K1, K2, Z };

struct TmaAddr {
  int off[5];
  int mul[5][kSrc];
  int panel[5];   // added once per panel (successive TMA boxes of one operand / successive 32-column output panels)
  int panel2[5];  // two-level panel walk: panel q sits at (q % period) * panel + (q / period) * panel2
  int period;     // 0 = single level
};

enum Epilogue { EPI_STORE = 0, EPI_BIAS_LRELU = 1, EPI_BIAS = 2, EPI_MASK = 3 };

struct alignas(64) GemmParams {
  CUtensorMap mapA, mapB;
  CUtensorMap mapD[4], mapX[4];  // output / mask-source maps (one per dgrad parity class, else only [0])
  TmaAddr a, b, d;
  int mt, nt, zt;      // tile grid: m-tiles x n-tiles x z (persistent CTAs walk tile = (z*nt + n)*mt + m)
  int e0, e1;          // m-tile index -> (m0, m1, m2) extents
  int f0;              // n-tile index -> (n0, n1)
  int g0, g1;          // k-iteration -> (k0, k1, k2)
  int k_iters;         // k-iterations per CTA
  int kz_stride;       // k-iteration offset per blockIdx.z (flat split-K), else 0
  int bk;              // K elements per k-iteration (32 for K-major operands; multiple of 8)
  int bn;              // tile N (multiple of 16, <= 256)
  int a_panels, b_panels, a_panel_bytes, b_panel_bytes;
  int a_bytes, b_bytes;  // per-stage region sizes
  int stages;
  int tmem_cols;
  int epilogue;
  int n_total;         // valid output columns (bias guard)
  int d_row_bytes;     // 128 (swizzled staging) or bn*4 when bn < 32 (unswizzled)
  int d_box_bytes;     // bytes of one output / mask TMA box (rows may be < 128)
  int nbuf;            // staging buffers: 2, or 4 with EPI_MASK
  int cols_per_map;    // > 0: output column c goes to map c / cols_per_map at channel c % cols_per_map (merged dgrad); power of 2
  int cpm_shift;       // log2(cols_per_map)
  int exp_a_off, exp_a_sbo, exp_a_baseoff;  // bring-up experiment hooks for the A descriptor (0 = normal)
  int exp_b_off, exp_b_lbo;                 // same for the B descriptor (MN-major shifted-view experiments)
  int pair;                                 // 1: CTA-pair mode (cta_group::2, see the kernel)
  int pair_b_dim, pair_b_off;               // B tensor-map coordinate that selects the second CTA's half of the B tile
  // patch mode: one TMA-loaded input patch per stage serves `taps` shifted A views (convolution taps); B (the whole
  // weight matrix) is loaded once per CTA and stays resident in shared memory.
  int taps;             // MMA groups per stage (1 = plain GEMM)
  int tap_off[4];       // byte offset of each tap's A view inside the stage
  int tap_acc[4];       // TMEM column offset of the accumulator block each tap adds to (sub-tiles of one staged patch)
  int tap_fresh;        // bit t: tap t is the first contribution to its accumulator block (0 = only tap 0)
  int sub_panels;       // > 0: the tile's 32-column panels are `sub_panels` channel panels per sub-tile, sub-tile after
                        // sub-tile (bias / channel index of panel q = (q % sub_panels) * 32); 0 = panels are channel panels
  int b_resident;       // 1: B slabs are loaded once per CTA
  int b_slabs;          // resident 32-wide K slabs
  int b_slab_bytes;     // bn * 128
  unsigned char b_tab[64];  // resident slab used by (k-iteration, tap): b_tab[k * taps + t]
  // slab mode (wgrad): one A slab per stage is multiplied with `ngroups` shifted views of the B patches, each an
  // independent mma_n-wide MMA into its own TMEM column block; the tile's bn columns are the union of the blocks.
  int ngroups;              // 0 = off
  int grp_b_off[16];        // byte offset of each group's B view inside the stage's B region
  int grp_acc[16];          // TMEM column offset of each group's accumulator block
  int mma_n;                // N of one MMA (0 = bn)
  int mma_m;                // M of one MMA: 0 / 128, or 64 when the problem has <= 64 rows (wgrad with Cout <= 64): the SS-mode
                            // MMA is bound by its shared-memory operand reads, and a 64-row A tile halves A's share.
                            // TMEM rows then sit in lanes 0-15 of each 32-lane quarter (row = 16*quarter + lane)
  int acc_stages;           // TMEM accumulator stages: 2, or 1 when bn > 256
  int panel_tab0[16], panel_tab1[16];  // slab mode: output panel q -> added to output coordinates 0 and 1
  // LeakyReLU' bitmask (1 bit per fp32 element of an activation tensor, same linear order): written by the
  // bias+LeakyReLU epilogue, read by EPI_MASK instead of TMA-loading the fp32 activation tile.
  unsigned* bits_out;        // nullable
  const unsigned* bits_in;   // nullable; with EPI_MASK replaces mapX
  int row_box[3];            // output tile row r -> (r % row_box[0], (r / row_box[0]) % row_box[1], r / (row_box[0]*row_box[1]));
                             // rows with the third coordinate >= row_box[2] are not part of the tile
  int row_ext[4][3];         // valid extents of the three row coordinates, per output map
  long bit_str[3];           // element strides of the three row coordinates in the tensor the bits describe
  long bit_base[4];          // element offset of each output map's origin in that tensor
  float slope;
  const float* bias;
  // Column sums of the stored (masked) output, accumulated into colsum_out[column & colsum_mask] (fp32 atomics): the bias
  // gradient of the layer whose pre-activation gradient this dgrad produces, taken from the staged tile instead of a
  // separate pass over the tensor.  Bit-mask epilogue only.  Rows whose coordinate `colsum_dim` (0..2, see row_box) is
  // >= colsum_limit are stored but not summed (gradient-penalty rows of the critic batch).
  float* colsum_out;         // nullable
  int colsum_mask, colsum_dim, colsum_limit;
  // debug (GC_UMMA_STATS=1): per-CTA clocks each role spent waiting on its barriers, 8 counters per CTA; nullptr = off
  long long* stats;
};

}  // namespace gcu
