// The tcgen05 / TMEM / TMA kernel behind every dense contraction (see gc_umma.cuh for the design).
#pragma once
#include "gc_common.cuh"
#include "gc_umma.cuh"

namespace gcu {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a barrier that never completes (a descriptor / byte-count bug) traps after ~4 s instead of
// hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 8000000000LL) __trap();
  }
}

// same, accumulating the clocks spent waiting when the debug counters are on
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool timed, long long& acc) {
  if (timed) {
    const long long t = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t;
  } else {
    mbar_wait(bar, parity);
  }
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, const int (&c)[5]) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
      : "memory");
}
// ---- CTA-pair (cta_group::2) helpers: the pair's leader is the even CTA of the 2-CTA cluster ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are accounted on the LEADER's mbarrier (cluster address)
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, const int (&c)[5]) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(leader_bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// arrive (when all prior MMAs of this thread retire) on the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, const int (&c)[5]) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"((uint64_t)map), "r"(src), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout type.
// K-major operands: SWIZZLE_128B (2), 8-row groups 1024 B apart.  MN-major tf32 operands only exist as
// SWIZZLE_128B_BASE32B (1): 128 B rows of 32 elements along M/N, 4-row K atoms 512 B apart (SBO), 32-element panels LBO apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// One lane of a converged warp.  Unlike `lane == 0`, elect.sync tells the compiler that exactly one thread runs the
// guarded region, so tcgen05.mma / TMA operands go straight to uniform registers instead of through a per-instruction
// ELECT / R2UR.BROADCAST / BRA.U.ANY serialisation loop (~100 clocks per MMA on the issuing thread).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// explicit shared-memory accesses with 32-bit addresses: through a generic pointer the compiler emits generic ST.E/LD.E
// with 64-bit address arithmetic for every 16-byte chunk (profiles/r01_ncu_conv1_fprop_source.txt)
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ uint64_t mk64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

__device__ __forceinline__ void coords(const TmaAddr& t, const int (&src)[kSrc], int lo, int hi, int (&c)[5]) {
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    int v = 0;
    for (int s = lo; s < hi; ++s) v += t.mul[d][s] * src[s];
    c[d] += v;
  }
}

// coordinates of panel q of an operand / output whose tile-level coordinates are `base`
__device__ __forceinline__ void panel_coords(const TmaAddr& t, const int (&base)[5], int q, int (&c)[5]) {
  const int q1 = t.period > 0 ? q % t.period : q;
  const int q2 = t.period > 0 ? q / t.period : 0;
#pragma unroll
  for (int d = 0; d < 5; ++d) c[d] = base[d] + q1 * t.panel[d] + q2 * t.panel2[d];
}

// ------------------------------------------------------------------ the kernel
__device__ __forceinline__ void decode_tile(const GemmParams& p, int tile, int (&src)[kSrc], int& n_tile) {
  const int mt = tile % p.mt;
  const int r = tile / p.mt;
  n_tile = r % p.nt;
  src[M0] = mt % p.e0;
  const int t = mt / p.e0;
  src[M1] = t % p.e1;
  src[M2] = t / p.e1;
  src[N0] = n_tile % p.f0;
  src[N1] = n_tile / p.f0;
  src[K0] = src[K1] = src[K2] = 0;
  src[Z] = r / p.nt;
}

__device__ __forceinline__ void tile_coords(const TmaAddr& t, const int (&src)[kSrc], int (&c)[5]) {
#pragma unroll
  for (int d = 0; d < 5; ++d) c[d] = t.off[d];
  coords(t, src, 0, 5, c);
  coords(t, src, Z, Z + 1, c);
}

// Division-free walk over this CTA's tiles.  A lone producer / MMA / epilogue thread pays ~200 clocks of dependent
// latency per runtime integer division (profiles/r01_umma_role_stats.txt: ~800-1000 clocks of scalar overhead per
// k-iteration before this), so the tile index is kept as mixed-radix digits (m0 m1 m2 | n0 n1 | z) and advanced by
// adding the digits of the stride with carries; the only divisions left run once per kernel.
struct TileWalk {
  int dig[6], step[6], rad[5];
  int tile, tstep;
};
__device__ __forceinline__ void walk_digits(const GemmParams& p, int x, int (&d)[6]) {
  const int m = x % p.mt, r = x / p.mt;
  const int n = r % p.nt;
  d[5] = r / p.nt;
  d[0] = m % p.e0;
  const int t = m / p.e0;
  d[1] = t % p.e1;
  d[2] = t / p.e1;
  d[3] = n % p.f0;
  d[4] = n / p.f0;
}
__device__ __forceinline__ void walk_init(TileWalk& w, const GemmParams& p, int start, int step) {
  walk_digits(p, start, w.dig);
  walk_digits(p, step, w.step);
  w.rad[0] = p.e0; w.rad[1] = p.e1; w.rad[2] = p.mt / (p.e0 * p.e1); w.rad[3] = p.f0; w.rad[4] = p.nt / p.f0;
  w.tile = start;
  w.tstep = step;
}
// returns the carries out of digits 0..4 as a bit mask
__device__ __forceinline__ unsigned walk_next(TileWalk& w) {
  int c = 0;
  unsigned carries = 0u;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int v = w.dig[i] + w.step[i] + c;
    c = v >= w.rad[i] ? 1 : 0;
    carries |= (unsigned)c << i;
    w.dig[i] = c ? v - w.rad[i] : v;
  }
  w.dig[5] += w.step[5] + c;
  w.tile += w.tstep;
  return carries;
}

// Tile-level TMA coordinates of one operand / output kept in registers and advanced together with the digits: the
// coordinates are linear in the digits, so a step adds a constant vector plus one correction vector per carry.  The
// per-tile recomputation (30 multiply-adds fed by constant-bank loads) was ~1000 of the ~2300 clocks the producer and the
// epilogue spent per 128x32 conv1 tile.
struct CoordTrack {
  int c[5], base[5], corr[5][5];
};
__device__ __forceinline__ void track_init(CoordTrack& t, const TmaAddr& a, const TileWalk& w) {
  constexpr int kDigitSrc[6] = {M0, M1, M2, N0, N1, Z};
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    int c = a.off[d], b = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      c += a.mul[d][kDigitSrc[i]] * w.dig[i];
      b += a.mul[d][kDigitSrc[i]] * w.step[i];
    }
    t.c[d] = c;
    t.base[d] = b;
#pragma unroll
    for (int i = 0; i < 5; ++i) t.corr[i][d] = a.mul[d][kDigitSrc[i + 1]] - w.rad[i] * a.mul[d][kDigitSrc[i]];
  }
}
__device__ __forceinline__ void track_next(CoordTrack& t, unsigned carries) {
  if (carries == 0u) {   // the common step (no digit wrapped): five additions instead of thirty select-adds (warp-uniform branch)
#pragma unroll
    for (int d = 0; d < 5; ++d) t.c[d] += t.base[d];
    return;
  }
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    int v = t.c[d] + t.base[d];
#pragma unroll
    for (int i = 0; i < 5; ++i) v += ((carries >> i) & 1u) ? t.corr[i][d] : 0;
    t.c[d] = v;
  }
}
__device__ __forceinline__ void walk_src(const TileWalk& w, const GemmParams& p, int (&src)[kSrc], int& n_tile) {
  src[M0] = w.dig[0]; src[M1] = w.dig[1]; src[M2] = w.dig[2];
  src[N0] = w.dig[3]; src[N1] = w.dig[4];
  src[K0] = src[K1] = src[K2] = 0;
  src[Z] = w.dig[5];
  n_tile = w.dig[4] * p.f0 + w.dig[3];
}
// per-k-iteration coordinate increments of an operand: k0 advances | k0 wraps, k1 advances | k0 and k1 wrap, k2 advances
__device__ __forceinline__ void k_deltas(const TmaAddr& t, int g0, int g1, int (&d0)[5], int (&d1)[5], int (&d2)[5]) {
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    const unsigned back0 = (unsigned)(g0 - 1) * (unsigned)t.mul[d][K0], back1 = (unsigned)(g1 - 1) * (unsigned)t.mul[d][K1];
    d0[d] = t.mul[d][K0];
    d1[d] = (int)((unsigned)t.mul[d][K1] - back0);
    d2[d] = (int)((unsigned)t.mul[d][K2] - back1 - back0);
  }
}

// Persistent, warp-specialised: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
//   warp 8   : TMA producer (one lane) + TMEM owner          -> smem ring (full/empty mbarriers)
//   warp 9   : tcgen05.mma issuer (one lane)                 -> two TMEM accumulator stages (tmem_full/tmem_empty)
//   warps 0-3, 4-7: two epilogue groups (TMEM lanes 32(w%4)..+31); group g drains accumulator stage g, i.e. every
//              other tile, -> staging smem -> TMA store.  The epilogue is a per-warp dependency chain (~3000 clocks
//              per 128x32 tile on one group, ncu: profiles/r01_ncu_conv1_fprop_details.txt), so two groups in
//              ping-pong double its throughput; with TMA-loaded mask tiles (EPI_MASK without bit masks) only group 0 runs.
// The epilogue of tile i overlaps the mainloop of tile i+1 and barrier/TMEM set-up is paid once per SM.
// CTA2: the CTAs of a 2-CTA cluster work as a pair on a 256 x N tile (tcgen05 cta_group::2): each stages its own 128 rows
// of A and HALF of the B tile, the leader (even CTA) issues M=256 MMAs that read both halves, each CTA's TMEM receives
// its 128 accumulator rows and each CTA runs its own epilogue.  TMA requests per CTA and k-iteration drop from 128 + N
// rows to 128 + N/2 - the per-SM TMA row rate (~0.5 x 128 B rows per clock) is what bounds the large layers
// (profiles/r01_operand_skip_experiment.txt).  Barriers: `full` lives in the leader and collects the bytes of both CTAs;
// `empty` / `tmem_full` are signalled in both CTAs by multicast commits; the leader's `tmem_empty` counts the epilogue
// warps of both CTAs.
template <bool A_MN, bool B_MN, bool SLAB, bool CTA2 = false>
__global__ void __launch_bounds__(320, 1) umma_gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = p.a_bytes + p.b_bytes;
  uint8_t* bres = smem + (size_t)p.stages * stage_bytes;       // resident B slabs (patch mode), 1024-aligned
  uint8_t* staging = bres + (size_t)p.b_slabs * p.b_slab_bytes;  // nbuf x 16 KB, 1024-aligned
  uint64_t* bars = (uint64_t*)(staging + (size_t)p.nbuf * 16384);
  // bars: full[stages], empty[stages], tmem_full[2], tmem_empty[2], aux[4], bres
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * p.stages + 9);
  // resident-weights mode: low descriptor word of the B slab used by (k-iteration, tap), 4 entries per k-iteration,
  // filled once by the MMA thread (the per-tap constant-bank lookups were ~100 clocks of its ~105 per MMA)
  uint32_t* s_btab = (uint32_t*)(((uintptr_t)(tmem_slot + 4) + 15) & ~(uintptr_t)15);
  float* s_bias = (float*)(s_btab + 64);  // nt*bn floats (<= 512) for bias epilogues
  float* s_colsum = s_bias + 512;         // <= 256 per-channel column sums of the stored tiles (p.colsum_out)

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * p.stages;
  const uint32_t tfull0 = empty0 + 8 * p.stages, tempty0 = tfull0 + 16, aux0 = tempty0 + 16, bres_bar = aux0 + 32;
  const int total_tiles = p.mt * p.nt * p.zt;
  const int acc_cols = ((p.bn + 31) >> 5) << 5;
  const int acc_stages = p.acc_stages;
#ifdef GC_UMMA_STATS_BUILD
  const bool timed = p.stats != nullptr;   // per-role wait counters (debug build only: python -m gail_carla_b200.build with GC_UMMA_STATS_BUILD=1)
#else
  constexpr bool timed = false;            // production build: the clock reads and counter updates are compiled out
#endif
  const uint32_t pair_rank = CTA2 ? cluster_ctarank() : 0u;   // 0 = leader

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, CTA2 ? 8 : 4);  // one arrival per epilogue warp (of both CTAs of a pair)
    }
    for (int q = 0; q < 4; ++q) mbar_init(aux0 + 8 * q, 1);
    mbar_init(bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == 8) {
    if constexpr (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if ((p.epilogue == EPI_BIAS_LRELU || p.epilogue == EPI_BIAS) && p.nt * p.bn <= 512) {
    for (int i = threadIdx.x; i < p.nt * p.bn; i += blockDim.x) s_bias[i] = i < p.n_total ? __ldg(p.bias + i) : 0.f;
  }
  if (p.colsum_out != nullptr) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_colsum[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (elect_one()) {
      const uint32_t tx = p.a_panels * p.a_panel_bytes + (p.b_resident ? 0 : p.b_panels * p.b_panel_bytes);
      if (p.b_resident) {  // the whole weight matrix, once per CTA
        mbar_expect_tx(bres_bar, (uint32_t)(p.b_slabs * p.b_slab_bytes));
        for (int s = 0; s < p.b_slabs; ++s) {
          const int cw[5] = {32 * s, 0, 0, 0, 0};
          tma_load_5d(smem_u32(bres) + s * p.b_slab_bytes, &p.mapB, bres_bar, cw);
        }
      }
      long long w_empty = 0;
      int s = 0;
      uint32_t ph = 0;
      int da0[5], da1[5], da2[5], db0[5], db1[5], db2[5];
      k_deltas(p.a, p.g0, p.g1, da0, da1, da2);
      k_deltas(p.b, p.g0, p.g1, db0, db1, db2);
      const int a_panels = p.a_panels, b_panels = p.b_resident ? 0 : p.b_panels, b_period = p.b.period;
      const int k_iters = p.k_iters, stages = p.stages, g0 = p.g0, g1 = p.g1;
      int pa[5], pb[5], pb2[5];
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        pa[d] = p.a.panel[d];
        pb[d] = p.b.panel[d];
        pb2[d] = p.b.panel2[d] - (b_period - 1) * p.b.panel[d];
      }
      TileWalk w;
      walk_init(w, p, blockIdx.x, gridDim.x);
      const bool tracked = p.kz_stride == 0;   // split-K tiles start at a z-dependent k-iteration: computed per tile
      CoordTrack ta_, tb_;
      track_init(ta_, p.a, w);
      track_init(tb_, p.b, w);
      for (; w.tile < total_tiles;) {
        int src[kSrc], n_tile, ca[5], cb[5];
        walk_src(w, p, src, n_tile);
        int k0 = 0, k1 = 0;
        if (p.kz_stride != 0) {  // flat split-K: this tile starts at k-iteration z * kz_stride
          const int kit = src[Z] * p.kz_stride;
          k0 = kit % g0;
          const int t = kit / g0;
          k1 = t % g1;
          src[K0] = k0; src[K1] = k1; src[K2] = t / g1;
        }
        if (tracked) {
#pragma unroll
          for (int d = 0; d < 5; ++d) { ca[d] = ta_.c[d]; cb[d] = tb_.c[d]; }
        } else {
#pragma unroll
          for (int d = 0; d < 5; ++d) { ca[d] = p.a.off[d]; cb[d] = p.b.off[d]; }
          coords(p.a, src, 0, kSrc, ca);
          coords(p.b, src, 0, kSrc, cb);
        }
        if constexpr (CTA2) {   // this CTA's half of the B tile
#pragma unroll
          for (int d = 0; d < 5; ++d) cb[d] += (d == p.pair_b_dim) ? (int)pair_rank * p.pair_b_off : 0;
        }
        for (int k = 0; k < k_iters; ++k) {
          mbar_wait_t(empty0 + 8 * s, ph ^ 1, timed, w_empty);
          const uint32_t full_s = CTA2 ? mapa_rank(full0 + 8 * s, 0) : full0 + 8 * s;   // pair mode: the leader's barrier
          if constexpr (CTA2) {
            if (pair_rank == 0) mbar_expect_tx(full0 + 8 * s, 2u * tx);   // the bytes of both CTAs land on the leader's barrier
          } else {
            mbar_expect_tx(full0 + 8 * s, tx);
          }
          const uint32_t sa = smem_u32(smem) + (uint32_t)s * (uint32_t)stage_bytes, sb = sa + p.a_bytes;
          if (a_panels == 1 && b_panels <= 1) {   // the common case, straight-line (the panel loops below unroll badly)
            if constexpr (CTA2) tma_load_5d_pair(sa, &p.mapA, full_s, ca); else tma_load_5d(sa, &p.mapA, full_s, ca);
            if (b_panels == 1) {
              if constexpr (CTA2) tma_load_5d_pair(sb, &p.mapB, full_s, cb); else tma_load_5d(sb, &p.mapB, full_s, cb);
            }
          } else {
            // panels are walked incrementally (no div/mod on the single producer thread's critical path)
            int cq[5];
#pragma unroll
            for (int d = 0; d < 5; ++d) cq[d] = ca[d];
#pragma unroll 1
            for (int q = 0; q < a_panels; ++q) {
              tma_load_5d(sa + q * p.a_panel_bytes, &p.mapA, full_s, cq);
#pragma unroll
              for (int d = 0; d < 5; ++d) cq[d] += pa[d];
            }
#pragma unroll
            for (int d = 0; d < 5; ++d) cq[d] = cb[d];
            int q1 = 0;
#pragma unroll 1
            for (int q = 0; q < b_panels; ++q) {
              tma_load_5d(sb + q * p.b_panel_bytes, &p.mapB, full_s, cq);
              if (++q1 == b_period) {
                q1 = 0;
#pragma unroll
                for (int d = 0; d < 5; ++d) cq[d] += pb2[d];
              } else {
#pragma unroll
                for (int d = 0; d < 5; ++d) cq[d] += pb[d];
              }
            }
          }
          // next k-iteration: (k0, k1, k2) advance as mixed-radix digits, the coordinates by the matching increments
          if (++k0 == g0) {
            k0 = 0;
            if (++k1 == g1) {
              k1 = 0;
#pragma unroll
              for (int d = 0; d < 5; ++d) { ca[d] += da2[d]; cb[d] += db2[d]; }
            } else {
#pragma unroll
              for (int d = 0; d < 5; ++d) { ca[d] += da1[d]; cb[d] += db1[d]; }
            }
          } else {
#pragma unroll
            for (int d = 0; d < 5; ++d) { ca[d] += da0[d]; cb[d] += db0[d]; }
          }
          if (++s == stages) { s = 0; ph ^= 1; }
        }
        const unsigned carries = walk_next(w);
        track_next(ta_, carries);
        track_next(tb_, carries);
      }
      if (timed) p.stats[blockIdx.x * 16 + 0] = w_empty;
    }
    __syncwarp();
  } else if (warp == 9) {
    if (pair_rank == 0 && elect_one()) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, majors, N>>3, M>>4 (M = 256 across a CTA pair)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)((p.mma_n ? p.mma_n : p.bn) >> 3) << 17) |
                             (((CTA2 ? 256u : (p.mma_m == 64 ? 64u : 128u)) >> 4) << 24);
      const int ksteps = p.bk >> 3;
      const uint32_t a_lbo = A_MN ? (uint32_t)p.a_panel_bytes : 0u;
      const uint32_t b_lbo = p.exp_b_lbo ? (uint32_t)p.exp_b_lbo : (B_MN ? (uint32_t)p.b_panel_bytes : 0u);
      int ti = 0;
      long long w_full = 0, w_tempty = 0;
      const long long t_start = timed ? clock64() : 0;
      const uint32_t a_sbo = p.exp_a_sbo ? (uint32_t)p.exp_a_sbo : (A_MN ? 512u : 1024u);
      // constant descriptor halves (see umma_desc): lo = start>>4 | LBO>>4 << 16, hi = SBO>>4 | version | base_offset | layout
      const uint64_t a_proto = umma_desc(0, a_lbo, a_sbo, A_MN ? 1 : 2) | ((uint64_t)(p.exp_a_baseoff & 7) << 49);
      const uint64_t b_proto = umma_desc(0, b_lbo, B_MN ? 512 : 1024, B_MN ? 1 : 2);
      const uint32_t a_lo_base = (uint32_t)a_proto, a_hi = (uint32_t)(a_proto >> 32);
      const uint32_t b_lo_base = (uint32_t)b_proto, b_hi = (uint32_t)(b_proto >> 32);
      const uint32_t a_step = A_MN ? 64u : 2u, b_step = B_MN ? 64u : 2u;  // per K-step: 1024 B or 32 B, in 16-byte units
      const int n_taps = p.taps;
      const bool b_res = p.b_resident != 0;
      const uint32_t bres_u32 = smem_u32(bres), b_slab = (uint32_t)p.b_slab_bytes, a_off0 = (uint32_t)p.exp_a_off;
      const uint32_t smem0 = smem_u32(smem), a_bytes = (uint32_t)p.a_bytes, b_off = (uint32_t)p.exp_b_off;
      const int k_iters = p.k_iters, stages = p.stages, ngroups = p.ngroups;
      uint32_t gb[8], ga[8];   // slab mode: per-group B descriptor constant and TMEM column offset (<= 8 groups, host-checked)
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        gb[g] = SLAB ? b_lo_base + ((uint32_t)p.grp_b_off[g] >> 4) : 0u;
        ga[g] = SLAB ? (uint32_t)p.grp_acc[g] : 0u;
      }
      uint32_t ta[4], tacc_off[4];   // per-tap A descriptor constants (taps are shifted views of one staged patch) and
                                     // accumulator column blocks (sub-tiles of the patch)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        ta[t] = a_lo_base + ((a_off0 + (uint32_t)p.tap_off[t]) >> 4);
        tacc_off[t] = (uint32_t)p.tap_acc[t];
      }
      const uint32_t fresh = p.tap_fresh ? (uint32_t)p.tap_fresh : 1u;
      const uint32_t s_btab_u32 = smem_u32(s_btab);
      if (p.b_resident) {
        for (int i = 0; i < k_iters; ++i)
          for (int t = 0; t < 4; ++t)
            s_btab[i * 4 + t] = t < n_taps ? b_lo_base + ((bres_u32 + (uint32_t)p.b_tab[i * n_taps + t] * b_slab) >> 4) : 0u;
        mbar_wait(bres_bar, 0);
        tc_fence_after();
      }
      // stage ring and accumulator ring positions are counters, not it % stages (no divisions on this thread)
      int s = 0, acc = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
        mbar_wait_t(tempty0 + 8 * acc, aph ^ 1, timed, w_tempty);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * acc_cols);
        for (int k = 0; k < k_iters; ++k) {
          mbar_wait_t(full0 + 8 * s, ph, timed, w_full);
          tc_fence_after();
          const uint32_t sa = smem0 + (uint32_t)s * (uint32_t)stage_bytes, sb = sa + a_bytes;
          // Descriptors differ only in their 14-bit start-address field, so the loop adds to precomputed 32-bit
          // halves; the 4 K-steps of a 128-byte K-major slab are unrolled (the MMA of a 128xN tile with small N takes
          // only N/2 cycles, so the single issuing thread must not spend more than that per instruction).
          if constexpr (SLAB) {
            // slab mode: every group multiplies the same A slab with its own shifted view of the B patches
            const uint32_t alo = a_lo_base + (sa >> 4), sb4 = sb >> 4;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (g < ngroups) {
                const uint32_t blo = sb4 + gb[g], tg = tacc + ga[g];
                for (int ks = 0; ks < ksteps; ++ks)
                  umma_tf32(tg, mk64(alo + ks * a_step, a_hi), mk64(blo + ks * b_step, b_hi), idesc, (k | ks) ? 1u : 0u);
              }
            }
          } else {
            if (b_res && ksteps == 4) {
              // patch modes: every descriptor is (stage base >> 4) + a per-tap constant (A) or a table entry (B)
              const uint32_t sa4 = sa >> 4;
              uint32_t bl[4];
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(bl[0]), "=r"(bl[1]), "=r"(bl[2]), "=r"(bl[3])
                           : "r"(s_btab_u32 + (uint32_t)k * 16u) : "memory");
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                if (t < n_taps) {
                  const uint32_t alo = sa4 + ta[t], blo = bl[t], tt = tacc + tacc_off[t];
                  umma_tf32(tt, mk64(alo, a_hi), mk64(blo, b_hi), idesc, (k == 0 && ((fresh >> t) & 1u)) ? 0u : 1u);
                  umma_tf32(tt, mk64(alo + a_step, a_hi), mk64(blo + b_step, b_hi), idesc, 1u);
                  umma_tf32(tt, mk64(alo + 2 * a_step, a_hi), mk64(blo + 2 * b_step, b_hi), idesc, 1u);
                  umma_tf32(tt, mk64(alo + 3 * a_step, a_hi), mk64(blo + 3 * b_step, b_hi), idesc, 1u);
                }
              }
            } else {
              const uint32_t alo = a_lo_base + ((sa + a_off0) >> 4);
              const uint32_t blo = b_lo_base + ((sb + b_off) >> 4);
              if constexpr (CTA2) {
                for (int ks = 0; ks < ksteps; ++ks)
                  umma_tf32_pair(tacc, mk64(alo + ks * a_step, a_hi), mk64(blo + ks * b_step, b_hi), idesc, (k | ks) ? 1u : 0u);
              } else if (ksteps == 4) {
                umma_tf32(tacc, mk64(alo, a_hi), mk64(blo, b_hi), idesc, k ? 1u : 0u);
                umma_tf32(tacc, mk64(alo + a_step, a_hi), mk64(blo + b_step, b_hi), idesc, 1u);
                umma_tf32(tacc, mk64(alo + 2 * a_step, a_hi), mk64(blo + 2 * b_step, b_hi), idesc, 1u);
                umma_tf32(tacc, mk64(alo + 3 * a_step, a_hi), mk64(blo + 3 * b_step, b_hi), idesc, 1u);
              } else {
                for (int ks = 0; ks < ksteps; ++ks)
                  umma_tf32(tacc, mk64(alo + ks * a_step, a_hi), mk64(blo + ks * b_step, b_hi), idesc, (k | ks) ? 1u : 0u);
              }
            }
          }
          if constexpr (CTA2) umma_commit_pair(empty0 + 8 * s);   // frees the stage in both CTAs
          else umma_commit(empty0 + 8 * s);  // frees the smem stage when these MMAs retire
          if (++s == stages) { s = 0; ph ^= 1; }
        }
        if constexpr (CTA2) umma_commit_pair(tfull0 + 8 * acc);
        else umma_commit(tfull0 + 8 * acc);
        if (++acc == acc_stages) { acc = 0; aph ^= 1; }
      }
      if (timed) {
        p.stats[blockIdx.x * 16 + 1] = w_full;
        p.stats[blockIdx.x * 16 + 2] = w_tempty;
        p.stats[blockIdx.x * 16 + 5] = clock64() - t_start;
        p.stats[blockIdx.x * 16 + 7] = ti;
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: thread t owns accumulator row 32*warp + lane ----------------
    // Output leaves in 32-column panels through `nbuf` 16 KB staging buffers (a ring over all panels of all tiles
    // of this CTA).  With EPI_MASK the panel's LeakyReLU' source tile (same box geometry as the output) is
    // TMA-loaded into the staging buffer two panels ahead, multiplied in place and stored from the same buffer.
    const int grp = warp >> 2, wq = warp & 3;      // epilogue group, TMEM lane quarter
    const bool m64 = p.mma_m == 64;                // M = 64 accumulators: 16 rows per lane quarter, in its lanes 0-15
    const int row = m64 ? wq * 16 + lane : wq * 32 + lane;
    const bool row_live = !m64 || lane < 16;
    // the group's TMA traffic is issued by the elected lane of its first warp (elect.sync picks the same lane every time,
    // so bulk-group waits see the stores that lane committed)
    auto leader = [&]() -> bool { return wq == 0 && elect_one(); };
    const int n_panels = (p.bn + 31) >> 5;
    const bool swz = p.d_row_bytes == 128;
    const bool masked = p.epilogue == EPI_MASK && p.bits_in == nullptr;  // TMA-loaded mask tiles
    const bool bitmask = p.epilogue == EPI_MASK && p.bits_in != nullptr;
    const bool two_groups = !masked && acc_stages == 2;  // ping-pong on the two accumulator stages, nbuf/2 staging buffers each
    const int nbuf = two_groups ? (p.nbuf >> 1) : p.nbuf;
    uint8_t* const my_staging = staging + (two_groups ? grp * nbuf * 16384 : 0);
    const int tile_step = two_groups ? 2 * (int)gridDim.x : (int)gridDim.x;
    const bool active = two_groups || grp == 0;
    const int r0 = row % p.row_box[0], r1 = (row / p.row_box[0]) % p.row_box[1], r2 = row / (p.row_box[0] * p.row_box[1]);
    int pf_tile = blockIdx.x, pf_q = 0, pf_count = 0;  // mask prefetch cursor (group leader only; single-group mode)
    // output panel q of a tile -> which tensor map and which coordinates
    const int op_cpm = p.cols_per_map, op_sh = p.cpm_shift, op_off0 = p.d.off[0], op_bn = p.bn;
    auto out_panel = [&](const int (&base)[5], int n_tile, int q, int (&c)[5]) -> int {
      if (op_cpm > 0) {
        const int col = n_tile * op_bn + q * 32;
#pragma unroll
        for (int d = 0; d < 5; ++d) c[d] = base[d];
        c[0] = op_off0 + (col & (op_cpm - 1));   // cols_per_map is a power of two (host-checked)
        return col >> op_sh;
      }
      if constexpr (SLAB) {
#pragma unroll
        for (int d = 0; d < 5; ++d) c[d] = base[d];
        c[0] += p.panel_tab0[q];
        c[1] += p.panel_tab1[q];
        return 0;
      }
      panel_coords(p.d, base, q, c);
      return 0;
    };
    auto prefetch_one = [&]() {
      if (pf_tile >= total_tiles) return;
      int src[kSrc], nt, cb_[5], cx[5];
      decode_tile(p, pf_tile, src, nt);
      tile_coords(p.d, src, cb_);
      const int mi = out_panel(cb_, nt, pf_q, cx);
      const int j = pf_count % nbuf;
      mbar_expect_tx(aux0 + 8 * j, (uint32_t)p.d_box_bytes);
      tma_load_5d(smem_u32(staging + j * 16384), &p.mapX[mi], aux0 + 8 * j, cx);
      ++pf_count;
      if (++pf_q == n_panels) { pf_q = 0; pf_tile += gridDim.x; }
    };
    if (masked && active && leader()) {
      for (int i = 0; i < nbuf - 2; ++i) prefetch_one();  // mask prefetch distance = nbuf - 2 panels
    }
    // this thread's eight 16-byte chunk offsets inside a staging buffer (row-major rows of d_row_bytes, 128-byte rows
    // XOR-swizzled like the TMA box), computed once
    const int nchunk = p.d_row_bytes >> 4;
    uint32_t soff[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t o = (uint32_t)row * (uint32_t)p.d_row_bytes + j * 16;
      if (swz) o ^= ((o >> 7) & 7) << 4;
      soff[j] = o;
    }
    const long bs0 = p.bit_str[0], bs1 = p.bit_str[1], bs2 = p.bit_str[2], bbase0 = p.bit_base[0];
    const long thread_bit_off = r0 * bs0 + r1 * bs1 + r2 * bs2;
    const int ext0[3] = {p.row_ext[0][0], p.row_ext[0][1], p.row_ext[0][2]};
    const bool r2_ok = r2 < p.row_box[2];
    const int epi = p.epilogue;
    const int sub_panels = p.sub_panels;
    const float slope = p.slope;
    const bool bias_smem = p.nt * p.bn <= 512;
    const bool want_bits = p.bits_out != nullptr && epi == EPI_BIAS_LRELU;
    const bool want_colsum = p.colsum_out != nullptr && bitmask && swz && !m64;
    const int cs_dim = p.colsum_dim, cs_limit = p.colsum_limit, cs_mask = p.colsum_mask;
    // fprop tiles: consecutive panels are consecutive 32-channel groups of the same pixels, i.e. consecutive bit words -
    // the word index and the clipping test are then formed once per tile instead of once per panel
    const bool simple_panels = !SLAB && p.cols_per_map == 0 && p.d.period == 0 && p.d.panel[0] == 32 && p.d.panel[1] == 0 &&
                               p.d.panel[2] == 0 && p.d.panel[3] == 0 && p.d.panel[4] == 0;
    // sub-tile mode with one channel panel per sub-tile: panel q = sub-tile q, the same channels `panel2` pixels further
    const bool sub_regular = !SLAB && p.cols_per_map == 0 && sub_panels == 1;
    const int sp1 = p.d.panel2[1], sp2 = p.d.panel2[2], sp3 = p.d.panel2[3];
    const long sub_words = (sp1 * bs0 + sp2 * bs1 + sp3 * bs2) >> 5;
    const uint32_t s_bias_u32 = smem_u32(s_bias), my_staging_u32 = smem_u32(my_staging);
    // Everything the per-panel code needs from the kernel parameters lives in registers: an ncu source-level capture of conv1
    // fprop (profiles/r02_ncu_conv1_fprop_bits_summary.txt) showed ~700 warp instructions per 128x32 panel of which ~150 were
    // the epilogue's arithmetic - the rest were constant-bank reloads of these fields (LDC: 14 % of the stall samples), the
    // q % period / q / period divisions of the panel walk (MUFU.RCP sequences) and the debug clock reads.
    const int d_period = p.d.period, cpm = p.cols_per_map, cpm_sh = p.cpm_shift, d_off0 = p.d.off[0], tile_bn = p.bn,
              n_total = p.n_total;
    int dpan[5], dwrap[5];         // panel walk: + dpan per panel, + dwrap when the first level wraps (two-level walk)
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      dpan[d] = p.d.panel[d];
      dwrap[d] = p.d.panel2[d] - (d_period > 0 ? d_period - 1 : 0) * p.d.panel[d];
    }
    unsigned* const bits_out = p.bits_out;
    const unsigned* const bits_in = p.bits_in;
    int pc = 0, bi = 0;            // panels done so far, staging buffer of the current panel (pc % nbuf without the division)
    uint32_t bph = 0;              // (pc / nbuf) & 1
    long long w_tfull = 0, e_free = 0, e_ld = 0, e_math = 0, e_store = 0, e_sts = 0, e_bar = 0;
    const long long t_epi0 = timed ? clock64() : 0;
    TileWalk w;
    walk_init(w, p, active ? (int)blockIdx.x + (two_groups ? grp * (int)gridDim.x : 0) : total_tiles, tile_step);
    int acc = two_groups ? grp : 0;   // accumulator stage drained by this group for the current tile, and its phase
    uint32_t aph = 0;
    CoordTrack td_;
    track_init(td_, p.d, w);
    for (; w.tile < total_tiles; track_next(td_, walk_next(w))) {
      int src[kSrc], n_tile, cd[5];
      walk_src(w, p, src, n_tile);
#pragma unroll
      for (int d = 0; d < 5; ++d) cd[d] = td_.c[d];
      // 32-bit word (in the bit tensor of the activation the mask describes) of this thread's row in panel q, or -1 if the
      // row is clipped.  Strides / extents of map 0 live in registers (hoisted above the tile loop).
      bool in_limit = true;   // side result of bit_word: this row counts towards the column sums (p.colsum_limit)
      auto bit_word = [&](int q) -> long {
        int cq[5];
        const int mi = out_panel(cd, n_tile, q, cq);
        const int c1 = cq[1] + r0, c2 = cq[2] + r1, c3 = cq[3] + r2;
        const int x0 = mi == 0 ? ext0[0] : p.row_ext[mi][0], x1 = mi == 0 ? ext0[1] : p.row_ext[mi][1],
                  x2 = mi == 0 ? ext0[2] : p.row_ext[mi][2];
        if (want_colsum) in_limit = (cs_dim == 0 ? c1 : (cs_dim == 1 ? c2 : c3)) < cs_limit;
        if (!r2_ok || c1 >= x0 || c2 >= x1 || c3 >= x2) return -1;
        const long base = mi == 0 ? bbase0 : p.bit_base[mi];
        return (base + cq[1] * bs0 + cq[2] * bs1 + cq[3] * bs2 + cq[0] + thread_bit_off) >> 5;
      };
      const long bw0 = (want_bits && simple_panels) ? bit_word(0) : -1;
      // sub-tile mode: word of panel 0 without the clipping test; each panel tests its own shifted rows
      const long sw0 = (bbase0 + cd[1] * bs0 + cd[2] * bs1 + cd[3] * bs2 + cd[0] + thread_bit_off) >> 5;
      unsigned mbits[8];
      unsigned cs_ok = 0u;   // bit q: this thread's row of panel q is stored and counts towards the column sums
      if (bitmask) {  // issued before waiting for the accumulator: the loads overlap the tile's mainloop
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          mbits[q] = 0u;
          if (q < n_panels) {
            const long wi = bit_word(q);
            if (wi >= 0) {
              mbits[q] = __ldg(bits_in + wi);
              cs_ok |= in_limit ? (1u << q) : 0u;
            }
          }
        }
      }
      mbar_wait_t(tfull0 + 8 * acc, aph, timed, w_tfull);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * acc_cols);
      int crun[5], q1 = 0;          // running panel coordinates of the plain / two-level panel walk (no division per panel)
#pragma unroll
      for (int d = 0; d < 5; ++d) crun[d] = cd[d];
      int col0 = sub_panels > 0 ? 0 : n_tile * tile_bn, subq = 0;   // first channel of the current panel (bias index / column guard)
      for (int q = 0; q < n_panels; ++q, ++pc) {
        const uint32_t buf = my_staging_u32 + (uint32_t)bi * 16384u;
        int cq_st[5];   // store coordinates of this panel, formed before the barriers so the leader only has to issue
        int mi_st = 0;
        if (cpm > 0) {
          const int col = n_tile * tile_bn + q * 32;
#pragma unroll
          for (int d = 0; d < 5; ++d) cq_st[d] = cd[d];
          cq_st[0] = d_off0 + (col & (cpm - 1));
          mi_st = col >> cpm_sh;
        } else if constexpr (SLAB) {
#pragma unroll
          for (int d = 0; d < 5; ++d) cq_st[d] = cd[d];
          cq_st[0] += p.panel_tab0[q];
          cq_st[1] += p.panel_tab1[q];
        } else {
#pragma unroll
          for (int d = 0; d < 5; ++d) cq_st[d] = crun[d];
          if (d_period > 0 && ++q1 == d_period) {
            q1 = 0;
#pragma unroll
            for (int d = 0; d < 5; ++d) crun[d] += dwrap[d];
          } else {
#pragma unroll
            for (int d = 0; d < 5; ++d) crun[d] += dpan[d];
          }
        }
        const long long c0 = timed ? clock64() : 0;
        if (leader()) {
          if (two_groups) {
            // the store that last used this staging buffer has drained it (ring of nbuf buffers per group)
            if (nbuf == 1) tma_wait_read<0>();
            else tma_wait_read<1>();
          } else {
            if (pc >= 2) tma_wait_read<1>();  // the store of panel pc-2 has drained its staging buffer
            if (masked) prefetch_one();       // panel pc+nbuf-2 -> the buffer the store of panel pc-2 just released
          }
        }
        epi_bar_sync(grp);
        const long long c1 = timed ? clock64() : 0;
        float v[32];
        tmem_ld32(tacc + (uint32_t)(q * 32), v);
        const long long c2 = timed ? clock64() : 0;
        if (q == n_panels - 1) {  // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CTA2) mbar_arrive_cluster(mapa_rank(tempty0 + 8 * acc, 0));   // the leader issues the pair's MMAs
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8 * acc) : "memory");
          }
        }
        // (col0: first channel of this panel; sub-tile mode repeats the channel panels per sub-tile)
        unsigned bits_w = 0u, bits4[4] = {0u, 0u, 0u, 0u};
        long bits_wi = -1;
        if (epi == EPI_BIAS_LRELU || epi == EPI_BIAS) {
          if (bias_smem) {  // bias staged in shared memory: 8 broadcast LDS.128 instead of 32 global loads
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = lds128(s_bias_u32 + (uint32_t)(col0 + 4 * j) * 4u);
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += (col0 + j < n_total) ? __ldg(p.bias + col0 + j) : 0.f;
          }
          if (epi == EPI_BIAS_LRELU) {
            if (want_bits) {
              // LeakyReLU and its derivative bit from ONE comparison per element: predicated multiply for the negative
              // side, predicated OR of the element's bit for the positive side (bit = output > 0 = pre-activation > 0
              // for a positive slope).  Four independent accumulators keep the OR chains 8 deep.
#pragma unroll
              for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const int i = 8 * c + j;
                  if (v[i] > 0.f) bits4[c] |= 1u << i;
                  else v[i] *= slope;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gc::leaky(v[j], slope);
            }
          }
        }
        if (want_bits) {
          unsigned w = (bits4[0] | bits4[1]) | (bits4[2] | bits4[3]);
          if (col0 + 32 > n_total) w &= (1u << (n_total - col0)) - 1u;   // partial last panel
          bits_w = w;
          if (simple_panels) {
            bits_wi = bw0 >= 0 ? bw0 + q : -1;
          } else if (sub_regular) {
            const bool ok = r2_ok && cd[1] + q * sp1 + r0 < ext0[0] && cd[2] + q * sp2 + r1 < ext0[1] && cd[3] + q * sp3 + r2 < ext0[2];
            bits_wi = ok ? sw0 + q * sub_words : -1;
          } else {
            bits_wi = bit_word(q);
          }
        }
        if (bitmask) {
          unsigned w = 0u;
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) w = (qq == q) ? mbits[qq] : w;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (!(w & (1u << j))) v[j] *= slope;   // one bit test + predicated multiply per element
        }
        if (masked) {
          mbar_wait(aux0 + 8 * bi, bph);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < nchunk) {
              const float4 m = lds128(buf + soff[j]);
              v[4 * j + 0] *= m.x > 0.f ? 1.f : slope;
              v[4 * j + 1] *= m.y > 0.f ? 1.f : slope;
              v[4 * j + 2] *= m.z > 0.f ? 1.f : slope;
              v[4 * j + 3] *= m.w > 0.f ? 1.f : slope;
            }
          }
        }
        const long long c2a = timed ? clock64() : 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < nchunk && row_live) sts128(buf + soff[j], v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        fence_async_smem();
        const long long c2b = timed ? clock64() : 0;
        epi_bar_sync(grp);
        const long long c3 = timed ? clock64() : 0;
        // the LeakyReLU' bit word goes out after the proxy fence so the fence never waits on a global store
        if (bits_wi >= 0) bits_out[bits_wi] = bits_w;
        if (leader()) {
          tma_store_5d(&p.mapD[mi_st], buf, cq_st);
          tma_commit();
        }
        if (want_colsum) {
          // The tile is complete in the staging buffer: warp wq sums column `lane` over its own 32 rows (the ballot is the
          // stored-and-counted flag of exactly those rows).  One 128-byte row per request: conflict-free with the swizzle.
          const unsigned rows_ok = __ballot_sync(0xffffffffu, (cs_ok >> q) & 1u);
          float cs = 0.f;
          const uint32_t rbase = buf + (uint32_t)(wq * 32) * 128u, cchunk = (uint32_t)(lane >> 2), cword = (uint32_t)(lane & 3) * 4u;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            if ((rows_ok >> rr) & 1u) {
              float t;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(rbase + (uint32_t)rr * 128u + ((cchunk ^ (uint32_t)(rr & 7)) << 4) + cword) : "memory");
              cs += t;
            }
          }
          if (rows_ok) atomicAdd(s_colsum + ((col0 + lane) & cs_mask), cs);
        }
        if (timed) { e_free += c1 - c0; e_ld += c2 - c1; e_math += c2a - c2; e_sts += c2b - c2a; e_bar += c3 - c2b; e_store += clock64() - c3; }
        if (++bi == nbuf) { bi = 0; bph ^= 1; }
        if (sub_panels == 0) col0 += 32;
        else if (sub_panels > 1) { col0 += 32; if (++subq == sub_panels) { subq = 0; col0 = 0; } }
      }
      if (two_groups) aph ^= 1;                                   // this group's stage is used by every other tile
      else if (++acc == acc_stages) { acc = 0; aph ^= 1; }
    }
    if (leader()) tma_wait_read<0>();
    if (timed && leader()) {
      p.stats[blockIdx.x * 16 + 3 + grp] = w_tfull;
      if (grp == 0) {
        p.stats[blockIdx.x * 16 + 6] = clock64() - t_epi0;
        p.stats[blockIdx.x * 16 + 8] = e_free;    // waiting for a free staging buffer + group barrier
        p.stats[blockIdx.x * 16 + 9] = e_ld;      // tcgen05.ld + wait::ld
        p.stats[blockIdx.x * 16 + 10] = e_math;   // bias / activation / bits / st.shared / proxy fence / group barrier
        p.stats[blockIdx.x * 16 + 11] = e_store;  // TMA store issue
        p.stats[blockIdx.x * 16 + 12] = pc;       // panels
        p.stats[blockIdx.x * 16 + 13] = e_sts;    // st.shared + proxy fence
        p.stats[blockIdx.x * 16 + 14] = e_bar;    // group barrier before the store
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.colsum_out != nullptr) {   // this CTA's share of the column sums
    for (int i = threadIdx.x; i <= p.colsum_mask; i += blockDim.x) {
      const float v = s_colsum[i];
      if (v != 0.f) atomicAdd(p.colsum_out + i, v);
    }
  }
  if constexpr (CTA2) cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the peer may still signal / read it
  if (warp == 8) {
    if constexpr (CTA2) {
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                   : "memory");
    } else {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                   : "memory");
    }
  }
}

}  // namespace gcu
