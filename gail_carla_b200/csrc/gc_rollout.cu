// HBM-bound kernels of the gail-carla learning hot path (sm_100a):
//   * GAE(lambda)/returns segmented reverse scan over time  (tools/storage.py:37-50)
//   * advantage statistics + normalisation                  (algo/ppo.py:47-49)
//   * fused PPO loss forward+backward on the head outputs   (tools/model.py:45-53,80-85; algo/ppo.py:80-85,104-113)
//   * Welford/Chan running mean-var merge                   (common/running_mean_std.py:10-31)
// All kernels enqueue on the caller's stream, never synchronise, and borrow their pointers.
#include <algorithm>
#include <cstdlib>

#include "gc_common.cuh"
#include "../../include/gail_carla_b200.h"

namespace {

using gc::warp_sum;

// --------------------------------------------------------------------------------------------
// GAE scan.  Storage is time-major [T(+1), N] (env index fastest, tools/storage.py:10-16).
// gae_t = delta_t + c_t * gae_{t+1} is an affine map x -> c_t x + delta_t; a mask of 0 resets the
// segment (c_t = 0).  A CTA owns NB<=32 envs (lanes run over envs => coalesced rows) and C time
// chunks of L steps.  Per super-chunk of C*L steps: (1) every thread folds its chunk into one affine
// map, (2) one warp per env does a warp-shuffle suffix scan over the C chunk maps to get each chunk's
// incoming gae, (3) every thread replays its chunk from registers with the right carry and writes
// returns (+ raw advantages and their sum / sum of squares).  One read of each input, one write.
// --------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(512) gae_scan_kernel(const float* __restrict__ rew, const float* __restrict__ val,
                                                       const float* __restrict__ mask, float* __restrict__ ret,
                                                       float* __restrict__ adv, double* __restrict__ stats, int T, int N,
                                                       float gamma, float lam, int NB, int C) {
  extern __shared__ float sm[];
  float* sA = sm;                 // [C][NB] chunk map slope
  float* sB = sA + C * NB;        // [C][NB] chunk map offset
  float* sCarry = sB + C * NB;    // [C][NB] gae entering each chunk (from later time)
  float* sIn = sCarry + C * NB;   // [NB]    gae entering the super-chunk
  __shared__ double red[64];

  const int tid = threadIdx.x;
  const int nl = tid % NB, c = tid / NB;
  const int n = blockIdx.x * NB + nl;
  const bool live = (c < C) && (n < N);
  const float gl = gamma * lam;
  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

  if (tid < NB) sIn[tid] = 0.f;
  float s1 = 0.f, s2 = 0.f;

  // register double buffering: the loads of super-chunk k+1 are issued before super-chunk k is folded / combined /
  // replayed, so HBM requests stay in flight across the two block barriers of every iteration
  float rn[L], vn[L + 1], mn[L];
  auto load_chunk = [&](int t_hi_, float (&r_)[L], float (&v_)[L + 1], float (&m_)[L]) {
    const int t0_ = t_hi_ - (C - c) * L;
#pragma unroll
    for (int i = 0; i < L; ++i) {
      const int t = t0_ + i;
      const bool ok = live && t >= 0;
      const size_t o = (size_t)(ok ? t : 0) * N + (live ? n : 0);
      r_[i] = ok ? __ldg(rew + o) : 0.f;
      v_[i] = ok ? __ldg(val + o) : 0.f;
      m_[i] = ok ? __ldg(mask + o + N) : 0.f;  // masks[t+1]
    }
    v_[L] = (live && t0_ + L >= 0) ? __ldg(val + (size_t)(t0_ + L) * N + n) : 0.f;
  };
  load_chunk(T, rn, vn, mn);

  for (int t_hi = T; t_hi > 0; t_hi -= C * L) {
    const int t0 = t_hi - (C - c) * L;  // first step of this thread's chunk (may be < 0)
    float r[L], v[L + 1], m[L];
#pragma unroll
    for (int i = 0; i < L; ++i) { r[i] = rn[i]; v[i] = vn[i]; m[i] = mn[i]; }
    v[L] = vn[L];
    if (t_hi - C * L > 0) load_chunk(t_hi - C * L, rn, vn, mn);
    float A = 1.f, B = 0.f;
    if (live) {
#pragma unroll
      for (int i = L - 1; i >= 0; --i) {
        if (t0 + i >= 0) {
          const float ci = gl * m[i];
          const float delta = r[i] + gamma * v[i + 1] * m[i] - v[i];
          B = delta + ci * B;
          A = ci * A;
        }
      }
    }
    if (c < C) { sA[c * NB + nl] = A; sB[c * NB + nl] = B; }
    __syncthreads();
    if (C <= 8) {
      // few chunks: one thread per env folds them serially (latest chunk first)
      if (tid < NB) {
        float carry = sIn[tid];
        for (int ci = C - 1; ci >= 0; --ci) {
          sCarry[ci * NB + tid] = carry;
          carry = sA[ci * NB + tid] * carry + sB[ci * NB + tid];
        }
        sIn[tid] = carry;
      }
    } else {
      // suffix scan over chunks, one warp per env, lanes over chunks (blocks of 32 chunks, high to low)
      for (int e = warp; e < NB; e += nwarps) {
        float carry = sIn[e];
        for (int cb = (C + 31) / 32 - 1; cb >= 0; --cb) {
          const int ci = cb * 32 + lane;
          float a = ci < C ? sA[ci * NB + e] : 1.f;
          float b = ci < C ? sB[ci * NB + e] : 0.f;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {  // inclusive suffix composition f_lane o f_{lane+1} o ...
            const float a2 = __shfl_down_sync(0xffffffffu, a, off);
            const float b2 = __shfl_down_sync(0xffffffffu, b, off);
            if (lane + off < 32) { b = a * b2 + b; a = a * a2; }
          }
          const float an = __shfl_down_sync(0xffffffffu, a, 1);
          const float bn = __shfl_down_sync(0xffffffffu, b, 1);
          const float mine = lane < 31 ? an * carry + bn : carry;  // gae entering chunk ci
          if (ci < C) sCarry[ci * NB + e] = mine;
          const float a0 = __shfl_sync(0xffffffffu, a, 0), b0 = __shfl_sync(0xffffffffu, b, 0);
          carry = a0 * carry + b0;
        }
        if (lane == 0) sIn[e] = carry;
      }
    }
    __syncthreads();
    if (live) {
      float x = sCarry[c * NB + nl];
#pragma unroll
      for (int i = L - 1; i >= 0; --i) {
        const int t = t0 + i;
        if (t >= 0) {
          const float ci = gl * m[i];
          const float delta = r[i] + gamma * v[i + 1] * m[i] - v[i];
          x = delta + ci * x;
          const size_t o = (size_t)t * N + n;
          ret[o] = x + v[i];
          if (adv) adv[o] = (x + v[i]) - v[i];  // returns - value_preds, as algo/ppo.py:47 forms it
          const float a_ = (x + v[i]) - v[i];
          s1 += a_;
          s2 += a_ * a_;
        }
      }
    }
    __syncthreads();
  }
  if (stats) {
    double acc[2] = {(double)s1, (double)s2};
    gc::block_sum<2>(acc, red);
    if (tid == 0) {
      atomicAdd(stats + 0, acc[0]);
      atomicAdd(stats + 1, acc[1]);
    }
  }
}

// mean / (unbiased std + 1e-5) from {sum, sumsq, count} (algo/ppo.py:48-49)
__device__ __forceinline__ void adv_moments(const double* __restrict__ stats, float& mean, float& inv) {
  const double s = stats[0], q = stats[1], cnt = stats[2];
  const double mu = s / cnt;
  double var = (q - s * mu) / (cnt - 1.0);
  var = var > 0.0 ? var : 0.0;
  mean = (float)mu;
  inv = 1.0f / ((float)sqrt(var) + 1e-5f);
}

__global__ void adv_normalize_kernel(const float* __restrict__ ret, const float* __restrict__ val,
                                     const double* __restrict__ stats, float* __restrict__ out, long n) {
  float mean, inv;
  adv_moments(stats, mean, inv);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    out[i] = ((ret[i] - val[i]) - mean) * inv;
}

__global__ void adv_stats_kernel(const float* __restrict__ ret, const float* __restrict__ val, double* __restrict__ stats,
                                 long n) {
  __shared__ double red[64];
  double acc[2] = {0.0, 0.0};
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float a = ret[i] - val[i];
    acc[0] += a;
    acc[1] += (double)a * a;
  }
  gc::block_sum<2>(acc, red);
  if (threadIdx.x == 0) {
    atomicAdd(stats + 0, acc[0]);
    atomicAdd(stats + 1, acc[1]);
  }
}

// --------------------------------------------------------------------------------------------
// Fused policy-head tail + PPO loss, forward and backward, one thread per sample.
// head[b] = {value, mu0_raw, mu1_raw, pad} (row pitch 4 floats).
// mode 0 (PPO, algo/ppo.py:80-85,104-111): writes d(vcoef*value_loss + w_act*action_loss)/d head.
// mode 1 (BC,  algo/ppo.py:88-99):          writes d(w_act * -mean(logp))/d head.
// mode 2 (forward only, tools/model.py:45-53 / :25-36): writes value/logp, no gradient.
// torch's min/max split the gradient evenly on ties and clamp passes it on the closed interval; so do we.
// acc[0..3] += {sum value-loss terms, sum -min(surr), sum -logp, n}
// --------------------------------------------------------------------------------------------
struct HeadTail {
  float v, mu0, mu1, logp, dmu0, dmu1;  // dmu = d logp / d mu
};

__device__ __forceinline__ HeadTail head_tail(const float4 h, const float a0, const float a1, const float ls0,
                                              const float ls1, const int activation) {
  HeadTail o;
  o.v = h.x;
  o.mu0 = activation ? tanhf(h.y) : h.y;
  o.mu1 = activation ? 1.f / (1.f + expf(-h.z)) : h.z;
  const float var0 = expf(ls0) * expf(ls0), var1 = expf(ls1) * expf(ls1);
  const float kHalfLog2Pi = 0.91893853320467274178f;
  const float d0 = a0 - o.mu0, d1 = a1 - o.mu1;
  o.logp = (-(d0 * d0) / (2.f * var0) - ls0 - kHalfLog2Pi) + (-(d1 * d1) / (2.f * var1) - ls1 - kHalfLog2Pi);
  o.dmu0 = d0 / var0;
  o.dmu1 = d1 / var1;
  return o;
}

// one sample of the fused head tail + PPO / BC loss: returns d loss / d head row, adds this sample's loss terms to part[]
struct PpoConsts {
  float ls0, ls1, clip, vcoef, w_act, inv_B, mean, inv;
  int activation, mode, clipped_value;
};
__device__ __forceinline__ float4 ppo_sample(const PpoConsts& c, const float4 h, const float2 a, const float olp, const float vo,
                                             const float R, const float adv_or_nan, const bool has_adv, double (&part)[3],
                                             float& out_v, float& out_lp) {
  const HeadTail t = head_tail(h, a.x, a.y, c.ls0, c.ls1, c.activation);
  out_v = t.v;
  out_lp = t.logp;
  float dlogp = 0.f, dv = 0.f;
  if (c.mode == 0) {
    const float A = has_adv ? adv_or_nan : ((R - vo) - c.mean) * c.inv;
    const float ratio = expf(t.logp - olp);
    const float lo = 1.f - c.clip, hi = 1.f + c.clip;
    const float rc = fminf(fmaxf(ratio, lo), hi);
    const float sa = ratio * A, sb = rc * A;
    const float in_rng = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
    float dr;  // d min(sa,sb) / d ratio
    if (sa < sb) dr = A;
    else if (sa > sb) dr = A * in_rng;
    else dr = A * (0.5f + 0.5f * in_rng);
    dlogp = -c.w_act * c.inv_B * dr * ratio;
    part[1] += (double)(-fminf(sa, sb));
    const float l1 = (t.v - R) * (t.v - R);
    if (c.clipped_value) {  // clipped value loss (algo/ppo.py:104-111)
      const float dvv = t.v - vo;
      const float vc = vo + fminf(fmaxf(dvv, -c.clip), c.clip);
      const float l2 = (vc - R) * (vc - R);
      const float vin = (dvv >= -c.clip && dvv <= c.clip) ? 1.f : 0.f;
      float g;  // d max(l1,l2) / d v
      if (l1 > l2) g = 2.f * (t.v - R);
      else if (l1 < l2) g = 2.f * (vc - R) * vin;
      else g = (t.v - R) + (vc - R) * vin;
      dv = c.vcoef * 0.5f * c.inv_B * g;
      part[0] += (double)(0.5f * fmaxf(l1, l2));
    } else {  // 0.5 * (return - value)^2 (algo/ppo.py:112-113)
      dv = c.vcoef * c.inv_B * (t.v - R);
      part[0] += (double)(0.5f * l1);
    }
  } else if (c.mode == 1) {
    dlogp = -c.w_act * c.inv_B;
    part[2] += (double)(-t.logp);
  }
  float4 g4;
  g4.x = dv;
  g4.y = dlogp * t.dmu0 * (c.activation ? (1.f - t.mu0 * t.mu0) : 1.f);
  g4.z = dlogp * t.dmu1 * (c.activation ? (t.mu1 * (1.f - t.mu1)) : 1.f);
  g4.w = 0.f;
  return g4;
}

// Each thread owns 4 consecutive samples: the five per-sample scalar streams are read as one 16-byte load each and the
// head / action / gradient rows as 4 + 2 + 4 of them, i.e. 13 x 128-bit requests per 4 samples instead of 28 narrower
// ones (the kernel is a pure 52 B/sample stream; request count and bytes in flight are what bound it).
__global__ void __launch_bounds__(256) ppo_loss_kernel(const float4* __restrict__ head, const float2* __restrict__ action,
                                                       const float* __restrict__ old_logp, const float* __restrict__ v_old,
                                                       const float* __restrict__ ret, const float* __restrict__ adv_in,
                                                       const double* __restrict__ stats, float4* __restrict__ d_head,
                                                       float* __restrict__ out_value, float* __restrict__ out_logp,
                                                       double* __restrict__ acc, int B, float ls0, float ls1, int activation,
                                                       float clip, float vcoef, float w_act, float inv_B, int mode,
                                                       int clipped_value) {
  __shared__ double red[32 * 3];
  PpoConsts c{ls0, ls1, clip, vcoef, w_act, inv_B, 0.f, 1.f, activation, mode, clipped_value};
  if (mode == 0 && adv_in == nullptr) adv_moments(stats, c.mean, c.inv);
  double part[3] = {0.0, 0.0, 0.0};
  const bool has_adv = adv_in != nullptr;
  const int B4 = B >> 2;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B4; i += gridDim.x * blockDim.x) {
    const float4 h0 = head[4 * i], h1 = head[4 * i + 1], h2 = head[4 * i + 2], h3 = head[4 * i + 3];
    const float4 a01 = reinterpret_cast<const float4*>(action)[2 * i], a23 = reinterpret_cast<const float4*>(action)[2 * i + 1];
    const float4 olp = mode == 0 ? reinterpret_cast<const float4*>(old_logp)[i] : z4;
    const float4 vo = mode == 0 ? reinterpret_cast<const float4*>(v_old)[i] : z4;
    const float4 R = mode == 0 ? reinterpret_cast<const float4*>(ret)[i] : z4;
    const float4 ad = (mode == 0 && has_adv) ? reinterpret_cast<const float4*>(adv_in)[i] : z4;
    float4 ov, ol;
    const float4 g0 = ppo_sample(c, h0, make_float2(a01.x, a01.y), olp.x, vo.x, R.x, ad.x, has_adv, part, ov.x, ol.x);
    const float4 g1 = ppo_sample(c, h1, make_float2(a01.z, a01.w), olp.y, vo.y, R.y, ad.y, has_adv, part, ov.y, ol.y);
    const float4 g2 = ppo_sample(c, h2, make_float2(a23.x, a23.y), olp.z, vo.z, R.z, ad.z, has_adv, part, ov.z, ol.z);
    const float4 g3 = ppo_sample(c, h3, make_float2(a23.z, a23.w), olp.w, vo.w, R.w, ad.w, has_adv, part, ov.w, ol.w);
    if (out_value) reinterpret_cast<float4*>(out_value)[i] = ov;
    if (out_logp) reinterpret_cast<float4*>(out_logp)[i] = ol;
    if (mode != 2) {
      d_head[4 * i] = g0; d_head[4 * i + 1] = g1; d_head[4 * i + 2] = g2; d_head[4 * i + 3] = g3;
    }
  }
  // ragged tail (B % 4 samples), one thread each
  for (int b = 4 * B4 + blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    float ov, ol;
    const float4 g = ppo_sample(c, head[b], action[b], mode == 0 ? old_logp[b] : 0.f, mode == 0 ? v_old[b] : 0.f,
                                mode == 0 ? ret[b] : 0.f, (mode == 0 && has_adv) ? adv_in[b] : 0.f, has_adv, part, ov, ol);
    if (out_value) out_value[b] = ov;
    if (out_logp) out_logp[b] = ol;
    if (mode != 2) d_head[b] = g;
  }
  if (acc && mode != 2) {
    gc::block_sum<3>(part, red);
    if (threadIdx.x == 0) {
      atomicAdd(acc + 0, part[0]);
      atomicAdd(acc + 1, part[1]);
      atomicAdd(acc + 2, part[2]);
    }
  }
}

// act(): action = mu (+ sigma*noise), logp of that action (tools/model.py:25-36)
__global__ void policy_act_kernel(const float4* __restrict__ head, const float2* __restrict__ noise, float* __restrict__ value,
                                  float2* __restrict__ action, float* __restrict__ logp, int B, float ls0, float ls1,
                                  int activation) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float4 h = head[b];
  HeadTail t = head_tail(h, 0.f, 0.f, ls0, ls1, activation);
  float2 a = make_float2(t.mu0, t.mu1);
  if (noise) {
    a.x += expf(ls0) * noise[b].x;
    a.y += expf(ls1) * noise[b].y;
  }
  t = head_tail(h, a.x, a.y, ls0, ls1, activation);
  value[b] = t.v;
  action[b] = a;
  logp[b] = t.logp;
}

// --------------------------------------------------------------------------------------------
// RunningMeanStd.update over a device vector (shape () statistics), float64 like the reference.
// pass 1: {sum, sumsq-centred via two-level} -> we use sum and sum of squares in double, then Chan merge.
// --------------------------------------------------------------------------------------------
__global__ void moments_kernel(const float* __restrict__ x, long n, double* __restrict__ acc) {
  __shared__ double red[64];
  double part[2] = {0.0, 0.0};
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double v = x[i];
    part[0] += v;
    part[1] += v * v;
  }
  gc::block_sum<2>(part, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc + 0, part[0]);
    atomicAdd(acc + 1, part[1]);
  }
}

// state = {mean, var, count}; acc = {sum, sumsq}; Chan et al. merge (common/running_mean_std.py:20-31)
__global__ void welford_merge_kernel(double* __restrict__ state, const double* __restrict__ acc, double batch_count) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double bm = acc[0] / batch_count;
  double bv = acc[1] / batch_count - bm * bm;  // population variance (np.var)
  bv = bv > 0.0 ? bv : 0.0;
  const double mean = state[0], var = state[1], count = state[2];
  const double delta = bm - mean, tot = count + batch_count;
  const double m2 = var * count + bv * batch_count + delta * delta * count * batch_count / tot;
  state[0] = mean + delta * batch_count / tot;
  state[1] = m2 / tot;
  state[2] = tot;
}

}  // namespace

// ================================= C ABI =================================
extern "C" {

const char* gc_last_error_string(void) { return gc::last_error().c_str(); }

int gc_abi_version(void) { return GC_ABI_VERSION; }

#ifndef GC_BUILD_DIGEST
#define GC_BUILD_DIGEST "unknown"
#endif
const char* gc_build_digest(void) { return GC_BUILD_DIGEST; }

int gc_gae_returns(const float* gail_rewards, const float* value_preds, const float* masks, float* returns, float* adv_raw,
                   double* stats, int T, int N, float gamma, float gae_lambda, void* stream) {
  GC_REQUIRE(T > 0 && N > 0, "gc_gae_returns: T=%d N=%d must be positive", T, N);
  GC_REQUIRE(gail_rewards && value_preds && masks && returns, "gc_gae_returns: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  // Launch shape.  Few envs: one CTA covers them all and many time chunks run in parallel (short serial chain).
  // Many envs: wide env groups per CTA (long contiguous row segments), few chunks, enough CTAs for every SM.
  int NB = N >= 32 ? 32 : N, threads = 512, L = 16;
  if (N >= 32 * 2 * gc::kNumSMs) { NB = 32; threads = 256; L = 16; }  // measured best on B200 (profiles/r01_gae_sweep.txt)
  if (const char* cfg = getenv("GC_GAE_CFG")) {  // tuning override: "NB,threads,L"
    int a = 0, b = 0, c = 0;
    if (sscanf(cfg, "%d,%d,%d", &a, &b, &c) == 3 && a > 0 && a <= N && b >= a && b <= 512 && (c == 4 || c == 8 || c == 16)) {
      NB = a; threads = b; L = c;
    }
  }
  const int grid = (N + NB - 1) / NB;
  int C = threads / NB;
  if ((long)C * 8 >= T && L > 8) L = 8;
  if ((long)C * 4 >= T) L = 4;
  // do not spawn chunks that would be entirely before t=0
  const int c_needed = (T + L - 1) / L;
  if (C > c_needed) C = c_needed;
  threads = ((C * NB + 31) / 32) * 32;
  const size_t smem = (size_t)(3 * C * NB + NB) * sizeof(float);
  if (stats) {
    GC_CUDA_OK(cudaMemsetAsync(stats, 0, 2 * sizeof(double), st));
    const double cnt = (double)T * (double)N;
    GC_CUDA_OK(cudaMemcpyAsync(stats + 2, &cnt, sizeof(double), cudaMemcpyHostToDevice, st));
  }
  if (L == 16)
    gae_scan_kernel<16><<<grid, threads, smem, st>>>(gail_rewards, value_preds, masks, returns, adv_raw, stats, T, N, gamma, gae_lambda, NB, C);
  else if (L == 8)
    gae_scan_kernel<8><<<grid, threads, smem, st>>>(gail_rewards, value_preds, masks, returns, adv_raw, stats, T, N, gamma, gae_lambda, NB, C);
  else
    gae_scan_kernel<4><<<grid, threads, smem, st>>>(gail_rewards, value_preds, masks, returns, adv_raw, stats, T, N, gamma, gae_lambda, NB, C);
  return gc::launch_status("gae_scan_kernel");
}

int gc_adv_stats(const float* returns, const float* value_preds, double* stats, long n, void* stream) {
  GC_REQUIRE(n > 1, "gc_adv_stats: need n > 1 (unbiased std), got %ld", n);
  cudaStream_t st = (cudaStream_t)stream;
  GC_CUDA_OK(cudaMemsetAsync(stats, 0, 2 * sizeof(double), st));
  const double cnt = (double)n;
  GC_CUDA_OK(cudaMemcpyAsync(stats + 2, &cnt, sizeof(double), cudaMemcpyHostToDevice, st));
  const int grid = (int)std::min<long>((n + 255) / 256, 4L * gc::kNumSMs);
  adv_stats_kernel<<<grid, 256, 0, st>>>(returns, value_preds, stats, n);
  return gc::launch_status("adv_stats_kernel");
}

int gc_adv_normalize(const float* returns, const float* value_preds, const double* stats, float* adv_out, long n,
                     void* stream) {
  GC_REQUIRE(n > 0, "gc_adv_normalize: n=%ld", n);
  const int grid = (int)std::min<long>((n + 255) / 256, 8L * gc::kNumSMs);
  adv_normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(returns, value_preds, stats, adv_out, n);
  return gc::launch_status("adv_normalize_kernel");
}

int gc_ppo_loss_fwd_bwd(const float* head_out, const float* actions, const float* old_logp, const float* value_old,
                        const float* returns, const float* adv, const double* adv_stats, float* d_head_out,
                        float* out_value, float* out_logp, double* loss_acc, int B, float logstd0, float logstd1,
                        int activation, float clip, float value_coef, float action_weight, int mode, int clipped_value,
                        float inv_norm, void* stream) {
  GC_REQUIRE(B > 0, "gc_ppo_loss_fwd_bwd: B=%d", B);
  GC_REQUIRE(mode >= 0 && mode <= 2, "gc_ppo_loss_fwd_bwd: mode %d not in {0,1,2}", mode);
  GC_REQUIRE(head_out && actions, "gc_ppo_loss_fwd_bwd: null head/actions");
  if (mode == 0) GC_REQUIRE(old_logp && value_old && returns && (adv || adv_stats) && d_head_out,
                            "gc_ppo_loss_fwd_bwd: PPO mode needs old_logp, value_old, returns, adv|adv_stats, d_head_out");
  if (mode == 1) GC_REQUIRE(d_head_out, "gc_ppo_loss_fwd_bwd: BC mode needs d_head_out");
  // the 128-bit path needs 16-byte aligned per-sample streams (torch allocations are; sliced views may not be)
  auto al16 = [](const void* q) { return q == nullptr || ((uintptr_t)q & 15) == 0; };
  GC_REQUIRE(al16(head_out) && al16(actions) && al16(old_logp) && al16(value_old) && al16(returns) && al16(adv) &&
                 al16(d_head_out) && al16(out_value) && al16(out_logp),
             "gc_ppo_loss_fwd_bwd: per-sample arrays must be 16-byte aligned");
  const int grid = std::min((B / 4 + 255) / 256 + 1, 8 * gc::kNumSMs);
  ppo_loss_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      (const float4*)head_out, (const float2*)actions, old_logp, value_old, returns, adv, adv_stats, (float4*)d_head_out,
      out_value, out_logp, loss_acc, B, logstd0, logstd1, activation, clip, value_coef, action_weight,
      inv_norm > 0.f ? inv_norm : 1.0f / (float)B, mode, clipped_value);
  return gc::launch_status("ppo_loss_kernel");
}

int gc_policy_act(const float* head_out, const float* noise, float* value, float* action, float* logp, int B, float logstd0,
                  float logstd1, int activation, void* stream) {
  GC_REQUIRE(B > 0 && head_out && value && action && logp, "gc_policy_act: bad arguments");
  policy_act_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const float4*)head_out, (const float2*)noise, value,
                                                                    (float2*)action, logp, B, logstd0, logstd1, activation);
  return gc::launch_status("policy_act_kernel");
}

int gc_welford_merge(double* state, const float* x, long n, double* scratch2, void* stream) {
  GC_REQUIRE(n > 0 && state && x && scratch2, "gc_welford_merge: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  GC_CUDA_OK(cudaMemsetAsync(scratch2, 0, 2 * sizeof(double), st));
  const int grid = (int)std::min<long>((n + 255) / 256, 4L * gc::kNumSMs);
  moments_kernel<<<grid, 256, 0, st>>>(x, n, scratch2);
  welford_merge_kernel<<<1, 32, 0, st>>>(state, scratch2, (double)n);
  return gc::launch_status("welford_merge_kernel");
}

}  // extern "C"
