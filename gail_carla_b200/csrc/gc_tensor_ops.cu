// Data-movement, small elementwise / reduction stages, parameter layout preparation and the fused
// clip-grad-norm + Adam step of the gail-carla hot path (sm_100a).  Entry points documented in
// include/gail_carla_b200.h; all of these are HBM-bound or tiny.
#include <algorithm>

#include "gc_common.cuh"
#include "../../include/gail_carla_b200.h"

namespace {

constexpr int kObsC = 3, kObsH = 192, kObsW = 192;
constexpr int kS2dH = 96, kS2dW = 96, kS2dC = 16;
constexpr int kFeat = 25600, kPix = 100, kC4 = 256;
__constant__ float c_mean[3] = {0.485f, 0.456f, 0.406f};  // tools/model.py:154
__constant__ float c_std[3] = {0.229f, 0.224f, 0.225f};   // tools/model.py:155

inline int grid_for(long n, int threads, int per_sm) {
  return (int)std::max<long>(1, std::min<long>((n + threads - 1) / threads, (long)per_sm * gc::kNumSMs));
}

// ---- image gather: CHW fp32 storage row -> normalised space-to-depth NHWC tile -----------------------------
// One CTA per (sample, group of 4 s2d rows): stage 4 x 3 channels x 2 image rows x 192 floats in smem with float4
// loads (3 independent 16-byte loads in flight per thread), then write the 4 x 96 x 16 output floats contiguously.
constexpr int kGatherRows = 4;
// source element -> 4 consecutive pixels of one image row as floats in [0,1]
__device__ __forceinline__ float4 load_px4(const float* row, int xv) { return __ldg(reinterpret_cast<const float4*>(row) + xv); }
__device__ __forceinline__ float4 load_px4(const unsigned char* row, int xv) {
  // expert images are stored as the PNGs' uint8 (algo/wdgail.py:222-227: ToTensor = uint8 / 255 in fp32)
  const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(row) + xv);
  return make_float4((float)u.x / 255.f, (float)u.y / 255.f, (float)u.z / 255.f, (float)u.w / 255.f);
}

template <typename SrcT>
__global__ void __launch_bounds__(384) gather_obs_s2d_kernel(const SrcT* __restrict__ src, const long long* __restrict__ idx,
                                                             float* __restrict__ out) {
  __shared__ __align__(16) float tile[kGatherRows][kObsC][2][kObsW + 4];
  const int b = blockIdx.y, Y0 = blockIdx.x * kGatherRows;
  const long row = idx ? idx[b] : b;
  const SrcT* img = src + row * (long)(kObsC * kObsH * kObsW);
  constexpr int kRowV = kObsW / 4;  // 4-pixel groups per image row
  for (int i = threadIdx.x; i < kGatherRows * kObsC * 2 * kRowV; i += blockDim.x) {
    const int xv = i % kRowV, dy = (i / kRowV) % 2, c = (i / (2 * kRowV)) % kObsC, yy = i / (2 * kRowV * kObsC);
    const float4 v = load_px4(img + ((long)c * kObsH + 2 * (Y0 + yy) + dy) * kObsW, xv);
    const float m = c_mean[c], sd = c_std[c];
    *reinterpret_cast<float4*>(&tile[yy][c][dy][4 * xv]) =
        make_float4((v.x - m) / sd, (v.y - m) / sd, (v.z - m) / sd, (v.w - m) / sd);
  }
  __syncthreads();
  float4* o = reinterpret_cast<float4*>(out + ((long)b * kS2dH + Y0) * (kS2dW * kS2dC));
  for (int i = threadIdx.x; i < kGatherRows * kS2dW * 4; i += blockDim.x) {  // one float4 = (c0,c1,c2,1) of one (Y,X,dy,dx)
    const int yy = i / (kS2dW * 4), j = i % (kS2dW * 4);
    const int X = j >> 2, dy = (j >> 1) & 1, dx = j & 1;
    const int x = 2 * X + dx;
    o[i] = make_float4(tile[yy][0][dy][x], tile[yy][1][dy][x], tile[yy][2][dy][x], 1.f);  // pad channel = 1: conv1 bias-grad column
  }
}

// uint8 source (byte store of the rollout, expert table): one CTA per (sample, 8 s2d rows).  The two IEEE divisions per
// element of ToTensor + Normalize ((u/255 - mean)/std, tools/model.py:154-161) cost more issue slots than the copy itself,
// so they are taken once per CTA into a 3 x 256 look-up table (bit-identical by construction: the same expression on
// the same 256 inputs); the 48 source rows are staged as bytes with 16-byte loads and every thread then writes whole
// (c0,c1,c2,1) float4 pixels of the space-to-depth image - 2.4 KB in, 36.9 KB out per CTA, fully coalesced.
constexpr int kG8Rows = 8;
__global__ void __launch_bounds__(256) gather_obs_u8_s2d_kernel(const unsigned char* __restrict__ src,
                                                                const long long* __restrict__ idx, float* __restrict__ out) {
  __shared__ float lut[kObsC][256];
  __shared__ __align__(16) unsigned char tile[kObsC][2 * kG8Rows][kObsW + 16];
  const int b = blockIdx.y, Y0 = blockIdx.x * kG8Rows;
  const long row = idx ? idx[b] : b;
  const unsigned char* img = src + row * (long)(kObsC * kObsH * kObsW);
  constexpr int kRowV = kObsW / 16;  // 16-byte groups per image row
  for (int i = threadIdx.x; i < kObsC * 2 * kG8Rows * kRowV; i += blockDim.x) {
    const int v = i % kRowV, r = (i / kRowV) % (2 * kG8Rows), c = i / (kRowV * 2 * kG8Rows);
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(img + ((long)c * kObsH + 2 * Y0 + r) * kObsW) + v);
    *reinterpret_cast<uint4*>(&tile[c][r][16 * v]) = q;
  }
  for (int i = threadIdx.x; i < kObsC * 256; i += blockDim.x) {
    const int c = i >> 8, u = i & 255;
    lut[c][u] = ((float)u / 255.f - c_mean[c]) / c_std[c];
  }
  __syncthreads();
  float4* o = reinterpret_cast<float4*>(out + ((long)b * kS2dH + Y0) * (kS2dW * kS2dC));
  for (int i = threadIdx.x; i < kG8Rows * kS2dW * 4; i += blockDim.x) {  // one float4 = (c0,c1,c2,1) of one (Y,X,dy,dx)
    const int yy = i / (kS2dW * 4), j = i % (kS2dW * 4);
    const int X = j >> 2, dy = (j >> 1) & 1, dx = j & 1;
    const int x = 2 * X + dx, r = 2 * yy + dy;
    o[i] = make_float4(lut[0][tile[0][r][x]], lut[1][tile[1][r][x]], lut[2][tile[2][r][x]], 1.f);
  }
}

// Critic minibatch in one pass (algo/wdgail.py:66-80,116,121): expert image b, policy image b and their mix-up
// alpha*e + (1-alpha)*p, all three written as normalised space-to-depth rows [0,B) | [B,2B) | [2B,3B) of `out`.  Same
// staging as gather_obs_u8_s2d_kernel; the mix is formed from the two normalised pixels in registers (the same two
// products and one sum gc_mixup computes from the stored images), so the separate mix-up pass - 2 reads and 1 write of the
// fp32 images - disappears.
__global__ void __launch_bounds__(256) gather_pair_mix_u8_s2d_kernel(const unsigned char* __restrict__ src_e,
                                                                     const long long* __restrict__ idx_e,
                                                                     const unsigned char* __restrict__ src_p,
                                                                     const long long* __restrict__ idx_p,
                                                                     const float* __restrict__ alpha, float* __restrict__ out, int B) {
  __shared__ float lut[kObsC][256];
  __shared__ __align__(16) unsigned char tile[2][kObsC][2 * kG8Rows][kObsW + 16];
  const int b = blockIdx.y, Y0 = blockIdx.x * kG8Rows;
  const long row_e = idx_e ? idx_e[b] : b, row_p = idx_p ? idx_p[b] : b;
  const unsigned char* img[2] = {src_e + row_e * (long)(kObsC * kObsH * kObsW), src_p + row_p * (long)(kObsC * kObsH * kObsW)};
  constexpr int kRowV = kObsW / 16;
  constexpr int kPer = kObsC * 2 * kG8Rows * kRowV;
  for (int i = threadIdx.x; i < 2 * kPer; i += blockDim.x) {
    const int which = i / kPer, k = i % kPer;
    const int v = k % kRowV, r = (k / kRowV) % (2 * kG8Rows), c = k / (kRowV * 2 * kG8Rows);
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(img[which] + ((long)c * kObsH + 2 * Y0 + r) * kObsW) + v);
    *reinterpret_cast<uint4*>(&tile[which][c][r][16 * v]) = q;
  }
  for (int i = threadIdx.x; i < kObsC * 256; i += blockDim.x) {
    const int c = i >> 8, u = i & 255;
    lut[c][u] = ((float)u / 255.f - c_mean[c]) / c_std[c];
  }
  __syncthreads();
  const float a = alpha[b], na = 1.f - a;
  const long per_row = kS2dW * kS2dC, per_img = (long)kS2dH * per_row;
  float4* oe = reinterpret_cast<float4*>(out + (long)b * per_img + Y0 * per_row);
  float4* op = reinterpret_cast<float4*>(out + ((long)B + b) * per_img + Y0 * per_row);
  float4* om = reinterpret_cast<float4*>(out + (2L * B + b) * per_img + Y0 * per_row);
  for (int i = threadIdx.x; i < kG8Rows * kS2dW * 4; i += blockDim.x) {
    const int yy = i / (kS2dW * 4), j = i % (kS2dW * 4);
    const int X = j >> 2, dy = (j >> 1) & 1, dx = j & 1;
    const int x = 2 * X + dx, r = 2 * yy + dy;
    const float e0 = lut[0][tile[0][0][r][x]], e1 = lut[1][tile[0][1][r][x]], e2 = lut[2][tile[0][2][r][x]];
    const float p0 = lut[0][tile[1][0][r][x]], p1 = lut[1][tile[1][1][r][x]], p2 = lut[2][tile[1][2][r][x]];
    oe[i] = make_float4(e0, e1, e2, 1.f);
    op[i] = make_float4(p0, p1, p2, 1.f);
    om[i] = make_float4(a * e0 + na * p0, a * e1 + na * p1, a * e2 + na * p2, a * 1.f + na * 1.f);
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx, float* __restrict__ out,
                                   int B, int width, long ldo) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)B * width) return;
  const int b = (int)(i / width), j = (int)(i % width);
  const long row = idx ? idx[b] : b;
  out[b * ldo + j] = src[row * width + j];
}

__global__ void mixup_kernel(const float4* __restrict__ xe, const float4* __restrict__ xp, const float* __restrict__ alpha,
                             float4* __restrict__ out, long per4) {
  const int b = blockIdx.y;
  const float a = alpha[b], na = 1.f - a;
  const float4* e = xe + b * per4;
  const float4* q = xp + b * per4;
  float4* o = out + b * per4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < per4; i += (long)gridDim.x * blockDim.x) {
    const float4 u = e[i], v = q[i];
    o[i] = make_float4(a * u.x + na * v.x, a * u.y + na * v.y, a * u.z + na * v.z, a * u.w + na * v.w);
  }
}

__device__ __forceinline__ float4 mixed_metrics(const float* m, const float* m2, const float* alpha, int b) {
  float4 v = *reinterpret_cast<const float4*>(m + 4 * b);
  if (m2) {
    const float a = alpha[b], na = 1.f - a;
    const float4 w = *reinterpret_cast<const float4*>(m2 + 4 * b);
    v = make_float4(a * v.x + na * w.x, a * v.y + na * w.y, a * v.z + na * w.z, a * v.w + na * w.w);
  }
  return v;
}

__global__ void metrics_features_kernel(const float* __restrict__ m, const float* __restrict__ m2, const float* __restrict__ act,
                                        const float* __restrict__ act2, const float* __restrict__ alpha,
                                        const float* __restrict__ emb, float* __restrict__ out, long ldo, int pad, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float4 v = mixed_metrics(m, m2, alpha, b);
  float* o = out + b * ldo;
  const float r = sqrtf(v.x * v.x + v.y * v.y);
  o[0] = 1000.f * v.x;
  o[1] = 1000.f * v.y;
  o[2] = 1000.f * r;
  o[3] = 0.3f * atan2f(v.y, v.x);
  o[4] = 0.1f * v.z;
  int c = (int)v.w;  // .long() truncates toward zero (tools/model.py:203-204)
  c = c < 0 ? 0 : (c > 9 ? 9 : c);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[5 + j] = emb[c * 8 + j];
  int n = 13;
  if (act) {
    float a0 = act[2 * b], a1 = act[2 * b + 1];
    if (act2) {
      const float a = alpha[b], na = 1.f - a;
      a0 = a * a0 + na * act2[2 * b];
      a1 = a * a1 + na * act2[2 * b + 1];
    }
    o[13] = a0;
    o[14] = a1;
    n = 15;
  }
  for (int j = n; j < pad; ++j) o[j] = 0.f;
}

__global__ void metrics_features_bwd_kernel(const float* __restrict__ m, const float* __restrict__ m2,
                                            const float* __restrict__ alpha, const float* __restrict__ dfeat, long ldf,
                                            float* __restrict__ demb, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 8) return;
  const int b = i >> 3, j = i & 7;
  const float4 v = mixed_metrics(m, m2, alpha, b);
  int c = (int)v.w;
  c = c < 0 ? 0 : (c > 9 ? 9 : c);
  atomicAdd(demb + c * 8 + j, dfeat[b * ldf + 5 + j]);
}

// ---- tiny-N linear layers (head.2: 256->3, trunk.2: 100->1) --------------------------------------------------
__global__ void small_linear_fwd_kernel(const float* __restrict__ x, long ldx, const float* __restrict__ w,
                                        const float* __restrict__ bias, float* __restrict__ y, long ldy, int B, int N, int K) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = lane; k < K; k += 32) {
    const float xv = x[b * ldx + k];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < N) acc[j] += xv * __ldg(w + j * K + k);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float s = gc::warp_sum(acc[j]);
    if (lane == 0 && j < N) y[b * ldy + j] = s + (bias ? bias[j] : 0.f);
  }
}

__global__ void small_linear_dx_kernel(const float* __restrict__ x, long ldx, const float* __restrict__ w,
                                       const float* __restrict__ dy, long lddy, float* __restrict__ dx, long lddx, int B, int N,
                                       int K, float slope) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)B * K) return;
  const int b = (int)(i / K), k = (int)(i % K);
  float s = 0.f;
  for (int j = 0; j < N; ++j) s += dy[b * lddy + j] * __ldg(w + j * K + k);
  if (slope >= 0.f) s *= (x[b * ldx + k] > 0.f ? 1.f : slope);
  dx[b * lddx + k] = s;
}

// dw[j][k] += sum_b dy[b][j]*x[b][k]; grid (k-chunks, row-chunks); db handled by blockIdx.x == 0
__global__ void small_linear_dw_kernel(const float* __restrict__ x, long ldx, const float* __restrict__ dy, long lddy,
                                       float* __restrict__ dw, float* __restrict__ db, int B, int N, int K, int rows_per) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int b0 = blockIdx.y * rows_per, b1 = min(B, b0 + rows_per);
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, accb[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b = b0; b < b1; ++b) {
    const float xv = k < K ? x[b * ldx + k] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < N) {
        const float d = dy[b * lddy + j];
        acc[j] += d * xv;
        accb[j] += d;
      }
  }
  if (k < K)
    for (int j = 0; j < N; ++j) atomicAdd(dw + j * K + k, acc[j]);
  if (db && k == 0)
    for (int j = 0; j < N; ++j) atomicAdd(db + j, accb[j]);
}

__global__ void disc_loss_seed_kernel(const float* __restrict__ d, float* __restrict__ dd, double* __restrict__ acc, int B,
                                      float invB) {
  __shared__ double red[32 * 4];
  double part[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * B; i += gridDim.x * blockDim.x) {
    const float v = d[i];
    if (i < 2 * B) {
      const float t = tanhf(v);
      const float g = (1.f - t * t) * invB;
      if (i < B) { dd[i] = -g; part[0] += v; part[2] += t; }
      else { dd[i] = g; part[1] += v; part[3] += t; }
    } else {
      dd[i] = 1.f;
    }
  }
  gc::block_sum<4>(part, red);
  if (threadIdx.x == 0)
    for (int j = 0; j < 4; ++j) atomicAdd(acc + j, part[j]);
}

// one CTA per sample: ||g_raw|| then u = lambda*2*(||g||-1)/(B*||g||) * s_c^2 * g
__global__ void __launch_bounds__(512) grad_penalty_kernel(const float4* __restrict__ g, float4* __restrict__ u,
                                                           double* __restrict__ acc, float inv_norm, long per4, float lambda_,
                                                           float s0, float s1, float s2) {
  __shared__ double red[32];
  __shared__ float s_coef;
  const int b = blockIdx.x;
  const float4* gb = g + b * per4;
  float4* ub = u + b * per4;
  float ss = 0.f;
  double part[1];
  double tot = 0.0;
  for (long i = threadIdx.x; i < per4; i += blockDim.x) {
    const float4 v = gb[i];
    const float a = v.x * s0, c = v.y * s1, d = v.z * s2;
    ss += a * a + c * c + d * d;
    if ((i & 1023) == 1023) { tot += ss; ss = 0.f; }
  }
  part[0] = tot + ss;
  gc::block_sum<1>(part, red);
  if (threadIdx.x == 0) {
    const float nrm = (float)sqrt(part[0]);
    const float diff = nrm - 1.f;
    atomicAdd(acc, (double)diff * diff);
    s_coef = nrm > 0.f ? lambda_ * 2.f * diff * inv_norm / nrm : 0.f;
  }
  __syncthreads();
  const float k0 = s_coef * s0 * s0, k1 = s_coef * s1 * s1, k2 = s_coef * s2 * s2;
  for (long i = threadIdx.x; i < per4; i += blockDim.x) {
    const float4 v = gb[i];
    ub[i] = make_float4(k0 * v.x, k1 * v.y, k2 * v.z, 0.f);
  }
}

__global__ void reward_kernel(const float* __restrict__ d, float* __restrict__ r, long n) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = 1.f / (1.f + expf(-d[i]));  // torch.sigmoid
  r[i] = -logf(1.f - s);                      // -(1 - s).log()  (algo/wdgail.py:185-186)
}

// out[c] += column sums; block = 32 x 8 (columns x row lanes), grid (col chunks, row chunks)
__global__ void colsum_kernel(const float* __restrict__ x, long ld, long rows, int C, float* __restrict__ out, long rows_per) {
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long r0 = blockIdx.y * rows_per, r1 = min(rows, r0 + rows_per);
  float acc = 0.f;
  if (c < C)
    for (long r = r0 + threadIdx.y; r < r1; r += 8) acc += x[r * ld + c];
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += sm[j][threadIdx.x];
    atomicAdd(out + c, s);
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, long M, int N, long ldp,
                                     const float* __restrict__ bias, const float* __restrict__ mask, long ldm,
                                     float* __restrict__ out, long ldo, int epi, float slope) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const long m = i / N;
  const int n = (int)(i % N);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[((long)z * M + m) * ldp + n];
  if (epi == 1 || epi == 2) s += bias[n];
  if (epi == 1) s = gc::leaky(s, slope);
  if (epi == 3) s *= (mask[m * ldm + n] > 0.f ? 1.f : slope);
  out[m * ldo + n] = s;
}

// ---- parameter layout preparation ----------------------------------------------------------------------------
// generic conv (layers 2-4): w[n][c][ky][kx]  ->  wf[n][ky][kx][c],  wd[cls][c][a][b'][n]
__global__ void prep_conv_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wd, int Cout, int Cin) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)Cout * Cin * 16;
  if (i >= total) return;
  const int kx = (int)(i & 3), ky = (int)((i >> 2) & 3);
  const int c = (int)((i >> 4) % Cin), n = (int)((i >> 4) / Cin);
  const float v = w[i];
  wf[((long)n * 16 + ky * 4 + kx) * Cin + c] = v;
  if (wd) {
    const int py = ky & 1, a = ky >> 1, px = kx & 1, bb = kx >> 1;
    wd[(((long)(py * 2 + px) * Cin + c) * 4 + a * 2 + bb) * Cout + n] = v;
  }
}
// conv1 in space-to-depth form: wf[n][ky2][px][q], q = dy*8+dx*4+c4;  wd[q][a=ky2][b'=px][n]
__global__ void prep_conv1_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over n(32) x ky2(2) x px(2) x q(16)
  if (i >= 32 * 64) return;
  const int q = i & 15, px = (i >> 4) & 1, ky2 = (i >> 5) & 1, n = i >> 6;
  const int c = q & 3, dx = (q >> 2) & 1, dy = (q >> 3) & 1;
  const float v = c < 3 ? w[((n * 3 + c) * 4 + (2 * ky2 + dy)) * 4 + (2 * px + dx)] : 0.f;
  wf[i] = v;
  if (wd) wd[((q * 2 + ky2) * 2 + px) * 32 + n] = v;
}
// Split-K partials -> parameter layout.  The partial sums (up to 296 splits) are walked in THEIR order (coalesced 256-byte
// reads per split), `kZ` threads per output share the splits and combine through shared memory; only the final value is
// scattered into dw[n][c][ky][kx].  (One thread per output walking all splits was ~36 us of pure load latency per call.)
constexpr int kUnX = 64, kUnZ = 8;
__global__ void __launch_bounds__(kUnX * kUnZ) unprep_conv_kernel(const float* __restrict__ part, int splits, float* __restrict__ dw,
                                                                  int Cout, int Cin) {
  __shared__ float red[kUnZ][kUnX + 1];
  const long total = (long)Cout * Cin * 16;
  const long j = blockIdx.x * (long)kUnX + threadIdx.x;   // index in the fprop operand layout [n][ky][kx][c]
  float s = 0.f;
  if (j < total)
    for (int z = threadIdx.y; z < splits; z += kUnZ) s += part[(long)z * total + j];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && j < total) {
#pragma unroll
    for (int y = 1; y < kUnZ; ++y) s += red[y][threadIdx.x];
    const int c = (int)(j % Cin);
    const int kyx = (int)((j / Cin) & 15), n = (int)(j / ((long)Cin * 16));
    dw[((long)n * Cin + c) * 16 + kyx] = s;
  }
}
__global__ void __launch_bounds__(kUnX * kUnZ) unprep_conv1_kernel(const float* __restrict__ part, int splits, float* __restrict__ dw,
                                                                   float* __restrict__ dbias) {
  __shared__ float red[kUnZ][kUnX + 1];
  const int j = blockIdx.x * kUnX + threadIdx.x;          // operand index [n][ky2][px][q = dy*8+dx*4+c4], 32*64 entries
  float s = 0.f;
  for (int z = threadIdx.y; z < splits; z += kUnZ) s += part[z * (32 * 64) + j];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int y = 1; y < kUnZ; ++y) s += red[y][threadIdx.x];
    const int q = j & 15, px = (j >> 4) & 1, ky2 = (j >> 5) & 1, n = j >> 6;
    const int c = q & 3, dx = (q >> 2) & 1, dy = (q >> 3) & 1;
    if (c < 3) dw[((n * 3 + c) * 4 + (2 * ky2 + dy)) * 4 + (2 * px + dx)] = s;
    // pad channel (q = 3) of tap (0,0): sum over pixels of dy[n] * 1.0 = bias gradient
    else if (dbias && (j & 63) == 3) dbias[n] = s;
  }
}
// FC1 column permutation: reference column c*100+p  <->  operand column p*256+c
__global__ void prep_fc1_kernel(const float* __restrict__ w, float* __restrict__ wg, int out, int tail, long ld) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)out * ld) return;
  const int o = (int)(i / ld);
  const long j = i % ld;
  const long K = kFeat + tail;
  float v = 0.f;
  if (j < kFeat) v = w[o * K + (j % kC4) * kPix + (j / kC4)];
  else if (j < K) v = w[o * K + j];
  wg[i] = v;
}
__global__ void unprep_fc1_kernel(const float* __restrict__ part, int splits, float* __restrict__ dw, int out, int tail, long ld) {
  const long K = kFeat + tail;
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)out * K) return;
  const int o = (int)(i / K);
  const long j = i % K;
  const long src = j < kFeat ? (j % kPix) * kC4 + (j / kPix) : j;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[((long)z * out + o) * ld + src];
  dw[i] = s;
}

// ---- clip_grad_norm_ + Adam ---------------------------------------------------------------------------------
__global__ void grad_sumsq_kernel(const float* __restrict__ g, long n, float scale, double* __restrict__ out) {
  __shared__ double red[32];
  float s = 0.f;
  double tot = 0.0;
  int cnt = 0;
  const long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    if (++cnt == 64) { tot += s; s = 0.f; cnt = 0; }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[(n4 << 2) + threadIdx.x]; s += v * v; }
  double part[1] = {(tot + s) * ((double)scale * (double)scale)};
  gc::block_sum<1>(part, red);
  if (threadIdx.x == 0) atomicAdd(out, part[0]);
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float coef, float lr_c, float b1, float b2, float eps,
                                      float inv_sqrt_bc2) {
  g *= coef;
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
  p -= lr_c * (m / denom);
}

__global__ void clip_adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
                                 const double* __restrict__ sumsq, float max_norm, float lr, float b1, float b2, float eps,
                                 float bc1, float bc2, float grad_scale, int zero_grad, const float* __restrict__ dev_hyper) {
  if (dev_hyper) {  // {lr, 1-beta1^t, 1-beta2^t} refreshed by the host between replays of a captured graph
    lr = dev_hyper[0];
    bc1 = dev_hyper[1];
    bc2 = dev_hyper[2];
  }
  float coef = 1.f;
  if (max_norm >= 0.f) {
    const float total = (float)sqrt(*sumsq);
    coef = fminf(max_norm / (total + 1e-6f), 1.f);
  }
  coef *= grad_scale;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float lr_c = lr / bc1, isb = 1.f / sqrtf(bc2);
  const long n4 = n >> 2;
  float4 *p4 = reinterpret_cast<float4*>(p), *m4 = reinterpret_cast<float4*>(m), *v4 = reinterpret_cast<float4*>(v);
  float4* g4 = reinterpret_cast<float4*>(g);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 P = p4[i], G = g4[i], M = m4[i], V = v4[i];
    adam1(P.x, G.x, M.x, V.x, coef, lr_c, b1, b2, eps, isb);
    adam1(P.y, G.y, M.y, V.y, coef, lr_c, b1, b2, eps, isb);
    adam1(P.z, G.z, M.z, V.z, coef, lr_c, b1, b2, eps, isb);
    adam1(P.w, G.w, M.w, V.w, coef, lr_c, b1, b2, eps, isb);
    p4[i] = P; m4[i] = M; v4[i] = V;
    if (zero_grad) g4[i] = z4;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long i = (n4 << 2) + threadIdx.x;
    adam1(p[i], g[i], m[i], v[i], coef, lr_c, b1, b2, eps, isb);
    if (zero_grad) g[i] = 0.f;
  }
}

}  // namespace

extern "C" {

int gc_gather_obs_s2d(const float* src, const long long* idx, float* out, int B, void* stream) {
  GC_REQUIRE(src && out && B > 0, "gc_gather_obs_s2d: bad arguments");
  GC_REQUIRE(B <= 65535, "gc_gather_obs_s2d: B=%d exceeds grid.y", B);
  GC_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)out & 15) == 0, "gc_gather_obs_s2d: pointers must be 16-byte aligned");
  gather_obs_s2d_kernel<float><<<dim3(kS2dH / kGatherRows, B), 384, 0, (cudaStream_t)stream>>>(src, idx, out);
  return gc::launch_status("gather_obs_s2d_kernel");
}

int gc_gather_obs_u8_s2d(const unsigned char* src, const long long* idx, float* out, int B, void* stream) {
  GC_REQUIRE(src && out && B > 0, "gc_gather_obs_u8_s2d: bad arguments");
  GC_REQUIRE(B <= 65535, "gc_gather_obs_u8_s2d: B=%d exceeds grid.y", B);
  GC_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)out & 15) == 0, "gc_gather_obs_u8_s2d: pointers must be 16-byte aligned");
  gather_obs_u8_s2d_kernel<<<dim3(kS2dH / kG8Rows, B), 256, 0, (cudaStream_t)stream>>>(src, idx, out);
  return gc::launch_status("gather_obs_u8_s2d_kernel");
}

int gc_gather_pair_mix_u8_s2d(const unsigned char* src_e, const long long* idx_e, const unsigned char* src_p, const long long* idx_p,
                              const float* alpha, float* out, int B, void* stream) {
  GC_REQUIRE(src_e && src_p && alpha && out && B > 0 && B <= 65535, "gc_gather_pair_mix_u8_s2d: bad arguments");
  GC_REQUIRE((((uintptr_t)src_e | (uintptr_t)src_p | (uintptr_t)out) & 15) == 0, "gc_gather_pair_mix_u8_s2d: pointers must be 16-byte aligned");
  gather_pair_mix_u8_s2d_kernel<<<dim3(kS2dH / kG8Rows, B), 256, 0, (cudaStream_t)stream>>>(src_e, idx_e, src_p, idx_p, alpha, out, B);
  return gc::launch_status("gather_pair_mix_u8_s2d_kernel");
}

int gc_gather_rows(const float* src, const long long* idx, float* out, int B, int width, long ldo, void* stream) {
  GC_REQUIRE(src && out && B > 0 && width > 0 && ldo >= width, "gc_gather_rows: bad arguments");
  const long n = (long)B * width;
  gather_rows_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, idx, out, B, width, ldo);
  return gc::launch_status("gather_rows_kernel");
}

int gc_mixup(const float* xe, const float* xp, const float* alpha, float* out, int B, long per_sample, void* stream) {
  GC_REQUIRE(xe && xp && alpha && out && B > 0 && per_sample > 0 && per_sample % 4 == 0 && B <= 65535, "gc_mixup: bad arguments");
  const long per4 = per_sample / 4;
  const int gx = (int)std::min<long>((per4 + 255) / 256, 64);
  mixup_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>((const float4*)xe, (const float4*)xp, alpha, (float4*)out, per4);
  return gc::launch_status("mixup_kernel");
}

int gc_metrics_features(const float* metrics, const float* metrics2, const float* action, const float* action2,
                        const float* alpha, const float* emb, float* out, long ldo, int pad, int B, void* stream) {
  GC_REQUIRE(metrics && emb && out && B > 0, "gc_metrics_features: bad arguments");
  GC_REQUIRE((metrics2 == nullptr) == (alpha == nullptr), "gc_metrics_features: metrics2 and alpha go together");
  GC_REQUIRE(pad >= (action ? 15 : 13) && ldo >= pad, "gc_metrics_features: pad/ldo too small");
  metrics_features_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(metrics, metrics2, action, action2, alpha, emb, out,
                                                                             ldo, pad, B);
  return gc::launch_status("metrics_features_kernel");
}

int gc_metrics_features_bwd(const float* metrics, const float* metrics2, const float* alpha, const float* d_feat, long ldf,
                            float* d_emb, int B, void* stream) {
  GC_REQUIRE(metrics && d_feat && d_emb && B > 0, "gc_metrics_features_bwd: bad arguments");
  metrics_features_bwd_kernel<<<(B * 8 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(metrics, metrics2, alpha, d_feat, ldf, d_emb, B);
  return gc::launch_status("metrics_features_bwd_kernel");
}

int gc_small_linear_fwd(const float* x, long ldx, const float* w, const float* bias, float* y, long ldy, int B, int N, int K,
                        void* stream) {
  GC_REQUIRE(x && w && y && B > 0 && N >= 1 && N <= 4 && K > 0, "gc_small_linear_fwd: bad arguments (N must be 1..4)");
  const long threads = (long)B * 32;
  small_linear_fwd_kernel<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, w, bias, y, ldy, B, N, K);
  return gc::launch_status("small_linear_fwd_kernel");
}

int gc_small_linear_bwd(const float* x, long ldx, const float* w, const float* dy, long lddy, float* dx, long lddx, float* dw,
                        float* db, int B, int B_params, int N, int K, float slope, void* stream) {
  GC_REQUIRE(x && w && dy && B > 0 && N >= 1 && N <= 4 && K > 0 && B_params <= B, "gc_small_linear_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (dx) {
    const long n = (long)B * K;
    small_linear_dx_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(x, ldx, w, dy, lddy, dx, lddx, B, N, K, slope);
    if (int e = gc::launch_status("small_linear_dx_kernel")) return e;
  }
  if (dw && B_params > 0) {
    const int rows_per = 256;
    dim3 grid((K + 127) / 128, (B_params + rows_per - 1) / rows_per);
    small_linear_dw_kernel<<<grid, 128, 0, st>>>(x, ldx, dy, lddy, dw, db, B_params, N, K, rows_per);
    if (int e = gc::launch_status("small_linear_dw_kernel")) return e;
  }
  return 0;
}

int gc_disc_loss_seed(const float* d, float* dd, double* acc, int B, float inv_norm, void* stream) {
  GC_REQUIRE(d && dd && acc && B > 0, "gc_disc_loss_seed: bad arguments");
  disc_loss_seed_kernel<<<grid_for(3L * B, 256, 2), 256, 0, (cudaStream_t)stream>>>(d, dd, acc, B,
                                                                                    inv_norm > 0.f ? inv_norm : 1.f / (float)B);
  return gc::launch_status("disc_loss_seed_kernel");
}

int gc_grad_penalty(const float* g, float* u, double* acc, int B, long per_sample, float lambda_, float s0, float s1, float s2,
                    float inv_norm, void* stream) {
  GC_REQUIRE(g && u && acc && B > 0 && per_sample % 4 == 0, "gc_grad_penalty: bad arguments");
  grad_penalty_kernel<<<B, 512, 0, (cudaStream_t)stream>>>((const float4*)g, (float4*)u, acc, inv_norm > 0.f ? inv_norm : 1.f / (float)B,
                                                           per_sample / 4, lambda_, s0, s1, s2);
  return gc::launch_status("grad_penalty_kernel");
}

int gc_reward_epilogue(const float* d, float* reward, long n, void* stream) {
  GC_REQUIRE(d && reward && n > 0, "gc_reward_epilogue: bad arguments");
  reward_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d, reward, n);
  return gc::launch_status("reward_kernel");
}

int gc_colsum(const float* x, long ld, long rows, int C, float* out, void* stream) {
  GC_REQUIRE(x && out && rows > 0 && C > 0 && ld >= C, "gc_colsum: bad arguments");
  const int cx = (C + 31) / 32;
  long chunks = std::max<long>(1, std::min<long>((rows + 63) / 64, (4L * gc::kNumSMs) / cx + 1));
  const long rows_per = (rows + chunks - 1) / chunks;
  chunks = (rows + rows_per - 1) / rows_per;
  colsum_kernel<<<dim3(cx, (unsigned)chunks), dim3(32, 8), 0, (cudaStream_t)stream>>>(x, ld, rows, C, out, rows_per);
  return gc::launch_status("colsum_kernel");
}

int gc_splitk_reduce(const float* part, int splits, long M, int N, long ldp, const float* bias, const float* mask_src, long ldm,
                     float* out, long ldo, int epilogue, float slope, void* stream) {
  GC_REQUIRE(part && out && splits >= 1 && M > 0 && N > 0, "gc_splitk_reduce: bad arguments");
  if (epilogue == 1 || epilogue == 2) GC_REQUIRE(bias, "gc_splitk_reduce: bias epilogue without bias");
  if (epilogue == 3) GC_REQUIRE(mask_src, "gc_splitk_reduce: mask epilogue without mask source");
  const long n = M * N;
  splitk_reduce_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(part, splits, M, N, ldp, bias, mask_src, ldm, out, ldo,
                                                                               epilogue, slope);
  return gc::launch_status("splitk_reduce_kernel");
}

int gc_prep_conv_weight(const float* w, float* w_fprop, float* w_dgrad, int Cout, int Cin, int layer1, void* stream) {
  GC_REQUIRE(w && w_fprop, "gc_prep_conv_weight: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (layer1) {
    GC_REQUIRE(Cout == 32 && Cin == 3, "gc_prep_conv_weight: layer1 expects 32x3x4x4");
    prep_conv1_kernel<<<8, 256, 0, st>>>(w, w_fprop, w_dgrad);
  } else {
    const long n = (long)Cout * Cin * 16;
    prep_conv_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(w, w_fprop, w_dgrad, Cout, Cin);
  }
  return gc::launch_status("prep_conv_kernel");
}

int gc_unprep_conv_wgrad(const float* part, int splits, float* dw, float* dbias, int Cout, int Cin, int layer1, void* stream) {
  GC_REQUIRE(part && dw && splits >= 1, "gc_unprep_conv_wgrad: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (layer1) {
    GC_REQUIRE(Cout == 32 && Cin == 3, "gc_unprep_conv_wgrad: layer1 expects 32x3x4x4");
    unprep_conv1_kernel<<<(32 * 64) / kUnX, dim3(kUnX, kUnZ), 0, st>>>(part, splits, dw, dbias);
  } else {
    GC_REQUIRE(dbias == nullptr, "gc_unprep_conv_wgrad: the bias-gradient column exists only for layer 1");
    const long n = (long)Cout * Cin * 16;
    unprep_conv_kernel<<<(int)((n + kUnX - 1) / kUnX), dim3(kUnX, kUnZ), 0, st>>>(part, splits, dw, Cout, Cin);
  }
  return gc::launch_status("unprep_conv_kernel");
}

int gc_prep_fc1_weight(const float* w, float* w_gemm, int out, int tail, long ld, void* stream) {
  GC_REQUIRE(w && w_gemm && out > 0 && tail >= 0 && ld >= kFeat + tail, "gc_prep_fc1_weight: bad arguments");
  const long n = (long)out * ld;
  prep_fc1_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, w_gemm, out, tail, ld);
  return gc::launch_status("prep_fc1_kernel");
}

int gc_unprep_fc1_wgrad(const float* part, int splits, float* dw, int out, int tail, long ld, void* stream) {
  GC_REQUIRE(part && dw && out > 0 && splits >= 1 && ld >= kFeat + tail, "gc_unprep_fc1_wgrad: bad arguments");
  const long n = (long)out * (kFeat + tail);
  unprep_fc1_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(part, splits, dw, out, tail, ld);
  return gc::launch_status("unprep_fc1_kernel");
}

int gc_zero_block(void* ptr, long pitch_bytes, long rows, long width_bytes, void* stream) {
  GC_REQUIRE(ptr && rows > 0 && width_bytes > 0 && pitch_bytes >= width_bytes, "gc_zero_block: bad arguments");
  GC_CUDA_OK(cudaMemset2DAsync(ptr, (size_t)pitch_bytes, 0, (size_t)width_bytes, (size_t)rows, (cudaStream_t)stream));
  return 0;
}

int gc_grad_sumsq(const float* grad, long n, float grad_scale, double* sumsq, void* stream) {
  GC_REQUIRE(grad && sumsq && n > 0, "gc_grad_sumsq: bad arguments");
  GC_REQUIRE(((uintptr_t)grad & 15) == 0, "gc_grad_sumsq: grad must be 16-byte aligned");
  GC_CUDA_OK(cudaMemsetAsync(sumsq, 0, sizeof(double), (cudaStream_t)stream));
  grad_sumsq_kernel<<<grid_for(n / 4 + 1, 256, 4), 256, 0, (cudaStream_t)stream>>>(grad, n, grad_scale, sumsq);
  return gc::launch_status("grad_sumsq_kernel");
}

int gc_clip_adam(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long n, const double* sumsq, float max_norm,
                 float lr, float beta1, float beta2, float eps, float bias_corr1, float bias_corr2, float grad_scale,
                 int zero_grad, const float* dev_hyper, void* stream) {
  GC_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0, "gc_clip_adam: bad arguments");
  GC_REQUIRE(max_norm < 0.f || sumsq, "gc_clip_adam: clipping needs sumsq");
  GC_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
             "gc_clip_adam: buffers must be 16-byte aligned");
  clip_adam_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, sumsq, max_norm,
                                                                               lr, beta1, beta2, eps, bias_corr1, bias_corr2, grad_scale, zero_grad,
                                                                               dev_hyper);
  return gc::launch_status("clip_adam_kernel");
}

}  // extern "C"
