// aux.cu — dice head, EAST RBOX loss, restore_rectangle, and housekeeping entry points.
//
//  dice      nets/model.py:145-159 == nets/model_vgg_16.py:179-193 dice_coefficient,
//            nets/model_vgg_16.py:196-225 loss (2*dice(pixel) + sum_d dice(link_d))  [L7, L8]
//  east      EAST RBOX loss — NOT in the reference; restated from upstream argman/EAST
//            model.loss (parity unpinned)
//  restore   datasets/icdar.py:410-483 restore_rectangle_rbox                        [E1]
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "raster.cuh"

namespace plh {

size_t loss_workspace_bytes(int B, int H, int W);             // loss.cu
size_t decode_workspace_bytes(int B, int H, int W, int K);    // decode.cu

constexpr int kRedThreads = 256;
constexpr int kRedMaxCTAs = kNumSMs * 4;

struct RedHeader {
  unsigned ticket;
  int pad[31];
};

// ------------------------------------------------------------------ dice
// Pass 1: per channel c: I_c = sum t*p*m, T_c = sum t*m, P_c = sum p*m  (fp32 products,
// fp32 per-thread partials, fp64 across CTAs in a fixed order).
// Channels: CA from tensor A ([M,CA]) followed by CB from tensor B ([M,CB]).
template <int CA, int CB>
__global__ void __launch_bounds__(kRedThreads)
dice_reduce_kernel(const float* __restrict__ tA, const float* __restrict__ pA, const float* __restrict__ tB,
                   const float* __restrict__ pB, const float* __restrict__ mask, long long M,
                   float* __restrict__ partials, RedHeader* __restrict__ hdr, const float wA, const float wB,
                   float* __restrict__ out) {
  constexpr int C = CA + CB;
  __shared__ float s_red[kRedThreads / 32][3 * C];
  __shared__ bool s_last;
  float acc[3 * C];
#pragma unroll
  for (int i = 0; i < 3 * C; ++i) acc[i] = 0.f;
  const long long stride = (long long)gridDim.x * kRedThreads;
  for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < M; i += stride) {
    const float m = __ldg(mask + i);
    float t[C], p[C];
#pragma unroll
    for (int c = 0; c < CA; ++c) t[c] = __ldg(tA + i * CA + c), p[c] = __ldg(pA + i * CA + c);
    if (CB == 8) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(tB) + i * 2), b = __ldg(reinterpret_cast<const float4*>(tB) + i * 2 + 1);
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(pB) + i * 2), d4 = __ldg(reinterpret_cast<const float4*>(pB) + i * 2 + 1);
      t[CA + 0] = a.x, t[CA + 1] = a.y, t[CA + 2] = a.z, t[CA + 3] = a.w;
      t[CA + 4] = b.x, t[CA + 5] = b.y, t[CA + 6] = b.z, t[CA + 7] = b.w;
      p[CA + 0] = c4.x, p[CA + 1] = c4.y, p[CA + 2] = c4.z, p[CA + 3] = c4.w;
      p[CA + 4] = d4.x, p[CA + 5] = d4.y, p[CA + 6] = d4.z, p[CA + 7] = d4.w;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      acc[3 * c + 0] += t[c] * p[c] * m;  // y_true * y_pred * training_mask (model.py:155)
      acc[3 * c + 1] += t[c] * m;
      acc[3 * c + 2] += p[c] * m;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 3 * C; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3 * C) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kRedThreads / 32; ++w) s += s_red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 32 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&hdr->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ double s_fin[3 * C];
  __shared__ double s_tmp[(kRedThreads / (3 * C)) * 3 * C];
  block_final_reduce<3 * C, kRedThreads>(partials, 32, gridDim.x, s_fin, s_tmp);
  if (threadIdx.x == 0) {
    // out: [0] total, then per channel: dice, I, U
    float total = 0.f;
    for (int c = 0; c < C; ++c) {
      const float I = (float)s_fin[3 * c];
      const float U = __fadd_rn(__fadd_rn((float)s_fin[3 * c + 1], (float)s_fin[3 * c + 2]), 1e-5f);  // model.py:154,156
      const float dice = __fsub_rn(1.f, __fdiv_rn(__fmul_rn(2.f, I), U));                               // model.py:157
      out[1 + 3 * c] = dice, out[2 + 3 * c] = I, out[3 + 3 * c] = U;
      total += (c < CA ? wA : wB) * dice;
    }
    out[0] = total;
    hdr->ticket = 0;
  }
}

// Pass 2: grad wrt pred = -2 w m (t U - I) / U^2
template <int CA, int CB>
__global__ void __launch_bounds__(kRedThreads)
dice_grad_kernel(const float* __restrict__ tA, const float* __restrict__ tB, const float* __restrict__ mask,
                 long long M, const float* __restrict__ out, const float wA, const float wB,
                 float* __restrict__ gA, float* __restrict__ gB) {
  constexpr int C = CA + CB;
  float I[C], U[C], k[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    I[c] = out[2 + 3 * c], U[c] = out[3 + 3 * c];
    k[c] = (c < CA ? wA : wB) * -2.f;
  }
  const long long stride = (long long)gridDim.x * kRedThreads;
  for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < M; i += stride) {
    const float m = __ldg(mask + i);
#pragma unroll
    for (int c = 0; c < CA; ++c) {
      const float t = __ldg(tA + i * CA + c);
      gA[i * CA + c] = k[c] * m * (t * U[c] - I[c]) / (U[c] * U[c]);
    }
    if (CB == 8) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(tB) + i * 2), b = __ldg(reinterpret_cast<const float4*>(tB) + i * 2 + 1);
      const float t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      float g[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) g[c] = k[CA + c] * m * (t[c] * U[CA + c] - I[CA + c]) / (U[CA + c] * U[CA + c]);
      stg_stream4(reinterpret_cast<float4*>(gB) + i * 2, make_float4(g[0], g[1], g[2], g[3]));
      stg_stream4(reinterpret_cast<float4*>(gB) + i * 2 + 1, make_float4(g[4], g[5], g[6], g[7]));
    }
  }
}

static size_t red_workspace_bytes() { return 256 + sizeof(float) * 32 * kRedMaxCTAs; }

template <int CA, int CB>
static int run_dice(const float* tA, const float* pA, const float* tB, const float* pB, const float* mask, long long M,
                    float wA, float wB, float* out, float* gA, float* gB, void* workspace, size_t workspace_bytes,
                    cudaStream_t s) {
  if (!workspace || workspace_bytes < red_workspace_bytes() || !aligned16(workspace)) return PLH_E_WORKSPACE;
  RedHeader* hdr = (RedHeader*)workspace;
  float* partials = (float*)((char*)workspace + 256);
  cudaError_t e = cudaMemsetAsync(hdr, 0, sizeof(RedHeader), s);
  if (e != cudaSuccess) return (int)e;
  const int grid = (int)std::min<long long>((M + kRedThreads - 1) / kRedThreads, kRedMaxCTAs);
  dice_reduce_kernel<CA, CB><<<grid, kRedThreads, 0, s>>>(tA, pA, tB, pB, mask, M, partials, hdr, wA, wB, out);
  int rc = launch_status();
  if (rc) return rc;
  if (gA || gB) {
    dice_grad_kernel<CA, CB><<<grid, kRedThreads, 0, s>>>(tA, tB, mask, M, out, wA, wB, gA, gB);
    rc = launch_status();
  }
  return rc;
}

// ------------------------------------------------------------------ EAST RBOX loss (upstream argman/EAST)
// L = mean(y_true * mask * (L_AABB + 20 L_theta)) + 0.01 * dice(y_true, y_pred, mask)
__global__ void __launch_bounds__(kRedThreads)
east_reduce_kernel(const float* __restrict__ sg, const float* __restrict__ sp, const float* __restrict__ gg,
                   const float* __restrict__ gp, const float* __restrict__ mask, long long M,
                   float* __restrict__ partials, RedHeader* __restrict__ hdr, float* __restrict__ out) {
  __shared__ float s_red[kRedThreads / 32][5];
  __shared__ bool s_last;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // I, T, P, sum aabb*w, sum theta*w
  const long long stride = (long long)gridDim.x * kRedThreads;
  for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < M; i += stride) {
    const float m = __ldg(mask + i), t = __ldg(sg + i), p = __ldg(sp + i);
    acc[0] += t * p * m, acc[1] += t * m, acc[2] += p * m;
    const float w = t * m;
    float g[5], q[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) g[c] = __ldg(gg + i * 5 + c), q[c] = __ldg(gp + i * 5 + c);
    const float area_gt = (g[0] + g[2]) * (g[1] + g[3]);
    const float area_pr = (q[0] + q[2]) * (q[1] + q[3]);
    const float w_union = fminf(g[1], q[1]) + fminf(g[3], q[3]);
    const float h_union = fminf(g[0], q[0]) + fminf(g[2], q[2]);
    const float ai = w_union * h_union;
    const float au = area_gt + area_pr - ai;
    const float l_aabb = -logf((ai + 1.f) / (au + 1.f));
    const float l_theta = 1.f - cosf(q[4] - g[4]);
    acc[3] += l_aabb * w, acc[4] += l_theta * w;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kRedThreads / 32; ++w) s += s_red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 32 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&hdr->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ double s_fin[5];
  __shared__ double s_tmp[(kRedThreads / 5) * 5];
  block_final_reduce<5, kRedThreads>(partials, 32, gridDim.x, s_fin, s_tmp);
  if (threadIdx.x == 0) {
    const float I = (float)s_fin[0];
    const float U = (float)s_fin[1] + (float)s_fin[2] + 1e-5f;
    const float dice = 1.f - 2.f * I / U;
    const float lg = (float)((s_fin[3] + 20.0 * s_fin[4]) / (double)M);
    out[0] = lg + 0.01f * dice;
    out[1] = dice, out[2] = lg, out[3] = I, out[4] = U;
    out[5] = (float)s_fin[3], out[6] = (float)s_fin[4], out[7] = 0.f;
    hdr->ticket = 0;
  }
}

__global__ void __launch_bounds__(kRedThreads)
east_grad_kernel(const float* __restrict__ sg, const float* __restrict__ gg, const float* __restrict__ gp,
                 const float* __restrict__ mask, long long M, const float* __restrict__ out,
                 float* __restrict__ grad_score, float* __restrict__ grad_geo) {
  const float I = out[3], U = out[4];
  const float invM = 1.f / (float)M;
  const long long stride = (long long)gridDim.x * kRedThreads;
  for (long long i = (long long)blockIdx.x * kRedThreads + threadIdx.x; i < M; i += stride) {
    const float m = __ldg(mask + i), t = __ldg(sg + i);
    grad_score[i] = 0.01f * -2.f * m * (t * U - I) / (U * U);
    const float w = t * m * invM;
    float g[5], q[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) g[c] = __ldg(gg + i * 5 + c), q[c] = __ldg(gp + i * 5 + c);
    const float w_union = fminf(g[1], q[1]) + fminf(g[3], q[3]);
    const float h_union = fminf(g[0], q[0]) + fminf(g[2], q[2]);
    const float ai = w_union * h_union;
    const float au = (g[0] + g[2]) * (g[1] + g[3]) + (q[0] + q[2]) * (q[1] + q[3]) - ai;
    const float ri = 1.f / (ai + 1.f), ru = 1.f / (au + 1.f);
    // tf.minimum(gt, pred): the gradient goes to pred only where pred < gt
    const float dAi[4] = {q[0] < g[0] ? w_union : 0.f, q[1] < g[1] ? h_union : 0.f, q[2] < g[2] ? w_union : 0.f,
                          q[3] < g[3] ? h_union : 0.f};
    const float dAp[4] = {q[1] + q[3], q[0] + q[2], q[1] + q[3], q[0] + q[2]};
#pragma unroll
    for (int c = 0; c < 4; ++c) grad_geo[i * 5 + c] = w * (-ri * dAi[c] + ru * (dAp[c] - dAi[c]));
    grad_geo[i * 5 + 4] = w * 20.f * sinf(q[4] - g[4]);
  }
}

// ------------------------------------------------------------------ restore_rectangle_rbox
// icdar.py:479 returns the theta >= 0 rows first, then the theta < 0 rows (quirk Q16):
// a stable partition, done with a block-count / scan / scatter triple.
constexpr int kRRBlock = 1024;

template <typename TG>
__global__ void __launch_bounds__(kRRBlock)
rr_count_kernel(const TG* __restrict__ geometry, int N, int* __restrict__ blockcnt) {
  __shared__ int s_c;
  if (threadIdx.x == 0) s_c = 0;
  __syncthreads();
  const int i = blockIdx.x * kRRBlock + threadIdx.x;
  const bool nn = i < N && geometry[(size_t)i * 5 + 4] >= (TG)0;
  const unsigned m = __ballot_sync(0xffffffffu, nn);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_c, __popc(m));
  __syncthreads();
  if (threadIdx.x == 0) blockcnt[blockIdx.x] = s_c;
}

__global__ void __launch_bounds__(1024) rr_scan_kernel(int* __restrict__ blockcnt, int nb, int* __restrict__ total) {
  // exclusive scan of blockcnt in place by one CTA
  __shared__ int s_w[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < nb; i0 += 1024) {
    const int i = i0 + tid;
    const int v = i < nb ? blockcnt[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const int wv = s_w[lane];
      int winc = wv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      s_w[lane] = winc - wv;
    }
    __syncthreads();
    const int excl = s_base + s_w[warp] + inc - v;
    if (i < nb) blockcnt[i] = excl;
    __syncthreads();
    if (tid == 1023) s_base = excl + v;
    __syncthreads();
  }
  if (tid == 0) *total = s_base;
}

// TG / TO: dtype of geometry / origin.  numpy computes the distance sums and cos / sin in geometry's dtype,
// everything else in float64 (np.zeros is float64); a float64 origin only enters the final translation.
__device__ __forceinline__ double rr_cos(float a) { return (double)(float)cos((double)a); }
__device__ __forceinline__ double rr_cos(double a) { return cos(a); }
__device__ __forceinline__ double rr_sin(float a) { return (double)(float)sin((double)a); }
__device__ __forceinline__ double rr_sin(double a) { return sin(a); }

template <typename TG, typename TO>
__global__ void __launch_bounds__(kRRBlock)
rr_scatter_kernel(const TO* __restrict__ origin, const TG* __restrict__ geometry, int N,
                  const int* __restrict__ blockoff, const int* __restrict__ total_nn, double* __restrict__ out,
                  int32_t* __restrict__ out_index) {
  __shared__ int s_w[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = blockIdx.x * kRRBlock + tid;
  const bool valid = i < N;
  TG d0 = 0, d1 = 0, d2 = 0, d3 = 0, th = -1;
  if (valid) {
    const TG* g = geometry + (size_t)i * 5;
    d0 = g[0], d1 = g[1], d2 = g[2], d3 = g[3], th = g[4];
  }
  const bool nn = valid && th >= (TG)0;
  const unsigned m = __ballot_sync(0xffffffffu, nn);
  if (lane == 0) s_w[warp] = __popc(m);
  __syncthreads();
  if (warp == 0) {
    const int wv = s_w[lane];
    int winc = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    s_w[lane] = winc - wv;
  }
  __syncthreads();
  if (!valid) return;
  const int nn_before = blockoff[blockIdx.x] + s_w[warp] + __popc(m & ((1u << lane) - 1u));
  const int row = nn ? nn_before : *total_nn + (i - nn_before);
  // sums of the distances and cos/sin of the angle in geometry's dtype (numpy arithmetic of that dtype),
  // products and sums in fp64 (numpy upcasts: np.zeros is float64), icdar.py:417-443 / :450-476
  const TG hh = -d0 - d2;
  double px[5], py[5];
  double c, s_;
  if (nn) {
    const TG ww = d1 + d3;
    px[0] = 0.0, py[0] = hh; px[1] = ww, py[1] = hh; px[2] = ww, py[2] = 0.0; px[3] = 0.0, py[3] = 0.0;
    px[4] = d3, py[4] = -d2;
    c = rr_cos(th), s_ = rr_sin(th);
    // rotate_matrix_x = [cos, sin], rotate_matrix_y = [-sin, cos]
  } else {
    const TG ww = -d1 - d3;
    px[0] = ww, py[0] = hh; px[1] = 0.0, py[1] = hh; px[2] = 0.0, py[2] = 0.0; px[3] = ww, py[3] = 0.0;
    px[4] = -d1, py[4] = -d2;
    const TG nth = -th;
    c = rr_cos(nth), s_ = -rr_sin(nth);
    // rotate_matrix_x = [cos(-a), -sin(-a)], rotate_matrix_y = [sin(-a), cos(-a)]
  }
  double rx[5], ry[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    rx[k] = __dadd_rn(__dmul_rn(c, px[k]), __dmul_rn(s_, py[k]));
    ry[k] = __dadd_rn(__dmul_rn(-s_, px[k]), __dmul_rn(c, py[k]));
  }
  const double ox = (double)origin[(size_t)i * 2] - rx[4], oy = (double)origin[(size_t)i * 2 + 1] - ry[4];
  double* o = out + (size_t)row * 8;
#pragma unroll
  for (int k = 0; k < 4; ++k) o[2 * k] = rx[k] + ox, o[2 * k + 1] = ry[k] + oy;
  if (out_index) out_index[row] = i;
}


// ------------------------------------------------------------------ link labels from a polygon-id map
// tool/pixellink_fn.py:81-111 (generate_rbox's double loop over polygons and their pixels, with
// valid_link :9-47): a pixel inside polygon v gets link label 1 in every direction if it lies on the map
// border, else 1 exactly in the directions whose neighbour carries the same id; background stays 0.
// Channel order = the reference's: left, left_down, left_up, right, right_down, right_up, up, down.
// One thread per pixel: 1 B read (+8 neighbour bytes from L1), 32 (+4) B written with 128-bit stores.
__constant__ int c_ldy[8] = {0, 1, -1, 0, 1, -1, -1, 1};
__constant__ int c_ldx[8] = {-1, -1, -1, 1, 1, 1, 0, 0};

__global__ void __launch_bounds__(256)
link_labels_kernel(const uint8_t* __restrict__ ids, int H, int W, long long total_px, float* __restrict__ link_lab,
                   float* __restrict__ pix_lab) {
  const int N = H * W;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total_px; g += stride) {
    const int v = (int)(g % N);
    const int y = v / W, x = v - y * W;
    const unsigned id = ids[g];
    float l[8];
    if (id == 0) {
#pragma unroll
      for (int d = 0; d < 8; ++d) l[d] = 0.f;
    } else if (x == 0 || y == 0 || x == W - 1 || y == H - 1) {
#pragma unroll
      for (int d = 0; d < 8; ++d) l[d] = 1.f;
    } else {
#pragma unroll
      for (int d = 0; d < 8; ++d) l[d] = ids[g + (long long)c_ldy[d] * W + c_ldx[d]] == id ? 1.f : 0.f;
    }
    float4* o = reinterpret_cast<float4*>(link_lab) + g * 2;
    stg_stream4(o, make_float4(l[0], l[1], l[2], l[3]));
    stg_stream4(o + 1, make_float4(l[4], l[5], l[6], l[7]));
    if (pix_lab) pix_lab[g] = id != 0 ? 1.f : 0.f;
  }
}

// ------------------------------------------------------------------ link labels, EAST-fork generator
// datasets/icdar.py:83-105 valid_link + :486-539 generate_rbox (the generator train.sh actually uses), as the
// code stands (quirk Q17): the direction names move the OTHER axis ('up' is x-1, 'left' is y-1, ...), the
// early return tests x against h-1 and y against w-1 (square maps only), an index of -1 wraps to the far side,
// and a link asks whether the neighbour is text in score_map AS IT IS when the pixel's polygon is processed:
// polygons are filled in order and a pixel is (re)labelled by every polygon that covers it, so the value that
// stays is the one of its LAST covering polygon k, for which the neighbour counts iff some polygon <= k covers
// it.  Inputs: last[y][x] = 1-based index of the last polygon covering the pixel (poly_mask), first[y][x] = of
// the first one (0 = background); outputs at every `stride`-th pixel (icdar.py:632-634 keeps [::4, ::4]).
// Channel order: left, left_down, left_up, right, right_down, right_up, up, down — as (dx, dy) of the code:
__constant__ int c_idx[8] = {0, 1, -1, 0, 1, -1, -1, 1};
__constant__ int c_idy[8] = {-1, -1, -1, 1, 1, 1, 0, 0};

__global__ void __launch_bounds__(256)
link_labels_icdar_kernel(const int32_t* __restrict__ last, const int32_t* __restrict__ first, int B, int H, int W,
                         int stride, float* __restrict__ link_lab, float* __restrict__ score) {
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  const long long total = (long long)B * Ho * Wo;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(g / ((long long)Ho * Wo));
    const int r = (int)(g - (long long)b * Ho * Wo);
    const int y = (r / Wo) * stride, x = (r % Wo) * stride;
    const int32_t* L = last + (size_t)b * H * W;
    const int32_t* F = first + (size_t)b * H * W;
    const int k = L[(size_t)y * W + x];
    float l[8];
    if (k == 0) {
#pragma unroll
      for (int d = 0; d < 8; ++d) l[d] = 0.f;
    } else if (x == H - 1 || y == W - 1) {   // icdar.py:84 `point[0] == h - 1 or point[1] == w - 1`
#pragma unroll
      for (int d = 0; d < 8; ++d) l[d] = 1.f;
    } else {
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        int qx = x + c_idx[d], qy = y + c_idy[d];
        qx = qx < 0 ? qx + W : qx, qy = qy < 0 ? qy + H : qy;   // numpy's negative index
        const int f = F[(size_t)qy * W + qx];
        l[d] = (f != 0 && f <= k) ? 1.f : 0.f;
      }
    }
    float4* o = reinterpret_cast<float4*>(link_lab) + g * 2;
    stg_stream4(o, make_float4(l[0], l[1], l[2], l[3]));
    stg_stream4(o + 1, make_float4(l[4], l[5], l[6], l[7]));
    if (score) score[g] = k != 0 ? 1.f : 0.f;
  }
}


// ------------------------------------------------------------------ N1: polygon rasterisation
// tool/pixellink_fn.py:66-79 and datasets/icdar.py:486-514 fill the image's quadrilaterals one after the other with
// cv2.fillPoly (then, in the pixel_link fork, shrink the canvas with cv2.resize(INTER_NEAREST)).  Here nothing is
// drawn: one CTA per OUTPUT row looks at the source row it samples; thread t turns polygon t into its (at most six)
// intervals on that row (raster.cuh: cv2's outline + scan-line spans, clipped like cv2 clips), then every thread
// walks the polygons in order for its pixels: the last one covering the pixel is cv2's result, the first one is
// what the EAST fork's "text so far" test needs, a flagged one clears the training mask.
constexpr int kFillChunk = 256;
constexpr int kFillPPT = 8;  // output pixels per thread: Wo <= 2048

__global__ void __launch_bounds__(kFillChunk)
fill_quads_kernel(const int32_t* __restrict__ quads, const int32_t* __restrict__ quad_off,
                  const uint8_t* __restrict__ zero_flags, int H, int W, int Ho, int Wo, int mode, int stride, double ify,
                  double ifx, int32_t* __restrict__ last_ids, int32_t* __restrict__ first_ids,
                  uint8_t* __restrict__ ids_u8, float* __restrict__ score, uint8_t* __restrict__ training_mask) {
  __shared__ Ivs s_iv[kFillChunk];
  __shared__ uint8_t s_zero[kFillChunk];
  const int b = blockIdx.y, oy = blockIdx.x, tid = threadIdx.x;
  // cv2.resize(INTER_NEAREST): source index = min(floor(dst * (1 / (dsize / ssize))), ssize - 1), in double
  const int sy = mode == 0 ? oy * stride : min((int)floor(__dmul_rn((double)oy, ify)), H - 1);
  const int q0 = quad_off[b], n = quad_off[b + 1] - q0;
  int last[kFillPPT], first[kFillPPT], sx[kFillPPT];
  bool keep[kFillPPT];
#pragma unroll
  for (int j = 0; j < kFillPPT; ++j) {
    const int ox = tid + j * kFillChunk;
    sx[j] = mode == 0 ? ox * stride : min((int)floor(__dmul_rn((double)ox, ifx)), W - 1);
    last[j] = 0, first[j] = 0, keep[j] = true;
  }
  for (int c0 = 0; c0 < n; c0 += kFillChunk) {
    const int m = min(n - c0, kFillChunk);
    __syncthreads();
    if (tid < m) {
      int qx[4], qy[4];
      int ymin = INT_MAX, ymax = INT_MIN;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        qx[k] = quads[(size_t)(q0 + c0 + tid) * 8 + 2 * k], qy[k] = quads[(size_t)(q0 + c0 + tid) * 8 + 2 * k + 1];
        ymin = min(ymin, qy[k]), ymax = max(ymax, qy[k]);
      }
      s_iv[tid].n = 0;
      if (sy >= ymin && sy <= ymax) {
        QuadEdges E;
        quad_edges(qx, qy, W, H, E);
        quad_row_intervals(E, sy, W, s_iv[tid]);
      }
      s_zero[tid] = zero_flags ? zero_flags[q0 + c0 + tid] : 0;
    }
    __syncthreads();
    for (int k = 0; k < m; ++k) {
      const int ni = s_iv[k].n;
      if (ni == 0) continue;
#pragma unroll
      for (int j = 0; j < kFillPPT; ++j) {
        bool in = false;
        for (int i = 0; i < ni; ++i) in = in || (sx[j] >= s_iv[k].a[i] && sx[j] <= s_iv[k].b[i]);
        if (in) {
          last[j] = c0 + k + 1;
          if (first[j] == 0) first[j] = c0 + k + 1;
          if (s_zero[k]) keep[j] = false;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kFillPPT; ++j) {
    const int ox = tid + j * kFillChunk;
    if (ox < Wo) {
      const size_t o = ((size_t)b * Ho + oy) * Wo + ox;
      last_ids[o] = last[j];
      if (first_ids) first_ids[o] = first[j];
      if (ids_u8) ids_u8[o] = (uint8_t)min(last[j], 255);  // the reference's canvas is uint8: cv2 saturates the fill value
      if (score) score[o] = last[j] ? 1.f : 0.f;
      if (training_mask) training_mask[o] = keep[j] ? 1 : 0;
    }
  }
}

}  // namespace plh

using namespace plh;

extern "C" int plh_fill_quads(const int32_t* quads, const int32_t* quad_off, const uint8_t* zero_flags, int B, int H, int W,
                              int Ho, int Wo, int mode, int stride, int32_t* last_ids, int32_t* first_ids, uint8_t* ids_u8,
                              float* score, uint8_t* training_mask, void* stream) {
  if (!quads || !quad_off || !last_ids) return PLH_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0 || Wo > kFillChunk * kFillPPT || H >= kMaxCoord || W >= kMaxCoord)
    return PLH_E_SHAPE;
  if (mode != 0 && mode != 1) return PLH_E_PARAM;
  if (mode == 0 && (stride <= 0 || (long long)(Ho - 1) * stride >= H || (long long)(Wo - 1) * stride >= W)) return PLH_E_PARAM;
  // cv::resize: inv_scale = (double)dsize / ssize; resizeNN: ifx = 1. / inv_scale
  const double ify = 1.0 / ((double)Ho / (double)H), ifx = 1.0 / ((double)Wo / (double)W);
  fill_quads_kernel<<<dim3(Ho, B), kFillChunk, 0, (cudaStream_t)stream>>>(quads, quad_off, zero_flags, H, W, Ho, Wo, mode,
                                                                          stride, ify, ifx, last_ids, first_ids, ids_u8,
                                                                          score, training_mask);
  return launch_status();
}

extern "C" int plh_link_labels_icdar(const int32_t* last_ids, const int32_t* first_ids, int B, int H, int W, int stride,
                                     float* link_lab, float* score, void* stream) {
  if (!last_ids || !first_ids || !link_lab) return PLH_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || stride <= 0 || (long long)B * H * W > (1ll << 31) - 1) return PLH_E_SHAPE;
  if (H != W) return PLH_E_SHAPE;   // the reference indexes out of range on non-square maps (quirk Q17)
  if (!aligned16(link_lab)) return PLH_E_ALIGN;
  const long long total = (long long)B * ((H + stride - 1) / stride) * ((W + stride - 1) / stride);
  const int grid = (int)std::min<long long>((total + 255) / 256, kNumSMs * 16);
  link_labels_icdar_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(last_ids, first_ids, B, H, W, stride, link_lab, score);
  return launch_status();
}

extern "C" int plh_dice(const float* y_true, const float* y_pred, const float* mask, long long M, float* out,
                        float* grad, void* workspace, size_t workspace_bytes, void* stream) {
  if (!y_true || !y_pred || !mask || !out) return PLH_E_NULL;
  if (M <= 0) return PLH_E_SHAPE;
  return run_dice<1, 0>(y_true, y_pred, nullptr, nullptr, mask, M, 1.f, 0.f, out, grad, nullptr, workspace,
                        workspace_bytes, (cudaStream_t)stream);
}

extern "C" int plh_dice_head(const float* t_pix, const float* p_pix, const float* t_link, const float* p_link,
                             const float* mask, long long M, float* out, float* grad_pix, float* grad_link,
                             void* workspace, size_t workspace_bytes, void* stream) {
  if (!t_pix || !p_pix || !t_link || !p_link || !mask || !out) return PLH_E_NULL;
  if ((grad_pix == nullptr) != (grad_link == nullptr)) return PLH_E_NULL;
  if (M <= 0) return PLH_E_SHAPE;
  if (!aligned16(t_link) || !aligned16(p_link) || (grad_link && !aligned16(grad_link))) return PLH_E_ALIGN;
  // model_vgg_16.py:205,223-225: classification (pixel) dice x2, the 8 link dice x1
  return run_dice<1, 8>(t_pix, p_pix, t_link, p_link, mask, M, 2.f, 1.f, out, grad_pix, grad_link, workspace,
                        workspace_bytes, (cudaStream_t)stream);
}

extern "C" int plh_east_loss(const float* score_gt, const float* score_pred, const float* geo_gt,
                             const float* geo_pred, const float* mask, long long M, float* out, float* grad_score,
                             float* grad_geo, void* workspace, size_t workspace_bytes, void* stream) {
  if (!score_gt || !score_pred || !geo_gt || !geo_pred || !mask || !out) return PLH_E_NULL;
  if ((grad_score == nullptr) != (grad_geo == nullptr)) return PLH_E_NULL;
  if (M <= 0) return PLH_E_SHAPE;
  if (!workspace || workspace_bytes < red_workspace_bytes() || !aligned16(workspace)) return PLH_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  RedHeader* hdr = (RedHeader*)workspace;
  float* partials = (float*)((char*)workspace + 256);
  cudaError_t e = cudaMemsetAsync(hdr, 0, sizeof(RedHeader), s);
  if (e != cudaSuccess) return (int)e;
  const int grid = (int)std::min<long long>((M + kRedThreads - 1) / kRedThreads, kRedMaxCTAs);
  east_reduce_kernel<<<grid, kRedThreads, 0, s>>>(score_gt, score_pred, geo_gt, geo_pred, mask, M, partials, hdr, out);
  int rc = launch_status();
  if (rc) return rc;
  if (grad_score) {
    east_grad_kernel<<<grid, kRedThreads, 0, s>>>(score_gt, geo_gt, geo_pred, mask, M, out, grad_score, grad_geo);
    rc = launch_status();
  }
  return rc;
}

template <typename TG, typename TO>
static int restore_rectangle_impl(const TO* origin, const TG* geometry, int N, double* out, int32_t* out_index,
                                  void* workspace, size_t workspace_bytes, cudaStream_t s) {
  const int nb = (N + kRRBlock - 1) / kRRBlock;
  if (!workspace || workspace_bytes < (size_t)(nb + 1) * 4 + 256) return PLH_E_WORKSPACE;
  int* blockcnt = (int*)workspace;
  int* total = blockcnt + nb;
  rr_count_kernel<TG><<<nb, kRRBlock, 0, s>>>(geometry, N, blockcnt);
  int rc = launch_status();
  if (rc) return rc;
  rr_scan_kernel<<<1, 1024, 0, s>>>(blockcnt, nb, total);
  if ((rc = launch_status())) return rc;
  rr_scatter_kernel<TG, TO><<<nb, kRRBlock, 0, s>>>(origin, geometry, N, blockcnt, total, out, out_index);
  return launch_status();
}

extern "C" int plh_restore_rectangle_ex(const void* origin, int origin_is_f64, const void* geometry, int geometry_is_f64,
                                        int N, double* out, int32_t* out_index, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  if (N == 0) return PLH_OK;
  if (!origin || !geometry || !out) return PLH_E_NULL;
  if (N < 0) return PLH_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  if (geometry_is_f64 && origin_is_f64)
    return restore_rectangle_impl((const double*)origin, (const double*)geometry, N, out, out_index, workspace, workspace_bytes, s);
  if (geometry_is_f64)
    return restore_rectangle_impl((const float*)origin, (const double*)geometry, N, out, out_index, workspace, workspace_bytes, s);
  if (origin_is_f64)
    return restore_rectangle_impl((const double*)origin, (const float*)geometry, N, out, out_index, workspace, workspace_bytes, s);
  return restore_rectangle_impl((const float*)origin, (const float*)geometry, N, out, out_index, workspace, workspace_bytes, s);
}

extern "C" int plh_restore_rectangle(const float* origin, const float* geometry, int N, double* out,
                                     int32_t* out_index, void* workspace, size_t workspace_bytes, void* stream) {
  return plh_restore_rectangle_ex(origin, 0, geometry, 0, N, out, out_index, workspace, workspace_bytes, stream);
}

extern "C" int plh_link_labels(const uint8_t* ids, int B, int H, int W, float* link_lab, float* pix_lab, void* stream) {
  if (!ids || !link_lab) return PLH_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || (long long)B * H * W > (1ll << 31) - 1) return PLH_E_SHAPE;
  if (!aligned16(link_lab)) return PLH_E_ALIGN;
  const long long total = (long long)B * H * W;
  const int grid = (int)std::min<long long>((total + 255) / 256, kNumSMs * 16);
  link_labels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ids, H, W, total, link_lab, pix_lab);
  return launch_status();
}

extern "C" size_t plh_workspace_bytes(int op, int B, int H, int W, int K) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  switch (op) {
    case PLH_OP_LOSS: return loss_workspace_bytes(B, H, W);
    case PLH_OP_DECODE: return decode_workspace_bytes(B, H, W, K > 0 ? K : 1);
    case PLH_OP_LOSS_DECODE:
      return align_up(loss_workspace_bytes(B, H, W), 256) + decode_workspace_bytes(B, H, W, K > 0 ? K : 1);
    case PLH_OP_DICE:
    case PLH_OP_EAST_LOSS: return red_workspace_bytes();
    case PLH_OP_RESTORE: return ((size_t)B * H * W / kRRBlock + 2) * 4 + 256;
    default: return 0;
  }
}

#ifdef PLH_TIMELINE
namespace plh { int tl_set_loss(unsigned long long*); int tl_set_decode(unsigned long long*); }
static unsigned long long* h_tl = nullptr;
extern "C" __attribute__((visibility("default"))) int plh_timeline_reset(void) {
  if (!h_tl) {
    if (cudaMalloc(&h_tl, 64 * 8) != cudaSuccess) return -1;
    plh::tl_set_loss(h_tl);
    plh::tl_set_decode(h_tl);
  }
  unsigned long long h[64];
  for (int i = 0; i < 32; ++i) h[2 * i] = ~0ull, h[2 * i + 1] = 0ull;
  return (int)cudaMemcpy(h_tl, h, sizeof(h), cudaMemcpyHostToDevice);
}
extern "C" __attribute__((visibility("default"))) int plh_timeline_read(unsigned long long* out) {
  return h_tl ? (int)cudaMemcpy(out, h_tl, 64 * 8, cudaMemcpyDeviceToHost) : -1;
}
#endif
extern "C" int plh_version(void) { return PLH_VERSION; }

extern "C" const char* plh_strerror(int code) {
  switch (code) {
    case PLH_OK: return "ok";
    case PLH_E_NULL: return "required pointer is NULL";
    case PLH_E_SHAPE: return "shape out of range";
    case PLH_E_ALIGN: return "pointer not 16-byte aligned";
    case PLH_E_WORKSPACE: return "workspace missing, misaligned or too small";
    case PLH_E_PARAM: return "bad parameter value";
    case PLH_E_DEVICE: return "no sm_100 CUDA device";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown plhead error";
  }
}

extern "C" long long plh_launch_count(void) { return g_launch_count.load(); }
