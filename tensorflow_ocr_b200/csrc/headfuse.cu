// headfuse.cu — the head's logit producer for sm_100a.
//
// Replaces the feature fusion that ends both networks: nets/pixellink.py:37-38,56-67 (PixelLink-4s over fc7,
// conv5_3, conv4_3, conv3_3) and nets/model.py:14-15,129-141 (the EAST fork over pool5..pool2): per level
//     y = up2(prev) + sum_f act_f(scale_f * (x_f W_f) + shift_f)          (1x1 convolutions, f = 1 or 2 features)
// and at the last level   logits = y W_out + b_out,   written as the [.,2] pixel and [.,16] link tensors the loss /
// decode kernels read.  The 2 pixel and 16 link channels go through together (18 columns).
//
// One launch per level.  A level is a skinny GEMM [pixels x K] x [K x 18]: 18 FMA per activation read, 4.5 FMA per
// byte — on a B200 (36 T fp32 FMA/s against 6.5 TB/s) the stream of activations and the fp32 pipe are about
// equally loaded, so the kernel is built to keep both busy: persistent CTAs, tiles of 128 consecutive pixels, the
// K axis in 32-channel chunks through a 3-stage cp.async ring that runs ACROSS tile boundaries (activations 16 B
// at a time, the chunk's weight rows — contiguous in memory — beside them), and plain fp32 FMAs on a 4 pixel x 18
// output register tile per thread (72 independent accumulators; the four warps of a CTA split the chunk's
// channels, their partial sums meet in shared memory once per tile and feature).  Per 8 channels a thread
// issues 8 + 72 shared-memory loads (the weight ones are warp-wide broadcasts) for 576 FMAs.
// A first version ran the products on the tensor cores (mma.sync m16n8k8 TF32 with the 3xTF32 split, needed for
// the 1e-5 contract): on this part the legacy MMA path saturated (`math_pipe_throttle`) at 0.34 of the HBM
// roofline — three TF32 MMAs per fp32 product, padded from 18 to 24 columns, is MORE pipe time than the FMAs
// (profiles/r02_headfuse.txt).  tcgen05 would need both halves of the split staged in shared memory in the
// canonical layout; not attempted.
// The epilogue (one thread per pixel) sums the partials, applies scale / shift / ReLU per feature, adds the
// bilinear x2 of the previous level (tf.image.resize_bilinear, align_corners = False: taps (y >> 1, x >> 1) and
// the next row / column, weight 0.5 on odd coordinates, clamped at the far edge), and the last level multiplies
// by the 18x18 output matrix and stores both tensors.
#include <algorithm>

#include "common.cuh"

namespace plh {

float prob_to_logit_threshold(float t);  // loss.cu

constexpr int kHfThreads = 128;                 // 4 warps: each takes 8 of a chunk's 32 channels
constexpr int kHfTile = 128;                    // pixels per tile: 4 per lane
constexpr int kHfKC = 32;                       // channels per chunk
constexpr int kHfStages = 3;
constexpr int kHfAStride = kHfKC + 4;           // floats per pixel row in shared memory (conflict-free 128-bit reads)
constexpr int kHfN = 18;                        // output columns
constexpr int kHfWFloats = kHfKC * kHfN;        // a chunk's weight rows as they lie in memory
constexpr int kHfStageFloats = kHfTile * kHfAStride + kHfWFloats;
constexpr int kHfPartial = 4 * kHfTile * (kHfN + 1);          // [warp][pixel][18 (+1: conflict-free column reads)]
constexpr int kHfMisc = kHfN * kHfN + kHfN + 4 * kHfN;        // output matrix, its bias, scale / shift of two features
constexpr size_t kHfSmem = (size_t)(kHfStages * kHfStageFloats + kHfPartial + kHfMisc + 2) * 4;

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int bytes) {  // bytes < 16: the rest is zero-filled
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct HfFeature {
  const float* x;      // [pixels, K]
  const float* w;      // [K, 18]
  const float* scale;  // [18] or null
  const float* shift;  // [18] or null
  int K, relu;
};

struct HfArgs {
  HfFeature f[2];
  int nf;
  const float* prev;   // [B, H/2, W/2, 18] or null
  const float* w_out;  // [18, 18] (in, out) or null
  const float* b_out;  // [18] or null
  int B, H, W;
  float* y18;          // [pixels, 18]  (levels before the last)
  float* pix;          // [pixels, 2]   (last level)
  float* link;         // [pixels, 16]
  uint16_t* flags;     // [pixels] or null: the decode's threshold word of the pixel just produced
  float tp_logit, tl_logit;
};

__global__ void __launch_bounds__(kHfThreads, 2) head_fuse_kernel(const HfArgs a) {
  extern __shared__ __align__(16) float hf_smem[];
  float* s_part = hf_smem + kHfStages * kHfStageFloats;  // [4][128][19]
  float* s_wout = s_part + kHfPartial;                   // [18][18] (8-byte aligned: kHfPartial is even)
  float* s_bout = s_wout + kHfN * kHfN;                  // [18]
  float* s_aff = s_bout + kHfN;                          // [feature][scale | shift][18]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long total = (long long)a.B * a.H * a.W;
  const long long ntiles = (total + kHfTile - 1) / kHfTile;
  for (int i = tid; i < kHfN * kHfN; i += kHfThreads) s_wout[i] = a.w_out ? a.w_out[i] : 0.f;
  if (tid < kHfN) s_bout[tid] = a.b_out ? a.b_out[tid] : 0.f;
  if (tid < 4 * kHfN) {
    const int fi = tid / (2 * kHfN), which = (tid / kHfN) & 1, c = tid % kHfN;
    const float* src = which ? a.f[fi].shift : a.f[fi].scale;
    s_aff[tid] = (fi < a.nf && src) ? src[c] : (which ? 0.f : 1.f);
  }
  __syncthreads();

  // The chunk stream: (tile, feature, 32-channel chunk) in the order they are consumed, issued kHfStages - 1
  // chunks ahead ACROSS tile boundaries, so the epilogue of a tile runs under the loads of the next one.
  long long i_tile = blockIdx.x;
  int i_f = 0, i_ch = 0, i_slot = 0;
  auto issue = [&]() {
    if (i_tile < ntiles) {
      const HfFeature& F = a.f[i_f];
      float* sA = hf_smem + i_slot * kHfStageFloats;
      float* sW = sA + kHfTile * kHfAStride;
      const int k0 = i_ch * kHfKC;
      const float* src = F.x + (i_tile * kHfTile) * F.K + k0;
      const int rows = (int)min((long long)kHfTile, total - i_tile * kHfTile);
#pragma unroll
      for (int i = 0; i < (kHfTile * kHfKC / 4) / kHfThreads; ++i) {   // activations: 8 copies of 16 B per thread
        const int idx = tid + i * kHfThreads, p = idx >> 3, q = idx & 7;
        const bool ok = p < rows && k0 + q * 4 < F.K;
        cp_async16(sA + p * kHfAStride + q * 4, ok ? src + (size_t)p * F.K + q * 4 : F.x, ok ? 16 : 0);
      }
      for (int i = tid; i < kHfWFloats / 4; i += kHfThreads) {         // weights: the chunk's rows are contiguous (2304 B)
        const int nb = max(min(min(kHfKC, F.K - k0) * kHfN * 4 - i * 16, 16), 0);
        cp_async16(sW + i * 4, nb ? F.w + (size_t)k0 * kHfN + i * 4 : F.w, nb);
      }
      const int nch = (F.K + kHfKC - 1) / kHfKC;
      if (++i_ch == nch) {
        i_ch = 0;
        if (++i_f == a.nf) i_f = 0, i_tile += gridDim.x;
      }
    }
    cp_async_commit();
    i_slot = i_slot + 1 == kHfStages ? 0 : i_slot + 1;
  };
  for (int s = 0; s < kHfStages - 1; ++s) issue();
  int c_slot = 0;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long px = tile * kHfTile + tid;   // the pixel this thread finishes in the epilogue
    float y[kHfN];
#pragma unroll
    for (int o = 0; o < kHfN; ++o) y[o] = 0.f;

    for (int fi = 0; fi < a.nf; ++fi) {
      const int nchunks = (a.f[fi].K + kHfKC - 1) / kHfKC;
      float acc[4][kHfN];   // pixels lane, lane + 32, lane + 64, lane + 96; this warp's 8 channels of every chunk
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int o = 0; o < kHfN; ++o) acc[j][o] = 0.f;
      for (int ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<kHfStages - 2>();
        __syncthreads();             // this chunk has landed for every thread; the slot consumed last is free
        issue();
        const float* sA = hf_smem + c_slot * kHfStageFloats + 8 * warp;
        const float* sW = hf_smem + c_slot * kHfStageFloats + kHfTile * kHfAStride + 8 * warp * kHfN;
        c_slot = c_slot + 1 == kHfStages ? 0 : c_slot + 1;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float4 av[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) av[j] = *reinterpret_cast<const float4*>(sA + (lane + 32 * j) * kHfAStride + 4 * half);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            float w[kHfN];
#pragma unroll
            for (int o = 0; o < kHfN; o += 2) {
              const float2 t2 = *reinterpret_cast<const float2*>(sW + (4 * half + kk) * kHfN + o);   // warp-wide broadcast
              w[o] = t2.x, w[o + 1] = t2.y;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float x = kk == 0 ? av[j].x : (kk == 1 ? av[j].y : (kk == 2 ? av[j].z : av[j].w));
#pragma unroll
              for (int o = 0; o < kHfN; ++o) acc[j][o] = fmaf(x, w[o], acc[j][o]);
            }
          }
        }
      }
      // the four warps' partial sums meet in shared memory; thread p finishes pixel p
      __syncthreads();   // (the partial buffer of the previous feature / tile has been read)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int o = 0; o < kHfN; ++o) s_part[(warp * kHfTile + lane + 32 * j) * (kHfN + 1) + o] = acc[j][o];
      __syncthreads();
      const float* sc = s_aff + fi * 2 * kHfN;
      const bool relu = a.f[fi].relu != 0;
#pragma unroll
      for (int o = 0; o < kHfN; ++o) {
        float v = 0.f;
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) v += s_part[(w4 * kHfTile + tid) * (kHfN + 1) + o];
        v = __fadd_rn(__fmul_rn(v, sc[o]), sc[kHfN + o]);
        if (relu) v = fmaxf(v, 0.f);
        y[o] += v;
      }
    }

    if (px < total) {
      // bilinear x2 of the previous level (TF: top + (bottom - top) * fy on rows interpolated the same way in x)
      if (a.prev) {
        const int hw = a.H * a.W, Hp = a.H >> 1, Wp = a.W >> 1;
        const int b = (int)(px / hw), r = (int)(px - (long long)b * hw), yy = r / a.W, xx = r - yy * a.W;
        const int ylo = yy >> 1, yhi = min(ylo + 1, Hp - 1), xlo = xx >> 1, xhi = min(xlo + 1, Wp - 1);
        const float fy = (yy & 1) ? 0.5f : 0.f, fx = (xx & 1) ? 0.5f : 0.f;
        const float* P = a.prev + (size_t)b * Hp * Wp * kHfN;
        const float2* tl = reinterpret_cast<const float2*>(P + ((size_t)ylo * Wp + xlo) * kHfN);
        const float2* tr = reinterpret_cast<const float2*>(P + ((size_t)ylo * Wp + xhi) * kHfN);
        const float2* bl = reinterpret_cast<const float2*>(P + ((size_t)yhi * Wp + xlo) * kHfN);
        const float2* br = reinterpret_cast<const float2*>(P + ((size_t)yhi * Wp + xhi) * kHfN);
#pragma unroll
        for (int o = 0; o < kHfN / 2; ++o) {
          const float2 q0 = tl[o], q1 = tr[o], q2 = bl[o], q3 = br[o];
          const float top0 = __fadd_rn(q0.x, __fmul_rn(__fsub_rn(q1.x, q0.x), fx)), top1 = __fadd_rn(q0.y, __fmul_rn(__fsub_rn(q1.y, q0.y), fx));
          const float bot0 = __fadd_rn(q2.x, __fmul_rn(__fsub_rn(q3.x, q2.x), fx)), bot1 = __fadd_rn(q2.y, __fmul_rn(__fsub_rn(q3.y, q2.y), fx));
          y[2 * o] += __fadd_rn(top0, __fmul_rn(__fsub_rn(bot0, top0), fy));
          y[2 * o + 1] += __fadd_rn(top1, __fmul_rn(__fsub_rn(bot1, top1), fy));
        }
      }
      if (a.w_out) {
        float z[kHfN];
#pragma unroll
        for (int o = 0; o < kHfN; ++o) z[o] = s_bout[o];
#pragma unroll
        for (int i = 0; i < kHfN; ++i)
#pragma unroll
          for (int o = 0; o < kHfN; o += 2) {
            const float2 t2 = *reinterpret_cast<const float2*>(s_wout + i * kHfN + o);
            z[o] = fmaf(y[i], t2.x, z[o]), z[o + 1] = fmaf(y[i], t2.y, z[o + 1]);
          }
        *reinterpret_cast<float2*>(a.pix + px * 2) = make_float2(z[0], z[1]);
        if (a.flags) {   // decode_flags_kernel's word (decode.cu): bit d = link d passes, bit 8 = the pixel passes, in logit space
          unsigned f = (z[1] - z[0]) > a.tp_logit ? 256u : 0u;
#pragma unroll
          for (int d = 0; d < 8; ++d) f |= ((z[3 + 2 * d] - z[2 + 2 * d]) > a.tl_logit ? 1u : 0u) << d;
          a.flags[px] = (uint16_t)f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(a.link + px * 16 + 4 * q) = make_float4(z[2 + 4 * q], z[3 + 4 * q], z[4 + 4 * q], z[5 + 4 * q]);
      } else {
#pragma unroll
        for (int o = 0; o < kHfN; o += 2) *reinterpret_cast<float2*>(a.y18 + px * kHfN + o) = make_float2(y[o], y[o + 1]);
      }
    }
  }
  cp_async_wait<0>();
}

}  // namespace plh

using namespace plh;

extern "C" int plh_head_fuse_level(const float* xa, int Ka, const float* wa, const float* scale_a, const float* shift_a,
                                   int relu_a, const float* xb, int Kb, const float* wb, const float* scale_b,
                                   const float* shift_b, int relu_b, const float* prev, const float* w_out,
                                   const float* b_out, int B, int H, int W, float* y18, float* pix_logits,
                                   float* link_logits, const plh_decode_params* flag_params, uint16_t* flags,
                                   void* stream) {
  if (!xa || !wa) return PLH_E_NULL;
  if (xb && !wb) return PLH_E_NULL;
  if (w_out ? (!pix_logits || !link_logits) : !y18) return PLH_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || (long long)B * H * W > (1ll << 31) - 1) return PLH_E_SHAPE;
  if (Ka <= 0 || (Ka & 3) || (xb && (Kb <= 0 || (Kb & 3)))) return PLH_E_SHAPE;   // 16-byte rows
  if (prev && ((H & 1) || (W & 1))) return PLH_E_SHAPE;                            // the level is exactly twice the previous one
  if (!aligned16(xa) || !aligned16(wa) || (xb && (!aligned16(xb) || !aligned16(wb))) || (link_logits && !aligned16(link_logits)))
    return PLH_E_ALIGN;
  HfArgs a;
  a.f[0] = HfFeature{xa, wa, scale_a, shift_a, Ka, relu_a};
  a.f[1] = HfFeature{xb, wb, scale_b, shift_b, xb ? Kb : 0, relu_b};
  a.nf = xb ? 2 : 1;
  a.prev = prev, a.w_out = w_out, a.b_out = b_out, a.B = B, a.H = H, a.W = W;
  a.y18 = y18, a.pix = pix_logits, a.link = link_logits;
  a.flags = nullptr, a.tp_logit = a.tl_logit = 0.f;
  if (flags) {
    if (!w_out || !flag_params) return PLH_E_NULL;   // the threshold word belongs to the logits of the last level
    a.flags = flags;
    a.tp_logit = prob_to_logit_threshold(flag_params->pixel_thresh);
    a.tl_logit = prob_to_logit_threshold(flag_params->link_thresh);
  }
  static SmemOptIn optin;
  int rc;
  if ((rc = ensure_dynamic_smem(optin, head_fuse_kernel, kHfSmem))) return rc;
  const long long ntiles = ((long long)B * H * W + kHfTile - 1) / kHfTile;
  const int grid = (int)std::min<long long>(ntiles, 2ll * kNumSMs);   // two resident CTAs per SM
  head_fuse_kernel<<<grid, kHfThreads, kHfSmem, (cudaStream_t)stream>>>(a);
  return launch_status();
}
