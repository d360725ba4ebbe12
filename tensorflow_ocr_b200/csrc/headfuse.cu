// headfuse.cu — the head's logit producer for sm_100a.
//
// Replaces the feature fusion that ends both networks: nets/pixellink.py:37-38,56-67 (PixelLink-4s over fc7,
// conv5_3, conv4_3, conv3_3) and nets/model.py:14-15,129-141 (the EAST fork over pool5..pool2): per level
//     y = up2(prev) + sum_f act_f(scale_f * (x_f W_f) + shift_f)          (1x1 convolutions, f = 1 or 2 features)
// and at the last level   logits = y W_out + b_out,   written as the [.,2] pixel and [.,16] link tensors the loss /
// decode kernels read.  The 2 pixel and 16 link channels go through together (18 columns, padded to 24).
//
// One launch per level.  A level is a skinny GEMM [pixels x K] x [K x 18] whose cost is reading the activations
// once (K * 4 B per pixel against 36 K flop: 9 flop/B, far below the machine balance), so the kernel is built to
// stream: persistent CTAs, tiles of 128 consecutive pixels, the K axis in 32-channel chunks through a 4-stage
// cp.async ring (activations 16 B at a time, the chunk's weight rows beside them), and the arithmetic on the
// tensor cores so that issue slots never limit the stream: mma.sync m16n8k8 TF32 with the 3xTF32 split
// (a = a_hi + a_lo, b = b_hi + b_lo, a_hi b_hi + a_hi b_lo + a_lo b_hi accumulated in fp32), which keeps fp32
// accuracy (dropped term ~2^-22 relative) — plain TF32 would miss the 1e-5 contract.  tcgen05 would need both
// operands of the split staged in shared memory; with an HBM-bound level there is nothing for it to win.
// The epilogue applies scale / shift / ReLU per feature, adds the bilinear x2 of the previous level
// (tf.image.resize_bilinear, align_corners = False: taps (y >> 1, x >> 1) and the next row / column, weight 0.5 on
// odd coordinates, clamped at the far edge), and the last level multiplies by the 18x18 output matrix out of
// shared memory and stores both tensors coalesced.
#include <algorithm>

#include "common.cuh"

namespace plh {

constexpr int kHfThreads = 256;                 // 8 warps x 16 pixels
constexpr int kHfTile = 128;                    // pixels per tile
constexpr int kHfKC = 32;                       // channels per chunk
constexpr int kHfStages = 4;
constexpr int kHfAStride = kHfKC + 4;           // floats per pixel row in shared memory (conflict-free fragments)
constexpr int kHfN = 18, kHfNP = 24;            // output columns, padded to three n8 tiles
constexpr int kHfStageFloats = kHfTile * kHfAStride + kHfKC * kHfNP;
constexpr size_t kHfSmem = (size_t)kHfStages * kHfStageFloats * 4 + (size_t)(kHfN * kHfN + kHfN + 8 * 16 * kHfNP) * 4;

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool pred) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int bytes = pred ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src, bool pred) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int bytes = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void split_tf32(float v, unsigned& hi, unsigned& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  const float r = v - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

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
};

__global__ void __launch_bounds__(kHfThreads, 2) head_fuse_kernel(const HfArgs a) {
  extern __shared__ __align__(16) float hf_smem[];
  float* s_wout = hf_smem + kHfStages * kHfStageFloats;  // [18][18]
  float* s_bout = s_wout + kHfN * kHfN;                  // [18]
  float* s_y = s_bout + kHfN;                            // [8 warps][16][24]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const long long total = (long long)a.B * a.H * a.W;
  const long long ntiles = (total + kHfTile - 1) / kHfTile;
  // the pad columns of the weight rows are never written by the copies: zero them once
  for (int i = tid; i < kHfStages * kHfKC * (kHfNP - kHfN); i += kHfThreads) {
    const int st = i / (kHfKC * (kHfNP - kHfN)), r = i % (kHfKC * (kHfNP - kHfN));
    hf_smem[st * kHfStageFloats + kHfTile * kHfAStride + (r / (kHfNP - kHfN)) * kHfNP + kHfN + r % (kHfNP - kHfN)] = 0.f;
  }
  if (a.w_out) {
    for (int i = tid; i < kHfN * kHfN; i += kHfThreads) s_wout[i] = a.w_out[i];
    if (tid < kHfN) s_bout[tid] = a.b_out ? a.b_out[tid] : 0.f;
  }
  __syncthreads();

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long px0 = tile * kHfTile;
    float y[3][4];
#pragma unroll
    for (int n = 0; n < 3; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) y[n][i] = 0.f;

    for (int fi = 0; fi < a.nf; ++fi) {
      const HfFeature F = a.f[fi];
      const int nchunks = (F.K + kHfKC - 1) / kHfKC;
      auto issue = [&](int c) {
        if (c < nchunks) {
          float* sA = hf_smem + (c % kHfStages) * kHfStageFloats;
          float* sW = sA + kHfTile * kHfAStride;
          const int k0 = c * kHfKC;
#pragma unroll
          for (int i = 0; i < (kHfTile * kHfKC / 4) / kHfThreads; ++i) {   // 4 copies of 16 B per thread
            const int idx = tid + i * kHfThreads, p = idx >> 3, q = idx & 7;
            const long long px = px0 + p;
            const bool ok = px < total && k0 + q * 4 < F.K;
            cp_async16(sA + p * kHfAStride + q * 4, F.x + (ok ? px * F.K + k0 + q * 4 : 0), ok);
          }
          for (int idx = tid; idx < kHfKC * kHfN; idx += kHfThreads) {
            const int r = idx / kHfN, cidx = idx - r * kHfN;
            const bool ok = k0 + r < F.K;
            cp_async4(sW + r * kHfNP + cidx, F.w + (ok ? (size_t)(k0 + r) * kHfN + cidx : 0), ok);
          }
        }
        cp_async_commit();
      };
      float c[3][4];
#pragma unroll
      for (int n = 0; n < 3; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[n][i] = 0.f;
      __syncthreads();  // the ring is free: the previous feature / tile has been consumed
      for (int s = 0; s < kHfStages - 1; ++s) issue(s);
      for (int ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<kHfStages - 2>();
        __syncthreads();             // chunk ch has landed for every thread; chunk ch-1's buffer is free
        issue(ch + kHfStages - 1);
        const float* sA = hf_smem + (ch % kHfStages) * kHfStageFloats + warp * 16 * kHfAStride;
        const float* sW = hf_smem + (ch % kHfStages) * kHfStageFloats + kHfTile * kHfAStride;
#pragma unroll
        for (int ks = 0; ks < kHfKC / 8; ++ks) {
          unsigned ah[4], al[4];
          split_tf32(sA[g * kHfAStride + ks * 8 + t], ah[0], al[0]);
          split_tf32(sA[(g + 8) * kHfAStride + ks * 8 + t], ah[1], al[1]);
          split_tf32(sA[g * kHfAStride + ks * 8 + t + 4], ah[2], al[2]);
          split_tf32(sA[(g + 8) * kHfAStride + ks * 8 + t + 4], ah[3], al[3]);
#pragma unroll
          for (int n = 0; n < 3; ++n) {
            unsigned bh0, bl0, bh1, bl1;
            split_tf32(sW[(ks * 8 + t) * kHfNP + n * 8 + g], bh0, bl0);
            split_tf32(sW[(ks * 8 + t + 4) * kHfNP + n * 8 + g], bh1, bl1);
            mma_tf32(c[n], al, bh0, bh1);
            mma_tf32(c[n], ah, bl0, bl1);
            mma_tf32(c[n], ah, bh0, bh1);
          }
        }
      }
      cp_async_wait<0>();
      // scale / shift / ReLU of this feature; c[n][0,1]: row g, columns n*8 + 2t, +1; c[n][2,3]: row g + 8
#pragma unroll
      for (int n = 0; n < 3; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int col = n * 8 + 2 * t + (i & 1);
          float v = c[n][i];
          if (col < kHfN) {
            if (F.scale) v = __fmul_rn(v, F.scale[col]);
            if (F.shift) v = __fadd_rn(v, F.shift[col]);
            if (F.relu) v = fmaxf(v, 0.f);
          }
          y[n][i] += v;
        }
    }

    // bilinear x2 of the previous level (TF: top + (bottom - top) * fy on rows interpolated the same way in x)
    if (a.prev) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long px = px0 + warp * 16 + g + 8 * h;
        if (px < total) {
          const int hw = a.H * a.W, b = (int)(px / hw), r = (int)(px - (long long)b * hw), yy = r / a.W, xx = r - yy * a.W;
          const int Hp = a.H >> 1, Wp = a.W >> 1;
          const int ylo = yy >> 1, yhi = min(ylo + 1, Hp - 1), xlo = xx >> 1, xhi = min(xlo + 1, Wp - 1);
          const float fy = (yy & 1) ? 0.5f : 0.f, fx = (xx & 1) ? 0.5f : 0.f;
          const float* P = a.prev + (size_t)b * Hp * Wp * kHfN;
#pragma unroll
          for (int n = 0; n < 3; ++n) {
            const int col = n * 8 + 2 * t;
            if (col < kHfN) {
              const float2 tl = *reinterpret_cast<const float2*>(P + ((size_t)ylo * Wp + xlo) * kHfN + col);
              const float2 tr = *reinterpret_cast<const float2*>(P + ((size_t)ylo * Wp + xhi) * kHfN + col);
              const float2 bl = *reinterpret_cast<const float2*>(P + ((size_t)yhi * Wp + xlo) * kHfN + col);
              const float2 br = *reinterpret_cast<const float2*>(P + ((size_t)yhi * Wp + xhi) * kHfN + col);
              const float top0 = __fadd_rn(tl.x, __fmul_rn(__fsub_rn(tr.x, tl.x), fx)), top1 = __fadd_rn(tl.y, __fmul_rn(__fsub_rn(tr.y, tl.y), fx));
              const float bot0 = __fadd_rn(bl.x, __fmul_rn(__fsub_rn(br.x, bl.x), fx)), bot1 = __fadd_rn(bl.y, __fmul_rn(__fsub_rn(br.y, bl.y), fx));
              y[n][2 * h] += __fadd_rn(top0, __fmul_rn(__fsub_rn(bot0, top0), fy));
              y[n][2 * h + 1] += __fadd_rn(top1, __fmul_rn(__fsub_rn(bot1, top1), fy));
            }
          }
        }
      }
    }

    // this warp's 16 x 24 result through shared memory: the stores (and the output matrix) want whole pixel rows
    float* sy = s_y + warp * 16 * kHfNP;
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      *reinterpret_cast<float2*>(sy + g * kHfNP + n * 8 + 2 * t) = make_float2(y[n][0], y[n][1]);
      *reinterpret_cast<float2*>(sy + (g + 8) * kHfNP + n * 8 + 2 * t) = make_float2(y[n][2], y[n][3]);
    }
    __syncwarp();
    const long long wpx0 = px0 + warp * 16;
    if (a.w_out) {
      // lane = (pixel, half): nine of the 18 outputs each
      const int p = lane >> 1, o0 = (lane & 1) * 9;
      float z[9];
#pragma unroll
      for (int o = 0; o < 9; ++o) z[o] = s_bout[o0 + o];
      for (int i = 0; i < kHfN; ++i) {
        const float v = sy[p * kHfNP + i];
#pragma unroll
        for (int o = 0; o < 9; ++o) z[o] = fmaf(v, s_wout[i * kHfN + o0 + o], z[o]);
      }
      __syncwarp();
#pragma unroll
      for (int o = 0; o < 9; ++o) sy[p * kHfNP + o0 + o] = z[o];
      __syncwarp();
      // pixel logits: 16 pixels x 2 floats = 32 consecutive floats; link logits: 16 x 16 = 64 float4
      if (wpx0 + (lane >> 1) < total) a.pix[wpx0 * 2 + lane] = sy[(lane >> 1) * kHfNP + (lane & 1)];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int q = lane + 32 * h, pp = q >> 2, c4 = (q & 3) * 4;
        if (wpx0 + pp < total) {
          const float* s = sy + pp * kHfNP + 2 + c4;
          *reinterpret_cast<float4*>(a.link + (wpx0 + pp) * 16 + c4) = make_float4(s[0], s[1], s[2], s[3]);
        }
      }
    } else {
      // [pixels, 18]: 16 pixels x 18 floats = 288 consecutive floats = 9 per lane
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int q = lane + 32 * i, pp = q / kHfN, cc = q - pp * kHfN;
        if (wpx0 + pp < total) a.y18[wpx0 * kHfN + q] = sy[pp * kHfNP + cc];
      }
    }
    __syncwarp();
  }
}

}  // namespace plh

using namespace plh;

extern "C" int plh_head_fuse_level(const float* xa, int Ka, const float* wa, const float* scale_a, const float* shift_a,
                                   int relu_a, const float* xb, int Kb, const float* wb, const float* scale_b,
                                   const float* shift_b, int relu_b, const float* prev, const float* w_out,
                                   const float* b_out, int B, int H, int W, float* y18, float* pix_logits,
                                   float* link_logits, void* stream) {
  if (!xa || !wa) return PLH_E_NULL;
  if (xb && !wb) return PLH_E_NULL;
  if (w_out ? (!pix_logits || !link_logits) : !y18) return PLH_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || (long long)B * H * W > (1ll << 31) - 1) return PLH_E_SHAPE;
  if (Ka <= 0 || (Ka & 3) || (xb && (Kb <= 0 || (Kb & 3)))) return PLH_E_SHAPE;   // 16-byte rows
  if (prev && ((H & 1) || (W & 1))) return PLH_E_SHAPE;                            // the level is exactly twice the previous one
  if (!aligned16(xa) || (xb && !aligned16(xb)) || (link_logits && !aligned16(link_logits))) return PLH_E_ALIGN;
  HfArgs a;
  a.f[0] = HfFeature{xa, wa, scale_a, shift_a, Ka, relu_a};
  a.f[1] = HfFeature{xb, wb, scale_b, shift_b, xb ? Kb : 0, relu_b};
  a.nf = xb ? 2 : 1;
  a.prev = prev, a.w_out = w_out, a.b_out = b_out, a.B = B, a.H = H, a.W = W;
  a.y18 = y18, a.pix = pix_logits, a.link = link_logits;
  static SmemOptIn optin;
  int rc;
  if ((rc = ensure_dynamic_smem(optin, head_fuse_kernel, kHfSmem))) return rc;
  const long long ntiles = ((long long)B * H * W + kHfTile - 1) / kHfTile;
  const int grid = (int)std::min<long long>(ntiles, 2ll * kNumSMs);
  head_fuse_kernel<<<grid, kHfThreads, kHfSmem, (cudaStream_t)stream>>>(a);
  return launch_status();
}
