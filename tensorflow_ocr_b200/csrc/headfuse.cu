// headfuse.cu — the head's logit producer for sm_100a.
//
// Replaces the feature fusion that ends both networks: nets/pixellink.py:37-38,56-67 (PixelLink-4s over fc7,
// conv5_3, conv4_3, conv3_3) and nets/model.py:14-15,129-141 (the EAST fork over pool5..pool2): per level
//     y = up2(prev) + sum_f act_f(scale_f * (x_f W_f) + shift_f)          (1x1 convolutions, f = 1 or 2 features)
// and at the last level   logits = y W_out + b_out,   written as the [.,2] pixel and [.,16] link tensors the loss /
// decode kernels read.  The 2 pixel and 16 link channels go through together (18 columns).
//
// One launch per level.  A level is a skinny GEMM [pixels x K] x [K x 18]: 18 FMA per activation read, 4.5 FMA per
// byte — on a B200 (36 T fp32 FMA/s against 6.5 TB/s) plain fp32 FMAs would load the fp32 pipe as much as the
// stream loads HBM.  Two kernels:
//  * head_fuse_tc_kernel (K % 32 == 0, i.e. every real backbone): the products run on the 5th-generation tensor
//    cores — tcgen05.mma kind::tf32 with the accumulator in tensor memory — with the 3xTF32 split that the 1e-5
//    contract needs done while the operands are staged into the swizzled shared-memory tiles; see its comment.
//  * head_fuse_kernel (any K % 4 == 0): fp32 FMAs on a 4 pixel x 18 output register tile per thread, persistent
//    CTAs, 32-channel chunks through a 3-stage cp.async ring that runs across tile boundaries, the four warps of a
//    CTA splitting a chunk's channels (partial sums meet in shared memory once per tile and feature).
// An earlier tensor-core version on the legacy path (mma.sync m16n8k8 TF32, 3xTF32) saturated that pipe
// (`math_pipe_throttle`) at 0.34 of the HBM roofline and was dropped (profiles/r02_headfuse.txt).
// The epilogue (one thread per pixel, shared by both kernels) applies scale / shift / ReLU per feature, adds the
// bilinear x2 of the previous level (tf.image.resize_bilinear, align_corners = False: taps (y >> 1, x >> 1) and
// the next row / column, weight 0.5 on odd coordinates, clamped at the far edge), and the last level multiplies
// by the 18x18 output matrix and stores both tensors (and, on request, the decode's threshold word).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace plh {

// -DPLH_HF_TRACE (make -C tensorflow_ocr_b200/csrc trace -> libplhead_trace.so): CTA 0 of the tensor-core kernel records
// %globaltimer at the pipeline's hand-overs for its first 512 chunks; tools/headfuse_trace.py prints them.
#ifdef PLH_HF_TRACE
__device__ unsigned long long g_hf_trace[8 * 512];
__device__ __forceinline__ unsigned long long hf_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define HF_TR(slot, n)                                                                 \
  do {                                                                                 \
    if (blockIdx.x == 0 && (n) < 512) g_hf_trace[(slot) * 512 + (n)] = hf_now();       \
  } while (0)
#else
#define HF_TR(slot, n) do { } while (0)
#endif

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

// per-pixel tail shared by both kernels: bilinear x2 of the previous level, output matrix, stores
__device__ __forceinline__ void hf_pixel_epilogue(const HfArgs& a, long long px, float (&y)[kHfN], const float* s_wout, const float* s_bout) {
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
  float z[kHfN];
  if (a.w_out) {
#pragma unroll
    for (int o = 0; o < kHfN; ++o) z[o] = s_bout[o];
#pragma unroll
    for (int i = 0; i < kHfN; ++i)
#pragma unroll
      for (int o = 0; o < kHfN; o += 2) {
        const float2 t2 = *reinterpret_cast<const float2*>(s_wout + i * kHfN + o);
        z[o] = fmaf(y[i], t2.x, z[o]), z[o + 1] = fmaf(y[i], t2.y, z[o + 1]);
      }
  } else {
#pragma unroll
    for (int o = 0; o < kHfN; ++o) z[o] = y[o];
  }
  if (a.pix) {   // the two logit tensors
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
  } else {       // [pixels, 18] for the next level
#pragma unroll
    for (int o = 0; o < kHfN; o += 2) *reinterpret_cast<float2*>(a.y18 + px * kHfN + o) = make_float2(z[o], z[o + 1]);
  }
}

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

    if (px < total) hf_pixel_epilogue(a, px, y, s_wout, s_bout);
  }
  cp_async_wait<0>();
}


// ------------------------------------------------------------------ the same level on the 5th-generation tensor cores
// tcgen05.mma kind::tf32, M = 128 (one tile of pixels), N = 32 (18 outputs, zero padded), K = 8 per instruction,
// accumulator in tensor memory (32 columns).  fp32 accuracy comes from the 3xTF32 split done while the operands are
// staged: a thread loads 16 bytes of activations from global memory into registers (coalesced, one chunk ahead),
// splits every value into hi = the top 19 bits (an exact TF32 number) and lo = v - hi (exact in fp32; the tensor
// core reads its top 19 bits), and writes both into shared memory in the canonical K-major SWIZZLE_128B layout the
// matrix descriptors name (a 32-channel chunk = one 128-byte row per pixel, 16-byte pieces XOR-ed with the row index
// modulo 8, 1024 bytes per group of 8 rows); the weights of the chunk go the same way, transposed to [output][channel].
// A fifth warp issues, per 8 channels, hi*[hi|lo] (one N = 64 MMA: the two weight tiles lie back to back) and lo*hi
// into the accumulator (64 columns; the epilogue adds the halves) and commits the stage's
// mbarrier; two operand stages and mbarriers both ways (operands ready: 128 arrivals; stage free / accumulator ready:
// tcgen05.commit; accumulator read: 128 arrivals), so the four staging warps never wait for the issue itself.  When
// the last chunk of a (tile, feature) is committed they wait for it, read their accumulator row (tcgen05.ld, lane =
// pixel) and run the same per-pixel epilogue as the FMA kernel.  Needs K % 32 == 0; other shapes take the FMA kernel.
// What bounds it (profiles/r02_headfuse.txt): SHARED-MEMORY bandwidth.  Per 16 KB chunk of activations the operand
// tiles are written once (hi + lo: 32 KB) and read by the tensor core once per MMA that uses them — the A operand is
// re-read by every MMA, whatever N is, which is why hi*hi and hi*lo share one MMA: 32 KB instead of 48 KB.  64 KB per
// chunk = 0.26 us at 128 B/clk, against 0.37 us for the chunk's HBM time (80 KB / 0.33 us with three MMAs).  A fully warp-specialised variant (cp.async loaders writing the swizzled layout directly so
// that the raw tile serves as the hi operand, converters producing only lo, two accumulators, separate epilogue
// warps) was built and is bit-compatible, but moves 96 KB per chunk through shared memory and measures the same
// (314 vs 305 us).  Feeding the A operand from TENSOR memory instead (staging warps transpose the tile through shared
// memory to thread = pixel, tcgen05.st of hi and lo into two A stages, MMAs with [a_tmem]) was built too and is
// bit-compatible: it takes the operand fetch off the tensor core but adds a barrier, 8 LDS and 2 tcgen05.st per chunk to
// the staging warps, which are the longer pole — 166 us on the conv3_3 level against 146 us; not kept.
constexpr int kTcThreads = 160;                  // 4 staging / epilogue warps + 1 warp that issues the MMAs
constexpr int kTcWorkers = 128;
constexpr int kTcTile = 128;
constexpr int kTcKC = 32;
constexpr int kTcNP = 32;                        // N of the MMA
constexpr int kTcABytes = kTcTile * kTcKC * 4;   // 16 KB: one operand tile of activations
constexpr int kTcBBytes = kTcNP * kTcKC * 4;     // 4 KB
constexpr int kTcStageBytes = 2 * kTcABytes + 2 * kTcBBytes;   // hi + lo of both operands
constexpr size_t kTcSmem = 1024 + 2 * kTcStageBytes + (kHfN * kHfN + kHfN + 4 * kHfN + 2) * 4 + 64;   // + 6 mbarriers

__device__ __forceinline__ unsigned long long tc_smem_desc(uint32_t saddr) {
  // start address [0,14) (>>4), leading byte offset [16,30) unused for a swizzled K-major tile, stride byte offset
  // [32,46) = 1024 B between groups of 8 rows, descriptor version 1 [46,48), layout SWIZZLE_128B = 2 at [61,64)
  return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((unsigned long long)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, unsigned long long da, unsigned long long db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra TC_DONE;\n"
      "bra TC_WAIT;\n"
      "TC_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(0x989680)
      : "memory");
}

__global__ void __launch_bounds__(kTcThreads, 2) head_fuse_tc_kernel(const HfArgs a) {
  extern __shared__ __align__(16) unsigned char tc_raw[];
  const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(tc_raw);
  unsigned char* base = tc_raw + ((1024u - (raw_s & 1023u)) & 1023u);     // SWIZZLE_128B tiles want 1024-byte alignment
  float* s_wout = reinterpret_cast<float*>(base + 2 * kTcStageBytes);
  float* s_bout = s_wout + kHfN * kHfN;
  float* s_aff = s_bout + kHfN;
  // mbarriers: [0,1] stage free (commit), [2] accumulator ready (commit), [3,4] operands ready (128), [5] accumulator read (128)
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_aff + 4 * kHfN + 2);
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const long long total = (long long)a.B * a.H * a.W;
  const long long ntiles = (total + kTcTile - 1) / kTcTile;
  for (int i = tid; i < kHfN * kHfN; i += kTcThreads) s_wout[i] = a.w_out ? a.w_out[i] : 0.f;
  if (tid < kHfN) s_bout[tid] = a.b_out ? a.b_out[tid] : 0.f;
  if (tid < 4 * kHfN) {
    const int fi = tid / (2 * kHfN), which = (tid / kHfN) & 1, c = tid % kHfN;
    const float* src = which ? a.f[fi].shift : a.f[fi].scale;
    s_aff[tid] = (fi < a.nf && src) ? src[c] : (which ? 0.f : 1.f);
  }
  // rows 18..31 of the weight tiles stay zero: clear both stages' B tiles once
  for (int i = tid; i < 2 * 2 * kTcBBytes / 16; i += kTcThreads) {
    const int st = i / (2 * kTcBBytes / 16), r = i % (2 * kTcBBytes / 16);
    *reinterpret_cast<float4*>(base + st * kTcStageBytes + 2 * kTcABytes + r * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tid == 0) {
    for (int i = 0; i < 6; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(s_bar + i)), "r"(i < 3 ? 1 : kTcWorkers));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the cleared weight rows
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(s_bar);
  const uint32_t base_s = (uint32_t)__cvta_generic_to_shared(base);
  // c_format F32 (bits 4-5 = 1), a / b format TF32 (2 at bits 7-9 and 10-12), both K-major, N >> 3 at 17, M >> 4 at 24
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcNP >> 3) << 17) | ((uint32_t)(kTcTile >> 4) << 24);
  const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * kTcNP >> 3) << 17) | ((uint32_t)(kTcTile >> 4) << 24);   // N = 64

  if (warp == 4) {
    // ---- the issuing warp: same (tile, feature, chunk) sequence as the staging warps
    uint32_t n_chunk = 0, n_acc = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      // DRAM -> L2 ahead of the staging warps: the activation block of this CTA's NEXT tile is one contiguous
      // range per feature (128 pixels x K floats); one bulk prefetch each.  The register-staged loads then find
      // their lines in L2, which takes the DRAM latency out of the 64 KB of loads an SM can keep outstanding.
      if ((tid & 31) == 0) {
        const long long nt = tile + gridDim.x;
        if (nt < ntiles)
          for (int fi = 0; fi < a.nf; ++fi) {
            const long long rows = min((long long)kTcTile, total - nt * kTcTile);
            const size_t bytes = (size_t)rows * a.f[fi].K * 4;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.f[fi].x + (size_t)nt * kTcTile * a.f[fi].K), "r"((uint32_t)bytes) : "memory");
          }
      }
      for (int fi = 0; fi < a.nf; ++fi) {
        const int nchunks = a.f[fi].K / kTcKC;
        if (n_acc >= 1) tc_mbar_wait(bar0 + 8 * 5, (n_acc - 1) & 1);   // the previous accumulator has been read
        for (int ch = 0; ch < nchunks; ++ch, ++n_chunk) {
          const int st = n_chunk & 1;
          if ((tid & 31) == 0) HF_TR(5, n_chunk);
          tc_mbar_wait(bar0 + 8 * (3 + st), (n_chunk >> 1) & 1);       // operands staged (and fenced) by all 128 threads
          if ((tid & 31) == 0) HF_TR(6, n_chunk);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if ((tid & 31) == 0) {
            const uint32_t sa_hi = base_s + st * kTcStageBytes, sa_lo = sa_hi + kTcABytes, sb_hi = sa_lo + kTcABytes;   // B_lo follows B_hi
#pragma unroll
            for (int j = 0; j < kTcKC / 8; ++j) {
              // the hi and lo weight tiles lie back to back: one N = 64 MMA gives hi*hi (columns 0-31) and hi*lo (columns
              // 32-63) for a single read of the hi activations; lo*hi goes into columns 0-31.  The epilogue adds the halves.
              const unsigned long long dah = tc_smem_desc(sa_hi + 32 * j), dal = tc_smem_desc(sa_lo + 32 * j);
              const unsigned long long dbh = tc_smem_desc(sb_hi + 32 * j);
              tc_mma(tmem, dah, dbh, idesc2, (ch | j) != 0);
              tc_mma(tmem, dal, dbh, idesc, 1);
            }
            tc_commit(bar0 + 8 * st);
            if (ch == nchunks - 1) tc_commit(bar0 + 16);
            HF_TR(7, n_chunk);
          }
          __syncwarp();
        }
        ++n_acc;
      }
    }
  } else {
  // this thread's share of a chunk: rows r0 + 16 i of the tile, 16-byte piece c of their 128 bytes; weights e, e + 128, ...
  // Two chunks are kept in flight in registers (the loads of chunk n + 2 are issued when chunk n has been staged),
  // following a cursor over the (tile, feature, chunk) stream that runs across tile boundaries.
  const int r0 = tid >> 3, cpiece = tid & 7;
  float4 av0[8], av1[8];
  float wv0[5], wv1[5];
  long long c_tile = blockIdx.x;
  int c_f = 0, c_ch = 0, c_rows = 0, c_nch = 1;
  const float* c_ptr = nullptr;   // this thread's 16 bytes of row r0 of the cursor's (tile, feature), chunk 0
  const float* c_w = nullptr;
  size_t c_stride = 0;            // 16 rows further
  auto retarget = [&]() {
    if (c_tile >= ntiles) return;
    const HfFeature& F = a.f[c_f];
    c_ptr = F.x + (size_t)(c_tile * kTcTile + r0) * F.K + 4 * cpiece;
    c_w = F.w;
    c_stride = (size_t)16 * F.K;
    c_rows = (int)min((long long)kTcTile, total - c_tile * kTcTile) - r0;   // row r0 + 16 i exists iff 16 i < c_rows
    c_nch = F.K / kTcKC;
  };
  retarget();
  auto load_next = [&](float4 (&av)[8], float (&wv)[5]) {
    if (c_tile >= ntiles) return;
    const float* p = c_ptr + c_ch * kTcKC;
#pragma unroll
    for (int i = 0; i < 8; ++i, p += c_stride)
      av[i] = 16 * i < c_rows ? ldg_stream4(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
    // weight element e = tid + 128 i is (channel e % 32, output e / 32): a warp then writes one 128-byte row of the
    // transposed tile (conflict-free), and reads with stride 18
    const float* q = c_w + (size_t)c_ch * (kTcKC * kHfN);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int e = tid + i * kTcWorkers;
      wv[i] = e < kTcKC * kHfN ? __ldg(q + (e & 31) * kHfN + (e >> 5)) : 0.f;
    }
    if (++c_ch == c_nch) {
      c_ch = 0;
      if (++c_f == a.nf) c_f = 0, c_tile += gridDim.x;
      retarget();
    }
  };
  auto split = [](float v, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    lo = v - hi;
  };
  // where this thread's pieces go in the swizzled tiles (the same in every chunk): rows r0 + 16 i keep r & 7, so the
  // activation offsets are a_off + 2048 i; weight element e = (k, n) -> row n, 16-byte piece k >> 2 XOR-ed with n & 7
  const int a_off = r0 * 128 + ((cpiece ^ (r0 & 7)) << 4);
  int b_off[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int e = tid + i * kTcWorkers, k = e & 31, n = e >> 5;
    b_off[i] = n * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4;
  }
  auto store_chunk = [&](int st, const float4 (&av)[8], const float (&wv)[5]) {
    unsigned char* A_hi = base + st * kTcStageBytes;
    unsigned char* A_lo = A_hi + kTcABytes;
    unsigned char* B_hi = A_lo + kTcABytes;
    unsigned char* B_lo = B_hi + kTcBBytes;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int off = a_off + 2048 * i;
      float4 h, l;
      split(av[i].x, h.x, l.x), split(av[i].y, h.y, l.y), split(av[i].z, h.z, l.z), split(av[i].w, h.w, l.w);
      *reinterpret_cast<float4*>(A_hi + off) = h;
      *reinterpret_cast<float4*>(A_lo + off) = l;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int e = tid + i * kTcWorkers;
      if (e < kTcKC * kHfN) {
        const int off = b_off[i];
        float h, l;
        split(wv[i], h, l);
        *reinterpret_cast<float*>(B_hi + off) = h;
        *reinterpret_cast<float*>(B_lo + off) = l;
      }
    }
  };

  uint32_t n_chunk = 0, n_acc = 0;   // chunks staged / accumulators finished by this CTA so far
  load_next(av0, wv0);
  load_next(av1, wv1);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long px = tile * kTcTile + tid;
    float y[kHfN];
#pragma unroll
    for (int o = 0; o < kHfN; ++o) y[o] = 0.f;
    for (int fi = 0; fi < a.nf; ++fi) {
      const int nchunks = a.f[fi].K / kTcKC;
      for (int ch = 0; ch < nchunks; ++ch, ++n_chunk) {
        const int st = n_chunk & 1;
        // the MMAs that read this stage two chunks ago are done (use u of a stage waits for completion u - 1)
        if (tid == 0) HF_TR(0, n_chunk);
        if (n_chunk >= 2) tc_mbar_wait(bar0 + 8 * st, ((n_chunk >> 1) - 1) & 1);
        if (tid == 0) HF_TR(1, n_chunk);
        if (st == 0) {
          store_chunk(0, av0, wv0);
          if (tid == 0) HF_TR(2, n_chunk);
          load_next(av0, wv0);
        } else {
          store_chunk(1, av1, wv1);
          if (tid == 0) HF_TR(2, n_chunk);
          load_next(av1, wv1);
        }
        if (tid == 0) HF_TR(3, n_chunk);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8 * (3 + st)) : "memory");
        if (tid == 0) HF_TR(4, n_chunk);
      }
      // the accumulator of this (tile, feature): lane = pixel, 32 columns (18 used)
      tc_mbar_wait(bar0 + 16, n_acc & 1);
      ++n_acc;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // 18 of the 32 columns of each half: x16 + x2 loads (columns 0-17 hold hi*hi + lo*hi, 32-49 hold hi*lo)
      uint32_t v[kHfN], v2[kHfN];
      const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
#define PLH_TLD16(arr, col)                                                                                                  \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"        \
               : "=r"(arr[0]), "=r"(arr[1]), "=r"(arr[2]), "=r"(arr[3]), "=r"(arr[4]), "=r"(arr[5]), "=r"(arr[6]), "=r"(arr[7]), \
                 "=r"(arr[8]), "=r"(arr[9]), "=r"(arr[10]), "=r"(arr[11]), "=r"(arr[12]), "=r"(arr[13]), "=r"(arr[14]),         \
                 "=r"(arr[15])                                                                                                 \
               : "r"(trow + (col)))
#define PLH_TLD2(arr, col) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(arr[16]), "=r"(arr[17]) : "r"(trow + (col) + 16))
      PLH_TLD16(v, 0);
      PLH_TLD2(v, 0);
      PLH_TLD16(v2, 32);
      PLH_TLD2(v2, 32);
#undef PLH_TLD16
#undef PLH_TLD2
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8 * 5) : "memory");   // the next accumulation may overwrite it
#pragma unroll
      for (int o = 0; o < kHfN; ++o) v[o] = __float_as_uint(__uint_as_float(v[o]) + __uint_as_float(v2[o]));
      const float* sc = s_aff + fi * 2 * kHfN;
      const bool relu = a.f[fi].relu != 0;
#pragma unroll
      for (int o = 0; o < kHfN; ++o) {
        float m = __fadd_rn(__fmul_rn(__uint_as_float(v[o]), sc[o]), sc[kHfN + o]);
        if (relu) m = fmaxf(m, 0.f);
        y[o] += m;
      }
    }
    if (px < total) hf_pixel_epilogue(a, px, y, s_wout, s_bout);
  }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}



}  // namespace plh

using namespace plh;

#ifdef PLH_HF_TRACE
extern "C" __attribute__((visibility("default"))) int plh_hf_trace_read(unsigned long long* host) {
  return (int)cudaMemcpyFromSymbol(host, plh::g_hf_trace, sizeof(unsigned long long) * 8 * 512);
}
#endif

extern "C" int plh_head_fuse_level(const float* xa, int Ka, const float* wa, const float* scale_a, const float* shift_a,
                                   int relu_a, const float* xb, int Kb, const float* wb, const float* scale_b,
                                   const float* shift_b, int relu_b, const float* prev, const float* w_out,
                                   const float* b_out, int B, int H, int W, float* y18, float* pix_logits,
                                   float* link_logits, const plh_decode_params* flag_params, uint16_t* flags,
                                   void* stream) {
  if (!xa || !wa) return PLH_E_NULL;
  if (xb && !wb) return PLH_E_NULL;
  if (y18 ? (pix_logits || link_logits) : (!pix_logits || !link_logits)) return PLH_E_NULL;   // one output form
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
    if (!pix_logits || !flag_params) return PLH_E_NULL;   // the threshold word belongs to the logits of the last level
    a.flags = flags;
    a.tp_logit = prob_to_logit_threshold(flag_params->pixel_thresh);
    a.tl_logit = prob_to_logit_threshold(flag_params->link_thresh);
  }
  static SmemOptIn optin, optin_tc;
  int rc;
  const long long ntiles = ((long long)B * H * W + kHfTile - 1) / kHfTile;
  static const bool force_fma = getenv("PLH_HEADFUSE_FMA") != nullptr;   // A/B: the fp32-FMA kernel for every shape
  if (!force_fma && Ka % kTcKC == 0 && (!xb || Kb % kTcKC == 0)) {
    if ((rc = ensure_dynamic_smem(optin_tc, head_fuse_tc_kernel, kTcSmem))) return rc;
    const int grid_tc = (int)std::min<long long>(ntiles, 2ll * kNumSMs);
    head_fuse_tc_kernel<<<grid_tc, kTcThreads, kTcSmem, (cudaStream_t)stream>>>(a);
    return launch_status();
  }
  if ((rc = ensure_dynamic_smem(optin, head_fuse_kernel, kHfSmem))) return rc;
  const int grid = (int)std::min<long long>(ntiles, 2ll * kNumSMs);   // two resident CTAs per SM
  head_fuse_kernel<<<grid, kHfThreads, kHfSmem, (cudaStream_t)stream>>>(a);
  return launch_status();
}
