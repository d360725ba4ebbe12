// loss.cu — PixelLink loss forward+backward for sm_100a.
//
// Replaces nets/model.py:145-261, nets/model_vgg_16.py:227-282 and
// nets/pixellink.py:88-263 of the reference.
//
// Pipeline at images <= 65536 px (two launches on one stream, no host sync):
//   K1 ohem_select_cluster   a cluster of 8 CTAs per image: pixel softmax score -> 32-bit radix key
//                    per pixel in REGISTERS, exact k-th smallest by MSB-first bisection of the fp32 bit
//                    pattern (two bits per round, counts exchanged over distributed shared memory),
//                    then the selected mask (1 B/px) and the 18 integer normalisers of the image from
//                    link labels that TMA staged into shared memory during the rounds.
//   K3 loss_main     all SMs, the HBM-bound pass: a producer warp streams the inputs through a
//                    shared-memory ring with TMA bulk copies; reads 108 B/px once, writes the 72 B/px of
//                    gradients once (normalisers are already known), accumulates the 17 loss sums, last
//                    CTA finalises the scalars in a fixed order (deterministic).
// General path (larger images, positives-only variant, or the split hint):
//   K0 score_keys (keys to HBM, 4 B/px) -> ohem_select (one CTA per image) -> K2 ohem_counts (mask +
//   normalisers by integer atomics) -> K3.
// Algorithmic bytes: 180 B/px.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include <cooperative_groups.h>
#include <vector>

namespace plh {

std::atomic<long long> g_launch_count{0};

// Optional instrumentation for bench.py's roofline: CUDA events around the dominant
// kernel (K3 loss_main) on the launching stream.  Off unless plh_profile_begin() was
// called; not used under CUDA-graph capture.
constexpr int kProfMax = 8192;
static cudaEvent_t g_prof_ev[2 * kProfMax];
static int g_prof_created = 0;
static int g_prof_n = -1;  // -1: disabled
static unsigned long long* g_prof_ts = nullptr;  // device: [start, end] of the kernel per profiled launch

// ------------------------------------------------------------------ workspace layout
constexpr int kSumReplicas = 16;
enum { HC_N_SEG_POS = 0, HC_CNT_P = 1, HC_CNT_N = 9, HC_N_SELECTED = 17, HC_COUNT = 18 };
struct LossHeader {  // zeroed by K1 (or a memset when there is no K1) at the start of every call
  unsigned ticket;   // K3 last-CTA election
  int pad[63];
  // 17 loss sums (s_pos[8], s_neg[8], s_pix): fp64 atomic adds of fp32 per-CTA partials.  The
  // partials carry 24 significant bits and similar exponents, so the fp64 additions are exact
  // and the result does not depend on the order in which the CTAs arrive (deterministic).
  // Same-address atomics serialise in one L2 slice, so CTA c adds into replica c % kSumReplicas
  // (256 B apart) and the last CTA adds the replicas up in a fixed order.
  double sums[kSumReplicas][32];
};
static_assert(sizeof(LossHeader) == 256 + kSumReplicas * 256, "header size");
constexpr int kHdrInts = (int)(sizeof(LossHeader) / sizeof(int));

struct ImageInfo {  // per image, 128 B; written by K1 (cnt: by the cluster form of K1, else by K2)
  int n_pos;
  int n_neg;
  unsigned thr_key;  // bit pattern of the threshold score
  int valid;         // 0: select no negatives
  // integer normalisers of this image: n_seg_pos, cntP[8], cntN[8], n_selected (HC_* indices); the
  // batch totals K3 needs are the sums over the images (integers: exact, order independent)
  int cnt[HC_COUNT];
  int pad[10];
};
static_assert(sizeof(ImageInfo) == 128, "one line per image");

// batch totals of the HC_COUNT counters: lane l < HC_COUNT returns counter l summed over the images
__device__ __forceinline__ int batch_count(const ImageInfo* info, int B, int lane) {
  int v = 0;
  if (lane < HC_COUNT)
    for (int b = 0; b < B; ++b) v += __ldcg(&info[b].cnt[lane]);
  return v;
}
constexpr int kKeysMaxCTAsPerImage = 64;  // K0 writes one (n_pos, n_neg) pair per CTA; K1 sums them (no atomics, no memset)

#ifndef PLH_K3_WARPS
#define PLH_K3_WARPS 12
#endif
#ifndef PLH_K3_CTAS_PER_SM
#define PLH_K3_CTAS_PER_SM 2
#endif
#ifndef PLH_K3_STAGES
#define PLH_K3_STAGES 4
#endif
// Gradients through shared memory + bulk async stores (shared -> global) instead of per-lane streaming stores.
// Measured on B200 (32 x 128x128): no gain alone (20.57 vs 20.6 us per launch: the main pass is not limited by
// the LSU / issue cost of its stores) and 2.6 us slower inside the loss chain (a stage is held one tile longer),
// so it is off by default and kept for A/B runs (-DPLH_K3_BULK_STORE=1).
#ifndef PLH_K3_BULK_STORE
#define PLH_K3_BULK_STORE 0
#endif
constexpr bool kBulkStore = PLH_K3_BULK_STORE != 0;
constexpr int kMainThreads = 32 * PLH_K3_WARPS;   // consumer threads
constexpr int kMainCTAsPerSM = PLH_K3_CTAS_PER_SM;
constexpr int kMainMaxCTAs = kNumSMs * kMainCTAsPerSM;
constexpr int kSelectThreads = 1024;
constexpr size_t kSmemKeysMaxBytes = 200 * 1024;   // keys of one image fit in shared memory up to 51200 px

struct LossWsLayout {
  size_t header, info, counts, mask, keys, total;
};

static LossWsLayout loss_ws_layout(int B, long long N) {
  LossWsLayout l;
  size_t off = 0;
  l.header = off;
  off += sizeof(LossHeader);
  l.info = off;
  off = align_up(off + sizeof(ImageInfo) * (size_t)B, 256);
  l.counts = off;
  off = align_up(off + sizeof(int2) * (size_t)B * kKeysMaxCTAsPerImage, 256);
  l.mask = off;
  off = align_up(off + (size_t)B * N, 256);
  l.keys = off;
  off = align_up(off + (size_t)B * N * 4, 256);
  l.total = off;
  return l;
}

size_t loss_workspace_bytes(int B, int H, int W) { return loss_ws_layout(B, (long long)H * W).total; }

// ------------------------------------------------------------------ K0: scores -> keys
// Key of a pixel: the fp32 bit pattern of its score (scores are >= 0, so the
// unsigned order of the bits is the numeric order).
//   KEYS_MODEL     (nets/model.py:175-176)     non-negatives are excluded (key 0x7FFFFFFF)
//   KEYS_PIXELLINK (nets/pixellink.py:124-125) non-negatives become score 0 (key 0)
enum { KEYS_MODEL = 0, KEYS_PIXELLINK = 1 };
constexpr uint32_t kExcluded = 0x7FFFFFFFu;  // > every real key, and (t - key) stays negative in int32
constexpr int kKeysThreads = 256;

// key + class of one pixel (shared by K0 and the cluster form of K1)
template <int KEYMODE>
__device__ __forceinline__ uint32_t class_key(float sc, bool p, bool n, bool& isp, bool& isn) {
  isp = p, isn = n;
  return n ? __float_as_uint(sc) : ((KEYMODE == KEYS_MODEL) ? kExcluded : 0u);
}
template <int KEYMODE>
__device__ __forceinline__ uint32_t logit_key(float2 x, float l, bool& isp, bool& isn) {
  bool p, n;
  if (KEYMODE == KEYS_MODEL) {  // int32 cast truncates (model.py:213), ==1 / ==0 (:199-202)
    const int li = (int)l;
    p = li == 1, n = li == 0;
  } else {                      // pixellink.py:98-99: pos = labels > 0, neg = !pos
    p = l > 0.f, n = !p;
  }
  return class_key<KEYMODE>(neg_class_score(x.x, x.y), p, n, isp, isn);
}
template <int KEYMODE, bool FROM_SCORES>
__device__ __forceinline__ uint32_t pixel_key(const float* __restrict__ pix_logits, const float* __restrict__ pix_lab,
                                              const float* __restrict__ scores, const uint8_t* __restrict__ pos_mask,
                                              const uint8_t* __restrict__ neg_mask, size_t px, bool& isp, bool& isn) {
  if (FROM_SCORES) return class_key<KEYMODE>(scores[px], pos_mask[px] != 0, neg_mask[px] != 0, isp, isn);
  return logit_key<KEYMODE>(__ldg(reinterpret_cast<const float2*>(pix_logits) + px), __ldg(pix_lab + px), isp, isn);
}

template <int KEYMODE, bool FROM_SCORES>
__global__ void __launch_bounds__(kKeysThreads)
score_keys_kernel(const float* __restrict__ pix_logits, const float* __restrict__ pix_lab,
                  const float* __restrict__ scores, const uint8_t* __restrict__ pos_mask,
                  const uint8_t* __restrict__ neg_mask, int N, uint32_t* __restrict__ keys,
                  int2* __restrict__ counts) {
  pdl_wait_and_release();
  tl_start(0);
  __shared__ int s_np, s_nn;
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  if (tid == 0) s_np = 0, s_nn = 0;
  __syncthreads();
  const size_t base = (size_t)b * N;
  int npos = 0, nneg = 0;
  for (int i = blockIdx.x * kKeysThreads + tid; i < N; i += gridDim.x * kKeysThreads) {
    bool isp, isn;
    keys[base + i] = pixel_key<KEYMODE, FROM_SCORES>(pix_logits, pix_lab, scores, pos_mask, neg_mask, base + i, isp, isn);
    npos += isp, nneg += isn;
  }
  npos = __reduce_add_sync(0xffffffffu, npos);
  nneg = __reduce_add_sync(0xffffffffu, nneg);
  if ((tid & 31) == 0) {
    if (npos) atomicAdd(&s_np, npos);
    if (nneg) atomicAdd(&s_nn, nneg);
  }
  __syncthreads();
  if (tid == 0) counts[(size_t)b * gridDim.x + blockIdx.x] = make_int2(s_np, s_nn);
  tl_end(0);
}

// ------------------------------------------------------------------ K1: per-image OHEM threshold
// Exact k-th smallest key, built MSB first: a bit of the answer is 1 iff fewer than k keys are
// <= (prefix | that bit clear | all lower bits set).  Keys of real scores are <= bits(1.0f) =
// 0x3F800000 < 2^30, so 30 bits; excluded keys (0x7FFFFFFF) never count.
//
// Cluster form (images up to 65536 px) — also does K0's work, so the loss chain is one launch
// shorter: a cluster of 8 CTAs owns one image, every thread keeps KPT keys in registers, and each
// round resolves TWO bits (three pivots).  Per round every warp posts its three counts, packed with
// the round number into one 64-bit word, straight into the shared memory of all 8 CTAs of the
// cluster (DSMEM store), then spins on its OWN shared memory until the 64 words of the round have
// landed: no __syncthreads and no cluster barrier inside the loop (a cluster barrier alone costs
// more than the counting).  Slots are double buffered by round parity; a slot is rewritten two
// rounds later, which every warp can only reach after all warps have consumed the older value.
constexpr int kClusterSize = 8;
constexpr int kClThreads = 256;
constexpr int kClWarps = kClThreads / 32;
constexpr int kClSlots = kClusterSize * kClWarps;  // 64: two per lane
constexpr int kClMaxKPT = 32;
constexpr int kClReleaseRound = 13;  // of 15

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Wait for a phase of an mbarrier.  The suspend-time hint lets the hardware park the thread until the
// phase completes instead of returning to a polling loop every few hundred cycles: waiting warps then
// issue (almost) nothing, which matters when another kernel shares the SM.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(0x989680)
      : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, int rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

// post (c0,c1,c2) of this warp for exchange `e`, return the cluster-wide sums (same on every lane).
// Barrier `e & 1` of every CTA expects the 64 words of the exchange (one arrival + 512 transaction
// bytes); st.async delivers a word and signals the destination's barrier in one operation.
__device__ __forceinline__ void cluster_exchange3(unsigned long long (*slot)[kClSlots], unsigned long long* mbar,
                                                  int e, int rank, int warp, int lane, unsigned c0, unsigned c1,
                                                  unsigned c2, int& t0, int& t1, int& t2) {
  const unsigned long long v = ((unsigned long long)c2 << 32) | ((unsigned long long)c1 << 16) | c0;
  unsigned long long* mine = slot[e & 1];
  const uint32_t bar = smem_u32(mbar + (e & 1));
  if (warp == 0 && lane == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kClSlots * 8) : "memory");
  if (lane < kClusterSize) {
    const uint32_t dst = mapa_u32(smem_u32(mine + rank * kClWarps + warp), lane);
    const uint32_t rbar = mapa_u32(bar, lane);
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst), "l"(v),
                 "r"(rbar)
                 : "memory");
  }
  mbar_wait(bar, (e >> 1) & 1);
  const unsigned long long v0 = mine[lane], v1 = mine[lane + 32];
  t0 = __reduce_add_sync(0xffffffffu, (unsigned)(v0 & 0xffffu) + (unsigned)(v1 & 0xffffu));
  t1 = __reduce_add_sync(0xffffffffu, (unsigned)((v0 >> 16) & 0xffffu) + (unsigned)((v1 >> 16) & 0xffffu));
  t2 = __reduce_add_sync(0xffffffffu, (unsigned)((v0 >> 32) & 0xffffu) + (unsigned)((v1 >> 32) & 0xffffu));
}

// COUNTS (the loss): the kernel also does K2's work for its image.  The link labels of the slice are
// packed to 2 bits per direction while the keys are loaded; once the threshold is known every thread
// derives the selected mask of its pixels and the 18 integer normalisers, which are reduced over the
// cluster into rank 0's shared memory (DSMEM atomics) and stored, not accumulated, into the image's
// ImageInfo row: nothing to zero, nothing for K3 to wait for but this one kernel.
template <int KEYMODE, bool FROM_SCORES, int KPT, bool COUNTS>
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kClThreads, KPT <= 8 ? 5 : 4)
ohem_select_cluster_kernel(const float* __restrict__ pix_logits, const float* __restrict__ pix_lab,
                           const float* __restrict__ link_lab, const float* __restrict__ scores,
                           const uint8_t* __restrict__ pos_mask, const uint8_t* __restrict__ neg_mask,
                           const int* __restrict__ n_pos_override, int N, int ratio, ImageInfo* __restrict__ info,
                           float* __restrict__ thr_out, uint8_t* __restrict__ mask, uint32_t* __restrict__ keys_out,
                           LossHeader* __restrict__ hdr, int overlap_prev) {
  static_assert(KPT <= kClMaxKPT && 32 * KPT < 65536, "per-warp counts are exchanged in 16-bit fields");
  static_assert(!(COUNTS && FROM_SCORES), "the standalone OHNM has no link labels");
  // overlap_prev: the caller vouches that the preceding kernel of the stream touches nothing this call
  // touches (the decode's tile pass): no need to wait for it, only for what came before it — which was
  // complete before that kernel could start.
  if (!overlap_prev) pdl_wait();
  tl_start(1);
  namespace cg = cooperative_groups;
  extern __shared__ __align__(128) unsigned char s_stage[];  // COUNTS: link labels of up to 8 chunks (64 KB)
  __shared__ unsigned long long s_slot[2][kClSlots];
  __shared__ __align__(8) unsigned long long s_mbar[3];      // two exchange barriers + the staging barrier
  __shared__ int s_c[HC_COUNT];    // this CTA's normalisers
  __shared__ int s_tot[HC_COUNT];  // rank 0: the image's normalisers
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = blockIdx.x % kClusterSize;  // == %cluster_ctarank for a 1-D cluster
  const int b = blockIdx.x / kClusterSize;
  if (tid < HC_COUNT) s_c[tid] = 0, s_tot[tid] = 0;
  if (!COUNTS && !FROM_SCORES && rank == 0 && tid < HC_COUNT) info[b].cnt[tid] = 0;  // the separate K2 accumulates into it
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_mbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_mbar[1])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_mbar[2])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();  // the staging barrier is used by this CTA right away
  cluster_arrive();  // peers may signal this CTA's barriers only after it has initialised them
  // the accumulators of K3 are cleared here (K1 precedes it on the stream)
  if (hdr && blockIdx.x == 0)
    for (int i = tid; i < kHdrInts; i += kClThreads) reinterpret_cast<int*>(hdr)[i] = 0;

  // ---- this CTA's slice of the image: KPT chunks of 256 consecutive pixels, keys into registers.
  // The link labels of the slice are needed only after the threshold is known: they are fetched into
  // shared memory by bulk async copies (TMA, 8 KB per chunk, up to 8 chunks per pass) that fly while
  // the selection rounds run, and cost no registers.
  const size_t base = (size_t)b * N;
  constexpr int KB = KPT < 8 ? KPT : 8;
  auto stage_pass = [&](int j0) {  // one thread: bulk copies of chunks j0 .. j0+KB-1
    const uint32_t bar = smem_u32(&s_mbar[2]);
    uint32_t bytes = 0;
#pragma unroll
    for (int j = 0; j < KB; ++j) bytes += (uint32_t)min(max(N - (rank * KPT + j0 + j) * kClThreads, 0), kClThreads) * 32u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      const int i0 = (rank * KPT + j0 + j) * kClThreads;
      const int valid = min(max(N - i0, 0), kClThreads);
      if (valid > 0)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(s_stage + j * (kClThreads * 32))),
                     "l"(link_lab + (base + i0) * 8), "r"(valid * 32), "r"(bar)
                     : "memory");
    }
  };
  uint32_t rk[KPT];
  uint32_t posbits = 0, negbits = 0;  // bit j: pixel j of this thread is positive / negative
  unsigned npos = 0, nneg = 0;
#pragma unroll
  for (int j0 = 0; j0 < KPT; j0 += KB) {
    // all loads of a batch are issued before the first score is computed
    float2 x[KB];
    float l[KB];
    uint8_t mp[KB], mn[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      const int i = (rank * KPT + j0 + j) * kClThreads + tid;
      const size_t px = base + (i < N ? i : N - 1);
      if (FROM_SCORES) {
        x[j].x = scores[px], mp[j] = pos_mask[px], mn[j] = neg_mask[px];
      } else {
        x[j] = __ldg(reinterpret_cast<const float2*>(pix_logits) + px);
        l[j] = __ldg(pix_lab + px);
      }
    }
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      const int i = (rank * KPT + j0 + j) * kClThreads + tid;
      bool isp, isn;
      const uint32_t key = FROM_SCORES ? class_key<KEYMODE>(x[j].x, mp[j] != 0, mn[j] != 0, isp, isn)
                                       : logit_key<KEYMODE>(x[j], l[j], isp, isn);
      const bool in = i < N;
      rk[j0 + j] = in ? key : kExcluded;
      if (!COUNTS && keys_out && in) keys_out[base + i] = key;  // for the separate K2
      isp = isp && in, isn = isn && in;
      npos += isp, nneg += isn;
      posbits |= (uint32_t)isp << (j0 + j), negbits |= (uint32_t)isn << (j0 + j);
    }
  }
  npos = __reduce_add_sync(0xffffffffu, npos);
  nneg = __reduce_add_sync(0xffffffffu, nneg);
  cluster_wait();
  int n_pos, n_neg, unused;
  cluster_exchange3(s_slot, s_mbar, 0, rank, warp, lane, npos, nneg, 0u, n_pos, n_neg, unused);
  tl_end(11);
  if (COUNTS && tid == 0) stage_pass(0);  // behind the logits: the first exchange does not wait for 64 KB of labels
  if (n_pos_override) n_pos = n_pos_override[b];  // OHNM_single_image(scores, n_pos, neg_mask): n_pos is an argument
  // ---- k (model.py:170-173 / pixellink.py:116-120)
  const long long kk = (long long)n_pos * ratio;
  const int cap = (KEYMODE == KEYS_MODEL) ? n_neg : max(n_neg, 1);
  const int k = (int)(kk < (long long)cap ? kk : (long long)cap);
  const bool none = (n_pos <= 0) || (k <= 0);  // n_pos == 0 -> no_pos(); k == 0 -> "select none" 

  uint32_t ans = 0;
  if (!none) {  // uniform over the cluster
    // c_lo < k <= c_hi: number of keys <= the lower / upper end of the interval the answer is known to lie in
    int c_lo = 0, c_hi = (KEYMODE == KEYS_MODEL) ? n_neg : N;
    for (int e = 1; e <= 15; ++e) {
      const int lo = 30 - 2 * e;  // this round resolves bits lo+1, lo
      const uint32_t low = (1u << lo) - 1u;
      const uint32_t p0 = ans | low, p1 = p0 | (1u << lo), p2 = p0 | (2u << lo);
      // count(key > p) from the sign of (p - key): keys and pivots are < 2^31, so no wrap
      unsigned g0 = 0, g1 = 0, g2 = 0;
#pragma unroll
      for (int j = 0; j < KPT; ++j) {
        g0 += (p0 - rk[j]) >> 31;
        g1 += (p1 - rk[j]) >> 31;
        g2 += (p2 - rk[j]) >> 31;
      }
      g0 = __reduce_add_sync(0xffffffffu, g0);
      g1 = __reduce_add_sync(0xffffffffu, g1);
      g2 = __reduce_add_sync(0xffffffffu, g2);
      int t0, t1, t2;  // #keys <= pivot over the whole image
      cluster_exchange3(s_slot, s_mbar, e, rank, warp, lane, 32u * KPT - g0, 32u * KPT - g1, 32u * KPT - g2, t0, t1, t2);
      const int q = (t0 < k) + (t1 < k) + (t2 < k);
      ans |= (uint32_t)q << lo;
      c_lo = q == 0 ? c_lo : (q == 1 ? t0 : (q == 2 ? t1 : t2));
      c_hi = q == 0 ? t0 : (q == 1 ? t1 : (q == 2 ? t2 : c_hi));
      if (c_hi - c_lo == 1 && e < 14) {
        // Exactly one key is left in the interval, so it is the answer: its owner posts it (one more
        // exchange instead of the 15 - e remaining rounds; scores without ties get here after ~8 rounds).
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < KPT; ++j) mine |= ((rk[j] >> lo) == (ans >> lo)) ? rk[j] : 0u;  // excluded keys: bit 30 set
        mine = __reduce_or_sync(0xffffffffu, mine);
        int v0, v1, v2;
        cluster_exchange3(s_slot, s_mbar, e + 1, rank, warp, lane, mine & 0xffffu, mine >> 16, 0u, v0, v1, v2);
        ans = (uint32_t)v0 | ((uint32_t)v1 << 16);
        break;
      }
    }
  }
  tl_end(12);
  if (COUNTS) {
    // ---- selected mask + normalisers of this thread's pixels (what K2 does in the general form)
    // Branch-free so that the pixels interleave; the class tests are float compares instead of the
    // reference's int cast (same result for every finite label: (int)l == 1 <=> 1 <= l < 2 and
    // (int)l == 0 <=> |l| < 1), which keeps the quarter-rate F2I unit out of the loop.
    // Two sweeps over the staged labels (directions 0-3, then 4-7): 8 live counters instead of 16 keep
    // the kernel at <= 48 registers, which is what lets the decode's kernels share the SMs with it.
    unsigned nsp = 0, nsel = 0;
#pragma unroll
    for (int j0 = 0; j0 < KPT; j0 += KB) {
      mbar_wait(smem_u32(&s_mbar[2]), (j0 / KB) & 1);  // the pass's labels have landed (pass 0 flew during the rounds)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        unsigned cP[4] = {0u, 0u, 0u, 0u}, cN[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int jj = 0; jj < KB; ++jj) {
          const int j = j0 + jj;
          const int i = (rank * KPT + j) * kClThreads + tid;
          const bool in = i < N;
          const bool pos = (posbits >> j) & 1u, neg = (negbits >> j) & 1u;  // both false outside the image
          const bool M = pos || (neg && !none && rk[j] <= ans);  // model.py:178 ties at the threshold are all selected
          if (h == 0) {
            nsel += M;
            nsp += (KEYMODE == KEYS_PIXELLINK) ? M : pos;  // pixellink.py:155 sum(selected) / model.py:221 sum(pos)
            if (in) mask[base + i] = M ? 1 : 0;
          }
          const bool Ml = (KEYMODE == KEYS_PIXELLINK) ? in : M;  // pixellink.py:193-194: not x the OHEM mask
          // (stale shared memory outside the image: masked by Ml)
          const float4 lq = *reinterpret_cast<const float4*>(s_stage + jj * (kClThreads * 32) + tid * 32 + h * 16);
          const float lv[4] = {lq.x, lq.y, lq.z, lq.w};
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            bool lp, ln;
            if (KEYMODE == KEYS_PIXELLINK) lp = lv[d] > 0.f, ln = !lp;          // pixellink.py:185-186
            else lp = lv[d] >= 1.f && lv[d] < 2.f, ln = fabsf(lv[d]) < 1.f;      // model.py:238-241
            cP[d] += (Ml && lp), cN[d] += (Ml && ln);
          }
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const int cp = __reduce_add_sync(0xffffffffu, cP[d]);
          const int cn = __reduce_add_sync(0xffffffffu, cN[d]);
          if (lane == 0) {
            if (cp) atomicAdd(&s_c[HC_CNT_P + 4 * h + d], cp);
            if (cn) atomicAdd(&s_c[HC_CNT_N + 4 * h + d], cn);
          }
        }
      }
      if (j0 + KB < KPT) {  // next pass reuses the staging buffer
        __syncthreads();
        if (tid == 0) stage_pass(j0 + KB);
      }
    }
    nsp = __reduce_add_sync(0xffffffffu, nsp);
    nsel = __reduce_add_sync(0xffffffffu, nsel);
    if (lane == 0) {
      if (nsp) atomicAdd(&s_c[HC_N_SEG_POS], nsp);
      if (nsel) atomicAdd(&s_c[HC_N_SELECTED], nsel);
    }
    tl_end(13);
    __syncthreads();
    if (tid < HC_COUNT && s_c[tid]) atomicAdd(cg::this_cluster().map_shared_rank(&s_tot[tid], 0), s_c[tid]);
  }
  tl_end(14);
  pdl_release();
  // no CTA may retire while a peer's store to it could still be in flight; rank 0 reads the totals after it
  cluster_arrive();
  cluster_wait();
  if (rank == 0) {
    if (COUNTS && tid < HC_COUNT) info[b].cnt[tid] = s_tot[tid];
    if (tid == 0) {
      info[b].n_pos = n_pos, info[b].n_neg = n_neg;
      info[b].thr_key = ans;
      info[b].valid = none ? 0 : 1;
      // NaN when nothing is selected: `score <= NaN` is false for every pixel
      if (thr_out) thr_out[b] = none ? __int_as_float(0x7fc00000) : __uint_as_float(ans);
    }
  }
  tl_end(1);
}

// General form (any image size): one CTA per image after K0; keys are re-read every round from a
// shared-memory copy if it fits, else from global/L2; one bit per round.
__global__ void __launch_bounds__(kSelectThreads, 1)
ohem_select_kernel(const uint32_t* __restrict__ keys_all, const int2* __restrict__ counts, int ncounts,
                   const int* __restrict__ n_pos_override, int N, int ratio, int keymode, int use_smem,
                   ImageInfo* __restrict__ info, float* __restrict__ thr_out, LossHeader* __restrict__ hdr) {
  pdl_wait();
  tl_start(1);
  extern __shared__ __align__(16) uint32_t skeys[];
  __shared__ int s_w[2][32];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t* gkeys = keys_all + (size_t)b * N;

  if (hdr && b == 0)
    for (int i = tid; i < kHdrInts; i += kSelectThreads) reinterpret_cast<int*>(hdr)[i] = 0;
  if (tid < HC_COUNT) info[b].cnt[tid] = 0;  // K2 accumulates into it
  int npos = 0, nneg = 0;
  if (warp == 0) {
    for (int i = lane; i < ncounts; i += 32) {
      const int2 c = counts[(size_t)b * ncounts + i];
      npos += c.x, nneg += c.y;
    }
    npos = __reduce_add_sync(0xffffffffu, npos);
    nneg = __reduce_add_sync(0xffffffffu, nneg);
    if (lane == 0) s_w[0][0] = npos, s_w[0][1] = nneg;
  }
  __syncthreads();
  npos = s_w[0][0], nneg = s_w[0][1];
  __syncthreads();
  if (n_pos_override) npos = n_pos_override[b];
  const long long kk = (long long)npos * ratio;
  const int cap = (keymode == KEYS_MODEL) ? nneg : max(nneg, 1);
  const int k = (int)(kk < (long long)cap ? kk : (long long)cap);
  const bool none = (npos <= 0) || (k <= 0);

  uint32_t ans = 0;
  if (!none) {
    const uint32_t* keys = gkeys;
    if (use_smem) {
      for (int i = tid; i < N; i += kSelectThreads) skeys[i] = __ldg(gkeys + i);
      keys = skeys;
      __syncthreads();
    }
    for (int bit = 29; bit >= 0; --bit) {
      const uint32_t t = ans | ((1u << bit) - 1u);
      int c = 0;
      for (int i = tid; i < N; i += kSelectThreads) c += (keys[i] <= t);
      c = __reduce_add_sync(0xffffffffu, c);
      int* sw = s_w[bit & 1];  // double-buffered: one barrier per round
      if (lane == 0) sw[warp] = c;
      __syncthreads();
      const int total = __reduce_add_sync(0xffffffffu, sw[lane]);
      if (total < k) ans |= (1u << bit);
    }
  }
  pdl_release();
  if (tid == 0) {
    info[b].n_pos = npos, info[b].n_neg = nneg;
    info[b].thr_key = ans;
    info[b].valid = none ? 0 : 1;
    if (thr_out) thr_out[b] = none ? __int_as_float(0x7fc00000) : __uint_as_float(ans);
  }
  tl_end(1);
}

// ------------------------------------------------------------------ K2: mask + integer normalisers
constexpr int kCountsThreads = 512;
template <int VARIANT>
__global__ void __launch_bounds__(kCountsThreads)
ohem_counts_kernel(const uint32_t* __restrict__ keys, const float* __restrict__ pix_lab,
                   const float* __restrict__ link_lab, ImageInfo* __restrict__ info, int N,
                   uint8_t* __restrict__ mask) {
  pdl_wait_and_release();
  tl_start(2);
  __shared__ int s_c[18];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  if (tid < 18) s_c[tid] = 0;
  __syncthreads();
  int cP[8], cN[8], nsp = 0, nsel = 0;
#pragma unroll
  for (int d = 0; d < 8; ++d) cP[d] = 0, cN[d] = 0;
  unsigned thr_key = 0;
  bool valid = false;
  if (VARIANT != PLH_VARIANT_POS_ONLY) {
    const ImageInfo ii = info[b];
    thr_key = ii.thr_key, valid = ii.valid != 0;
  }
  const size_t base = (size_t)b * N;
  for (int i = blockIdx.x * kCountsThreads + tid; i < N; i += gridDim.x * kCountsThreads) {
    const size_t px = base + i;
    const float4 la = __ldg(reinterpret_cast<const float4*>(link_lab) + px * 2);
    const float4 lb = __ldg(reinterpret_cast<const float4*>(link_lab) + px * 2 + 1);
    const float l = __ldg(pix_lab + px);
    const unsigned key = (VARIANT != PLH_VARIANT_POS_ONLY) ? __ldg(keys + px) : 0u;
    bool pos, neg;
    if (VARIANT == PLH_VARIANT_PIXELLINK) pos = l > 0.f, neg = !pos;
    else { const int li = (int)l; pos = li == 1, neg = li == 0; }
    bool M = pos;
    if (VARIANT != PLH_VARIANT_POS_ONLY && neg)
      M = valid && key <= thr_key;  // model.py:178 ties at the threshold are all selected
    mask[px] = M ? 1 : 0;
    nsel += M;
    if (VARIANT == PLH_VARIANT_PIXELLINK) nsp += M;  // pixellink.py:155 n_seg_pos = sum(selected mask)
    else nsp += pos;                                 // model.py:221    n_seg_pos = sum(pos mask)
    const bool Ml = (VARIANT == PLH_VARIANT_PIXELLINK) ? true : M;  // pixellink.py:193-194: not x OHEM mask
    const float lv[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      bool lp, ln;
      if (VARIANT == PLH_VARIANT_PIXELLINK) lp = lv[d] > 0.f, ln = !lp;
      else { const int li = (int)lv[d]; lp = li == 1, ln = li == 0; }
      cP[d] += (lp && Ml), cN[d] += (ln && Ml);
    }
  }
  const int lane = tid & 31;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    const int a = __reduce_add_sync(0xffffffffu, cP[d]);
    const int c = __reduce_add_sync(0xffffffffu, cN[d]);
    if (lane == 0) {
      if (a) atomicAdd(&s_c[1 + d], a);
      if (c) atomicAdd(&s_c[9 + d], c);
    }
  }
  nsp = __reduce_add_sync(0xffffffffu, nsp);
  nsel = __reduce_add_sync(0xffffffffu, nsel);
  if (lane == 0) {
    if (nsp) atomicAdd(&s_c[0], nsp);
    if (nsel) atomicAdd(&s_c[17], nsel);
  }
  __syncthreads();
  // integer atomics: exact and order independent
  if (tid < HC_COUNT && s_c[tid]) atomicAdd(&info[b].cnt[tid], s_c[tid]);  // row zeroed by K1 / the memset
  tl_end(2);
}

// ------------------------------------------------------------------ K3: main fused pass
struct MainArgs {
  const float* pix_logits;
  const float* link_logits;
  const float* pix_lab;
  const float* link_lab;
  const uint8_t* mask;
  LossHeader* hdr;
  const ImageInfo* info;
  unsigned long long* ts;  // profiling only: [first CTA start, last CTA end] in %globaltimer ns, else null
  float* stats;
  float* grad_pix;
  float* grad_link;
  uint16_t* flags;
  long long total_px;
  float alpha, gamma;
  float tp_logit, tl_logit;  // decode thresholds moved to logit-difference space
};

// One 2-way softmax term.  Returns the per-element loss term and d(term)/d(logit 1);
// d/d(logit 0) is its negative.  Fast intrinsics: nothing here feeds a mask decision.
//   z = x_other - x_label;  CE = softplus(z) = max(z,0) + log(1 + exp(-|z|));
//   d CE / d logit1 = +sigmoid(z) for label 0, -sigmoid(z) for label 1.
template <int TERM>
__device__ __forceinline__ void term_and_grad(float x0, float x1, bool lab1, float alpha, float gamma,
                                              float& term, float& g1) {
  const float d = x1 - x0;
  const float z = lab1 ? -d : d;
  const float e = __expf(-fabsf(z));
  const float den = 1.f + e;
  const float r = __fdividef(1.f, den);
  const float sig = z >= 0.f ? r : e * r;         // sigmoid(z) = 1 - p_label
  const float ce = fmaxf(z, 0.f) + __logf(den);
  if (TERM == PLH_TERM_CE) {
    term = ce;
    g1 = lab1 ? -sig : sig;
  } else {
    const float pt = z >= 0.f ? e * r : r;        // p_label = sigmoid(-z)
    const float logpt = -ce;
    const float at = lab1 ? alpha : 1.f - alpha;
    const float om = sig;                          // 1 - p_label
    const float mod = (gamma == 2.f) ? om * om : __powf(om, gamma);
    term = -(at * mod * logpt);
    const float gt = at * mod * (gamma * pt * logpt - om);   // d/d x_label
    g1 = lab1 ? gt : -gt;
  }
}

// model.py:213 casts the labels to int32 (truncation) before `== 1` / `== 0` (:199-202).  Same decisions from float
// compares, which keeps the quarter-rate conversion unit out of the main pass: (int)l == 1 <=> 1 <= l < 2;
// (int)l == 0 <=> |l| < 1, or l is NaN (the device cast gives 0 for NaN; +-inf saturate and match neither).
template <int VARIANT>
__device__ __forceinline__ void classify(float l, bool& p, bool& n) {
  if (VARIANT == PLH_VARIANT_PIXELLINK) p = l > 0.f, n = !p;
  else p = l >= 1.f && l < 2.f, n = !(fabsf(l) >= 1.f);
}

// ---- CTA totals -> header (fp64 atomics), last CTA -> the scalars.  s_red holds the consumer warps' sums.
template <int VARIANT>
__device__ __forceinline__ void main_epilogue(const MainArgs& a, int B, int N, float (*s_red)[4][5], const int* s_cnt,
                                              bool* s_last) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  if (tid < 17) {
    // sum index: 0..7 s_pos[d], 8..15 s_neg[d], 16 s_pix ; d = 2*jj + c
    int jj, slot;
    if (tid < 16) { const int d = tid & 7; jj = d >> 1; slot = (tid < 8 ? 0 : 2) + (d & 1); }
    else { jj = 0; slot = 4; }
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kMainThreads / 32; ++w) s += s_red[w][jj][slot];
    atomicAdd(&a.hdr->sums[blockIdx.x % kSumReplicas][tid], (double)s);  // exact in fp64: order independent
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) *s_last = (atomicAdd(&a.hdr->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!*s_last) return;

  // ---- last CTA: the scalars.  One warp: lane d owns link direction d, so the header loads and the
  // IEEE divisions run side by side instead of as one thread's chain of dependent L2 round trips.
  __threadfence();
  if (warp == 0) {
    float* st = a.stats;
    const float nsp = (float)s_cnt[HC_N_SEG_POS], nsel = (float)s_cnt[HC_N_SELECTED];
    const int dd = lane & 7;
    const float cp = (float)s_cnt[HC_CNT_P + dd], cn = (float)s_cnt[HC_CNT_N + dd];
    double sum = 0.0;  // lane l < 17: sum l, replicas added in a fixed order (independent loads: one round trip)
    if (lane < 17) {
#pragma unroll
      for (int r = 0; r < kSumReplicas; ++r) sum += __ldcg(&a.hdr->sums[r][lane]);
    }
    const float s_pos = (float)__shfl_sync(0xffffffffu, sum, dd), s_neg = (float)__shfl_sync(0xffffffffu, sum, 8 + dd);
    const double s_pix_d = __shfl_sync(0xffffffffu, sum, 16);
    const float s_pix = (float)s_pix_d;
    float Ld;
    if (VARIANT == PLH_VARIANT_PIXELLINK)
      Ld = (cp != 0.f ? s_pos * __fdiv_rn(1.f, cp) : 0.f) + (cn != 0.f ? s_neg * __fdiv_rn(1.f, cn) : 0.f);
    else
      Ld = __fdiv_rn(s_pos, cp) + __fdiv_rn(s_neg, cn);
    if (lane < 8) {
      st[PLH_ST_L_LINK + dd] = Ld;
      st[PLH_ST_SUM_WP + dd] = cp;
      st[PLH_ST_SUM_WN + dd] = cn;
      st[PLH_ST_S_POS + dd] = s_pos;
      st[PLH_ST_S_NEG + dd] = s_neg;
    }
    double link_total = 0.0;
#pragma unroll
    for (int d = 0; d < 8; ++d) link_total += (double)__shfl_sync(0xffffffffu, Ld, d);  // fixed order
    float L_pix;
    if (VARIANT == PLH_VARIANT_MODEL) L_pix = nsp > 0.f ? __fdiv_rn(s_pix, nsp) : 0.f;
    else if (VARIANT == PLH_VARIANT_POS_ONLY) L_pix = __fdiv_rn(s_pix, nsp);
    else L_pix = (float)(s_pix_d / (double)((long long)B * N));
    if (lane == 0) {
      st[PLH_ST_L_PIX] = L_pix;
      st[PLH_ST_N_SEG_POS] = nsp;
      st[PLH_ST_S_PIX] = s_pix;
      st[PLH_ST_LINK_TOTAL] = (float)link_total;
      st[PLH_ST_N_SELECTED] = nsel;
      st[PLH_ST_TOTAL] = (float)link_total + 2.f * L_pix;  // model.py:261 / pixellink.py:170,254
    }
    if (lane >= 14) st[32 + lane] = 0.f;  // stats[46..63] reserved
    // leave the accumulators clean, so that this kernel can be relaunched on the same prepared
    // workspace (plh_loss_params.reserved[0] = 1: measurement of the main pass alone)
    __syncwarp();
    if (lane < 17) {
#pragma unroll
      for (int r = 0; r < kSumReplicas; ++r) a.hdr->sums[r][lane] = 0.0;
    }
    if (lane == 0) a.hdr->ticket = 0u;
  }
}

// ------------------------------------------------------------------ K3: main fused pass, TMA-staged
// Work unit = 16 consecutive pixels, handled by one warp:
//   link phase : 2 iterations x 32 lanes, lane = (pixel-in-iteration, quarter j): 128-bit operands of the
//                two link directions 2j, 2j+1 of that pixel (logits) + 64 bits of their labels;
//   pixel phase: lanes 0..15 own one pixel each (lanes 16..31 shadow them) — the pixel term is computed
//                once per pixel instead of once per quarter.
// Every CTA owns a contiguous range of units and walks it in tiles of 12 units (one per consumer warp).  A
// producer warp streams the five input arrays of a tile into a ring of shared-memory stages with bulk
// async copies (TMA; 20.9 KB per tile, completion on an mbarrier), several tiles ahead of the
// consumers, so the bytes in flight no longer live in registers and no warp ever waits on a global
// load; consumers read their unit from shared memory (conflict-free 128-bit reads), hand the stage
// back at once and compute.  Gradients leave through streaming 128-bit stores as before.
constexpr int kTileUnits = kMainThreads / 32;          // 12 consumer warps, one unit each
constexpr int kTilePx = kTileUnits * 16;               // 192
constexpr int kOffLL = 0;                              // link logits  64 B/px
constexpr int kOffLB = kOffLL + kTilePx * 64;          // link labels  32 B/px
constexpr int kOffPL = kOffLB + kTilePx * 32;          // pixel logits  8 B/px
constexpr int kOffPB = kOffPL + kTilePx * 8;           // pixel labels  4 B/px
constexpr int kOffMK = kOffPB + kTilePx * 4;           // selected mask 1 B/px
constexpr int kStageBytes = ((kOffMK + kTilePx + 127) / 128) * 128;
constexpr int kStages = PLH_K3_STAGES;
constexpr int kMainBlock = kMainThreads + 32;          // + the producer warp
constexpr size_t kMainSmem = (size_t)kStages * kStageBytes;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// The main pass streams 94 MB through the 126 MB L2 exactly once while, on the other stream, the decode's
// latency-bound kernels live on ~10 MB of L2-resident state (forest, sizes, flags, records): the streamed lines
// are marked evict-first so that they do not push that state out (PLH_K3_L2_HINT=0 builds without the hints).
#ifndef PLH_K3_L2_HINT
#define PLH_K3_L2_HINT 1
#endif
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                              unsigned long long pol) {
#if PLH_K3_L2_HINT
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "l"(pol)
               : "memory");
#else
  bulk_g2s(dst, src, bytes, bar);
#endif
}
__device__ __forceinline__ void stg_stream4_hint(float4* p, float4 v, unsigned long long pol) {
#if PLH_K3_L2_HINT
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
#else
  stg_stream4(p, v);
#endif
}
__device__ __forceinline__ void stg_stream2_hint(float2* p, float2 v, unsigned long long pol) {
#if PLH_K3_L2_HINT
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol)
               : "memory");
#else
  stg_stream2(p, v);
#endif
}

// one 16-pixel unit, operands already in registers (see loss_main_kernel for the lane mapping)
// sgl / sgp: when non-null the gradients of the unit go to shared memory (the unit's own slots of its stage: the
// gradient has the layout of the logits it was computed from) and leave with bulk async stores; else straight to
// global memory with streaming stores.
template <int VARIANT, int TERM, bool GRAD, bool FLAGS>
__device__ __forceinline__ void main_unit(const MainArgs& a, int px0, int total_px, const float4 (&L)[2],
                                          const float2 (&LB)[2], float2 P, float PLB, float MF, float pix_scale,
                                          const float (&invP)[2], const float (&invN)[2], int lane, float (&sp)[2],
                                          float (&sn)[2], float& spx, float4* sgl = nullptr, float2* sgp = nullptr,
                                          unsigned long long pol = 0ull) {
  const int j = lane & 3, q_pix = lane >> 2, pl_lane = lane & 15;
  const int pp = px0 + pl_lane;
  {
    bool ppos, pneg;
    classify<VARIANT>(PLB, ppos, pneg);
    float t, g1;
    term_and_grad<TERM>(P.x, P.y, ppos, a.alpha, a.gamma, t, g1);
    if (lane < 16 && pp < total_px) {
      spx += t * MF;
      if (GRAD) {
        const float gp = (MF * pix_scale) * g1;
        if (sgp) sgp[pl_lane] = make_float2(-gp, gp);
        else stg_stream2_hint(reinterpret_cast<float2*>(a.grad_pix) + pp, make_float2(-gp, gp), pol);
      }
    }
  }
  unsigned lbits[2] = {0u, 0u};
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int px = px0 + it * 8 + q_pix;
    const float Mf = __shfl_sync(0xffffffffu, MF, it * 8 + q_pix);  // the pixel's selected-mask weight
    bool p0, n0, p1, n1;
    classify<VARIANT>(LB[it].x, p0, n0);
    classify<VARIANT>(LB[it].y, p1, n1);
    float t0, g0, t1, g1;
    term_and_grad<TERM>(L[it].x, L[it].y, p0, a.alpha, a.gamma, t0, g0);
    term_and_grad<TERM>(L[it].z, L[it].w, p1, a.alpha, a.gamma, t1, g1);
    if (px < total_px) {
      const float tm0 = t0 * Mf, tm1 = t1 * Mf;
      sp[0] += p0 ? tm0 : 0.f, sn[0] += n0 ? tm0 : 0.f;
      sp[1] += p1 ? tm1 : 0.f, sn[1] += n1 ? tm1 : 0.f;
      if (GRAD) {
        const float a0 = (Mf * (p0 ? invP[0] : (n0 ? invN[0] : 0.f))) * g0;
        const float a1 = (Mf * (p1 ? invP[1] : (n1 ? invN[1] : 0.f))) * g1;
        if (sgl) sgl[it * 32 + lane] = make_float4(-a0, a0, -a1, a1);
        else stg_stream4_hint(reinterpret_cast<float4*>(a.grad_link) + ((size_t)px * 4 + j), make_float4(-a0, a0, -a1, a1), pol);
      }
    }
    if (FLAGS) {
      unsigned bits = ((L[it].y - L[it].x) > a.tl_logit ? 1u : 0u) << (2 * j) |
                      ((L[it].w - L[it].z) > a.tl_logit ? 1u : 0u) << (2 * j + 1);
      bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
      bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
      lbits[it] = bits;
    }
  }
  if (FLAGS) {
    const unsigned b0 = __shfl_sync(0xffffffffu, lbits[0], (lane & 7) << 2);
    const unsigned b1 = __shfl_sync(0xffffffffu, lbits[1], (lane & 7) << 2);
    if (lane < 16 && pp < total_px)
      a.flags[pp] = (uint16_t)(((lane & 8) ? b1 : b0) | (((P.y - P.x) > a.tp_logit ? 1u : 0u) << 8));
  }
}

template <int VARIANT, int TERM, bool GRAD, bool FLAGS>
__global__ void __launch_bounds__(kMainBlock, kMainCTAsPerSM)
loss_main_kernel(const MainArgs a, const int B, const int N) {
  // No dependency wait at the top: the logits and labels are not written by the preceding kernel (the selection
  // kernel, or this kernel's previous launch), so the producer fetches the first tiles of them BEFORE it waits;
  // only the selected mask and the normalisers are the predecessor's, and they are read after the wait.
  tl_start(3);
  if (a.ts && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    atomicMin(a.ts, t);
  }
  extern __shared__ __align__(128) unsigned char stages[];
  __shared__ __align__(8) unsigned long long s_full[kStages], s_empty[kStages];
  __shared__ float s_red[kMainThreads / 32][4][5];
  __shared__ int s_cnt[HC_COUNT];
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int j = tid & 3;
  const bool producer = warp == kTileUnits;

  // ---- this CTA's contiguous range of full units, in tiles of kTileUnits
  const int total_px = (int)a.total_px;
  const int nfull = total_px >> 4;
  const int per = nfull / (int)gridDim.x, rem = nfull % (int)gridDim.x;
  const int u_begin = (int)blockIdx.x * per + min((int)blockIdx.x, rem);
  const int u_count = per + ((int)blockIdx.x < rem ? 1 : 0);
  const int ntiles = (u_count + kTileUnits - 1) / kTileUnits;
  constexpr bool kNeedMask = VARIANT != PLH_VARIANT_PIXELLINK;

  // ---- producer: one thread initialises the barriers and starts streaming at once, while the
  // consumer warps are still fetching the normalisers (they meet the barriers behind a named barrier)
  if (producer) {
    if (lane == 0) {
#pragma unroll
      for (int st = 0; st < kStages; ++st) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_full[st])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_empty[st])), "r"(kTileUnits));
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("bar.arrive 1, %0;" ::"r"(kMainBlock) : "memory");  // barriers are live
    if (lane == 0) {
      const unsigned long long pol = l2_evict_first_policy();
      auto fetch = [&](int k, bool inputs, bool mask_too) {
        const int st = k % kStages;
        const int nu = min(kTileUnits, u_count - k * kTileUnits);
        const size_t px0 = (size_t)(u_begin + k * kTileUnits) << 4;
        const uint32_t npx = (uint32_t)nu * 16u;
        const uint32_t bar = smem_u32(&s_full[st]);
        const uint32_t base = smem_u32(stages + (size_t)st * kStageBytes);
        if (inputs) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                       "r"(npx * (kNeedMask ? 109u : 108u))
                       : "memory");
          bulk_g2s_hint(base + kOffLL, a.link_logits + px0 * 16, npx * 64u, bar, pol);
          bulk_g2s_hint(base + kOffLB, a.link_lab + px0 * 8, npx * 32u, bar, pol);
          bulk_g2s_hint(base + kOffPL, a.pix_logits + px0 * 2, npx * 8u, bar, pol);
          bulk_g2s_hint(base + kOffPB, a.pix_lab + px0, npx * 4u, bar, pol);
        }
        if (kNeedMask && mask_too) bulk_g2s_hint(base + kOffMK, a.mask + px0, npx, bar, pol);
      };
      const int nearly = min(ntiles, kStages);
      for (int k = 0; k < nearly; ++k) fetch(k, true, false);  // the ring is empty: no wait
      pdl_wait();                                               // the mask is the predecessor's
      pdl_release();
      for (int k = 0; k < nearly; ++k) fetch(k, false, true);
      for (int k = nearly; k < ntiles; ++k) {
        mbar_wait(smem_u32(&s_empty[k % kStages]), ((k / kStages) & 1) ^ 1);
        fetch(k, true, true);
      }
    }
  } else {
    // ---- normalisers (final: the selection/count kernels completed before this launch)
    pdl_wait();
    if (warp == 0) {
      const int c = batch_count(a.info, B, lane);
      if (lane < HC_COUNT) s_cnt[lane] = c;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(kMainBlock) : "memory");  // s_cnt written, mbarriers initialised
  }
  float pix_scale, invP[2], invN[2];
  {
    const float nsp = (float)s_cnt[HC_N_SEG_POS];
    if (VARIANT == PLH_VARIANT_MODEL) pix_scale = nsp > 0.f ? __fdiv_rn(2.f, nsp) : 0.f;  // model.py:226-233
    else if (VARIANT == PLH_VARIANT_POS_ONLY) pix_scale = __fdiv_rn(2.f, nsp);            // vgg16 :267 unguarded
    else pix_scale = __fdiv_rn(2.f, (float)((long long)B * N));                          // pixellink.py:160,170
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float cp = (float)s_cnt[HC_CNT_P + 2 * j + c], cn = (float)s_cnt[HC_CNT_N + 2 * j + c];
      if (VARIANT == PLH_VARIANT_PIXELLINK) {  // pixellink.py:198-211 zero guards
        invP[c] = cp != 0.f ? __fdiv_rn(1.f, cp) : 0.f;
        invN[c] = cn != 0.f ? __fdiv_rn(1.f, cn) : 0.f;
      } else {                                 // model.py:252-253 unguarded: 0/0 = NaN is data
        invP[c] = __fdiv_rn(1.f, cp);
        invN[c] = __fdiv_rn(1.f, cn);
      }
    }
  }

  // ---- consumers: operands of this warp's unit from the stage, stage back to the producer, compute; gradients
  // leave through streaming 128-bit stores.  (kBulkStore: the gradients are written back INTO the unit's slots of
  // the stage — same layout as the logits — and leave through two bulk async stores, shared -> global, 1024 +
  // 128 B per unit, issued by lane 0; the stage then goes back to the producer one tile later, once the bulk
  // stores of the tile have read their source.)
  float sp[2] = {0.f, 0.f}, sn[2] = {0.f, 0.f}, spx = 0.f;
  if (!producer) {
    const int pl_lane = lane & 15;
    const unsigned long long gpol = l2_evict_first_policy();   // gradients: written once, read by nobody here
    int held = -1;  // stage whose bulk stores are still reading shared memory
    for (int k = 0; k < ntiles; ++k) {
      const int st = k % kStages;
      const int nu = min(kTileUnits, u_count - k * kTileUnits);
      mbar_wait(smem_u32(&s_full[st]), (k / kStages) & 1);
      unsigned char* sb = stages + (size_t)st * kStageBytes;
      float4 L[2];
      float2 LB[2], P;
      float PLB, MF;
      const bool mine = warp < nu;
      if (mine) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          L[it] = *reinterpret_cast<const float4*>(sb + kOffLL + warp * 1024 + (it * 32 + lane) * 16);
          LB[it] = *reinterpret_cast<const float2*>(sb + kOffLB + warp * 512 + (it * 32 + lane) * 8);
        }
        P = *reinterpret_cast<const float2*>(sb + kOffPL + (warp * 16 + pl_lane) * 8);
        PLB = *reinterpret_cast<const float*>(sb + kOffPB + (warp * 16 + pl_lane) * 4);
        MF = kNeedMask ? (float)sb[kOffMK + warp * 16 + pl_lane] : 1.f;
      }
      __syncwarp();
      if (GRAD && kBulkStore) {
        // the previous tile's bulk stores have read their source by now: hand that stage back
        if (held >= 0 && lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_empty[held])) : "memory");
        }
        held = st;
        if (mine) {
          const int px0 = (u_begin + k * kTileUnits + warp) << 4;
          float4* sgl = reinterpret_cast<float4*>(sb + kOffLL + warp * 1024);
          float2* sgp = reinterpret_cast<float2*>(sb + kOffPL + warp * 128);
          main_unit<VARIANT, TERM, GRAD, FLAGS>(a, px0, total_px, L, LB, P, PLB, MF, pix_scale, invP, invN, lane, sp, sn,
                                                spx, sgl, sgp);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk copy
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.grad_link + (size_t)px0 * 16),
                         "r"(smem_u32(sgl)), "r"(1024)
                         : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.grad_pix + (size_t)px0 * 2),
                         "r"(smem_u32(sgp)), "r"(128)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else {
        if (lane == 0)  // the stage goes back to the producer as soon as this warp holds its operands
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_empty[st])) : "memory");
        if (mine)
          main_unit<VARIANT, TERM, GRAD, FLAGS>(a, (u_begin + k * kTileUnits + warp) << 4, total_px, L, LB, P, PLB, MF,
                                                pix_scale, invP, invN, lane, sp, sn, spx, nullptr, nullptr, gpol);
      }
    }
    if (GRAD && kBulkStore && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all gradient stores complete
    // the ragged last unit of the batch (total_px % 16 pixels): guarded global loads, one warp
    if ((total_px & 15) && blockIdx.x == gridDim.x - 1 && warp == 0) {
      const int px0 = nfull << 4;
      const int q_pix = lane >> 2;
      float4 L[2];
      float2 LB[2], P = make_float2(0.f, 0.f);
      float PLB = 0.f, MF = 0.f;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int px = px0 + it * 8 + q_pix;
        L[it] = make_float4(0.f, 0.f, 0.f, 0.f), LB[it] = make_float2(0.f, 0.f);
        if (px < total_px) {
          const size_t q = (size_t)px * 4 + j;
          L[it] = ldg_stream4(reinterpret_cast<const float4*>(a.link_logits) + q);
          LB[it] = ldg_stream2(reinterpret_cast<const float2*>(a.link_lab) + q);
        }
      }
      const int pp = px0 + pl_lane;
      if (pp < total_px) {
        P = __ldg(reinterpret_cast<const float2*>(a.pix_logits) + pp);
        PLB = __ldg(a.pix_lab + pp);
        MF = kNeedMask ? (float)__ldg(a.mask + pp) : 1.f;
      }
      main_unit<VARIANT, TERM, GRAD, FLAGS>(a, px0, total_px, L, LB, P, PLB, MF, pix_scale, invP, invN, lane, sp, sn,
                                            spx, nullptr, nullptr, gpol);
    }
  }

  // ---- block reduction of the 17 sums (consumer warps only) and the last-CTA epilogue
  if (!producer) {
    spx += __shfl_xor_sync(0xffffffffu, spx, 1);
    spx += __shfl_xor_sync(0xffffffffu, spx, 2);
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
      sp[0] += __shfl_xor_sync(0xffffffffu, sp[0], o);
      sp[1] += __shfl_xor_sync(0xffffffffu, sp[1], o);
      sn[0] += __shfl_xor_sync(0xffffffffu, sn[0], o);
      sn[1] += __shfl_xor_sync(0xffffffffu, sn[1], o);
      spx += __shfl_xor_sync(0xffffffffu, spx, o);
    }
    if (lane < 4) {
      s_red[warp][lane][0] = sp[0], s_red[warp][lane][1] = sp[1];
      s_red[warp][lane][2] = sn[0], s_red[warp][lane][3] = sn[1];
      s_red[warp][lane][4] = spx;
    }
  }
  __syncthreads();
  main_epilogue<VARIANT>(a, B, N, s_red, s_cnt, &s_last);
  if (a.ts && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    atomicMax(a.ts + 1, t);
  }
  tl_end(3);
}

// ------------------------------------------------------------------ standalone OHNM_batch apply
__global__ void ohnm_apply_kernel(const float* __restrict__ scores, const uint8_t* __restrict__ pos,
                                  const uint8_t* __restrict__ neg, const float* __restrict__ thr, int N,
                                  long long total, float* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const float t = thr[i / N];
    const bool sel = neg[i] && (scores[i] <= t);
    out[i] = (pos[i] ? 1.f : 0.f) + (sel ? 1.f : 0.f);  // model.py:196 float(pos) + selected_neg
  }
}

// ------------------------------------------------------------------ host side
// Probability threshold -> threshold on the logit difference d = x1 - x0.
// The 2-way softmax score exp(x1-m)/(exp(x0-m)+exp(x1-m)) is a monotone function
// of fl(x1-x0) only, so `score > t` is decided exactly by `d > d*`, where d* is the
// largest fp32 difference whose fp32 softmax score is still <= t (found by bisection
// over the fp32 bit patterns with the same TF formula).  No transcendental per pixel.
static float softmax1_from_diff(float d) {
  // x0 = 0, x1 = d (TF: subtract the max, exp, normalise), all in fp32
  const float m = d > 0.f ? d : 0.f;
  const float e0 = expf(0.f - m), e1 = expf(d - m);
  return e1 / (e0 + e1);
}
static inline int32_t ordered_bits(float f) {
  int32_t i;
  memcpy(&i, &f, 4);
  return i >= 0 ? i : (int32_t)(0x80000000u - (uint32_t)i);  // monotone map float -> int
}
static inline float from_ordered_bits(int32_t i) {
  const int32_t r = i >= 0 ? i : (int32_t)(0x80000000u - (uint32_t)i);
  float f;
  memcpy(&f, &r, 4);
  return f;
}
float prob_to_logit_threshold(float t) {
  // largest d with softmax1(d) <= t  (then score > t  <=>  d > d*)
  if (!(t > 0.f)) return -INFINITY;  // every score > t unless score == 0; keep simple: all pass
  if (t >= 1.f) return INFINITY;
  long long lo = ordered_bits(-200.f), hi = ordered_bits(200.f);  // softmax1(lo) = 0 <= t, softmax1(hi) = 1 > t
  while (hi - lo > 1) {
    const long long mid = lo + (hi - lo) / 2;
    if (softmax1_from_diff(from_ordered_bits((int32_t)mid)) <= t) lo = mid;
    else hi = mid;
  }
  return from_ordered_bits((int32_t)lo);
}

template <int VARIANT, int TERM, bool GRAD, bool FLAGS>
static int launch_main_one(int grid, cudaStream_t s, const MainArgs& a, int B, int N) {
  auto kern = loss_main_kernel<VARIANT, TERM, GRAD, FLAGS>;
  static SmemOptIn optin;
  if (int rc = ensure_dynamic_smem(optin, kern, kMainSmem)) return rc;
  return launch(kern, grid, kMainBlock, kMainSmem, s, a, B, N);
}

template <int VARIANT, int TERM>
static int launch_main(bool grad, bool flags, int grid, cudaStream_t s, const MainArgs& a, int B, int N) {
  if (grad && flags) return launch_main_one<VARIANT, TERM, true, true>(grid, s, a, B, N);
  if (grad) return launch_main_one<VARIANT, TERM, true, false>(grid, s, a, B, N);
  if (flags) return launch_main_one<VARIANT, TERM, false, true>(grid, s, a, B, N);
  return launch_main_one<VARIANT, TERM, false, false>(grid, s, a, B, N);
}

static inline bool select_uses_cluster(int N) { return N <= kClusterSize * kClThreads * kClMaxKPT; }

// K0 + K1
template <int KEYMODE, bool FROM_SCORES>
static int launch_keys_and_select(const float* pix_logits, const float* pix_lab, const float* link_lab,
                                  const float* scores, const uint8_t* pos, const uint8_t* neg,
                                  const int* n_pos_override, int B, int N, int ratio, uint32_t* keys, int2* counts,
                                  ImageInfo* info, float* thr_out, uint8_t* mask, LossHeader* hdr, bool fuse_counts,
                                  bool head_pdl, cudaStream_t s) {
  int rc;
  if (select_uses_cluster(N)) {
    // cluster form: scores, keys, threshold (and for the loss: mask + normalisers) in one launch
#define PLH_CLUSTER(KPT)                                                                                         \
  {                                                                                                              \
    const bool fused = !FROM_SCORES && fuse_counts;                                                            \
    auto kern = fused ? ohem_select_cluster_kernel<KEYMODE, FROM_SCORES, KPT, !FROM_SCORES>                      \
                      : ohem_select_cluster_kernel<KEYMODE, FROM_SCORES, KPT, false>;                            \
    const size_t smem = fused ? (size_t)(KPT < 8 ? KPT : 8) * kClThreads * 32 : 0;                               \
    static SmemOptIn optin; /* only the fused kernel asks for more than 48 KB */                               \
    if ((rc = ensure_dynamic_smem(optin, kern, smem))) return rc;                                                \
    /* first kernel of the chain: plain stream order unless the caller vouches for the predecessor (its CTAs   \
       parked under a foreign kernel only get in that kernel's way) */                                        \
    rc = head_pdl ? launch(kern, B * kClusterSize, kClThreads, smem, s, pix_logits, pix_lab, link_lab, scores, pos, neg, \
                n_pos_override, N, ratio, info, thr_out, mask, FROM_SCORES ? (uint32_t*)nullptr : keys, hdr, 1) \
         : launch_plain(kern, B * kClusterSize, kClThreads, smem, s, pix_logits, pix_lab, link_lab, scores, pos, neg, \
                n_pos_override, N, ratio, info, thr_out, mask, FROM_SCORES ? (uint32_t*)nullptr : keys, hdr, 0); \
  }
    if (N <= kClusterSize * kClThreads * 2) PLH_CLUSTER(2)
    else if (N <= kClusterSize * kClThreads * 8) PLH_CLUSTER(8)
    else PLH_CLUSTER(kClMaxKPT)
#undef PLH_CLUSTER
    return rc;
  }
  const int per_image = std::max(1, std::min({(N + kKeysThreads - 1) / kKeysThreads, (kNumSMs * 8 + B - 1) / B,
                                              kKeysMaxCTAsPerImage}));
  rc = launch_plain(score_keys_kernel<KEYMODE, FROM_SCORES>, dim3(per_image, B), kKeysThreads, 0, s, pix_logits, pix_lab,
              scores, pos, neg, N, keys, counts);
  if (rc) return rc;
  const bool use_smem = (size_t)N * 4 <= kSmemKeysMaxBytes;
  static SmemOptIn optin;
  if (use_smem && (rc = ensure_dynamic_smem(optin, ohem_select_kernel, kSmemKeysMaxBytes))) return rc;
  return launch(ohem_select_kernel, B, kSelectThreads, use_smem ? (size_t)N * 4 : 0, s, keys, counts, per_image,
                n_pos_override, N, ratio, KEYMODE, use_smem ? 1 : 0, info, thr_out, hdr);
}

#ifdef PLH_TIMELINE
int tl_set_loss(unsigned long long* p) { return tl_set_ptr(p); }
#endif

}  // namespace plh

using namespace plh;

extern "C" int plh_pixellink_loss(const float* pix_logits, const float* link_logits, const float* pix_lab,
                                  const float* link_lab, const float* /*train_mask: never read, model.py Q3*/,
                                  int B, int H, int W, const plh_loss_params* p, float* stats, float* grad_pix,
                                  float* grad_link, uint8_t* ohem_mask, uint16_t* decode_flags,
                                  const plh_decode_params* dp, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  if (!pix_logits || !link_logits || !pix_lab || !link_lab || !p || !stats) return PLH_E_NULL;
  if ((grad_pix == nullptr) != (grad_link == nullptr)) return PLH_E_NULL;
  if (decode_flags && !dp) return PLH_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || (long long)B * H * W > (1ll << 29)) return PLH_E_SHAPE;
  if (p->variant < 0 || p->variant > 2 || p->term < 0 || p->term > 1 || p->neg_pos_ratio < 0) return PLH_E_PARAM;
  if (!aligned16(pix_logits) || !aligned16(link_logits) || !aligned16(pix_lab) || !aligned16(link_lab) ||
      !aligned16(stats) || (grad_pix && (!aligned16(grad_pix) || !aligned16(grad_link))) || !aligned16(workspace) ||
      (ohem_mask && !aligned16(ohem_mask)))
    return PLH_E_ALIGN;
  const int N = H * W;
  const LossWsLayout l = loss_ws_layout(B, N);
  if (!workspace || workspace_bytes < l.total) return PLH_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  LossHeader* hdr = (LossHeader*)(ws + l.header);
  ImageInfo* info = (ImageInfo*)(ws + l.info);
  uint8_t* mask = ohem_mask ? ohem_mask : (uint8_t*)(ws + l.mask);
  uint32_t* keys = (uint32_t*)(ws + l.keys);
  int2* counts = (int2*)(ws + l.counts);
  const int total_px = B * N;
  cudaError_t e;
  int rc = PLH_OK;
  const bool main_only = (p->reserved[0] & 1) != 0;  // K0-K2 already ran on this workspace
  const bool fuse_counts = (p->reserved[0] & 2) == 0;  // bit 1: keep the normalisers in their own pass (K2)
  const bool head_pdl = (p->reserved[0] & 4) != 0;     // bit 2: the stream's preceding kernel is this library's
  // K0 + K1 (not needed for the positives-only variant: there is no mining, vgg16 :265)
  if (main_only) {
  } else if (p->variant == PLH_VARIANT_MODEL)
    rc = launch_keys_and_select<KEYS_MODEL, false>(pix_logits, pix_lab, link_lab, nullptr, nullptr, nullptr, nullptr, B,
                                                   N, p->neg_pos_ratio, keys, counts, info, stats + PLH_ST_THR, mask,
                                                   hdr, fuse_counts, head_pdl, s);
  else if (p->variant == PLH_VARIANT_PIXELLINK)
    rc = launch_keys_and_select<KEYS_PIXELLINK, false>(pix_logits, pix_lab, link_lab, nullptr, nullptr, nullptr,
                                                       nullptr, B, N, p->neg_pos_ratio, keys, counts, info,
                                                       stats + PLH_ST_THR, mask, hdr, fuse_counts, head_pdl, s);
  else {
    e = cudaMemsetAsync(stats + PLH_ST_THR, 0xff, sizeof(float) * B, s);  // thr[b] = NaN
    // header and the ImageInfo rows (adjacent in the workspace): K2 accumulates into the rows
    if (e == cudaSuccess) e = cudaMemsetAsync(ws + l.header, 0, l.counts - l.header, s);
    rc = e == cudaSuccess ? PLH_OK : (int)e;
  }
  if (rc) return rc;
  // K2 (the cluster form of K1 has already done this)
  if (!main_only && (p->variant == PLH_VARIANT_POS_ONLY || !select_uses_cluster(N) || !fuse_counts)) {
    const int per_image = std::max(1, std::min((N + kCountsThreads - 1) / kCountsThreads, (kNumSMs * 2 + B - 1) / B));
    const dim3 grid(per_image, B);
    if (p->variant == PLH_VARIANT_MODEL)
      rc = launch(ohem_counts_kernel<PLH_VARIANT_MODEL>, grid, kCountsThreads, 0, s, keys, pix_lab, link_lab, info, N, mask);
    else if (p->variant == PLH_VARIANT_POS_ONLY)
      rc = launch(ohem_counts_kernel<PLH_VARIANT_POS_ONLY>, grid, kCountsThreads, 0, s, keys, pix_lab, link_lab, info, N, mask);
    else
      rc = launch(ohem_counts_kernel<PLH_VARIANT_PIXELLINK>, grid, kCountsThreads, 0, s, keys, pix_lab, link_lab, info, N, mask);
    if (rc) return rc;
  }
  // K3
  {
    MainArgs a;
    a.pix_logits = pix_logits, a.link_logits = link_logits, a.pix_lab = pix_lab, a.link_lab = link_lab;
    a.mask = mask, a.hdr = hdr, a.info = info, a.stats = stats;
    a.ts = nullptr;
    a.grad_pix = grad_pix, a.grad_link = grad_link, a.flags = decode_flags;
    a.total_px = total_px, a.alpha = p->focal_alpha, a.gamma = p->focal_gamma;
    a.tp_logit = dp ? prob_to_logit_threshold(dp->pixel_thresh) : 0.f;
    a.tl_logit = dp ? prob_to_logit_threshold(dp->link_thresh) : 0.f;
    const long long Q = (long long)total_px * 4;
    const int grid = (int)std::min<long long>((Q + kMainThreads * 2 - 1) / (kMainThreads * 2), kMainMaxCTAs);
    const bool g = grad_pix != nullptr, f = decode_flags != nullptr;
    const bool prof = g_prof_n >= 0 && g_prof_n < g_prof_created / 2;
    if (prof) {
      a.ts = g_prof_ts + 2 * g_prof_n;
      cudaEventRecord(g_prof_ev[2 * g_prof_n], s);
    }
#define PLH_DISPATCH(V, T) rc = launch_main<V, T>(g, f, grid, s, a, B, N)
    if (p->term == PLH_TERM_CE) {
      if (p->variant == PLH_VARIANT_MODEL) PLH_DISPATCH(PLH_VARIANT_MODEL, PLH_TERM_CE);
      else if (p->variant == PLH_VARIANT_POS_ONLY) PLH_DISPATCH(PLH_VARIANT_POS_ONLY, PLH_TERM_CE);
      else PLH_DISPATCH(PLH_VARIANT_PIXELLINK, PLH_TERM_CE);
    } else {
      if (p->variant == PLH_VARIANT_MODEL) PLH_DISPATCH(PLH_VARIANT_MODEL, PLH_TERM_FOCAL);
      else if (p->variant == PLH_VARIANT_POS_ONLY) PLH_DISPATCH(PLH_VARIANT_POS_ONLY, PLH_TERM_FOCAL);
      else PLH_DISPATCH(PLH_VARIANT_PIXELLINK, PLH_TERM_FOCAL);
    }
#undef PLH_DISPATCH
    if (rc) return rc;
    if (prof) {
      cudaEventRecord(g_prof_ev[2 * g_prof_n + 1], s);
      ++g_prof_n;
    }
  }
  return PLH_OK;
}

extern "C" int plh_profile_begin(int max_launches) {
  if (max_launches <= 0 || max_launches > kProfMax) return PLH_E_PARAM;
  for (; g_prof_created < 2 * max_launches; ++g_prof_created) {
    cudaError_t e = cudaEventCreate(&g_prof_ev[g_prof_created]);
    if (e != cudaSuccess) return (int)e;
  }
  if (!g_prof_ts && cudaMalloc(&g_prof_ts, sizeof(unsigned long long) * 2 * kProfMax) != cudaSuccess) return PLH_E_DEVICE;
  {
    std::vector<unsigned long long> init(2 * (size_t)max_launches);
    for (int i = 0; i < max_launches; ++i) init[2 * i] = ~0ull, init[2 * i + 1] = 0ull;
    cudaError_t e = cudaMemcpy(g_prof_ts, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
  }
  g_prof_n = 0;
  return PLH_OK;
}

// sum over the profiled launches of (last CTA end - first CTA start); call before plh_profile_end
extern "C" int plh_profile_kernel_window(float* total_ms) {
  if (!total_ms) return PLH_E_NULL;
  if (g_prof_n < 0 || !g_prof_ts) return PLH_E_PARAM;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return (int)e;
  std::vector<unsigned long long> h(2 * (size_t)std::max(g_prof_n, 1));
  e = cudaMemcpy(h.data(), g_prof_ts, sizeof(unsigned long long) * 2 * g_prof_n, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return (int)e;
  double ns = 0.0;
  for (int i = 0; i < g_prof_n; ++i)
    if (h[2 * i + 1] > h[2 * i]) ns += (double)(h[2 * i + 1] - h[2 * i]);
  *total_ms = (float)(ns * 1e-6);
  return PLH_OK;
}

extern "C" int plh_profile_end(float* total_ms, int* n_launches) {
  if (!total_ms || !n_launches) return PLH_E_NULL;
  if (g_prof_n < 0) return PLH_E_PARAM;
  double tot = 0.0;
  for (int i = 0; i < g_prof_n; ++i) {
    cudaError_t e = cudaEventSynchronize(g_prof_ev[2 * i + 1]);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, g_prof_ev[2 * i], g_prof_ev[2 * i + 1]);
    if (e != cudaSuccess) { g_prof_n = -1; return (int)e; }
    tot += ms;
  }
  *total_ms = (float)tot;
  *n_launches = g_prof_n;
  g_prof_n = -1;
  return PLH_OK;
}

extern "C" int plh_ohnm_batch(const float* scores, const uint8_t* pos_mask, const uint8_t* neg_mask,
                              const int32_t* n_pos, int B, int N, int variant, int neg_pos_ratio,
                              float* selected_mask, float* thr, void* workspace, size_t workspace_bytes,
                              void* stream) {
  if (!scores || !pos_mask || !neg_mask || !selected_mask || !thr) return PLH_E_NULL;
  if (B <= 0 || N <= 0 || (long long)B * N > (1ll << 29)) return PLH_E_SHAPE;
  if (variant != PLH_VARIANT_MODEL && variant != PLH_VARIANT_PIXELLINK) return PLH_E_PARAM;
  // workspace: ImageInfo[B] | counts[B * kKeysMaxCTAsPerImage] | keys[B*N]
  const size_t info_bytes = align_up(sizeof(ImageInfo) * (size_t)B, 256);
  const size_t cnt_bytes = align_up(sizeof(int2) * (size_t)B * kKeysMaxCTAsPerImage, 256);
  if (!workspace || !aligned16(workspace) || workspace_bytes < info_bytes + cnt_bytes + (size_t)B * N * 4)
    return PLH_E_WORKSPACE;
  ImageInfo* info = (ImageInfo*)workspace;
  int2* counts = (int2*)((char*)workspace + info_bytes);
  uint32_t* keys = (uint32_t*)((char*)workspace + info_bytes + cnt_bytes);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = variant == PLH_VARIANT_MODEL
               ? launch_keys_and_select<KEYS_MODEL, true>(nullptr, nullptr, nullptr, scores, pos_mask, neg_mask, n_pos, B,
                                                          N, neg_pos_ratio, keys, counts, info, thr, nullptr, nullptr, false, false, s)
               : launch_keys_and_select<KEYS_PIXELLINK, true>(nullptr, nullptr, nullptr, scores, pos_mask, neg_mask,
                                                              n_pos, B, N, neg_pos_ratio, keys, counts, info, thr,
                                                              nullptr, nullptr, false, false, s);
  if (rc) return rc;
  const long long total = (long long)B * N;
  const int grid = (int)std::min<long long>((total + 255) / 256, kNumSMs * 8);
  ohnm_apply_kernel<<<grid, 256, 0, s>>>(scores, pos_mask, neg_mask, thr, N, total, selected_mask);
  return launch_status();
}
