// decode.cu — PixelLink inference decode for sm_100a.
//
// Replaces test_pixellink_fast.py:110-202 / test_pixellink.py:107-217 (thresholds,
// directed 8-neighbour link graph from interior pixels, connected components,
// size filter, per-component minAreaRect -> boxPoints -> int0) and
// tool/pixellink_fn.py:120-154 pixel_detect.
//
// Pipeline (one stream, no host sync):
//   D0 flags      logits -> uint16 flags (bit d = link_d score > thr_l, bit 8 = pixel
//                 score > thr_p), decided in logit-difference space (no exp), and
//                 parent[v] = v / -1, size[v] = 0.       reads 72 B/px, writes 10 B/px
//                 (or, in the fused path, flags come from the loss kernel and D0'
//                 only initialises parent/size)
//   D1a tile cc   one CTA per 32x16 tile: union-find in SHARED memory over the tile's
//                 intra-tile edges (flags staged with a 1-pixel halo), hooking the larger
//                 index under the smaller (atomicMin) -> tile root = min pixel index
//   D1b cross     only tile-border pixels: lock-free global union of the tile roots
//                 across tile boundaries (atomic pointer jumping in L2)
//   D2 flatten    parent[v] = root, component sizes (warp-aggregated atomics),
//                 border pixels that no edge reaches are dropped (not graph nodes)
//   D3 roots      roots with size > min_size take a box slot (ascending-label order is
//                 restored by ranking in D5)
//   D4 labels     final label map + per-(component,row) min/max x (run-end atomics)
//   D5 rects      one CTA per component: row extremes -> sort -> OpenCV-exact hull +
//                 rotating calipers -> 4 integer corners (rect.cuh)
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "rect.cuh"

namespace plh {

float prob_to_logit_threshold(float t);  // loss.cu

// neighbour table (tool/pixellink_fn.py:93-108; test_pixellink_fast.py:124-146)
__constant__ int c_dy[8] = {0, 1, -1, 0, 1, -1, -1, 1};
__constant__ int c_dx[8] = {-1, -1, -1, 1, 1, 1, 0, 0};

constexpr int kFlagP = 1 << 8;

// run-segment record handed to the box kernel: component label | row | first x | last x  (H <= 1024, W <= 2048)
__device__ __forceinline__ unsigned long long make_rec(int root, int y, int x0, int x1) {
  return ((unsigned long long)(unsigned)root << 32) | ((unsigned long long)y << 22) | ((unsigned long long)x0 << 11) |
         (unsigned long long)x1;
}

struct DecodeWsLayout {
  size_t flags, parent, size, comp_root, comp_size, nrec, recs, total;
};

static DecodeWsLayout decode_ws_layout(int B, int H, int W, int K) {
  DecodeWsLayout l;
  const size_t px = (size_t)B * H * W;
  size_t off = 0;
  l.flags = off; off = align_up(off + px * 2, 256);
  l.parent = off; off = align_up(off + px * 4, 256);
  l.size = off; off = align_up(off + px * 4, 256);
  l.comp_root = off; off = align_up(off + (size_t)B * K * 4, 256);
  l.comp_size = off; off = align_up(off + (size_t)B * K * 4, 256);
  l.nrec = off; off = align_up(off + (size_t)B * 8 * 4, 256);   // one counter per (image, record sub-list)
  l.recs = off; off = align_up(off + px * 8, 256);   // one run-boundary record per pixel at most
  l.total = off;
  return l;
}
size_t decode_workspace_bytes(int B, int H, int W, int K) { return decode_ws_layout(B, H, W, K).total; }

// ------------------------------------------------------------------ D0: flags from logits (+ init)
// Work unit = 32 consecutive pixels per warp: four link iterations (lane = pixel-in-iteration x quarter,
// one 128-bit load of two directions' logits each) and one pixel phase (lane = pixel: its 2 pixel logits,
// the 16-bit flag word store).  ~3 instructions per pixel: the kernel is bound by the 72 B/px it reads.
// PLANES: the flag words leave as BIT PLANES for the resident component kernel instead (W % 32 == 0: a unit of
// 32 consecutive pixels is then one word of a row): planes[b][y][k][w], k = 0..7 link_k passes, k = 8 pixel
// passes, w = word of the row (rows of a strip are contiguous); 1.125 B/px instead of 2 B/px.
template <bool PLANES>
__global__ void __launch_bounds__(256)
decode_flags_kernel(const float* __restrict__ pix_logits, const float* __restrict__ link_logits, long long total_px,
                    float tp_logit, float tl_logit, uint16_t* __restrict__ flags, int* __restrict__ n_boxes, int B,
                    unsigned* __restrict__ planes, int words_per_row) {
  pdl_wait_and_release();
  tl_start(4);
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < B; i += blockDim.x) n_boxes[i] = 0;
  const int j = lane & 3, qp = lane >> 2;
  const int total = (int)total_px;
  const int nunits = (total + 31) >> 5;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const float4* ll4 = reinterpret_cast<const float4*>(link_logits);
  const float2* pl2 = reinterpret_cast<const float2*>(pix_logits);
  for (int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < nunits; u += nwarps) {
    const int px0 = u << 5;
    float4 L[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int px = px0 + it * 8 + qp;
      L[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (px < total) L[it] = ldg_stream4(ll4 + ((size_t)px * 4 + j));
    }
    const int pp = px0 + lane;
    float2 P = make_float2(0.f, 0.f);
    if (pp < total) P = ldg_stream2(pl2 + pp);
    unsigned mine = 0;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      unsigned bits = ((L[it].y - L[it].x) > tl_logit ? 1u : 0u) << (2 * j) |
                      ((L[it].w - L[it].z) > tl_logit ? 1u : 0u) << (2 * j + 1);
      bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
      bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
      // pixel-phase lane l owns pixel (l >> 3) * 8 + (l & 7): iteration l >> 3, quad l & 7
      const unsigned got = __shfl_sync(0xffffffffu, bits, (lane & 7) << 2);
      if ((lane >> 3) == it) mine = got;
    }
    const unsigned fl = pp < total ? (mine | ((P.y - P.x) > tp_logit ? kFlagP : 0)) : 0u;
    if (PLANES) {
      unsigned out = 0;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const unsigned bk = __ballot_sync(0xffffffffu, (fl >> k) & 1u);
        out = lane == k ? bk : out;
      }
      const int row = u / words_per_row;  // global row index b * H + y
      if (lane < 9) planes[((size_t)row * 9 + lane) * words_per_row + (u - row * words_per_row)] = out;
    } else if (pp < total) {
      flags[pp] = (uint16_t)fl;
    }
  }
  tl_end(4);
}

// ------------------------------------------------------------------ D1: union-find
__device__ __forceinline__ int find_root(int* parent, int v) {
  volatile int* P = parent;
  int p = P[v];
  while (p != v) {
    const int gp = P[p];
    if (gp != p) P[v] = gp;  // path halving; benign race (pointers only move rootwards)
    v = p;
    p = gp;
  }
  return v;
}

// Read-only find (no path compression): used by the flatten pass, where a compressing
// store could overwrite another thread's final `parent[v] = root`.
__device__ __forceinline__ int find_root_ro(const int* parent, int v) {
  const volatile int* P = parent;
  int p = P[v];
  while (p != v) {
    v = p;
    p = P[v];
  }
  return v;
}

__device__ __forceinline__ void unite(int* parent, int a, int b) {
  while (true) {
    a = find_root(parent, a);
    b = find_root(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }  // hook the larger root under the smaller
    const int old = atomicMin(&parent[a], b);
    if (old == a) return;
    a = old;  // somebody re-parented a meanwhile: retry from there
  }
}

// Edge set of the reference (test_pixellink_fast.py:119-146): only INTERIOR pixels
// (1 <= x <= W-2, 1 <= y <= H-2) that pass the pixel threshold emit edges, to the
// neighbour in direction d iff link_d passes AND the neighbour passes the pixel
// threshold.  Connectivity is taken undirected (weakly-connected components).
constexpr int kTW = 32, kTH = 16;  // tile: 512 pixels, one thread each

// Each undirected neighbour pair is visited once, from its first pixel in row-major order, through the
// four "forward" directions; the pair is connected if either endpoint emits the edge:
//   v -> u via direction d (v interior, link_d[v])   or   u -> v via opp(d) (u interior, link_opp(d)[u]).
__constant__ int c_fwd[4] = {3, 1, 7, 4};   // right, left_down, down, right_down
__constant__ int c_opp[4] = {0, 5, 6, 2};   // left,  right_up,  up,   left_up

__device__ __forceinline__ int find_root_s(volatile int* lab, int v) {
  int p = lab[v];
  while (p != v) {
    const int gp = lab[p];
    if (gp != p) lab[v] = gp;
    v = p;
    p = gp;
  }
  return v;
}

__device__ __forceinline__ void unite_s(int* lab, int a, int b) {
  while (true) {
    a = find_root_s(lab, a);
    b = find_root_s(lab, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&lab[a], b);
    if (old == a) return;
    a = old;
  }
}

// FROM_LOGITS: the tile thresholds its own pixels (D0 folded in: one thread = one pixel = 72 B of logits,
// a warp = 32 consecutive pixels of a row) and also writes the flag words for the later passes.
template <bool FROM_LOGITS>
__global__ void __launch_bounds__(kTW * kTH)
decode_tile_cc_kernel(const float* __restrict__ pix_logits, const float* __restrict__ link_logits, float tp_logit,
                      float tl_logit, uint16_t* __restrict__ flags, int H, int W, int* __restrict__ parent,
                      int* __restrict__ size, int* __restrict__ n_boxes, int* __restrict__ nrec) {
  pdl_wait_and_release();
  tl_start(5);
  __shared__ uint16_t sf[kTH + 2][kTW + 2];
  __shared__ int slab[kTW * kTH];
  const int tid = threadIdx.x;
  const int tx0 = blockIdx.x * kTW, ty0 = blockIdx.y * kTH, b = blockIdx.z;
  const size_t base = (size_t)b * H * W;
  if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) n_boxes[b] = 0, nrec[b] = 0;  // (tiled form: one record list per image)
  const int ly = tid / kTW, lx = tid - ly * kTW;
  const int gy = ty0 + ly, gx = tx0 + lx;
  const bool inimg = gy < H && gx < W;
  if (FROM_LOGITS) {
    // A warp is one tile row = 32 consecutive pixels.  Link logits are read like decode_flags_kernel does:
    // four iterations, lane = (pixel-in-iteration, quarter), one fully coalesced 128-bit load each (512 B
    // per warp instruction); the quad's bits are OR-ed and handed to the lane that owns the pixel.
    static_assert(kTW == 32, "one warp per tile row");
    const int lane = tid & 31;
    const int j = lane & 3, qp = lane >> 2;
    const bool rowin = gy < H;
    const size_t row0 = base + (size_t)(rowin ? gy : 0) * W + tx0;  // first pixel of this warp's row
    const float4* ll4 = reinterpret_cast<const float4*>(link_logits);
    float4 L[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int px = it * 8 + qp;
      L[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rowin && tx0 + px < W) L[it] = ldg_stream4(ll4 + ((row0 + px) * 4 + j));
    }
    float2 pp = make_float2(0.f, 0.f);
    if (inimg) pp = ldg_stream2(reinterpret_cast<const float2*>(pix_logits) + row0 + lx);
    unsigned fl = 0;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      unsigned bits = ((L[it].y - L[it].x) > tl_logit ? 1u : 0u) << (2 * j) |
                      ((L[it].w - L[it].z) > tl_logit ? 1u : 0u) << (2 * j + 1);
      bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
      bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
      // pixel lane l owns pixel (l >> 3) * 8 + (l & 7): iteration l >> 3, quad l & 7
      const unsigned got = __shfl_sync(0xffffffffu, bits, (lane & 7) << 2);
      if ((lane >> 3) == it) fl = got;
    }
    fl = inimg ? (fl | ((pp.y - pp.x) > tp_logit ? kFlagP : 0u)) : 0u;
    if (inimg) flags[row0 + lx] = (uint16_t)fl;
    sf[ly + 1][lx + 1] = (uint16_t)fl;
  } else {
    for (int i = tid; i < (kTH + 2) * (kTW + 2); i += kTW * kTH) {
      const int r = i / (kTW + 2), c = i - r * (kTW + 2);
      const int yy = ty0 - 1 + r, xx = tx0 - 1 + c;
      sf[r][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? flags[base + (size_t)yy * W + xx] : (uint16_t)0;
    }
  }
  __syncthreads();
  // Run-based labelling inside the tile.  A warp is one tile row: pixel x continues the run of x-1 iff the pair
  // is linked (either endpoint emits the edge) — one ballot gives every lane its run head; the union-find nodes
  // are the RUN HEADS, and only the three downward directions need unions, at most one per pair of runs (an
  // edge is dropped when the lane to its left has the same edge between the same two runs, or when the
  // previous direction already reaches the same run below).  A per-pixel union-find did up to four unions
  // per pixel, and the tile's barrier waited for the thread with the longest chain.
  __shared__ unsigned s_cl[kTH];
  static_assert(kTW == 32, "one warp per tile row");
  const int lane = tid & 31;
  const unsigned f = sf[ly + 1][lx + 1];
  const bool P = inimg && (f & kFlagP);
  const bool yin = gy >= 1 && gy <= H - 2;
  const bool vin = yin && gx >= 1 && gx <= W - 2;
  const unsigned fl = sf[ly + 1][lx];   // left neighbour (for lx == 0 it lies in the next tile: decode_cross_kernel)
  const bool lin = yin && gx >= 2 && gx <= W - 1;
  const bool cl = P && lx > 0 && (fl & kFlagP) && ((vin && (f & 1u)) || (lin && (fl & 8u)));
  const unsigned m = __ballot_sync(0xffffffffu, cl);
  const int hv = ly * kTW + 31 - __clz(~m & (0xffffffffu >> (31 - lane)));   // bit 0 of m is never set
  slab[tid] = P ? hv : -1;
  if (lane == 0) s_cl[ly] = m;
  __syncthreads();
  {
    // downward edges (all lanes take part in the ballots): left_down 1 <-> right_up 5, down 7 <-> up 6, right_down 4 <-> left_up 2
    const bool row = ly + 1 < kTH;      // the row below is in this tile, else decode_cross_kernel
    const bool y1in = gy + 1 <= H - 2;
    const unsigned fu1 = (P && row && lx >= 1) ? sf[ly + 2][lx] : 0u;
    const unsigned fu7 = (P && row) ? sf[ly + 2][lx + 1] : 0u;
    const unsigned fu4 = (P && row && lx + 1 < kTW) ? sf[ly + 2][lx + 2] : 0u;
    const bool u1in = y1in && gx - 1 >= 1 && gx - 1 <= W - 2, u7in = y1in && gx >= 1 && gx <= W - 2, u4in = y1in && gx + 1 <= W - 2;
    const bool e1 = (fu1 & kFlagP) && ((vin && (f & (1u << 1))) || (u1in && (fu1 & (1u << 5))));
    const bool e7 = (fu7 & kFlagP) && ((vin && (f & (1u << 7))) || (u7in && (fu7 & (1u << 6))));
    const bool e4 = (fu4 & kFlagP) && ((vin && (f & (1u << 4))) || (u4in && (fu4 & (1u << 2))));
    const unsigned E1 = __ballot_sync(0xffffffffu, e1), E7 = __ballot_sync(0xffffffffu, e7), E4 = __ballot_sync(0xffffffffu, e4);
    if (E1 | E7 | E4) {
      const unsigned md = row ? s_cl[ly + 1] : 0u;                  // run continuation bits of the row below
      const bool cu1 = lane >= 1 && ((md >> (lane - 1)) & 1u), cu7 = (md >> lane) & 1u, cu4 = lane < 31 && ((md >> (lane + 1)) & 1u);
      const bool prev = lane > 0 && cl;
      const bool s1 = prev && ((E1 >> (lane - 1)) & 1u) && cu1;
      const bool s7 = (prev && ((E7 >> (lane - 1)) & 1u) && cu7) || (e1 && cu7);
      const bool s4 = (prev && ((E4 >> (lane - 1)) & 1u) && cu4) || (e7 && cu4);
      auto head_below = [&](int x) { return (ly + 1) * kTW + 31 - __clz(~md & (0xffffffffu >> (31 - x))); };
      if (e1 && !s1) unite_s(slab, hv, head_below(lane - 1));
      if (e7 && !s7) unite_s(slab, hv, head_below(lane));
      if (e4 && !s4) unite_s(slab, hv, head_below(lane + 1));
    }
  }
  __syncthreads();
  if (inimg) {
    const size_t g = base + (size_t)gy * W + gx;
    int root = -1;
    if (P) {
      const int r = find_root_s(slab, hv);
      const int ry = r / kTW, rx = r - ry * kTW;
      root = (int)(base + (size_t)(ty0 + ry) * W + tx0 + rx);
    }
    parent[g] = root;
    size[g] = 0;
  }
  tl_end(5);
}

// Cross-tile pairs: only pixels on the right / bottom / left border of a tile have a forward
// neighbour in another tile.
__global__ void __launch_bounds__(256)
decode_cross_kernel(const uint16_t* __restrict__ flags, int H, int W, int total_px, int* __restrict__ parent) {
  pdl_wait_and_release();
  tl_start(6);
  const int N = H * W;
  const int stride = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  // Warp-converged loop: a warp covers 32 consecutive pixels of a row — for the bottom row of a tile that
  // is one whole tile edge, whose pixels mostly ask for the SAME union (tile root above, tile root below).
  // Lanes with the same pair of tile roots elect one to do it: up to 32x fewer atomic chains on the two
  // hottest words of the forest.
  for (int g0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); g0 < total_px; g0 += stride) {
    const int g = g0 + lane;
    bool act = g < total_px;
    int x = 0, y = 0;
    unsigned f = 0;
    if (act) {
      const int v = g % N;
      y = v / W, x = v - y * W;
      const int lx = x % kTW, ly = y % kTH;
      act = lx == 0 || lx == kTW - 1 || ly == kTH - 1;
      if (act) {
        f = flags[g];
        act = (f & kFlagP) != 0;
      }
    }
    if (!__any_sync(0xffffffffu, act)) continue;
    const bool vin = x >= 1 && x <= W - 2 && y >= 1 && y <= H - 2;
    // all neighbour flags first, then all tile roots: two L2 round trips for the four directions together
    // instead of a dependent chain per direction
    bool cand[4];
    int uidx[4];
    unsigned fu[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int d = c_fwd[k];
      const int ux = x + c_dx[d], uy = y + c_dy[d];
      cand[k] = act && !(ux < 0 || ux >= W || uy >= H) &&
                !(ux / kTW == x / kTW && uy / kTH == y / kTH);  // intra-tile pairs were done in shared memory
      uidx[k] = g + c_dy[d] * W + c_dx[d];
      fu[k] = cand[k] ? flags[uidx[k]] : 0u;
    }
    int rb[4];
    bool want[4];
    bool any_want = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int d = c_fwd[k];
      const int ux = x + c_dx[d], uy = y + c_dy[d];
      const bool uin = ux >= 1 && ux <= W - 2 && uy >= 1 && uy <= H - 2;
      want[k] = cand[k] && (fu[k] & kFlagP) && ((vin && (f & (1u << d))) || (uin && (fu[k] & (1u << c_opp[k]))));
      rb[k] = want[k] ? parent[uidx[k]] : 0;
      any_want |= want[k];
    }
    // tile roots of the pixels (the tile pass left parent = tile root; later hooks only add hops)
    const int ra = any_want ? parent[g] : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned wm = __ballot_sync(0xffffffffu, want[k]);
      if (want[k]) {
        const unsigned long long key = ((unsigned long long)(unsigned)min(ra, rb[k]) << 32) | (unsigned)max(ra, rb[k]);
        const unsigned peers = __match_any_sync(wm, key);
        if (lane == __ffs(peers) - 1) unite(parent, ra, rb[k]);
      }
    }
  }
  tl_end(6);
}

// ------------------------------------------------------------------ D2: flatten + sizes
__global__ void __launch_bounds__(256)
decode_flatten_kernel(const uint16_t* __restrict__ flags, int H, int W, long long total_px, int* __restrict__ parent,
                      int* __restrict__ size) {
  pdl_wait_and_release();
  tl_start(7);
  const int N = H * W;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long start = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  const int lane = threadIdx.x & 31;
  for (long long w0 = start; w0 < total_px; w0 += stride) {
    const long long g = w0 + lane;
    int root = -1;
    if (g < total_px && (flags[g] & kFlagP)) {
      const int v = (int)(g % N);
      const int y = v / W, x = v - y * W;
      bool node = true;
      if (x < 1 || x > W - 2 || y < 1 || y > H - 2) {
        // a border pixel is a graph node only if some interior neighbour links to it
        node = false;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          const int sy = y - c_dy[d], sx = x - c_dx[d];  // source s with s + off(d) == this pixel
          if (sx >= 1 && sx <= W - 2 && sy >= 1 && sy <= H - 2) {
            const unsigned fs = flags[g - (long long)c_dy[d] * W - c_dx[d]];
            if ((fs & kFlagP) && (fs & (1u << d))) node = true;
          }
        }
      }
      if (node) {
        root = find_root_ro(parent, (int)g);
        parent[g] = root;
      } else {
        parent[g] = -1;
      }
    }
    // warp-aggregated size count
    const unsigned active = __ballot_sync(0xffffffffu, root >= 0);
    if (root >= 0) {
      const unsigned peers = __match_any_sync(active, root);
      if ((__ffs(peers) - 1) == lane) atomicAdd(&size[root], __popc(peers));
    }
  }
  tl_end(7);
}

// ------------------------------------------------------------------ D4: labels, components, run records
// One pass over the flattened forest does everything the boxes need:
//  * label map: the component's minimum pixel index if its size > min_size, else -1;
//  * every kept root takes a box slot (arrival order; D5 restores the ascending-label order by ranking);
//  * every kept pixel that starts or ends a horizontal run of its component appends one record
//    (root, y, x) to its image's list: the row extremes D5 needs are the min / max x over the records
//    of a (component, row), and a list needs no initialisation, unlike a per-slot row table.
__global__ void __launch_bounds__(256)
decode_labels_kernel(const int* __restrict__ parent, const int* __restrict__ size, int H, int W, long long total_px,
                     int min_size, int K, int32_t* __restrict__ labels, int* __restrict__ comp_root,
                     int* __restrict__ comp_size, int* __restrict__ n_boxes, int* __restrict__ nrec,
                     unsigned long long* __restrict__ recs) {
  pdl_wait_and_release();
  tl_start(9);
  const int N = H * W;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long start = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  for (long long w0 = start; w0 < total_px; w0 += stride) {
    const long long g = w0 + lane;
    int b = -1;
    bool emit = false;
    unsigned long long rec = 0;
    if (g < total_px) {
      b = (int)(g / N);
      const int v = (int)(g - (long long)b * N);
      const int y = v / W, x = v - y * W;
      const int r = parent[g];
      int sz = 0;
      if (r >= 0) sz = size[r];
      const bool kept = r >= 0 && sz > min_size;  // test_pixellink_fast.py:174 `len(index_list) > 10`
      const int rl = r - b * N;
      labels[g] = kept ? rl : -1;
      if (kept) {
        const bool run_start = (x == 0) || (parent[g - 1] != r);
        const bool run_end = (x == W - 1) || (parent[g + 1] != r);
        emit = run_start || run_end;
        rec = make_rec(rl, y, x, x);
        if (r == (int)g) {
          const int slot = atomicAdd(&n_boxes[b], 1);
          if (slot < K) comp_root[(size_t)b * K + slot] = rl, comp_size[(size_t)b * K + slot] = sz;
        }
      }
    }
    // warp-aggregated append: one atomic per (warp, image)
    const unsigned em = __ballot_sync(0xffffffffu, emit);
    if (emit) {
      const unsigned peers = __match_any_sync(em, b);
      const int leader = __ffs(peers) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&nrec[b], __popc(peers));
      base = __shfl_sync(peers, base, leader);
      recs[(size_t)b * N + base + __popc(peers & ((1u << lane) - 1u))] = rec;
    }
  }
  tl_end(9);
}


// ------------------------------------------------------------------ resident form: one cluster per image
// Phases D1a-D4 (component labelling from the threshold bits, sizes, label map, box slots, run records) in
// ONE launch with the whole map ON CHIP: a thread-block cluster of 8 CTAs owns one image, CTA r holds the
// strip of rows [r*RS, (r+1)*RS) (+ a 2-row halo of the bit planes) in its shared memory, and the union-find
// words of the strips form one array in DISTRIBUTED shared memory (every CTA can read / atomicMin any word
// through the cluster window).  The tiled kernels above spend most of their time in launch / ramp / drain
// and dependent L2 round trips, and a pixel-per-thread labelling costs ~340 instructions per pixel.  Here
// the map is held as BIT PLANES (one 32-bit word = 32 pixels of a row; 8 link planes + the pixel plane)
// and everything that can be is word-wide bit arithmetic, one lane per word:
//   1. node filter (border pixels no interior neighbour links to are not graph nodes), then the
//      "continues the run of its left neighbour" mask CL: pixel x joins x-1 iff the pair is linked (either
//      endpoint emits the edge, same rule as the tiled form).  A run = a maximal stretch of CL; its head is
//      its first pixel.  Runs, not pixels, are the union-find nodes;
//   2. row y / row y+1: the edge masks of the three downward directions by shifts and ANDs; an edge is
//      dropped when the bit to its left is the same edge between the same two runs (or the previous
//      direction already reaches the same run), which leaves about one union per PAIR OF RUNS.  The edges go
//      to a list so that all threads share the unions: atomicMin on the run heads (larger index under
//      smaller: root = minimum pixel index), local or remote shared memory alike;
//   3. flatten the heads; sizes accumulate IN the root's own word as -(count) - 1 (a root needs no
//      pointer), one atomic per run segment;
//   4. label map out (one warp per word, coalesced), kept roots take box slots, kept run segments append
//      (root, y, x0, x1) records for the box kernel.
// Cluster barriers (release / acquire) separate the phases that touch other CTAs' words.
constexpr int kImgCluster = 8;
#ifndef PLH_IMG_THREADS
#define PLH_IMG_THREADS 256
#endif
constexpr int kImgThreads = PLH_IMG_THREADS;
constexpr int kImgWarps = kImgThreads / 32;
constexpr size_t kImgSmemMax = 227 * 1024 - 256;
constexpr int kImgEdgeCapMax = 2048, kImgEdgeCapMin = 512;

struct ImgGeom {
  int RS, WW, nl, np, nc, nx;  // rows per strip, words per row, words of: labels / planes / CL (= CH ...) ; cross-edge slots
};
__host__ __device__ inline ImgGeom image_geom(int H, int W) {
  ImgGeom g;
  g.RS = (H + kImgCluster - 1) / kImgCluster;
  g.WW = (W + 31) / 32;
  g.nl = g.RS * W;
  g.np = ((9 * (g.RS + 4) * g.WW + 3) / 4) * 4;
  g.nc = (g.RS + 1) * g.WW;
  g.nx = 3 * W;  // at most three downward edges per pixel of the strip's last row
  return g;
}
__host__ __device__ inline size_t image_smem_base_bytes(int H, int W) {
  const ImgGeom g = image_geom(H, W);
  return 4 * ((size_t)g.nl + g.np + 6 * (size_t)g.nc) + 8 + 8 * (size_t)g.nx;
}
// capacity of the shared-memory edge list: whatever is left, within [kImgEdgeCapMin, kImgEdgeCapMax]; 0 = does not fit
inline int image_edge_cap(int H, int W) {
  const size_t base = image_smem_base_bytes(H, W);
  if (base + (size_t)kImgEdgeCapMin * 8 > kImgSmemMax) return 0;
  return (int)std::min<size_t>(kImgEdgeCapMax, (kImgSmemMax - base) / 8);
}

// bits of word w (pixels 32w .. 32w+31) whose x lies in [lo, hi]
__device__ __forceinline__ unsigned xmask(int lo, int hi, int w) {
  const int a = max(lo - (w << 5), 0), b = min(hi - (w << 5), 31);
  return b < a ? 0u : ((0xffffffffu >> (31 - b)) & (0xffffffffu << a));
}
__device__ __forceinline__ void cluster_barrier() {
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// the union-find words of the image, spread over the cluster: word px lives in CTA px / (RS*W)
struct DLab {
  int* loc;
  unsigned long long magic;  // ceil(2^40 / rsw)
  int rsw;
  __device__ __forceinline__ uint32_t at(int px) const {  // shared::cluster address of word px
    const int r = (int)(((unsigned long long)(unsigned)px * magic) >> 40);
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(loc + (px - r * rsw));
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(r));
    return ra;
  }
  __device__ __forceinline__ int ld(int px) const {
    int v;
    asm volatile("ld.volatile.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(at(px)) : "memory");
    return v;
  }
  __device__ __forceinline__ void st(int px, int v) const {
    asm volatile("st.volatile.shared::cluster.s32 [%0], %1;" ::"r"(at(px)), "r"(v) : "memory");
  }
  __device__ __forceinline__ int atom_min(int px, int v) const {
    int o;
    asm volatile("atom.relaxed.cluster.shared::cluster.min.s32 %0, [%1], %2;" : "=r"(o) : "r"(at(px)), "r"(v) : "memory");
    return o;
  }
  __device__ __forceinline__ void red_add(int px, int v) const {
    asm volatile("red.relaxed.cluster.shared::cluster.add.s32 [%0], %1;" ::"r"(at(px)), "r"(v) : "memory");
  }
  __device__ __forceinline__ int find(int v) const {  // with path halving
    int p = ld(v);
    while (p != v) {
      const int gp = ld(p);
      if (gp != p) st(v, gp);
      v = p;
      p = gp;
    }
    return v;
  }
  __device__ __forceinline__ void unite(int a, int c) const {
    while (true) {
      a = find(a);
      c = find(c);
      if (a == c) return;
      if (a < c) { const int t = a; a = c; c = t; }  // hook the larger root under the smaller
      const int old = atom_min(a, c);
      if (old == a) return;
      a = old;
    }
  }
};

template <bool FROM_PLANES>
__global__ void __cluster_dims__(kImgCluster, 1, 1) __launch_bounds__(kImgThreads, 2)
decode_image_kernel(const uint16_t* __restrict__ flags, const unsigned* __restrict__ planes, int H, int W, int min_size,
                    int K, int edge_cap, unsigned long long magic, int32_t* __restrict__ labels,
                    int* __restrict__ comp_root, int* __restrict__ comp_size, int* __restrict__ n_boxes,
                    int* __restrict__ nrec, unsigned long long* __restrict__ recs) {
  tl_end(23);  // latest CTA arrival (before the dependency wait)
  pdl_wait();
  tl_start(5);
  tl_end(22);  // latest CTA start
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_nedge, s_nxedge, s_nrec;
  const ImgGeom g = image_geom(H, W);
  const int N = H * W, WW = g.WW, RS = g.RS, PR = RS + 4;  // PR: plane rows held (2-row halo each side)
  int* lab = reinterpret_cast<int*>(smem);                 // [RS*W] this strip's union-find words / root counters
  unsigned* PL = reinterpret_cast<unsigned*>(lab + g.nl);  // [PR][9][WW] planes: 0..7 link_d passes, 8 pixel passes (then: is a node)
  unsigned* CL = PL + g.np;                                // [RS+1][WW] continues the run of the left neighbour
  int* CH = reinterpret_cast<int*>(CL + g.nc);             // [RS+1][WW] head of the run that enters the word at bit 0
  unsigned* ER = reinterpret_cast<unsigned*>(CH + g.nc);   // [3][RS+1][WW] de-duplicated downward edges per direction
  int* EO = reinterpret_cast<int*>(ER + 3 * g.nc);         // [RS+1][WW] first edge-list slot of the word
  int2* EX = reinterpret_cast<int2*>((reinterpret_cast<uintptr_t>(EO + g.nc) + 7) & ~(uintptr_t)7);  // [3W] edges into the next strip
  int2* EL = EX + g.nx;                                    // [edge_cap] edges inside the strip
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = blockIdx.x % kImgCluster, b = blockIdx.x / kImgCluster;
  const size_t base = (size_t)b * N;
  const int y0 = rank * RS, y1 = min(H, y0 + RS);          // this strip's rows [y0, y1) (may be empty)
  const int nrows = max(y1 - y0, 0), p00 = y0 * W;         // p00: first pixel of the strip
  const DLab L{lab, magic, RS * W};
  auto pl = [&](int k, int y, int w) -> unsigned& { return PL[((y - y0 + 2) * 9 + k) * WW + w]; };
  if (tid == 0) s_nedge = 0, s_nxedge = 0, s_nrec = 0;
  if (rank == 0 && tid == 0) n_boxes[b] = 0;

  // ---- bit planes of rows y0-2 .. y1+1 into shared memory (copied: the rows of a strip are contiguous; or built
  // from the flag words: one warp per word of 32 pixels); union-find initialised to singletons
  // (strip-relative indices until the strip is flat: the long pointer chains are walked with plain shared-memory loads)
  for (int i = tid; i < nrows * W; i += kImgThreads) lab[i] = i;
  if (FROM_PLANES) {
    const int rw = 9 * WW;                                 // words per row of planes
    const long long off = ((long long)b * H + (y0 - 2)) * rw;
    for (int i = tid; i < PR * rw; i += kImgThreads) {
      const int y = y0 - 2 + i / rw;
      PL[i] = (y >= 0 && y < H) ? __ldg(planes + (off + i)) : 0u;
    }
  } else {
    for (int j = warp; j < PR * WW; j += kImgWarps) {
      const int r = j / WW, y = y0 - 2 + r, w = j - r * WW, x = (w << 5) + lane;
      const unsigned f = (y >= 0 && y < H && x < W) ? flags[base + (size_t)y * W + x] : 0u;
      unsigned mine = 0;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const unsigned bk = __ballot_sync(0xffffffffu, (f >> k) & 1u);
        mine = lane == k ? bk : mine;
      }
      if (lane < 9) PL[(r * 9 + lane) * WW + w] = mine;
    }
  }
  __syncthreads();
  tl_end(24);

  // ---- 1a. node filter for rows y0-1 .. y1: a border pixel is a graph node only if some interior neighbour links to it
  {
    const int ya = max(y0 - 1, 0), yb = min(y1 + 1, H);
    for (int i = tid; i < max(yb - ya, 0) * WW; i += kImgThreads) {
      const int y = ya + i / WW, w = i % WW;
      const unsigned P = pl(8, y, w);
      const unsigned vin = (y >= 1 && y <= H - 2) ? xmask(1, W - 2, w) : 0u;
      if (P & ~vin) {
        unsigned reached = 0;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          const int sy = y - c_dy[d];  // source row: must be interior
          if (sy < 1 || sy > H - 2) continue;
          const unsigned S = pl(8, sy, w) & pl(d, sy, w) & xmask(1, W - 2, w);
          if (c_dx[d] == 0) reached |= S;
          else if (c_dx[d] > 0) reached |= (S << 1) | (w > 0 ? (pl(8, sy, w - 1) & pl(d, sy, w - 1) & xmask(1, W - 2, w - 1)) >> 31 : 0u);
          else reached |= (S >> 1) | (w < WW - 1 ? (pl(8, sy, w + 1) & pl(d, sy, w + 1) & xmask(1, W - 2, w + 1)) << 31 : 0u);
        }
        pl(8, y, w) = P & (vin | reached);  // only border bits change; the sources read above are interior bits
      }
    }
  }
  __syncthreads();
  // ---- 1b. CL for rows y0 .. y1: linked to the left neighbour (v emits "left", plane 0, or the neighbour emits "right", plane 3)
  const int yc = min(y1 + 1, H);  // CL / CH rows [y0, yc)
  for (int i = tid; i < max(yc - y0, 0) * WW; i += kImgThreads) {
    const int y = y0 + i / WW, w = i % WW;
    const unsigned P = pl(8, y, w);
    unsigned cl = 0;
    if (P) {
      const bool yin = y >= 1 && y <= H - 2;
      const unsigned Pl = (P << 1) | (w > 0 ? pl(8, y, w - 1) >> 31 : 0u);
      const unsigned L3l = (pl(3, y, w) << 1) | (w > 0 ? pl(3, y, w - 1) >> 31 : 0u);
      const unsigned vin = yin ? xmask(1, W - 2, w) : 0u, vinl = yin ? xmask(2, W - 1, w) : 0u;
      cl = P & Pl & ((vin & pl(0, y, w)) | (vinl & L3l));
    }
    CL[i] = cl;
  }
  __syncthreads();
  // ---- 1c. head of the run that enters a word from the left (bit 0 of word 0 is never set in CL)
  for (int i = tid; i < max(yc - y0, 0) * WW; i += kImgThreads)
    if (CL[i] & 1u) {
      const int y = y0 + i / WW;
      int ii = i - 1;
      unsigned nz = ~CL[ii];
      while (nz == 0u) nz = ~CL[--ii];
      CH[i] = y * W + ((ii % WW) << 5) + 31 - __clz(nz);
    }
  __syncthreads();
  tl_end(17);
  // run head of pixel `bit` of word w of row y (y in [y0, yc))
  auto head_of = [&](int y, int w, int bit) -> int {
    const int i = (y - y0) * WW + w;
    const unsigned z = ~CL[i] & (0xffffffffu >> (31 - bit));
    return z ? y * W + (w << 5) + 31 - __clz(z) : CH[i];
  };

  // ---- 2. edges between the runs of row y and row y+1, y in this strip (the pair y1-1 / y1 reaches into the
  // next strip).  2a, one lane per word: the three de-duplicated edge masks and the word's slots in the edge list
  for (int i = tid; i < nrows * WW; i += kImgThreads) {
    const int y = y0 + i / WW, w = i % WW;
    unsigned R1 = 0, R7 = 0, R4 = 0;
    const unsigned P = (y + 1 < H) ? pl(8, y, w) : 0u;
    if (P) {
      const bool hasl = w > 0, hasr = w < WW - 1;
      const unsigned Pd = pl(8, y + 1, w), Pd1 = (Pd << 1) | (hasl ? pl(8, y + 1, w - 1) >> 31 : 0u),
                     Pd4 = (Pd >> 1) | (hasr ? pl(8, y + 1, w + 1) << 31 : 0u);
      if (P & (Pd | Pd1 | Pd4)) {
        const bool yin = y >= 1 && y <= H - 2, y1in = y + 1 <= H - 2;
        const unsigned vin = yin ? xmask(1, W - 2, w) : 0u;
        const unsigned vind = y1in ? xmask(1, W - 2, w) : 0u;    // (y+1, x)   interior
        const unsigned vind1 = y1in ? xmask(2, W - 1, w) : 0u;   // (y+1, x-1) interior
        const unsigned vind4 = y1in ? xmask(0, W - 3, w) : 0u;   // (y+1, x+1) interior
        // neighbour's answering link, aligned to this word: left_down 1 <-> right_up 5, down 7 <-> up 6, right_down 4 <-> left_up 2
        const unsigned L5d1 = (pl(5, y + 1, w) << 1) | (hasl ? pl(5, y + 1, w - 1) >> 31 : 0u);
        const unsigned L2d4 = (pl(2, y + 1, w) >> 1) | (hasr ? pl(2, y + 1, w + 1) << 31 : 0u);
        const unsigned E1 = P & Pd1 & ((vin & pl(1, y, w)) | (vind1 & L5d1));
        const unsigned E7 = P & Pd & ((vin & pl(7, y, w)) | (vind & pl(6, y + 1, w)));
        const unsigned E4 = P & Pd4 & ((vin & pl(4, y, w)) | (vind4 & L2d4));
        const int id = i + WW;  // row y+1 in CL / CH
        const unsigned cl = CL[i], cld = CL[id];
        const unsigned cld1 = (cld << 1) | (hasl ? CL[id - 1] >> 31 : 0u), cld4 = (cld >> 1) | (hasr ? CL[id + 1] << 31 : 0u);
        R1 = E1 & ~((E1 << 1) & cl & cld1);
        R7 = E7 & ~((E7 << 1) & cl & cld) & ~(E1 & cld);
        R4 = E4 & ~((E4 << 1) & cl & cld4) & ~(E7 & cld4);
      }
    }
    ER[i] = R1, ER[g.nc + i] = R7, ER[2 * g.nc + i] = R4;
    const int ne = __popc(R1) + __popc(R7) + __popc(R4);
    // the strip's last row pair goes to the cross list (always large enough), the others to the local list
    EO[i] = ne ? atomicAdd(y + 1 == y1 ? &s_nxedge : &s_nedge, ne) : 0;
  }
  __syncthreads();
  // 2b, one warp per word, lane = pixel: the (head, head) pair of every edge into its list, so that ALL threads
  // share the unions afterwards; local edges that do not fit the list are united here (shared-memory atomics)
  const unsigned lt = (1u << lane) - 1u;
  for (int i = warp; i < nrows * WW; i += kImgWarps) {
    const unsigned R1 = ER[i], R7 = ER[g.nc + i], R4 = ER[2 * g.nc + i];
    if (!(R1 | R7 | R4)) continue;
    const int y = y0 + i / WW, w = i % WW, x = (w << 5) + lane;
    const int pos = EO[i];
    const bool cross = y + 1 == y1;
    const int hv = head_of(y, w, lane);
    auto emit = [&](int idx, int c) {
      if (cross) EX[idx] = make_int2(hv, c);
      else if (idx < edge_cap) EL[idx] = make_int2(hv, c);
      else unite_s(lab, hv - p00, c - p00);
    };
    if ((R1 >> lane) & 1u) emit(pos + __popc(R1 & lt), head_of(y + 1, (x - 1) >> 5, (x - 1) & 31));
    if ((R7 >> lane) & 1u) emit(pos + __popc(R1) + __popc(R7 & lt), head_of(y + 1, w, lane));
    if ((R4 >> lane) & 1u) emit(pos + __popc(R1) + __popc(R7) + __popc(R4 & lt), head_of(y + 1, (x + 1) >> 5, (x + 1) & 31));
  }
  __syncthreads();
  tl_end(18);

  // ---- 3. unions, in two levels so that the long pointer chains stay in LOCAL shared memory.
  // 3a: edges inside the strip, on strip-relative indices (plain shared-memory loads / atomicMin)
  {
    const int ne = min(s_nedge, edge_cap);
    for (int i = tid; i < ne; i += kImgThreads) unite_s(lab, EL[i].x - p00, EL[i].y - p00);
  }
  __syncthreads();
  // 3b: every head of the strip -> its strip-local root (read-only walks), absolute indices again
  for (int i = warp; i < nrows * WW; i += kImgWarps) {
    const int y = y0 + i / WW, w = i % WW;
    const unsigned P = pl(8, y, w);
    const int q = (y - y0) * W + (w << 5) + lane;          // strip-relative pixel
    int r = -1;
    if ((((P & ~CL[i]) >> lane) & 1u)) {
      r = q;
      for (int t = reinterpret_cast<volatile int*>(lab)[r]; t != r; t = reinterpret_cast<volatile int*>(lab)[r]) r = t;
    }
    __syncwarp();
    if (r >= 0) lab[q] = r;                                // (a root's word is only written by its own lane: r == q)
  }
  __syncthreads();
  for (int i = tid; i < nrows * W; i += kImgThreads) lab[i] += p00;
  cluster_barrier();  // every strip is flat and absolute: heads point at strip-local roots
  tl_end(19);
  // 3c: the edges into the next strip, between strip-local roots, through the cluster window (chains: one hop per strip)
  {
    const int nx = s_nxedge;
    for (int i = tid; i < nx; i += kImgThreads) L.unite(lab[EX[i].x - p00], L.ld(EX[i].y));
  }
  cluster_barrier();  // all unions of the image are done
  // 3d: heads -> the image-wide root (their strip-local root, then at most one hop per strip)
  for (int i = warp; i < nrows * WW; i += kImgWarps) {
    const int y = y0 + i / WW, w = i % WW;
    const unsigned hd = pl(8, y, w) & ~CL[i];
    if ((hd >> lane) & 1u) {
      const int h = y * W + (w << 5) + lane;
      int r = lab[h - p00];                                // strip-local root (or already further: pointers only move rootwards)
      for (int t = L.ld(r); t != r; t = L.ld(r)) r = t;
      lab[h - p00] = r;
    }
  }
  cluster_barrier();  // nobody walks pointers any more
  tl_end(20);
  // roots start counting: -(count) - 1
  for (int i = warp; i < nrows * WW; i += kImgWarps) {
    const int y = y0 + i / WW, w = i % WW;
    const unsigned hd = pl(8, y, w) & ~CL[i];
    const int h = y * W + (w << 5) + lane;
    if (((hd >> lane) & 1u) && lab[h - p00] == h) lab[h - p00] = -1;
  }
  cluster_barrier();
  // one atomic per run segment (a run that crosses a word boundary counts once per word)
  for (int i = warp; i < nrows * WW; i += kImgWarps) {
    const int y = y0 + i / WW, w = i % WW;
    const unsigned P = pl(8, y, w);
    if (!P) continue;
    const unsigned cl = CL[i];
    const bool entering = lane == 0 && (cl & 1u);  // its head lives further left in this row, hence in this strip
    if ((((P & ~cl) >> lane) & 1u) || entering) {
      const unsigned rest = lane < 31 ? (~cl & (0xffffffffu << (lane + 1))) : 0u;
      const int h = entering ? CH[i] : y * W + (w << 5) + lane, v = lab[h - p00];
      L.red_add(v < 0 ? h : v, -((rest ? __ffs(rest) - 1 : 32) - lane));
    }
  }
  cluster_barrier();
  tl_end(21);

  // ---- 4. label map, box slots, run records (one list per strip): one warp per word, lane = pixel
  unsigned long long* myrecs = recs + base + (size_t)rank * RS * W;
  for (int i = warp; i < nrows * WW; i += kImgWarps) {
    const int y = y0 + i / WW, w = i % WW, x = (w << 5) + lane, px = y * W + x;
    const unsigned P = pl(8, y, w);
    if (!P) {
      if (x < W) labels[base + px] = -1;
      continue;
    }
    const unsigned cl = CL[i];
    const bool member = (P >> lane) & 1u;
    const unsigned z = ~cl & (0xffffffffu >> (31 - lane));
    const int h = z ? y * W + (w << 5) + 31 - __clz(z) : CH[i];
    int root = -1, cnt = 0, v = 0;
    if (member) {
      v = lab[h - p00];
      root = v < 0 ? h : v;
      cnt = -(v < 0 ? v : L.ld(root)) - 1;
    }
    const bool kept = member && cnt > min_size;  // test_pixellink_fast.py:174 `len(index_list) > 10`
    if (x < W) labels[base + px] = kept ? root : -1;
    if (kept && v < 0 && h == px) {
      const int slot = atomicAdd(&n_boxes[b], 1);
      if (slot < K) comp_root[(size_t)b * K + slot] = px, comp_size[(size_t)b * K + slot] = cnt;
    }
    const bool seg = kept && (lane == 0 || !((cl >> lane) & 1u));  // a head, or the run entering at bit 0
    const unsigned sm = __ballot_sync(0xffffffffu, seg);
    if (sm == 0u) continue;
    int pos = 0;
    if (lane == 0) pos = atomicAdd(&s_nrec, __popc(sm));
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (seg) {
      const unsigned rest = lane < 31 ? (~cl & (0xffffffffu << (lane + 1))) : 0u;
      const int len = (rest ? __ffs(rest) - 1 : 32) - lane;
      myrecs[pos + __popc(sm & ((1u << lane) - 1u))] = make_rec(root, y, x, x + len - 1);
    }
  }
  pdl_release();
  cluster_barrier();  // no CTA retires while a peer may still read its words
  if (tid == 0) nrec[b * kImgCluster + rank] = s_nrec;
  tl_end(5);
}

// ------------------------------------------------------------------ D5: one CTA per component
// Persistent grid: the components of the whole batch form one work list (image-major); CTA c takes
// items c, c + gridDim.x, ...  (a grid of B x K CTAs would be ~95% empty CTAs that still have to be
// scheduled ahead of the real ones).
__global__ void __launch_bounds__(256)
decode_rects_kernel(const int* __restrict__ n_boxes, const int* __restrict__ comp_root,
                    const int* __restrict__ comp_size, const int* __restrict__ nrec,
                    const unsigned long long* __restrict__ recs,
                    int B, int H, int W, int K, double sx, double sy, int npad, int nsub, int sub_stride,
                    int32_t* __restrict__ boxes, float* __restrict__ rects, int32_t* __restrict__ comp) {
  pdl_wait_and_release();
  tl_start(10);
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_n, s_rank, s_b, s_slot;
  RectSmem S = rect_carve(smem, npad);
  for (int item = blockIdx.x;; item += gridDim.x) {
    // locate the item in the image-major list: warp 0 loads 32 counts at a time (independent loads,
    // one L2 round trip) and scans them with shuffles — a one-thread walk would be B dependent round trips
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      int acc = 0, bb = -1, sl = 0;
      for (int i0 = 0; i0 < B && bb < 0; i0 += 32) {
        const int nbi = (i0 + lane < B) ? min(n_boxes[i0 + lane], K) : 0;
        int inc = nbi;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const int lo = acc + inc - nbi, hi = acc + inc;          // this image covers items [lo, hi)
        const unsigned hit = __ballot_sync(0xffffffffu, item >= lo && item < hi);
        if (hit) {
          const int src = __ffs(hit) - 1;
          bb = i0 + src;
          sl = item - __shfl_sync(0xffffffffu, lo, src);
        }
        acc = __shfl_sync(0xffffffffu, hi, 31);
      }
      if (lane == 0) s_b = bb, s_slot = sl, s_n = 0, s_rank = 0;
    }
    __syncthreads();
    const int b = s_b, slot = s_slot;
    if (b < 0) { tl_end(10); return; }
    const int nb = min(n_boxes[b], K);
    const int root = comp_root[(size_t)b * K + slot];
    // row extremes of this component from the image's run records (the table lives where the explicit-list
    // path keeps its sort keys, unused here)
    int* s_rmin = reinterpret_cast<int*>(S.keys);
    int* s_rmax = s_rmin + H;
    const int y0 = root / W;  // the component's first row: its minimum pixel index lives there
    for (int y = y0 + threadIdx.x; y < H; y += blockDim.x) s_rmin[y] = 0x7fffffff, s_rmax[y] = -1;
    __syncthreads();
    for (int sub = 0; sub < nsub; ++sub) {  // tiled form: one record list per image; resident form: one per strip
      const int nr = nrec[b * nsub + sub];
      const unsigned long long* rb = recs + (size_t)b * H * W + (size_t)sub * sub_stride;
#pragma unroll 4
      for (int i = threadIdx.x; i < nr; i += blockDim.x) {
        const unsigned long long rc = rb[i];
        if ((int)(rc >> 32) == root) {
          const int y = (int)(rc >> 22) & 0x3ff;
          atomicMin(&s_rmin[y], (int)(rc >> 11) & 0x7ff);
          atomicMax(&s_rmax[y], (int)rc & 0x7ff);
        }
      }
    }
    // output position = rank of this component's label among the image's kept components
    {
      int r = 0;
      for (int j = threadIdx.x; j < nb; j += blockDim.x) r += comp_root[(size_t)b * K + j] < root;
      r = __reduce_add_sync(0xffffffffu, r);
      if ((threadIdx.x & 31) == 0 && r) atomicAdd(&s_rank, r);
    }
    // candidates: (minx, y) and (maxx, y) of every occupied row; input index = row-major order
    // (test_pixellink_fast.py:194-197: x*scale_x, y*scale_y assigned into an int64 array -> trunc).
    // Scales >= 1 keep the candidates distinct, so the sort-free hull applies.
    __syncthreads();
    for (int y = y0 + threadIdx.x; y < H; y += blockDim.x) {
      const int mn = s_rmin[y], mx = s_rmax[y];
      if (mx >= 0) {
        const int py = (int)((double)y * sy);
        const int pos = atomicAdd(&s_n, mn == mx ? 1 : 2);
        S.X[pos] = (int)((double)mn * sx), S.Y[pos] = py, S.I[pos] = 2 * y;
        if (mn != mx) S.X[pos + 1] = (int)((double)mx * sx), S.Y[pos + 1] = py, S.I[pos + 1] = 2 * y + 1;
      }
    }
    __syncthreads();
    const int total = s_n;
    const int rank = s_rank;
    int box[8];
    float rect[5];
    min_area_box_distinct(S, total, box, rect);
    if (threadIdx.x == 0) {
      int32_t* ob = boxes + ((size_t)b * K + rank) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) ob[i] = box[i];
      if (rects) {
        float* orc = rects + ((size_t)b * K + rank) * 5;
#pragma unroll
        for (int i = 0; i < 5; ++i) orc[i] = rect[i];
      }
      if (comp) {
        comp[((size_t)b * K + rank) * 2] = root;
        comp[((size_t)b * K + rank) * 2 + 1] = comp_size[(size_t)b * K + slot];
      }
    }
    __syncthreads();  // shared state is reused by the next item
  }
  tl_end(10);
}

// explicit point lists: one CTA per list
__global__ void __launch_bounds__(256)
min_area_boxes_kernel(const int32_t* __restrict__ pts, const int32_t* __restrict__ offsets, int npad,
                      int32_t* __restrict__ boxes, float* __restrict__ rects) {
  extern __shared__ __align__(16) unsigned char smem[];
  RectSmem S = rect_carve(smem, npad);
  const int s = blockIdx.x;
  const int o0 = offsets[s], total = offsets[s + 1] - o0;
  if (total > npad) {  // more points than the shared-memory carve holds: sentinel box, no out-of-bounds write
    if (threadIdx.x == 0) {
      for (int i = 0; i < 8; ++i) boxes[(size_t)s * 8 + i] = INT32_MIN;
      if (rects)
        for (int i = 0; i < 5; ++i) rects[(size_t)s * 5 + i] = __int_as_float(0x7fc00000);
    }
    return;
  }
  int nsort = 32;
  while (nsort < total) nsort <<= 1;
  for (int i = threadIdx.x; i < nsort; i += blockDim.x)
    S.keys[i] = i < total ? make_key(pts[2 * (o0 + i)], pts[2 * (o0 + i) + 1], i) : ~0ull;
  __syncthreads();
  bitonic_sort(S.keys, nsort);
  int box[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float rect[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (total > 0) min_area_box_sorted(S, total, npad, box, rect);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) boxes[(size_t)s * 8 + i] = box[i];
    if (rects)
      for (int i = 0; i < 5; ++i) rects[(size_t)s * 5 + i] = rect[i];
  }
}

// ------------------------------------------------------------------ D1 of the survey: pixel_detect
// tool/pixellink_fn.py:120-154: mask = score > thr_p, cleared wherever any link_d[...,1] < thr_l.
__global__ void pixel_detect_kernel(const float* __restrict__ score, const float* __restrict__ link, int N,
                                    float thr_p, float thr_l, uint8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  bool m = score[i] > thr_p;
#pragma unroll
  for (int d = 0; d < 8; ++d) m = m && !(link[((size_t)d * N + i) * 2 + 1] < thr_l);
  out[i] = m ? 1 : 0;
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

static int decode_common(const uint16_t* flags_in, const float* pix_logits, const float* link_logits, int B, int H,
                         int W, const plh_decode_params* p, int32_t* labels, int32_t* boxes, int32_t* n_boxes,
                         float* rects, int32_t* comp, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  if (!p || !labels || !boxes || !n_boxes) return PLH_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || H > 1024 || W > 2048 || (long long)B * H * W > (1ll << 31) - 1) return PLH_E_SHAPE;
  if (p->max_boxes <= 0 || p->min_size < 0 || !(p->scale_x >= 1.0) || !(p->scale_y >= 1.0) ||
      W * p->scale_x >= 32768.0 || H * p->scale_y >= 32768.0)
    return PLH_E_PARAM;
  if (!aligned16(workspace) || !aligned16(labels)) return PLH_E_ALIGN;
  const int K = p->max_boxes;
  const DecodeWsLayout l = decode_ws_layout(B, H, W, K);
  if (!workspace || workspace_bytes < l.total) return PLH_E_WORKSPACE;
  char* ws = (char*)workspace;
  uint16_t* flags = (uint16_t*)(ws + l.flags);
  int* parent = (int*)(ws + l.parent);
  int* size = (int*)(ws + l.size);
  int* comp_root = (int*)(ws + l.comp_root);
  int* comp_size = (int*)(ws + l.comp_size);
  int* nrec = (int*)(ws + l.nrec);
  unsigned long long* recs = (unsigned long long*)(ws + l.recs);
  const long long total_px = (long long)B * H * W;
  const int N = H * W;
  int rc;
  // pixel-parallel passes: a thread's grid-stride iterations are dependent L2 round trips in sequence, so large
  // batches get more CTAs rather than more iterations (at most ~2 pixels per thread)
  const int grid_px = (int)std::min<long long>((total_px + 255) / 256, std::max<long long>(kNumSMs * 16, (total_px + 511) / 512));
  const bool skip_rects = (p->reserved[0] & 1) != 0;  // components + label map only
  const bool rects_only = (p->reserved[0] & 2) != 0;  // boxes from a workspace prepared by a skip_rects call
  const bool tile_only = (p->reserved[0] & 4) != 0;   // stop after the threshold + tile-labelling kernel
  const bool after_tile = (p->reserved[0] & 8) != 0;  // resume on a workspace a tile_only call prepared
  // Two forms, same results.  Tiled (default): five launches, the forest in global memory (L2).  Resident (bit 5,
  // opt-in): three launches — threshold bits as bit planes, components with the map in the shared memory of an
  // 8-CTA cluster (union-find words in distributed shared memory), boxes.  Measured on B200 at 32 x 128x128
  // the resident form is NOT faster (61 us vs 50 us alone: its ~10 phases are each bound by the latency of one
  // warp's dependent shared-memory chain and by the busiest strip of a cluster), so it is
  // kept as a cross-check of the tiled form in the tests, not as the default.
  const int edge_cap = image_edge_cap(H, W);
  const bool fits = edge_cap > 0;
  const size_t img_smem = image_smem_base_bytes(H, W) + (size_t)edge_cap * 8;
  const bool planes_ok = !flags_in && (W % 32) == 0;   // the threshold kernel writes bit planes (into the forest's region)
  unsigned* planes = reinterpret_cast<unsigned*>(parent);
  const unsigned long long rsw = (unsigned long long)image_geom(H, W).RS * W;
  const unsigned long long magic = ((1ull << 40) + rsw - 1) / rsw;  // px / rsw == (px * magic) >> 40 for px < 2^31
  if ((p->reserved[0] & 32) != 0 && !fits) return PLH_E_SHAPE;  // the resident form was demanded
  const bool legacy = (p->reserved[0] & 32) == 0 || tile_only || after_tile;
  if (rects_only) goto rects;
  if (!legacy) {
    if (flags_in) {
      flags = const_cast<uint16_t*>(flags_in);
    } else {
      const int grid = (int)std::min<long long>((total_px / 32 + 7) / 8 + 1, kNumSMs * 8);
      if (planes_ok)
        rc = launch_plain(decode_flags_kernel<true>, grid, 256, 0, s, pix_logits, link_logits, total_px,
                          prob_to_logit_threshold(p->pixel_thresh), prob_to_logit_threshold(p->link_thresh),
                          (uint16_t*)nullptr, (int*)nullptr, 0, planes, W / 32);
      else
        rc = launch_plain(decode_flags_kernel<false>, grid, 256, 0, s, pix_logits, link_logits, total_px,
                          prob_to_logit_threshold(p->pixel_thresh), prob_to_logit_threshold(p->link_thresh), flags,
                          (int*)nullptr, 0, (unsigned*)nullptr, 0);
      if (rc) return rc;
    }
    static SmemOptIn optin_p, optin_f;
    static const bool img_plain = getenv("PLH_IMG_PLAIN") != nullptr;  // A/B: no programmatic early launch
    if (planes_ok && img_plain) {
      if ((rc = ensure_dynamic_smem(optin_p, decode_image_kernel<true>, img_smem))) return rc;
      rc = launch_plain(decode_image_kernel<true>, B * kImgCluster, kImgThreads, img_smem, s, (const uint16_t*)nullptr,
                  (const unsigned*)planes, H, W, p->min_size, K, edge_cap, magic, labels, comp_root, comp_size, n_boxes,
                  nrec, recs);
    } else if (planes_ok) {
      if ((rc = ensure_dynamic_smem(optin_p, decode_image_kernel<true>, img_smem))) return rc;
      rc = launch(decode_image_kernel<true>, B * kImgCluster, kImgThreads, img_smem, s, (const uint16_t*)nullptr,
                  (const unsigned*)planes, H, W, p->min_size, K, edge_cap, magic, labels, comp_root, comp_size, n_boxes,
                  nrec, recs);
    } else {
      if ((rc = ensure_dynamic_smem(optin_f, decode_image_kernel<false>, img_smem))) return rc;
      rc = launch(decode_image_kernel<false>, B * kImgCluster, kImgThreads, img_smem, s, (const uint16_t*)flags,
                  (const unsigned*)nullptr, H, W, p->min_size, K, edge_cap, magic, labels, comp_root, comp_size, n_boxes,
                  nrec, recs);
    }
    if (rc) return rc;
    if (skip_rects) return PLH_OK;
    goto rects;
  }
  if (after_tile) goto merge;
  {
    const dim3 tiles((W + kTW - 1) / kTW, (H + kTH - 1) / kTH, B);
    if (flags_in) {
      flags = const_cast<uint16_t*>(flags_in);
      rc = launch(decode_tile_cc_kernel<false>, tiles, kTW * kTH, 0, s, (const float*)nullptr, (const float*)nullptr,
                  0.f, 0.f, flags, H, W, parent, size, n_boxes, nrec);
    } else {  // first kernel of the chain: thresholds + tile labelling in one pass over the logits
      rc = launch_plain(decode_tile_cc_kernel<true>, tiles, kTW * kTH, 0, s, pix_logits, link_logits,
                        prob_to_logit_threshold(p->pixel_thresh), prob_to_logit_threshold(p->link_thresh), flags, H,
                        W, parent, size, n_boxes, nrec);
    }
    if (rc) return rc;
  }
  if (tile_only) return PLH_OK;
merge:
  rc = launch(decode_cross_kernel, grid_px, 256, 0, s, flags, H, W, (int)total_px, parent);
  if (rc) return rc;
  rc = launch(decode_flatten_kernel, grid_px, 256, 0, s, flags, H, W, total_px, parent, size);
  if (rc) return rc;
  rc = launch(decode_labels_kernel, grid_px, 256, 0, s, parent, size, H, W, total_px, p->min_size, K, labels, comp_root,
              comp_size, n_boxes, nrec, recs);
  if (rc) return rc;
  if (skip_rects) return PLH_OK;
rects:
  {
    const int npad = next_pow2(std::max(2 * H, 32));
    const size_t smem = rect_smem_bytes(npad);
    static SmemOptIn optin;
    if ((rc = ensure_dynamic_smem(optin, decode_rects_kernel, smem))) return rc;
    const int rect_grid = (int)std::min<long long>((long long)B * K, kNumSMs * 8);
    rc = launch(decode_rects_kernel, rect_grid, 256, smem, s, n_boxes, comp_root, comp_size, nrec, recs, B, H, W, K,
                p->scale_x, p->scale_y, npad, legacy ? 1 : kImgCluster, legacy ? 0 : image_geom(H, W).RS * W, boxes, rects,
                comp);
    if (rc) return rc;
  }
  return PLH_OK;
}

// Component labels of ready-made flag maps, nothing else (the contour path labels the foreground and the
// background of a binary mask with it): tiled form, every component kept, no boxes.
int decode_labels_only(const uint16_t* flags, int B, int H, int W, int32_t* labels, int32_t* n_boxes, int32_t* scratch_boxes,
                       void* workspace, size_t workspace_bytes, cudaStream_t s) {
  plh_decode_params p = {};
  p.pixel_thresh = 0.5f, p.link_thresh = 0.5f, p.min_size = 0, p.max_boxes = 1, p.scale_x = 1.0, p.scale_y = 1.0;
  p.reserved[0] = 1;  // components + label map only
  return decode_common(flags, nullptr, nullptr, B, H, W, &p, labels, scratch_boxes, n_boxes, nullptr, nullptr, workspace,
                       workspace_bytes, s);
}

#ifdef PLH_TIMELINE
int tl_set_decode(unsigned long long* p) { return tl_set_ptr(p); }
#endif
int debug_image_cluster_occupancy(int H, int W) {
  const int edge_cap = image_edge_cap(H, W);
  if (!edge_cap) return -1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(32 * kImgCluster);
  cfg.blockDim = dim3(kImgThreads);
  cfg.dynamicSmemBytes = image_smem_base_bytes(H, W) + (size_t)edge_cap * 8;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kImgCluster, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  cfg.attrs = at, cfg.numAttrs = 1;
  cudaFuncSetAttribute(decode_image_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
  int n = -2;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, decode_image_kernel<true>, &cfg);
  return e == cudaSuccess ? n : -1000 - (int)e;
}

}  // namespace plh

using namespace plh;

#ifdef PLH_TIMELINE
extern "C" __attribute__((visibility("default"))) int plh_debug_image_cluster_occupancy(int H, int W) {
  return debug_image_cluster_occupancy(H, W);
}
#endif

extern "C" int plh_decode(const float* pix_logits, const float* link_logits, int B, int H, int W,
                          const plh_decode_params* p, int32_t* labels, int32_t* boxes, int32_t* n_boxes, float* rects,
                          int32_t* comp, void* workspace, size_t workspace_bytes, void* stream) {
  if (!pix_logits || !link_logits) return PLH_E_NULL;
  if (!aligned16(pix_logits) || !aligned16(link_logits)) return PLH_E_ALIGN;
  return decode_common(nullptr, pix_logits, link_logits, B, H, W, p, labels, boxes, n_boxes, rects, comp, workspace,
                       workspace_bytes, (cudaStream_t)stream);
}

extern "C" int plh_decode_from_flags(const uint16_t* flags, int B, int H, int W, const plh_decode_params* p,
                                     int32_t* labels, int32_t* boxes, int32_t* n_boxes, float* rects, int32_t* comp,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  if (!flags) return PLH_E_NULL;
  return decode_common(flags, nullptr, nullptr, B, H, W, p, labels, boxes, n_boxes, rects, comp, workspace,
                       workspace_bytes, (cudaStream_t)stream);
}

extern "C" int plh_decode_flags(const float* pix_logits, const float* link_logits, int B, int H, int W,
                                const plh_decode_params* p, uint16_t* flags, void* stream) {
  if (!pix_logits || !link_logits || !p || !flags) return PLH_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || (long long)B * H * W > (1ll << 31) - 1) return PLH_E_SHAPE;
  if (!aligned16(pix_logits) || !aligned16(link_logits) || !aligned16(flags)) return PLH_E_ALIGN;
  const long long total_px = (long long)B * H * W;
  const int grid = (int)std::min<long long>((total_px / 32 + 7) / 8 + 1, kNumSMs * 8);
  return launch_plain(decode_flags_kernel<false>, grid, 256, 0, (cudaStream_t)stream, pix_logits, link_logits, total_px,
                prob_to_logit_threshold(p->pixel_thresh), prob_to_logit_threshold(p->link_thresh), flags,
                (int*)nullptr, 0, (unsigned*)nullptr, 0);
}

extern "C" int plh_min_area_boxes(const int32_t* pts, const int32_t* offsets, int n_sets, int32_t* boxes, float* rects,
                                  void* stream) {
  if (!pts || !offsets || !boxes) return PLH_E_NULL;
  if (n_sets <= 0) return PLH_E_SHAPE;
  const int npad = 2048;  // max points per set
  const size_t smem = rect_smem_bytes(npad);
  static SmemOptIn optin;
  if (int rc = ensure_dynamic_smem(optin, min_area_boxes_kernel, smem)) return rc;
  min_area_boxes_kernel<<<n_sets, 256, smem, (cudaStream_t)stream>>>(pts, offsets, npad, boxes, rects);
  return launch_status();
}

extern "C" int plh_pixel_detect(const float* score, const float* link, int H, int W, float score_map_thresh,
                                float link_thresh, uint8_t* out, void* stream) {
  if (!score || !link || !out) return PLH_E_NULL;
  if (H <= 0 || W <= 0) return PLH_E_SHAPE;
  const int N = H * W;
  pixel_detect_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(score, link, N, score_map_thresh,
                                                                         link_thresh, out);
  return launch_status();
}
