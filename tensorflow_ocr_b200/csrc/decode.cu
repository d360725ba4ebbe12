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
  l.nrec = off; off = align_up(off + (size_t)B * 4, 256);
  l.recs = off; off = align_up(off + px * 8, 256);   // one run-boundary record per pixel at most
  l.total = off;
  return l;
}
size_t decode_workspace_bytes(int B, int H, int W, int K) { return decode_ws_layout(B, H, W, K).total; }

// ------------------------------------------------------------------ D0: flags from logits (+ init)
// Work unit = 32 consecutive pixels per warp: four link iterations (lane = pixel-in-iteration x quarter,
// one 128-bit load of two directions' logits each) and one pixel phase (lane = pixel: its 2 pixel logits,
// the 16-bit flag word store).  ~3 instructions per pixel: the kernel is bound by the 72 B/px it reads.
__global__ void __launch_bounds__(256)
decode_flags_kernel(const float* __restrict__ pix_logits, const float* __restrict__ link_logits, long long total_px,
                    float tp_logit, float tl_logit, uint16_t* __restrict__ flags, int* __restrict__ n_boxes, int B) {
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
    if (pp < total) flags[pp] = (uint16_t)(mine | ((P.y - P.x) > tp_logit ? kFlagP : 0));
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
  if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) n_boxes[b] = 0, nrec[b] = 0;
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
  const unsigned f = sf[ly + 1][lx + 1];
  const bool P = inimg && (f & kFlagP);
  slab[tid] = P ? tid : -1;
  __syncthreads();
  if (P) {
    const bool vin = gx >= 1 && gx <= W - 2 && gy >= 1 && gy <= H - 2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int d = c_fwd[k];
      const int uy = ly + c_dy[d], ux = lx + c_dx[d];
      if (uy < 0 || uy >= kTH || ux < 0 || ux >= kTW) continue;  // cross-tile: decode_cross_kernel
      const unsigned fu = sf[uy + 1][ux + 1];
      if (!(fu & kFlagP)) continue;
      const int ugx = gx + c_dx[d], ugy = gy + c_dy[d];
      const bool uin = ugx >= 1 && ugx <= W - 2 && ugy >= 1 && ugy <= H - 2;
      if ((vin && (f & (1u << d))) || (uin && (fu & (1u << c_opp[k])))) unite_s(slab, tid, uy * kTW + ux);
    }
  }
  __syncthreads();
  if (inimg) {
    const size_t g = base + (size_t)gy * W + gx;
    int root = -1;
    if (P) {
      const int r = find_root_s(slab, tid);
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
        rec = ((unsigned long long)(unsigned)rl << 32) | ((unsigned long long)y << 16) | (unsigned long long)x;
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

// ------------------------------------------------------------------ D5: one CTA per component
// Persistent grid: the components of the whole batch form one work list (image-major); CTA c takes
// items c, c + gridDim.x, ...  (a grid of B x K CTAs would be ~95% empty CTAs that still have to be
// scheduled ahead of the real ones).
__global__ void __launch_bounds__(256)
decode_rects_kernel(const int* __restrict__ n_boxes, const int* __restrict__ comp_root,
                    const int* __restrict__ comp_size, const int* __restrict__ nrec,
                    const unsigned long long* __restrict__ recs,
                    int B, int H, int W, int K, double sx, double sy, int npad, int32_t* __restrict__ boxes,
                    float* __restrict__ rects, int32_t* __restrict__ comp) {
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
    {
      const int nr = nrec[b];
      const unsigned long long* rb = recs + (size_t)b * H * W;
#pragma unroll 4
      for (int i = threadIdx.x; i < nr; i += blockDim.x) {
        const unsigned long long rc = rb[i];
        if ((int)(rc >> 32) == root) {
          const int y = (int)(rc >> 16) & 0xffff, x = (int)rc & 0xffff;
          atomicMin(&s_rmin[y], x);
          atomicMax(&s_rmax[y], x);
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
  if (B <= 0 || H <= 0 || W <= 0 || H > 1024 || (long long)B * H * W > (1ll << 31) - 1) return PLH_E_SHAPE;
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
  const int grid_px = (int)std::min<long long>((total_px + 255) / 256, kNumSMs * 16);
  const bool skip_rects = (p->reserved[0] & 1) != 0;  // components + label map only
  const bool rects_only = (p->reserved[0] & 2) != 0;  // boxes from a workspace prepared by a skip_rects call
  const bool tile_only = (p->reserved[0] & 4) != 0;   // stop after the threshold + tile-labelling kernel
  const bool after_tile = (p->reserved[0] & 8) != 0;  // resume on a workspace a tile_only call prepared
  if (rects_only) goto rects;
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
                                                      p->scale_x, p->scale_y, npad, boxes, rects, comp);
    if (rc) return rc;
  }
  return PLH_OK;
}

#ifdef PLH_TIMELINE
int tl_set_decode(unsigned long long* p) { return tl_set_ptr(p); }
#endif

}  // namespace plh

using namespace plh;

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
  return launch_plain(decode_flags_kernel, grid, 256, 0, (cudaStream_t)stream, pix_logits, link_logits, total_px,
                prob_to_logit_threshold(p->pixel_thresh), prob_to_logit_threshold(p->link_thresh), flags,
                (int*)nullptr, 0);
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
