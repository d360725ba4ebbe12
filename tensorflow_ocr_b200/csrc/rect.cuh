// rect.cuh — cv2.minAreaRect -> cv2.boxPoints -> np.int0 for one integer point set,
// restated step by step for one CTA (model: oracle/minarearect.py, which is pinned
// bit-exactly against OpenCV 4.13.0; reference call sites test_pixellink_fast.py:199-200,
// test.py:190-191).  OpenCV's convexHull is Sklansky's scan over points sorted by
// (x, y, input index); minAreaRect runs fp32 rotating calipers over that hull.
//
// Every fp32 operation uses the _rn intrinsics so that nvcc never contracts a
// multiply-add: the rounding sequence is OpenCV's.
//
// Work split inside the CTA: bitonic sort (all threads) -> coordinates decoded to int
// arrays (all threads) -> the four Sklansky scans on four warps concurrently, each
// carrying the current/previous point in registers -> hull assembly (thread 0) -> edge
// vectors / inverse lengths / unit leads (all threads) -> calipers loop (thread 0).
#pragma once
#include "common.cuh"

namespace plh {

struct RectSmem {
  unsigned long long* keys;  // [npad] sort keys (x, y, input index)
  int* X;                    // [npad] sorted coordinates / input indices
  int* Y;
  int* I;
  int* stack;                // 4 regions of (npad + 4): one per Sklansky scan
  int* hullbuf;              // [npad] hull as sorted positions
  float* hx;                 // [npad] hull points, edge vectors, inverse lengths, unit leads
  float* hy;
  float* vx;
  float* vy;
  float* inv;
  float* lx;
  float* ly;
};

__host__ __device__ inline size_t rect_smem_bytes(int npad) {
  return (size_t)npad * 8 + (size_t)npad * 4 * 3 + (size_t)(npad + 4) * 4 * 4 + (size_t)npad * 4 +
         (size_t)npad * 4 * 7 + 64;
}

__device__ inline RectSmem rect_carve(unsigned char* base, int npad) {
  RectSmem s;
  s.keys = reinterpret_cast<unsigned long long*>(base);
  base += (size_t)npad * 8;
  s.X = reinterpret_cast<int*>(base);
  s.Y = s.X + npad;
  s.I = s.Y + npad;
  base += (size_t)npad * 12;
  s.stack = reinterpret_cast<int*>(base);
  base += (size_t)(npad + 4) * 16;
  s.hullbuf = reinterpret_cast<int*>(base);
  base += (size_t)npad * 4;
  s.hx = reinterpret_cast<float*>(base);
  s.hy = s.hx + npad;
  s.vx = s.hy + npad;
  s.vy = s.vx + npad;
  s.inv = s.vy + npad;
  s.lx = s.inv + npad;
  s.ly = s.lx + npad;
  return s;
}

// key orders by (x, y, index); x,y biased by 32768 into 24/20-bit fields, index < 2^20
__device__ __forceinline__ unsigned long long make_key(int x, int y, int idx) {
  return ((unsigned long long)(unsigned)(x + 32768) << 40) | ((unsigned long long)(unsigned)(y + 32768) << 20) |
         (unsigned long long)(unsigned)idx;
}
__device__ __forceinline__ int key_x(unsigned long long k) { return (int)(k >> 40) - 32768; }
__device__ __forceinline__ int key_y(unsigned long long k) { return (int)((k >> 20) & 0xFFFFFu) - 32768; }
__device__ __forceinline__ int key_idx(unsigned long long k) { return (int)(k & 0xFFFFFu); }

// In-place bitonic sort of keys[0..n) by the whole CTA (n a power of two; pad = ~0).
__device__ inline void bitonic_sort(unsigned long long* keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) keys[i] = b, keys[ixj] = a;
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int sgn(long long v) { return (v > 0) - (v < 0); }

// OpenCV convhull.cpp Sklansky_ over the sorted points; returns the stack size.
// The coordinates of pprev / pcur are carried in registers: one shared-memory round trip
// per step instead of five.
__device__ inline int sklansky(const int* __restrict__ X, const int* __restrict__ Y, int start, int end, int* stack,
                               int nsign, int sign2) {
  const int incr = end > start ? 1 : -1;
  int pprev = start, pcur = pprev + incr, pnext = pcur + incr;
  int stacksize = 3;
  if (start == end || (X[start] == X[end] && Y[start] == Y[end])) {
    stack[0] = start;
    return 1;
  }
  stack[0] = pprev, stack[1] = pcur, stack[2] = pnext;
  end += incr;
  int xprev = X[pprev], yprev = Y[pprev], xcur = X[pcur], ycur = Y[pcur];
  while (pnext != end) {
    const int xnext = X[pnext], ynext = Y[pnext];
    const int by = ynext - ycur;
    if (sgn(by) != nsign) {
      const int ax = xcur - xprev;
      const int bx = xnext - xcur;
      const int ay = ycur - yprev;
      const long long convexity = (long long)ay * bx - (long long)ax * by;
      if (sgn(convexity) == sign2 && (ax != 0 || ay != 0)) {
        pprev = pcur, xprev = xcur, yprev = ycur;
        pcur = pnext, xcur = xnext, ycur = ynext;
        pnext += incr;
        stack[stacksize] = pnext;
        stacksize++;
      } else {
        if (pprev == start) {
          pcur = pnext, xcur = xnext, ycur = ynext;
          stack[1] = pcur;
          pnext += incr;
          stack[2] = pnext;
        } else {
          stack[stacksize - 2] = pnext;
          pcur = pprev, xcur = xprev, ycur = yprev;
          pprev = stack[stacksize - 4];
          xprev = X[pprev], yprev = Y[pprev];
          stacksize--;
        }
      }
    } else {
      pnext += incr;
      stack[stacksize - 1] = pnext;
    }
  }
  return --stacksize;
}

// Cyclic shift of the hull towards a monotone sequence of INPUT indices (convhull.cpp, the block after
// the four scans).  Thread 0 only.  hullbuf holds positions into X/Y/I; tmp is scratch of >= nout ints.
__device__ inline void hull_index_shift(int* hullbuf, int nout, const int* I, int* tmp) {
  if (nout < 3) return;
  int min_idx = 0, max_idx = 0, lt = 0;
  int prev = I[hullbuf[0]], vmin = prev, vmax = prev;
  for (int i = 1; i < nout; ++i) {
    const int idx = I[hullbuf[i]];
    lt += prev < idx;
    if (lt > 1 && lt <= i - 2) break;
    if (idx < vmin) vmin = idx, min_idx = i;
    if (idx > vmax) vmax = idx, max_idx = i;
    prev = idx;
  }
  const int mmdist = abs(max_idx - min_idx);
  if ((mmdist == 1 || mmdist == nout - 1) && (lt <= 1 || lt >= nout - 2)) {
    const int ascending = (max_idx + 1) % nout == min_idx;
    const int i0 = ascending ? min_idx : max_idx;
    int j = i0;
    if (i0 > 0) {
      int i;
      int curr_idx = I[hullbuf[j]];
      for (i = 0; i < nout; ++i) {
        tmp[i] = hullbuf[j];
        const int next_j = j + 1 < nout ? j + 1 : 0;
        const int next_idx = I[hullbuf[next_j]];
        if (i < nout - 1 && (ascending != (curr_idx < next_idx))) break;
        j = next_j;
        curr_idx = next_idx;
      }
      if (i == nout)
        for (i = 0; i < nout; ++i) hullbuf[i] = tmp[i];
    }
  }
}

__device__ inline void hull_to_box(const RectSmem& S, int n, int* out_box, float* out_rect);

// Hull of DISTINCT points without sorting and without the sequential Sklansky scans (decode path:
// row extremes are distinct).  Same vertex sequence as cv::convexHull(clockwise=false) before the
// index shift (oracle/minarearect.py::convex_hull_giftwrap, pinned against the Sklansky restatement
// and cv2): start at the lexicographic maximum (x, then y); the successor of p is the point q with every
// other point on the clockwise side of p->q (x, y-down), the farthest among collinear candidates.
// Successors of all points are computed in parallel (O(n^2) integer cross products), then thread 0
// follows the pointers.  X/Y/I hold the points in any order; called by the whole CTA.
__device__ inline void min_area_box_distinct(const RectSmem& S, int total, int* out_box, float* out_rect) {
  __shared__ int s_start, s_nh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* X = S.X;
  const int* Y = S.Y;
  int* succ = S.stack;  // [total]
  if (warp == 0) {
    int bx = -0x7fffffff, by = -0x7fffffff, bi = 0;
    for (int i = lane; i < total; i += 32) {
      const int x = X[i], y = Y[i];
      if (x > bx || (x == bx && y > by)) bx = x, by = y, bi = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int ox = __shfl_xor_sync(0xffffffffu, bx, o), oy = __shfl_xor_sync(0xffffffffu, by, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ox > bx || (ox == bx && oy > by)) bx = ox, by = oy, bi = oi;
    }
    if (lane == 0) s_start = bi;
  }
  // coordinates are < 2^15 in magnitude (checked on the host), so every product below fits in int32.
  // Branch-free body: j == i gives b = 0 (cross 0, dot 0 -> not taken) and the initial candidate
  // compares equal to itself (not farther -> not taken), so neither needs a special case.
  for (int i = tid; i < total; i += blockDim.x) {
    const int xi = X[i], yi = Y[i];
    int best = (i == 0 && total > 1) ? 1 : 0;
    int ax = X[best] - xi, ay = Y[best] - yi;
#pragma unroll 4
    for (int j = 0; j < total; ++j) {
      const int bx = X[j] - xi, by = Y[j] - yi;
      const int c = ax * by - ay * bx;
      const bool far = (ax * bx + ay * by) > 0 && (bx * bx + by * by) > (ax * ax + ay * ay);
      const bool take = c < 0 || (c == 0 && far);
      best = take ? j : best, ax = take ? bx : ax, ay = take ? by : ay;
    }
    if (total == 1) best = -1;
    succ[i] = best;
  }
  __syncthreads();
  if (tid == 0) {
    int* hullbuf = S.hullbuf;
    const int start = s_start;
    int nh = 0, cur = start;
    hullbuf[nh++] = start;
    if (total > 1) {
      while (true) {
        const int nx = succ[cur];
        if (nx == start || nh > total) break;
        hullbuf[nh++] = nx;
        cur = nx;
      }
    }
    hull_index_shift(hullbuf, nh, S.I, S.stack + total);
    s_nh = nh;
  }
  __syncthreads();
  hull_to_box(S, s_nh, out_box, out_rect);
}

// Called by the WHOLE CTA (>= 128 threads) after the keys are sorted (explicit point lists, which may
// contain duplicates: OpenCV's own Sklansky scans are followed literally).  total = number of real
// points (> 0).  out_box: 8 ints (x0,y0..x3,y3), out_rect: 5 floats or nullptr — valid in thread 0 only.
__device__ inline void min_area_box_sorted(const RectSmem& S, int total, int npad, int* out_box, float* out_rect) {
  __shared__ int s_ind[2];     // miny_ind, maxy_ind
  __shared__ int s_count[4];   // tl, tr, bl, br stack sizes
  __shared__ int s_nout;
  int* hullbuf = S.hullbuf;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < total; i += blockDim.x) {
    const unsigned long long k = S.keys[i];
    S.X[i] = key_x(k), S.Y[i] = key_y(k), S.I[i] = key_idx(k);
  }
  __syncthreads();
  const int* X = S.X;
  const int* Y = S.Y;
  // ---- cv::convexHull(points, clockwise=false, returnPoints=true)
  // first sorted position with the minimum y / with the maximum y (strict comparisons in OpenCV's loop)
  if (warp == 0) {
    int miny = 0x7fffffff, mini = 0x7fffffff, maxy = -0x7fffffff, maxi = 0x7fffffff;
    for (int i = lane; i < total; i += 32) {
      const int y = Y[i];
      if (y < miny) miny = y, mini = i;
      if (y > maxy) maxy = y, maxi = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int oy = __shfl_xor_sync(0xffffffffu, miny, o), oi = __shfl_xor_sync(0xffffffffu, mini, o);
      if (oy < miny || (oy == miny && oi < mini)) miny = oy, mini = oi;
      const int py = __shfl_xor_sync(0xffffffffu, maxy, o), pi = __shfl_xor_sync(0xffffffffu, maxi, o);
      if (py > maxy || (py == maxy && pi < maxi)) maxy = py, maxi = pi;
    }
    if (lane == 0) s_ind[0] = mini, s_ind[1] = maxi;
  }
  __syncthreads();
  const int miny_ind = s_ind[0], maxy_ind = s_ind[1];
  const bool single = X[0] == X[total - 1] && Y[0] == Y[total - 1];
  const int SS = npad + 4;  // stack stride: one region per scan
  if (!single && lane == 0 && warp < 4) {
    int* st = S.stack + warp * SS;
    int c;
    if (warp == 0) c = sklansky(X, Y, 0, maxy_ind, st, -1, 1);               // tl
    else if (warp == 1) c = sklansky(X, Y, total - 1, maxy_ind, st, -1, -1);  // tr
    else if (warp == 2) c = sklansky(X, Y, 0, miny_ind, st, 1, -1);           // bl
    else c = sklansky(X, Y, total - 1, miny_ind, st, 1, 1);                   // br
    s_count[warp] = c;
  }
  __syncthreads();
  if (tid == 0) {
    int nout = 0;
    if (single) {
      hullbuf[nout++] = 0;
    } else {
      int* tl_stack = S.stack;
      int tl_count = s_count[0];
      int* tr_stack = S.stack + SS;
      int tr_count = s_count[1];
      {  // !clockwise: swap
        int* t = tl_stack; tl_stack = tr_stack; tr_stack = t;
        int c = tl_count; tl_count = tr_count; tr_count = c;
      }
      for (int i = 0; i < tl_count - 1; ++i) hullbuf[nout++] = tl_stack[i];
      for (int i = tr_count - 1; i > 0; --i) hullbuf[nout++] = tr_stack[i];
      const int stop_idx = tr_count > 2 ? tr_stack[1] : (tl_count > 2 ? tl_stack[tl_count - 2] : -1);
      int* bl_stack = S.stack + 2 * SS;
      int bl_count = s_count[2];
      int* br_stack = S.stack + 3 * SS;
      int br_count = s_count[3];
      if (stop_idx >= 0) {
        const int check_idx = bl_count > 2 ? bl_stack[1] : (bl_count + br_count > 2 ? br_stack[2 - bl_count] : -1);
        if (check_idx == stop_idx ||
            (check_idx >= 0 && X[check_idx] == X[stop_idx] && Y[check_idx] == Y[stop_idx])) {
          bl_count = min(bl_count, 2);
          br_count = min(br_count, 2);
        }
      }
      for (int i = 0; i < bl_count - 1; ++i) hullbuf[nout++] = bl_stack[i];
      for (int i = br_count - 1; i > 0; --i) hullbuf[nout++] = br_stack[i];
      hull_index_shift(hullbuf, nout, S.I, S.stack);  // the scan stacks are dead now
    }
    s_nout = nout;
  }
  __syncthreads();
  hull_to_box(S, s_nout, out_box, out_rect);
}

// Common tail, called by the whole CTA: hull (positions in S.hullbuf) -> cv::minAreaRect -> boxPoints -> int.
__device__ inline void hull_to_box(const RectSmem& S, int n, int* out_box, float* out_rect) {
  const int tid = threadIdx.x;
  const int* X = S.X;
  const int* Y = S.Y;
  const int* hullbuf = S.hullbuf;
  float* hx = S.hx; float* hy = S.hy; float* vx = S.vx; float* vy = S.vy; float* inv = S.inv;
  float* ldx = S.lx; float* ldy = S.ly;
  for (int i = tid; i < n; i += blockDim.x) hx[i] = (float)X[hullbuf[i]], hy[i] = (float)Y[hullbuf[i]];
  __syncthreads();
  // edge vectors (rotcalipers.cpp: differences in double, stored as float; 1/length via double)
  for (int i = tid; i < n; i += blockDim.x) {
    const int nx = i + 1 < n ? i + 1 : 0;
    const double dx = (double)hx[nx] - (double)hx[i];
    const double dy = (double)hy[nx] - (double)hy[i];
    const float fx = (float)dx, fy = (float)dy;
    const float iv = (float)__ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
    vx[i] = fx, vy[i] = fy, inv[i] = iv;
    ldx[i] = __fmul_rn(fx, iv), ldy[i] = __fmul_rn(fy, iv);
  }
  __syncthreads();
  if (tid != 0) return;
  // ---- cv::minAreaRect
  float cx = 0.f, cy = 0.f, w = 0.f, h = 0.f;
  double angle = 0.0;
  if (n > 2) {
    // rotatingCalipers(CALIPERS_MINAREARECT): extreme points with strict comparisons
    int left = 0, bottom = 0, right = 0, top = 0;
    float left_x, right_x, top_y, bottom_y;
    left_x = right_x = hx[0];
    top_y = bottom_y = hy[0];
    for (int i = 1; i < n; ++i) {
      const float px = hx[i], py = hy[i];
      if (px < left_x) left_x = px, left = i;
      if (px > right_x) right_x = px, right = i;
      if (py > top_y) top_y = py, top = i;
      if (py < bottom_y) bottom_y = py, bottom = i;
    }
    float orientation = 0.f;
    {
      double ax = vx[n - 1], ay = vy[n - 1];
      for (int i = 0; i < n; ++i) {
        const double bx = vx[i], by = vy[i];
        const double convexity = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
        if (convexity != 0) {
          orientation = convexity > 0 ? 1.f : -1.f;
          break;
        }
        ax = bx, ay = by;
      }
    }
    float base_a = orientation, base_b = 0.f;
    int seq[4] = {bottom, right, top, left};
    // state carried in registers: edge vector and hull point at each of the four calipers positions;
    // one position advances per step, so one edge vector, one point and one unit lead are (re)loaded
    float ex[4], ey[4], qx[4], qy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ex[i] = vx[seq[i]], ey[i] = vy[seq[i]], qx[i] = hx[seq[i]], qy[i] = hy[seq[i]];
    float minarea = 3.402823466e+38f;
    int b_left = 0, b_bottom = 0;
    float b_a = 0.f, b_w = 0.f, b_b = 0.f, b_h = 0.f;
    for (int k = 0; k < n; ++k) {
      // edge vectors rotated into a common frame; the first one met when rotating is the main edge
      float rx[4], ry[4];
      rx[0] = ex[0], ry[0] = ey[0];
      rx[1] = ey[1], ry[1] = -ex[1];
      rx[2] = -ex[2], ry[2] = -ey[2];
      rx[3] = -ey[3], ry[3] = ex[3];
      int main_element = 0;
      float mx = rx[0], my = ry[0];
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        const float tx = ry[i], ty = -rx[i];  // rotate90CW(rv[i])
        if (__fadd_rn(__fmul_rn(tx, mx), __fmul_rn(ty, my)) < 0.f) main_element = i, mx = rx[i], my = ry[i];
      }
      int pindex = seq[0];
#pragma unroll
      for (int i = 1; i < 4; ++i) pindex = main_element == i ? seq[i] : pindex;
      int nidx = pindex + 1;
      nidx = nidx == n ? 0 : nidx;
      const float lead_x = ldx[pindex], lead_y = ldy[pindex];
      const float nex = vx[nidx], ney = vy[nidx], nqx = hx[nidx], nqy = hy[nidx];
      switch (main_element) {
        case 0: base_a = lead_x, base_b = lead_y; break;
        case 1: base_a = lead_y, base_b = -lead_x; break;
        case 2: base_a = -lead_x, base_b = -lead_y; break;
        default: base_a = -lead_y, base_b = lead_x; break;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (main_element == i) seq[i] = nidx, ex[i] = nex, ey[i] = ney, qx[i] = nqx, qy[i] = nqy;
      float dx = __fsub_rn(qx[1], qx[3]);
      float dy = __fsub_rn(qy[1], qy[3]);
      const float width = __fadd_rn(__fmul_rn(dx, base_a), __fmul_rn(dy, base_b));
      dx = __fsub_rn(qx[2], qx[0]);
      dy = __fsub_rn(qy[2], qy[0]);
      const float height = __fadd_rn(__fmul_rn(-dx, base_b), __fmul_rn(dy, base_a));
      const float area = __fmul_rn(width, height);
      if (area <= minarea) {
        minarea = area;
        b_left = seq[3], b_a = base_a, b_w = width, b_b = base_b, b_h = height, b_bottom = seq[0];
      }
    }
    const float A1 = b_a, B1 = b_b, A2 = -b_b, B2 = b_a;
    const float C1 = __fadd_rn(__fmul_rn(A1, hx[b_left]), __fmul_rn(hy[b_left], B1));
    const float C2 = __fadd_rn(__fmul_rn(A2, hx[b_bottom]), __fmul_rn(hy[b_bottom], B2));
    const float idet = __fdiv_rn(1.f, __fsub_rn(__fmul_rn(A1, B2), __fmul_rn(A2, B1)));
    const float ox = __fmul_rn(__fsub_rn(__fmul_rn(C1, B2), __fmul_rn(C2, B1)), idet);
    const float oy = __fmul_rn(__fsub_rn(__fmul_rn(A1, C2), __fmul_rn(A2, C1)), idet);
    const float o2 = __fmul_rn(A1, b_w), o3 = __fmul_rn(B1, b_w);
    const float o4 = __fmul_rn(A2, b_h), o5 = __fmul_rn(B2, b_h);
    cx = __fadd_rn(ox, __fmul_rn(__fadd_rn(o2, o4), 0.5f));
    cy = __fadd_rn(oy, __fmul_rn(__fadd_rn(o3, o5), 0.5f));
    w = (float)__dsqrt_rn(__dadd_rn(__dmul_rn((double)o2, (double)o2), __dmul_rn((double)o3, (double)o3)));
    h = (float)__dsqrt_rn(__dadd_rn(__dmul_rn((double)o4, (double)o4), __dmul_rn((double)o5, (double)o5)));
    angle = atan2((double)o3, (double)o2);
  } else if (n == 2) {
    cx = __fmul_rn(__fadd_rn(hx[0], hx[1]), 0.5f);
    cy = __fmul_rn(__fadd_rn(hy[0], hy[1]), 0.5f);
    const double dx = (double)__fsub_rn(hx[1], hx[0]);
    const double dy = (double)__fsub_rn(hy[1], hy[0]);
    w = (float)__dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    h = 0.f;
    angle = atan2(dy, dx);
  } else if (n == 1) {
    cx = hx[0], cy = hy[0];
  }
  angle = __ddiv_rn(__dmul_rn(angle, 180.0), 3.1415926535897932384626433832795);
  // OpenCV 4.13 reports the angle in [-90, 0): fold by 180 into [-90, 90), then a >= 0 -> a - 90, swap w/h
  if (angle >= 90.0) angle -= 180.0;
  else if (angle < -90.0) angle += 180.0;
  if (angle >= 0.0) {
    angle -= 90.0;
    const float t = w; w = h; h = t;
  }
  const float fangle = (float)angle;
  if (out_rect) out_rect[0] = cx, out_rect[1] = cy, out_rect[2] = w, out_rect[3] = h, out_rect[4] = fangle;
  // ---- cv::boxPoints (RotatedRect::points) then np.int0 (truncate toward zero)
  const double ra = __ddiv_rn(__dmul_rn((double)fangle, 3.1415926535897932384626433832795), 180.0);
  const float b = __fmul_rn((float)cos(ra), 0.5f);
  const float a = __fmul_rn((float)sin(ra), 0.5f);
  const float p0x = __fsub_rn(__fsub_rn(cx, __fmul_rn(a, h)), __fmul_rn(b, w));
  const float p0y = __fsub_rn(__fadd_rn(cy, __fmul_rn(b, h)), __fmul_rn(a, w));
  const float p1x = __fsub_rn(__fadd_rn(cx, __fmul_rn(a, h)), __fmul_rn(b, w));
  const float p1y = __fsub_rn(__fsub_rn(cy, __fmul_rn(b, h)), __fmul_rn(a, w));
  const float p2x = __fsub_rn(__fmul_rn(2.f, cx), p0x);
  const float p2y = __fsub_rn(__fmul_rn(2.f, cy), p0y);
  const float p3x = __fsub_rn(__fmul_rn(2.f, cx), p1x);
  const float p3y = __fsub_rn(__fmul_rn(2.f, cy), p1y);
  out_box[0] = (int)p0x, out_box[1] = (int)p0y, out_box[2] = (int)p1x, out_box[3] = (int)p1y;
  out_box[4] = (int)p2x, out_box[5] = (int)p2y, out_box[6] = (int)p3x, out_box[7] = (int)p3y;
}

}  // namespace plh
