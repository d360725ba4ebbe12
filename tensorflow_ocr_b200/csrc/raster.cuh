// raster.cuh — cv2's filled polygon (cv2.fillPoly / cv2.drawContours(thickness = -1), OpenCV 4.13) as row intervals.
//
// A filled quadrilateral is, row by row, a union of at most six intervals (oracle/evaluation.py::filled_quad_rows,
// pinned against cv2):
//   * the outline, drawn with cv::line — LineIterator(8-connected, leftToRight): Bresenham from the endpoint
//     with the smaller x, err0 = dx - 2dy on the major axis, so the minor coordinate after i major steps is
//     m_i = (2 dminor i + dmajor - 1) / (2 dmajor); the pixels of an edge on one row are therefore one run
//     with closed-form ends; a segment leaving the image is clipped by cv::clipLine before it is walked;
//   * the scan-line spans: for every row in [ymin, ymax) of the non-horizontal edges, the active edges'
//     abscissae x_e(y) = x_e(y_top) + (y - y_top) * dx_e (16.16 fixed point, dx_e by truncating division)
//     are sorted and consecutive pairs filled from ceil(x_left) to floor(x_right), cut to the image; an edge
//     with an endpoint outside the image takes its slope from the clipped segment.
#pragma once
#include "common.cuh"

namespace plh {

constexpr int kMaxIv = 6;  // 4 outline runs + 2 spans per row of a quadrilateral
constexpr int kMaxCoord = 1 << 20;

struct Ivs {
  int n;
  int a[kMaxIv], b[kMaxIv];
};

__device__ __forceinline__ long long ceil_div_pos(long long a, long long b) {  // b > 0
  return a >= 0 ? (a + b - 1) / b : -((-a) / b);
}

// cv::clipLine(Size(w, h), pt1, pt2): the y side first, then x, each intersection computed in double and
// truncated (imgproc/drawing.cpp; oracle/evaluation.py::clip_line, pinned against cv2.line).
__device__ inline bool clip_line(long long w, long long h, long long& x1, long long& y1, long long& x2, long long& y2) {
  const long long right = w - 1, bottom = h - 1;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    if (c1 & 12) {
      const long long a = c1 < 8 ? 0 : bottom;
      x1 += (long long)__ddiv_rn(__dmul_rn((double)(a - y1), (double)(x2 - x1)), (double)(y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      const long long a = c2 < 8 ? 0 : bottom;
      x2 += (long long)__ddiv_rn(__dmul_rn((double)(a - y2), (double)(x2 - x1)), (double)(y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        const long long a = c1 == 1 ? 0 : right;
        y1 += (long long)__ddiv_rn(__dmul_rn((double)(a - x1), (double)(y2 - y1)), (double)(x2 - x1));
        x1 = a, c1 = 0;
      }
      if (c2) {
        const long long a = c2 == 1 ? 0 : right;
        y2 += (long long)__ddiv_rn(__dmul_rn((double)(a - x2), (double)(y2 - y1)), (double)(x2 - x1));
        x2 = a, c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

// Row-independent description of a quadrilateral's four edges.
struct QuadEdges {
  // outline (after clipping, walked from the endpoint with the smaller x)
  int lx0[4], ly0[4], ldx[4], lady[4], lsy[4];
  bool lvis[4];
  // scan-line edge: active for rows [yt, yb), abscissa xs0 + (y - yt) * dxe in 16.16 fixed point
  int yt[4], yb[4];
  long long xs0[4], dxe[4];
};

__device__ inline void quad_edges(const int* qx, const int* qy, long long w, long long h, QuadEdges& E) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p0x = qx[(i + 3) & 3], p0y = qy[(i + 3) & 3], p1x = qx[i], p1y = qy[i];
    long long a0 = p0x, b0 = p0y, a1 = p1x, b1 = p1y;
    const bool oob = p0x < 0 || p0y < 0 || p1x < 0 || p1y < 0 || p0x >= w || p1x >= w || p0y >= h || p1y >= h;
    const bool vis = oob ? clip_line(w, h, a0, b0, a1, b1) : true;
    {
      int x0 = (int)a0, y0 = (int)b0, dx = (int)(a1 - a0), dy = (int)(b1 - b0);
      if (dx < 0) x0 = (int)a1, y0 = (int)b1, dx = -dx, dy = -dy;
      E.lvis[i] = vis, E.lx0[i] = x0, E.ly0[i] = y0, E.ldx[i] = dx, E.lsy[i] = dy < 0 ? -1 : 1, E.lady[i] = dy < 0 ? -dy : dy;
    }
    E.yt[i] = 0, E.yb[i] = 0, E.xs0[i] = 0, E.dxe[i] = 0;
    if (p0y != p1y) {
      // an edge with an endpoint outside the image takes its abscissae (and, unless the clipped segment is horizontal,
      // its ordinates) from the clipped segment, and is extrapolated back to its own top row
      long long c0x = (long long)p0x << 16, c0y = p0y, c1x = (long long)p1x << 16, c1y = p1y;
      if (oob) {
        if (b0 != b1) c0y = b0, c1y = b1;
        c0x = a0 << 16, c1x = a1 << 16;
      }
      const long long dxe = (c1x - c0x) / (c1y - c0y);  // C++ division truncates, as OpenCV's
      E.dxe[i] = dxe;
      if (p0y < p1y) E.yt[i] = p0y, E.yb[i] = p1y, E.xs0[i] = c0x + (p0y - c0y) * dxe;
      else E.yt[i] = p1y, E.yb[i] = p0y, E.xs0[i] = c1x + (p1y - c1y) * dxe;
    }
  }
}

// Intervals of the filled quadrilateral on row y (0 <= y < h) of a w-wide image, unsorted, possibly overlapping.
__device__ inline void quad_row_intervals(const QuadEdges& E, int y, long long w, Ivs& out) {
  out.n = 0;
  long long xs[4];
  int nx = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (E.lvis[i]) {  // outline run of this edge on row y
      const int dx = E.ldx[i], ady = E.lady[i];
      const int r = (y - E.ly0[i]) * E.lsy[i];
      if (r >= 0 && r <= ady) {
        int lo, hi;
        if (ady > dx) {
          lo = hi = (int)((2ll * dx * r + ady - 1) / (2ll * ady));
        } else if (ady == 0) {
          lo = 0, hi = dx;
        } else {
          lo = (int)max(ceil_div_pos(2ll * dx * r - dx + 1, 2ll * ady), 0ll);
          hi = (int)min(ceil_div_pos(2ll * dx * (r + 1) - dx + 1, 2ll * ady) - 1, (long long)dx);
        }
        if (lo <= hi) out.a[out.n] = E.lx0[i] + lo, out.b[out.n] = E.lx0[i] + hi, ++out.n;
      }
    }
    if (y >= E.yt[i] && y < E.yb[i]) xs[nx++] = E.xs0[i] + (long long)(y - E.yt[i]) * E.dxe[i];
  }
  // sort the (at most 4) abscissae, fill between consecutive pairs
  for (int i = 1; i < nx; ++i) {
    const long long v = xs[i];
    int j = i - 1;
    while (j >= 0 && xs[j] > v) xs[j + 1] = xs[j], --j;
    xs[j + 1] = v;
  }
  for (int k = 0; k + 1 < nx; k += 2) {
    const int x1 = (int)max((xs[k] + 65535) >> 16, 0ll), x2 = (int)min(xs[k + 1] >> 16, w - 1);
    if (x1 <= x2) out.a[out.n] = x1, out.b[out.n] = x2, ++out.n;
  }
}

}  // namespace plh
