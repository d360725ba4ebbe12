// lanms.cu — locality-aware NMS for the EAST decode.
//
// NOT IN THE REFERENCE (no `lanms`, no `nms_locality` in the reference tree): restated from
// upstream argman/EAST `locality_aware_nms.py` — PARITY UNPINNED (checked against oracle/east.py).
//   nms_locality(polys, thres): scan the boxes in row-major order; while the next box overlaps the
//   running merged box (IoU > thres) fold it in by score-weighted averaging of the 8 coordinates
//   (scores add), else flush; then standard greedy NMS (descending score) on the flushed boxes.
// The fold is a sequential dependency chain per image: one CTA per image, the fold on one thread,
// the greedy NMS with all threads computing IoUs in parallel.  It is latency-bound integer/fp64
// work and is kept out of every roofline claim.
#include "common.cuh"

namespace plh {

struct Pt {
  double x, y;
};

__device__ inline double poly_area(const Pt* p, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) {
    const int k = i + 1 < n ? i + 1 : 0;
    s += p[i].x * p[k].y - p[k].x * p[i].y;
  }
  return 0.5 * s;
}

// IoU of two convex quadrilaterals (8 doubles each), Sutherland-Hodgman clipping in fp64.
// Same operation order as oracle/east.py::quad_iou.
__device__ inline double quad_iou(const double* g, const double* p) {
  Pt G[4], P[4];
  for (int i = 0; i < 4; ++i) G[i] = Pt{g[2 * i], g[2 * i + 1]}, P[i] = Pt{p[2 * i], p[2 * i + 1]};
  double ag = poly_area(G, 4), ap = poly_area(P, 4);
  if (ag < 0) {
    Pt t = G[0]; G[0] = G[3]; G[3] = t; t = G[1]; G[1] = G[2]; G[2] = t;
    ag = -ag;
  }
  if (ap < 0) {
    Pt t = P[0]; P[0] = P[3]; P[3] = t; t = P[1]; P[1] = P[2]; P[2] = t;
    ap = -ap;
  }
  if (ag == 0 || ap == 0) return 0.0;
  Pt a[12], b[12];
  int na = 4;
  for (int i = 0; i < 4; ++i) a[i] = G[i];
  for (int e = 0; e < 4 && na > 0; ++e) {
    const Pt A = P[e], Bp = P[(e + 1) & 3];
    int nb = 0;
    for (int i = 0; i < na; ++i) {
      const Pt cur = a[i], nxt = a[i + 1 < na ? i + 1 : 0];
      const double sc = (Bp.x - A.x) * (cur.y - A.y) - (Bp.y - A.y) * (cur.x - A.x);
      const double sn = (Bp.x - A.x) * (nxt.y - A.y) - (Bp.y - A.y) * (nxt.x - A.x);
      if (sc >= 0) b[nb++] = cur;
      if ((sc >= 0) != (sn >= 0)) {
        const double t = sc / (sc - sn);
        b[nb++] = Pt{cur.x + t * (nxt.x - cur.x), cur.y + t * (nxt.y - cur.y)};
      }
    }
    na = nb;
    for (int i = 0; i < nb; ++i) a[i] = b[i];
  }
  const double inter = na >= 3 ? fabs(poly_area(a, na)) : 0.0;
  const double uni = ag + ap - inter;
  return uni != 0 ? inter / uni : 0.0;
}

// workspace per image: S [n,9] doubles (merged boxes), order [n] ints, supp [n] bytes
__global__ void __launch_bounds__(256)
lanms_kernel(const double* __restrict__ polys, const int* __restrict__ offsets, double thres,
             double* __restrict__ S_all, int* __restrict__ order_all, double* __restrict__ out,
             int* __restrict__ n_out) {
  __shared__ int s_m;
  __shared__ int s_cur;
  const int b = blockIdx.x;
  const int o0 = offsets[b], n = offsets[b + 1] - o0;
  double* S = S_all + (size_t)o0 * 9;
  int* order = order_all + o0;
  const int tid = threadIdx.x;
  // ---- phase 1: the locality-aware fold (sequential by construction)
  if (tid == 0) {
    int m = 0;
    double p[9];
    bool have = false;
    for (int i = 0; i < n; ++i) {
      const double* g = polys + (size_t)(o0 + i) * 9;
      if (have && quad_iou(g, p) > thres) {
        // weighted_merge(g, p): g[:8] = (g[8] g[:8] + p[8] p[:8]) / (g[8] + p[8]); g[8] += p[8]
        const double wg = g[8], wp = p[8];
        for (int k = 0; k < 8; ++k) p[k] = (wg * g[k] + wp * p[k]) / (wg + wp);
        p[8] = wg + wp;
      } else {
        if (have) {
          for (int k = 0; k < 9; ++k) S[(size_t)m * 9 + k] = p[k];
          ++m;
        }
        for (int k = 0; k < 9; ++k) p[k] = g[k];
        have = true;
      }
    }
    if (have) {
      for (int k = 0; k < 9; ++k) S[(size_t)m * 9 + k] = p[k];
      ++m;
    }
    s_m = m;
  }
  __syncthreads();
  const int m = s_m;
  // ---- phase 2: standard NMS.  order = indices by descending score (ties: ascending index);
  // rank by counting (m is a few hundred at most).
  for (int i = tid; i < m; i += blockDim.x) {
    const double si = S[(size_t)i * 9 + 8];
    int r = 0;
    for (int k = 0; k < m; ++k) {
      const double sk = S[(size_t)k * 9 + 8];
      r += (sk > si) || (sk == si && k < i);
    }
    order[r] = i;
  }
  __syncthreads();
  // greedy: walk the order; every kept box suppresses the later ones it overlaps (IoUs in parallel).
  // order[] entries are negated-minus-one when suppressed.
  int kept = 0;
  for (int c = 0; c < m; ++c) {
    if (tid == 0) s_cur = order[c];
    __syncthreads();
    const int cur = s_cur;
    if (cur >= 0) {
      for (int t = c + 1 + tid; t < m; t += blockDim.x) {
        const int o = order[t];
        if (o >= 0 && quad_iou(S + (size_t)cur * 9, S + (size_t)o * 9) > thres) order[t] = -1 - o;
      }
      if (tid == 0)
        for (int k = 0; k < 9; ++k) out[(size_t)(o0 + kept) * 9 + k] = S[(size_t)cur * 9 + k];
      ++kept;
    }
    __syncthreads();
  }
  if (tid == 0) n_out[b] = kept;
}

}  // namespace plh

using namespace plh;

extern "C" int plh_lanms(const double* polys, const int32_t* offsets, int B, int total, double thres, double* out,
                         int32_t* n_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!offsets || !n_out || (total > 0 && (!polys || !out))) return PLH_E_NULL;
  if (B <= 0 || total < 0) return PLH_E_SHAPE;
  const size_t need = align_up((size_t)total * 9 * sizeof(double), 256) + (size_t)total * sizeof(int) + 256;
  if (!workspace || workspace_bytes < need || !aligned16(workspace)) return PLH_E_WORKSPACE;
  double* S = (double*)workspace;
  int* order = (int*)((char*)workspace + align_up((size_t)total * 9 * sizeof(double), 256));
  lanms_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(polys, offsets, thres, S, order, out, n_out);
  return launch_status();
}
