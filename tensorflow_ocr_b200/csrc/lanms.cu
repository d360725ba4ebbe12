// lanms.cu — locality-aware NMS for the EAST decode.
//
// NOT IN THE REFERENCE (no `lanms`, no `nms_locality` in the reference tree): restated from
// upstream argman/EAST `locality_aware_nms.py` — PARITY UNPINNED (checked against oracle/east.py).
//   nms_locality(polys, thres): scan the boxes in row-major order; while the next box overlaps the
//   running merged box (IoU > thres) fold it in by score-weighted averaging of the 8 coordinates
//   (scores add), else flush; then standard greedy NMS (descending score) on the flushed boxes.
// The fold is a sequential dependency chain per image: one CTA per image, the fold on one WARP (each IoU
// computed cooperatively by its lanes), the greedy NMS with all threads computing IoUs in parallel.  It is latency-bound integer/fp64
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

// Axis-aligned bounding boxes strictly apart => the quadrilaterals are disjoint => quad_iou is exactly 0.0 (the
// clipped polygon is empty): the expensive clipping is skipped for such pairs with no change of any result.
__device__ inline bool aabb_apart(const double* g, const double* p) {
  const double gx0 = fmin(fmin(g[0], g[2]), fmin(g[4], g[6])), gx1 = fmax(fmax(g[0], g[2]), fmax(g[4], g[6]));
  const double px0 = fmin(fmin(p[0], p[2]), fmin(p[4], p[6])), px1 = fmax(fmax(p[0], p[2]), fmax(p[4], p[6]));
  if (gx1 < px0 || px1 < gx0) return true;
  const double gy0 = fmin(fmin(g[1], g[3]), fmin(g[5], g[7])), gy1 = fmax(fmax(g[1], g[3]), fmax(g[5], g[7]));
  const double py0 = fmin(fmin(p[1], p[3]), fmin(p[5], p[7])), py1 = fmax(fmax(p[1], p[3]), fmax(p[5], p[7]));
  return gy1 < py0 || py1 < gy0;
}

// The same IoU computed by ONE WARP (all lanes call it with the same g, p; the result is uniform): the subject
// polygon of the Sutherland-Hodgman clipping lives one vertex per lane, every clip edge is one parallel step
// (side tests, the intersection point with its fp64 division) followed by a ballot compaction through a small
// shared-memory scratch.  Same formulas and the same summation order as quad_iou, so both give the same bits;
// the sequential fold of the locality-aware NMS is ~7x shorter with it (an fp64 division chain per vertex per
// edge is what the one-thread version spends its 6 us per box on).
__device__ inline double quad_iou_warp(const double* g, const double* p, Pt* scratch /*[16] per warp*/, int lane) {
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
  int na = 4;
  Pt cur = G[lane & 3];                               // lane i < na holds vertex i
  const unsigned lt = (1u << lane) - 1u;
  for (int e = 0; e < 4 && na > 0; ++e) {
    const Pt A = P[e], Bp = P[(e + 1) & 3];
    const int nl = lane + 1 < na ? lane + 1 : 0;
    Pt nxt;
    nxt.x = __shfl_sync(0xffffffffu, cur.x, nl), nxt.y = __shfl_sync(0xffffffffu, cur.y, nl);
    const bool live = lane < na;
    const double sc = (Bp.x - A.x) * (cur.y - A.y) - (Bp.y - A.y) * (cur.x - A.x);
    const double sn = (Bp.x - A.x) * (nxt.y - A.y) - (Bp.y - A.y) * (nxt.x - A.x);
    const bool keep = live && sc >= 0, cross = live && ((sc >= 0) != (sn >= 0));
    const unsigned K = __ballot_sync(0xffffffffu, keep), X = __ballot_sync(0xffffffffu, cross);
    const int pos = __popc(K & lt) + __popc(X & lt);
    if (keep) scratch[pos] = cur;
    if (cross) {
      const double t = sc / (sc - sn);
      scratch[pos + (keep ? 1 : 0)] = Pt{cur.x + t * (nxt.x - cur.x), cur.y + t * (nxt.y - cur.y)};
    }
    na = __popc(K) + __popc(X);
    __syncwarp();
    if (lane < na) cur = scratch[lane];
    __syncwarp();
  }
  double inter = 0.0;
  if (na >= 3) {
    const int nl = lane + 1 < na ? lane + 1 : 0;
    const double nx = __shfl_sync(0xffffffffu, cur.x, nl), ny = __shfl_sync(0xffffffffu, cur.y, nl);
    const double term = cur.x * ny - nx * cur.y;
    double s = 0.0;
    for (int i = 0; i < na; ++i) s += __shfl_sync(0xffffffffu, term, i);   // the order of poly_area
    inter = fabs(0.5 * s);
  }
  const double uni = ag + ap - inter;
  return uni != 0 ? inter / uni : 0.0;
}

// workspace per image: S [n,9] doubles (merged boxes), order [n] ints, supp [n] bytes
__global__ void __launch_bounds__(256)
lanms_kernel(const double* __restrict__ polys, const int* __restrict__ offsets, double thres, int cap,
             double* __restrict__ S_all, int* __restrict__ order_all, double* __restrict__ out,
             int* __restrict__ n_out) {
  __shared__ int s_m;
  __shared__ int s_cur;
  extern __shared__ __align__(16) double s_boxes[];   // [cap][9]: the image's boxes, then (in place) the merged ones
  const int b = blockIdx.x;
  const int o0 = offsets[b], n = offsets[b + 1] - o0;
  const int tid = threadIdx.x;
  // Every step of both phases is a short dependent chain, so the boxes live in shared memory whenever the image's
  // n fits (a global / L2 round trip per step is what the time went to otherwise); the fold writes merged box m
  // over input box m <= i - 1, which has been read already.
  const bool resident = n <= cap;
  const double* in = polys + (size_t)o0 * 9;
  double* S = S_all + (size_t)o0 * 9;
  int* order = order_all + o0;
  if (resident) {
    for (int k = tid; k < n * 9; k += blockDim.x) s_boxes[k] = in[k];
    in = s_boxes, S = s_boxes;
    order = reinterpret_cast<int*>(s_boxes + (size_t)cap * 9);
  }
  __syncthreads();
  // ---- phase 1: the locality-aware fold (sequential over the boxes by construction), on warp 0: every IoU is
  // computed by the whole warp, the merge one coordinate per lane; the next box is fetched one step ahead
  __shared__ Pt s_scratch[16];
  if (tid < 32) {
    const int lane = tid;
    int m = 0;
    double p[9];
    bool have = false;
    double gn = (n > 0 && lane < 9) ? in[lane] : 0.0;   // lane k holds g[k] of the next box
    double pl = 0.0;                                    // lane k holds p[k] (besides the replicated copy p[])
    for (int i = 0; i < n; ++i) {
      const double gl = gn;                             // lane k: g[k]
      double g[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) g[k] = __shfl_sync(0xffffffffu, gl, k);
      if (i + 1 < n && lane < 9) gn = in[(size_t)(i + 1) * 9 + lane];
      if (have && !aabb_apart(g, p) && quad_iou_warp(g, p, s_scratch, lane) > thres) {
        // weighted_merge(g, p): g[:8] = (g[8] g[:8] + p[8] p[:8]) / (g[8] + p[8]); g[8] += p[8] — lane k does coordinate k
        const double wg = g[8], wp = p[8];
        pl = lane < 8 ? (wg * gl + wp * pl) / (wg + wp) : wg + wp;
#pragma unroll
        for (int k = 0; k < 9; ++k) p[k] = __shfl_sync(0xffffffffu, pl, k < 8 ? k : 8);
      } else {
        if (have) {
          if (lane < 9) S[(size_t)m * 9 + lane] = pl;
          ++m;
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) p[k] = g[k];
        pl = gl;
        have = true;
      }
    }
    if (have) {
      if (lane < 9) S[(size_t)m * 9 + lane] = pl;
      ++m;
    }
    if (lane == 0) s_m = m;
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
      double cb[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) cb[k] = S[(size_t)cur * 9 + k];
      for (int t = c + 1 + tid; t < m; t += blockDim.x) {
        const int o = order[t];
        if (o < 0) continue;
        double ob[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) ob[k] = S[(size_t)o * 9 + k];
        if (!aabb_apart(cb, ob) && quad_iou(cb, ob) > thres) order[t] = -1 - o;
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
  constexpr int kCap = 2800;                        // boxes per image held in shared memory (72 B each + 4 B of order: 212.8 KB)
  static SmemOptIn optin;
  if (int rc = ensure_dynamic_smem(optin, lanms_kernel, (size_t)kCap * 76)) return rc;
  lanms_kernel<<<B, 256, (size_t)kCap * 76, (cudaStream_t)stream>>>(polys, offsets, thres, kCap, S, order, out, n_out);
  return launch_status();
}
