// evalbox.cu — detection evaluation for sm_100a.
//
// Replaces tool/bboxes.py:252-282 np_bboxes_jaccard (Jaccard of quadrilaterals by rasterising both with
// cv2.drawContours(thickness = -1) and counting mask pixels) and tool/bboxes.py:158-246 bboxes_matching (greedy
// Pascal-VOC matching in detection order, `ignored` ground truth).
//
// No mask is ever materialised: cv2's filled contour is, row by row, a union of at most six intervals per
// quadrilateral (raster.cuh; the reference's mask starts at (0, 0) and is 10 pixels larger than the largest
// coordinate, so only its x < 0 / y < 0 sides clip).
// One warp per (detection, ground truth) pair, one lane per row: the lane builds both interval sets, merges
// each into disjoint sorted intervals and counts |A|, |B|, |A ∩ B| with a two-pointer sweep; integer counts
// are reduced over the warp and iou = (float)((double)inter / (double)union) exactly as numpy does it.
// Pairs whose bounding boxes are disjoint are 0 without looking at a row.
#include <algorithm>

#include "common.cuh"
#include "raster.cuh"

namespace plh {

// sort by start, merge overlaps: disjoint ascending intervals; returns the number of pixels covered
__device__ int normalise(Ivs& s) {
  for (int i = 1; i < s.n; ++i) {
    const int a = s.a[i], b = s.b[i];
    int j = i - 1;
    while (j >= 0 && s.a[j] > a) s.a[j + 1] = s.a[j], s.b[j + 1] = s.b[j], --j;
    s.a[j + 1] = a, s.b[j + 1] = b;
  }
  int m = 0, count = 0;
  for (int i = 0; i < s.n; ++i) {
    if (m > 0 && s.a[i] <= s.b[m - 1]) {
      s.b[m - 1] = max(s.b[m - 1], s.b[i]);
    } else {
      s.a[m] = s.a[i], s.b[m] = s.b[i], ++m;
    }
  }
  s.n = m;
  for (int i = 0; i < m; ++i) count += s.b[i] - s.a[i] + 1;
  return count;
}

__device__ int intersect_count(const Ivs& A, const Ivs& B) {
  int i = 0, j = 0, c = 0;
  while (i < A.n && j < B.n) {
    const int lo = max(A.a[i], B.a[j]), hi = min(A.b[i], B.b[j]);
    if (lo <= hi) c += hi - lo + 1;
    if (A.b[i] < B.b[j]) ++i; else ++j;
  }
  return c;
}

// pair p of image b: detection det_off[b] + q / G_b, ground truth gt_off[b] + q % G_b, q = p - pair_off[b]
__global__ void __launch_bounds__(256)
quad_jaccard_kernel(const int32_t* __restrict__ dets, const int32_t* __restrict__ gts, const int32_t* __restrict__ det_off,
                    const int32_t* __restrict__ gt_off, const long long* __restrict__ pair_off, int B, long long total_pairs,
                    float* __restrict__ jaccard) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < total_pairs; p += nwarps) {
    int lo = 0, hi = B;  // last image whose first pair is <= p
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (pair_off[mid] <= p) lo = mid; else hi = mid;
    }
    const int b = lo;
    const int G = gt_off[b + 1] - gt_off[b];
    const long long q = p - pair_off[b];
    const int di = det_off[b] + (int)(q / G), gi = gt_off[b] + (int)(q % G);
    int ax[4], ay[4], bx[4], by[4];
    bool ok = true;
    int aminx = INT_MAX, amaxx = INT_MIN, aminy = INT_MAX, amaxy = INT_MIN;
    int bminx = INT_MAX, bmaxx = INT_MIN, bminy = INT_MAX, bmaxy = INT_MIN;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ax[k] = dets[(size_t)di * 8 + 2 * k], ay[k] = dets[(size_t)di * 8 + 2 * k + 1];
      bx[k] = gts[(size_t)gi * 8 + 2 * k], by[k] = gts[(size_t)gi * 8 + 2 * k + 1];
      ok = ok && ax[k] > -kMaxCoord && ay[k] > -kMaxCoord && bx[k] > -kMaxCoord && by[k] > -kMaxCoord &&
           ax[k] < kMaxCoord && ay[k] < kMaxCoord && bx[k] < kMaxCoord && by[k] < kMaxCoord;
      aminx = min(aminx, ax[k]), amaxx = max(amaxx, ax[k]), aminy = min(aminy, ay[k]), amaxy = max(amaxy, ay[k]);
      bminx = min(bminx, bx[k]), bmaxx = max(bmaxx, bx[k]), bminy = min(bminy, by[k]), bmaxy = max(bmaxy, by[k]);
    }
    float iou;
    if (!ok) {
      iou = __int_as_float(0x7fc00000);  // |coordinate| >= 2^20
    } else if (amaxx < bminx || bmaxx < aminx || amaxy < bminy || bmaxy < aminy) {
      iou = 0.f;  // every drawn pixel lies inside its quadrilateral's bounding box
    } else {
      long long cA = 0, cB = 0, cI = 0;
      QuadEdges EA, EB;
      // the reference's mask starts at (0, 0) and is 10 pixels larger than the largest coordinate of the image's
      // boxes (tool/bboxes.py:258-262): its right and bottom sides never clip
      const long long far = 4ll * kMaxCoord;
      quad_edges(ax, ay, far, far, EA);
      quad_edges(bx, by, far, far, EB);
      for (int y = max(min(aminy, bminy), 0) + lane; y <= max(amaxy, bmaxy); y += 32) {  // rows above the mask are not drawn
        Ivs A, Bv;
        quad_row_intervals(EA, y, far, A);
        quad_row_intervals(EB, y, far, Bv);
        cA += normalise(A);
        cB += normalise(Bv);
        cI += intersect_count(A, Bv);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cA += __shfl_xor_sync(0xffffffffu, cA, o);
        cB += __shfl_xor_sync(0xffffffffu, cB, o);
        cI += __shfl_xor_sync(0xffffffffu, cI, o);
      }
      iou = (float)__ddiv_rn((double)cI, (double)(cA + cB - cI));  // tool/bboxes.py:279-280
    }
    if (lane == 0) jaccard[p] = iou;
  }
}

// tool/bboxes.py:198-226: one warp per image walks its detections in order.
__global__ void __launch_bounds__(32)
bboxes_matching_kernel(const float* __restrict__ jaccard, const int32_t* __restrict__ det_off,
                       const int32_t* __restrict__ gt_off, const long long* __restrict__ pair_off,
                       const uint8_t* __restrict__ gignored, float thr, uint8_t* __restrict__ gmatch,
                       uint8_t* __restrict__ tp, uint8_t* __restrict__ fp, int32_t* __restrict__ n_gbboxes) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int d0 = det_off[b], D = det_off[b + 1] - d0, g0 = gt_off[b], G = gt_off[b + 1] - g0;
  int cnt = 0;
  for (int g = lane; g < G; g += 32) gmatch[g0 + g] = 0, cnt += gignored[g0 + g] == 0;
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) n_gbboxes[b] = cnt;  // :180 count_nonzero(not ignored)
  __syncwarp();
  const float* J = jaccard + pair_off[b];
  for (int i = 0; i < D; ++i) {
    float best = -1.f;
    int bi = INT_MAX;
    for (int g = lane; g < G; g += 32) {
      const float v = J[(size_t)i * G + g];
      if (v > best) best = v, bi = g;  // strict: the first maximum (tf.argmax)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
    }
    if (lane == 0) {
      bool t = false, f = false;
      if (G > 0) {
        const bool match = best > thr;                        // :208
        const bool existing = gmatch[g0 + bi] != 0;
        const bool not_ignored = gignored[g0 + bi] == 0;
        t = not_ignored && match && !existing;                // :215
        f = not_ignored && (existing || !match);              // :218
        if (not_ignored && match) gmatch[g0 + bi] = 1;        // :222-223
      }
      tp[d0 + i] = t, fp[d0 + i] = f;
    }
    __syncwarp();
  }
}

}  // namespace plh

using namespace plh;

extern "C" int plh_quad_jaccard(const int32_t* dets, const int32_t* gts, const int32_t* det_off, const int32_t* gt_off,
                                const int64_t* pair_off, int B, long long total_pairs, float* jaccard, void* stream) {
  if (B <= 0 || total_pairs < 0) return PLH_E_SHAPE;
  if (total_pairs == 0) return PLH_OK;   // no detections or no ground truth anywhere: nothing to compute
  if (!dets || !gts || !det_off || !gt_off || !pair_off || !jaccard) return PLH_E_NULL;
  const int grid = (int)std::min<long long>((total_pairs + 7) / 8, kNumSMs * 8);
  quad_jaccard_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dets, gts, det_off, gt_off, (const long long*)pair_off, B,
                                                               total_pairs, jaccard);
  return launch_status();
}

extern "C" int plh_bboxes_matching(const float* jaccard, const int32_t* det_off, const int32_t* gt_off,
                                   const int64_t* pair_off, int B, const uint8_t* gignored, float matching_threshold,
                                   uint8_t* gmatch, uint8_t* tp, uint8_t* fp, int32_t* n_gbboxes, void* stream) {
  if (!det_off || !gt_off || !pair_off || !n_gbboxes) return PLH_E_NULL;   // the per-box arrays may be empty
  if (B <= 0) return PLH_E_SHAPE;
  bboxes_matching_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(jaccard, det_off, gt_off, (const long long*)pair_off, gignored,
                                                             matching_threshold, gmatch, tp, fp, n_gbboxes);
  return launch_status();
}
