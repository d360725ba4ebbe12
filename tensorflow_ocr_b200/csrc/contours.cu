// contours.cu — the contour path of the decode for sm_100a.
//
// Replaces test.py:182-218 of the reference: cv2.findContours(mask, RETR_TREE, CHAIN_APPROX_SIMPLE) — hole
// borders included — then per contour cv2.minAreaRect -> cv2.boxPoints -> np.int0, x4, /ratio (assigned into an
// integer array: truncation), order_points (test.py:24-35).
//
// OpenCV's border following is a sequential raster scan whose marks decide where later borders start.  The
// parallel statement used here (model: oracle/contours.py, pinned against cv2 4.13 point by point): an OUTER
// border starts exactly at the raster-first pixel of every 8-connected foreground component, a HOLE border
// exactly at the left neighbour of the raster-first pixel of every 4-connected background region that does not
// touch the frame; given the starts every border is traced independently on the unmarked binary image.
//   C1 contour_flags    mask -> two flag maps per image, padded by 2: foreground with all 8 links set, background
//                       with the 4 axis links set (the padding ring joins everything that touches the frame)
//   C2 (decode.cu)      component labels of the 2B maps with the link-graph kernels (label = minimum pixel index)
//   C3 contour_trace    one thread per start: trace once to count the CHAIN_APPROX_SIMPLE points, reserve a span
//                       of the image's point pool, trace again to write them; border-tree parent from the labels
//   C4 contour_boxes    one CTA per contour: OpenCV-exact hull + rotating calipers (rect.cuh), x4, /ratio,
//                       order_points
// The host mirror (tensorflow_ocr_b200/decode.py) puts the contours into OpenCV's output order (pre-order of the
// border tree, siblings in reverse discovery order).
#include <algorithm>

#include "common.cuh"
#include "rect.cuh"

namespace plh {

int decode_labels_only(const uint16_t* flags, int B, int H, int W, int32_t* labels, int32_t* n_boxes, int32_t* scratch_boxes,
                       void* workspace, size_t workspace_bytes, cudaStream_t s);  // decode.cu
size_t decode_workspace_bytes(int B, int H, int W, int K);

constexpr int kPad = 2;
constexpr int kCtFlagP = 1 << 8;

struct ContourWs {
  size_t flags, labels, nb, scratch, npts, pool, decode, total;
};
static ContourWs contour_ws_layout(int B, int H, int W) {
  const int Hp = H + 2 * kPad, Wp = W + 2 * kPad;
  const size_t px2 = (size_t)2 * B * Hp * Wp;
  ContourWs l;
  size_t off = 0;
  l.flags = off; off = align_up(off + px2 * 2, 256);
  l.labels = off; off = align_up(off + px2 * 4, 256);
  l.nb = off; off = align_up(off + (size_t)2 * B * 4, 256);
  l.scratch = off; off = align_up(off + (size_t)2 * B * 8 * 4, 256);
  l.npts = off; off = align_up(off + (size_t)B * 4, 256);
  l.pool = off; off = align_up(off + (size_t)B * 2 * H * W * 8, 256);      // <= 2 points per pixel, int2 each
  l.decode = off; off = align_up(off + decode_workspace_bytes(2 * B, Hp, Wp, 1), 256);
  l.total = off;
  return l;
}
size_t contour_workspace_bytes(int B, int H, int W) { return contour_ws_layout(B, H, W).total; }

// ------------------------------------------------------------------ C1
__global__ void __launch_bounds__(256)
contour_flags_kernel(const uint8_t* __restrict__ mask, int B, int H, int W, uint16_t* __restrict__ flags,
                     int* __restrict__ n_contours, int* __restrict__ npts) {
  const int Hp = H + 2 * kPad, Wp = W + 2 * kPad;
  const long long total = (long long)B * Hp * Wp;
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < B; i += blockDim.x) n_contours[i] = 0, npts[i] = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / ((long long)Hp * Wp));
    const int r = (int)(i - (long long)b * Hp * Wp);
    const int y = r / Wp - kPad, x = r % Wp - kPad;
    const bool in = y >= 0 && y < H && x >= 0 && x < W;
    const bool fg = in && mask[((size_t)b * H + y) * W + x] != 0;
    // foreground: 8-connected (all links set); background: 4-connected (left 0, right 3, up 6, down 7)
    flags[((size_t)2 * b) * Hp * Wp + r] = fg ? (uint16_t)(kCtFlagP | 0xFF) : (uint16_t)0;
    flags[((size_t)2 * b + 1) * Hp * Wp + r] = fg ? (uint16_t)0 : (uint16_t)(kCtFlagP | 0xC9);
  }
}

// ------------------------------------------------------------------ C3
// OpenCV direction codes (CV_INIT_3X3_DELTAS): 0 = right, then counter-clockwise in image coordinates
__constant__ int c_cdx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
__constant__ int c_cdy[8] = {0, -1, -1, -1, 0, 1, 1, 1};

// icvFetchContour on the padded foreground flag map: CHAIN_APPROX_SIMPLE points of the border that starts at
// padded pixel (x0, y0).  out == nullptr: count only.  Returns the number of points.
__device__ int trace_border(const uint16_t* __restrict__ f, int Wp, int x0, int y0, bool is_hole, int2* out, int cap) {
  auto on = [&](int x, int y) { return (f[y * Wp + x] & kCtFlagP) != 0; };
  int s_end = is_hole ? 0 : 4, s = s_end;
  int x1, y1;
  do {
    s = (s - 1) & 7;
    x1 = x0 + c_cdx[s], y1 = y0 + c_cdy[s];
  } while (!on(x1, y1) && s != s_end);
  if (!on(x1, y1)) {  // isolated pixel
    if (out && cap > 0) out[0] = make_int2(x0 - kPad, y0 - kPad);
    return 1;
  }
  int n = 0;
  int x3 = x0, y3 = y0, prev_s = s ^ 4, px = x0, py = y0;
  while (true) {
    int x4, y4;
    do {
      ++s;
      x4 = x3 + c_cdx[s & 7], y4 = y3 + c_cdy[s & 7];
    } while (!on(x4, y4));
    s &= 7;
    if (s != prev_s) {  // the direction changes here
      if (out && n < cap) out[n] = make_int2(px - kPad, py - kPad);
      ++n;
      prev_s = s;
    }
    px += c_cdx[s], py += c_cdy[s];
    if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) break;
    x3 = x4, y3 = y4;
    s = (s + 4) & 7;
  }
  return n;
}

// info per contour: scan position (y * W + x of the pixel the raster scan is at when the border is found), hole
// flag, own key, parent key (-1: none), first point, number of points.  Keys: 2 * label for outer borders,
// 2 * label + 1 for hole borders (label = the region's minimum padded pixel index).
constexpr int kInfoInts = 6;
__global__ void __launch_bounds__(128)
contour_trace_kernel(const uint16_t* __restrict__ flags, const int32_t* __restrict__ labels, int B, int H, int W, int K,
                     int* __restrict__ n_contours, int* __restrict__ npts, int32_t* __restrict__ info,
                     int2* __restrict__ pool) {
  pdl_wait_and_release();
  const int Hp = H + 2 * kPad, Wp = W + 2 * kPad;
  const long long total = (long long)B * H * W;
  const int cap = 2 * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / ((long long)H * W));
    const int r = (int)(i - (long long)b * H * W);
    const int y = r / W, x = r - y * W;
    const int pp = (y + kPad) * Wp + x + kPad;                       // padded pixel index
    const uint16_t* ffg = flags + ((size_t)2 * b) * Hp * Wp;
    const int32_t* lfg = labels + ((size_t)2 * b) * Hp * Wp;
    const int32_t* lbg = lfg + (size_t)Hp * Wp;
    const int outside = lbg[1 * Wp + 1];                               // the region of the padding ring
    const bool fg = (ffg[pp] & kCtFlagP) != 0;
    bool start = false, hole = false;
    int key = 0, parent = -1;
    if (fg) {
      if (lfg[pp] == pp) {                                             // raster-first pixel of a foreground component
        start = true;
        const int left = lbg[pp - 1];
        key = 2 * pp, parent = left == outside ? -1 : 2 * left + 1;
      }
    } else if (lbg[pp] == pp && pp != outside) {                       // raster-first pixel of a hole
      start = hole = true;
      key = 2 * pp + 1, parent = 2 * lfg[pp - 1];
    }
    if (!start) continue;
    const int x0 = x + kPad - (hole ? 1 : 0), y0 = y + kPad;
    const int n = trace_border(ffg, Wp, x0, y0, hole, nullptr, 0);
    const int slot = atomicAdd(&n_contours[b], 1);
    const int off = atomicAdd(&npts[b], n);
    if (slot < K) {
      int32_t* q = info + ((size_t)b * K + slot) * kInfoInts;
      const bool fits = off + n <= cap;
      q[0] = r, q[1] = hole, q[2] = key, q[3] = parent, q[4] = fits ? off : -1, q[5] = n;
      if (fits) trace_border(ffg, Wp, x0, y0, hole, pool + (size_t)b * cap + off, n);
    }
  }
}

// ------------------------------------------------------------------ C4
// test.py:24-35 order_points on one integer box (x_sorted by a stable sort, left pair by y, the right pair by
// distance from the top-left: squared integer distances order like scipy's euclidean cdist).
__device__ void order_points_dev(const long long (&b)[8], int32_t* out) {
  int idx[4] = {0, 1, 2, 3};
  for (int i = 1; i < 4; ++i)  // insertion sort by x: stable, like numpy's sort of a 4-element array
    for (int j = i; j > 0 && b[2 * idx[j]] < b[2 * idx[j - 1]]; --j) { const int t = idx[j]; idx[j] = idx[j - 1]; idx[j - 1] = t; }
  int l0 = idx[0], l1 = idx[1], r0 = idx[2], r1 = idx[3];
  if (b[2 * l1 + 1] < b[2 * l0 + 1]) { const int t = l0; l0 = l1; l1 = t; }   // tl, bl by y (stable)
  const long long tx = b[2 * l0], ty = b[2 * l0 + 1];
  const long long d0 = (b[2 * r0] - tx) * (b[2 * r0] - tx) + (b[2 * r0 + 1] - ty) * (b[2 * r0 + 1] - ty);
  const long long d1 = (b[2 * r1] - tx) * (b[2 * r1] - tx) + (b[2 * r1 + 1] - ty) * (b[2 * r1 + 1] - ty);
  // (br, tr) = rightMost[np.argsort(D)[::-1]]: argsort ascending (ties keep order), reversed
  const int br = d0 > d1 ? r0 : r1, tr = d0 > d1 ? r1 : r0;
  const int o[4] = {l0, tr, br, l1};
  for (int i = 0; i < 4; ++i) out[2 * i] = (int32_t)b[2 * o[i]], out[2 * i + 1] = (int32_t)b[2 * o[i] + 1];
}

template <int NPAD, bool LAST>
__global__ void __launch_bounds__(256)
contour_boxes_kernel(const int* __restrict__ n_contours, const int32_t* __restrict__ info, const int2* __restrict__ pool,
                     int B, int H, int W, int K, int min_pts, double ratio_w, double ratio_h, int32_t* __restrict__ boxes,
                     int32_t* __restrict__ raw_boxes) {
  pdl_wait_and_release();
  extern __shared__ __align__(16) unsigned char smem[];
  RectSmem S = rect_carve(smem, NPAD);
  const int cap = 2 * H * W;
  for (int g = blockIdx.x; g < B * K; g += gridDim.x) {
    const int b = g / K, slot = g - b * K;
    if (slot >= min(n_contours[b], K)) continue;
    const int32_t* q = info + ((size_t)b * K + slot) * kInfoInts;
    const int off = q[4], total = q[5];
    if (total <= min_pts || (total > NPAD && !LAST)) continue;   // the other launch of this kernel takes it
    int32_t* ob = boxes + ((size_t)b * K + slot) * 8;
    if (off < 0 || total > NPAD) {                    // point pool overflow / more points than the kernel holds: sentinel
      if (threadIdx.x < 8) ob[threadIdx.x] = INT32_MIN;
      continue;
    }
    const int2* pts = pool + (size_t)b * cap + off;
    int nsort = 32;
    while (nsort < total) nsort <<= 1;
    for (int i = threadIdx.x; i < nsort; i += blockDim.x)
      S.keys[i] = i < total ? make_key(pts[i].x, pts[i].y, i) : ~0ull;
    __syncthreads();
    bitonic_sort(S.keys, nsort);
    int box[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float rect[5];
    min_area_box_sorted(S, total, NPAD, box, rect);
    if (threadIdx.x == 0) {
      long long v[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // box[:,0] = box[:,0] * 4; box[:,0] = box[:,0] / ratio_w  (float64 division assigned into an int array)
        v[2 * i] = (long long)((double)((long long)box[2 * i] * 4) / ratio_w);
        v[2 * i + 1] = (long long)((double)((long long)box[2 * i + 1] * 4) / ratio_h);
      }
      if (raw_boxes)
        for (int i = 0; i < 8; ++i) raw_boxes[((size_t)b * K + slot) * 8 + i] = (int32_t)v[i];
      order_points_dev(v, ob);
    }
    __syncthreads();
  }
}

}  // namespace plh

using namespace plh;

extern "C" size_t plh_contour_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return contour_workspace_bytes(B, H, W);
}

extern "C" int plh_contour_boxes(const uint8_t* mask, int B, int H, int W, double ratio_w, double ratio_h, int K,
                                 int32_t* boxes, int32_t* raw_boxes, int32_t* info, int32_t* n_contours, int32_t* points,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!mask || !boxes || !info || !n_contours) return PLH_E_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || K <= 0 || H + 2 * kPad > 1024 || W + 2 * kPad > 2048 ||
      (long long)2 * B * (H + 2 * kPad) * (W + 2 * kPad) > (1ll << 30))
    return PLH_E_SHAPE;
  if (!(ratio_w > 0.0) || !(ratio_h > 0.0)) return PLH_E_PARAM;
  const ContourWs l = contour_ws_layout(B, H, W);
  if (!workspace || !aligned16(workspace) || workspace_bytes < l.total) return PLH_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  uint16_t* flags = (uint16_t*)(ws + l.flags);
  int32_t* labels = (int32_t*)(ws + l.labels);
  int* npts = (int*)(ws + l.npts);
  int2* pool = points ? (int2*)points : (int2*)(ws + l.pool);
  const int Hp = H + 2 * kPad, Wp = W + 2 * kPad;
  const long long total_p = (long long)B * Hp * Wp;
  int rc = launch_plain(contour_flags_kernel, (int)std::min<long long>((total_p + 255) / 256, kNumSMs * 32), 256, 0, s, mask, B,
                        H, W, flags, (int*)n_contours, npts);
  if (rc) return rc;
  rc = decode_labels_only(flags, 2 * B, Hp, Wp, labels, (int32_t*)(ws + l.nb), (int32_t*)(ws + l.scratch), ws + l.decode,
                          l.total - l.decode, s);
  if (rc) return rc;
  const long long total = (long long)B * H * W;
  rc = launch(contour_trace_kernel, (int)std::min<long long>((total + 127) / 128, kNumSMs * 64), 128, 0, s,
              (const uint16_t*)flags, (const int32_t*)labels, B, H, W, K, (int*)n_contours, npts, info, pool);
  if (rc) return rc;
  // two launches of the box kernel: short borders (the many) with a small shared-memory footprint, long ones with a large
  static SmemOptIn optin_small, optin_big;
  constexpr int kSmall = 128, kBig = 2048;
  if ((rc = ensure_dynamic_smem(optin_small, contour_boxes_kernel<kSmall, false>, rect_smem_bytes(kSmall)))) return rc;
  if ((rc = ensure_dynamic_smem(optin_big, contour_boxes_kernel<kBig, true>, rect_smem_bytes(kBig)))) return rc;
  const int grid = (int)std::min<long long>((long long)B * K, kNumSMs * 16);
  rc = launch(contour_boxes_kernel<kSmall, false>, grid, 256, rect_smem_bytes(kSmall), s, (const int*)n_contours, (const int32_t*)info,
              (const int2*)pool, B, H, W, K, 0, ratio_w, ratio_h, boxes, raw_boxes);
  if (rc) return rc;
  return launch(contour_boxes_kernel<kBig, true>, grid, 256, rect_smem_bytes(kBig), s, (const int*)n_contours, (const int32_t*)info,
                (const int2*)pool, B, H, W, K, kSmall, ratio_w, ratio_h, boxes, raw_boxes);
}
