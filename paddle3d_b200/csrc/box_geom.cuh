// Rotated-rectangle BEV geometry shared by iou3d_nms.cu and postprocess.cu.
//
// Arithmetic follows paddle3d/ops/iou3d_nms/iou3d_nms_kernel.cu:22-273 operation for operation
// (same fp32 expression trees, so nvcc's FMA contraction lands on the same places and the result is
// bit-identical to the reference kernels on the same GPU): clip-free polygon intersection —
// up to 16 edge/edge crossings + corners of one box inside the other, centroid, polar-angle
// ordering, fan shoelace.  A box is 7 floats: x, y, z, dx, dy, dz, heading.
#pragma once
#include <cuda_runtime.h>

namespace p3d {
namespace geom {

struct V2 {
  float x, y;
};

constexpr float kEps = 1e-8f;
constexpr float kInsideMargin = 1e-2f;

__device__ __forceinline__ V2 mk(float x, float y) {
  V2 r;
  r.x = x;
  r.y = y;
  return r;
}

// z of (p1 - p0) x (p2 - p0)
__device__ __forceinline__ float turn(const V2 &p1, const V2 &p2, const V2 &p0) {
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

__device__ __forceinline__ bool bbox_touch(const V2 &p1, const V2 &p2, const V2 &q1, const V2 &q2) {
  return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
         fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}

__device__ __forceinline__ bool inside(const float *box, const V2 &p) {
  const float cx = box[0], cy = box[1];
  const float ca = cosf(-box[6]), sa = sinf(-box[6]);
  const float rx = (p.x - cx) * ca + (p.y - cy) * (-sa);
  const float ry = (p.x - cx) * sa + (p.y - cy) * ca;
  return fabsf(rx) < box[3] / 2 + kInsideMargin && fabsf(ry) < box[4] / 2 + kInsideMargin;
}

// Proper crossing of segments (p0,p1) and (q0,q1); writes the crossing point.
__device__ __forceinline__ bool cross_point(const V2 &p1, const V2 &p0, const V2 &q1, const V2 &q0, V2 *out) {
  if (!bbox_touch(p0, p1, q0, q1)) return false;
  const float s1 = turn(q0, p1, p0);
  const float s2 = turn(p1, q1, p0);
  const float s3 = turn(p0, q1, q0);
  const float s4 = turn(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
  const float s5 = turn(q1, p1, p0);
  if (fabsf(s5 - s1) > kEps) {
    out->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    out->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    const float D = a0 * b1 - a1 * b0;
    out->x = (b0 * c1 - b1 * c0) / D;
    out->y = (a1 * c0 - a0 * c1) / D;
  }
  return true;
}

__device__ __forceinline__ V2 spin(const V2 &c, float ca, float sa, const V2 &p) {
  const float nx = (p.x - c.x) * ca + (p.y - c.y) * (-sa) + c.x;
  const float ny = (p.x - c.x) * sa + (p.y - c.y) * ca + c.y;
  return mk(nx, ny);
}

__device__ inline float overlap_area(const float *a, const float *b) {
  const float ahx = a[3] / 2, bhx = b[3] / 2, ahy = a[4] / 2, bhy = b[4] / 2;
  const float ax1 = a[0] - ahx, ay1 = a[1] - ahy, ax2 = a[0] + ahx, ay2 = a[1] + ahy;
  const float bx1 = b[0] - bhx, by1 = b[1] - bhy, bx2 = b[0] + bhx, by2 = b[1] + bhy;
  const V2 ca = mk(a[0], a[1]), cb = mk(b[0], b[1]);
  V2 A[5], B[5];
  A[0] = mk(ax1, ay1); A[1] = mk(ax2, ay1); A[2] = mk(ax2, ay2); A[3] = mk(ax1, ay2);
  B[0] = mk(bx1, by1); B[1] = mk(bx2, by1); B[2] = mk(bx2, by2); B[3] = mk(bx1, by2);
  const float cosa = cosf(a[6]), sina = sinf(a[6]);
  const float cosb = cosf(b[6]), sinb = sinf(b[6]);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    A[k] = spin(ca, cosa, sina, A[k]);
    B[k] = spin(cb, cosb, sinb, B[k]);
  }
  A[4] = A[0];
  B[4] = B[0];

  V2 poly[16];
  V2 ctr = mk(0.f, 0.f);
  int cnt = 0;
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 4; ++j) {
      if (cross_point(A[i + 1], A[i], B[j + 1], B[j], &poly[cnt])) {
        ctr = mk(ctr.x + poly[cnt].x, ctr.y + poly[cnt].y);
        ++cnt;
      }
    }
  }
  for (int k = 0; k < 4; ++k) {
    if (inside(a, B[k])) {
      ctr = mk(ctr.x + B[k].x, ctr.y + B[k].y);
      poly[cnt++] = B[k];
    }
    if (inside(b, A[k])) {
      ctr = mk(ctr.x + A[k].x, ctr.y + A[k].y);
      poly[cnt++] = A[k];
    }
  }
  ctr.x /= cnt;
  ctr.y /= cnt;
  // order by polar angle about the centroid (bubble sort, as the reference: the comparison is not a
  // strict weak order under NaN/ties, so the exact pass structure matters for bit parity)
  for (int j = 0; j < cnt - 1; ++j) {
    for (int i = 0; i < cnt - j - 1; ++i) {
      if (atan2f(poly[i].y - ctr.y, poly[i].x - ctr.x) > atan2f(poly[i + 1].y - ctr.y, poly[i + 1].x - ctr.x)) {
        const V2 t = poly[i];
        poly[i] = poly[i + 1];
        poly[i + 1] = t;
      }
    }
  }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) {
    const V2 u = mk(poly[k].x - poly[0].x, poly[k].y - poly[0].y);
    const V2 w = mk(poly[k + 1].x - poly[0].x, poly[k + 1].y - poly[0].y);
    area += u.x * w.y - u.y * w.x;
  }
  return fabsf(area) / 2.0f;
}

__device__ inline float iou_rotated(const float *a, const float *b) {
  const float sa = a[3] * a[4];
  const float sb = b[3] * b[4];
  const float so = overlap_area(a, b);
  return so / fmaxf(sa + sb - so, kEps);
}

// axis-aligned variant, iou3d_nms_kernel.cu:365-378
__device__ inline float iou_axis_aligned(const float *a, const float *b) {
  const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
  const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
  const float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
  const float inter = width * height;
  const float Sa = a[3] * a[4];
  const float Sb = b[3] * b[4];
  return inter / fmaxf(Sa + Sb - inter, kEps);
}

}  // namespace geom
}  // namespace p3d
