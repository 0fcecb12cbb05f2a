// dmath.cuh — device math for libcannon_cuda.so.
//
// Numeric contract: the reference stores every vector in a Float32List and evaluates
// every expression in IEEE double without fused multiply-add. All helpers here widen float -> double,
// compute in double in the reference's association order and round once at the store. The translation
// unit is compiled with -fmad=false so nvcc never contracts a*b+c; double division and sqrt are IEEE
// correctly rounded on sm_100a.
//
// Reference anchors: lib/math/vec3.dart, lib/math/quaternion.dart, lib/math/mat3.dart,
// lib/math/transform.dart and package:vector_math's Vector3/Quaternion instance members
// (dot / length2 / normalize accumulate left to right; normalize multiplies by 1/len).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define HD __host__ __device__ __forceinline__

struct f3 { float x, y, z; };
struct q4 { float x, y, z, w; };
struct m33 { float e[9]; };

HD double W(float v) { return (double)v; }
HD f3 mk3(double x, double y, double z) { f3 r; r.x = (float)x; r.y = (float)y; r.z = (float)z; return r; }
HD f3 ld3(const float4& v) { f3 r; r.x = v.x; r.y = v.y; r.z = v.z; return r; }
HD q4 ldq(const float4& v) { q4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
HD float4 st3(const f3& v, float w = 0.f) { return make_float4(v.x, v.y, v.z, w); }
HD float4 stq(const q4& v) { return make_float4(v.x, v.y, v.z, v.w); }

HD f3 vadd(const f3& a, const f3& b) { return mk3(W(a.x) + W(b.x), W(a.y) + W(b.y), W(a.z) + W(b.z)); }    // vec3.dart:18
HD f3 vsub(const f3& a, const f3& b) { return mk3(W(a.x) - W(b.x), W(a.y) - W(b.y), W(a.z) - W(b.z)); }    // vec3.dart:26
HD f3 vscale(double s, const f3& a) { return mk3(s * W(a.x), s * W(a.y), s * W(a.z)); }                    // vec3.dart:50
HD f3 vneg(const f3& a) { f3 r; r.x = -a.x; r.y = -a.y; r.z = -a.z; return r; }
HD f3 vcross(const f3& a, const f3& b) {                                                                     // vec3.dart:59
  return mk3(W(a.y) * W(b.z) - W(a.z) * W(b.y), W(a.z) * W(b.x) - W(a.x) * W(b.z), W(a.x) * W(b.y) - W(a.y) * W(b.x));
}
HD f3 vmulc(const f3& a, const f3& b) { return mk3(W(b.x) * W(a.x), W(b.y) * W(a.y), W(b.z) * W(a.z)); }   // vec3.dart:73
HD double vdot(const f3& a, const f3& b) {
  double s = W(a.x) * W(b.x);
  s += W(a.y) * W(b.y);
  s += W(a.z) * W(b.z);
  return s;
}
HD double vlen2(const f3& a) { return vdot(a, a); }
HD double vlen(const f3& a) { return sqrt(vlen2(a)); }
HD double vnormalize(f3& a) {  // Vector3.normalize()
  double l = vlen(a);
  if (l == 0.0) return 0.0;
  double d = 1.0 / l;
  a = mk3(W(a.x) * d, W(a.y) * d, W(a.z) * d);
  return l;
}
HD double vdist(const f3& a, const f3& b) {  // Vector3.distanceTo
  double dx = W(a.x) - W(b.x), dy = W(a.y) - W(b.y), dz = W(a.z) - W(b.z);
  return sqrt(dx * dx + dy * dy + dz * dz);
}
HD f3 vlerp(const f3& a, const f3& b, double t) {                                                            // vec3.dart:42
  return mk3(W(a.x) + (W(b.x) - W(a.x)) * t, W(a.y) + (W(b.y) - W(a.y)) * t, W(a.z) + (W(b.z) - W(a.z)) * t);
}
HD f3 vunit(const f3& a) {                                                                                   // vec3.dart:122
  double n = sqrt(W(a.x) * W(a.x) + W(a.y) * W(a.y) + W(a.z) * W(a.z));
  if (n > 0.0) {
    n = 1.0 / n;
    return mk3(W(a.x) * n, W(a.y) * n, W(a.z) * n);
  }
  f3 r; r.x = 1.f; r.y = 0.f; r.z = 0.f;
  return r;
}
HD f3 vaddscaled(const f3& a, double s, const f3& b) {                                                       // vec3.dart:140
  return mk3(W(a.x) + s * W(b.x), W(a.y) + s * W(b.y), W(a.z) + s * W(b.z));
}
HD bool valmost_eq(const f3& a, const f3& b) {                                                               // vec3.dart:148
  const double p = 1e-6;
  return !(fabs(W(a.x) - W(b.x)) > p || fabs(W(a.y) - W(b.y)) > p || fabs(W(a.z) - W(b.z)) > p);
}
HD bool valmost_zero(const f3& a) {                                                                          // vec3.dart:161
  const double p = 1e-6;
  return !(fabs(W(a.x)) > p || fabs(W(a.y)) > p || fabs(W(a.z)) > p);
}
HD void vtangents(const f3& a, f3& t1, f3& t2) {                                                             // vec3.dart:97
  double norm = vlen(a);
  if (norm > 0.0) {
    double inorm = 1 / norm;
    f3 n = mk3(W(a.x) * inorm, W(a.y) * inorm, W(a.z) * inorm);
    f3 r;
    if (fabs(W(n.x)) < 0.9) { r.x = 1.f; r.y = 0.f; r.z = 0.f; }
    else { r.x = 0.f; r.y = 1.f; r.z = 0.f; }
    t1 = vcross(n, r);
    t2 = vcross(n, t1);
  } else {
    t1.x = 1.f; t1.y = 0.f; t1.z = 0.f;
    t2.x = 0.f; t2.y = 1.f; t2.z = 0.f;
  }
}

HD f3 qrot(const q4& q, const f3& v) {                                                                       // quaternion.dart:21
  double x = W(v.x), y = W(v.y), z = W(v.z);
  double qx = W(q.x), qy = W(q.y), qz = W(q.z), qw = W(q.w);
  double ix = qw * x + qy * z - qz * y;
  double iy = qw * y + qz * x - qx * z;
  double iz = qw * z + qx * y - qy * x;
  double iw = -qx * x - qy * y - qz * z;
  return mk3(ix * qw + iw * -qx + iy * -qz - iz * -qy, iy * qw + iw * -qy + iz * -qx - ix * -qz,
             iz * qw + iw * -qz + ix * -qy - iy * -qx);
}
HD q4 qconj(const q4& q) { q4 r; r.x = -q.x; r.y = -q.y; r.z = -q.z; r.w = q.w; return r; }
HD q4 qnegw(const q4& q) { q4 r; r.x = q.x; r.y = q.y; r.z = q.z; r.w = -q.w; return r; }  // transform.dart:65-71
HD q4 qmul(const q4& a, const q4& b) {  // Quaternion.multiply2, quaternion.dart:65-83
  const double ax = W(a.x), ay = W(a.y), az = W(a.z), aw = W(a.w), bx = W(b.x), by = W(b.y), bz = W(b.z), bw = W(b.w);
  q4 t;
  t.x = (float)(ax * bw + aw * bx + ay * bz - az * by);
  t.y = (float)(ay * bw + aw * by + az * bx - ax * bz);
  t.z = (float)(az * bw + aw * bz + ax * by - ay * bx);
  t.w = (float)(aw * bw - ax * bx - ay * by - az * bz);
  return t;
}
HD f3 to_local_point(const f3& pos, const q4& q, const f3& wp) { return qrot(qconj(q), vsub(wp, pos)); }     // transform.dart:43
HD f3 to_world_point(const f3& pos, const q4& q, const f3& lp) { return vadd(qrot(q, lp), pos); }            // transform.dart:52

HD f3 mrow_mul(const float4& r0, const float4& r1, const float4& r2, const f3& v) {                          // mat3.dart:9
  double x = W(v.x), y = W(v.y), z = W(v.z);
  return mk3(W(r0.x) * x + W(r0.y) * y + W(r0.z) * z, W(r1.x) * x + W(r1.y) * y + W(r1.z) * z,
             W(r2.x) * x + W(r2.y) * y + W(r2.z) * z);
}

// Body.updateInertiaWorld (rigid_body.dart:450-466): R * diag(I) * R^T with float rounding after each stage
HD void inertia_world(const q4& q, const f3& I, float4& r0, float4& r1, float4& r2) {
  double x = W(q.x), y = W(q.y), z = W(q.z), w = W(q.w);
  double x2 = x + x, y2 = y + y, z2 = z + z;
  double xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2;
  double wx = w * x2, wy = w * y2, wz = w * z2;
  float m[9];  // setRotationFromQuaternion, mat3.dart:22
  m[0] = (float)(1 - (yy + zz)); m[1] = (float)(xy - wz);       m[2] = (float)(xz + wy);
  m[3] = (float)(xy + wz);       m[4] = (float)(1 - (xx + zz)); m[5] = (float)(yz - wx);
  m[6] = (float)(xz - wy);       m[7] = (float)(yz + wx);       m[8] = (float)(1 - (xx + yy));
  float t[9];  // m2 = transpose(m1)
  t[0] = m[0]; t[1] = m[3]; t[2] = m[6]; t[3] = m[1]; t[4] = m[4]; t[5] = m[7]; t[6] = m[2]; t[7] = m[5]; t[8] = m[8];
  float s[9];  // vscale: scale columns, mat3.dart:100
  for (int i = 0; i < 3; i++) {
    s[3 * i + 0] = (float)(W(I.x) * W(m[3 * i + 0]));
    s[3 * i + 1] = (float)(W(I.y) * W(m[3 * i + 1]));
    s[3 * i + 2] = (float)(W(I.z) * W(m[3 * i + 2]));
  }
  float o[9];  // multiply2, mat3.dart:57
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++)
      o[3 * r + c] = (float)(W(s[3 * r]) * W(t[c]) + W(s[3 * r + 1]) * W(t[3 + c]) + W(s[3 * r + 2]) * W(t[6 + c]));
  r0 = make_float4(o[0], o[1], o[2], 0.f);
  r1 = make_float4(o[3], o[4], o[5], 0.f);
  r2 = make_float4(o[6], o[7], o[8], 0.f);
}
