/*
 * oracle_math.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Float32-stored / float64-computed vector math exactly as the reference evaluates it on the Dart VM:
 * every vector lives in a Float32List (package:vector_math, see lib/math/vec3.dart:2), every Dart
 * expression is evaluated in double without FMA, every store rounds to float.
 * Compile with -ffp-contract=off.
 *
 * PARITY UNPINNED: the reference has no tests / golden vectors and cannot run in this container
 * (no Dart SDK), so these restatements are pinned only by source-derived known answers
 * (tests/test_oracle_kat.py).
 *
 * Third-party arithmetic restated here: package `vector_math` ^2.1.4 (pubspec.yaml:14, un-vendored,
 * no lock file): Vector3.dot / length2 / length / normalize / negate / distanceTo, Quaternion.normalize /
 * conjugate, Matrix3.transpose — sums are accumulated left to right, normalize multiplies by 1/len and
 * is a no-op on zero length.
 */
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

struct V3 { float x, y, z; };
struct Q4 { float x, y, z, w; };
struct M3 { float e[9]; };  // used row-major by the reference (lib/math/mat3.dart:11-19)

static inline double D(float f) { return (double)f; }
static inline V3 v3(double x, double y, double z) { return V3{(float)x, (float)y, (float)z}; }

// lib/math/vec3.dart:18-23 (target = vector + this)
static inline V3 add(const V3& a, const V3& b) { return v3(D(a.x) + D(b.x), D(a.y) + D(b.y), D(a.z) + D(b.z)); }
// lib/math/vec3.dart:26-31 (target = this - vector)
static inline V3 sub(const V3& a, const V3& b) { return v3(D(a.x) - D(b.x), D(a.y) - D(b.y), D(a.z) - D(b.z)); }
// lib/math/vec3.dart:50-56 (target = scalar * this)
static inline V3 scale(double s, const V3& a) { return v3(s * D(a.x), s * D(a.y), s * D(a.z)); }
// lib/math/vec3.dart:59-70
static inline V3 cross(const V3& a, const V3& b) {
  return v3(D(a.y) * D(b.z) - D(a.z) * D(b.y), D(a.z) * D(b.x) - D(a.x) * D(b.z), D(a.x) * D(b.y) - D(a.y) * D(b.x));
}
// lib/math/vec3.dart:73-79
static inline V3 mulc(const V3& a, const V3& b) { return v3(D(b.x) * D(a.x), D(b.y) * D(a.y), D(b.z) * D(a.z)); }
// vector_math Vector3.dot: sum = a0*b0; sum += a1*b1; sum += a2*b2
static inline double dot(const V3& a, const V3& b) {
  double s = D(a.x) * D(b.x);
  s += D(a.y) * D(b.y);
  s += D(a.z) * D(b.z);
  return s;
}
static inline double length2(const V3& a) { return dot(a, a); }
static inline double length(const V3& a) { return std::sqrt(length2(a)); }
// vector_math Vector3.normalize(): returns the old length, in place, no-op on zero
static inline double normalize(V3& a) {
  double l = length(a);
  if (l == 0.0) return 0.0;
  double d = 1.0 / l;
  a = v3(D(a.x) * d, D(a.y) * d, D(a.z) * d);
  return l;
}
static inline V3 neg(const V3& a) { return V3{-a.x, -a.y, -a.z}; }
// vector_math Vector3.distanceTo (shadows the extension of vec3.dart:81-86; same value)
static inline double distance_to(const V3& a, const V3& b) {
  double dx = D(a.x) - D(b.x), dy = D(a.y) - D(b.y), dz = D(a.z) - D(b.z);
  return std::sqrt(dx * dx + dy * dy + dz * dz);
}
// lib/math/vec3.dart:42-46
static inline V3 lerp(const V3& a, const V3& b, double t) {
  return v3(D(a.x) + (D(b.x) - D(a.x)) * t, D(a.y) + (D(b.y) - D(a.y)) * t, D(a.z) + (D(b.z) - D(a.z)) * t);
}
// lib/math/vec3.dart:122-137 (extension `unit`)
static inline V3 unit(const V3& a) {
  double n = std::sqrt(D(a.x) * D(a.x) + D(a.y) * D(a.y) + D(a.z) * D(a.z));
  if (n > 0.0) {
    n = 1.0 / n;
    return v3(D(a.x) * n, D(a.y) * n, D(a.z) * n);
  }
  return V3{1, 0, 0};
}
// lib/math/vec3.dart:140-146 (target = this + scalar*vector)
static inline V3 add_scaled(const V3& a, double s, const V3& b) {
  return v3(D(a.x) + s * D(b.x), D(a.y) + s * D(b.y), D(a.z) + s * D(b.z));
}
// lib/math/vec3.dart:148-158
static inline bool almost_equals(const V3& a, const V3& b, double prec = 1e-6) {
  return !(std::fabs(D(a.x) - D(b.x)) > prec || std::fabs(D(a.y) - D(b.y)) > prec || std::fabs(D(a.z) - D(b.z)) > prec);
}
// lib/math/vec3.dart:161-166
static inline bool almost_zero(const V3& a, double prec = 1e-6) {
  return !(std::fabs(D(a.x)) > prec || std::fabs(D(a.y)) > prec || std::fabs(D(a.z)) > prec);
}
// lib/math/vec3.dart:97-117
static inline void tangents(const V3& a, V3& t1, V3& t2) {
  double norm = length(a);
  if (norm > 0.0) {
    double inorm = 1 / norm;
    V3 n = v3(D(a.x) * inorm, D(a.y) * inorm, D(a.z) * inorm);
    if (std::fabs(D(n.x)) < 0.9) t1 = cross(n, V3{1, 0, 0});
    else t1 = cross(n, V3{0, 1, 0});
    t2 = cross(n, t1);
  } else {
    t1 = V3{1, 0, 0};
    t2 = V3{0, 1, 0};
  }
}

// lib/math/quaternion.dart:21-45
static inline V3 qvmult(const Q4& q, const V3& v) {
  double x = D(v.x), y = D(v.y), z = D(v.z);
  double qx = D(q.x), qy = D(q.y), qz = D(q.z), qw = D(q.w);
  double ix = qw * x + qy * z - qz * y;
  double iy = qw * y + qz * x - qx * z;
  double iz = qw * z + qx * y - qy * x;
  double iw = -qx * x - qy * y - qz * z;
  return v3(ix * qw + iw * -qx + iy * -qz - iz * -qy,
            iy * qw + iw * -qy + iz * -qx - ix * -qz,
            iz * qw + iw * -qz + ix * -qy - iy * -qx);
}
// lib/math/quaternion.dart:65-83
static inline Q4 qmul(const Q4& a, const Q4& b) {
  double ax = D(a.x), ay = D(a.y), az = D(a.z), aw = D(a.w);
  double bx = D(b.x), by = D(b.y), bz = D(b.z), bw = D(b.w);
  Q4 t;
  t.x = (float)(ax * bw + aw * bx + ay * bz - az * by);
  t.y = (float)(ay * bw + aw * by + az * bx - ax * bz);
  t.z = (float)(az * bw + aw * bz + ax * by - ay * bx);
  t.w = (float)(aw * bw - ax * bx - ay * by - az * bz);
  return t;
}
static inline Q4 qconj(const Q4& q) { return Q4{-q.x, -q.y, -q.z, q.w}; }
// lib/math/quaternion.dart:191-232, Order.xyz
static inline Q4 q_from_euler_xyz(double x, double y, double z) {
  double c1 = std::cos(x / 2), c2 = std::cos(y / 2), c3 = std::cos(z / 2);
  double s1 = std::sin(x / 2), s2 = std::sin(y / 2), s3 = std::sin(z / 2);
  Q4 q;
  q.x = (float)(s1 * c2 * c3 + c1 * s2 * s3);
  q.y = (float)(c1 * s2 * c3 - s1 * c2 * s3);
  q.z = (float)(c1 * c2 * s3 + s1 * s2 * c3);
  q.w = (float)(c1 * c2 * c3 - s1 * s2 * s3);
  return q;
}
// lib/math/transform.dart:43-50
static inline V3 point_to_local_frame(const V3& pos, const Q4& q, const V3& world_point) {
  V3 r = sub(world_point, pos);
  return qvmult(qconj(q), r);
}
// lib/math/transform.dart:52-57
static inline V3 point_to_world_frame(const V3& pos, const Q4& q, const V3& local_point) {
  return add(qvmult(q, local_point), pos);
}
// lib/math/transform.dart:65-71 (w negated, not a conjugate)
static inline V3 vector_to_local_frame(const Q4& q, const V3& world_vector) {
  Q4 t = q;
  t.w = -t.w;
  return qvmult(t, world_vector);
}

// lib/math/mat3.dart:9-20
static inline V3 mvmult(const M3& m, const V3& v) {
  const float* e = m.e;
  double x = D(v.x), y = D(v.y), z = D(v.z);
  return v3(D(e[0]) * x + D(e[1]) * y + D(e[2]) * z, D(e[3]) * x + D(e[4]) * y + D(e[5]) * z,
            D(e[6]) * x + D(e[7]) * y + D(e[8]) * z);
}
// lib/math/mat3.dart:22-54
static inline M3 m_from_quat(const Q4& q) {
  double x = D(q.x), y = D(q.y), z = D(q.z), w = D(q.w);
  double x2 = x + x, y2 = y + y, z2 = z + z;
  double xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2;
  double wx = w * x2, wy = w * y2, wz = w * z2;
  M3 m;
  m.e[0] = (float)(1 - (yy + zz));
  m.e[1] = (float)(xy - wz);
  m.e[2] = (float)(xz + wy);
  m.e[3] = (float)(xy + wz);
  m.e[4] = (float)(1 - (xx + zz));
  m.e[5] = (float)(yz - wx);
  m.e[6] = (float)(xz - wy);
  m.e[7] = (float)(yz + wx);
  m.e[8] = (float)(1 - (xx + yy));
  return m;
}
static inline M3 m_transpose(const M3& a) {
  M3 t = a;
  t.e[1] = a.e[3]; t.e[3] = a.e[1];
  t.e[2] = a.e[6]; t.e[6] = a.e[2];
  t.e[5] = a.e[7]; t.e[7] = a.e[5];
  return t;
}
// lib/math/mat3.dart:57-97 (this * matrix)
static inline M3 m_mul(const M3& A, const M3& B) {
  M3 T;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++)
      T.e[3 * r + c] = (float)(D(A.e[3 * r + 0]) * D(B.e[0 + c]) + D(A.e[3 * r + 1]) * D(B.e[3 + c]) +
                               D(A.e[3 * r + 2]) * D(B.e[6 + c]));
  return T;
}
// lib/math/mat3.dart:100-109 (scale each column)
static inline M3 m_vscale(const M3& a, const V3& v) {
  M3 t;
  for (int i = 0; i < 3; i++) {
    t.e[3 * i + 0] = (float)(D(v.x) * D(a.e[3 * i + 0]));
    t.e[3 * i + 1] = (float)(D(v.y) * D(a.e[3 * i + 1]));
    t.e[3 * i + 2] = (float)(D(v.z) * D(a.e[3 * i + 2]));
  }
  return t;
}

}  // namespace orc
