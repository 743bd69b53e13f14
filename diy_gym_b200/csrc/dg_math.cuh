// dg_math.cuh - fp32 3-vector / 3x3 / quaternion helpers for the per-environment device code.
// Compiles for sm_100a with nvcc and (for the test-only host emulation under tests/emul) with g++.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DG_HD __host__ __device__ __forceinline__
#define DG_FN __host__ __device__ inline
#define DG_NOINLINE __noinline__
#else
#define DG_NOINLINE
#define DG_HD inline
#define DG_FN inline
#endif

namespace dg {

constexpr float kPi = 3.14159265358979323846f;

DG_HD void v_set(float* o, float x, float y, float z) { o[0] = x; o[1] = y; o[2] = z; }
DG_HD void v_cpy(float* o, const float* a) { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
DG_HD void v_add(float* o, const float* a, const float* b) { o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; }
DG_HD void v_sub(float* o, const float* a, const float* b) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
DG_HD void v_scale(float* o, const float* a, float s) { o[0] = a[0] * s; o[1] = a[1] * s; o[2] = a[2] * s; }
DG_HD void v_madd(float* o, const float* a, float s) { o[0] = fmaf(a[0], s, o[0]); o[1] = fmaf(a[1], s, o[1]); o[2] = fmaf(a[2], s, o[2]); }
DG_HD float v_dot(const float* a, const float* b) { return fmaf(a[0], b[0], fmaf(a[1], b[1], a[2] * b[2])); }
DG_HD float v_len(const float* a) { return sqrtf(v_dot(a, a)); }
DG_HD void v_cross(float* o, const float* a, const float* b) {
  float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
// 3x3 row-major
DG_HD void m_vec(float* o, const float* m, const float* v) {
  float x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
DG_HD void mT_vec(float* o, const float* m, const float* v) {
  float x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2], z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
DG_HD void m_mul(float* o, const float* a, const float* b) {  // o = a b   (o may alias neither)
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
DG_HD void m_mulT(float* o, const float* a, const float* b) {  // o = a b^T
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) o[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
}
DG_HD void mT_mul(float* o, const float* a, const float* b) {  // o = a^T b
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) o[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
DG_HD void m_cpy(float* o, const float* a) {
#pragma unroll
  for (int i = 0; i < 9; i++) o[i] = a[i];
}
// quaternions xyzw
DG_HD void q_to_mat(float* m, const float* q) {
  float x = q[0], y = q[1], z = q[2], w = q[3];
  float n = 1.0f / sqrtf(x * x + y * y + z * z + w * w);
  x *= n; y *= n; z *= n; w *= n;
  m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - z * w); m[2] = 2 * (x * z + y * w);
  m[3] = 2 * (x * y + z * w); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - x * w);
  m[6] = 2 * (x * z - y * w); m[7] = 2 * (y * z + x * w); m[8] = 1 - 2 * (x * x + y * y);
}
DG_HD void q_mul(float* o, const float* a, const float* b) {
  float x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  float y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  float z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
  float w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
DG_HD void q_norm(float* q) {
  float n = 1.0f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] *= n; q[1] *= n; q[2] *= n; q[3] *= n;
}
DG_HD void mat_to_q(float* q, const float* m) {
  float t = m[0] + m[4] + m[8];
  if (t > 0) { float s = sqrtf(t + 1.0f) * 2; q[3] = 0.25f * s; q[0] = (m[7] - m[5]) / s; q[1] = (m[2] - m[6]) / s; q[2] = (m[3] - m[1]) / s; }
  else if (m[0] > m[4] && m[0] > m[8]) { float s = sqrtf(1.0f + m[0] - m[4] - m[8]) * 2; q[3] = (m[7] - m[5]) / s; q[0] = 0.25f * s; q[1] = (m[1] + m[3]) / s; q[2] = (m[2] + m[6]) / s; }
  else if (m[4] > m[8]) { float s = sqrtf(1.0f + m[4] - m[0] - m[8]) * 2; q[3] = (m[2] - m[6]) / s; q[0] = (m[1] + m[3]) / s; q[1] = 0.25f * s; q[2] = (m[5] + m[7]) / s; }
  else { float s = sqrtf(1.0f + m[8] - m[0] - m[4]) * 2; q[3] = (m[3] - m[1]) / s; q[0] = (m[2] + m[6]) / s; q[1] = (m[5] + m[7]) / s; q[2] = 0.25f * s; }
  q_norm(q);
}
// rotation about unit axis a by angle (Rodrigues), row-major
DG_HD void axis_angle_mat(float* m, const float* a, float ang, bool precise = false) {
  float s, c;
#if defined(__CUDA_ARCH__)
  // SFU sine / cosine after reduction to [-pi, pi] (absolute error ~4e-7 there, the size of a few fp32 ulps of the
  // rotation entries); the libm-accurate sincosf costs ~10x the instructions in the FK / IK inner loops.  `precise`:
  // scenes with fixed constraints between models, whose error-reduction term multiplies a link-position error by
  // erp / h ~ 100 / s - the SFU error then shows up as 1e-3 rad/s on light wrist joints (measured, ur_gripper).
  if (precise) sincosf(ang, &s, &c);
  else { ang = fmaf(-6.283185307179586f, rintf(ang * 0.15915494309189535f), ang); __sincosf(ang, &s, &c); }
#else
  (void)precise;
  sincosf(ang, &s, &c);
#endif
  float t = 1 - c, x = a[0], y = a[1], z = a[2];
  m[0] = t * x * x + c; m[1] = t * x * y - s * z; m[2] = t * x * z + s * y;
  m[3] = t * x * y + s * z; m[4] = t * y * y + c; m[5] = t * y * z - s * x;
  m[6] = t * x * z - s * y; m[7] = t * y * z + s * x; m[8] = t * z * z + c;
}
// R = Rz(yaw) Ry(pitch) Rx(roll)  (pybullet getQuaternionFromEuler; diy_gym/model.py:54, misc/respawn.py:39)
DG_HD void q_from_euler(float* q, const float* rpy) {
  float sr, cr, sp, cp, sy, cy;
  sincosf(0.5f * rpy[0], &sr, &cr); sincosf(0.5f * rpy[1], &sp, &cp); sincosf(0.5f * rpy[2], &sy, &cy);
  q[0] = sr * cp * cy - cr * sp * sy; q[1] = cr * sp * cy + sr * cp * sy; q[2] = cr * cp * sy - sr * sp * cy; q[3] = cr * cp * cy + sr * sp * sy;
}
// pybullet getEulerFromQuaternion (diy_gym/addons/sensors/object_state_sensor.py:70)
DG_HD void euler_from_q(float* rpy, const float* q) {
  float x = q[0], y = q[1], z = q[2], w = q[3];
  float sarg = -2 * (x * z - w * y);
  if (sarg <= -0.99999f) { rpy[0] = 0; rpy[1] = -0.5f * kPi; rpy[2] = 2 * atan2f(x, -y); }
  else if (sarg >= 0.99999f) { rpy[0] = 0; rpy[1] = 0.5f * kPi; rpy[2] = 2 * atan2f(-x, y); }
  else {
    rpy[0] = atan2f(2 * (y * z + w * x), w * w - x * x - y * y + z * z);
    rpy[1] = asinf(sarg);
    rpy[2] = atan2f(2 * (x * y + w * z), w * w + x * x - y * y - z * z);
  }
}
DG_HD uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
// counter-based uniform in [0,1) with 24 random bits: identical integers on host and device
DG_HD float urand(uint32_t seed, uint32_t env, uint32_t epoch, uint32_t stream) {
  uint32_t h = hash32(seed ^ hash32(env + 0x9e3779b9U * (epoch + 1)) ^ hash32(stream * 0x85ebca6bU + 0xc2b2ae35U));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

}  // namespace dg
