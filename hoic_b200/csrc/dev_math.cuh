// dev_math.cuh -- float32 vector / quaternion / spatial-algebra helpers and warp-level linear algebra.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "dev_model.h"

#define FULLMASK 0xffffffffu
#define F_MINVAL 1e-15f
#define F_MAXVAL 1e10f
#define F_PI 3.14159265358979323846f

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float norm(V3 a) { return sqrtf(dot(a, a)); }
// mju_normalize3 semantics: returns the norm, tiny vectors become (1,0,0)
__device__ __forceinline__ float normalize(V3& a) {
  float n = norm(a);
  if (n < F_MINVAL) { a = v3(1.f, 0.f, 0.f); }
  else { float s = 1.0f / n; a = a * s; }
  return n;
}
__device__ __forceinline__ float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
__device__ __forceinline__ void setcomp(V3& a, int i, float v) { if (i == 0) a.x = v; else if (i == 1) a.y = v; else a.z = v; }

struct Q4 { float w, x, y, z; };
__device__ __forceinline__ Q4 q4(float w, float x, float y, float z) { Q4 q; q.w = w; q.x = x; q.y = y; q.z = z; return q; }
__device__ __forceinline__ Q4 ldq(const float* p) { return q4(p[0], p[1], p[2], p[3]); }
__device__ __forceinline__ void stq(float* p, Q4 q) { p[0] = q.w; p[1] = q.x; p[2] = q.y; p[3] = q.z; }
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
  return q4(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w);
}
__device__ __forceinline__ Q4 qconj(Q4 a) { return q4(a.w, -a.x, -a.y, -a.z); }
__device__ __forceinline__ Q4 qnormalize(Q4 q) {
  float n = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  if (n < F_MINVAL) return q4(1.f, 0.f, 0.f, 0.f);
  float s = 1.0f / n;
  return q4(q.w * s, q.x * s, q.y * s, q.z * s);
}
// row-major 3x3 in registers
struct M3 { float m[9]; };
__device__ __forceinline__ M3 q2m(Q4 q) {
  M3 r;
  float w = q.w, x = q.x, y = q.y, z = q.z;
  r.m[0] = w * w + x * x - y * y - z * z; r.m[1] = 2 * (x * y - w * z); r.m[2] = 2 * (x * z + w * y);
  r.m[3] = 2 * (x * y + w * z); r.m[4] = w * w - x * x + y * y - z * z; r.m[5] = 2 * (y * z - w * x);
  r.m[6] = 2 * (x * z - w * y); r.m[7] = 2 * (y * z + w * x); r.m[8] = w * w - x * x - y * y + z * z;
  return r;
}
__device__ __forceinline__ M3 ldm(const float* p) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = p[i]; return r; }
__device__ __forceinline__ void stm(float* p, const M3& a) { for (int i = 0; i < 9; i++) p[i] = a.m[i]; }
__device__ __forceinline__ V3 mulv(const M3& a, V3 v) {
  return v3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
            a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}
__device__ __forceinline__ V3 mulTv(const M3& a, V3 v) {
  return v3(a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z,
            a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z);
}
__device__ __forceinline__ M3 mulm(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
  return r;
}
__device__ __forceinline__ M3 mulmT(const M3& a, const M3& b) {  // a * b^T
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * i] * b.m[3 * j] + a.m[3 * i + 1] * b.m[3 * j + 1] + a.m[3 * i + 2] * b.m[3 * j + 2];
  return r;
}
__device__ __forceinline__ V3 col(const M3& a, int k) { return v3(a.m[k], a.m[3 + k], a.m[6 + k]); }
__device__ __forceinline__ V3 rotq(Q4 q, V3 v) { return mulv(q2m(q), v); }
__device__ __forceinline__ Q4 aa2q(V3 axis, float angle) {
  float s, c;
  sincosf(0.5f * angle, &s, &c);
  return q4(c, axis.x * s, axis.y * s, axis.z * s);
}

// ---- warp reductions (all lanes receive the result)
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULLMASK, v, o));
  return v;
}
__device__ __forceinline__ double wsumd(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
__device__ __forceinline__ int wsumi(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
// exclusive prefix sum of small ints across the warp
__device__ __forceinline__ int wexscan(int v, int lane) {
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(FULLMASK, x, o);
    if (lane >= o) x += y;
  }
  return x - v;
}

// ---- spatial algebra: 6-vectors (rot; lin) at the tree root's CoM, 10-vector inertia
__device__ __forceinline__ void inert_com(float* res, const float* inert, const M3& mat, V3 d, float mass) {
  float I[6];  // xx yy zz xy xz yz of mat * diag(inert) * mat^T
  const float* m = mat.m;
  I[0] = m[0] * m[0] * inert[0] + m[1] * m[1] * inert[1] + m[2] * m[2] * inert[2];
  I[1] = m[3] * m[3] * inert[0] + m[4] * m[4] * inert[1] + m[5] * m[5] * inert[2];
  I[2] = m[6] * m[6] * inert[0] + m[7] * m[7] * inert[1] + m[8] * m[8] * inert[2];
  I[3] = m[0] * m[3] * inert[0] + m[1] * m[4] * inert[1] + m[2] * m[5] * inert[2];
  I[4] = m[0] * m[6] * inert[0] + m[1] * m[7] * inert[1] + m[2] * m[8] * inert[2];
  I[5] = m[3] * m[6] * inert[0] + m[4] * m[7] * inert[1] + m[5] * m[8] * inert[2];
  res[0] = I[0] + mass * (d.y * d.y + d.z * d.z);
  res[1] = I[1] + mass * (d.x * d.x + d.z * d.z);
  res[2] = I[2] + mass * (d.x * d.x + d.y * d.y);
  res[3] = I[3] - mass * d.x * d.y;
  res[4] = I[4] - mass * d.x * d.z;
  res[5] = I[5] - mass * d.y * d.z;
  res[6] = mass * d.x; res[7] = mass * d.y; res[8] = mass * d.z; res[9] = mass;
}
__device__ __forceinline__ void mul_inert_vec(float* res, const float* i, const float* v) {
  res[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  res[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  res[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  res[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  res[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  res[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
__device__ __forceinline__ void cross_motion(float* res, const float* vel, const float* v) {
  V3 w = ld3(vel), u = ld3(vel + 3), a = ld3(v), b = ld3(v + 3);
  st3(res, cross(w, a));
  st3(res + 3, cross(w, b) + cross(u, a));
}
__device__ __forceinline__ void cross_force(float* res, const float* vel, const float* f) {
  V3 w = ld3(vel), u = ld3(vel + 3), a = ld3(f), b = ld3(f + 3);
  st3(res, cross(w, a) + cross(u, b));
  st3(res + 3, cross(w, b));
}

// ---- 32x32 SPD solve, one warp, written as rolled loops over 4-wide panels so that the code stays a few hundred
// instructions (the step kernel is instruction-fetch bound: an unrolled register-resident factorisation costs more in
// fetch stalls than its arithmetic, profiles/README.md).  Storage: lower triangle in 4x4 blocks, block (I, J), J <= I, at
// 16 * (I (I + 1) / 2 + J), row-major inside the block: 36 blocks = 576 floats, every row segment a float4.
// Lane i owns row i.  Per panel p every lane factors the 4x4 diagonal block redundantly in registers (no shuffles on
// the pivot chain), solves its own row segment against it, and applies the rank-4 update to the rest of its row.
// The forward substitution rides along; the reciprocals of the diagonal land in the unused upper-triangle slots of
// the diagonal blocks and feed the blocked back-substitution.  Rows / columns >= n must be identity.
#define HB_OFF(I, J) (16 * ((((I) * ((I) + 1)) >> 1) + (J)))
#ifndef CHOL_TRAIL_UNROLL
#define CHOL_TRAIL_UNROLL 2    // A/B knob (tools/build_variant.sh): unroll factor of the rank-4 trailing update
#endif
#define H_PRAGMA_(x) _Pragma(#x)
#define H_UNROLL(n) H_PRAGMA_(unroll n)
#define H_FLOATS 576
__device__ __forceinline__ int hidx(int i, int j) { return HB_OFF(i >> 2, j >> 2) + 4 * (i & 3) + (j & 3); }
__device__ __forceinline__ float rsqrt_nr(float x) {
  float y = rsqrtf(x);
  return y * fmaf(-0.5f * x * y, y, 1.5f);   // one Newton step on the MUFU estimate
}
#ifdef CHOL_REG
// Register-resident variant: lane i loads row i of the lower triangle into 32 registers and the factorisation runs right-looking,
// one column per step, fully unrolled: pivot by shuffle, the column of L goes to shared memory (column-major, columns in groups
// of four of equal padded length: 4 * (32 + 28 + ... + 4) = 576 floats, the space H occupied) and comes back as broadcast float4
// loads for the rank-1 update of the rows in registers.  Forward substitution rides along; the back substitution reads
// L[i][lane] from the lane's own column.  About half the instructions of the blocked form and a third of its dependent chain
// (a pivot step is shuffle + rsqrt + one shared-memory round trip + one FMA).  A/B knob only (-DCHOL_REG), MEASURED NO BETTER: the
// straight-line code is 63 KB (the blocked loops: 6.6 KB) and is fetched once per call, so the factorisation + substitutions take
// 13.5 k instead of 15.5 k cycles per Newton iteration while every other phase of the iteration slows down by 5-10 % (instruction
// cache); box, 4 096 envs: 4.274 vs 4.249 ms per step, 16 384 envs: 13.54 vs 13.04 ms.  All GPU parity tests pass with it.
#define HC_GROUP(g) (128 * (g) - 8 * (g) * ((g) - 1))                         // first float of column group g (columns 4g .. 4g+3)
#define HC_COL(k) (HC_GROUP((k) >> 2) + ((k) & 3) * (32 - 4 * ((k) >> 2)))    // column k, entry for row 4 * (k >> 2)
static __device__ __noinline__ bool warp_chol_solve(float* A, float b, float* xout, int lane) {
  const int I = lane >> 2, r4 = lane & 3;
  float a[32];
#pragma unroll
  for (int J = 0; J < 8; J++) {
    const float4 v = *reinterpret_cast<const float4*>(A + HB_OFF(I, min(J, I)) + 4 * r4);
    const bool in = J <= I;
    a[4 * J] = in ? v.x : 0.f; a[4 * J + 1] = in ? v.y : 0.f; a[4 * J + 2] = in ? v.z : 0.f; a[4 * J + 3] = in ? v.w : 0.f;
  }
  __syncwarp();                     // every row is in registers: the columns of L may overwrite H
  float y = b, rinv = 1.f;
  bool ok = true;
  const int g_l = lane >> 2;
  float* mycol = A + HC_GROUP(g_l) + (lane & 3) * (32 - 4 * g_l) - 4 * g_l;   // + i = L[i][lane]
#pragma unroll
  for (int k = 0; k < 32; k++) {
    const float d = __shfl_sync(FULLMASK, a[k], k);
    ok = ok && d > F_MINVAL;
    const float r = rsqrt_nr(fmaxf(d, F_MINVAL));
    const float l = a[k] * r;                                   // lane i >= k: L[i][k]
    rinv = lane == k ? r : rinv;
    const float yk = __shfl_sync(FULLMASK, y, k) * r;           // forward substitution: y_k / L_kk
    y = lane == k ? yk : (lane > k ? fmaf(-l, yk, y) : y);
    if (k < 31) {
      if (lane >= 4 * (k >> 2)) A[HC_COL(k) + lane - 4 * (k >> 2)] = l;
      __syncwarp();
      const float4* col = reinterpret_cast<const float4*>(A + HC_COL(k));
#pragma unroll
      for (int m = (k + 1) >> 2; m < 8; m++) {
        const float4 q = col[m - (k >> 2)];
        if (4 * m > k) a[4 * m] = fmaf(-l, q.x, a[4 * m]);
        if (4 * m + 1 > k) a[4 * m + 1] = fmaf(-l, q.y, a[4 * m + 1]);
        if (4 * m + 2 > k) a[4 * m + 2] = fmaf(-l, q.z, a[4 * m + 2]);
        a[4 * m + 3] = fmaf(-l, q.w, a[4 * m + 3]);
      }
    }
  }
  __syncwarp();
  // backward: L^T x = y, one unknown per step from the bottom; lane k < i takes L[i][k] from its own column
  float x = 0.f;
#pragma unroll
  for (int i = 31; i >= 0; i--) {
    const float xi = __shfl_sync(FULLMASK, y * rinv, i);
    x = lane == i ? xi : x;
    if (lane < i) y = fmaf(-mycol[i], xi, y);
  }
  *xout = x;
  return ok;
}
#else
static __device__ __noinline__ bool warp_chol_solve(float* A, float b, float* xout, int lane) {
  // Branch-free on purpose: lanes outside the active part of a panel read clamped (valid, irrelevant) addresses and
  // only their stores are predicated, so the warp never diverges and nothing waits on a reconvergence barrier.
  const int I = lane >> 2, r = lane & 3;
  float yb = b;
  bool ok = true;
#pragma unroll 1
  for (int p = 0; p < 8; p++) {
    const float4* Dp = reinterpret_cast<const float4*>(A + HB_OFF(p, p));
    const float4 r0 = Dp[0], r1 = Dp[1], r2 = Dp[2], r3 = Dp[3];
    float4* own = reinterpret_cast<float4*>(A + HB_OFF(max(I, p), p) + 4 * r);
    const float4 a = *own;
    const float z0 = __shfl_sync(FULLMASK, yb, 4 * p), z1 = __shfl_sync(FULLMASK, yb, 4 * p + 1),
                z2 = __shfl_sync(FULLMASK, yb, 4 * p + 2), z3 = __shfl_sync(FULLMASK, yb, 4 * p + 3);
    const float d0 = r0.x;
    const float i0 = rsqrt_nr(fmaxf(d0, F_MINVAL));
    const float l10 = r1.x * i0, l20 = r2.x * i0, l30 = r3.x * i0;
    const float d1 = fmaf(-l10, l10, r1.y);
    const float i1 = rsqrt_nr(fmaxf(d1, F_MINVAL));
    const float l21 = fmaf(-l20, l10, r2.y) * i1, l31 = fmaf(-l30, l10, r3.y) * i1;
    const float d2 = fmaf(-l21, l21, fmaf(-l20, l20, r2.z));
    const float i2 = rsqrt_nr(fmaxf(d2, F_MINVAL));
    const float l32 = fmaf(-l31, l21, fmaf(-l30, l20, r3.z)) * i2;
    const float d3 = fmaf(-l32, l32, fmaf(-l31, l31, fmaf(-l30, l30, r3.w)));
    const float i3 = rsqrt_nr(fmaxf(d3, F_MINVAL));
    ok = ok && fminf(fminf(d0, d1), fminf(d2, d3)) > F_MINVAL;
    const float t0 = z0 * i0, t1 = fmaf(-l10, t0, z1) * i1, t2 = fmaf(-l21, t1, fmaf(-l20, t0, z2)) * i2,
                t3 = fmaf(-l32, t2, fmaf(-l31, t1, fmaf(-l30, t0, z3))) * i3;
    const float x0 = a.x * i0;
    const float x1 = fmaf(-x0, l10, a.y) * i1;
    const float x2 = fmaf(-x1, l21, fmaf(-x0, l20, a.z)) * i2;
    const float x3 = fmaf(-x2, l32, fmaf(-x1, l31, fmaf(-x0, l30, a.w))) * i3;
    const bool inblk = I == p, below = I > p;
    // rows of the diagonal block: entries right of the diagonal carry 1 / L_cc for the back-substitution
    const float s1 = inblk && r == 0 ? i1 : x1;
    const float s2 = inblk && r < 2 ? (r == 0 ? i2 : i0) : x2;
    const float s3 = inblk && r < 3 ? (r == 0 ? i3 : 0.f) : x3;
    const float tr = r == 0 ? t0 : (r == 1 ? t1 : (r == 2 ? t2 : t3));
    const float yd = fmaf(-x3, t3, fmaf(-x2, t2, fmaf(-x1, t1, fmaf(-x0, t0, yb))));
    yb = inblk ? tr : (below ? yd : yb);
    if (I >= p) *own = make_float4(x0, s1, s2, s3);
    __syncwarp();
    // rank-4 update of the row segments right of the panel: A[i][4 Jb + c] -= sum_k L[i][4p + k] L[4 Jb + c][4p + k]
H_UNROLL(CHOL_TRAIL_UNROLL)
    for (int Jb = p + 1; Jb < 8; Jb++) {
      const float4* Lr = reinterpret_cast<const float4*>(A + HB_OFF(Jb, p));
      float4* o = reinterpret_cast<float4*>(A + HB_OFF(I, min(Jb, I)) + 4 * r);
      float4 v = *o;
      const float4 q0 = Lr[0], q1 = Lr[1], q2 = Lr[2], q3 = Lr[3];
      v.x = fmaf(-x3, q0.w, fmaf(-x2, q0.z, fmaf(-x1, q0.y, fmaf(-x0, q0.x, v.x))));
      v.y = fmaf(-x3, q1.w, fmaf(-x2, q1.z, fmaf(-x1, q1.y, fmaf(-x0, q1.x, v.y))));
      v.z = fmaf(-x3, q2.w, fmaf(-x2, q2.z, fmaf(-x1, q2.y, fmaf(-x0, q2.x, v.z))));
      v.w = fmaf(-x3, q3.w, fmaf(-x2, q3.z, fmaf(-x1, q3.y, fmaf(-x0, q3.x, v.w))));
      if (Jb <= I) *o = v;
    }
    __syncwarp();
  }
  // backward: L^T x = y, panel by panel from the bottom
  float xb = yb;
#pragma unroll 1
  for (int p = 7; p >= 0; p--) {
    const float4* Dp = reinterpret_cast<const float4*>(A + HB_OFF(p, p));
    const float4 r0 = Dp[0], r1 = Dp[1], r2 = Dp[2], r3 = Dp[3];   // r0 = (L00, 1/L11, 1/L22, 1/L33), r1 = (L10, L11, 1/L00, 0)
    const float* cb = A + HB_OFF(p, min(I, p)) + r;               // column r of block (p, I)
    const float c0 = cb[0], c1 = cb[4], c2 = cb[8], c3 = cb[12];
    const float z0 = __shfl_sync(FULLMASK, xb, 4 * p), z1 = __shfl_sync(FULLMASK, xb, 4 * p + 1),
                z2 = __shfl_sync(FULLMASK, xb, 4 * p + 2), z3 = __shfl_sync(FULLMASK, xb, 4 * p + 3);
    const float s3 = z3 * r0.w;
    const float s2 = fmaf(-r3.z, s3, z2) * r0.z;
    const float s1 = fmaf(-r3.y, s3, fmaf(-r2.y, s2, z1)) * r0.y;
    const float s0 = fmaf(-r3.x, s3, fmaf(-r2.x, s2, fmaf(-r1.x, s1, z0))) * r1.z;
    const float sr = r == 0 ? s0 : (r == 1 ? s1 : (r == 2 ? s2 : s3));
    const float xu = fmaf(-c3, s3, fmaf(-c2, s2, fmaf(-c1, s1, fmaf(-c0, s0, xb))));
    xb = I == p ? sr : (I < p ? xu : xb);
  }
  *xout = xb;
  return ok;
}
#endif

// ---- tree-sparse SPD solve (mj_factorM / mj_solveLD of MuJoCo: M = L' D L over the dof tree), one warp.
// Solves (M + diag(diag_add)) x = b where M is the joint-space inertia stored by (dof, ancestor depth) in S_MS:
// entry d of row i is M[i, a_i(d)], a_i(d) = ancestor of dof i at depth d, entry depth(i) the diagonal.  Lane = dof.
// Levels are eliminated deepest first; inside a chain a dof has one descendant per level, so the right-looking update
// of a row is one scaled subtraction per level.  The only fan-out of the supported topology (the hub: the palm's last
// dof with the finger chains below it) is handled as a block: the trunk rows receive the Schur complement of all
// branch dofs at once before the hub level is eliminated.  All level loops are static, the row lives in registers.
// Scratch (in S_H): scaled rows L[k][d] = M'[k, a_k(d)] / D_k, pivots D, published solution entries, trunk block.
#define T_MAXD 10
#define T_HUBD 5     // depth of the hub dof when there is one (compile-time so that the block step is emitted once)
#if H_KERNEL == HK_P1 || H_KERNEL == HK_P12
#define TS_LS (S_UNION + UN_TS)     // k_p1 holds no Hessian: the factor scratch follows the branch products in the union
#else
#define TS_LS (S_H)                 // 32 x MS_STRIDE scaled rows
#endif
#define TS_D (TS_LS + 32 * MS_STRIDE)
#define TS_X (TS_D + 32)
#define TS_T (TS_X + 32)            // 21 trunk-block sums (TS_T + 21 <= TS_LS + 480)
#define TS_P (S_UNION)              // 32 x 21 products of the branch dofs (default scratch: callers run while the union region is free)
__device__ __forceinline__ float rcp_nr(float x) {
  const float y = 1.0f / x;
  return y * fmaf(-x, y, 2.0f);
}
__device__ __forceinline__ int tbyte(unsigned w0, unsigned w1, unsigned w2, int i) {   // i is a compile-time constant
  const unsigned w = i < 4 ? w0 : (i < 8 ? w1 : w2);
  return (int)((w >> (8 * (i & 3))) & 0xffu);
}
__device__ __forceinline__ void ld_row12(float (&r)[12], const float* p) {
  const float4 v0 = *reinterpret_cast<const float4*>(p), v1 = *reinterpret_cast<const float4*>(p + 4), v2 = *reinterpret_cast<const float4*>(p + 8);
  r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
  r[8] = v2.x; r[9] = v2.y; r[10] = v2.z; r[11] = v2.w;
}
// `rows` = row storage to factorise (S_MS layout; may alias TS_LS: every lane reads only its own row before anything is
// written), `tsp` = 32 x 21 floats of scratch for the hub block.
static __device__ __noinline__ bool tree_solve_at(float* s, const DevModel& m, int lane, const float* rows, float* tsp, float diag_add, float b,
                                           float* xout) {
  const int dep = lane < m.nv ? m.dof_depth[lane] : -1;
  const int hd = m.hub_depth, maxd = m.tree_maxdepth;
  const bool trunk = lane < m.nv && m.dof_is_trunk[lane];
  const bool branch = (m.branch_mask >> lane) & 1u;
  const int blo = m.branch_lo, bn = m.branch_n;
  const unsigned a0 = m.dof_anc_at[lane][0], a1 = m.dof_anc_at[lane][1], a2 = m.dof_anc_at[lane][2];
  const unsigned c0 = m.dof_desc_at[lane][0], c1 = m.dof_desc_at[lane][1], c2 = m.dof_desc_at[lane][2];
  float* Ls = s + TS_LS;
  float r[12];
  ld_row12(r, rows + lane * MS_STRIDE);
#pragma unroll
  for (int d = 0; d < T_MAXD; d++) r[d] += d == dep ? diag_add : 0.f;
  __syncwarp();
  bool ok = true;
  float dk = 1.f, dinv = 1.f;
  // ---- factorisation, deepest level first.  Entries of r beyond the lane's depth are never used.
#pragma unroll
  for (int l = T_MAXD - 1; l >= 1; l--) {
    if (l > maxd) continue;
    if (l == T_HUBD && hd == T_HUBD) {
      // Schur complement of the (already eliminated) branch dofs on the trunk block: each branch lane forms its 21
      // products L[k][ti] D_k L[k][td] from registers, 21 lanes add them up, trunk lanes subtract their row.
      if (branch) {
        float w[T_HUBD + 1];
#pragma unroll
        for (int t = 0; t <= T_HUBD; t++) w[t] = r[t] * dk;
        int e = 0;
#pragma unroll
        for (int ti = 0; ti <= T_HUBD; ti++)
#pragma unroll
          for (int td = 0; td <= ti; td++) tsp[lane * 21 + e++] = w[ti] * r[td];
      }
      __syncwarp();
      if (lane < 21) {
        float acc0 = 0.f, acc1 = 0.f;
        const float* P = tsp + blo * 21 + lane;
        int q = 0;
#pragma unroll 4
        for (; q + 1 < bn; q += 2) { acc0 += P[q * 21]; acc1 += P[(q + 1) * 21]; }
        if (q < bn) acc0 += P[q * 21];
        s[TS_T + lane] = acc0 + acc1;
      }
      __syncwarp();
      if (trunk) {
        const float* T = s + TS_T + dep * (dep + 1) / 2;
#pragma unroll
        for (int d = 0; d <= T_HUBD; d++)
          if (d <= dep) r[d] -= T[d];
      }
    }
    if (dep == l) {
      dk = r[l];
      ok = dk > F_MINVAL;
      dinv = rcp_nr(fmaxf(dk, F_MINVAL));
#pragma unroll
      for (int d = 0; d < l; d++) r[d] *= dinv;
      float4* o = reinterpret_cast<float4*>(Ls + lane * MS_STRIDE);
      o[0] = make_float4(r[0], r[1], r[2], r[3]);
      if (l > 4) o[1] = make_float4(r[4], r[5], r[6], r[7]);
      if (l > 8) o[2] = make_float4(r[8], r[9], r[10], r[11]);
      s[TS_D + lane] = dk;
    }
    __syncwarp();
    const int k = tbyte(c0, c1, c2, l);
    if (k != 0xff) {
      float kr[12];
      const float* kp = Ls + k * MS_STRIDE;
      {
        const float4 v0 = *reinterpret_cast<const float4*>(kp);
        kr[0] = v0.x; kr[1] = v0.y; kr[2] = v0.z; kr[3] = v0.w;
        if (l > 4) { const float4 v1 = *reinterpret_cast<const float4*>(kp + 4); kr[4] = v1.x; kr[5] = v1.y; kr[6] = v1.z; kr[7] = v1.w; }
        if (l > 8) { const float4 v2 = *reinterpret_cast<const float4*>(kp + 8); kr[8] = v2.x; kr[9] = v2.y; kr[10] = v2.z; kr[11] = v2.w; }
      }
      const float t = kp[dep] * s[TS_D + k];
#pragma unroll
      for (int d = 0; d < l; d++) r[d] = fmaf(-t, kr[d], r[d]);
    }
  }
  if (dep == 0) {
    ok = r[0] > F_MINVAL;
    dinv = rcp_nr(fmaxf(r[0], F_MINVAL));
  }
  // ---- x <- inv(L') x, deepest level first
  float x = b;
#pragma unroll
  for (int l = T_MAXD - 1; l >= 1; l--) {
    if (l == T_HUBD) {   // static; without a hub there are no branch lanes and the correction is zero
      float corr = 0.f;
#pragma unroll
      for (int t = 0; t <= T_HUBD; t++) {
        const float q = wsum(branch ? r[t] * x : 0.f);
        if (t == dep) corr = q;
      }
      if (trunk) x -= corr;
    }
    if (l > maxd) continue;
    if (dep == l) s[TS_X + lane] = x;
    __syncwarp();
    const int k = tbyte(c0, c1, c2, l);
    if (k != 0xff) x = fmaf(-Ls[k * MS_STRIDE + dep], s[TS_X + k], x);
  }
  // ---- x <- inv(D) x ; x <- inv(L) x, root level first
  x *= dinv;
#pragma unroll
  for (int l = 0; l < T_MAXD - 1; l++) {
    if (l >= maxd) continue;
    if (dep == l) s[TS_X + lane] = x;
    __syncwarp();
    if (dep > l) x = fmaf(-r[l], s[TS_X + tbyte(a0, a1, a2, l)], x);
  }
  __syncwarp();
  *xout = x;
  return __all_sync(FULLMASK, ok);
}
__device__ __forceinline__ bool tree_solve(float* s, const DevModel& m, int lane, float diag_add, float b, float* xout) {
  return tree_solve_at(s, m, lane, s + S_MS, s + TS_P, diag_add, b, xout);
}
