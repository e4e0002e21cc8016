// dev_collide.cuh -- contact generation, one warp per environment.
// Analytic pairs (plane/capsule/box) run one candidate per lane; convex-mesh pairs run libccd's MPR with the whole
// warp scanning the hull vertices for each support query.  Restates MuJoCo 2.1 mj_collision for this model class
// (SURVEY.md Appendix B.6-B.7); the env reads the result at uhc/envs/ho_im4.py:883-889 (record_contact).
// Contact convention: dist < 0 penetration, pos midway between the surfaces, frame row 0 = normal geom1 -> geom2.
#pragma once
#include <float.h>
#include "dev_math.cuh"

struct RawCon { float dist; V3 pos, n, t; };

__device__ __forceinline__ int sphere_sphere(RawCon& c, float margin, V3 p1, float r1, V3 p2, float r2) {
  V3 d = p2 - p1;
  float cd = norm(d);
  if (cd - r1 - r2 > margin) return 0;
  c.dist = cd - r1 - r2;
  c.n = cd < F_MINVAL ? v3(1.f, 0.f, 0.f) : d * (1.0f / cd);
  c.pos = p1 + c.n * (r1 + 0.5f * c.dist);
  c.t = v3(0, 0, 0);
  return 1;
}
__device__ __forceinline__ int plane_sphere(RawCon& c, float margin, V3 pp, V3 pn, V3 sp, float r) {
  float dist = dot(sp - pp, pn) - r;
  if (dist > margin) return 0;
  c.dist = dist; c.n = pn; c.pos = sp - pn * (r + 0.5f * dist); c.t = v3(0, 0, 0);
  return 1;
}
static __device__ int plane_capsule(RawCon* c, float margin, V3 pp, const M3& pm, V3 cp, const M3& cm, const float* size) {
  V3 pn = col(pm, 2), ax = col(cm, 2);
  int n = 0;
  n += plane_sphere(c[n], margin, pp, pn, cp + ax * size[1], size[0]);
  n += plane_sphere(c[n], margin, pp, pn, cp - ax * size[1], size[0]);
  for (int i = 0; i < n; i++) c[i].t = ax;
  return n;
}
static __device__ int plane_box(RawCon* c, float margin, V3 pp, const M3& pm, V3 bp, const M3& bm, const float* size) {
  V3 pn = col(pm, 2);
  float dist = dot(bp - pp, pn);
  int n = 0;
  for (int i = 0; i < 8; i++) {
    V3 v = v3((i & 1) ? size[0] : -size[0], (i & 2) ? size[1] : -size[1], (i & 4) ? size[2] : -size[2]);
    V3 cw = mulv(bm, v);
    float ld = dot(pn, cw);
    if (dist + ld > margin || ld > 0) continue;
    c[n].dist = dist + ld;
    c[n].pos = cw + bp - pn * (0.5f * c[n].dist);
    c[n].n = pn; c[n].t = v3(0, 0, 0);
    if (++n >= 4) break;
  }
  return n;
}
static __device__ int capsule_capsule(RawCon* c, float margin, V3 p1, const M3& m1, const float* s1, V3 p2, const M3& m2, const float* s2) {
  V3 a1 = col(m1, 2), a2 = col(m2, 2), dif = p1 - p2;
  float ma = dot(a1, a1), mb = -dot(a1, a2), mc = dot(a2, a2), u = -dot(a1, dif), v = dot(a2, dif);
  float det = ma * mc - mb * mb, x1, x2;
  if (fabsf(det) >= F_MINVAL) {
    x1 = (mc * u - mb * v) / det;
    x2 = (ma * v - mb * u) / det;
    if (x1 > s1[1]) { x1 = s1[1]; x2 = (v - mb * s1[1]) / mc; }
    else if (x1 < -s1[1]) { x1 = -s1[1]; x2 = (v + mb * s1[1]) / mc; }
    if (x2 > s2[1]) { x2 = s2[1]; x1 = (u - mb * s2[1]) / ma; }
    else if (x2 < -s2[1]) { x2 = -s2[1]; x1 = (u + mb * s2[1]) / ma; }
    x1 = fminf(fmaxf(x1, -s1[1]), s1[1]);
    return sphere_sphere(c[0], margin, p1 + a1 * x1, s1[0], p2 + a2 * x2, s2[0]);
  }
  int n = 0;
  for (int k = 0; k < 4 && n < 2; k++) {
    if (k < 2) {
      x1 = k == 0 ? s1[1] : -s1[1];
      x2 = (v - mb * x1) / mc;
      if (x2 < -s2[1] || x2 > s2[1]) continue;
    } else {
      x2 = k == 2 ? s2[1] : -s2[1];
      x1 = (u - mb * x2) / ma;
      if (x1 < -s1[1] || x1 > s1[1]) continue;
    }
    n += sphere_sphere(c[n], margin, p1 + a1 * x1, s1[0], p2 + a2 * x2, s2[0]);
  }
  return n;
}

// sphere (centre p, box frame) against the solid box; normal points sphere -> box
static __device__ int sphere_box_local(RawCon& c, float margin, V3 p, float r, const float* s) {
  V3 q = v3(fminf(fmaxf(p.x, -s[0]), s[0]), fminf(fmaxf(p.y, -s[1]), s[1]), fminf(fmaxf(p.z, -s[2]), s[2]));
  V3 d = q - p;
  float d2 = dot(d, d), dist;
  V3 n;
  if (d2 > 0.f) {
    float dd = sqrtf(d2);
    dist = dd - r;
    if (dist > margin) return 0;
    n = d * (1.0f / dd);
  } else {
    int best = 0;
    float bd = FLT_MAX, sign = 1.f;
    for (int i = 0; i < 3; i++) {
      float pi = comp(p, i), dpos = s[i] - pi, dneg = pi + s[i];
      if (dpos < bd) { bd = dpos; best = i; sign = 1.f; }
      if (dneg < bd) { bd = dneg; best = i; sign = -1.f; }
    }
    n = v3(0, 0, 0);
    setcomp(n, best, -sign);
    dist = -bd - r;
  }
  c.dist = dist; c.n = n; c.pos = p + n * (r + 0.5f * dist); c.t = v3(0, 0, 0);
  return 1;
}

// capsule (geom1) vs box (geom2): exact minimiser of the segment-box distance, two contacts at the ends of the
// portion above a face when the closest feature is a face (same construction as the oracle).
static __device__ __noinline__ int capsule_box(RawCon* c, float margin, V3 cp, const M3& cm, const float* cs, V3 bp, const M3& bm, const float* bs) {
  const float h = cs[1], r = cs[0];
  V3 p0 = mulTv(bm, cp - bp), a = mulTv(bm, col(cm, 2));
  float tb[8];
  int nb = 0, n = 0;
  tb[nb++] = -h; tb[nb++] = h;
  for (int i = 0; i < 3; i++) {
    float ai = comp(a, i);
    if (fabsf(ai) > 1e-12f)
      for (int sgn = -1; sgn <= 1; sgn += 2) {
        float t = (sgn * bs[i] - comp(p0, i)) / ai;
        if (t > -h && t < h) tb[nb++] = t;
      }
  }
  for (int i = 1; i < nb; i++) { float x = tb[i]; int j = i - 1; while (j >= 0 && tb[j] > x) { tb[j + 1] = tb[j]; j--; } tb[j + 1] = x; }
  float best_t = 0.f, best_d2 = FLT_MAX, zlo = 0.f, zhi = 0.f;
  bool have_z = false;
  for (int k = 0; k + 1 < nb; k++) {
    float t0 = tb[k], t1 = tb[k + 1], tm = 0.5f * (t0 + t1), A = 0.f, B = 0.f, C = 0.f, t, d2;
    for (int i = 0; i < 3; i++) {
      float pi = comp(p0, i), ai = comp(a, i), x = pi + tm * ai, e;
      if (x > bs[i]) e = pi - bs[i]; else if (x < -bs[i]) e = pi + bs[i]; else continue;
      A += ai * ai; B += ai * e; C += e * e;
    }
    if (A > 0.f) { t = fminf(fmaxf(-B / A, t0), t1); d2 = A * t * t + 2.f * B * t + C; }
    else { t = t0; d2 = C; }
    if (d2 < best_d2) { best_d2 = d2; best_t = t; }
    if (A == 0.f && C == 0.f) { if (!have_z) { zlo = t0; have_z = true; } zhi = t1; }   // this stretch of the axis is inside the box
  }
  // axis through the box interior: every point of the inside stretch is at distance zero; take its midpoint (the entry point
  // lies on a face and whether it counts as inside would be decided by rounding)
  if (have_z) best_t = 0.5f * (zlo + zhi);
  V3 p = p0 + a * best_t;
  int nout = 0, kface = -1;
  for (int i = 0; i < 3; i++)
    if (fabsf(comp(p, i)) > bs[i]) { nout++; kface = i; }
  if (nout == 1) {
    float lo = -h, hi = h;
    for (int i = 0; i < 3; i++) {
      float ai = comp(a, i);
      if (i == kface || fabsf(ai) <= 1e-12f) continue;
      float ta = (-bs[i] - comp(p0, i)) / ai, tc = (bs[i] - comp(p0, i)) / ai;
      if (ta > tc) { float x = ta; ta = tc; tc = x; }
      lo = fmaxf(lo, ta);
      hi = fminf(hi, tc);
    }
    lo = fminf(lo, best_t);
    hi = fmaxf(hi, best_t);
    n += sphere_box_local(c[n], margin, p0 + a * lo, r, bs);
    if (hi - lo > 1e-6f * (h + r)) n += sphere_box_local(c[n], margin, p0 + a * hi, r, bs);
  } else {
    n += sphere_box_local(c[n], margin, p, r, bs);
  }
  for (int i = 0; i < n; i++) { c[i].pos = mulv(bm, c[i].pos) + bp; c[i].n = mulv(bm, c[i].n); }
  return n;
}

// box (geom1) vs box (geom2): 15-axis separating-axis test, reference/incident face clipping or edge-edge point
static __device__ __noinline__ int box_box(RawCon* c, float margin, V3 p1, const M3& R1, const float* A, V3 p2, const M3& R2, const float* B) {
  V3 p = p2 - p1, pp = mulTv(R1, p);
  float R[9], Q[9], s, best = -FLT_MAX;
  V3 nrm = v3(0, 0, 0);
  int code = 0, invert = 0, n = 0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { R[3 * i + j] = dot(col(R1, i), col(R2, j)); Q[3 * i + j] = fabsf(R[3 * i + j]); }
  for (int i = 0; i < 3; i++) {
    s = fabsf(comp(pp, i)) - (A[i] + B[0] * Q[3 * i] + B[1] * Q[3 * i + 1] + B[2] * Q[3 * i + 2]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = 1 + i; invert = comp(pp, i) < 0; nrm = col(R1, i); }
  }
  for (int j = 0; j < 3; j++) {
    float pj = dot(col(R2, j), p);
    s = fabsf(pj) - (A[0] * Q[j] + A[1] * Q[3 + j] + A[2] * Q[6 + j] + B[j]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = 4 + j; invert = pj < 0; nrm = col(R2, j); }
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      float ex = comp(pp, i2) * R[3 * i1 + j] - comp(pp, i1) * R[3 * i2 + j];
      float ra = A[i1] * Q[3 * i2 + j] + A[i2] * Q[3 * i1 + j];
      float rb = B[j1] * Q[3 * i + j2] + B[j2] * Q[3 * i + j1];
      float l2 = R[3 * i1 + j] * R[3 * i1 + j] + R[3 * i2 + j] * R[3 * i2 + j];
      if (l2 < 1e-12f) continue;
      float l = sqrtf(l2);
      s = (fabsf(ex) - (ra + rb)) / l;
      if (s > margin) return 0;
      if (s * 1.05f > best + 1e-12f && s > best) {
        V3 nl = v3(0, 0, 0);
        best = s; code = 7 + 3 * i + j; invert = ex < 0;
        setcomp(nl, i1, -R[3 * i2 + j] / l);
        setcomp(nl, i2, R[3 * i1 + j] / l);
        nrm = mulv(R1, nl);
      }
    }
  if (!code) return 0;
  if (invert) nrm = -nrm;
  if (code > 6) {
    int i = (code - 7) / 3, j = (code - 7) % 3;
    V3 pa = p1, pb = p2, ua = col(R1, i), ub = col(R2, j);
    for (int k = 0; k < 3; k++) {
      V3 ck = col(R1, k);
      if (k != i) pa = pa + ck * ((dot(nrm, ck) > 0 ? 1.f : -1.f) * A[k]);
    }
    for (int k = 0; k < 3; k++) {
      V3 ck = col(R2, k);
      if (k != j) pb = pb + ck * ((dot(nrm, ck) > 0 ? -1.f : 1.f) * B[k]);
    }
    V3 d = pb - pa;
    float uaub = dot(ua, ub), q1 = dot(ua, d), q2 = -dot(ub, d), den = 1.f - uaub * uaub, al = 0.f, be = 0.f;
    if (den > 1e-12f) { al = (q1 + uaub * q2) / den; be = (uaub * q1 + q2) / den; }
    pa = pa + ua * al; pb = pb + ub * be;
    c[0].dist = best; c[0].pos = (pa + pb) * 0.5f; c[0].n = nrm; c[0].t = v3(0, 0, 0);
    return 1;
  }
  const bool first = code <= 3;
  const M3& Ra = first ? R1 : R2;
  const M3& Rb = first ? R2 : R1;
  const V3 pa = first ? p1 : p2, pb = first ? p2 : p1;
  const float* Sa = first ? A : B;
  const float* Sb = first ? B : A;
  const V3 nR = first ? nrm : -nrm;
  V3 nr = mulTv(Rb, nR);
  float anr[3] = {fabsf(nr.x), fabsf(nr.y), fabsf(nr.z)};
  int lanr, a1, a2;
  if (anr[1] > anr[0]) { if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  else { if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
  V3 center = pb - pa + col(Rb, lanr) * (comp(nr, lanr) < 0 ? Sb[lanr] : -Sb[lanr]);
  int codeN = first ? code - 1 : code - 4, code1, code2;
  if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
  V3 ca1 = col(Ra, code1), ca2 = col(Ra, code2), cb1 = col(Rb, a1), cb2 = col(Rb, a2);
  float c1 = dot(center, ca1), c2 = dot(center, ca2);
  float m11 = dot(ca1, cb1), m12 = dot(ca1, cb2), m21 = dot(ca2, cb1), m22 = dot(ca2, cb2);
  float k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
  float bufA[16], bufB[16];
  bufA[0] = c1 - k1 - k3; bufA[1] = c2 - k2 - k4; bufA[2] = c1 - k1 + k3; bufA[3] = c2 - k2 + k4;
  bufA[4] = c1 + k1 + k3; bufA[5] = c2 + k2 + k4; bufA[6] = c1 + k1 - k3; bufA[7] = c2 + k2 - k4;
  float rect[2] = {Sa[code1], Sa[code2]};
  float *qv = bufA, *rv = bufB;
  int nq = 4, nret = 0;
  bool full = false;
  for (int dir = 0; dir <= 1 && !full; dir++)
    for (int sign = -1; sign <= 1 && !full; sign += 2) {
      float *pq = qv, *pr = rv;
      nret = 0;
      for (int i = nq; i > 0; i--) {
        float* nextq = (i > 1) ? pq + 2 : qv;
        bool in0 = sign * pq[dir] < rect[dir], in1 = sign * nextq[dir] < rect[dir];
        if (in0) { pr[0] = pq[0]; pr[1] = pq[1]; pr += 2; nret++; if (nret & 8) { full = true; break; } }
        if (in0 != in1) {
          pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * rect[dir] - pq[dir]);
          pr[dir] = sign * rect[dir];
          pr += 2; nret++;
          if (nret & 8) { full = true; break; }
        }
        pq += 2;
      }
      float* t = qv; qv = rv; rv = t;
      nq = nret;
    }
  if (nret < 1) return 0;
  float det1 = 1.0f / (m11 * m22 - m12 * m21);
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  for (int j = 0; j < nret && n < 8; j++) {
    float kk1 = m22 * (qv[2 * j] - c1) - m12 * (qv[2 * j + 1] - c2);
    float kk2 = -m21 * (qv[2 * j] - c1) + m11 * (qv[2 * j + 1] - c2);
    V3 pt = center + cb1 * kk1 + cb2 * kk2;
    float depth = Sa[codeN] - dot(nR, pt);
    if (-depth > margin) continue;
    c[n].dist = -depth;
    c[n].pos = pa + pt + nR * (0.5f * depth);
    c[n].n = nrm; c[n].t = v3(0, 0, 0);
    n++;
  }
  return n;
}

// unit directions of the mesh 26-DOP used for pair culling (dev_model.h: geom_dop)
#define H_DOP_TABLE { \
    {1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}, \
    {0.70710678f, 0.70710678f, 0.f}, {0.70710678f, -0.70710678f, 0.f}, {0.70710678f, 0.f, 0.70710678f}, {0.70710678f, 0.f, -0.70710678f}, \
    {0.f, 0.70710678f, 0.70710678f}, {0.f, 0.70710678f, -0.70710678f}, \
    {0.57735027f, 0.57735027f, 0.57735027f}, {0.57735027f, 0.57735027f, -0.57735027f}, {0.57735027f, -0.57735027f, 0.57735027f}, \
    {0.57735027f, -0.57735027f, -0.57735027f}}
static __constant__ float c_dop[H_NDOP][3] = H_DOP_TABLE;
static const float h_dop[H_NDOP][3] = H_DOP_TABLE;   // host copy for the model build

// ---------------------------------------------------------------------------------------------- MPR (libccd)
// The two objects of the pair being solved live in shared memory (U_COBJ): mpr_penetration / support_w are out of line and
// would otherwise read them from the local-memory stack.
struct CObj { int type, nvert, ncl, pad; V3 pos; M3 mat; float size[3]; float pad2; const float4* vert; const float4* cl; };
#define U_COBJ (U_AREF)      // 2 x 24 words over the head of the row arrays: build_rows writes those after collide() is done
static_assert(sizeof(CObj) == 96 && U_COBJ * 4 % 8 == 0 && 48 <= H_NROW, "CObj scratch");
struct Supp { V3 v, v1, v2; };
// per-lane view of a mesh for the support function: lane k keeps the AABB of vertex cluster k in registers for the whole
// MPR call, `warm` (uniform) is the last support vertex in local coordinates = a lower bound for the next query
struct MeshLane { float lox, loy, loz, hix, hiy, hiz, rx, ry, rz; V3 warm; int valid; };
struct SuppPt { V3 world, local; };

// order-preserving float -> uint key (for redux.sync, one instruction instead of a 5-step shuffle tree)
__device__ __forceinline__ unsigned fkey(float f) {
  const unsigned u = __float_as_uint(f + 0.0f);   // -0 -> +0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ MeshLane mesh_lane(const CObj& o, int lane) {
  MeshLane ml;
  ml.lox = ml.loy = ml.loz = ml.hix = ml.hiy = ml.hiz = ml.rx = ml.ry = ml.rz = 0.f;
  ml.warm = v3(0, 0, 0);
  ml.valid = 0;
  if (o.type == G_MESH && o.nvert > 0) {
    if (lane < o.ncl) {
      const float4 lo = __ldg(o.cl + 2 * lane), hi = __ldg(o.cl + 2 * lane + 1), rp = __ldg(o.vert + 32 * lane);
      ml.lox = lo.x; ml.loy = lo.y; ml.loz = lo.z; ml.hix = hi.x; ml.hiy = hi.y; ml.hiz = hi.z;
      ml.rx = rp.x; ml.ry = rp.y; ml.rz = rp.z;
    }
    const float4 v0 = __ldg(o.vert);
    ml.warm = v3(v0.x, v0.y, v0.z);
    ml.valid = 1;
  }
  return ml;
}

// support point of a convex object along world direction `dir`; the whole warp calls this with identical
// arguments.  Meshes: clusters of 32 vertices with a local AABB each (model build).  Upper bound of a cluster = AABB
// support, lower bound of the answer = the best projection of a few known vertices; only clusters
// whose upper bound reaches it are scanned, lane-strided, and reduced with redux.sync.  The dot products and the tie
// rule (lowest ORIGINAL index) are those of a serial scan over all vertices.
static __device__ __noinline__ SuppPt support_w(const CObj& o, V3 dirw, int lane, MeshLane ml) {
  V3 d = mulTv(o.mat, dirw), l;
  if (o.type == G_CAPSULE) {
    float nn = norm(d);
    l = nn < F_MINVAL ? v3(0, 0, 0) : d * (o.size[0] / nn);
    l.z += d.z >= 0.f ? o.size[1] : -o.size[1];
  } else if (o.type == G_BOX) {
    l = v3(d.x >= 0.f ? o.size[0] : -o.size[0], d.y >= 0.f ? o.size[1] : -o.size[1], d.z >= 0.f ? o.size[2] : -o.size[2]);
  } else if (o.type == G_SPHERE) {
    float nn = norm(d);
    l = nn < F_MINVAL ? v3(0, 0, 0) : d * (o.size[0] / nn);
  } else {
    if (!ml.valid) ml = mesh_lane(o, lane);
    // lower bound of the answer: the best of the previous support vertex and every cluster's first vertex
    float ub = -FLT_MAX, lb = ml.warm.x * d.x + ml.warm.y * d.y + ml.warm.z * d.z;
    if (lane < o.ncl) {
      ub = (d.x >= 0.f ? ml.hix : ml.lox) * d.x + (d.y >= 0.f ? ml.hiy : ml.loy) * d.y + (d.z >= 0.f ? ml.hiz : ml.loz) * d.z;
      lb = fmaxf(lb, ml.rx * d.x + ml.ry * d.y + ml.rz * d.z);
    }
    const unsigned k0 = __reduce_max_sync(FULLMASK, fkey(lb));
    const float best0 = __uint_as_float((k0 & 0x80000000u) ? (k0 & 0x7fffffffu) : ~k0);
    const float slack = 1e-6f * (fabsf(d.x) + fabsf(d.y) + fabsf(d.z));   // >> rounding of a 3-term dot of |v| < 1 m
    unsigned surv = __ballot_sync(FULLMASK, ub >= best0 - slack);
    float best = -FLT_MAX, bx = 0.f, by = 0.f, bz = 0.f;
    int bi = 0x7fffffff;
    auto scan = [&](int i) {
      if (i < o.nvert) {
        const float4 v = __ldg(o.vert + i);
        const float sc = v.x * d.x + v.y * d.y + v.z * d.z;
        const int oi = __float_as_int(v.w);
        if (sc > best || (sc == best && oi < bi)) { best = sc; bi = oi; bx = v.x; by = v.y; bz = v.z; }
      }
    };
    for (; surv; surv &= surv - 1) scan(32 * (__ffs(surv) - 1) + lane);
    for (int k = 32; k < o.ncl; k++) scan(32 * k + lane);   // meshes beyond 1024 vertices: the tail is scanned unpruned
    const unsigned key = fkey(best);
    const unsigned kmax = __reduce_max_sync(FULLMASK, key);
    const bool top = key == kmax;
    const unsigned imin = __reduce_min_sync(FULLMASK, top ? (unsigned)bi : 0xffffffffu);
    const int src = __ffs(__ballot_sync(FULLMASK, top && (unsigned)bi == imin)) - 1;
    l = v3(__shfl_sync(FULLMASK, bx, src), __shfl_sync(FULLMASK, by, src), __shfl_sync(FULLMASK, bz, src));
  }
  SuppPt r;
  r.local = l;
  r.world = mulv(o.mat, l) + o.pos;
  return r;
}
// o1's mesh view is not cached (no mesh is ever geom1 against a mesh in these models: it would be rebuilt per query)
__device__ __forceinline__ void mpr_support(const CObj& o1, const CObj& o2, V3 dir, Supp& s, int lane, MeshLane& ml2) {
  MeshLane none;
  none.lox = none.loy = none.loz = none.hix = none.hiy = none.hiz = none.rx = none.ry = none.rz = 0.f; none.warm = v3(0, 0, 0); none.valid = 0;
  s.v1 = support_w(o1, dir, lane, none).world;
  const SuppPt p2 = support_w(o2, -dir, lane, ml2);
  ml2.warm = ml2.valid ? p2.local : ml2.warm;
  s.v2 = p2.world;
  s.v = s.v1 - s.v2;
}
// libccd's absolute epsilon tests (ccdIsZero on squared lengths / dot products) are kept at the DOUBLE build's value:
// MuJoCo links libccd in double precision, and with FLT_EPSILON the 'origin on the v0-v1 segment' shortcut of
// discoverPortal would fire for offsets of several millimetres.
#define CCDF_EPS 2.220446049250313e-16f
__device__ __forceinline__ bool ccd_zero(float x) { return fabsf(x) < CCDF_EPS; }
__device__ __forceinline__ bool ccd_eq(float a, float b) {
  float ab = fabsf(a - b);
  if (ab < CCDF_EPS) return true;
  return ab < CCDF_EPS * fmaxf(fabsf(a), fabsf(b));
}
static __device__ float point_seg_dist2(V3 P, V3 x0, V3 b, V3& wit) {
  V3 d = b - x0, a = x0 - P;
  float t = -dot(a, d) / dot(d, d);
  if (t < 0.f || ccd_zero(t)) { wit = x0; V3 e = x0 - P; return dot(e, e); }
  if (t > 1.f || ccd_eq(t, 1.f)) { wit = b; V3 e = b - P; return dot(e, e); }
  wit = x0 + d * t;
  V3 e = wit - P;
  return dot(e, e);
}
static __device__ __noinline__ float point_tri_dist2(V3 P, V3 x0, V3 B, V3 C, V3& wit, bool& inside) {
  V3 d1 = B - x0, d2 = C - x0, a = x0 - P;
  float v = dot(d1, d1), w = dot(d2, d2), p = dot(a, d1), q = dot(a, d2), r = dot(d1, d2);
  float s = (q * r - w * p) / (w * v - r * r), t = (-s * r - q) / w;
  if ((ccd_zero(s) || s > 0.f) && (ccd_eq(s, 1.f) || s < 1.f) && (ccd_zero(t) || t > 0.f) && (ccd_eq(t, 1.f) || t < 1.f) &&
      (ccd_eq(t + s, 1.f) || t + s < 1.f)) {
    wit = x0 + d1 * s + d2 * t;
    V3 e = wit - P;
    inside = true;
    return dot(e, e);
  }
  inside = false;
  V3 w2;
  float dist = point_seg_dist2(P, x0, B, wit), d2_ = point_seg_dist2(P, x0, C, w2);
  if (d2_ < dist) { dist = d2_; wit = w2; }
  d2_ = point_seg_dist2(P, B, C, w2);
  if (d2_ < dist) { dist = d2_; wit = w2; }
  return dist;
}
__device__ __forceinline__ V3 portal_dir(const Supp* s) {
  V3 d = cross(s[2].v - s[1].v, s[3].v - s[1].v);
  normalize(d);
  return d;
}
__device__ __forceinline__ bool portal_reach_tol(const Supp* s, const Supp& v4, V3 dir, float tol) {
  float dv4 = dot(v4.v, dir);
  float d1 = fminf(fminf(dv4 - dot(s[1].v, dir), dv4 - dot(s[2].v, dir)), dv4 - dot(s[3].v, dir));
  return ccd_eq(d1, tol) || d1 < tol;
}
__device__ __forceinline__ void expand_portal(Supp* s, const Supp& v4) {
  V3 c = cross(v4.v, s[0].v);
  if (dot(s[1].v, c) > 0.f) { if (dot(s[2].v, c) > 0.f) s[1] = v4; else s[3] = v4; }
  else { if (dot(s[3].v, c) > 0.f) s[2] = v4; else s[1] = v4; }
}
// returns 0 when intersecting; all lanes run the same scalar logic (uniform), only support scans are split
static __device__ __noinline__ int mpr_penetration(const CObj& o1, const CObj& o2, float tol, int max_iter, float& depth, V3& dir, V3& pos,
                               int& iters, int lane, MeshLane ml2) {
  Supp s[4], v4;
  V3 d;
  float dt;
  int res = 0;
  const V3 origin = v3(0, 0, 0);
  s[0].v1 = o1.pos; s[0].v2 = o2.pos; s[0].v = o1.pos - o2.pos;
  if (ccd_eq(s[0].v.x, 0.f) && ccd_eq(s[0].v.y, 0.f) && ccd_eq(s[0].v.z, 0.f)) s[0].v.x += CCDF_EPS * 10.f;
  d = -s[0].v; normalize(d);
  mpr_support(o1, o2, d, s[1], lane, ml2);
  dt = dot(s[1].v, d);
  if (ccd_zero(dt) || dt < 0.f) return -1;
  d = cross(s[0].v, s[1].v);
  if (ccd_zero(dot(d, d))) {
    res = (ccd_eq(s[1].v.x, 0.f) && ccd_eq(s[1].v.y, 0.f) && ccd_eq(s[1].v.z, 0.f)) ? 1 : 2;
  } else {
    normalize(d);
    mpr_support(o1, o2, d, s[2], lane, ml2);
    dt = dot(s[2].v, d);
    if (ccd_zero(dt) || dt < 0.f) return -1;
    d = cross(s[1].v - s[0].v, s[2].v - s[0].v); normalize(d);
    if (dot(d, s[0].v) > 0.f) { Supp t = s[1]; s[1] = s[2]; s[2] = t; d = -d; }
    int size = 3, guard = 0;
    while (size < 4 && guard++ < 64) {
      bool cont = false;
      mpr_support(o1, o2, d, s[3], lane, ml2);
      iters++;
      dt = dot(s[3].v, d);
      if (ccd_zero(dt) || dt < 0.f) return -1;
      dt = dot(cross(s[1].v, s[3].v), s[0].v);
      if (dt < 0.f && !ccd_zero(dt)) { s[2] = s[3]; cont = true; }
      if (!cont) {
        dt = dot(cross(s[3].v, s[2].v), s[0].v);
        if (dt < 0.f && !ccd_zero(dt)) { s[1] = s[3]; cont = true; }
      }
      if (cont) { d = cross(s[1].v - s[0].v, s[2].v - s[0].v); normalize(d); }
      else size = 4;
    }
    if (size < 4) return -1;
  }
  if (res == 1) { depth = 0.f; dir = v3(0, 0, 0); pos = (s[1].v1 + s[1].v2) * 0.5f; return 0; }
  if (res == 2) { pos = (s[1].v1 + s[1].v2) * 0.5f; depth = norm(s[1].v); dir = s[1].v; normalize(dir); return 0; }
  for (int guard = 0;; guard++) {
    d = portal_dir(s);
    dt = dot(d, s[1].v);
    if (ccd_zero(dt) || dt > 0.f) break;
    mpr_support(o1, o2, d, v4, lane, ml2);
    iters++;
    dt = dot(v4.v, d);
    if (!(ccd_zero(dt) || dt > 0.f) || portal_reach_tol(s, v4, d, tol) || guard > 128) return -1;
    expand_portal(s, v4);
  }
  for (int it = 0;; it++) {
    d = portal_dir(s);
    mpr_support(o1, o2, d, v4, lane, ml2);
    iters++;
    if (portal_reach_tol(s, v4, d, tol) || it > max_iter) {
      V3 wit;
      bool inside;
      depth = sqrtf(point_tri_dist2(origin, s[1].v, s[2].v, s[3].v, wit, inside));
      if (inside) {
        // the closest point is the orthogonal projection of the origin on the portal: its direction IS the portal
        // normal and its length n.v1.  Taking both from the normal (cross product of O(size) edges) instead of from
        // the O(depth) witness vector keeps shallow contacts accurate in float32 (same value in exact arithmetic).
        float dn = dot(d, s[1].v);
        wit = dn < 0.f ? -d : d;
        depth = fabsf(dn);
      } else {
        if (ccd_zero(wit.x) && ccd_zero(wit.y) && ccd_zero(wit.z)) wit = d;
        normalize(wit);
      }
      dir = wit;
      // barycentric position (findPos)
      float b[4], sum;
      b[0] = dot(cross(s[1].v, s[2].v), s[3].v);
      b[1] = dot(cross(s[3].v, s[2].v), s[0].v);
      b[2] = dot(cross(s[0].v, s[1].v), s[3].v);
      b[3] = dot(cross(s[2].v, s[1].v), s[0].v);
      sum = b[0] + b[1] + b[2] + b[3];
      if (ccd_zero(sum) || sum < 0.f) {
        V3 pd = portal_dir(s);
        b[0] = 0.f;
        b[1] = dot(cross(s[2].v, s[3].v), pd);
        b[2] = dot(cross(s[3].v, s[1].v), pd);
        b[3] = dot(cross(s[1].v, s[2].v), pd);
        sum = b[1] + b[2] + b[3];
      }
      float inv = 1.0f / sum;
      V3 p1 = v3(0, 0, 0), p2 = v3(0, 0, 0);
      for (int i = 0; i < 4; i++) { p1 = p1 + s[i].v1 * b[i]; p2 = p2 + s[i].v2 * b[i]; }
      pos = (p1 + p2) * (0.5f * inv);
      return 0;
    }
    expand_portal(s, v4);
  }
}

__device__ __forceinline__ void make_cobj(const float* s, const DevModel& m, int g, CObj& o) {
  o.type = m.geom_type[g];
  o.pad = 0; o.pad2 = 0.f;
  o.pos = ld3(s + S_GPOS + 3 * g);
  o.mat = ldm(s + S_GMAT + 9 * g);
  o.size[0] = m.geom_size[g][0]; o.size[1] = m.geom_size[g][1]; o.size[2] = m.geom_size[g][2];
  o.nvert = m.geom_vertnum[g];
  o.vert = m.mesh_vert + (m.geom_vertadr[g] >= 0 ? m.geom_vertadr[g] : 0);
  o.ncl = m.geom_ncl[g];
  o.cl = m.mesh_cl + m.geom_cladr[g];   // float4 units
}

// mju_makeFrame
__device__ __forceinline__ void make_frame(V3& x, V3& y, V3& z) {
  normalize(x);
  if (norm(y) < 0.5f) { y = (x.y < 0.5f && x.y > -0.5f) ? v3(0.f, 1.f, 0.f) : v3(0.f, 0.f, 1.f); }
  y = y - x * dot(x, y);
  normalize(y);
  z = cross(x, y);
}

__device__ __forceinline__ void write_contact(float* s, int idx, const RawCon& rc, int cand) {
  float* o = s + U_CON + 16 * idx;
  V3 x = rc.n, y = rc.t, z;
  make_frame(x, y, z);
  st3(o, rc.pos); st3(o + 3, x); st3(o + 6, y); st3(o + 9, z);
  o[12] = rc.dist;
  o[13] = __int_as_float(cand);
}

// Fills U_CON / MI_NCON from S_GPOS / S_GMAT.  Counters: mpr calls / support iterations.
static __device__ __noinline__ void collide(float* s, const DevModel& m, int lane, unsigned& mpr_calls, unsigned& mpr_iters) {
  int ncon = 0, nq = 0;
  bool overflow = false;
  unsigned short* mprq = (unsigned short*)(s + U_MPRQ);   // candidate indices < H_MAXCAND = 128
  for (int base = 0; base < m.ncand; base += 32) {
    const int ci = base + lane;
    int n = 0, kind = -1;
    RawCon rc[8];
    bool active = false;
    if (ci < m.ncand) {
      const DevCand& cd = m.cand[ci];
      const int g1 = cd.g1, g2 = cd.g2;
      kind = cd.kind;
      const V3 p1 = ld3(s + S_GPOS + 3 * g1), p2 = ld3(s + S_GPOS + 3 * g2);
      const float rb1 = m.geom_rbound[g1], rb2 = m.geom_rbound[g2];
      active = true;
      if (rb1 > 0.f && rb2 > 0.f) {
        if (norm(p1 - p2) > cd.margin + rb1 + rb2) active = false;
      } else if (m.geom_type[g1] == G_PLANE && rb2 > 0.f) {
        const M3 pm = ldm(s + S_GMAT + 9 * g1);
        if (dot(p2 - p1, col(pm, 2)) > cd.margin + rb2) active = false;
      }
      if (active && kind == K_CONVEX && m.geom_type[g2] == G_MESH && (m.geom_type[g1] == G_CAPSULE || m.geom_type[g1] == G_BOX)) {
        // conservative cull before MPR: separating-axis tests along g1's own axes (against the mesh's local AABB) and
        // along 13 fixed mesh-local directions whose vertex extents were tabulated at model build (a 26-DOP of the
        // mesh).  A separated pair cannot penetrate, so MPR would report no contact as well.
        const M3 R1 = ldm(s + S_GMAT + 9 * g1), R2 = ldm(s + S_GMAT + 9 * g2);
        const float* sz = m.geom_size[g1];
        const bool cap = m.geom_type[g1] == G_CAPSULE;
        const float e1[3] = {sz[0], cap ? sz[0] : sz[1], cap ? sz[1] + sz[0] : sz[2]};
        const float* h2 = m.geom_aabb_h[g2];
        const float tol = cd.margin + 1e-6f;
        {
          const V3 d1 = mulTv(R1, p2 + mulv(R2, ld3(m.geom_aabb_c[g2])) - p1);
          for (int i = 0; i < 3; i++) {
            const V3 a = col(R1, i);
            const float ext = fabsf(dot(a, col(R2, 0))) * h2[0] + fabsf(dot(a, col(R2, 1))) * h2[1] + fabsf(dot(a, col(R2, 2))) * h2[2];
            if (fabsf(comp(d1, i)) > e1[i] + ext + tol) active = false;
          }
        }
        const V3 c1 = mulTv(R2, p1 - p2);                       // g1's centre in the mesh frame
        const V3 a0 = mulTv(R2, col(R1, 0)), a1 = mulTv(R2, col(R1, 1)), a2 = mulTv(R2, col(R1, 2));   // g1's axes there
#pragma unroll 1
        for (int k = 0; k < H_NDOP; k++) {
          const V3 u = v3(c_dop[k][0], c_dop[k][1], c_dop[k][2]);
          const float ext = cap ? fabsf(dot(a2, u)) * sz[1] + sz[0]
                                : fabsf(dot(a0, u)) * e1[0] + fabsf(dot(a1, u)) * e1[1] + fabsf(dot(a2, u)) * e1[2];
          const float c = dot(c1, u);
          if (c - ext > m.geom_dop[g2][k][1] + tol || c + ext < m.geom_dop[g2][k][0] - tol) active = false;
        }
      }
      if (active && kind != K_CONVEX && kind != K_PLANE_MESH) {
        const M3 m1 = ldm(s + S_GMAT + 9 * g1), m2 = ldm(s + S_GMAT + 9 * g2);
        // conservative separating-axis culls in front of the two large out-of-line routines (whose code runs for the whole
        // warp as soon as one lane needs it): a pair separated by more than the margin along a face axis of the box(es) has
        // no contact, so the routine would return 0 as well.
        const float tol = cd.margin + 1e-6f;
        if (kind == K_CAP_BOX) {
          const float* cs = m.geom_size[g1];
          const float* bs = m.geom_size[g2];
          const V3 c = mulTv(m2, p1 - p2), a = mulTv(m2, col(m1, 2)) * cs[1];      // capsule centre / half axis in the box frame
          if (fabsf(c.x) - fabsf(a.x) > bs[0] + cs[0] + tol || fabsf(c.y) - fabsf(a.y) > bs[1] + cs[0] + tol ||
              fabsf(c.z) - fabsf(a.z) > bs[2] + cs[0] + tol) kind = -1;
        } else if (kind == K_BOX_BOX) {
          const float* A = m.geom_size[g1];
          const float* Bz = m.geom_size[g2];
          const V3 d = p2 - p1;
          bool sep = false;
#pragma unroll
          for (int i = 0; i < 3; i++) {
            const V3 u = col(m1, i), w = col(m2, i);
            const float e1 = comp(v3(A[0], A[1], A[2]), i) + fabsf(dot(u, col(m2, 0))) * Bz[0] + fabsf(dot(u, col(m2, 1))) * Bz[1] + fabsf(dot(u, col(m2, 2))) * Bz[2];
            const float e2 = comp(v3(Bz[0], Bz[1], Bz[2]), i) + fabsf(dot(w, col(m1, 0))) * A[0] + fabsf(dot(w, col(m1, 1))) * A[1] + fabsf(dot(w, col(m1, 2))) * A[2];
            sep = sep || fabsf(dot(d, u)) > e1 + tol || fabsf(dot(d, w)) > e2 + tol;
          }
          if (sep) kind = -1;
        }
        switch (kind) {
          case K_PLANE_CAPSULE: n = plane_capsule(rc, cd.margin, p1, m1, p2, m2, m.geom_size[g2]); break;
          case K_PLANE_BOX: n = plane_box(rc, cd.margin, p1, m1, p2, m2, m.geom_size[g2]); break;
          case K_CAP_CAP: n = capsule_capsule(rc, cd.margin, p1, m1, m.geom_size[g1], p2, m2, m.geom_size[g2]); break;
          case K_CAP_BOX: n = capsule_box(rc, cd.margin, p1, m1, m.geom_size[g1], p2, m2, m.geom_size[g2]); break;
          case K_BOX_BOX: n = box_box(rc, cd.margin, p1, m1, m.geom_size[g1], p2, m2, m.geom_size[g2]); break;
          default: break;
        }
      }
    }
    // ordered append of the analytic contacts
    const int off = ncon + wexscan(n, lane);
    for (int k = 0; k < n; k++) {
      if (off + k < H_NCON) write_contact(s, off + k, rc[k], ci);
    }
    const int tot = wsumi(n);
    if (ncon + tot > H_NCON) overflow = true;
    ncon = min(ncon + tot, H_NCON);
    // queue warp-cooperative candidates
    const bool qd = active && (kind == K_CONVEX || kind == K_PLANE_MESH);
    const unsigned bal = __ballot_sync(FULLMASK, qd);
    if (qd) mprq[nq + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)ci;
    nq += __popc(bal);
    __syncwarp();
  }
  for (int k = 0; k < nq; k++) {
    const int ci = mprq[k];
    const DevCand& cd = m.cand[ci];
    RawCon rc;
    int n = 0;
    CObj& o1 = *reinterpret_cast<CObj*>(s + U_COBJ);
    CObj& o2 = *reinterpret_cast<CObj*>(s + U_COBJ + 24);
    __syncwarp();
    if (lane == 0) { if (cd.kind == K_CONVEX) make_cobj(s, m, cd.g1, o1); make_cobj(s, m, cd.g2, o2); }
    __syncwarp();
    const MeshLane ml2 = mesh_lane(o2, lane);
    if (cd.kind == K_CONVEX) {
      float depth;
      V3 dir, pos;
      int it = 0;
      mpr_calls++;
      if (mpr_penetration(o1, o2, m.mpr_tolerance, m.mpr_iterations, depth, dir, pos, it, lane, ml2) == 0 &&
          !(dir.x == 0.f && dir.y == 0.f && dir.z == 0.f)) {
        rc.dist = -depth; rc.n = dir; rc.pos = pos; rc.t = v3(0, 0, 0);
        n = 1;
      }
      mpr_iters += it;
    } else {  // plane - mesh: support vertex along -normal
      const V3 pn = col(ldm(s + S_GMAT + 9 * cd.g1), 2);
      const V3 v = support_w(o2, -pn, lane, ml2).world;
      const float dist = dot(v - ld3(s + S_GPOS + 3 * cd.g1), pn);
      if (dist <= cd.margin) { rc.dist = dist; rc.pos = v - pn * (0.5f * dist); rc.n = pn; rc.t = v3(0, 0, 0); n = 1; }
    }
    if (n) {
      if (ncon < H_NCON) { if (lane == 0) write_contact(s, ncon, rc, ci); ncon++; }
      else overflow = true;
    }
  }
  if (lane == 0) {
    ((int*)(s + S_MISC))[MI_NCON] = ncon;
    if (overflow) ((int*)(s + S_MISC))[MI_FLAGS] |= 2;
  }
  __syncwarp();
}
