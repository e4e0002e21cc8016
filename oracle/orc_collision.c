/* orc_collision.c -- CPU ORACLE (test infrastructure): contact generation for the HandMimic model class.
 * Restates MuJoCo 2.1 `mj_collision` (third-party binary, un-vendored): pair driver,
 * analytic plane/capsule/box routines and libccd's MPR for convex meshes (Appendix B.7).  The capsule-box and
 * box-box routines are closest-feature / SAT+clipping re-derivations (MuJoCo's own routines are not public
 * algorithms); contact convention is MuJoCo's: dist<0 penetration, pos midway between the surfaces,
 * frame[0:3] = normal pointing from geom1 to geom2.  PARITY UNPINNED (no MuJoCo to diff against here).
 */
#include <float.h>
#include <stdlib.h>
#include "orc_math.h"
#include "orc_model.h"

typedef struct { double dist, pos[3], normal[3], tangent[3]; } rawcon;

/* ------------------------------------------------------------------------------------------ primitives */

static int sphere_sphere(rawcon* c, double margin, const double* p1, double r1, const double* p2, double r2) {
  double dif[3], cd;
  sub3(dif, p2, p1);
  cd = norm3(dif);
  if (cd - r1 - r2 > margin) return 0;
  c->dist = cd - r1 - r2;
  if (cd < ORC_MINVAL) { c->normal[0] = 1; c->normal[1] = c->normal[2] = 0; }
  else scl3(c->normal, dif, 1.0 / cd);
  addscl3(c->pos, p1, c->normal, r1 + 0.5 * c->dist);
  zero3(c->tangent);
  return 1;
}

static int plane_sphere(rawcon* c, double margin, const double* ppos, const double* pn, const double* spos, double r) {
  double dif[3], dist;
  sub3(dif, spos, ppos);
  dist = dot3(dif, pn) - r;
  if (dist > margin) return 0;
  c->dist = dist;
  copy3(c->normal, pn);
  addscl3(c->pos, spos, pn, -(r + 0.5 * dist));
  zero3(c->tangent);
  return 1;
}

static int plane_capsule(rawcon* c, double margin, const double* ppos, const double* pmat, const double* cpos,
                         const double* cmat, const double* size) {
  double pn[3] = {pmat[2], pmat[5], pmat[8]}, ax[3] = {cmat[2], cmat[5], cmat[8]}, e[3];
  int n = 0;
  addscl3(e, cpos, ax, size[1]);
  n += plane_sphere(c + n, margin, ppos, pn, e, size[0]);
  addscl3(e, cpos, ax, -size[1]);
  n += plane_sphere(c + n, margin, ppos, pn, e, size[0]);
  for (int i = 0; i < n; i++) copy3(c[i].tangent, ax);
  return n;
}

static int plane_box(rawcon* c, double margin, const double* ppos, const double* pmat, const double* bpos,
                     const double* bmat, const double* size) {
  double pn[3] = {pmat[2], pmat[5], pmat[8]}, dif[3], dist;
  int n = 0;
  sub3(dif, bpos, ppos);
  dist = dot3(dif, pn);
  for (int i = 0; i < 8; i++) {
    double v[3] = {(i & 1 ? size[0] : -size[0]), (i & 2 ? size[1] : -size[1]), (i & 4 ? size[2] : -size[2])}, cw[3], ld;
    mulmat3(cw, bmat, v);
    ld = dot3(pn, cw);
    if (dist + ld > margin || ld > 0) continue;
    c[n].dist = dist + ld;
    add3(c[n].pos, cw, bpos);
    addscl3(c[n].pos, c[n].pos, pn, -0.5 * c[n].dist);
    copy3(c[n].normal, pn);
    zero3(c[n].tangent);
    if (++n >= 4) break;
  }
  return n;
}

static int capsule_capsule(rawcon* c, double margin, const double* p1, const double* m1, const double* s1,
                           const double* p2, const double* m2, const double* s2) {
  double a1[3] = {m1[2], m1[5], m1[8]}, a2[3] = {m2[2], m2[5], m2[8]}, dif[3], v1[3], v2[3];
  double ma, mb, mc, u, v, det, x1, x2;
  sub3(dif, p1, p2);
  ma = dot3(a1, a1); mb = -dot3(a1, a2); mc = dot3(a2, a2);
  u = -dot3(a1, dif); v = dot3(a2, dif);
  det = ma * mc - mb * mb;
  if (fabs(det) >= ORC_MINVAL) {
    x1 = (mc * u - mb * v) / det;
    x2 = (ma * v - mb * u) / det;
    if (x1 > s1[1]) { x1 = s1[1]; x2 = (v - mb * s1[1]) / mc; }
    else if (x1 < -s1[1]) { x1 = -s1[1]; x2 = (v + mb * s1[1]) / mc; }
    if (x2 > s2[1]) { x2 = s2[1]; x1 = (u - mb * s2[1]) / ma; }
    else if (x2 < -s2[1]) { x2 = -s2[1]; x1 = (u + mb * s2[1]) / ma; }
    if (x1 > s1[1]) x1 = s1[1]; else if (x1 < -s1[1]) x1 = -s1[1];
    addscl3(v1, p1, a1, x1);
    addscl3(v2, p2, a2, x2);
    return sphere_sphere(c, margin, v1, s1[0], v2, s2[0]);
  } else {
    /* parallel axes: up to two contacts from the four end-point projections */
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
      addscl3(v1, p1, a1, x1);
      addscl3(v2, p2, a2, x2);
      n += sphere_sphere(c + n, margin, v1, s1[0], v2, s2[0]);
    }
    return n;
  }
}

/* squared distance from point p (box frame) to the solid box, with closest point q */
static double point_box(const double* p, const double* s, double* q) {
  double d2 = 0;
  for (int i = 0; i < 3; i++) {
    q[i] = p[i] < -s[i] ? -s[i] : (p[i] > s[i] ? s[i] : p[i]);
    d2 += (p[i] - q[i]) * (p[i] - q[i]);
  }
  return d2;
}

/* sphere (centre p in box frame) against the box -> contact with geom1 = sphere side, geom2 = box */
static int sphere_box_local(rawcon* c, double margin, const double* p, double r, const double* s) {
  double q[3], d2 = point_box(p, s, q), dist, n[3];
  if (d2 > 0) {
    double dd = sqrt(d2);
    dist = dd - r;
    if (dist > margin) return 0;
    sub3(n, q, p);
    scl3(n, n, 1.0 / dd); /* from sphere centre towards the box */
  } else {
    /* centre inside the box: leave through the nearest face */
    int best = 0; double bd = DBL_MAX, sign = 1;
    for (int i = 0; i < 3; i++) {
      double dpos = s[i] - p[i], dneg = p[i] + s[i];
      if (dpos < bd) { bd = dpos; best = i; sign = 1; }
      if (dneg < bd) { bd = dneg; best = i; sign = -1; }
    }
    zero3(n);
    n[best] = -sign; /* sphere must move along +sign axis to leave; normal points sphere -> box = -sign */
    dist = -bd - r;
    copy3(q, p);
    q[best] = sign * s[best];
  }
  c->dist = dist;
  copy3(c->normal, n);
  addscl3(c->pos, p, n, r + 0.5 * dist);
  zero3(c->tangent);
  return 1;
}

/* capsule (geom1) vs box (geom2).  The distance from the capsule segment p0 + t a, t in [-h,h], to the solid box is
 * convex and piecewise quadratic in t: minimise it exactly (breakpoints where a coordinate crosses a slab face).
 * If the closest feature at the minimiser is a box FACE, the distance is linear over the part of the segment that lies
 * above that face, so contacts are placed at both ends of that part (one if it degenerates); if it is an edge or a
 * vertex, or the segment point is inside the box, a single contact at the minimiser. */
static int capsule_box(rawcon* c, double margin, const double* cp, const double* cm, const double* cs,
                       const double* bp, const double* bm, const double* bs) {
  double dif[3], p0[3], a[3], ax[3] = {cm[2], cm[5], cm[8]}, h = cs[1], r = cs[0];
  double tb[8], best_t = 0, best_d2 = DBL_MAX, p[3], zlo = 0, zhi = 0;
  int nb = 0, n = 0, nout = 0, kface = -1, have_z = 0;
  sub3(dif, cp, bp);
  mulmatT3(p0, bm, dif);
  mulmatT3(a, bm, ax);
  tb[nb++] = -h; tb[nb++] = h;
  for (int i = 0; i < 3; i++)
    if (fabs(a[i]) > 1e-12)
      for (int sgn = -1; sgn <= 1; sgn += 2) {
        double t = (sgn * bs[i] - p0[i]) / a[i];
        if (t > -h && t < h) tb[nb++] = t;
      }
  for (int i = 1; i < nb; i++) { double x = tb[i]; int j = i - 1; while (j >= 0 && tb[j] > x) { tb[j + 1] = tb[j]; j--; } tb[j + 1] = x; }
  for (int k = 0; k + 1 < nb; k++) {
    double t0 = tb[k], t1 = tb[k + 1], tm = 0.5 * (t0 + t1), A = 0, B = 0, C = 0, t, d2;
    /* on [t0,t1]: d2(t) = sum over axes outside their slab of (p0_i + t a_i -/+ s_i)^2 = A t^2 + 2 B t + C */
    for (int i = 0; i < 3; i++) {
      double x = p0[i] + tm * a[i], e;
      if (x > bs[i]) e = p0[i] - bs[i]; else if (x < -bs[i]) e = p0[i] + bs[i]; else continue;
      A += a[i] * a[i]; B += a[i] * e; C += e * e;
    }
    if (A > 0) { t = -B / A; if (t < t0) t = t0; if (t > t1) t = t1; d2 = A * t * t + 2 * B * t + C; }
    else { t = t0; d2 = C; }
    if (d2 < best_d2) { best_d2 = d2; best_t = t; }
    if (A == 0 && C == 0) { if (!have_z) { zlo = t0; have_z = 1; } zhi = t1; }   /* this stretch of the axis is inside the box */
  }
  /* axis through the box interior: every point of the inside stretch is at distance zero; take its midpoint (the entry
     point lies on a face and whether it counts as inside would be decided by rounding) */
  if (have_z) best_t = 0.5 * (zlo + zhi);
  addscl3(p, p0, a, best_t);
  for (int i = 0; i < 3; i++)
    if (fabs(p[i]) > bs[i]) { nout++; kface = i; }
  if (nout == 1) {
    double lo = -h, hi = h;
    for (int i = 0; i < 3; i++) {
      if (i == kface || fabs(a[i]) <= 1e-12) continue;
      double ta = (-bs[i] - p0[i]) / a[i], tc = (bs[i] - p0[i]) / a[i];
      if (ta > tc) { double x = ta; ta = tc; tc = x; }
      if (ta > lo) lo = ta;
      if (tc < hi) hi = tc;
    }
    if (lo > best_t) lo = best_t;
    if (hi < best_t) hi = best_t;
    addscl3(p, p0, a, lo);
    n += sphere_box_local(c + n, margin, p, r, bs);
    if (hi - lo > 1e-6 * (h + r)) {
      addscl3(p, p0, a, hi);
      n += sphere_box_local(c + n, margin, p, r, bs);
    }
  } else {
    n += sphere_box_local(c + n, margin, p, r, bs);
  }
  for (int i = 0; i < n; i++) { /* back to world */
    double t[3];
    mulmat3(t, bm, c[i].pos); add3(c[i].pos, t, bp);
    mulmat3(t, bm, c[i].normal); copy3(c[i].normal, t);
  }
  return n;
}

/* box (geom1) vs box (geom2): separating-axis test over 15 axes, then reference/incident face clipping (<= 8 points)
 * or the closest points of the two edges (1 point). */
static int box_box(rawcon* c, double margin, const double* p1, const double* R1, const double* A, const double* p2,
                   const double* R2, const double* B) {
  double p[3], pp[3], R[9], Q[9], s, best = -DBL_MAX, nrm[3] = {0, 0, 0}, nR[3];
  int code = 0, invert = 0, n = 0;
  sub3(p, p2, p1);
  mulmatT3(pp, R1, p);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      R[3 * i + j] = R1[i] * R2[j] + R1[3 + i] * R2[3 + j] + R1[6 + i] * R2[6 + j];
      Q[3 * i + j] = fabs(R[3 * i + j]);
    }
  /* face axes of box 1 */
  for (int i = 0; i < 3; i++) {
    s = fabs(pp[i]) - (A[i] + B[0] * Q[3 * i] + B[1] * Q[3 * i + 1] + B[2] * Q[3 * i + 2]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = 1 + i; invert = pp[i] < 0; nrm[0] = R1[i]; nrm[1] = R1[3 + i]; nrm[2] = R1[6 + i]; }
  }
  /* face axes of box 2 */
  for (int j = 0; j < 3; j++) {
    double col[3] = {R2[j], R2[3 + j], R2[6 + j]}, pj = dot3(col, p);
    s = fabs(pj) - (A[0] * Q[j] + A[1] * Q[3 + j] + A[2] * Q[6 + j] + B[j]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = 4 + j; invert = pj < 0; copy3(nrm, col); }
  }
  /* edge x edge axes (normalised; slight preference for face axes) */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      double ex = pp[i2] * R[3 * i1 + j] - pp[i1] * R[3 * i2 + j];
      double ra = A[i1] * Q[3 * i2 + j] + A[i2] * Q[3 * i1 + j];
      double rb = B[j1] * Q[3 * i + j2] + B[j2] * Q[3 * i + j1];
      double l2 = R[3 * i1 + j] * R[3 * i1 + j] + R[3 * i2 + j] * R[3 * i2 + j], l;
      s = fabs(ex) - (ra + rb);
      if (l2 < 1e-12) continue; /* parallel edges: covered by face axes */
      l = sqrt(l2);
      s /= l;
      if (s > margin) return 0;
      if (s * 1.05 > best + 1e-12 && s > best) { /* require a real improvement over the face axes */
        double nl[3] = {0, 0, 0};
        best = s; code = 7 + 3 * i + j; invert = ex < 0;
        nl[i1] = -R[3 * i2 + j] / l; nl[i2] = R[3 * i1 + j] / l;
        mulmat3(nrm, R1, nl);
      }
    }
  if (!code) return 0;
  if (invert) scl3(nrm, nrm, -1); /* nrm now points from box 1 towards box 2 */
  if (code > 6) {
    /* edge-edge: closest points between the supporting edges */
    int i = (code - 7) / 3, j = (code - 7) % 3;
    double pa[3], pb[3], ua[3] = {R1[i], R1[3 + i], R1[6 + i]}, ub[3] = {R2[j], R2[3 + j], R2[6 + j]};
    double d[3], uaub, q1, q2, den, al, be;
    copy3(pa, p1);
    for (int k = 0; k < 3; k++) {
      double col[3] = {R1[k], R1[3 + k], R1[6 + k]};
      double sg = dot3(nrm, col) > 0 ? 1 : -1;
      if (k != i) addscl3(pa, pa, col, sg * A[k]);
    }
    copy3(pb, p2);
    for (int k = 0; k < 3; k++) {
      double col[3] = {R2[k], R2[3 + k], R2[6 + k]};
      double sg = dot3(nrm, col) > 0 ? -1 : 1;
      if (k != j) addscl3(pb, pb, col, sg * B[k]);
    }
    sub3(d, pb, pa);
    uaub = dot3(ua, ub); q1 = dot3(ua, d); q2 = -dot3(ub, d);
    den = 1 - uaub * uaub;
    if (den <= 1e-12) { al = be = 0; }
    else { al = (q1 + uaub * q2) / den; be = (uaub * q1 + q2) / den; }
    addscl3(pa, pa, ua, al);
    addscl3(pb, pb, ub, be);
    c[0].dist = best;
    add3(c[0].pos, pa, pb); scl3(c[0].pos, c[0].pos, 0.5);
    copy3(c[0].normal, nrm);
    zero3(c[0].tangent);
    return 1;
  }
  /* face contact.  Reference box a (face normal nR pointing to the other box), incident box b. */
  {
    const double *Ra, *Rb, *pa, *pb, *Sa, *Sb;
    double nr[3], anr[3], center[3], quad[8], ret[16], m11, m12, m21, m22, k1, k2, k3, k4, c1, c2, rect[2];
    double normal2[3];
    int lanr, a1, a2, codeN, code1, code2, nret;
    if (code <= 3) { Ra = R1; Rb = R2; pa = p1; pb = p2; Sa = A; Sb = B; copy3(nR, nrm); }
    else { Ra = R2; Rb = R1; pa = p2; pb = p1; Sa = B; Sb = A; scl3(nR, nrm, -1); }
    mulmatT3(nr, Rb, nR); /* reference normal in incident-box frame */
    for (int k = 0; k < 3; k++) anr[k] = fabs(nr[k]);
    if (anr[1] > anr[0]) { if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
    else { if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; } }
    /* centre of the incident face, relative to pa */
    {
      double col[3] = {Rb[lanr], Rb[3 + lanr], Rb[6 + lanr]};
      sub3(center, pb, pa);
      if (nr[lanr] < 0) addscl3(center, center, col, Sb[lanr]); else addscl3(center, center, col, -Sb[lanr]);
    }
    codeN = (code <= 3) ? code - 1 : code - 4;
    if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
    {
      double ca1[3] = {Ra[code1], Ra[3 + code1], Ra[6 + code1]}, ca2[3] = {Ra[code2], Ra[3 + code2], Ra[6 + code2]};
      double cb1[3] = {Rb[a1], Rb[3 + a1], Rb[6 + a1]}, cb2[3] = {Rb[a2], Rb[3 + a2], Rb[6 + a2]};
      c1 = dot3(center, ca1); c2 = dot3(center, ca2);
      m11 = dot3(ca1, cb1); m12 = dot3(ca1, cb2); m21 = dot3(ca2, cb1); m22 = dot3(ca2, cb2);
    }
    k1 = m11 * Sb[a1]; k2 = m21 * Sb[a1]; k3 = m12 * Sb[a2]; k4 = m22 * Sb[a2];
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
    quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
    quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
    rect[0] = Sa[code1]; rect[1] = Sa[code2];
    /* clip the quad against the reference rectangle (Sutherland-Hodgman) */
    {
      double buf[16], *qv = quad, *rv = ret;
      int nq = 4;
      nret = 0;
      for (int dir = 0; dir <= 1; dir++)
        for (int sign = -1; sign <= 1; sign += 2) {
          double* pq = qv; double* pr = rv;
          nret = 0;
          for (int i = nq; i > 0; i--) {
            double* nextq = (i > 1) ? pq + 2 : qv;
            if (sign * pq[dir] < rect[dir]) { pr[0] = pq[0]; pr[1] = pq[1]; pr += 2; nret++; if (nret & 8) { qv = rv; goto clipdone; } }
            if ((sign * pq[dir] < rect[dir]) ^ (sign * nextq[dir] < rect[dir])) {
              pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * rect[dir] - pq[dir]);
              pr[dir] = sign * rect[dir];
              pr += 2; nret++;
              if (nret & 8) { qv = rv; goto clipdone; }
            }
            pq += 2;
          }
          qv = rv;
          rv = (qv == ret) ? buf : ret;
          nq = nret;
        }
    clipdone:
      if (qv != ret) memcpy(ret, qv, nret * 2 * sizeof(double));
    }
    if (nret < 1) return 0;
    /* back to 3-D; depth along the reference normal */
    {
      double det1 = 1.0 / (m11 * m22 - m12 * m21);
      double cb1[3] = {Rb[a1], Rb[3 + a1], Rb[6 + a1]}, cb2[3] = {Rb[a2], Rb[3 + a2], Rb[6 + a2]};
      copy3(normal2, nR);
      m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
      for (int j = 0; j < nret && n < 8; j++) {
        double kk1 = m22 * (ret[2 * j] - c1) - m12 * (ret[2 * j + 1] - c2);
        double kk2 = -m21 * (ret[2 * j] - c1) + m11 * (ret[2 * j + 1] - c2);
        double pt[3], depth;
        for (int k = 0; k < 3; k++) pt[k] = center[k] + kk1 * cb1[k] + kk2 * cb2[k];
        depth = Sa[codeN] - dot3(normal2, pt);
        if (-depth > margin) continue;
        /* pt is on the incident face (relative to pa); the reference surface is depth along +nR */
        c[n].dist = -depth;
        for (int k = 0; k < 3; k++) c[n].pos[k] = pa[k] + pt[k] + 0.5 * depth * normal2[k];
        copy3(c[n].normal, nrm);
        zero3(c[n].tangent);
        n++;
      }
    }
  }
  return n;
}

/* ------------------------------------------------------------------------------------------ MPR (libccd) */

typedef struct { int type; const double *pos, *mat, *size, *vert; int nvert; } cobj;
typedef struct { double v[3], v1[3], v2[3]; } supp;

static void support_local(const cobj* o, const double* dirw, double* out) {
  double d[3], l[3];
  mulmatT3(d, o->mat, dirw);
  switch (o->type) {
    case GEOM_SPHERE: { double n = norm3(d); if (n < ORC_MINVAL) { zero3(l); } else scl3(l, d, o->size[0] / n); break; }
    case GEOM_CAPSULE: {
      double n = norm3(d);
      if (n < ORC_MINVAL) zero3(l); else scl3(l, d, o->size[0] / n);
      l[2] += d[2] >= 0 ? o->size[1] : -o->size[1];
      break;
    }
    case GEOM_BOX:
      for (int i = 0; i < 3; i++) l[i] = d[i] >= 0 ? o->size[i] : -o->size[i];
      break;
    default: { /* mesh: exhaustive scan over hull vertices */
      double best = -DBL_MAX; int bi = 0;
      for (int i = 0; i < o->nvert; i++) {
        double s = dot3(o->vert + 3 * i, d);
        if (s > best) { best = s; bi = i; }
      }
      copy3(l, o->vert + 3 * bi);
    }
  }
  mulmat3(out, o->mat, l);
  add3(out, out, o->pos);
}

static void mpr_support(const cobj* o1, const cobj* o2, const double* dir, supp* s) {
  double nd[3] = {-dir[0], -dir[1], -dir[2]};
  support_local(o1, dir, s->v1);
  support_local(o2, nd, s->v2);
  sub3(s->v, s->v1, s->v2);
}

#define CCD_EPS DBL_EPSILON
static int ccd_zero(double x) { return fabs(x) < CCD_EPS; }
static int ccd_eq(double a, double b) {
  double ab = fabs(a - b), m;
  if (ab < CCD_EPS) return 1;
  m = fabs(a) > fabs(b) ? fabs(a) : fabs(b);
  return ab < CCD_EPS * m;
}
static int v3_eq(const double* a, const double* b) { return ccd_eq(a[0], b[0]) && ccd_eq(a[1], b[1]) && ccd_eq(a[2], b[2]); }

static double point_seg_dist2(const double* P, const double* x0, const double* b, double* wit) {
  double d[3], a[3], t, dist;
  sub3(d, b, x0);
  sub3(a, x0, P);
  t = -dot3(a, d) / dot3(d, d);
  if (t < 0 || ccd_zero(t)) { double e[3]; sub3(e, x0, P); dist = dot3(e, e); if (wit) copy3(wit, x0); }
  else if (t > 1 || ccd_eq(t, 1)) { double e[3]; sub3(e, b, P); dist = dot3(e, e); if (wit) copy3(wit, b); }
  else {
    double w[3], e[3];
    addscl3(w, x0, d, t);
    sub3(e, w, P);
    dist = dot3(e, e);
    if (wit) copy3(wit, w);
  }
  return dist;
}

static double point_tri_dist2(const double* P, const double* x0, const double* B, const double* C, double* wit) {
  double d1[3], d2[3], a[3], u, v, w, p, q, r, s, t, dist, dist2, w2[3];
  sub3(d1, B, x0); sub3(d2, C, x0); sub3(a, x0, P);
  u = dot3(a, a); v = dot3(d1, d1); w = dot3(d2, d2); p = dot3(a, d1); q = dot3(a, d2); r = dot3(d1, d2);
  s = (q * r - w * p) / (w * v - r * r);
  t = (-s * r - q) / w;
  (void)u;
  if ((ccd_zero(s) || s > 0) && (ccd_eq(s, 1) || s < 1) && (ccd_zero(t) || t > 0) && (ccd_eq(t, 1) || t < 1) &&
      (ccd_eq(t + s, 1) || t + s < 1)) {
    double e[3];
    for (int k = 0; k < 3; k++) wit[k] = x0[k] + s * d1[k] + t * d2[k];
    sub3(e, wit, P);
    dist = dot3(e, e);
  } else {
    dist = point_seg_dist2(P, x0, B, wit);
    dist2 = point_seg_dist2(P, x0, C, w2);
    if (dist2 < dist) { dist = dist2; copy3(wit, w2); }
    dist2 = point_seg_dist2(P, B, C, w2);
    if (dist2 < dist) { dist = dist2; copy3(wit, w2); }
  }
  return dist;
}

static void portal_dir(const supp* s, double* dir) {
  double a[3], b[3];
  sub3(a, s[2].v, s[1].v);
  sub3(b, s[3].v, s[1].v);
  cross3(dir, a, b);
  normalize3(dir);
}
static int portal_reach_tol(const supp* s, const supp* v4, const double* dir, double tol) {
  double dv1 = dot3(s[1].v, dir), dv2 = dot3(s[2].v, dir), dv3 = dot3(s[3].v, dir), dv4 = dot3(v4->v, dir);
  double d1 = dv4 - dv1, d2 = dv4 - dv2, d3 = dv4 - dv3;
  if (d2 < d1) d1 = d2;
  if (d3 < d1) d1 = d3;
  return ccd_eq(d1, tol) || d1 < tol;
}
static void expand_portal(supp* s, const supp* v4) {
  double c[3], d;
  cross3(c, v4->v, s[0].v);
  d = dot3(s[1].v, c);
  if (d > 0) {
    d = dot3(s[2].v, c);
    if (d > 0) s[1] = *v4; else s[3] = *v4;
  } else {
    d = dot3(s[3].v, c);
    if (d > 0) s[2] = *v4; else s[1] = *v4;
  }
}

static void find_pos(const supp* s, double* pos) {
  double dir[3], b[4], sum, inv, c[3], p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
  portal_dir(s, dir);
  cross3(c, s[1].v, s[2].v); b[0] = dot3(c, s[3].v);
  cross3(c, s[3].v, s[2].v); b[1] = dot3(c, s[0].v);
  cross3(c, s[0].v, s[1].v); b[2] = dot3(c, s[3].v);
  cross3(c, s[2].v, s[1].v); b[3] = dot3(c, s[0].v);
  sum = b[0] + b[1] + b[2] + b[3];
  if (ccd_zero(sum) || sum < 0) {
    b[0] = 0;
    cross3(c, s[2].v, s[3].v); b[1] = dot3(c, dir);
    cross3(c, s[3].v, s[1].v); b[2] = dot3(c, dir);
    cross3(c, s[1].v, s[2].v); b[3] = dot3(c, dir);
    sum = b[1] + b[2] + b[3];
  }
  inv = 1.0 / sum;
  for (int i = 0; i < 4; i++) { addscl3(p1, p1, s[i].v1, b[i]); addscl3(p2, p2, s[i].v2, b[i]); }
  for (int k = 0; k < 3; k++) pos[k] = 0.5 * inv * (p1[k] + p2[k]);
}

/* returns 0 when intersecting (depth, dir, pos filled), -1 otherwise */
static int mpr_penetration(const cobj* o1, const cobj* o2, double tol, int max_iter, double* depth, double* dir,
                           double* pos, int* iters) {
  supp s[4], v4;
  double d[3], va[3], vb[3], dt;
  int res = 0, size;
  static const double origin[3] = {0, 0, 0};
  /* discover portal */
  copy3(s[0].v1, o1->pos); copy3(s[0].v2, o2->pos); sub3(s[0].v, s[0].v1, s[0].v2);
  if (v3_eq(s[0].v, origin)) s[0].v[0] += CCD_EPS * 10;
  scl3(d, s[0].v, -1); normalize3(d);
  mpr_support(o1, o2, d, &s[1]);
  dt = dot3(s[1].v, d);
  if (ccd_zero(dt) || dt < 0) return -1;
  cross3(d, s[0].v, s[1].v);
  if (ccd_zero(dot3(d, d))) {
    res = v3_eq(s[1].v, origin) ? 1 : 2;
  } else {
    normalize3(d);
    mpr_support(o1, o2, d, &s[2]);
    dt = dot3(s[2].v, d);
    if (ccd_zero(dt) || dt < 0) return -1;
    sub3(va, s[1].v, s[0].v); sub3(vb, s[2].v, s[0].v);
    cross3(d, va, vb); normalize3(d);
    dt = dot3(d, s[0].v);
    if (dt > 0) { supp t = s[1]; s[1] = s[2]; s[2] = t; scl3(d, d, -1); }
    size = 3;
    while (size < 4) {
      int cont = 0;
      mpr_support(o1, o2, d, &s[3]);
      (*iters)++;
      dt = dot3(s[3].v, d);
      if (ccd_zero(dt) || dt < 0) return -1;
      cross3(va, s[1].v, s[3].v);
      dt = dot3(va, s[0].v);
      if (dt < 0 && !ccd_zero(dt)) { s[2] = s[3]; cont = 1; }
      if (!cont) {
        cross3(va, s[3].v, s[2].v);
        dt = dot3(va, s[0].v);
        if (dt < 0 && !ccd_zero(dt)) { s[1] = s[3]; cont = 1; }
      }
      if (cont) {
        sub3(va, s[1].v, s[0].v); sub3(vb, s[2].v, s[0].v);
        cross3(d, va, vb); normalize3(d);
      } else size = 4;
    }
  }
  if (res == 1) { /* touching contact on v1 */
    *depth = 0; zero3(dir);
    add3(pos, s[1].v1, s[1].v2); scl3(pos, pos, 0.5);
    return 0;
  }
  if (res == 2) { /* origin on the v0-v1 segment */
    add3(pos, s[1].v1, s[1].v2); scl3(pos, pos, 0.5);
    *depth = norm3(s[1].v);
    copy3(dir, s[1].v); normalize3(dir);
    return 0;
  }
  /* refine portal */
  for (;;) {
    portal_dir(s, d);
    dt = dot3(d, s[1].v);
    if (ccd_zero(dt) || dt > 0) break; /* portal encapsulates origin */
    mpr_support(o1, o2, d, &v4);
    (*iters)++;
    dt = dot3(v4.v, d);
    if (!(ccd_zero(dt) || dt > 0) || portal_reach_tol(s, &v4, d, tol)) return -1;
    expand_portal(s, &v4);
  }
  /* penetration info */
  for (int it = 0;; it++) {
    portal_dir(s, d);
    mpr_support(o1, o2, d, &v4);
    (*iters)++;
    if (portal_reach_tol(s, &v4, d, tol) || it > max_iter) {
      double wit[3];
      *depth = sqrt(point_tri_dist2(origin, s[1].v, s[2].v, s[3].v, wit));
      if (ccd_zero(wit[0]) && ccd_zero(wit[1]) && ccd_zero(wit[2])) copy3(wit, d);
      normalize3(wit);
      copy3(dir, wit);
      find_pos(s, pos);
      return 0;
    }
    expand_portal(s, &v4);
  }
}

static void make_cobj(const orc_model* m, const orc_data* d, int g, cobj* o) {
  o->type = m->geom_type[g]; o->pos = d->geom_xpos[g]; o->mat = d->geom_xmat[g]; o->size = m->geom_size[g];
  o->vert = m->geom_vertadr[g] >= 0 ? m->mesh_vert + 3 * m->geom_vertadr[g] : NULL;
  o->nvert = m->geom_vertnum[g];
}

static int convex_convex(const orc_model* m, orc_data* d, rawcon* c, double margin, int g1, int g2) {
  cobj o1, o2;
  double depth, dir[3], pos[3];
  int iters = 0;
  (void)margin;
  make_cobj(m, d, g1, &o1);
  make_cobj(m, d, g2, &o2);
  d->mpr_calls++;
  if (mpr_penetration(&o1, &o2, m->mpr_tolerance, m->mpr_iterations, &depth, dir, pos, &iters) != 0) {
    d->mpr_iters += iters;
    return 0;
  }
  d->mpr_iters += iters;
  if (dir[0] == 0 && dir[1] == 0 && dir[2] == 0) return 0;
  c->dist = -depth;
  copy3(c->normal, dir);
  copy3(c->pos, pos);
  zero3(c->tangent);
  return 1;
}

static int plane_convex(const orc_model* m, const orc_data* d, rawcon* c, double margin, int g1, int g2) {
  /* support vertex of the mesh along -normal (single point; see DESIGN.md "known deviations") */
  cobj o;
  const double* pm = d->geom_xmat[g1];
  double pn[3] = {pm[2], pm[5], pm[8]}, nd[3] = {-pm[2], -pm[5], -pm[8]}, v[3], dif[3], dist;
  make_cobj(m, d, g2, &o);
  support_local(&o, nd, v);
  sub3(dif, v, d->geom_xpos[g1]);
  dist = dot3(dif, pn);
  if (dist > margin) return 0;
  c->dist = dist;
  addscl3(c->pos, v, pn, -0.5 * dist);
  copy3(c->normal, pn);
  zero3(c->tangent);
  return 1;
}

/* ------------------------------------------------------------------------------------------ driver */

static void make_frame(double* f) {
  double t;
  normalize3(f);
  if (norm3(f + 3) < 0.5) {
    zero3(f + 3);
    if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  }
  t = dot3(f, f + 3);
  addscl3(f + 3, f + 3, f, -t);
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}

static void collide_geoms(const orc_model* m, orc_data* d, int g1, int g2, int pair) {
  rawcon rc[8];
  int n = 0, t1, t2;
  double margin, gap, solref[2], solimp[5], friction[5];
  int condim;
  if (pair >= 0) { g1 = m->pair_geom1[pair]; g2 = m->pair_geom2[pair]; }
  if (m->geom_type[g1] > m->geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
  t1 = m->geom_type[g1]; t2 = m->geom_type[g2];
  if (pair < 0) {
    if (m->geom_bodyid[g1] == m->geom_bodyid[g2]) return;
    if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]))) return;
    margin = fmax(m->geom_margin[g1], m->geom_margin[g2]);
    gap = fmax(m->geom_gap[g1], m->geom_gap[g2]);
  } else { margin = m->pair_margin[pair]; gap = m->pair_gap[pair]; }
  /* bounding-sphere / plane-sphere filter */
  if (m->geom_rbound[g1] > 0 && m->geom_rbound[g2] > 0) {
    double dif[3];
    sub3(dif, d->geom_xpos[g1], d->geom_xpos[g2]);
    if (norm3(dif) > margin + m->geom_rbound[g1] + m->geom_rbound[g2]) return;
  } else if (t1 == GEOM_PLANE && m->geom_rbound[g2] > 0) {
    const double* pm = d->geom_xmat[g1];
    double pn[3] = {pm[2], pm[5], pm[8]}, dif[3];
    sub3(dif, d->geom_xpos[g2], d->geom_xpos[g1]);
    if (dot3(dif, pn) > margin + m->geom_rbound[g2]) return;
  }
  if (t1 == GEOM_PLANE && t2 == GEOM_CAPSULE)
    n = plane_capsule(rc, margin, d->geom_xpos[g1], d->geom_xmat[g1], d->geom_xpos[g2], d->geom_xmat[g2], m->geom_size[g2]);
  else if (t1 == GEOM_PLANE && t2 == GEOM_BOX)
    n = plane_box(rc, margin, d->geom_xpos[g1], d->geom_xmat[g1], d->geom_xpos[g2], d->geom_xmat[g2], m->geom_size[g2]);
  else if (t1 == GEOM_PLANE && t2 == GEOM_MESH)
    n = plane_convex(m, d, rc, margin, g1, g2);
  else if (t1 == GEOM_CAPSULE && t2 == GEOM_CAPSULE)
    n = capsule_capsule(rc, margin, d->geom_xpos[g1], d->geom_xmat[g1], m->geom_size[g1], d->geom_xpos[g2], d->geom_xmat[g2], m->geom_size[g2]);
  else if (t1 == GEOM_CAPSULE && t2 == GEOM_BOX)
    n = capsule_box(rc, margin, d->geom_xpos[g1], d->geom_xmat[g1], m->geom_size[g1], d->geom_xpos[g2], d->geom_xmat[g2], m->geom_size[g2]);
  else if (t1 == GEOM_BOX && t2 == GEOM_BOX)
    n = box_box(rc, margin, d->geom_xpos[g1], d->geom_xmat[g1], m->geom_size[g1], d->geom_xpos[g2], d->geom_xmat[g2], m->geom_size[g2]);
  else if (t2 == GEOM_MESH && t1 != GEOM_PLANE)
    n = convex_convex(m, d, rc, margin, g1, g2);
  else
    return; /* plane-plane etc.: not collidable */
  if (n <= 0) return;
  /* contact parameters */
  if (pair >= 0) {
    condim = m->pair_dim[pair];
    memcpy(friction, m->pair_friction[pair], sizeof friction);
    memcpy(solref, m->pair_solref[pair], sizeof solref);
    memcpy(solimp, m->pair_solimp[pair], sizeof solimp);
  } else {
    double s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2], mix = (s1 + s2 > 0) ? s1 / (s1 + s2) : 0.5, f[3];
    condim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
    for (int k = 0; k < 3; k++) f[k] = fmax(m->geom_friction[g1][k], m->geom_friction[g2][k]);
    friction[0] = friction[1] = f[0]; friction[2] = f[1]; friction[3] = friction[4] = f[2];
    if (m->geom_solref[g1][0] > 0 && m->geom_solref[g2][0] > 0)
      for (int k = 0; k < 2; k++) solref[k] = mix * m->geom_solref[g1][k] + (1 - mix) * m->geom_solref[g2][k];
    else
      for (int k = 0; k < 2; k++) solref[k] = fmin(m->geom_solref[g1][k], m->geom_solref[g2][k]);
    for (int k = 0; k < 5; k++) solimp[k] = mix * m->geom_solimp[g1][k] + (1 - mix) * m->geom_solimp[g2][k];
  }
  for (int i = 0; i < n; i++) {
    orc_contact* con;
    if (d->ncon >= m->nconmax) { d->warn_contactfull = 1; return; }
    con = &d->contact[d->ncon++];
    con->dist = rc[i].dist;
    copy3(con->pos, rc[i].pos);
    copy3(con->frame, rc[i].normal);
    copy3(con->frame + 3, rc[i].tangent);
    make_frame(con->frame);
    con->includemargin = margin - gap;
    con->dim = condim;
    con->geom1 = g1; con->geom2 = g2;
    memcpy(con->friction, friction, sizeof friction);
    memcpy(con->solref, solref, sizeof solref);
    memcpy(con->solimp, solimp, sizeof solimp);
    con->exclude = 0;
    con->mu = 0;
    con->efc_address = -1;
  }
}

void orc_collision(const orc_model* m, orc_data* d) {
  int pairadr = 0;
  d->ncon = 0;
  for (int b1 = 0; b1 < m->nbody; b1++)
    for (int b2 = b1 + 1; b2 < m->nbody; b2++) {
      int sig = (b1 << 16) + b2, merged = 0;
      while (pairadr < m->npair && m->pair_signature[pairadr] <= sig) {
        if (m->pair_signature[pairadr] == sig) merged = 1;
        collide_geoms(m, d, -1, -1, pairadr++);
      }
      if (merged) continue;
      if (!m->body_geomnum[b1] || !m->body_geomnum[b2]) continue;
      if (m->body_weldid[b1] == 0 && m->body_weldid[b2] == 0) continue;      /* both static */
      if (m->body_weldid[b1] == m->body_weldid[b2]) continue;                /* welded together */
      {
        /* parent-child filter (on weld bodies), unless the parent is static */
        int w1 = m->body_weldid[b1], w2 = m->body_weldid[b2];
        int pw1 = m->body_weldid[m->body_parentid[w1]], pw2 = m->body_weldid[m->body_parentid[w2]];
        if ((w1 != 0 && w2 != 0) && (pw1 == w2 || pw2 == w1)) continue;
      }
      for (int g1 = m->body_geomadr[b1]; g1 < m->body_geomadr[b1] + m->body_geomnum[b1]; g1++)
        for (int g2 = m->body_geomadr[b2]; g2 < m->body_geomadr[b2] + m->body_geomnum[b2]; g2++)
          collide_geoms(m, d, g1, g2, -1);
    }
  while (pairadr < m->npair) collide_geoms(m, d, -1, -1, pairadr++);
}
