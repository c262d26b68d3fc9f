// lifting_3d on the device: monocular 3D box fitting for one (bounding box, motion direction, class)
// problem per thread.  Follows the reference's arithmetic operation by operation (float geometry,
// double norms / trigonometry / errors, OpenCV's per-operator rounding of Point2f expressions):
//   get_intersect          cpp/src/lifting_3d.cpp:63-88
//   warp_perspective       cpp/src/lifting_3d.cpp:112-121
//   get_bottom             cpp/src/lifting_3d.cpp:162-217
//   get_motion_direction   cpp/src/lifting_3d.cpp:219-253
//   get_upper_face         cpp/src/lifting_3d.cpp:290-348
//   get_bottom_variants    cpp/src/lifting_3d.cpp:350-439
//   get_score              cpp/src/graph.cpp:241-270
#pragma once
#include "dofs_common.cuh"

struct P2f {
    float x, y;
};

struct LiftSolution {
    int cls;
    int has_rect;  // !Solution::rectangle.empty()
    P2f ps_bev[4], lower[4], upper[4], rect[4];
    double w_error, h_error, orient;
};

DOFS_D P2f p2(float x, float y) {
    P2f p;
    p.x = x;
    p.y = y;
    return p;
}
DOFS_D P2f p2sub(P2f a, P2f b) { return p2(xfsub(a.x, b.x), xfsub(a.y, b.y)); }
DOFS_D P2f p2add(P2f a, P2f b) { return p2(xfadd(a.x, b.x), xfadd(a.y, b.y)); }
DOFS_D double p2norm(P2f a) { return norm2d(a.x, a.y); }

DOFS_D float dofs_nanf() { return __int_as_float(0x7fc00000); }
DOFS_D float dofs_inff() { return __int_as_float(0x7f800000); }

// Intersection of lines AB and CD, all float; (NaN, NaN) when |det| < 1e-9.
DOFS_D P2f lift_isect(P2f A, P2f B, P2f C, P2f D) {
    float a1 = xfsub(B.y, A.y);
    float b1 = xfsub(A.x, B.x);
    float c1 = xfadd(xfmul(a1, A.x), xfmul(b1, A.y));
    float a2 = xfsub(D.y, C.y);
    float b2 = xfsub(C.x, D.x);
    float c2 = xfadd(xfmul(a2, C.x), xfmul(b2, C.y));
    float det = xfsub(xfmul(a1, b2), xfmul(a2, b1));
    if ((double)fabsf(det) < 1e-9) return p2(dofs_nanf(), dofs_nanf());
    float x = xfdiv(xfsub(xfmul(b2, c1), xfmul(b1, c2)), det);
    float y = xfdiv(xfsub(xfmul(a1, c2), xfmul(a2, c1)), det);
    return p2(x, y);
}

// Homography applied to a point, float, summed left to right.
DOFS_D P2f lift_warp(P2f p, const float* M) {
    float den = xfadd(xfadd(xfmul(M[6], p.x), xfmul(M[7], p.y)), M[8]);
    float nx = xfadd(xfadd(xfmul(M[0], p.x), xfmul(M[1], p.y)), M[2]);
    float ny = xfadd(xfadd(xfmul(M[3], p.x), xfmul(M[4], p.y)), M[5]);
    return p2(xfdiv(nx, den), xfdiv(ny, den));
}

// Ground rectangle of prior size (w, h) = (dim_l, dim_w) anchored at the warped box corners.
DOFS_D bool lift_bottom(const P2f* bev, double orient, double w, double h, double* error, P2f* out) {
    P2f a0 = p2(bev[0].x, -bev[0].y), a1 = p2(bev[1].x, -bev[1].y);
    P2f a2 = p2(bev[2].x, -bev[2].y), a3 = p2(bev[3].x, -bev[3].y);
    const double co = cos(orient), si = sin(orient);
    P2f k = lift_isect(a3, p2((float)xdadd((double)a3.x, co), (float)xdadd((double)a3.y, si)), a0, a1);
    if (k.x == dofs_inff() || k.y == dofs_inff()) return false;
    double l = p2norm(p2sub(a3, k));
    if (l == 0) return false;
    const double lw = xdsub(l, w);
    P2f t0 = p2((float)xdmul((double)a0.x, lw), (float)xdmul((double)a0.y, lw));
    P2f t1 = p2((float)xdmul((double)a3.x, w), (float)xdmul((double)a3.y, w));
    P2f s = p2add(t0, t1);
    P2f c = p2((float)xddiv((double)s.x, l), (float)xddiv((double)s.y, l));
    P2f b = lift_isect(c, p2((float)xdadd((double)c.x, co), (float)xdadd((double)c.y, si)), a0, a1);
    if (b.x == dofs_inff()) return false;
    double ew = p2norm(p2sub(c, b));
    double error_w = (ew < w) ? xddiv(ew, w) : xddiv(w, ew);
    P2f d = lift_isect(c, p2((float)xdsub((double)c.x, si), (float)xdadd((double)c.y, co)), a3, a2);
    if (d.x == dofs_inff()) return false;
    double el = p2norm(p2sub(c, d));
    double error_l = (el < h) ? xddiv(el, h) : xddiv(h, el);
    P2f bd = p2add(b, d);
    P2f center = p2(xfdiv(bd.x, 2.0f), xfdiv(bd.y, 2.0f));
    P2f f = p2sub(p2(xfmul(center.x, 2.0f), xfmul(center.y, 2.0f)), c);
    out[0] = p2(c.x, -c.y);
    out[1] = p2(b.x, -b.y);
    out[2] = p2(f.x, -f.y);
    out[3] = p2(d.x, -d.y);
    *error = xdmul(error_w, error_l);
    return true;
}

// box = xmin, ymin, xmax, ymax
DOFS_D void lift_bottom_variants(float dirx, float diry, int xmin, int ymin, int xmax, int ymax, const float* mat,
                                 const float* inv_mat, const float* inv_upper, int cls, int dim_l, int dim_w,
                                 LiftSolution* s) {
    s->cls = 0;
    s->has_rect = 0;
    s->w_error = -1.0;
    s->h_error = -1.0;
    s->orient = 0.0;
    P2f center = p2((float)((xmin + xmax) / 2), (float)((ymin + ymax) / 2));
    double n = norm2d(dirx, diry);
    P2f nd = p2((float)xddiv((double)dirx, n), (float)xddiv((double)diry, n));
    P2f t1 = lift_warp(center, mat);
    P2f t2 = lift_warp(p2add(center, nd), mat);
    double vx = (double)xfsub(t2.x, t1.x);
    double vy = (double)xfsub(t1.y, t2.y);
    double orient = atan2(vy, vx);
    if (isinf(orient)) return;
    s->cls = cls;
    P2f ps[4] = {p2((float)xmin, (float)ymax), p2((float)xmin, (float)ymin), p2((float)xmax, (float)ymin),
                 p2((float)xmax, (float)ymax)};
#pragma unroll
    for (int i = 0; i < 4; ++i) s->ps_bev[i] = lift_warp(ps[i], mat);
    double error;
    if (!lift_bottom(s->ps_bev, orient, (double)dim_l, (double)dim_w, &error, s->rect)) {
        s->w_error = 0.0;
        s->h_error = 0.0;
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) s->lower[i] = lift_warp(s->rect[i], inv_mat);
    const P2f l0 = s->lower[0], l1 = s->lower[1], l2 = s->lower[2], l3 = s->lower[3];
    P2f u2 = p2sub(l2, p2(0.0f, xfsub(l2.y, (float)ymin)));
    P2f right_van = lift_isect(l1, l2, l0, l3);
    P2f u1 = lift_isect(u2, right_van, p2((float)xmin, (float)ymin), p2((float)xmin, (float)ymax));
    P2f left_van = lift_isect(l2, l3, l0, l1);
    P2f u3 = lift_isect(u2, left_van, p2((float)xmax, (float)ymin), p2((float)xmax, (float)ymax));
    P2f u0 = lift_isect(left_van, u1, right_van, u3);
    s->upper[0] = u0;
    s->upper[1] = u1;
    s->upper[2] = u2;
    s->upper[3] = u3;
    s->has_rect = 1;
    s->w_error = error;
    P2f expected_edge = lift_warp(s->rect[0], inv_upper);
    double expected_h = p2norm(p2sub(l0, expected_edge));
    double computed_h = p2norm(p2sub(u0, l0));
    s->h_error = (computed_h < expected_h) ? xddiv(computed_h, expected_h) : xddiv(expected_h, computed_h);
    s->orient = orient;
}

// Best of the three classes by (w_error + h_error) / 2; strict '<' so the first maximum wins and NaN
// never wins; returns -1 when no class produced a rectangle.
DOFS_D double lift_get_score(float dirx, float diry, int xmin, int ymin, int xmax, int ymax, const SegParams& P,
                             LiftSolution* best) {
    double max_score = -1.0;
    for (int cls = 0; cls < 3; ++cls) {
        LiftSolution s;
        lift_bottom_variants(dirx, diry, xmin, ymin, xmax, ymax, P.hg.persp, P.hg.inv, P.hg.upper[cls], cls,
                             P.cls_size[cls][0], P.cls_size[cls][1], &s);
        double sc = xddiv(xdadd(s.w_error, s.h_error), 2.0);
        if (s.has_rect && max_score < sc) {
            max_score = sc;
            *best = s;
        }
    }
    return max_score;
}
