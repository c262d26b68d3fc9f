// TEST INFRASTRUCTURE (oracle) — not part of the product.
//
// CPU restatement ("port") of the reference's segmentation + lifting path, written from the
// behaviour of the reference sources (each function cites the file:line it follows).  It exists
// so that parity tests have a checker that (a) runs on the GPU box, where /root/reference is
// absent, and (b) finishes a 1080p pair in about a second instead of the reference's ~30 s
// (flat arrays + one stable sort instead of std::multiset / std::set).
//
// Pinned against: the reference's three known-answer tests (cpp/tests/test_liftig_3d.cpp:69-89,
// 179-227) and against oracle/_ref (the unchanged reference sources) on data/frame_1052-1053 and
// seeded synthetic fields — see tests/test_oracle.py.
//
// Compile with -ffp-contract=off: every float/double operation below is meant to round exactly
// once, as the reference's plain x86-64 build does.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

namespace {

struct P2 {
    float x, y;
};

struct Mat3 {
    float m[9];
};

struct Sol {
    int cls = 0;
    bool has_rect = false;
    P2 ps_bev[4], lower[4], upper[4], rect[4];
    double w_error = -1.0, h_error = -1.0, orient = 0.0;
};

const float NANF = std::numeric_limits<float>::quiet_NaN();
const float INFF = std::numeric_limits<float>::infinity();

inline P2 sub(P2 a, P2 b) { return {a.x - b.x, a.y - b.y}; }
inline P2 add(P2 a, P2 b) { return {a.x + b.x, a.y + b.y}; }
inline double norm2(P2 p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }

// get_intersect, lifting_3d.cpp:63-88 — all float; parallel lines give (NaN, NaN).
P2 isect(P2 A, P2 B, P2 C, P2 D) {
    float a1 = B.y - A.y;
    float b1 = A.x - B.x;
    float c1 = a1 * A.x + b1 * A.y;
    float a2 = D.y - C.y;
    float b2 = C.x - D.x;
    float c2 = a2 * C.x + b2 * C.y;
    float det = a1 * b2 - a2 * b1;
    if (std::fabs(det) < 1e-9) return {NANF, NANF};
    return {(b2 * c1 - b1 * c2) / det, (a1 * c2 - a2 * c1) / det};
}

// warp_perspective, lifting_3d.cpp:112-121 — float, left to right.
P2 warp(P2 p, const Mat3& M) {
    float den = M.m[6] * p.x + M.m[7] * p.y + M.m[8];
    float px = (M.m[0] * p.x + M.m[1] * p.y + M.m[2]) / den;
    float py = (M.m[3] * p.x + M.m[4] * p.y + M.m[5]) / den;
    return {px, py};
}

const int kObjSize[3][2] = {{258, 84}, {349, 165}, {370, 180}};  // getObjSize, lifting_3d.cpp:255-259

// get_bottom, lifting_3d.cpp:162-217 (note its parameters (w, h) receive (dim_l, dim_w), :399).
bool get_bottom(const P2 bev[4], double orient, double w, double h, double* error, P2 out[4]) {
    P2 a[4];
    for (int i = 0; i < 4; ++i) a[i] = {bev[i].x, -bev[i].y};
    const double co = std::cos(orient), si = std::sin(orient);
    P2 k = isect(a[3], {(float)(a[3].x + co), (float)(a[3].y + si)}, a[0], a[1]);
    if (k.x == INFF || k.y == INFF) return false;
    double l = norm2(sub(a[3], k));
    if (l == 0) return false;
    // ((l - w) * a0 + w * a3) / l with OpenCV's per-operator float rounding
    P2 t0 = {(float)(a[0].x * (l - w)), (float)(a[0].y * (l - w))};
    P2 t1 = {(float)(a[3].x * w), (float)(a[3].y * w)};
    P2 s = add(t0, t1);
    P2 c = {(float)(s.x / l), (float)(s.y / l)};
    P2 b = isect(c, {(float)(c.x + co), (float)(c.y + si)}, a[0], a[1]);
    if (b.x == INFF) return false;
    double ew = norm2(sub(c, b));
    double error_w = (ew < w) ? ew / w : w / ew;
    P2 d = isect(c, {(float)(c.x - si), (float)(c.y + co)}, a[3], a[2]);
    if (d.x == INFF) return false;
    double el = norm2(sub(c, d));
    double error_l = (el < h) ? el / h : h / el;
    P2 bd = add(b, d);
    P2 center = {bd.x / 2, bd.y / 2};
    P2 f = sub(P2{center.x * 2, center.y * 2}, c);
    P2 r[4] = {c, b, f, d};
    for (int i = 0; i < 4; ++i) out[i] = {r[i].x, -r[i].y};
    *error = error_w * error_l;
    return true;
}

// get_bottom_variants, lifting_3d.cpp:350-439 (+ get_motion_direction :219-253, get_upper_face :290-348).
Sol bottom_variants(P2 dir, const int box[4], const Mat3& mat, const Mat3& inv_mat, const Mat3& inv_upper, int cls) {
    const int xmin = box[0], ymin = box[1], xmax = box[2], ymax = box[3];
    // motion direction in the bird's-eye view
    P2 center = {(float)((xmin + xmax) / 2), (float)((ymin + ymax) / 2)};
    double n = norm2(dir);
    P2 nd = {(float)(dir.x / n), (float)(dir.y / n)};
    P2 t1 = warp(center, mat);
    P2 t2 = warp(add(center, nd), mat);
    double vx = t2.x - t1.x;
    double vy = t1.y - t2.y;
    double orient = std::atan2(vy, vx);
    Sol s;
    if (std::isinf(orient)) return s;  // default Solution: errors -1, no rectangle
    s.cls = cls;
    P2 ps[4] = {{(float)xmin, (float)ymax}, {(float)xmin, (float)ymin}, {(float)xmax, (float)ymin}, {(float)xmax, (float)ymax}};
    for (int i = 0; i < 4; ++i) s.ps_bev[i] = warp(ps[i], mat);
    double error;
    P2 corners[4];
    if (!get_bottom(s.ps_bev, orient, kObjSize[cls][0], kObjSize[cls][1], &error, corners)) {
        s.w_error = s.h_error = 0.0;  // Solution(cls, {}, {}, {}, {}, 0, 0, 0), lifting_3d.cpp:405
        s.orient = 0.0;
        return s;
    }
    for (int i = 0; i < 4; ++i) s.lower[i] = warp(corners[i], inv_mat);
    const P2* lf = s.lower;
    P2 u[4];
    u[2] = sub(lf[2], P2{0.f, lf[2].y - (float)ymin});
    P2 right_van = isect(lf[1], lf[2], lf[0], lf[3]);
    u[1] = isect(u[2], right_van, {(float)xmin, (float)ymin}, {(float)xmin, (float)ymax});
    P2 left_van = isect(lf[2], lf[3], lf[0], lf[1]);
    u[3] = isect(u[2], left_van, {(float)xmax, (float)ymin}, {(float)xmax, (float)ymax});
    u[0] = isect(left_van, u[1], right_van, u[3]);
    for (int i = 0; i < 4; ++i) {
        s.upper[i] = u[i];
        s.rect[i] = corners[i];
    }
    s.has_rect = true;
    s.w_error = error;
    P2 expected_edge = warp(corners[0], inv_upper);
    double expected_h = norm2(sub(lf[0], expected_edge));
    double computed_h = norm2(sub(u[0], lf[0]));
    s.h_error = (computed_h < expected_h) ? computed_h / expected_h : expected_h / computed_h;
    s.orient = orient;
    return s;
}

// get_score, graph.cpp:241-270 — first maximum over cls 0..2 wins, empty rectangles skipped.
double get_score(const int box[4], P2 dir, const Mat3& persp, const Mat3& inv, const Mat3 upper[3], Sol* best) {
    double max_score = -1.0;
    for (int cls = 0; cls < 3; ++cls) {
        Sol s = bottom_variants(dir, box, persp, inv, upper[cls], cls);
        double sc = (s.w_error + s.h_error) / 2;
        if (s.has_rect && max_score < sc) {
            max_score = sc;
            *best = s;
        }
    }
    return max_score;
}

// getPerspectiveTransform as used by get_mat/get_mat_upper (lifting_3d.cpp:479,510-511):
// 8x8 system, LU with partial pivoting in double, M[8] = 1, rounded to float.
Mat3 perspective(const double src[4][2], const double dst[4][2]) {
    double a[8][9];
    for (int i = 0; i < 4; ++i) {
        double x = (float)src[i][0], y = (float)src[i][1], u = (float)dst[i][0], v = (float)dst[i][1];
        double r0[9] = {x, y, 1, 0, 0, 0, -x * u, -y * u, u};
        double r1[9] = {0, 0, 0, x, y, 1, -x * v, -y * v, v};
        std::memcpy(a[i], r0, sizeof r0);
        std::memcpy(a[i + 4], r1, sizeof r1);
    }
    for (int i = 0; i < 8; ++i) {
        int k = i;
        for (int j = i + 1; j < 8; ++j)
            if (std::fabs(a[j][i]) > std::fabs(a[k][i])) k = j;
        if (k != i)
            for (int j = i; j < 9; ++j) std::swap(a[i][j], a[k][j]);
        double d = -1 / a[i][i];
        for (int j = i + 1; j < 8; ++j) {
            double alpha = a[j][i] * d;
            for (int c = i + 1; c < 9; ++c) a[j][c] += alpha * a[i][c];
        }
    }
    double xs[8];
    for (int i = 7; i >= 0; --i) {
        double s = a[i][8];
        for (int k = i + 1; k < 8; ++k) s -= a[i][k] * xs[k];
        xs[i] = s / a[i][i];
    }
    Mat3 m;
    for (int i = 0; i < 8; ++i) m.m[i] = (float)xs[i];
    m.m[8] = 1.f;
    return m;
}

struct Entry {
    int root;
    int size;     // snapshot size
    int time;     // index of the merge (0-based, in merge order) that produced the snapshot
    double score, move;
    float flow[2];
    int bbox[4];
    Sol sol;
};

struct Trace {  // one record per merge, in merge order
    std::vector<int32_t> loser, winner, size, edge_pos;
    std::vector<int32_t> bbox;  // 4 per merge
    std::vector<float> flow;    // 2 per merge
};

struct Result {
    std::vector<Entry> entries;    // ascending root
    std::vector<int32_t> list_next;  // pixel lists: snapshot of root r = first `size` items from r
    long n_edges = 0;
    int num_sets = 0;
    long counters[8] = {0};  // merges, fail_size, fail_row, fail_move, get_score, fail_no_rect, fail_convexity, fail_score
    long history_writes = 0;
    double t_build = 0, t_segment = 0;
    Trace trace;
    bool keep_trace = false;
};

struct Params {
    double score_threshold = 0.3;  // graph.hpp:93
    int min_size = 500;            // graph.hpp:94
};

// Edge enumeration of build_graph (graph.cpp:62-93) with weights of diff (segment.cpp:20-32),
// ordered as the std::multiset orders them (graph.cpp:55-60,96-99): ascending weight, equal
// weights in insertion order == stable sort by weight of the insertion sequence.
void sorted_edges(const float* flow, int W, int H, bool n8, std::vector<int32_t>& start, std::vector<int32_t>& end,
                  std::vector<double>& weight) {
    std::vector<int32_t> s, e;
    std::vector<double> w;
    size_t cap = (size_t)4 * W * H;
    s.reserve(cap);
    e.reserve(cap);
    w.reserve(cap);
    auto push = [&](int x, int y, int x1, int y1) {
        const float* a = flow + 2 * ((size_t)y * W + x);
        const float* b = flow + 2 * ((size_t)y1 * W + x1);
        double dx = a[0] - b[0];
        double dy = a[1] - b[1];
        s.push_back(y * W + x);
        e.push_back(y1 * W + x1);
        w.push_back(std::sqrt(dx * dx + dy * dy));
    };
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            if (x > 0) push(x, y, x - 1, y);
            if (y > 0) push(x, y, x, y - 1);
            if (n8) {
                if (x > 0 && y > 0) push(x, y, x - 1, y - 1);
                if (x > 0 && y < H - 1) push(x, y, x - 1, y + 1);
            }
        }
    std::vector<uint32_t> idx(s.size());
    std::iota(idx.begin(), idx.end(), 0u);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t i, uint32_t j) { return w[i] < w[j]; });
    start.resize(idx.size());
    end.resize(idx.size());
    weight.resize(idx.size());
    for (size_t i = 0; i < idx.size(); ++i) {
        start[i] = s[idx[i]];
        end[i] = e[idx[i]];
        weight[i] = w[idx[i]];
    }
}

// segment_graph (graph.cpp:503-536) + Forest::merge (:170-218) + Forest::new_merge (:272-384).
Result* run_segment(const float* flow, int W, int H, int neighbors, const Mat3& persp, const Mat3& inv,
                    const Mat3 upper[3], const Params& prm, bool keep_trace) {
    auto* res = new Result();
    res->keep_trace = keep_trace;
    const int N = W * H;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<int32_t> es, ee;
    std::vector<double> ew;
    sorted_edges(flow, W, H, neighbors == 8, es, ee, ew);
    auto t1 = std::chrono::steady_clock::now();
    res->n_edges = (long)es.size();

    // Forest::Forest, graph.cpp:129-148
    std::vector<int32_t> parent(N), size(N, 1), tail(N);
    std::vector<uint8_t> rank(N, 0);
    std::vector<float> fl(flow, flow + 2 * (size_t)N);
    std::vector<int32_t> bb(4 * (size_t)N);
    res->list_next.assign(N, -1);
    std::vector<int32_t> hist(N, -1);  // root -> index into entries
    for (int i = 0; i < N; ++i) {
        parent[i] = i;
        tail[i] = i;
        bb[4 * i] = bb[4 * i + 2] = i % W;
        bb[4 * i + 1] = bb[4 * i + 3] = i / W;
    }
    auto find = [&](int n) {  // Forest::find, graph.cpp:150-157 (full path compression)
        int r = n;
        while (parent[r] != r) r = parent[r];
        while (parent[n] != r) {
            int nx = parent[n];
            parent[n] = r;
            n = nx;
        }
        return r;
    };
    int num_sets = N;
    int merges = 0;
    long* cnt = res->counters;
    for (size_t k = 0; k < es.size(); ++k) {
        int a = find(es[k]);
        int b = find(ee[k]);
        if (a == b) continue;
        // Forest::merge
        if (rank[a] > rank[b]) std::swap(a, b);
        parent[a] = b;
        int sa = size[a], sb = size[b];
        float wax = fl[2 * a] * (float)sa, way = fl[2 * a + 1] * (float)sa;  // Vec2f * int
        float wbx = fl[2 * b] * (float)sb, wby = fl[2 * b + 1] * (float)sb;
        float sx = wax + wbx, sy = way + wby;                                  // Vec2f + Vec2f
        double inv_n = 1. / (sa + sb);                                         // Vec2f / int
        fl[2 * b] = (float)(sx * inv_n);
        fl[2 * b + 1] = (float)(sy * inv_n);
        res->list_next[tail[b]] = a;  // pixel list of b gets a's list appended
        tail[b] = tail[a];
        size[b] = sa + sb;
        size[a] = 0;
        bb[4 * b] = std::min(bb[4 * b], bb[4 * a]);
        bb[4 * b + 1] = std::min(bb[4 * b + 1], bb[4 * a + 1]);
        bb[4 * b + 2] = std::max(bb[4 * b + 2], bb[4 * a + 2]);
        bb[4 * b + 3] = std::max(bb[4 * b + 3], bb[4 * a + 3]);
        if (rank[a] == rank[b]) rank[b] += 1;
        num_sets -= 1;
        const int t = merges++;
        cnt[0]++;
        if (keep_trace) {
            Trace& tr = res->trace;
            tr.loser.push_back(a);
            tr.winner.push_back(b);
            tr.size.push_back(size[b]);
            tr.edge_pos.push_back((int32_t)k);
            tr.bbox.insert(tr.bbox.end(), &bb[4 * b], &bb[4 * b] + 4);
            tr.flow.push_back(fl[2 * b]);
            tr.flow.push_back(fl[2 * b + 1]);
        }
        // Forest::new_merge gates
        if (size[b] < prm.min_size) { cnt[1]++; continue; }
        int y = b / W;
        if (y < H / 10) { cnt[2]++; continue; }
        double move = std::sqrt((double)fl[2 * b] * fl[2 * b] + (double)fl[2 * b + 1] * fl[2 * b + 1]);
        if (move < 3 * (y + 1) / static_cast<double>(H)) { cnt[3]++; continue; }
        const int* bx = &bb[4 * b];
        double rect_area = ((bx[2] - bx[0] + 1) * (bx[3] - bx[1] + 1));
        double convexity = size[b] / rect_area;
        Sol sol;
        cnt[4]++;
        double score = get_score(bx, P2{fl[2 * b], fl[2 * b + 1]}, persp, inv, upper, &sol);
        if (score == -1) { cnt[5]++; continue; }
        double min_convexity = sol.cls == 0 ? 3.0 / 4.0 : sol.cls == 1 ? 1.0 / 2.0 : 20.0 / 29.0;
        if (convexity < min_convexity) { cnt[6]++; continue; }
        if (!(score > prm.score_threshold)) { cnt[7]++; continue; }
        double prev = hist[b] < 0 ? -1.0 : res->entries[hist[b]].score;
        if (prev < score) {
            Entry e;
            e.root = b;
            e.size = size[b];
            e.time = t;
            e.score = score;
            e.move = move;
            e.flow[0] = fl[2 * b];
            e.flow[1] = fl[2 * b + 1];
            std::memcpy(e.bbox, bx, sizeof e.bbox);
            e.sol = sol;
            if (hist[b] < 0) {
                hist[b] = (int)res->entries.size();
                res->entries.push_back(e);
            } else {
                res->entries[hist[b]] = e;
            }
            res->history_writes++;
        }
    }
    std::sort(res->entries.begin(), res->entries.end(), [](const Entry& x, const Entry& y) { return x.root < y.root; });
    res->num_sets = num_sets;
    auto t2 = std::chrono::steady_clock::now();
    res->t_build = std::chrono::duration<double>(t1 - t0).count();
    res->t_segment = std::chrono::duration<double>(t2 - t1).count();
    return res;
}

Mat3 to_mat(const float* m) {
    Mat3 r;
    std::memcpy(r.m, m, sizeof r.m);
    return r;
}

void put4(const P2 p[4], bool valid, float* out) {
    for (int i = 0; i < 4; ++i) {
        out[2 * i] = valid ? p[i].x : NANF;
        out[2 * i + 1] = valid ? p[i].y : NANF;
    }
}

}  // namespace

extern "C" {

struct ref_solution {  // same layout as oracle/ref_driver.cpp
    int cls;
    int has_rectangle;
    float ps_bev[8], lower_face[8], upper_face[8], rectangle[8];
    double w_error, h_error, orient;
};

static void fill_solution(const Sol& s, ref_solution* out) {
    out->cls = s.cls;
    out->has_rectangle = s.has_rect;
    put4(s.ps_bev, s.has_rect, out->ps_bev);
    put4(s.lower, s.has_rect, out->lower_face);
    put4(s.upper, s.has_rect, out->upper_face);
    put4(s.rect, s.has_rect, out->rectangle);
    out->w_error = s.w_error;
    out->h_error = s.h_error;
    out->orient = s.orient;
}

// get_mat / get_mat_upper calibration constants, lifting_3d.cpp:441-514
void ref_get_mats(float* persp, float* inv, float* upper) {
    const double img[4][2] = {{215, 265}, {90, 121}, {294, 120}, {625, 265}};
    const double bev[4][2] = {{100, 13000}, {100, 6000}, {800, 6000}, {800, 13000}};
    std::memcpy(persp, perspective(img, bev).m, 9 * sizeof(float));
    std::memcpy(inv, perspective(bev, img).m, 9 * sizeof(float));
    const double ys[3][2] = {{176, 85}, {185, 80}, {140, 55}};
    for (int c = 0; c < 3; ++c) {
        const double roof[4][2] = {{215, ys[c][0]}, {90, ys[c][1]}, {294, ys[c][1]}, {625, ys[c][0]}};
        std::memcpy(upper + 9 * c, perspective(bev, roof).m, 9 * sizeof(float));
    }
}

void ref_get_intersect(const float* a, const float* b, const float* c, const float* d, float* out) {
    P2 p = isect({a[0], a[1]}, {b[0], b[1]}, {c[0], c[1]}, {d[0], d[1]});
    out[0] = p.x;
    out[1] = p.y;
}

void ref_get_bottom_variants(const float* dir, const int* box, const float* mat, const float* inv_mat,
                             const float* inv_upper, int cls, ref_solution* out) {
    Sol s = bottom_variants({dir[0], dir[1]}, box, to_mat(mat), to_mat(inv_mat), to_mat(inv_upper), cls);
    fill_solution(s, out);
}

long ref_build_graph(const float* flow, int width, int height, int neighbors8, int32_t* start, int32_t* end,
                     double* weight) {
    std::vector<int32_t> s, e;
    std::vector<double> w;
    sorted_edges(flow, width, height, neighbors8 != 0, s, e, w);
    std::memcpy(start, s.data(), s.size() * sizeof(int32_t));
    std::memcpy(end, e.data(), e.size() * sizeof(int32_t));
    std::memcpy(weight, w.data(), w.size() * sizeof(double));
    return (long)s.size();
}

void* oracle_segment_ex(const float* flow_blurred, int width, int height, int neighbors, const float* persp,
                        const float* inv, const float* upper, double score_threshold, int min_size, int keep_trace) {
    Mat3 up[3] = {to_mat(upper), to_mat(upper + 9), to_mat(upper + 18)};
    Params prm;
    prm.score_threshold = score_threshold;
    prm.min_size = min_size;
    return run_segment(flow_blurred, width, height, neighbors, to_mat(persp), to_mat(inv), up, prm, keep_trace != 0);
}

void* ref_segment(const float* flow_blurred, int width, int height, int neighbors, const float* persp,
                  const float* inv, const float* upper) {
    return oracle_segment_ex(flow_blurred, width, height, neighbors, persp, inv, upper, 0.3, 500, 0);
}

int ref_result_count(void* h) { return (int)static_cast<Result*>(h)->entries.size(); }
long ref_result_edges(void* h) { return static_cast<Result*>(h)->n_edges; }
int ref_result_num_sets(void* h) { return static_cast<Result*>(h)->num_sets; }
void ref_result_times(void* h, double* tb, double* ts) {
    *tb = static_cast<Result*>(h)->t_build;
    *ts = static_cast<Result*>(h)->t_segment;
}
int ref_result_entry(void* h, int i, int* root, double* score, double* move, ref_solution* sol) {
    const Entry& e = static_cast<Result*>(h)->entries[i];
    *root = e.root;
    *score = e.score;
    *move = e.move;
    fill_solution(e.sol, sol);
    return e.size;
}
void ref_result_pixels(void* h, int i, int32_t* out) {
    auto* r = static_cast<Result*>(h);
    const Entry& e = r->entries[i];
    int p = e.root;
    for (int k = 0; k < e.size; ++k) {
        out[k] = p;
        p = r->list_next[p];
    }
    std::sort(out, out + e.size);
}
// extras of the port: snapshot time / mean flow / bbox of entry i, gate counters, merge trace
void oracle_result_entry_extra(void* h, int i, int* time, float* flow2, int* bbox4) {
    const Entry& e = static_cast<Result*>(h)->entries[i];
    *time = e.time;
    flow2[0] = e.flow[0];
    flow2[1] = e.flow[1];
    std::memcpy(bbox4, e.bbox, sizeof e.bbox);
}
void oracle_result_counters(void* h, long* out9) {
    auto* r = static_cast<Result*>(h);
    for (int i = 0; i < 8; ++i) out9[i] = r->counters[i];
    out9[8] = r->history_writes;
}
long oracle_result_trace(void* h, int32_t* loser, int32_t* winner, int32_t* size, int32_t* edge_pos, int32_t* bbox,
                         float* flow) {
    const Trace& t = static_cast<Result*>(h)->trace;
    size_t n = t.loser.size();
    if (loser) {
        std::memcpy(loser, t.loser.data(), n * 4);
        std::memcpy(winner, t.winner.data(), n * 4);
        std::memcpy(size, t.size.data(), n * 4);
        std::memcpy(edge_pos, t.edge_pos.data(), n * 4);
        std::memcpy(bbox, t.bbox.data(), n * 16);
        std::memcpy(flow, t.flow.data(), n * 8);
    }
    return (long)n;
}
void ref_result_free(void* h) { delete static_cast<Result*>(h); }

void ref_set_counting(int) {}
int ref_get_counts(char* buf, int cap) {
    if (buf && cap > 0) buf[0] = 0;
    return 1;
}

}  // extern "C"
