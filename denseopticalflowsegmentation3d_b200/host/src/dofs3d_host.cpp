// C++ host shim: the reference's cpp/inc entry points (build_graph, segment_graph, Forest,
// get_segmented_array, get_bottom_variants, get_mat, get_mat_upper, get_intersect) implemented on top of
// the C ABI of libdofs3d.so (include/dofs3d.h).  Nothing here computes on the CPU except argument
// marshalling, the setup-sized get_intersect, and turning label images into std::set<int> pixel sets.
//
// Errors: the reference's hot path never throws; failures are sentinels (Solution() with errors -1,
// score -1).  Device failures have no sentinel in that vocabulary, so they throw std::runtime_error.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>

#include "../../../include/dofs3d.h"
#include "../inc/graph.hpp"
#include "../inc/lifting_3d.hpp"
#include "../inc/segment.hpp"

bool debug = false;

namespace {

struct CtxKey {
    int w, h, n, neighbors;
    float mats[45];
    bool operator<(const CtxKey& o) const {
        if (std::tie(w, h, n, neighbors) != std::tie(o.w, o.h, o.n, o.neighbors))
            return std::tie(w, h, n, neighbors) < std::tie(o.w, o.h, o.n, o.neighbors);
        return std::memcmp(mats, o.mats, sizeof mats) < 0;
    }
};

std::mutex g_mu;
std::map<CtxKey, dofs3d_ctx*> g_ctx;  // contexts are reused across calls (allocation is the expensive part)

[[noreturn]] void fail(dofs3d_ctx* c, int rc, const char* what) {
    throw std::runtime_error(std::string(what) + ": dofs3d status " + std::to_string(rc) + " (" +
                             (c ? dofs3d_last_error(c) : "no context") + ")");
}

void fill_mats(dofs3d_params& p, const cv::Matx33f& persp, const cv::Matx33f& inv, const std::vector<cv::Matx33f>& up) {
    std::memcpy(p.persp, persp.val, sizeof p.persp);
    std::memcpy(p.inv, inv.val, sizeof p.inv);
    for (size_t c = 0; c < 3 && c < up.size(); ++c) std::memcpy(p.inv_upper[c], up[c].val, sizeof p.inv_upper[c]);
}

dofs3d_ctx* context_for(int w, int h, int n, const dofs3d_params& p) {
    CtxKey k;
    k.w = w;
    k.h = h;
    k.n = n;
    k.neighbors = p.neighbors;
    std::memcpy(k.mats, p.persp, sizeof p.persp);
    std::memcpy(k.mats + 9, p.inv, sizeof p.inv);
    std::memcpy(k.mats + 18, p.inv_upper, sizeof p.inv_upper);
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_ctx.find(k);
    if (it != g_ctx.end()) return it->second;
    dofs3d_ctx* c = nullptr;
    int rc = dofs3d_create(&c, 0, w, h, n, &p);
    if (rc != 0) {
        std::string msg = c ? dofs3d_last_error(c) : "no CUDA device";
        if (c) dofs3d_destroy(c);
        throw std::runtime_error("dofs3d_create failed (" + msg + "); there is no CPU fallback");
    }
    g_ctx[k] = c;
    return c;
}

std::vector<cv::Point2f> points4(const float* p) {
    std::vector<cv::Point2f> v(4);
    for (int i = 0; i < 4; ++i) v[i] = cv::Point2f(p[2 * i], p[2 * i + 1]);
    return v;
}

Solution solution_of(const dofs3d_box& b, bool has_rectangle) {
    if (!has_rectangle) {
        // get_bottom (lifting_3d.cpp:404-406) failed: Solution(cls, {}, {}, {}, {}, 0, 0, 0)
        return Solution(b.cls, {}, {}, {}, {}, b.w_error, b.h_error, b.orient);
    }
    return Solution(b.cls, points4(b.ps_bev), points4(b.lower_face), points4(b.upper_face), points4(b.rectangle), b.w_error,
                    b.h_error, b.orient);
}

void check_flow(const cv::Mat& flow, const char* who) {
    if (flow.empty() || flow.type() != CV_32FC2 || !flow.isContinuous())
        throw std::invalid_argument(std::string(who) + ": flow must be a continuous CV_32FC2 matrix");
}

}  // namespace

struct Forest::Result {
    int n_boxes = 0;
    int final_root = 0;
    std::vector<dofs3d_box> boxes;
    std::vector<int32_t> labels;
    dofs3d_stats stats;
    // pixel sets, built on first use
    std::vector<std::set<int>> sets;
    bool sets_built = false;
    void build_sets() {
        if (sets_built) return;
        sets.assign(n_boxes, {});
        std::vector<std::vector<int>> own(n_boxes);
        for (size_t p = 0; p < labels.size(); ++p)
            if (labels[p] >= 0) own[labels[p]].push_back((int)p);
        for (int b = 0; b < n_boxes; ++b)
            for (int a = b; a >= 0; a = boxes[a].parent_box) sets[a].insert(own[b].begin(), own[b].end());
        sets_built = true;
    }
};

// ------------------------------------------------------------------------------------------------
double diff(const cv::Mat& flow, int x1, int y1, int x2, int y2) {
    const cv::Point2f a = flow.at<cv::Point2f>(y1, x1), b = flow.at<cv::Point2f>(y2, x2);
    const double dx = a.x - b.x, dy = a.y - b.y;
    return std::sqrt(dx * dx + dy * dy);
}

std::vector<Edge> build_graph(const cv::Mat& img, int width, int height, const DiffFunction& diff_fn, bool neighborhood_8) {
    check_flow(img, "build_graph");
    if (img.cols != width || img.rows != height) throw std::invalid_argument("build_graph: size mismatch");
    dofs3d_params p;
    dofs3d_default_params(&p);
    p.neighbors = neighborhood_8 ? 8 : 4;
    dofs3d_ctx* c = context_for(width, height, 1, p);
    const size_t slots = 4 * (size_t)width * height;
    std::vector<int32_t> s(slots), e(slots);
    std::vector<uint64_t> w(slots);
    long long n = dofs3d_edges_sorted(c, img.ptr<float>(), s.data(), e.data(), w.data());
    if (n < 0) fail(c, (int)n, "build_graph");
    std::vector<Edge> edges((size_t)n);
    for (long long i = 0; i < n; ++i) {
        double wt;
        std::memcpy(&wt, &w[i], 8);
        edges[i] = Edge{s[i], e[i], wt};
    }
    if (diff_fn) {  // spot check: the callback must be the reference's flow distance
        const long long probes[3] = {0, n / 2, n - 1};
        for (long long i : probes) {
            if (i < 0 || i >= n) continue;
            const Edge& ed = edges[i];
            const double d = diff_fn(img, ed.start % width, ed.start / width, ed.end % width, ed.end / width);
            if (!(d == ed.weight))
                throw std::invalid_argument("build_graph: only the reference's diff (segment.cpp:20-32) is supported on the device");
        }
    }
    return edges;
}

static Forest run_segment(const cv::Mat& flow, int already_blurred, const cv::Mat& bev, const cv::Matx33f& persp_mat,
                          const cv::Matx33f& inv_mat, const std::vector<cv::Matx33f>& inv_mat_upper, int neighbors,
                          float* blurred_out) {
    check_flow(flow, "segment_graph");
    if (inv_mat_upper.size() < 3) throw std::invalid_argument("segment_graph: inv_mat_upper needs one matrix per class (3)");
    dofs3d_params p;
    dofs3d_default_params(&p);
    p.neighbors = neighbors;
    fill_mats(p, persp_mat, inv_mat, inv_mat_upper);
    const int W = flow.cols, H = flow.rows;
    dofs3d_ctx* c = context_for(W, H, 1, p);
    auto res = std::make_shared<Forest::Result>();
    const int cap = 4096;
    res->boxes.resize(cap);
    res->labels.resize((size_t)W * H);
    int32_t nb = 0;
    int rc = dofs3d_segment(c, flow.ptr<float>(), already_blurred, 1, res->labels.data(), res->boxes.data(), &nb, cap,
                            &res->stats, blurred_out);
    if (rc != 0) fail(c, rc, "segment_graph");
    res->n_boxes = nb;
    res->boxes.resize(nb);
    res->final_root = res->stats.final_root;
    Forest f;
    f.num_sets = W * H - res->stats.n_merges;
    f.width = W;
    f.height = H;
    f.min_move = 5;
    f.bev = bev;
    f.persp_mat = persp_mat;
    f.inv_mat = inv_mat;
    f.inv_mat_upper = inv_mat_upper;
    f.result = res;
    return f;
}

Forest segment_graph(const cv::Mat& flow, const std::vector<Edge>& graph_edges, const cv::Mat& bev, const cv::Matx33f& persp_mat,
                     const cv::Matx33f& inv_mat, const std::vector<cv::Matx33f>& inv_mat_upper) {
    const long long W = flow.cols, H = flow.rows;
    const long long e8 = 4 * W * H - 3 * W - 3 * H + 2, e4 = 2 * W * H - W - H;
    const long long n = (long long)graph_edges.size();
    if (n != e8 && n != e4) throw std::invalid_argument("segment_graph: graph_edges was not built from this flow field");
    return run_segment(flow, 1, bev, persp_mat, inv_mat, inv_mat_upper, n == e8 ? 8 : 4, nullptr);
}

Forest get_segmented_array(const cv::Mat& flow, const cv::Mat& bev, const cv::Matx33f& persp_mat, const cv::Matx33f& inv_mat,
                           const std::vector<cv::Matx33f>& inv_mat_upper, int neighbor) {
    if (neighbor != 4 && neighbor != 8) {
        std::fprintf(stderr, "[seg] [error] Invalid neighborhood chosen. The acceptable values are 4 or 8.\n");
        std::fprintf(stderr, "[seg] [error] Segmenting with 4-neighborhood...\n");
        neighbor = 4;
    }
    check_flow(flow, "get_segmented_array");
    // the reference blurs the caller's matrix in place (segment.cpp:52): write the blurred field back
    float* inplace = const_cast<float*>(flow.ptr<float>());
    return run_segment(flow, 0, bev, persp_mat, inv_mat, inv_mat_upper, neighbor, inplace);
}

// ------------------------------------------------------------------------------------------------
int Forest::find(int) const {
    if (!result) throw std::logic_error("Forest::find on an empty forest");
    return result->final_root;
}

std::vector<std::pair<int, SegmentData>> Forest::get_best_segments_sparse() const {
    std::vector<std::pair<int, SegmentData>> out;
    if (!result) return out;
    result->build_sets();
    for (int b = 0; b < result->n_boxes; ++b) {
        const dofs3d_box& bx = result->boxes[b];
        out.emplace_back(bx.root, SegmentData(bx.score, result->sets[b], solution_of(bx, true), bx.move));
    }
    return out;
}

std::vector<SegmentData> Forest::get_best_segments() {
    std::vector<SegmentData> hist((size_t)width * height);
    for (auto& kv : get_best_segments_sparse()) hist[kv.first] = kv.second;
    return hist;
}

std::vector<cv::Point2i> Forest::get_bounding_box(int node_id) const {
    if (result)
        for (const dofs3d_box& b : result->boxes)
            if (b.root == node_id) return {cv::Point2i(b.bbox[0], b.bbox[1]), cv::Point2i(b.bbox[2], b.bbox[3])};
    if (result && node_id == result->final_root) return {cv::Point2i(0, 0), cv::Point2i(width - 1, height - 1)};
    return {};
}

double Forest::get_segment_best_score(int node_id) const {
    if (result)
        for (const dofs3d_box& b : result->boxes)
            if (b.root == node_id) return b.score;
    return -1.0;
}

int Forest::merge(int, int) { throw std::logic_error("Forest::merge: the merge loop runs on the device as a whole (segment_graph)"); }
void Forest::new_merge(int, int, double, int, double, double) {
    throw std::logic_error("Forest::new_merge: the merge loop runs on the device as a whole (segment_graph)");
}

// ------------------------------------------------------------------------------------------------
std::vector<Solution> get_bottom_variants_batch(const std::vector<cv::Point2f>& dirs,
                                                const std::vector<std::vector<cv::Point2i>>& boxes, const cv::Matx33f& mat,
                                                const cv::Matx33f& inv_mat, const std::vector<cv::Matx33f>& inv_matrix_upper,
                                                const std::vector<int>& cls) {
    const size_t n = dirs.size();
    if (boxes.size() != n || cls.size() != n || inv_matrix_upper.size() < 3)
        throw std::invalid_argument("get_bottom_variants_batch: argument sizes");
    dofs3d_params p;
    dofs3d_default_params(&p);
    fill_mats(p, mat, inv_mat, inv_matrix_upper);
    dofs3d_ctx* c = context_for(64, 64, 1, p);
    std::vector<float> d(2 * n);
    std::vector<int32_t> b(4 * n), k(n);
    for (size_t i = 0; i < n; ++i) {
        if (boxes[i].size() != 2) throw std::invalid_argument("get_bottom_variants: box_2d must hold 2 points");
        d[2 * i] = dirs[i].x;
        d[2 * i + 1] = dirs[i].y;
        b[4 * i] = boxes[i][0].x;
        b[4 * i + 1] = boxes[i][0].y;
        b[4 * i + 2] = boxes[i][1].x;
        b[4 * i + 3] = boxes[i][1].y;
        k[i] = cls[i];
    }
    std::vector<dofs3d_box> out(n);
    int rc = dofs3d_lift(c, d.data(), b.data(), k.data(), (int)n, out.data());
    if (rc != 0) fail(c, rc, "get_bottom_variants");
    std::vector<Solution> sols(n);
    for (size_t i = 0; i < n; ++i) {
        if (out[i].size) sols[i] = solution_of(out[i], true);
        else if (out[i].w_error == 0.0) sols[i] = solution_of(out[i], false);  // get_bottom failed (lifting_3d.cpp:405)
        else sols[i] = Solution();                                               // infinite orientation (lifting_3d.cpp:367)
    }
    return sols;
}

Solution get_bottom_variants(const cv::Point2f& orig_mov_dir, const std::vector<cv::Point2i>& box_2d, const cv::Matx33f& mat,
                             const cv::Matx33f& inv_mat, const cv::Matx33f& inv_matrix_upper, int cls) {
    if (cls < 0 || cls > 2) throw std::invalid_argument("get_bottom_variants: cls must be 0, 1 or 2");
    std::vector<cv::Matx33f> up(3, inv_matrix_upper);
    return get_bottom_variants_batch({orig_mov_dir}, {box_2d}, mat, inv_mat, up, {cls})[0];
}

std::pair<cv::Matx33f, cv::Matx33f> get_mat() {
    dofs3d_params p;
    dofs3d_default_params(&p);
    cv::Matx33f a, b;
    std::memcpy(a.val, p.persp, sizeof p.persp);
    std::memcpy(b.val, p.inv, sizeof p.inv);
    return {a, b};
}

std::pair<cv::Matx33f, cv::Matx33f> get_mat(int width, int height) {
    dofs3d_params p;
    dofs3d_params_for_size(&p, width, height);
    cv::Matx33f a, b;
    std::memcpy(a.val, p.persp, sizeof p.persp);
    std::memcpy(b.val, p.inv, sizeof p.inv);
    return {a, b};
}

cv::Matx33f get_mat_upper(int cls, int width, int height) {
    if (cls < 0 || cls > 2) throw std::invalid_argument("get_mat_upper: cls must be 0, 1 or 2");
    dofs3d_params p;
    dofs3d_params_for_size(&p, width, height);
    cv::Matx33f a;
    std::memcpy(a.val, p.inv_upper[cls], sizeof p.inv_upper[cls]);
    return a;
}

cv::Matx33f get_mat_upper(int cls) {
    if (cls < 0 || cls > 2) throw std::invalid_argument("get_mat_upper: cls must be 0, 1 or 2");
    dofs3d_params p;
    dofs3d_default_params(&p);
    cv::Matx33f a;
    std::memcpy(a.val, p.inv_upper[cls], sizeof p.inv_upper[cls]);
    return a;
}

cv::Point2f get_intersect(cv::Point2f A, cv::Point2f B, cv::Point2f C, cv::Point2f D) {
    const float a1 = B.y - A.y, b1 = A.x - B.x, c1 = a1 * A.x + b1 * A.y;
    const float a2 = D.y - C.y, b2 = C.x - D.x, c2 = a2 * C.x + b2 * C.y;
    const float det = a1 * b2 - a2 * b1;
    if (std::fabs(det) < 1e-9) {
        const float nan = std::numeric_limits<float>::quiet_NaN();
        return cv::Point2f(nan, nan);
    }
    return cv::Point2f((b2 * c1 - b1 * c2) / det, (a1 * c2 - a2 * c1) / det);
}

// ------------------------------------------------------------------------------------------------
cv::Mat dense_flow(const cv::Mat& f1, const cv::Mat& f2) {
    if (f1.empty() || f1.type() != CV_8UC3 || f2.type() != CV_8UC3 || f1.rows != f2.rows || f1.cols != f2.cols)
        throw std::invalid_argument("dense_flow: two CV_8UC3 frames of equal size expected");
    const int W = f1.cols, H = f1.rows;
    dofs3d_params p;
    dofs3d_default_params(&p);
    dofs3d_ctx* c = context_for(W, H, 1, p);
    const size_t N = (size_t)W * H;
    std::vector<uint8_t> bgr(2 * N * 3), gray(2 * N);
    std::memcpy(bgr.data(), f1.ptr<uint8_t>(), N * 3);
    std::memcpy(bgr.data() + N * 3, f2.ptr<uint8_t>(), N * 3);
    int rc = dofs3d_gray(c, bgr.data(), 2, gray.data());
    if (rc != 0) fail(c, rc, "dense_flow/gray");
    cv::Mat flow(H, W, CV_32FC2);
    rc = dofs3d_flow(c, gray.data(), gray.data() + N, 1, flow.ptr<float>());
    if (rc != 0) fail(c, rc, "dense_flow");
    return flow;
}

std::vector<Forest> process_video(const std::vector<cv::Mat>& frames, const cv::Matx33f& persp_mat, const cv::Matx33f& inv_mat,
                                  const std::vector<cv::Matx33f>& inv_mat_upper, int neighbor) {
    std::vector<Forest> out;
    if (frames.size() < 2) return out;
    if (neighbor != 4 && neighbor != 8) neighbor = 4;
    const int W = frames[0].cols, H = frames[0].rows, n = (int)frames.size() - 1;
    const size_t N = (size_t)W * H;
    dofs3d_params p;
    dofs3d_default_params(&p);
    p.neighbors = neighbor;
    fill_mats(p, persp_mat, inv_mat, inv_mat_upper);
    dofs3d_ctx* c = context_for(W, H, n, p);
    std::vector<uint8_t> bgr((size_t)(n + 1) * N * 3);
    for (int i = 0; i <= n; ++i) {
        if (frames[i].type() != CV_8UC3 || frames[i].cols != W || frames[i].rows != H)
            throw std::invalid_argument("process_video: CV_8UC3 frames of equal size expected");
        std::memcpy(bgr.data() + (size_t)i * N * 3, frames[i].ptr<uint8_t>(), N * 3);
    }
    const int cap = 1024;
    std::vector<int32_t> labels((size_t)n * N), nb(n);
    std::vector<dofs3d_box> boxes((size_t)n * cap);
    std::vector<dofs3d_stats> stats(n);
    int rc = dofs3d_process(c, bgr.data(), n + 1, labels.data(), boxes.data(), nb.data(), cap, stats.data());
    if (rc != 0) fail(c, rc, "process_video");
    for (int i = 0; i < n; ++i) {
        auto res = std::make_shared<Forest::Result>();
        res->n_boxes = nb[i];
        res->boxes.assign(boxes.begin() + (size_t)i * cap, boxes.begin() + (size_t)i * cap + nb[i]);
        res->labels.assign(labels.begin() + (size_t)i * N, labels.begin() + (size_t)(i + 1) * N);
        res->stats = stats[i];
        res->final_root = stats[i].final_root;
        Forest f;
        f.num_sets = (int)N - stats[i].n_merges;
        f.width = W;
        f.height = H;
        f.persp_mat = persp_mat;
        f.inv_mat = inv_mat;
        f.inv_mat_upper = inv_mat_upper;
        f.result = res;
        out.push_back(f);
    }
    return out;
}
