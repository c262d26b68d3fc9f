// C++ host shim: the reference's cpp/inc entry points (build_graph, segment_graph, Forest,
// get_segmented_array, get_bottom_variants, get_mat, get_mat_upper, get_intersect) implemented on top of
// the C ABI of libdofs3d.so (include/dofs3d.h).  Nothing here computes on the CPU except argument
// marshalling, the setup-sized get_intersect, and turning label images into std::set<int> pixel sets.
//
// Errors: the reference's hot path never throws; failures are sentinels (Solution() with errors -1,
// score -1).  Device failures have no sentinel in that vocabulary, so they throw std::runtime_error.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
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
    int w, h, n, neighbors, device;
    float mats[45];
    bool operator<(const CtxKey& o) const {
        if (std::tie(w, h, n, neighbors, device) != std::tie(o.w, o.h, o.n, o.neighbors, o.device))
            return std::tie(w, h, n, neighbors, device) < std::tie(o.w, o.h, o.n, o.neighbors, o.device);
        return std::memcmp(mats, o.mats, sizeof mats) < 0;
    }
};

// Contexts are reused across calls (allocation is the expensive part).  A context serves one call at a time: every
// shim function holds the entry's mutex for its whole duration, so concurrent callers with the same key queue up
// instead of racing on one stream and one set of buffers (the reference functions are re-entrant).
struct CtxEntry {
    dofs3d_ctx* c = nullptr;
    std::mutex mu;
};

std::mutex g_mu;
std::map<CtxKey, std::unique_ptr<CtxEntry>> g_ctx;
int g_device = -1;  // -1: DOFS3D_DEVICE from the environment, else 0

[[noreturn]] void fail(dofs3d_ctx* c, int rc, const char* what) {
    throw std::runtime_error(std::string(what) + ": dofs3d status " + std::to_string(rc) + " (" +
                             (c ? dofs3d_last_error(c) : "no context") + ")");
}

void fill_mats(dofs3d_params& p, const cv::Matx33f& persp, const cv::Matx33f& inv, const std::vector<cv::Matx33f>& up) {
    std::memcpy(p.persp, persp.val, sizeof p.persp);
    std::memcpy(p.inv, inv.val, sizeof p.inv);
    for (size_t c = 0; c < 3 && c < up.size(); ++c) std::memcpy(p.inv_upper[c], up[c].val, sizeof p.inv_upper[c]);
}

int device_index() {
    if (g_device >= 0) return g_device;
    const char* e = std::getenv("DOFS3D_DEVICE");
    return e ? std::atoi(e) : 0;
}

CtxEntry* context_for(int w, int h, int n, const dofs3d_params& p) {
    CtxKey k;
    k.w = w;
    k.h = h;
    k.n = n;
    k.neighbors = p.neighbors;
    std::memcpy(k.mats, p.persp, sizeof p.persp);
    std::memcpy(k.mats + 9, p.inv, sizeof p.inv);
    std::memcpy(k.mats + 18, p.inv_upper, sizeof p.inv_upper);
    k.device = device_index();
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_ctx.find(k);
    if (it != g_ctx.end()) return it->second.get();
    dofs3d_ctx* c = nullptr;
    int rc = dofs3d_create(&c, k.device, w, h, n, &p);
    if (rc != 0) {
        std::string msg = c ? dofs3d_last_error(c) : "no CUDA device";
        if (c) dofs3d_destroy(c);
        throw std::runtime_error("dofs3d_create failed (" + msg + "); there is no CPU fallback");
    }
    auto e = std::make_unique<CtxEntry>();
    e->c = c;
    CtxEntry* out = e.get();
    g_ctx[k] = std::move(e);
    return out;
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
    // segment_scores of the reference (graph.cpp:326): the score of the LATEST merge of every root whose get_score
    // was not -1, whatever the convexity / threshold gates said afterwards
    std::map<int, std::pair<uint32_t, double>> last_score;
    void fetch_scores(dofs3d_ctx* c, int pair_index) {
        int n = dofs3d_scored_merges(c, pair_index, 0, nullptr, nullptr, nullptr, nullptr);
        if (n < 0) fail(c, n, "Forest (scored merges)");
        std::vector<int32_t> root((size_t)n);
        std::vector<uint32_t> time((size_t)n);
        std::vector<double> score((size_t)n);
        if (n) n = dofs3d_scored_merges(c, pair_index, n, root.data(), time.data(), score.data(), nullptr);
        if (n < 0) fail(c, n, "Forest (scored merges)");
        for (int i = 0; i < n; ++i) {
            auto it = last_score.find(root[i]);
            if (it == last_score.end() || it->second.first < time[i]) last_score[root[i]] = {time[i], score[i]};
        }
    }
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
    CtxEntry* ent = context_for(width, height, 1, p);
    std::lock_guard<std::mutex> call(ent->mu);
    dofs3d_ctx* c = ent->c;
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
    CtxEntry* ent = context_for(W, H, 1, p);
    std::lock_guard<std::mutex> call(ent->mu);
    dofs3d_ctx* c = ent->c;
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
    res->fetch_scores(c, 0);
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
    // the device rebuilds the list from `flow`: a list that was re-weighted or re-ordered by the caller would be silently
    // ignored, so a few entries are checked against the reference's weight function (segment.cpp:20-32) and order
    check_flow(flow, "segment_graph");
    const long long probes[5] = {0, n / 4, n / 2, (3 * n) / 4, n - 1};
    for (long long i : probes) {
        const Edge& ed = graph_edges[(size_t)i];
        if (ed.start < 0 || ed.start >= W * H || ed.end < 0 || ed.end >= W * H ||
            !(diff(flow, ed.start % (int)W, ed.start / (int)W, ed.end % (int)W, ed.end / (int)W) == ed.weight) ||
            (i > 0 && graph_edges[(size_t)i - 1].weight > ed.weight))
            throw std::invalid_argument("segment_graph: graph_edges is not build_graph(flow) (weights or order differ)");
    }
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
struct Forest::Incremental {
    CtxEntry* ent = nullptr;
    dofs3d_forest* f = nullptr;
    ~Incremental() {
        if (f) {
            std::lock_guard<std::mutex> call(ent->mu);
            dofs3d_forest_destroy(f);
        }
    }
};

Forest::Forest(const cv::Mat& flow, const cv::Mat& bev_, const cv::Matx33f& persp, const cv::Matx33f& inv,
               const std::vector<cv::Matx33f>& upper, int min_move_)
    : num_sets(flow.rows * flow.cols), width(flow.cols), height(flow.rows), min_move(min_move_), bev(bev_), persp_mat(persp),
      inv_mat(inv), inv_mat_upper(upper) {
    check_flow(flow, "Forest");
    if (upper.size() < 3) throw std::invalid_argument("Forest: inv_mat_upper needs one matrix per class (3)");
    dofs3d_params p;
    dofs3d_default_params(&p);
    fill_mats(p, persp, inv, upper);
    auto inc = std::make_shared<Incremental>();
    inc->ent = context_for(width, height, 1, p);
    std::lock_guard<std::mutex> call(inc->ent->mu);
    int rc = dofs3d_forest_create(inc->ent->c, flow.ptr<float>(), &inc->f);
    if (rc != 0) fail(inc->ent->c, rc, "Forest");
    incremental = inc;
}

int Forest::find(int n) const {
    if (incremental) {
        std::lock_guard<std::mutex> call(incremental->ent->mu);
        int32_t r = 0;
        int rc = dofs3d_forest_find(incremental->f, n, &r);
        if (rc != 0) fail(incremental->ent->c, rc, "Forest::find");
        return r;
    }
    if (!result) throw std::logic_error("Forest::find on an empty forest");
    return result->final_root;
}

std::vector<std::pair<int, SegmentData>> Forest::get_best_segments_sparse() const {
    std::vector<std::pair<int, SegmentData>> out;
    if (incremental) {
        std::lock_guard<std::mutex> call(incremental->ent->mu);
        dofs3d_ctx* c = incremental->ent->c;
        std::vector<dofs3d_box> boxes(4096);
        int n = dofs3d_forest_boxes(incremental->f, (int)boxes.size(), boxes.data());
        if (n < 0) fail(c, n, "Forest::get_best_segments");
        for (int b = 0; b < n; ++b) {
            const dofs3d_box& bx = boxes[b];
            std::vector<int32_t> px((size_t)bx.size);
            int m = dofs3d_forest_pixels(incremental->f, bx.root, bx.size, px.data());
            if (m < 0) fail(c, m, "Forest::get_best_segments");
            out.emplace_back(bx.root, SegmentData(bx.score, std::set<int>(px.begin(), px.begin() + std::min(m, bx.size)),
                                                  solution_of(bx, true), bx.move));
        }
        return out;
    }
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

// graph.cpp:446-452 returns bboxes[node_id], and Forest::merge clears the box of every absorbed root (graph.cpp:207):
// after the full Kruskal pass only the final root still has one — the whole frame.  Every other node: empty.
// (The box a segment had at its best snapshot is SegmentData::sol / dofs3d_box::bbox; the state a node had when it was
// absorbed is dofs3d_node_state.)
std::vector<cv::Point2i> Forest::get_bounding_box(int node_id) const {
    if (incremental) {
        std::lock_guard<std::mutex> call(incremental->ent->mu);
        int32_t bb[4];
        int rc = dofs3d_forest_bbox(incremental->f, node_id, bb);
        if (rc < 0) fail(incremental->ent->c, rc, "Forest::get_bounding_box");
        if (rc == 0) return {};
        return {cv::Point2i(bb[0], bb[1]), cv::Point2i(bb[2], bb[3])};
    }
    if (!result) return {};
    if (node_id < 0 || node_id >= width * height) throw std::out_of_range("Forest::get_bounding_box: node id");
    if (node_id == result->final_root) return {cv::Point2i(0, 0), cv::Point2i(width - 1, height - 1)};
    return {};
}

// graph.cpp:386-389: segment_scores[node_id] — the score of the node's latest merge that produced a rectangle
// (graph.cpp:326, written before the convexity and threshold gates), 0.0 if there was none.
double Forest::get_segment_best_score(int node_id) const {
    if (incremental) {
        std::lock_guard<std::mutex> call(incremental->ent->mu);
        double sc = 0.0;
        int rc = dofs3d_forest_last_score(incremental->f, node_id, &sc);
        if (rc != 0) fail(incremental->ent->c, rc, "Forest::get_segment_best_score");
        return sc;
    }
    if (!result) return 0.0;
    auto it = result->last_score.find(node_id);
    return it == result->last_score.end() ? 0.0 : it->second.second;
}

int Forest::merge(int a, int b) {
    if (!incremental)
        throw std::logic_error("Forest::merge: this forest is the finished result of segment_graph (its merge loop ran on the "
                               "device as a whole); build one with Forest(flow, ...) to merge one call at a time");
    std::lock_guard<std::mutex> call(incremental->ent->mu);
    int32_t r = 0, ns = 0;
    int rc = dofs3d_forest_merge(incremental->f, a, b, &r);
    if (rc == 0) rc = dofs3d_forest_num_sets(incremental->f, &ns);
    if (rc != 0) fail(incremental->ent->c, rc, "Forest::merge");
    num_sets = ns;
    return r;
}

void Forest::new_merge(int a, int b, double score_threshold, int min_size, double, double) {
    if (!incremental)
        throw std::logic_error("Forest::new_merge: this forest is the finished result of segment_graph (its merge loop ran on "
                               "the device as a whole); build one with Forest(flow, ...) to merge one call at a time");
    std::lock_guard<std::mutex> call(incremental->ent->mu);
    int32_t ns = 0;
    int rc = dofs3d_forest_new_merge(incremental->f, a, b, score_threshold, min_size);
    if (rc == 0) rc = dofs3d_forest_num_sets(incremental->f, &ns);
    if (rc != 0) fail(incremental->ent->c, rc, "Forest::new_merge");
    num_sets = ns;
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
    CtxEntry* ent = context_for(64, 64, 1, p);
    std::lock_guard<std::mutex> call(ent->mu);
    dofs3d_ctx* c = ent->c;
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
    CtxEntry* ent = context_for(W, H, 1, p);
    std::lock_guard<std::mutex> call(ent->mu);
    dofs3d_ctx* c = ent->c;
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

void set_device(int device) { g_device = device; }

// main1 (segment.cpp:174-275) over the streaming entry points: the clip goes through the device in chunks of at most
// `chunk_pairs` pairs; the last frame of a chunk is carried on the device (prev_frame, segment.cpp:268), the upload of
// chunk k+1 runs under the kernels of chunk k, and device memory is bounded by the chunk size whatever the clip length.
std::vector<Forest> process_video(const std::vector<cv::Mat>& frames, const cv::Matx33f& persp_mat, const cv::Matx33f& inv_mat,
                                  const std::vector<cv::Matx33f>& inv_mat_upper, int neighbor, int chunk_pairs) {
    std::vector<Forest> out;
    if (frames.size() < 2) return out;
    if (neighbor != 4 && neighbor != 8) neighbor = 4;
    const int W = frames[0].cols, H = frames[0].rows, n = (int)frames.size() - 1;
    const size_t N = (size_t)W * H;
    const int CH = std::max(1, std::min(chunk_pairs, n));
    dofs3d_params p;
    dofs3d_default_params(&p);
    p.neighbors = neighbor;
    fill_mats(p, persp_mat, inv_mat, inv_mat_upper);
    CtxEntry* ent = context_for(W, H, CH, p);
    std::lock_guard<std::mutex> call(ent->mu);
    dofs3d_ctx* c = ent->c;
    for (const cv::Mat& f : frames)
        if (f.type() != CV_8UC3 || f.cols != W || f.rows != H || !f.isContinuous())
            throw std::invalid_argument("process_video: continuous CV_8UC3 frames of equal size expected");
    const int cap = 1024;
    struct Slot {  // two chunks are in flight: pinned staging for the frames going in and the results coming out
        uint8_t* bgr = nullptr;
        uint16_t* labels = nullptr;
        dofs3d_box* boxes = nullptr;
        int32_t* nb = nullptr;
        dofs3d_stats* stats = nullptr;
        int first_pair = 0;
    } slot[2];
    auto release = [&]() {
        for (Slot& sl : slot) {
            dofs3d_pinned_free(sl.bgr);
            dofs3d_pinned_free(sl.labels);
            dofs3d_pinned_free(sl.boxes);
            dofs3d_pinned_free(sl.nb);
            dofs3d_pinned_free(sl.stats);
        }
    };
    for (Slot& sl : slot) {
        sl.bgr = static_cast<uint8_t*>(dofs3d_pinned_alloc((size_t)(CH + 1) * N * 3));
        sl.labels = static_cast<uint16_t*>(dofs3d_pinned_alloc((size_t)CH * N * sizeof(uint16_t)));
        sl.boxes = static_cast<dofs3d_box*>(dofs3d_pinned_alloc((size_t)CH * cap * sizeof(dofs3d_box)));
        sl.nb = static_cast<int32_t*>(dofs3d_pinned_alloc((size_t)CH * sizeof(int32_t)));
        sl.stats = static_cast<dofs3d_stats*>(dofs3d_pinned_alloc((size_t)CH * sizeof(dofs3d_stats)));
        if (!sl.bgr || !sl.labels || !sl.boxes || !sl.nb || !sl.stats) {
            release();
            throw std::runtime_error("process_video: pinned host allocation failed");
        }
    }
    auto collect = [&](Slot& sl) {
        int got = 0;
        int rc = dofs3d_stream_collect(c, &got);
        if (rc != 0) {
            release();
            fail(c, rc, "process_video");
        }
        for (int i = 0; i < got; ++i) {
            auto res = std::make_shared<Forest::Result>();
            res->n_boxes = sl.nb[i];
            res->boxes.assign(sl.boxes + (size_t)i * cap, sl.boxes + (size_t)i * cap + sl.nb[i]);
            res->labels.resize(N);
            const uint16_t* lab = sl.labels + (size_t)i * N;
            for (size_t px = 0; px < N; ++px) res->labels[px] = lab[px] == 0xFFFF ? -1 : (int32_t)lab[px];
            res->stats = sl.stats[i];
            res->final_root = sl.stats[i].final_root;
            Forest f;
            f.num_sets = (int)N - sl.stats[i].n_merges;
            f.width = W;
            f.height = H;
            f.persp_mat = persp_mat;
            f.inv_mat = inv_mat;
            f.inv_mat_upper = inv_mat_upper;
            f.result = res;
            out.push_back(f);
        }
    };
    int rc = dofs3d_stream_begin(c);
    if (rc != 0) {
        release();
        fail(c, rc, "process_video");
    }
    int next_frame = 0, submitted = 0, collected = 0;
    while (next_frame <= n) {
        const bool first = next_frame == 0;
        const int nf = std::min(first ? CH + 1 : CH, n + 1 - next_frame);
        if (nf < 1 || (first && nf < 2)) break;
        Slot& sl = slot[submitted & 1];
        if (submitted - collected == 2) collect(slot[collected++ & 1]);
        for (int i = 0; i < nf; ++i) std::memcpy(sl.bgr + (size_t)i * N * 3, frames[next_frame + i].ptr<uint8_t>(), N * 3);
        dofs3d_outputs o;
        std::memset(&o, 0, sizeof o);
        o.label_format = DOFS3D_LABELS_U16;
        o.labels = sl.labels;
        o.boxes = sl.boxes;
        o.n_boxes = sl.nb;
        o.max_boxes = cap;
        o.stats = sl.stats;
        rc = dofs3d_stream_submit(c, sl.bgr, nf, &o);
        if (rc != 0) {
            release();
            fail(c, rc, "process_video");
        }
        ++submitted;
        next_frame += nf;
    }
    while (collected < submitted) collect(slot[collected++ & 1]);
    release();
    return out;
}
