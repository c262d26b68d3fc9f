// TEST INFRASTRUCTURE (oracle) — not part of the product.
//
// C-callable driver around the UNCHANGED reference sources
//   /root/reference/cpp/src/graph.cpp  and  /root/reference/cpp/src/lifting_3d.cpp
// which oracle/Makefile compiles in place (never copied into this repo) against
// oracle/cv_compat (OpenCV stand-in) and the spdlog headers bundled with the
// image.  Output: oracle/_ref/libdofs3d_ref.so (git-ignored, travels to the GPU
// box).  It provides what the reference executables provide to the library:
//   * the global `logger`                        (segment.cpp:18, graph.cpp:15)
//   * the edge-weight function `diff`            (restated from segment.cpp:20-32)
//   * `generate_image`                           (draw.cpp:12-46; only reached from the
//                                                 one-shot debug block graph.cpp:357-371)
// and the order of calls of get_segmented_array minus the blur
// (segment.cpp:54-63: build_graph -> segment_graph).
#include <chrono>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include <spdlog/sinks/base_sink.h>
#include <spdlog/sinks/null_sink.h>
#include <spdlog/spdlog.h>

#include "graph.hpp"
#include "lifting_3d.hpp"

std::shared_ptr<spdlog::logger> logger;

namespace {

// Counts log lines by their leading words: a cheap probe of how often each
// gate of Forest::new_merge (graph.cpp:277-375) fires.
class counting_sink : public spdlog::sinks::base_sink<std::mutex> {
public:
    std::map<std::string, long> counts;

protected:
    void sink_it_(const spdlog::details::log_msg& msg) override {
        std::string s(msg.payload.data(), msg.payload.size());
        size_t cut = s.find_first_of("0123456789:{(");
        if (cut != std::string::npos) s.resize(cut);
        while (!s.empty() && s.back() == ' ') s.pop_back();
        counts[s] += 1;
    }
    void flush_() override {}
};

std::shared_ptr<counting_sink> g_counter;

void ensure_logger() {
    if (logger) return;
    auto sink = std::make_shared<spdlog::sinks::null_sink_mt>();
    logger = std::make_shared<spdlog::logger>("seg", sink);
    logger->set_level(spdlog::level::off);
}

cv::Matx33f to_matx(const float* m) { return cv::Matx33f(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]); }

void put_points(const std::vector<cv::Point2f>& v, float* out) {
    for (int i = 0; i < 4; ++i) {
        out[2 * i] = i < (int)v.size() ? v[i].x : NAN;
        out[2 * i + 1] = i < (int)v.size() ? v[i].y : NAN;
    }
}

struct RefResult {
    std::vector<SegmentData> history;  // Forest::get_best_segments(), all N entries
    std::vector<int> kept;             // indices with score >= 0
    std::vector<double> last_score;    // Forest::get_segment_best_score(node), all N nodes (graph.cpp:386-389)
    std::vector<int32_t> node_bbox;    // Forest::get_bounding_box(node) as xmin,ymin,xmax,ymax, all N nodes (graph.cpp:446-452)
    double t_build = 0, t_segment = 0;
    long n_edges = 0;
    int num_sets = 0;
};

}  // namespace

// Euclidean distance between two flow vectors: component differences taken in
// float, squared and summed in double (segment.cpp:20-32).
double diff(const cv::Mat& flow, int x1, int y1, int x2, int y2) {
    const cv::Point2f a = flow.at<cv::Point2f>(y1, x1);
    const cv::Point2f b = flow.at<cv::Point2f>(y2, x2);
    const double dx = a.x - b.x;
    const double dy = a.y - b.y;
    return std::sqrt(dx * dx + dy * dy);
}

// draw.cpp is not compiled (it needs highgui); the debug block only displays the image.
cv::Mat generate_image(const Forest&, int, int) { return cv::Mat(); }

extern "C" {

struct ref_solution {
    int cls;
    int has_rectangle;
    float ps_bev[8], lower_face[8], upper_face[8], rectangle[8];
    double w_error, h_error, orient;
};

// mode 0: logging off; mode 1: count info/warn lines by prefix (slow).
void ref_set_counting(int mode) {
    if (mode) {
        g_counter = std::make_shared<counting_sink>();
        logger = std::make_shared<spdlog::logger>("seg", g_counter);
        logger->set_level(spdlog::level::info);
    } else {
        g_counter.reset();
        logger.reset();
        ensure_logger();
    }
}

// Writes "prefix=count\n" lines into buf; returns bytes needed.
int ref_get_counts(char* buf, int cap) {
    std::string s;
    if (g_counter)
        for (auto& kv : g_counter->counts) s += kv.first + "=" + std::to_string(kv.second) + "\n";
    if (buf && cap > 0) {
        std::strncpy(buf, s.c_str(), cap - 1);
        buf[cap - 1] = 0;
    }
    return (int)s.size() + 1;
}

// get_mat / get_mat_upper (lifting_3d.cpp:482-514, 441-480)
void ref_get_mats(float* persp, float* inv, float* upper /* 3x9 */) {
    ensure_logger();
    auto pr = get_mat();
    std::memcpy(persp, pr.first.val, 9 * sizeof(float));
    std::memcpy(inv, pr.second.val, 9 * sizeof(float));
    for (int c = 0; c < 3; ++c) {
        cv::Matx33f u = get_mat_upper(c);
        std::memcpy(upper + 9 * c, u.val, 9 * sizeof(float));
    }
}

// get_intersect (lifting_3d.cpp:63-88)
void ref_get_intersect(const float* a, const float* b, const float* c, const float* d, float* out) {
    ensure_logger();
    cv::Point2f p = get_intersect(cv::Point2f(a[0], a[1]), cv::Point2f(b[0], b[1]), cv::Point2f(c[0], c[1]),
                                  cv::Point2f(d[0], d[1]));
    out[0] = p.x;
    out[1] = p.y;
}

static void fill_solution(const Solution& s, ref_solution* out) {
    out->cls = s.cls;
    out->has_rectangle = s.rectangle.empty() ? 0 : 1;
    put_points(s.ps_bev, out->ps_bev);
    put_points(s.lower_face, out->lower_face);
    put_points(s.upper_face, out->upper_face);
    put_points(s.rectangle, out->rectangle);
    out->w_error = s.w_error;
    out->h_error = s.h_error;
    out->orient = s.orient;
}

// get_bottom_variants (lifting_3d.cpp:350-439); box = xmin,ymin,xmax,ymax
void ref_get_bottom_variants(const float* dir, const int* box, const float* mat, const float* inv_mat,
                             const float* inv_upper, int cls, ref_solution* out) {
    ensure_logger();
    std::vector<cv::Point2i> box_2d = {cv::Point2i(box[0], box[1]), cv::Point2i(box[2], box[3])};
    Solution s = get_bottom_variants(cv::Point2f(dir[0], dir[1]), box_2d, to_matx(mat), to_matx(inv_mat),
                                     to_matx(inv_upper), cls);
    fill_solution(s, out);
}

// build_graph (graph.cpp:51-103) on an interleaved H x W x 2 float flow field.
// Outputs must hold 4*W*H entries; returns the number of edges.
long ref_build_graph(const float* flow, int width, int height, int neighbors8, int32_t* start, int32_t* end,
                     double* weight) {
    ensure_logger();
    cv::Mat f(height, width, CV_32FC2, const_cast<float*>(flow));
    std::vector<Edge> edges = build_graph(f, width, height, diff, neighbors8 != 0);
    for (size_t i = 0; i < edges.size(); ++i) {
        start[i] = edges[i].start;
        end[i] = edges[i].end;
        weight[i] = edges[i].weight;
    }
    return (long)edges.size();
}

// get_segmented_array minus the blur (segment.cpp:54-63): build_graph + segment_graph.
void* ref_segment(const float* flow_blurred, int width, int height, int neighbors, const float* persp,
                  const float* inv, const float* upper /* 3x9 */) {
    ensure_logger();
    cv::Mat f(height, width, CV_32FC2, const_cast<float*>(flow_blurred));
    std::vector<cv::Matx33f> up = {to_matx(upper), to_matx(upper + 9), to_matx(upper + 18)};
    auto* res = new RefResult();
    auto t0 = std::chrono::steady_clock::now();
    const std::vector<Edge> edges = build_graph(f, width, height, diff, neighbors == 8);
    auto t1 = std::chrono::steady_clock::now();
    Forest forest = segment_graph(f, edges, cv::Mat(), to_matx(persp), to_matx(inv), up);
    auto t2 = std::chrono::steady_clock::now();
    res->t_build = std::chrono::duration<double>(t1 - t0).count();
    res->t_segment = std::chrono::duration<double>(t2 - t1).count();
    res->n_edges = (long)edges.size();
    res->num_sets = forest.num_sets;
    res->history = forest.get_best_segments();
    for (size_t i = 0; i < res->history.size(); ++i)
        if (res->history[i].score >= 0) res->kept.push_back((int)i);
    const int n_nodes = width * height;
    res->last_score.resize(n_nodes);
    res->node_bbox.resize((size_t)4 * n_nodes);
    for (int i = 0; i < n_nodes; ++i) {
        res->last_score[i] = forest.get_segment_best_score(i);
        const std::vector<cv::Point2i> bb = forest.get_bounding_box(i);  // empty for every absorbed node (graph.cpp:207)
        const bool has = bb.size() == 2;
        res->node_bbox[4 * (size_t)i + 0] = has ? bb[0].x : -1;
        res->node_bbox[4 * (size_t)i + 1] = has ? bb[0].y : -1;
        res->node_bbox[4 * (size_t)i + 2] = has ? bb[1].x : -1;
        res->node_bbox[4 * (size_t)i + 3] = has ? bb[1].y : -1;
    }
    return res;
}

// the two per-node accessors of the finished forest, for all N nodes: score_out[N], bbox_out[N][4]
void ref_result_nodes(void* h, double* score_out, int32_t* bbox_out) {
    auto* r = static_cast<RefResult*>(h);
    if (score_out) std::memcpy(score_out, r->last_score.data(), r->last_score.size() * sizeof(double));
    if (bbox_out) std::memcpy(bbox_out, r->node_bbox.data(), r->node_bbox.size() * sizeof(int32_t));
}

int ref_result_count(void* h) { return (int)static_cast<RefResult*>(h)->kept.size(); }
long ref_result_edges(void* h) { return static_cast<RefResult*>(h)->n_edges; }
int ref_result_num_sets(void* h) { return static_cast<RefResult*>(h)->num_sets; }
void ref_result_times(void* h, double* t_build, double* t_segment) {
    *t_build = static_cast<RefResult*>(h)->t_build;
    *t_segment = static_cast<RefResult*>(h)->t_segment;
}

// entry i (ascending root id): returns the segment size; fills root, score, move, solution.
int ref_result_entry(void* h, int i, int* root, double* score, double* move, ref_solution* sol) {
    auto* r = static_cast<RefResult*>(h);
    const SegmentData& d = r->history[r->kept[i]];
    *root = r->kept[i];
    *score = d.score;
    *move = d.move;
    fill_solution(d.sol, sol);
    return (int)d.seg.size();
}

// sorted pixel ids (y*W+x) of entry i
void ref_result_pixels(void* h, int i, int32_t* out) {
    auto* r = static_cast<RefResult*>(h);
    const SegmentData& d = r->history[r->kept[i]];
    size_t k = 0;
    for (int p : d.seg) out[k++] = p;
}

void ref_result_free(void* h) { delete static_cast<RefResult*>(h); }

}  // extern "C"
