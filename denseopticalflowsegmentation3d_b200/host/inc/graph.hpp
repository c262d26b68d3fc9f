// Drop-in replacement for the reference's cpp/inc/graph.hpp (same names, argument meaning and error
// behaviour for the hot path) on top of the C ABI include/dofs3d.h.  The Kruskal loop, the per-merge
// gates and the scoring run on the GPU as one batch; this header only carries the results.
//
//   reference cpp/inc/graph.hpp:13-17   Edge
//   reference cpp/inc/graph.hpp:20-23   DiffFunction, build_graph
//   reference cpp/inc/graph.hpp:25-57   Solution, SegmentData
//   reference cpp/inc/graph.hpp:72-114  Forest
//   reference cpp/inc/graph.hpp:120-122 segment_graph
#ifndef DOFS3D_HOST_GRAPH_HPP
#define DOFS3D_HOST_GRAPH_HPP

#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#else
#include "cv_min.hpp"
#endif
#include <memory>
#include <set>
#include <vector>

struct Edge {
    int start;
    int end;
    double weight;
};

typedef double (*DiffFunction)(const cv::Mat& flow, int x0, int y0, int x1, int y1);

// Sorted edge list of the flow graph: ascending weight, equal weights in insertion order — what the
// reference's std::multiset yields (graph.cpp:51-103).  The weights are computed on the device with the
// arithmetic of the reference's `diff` (segment.cpp:20-32); `diff` itself is only spot-checked against
// the device result (a mismatch throws std::invalid_argument): arbitrary host callbacks cannot run there.
std::vector<Edge> build_graph(const cv::Mat& img, int width, int height, const DiffFunction& diff,
                              bool neighborhood_8 = false);

class Solution {
public:
    int cls;
    std::vector<cv::Point2f> ps_bev;
    std::vector<cv::Point2f> lower_face;
    std::vector<cv::Point2f> upper_face;
    std::vector<cv::Point2f> rectangle;
    double w_error;
    double h_error;
    double orient;
    Solution() : cls(0), w_error(-1.0), h_error(-1.0), orient(0.0) {}
    Solution(int cls, const std::vector<cv::Point2f>& ps_bev, const std::vector<cv::Point2f>& lower_face,
             const std::vector<cv::Point2f>& upper_face, const std::vector<cv::Point2f>& rectangle, double w_error,
             double h_error, double orient)
        : cls(cls), ps_bev(ps_bev), lower_face(lower_face), upper_face(upper_face), rectangle(rectangle),
          w_error(w_error), h_error(h_error), orient(orient) {}
};

class SegmentData {
public:
    double score;
    std::set<int> seg;
    Solution sol;
    double move;
    SegmentData(const double score, const std::set<int>& seg, const Solution& sol, const double move)
        : score(score), seg(seg), sol(sol), move(move) {}
    SegmentData() : score(-1), move(0) {}
};

class Forest {
public:
    int num_sets;
    int width;
    int height;
    int min_move;
    cv::Mat bev;
    cv::Matx33f persp_mat;
    cv::Matx33f inv_mat;
    std::vector<cv::Matx33f> inv_mat_upper;

    Forest() : num_sets(0), width(0), height(0), min_move(5) {}
    // The reference's own constructor (graph.cpp:129-148): N singleton sets over `flow`, to be merged one call at a
    // time with find / merge / new_merge below — every call is a one-thread kernel over device state and a wait.
    // segment_graph / get_segmented_array run the whole loop on the device at once and are the path for throughput.
    Forest(const cv::Mat& flow, const cv::Mat& bev, const cv::Matx33f& persp_mat, const cv::Matx33f& inv_mat,
           const std::vector<cv::Matx33f>& inv_mat_upper, int min_move = 5);

    // Root of n's set (graph.cpp:150-168).  For a forest returned by segment_graph: the final forest, one set.
    int find(int n) const;
    // The whole history vector, index = root id, score == -1 where empty (graph.cpp:391-429).
    std::vector<SegmentData> get_best_segments();
    // Only the kept entries, as (root id, data) in ascending root order — what callers iterate for.
    std::vector<std::pair<int, SegmentData>> get_best_segments_sparse() const;
    // bboxes[node_id] of the finished forest (graph.cpp:446-452): Forest::merge clears the box of every absorbed root
    // (graph.cpp:207), so this is the whole frame for the final root and empty for every other node.
    std::vector<cv::Point2i> get_bounding_box(int node_id) const;
    // segment_scores[node_id] (graph.cpp:386-389): the score of the node's LATEST merge whose get_score produced a
    // rectangle, written before the convexity and threshold gates (graph.cpp:326); 0.0 if it never had one.
    // (Forests returned by process_video carry no per-merge scores: 0.0.)
    double get_segment_best_score(int node_id) const;
    // merge / new_merge (graph.cpp:170-218, 272-384): on a forest built with the constructor above; a forest returned by
    // segment_graph is finished (its loop ran on the device as a whole) and throws std::logic_error.  min_move and
    // min_convexity are accepted and ignored like in the reference (its gates use the row-adaptive motion threshold
    // and the per-class convexity bounds, graph.cpp:296, 328-339).
    int merge(int a, int b);
    void new_merge(int a, int b, double score_threshold = 0.3, int min_size = 500, double min_move = 1.0,
                   double min_convexity = 1.0 / 2.0);

    struct Result;  // labels + boxes returned by dofs3d_segment
    std::shared_ptr<Result> result;
    struct Incremental;  // device forest driven one call at a time
    std::shared_ptr<Incremental> incremental;
};

// Kruskal / union-find clustering with per-merge lifting (graph.cpp:503-536).  `graph_edges` is what
// build_graph returned for the same `flow`; the device rebuilds it from `flow`, so only its size is checked.
Forest segment_graph(const cv::Mat& flow, const std::vector<Edge>& graph_edges, const cv::Mat& bev,
                     const cv::Matx33f& persp_mat, const cv::Matx33f& inv_mat,
                     const std::vector<cv::Matx33f>& inv_mat_upper);

#endif
