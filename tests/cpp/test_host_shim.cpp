// Exercises the C++ host shim (denseopticalflowsegmentation3d_b200/host) the way the reference's own code does:
//   * the three gtest cases of cpp/tests/test_liftig_3d.cpp (:69-89, :179-227), same inputs and tolerances;
//   * the call sequence of main() (cpp/src/segment.cpp:125-166): get_mat / get_mat_upper ->
//     get_segmented_array(flow, bev, persp, inv, upper, 8) -> forest.get_best_segments();
//   * build_graph + segment_graph called directly, as get_segmented_array does internally (:54-63).
// Input: a raw file of W*H*2 floats (unblurred flow).  Prints one line per kept segment for the Python test
// to compare with the oracle.  Exit code 0 = all assertions held.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

#include "graph.hpp"
#include "lifting_3d.hpp"
#include "segment.hpp"

#define CHECK(c)                                                       \
    do {                                                               \
        if (!(c)) {                                                    \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                  \
        }                                                              \
    } while (0)

static bool near(double a, double b, double tol) { return std::fabs(a - b) <= tol; }

int main(int argc, char** argv) {
    // IntersectionTest.IntersectionExists / IntersectionDoesNotExist
    cv::Point2f p = get_intersect({1.f, 1.f}, {4.f, 4.f}, {1.f, 8.f}, {2.f, 4.f});
    CHECK(near(p.x, 2.4, 1e-2) && near(p.y, 2.4, 1e-2));
    p = get_intersect({1.f, 1.f}, {1.f, 2.f}, {3.f, 3.f}, {3.f, 4.f});
    CHECK(std::isnan(p.x) && std::isnan(p.y));

    // GetBottomVariantsTest.Test1
    {
        cv::Point2f dir(2.5470946f, 1.9316475f);
        std::vector<cv::Point2i> box = {cv::Point2i(375, 92), cv::Point2i(576, 286)};
        cv::Matx33f mat(20.1377838f, -13.4744920f, 402.174272f, 5.11635077f, 800.335022f, -62251.3321f, 0.000393565444f,
                        0.0397205947f, 1.0f);
        cv::Matx33f inv(0.202212552f, 0.00181942728f, 31.9370859f, -0.00182975914f, 0.00123437589f, 77.5774258f,
                        -6.90475148e-06f, -4.97462083e-05f, 1.0f);
        cv::Matx33f up(0.203701900f, 0.00169508037f, 32.3672674f, 0.0f, 0.00146371164f, 29.6614710f, 0.0f, -5.01704822e-05f, 1.0f);
        Solution s = get_bottom_variants(dir, box, mat, inv, up, 2);
        const double ps_bev[4][2] = {{327.809749190909, 13476.772230116465}, {1398.2414179174136, 2769.3562851313454},
                                     {2204.8576245955073, 2935.1653236816246}, {647.3324083001849, 13473.77576520306}};
        const double lower[4][2] = {{385.305, 286.0}, {375.0, 269.47327}, {555.45557, 270.2111}, {576.0, 286.92706}};
        const double upper[4][2] = {{385.305, 99.43571}, {375.0, 92.75792}, {555.45557, 92.0}, {576.0, 98.69487}};
        CHECK(s.cls == 2 && s.ps_bev.size() == 4 && s.lower_face.size() == 4 && s.upper_face.size() == 4 && s.rectangle.size() == 4);
        for (int i = 0; i < 4; ++i) {
            CHECK(near(s.ps_bev[i].x, ps_bev[i][0], 1e-1) && near(s.ps_bev[i].y, ps_bev[i][1], 1e-1));
            CHECK(near(s.lower_face[i].x, lower[i][0], 1e-1) && near(s.lower_face[i].y, lower[i][1], 1e-1));
            CHECK(near(s.upper_face[i].x, upper[i][0], 1e-1) && near(s.upper_face[i].y, upper[i][1], 1e-1));
        }
        CHECK(near(s.w_error, 0.5987518562843858, 1e-1) && near(s.h_error, 0.7156805292391223, 1e-1));
        CHECK(near(s.orient, -1.6261444189491607, 1e-1));
    }
    if (argc < 4) {
        std::printf("kat ok\n");
        return 0;
    }
    const int W = std::atoi(argv[2]), H = std::atoi(argv[3]);
    cv::Mat flow(H, W, CV_32FC2);
    FILE* f = std::fopen(argv[1], "rb");
    CHECK(f != nullptr);
    CHECK(std::fread(flow.ptr<float>(), sizeof(float), (size_t)W * H * 2, f) == (size_t)W * H * 2);
    std::fclose(f);

    // main(): segment.cpp:125-166
    auto mats = get_mat();
    std::vector<cv::Matx33f> upper = {get_mat_upper(0), get_mat_upper(1), get_mat_upper(2)};
    cv::Mat bev;
    cv::Mat original = flow.clone();
    Forest forest = get_segmented_array(flow, bev, mats.first, mats.second, upper, 8);
    // the caller's flow was blurred in place (segment.cpp:52)
    bool changed = false;
    for (size_t i = 0; i < (size_t)W * H * 2 && !changed; ++i) changed = flow.ptr<float>()[i] != original.ptr<float>()[i];
    CHECK(changed);
    CHECK(forest.num_sets == 1 && forest.width == W && forest.height == H);
    std::vector<SegmentData> best = forest.get_best_segments();
    CHECK((int)best.size() == W * H);
    int kept = 0;
    for (int root = 0; root < W * H; ++root) {
        const SegmentData& sd = best[root];
        if (sd.score < 0) continue;
        ++kept;
        CHECK(sd.seg.count(root) == 1);  // every root pixel is a member of its own set
        long long h = 0;
        for (int px : sd.seg) h = (h * 1000003LL + px) % 2147483647LL;
        // segment_scores[root] (graph.cpp:386-389) is the LATEST scored merge of the root, written before the convexity and
        // threshold gates: it can be lower OR higher than the kept score; the Python test compares it with the unchanged
        // reference
        const double last = forest.get_segment_best_score(root);
        CHECK(last > 0.0);
        std::printf("segment root=%d size=%zu hash=%lld cls=%d score=%.17g move=%.17g orient=%.17g last=%.17g\n", root,
                    sd.seg.size(), h, sd.sol.cls, sd.score, sd.move, sd.sol.orient, last);
        // every absorbed root has lost its box (graph.cpp:207)
        CHECK(forest.get_bounding_box(root).empty() == (root != forest.find(0)));
    }
    CHECK(forest.find(0) == forest.find(W * H - 1));
    {
        auto bb = forest.get_bounding_box(forest.find(0));
        CHECK(bb.size() == 2 && bb[0].x == 0 && bb[0].y == 0 && bb[1].x == W - 1 && bb[1].y == H - 1);
        CHECK(forest.get_segment_best_score(0) == 0.0 || forest.get_segment_best_score(0) > 0.0);
    }
    // a re-weighted edge list is refused, not silently ignored
    {
        std::vector<Edge> tampered = build_graph(flow, W, H, diff, true);
        tampered[tampered.size() / 2].weight += 1.0;
        bool threw = false;
        try {
            segment_graph(flow, tampered, bev, mats.first, mats.second, upper);
        } catch (const std::invalid_argument&) {
            threw = true;
        }
        CHECK(threw);
    }

    // build_graph + segment_graph on the (now blurred) flow must give the same forest
    std::vector<Edge> edges = build_graph(flow, W, H, diff, true);
    CHECK((long long)edges.size() == 4LL * W * H - 3 * W - 3 * H + 2);
    for (size_t i = 1; i < edges.size(); ++i) CHECK(edges[i - 1].weight <= edges[i].weight);
    Forest again = segment_graph(flow, edges, bev, mats.first, mats.second, upper);
    auto a = forest.get_best_segments_sparse(), b = again.get_best_segments_sparse();
    CHECK(a.size() == b.size() && (int)a.size() == kept);
    for (size_t i = 0; i < a.size(); ++i) CHECK(a[i].first == b[i].first && a[i].second.seg == b[i].second.seg);
    // The reference's own Forest, driven one merge at a time (graph.cpp:520-531 is this loop) on a crop of the field:
    // it must end with the history the whole-pass segment_graph returns for the same crop.
    {
        const int cw = 72, ch = 56, x0 = W / 2 - cw / 2, y0 = H / 2 - ch / 2;
        cv::Mat crop(ch, cw, CV_32FC2);
        for (int y = 0; y < ch; ++y)
            for (int x = 0; x < cw; ++x) {
                const cv::Point2f v = flow.at<cv::Point2f>(y0 + y, x0 + x);
                crop.at<cv::Point2f>(y, x) = cv::Point2f(v.x * 3.0f, v.y * 3.0f);
            }
        std::vector<Edge> ce = build_graph(crop, cw, ch, diff, true);
        Forest whole = segment_graph(crop, ce, bev, mats.first, mats.second, upper);
        Forest step(crop, bev, mats.first, mats.second, upper);
        CHECK(step.num_sets == cw * ch);
        for (const Edge& ed : ce) {
            const int ra = step.find(ed.start), rb = step.find(ed.end);
            if (ra != rb) step.new_merge(ra, rb);
        }
        CHECK(step.num_sets == 1 && step.find(0) == whole.find(0));
        auto hs = step.get_best_segments_sparse(), hw = whole.get_best_segments_sparse();
        CHECK(hs.size() == hw.size());
        for (size_t i = 0; i < hs.size(); ++i) {
            CHECK(hs[i].first == hw[i].first && hs[i].second.seg == hw[i].second.seg);
            CHECK(near(hs[i].second.score, hw[i].second.score, 1e-12) && hs[i].second.sol.cls == hw[i].second.sol.cls);
            CHECK(step.get_segment_best_score(hs[i].first) == whole.get_segment_best_score(hw[i].first));
        }
        std::printf("incremental forest: %zu kept segments equal the whole-pass result\n", hs.size());
        bool threw = false;
        try {
            whole.merge(0, 1);
        } catch (const std::logic_error&) {
            threw = true;
        }
        CHECK(threw);
    }
    std::printf("edges=%zu kept=%d\n", edges.size(), kept);
    return 0;
}
