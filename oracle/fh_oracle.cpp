// TEST INFRASTRUCTURE (oracle) — not part of the product.
//
// CPU restatement of the reference's Felzenszwalb-style flow segmentation, the Python twin of the hot path
// (/root/reference/graph.py): build_graph (:77-96), segment_graph_flow (:156-177) = the adaptive-threshold Kruskal loop
// (:163-172), remove_small_components (:98-106) and merge_components (:108-130), with Forest.find / Forest.merge
// (:27-64) and main.py's diff (:310-312) and threshold (:314-315).
//
// Arithmetic follows what the reference computes under NumPy 2 (NEP 50 promotion; the reference pins no version) when
// `flow` is a float32 array:
//   * edge weight = np.sqrt(np.sum((img[p] - img[q]) ** 2)): every step in float32;
//   * threshold[root] = weight + K / size: K / size in double (Python floats), added to the float32 weight as a float32;
//     the initial threshold K / 1 stays a Python float (compared exactly);
//   * Node.color = (size_a * color_a + size_b * color_b) / (size_a + size_b): float32 products, sum and true division.
// sorted() is stable, so edges of equal weight keep build_graph's insertion order.
// The array is addressed as img[x][y] by the reference (graph.py:19, main.py:311); callers pass it so that x is the
// fastest-varying pixel index of node id = y * width + x (see oracle/fh.py).
//
// Pinned by tests/test_fh_oracle.py against outputs of the reference's own Python, generated in the authoring container
// by tools/make_golden_fh.py (tests/golden/fh_*.npz).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace {
struct Forest {
    std::vector<int> parent, rank, size;
    std::vector<float> color;  // 2 per node
    explicit Forest(int n, const float* flow) : parent(n), rank(n, 0), size(n, 1), color(flow, flow + 2 * (size_t)n) {
        std::iota(parent.begin(), parent.end(), 0);
    }
    int find(int n) {
        int t = n;
        while (t != parent[t]) t = parent[t];
        parent[n] = t;  // graph.py:31 (only n itself is re-pointed)
        return t;
    }
    void merge(int a, int b) {  // graph.py:44-61
        const int keep = rank[a] > rank[b] ? a : b, gone = rank[a] > rank[b] ? b : a;
        parent[gone] = keep;
        const float sa = (float)size[a], sb = (float)size[b];
        const float tot = (float)(size[a] + size[b]);
        for (int c = 0; c < 2; ++c) {
            const float pa = sa * color[2 * a + c], pb = sb * color[2 * b + c];
            color[2 * keep + c] = (pa + pb) / tot;
        }
        size[keep] = size[a] + size[b];
        if (!(rank[a] > rank[b]) && rank[a] == rank[b]) rank[b] += 1;
    }
};
}  // namespace

extern "C" {

// flow: [height][width][2] float32 with node id = y * width + x.  labels_out[id] = Forest.find(id) after the three
// passes.  Returns the number of components.  stage: 1 = after the threshold loop, 2 = after remove_small_components,
// 3 = after merge_components (the reference's segment_graph_flow).
int fh_segment_flow(const float* flow, int width, int height, int neighbors8, double K, int min_size, double flow_dist,
                    double edge_dist, int stage, int32_t* labels_out) {
    const int n = width * height;
    struct E {
        int a, b;
        float w;
    };
    std::vector<E> edges;
    edges.reserve((size_t)4 * n);
    auto w_of = [&](int p, int q) {
        const float dx = flow[2 * p] - flow[2 * q], dy = flow[2 * p + 1] - flow[2 * q + 1];
        return std::sqrt(dx * dx + dy * dy);  // float overloads: every step rounds to float32
    };
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            const int p = y * width + x;
            if (x > 0) edges.push_back({p, p - 1, w_of(p, p - 1)});
            if (y > 0) edges.push_back({p, p - width, w_of(p, p - width)});
            if (neighbors8) {
                if (x > 0 && y > 0) edges.push_back({p, p - width - 1, w_of(p, p - width - 1)});
                if (x > 0 && y < height - 1) edges.push_back({p, p + width - 1, w_of(p, p + width - 1)});
            }
        }
    std::stable_sort(edges.begin(), edges.end(), [](const E& l, const E& r) { return l.w < r.w; });
    Forest f(n, flow);
    // thresholds: a Python float until the first update, a float32 afterwards; both compare exactly as doubles
    std::vector<double> thr((size_t)n, K / 1.0);
    for (const E& e : edges) {
        const int a = f.find(e.a), b = f.find(e.b);
        if (a != b && (double)e.w <= thr[a] && (double)e.w <= thr[b]) {
            f.merge(a, b);
            const int r = f.find(a);
            thr[r] = (double)(e.w + (float)(K * 1.0 / f.size[r]));
        }
    }
    if (stage >= 2)
        for (const E& e : edges) {
            const int a = f.find(e.a), b = f.find(e.b);
            if (a != b && (f.size[a] < min_size || f.size[b] < min_size)) f.merge(a, b);
        }
    if (stage >= 3)
        for (const E& e : edges) {
            const int a = f.find(e.a), b = f.find(e.b);
            if (a == b) continue;
            const float dx = f.color[2 * a] - f.color[2 * b], dy = f.color[2 * a + 1] - f.color[2 * b + 1];
            const float d = std::sqrt(dx * dx + dy * dy);
            if ((double)d < flow_dist && (double)e.w < edge_dist) f.merge(a, b);
        }
    int comps = 0;
    for (int i = 0; i < n; ++i) {
        int t = i;
        while (t != f.parent[t]) t = f.parent[t];
        labels_out[i] = t;
        comps += t == i;
    }
    return comps;
}

}  // extern "C"
