// Streaming driver in the shape of the reference's video loop (cpp/src/segment.cpp:174-275, `main1`): consecutive
// frames in, one Forest per frame pair out — but a whole clip goes through the GPU in one batched call
// (process_video of the host shim) instead of pair by pair.  Frames come from a raw BGR file (W*H*3 bytes per frame)
// because OpenCV's imgcodecs/videoio are not part of this build; with OpenCV present the same code takes cv::VideoCapture
// frames.  Prints, per pair, what plot_best_segments_simple would draw (score > 0.7): class, score, yaw, lower face.
//
//   video_driver <frames.bgr> <width> <height> [min_score]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "graph.hpp"
#include "lifting_3d.hpp"
#include "segment.hpp"

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <frames.bgr> <width> <height> [min_score]\n", argv[0]);
        return 2;
    }
    const int W = std::atoi(argv[2]), H = std::atoi(argv[3]);
    const double min_score = argc > 4 ? std::atof(argv[4]) : 0.7;  // segment.cpp:166
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) {
        std::perror(argv[1]);
        return 2;
    }
    std::vector<cv::Mat> frames;
    for (;;) {
        cv::Mat fr(H, W, CV_8UC3);
        if (std::fread(fr.ptr<unsigned char>(), 1, (size_t)W * H * 3, f) != (size_t)W * H * 3) break;
        frames.push_back(fr);
    }
    std::fclose(f);
    if (frames.size() < 2) {
        std::fprintf(stderr, "need at least two frames\n");
        return 2;
    }
    auto mats = get_mat();                                                                   // segment.cpp:125
    std::vector<cv::Matx33f> upper = {get_mat_upper(0), get_mat_upper(1), get_mat_upper(2)};  // segment.cpp:128-133
    std::vector<Forest> forests = process_video(frames, mats.first, mats.second, upper, 8);
    for (size_t i = 0; i < forests.size(); ++i) {
        int drawn = 0;
        for (auto& kv : forests[i].get_best_segments_sparse()) {
            const SegmentData& sd = kv.second;
            if (!(sd.score > min_score)) continue;  // draw.cpp:127
            ++drawn;
            std::printf("pair %zu root %d cls %d score %.6f move %.4f yaw %.6f size %zu lower", i, kv.first, sd.sol.cls, sd.score,
                        sd.move, sd.sol.orient, sd.seg.size());
            for (const auto& p : sd.sol.lower_face) std::printf(" (%.2f,%.2f)", p.x, p.y);
            std::printf("\n");
        }
        std::printf("pair %zu: %d boxes drawn\n", i, drawn);
    }
    return 0;
}
