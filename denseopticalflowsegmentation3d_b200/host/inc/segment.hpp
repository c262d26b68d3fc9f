// The hot-path functions the reference defines in cpp/src/segment.cpp without declaring them in a header.
//   reference cpp/src/segment.cpp:20-32  diff
//   reference cpp/src/segment.cpp:34-72  get_segmented_array
//   reference cpp/src/segment.cpp:97-101 cvtColor + calcOpticalFlowFarneback call site -> dense_flow below
#ifndef DOFS3D_HOST_SEGMENT_HPP
#define DOFS3D_HOST_SEGMENT_HPP
#include <vector>

#include "graph.hpp"

extern bool debug;  // reference cpp/inc/segment.hpp:11

double diff(const cv::Mat& flow, int x1, int y1, int x2, int y2);

// Blurs `flow` IN PLACE (GaussianBlur sigma 3, segment.cpp:52 — the caller's Mat is modified, as in the
// reference), builds the graph and segments it.  neighbor other than 4 or 8 logs an error and uses 4 (segment.cpp:38-43).
Forest get_segmented_array(const cv::Mat& flow, const cv::Mat& bev, const cv::Matx33f& persp_mat, const cv::Matx33f& inv_mat,
                           const std::vector<cv::Matx33f>& inv_mat_upper, int neighbor = 8);

// cvtColor(BGR2GRAY) x2 + calcOpticalFlowFarneback(g1, g2, flow, 0.5, 3, 15, 3, 5, 1.2, 0) of main()
// (segment.cpp:97-101): two CV_8UC3 frames in, CV_32FC2 flow out.
cv::Mat dense_flow(const cv::Mat& frame1_bgr, const cv::Mat& frame2_bgr);

// Whole path for a video (main1, segment.cpp:209-269): frames[i], frames[i+1] -> one Forest per pair.  The clip is streamed
// through the device in chunks of at most chunk_pairs pairs (dofs3d_stream_*): the boundary frame of a chunk stays on the
// device (prev_frame, segment.cpp:268), uploads overlap the kernels of the previous chunk, memory is bounded by the chunk.
std::vector<Forest> process_video(const std::vector<cv::Mat>& frames_bgr, const cv::Matx33f& persp_mat,
                                  const cv::Matx33f& inv_mat, const std::vector<cv::Matx33f>& inv_mat_upper,
                                  int neighbor = 8, int chunk_pairs = 32);

// GPU the shim's contexts are created on (default: environment variable DOFS3D_DEVICE, else 0).
void set_device(int device);
#endif
