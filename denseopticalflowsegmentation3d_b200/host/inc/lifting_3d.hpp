// Drop-in replacement for the reference's cpp/inc/lifting_3d.hpp (hot-path subset) over include/dofs3d.h.
//   reference cpp/inc/lifting_3d.hpp:13-16  get_bottom_variants
//   reference cpp/inc/lifting_3d.hpp:17-18  get_mat, get_mat_upper
//   reference cpp/inc/lifting_3d.hpp:26     get_intersect
#ifndef DOFS3D_HOST_LIFTING_3D_HPP
#define DOFS3D_HOST_LIFTING_3D_HPP
#include <utility>
#include <vector>

#include "graph.hpp"

// One (direction, box, class) lifting problem on the device (lifting_3d.cpp:350-439).  box_2d = {(xmin,ymin),(xmax,ymax)}.
Solution get_bottom_variants(const cv::Point2f& orig_mov_dir, const std::vector<cv::Point2i>& box_2d,
                             const cv::Matx33f& mat, const cv::Matx33f& inv_mat, const cv::Matx33f& inv_matrix_upper,
                             int cls);
// Batched form: n problems in one launch (no reference equivalent; this is how the device wants it).
std::vector<Solution> get_bottom_variants_batch(const std::vector<cv::Point2f>& dirs,
                                                const std::vector<std::vector<cv::Point2i>>& boxes,
                                                const cv::Matx33f& mat, const cv::Matx33f& inv_mat,
                                                const std::vector<cv::Matx33f>& inv_matrix_upper,
                                                const std::vector<int>& cls);
std::pair<cv::Matx33f, cv::Matx33f> get_mat();  // lifting_3d.cpp:482-514
cv::Matx33f get_mat_upper(int cls);             // lifting_3d.cpp:441-480
// The same calibration for a width x height frame of the same view (the reference's quads are 640x360 pixel coordinates).
std::pair<cv::Matx33f, cv::Matx33f> get_mat(int width, int height);
cv::Matx33f get_mat_upper(int cls, int width, int height);
// Intersection of lines a1a2 and b1b2 in float; (NaN, NaN) for parallel lines (lifting_3d.cpp:63-88).  Setup-sized
// scalar helper, evaluated on the host.
cv::Point2f get_intersect(cv::Point2f a1, cv::Point2f a2, cv::Point2f b1, cv::Point2f b2);
#endif
