/* dofs3d.h — C ABI of the B200 (sm_100a) implementation of the hot path of
 * DmitriyZhuravlev/DenseOpticalFlowSegmentation3D:
 *
 *     frame pair -> dense Farneback flow -> flow-graph clustering -> per-cluster 3D box lifting
 *
 * The reference has no FFI layer; its boundary for this path is the set of C++ symbols in cpp/inc
 * (see SURVEY.md section 8b).  Each entry point below names the reference interface it stands in
 * for (file:line under the reference tree).  All pointers are plain HOST pointers unless the
 * function name ends in `_dev` (then every data pointer is a DEVICE pointer on the context's GPU and
 * the call is asynchronous on the context's stream; dofs3d_sync() waits for it).
 *
 * All functions return 0 on success and a negative dofs3d_status on failure; there is NO CPU
 * fallback: without a CUDA device every call fails with DOFS3D_ERR_CUDA.
 * A context is bound to one GPU and is thread-compatible (one thread at a time per context).
 */
#ifndef DOFS3D_H
#define DOFS3D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dofs3d_ctx dofs3d_ctx;

typedef enum {
    DOFS3D_OK = 0,
    DOFS3D_ERR_ARG = -1,      /* bad argument (null pointer, size out of range, n_pairs > max_pairs) */
    DOFS3D_ERR_CUDA = -2,     /* CUDA runtime error, or no CUDA device present */
    DOFS3D_ERR_NOMEM = -3,    /* device allocation failed */
    DOFS3D_ERR_OVERFLOW = -4, /* more boxes/candidates than the caller-provided capacity */
    DOFS3D_ERR_INTERNAL = -5
} dofs3d_status;

/* Every constant the reference hard-codes on this path, in one POD (SURVEY.md section 5 "Config"). */
typedef struct {
    float persp[9];        /* get_mat().first            lifting_3d.cpp:482-514 */
    float inv[9];          /* get_mat().second           lifting_3d.cpp:482-514 */
    float inv_upper[3][9]; /* get_mat_upper(cls)         lifting_3d.cpp:441-480 */
    double pyr_scale;      /* 0.5   calcOpticalFlowFarneback arguments, segment.cpp:101 */
    int levels;            /* 3   */
    int winsize;           /* 15  */
    int iters;             /* 3   */
    int poly_n;            /* 5   */
    double poly_sigma;     /* 1.2 */
    double blur_sigma;     /* 3.0   GaussianBlur(flow, Size(0,0), 3.0), segment.cpp:52 */
    int neighbors;         /* 8     get_segmented_array(..., neighbor = 8), segment.cpp:36,154 */
    int min_size;          /* 500   Forest::new_merge default, graph.hpp:94 */
    double score_threshold;/* 0.3   Forest::new_merge default, graph.hpp:93 */
    int cls_size[3][2];    /* {258,84},{349,165},{370,180}   getObjSize, lifting_3d.cpp:255-259 */
    double cls_min_convexity[3]; /* 3/4, 1/2, 20/29          graph.cpp:328-339 */
} dofs3d_params;

/* One entry of Forest::get_best_segments() (graph.cpp:391-429) without its pixel set: a
 * SegmentData{score, seg, sol, move} (graph.hpp:48-57) with sol = Solution (graph.hpp:25-46).
 * The pixel set is carried by the label image, see dofs3d_segment. */
typedef struct {
    int32_t root;        /* union-find root id = y*W + x of the representative pixel (index into history) */
    int32_t size;        /* |seg| at the snapshot */
    int32_t cls;         /* Solution::cls */
    int32_t parent_box;  /* index (in this frame's box list) of the smallest box whose pixel set strictly
                            contains this one, or -1; segments nest because they are merge-tree nodes */
    int32_t bbox[4];     /* xmin, ymin, xmax, ymax of the pixel set (Forest::get_bounding_box, graph.cpp:446) */
    uint32_t time;       /* index of the merge in the reference's sequence of accepted edges (snapshot instant) */
    float mean_flow[2];  /* Node::flow_value of the root at the snapshot (graph.cpp:184-190) */
    float pad_;
    double score;        /* (w_error + h_error) / 2            graph.cpp:257-260 */
    double move;         /* |mean_flow|                        graph.cpp:294 */
    double orient;       /* Solution::orient (yaw, rad, BEV)   lifting_3d.cpp:244 */
    double w_error;      /* Solution::w_error */
    double h_error;      /* Solution::h_error */
    float ps_bev[8];     /* Solution::ps_bev      4 x (x, y) */
    float rectangle[8];  /* Solution::rectangle   4 x (x, y), BEV ground rectangle */
    float lower_face[8]; /* Solution::lower_face  4 x (x, y), image */
    float upper_face[8]; /* Solution::upper_face  4 x (x, y), image */
} dofs3d_box;

/* Work counters of one frame pair: the gates of Forest::new_merge (graph.cpp:280-375). */
typedef struct {
    int32_t n_edges;       /* edges built by build_graph (graph.cpp:62-93) */
    int32_t n_merges;      /* calls of Forest::new_merge that merged (== W*H - 1) */
    int32_t n_levels;      /* number of Boruvka levels == maximum union-find rank reached */
    int32_t n_candidates;  /* merges that passed the size, row and move gates (== get_score calls) */
    int32_t n_scored;      /* ... of those with a valid rectangle that passed convexity and score gates */
    int32_t n_boxes;       /* history entries (roots with a kept snapshot) */
    int32_t longest_chain; /* longest run of merges won by one root (serial depth of the replay) */
    int32_t final_root;    /* root id of the single set the forest ends as (Forest::find of any pixel) */
    int32_t sort_fallback; /* 1 if the merge times of the batch needed the exact 64-bit fallback sort */
    int32_t replay_exact_chunks; /* batch-wide: 32-event chunks of the mean-flow replay whose fast path failed its
                                    exact check and were replayed in double arithmetic */
} dofs3d_stats;

/* Fills *p with the reference's constants (homographies from get_mat/get_mat_upper, etc.). */
void dofs3d_default_params(dofs3d_params* p);
/* The same constants with the calibration quads of get_mat / get_mat_upper (pixel coordinates of the reference's
 * 640x360 camera, lifting_3d.cpp:441-514) rescaled to a width x height frame of the same view: the homographies then map
 * that frame to the same bird's-eye-view rectangle.  The pixel-valued gates (min_size, the motion threshold) are left
 * as the reference has them.  ("calibration generalised to arbitrary resolution".) */
void dofs3d_params_for_size(dofs3d_params* p, int width, int height);

/* Context for frames of width x height (2..65535 each, at most 2^26 pixels) on GPU `device`; at most max_pairs frame
 * pairs per call.  params == NULL selects dofs3d_default_params. */
int dofs3d_create(dofs3d_ctx** out, int device, int width, int height, int max_pairs, const dofs3d_params* params);
void dofs3d_destroy(dofs3d_ctx* ctx);
/* Waits for the context's stream.  After an asynchronous (_dev) segment/process call it also reports that call's
 * deferred conditions: DOFS3D_ERR_OVERFLOW (more boxes than max_boxes, candidate queue full), DOFS3D_ERR_INTERNAL. */
int dofs3d_sync(dofs3d_ctx* ctx);
const char* dofs3d_last_error(const dofs3d_ctx* ctx);
/* Raw cudaStream_t of the context (as void*), for callers that time with CUDA events. */
void* dofs3d_stream(dofs3d_ctx* ctx);
/* Number of kernel launches issued by this context so far. */
long long dofs3d_launch_count(const dofs3d_ctx* ctx);
/* Bytes of device memory held by the context. */
long long dofs3d_device_bytes(const dofs3d_ctx* ctx);

/* cv::cvtColor(BGR2GRAY) (segment.cpp:97-98) for n frames: bgr [n][H][W][3] -> gray [n][H][W]. */
int dofs3d_gray(dofs3d_ctx* ctx, const uint8_t* bgr, int n_frames, uint8_t* gray_out);
int dofs3d_gray_dev(dofs3d_ctx* ctx, const uint8_t* d_bgr, int n_frames, uint8_t* d_gray_out);

/* cv::calcOpticalFlowFarneback(gray0, gray1, flow, pyr_scale, levels, winsize, iters, poly_n,
 * poly_sigma, 0) (segment.cpp:101) for n_pairs independent pairs:
 * gray0/gray1 [n][H][W] u8 -> flow_out [n][H][W][2] f32 (x, y interleaved, CV_32FC2). */
int dofs3d_flow(dofs3d_ctx* ctx, const uint8_t* gray0, const uint8_t* gray1, int n_pairs, float* flow_out);

/* cv::GaussianBlur(flow, flow, Size(0,0), blur_sigma) (segment.cpp:52): [n][H][W][2] f32. */
int dofs3d_blur(dofs3d_ctx* ctx, const float* flow_in, int n_pairs, float* flow_out);

/* get_segmented_array (segment.cpp:34-72) + Forest::get_best_segments (graph.cpp:391-429) for
 * n_pairs flow fields.  already_blurred != 0 skips the GaussianBlur of segment.cpp:52 (the parity
 * tests feed both sides the same blurred bits).  Outputs, all optional (NULL to skip):
 *   labels_out   [n][H][W] int32: index of the SMALLEST box of that frame containing the pixel, -1 if
 *                none.  Pixel set of box b = pixels labelled b or labelled with a box whose parent_box
 *                chain reaches b.
 *   boxes_out    [n][max_boxes] boxes sorted by ascending root (the order of the history vector)
 *   n_boxes_out  [n]
 *   stats_out    [n]
 *   flow_blurred_out [n][H][W][2]: the blurred field the graph was built on. */
int dofs3d_segment(dofs3d_ctx* ctx, const float* flow, int already_blurred, int n_pairs, int32_t* labels_out,
                   dofs3d_box* boxes_out, int32_t* n_boxes_out, int max_boxes, dofs3d_stats* stats_out,
                   float* flow_blurred_out);

/* What plot_best_segments_simple (draw.cpp:102-160) leaves in every pixel, for the n_pairs results of the LAST
 * segment/process call of this context: segments are painted in ascending root order when score > min_score, later over
 * earlier (draw.cpp:120-147).  painted_out [n][H][W] = index of the box whose colour the pixel ends with, -1 if unpainted.
 * bgr_inout (optional, [n][H][W][3]) receives the class colour of draw.cpp:130-141 in the painted pixels (the "seg"
 * image before the cubes and the blending; dofs3d_render is the whole display). */
int dofs3d_paint(dofs3d_ctx* ctx, int n_pairs, double min_score, int32_t* painted_out, uint8_t* bgr_inout);

/* The frame plot_best_segments_simple (draw.cpp:102-160) returns, for the n_pairs results of the LAST segment/process call:
 * bgr_inout [n][H][W][3] holds the frames to draw on (segment.cpp:166 passes the second frame of the pair) and receives
 * the result: segments with score > min_score painted in ascending root order (draw.cpp:120-147), the wireframe cube of
 * each (draw_cube, draw.cpp:85-100: cv::line, colour (255,0,0), thickness 1) on the frame and on the painted copy, both
 * blended with cv::addWeighted(frame, 3/5, seg, 2/5) (draw.cpp:157-158).  Bit-exact against the same calls of cv2 4.13. */
int dofs3d_render(dofs3d_ctx* ctx, int n_pairs, double min_score, uint8_t* bgr_inout);

/* get_bottom_variants (lifting_3d.cpp:350-439) for n independent (direction, box, cls) problems:
 * dir2 [n][2], bbox4 [n][4] = xmin,ymin,xmax,ymax, cls [n].  out[i].score = (w_error+h_error)/2,
 * out[i].size = 1 when the solution has a rectangle, 0 otherwise (Solution::rectangle.empty()). */
int dofs3d_lift(dofs3d_ctx* ctx, const float* dir2, const int32_t* bbox4, const int32_t* cls, int n, dofs3d_box* out);

/* build_graph (graph.cpp:51-103) on one blurred flow field: the sorted edge list.  Outputs hold
 * 4*W*H entries; returns the number of edges (>= 0) or a negative status. */
long long dofs3d_edges_sorted(dofs3d_ctx* ctx, const float* flow_blurred, int32_t* start, int32_t* end,
                              uint64_t* weight_bits);

/* The whole path of main()/main1() (segment.cpp:97-101,154 / 222-226,250): n_frames consecutive BGR
 * frames [n][H][W][3] -> n_frames-1 pairs (pair i = frames i, i+1): gray, Farneback, blur, graph,
 * segmentation, lifting.  Outputs as dofs3d_segment with n = n_frames - 1. */
int dofs3d_process(dofs3d_ctx* ctx, const uint8_t* bgr_frames, int n_frames, int32_t* labels_out,
                   dofs3d_box* boxes_out, int32_t* n_boxes_out, int max_boxes, dofs3d_stats* stats_out);

/* ---- compact results ("a compact on-wire result format (labels RLE + boxes)") ------------------
 * What the reference's consumers take from a frame's segments is which pixels each kept segment covers (draw.cpp:120-147
 * paints them) and the Solution of each (draw.cpp:85-100 draws its cube).  The int32 label image above carries the pixel
 * sets at 4 bytes per pixel; the two formats below carry the same information in 2 bytes per pixel or in a few KB. */
typedef enum {
    DOFS3D_LABELS_I32 = 0, /* int32 [n][H][W], -1 = no box (the format of dofs3d_segment / dofs3d_process) */
    DOFS3D_LABELS_U16 = 1, /* uint16 [n][H][W], 0xFFFF = no box (a frame has fewer than 4096 boxes) */
    DOFS3D_LABELS_RLE = 2  /* dofs3d_run [n][max_runs]: the label image as runs in raster order */
} dofs3d_label_format;

/* One run of equal labels in raster order: pixels start .. (start of the next run of the frame) - 1, the last run of a
 * frame ends at W*H - 1.  label = box index or -1.  Every pixel belongs to exactly one run (background runs included). */
typedef struct {
    uint32_t start;
    int32_t label;
} dofs3d_run;

/* Where the results of one call go; every pointer is optional (NULL to skip).  Host pointers for the host entry points,
 * device pointers for the _dev ones. */
typedef struct {
    int label_format;     /* dofs3d_label_format */
    void* labels;         /* I32 / U16: the dense image; RLE: dofs3d_run [n][max_runs] */
    int32_t* n_runs;      /* RLE: [n] number of runs of each frame */
    int max_runs;         /* RLE: capacity per frame; a frame with more runs fails with DOFS3D_ERR_OVERFLOW */
    dofs3d_box* boxes;    /* [n][max_boxes] */
    int32_t* n_boxes;     /* [n] */
    int max_boxes;
    dofs3d_stats* stats;  /* [n] */
} dofs3d_outputs;

/* dofs3d_process / dofs3d_segment with the result formats above. */
int dofs3d_process_ex(dofs3d_ctx* ctx, const uint8_t* bgr_frames, int n_frames, const dofs3d_outputs* out);
int dofs3d_process_ex_dev(dofs3d_ctx* ctx, const uint8_t* d_bgr_frames, int n_frames, const dofs3d_outputs* d_out);
int dofs3d_segment_ex(dofs3d_ctx* ctx, const float* flow, int already_blurred, int n_pairs, const dofs3d_outputs* out);

/* ---- streaming (the video loop of main1, segment.cpp:174-275: prev_frame is the only state carried, :268) ------------
 * A stream is a sequence of frames pushed in chunks; pair i = frames (i, i+1) of the WHOLE stream, so the first chunk of
 * n frames yields n-1 pairs and every later chunk of n frames yields n pairs (n <= max_pairs).  The context keeps the
 * last frame of a chunk on the device — its gray image and its polynomial expansion at every pyramid level — so no frame
 * is uploaded or expanded twice, and memory does not grow with the length of the clip.
 *   dofs3d_stream_begin    forgets the carried frame (a new clip starts)
 *   dofs3d_stream_submit   asynchronous: the chunk is copied to one of two device staging buffers on a copy stream
 *                          (under the kernels of the previous chunk), its work is enqueued behind that copy and its
 *                          results are copied to `out` (HOST pointers) behind the work.  The frame buffer and the
 *                          output buffers must stay valid until the matching collect; pinned memory makes the copies
 *                          truly asynchronous.  At most two chunks may be outstanding; the first chunk of a stream needs
 *                          at least two frames.
 *   dofs3d_stream_collect  waits for the OLDEST outstanding chunk (its outputs are complete then) and reports its
 *                          deferred errors; *n_pairs_out = number of pairs of that chunk.
 * Results are bit-identical to one dofs3d_process call over the whole clip. */
int dofs3d_stream_begin(dofs3d_ctx* ctx);
int dofs3d_stream_submit(dofs3d_ctx* ctx, const uint8_t* bgr_frames, int n_frames, const dofs3d_outputs* out);
int dofs3d_stream_collect(dofs3d_ctx* ctx, int* n_pairs_out);

/* ---- Felzenszwalb adaptive-threshold mode --------------------------------------------------------
 * The segmentation of the reference's Python twin on a flow field: graph.py:156-177 segment_graph_flow = sorted edges
 * (graph.py:77-96 build_graph, main.py:310-312 diff in float32), the threshold loop `w <= thr[a] && w <= thr[b]`,
 * thr = w + K / size (graph.py:163-172, main.py:314-315), remove_small_components (graph.py:98-106) and merge_components
 * (graph.py:108-130: mean-flow distance < flow_dist and edge weight < edge_dist, both 5 in the reference).
 * Node ids are row * W + col (the reference addresses its array as img[x][y]; give it the transposed field to compare).
 * flow: ONE field [H][W][2] f32 (host), taken as it is (blur it first with dofs3d_blur if wanted).
 * labels_out [H][W] int32 = Forest.find(pixel), the root id of the pixel's component; *n_components_out = their number.
 * stage: 3 = the whole of segment_graph_flow; 2 = stop after remove_small_components (the reference's segment_graph,
 * graph.py:133-153); 1 = the threshold loop only. */
typedef struct {
    double k;          /* 10.0  --K of main.py:483 */
    int min_size;      /* 100   --min-comp-size of main.py:485 */
    int neighbors;     /* 8     --neighbor of main.py:480 */
    double flow_dist;  /* 5     graph.py:126 */
    double edge_dist;  /* 5     graph.py:126 */
    int stage;         /* 3 */
} dofs3d_fh_params;
void dofs3d_fh_default_params(dofs3d_fh_params* p);
int dofs3d_segment_fh(dofs3d_ctx* ctx, const float* flow, const dofs3d_fh_params* params, int32_t* labels_out,
                      int32_t* n_components_out);

/* ---- bird's-eye-view warp -----------------------------------------------------------------------
 * cv::warpPerspective(img, result, mat, Size(out_w, out_h), INTER_CUBIC, BORDER_REPLICATE) on an 8-bit image of 1, 3 or 4
 * interleaved channels (host pointers): what the reference's `transform` (lifting_3d.cpp:516-522) computes with
 * out_w x out_h = 2500 x 14000 (lifting_3d.cpp:34-35) to build the `bev` image of segment.cpp:143,196.  Bit-exact against
 * cv2 4.13.  dofs3d_bev_transform = transform(frame, get_mat().first) for a frame of the context's size. */
#define DOFS3D_BEV_WIDTH 2500
#define DOFS3D_BEV_HEIGHT 14000
int dofs3d_warp_perspective(dofs3d_ctx* ctx, const uint8_t* img, int width, int height, int channels, const float* mat9,
                            int out_w, int out_h, uint8_t* out);
int dofs3d_bev_transform(dofs3d_ctx* ctx, const uint8_t* bgr_frame, uint8_t* bev_out);

/* The boxes of the LAST segment/process call as one dense DEVICE array (frame after frame, root order inside a frame) and
 * their number: the payload of the NCCL gather of per-frame results without a host round trip.
 * Asynchronous on the context's stream.  Boxes beyond `capacity` are dropped (*d_total_out still counts them). */
int dofs3d_pack_boxes_dev(dofs3d_ctx* ctx, int n_pairs, dofs3d_box* d_out, int capacity, int32_t* d_total_out);

/* ---- the incremental Forest, one call at a time (graph.hpp:72-114) --------------------------------------------------------
 * For callers that drive the merge loop themselves, the way the reference's segment_graph does (graph.cpp:520-531):
 * Forest::Forest (graph.cpp:129-148), find (:150-157), merge (:170-218), new_merge (:272-384) over device state.  Every call
 * is a one-thread kernel and a wait (tens of microseconds): dofs3d_segment computes the same merge sequence for a whole
 * Kruskal pass at once and is the path to use for throughput.  The forest uses the context's GPU, stream and parameters
 * (homographies, class sizes, convexity bounds); flow = one field [H][W][2] of the context's size (host).
 *   dofs3d_forest_boxes   the history (Forest::get_best_segments, graph.cpp:391-429) as box records in ascending root
 *                         order (parent_box is not computed: -1); returns their number or a negative status
 *   dofs3d_forest_pixels  the pixel set of a kept root's snapshot (SegmentData::seg), unsorted; returns its size */
typedef struct dofs3d_forest dofs3d_forest;
int dofs3d_forest_create(dofs3d_ctx* ctx, const float* flow, dofs3d_forest** out);
void dofs3d_forest_destroy(dofs3d_forest* f);
int dofs3d_forest_find(dofs3d_forest* f, int n, int32_t* root_out);
int dofs3d_forest_merge(dofs3d_forest* f, int a, int b, int32_t* root_out);
int dofs3d_forest_new_merge(dofs3d_forest* f, int a, int b, double score_threshold, int min_size);
int dofs3d_forest_num_sets(dofs3d_forest* f, int32_t* num_sets_out);
int dofs3d_forest_last_score(dofs3d_forest* f, int node, double* score_out);   /* Forest::get_segment_best_score */
int dofs3d_forest_bbox(dofs3d_forest* f, int node, int32_t* bbox4_out);         /* Forest::get_bounding_box: returns 0 when the box was cleared, 1 when bbox4_out holds it */
int dofs3d_forest_boxes(dofs3d_forest* f, int max_boxes, dofs3d_box* boxes_out);
int dofs3d_forest_pixels(dofs3d_forest* f, int root, int cap, int32_t* pixels_out);

/* Page-locked host memory for the frame and result buffers of the streaming entry points (cudaMallocHost / cudaFreeHost,
 * so that callers need no CUDA headers); NULL on failure. */
void* dofs3d_pinned_alloc(size_t bytes);
void dofs3d_pinned_free(void* p);

/* ---- per-node state of the finished forest (what the Forest accessors of graph.hpp:96,103 are made of) ---------------
 * For pair `pair` of the LAST segment/process call:
 *   dofs3d_node_state     the set `node` was the root of at the moment it was absorbed (for the final root: at the end):
 *                         its size, mean flow (Node::flow_value, graph.cpp:184-190) and bounding box xmin,ymin,xmax,ymax.
 *                         The reference clears that state when the node is absorbed (graph.cpp:195,207), so
 *                         Forest::get_bounding_box (graph.cpp:446-452) only ever shows it for the final root.
 *   dofs3d_scored_merges  every merge whose get_score was not -1 (graph.cpp:318-326), in no particular order: the root,
 *                         the merge time, the score, and whether it passed the convexity and threshold gates.
 *                         Forest::get_segment_best_score(node) (graph.cpp:386-389) is the score of the LATEST such merge
 *                         of that root (0.0 if none).  Returns the number of such merges (may exceed cap; only cap are
 *                         written) or a negative status. */
int dofs3d_node_state(dofs3d_ctx* ctx, int pair, int node, int32_t* size_out, float* mean_flow2_out, int32_t* bbox4_out);
int dofs3d_scored_merges(dofs3d_ctx* ctx, int pair, int cap, int32_t* root_out, uint32_t* time_out, double* score_out,
                         uint8_t* kept_out);

/* Device-pointer variants (inputs and outputs resident in HBM, asynchronous on dofs3d_stream). */
int dofs3d_process_dev(dofs3d_ctx* ctx, const uint8_t* d_bgr_frames, int n_frames, int32_t* d_labels_out,
                       dofs3d_box* d_boxes_out, int32_t* d_n_boxes_out, int max_boxes, dofs3d_stats* d_stats_out);
int dofs3d_segment_dev(dofs3d_ctx* ctx, const float* d_flow, int already_blurred, int n_pairs, int32_t* d_labels_out,
                       dofs3d_box* d_boxes_out, int32_t* d_n_boxes_out, int max_boxes, dofs3d_stats* d_stats_out);
int dofs3d_flow_dev(dofs3d_ctx* ctx, const uint8_t* d_gray0, const uint8_t* d_gray1, int n_pairs, float* d_flow_out);

/* Synthetic video of SURVEY.md section 8d (integer-defined, bit-identical to the host generator
 * denseopticalflowsegmentation3d_b200/synth.py): frames first_frame .. first_frame+n_frames-1 of the
 * stream `seed`, written as BGR u8 [n][H][W][3] to DEVICE memory. */
int dofs3d_synth_frames_dev(dofs3d_ctx* ctx, uint32_t seed, int n_objects, int first_frame, int n_frames,
                            uint8_t* d_bgr_out);

/* Per-stage device time (ms, CUDA events on the context's stream) of the last process/segment/flow
 * call when timing was enabled with dofs3d_set_timing(ctx, 1): name, total ms and number of timed
 * intervals of up to cap stages (the radix-sort kernels are timed one by one). */
int dofs3d_set_timing(dofs3d_ctx* ctx, int enabled);
int dofs3d_get_timing(dofs3d_ctx* ctx, const char** names, float* ms, int* counts, int cap);

#ifdef __cplusplus
}
#endif
#endif /* DOFS3D_H */
