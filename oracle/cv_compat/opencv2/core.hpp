// TEST INFRASTRUCTURE (oracle) — not part of the product.
//
// Header-only stand-in for the handful of OpenCV types that the reference's
// cpp/src/graph.cpp and cpp/src/lifting_3d.cpp touch, so that those two files
// compile UNCHANGED from /root/reference without an OpenCV C++ installation
// (absent in this image).  Only the operators the reference actually uses are
// provided; each one follows OpenCV 4.x's per-operator rounding rule
// (compute in the promoted type of the operands, saturate_cast to _Tp once).
//
// Users in the reference:
//   cv::Point_/Vec/Matx33f/Mat/norm      graph.cpp:120-148,184-208,294   lifting_3d.cpp:63-253
//   cv::getPerspectiveTransform          lifting_3d.cpp:479,510-511
//   cv::imshow/waitKey/warpPerspective   graph.cpp:365-367, lifting_3d.cpp:520  (no-ops here)
#ifndef DOFS3D_ORACLE_CV_COMPAT_CORE_HPP
#define DOFS3D_ORACLE_CV_COMPAT_CORE_HPP

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <exception>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)

namespace cv {

template <typename T, int n>
struct Vec;

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U>
    Point_(const Point_<U>& p) : x(static_cast<T>(p.x)), y(static_cast<T>(p.y)) {}
    Point_(const Vec<T, 2>& v);
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T>
inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) {
    return Point_<T>(static_cast<T>(a.x + b.x), static_cast<T>(a.y + b.y));
}
template <typename T>
inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) {
    return Point_<T>(static_cast<T>(a.x - b.x), static_cast<T>(a.y - b.y));
}
template <typename T>
inline Point_<T> operator*(const Point_<T>& a, int b) {
    return Point_<T>(static_cast<T>(a.x * b), static_cast<T>(a.y * b));
}
template <typename T>
inline Point_<T> operator*(int a, const Point_<T>& b) {
    return Point_<T>(static_cast<T>(b.x * a), static_cast<T>(b.y * a));
}
template <typename T>
inline Point_<T> operator*(const Point_<T>& a, float b) {
    return Point_<T>(static_cast<T>(a.x * b), static_cast<T>(a.y * b));
}
template <typename T>
inline Point_<T> operator*(float a, const Point_<T>& b) {
    return Point_<T>(static_cast<T>(b.x * a), static_cast<T>(b.y * a));
}
template <typename T>
inline Point_<T> operator*(const Point_<T>& a, double b) {
    return Point_<T>(static_cast<T>(a.x * b), static_cast<T>(a.y * b));
}
template <typename T>
inline Point_<T> operator*(double a, const Point_<T>& b) {
    return Point_<T>(static_cast<T>(b.x * a), static_cast<T>(b.y * a));
}
template <typename T>
inline Point_<T> operator/(const Point_<T>& a, int b) {
    return Point_<T>(static_cast<T>(a.x / b), static_cast<T>(a.y / b));
}
template <typename T>
inline Point_<T> operator/(const Point_<T>& a, float b) {
    return Point_<T>(static_cast<T>(a.x / b), static_cast<T>(a.y / b));
}
template <typename T>
inline Point_<T> operator/(const Point_<T>& a, double b) {
    return Point_<T>(static_cast<T>(a.x / b), static_cast<T>(a.y / b));
}
template <typename T>
inline bool operator==(const Point_<T>& a, const Point_<T>& b) {
    return a.x == b.x && a.y == b.y;
}

template <typename T>
inline double norm(const Point_<T>& p) {
    return std::sqrt((double)p.x * p.x + (double)p.y * p.y);
}

template <typename T, int n>
struct Vec {
    T val[n];
    Vec() {
        for (int i = 0; i < n; ++i) val[i] = T(0);
    }
    Vec(T a, T b) {
        static_assert(n >= 2, "Vec(a,b)");
        for (int i = 0; i < n; ++i) val[i] = T(0);
        val[0] = a;
        val[1] = b;
    }
    Vec(T a, T b, T c) {
        static_assert(n >= 3, "Vec(a,b,c)");
        for (int i = 0; i < n; ++i) val[i] = T(0);
        val[0] = a;
        val[1] = b;
        val[2] = c;
    }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 2> Vec2f;
typedef Vec<unsigned char, 3> Vec3b;

template <typename T>
Point_<T>::Point_(const Vec<T, 2>& v) : x(v[0]), y(v[1]) {}

// Vec * int : every component is T(a[i] * alpha) with alpha kept as int.
template <typename T, int n>
inline Vec<T, n> operator*(const Vec<T, n>& a, int alpha) {
    Vec<T, n> r;
    for (int i = 0; i < n; ++i) r.val[i] = static_cast<T>(a.val[i] * alpha);
    return r;
}
template <typename T, int n>
inline Vec<T, n> operator+(const Vec<T, n>& a, const Vec<T, n>& b) {
    Vec<T, n> r;
    for (int i = 0; i < n; ++i) r.val[i] = static_cast<T>(a.val[i] + b.val[i]);
    return r;
}
// Vec / int multiplies by the double reciprocal (OpenCV: Matx_ScaleOp with 1./alpha).
template <typename T, int n>
inline Vec<T, n> operator/(const Vec<T, n>& a, int alpha) {
    const double inv = 1. / alpha;
    Vec<T, n> r;
    for (int i = 0; i < n; ++i) r.val[i] = static_cast<T>(a.val[i] * inv);
    return r;
}
template <typename T, int n>
inline double norm(const Vec<T, n>& v) {
    double s = 0;
    for (int i = 0; i < n; ++i) {
        double t = v.val[i];
        s += t * t;
    }
    return std::sqrt(s);
}

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
    double operator[](int i) const { return val[i]; }
};

class Exception : public std::exception {
public:
    std::string msg;
    explicit Exception(const std::string& m = "") : msg(m) {}
    const char* what() const noexcept override { return msg.c_str(); }
};

// Dense row-major matrix view with shared ownership — just enough for the
// reference (rows, cols, at<T>, clone, empty) and for the oracle driver to
// wrap a raw interleaved CV_32FC2 flow buffer.
class Mat {
public:
    int rows, cols;
    Mat() : rows(0), cols(0), elem_(0) {}
    Mat(int r, int c, int type, const Scalar& = Scalar()) : rows(r), cols(c), elem_(elem_size(type)) {
        buf_ = std::shared_ptr<unsigned char>(new unsigned char[(size_t)r * c * elem_](),
                                              std::default_delete<unsigned char[]>());
    }
    // wrap external memory (not owned)
    Mat(int r, int c, int type, void* data) : rows(r), cols(c), elem_(elem_size(type)) {
        buf_ = std::shared_ptr<unsigned char>(static_cast<unsigned char*>(data), [](unsigned char*) {});
    }
    bool empty() const { return rows == 0 || cols == 0 || !buf_; }
    Size size() const { return Size(cols, rows); }
    Mat clone() const {
        Mat m;
        m.rows = rows;
        m.cols = cols;
        m.elem_ = elem_;
        if (buf_) {
            size_t bytes = (size_t)rows * cols * elem_;
            m.buf_ = std::shared_ptr<unsigned char>(new unsigned char[bytes], std::default_delete<unsigned char[]>());
            std::memcpy(m.buf_.get(), buf_.get(), bytes);
        }
        return m;
    }
    template <typename T>
    T& at(int r, int c) {
        return *reinterpret_cast<T*>(buf_.get() + ((size_t)r * cols + c) * elem_);
    }
    template <typename T>
    const T& at(int r, int c) const {
        return *reinterpret_cast<const T*>(buf_.get() + ((size_t)r * cols + c) * elem_);
    }
    unsigned char* ptr() { return buf_.get(); }

private:
    static size_t elem_size(int type) {
        static const size_t depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
        return depth_bytes[type & 7] * (size_t)((type >> 3) + 1);
    }
    std::shared_ptr<unsigned char> buf_;
    size_t elem_;
};

struct Matx33d {
    double val[9];
};

struct Matx33f {
    float val[9];
    Matx33f() {
        for (float& v : val) v = 0.f;
    }
    Matx33f(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7, float a8)
        : val{a0, a1, a2, a3, a4, a5, a6, a7, a8} {}
    // cv::Mat(CV_64F 3x3) -> Matx33f conversion used at lifting_3d.cpp:479,510-511
    Matx33f(const Matx33d& m) {
        for (int i = 0; i < 9; ++i) val[i] = static_cast<float>(m.val[i]);
    }
    float& operator()(int r, int c) { return val[r * 3 + c]; }
    const float& operator()(int r, int c) const { return val[r * 3 + c]; }
};

// getPerspectiveTransform: the 8x8 system OpenCV builds (rows i and i+4:
// [x y 1 0 0 0 -x*u -y*u | u], [0 0 0 x y 1 -x*v -y*v | v]), solved by LU with
// partial pivoting in double, M[8] = 1.
inline Matx33d getPerspectiveTransform(const Point2f src[], const Point2f dst[]) {
    double a[8][9];
    for (int i = 0; i < 4; ++i) {
        double x = src[i].x, y = src[i].y, u = dst[i].x, v = dst[i].y;
        double r0[9] = {x, y, 1, 0, 0, 0, -x * u, -y * u, u};
        double r1[9] = {0, 0, 0, x, y, 1, -x * v, -y * v, v};
        for (int j = 0; j < 9; ++j) {
            a[i][j] = r0[j];
            a[i + 4][j] = r1[j];
        }
    }
    const int n = 8;
    for (int i = 0; i < n; ++i) {
        int k = i;
        for (int j = i + 1; j < n; ++j)
            if (std::abs(a[j][i]) > std::abs(a[k][i])) k = j;
        if (k != i)
            for (int j = i; j <= n; ++j) std::swap(a[i][j], a[k][j]);
        double d = -1 / a[i][i];
        for (int j = i + 1; j < n; ++j) {
            double alpha = a[j][i] * d;
            for (int c = i + 1; c <= n; ++c) a[j][c] += alpha * a[i][c];
        }
    }
    double xs[8];
    for (int i = n - 1; i >= 0; --i) {
        double s = a[i][n];
        for (int k = i + 1; k < n; ++k) s -= a[i][k] * xs[k];
        xs[i] = s / a[i][i];
    }
    Matx33d m;
    for (int i = 0; i < 8; ++i) m.val[i] = xs[i];
    m.val[8] = 1.0;
    return m;
}

enum { INTER_CUBIC = 2, BORDER_REPLICATE = 1, WINDOW_NORMAL = 0 };

inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return 0; }
inline void warpPerspective(const Mat&, Mat&, const Matx33f&, Size, int = 0, int = 0) {}

}  // namespace cv

inline int cvIsNaN(double v) { return std::isnan(v) ? 1 : 0; }

#endif
