// Stand-in for the handful of OpenCV core types the reference's cpp/inc signatures mention, used ONLY
// when <opencv2/core.hpp> is not installed (it is not in this image).  With OpenCV present the shim
// headers include the real thing and this file is not used.  Plain data holders: no arithmetic of the
// hot path lives here (that is in csrc/, on the device).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define DOFS3D_CV_MIN 1
#ifndef CV_32FC2
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC1 5
#define CV_32FC2 13
#endif

namespace cv {

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;

struct Vec2f {
    float val[2];
    Vec2f() : val{0.f, 0.f} {}
    Vec2f(float a, float b) : val{a, b} {}
    float& operator[](int i) { return val[i]; }
    const float& operator[](int i) const { return val[i]; }
};

struct Matx33f {
    float val[9];
    Matx33f() : val{0, 0, 0, 0, 0, 0, 0, 0, 0} {}
    Matx33f(float a, float b, float c, float d, float e, float f, float g, float h, float i) : val{a, b, c, d, e, f, g, h, i} {}
    float& operator()(int r, int c) { return val[r * 3 + c]; }
    const float& operator()(int r, int c) const { return val[r * 3 + c]; }
};

// reference-counted dense 2D array, interleaved channels, like cv::Mat for the types above
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* external) : rows(r), cols(c), data(static_cast<unsigned char*>(external)), type_(type) {}
    void create(int r, int c, int type) {
        rows = r;
        cols = c;
        type_ = type;
        store_ = std::shared_ptr<unsigned char>(new unsigned char[total() * elemSize()](), std::default_delete<unsigned char[]>());
        data = store_.get();
    }
    int type() const { return type_; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * ((type_ & 7) == 5 ? 4 : 1); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || total() == 0; }
    bool isContinuous() const { return true; }
    Mat clone() const {
        Mat m;
        if (!empty()) {
            m.create(rows, cols, type_);
            std::memcpy(m.data, data, total() * elemSize());
        }
        return m;
    }
    template <typename T>
    T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * cols * elemSize()); }
    template <typename T>
    const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * cols * elemSize()); }
    template <typename T>
    T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T>
    const T& at(int r, int c) const { return ptr<T>(r)[c]; }

private:
    int type_ = 0;
    std::shared_ptr<unsigned char> store_;
};

}  // namespace cv
