// OpenCV core stand-in: the handful of cv::Mat operations RadarCPU.cpp:36-46,402-542 and image_algorithms.h use.
// Arithmetic restated from OpenCV 4: `Mat *= double` is convertTo(self, -1, alpha), which for CV_32F works in float
// (dst = src * (float)alpha); convertTo(CV_8UC1) is saturate_cast<uchar>(cvRound(x)) with cvRound = lrint
// (round-half-even; NaN / out-of-int-range -> INT_MIN -> 0).
#ifndef RR_SHIM_OPENCV_CORE_HPP
#define RR_SHIM_OPENCV_CORE_HPP
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
typedef unsigned char uchar;
#define CV_8UC1 0
#define CV_32FC1 5
#define CV_64FC1 6
namespace cv {
struct Scalar { double v[4]; Scalar(double a = 0) : v{a, 0, 0, 0} {} };
struct Size { int width = 0, height = 0; Size() = default; Size(int w, int h) : width(w), height(h) {} };
inline int cvRound(float x) { return (std::fabs(x) < 2147483648.0f) ? (int)lrintf(x) : INT32_MIN; }
class Mat {
public:
    int rows = 0, cols = 0;
    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; m_type = type; m_esz = type == CV_8UC1 ? 1 : type == CV_32FC1 ? 4 : 8;
        m_step = (size_t)cols * m_esz;
        m_store = std::make_shared<std::vector<uint8_t>>((size_t)rows * m_step + 8);
        m_data = m_store->data();
    }
    void resize(size_t nrows)
    {
        auto ns = std::make_shared<std::vector<uint8_t>>(nrows * m_step + 8);
        if (m_store) std::memcpy(ns->data(), m_data, std::min<size_t>(nrows, (size_t)rows) * m_step);
        m_store = ns; m_data = ns->data(); rows = (int)nrows;
    }
    Mat& setTo(const Scalar& s)
    {
        for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) {
            if (m_type == CV_8UC1) *ptr<uchar>(r, c) = (uchar)s.v[0];
            else if (m_type == CV_32FC1) *ptr<float>(r, c) = (float)s.v[0];
            else *ptr<double>(r, c) = s.v[0];
        }
        return *this;
    }
    Mat col(int x) const { Mat m = *this; m.cols = 1; m.m_data = m_data + (size_t)x * m_esz; return m; }
    template <typename T> T& at(int r, int c = 0) { return *ptr<T>(r, c); }
    template <typename T> const T& at(int r, int c = 0) const { return *ptr<T>(r, c); }
    void convertTo(const Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const
    {
        Mat& d = const_cast<Mat&>(dst);
        if (rtype < 0) rtype = m_type;
        for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) {
            if (m_type == CV_32FC1 && rtype == CV_8UC1) {
                const int v = cvRound(*ptr<float>(r, c));
                *d.ptr<uchar>(r, c) = (uchar)(v < 0 ? 0 : v > 255 ? 255 : v);
            } else if (m_type == CV_32FC1 && rtype == CV_32FC1) {
                *d.ptr<float>(r, c) = *ptr<float>(r, c) * (float)alpha + (float)beta;
            }
        }
    }
    int type() const { return m_type; }
    const uint8_t* raw() const { return m_data; }
    size_t step() const { return m_step; }
protected:
    template <typename T> T* ptr(int r, int c) const { return reinterpret_cast<T*>(m_data + (size_t)r * m_step + (size_t)c * m_esz); }
    int m_type = 0; size_t m_esz = 1, m_step = 0;
    std::shared_ptr<std::vector<uint8_t>> m_store;
    uint8_t* m_data = nullptr;
};
inline Mat& operator*=(Mat& a, double s) { a.convertTo(a, -1, s); return a; }
template <typename T> struct MatType;
template <> struct MatType<uchar> { static const int value = CV_8UC1; };
template <> struct MatType<float> { static const int value = CV_32FC1; };
template <> struct MatType<double> { static const int value = CV_64FC1; };
template <typename T>
class Mat_ : public Mat {
public:
    Mat_() { m_type = MatType<T>::value; }
    Mat_(int r, int c) : Mat(r, c, MatType<T>::value) {}
    Mat_(int r, int c, const T& v) : Mat(r, c, MatType<T>::value) { for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) *this->template ptr<T>(i, j) = v; }
    explicit Mat_(const Size& s) : Mat(s.height, s.width, MatType<T>::value) {}
    T& operator()(int r, int c = 0) { return *this->template ptr<T>(r, c); }
};
} // namespace cv
#endif
