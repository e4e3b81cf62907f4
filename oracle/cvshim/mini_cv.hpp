// ============================================================================
// mini_cv.hpp — TEST INFRASTRUCTURE.  A minimal stand-in for the slice of OpenCV 3.1 that
// the reference's hot-path sources use (MultiH/MultiH/MultiH.cpp, MultiH.h,
// moduls/mode_seeking/MeanShiftClustering.h, moduls/homographies/*.h), so that those
// files compile UNMODIFIED, in place from /root/reference, into oracle/_ref/libmultih_ref.so
// (oracle/Makefile, oracle/ref_multih_wrapper.cpp) and the oracle's restatement can be
// checked against the reference's own source.  Nothing here is used by the product.
//
// What is implemented is what those files call; semantics follow OpenCV where they matter:
//   * cv::Mat is a reference-counted header over row-major data; row()/col() are views;
//     assigning a Mat copies the header, assigning an expression (MatExpr: a + b, a * s,
//     t(), inv() ...) writes INTO the existing buffer when size and type match — the
//     reference relies on that (`resultMat.row(i) = resultMat.row(i) * avgRatio`,
//     Homography_Refine3PTCallback.h:183);
//   * cv::eigen on a symmetric matrix = cyclic Jacobi, eigenvalues descending, eigenvectors
//     in rows (what OpenCV's eigen does; the reference takes the last row, MultiH.cpp:893);
//   * Mat::inv() = LU with partial pivoting, inv(DECOMP_SVD) = pseudo-inverse from the
//     one-sided Jacobi SVD with OpenCV's threshold (singular values <= DBL_EPSILON * 2 * sum
//     are dropped... see inv());
//   * cv::solvePoly = Durand-Kerner as in OpenCV (same start values, same stopping rule);
//   * concurrency::parallel_for runs its body sequentially, in index order.
// What the reference's pre-path calls but the oracle does not cover (findFundamentalMat,
// findHomography, FlannBasedMatcher, drawing, the LM solver) is declared and aborts when
// reached, with the exception of FlannBasedMatcher::radiusMatch, which implements the
// oracle's documented neighbourhood definition (see there).
// ============================================================================
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_64FC1 CV_64F
#define CV_32FC1 CV_32F
#define CV_32FC2 13
#define CV_64FC2 14
#define CV_32FC3 21
#define CV_64FC3 22
#define CV_FM_RANSAC 8
#define CV_RANSAC 8
#define CV_Assert(expr) do { if (!(expr)) { std::fprintf(stderr, "mini_cv: CV_Assert(%s) failed at %s:%d\n", #expr, __FILE__, __LINE__); std::abort(); } } while (0)
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif
#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif

namespace cv {

typedef unsigned char uchar;
enum { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_EIG = 2, DECOMP_CHOLESKY = 3 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_L2SQR = 5 };

[[noreturn]] inline void mini_cv_unsupported(const char* what) {
  std::fprintf(stderr, "mini_cv: %s is outside the slice of OpenCV this shim implements\n", what);
  std::abort();
}

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T _x, T _y) : x(_x), y(_y) {}
  Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
};
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> inline Point_<T> operator*(const Point_<T>& a, double s) { return Point_<T>((T)(a.x * s), (T)(a.y * s)); }
template <typename T> inline Point_<T> operator*(double s, const Point_<T>& a) { return a * s; }
template <typename T> inline Point_<T> operator/(const Point_<T>& a, double s) { return Point_<T>((T)(a.x / s), (T)(a.y / s)); }
template <typename T> inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
typedef Point_<double> Point2d;
typedef Point_<float> Point2f;
typedef Point_<int> Point;

template <typename T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
  Point3_& operator+=(const Point3_& o) { x += o.x; y += o.y; z += o.z; return *this; }
};
template <typename T> inline Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> inline Point3_<T> operator*(const Point3_<T>& a, double s) { return Point3_<T>((T)(a.x * s), (T)(a.y * s), (T)(a.z * s)); }
template <typename T> inline Point3_<T> operator*(double s, const Point3_<T>& a) { return a * s; }
template <typename T> inline Point3_<T> operator/(const Point3_<T>& a, double s) { return Point3_<T>((T)(a.x / s), (T)(a.y / s), (T)(a.z / s)); }
template <typename T> inline double norm(const Point3_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z); }

struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} bool operator!=(const Size& o) const { return width != o.width || height != o.height; } bool operator==(const Size& o) const { return !(*this != o); } };
struct Range { int start, end; };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; };
class ParallelLoopBody { public: virtual ~ParallelLoopBody() {} virtual void operator()(const Range& r) const = 0; };

template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8U }; };
template <> struct DataType<int> { enum { type = CV_32S }; };
template <> struct DataType<float> { enum { type = CV_32F }; };
template <> struct DataType<double> { enum { type = CV_64F }; };

inline size_t mini_cv_elem_size(int type) { return type == CV_64F ? 8 : type == CV_8U || type == CV_8S ? 1 : type == CV_16U || type == CV_16S ? 2 : 4; }

class MatExpr;
template <typename T> class Mat_;

class Mat {
 public:
  int rows = 0, cols = 0;
  uchar* data = nullptr;
  size_t step = 0;   // bytes per row
  struct SizeProxy {   // `param0.size != x.size` in the LM solver
    const Mat* m;
    Size operator()() const { return Size(m->cols, m->rows); }
    bool operator!=(const SizeProxy& o) const { return m->rows != o.m->rows || m->cols != o.m->cols; }
    bool operator==(const SizeProxy& o) const { return !(*this != o); }
    int width() const { return m->cols; }
  };
  SizeProxy size{this};

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* external, size_t _step = 0) : rows(r), cols(c), data((uchar*)external), type_(type) { step = _step ? _step : (size_t)c * mini_cv_elem_size(type); }
  Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); setTo(s.val[0]); }
  Mat(const Mat& m) : rows(m.rows), cols(m.cols), data(m.data), step(m.step), type_(m.type_), owner_(m.owner_) {}
  template <typename T> explicit Mat(const Point_<T>& p) { create(2, 1, DataType<T>::type); at<T>(0) = p.x; at<T>(1) = p.y; }
  Mat(const MatExpr& e);
  Mat& operator=(const Mat& m) {
    rows = m.rows; cols = m.cols; data = m.data; step = m.step; type_ = m.type_; owner_ = m.owner_;
    return *this;
  }
  Mat& operator=(const MatExpr& e);

  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return;
    rows = r; cols = c; type_ = type;
    step = (size_t)c * mini_cv_elem_size(type);
    const size_t bytes = std::max<size_t>(step * (size_t)r, 1);
    owner_ = std::shared_ptr<uchar>(new uchar[bytes], std::default_delete<uchar[]>());
    data = owner_.get();
  }
  int type() const { return type_; }
  int depth() const { return type_; }
  int channels() const { return 1; }
  size_t elemSize() const { return mini_cv_elem_size(type_); }
  size_t total() const { return (size_t)rows * cols; }
  bool empty() const { return data == nullptr || total() == 0; }
  bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }
  void release() { owner_.reset(); data = nullptr; rows = cols = 0; }

  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + step * (size_t)r); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + step * (size_t)r); }
  template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
  // single index: element i of a vector (row or column), or of a continuous matrix seen as elements of type T
  template <typename T> T& at(int i) { return rows == 1 ? ptr<T>(0)[i] : cols == 1 ? ptr<T>(i)[0] : ptr<T>(i / cols_of<T>())[i % cols_of<T>()]; }
  template <typename T> const T& at(int i) const { return const_cast<Mat*>(this)->at<T>(i); }

  Mat row(int r) const { Mat m(*this); m.rows = 1; m.data = data + step * (size_t)r; return m; }
  Mat col(int c) const { Mat m(*this); m.cols = 1; m.data = data + elemSize() * (size_t)c; return m; }
  Mat rowRange(int a, int b) const { Mat m(*this); m.rows = b - a; m.data = data + step * (size_t)a; return m; }
  Mat colRange(int a, int b) const { Mat m(*this); m.cols = b - a; m.data = data + elemSize() * (size_t)a; return m; }
  Mat diag() const { Mat d(std::min(rows, cols), 1, type_); for (int i = 0; i < d.rows; ++i) std::memcpy(d.ptr<uchar>(i), data + step * (size_t)i + elemSize() * (size_t)i, elemSize()); return d; }
  Mat clone() const { Mat m; copyTo(m); return m; }
  void copyTo(Mat& dst) const {
    dst.create(rows, cols, type_);
    for (int r = 0; r < rows; ++r) std::memcpy(dst.data + dst.step * (size_t)r, data + step * (size_t)r, (size_t)cols * elemSize());
  }
  void copyTo(Mat&& dst) const { copyTo(dst); }   // a temporary row()/col() header: written in place
  void setTo(double v) {
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) set(r, c, v);
  }
  double get(int r, int c) const {
    switch (type_) {
      case CV_64F: return at<double>(r, c);
      case CV_32F: return at<float>(r, c);
      case CV_32S: return at<int>(r, c);
      case CV_8U: return at<uchar>(r, c);
      default: mini_cv_unsupported("Mat element type");
    }
  }
  void set(int r, int c, double v) {
    switch (type_) {
      case CV_64F: at<double>(r, c) = v; break;
      case CV_32F: at<float>(r, c) = (float)v; break;
      case CV_32S: at<int>(r, c) = (int)std::lrint(v); break;
      case CV_8U: at<uchar>(r, c) = (uchar)std::lrint(v); break;
      default: mini_cv_unsupported("Mat element type");
    }
  }
  void convertTo(Mat& dst, int type) const {
    Mat out(rows, cols, type);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) out.set(r, c, get(r, c));
    if (dst.data && dst.rows == rows && dst.cols == cols && dst.type() == type) out.copyTo(dst);
    else dst = out;
  }
  static Mat zeros(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, m.step * (size_t)r); return m; }
  static Mat ones(int r, int c, int type) { Mat m(r, c, type); m.setTo(1.0); return m; }
  static Mat eye(int r, int c, int type) { Mat m = zeros(r, c, type); for (int i = 0; i < std::min(r, c); ++i) m.set(i, i, 1.0); return m; }

  MatExpr t() const;
  MatExpr inv(int method = DECOMP_LU) const;
  MatExpr mul(const Mat& o, double scale = 1.0) const;
  double dot(const Mat& o) const {
    double s = 0;
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) s += get(r, c) * o.get(r, c);
    return s;
  }

 protected:
  template <typename T> int cols_of() const { return (int)(step / sizeof(T)) ? (int)((size_t)cols * elemSize() / sizeof(T)) : 1; }
  int type_ = CV_8U;
  std::shared_ptr<uchar> owner_;
};

// The value of an expression; distinct from Mat only so that assignment can write in place (see the header).
class MatExpr {
 public:
  Mat m;
  MatExpr() {}
  explicit MatExpr(const Mat& _m) : m(_m) {}
  template <typename T> T& at(int r, int c) { return m.at<T>(r, c); }
  template <typename T> T& at(int i) { return m.at<T>(i); }
  MatExpr t() const { return m.t(); }
  MatExpr inv(int method = DECOMP_LU) const { return m.inv(method); }
  double dot(const Mat& o) const { return m.dot(o); }
  Mat row(int r) const { return m.row(r); }
};
inline Mat::Mat(const MatExpr& e) : Mat(e.m) {}
inline Mat& Mat::operator=(const MatExpr& e) {
  if (data && rows == e.m.rows && cols == e.m.cols && type_ == e.m.type()) {
    if (data != e.m.data) e.m.copyTo(*this);
  } else {
    *this = e.m;
  }
  return *this;
}

template <typename T> class MatCommaInitializer_;
template <typename T> class Mat_ : public Mat {
 public:
  Mat_() { type_ = DataType<T>::type; }
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
  Mat_(int r, int c, const T& v) : Mat(r, c, DataType<T>::type) { setTo((double)v); }
  Mat_(const Mat& m) { assign(m); }
  Mat_(const MatExpr& e) { assign(e.m); }
  Mat_& operator=(const Mat& m) { assign(m); return *this; }
  T& operator()(int r, int c) { return at<T>(r, c); }
  const T& operator()(int r, int c) const { return at<T>(r, c); }
  T& operator()(int i) { return at<T>(i); }
  static Mat_ zeros(int r, int c) { return Mat_(Mat::zeros(r, c, DataType<T>::type)); }
  static Mat_ ones(int r, int c) { return Mat_(Mat::ones(r, c, DataType<T>::type)); }
  static Mat_ eye(int r, int c) { return Mat_(Mat::eye(r, c, DataType<T>::type)); }

 private:
  void assign(const Mat& m) {
    if (m.empty() || m.type() == DataType<T>::type) Mat::operator=(m);
    else { Mat c; m.convertTo(c, DataType<T>::type); Mat::operator=(c); }
    type_ = DataType<T>::type;
  }
};

// (Mat_<T>(r, c) << a, b, c ...)
template <typename T> class MatCommaInitializer_ {
 public:
  MatCommaInitializer_(const Mat_<T>& m, T first) : m_(m), i_(0) { put(first); }
  template <typename V> MatCommaInitializer_& operator,(V v) { put((T)v); return *this; }
  operator Mat_<T>() const { return m_; }   // the only conversion, as in OpenCV: `Mat = (Mat_<T>(r, c) << ...)` rebinds the header
  const Mat_<T>& mat() const { return m_; }

 private:
  void put(T v) { CV_Assert(i_ < (int)m_.total()); m_.template at<T>(i_ / m_.cols, i_ % m_.cols) = v; ++i_; }
  Mat_<T> m_;
  int i_;
};
template <typename T, typename V> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, V v) { return MatCommaInitializer_<T>(m, (T)v); }

// ---- arithmetic (eager; computed in double, stored in the operands' type) -------------------------------------------------------
inline MatExpr mini_cv_binary(const Mat& a, const Mat& b, int op) {
  CV_Assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
  Mat out(a.rows, a.cols, a.type());
  if (a.type() == CV_64F) {
    for (int r = 0; r < a.rows; ++r) {
      const double *pa = a.ptr<double>(r), *pb = b.ptr<double>(r);
      double* po = out.ptr<double>(r);
      for (int c = 0; c < a.cols; ++c) po[c] = op == 0 ? pa[c] + pb[c] : pa[c] - pb[c];
    }
  } else {
    for (int r = 0; r < a.rows; ++r)
      for (int c = 0; c < a.cols; ++c) out.set(r, c, op == 0 ? a.get(r, c) + b.get(r, c) : a.get(r, c) - b.get(r, c));
  }
  return MatExpr(out);
}
inline MatExpr mini_cv_scale(const Mat& a, double s, bool divide) {
  Mat out(a.rows, a.cols, a.type());
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols; ++c) out.set(r, c, divide ? a.get(r, c) / s : a.get(r, c) * s);
  return MatExpr(out);
}
inline MatExpr mini_cv_matmul(const Mat& a, const Mat& b) {
  CV_Assert(a.cols == b.rows && a.type() == CV_64F && b.type() == CV_64F);
  Mat out = Mat::zeros(a.rows, b.cols, CV_64F);
  for (int i = 0; i < a.rows; ++i)
    for (int k = 0; k < a.cols; ++k) {
      const double aik = a.at<double>(i, k);
      const double* pb = b.ptr<double>(k);
      double* po = out.ptr<double>(i);
      for (int j = 0; j < b.cols; ++j) po[j] += aik * pb[j];
    }
  return MatExpr(out);
}
inline MatExpr operator+(const Mat& a, const Mat& b) { return mini_cv_binary(a, b, 0); }
inline MatExpr operator+(const MatExpr& a, const Mat& b) { return mini_cv_binary(a.m, b, 0); }
inline MatExpr operator+(const Mat& a, const MatExpr& b) { return mini_cv_binary(a, b.m, 0); }
inline MatExpr operator+(const MatExpr& a, const MatExpr& b) { return mini_cv_binary(a.m, b.m, 0); }
inline MatExpr operator-(const Mat& a, const Mat& b) { return mini_cv_binary(a, b, 1); }
inline MatExpr operator-(const MatExpr& a, const Mat& b) { return mini_cv_binary(a.m, b, 1); }
inline MatExpr operator-(const Mat& a, const MatExpr& b) { return mini_cv_binary(a, b.m, 1); }
inline MatExpr operator-(const MatExpr& a, const MatExpr& b) { return mini_cv_binary(a.m, b.m, 1); }
inline MatExpr operator-(const Mat& a) { return mini_cv_scale(a, -1.0, false); }
inline MatExpr operator*(const Mat& a, const Mat& b) { return mini_cv_matmul(a, b); }
inline MatExpr operator*(const MatExpr& a, const Mat& b) { return mini_cv_matmul(a.m, b); }
inline MatExpr operator*(const Mat& a, const MatExpr& b) { return mini_cv_matmul(a, b.m); }
inline MatExpr operator*(const MatExpr& a, const MatExpr& b) { return mini_cv_matmul(a.m, b.m); }
inline MatExpr operator*(const Mat& a, double s) { return mini_cv_scale(a, s, false); }
inline MatExpr operator*(double s, const Mat& a) { return mini_cv_scale(a, s, false); }
inline MatExpr operator*(const MatExpr& a, double s) { return mini_cv_scale(a.m, s, false); }
inline MatExpr operator*(double s, const MatExpr& a) { return mini_cv_scale(a.m, s, false); }
inline MatExpr operator/(const Mat& a, double s) { return mini_cv_scale(a, s, true); }
inline MatExpr operator/(const MatExpr& a, double s) { return mini_cv_scale(a.m, s, true); }
template <typename T> inline MatExpr operator*(const MatCommaInitializer_<T>& a, const Mat& b) { return mini_cv_matmul(a.mat(), b); }
template <typename T> inline MatExpr operator*(const Mat& a, const MatCommaInitializer_<T>& b) { return mini_cv_matmul(a, b.mat()); }
template <typename T> inline MatExpr operator*(const MatExpr& a, const MatCommaInitializer_<T>& b) { return mini_cv_matmul(a.m, b.mat()); }
template <typename T> inline MatExpr operator*(double s, const MatCommaInitializer_<T>& b) { return mini_cv_scale(b.mat(), s, false); }
template <typename T> inline MatExpr operator-(const MatCommaInitializer_<T>& a, const Mat& b) { return mini_cv_binary(a.mat(), b, 1); }
template <typename T> inline MatExpr operator+(const MatCommaInitializer_<T>& a, const Mat& b) { return mini_cv_binary(a.mat(), b, 0); }
inline Mat& operator+=(Mat& a, const Mat& b) { a = a + b; return a; }
inline Mat& operator-=(Mat& a, const Mat& b) { a = a - b; return a; }
inline Mat& operator*=(Mat& a, double s) { a = a * s; return a; }

inline MatExpr Mat::t() const {
  Mat out(cols, rows, type_);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) std::memcpy(out.data + out.step * (size_t)c + elemSize() * (size_t)r, data + step * (size_t)r + elemSize() * (size_t)c, elemSize());
  return MatExpr(out);
}
inline MatExpr Mat::mul(const Mat& o, double scale) const {
  Mat out(rows, cols, type_);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) out.set(r, c, get(r, c) * o.get(r, c) * scale);
  return MatExpr(out);
}

// ---- InputArray / OutputArray ----------------------------------------------------------------------------------------------------
class _InputArray {
 public:
  _InputArray() {}
  _InputArray(const Mat& m) : m_(m), has_(true) {}
  _InputArray(const MatExpr& e) : m_(e.m), has_(true) {}
  template <typename T> _InputArray(const Mat_<T>& m) : m_(m), has_(true) {}
  template <typename T> _InputArray(const MatCommaInitializer_<T>& m) : m_(m.mat()), has_(true) {}
  template <typename T> _InputArray(const std::vector<Point_<T>>& v) : m_((int)v.size(), 2, DataType<T>::type, (void*)v.data()), has_(true) {}
  _InputArray(const std::vector<double>& v) : m_((int)v.size(), 1, CV_64F, (void*)v.data()), has_(true) {}
  Mat getMat() const { return m_; }
  bool empty() const { return !has_ || m_.empty(); }
  bool needed() const { return has_; }

 protected:
  Mat m_;
  bool has_ = false;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray() {}
  _OutputArray(Mat& m) : ref_(&m) { has_ = true; }
  _OutputArray(Mat&& m) : tmp_(m), ref_(&tmp_) { has_ = true; }   // a temporary row()/col() header
  template <typename T> _OutputArray(Mat_<T>& m) : ref_(&m) { has_ = true; }
  _OutputArray(std::vector<uchar>& v) : bytes_(&v) { has_ = true; }
  Mat getMat() const { return ref_ ? *ref_ : Mat(); }
  Mat& getMatRef() const { CV_Assert(ref_); return *ref_; }
  void create(int r, int c, int type) const { CV_Assert(ref_); ref_->create(r, c, type); }
  void create(Size sz, int type) const { create(sz.height, sz.width, type); }
  bool needed() const { return ref_ != nullptr || bytes_ != nullptr; }
  std::vector<uchar>* bytes() const { return bytes_; }

 private:
  mutable Mat tmp_;
  Mat* ref_ = nullptr;
  std::vector<uchar>* bytes_ = nullptr;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
inline const _OutputArray& noArray() { static _OutputArray none; return none; }

// ---- linear algebra ----------------------------------------------------------------------------------------------------------------
inline double norm(InputArray a, int normType = NORM_L2) {
  const Mat m = a.getMat();
  double s = 0;
  for (int r = 0; r < m.rows; ++r)
    for (int c = 0; c < m.cols; ++c) {
      const double v = m.get(r, c);
      if (normType == NORM_INF) s = std::max(s, std::fabs(v));
      else if (normType == NORM_L1) s += std::fabs(v);
      else s += v * v;
    }
  return normType == NORM_L2 ? std::sqrt(s) : s;
}
inline double norm(const Mat& a, const Mat& b) { return norm(Mat(a - b)); }

// symmetric eigen-decomposition: cyclic Jacobi (what cv::eigen runs), eigenvalues descending, eigenvectors in rows
inline bool eigen(InputArray _src, OutputArray _evals, OutputArray _evecs = noArray()) {
  const Mat src = _src.getMat();
  CV_Assert(src.rows == src.cols && src.type() == CV_64F);
  const int n = src.rows;
  std::vector<double> A((size_t)n * n), V((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = src.at<double>(i, j);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vpk = V[(size_t)p * n + k], vqk = V[(size_t)q * n + k];
          V[(size_t)p * n + k] = c * vpk - s * vqk;
          V[(size_t)q * n + k] = s * vpk + c * vqk;
        }
      }
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return A[(size_t)a * n + a] > A[(size_t)b * n + b]; });
  Mat evals(n, 1, CV_64F), evecs(n, n, CV_64F);
  for (int i = 0; i < n; ++i) {
    evals.at<double>(i) = A[(size_t)order[i] * n + order[i]];
    for (int k = 0; k < n; ++k) evecs.at<double>(i, k) = V[(size_t)order[i] * n + k];
  }
  _evals.getMatRef() = evals;
  if (_evecs.needed()) _evecs.getMatRef() = evecs;
  return true;
}

// one-sided Jacobi SVD of an m x n matrix (m >= n): A = U diag(w) V^T
inline void mini_cv_svd(const Mat& A, std::vector<double>& U, std::vector<double>& w, std::vector<double>& V) {
  const int m = A.rows, n = A.cols;
  U.assign((size_t)m * n, 0.0); V.assign((size_t)n * n, 0.0); w.assign(n, 0.0);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) U[(size_t)i * n + j] = A.at<double>(i, j);
  for (int j = 0; j < n; ++j) V[(size_t)j * n + j] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        double a = 0, b = 0, g = 0;
        for (int i = 0; i < m; ++i) {
          const double up = U[(size_t)i * n + p], uq = U[(size_t)i * n + q];
          a += up * up; b += uq * uq; g += up * uq;
        }
        if (std::fabs(g) <= DBL_EPSILON * std::sqrt(a * b)) continue;
        rotated = true;
        const double zeta = (b - a) / (2.0 * g);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < m; ++i) {
          const double up = U[(size_t)i * n + p], uq = U[(size_t)i * n + q];
          U[(size_t)i * n + p] = c * up - s * uq;
          U[(size_t)i * n + q] = s * up + c * uq;
        }
        for (int i = 0; i < n; ++i) {
          const double vp = V[(size_t)i * n + p], vq = V[(size_t)i * n + q];
          V[(size_t)i * n + p] = c * vp - s * vq;
          V[(size_t)i * n + q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  for (int j = 0; j < n; ++j) {
    double s = 0;
    for (int i = 0; i < m; ++i) s += U[(size_t)i * n + j] * U[(size_t)i * n + j];
    w[j] = std::sqrt(s);
    if (w[j] > 0)
      for (int i = 0; i < m; ++i) U[(size_t)i * n + j] /= w[j];
  }
}

inline bool mini_cv_lu_inverse(const Mat& A, Mat& out) {
  const int n = A.rows;
  std::vector<double> a((size_t)n * n), inv((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) a[(size_t)i * n + j] = A.at<double>(i, j);
  for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(a[(size_t)r * n + c]) > std::fabs(a[(size_t)piv * n + c])) piv = r;
    if (std::fabs(a[(size_t)piv * n + c]) < DBL_EPSILON) return false;
    if (piv != c)
      for (int k = 0; k < n; ++k) { std::swap(a[(size_t)piv * n + k], a[(size_t)c * n + k]); std::swap(inv[(size_t)piv * n + k], inv[(size_t)c * n + k]); }
    const double d = 1.0 / a[(size_t)c * n + c];
    for (int r = c + 1; r < n; ++r) {
      const double f = a[(size_t)r * n + c] * d;
      for (int k = c + 1; k < n; ++k) a[(size_t)r * n + k] -= f * a[(size_t)c * n + k];
      for (int k = 0; k < n; ++k) inv[(size_t)r * n + k] -= f * inv[(size_t)c * n + k];
    }
  }
  for (int c = n - 1; c >= 0; --c)
    for (int k = 0; k < n; ++k) {
      double s = inv[(size_t)c * n + k];
      for (int j = c + 1; j < n; ++j) s -= a[(size_t)c * n + j] * inv[(size_t)j * n + k];
      inv[(size_t)c * n + k] = s / a[(size_t)c * n + c];
    }
  out.create(n, n, CV_64F);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) out.at<double>(i, j) = inv[(size_t)i * n + j];
  return true;
}

inline double invert(InputArray _src, OutputArray _dst, int method = DECOMP_LU) {
  const Mat A = _src.getMat();
  CV_Assert(A.type() == CV_64F);
  Mat out;
  if (method == DECOMP_LU) {
    CV_Assert(A.rows == A.cols);
    if (!mini_cv_lu_inverse(A, out)) out = Mat::zeros(A.rows, A.cols, CV_64F);
  } else {   // DECOMP_SVD / DECOMP_EIG: pseudo-inverse; singular values below DBL_EPSILON * 2 * sum(w) count as zero (OpenCV's SVBkSb)
    const bool wide = A.rows < A.cols;
    const Mat B = wide ? Mat(A.t()) : A;
    std::vector<double> U, w, V;
    mini_cv_svd(B, U, w, V);
    const int m = B.rows, n = B.cols;
    double sum = 0;
    for (double v : w) sum += v;
    const double thr = DBL_EPSILON * 2 * sum;
    Mat P = Mat::zeros(n, m, CV_64F);
    for (int k = 0; k < n; ++k) {
      if (!(w[k] > thr)) continue;
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) P.at<double>(i, j) += V[(size_t)i * n + k] * U[(size_t)j * n + k] / w[k];
    }
    out = wide ? Mat(P.t()) : P;
  }
  _dst.getMatRef() = out;
  return 1.0;
}
inline MatExpr Mat::inv(int method) const { Mat out; invert(*this, out, method); return MatExpr(out); }

inline bool solve(InputArray A, InputArray b, OutputArray x, int method = DECOMP_LU) {
  Mat Ai;
  invert(A, Ai, method == DECOMP_LU || method == DECOMP_CHOLESKY ? DECOMP_LU : DECOMP_SVD);
  x.getMatRef() = Mat(Ai * b.getMat());
  return true;
}
inline void transpose(InputArray a, OutputArray b) { const Mat t = a.getMat().t(); b.getMatRef() = t; }
inline void subtract(InputArray a, InputArray b, OutputArray c) { const Mat r = a.getMat() - b.getMat(); c.getMatRef() = r; }
inline void mulTransposed(InputArray a, OutputArray dst, bool aTa) { const Mat A = a.getMat(); const Mat r = aTa ? Mat(A.t() * A) : Mat(A * A.t()); dst.getMatRef() = r; }
enum { GEMM_1_T = 1, GEMM_2_T = 2, GEMM_3_T = 4 };
inline void gemm(InputArray a, InputArray b, double alpha, InputArray c, double beta, OutputArray dst, int flags = 0) {
  Mat A = a.getMat(), B = b.getMat();
  if (flags & GEMM_1_T) A = A.t();
  if (flags & GEMM_2_T) B = B.t();
  Mat r = Mat(A * B) * alpha;
  if (!c.empty() && beta != 0) { Mat C = c.getMat(); if (flags & GEMM_3_T) C = C.t(); r = r + C * beta; }
  dst.getMatRef() = r;
}

// polynomial roots (coefficients from the constant term up): Durand-Kerner with OpenCV's start values and stopping rule
inline double solvePoly(InputArray _coeffs, OutputArray _roots, int maxIters = 300) {
  typedef std::complex<double> C;
  const Mat coeffs0 = _coeffs.getMat();
  const int n0 = (int)coeffs0.total() - 1;
  CV_Assert(n0 >= 0 && coeffs0.type() == CV_64F);
  std::vector<C> coeffs(n0 + 1), roots(std::max(n0, 1));
  for (int i = 0; i <= n0; ++i) coeffs[i] = C(coeffs0.at<double>(i), 0.0);
  int n = n0;
  while (n > 0 && std::abs(coeffs[n]) <= DBL_EPSILON) --n;   // (OpenCV: leading zero coefficients are dropped)
  C p(1, 0), r(1, 1);
  for (int i = 0; i < n; ++i) { roots[i] = p; p = p * r; }
  double maxDiff = 0;
  for (int iter = 0; iter < maxIters; ++iter) {
    maxDiff = 0;
    for (int i = 0; i < n; ++i) {
      p = roots[i];
      C num = coeffs[n], denom = coeffs[n];
      for (int j = 0; j < n; ++j) {
        num = num * p + coeffs[n - j - 1];
        if (j != i) denom = denom * (p - roots[j]);
      }
      num /= denom;
      roots[i] = p - num;
      maxDiff = std::max(maxDiff, std::abs(num));
    }
    if (maxDiff <= 0) break;
  }
  for (int i = 0; i < n; ++i)
    if (std::fabs(roots[i].imag()) < 100 * DBL_EPSILON) roots[i] = C(roots[i].real(), 0);   // OpenCV: verySmallEps
  for (int i = n; i < n0; ++i) roots[i] = C(0, 0);
  Mat out(n0, 2, CV_64F);   // n0 complex numbers (CV_64FC2 in OpenCV: re, im interleaved)
  for (int i = 0; i < n0; ++i) { out.at<double>(i, 0) = roots[i].real(); out.at<double>(i, 1) = roots[i].imag(); }
  _roots.getMatRef() = out;
  return maxDiff;
}
typedef Point_<double> Vec2d_;   // roots.at<cv::Vec2d>(i)[k]
struct Vec2d {
  double v[2];
  double& operator[](int i) { return v[i]; }
  const double& operator[](int i) const { return v[i]; }
};

// ---- things the pre-path calls; only what the oracle also covers is implemented ---------------------------------------------------
inline Mat findFundamentalMat(InputArray, InputArray, int, double, double, OutputArray) { mini_cv_unsupported("findFundamentalMat (RANSAC, upstream of the hot path)"); }
inline Mat findHomography(InputArray, InputArray, int, double, OutputArray) { mini_cv_unsupported("findHomography (HandleDegenerateCase)"); }
inline void circle(Mat&, Point2d, int, const Scalar&, int) {}
class FlannBasedMatcher {
 public:
  // The reference calls radiusMatch with FLANN's defaults: 4 randomised KD-trees and checks = 32, so a query returns about the
  // 31 nearest other points inside the radius, not the ball.  This shim implements the oracle's exactly-defined
  // stand-in: the query itself + its 31 nearest other points (ties by index) within the radius, in increasing distance.
  void radiusMatch(InputArray _q, InputArray _t, std::vector<std::vector<DMatch>>& out, float radius) {
    const Mat q = _q.getMat(), t = _t.getMat();
    CV_Assert(q.type() == CV_32F && t.type() == CV_32F && q.cols == t.cols);
    out.assign(q.rows, std::vector<DMatch>());
    for (int i = 0; i < q.rows; ++i) {
      std::vector<std::pair<float, int>> cand;
      for (int j = 0; j < t.rows; ++j) {
        float d2 = 0;
        for (int k = 0; k < q.cols; ++k) { const float d = q.at<float>(i, k) - t.at<float>(j, k); d2 += d * d; }
        if (d2 <= radius * radius) cand.push_back({d2, j});
      }
      std::sort(cand.begin(), cand.end());
      if (cand.size() > 32) cand.resize(32);
      for (auto& c : cand) out[i].push_back(DMatch{i, c.second, 0, std::sqrt(c.first)});
    }
  }
};

template <typename T> class Ptr : public std::shared_ptr<T> {
 public:
  Ptr() {}
  Ptr(T* p) : std::shared_ptr<T>(p) {}
  template <typename U> Ptr(const Ptr<U>& o) : std::shared_ptr<T>(std::static_pointer_cast<T>(std::shared_ptr<U>(o))) {}
  operator T*() const { return this->get(); }
};
class Algorithm { public: virtual ~Algorithm() {} };
class LMSolver : public Algorithm {
 public:
  class Callback {
   public:
    virtual ~Callback() {}
    virtual bool compute(InputArray param, OutputArray err, OutputArray J) const = 0;
  };
  virtual int run(InputOutputArray param) const = 0;
};

}  // namespace cv

// MSVC's Parallel Patterns Library: sequential here, in index order (deterministic)
namespace concurrency {
template <typename I, typename F> inline void parallel_for(I first, I last, const F& body) {
  for (I i = first; i < last; ++i) body(i);
}
}  // namespace concurrency
