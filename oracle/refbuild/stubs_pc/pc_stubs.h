// pc_stubs.h -- TEST INFRASTRUCTURE.  Minimal stand-ins for the ROS / OpenCV / PCL / cv_bridge
// declarations that the reference's two node sources name, so that the UNMODIFIED reference sources
// compile on a box with none of those libraries:
//   * src/pointcloud/getPointCloud.cpp: its triangulation code (PC:697-884 cloud loops, PC:886-1019
//     closed-form point, PC:375-484 fovea coordinate mapping) is executed by refpc_main.cpp to pin the oracle;
//   * src/gpu_matcher/UG_GPU_matcher.cpp: the matcher node, compiled against include/MatchGPULib.h and run
//     by refnode_main.cpp to show that it consumes this repo's library unchanged.
// Messages a node publishes are captured in ugsm_stub::published (topic -> last image).
//
// Only data containers and no-op plumbing live here: every arithmetic statement that is executed
// comes from the reference file itself.  Nothing here is copied from ROS, OpenCV or PCL; the types
// only mimic the member names the reference source touches.  cv::resize is a stub that aborts: the
// "resized" clouds (PC:784-884) go through OpenCV's bicubic resampler and are out of scope.
#ifndef UGSM_PC_STUBS_H
#define UGSM_PC_STUBS_H

#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

// ---------------------------------------------------------------- logging: swallowed
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_ERROR_STREAM(x) ((void)0)

// ---------------------------------------------------------------- boost bits reached through ROS
namespace boost {
template <class T>
using shared_ptr = std::shared_ptr<T>;
template <class... A>
inline int bind(A...) {
    return 0;
}
}  // namespace boost
struct ugsm_stub_placeholder {};
static const ugsm_stub_placeholder _1 = {}, _2 = {}, _3 = {}, _4 = {};

// ---------------------------------------------------------------- OpenCV containers
typedef long long int64;
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_8UC3 16

namespace cv {
struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
    Size(double w, double h) : width((int)w), height((int)h) {}
    Size(float w, float h) : width((int)w), height((int)h) {}
    bool operator!=(const Size& o) const { return width != o.width || height != o.height; }
    bool operator==(const Size& o) const { return !(*this != o); }
};
struct Vec3f {
    float v[3];
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};
inline size_t stub_elem_size(int type) {
    const int depth = type % 8, ch = type / 8 + 1;
    const size_t d = (depth <= 1) ? 1 : (depth <= 3) ? 2 : (depth <= 5) ? 4 : 8;
    return d * ch;
}
// Dense row-major matrix sharing its buffer between copies (like cv::Mat); clone() deep-copies.
class Mat {
public:
    int rows, cols;
    unsigned char* data;  // first byte (shared between copies, like cv::Mat)
    size_t step;          // bytes per row
    Mat() : rows(0), cols(0), data(nullptr), step(0), type_(0) {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    void create(int r, int c, int type) {
        rows = r;
        cols = c;
        type_ = type;
        buf_ = std::make_shared<std::vector<unsigned char> >((size_t)r * c * stub_elem_size(type), 0);
        data = buf_->data();
        step = (size_t)c * stub_elem_size(type);
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    Mat clone() const {
        Mat m;
        m.rows = rows, m.cols = cols, m.type_ = type_, m.step = step;
        if (buf_) {
            m.buf_ = std::make_shared<std::vector<unsigned char> >(*buf_);
            m.data = m.buf_->data();
        }
        return m;
    }
    void release() {
        buf_.reset();
        data = nullptr;
        rows = cols = 0;
    }
    bool empty() const { return !buf_ || rows == 0 || cols == 0; }
    int type() const { return type_; }
    Size size() const { return Size(cols, rows); }
    template <class T>
    T* ptr(int r) {
        return reinterpret_cast<T*>(buf_->data() + (size_t)r * cols * stub_elem_size(type_));
    }
    template <class T>
    T& at(int r, int c) {
        return ptr<T>(r)[c];
    }
    template <class T>
    const T& at(int r, int c) const {
        return const_cast<Mat*>(this)->ptr<T>(r)[c];
    }

protected:
    int type_;
    std::shared_ptr<std::vector<unsigned char> > buf_;
};
template <class T>
class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(const Mat& m) : Mat(m) {}
    Mat_& operator=(const Mat& m) {
        Mat::operator=(m);
        return *this;
    }
};
enum { INTER_CUBIC = 2 };
inline void resize(const Mat&, Mat&, Size, double, double, int) {
    fprintf(stderr, "pc_stubs: cv::resize is not available (resized clouds are out of scope)\n");
    abort();
}
inline bool imwrite(const std::string&, const Mat&, const std::vector<int>& = std::vector<int>()) { return false; }
inline int64 getTickCount() { return 0; }
inline double getTickFrequency() { return 1.0; }
}  // namespace cv

// ---------------------------------------------------------------- messages
namespace std_msgs {
struct Header {
    std::string frame_id;
};
}  // namespace std_msgs
namespace sensor_msgs {
struct Image {
    std_msgs::Header header;
    unsigned height, width;
    std::string encoding;
    cv::Mat payload;  // the stub carries the pixels directly
    Image() : height(0), width(0) {}
};
typedef std::shared_ptr<const Image> ImageConstPtr;
typedef std::shared_ptr<Image> ImagePtr;
struct CameraInfo {
    unsigned height, width;
    std::vector<double> D;
    std::array<double, 9> K;
    std::array<double, 12> P;
    CameraInfo() : height(0), width(0), D(5, 0.0) {
        K.fill(0.0);
        P.fill(0.0);
    }
};
typedef std::shared_ptr<const CameraInfo> CameraInfoConstPtr;
namespace image_encodings {
static const std::string BGR8 = "bgr8";
static const std::string RGB8 = "rgb8";
static const std::string TYPE_32FC1 = "32FC1";
}  // namespace image_encodings
}  // namespace sensor_msgs
namespace stereo_msgs {
struct DisparityImage {
    std_msgs::Header header;
    sensor_msgs::Image image;
};
typedef std::shared_ptr<const DisparityImage> DisparityImageConstPtr;
typedef std::shared_ptr<DisparityImage> DisparityImagePtr;
}  // namespace stereo_msgs
namespace ug_stereomatcher {
struct foveatedstack {
    std_msgs::Header header;
    int im_height, im_width, roi_width, roi_height, num_levels;
    sensor_msgs::Image image_stack;
};
typedef std::shared_ptr<const foveatedstack> foveatedstackConstPtr;
struct GetDisparitiesGPU {  // srv/GetDisparitiesGPU.srv
    struct Request {
        sensor_msgs::Image imL, imR;
    };
    struct Response {
        stereo_msgs::DisparityImage dispH, dispV, dispC;
        foveatedstack fdispH, fdispV, fdispC;
    };
};
}  // namespace ug_stereomatcher

namespace ugsm_stub {
inline std::map<std::string, cv::Mat>& published() {
    static std::map<std::string, cv::Mat> m;
    return m;
}
inline std::map<std::string, int>& params() {
    static std::map<std::string, int> m;
    return m;
}
}  // namespace ugsm_stub

// ---------------------------------------------------------------- cv_bridge
namespace cv_bridge {
struct CvImage {
    std_msgs::Header header;
    std::string encoding;
    cv::Mat image;
    sensor_msgs::ImagePtr toImageMsg() const {
        sensor_msgs::ImagePtr m(new sensor_msgs::Image);
        m->header = header;
        m->encoding = encoding;
        m->height = image.rows;
        m->width = image.cols;
        m->payload = image.clone();
        return m;
    }
};
typedef std::shared_ptr<CvImage> CvImagePtr;
struct Exception : std::exception {};
inline CvImagePtr toCvCopy(const sensor_msgs::Image& m, const std::string& encoding) {
    CvImagePtr p(new CvImage);
    p->header = m.header;
    p->encoding = encoding;
    p->image = m.payload.clone();
    return p;
}
inline CvImagePtr toCvCopy(const sensor_msgs::ImageConstPtr& m, const std::string& e) { return toCvCopy(*m, e); }
}  // namespace cv_bridge

// ---------------------------------------------------------------- ROS plumbing (inert)
namespace ros {
struct Publisher {
    std::string topic;
    int getNumSubscribers() const { return 0; }
    void publish(const stereo_msgs::DisparityImagePtr& m) const { ugsm_stub::published()[topic] = m->image.payload; }
    void publish(const ug_stereomatcher::foveatedstack& m) const {
        ugsm_stub::published()[topic] = m.image_stack.payload;
    }
    template <class T>
    void publish(const T&) const {}
};
struct ServiceServer {};
struct NodeHandle {
    bool hasParam(const std::string& name) const { return ugsm_stub::params().count(name) != 0; }
    bool getParam(const std::string& name, int& v) const {
        if (!hasParam(name)) return false;
        v = ugsm_stub::params()[name];
        return true;
    }
    template <class T>
    Publisher advertise(const std::string& topic, int) {
        Publisher p;
        p.topic = topic;
        return p;
    }
    template <class M, class O>
    ServiceServer advertiseService(const std::string&, M, O*) {
        return ServiceServer();
    }
};
struct WallDuration {
    double toSec() const { return 0.0; }
};
struct WallTime {
    static WallTime now() { return WallTime(); }
    WallDuration operator-(const WallTime&) const { return WallDuration(); }
};
inline bool ok() { return false; }
inline void init(int&, char**, const std::string&) {}
struct MultiThreadedSpinner {
    explicit MultiThreadedSpinner(int) {}
};
inline void spin(MultiThreadedSpinner&) {}
inline void spin() {}
namespace package {
inline std::string getPath(const std::string&) { return "."; }
}  // namespace package
}  // namespace ros
namespace image_transport {
struct ImageTransport {
    explicit ImageTransport(ros::NodeHandle&) {}
};
struct SubscriberFilter {
    SubscriberFilter(ImageTransport&, const std::string&, int) {}
};
}  // namespace image_transport
namespace message_filters {
template <class M>
struct Subscriber {
    Subscriber(ros::NodeHandle&, const std::string&, int) {}
};
namespace sync_policies {
template <class... M>
struct ApproximateTime {
    explicit ApproximateTime(int) {}
};
}  // namespace sync_policies
template <class Policy>
struct Synchronizer {
    template <class... S>
    Synchronizer(const Policy&, S&...) {}
    template <class F>
    void registerCallback(const F&) {}
};
}  // namespace message_filters

// ---------------------------------------------------------------- PCL containers
namespace pcl {
struct PointXYZRGB {
    float x, y, z;
    float rgb;
    PointXYZRGB() : x(0), y(0), z(0), rgb(0) {}
};
struct PCLHeader {
    std::string frame_id;
};
template <class P>
struct PointCloud {
    typedef boost::shared_ptr<PointCloud<P> > Ptr;
    PCLHeader header;
    std::vector<P> points;
    unsigned width, height;
    PointCloud() : width(0), height(0) {}
};
namespace io {
template <class C>
inline int savePCDFileASCII(const std::string&, const C&) {
    return 0;
}
}  // namespace io
}  // namespace pcl

#endif  // UGSM_PC_STUBS_H
