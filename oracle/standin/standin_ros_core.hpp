// Stand-ins for the third-party middleware the reference's hot-path sources include but this image lacks:
// roscpp, tf, message_filters, laser_geometry, angles, the ROS message headers and the few boost facilities
// (TEST INFRASTRUCTURE - see oracle/standin/README.md).  With these, oracle/Makefile compiles the reference's own
// move_control/src/{map_provider,steerer,laser_map_updater,range_map_updater,vfh}.cpp UNCHANGED, in place, into
// oracle/_ref/libnav_ref.so, and tests/cpp compiles the SAME map_provider.cpp / steerer.cpp against the drop-in
// headers of include/move_control/ (the product).  Nothing here is reference code; nothing here is product code.
//
// What the stand-ins do:
//   * a "world" per harness instance (standin::World): a settable clock, named frames with 2-D poses in the map
//     frame (what tf would answer), a parameter table and a topic bus that delivers a published message
//     synchronously to the subscribers of that topic and remembers the last message of every topic.
//   * boost::thread does NOT run its function (the harness calls MapProvider::updateMap / Steerer::update itself,
//     deterministically); mutexes are std::mutex.
//   * tf::TransformListener answers from the world's frames: p_map = R(yaw) p + t, fp64, evaluated as
//     (c*px - s*py) + x0 / (s*px + c*py) + y0 with the fixed-sequence sincos of the scan-form specification
//     (ros_navigation_b200/csrc/scan_project.h), so that the device scan form can be compared bit for bit.
//   * laser_geometry::LaserProjection::transformLaserScanToPointCloud = that same specification: readings with
//     range_min <= r < range_max, polar -> sensor frame in float32, ONE rigid transform per scan (static sensor
//     during the scan; the per-point time interpolation of the published high-fidelity projection is not modelled),
//     float32 x / y / "index" channel.
#ifndef B200NAV_STANDIN_ROS_CORE_HPP
#define B200NAV_STANDIN_ROS_CORE_HPP

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <sys/types.h>

#include <cmath>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <typeindex>
#include <vector>

// ------------------------------------------------------------------------------------------------------------------
// boost
// ------------------------------------------------------------------------------------------------------------------
namespace boost {
template <typename T>
using shared_ptr = std::shared_ptr<T>;
template <typename T, typename U>
inline std::shared_ptr<T> const_pointer_cast(const std::shared_ptr<U>& p) {
  return std::const_pointer_cast<T>(p);
}
template <typename T, typename... A>
inline std::shared_ptr<T> make_shared(A&&... a) {
  return std::make_shared<T>(std::forward<A>(a)...);
}
template <typename... A>
inline auto bind(A&&... a) -> decltype(std::bind(std::forward<A>(a)...)) {
  return std::bind(std::forward<A>(a)...);
}
struct mutex {
  std::mutex m;
  void lock() { m.lock(); }
  void unlock() { m.unlock(); }
  bool try_lock() { return m.try_lock(); }
  typedef std::unique_lock<mutex> scoped_lock;
};
struct shared_mutex : mutex {
  void lock_shared() { lock(); }
  void unlock_shared() { unlock(); }
};
template <typename M>
using unique_lock = std::unique_lock<M>;
template <typename M>
struct shared_lock {
  M& m;
  bool owns;
  explicit shared_lock(M& mm) : m(mm), owns(true) { m.lock_shared(); }
  void unlock() {
    if (owns) m.unlock_shared();
    owns = false;
  }
  ~shared_lock() { unlock(); }
};
// The reference starts its 5 Hz loops on boost::threads; the harness drives the loop bodies itself.
struct thread {
  thread() {}
  template <typename F>
  explicit thread(F) {}
  void join() {}
  void detach() {}
};
}  // namespace boost
// boost/bind.hpp puts the placeholders into the global namespace
using std::placeholders::_1;
using std::placeholders::_2;

// ------------------------------------------------------------------------------------------------------------------
// the world behind roscpp / tf
// ------------------------------------------------------------------------------------------------------------------
namespace standin {

// sin / cos as specified in ros_navigation_b200/csrc/scan_project.h (fixed sequence of rounded fp64 operations).
inline void spec_sincos(double x, double& s, double& c) {
  const double inv_pio2 = 6.36619772367581382433e-01;
  const double p1 = 1.57079632673412561417e+00, p2 = 6.07710050630396597660e-11, p3 = 2.02226624871116645580e-21;
  const double magic = 6755399441055744.0;
  const double fn = (x * inv_pio2 + magic) - magic;
  const double r = ((x - fn * p1) - fn * p2) - fn * p3;
  const double z = r * r;
  const double ps =
      -1.66666666666666324348e-01 +
      z * (8.33333333332248946124e-03 +
           z * (-1.98412698298579493134e-04 +
                z * (2.75573137070700676789e-06 + z * (-2.50507602534068634195e-08 + z * 1.58969099521155010221e-10))));
  const double sk = r + r * (z * ps);
  const double pc =
      4.16666666666666019037e-02 +
      z * (-1.38888888888741095749e-03 +
           z * (2.48015872894767294178e-05 +
                z * (-2.75573143513906633035e-07 + z * (2.08757232129817482790e-09 + z * -1.13596475577881948265e-11))));
  const double ck = 1.0 - (0.5 * z - (z * z) * pc);
  const int q = (int)(((long long)fn) & 3);
  s = (q == 0) ? sk : (q == 1) ? ck : (q == 2) ? -sk : -ck;
  c = (q == 0) ? ck : (q == 1) ? -sk : (q == 2) ? -ck : sk;
}

struct Frame {
  double x = 0, y = 0, yaw = 0;
};

struct World {
  int64_t now_ns = 0;
  std::map<std::string, Frame> frames;  // pose of every frame in the map ("odom") frame
  std::map<std::string, double> params;
  typedef std::function<void(const std::shared_ptr<const void>&)> Handler;
  struct Sub {
    int id;
    std::type_index type;
    Handler fn;
  };
  std::map<std::string, std::vector<Sub>> subs;
  std::map<std::string, std::shared_ptr<const void>> last;  // last message per topic
  std::map<std::string, long> published;                    // message count per topic
  int next_id = 1;

  static std::string norm(const std::string& f) { return (!f.empty() && f[0] == '/') ? f.substr(1) : f; }
  bool frame(const std::string& name, Frame& out) const {
    auto it = frames.find(norm(name));
    if (it == frames.end()) return false;
    out = it->second;
    return true;
  }
  template <typename M>
  int subscribe(const std::string& topic, std::function<void(const std::shared_ptr<const M>&)> f) {
    Sub s{next_id++, std::type_index(typeid(M)),
          [f](const std::shared_ptr<const void>& p) { f(std::static_pointer_cast<const M>(p)); }};
    subs[norm(topic)].push_back(s);
    return s.id;
  }
  void unsubscribe(int id) {
    for (auto& kv : subs)
      for (size_t i = 0; i < kv.second.size(); ++i)
        if (kv.second[i].id == id) {
          kv.second.erase(kv.second.begin() + i);
          return;
        }
  }
  template <typename M>
  void publish(const std::string& topic, const M& msg) {
    std::shared_ptr<const M> p = std::make_shared<M>(msg);
    const std::string t = norm(topic);
    last[t] = p;
    published[t] += 1;
    auto it = subs.find(t);
    if (it == subs.end()) return;
    std::vector<Sub> copy = it->second;
    for (auto& s : copy)
      if (s.type == std::type_index(typeid(M))) s.fn(p);
  }
  template <typename M>
  std::shared_ptr<const M> last_of(const std::string& topic) const {
    auto it = last.find(norm(topic));
    if (it == last.end()) return std::shared_ptr<const M>();
    return std::static_pointer_cast<const M>(it->second);
  }
};

// The world the calling thread is working in; the harness sets it around every call into the code under test.
inline World*& current() {
  static thread_local World* w = nullptr;
  return w;
}
inline World& world() {
  if (!current()) throw std::runtime_error("standin: no current world on this thread");
  return *current();
}
struct Scope {
  World* prev;
  explicit Scope(World* w) : prev(current()) { current() = w; }
  ~Scope() { current() = prev; }
};
}  // namespace standin

// ------------------------------------------------------------------------------------------------------------------
// roscpp
// ------------------------------------------------------------------------------------------------------------------
#define ROS_DEBUG(...) ((void)0)
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_ERROR_THROTTLE(...) ((void)0)
#define ROS_WARN_THROTTLE(...) ((void)0)
#define ROS_INFO_STREAM(...) ((void)0)
#define ROS_INFO_STREAM_ONCE(...) ((void)0)
#define ROS_WARN_STREAM(...) ((void)0)

namespace ros {
struct Duration {
  int64_t ns;
  Duration() : ns(0) {}
  Duration(double s) : ns((int64_t)floor(s * 1e9 + 0.5)) {}
  double toSec() const { return (double)ns * 1e-9; }
  bool operator>(const Duration& o) const { return ns > o.ns; }
  bool operator<(const Duration& o) const { return ns < o.ns; }
  bool sleep() const { return true; }
};
struct Time {
  int64_t ns;
  Time() : ns(0) {}
  Time(double s) : ns((int64_t)floor(s * 1e9 + 0.5)) {}
  static Time now() {
    Time t;
    t.ns = standin::world().now_ns;
    return t;
  }
  double toSec() const { return (double)ns * 1e-9; }
  uint64_t toNSec() const { return (uint64_t)ns; }
  Time& fromNSec(uint64_t n) {
    ns = (int64_t)n;
    return *this;
  }
  Time operator+(const Duration& d) const {
    Time t;
    t.ns = ns + d.ns;
    return t;
  }
  Duration operator-(const Time& o) const {
    Duration d;
    d.ns = ns - o.ns;
    return d;
  }
  bool operator<(const Time& o) const { return ns < o.ns; }
  bool operator>(const Time& o) const { return ns > o.ns; }
  bool operator<=(const Time& o) const { return ns <= o.ns; }
  bool operator>=(const Time& o) const { return ns >= o.ns; }
  bool operator==(const Time& o) const { return ns == o.ns; }
};
struct Rate {
  Duration period;
  Rate(double hz) : period(1.0 / hz) {}
  bool sleep() { return true; }
  Duration cycleTime() const { return Duration(0.0); }
};
inline bool ok() { return false; } /* loops of the code under test never spin here */

struct Publisher {
  standin::World* w = nullptr;
  std::string topic;
  template <typename M>
  void publish(const M& m) const {
    if (w) w->publish<M>(topic, m);
  }
};
struct Subscriber {
  standin::World* w = nullptr;
  int id = 0;
  void shutdown() {
    if (w && id) w->unsubscribe(id);
    id = 0;
  }
};
struct NodeHandle {
  standin::World* w;
  NodeHandle() : w(&standin::world()) {}
  explicit NodeHandle(const std::string&) : w(&standin::world()) {}
  bool ok() const { return false; }
  template <typename M>
  Publisher advertise(const std::string& topic, int, bool = false) {
    Publisher p;
    p.w = w;
    p.topic = topic;
    return p;
  }
  template <typename M, typename T>
  Subscriber subscribe(const std::string& topic, int, void (T::*fp)(const std::shared_ptr<const M>&), T* obj) {
    Subscriber s;
    s.w = w;
    s.id = w->subscribe<M>(topic, [obj, fp](const std::shared_ptr<const M>& m) { (obj->*fp)(m); });
    return s;
  }
  template <typename V>
  bool getParam(const std::string& key, V& out) const {
    auto it = w->params.find(key);
    if (it == w->params.end()) return false;
    out = (V)it->second;
    return true;
  }
  template <typename V>
  void param(const std::string& key, V& out, const V& def) const {
    if (!getParam(key, out)) out = def;
  }
};
inline void init(int&, char**, const std::string&) {}
inline void spin() {}
inline void spinOnce() {}
}  // namespace ros

// ------------------------------------------------------------------------------------------------------------------
// messages
// ------------------------------------------------------------------------------------------------------------------
#define B200NAV_STANDIN_MSG(NS, NAME)                 \
  namespace NS {                                      \
  typedef std::shared_ptr<NAME> NAME##Ptr;            \
  typedef std::shared_ptr<const NAME> NAME##ConstPtr; \
  }

namespace std_msgs {
struct Header {
  uint32_t seq = 0;
  ros::Time stamp;
  std::string frame_id;
};
}  // namespace std_msgs

namespace geometry_msgs {
struct Point {
  double x = 0, y = 0, z = 0;
};
struct Vector3 {
  double x = 0, y = 0, z = 0;
};
struct Quaternion {
  double x = 0, y = 0, z = 0, w = 0;
};
struct Pose {
  Point position;
  Quaternion orientation;
};
struct PointStamped {
  std_msgs::Header header;
  Point point;
  typedef std::shared_ptr<const PointStamped> ConstPtr;
  typedef std::shared_ptr<PointStamped> Ptr;
};
struct PoseStamped {
  std_msgs::Header header;
  Pose pose;
  typedef std::shared_ptr<const PoseStamped> ConstPtr;
  typedef std::shared_ptr<PoseStamped> Ptr;
};
struct Twist {
  Vector3 linear, angular;
  typedef std::shared_ptr<const Twist> ConstPtr;
  typedef std::shared_ptr<Twist> Ptr;
};
struct PoseWithCovariance {
  Pose pose;
  double covariance[36] = {0};
};
struct TwistWithCovariance {
  Twist twist;
  double covariance[36] = {0};
};
}  // namespace geometry_msgs
B200NAV_STANDIN_MSG(geometry_msgs, PointStamped)
B200NAV_STANDIN_MSG(geometry_msgs, PoseStamped)
B200NAV_STANDIN_MSG(geometry_msgs, Twist)

namespace nav_msgs {
struct MapMetaData {
  ros::Time map_load_time;
  float resolution = 0;
  uint32_t width = 0, height = 0;
  geometry_msgs::Pose origin;
};
struct OccupancyGrid {
  std_msgs::Header header;
  MapMetaData info;
  std::vector<int8_t> data;
  typedef std::shared_ptr<const OccupancyGrid> ConstPtr;
  typedef std::shared_ptr<OccupancyGrid> Ptr;
};
struct Odometry {
  std_msgs::Header header;
  std::string child_frame_id;
  geometry_msgs::PoseWithCovariance pose;
  geometry_msgs::TwistWithCovariance twist;
  typedef std::shared_ptr<const Odometry> ConstPtr;
  typedef std::shared_ptr<Odometry> Ptr;
};
struct Path {
  std_msgs::Header header;
  std::vector<geometry_msgs::PoseStamped> poses;
  typedef std::shared_ptr<const Path> ConstPtr;
  typedef std::shared_ptr<Path> Ptr;
};
}  // namespace nav_msgs
B200NAV_STANDIN_MSG(nav_msgs, OccupancyGrid)
B200NAV_STANDIN_MSG(nav_msgs, Odometry)
B200NAV_STANDIN_MSG(nav_msgs, Path)

namespace sensor_msgs {
struct LaserScan {
  std_msgs::Header header;
  float angle_min = 0, angle_max = 0, angle_increment = 0, time_increment = 0, scan_time = 0, range_min = 0,
        range_max = 0;
  std::vector<float> ranges, intensities;
  typedef std::shared_ptr<const LaserScan> ConstPtr;
  typedef std::shared_ptr<LaserScan> Ptr;
};
struct Range {
  std_msgs::Header header;
  uint8_t radiation_type = 0;
  float field_of_view = 0, min_range = 0, max_range = 0, range = 0;
  typedef std::shared_ptr<const Range> ConstPtr;
  typedef std::shared_ptr<Range> Ptr;
};
// Only the three channels the reference reads exist ("x", "y", "index"), as plain arrays.
struct PointCloud2 {
  std_msgs::Header header;
  std::vector<float> x, y, z;
  std::vector<int> index;
  typedef std::shared_ptr<const PointCloud2> ConstPtr;
  typedef std::shared_ptr<PointCloud2> Ptr;
};
namespace standin_detail {
inline std::vector<float>& channel(PointCloud2& c, const std::string& n, float*) {
  return n == "x" ? c.x : (n == "y" ? c.y : c.z);
}
inline std::vector<int>& channel(PointCloud2& c, const std::string&, int*) { return c.index; }
}  // namespace standin_detail
template <typename T>
struct PointCloud2Iterator {
  T* p;
  T* e;
  PointCloud2Iterator() : p(nullptr), e(nullptr) {}
  PointCloud2Iterator(PointCloud2& c, const std::string& name) {
    std::vector<T>& v = standin_detail::channel(c, name, (T*)nullptr);
    p = v.data();
    e = v.data() + v.size();
  }
  T& operator*() const { return *p; }
  PointCloud2Iterator& operator++() {
    ++p;
    return *this;
  }
  PointCloud2Iterator end() const {
    PointCloud2Iterator r;
    r.p = e;
    r.e = e;
    return r;
  }
  bool operator!=(const PointCloud2Iterator& o) const { return p != o.p; }
};
}  // namespace sensor_msgs
B200NAV_STANDIN_MSG(sensor_msgs, LaserScan)
B200NAV_STANDIN_MSG(sensor_msgs, Range)
B200NAV_STANDIN_MSG(sensor_msgs, PointCloud2)

namespace move_control {
// generated from move_control/msg/Histogram.msg
struct Histogram {
  uint8_t num_bin = 0;
  std::vector<uint16_t> xData, yData;
  uint16_t yLowThreshold = 0, yHighThreshold = 0;
  std::vector<uint16_t> yBinData;
  typedef std::shared_ptr<const Histogram> ConstPtr;
  typedef std::shared_ptr<Histogram> Ptr;
};
}  // namespace move_control

// ------------------------------------------------------------------------------------------------------------------
// angles (header-only upstream; published definitions)
// ------------------------------------------------------------------------------------------------------------------
namespace angles {
static inline double from_degrees(double d) { return d * M_PI / 180.0; }
static inline double to_degrees(double r) { return r * 180.0 / M_PI; }
static inline double normalize_angle_positive(double a) { return fmod(fmod(a, 2.0 * M_PI) + 2.0 * M_PI, 2.0 * M_PI); }
static inline double normalize_angle(double a) {
  double r = normalize_angle_positive(a);
  if (r > M_PI) r -= 2.0 * M_PI;
  return r;
}
}  // namespace angles

// ------------------------------------------------------------------------------------------------------------------
// tf, message_filters
// ------------------------------------------------------------------------------------------------------------------
namespace tf {
struct TransformException : public std::runtime_error {
  explicit TransformException(const std::string& s) : std::runtime_error(s) {}
};
inline double getYaw(const geometry_msgs::Quaternion& q) {
  return atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));
}
struct TransformListener {
  standin::World* w;
  TransformListener() : w(&standin::world()) {}
  explicit TransformListener(ros::Duration) : w(&standin::world()) {}
  // Only target = the map frame is supported: every frame's pose is given in it.
  bool waitForTransform(const std::string&, const std::string& source, const ros::Time&, const ros::Duration&) const {
    standin::Frame f;
    return w->frame(source, f);
  }
  void transformPoint(const std::string& target, const geometry_msgs::PointStamped& in,
                      geometry_msgs::PointStamped& out) const {
    standin::Frame f;
    if (!w->frame(in.header.frame_id, f)) throw TransformException("unknown frame " + in.header.frame_id);
    double s, c;
    standin::spec_sincos(f.yaw, s, c);
    out.header = in.header;
    out.header.frame_id = target;
    out.point.x = (c * in.point.x - s * in.point.y) + f.x;
    out.point.y = (s * in.point.x + c * in.point.y) + f.y;
    out.point.z = in.point.z;
  }
  void transformPose(const std::string& target, const geometry_msgs::PoseStamped& in,
                     geometry_msgs::PoseStamped& out) const {
    standin::Frame f;
    if (!w->frame(in.header.frame_id, f)) throw TransformException("unknown frame " + in.header.frame_id);
    double s, c;
    standin::spec_sincos(f.yaw, s, c);
    out.header = in.header;
    out.header.frame_id = target;
    out.pose.position.x = (c * in.pose.position.x - s * in.pose.position.y) + f.x;
    out.pose.position.y = (s * in.pose.position.x + c * in.pose.position.y) + f.y;
    out.pose.position.z = in.pose.position.z;
    // orientation: yaw of the frame composed with the (identity) input orientation
    out.pose.orientation.x = 0.0;
    out.pose.orientation.y = 0.0;
    out.pose.orientation.z = sin(f.yaw * 0.5);
    out.pose.orientation.w = cos(f.yaw * 0.5);
  }
};
}  // namespace tf

namespace message_filters {
template <typename M>
struct Subscriber {
  typedef std::function<void(const std::shared_ptr<const M>&)> Cb;
  standin::World* w;
  int id;
  std::vector<Cb> cbs;
  Subscriber(ros::NodeHandle& nh, const std::string& topic, uint32_t) : w(nh.w) {
    id = w->template subscribe<M>(topic, [this](const std::shared_ptr<const M>& m) {
      for (auto& c : cbs) c(m);
    });
  }
  ~Subscriber() { w->unsubscribe(id); }
  Subscriber(const Subscriber&) = delete;
  Subscriber& operator=(const Subscriber&) = delete;
};
}  // namespace message_filters

namespace tf {
// Delivers a message once its frame can be transformed into the target frame; frames are always known here.
template <typename M>
struct MessageFilter {
  typedef std::function<void(const std::shared_ptr<const M>&)> Cb;
  std::shared_ptr<std::vector<Cb>> cbs;
  MessageFilter(message_filters::Subscriber<M>& sub, TransformListener&, const std::string&, uint32_t)
      : cbs(std::make_shared<std::vector<Cb>>()) {
    std::shared_ptr<std::vector<Cb>> mine = cbs;
    sub.cbs.push_back([mine](const std::shared_ptr<const M>& m) {
      for (auto& c : *mine) c(m);
    });
  }
  template <typename F>
  void registerCallback(F f) {
    cbs->push_back(Cb(f));
  }
};
}  // namespace tf

// ------------------------------------------------------------------------------------------------------------------
// laser_geometry
// ------------------------------------------------------------------------------------------------------------------
namespace laser_geometry {
struct LaserProjection {
  void transformLaserScanToPointCloud(const std::string& target, const sensor_msgs::LaserScan& scan,
                                      sensor_msgs::PointCloud2& cloud, tf::TransformListener& tfl) {
    standin::Frame f;
    if (!tfl.w->frame(scan.header.frame_id, f)) throw tf::TransformException("unknown frame " + scan.header.frame_id);
    double sy, cy;
    standin::spec_sincos(f.yaw, sy, cy);
    cloud.header = scan.header;
    cloud.header.frame_id = target;
    cloud.x.clear();
    cloud.y.clear();
    cloud.z.clear();
    cloud.index.clear();
    for (size_t i = 0; i < scan.ranges.size(); ++i) {
      const float r = scan.ranges[i];
      if (!(r >= scan.range_min && r < scan.range_max)) continue;
      const double angle = (double)scan.angle_min + (double)i * (double)scan.angle_increment;
      double sa, ca;
      standin::spec_sincos(angle, sa, ca);
      const float lx = (float)((double)r * ca), ly = (float)((double)r * sa);
      cloud.x.push_back((float)((cy * (double)lx - sy * (double)ly) + f.x));
      cloud.y.push_back((float)((sy * (double)lx + cy * (double)ly) + f.y));
      cloud.z.push_back(0.f);
      cloud.index.push_back((int)i);
    }
  }
};
}  // namespace laser_geometry

#endif
