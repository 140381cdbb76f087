// Minimal roscpp stand-in: just what Radar.cpp / RadarCPU.cpp / ros_helper.h touch.
#ifndef RR_SHIM_ROS_H
#define RR_SHIM_ROS_H
#include <cstdint>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include <radarays_ros/RadarMaterials.h>
namespace XmlRpc { class XmlRpcValue {}; }
namespace ros {
struct Duration { double sec = 0; double toSec() const { return sec; } };
struct Time {
    double sec = 0;
    Time() = default;
    explicit Time(double s) : sec(s) {}
    static Time now() { return Time(0.0); }
    uint64_t toNSec() const { return (uint64_t)(sec * 1e9); }
    Duration operator-(const Time& o) const { return Duration{sec - o.sec}; }
    Time operator+(const Duration& d) const { return Time(sec + d.sec); }
};
inline void spinOnce() {}
class NodeHandle {
public:
    NodeHandle() = default;
    explicit NodeHandle(const std::string&) {}
    // the three parameters Radar::loadParams reads (Radar.cpp:220-226), filled by the harness
    radarays_ros::RadarMaterials materials;
    std::vector<int> object_materials;
    int material_id_air = 0;
    bool getParam(const std::string& key, std::vector<int>& v) const { if (key == "object_materials") { v = object_materials; return true; } return false; }
    bool getParam(const std::string& key, int& v) const { if (key == "material_id_air") { v = material_id_air; return true; } return false; }
    bool getParam(const std::string&, std::string&) const { return false; }
    bool getParam(const std::string&, XmlRpc::XmlRpcValue&) const { return false; }
};
} // namespace ros
#define ROS_INFO(...) do { } while (0)
#define ROS_INFO_STREAM(x) do { } while (0)
#define ROS_WARN_STREAM(x) do { } while (0)
#endif
