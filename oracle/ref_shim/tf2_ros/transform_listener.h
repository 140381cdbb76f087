#pragma once
// tf2_ros stand-in: a buffer that replays the poses the harness queued (one per lookup, last one sticks).
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include <ros/ros.h>
#include <sensor_msgs/Image.h>
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::Header header; Transform transform; };
}
namespace tf2 { struct TransformException : public std::runtime_error { using std::runtime_error::runtime_error; }; }
namespace tf2_ros {
class Buffer {
public:
    std::vector<geometry_msgs::TransformStamped> queue;   // filled by the harness
    size_t next = 0;
    bool available = true;
    geometry_msgs::TransformStamped lookupTransform(const std::string&, const std::string&, const ros::Time&)
    {
        if (!available || queue.empty()) throw tf2::TransformException("no transform");
        const size_t i = next < queue.size() ? next : queue.size() - 1;
        next++;
        return queue[i];
    }
};
class TransformListener { public: explicit TransformListener(Buffer&) {} };
}
