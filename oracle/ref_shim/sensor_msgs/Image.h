#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include <ros/ros.h>
namespace std_msgs { struct Header { ros::Time stamp; std::string frame_id; uint32_t seq = 0; }; }
namespace sensor_msgs {
struct Image { std_msgs::Header header; uint32_t height = 0, width = 0, step = 0; std::string encoding; std::vector<uint8_t> data; };
using ImagePtr = std::shared_ptr<Image>;
}
