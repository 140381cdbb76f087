#pragma once
#include <string>
namespace sensor_msgs { namespace image_encodings { static const std::string MONO8 = "mono8"; } }
