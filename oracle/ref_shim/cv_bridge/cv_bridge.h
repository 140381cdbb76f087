#pragma once
#include <opencv2/core.hpp>
#include <sensor_msgs/Image.h>
namespace cv_bridge {
class CvImage {
public:
    CvImage(const std_msgs::Header& h, const std::string& enc, const cv::Mat& img) : header(h), encoding(enc), image(img) {}
    sensor_msgs::ImagePtr toImageMsg() const      // RadarCPU.cpp:555-558: mono8, rows = range bins, cols = azimuths
    {
        auto msg = std::make_shared<sensor_msgs::Image>();
        msg->header = header; msg->encoding = encoding; msg->height = image.rows; msg->width = image.cols; msg->step = image.cols;
        msg->data.resize((size_t)image.rows * image.cols);
        for (int r = 0; r < image.rows; r++) std::memcpy(&msg->data[(size_t)r * image.cols], image.raw() + (size_t)r * image.step(), image.cols);
        return msg;
    }
    std_msgs::Header header; std::string encoding; cv::Mat image;
};
}
