// opencv2/opencv.hpp — stand-in: cv::Mat only appears in the signature of Session::getReloInfo
// (reference include/session.h:21), which has no definition anywhere.
#pragma once
namespace cv {
class Mat {};
}  // namespace cv
