// stub: stands in for the OpenCV module header of the same name; the hot-path sources only need the core types
#pragma once
#include "opencv4/opencv2/core.hpp"
