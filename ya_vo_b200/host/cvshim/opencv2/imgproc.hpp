// cv::cvtColor(GRAY2RGB) and cv::line for Brief::drawMatches (debug drawing only) — shim, see core.hpp
#ifndef YAVO_CVSHIM_IMGPROC_HPP
#define YAVO_CVSHIM_IMGPROC_HPP
#include <cstdlib>

#include "core.hpp"
namespace cv {
inline void cvtColor(const Mat &src, Mat &dst, int /*code: GRAY2RGB*/) {
    Mat out(src.rows, src.cols, CV_8UC3);
    for (int r = 0; r < src.rows; r++)
        for (int c = 0; c < src.cols; c++) {
            const uchar v = src.at<uchar>(r, c);
            out.at<Vec3b>(r, c) = Vec3b(v, v, v);
        }
    dst = out;
}
inline void line(Mat &img, Point p1, Point p2, const Scalar &color, int /*thickness*/ = 1) {
    int x0 = p1.x, y0 = p1.y, x1 = p2.x, y1 = p2.y;
    const int dx = std::abs(x1 - x0), sx = x0 < x1 ? 1 : -1, dy = -std::abs(y1 - y0), sy = y0 < y1 ? 1 : -1;
    int err = dx + dy;
    for (;;) {
        if (x0 >= 0 && y0 >= 0 && x0 < img.cols && y0 < img.rows) {
            if (img.channels() == 3) img.at<Vec3b>(y0, x0) = Vec3b((uchar)color[0], (uchar)color[1], (uchar)color[2]);
            else img.at<uchar>(y0, x0) = (uchar)color[0];
        }
        if (x0 == x1 && y0 == y1) break;
        const int e2 = 2 * err;
        if (e2 >= dy) { err += dy; x0 += sx; }
        if (e2 <= dx) { err += dx; y0 += sy; }
    }
}
}  // namespace cv
#endif
