// Stand-in for <opencv2/opencv.hpp>, as wide as environment/atari/atari.cpp:143-150 needs it: one 8-bit 3-channel matrix over
// caller memory and cv::resize with INTER_AREA. The synthetic screen already has the target resolution, for which OpenCV's
// INTER_AREA resize is a copy; other sizes fall back to a box average (never reached with the shim's emulator).
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstring>
#include <vector>

#define CV_8UC3 16

namespace cv {

enum { INTER_AREA = 3 };

struct Size {
    int width, height;
    Size(int w, int h) : width(w), height(h) {}
};

class Mat {
public:
    Mat() {}
    Mat(int rows, int cols, int /*type*/, void* data) : rows(rows), cols(cols), data(static_cast<unsigned char*>(data)) {}
    template <class T>
    T& at(int i)
    {
        return reinterpret_cast<T*>(data)[i];
    }
    void create(int r, int c)
    {
        rows = r, cols = c;
        own_.assign(static_cast<size_t>(r) * c * 3, 0);
        data = own_.data();
    }
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;

private:
    std::vector<unsigned char> own_;
};

inline void resize(const Mat& src, Mat& dst, Size size, double = 0, double = 0, int = INTER_AREA)
{
    dst.create(size.height, size.width);
    if (src.rows == size.height && src.cols == size.width) {
        std::memcpy(dst.data, src.data, static_cast<size_t>(src.rows) * src.cols * 3);
        return;
    }
    for (int y = 0; y < size.height; ++y) {
        const int y0 = y * src.rows / size.height, y1 = ((y + 1) * src.rows + size.height - 1) / size.height;
        for (int x = 0; x < size.width; ++x) {
            const int x0 = x * src.cols / size.width, x1 = ((x + 1) * src.cols + size.width - 1) / size.width;
            for (int c = 0; c < 3; ++c) {
                int sum = 0;
                for (int yy = y0; yy < y1; ++yy) {
                    for (int xx = x0; xx < x1; ++xx) { sum += src.data[(yy * src.cols + xx) * 3 + c]; }
                }
                const int n = (y1 - y0) * (x1 - x0);
                dst.data[(y * size.width + x) * 3 + c] = static_cast<unsigned char>((sum + n / 2) / n);
            }
        }
    }
}

} // namespace cv
