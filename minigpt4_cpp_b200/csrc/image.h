// image.h — host-side image ingestion behind minigpt4_image_load_from_file / minigpt4_preprocess_image
// (reference minigpt4.cpp:2576-2651; there: OpenCV imread + zurutech/pillow-resize, compiled only with MINIGPT4_BUILD_WITH_OPENCV).
// No third-party decoder is linked: PNG (all colour types, 1-16 bit, Adam7), JPEG (baseline and progressive Huffman, jpeg.cpp) and binary
// PPM/PGM are decoded here.
#pragma once
#include <stdint.h>
#include <string>
#include <thread>
#include <vector>

namespace mg4 {

struct RgbImage { std::vector<uint8_t> px; int w = 0, h = 0; };   // interleaved R,G,B rows, top to bottom

// Decodes a file into 8-bit RGB the way cv::imread(path, IMREAD_COLOR) + BGR2RGB presents it (alpha dropped, grey replicated, 16-bit
// samples reduced to their high byte, palette expanded).  false + message on anything else.
bool decode_image_file(const char *path, RgbImage &out, std::string &err);
bool decode_png(const uint8_t *data, size_t n, RgbImage &out, std::string &err);
bool decode_jpeg(const uint8_t *data, size_t n, RgbImage &out, std::string &err);   // EXIF orientation applied, like cv::imread
bool inflate_zlib(const uint8_t *src, size_t n, std::vector<uint8_t> &dst, size_t expected, std::string &err);

// Pillow's Image.resize(size, BICUBIC) for 8-bit RGB (= ImagingResample, 8bpc path: normalised double coefficients rounded to 22-bit
// fixed point, horizontal pass to uint8, then vertical pass to uint8).  Bit-exact with Pillow (tests/test_image_cpu.py).
void resize_bicubic_u8(const uint8_t *src, int w, int h, uint8_t *dst, int ow, int oh);

// (u8 / 255 - mean) / std, HWC -> planar CHW, with the roundings of the reference's OpenCV expressions (minigpt4.cpp:2621-2636):
// convertTo(CV_32F, 1/255) in float, the two Scalar operations in double, each stored back to float.
void normalize_to_chw(const uint8_t *rgb, int w, int h, float *out);

// f(first_row, last_row) over [0, rows) on up to 16 host threads when the work is worth a thread start (rows are independent)
template <class F> void parallel_rows(int rows, size_t work_per_row, F f) {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt > 16) nt = 16;
    if (nt < 2 || (size_t)rows * work_per_row < (1u << 22)) { f(0, rows); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) { const int a = (int)((size_t)rows * t / nt), b = (int)((size_t)rows * (t + 1) / nt); if (b > a) th.emplace_back(f, a, b); }
    for (auto &x : th) x.join();
}

}  // namespace mg4
