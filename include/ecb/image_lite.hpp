// 8-bit 3-channel image with the two OpenCV operations the reference's debug output uses — at<Vec3b>(loc) = colour and
// cv::circle(img, centre, radius, colour) with thickness 1 — and a dependency-free PNG writer (stored deflate blocks), for the
// per-key-frame images of the CLI (ECC/test/eventCameraCalib.cpp:214-227: SavePath/image/<timestamp>.png = cf->image()).
// Pixels are B, G, R like cv::Mat CV_8UC3; the writer swaps to RGB like cv::imwrite.  Differences from OpenCV: the circle is a
// plain midpoint circle (cv::circle's rasteriser differs in a few pixels), drawChessboardCorners is not drawn.
#ifndef ECB_IMAGE_LITE_HPP
#define ECB_IMAGE_LITE_HPP

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace ecb {

struct Image8UC3 {
    int rows = 0, cols = 0;
    std::vector<uint8_t> data;  // rows x cols x (B, G, R)
    Image8UC3() {}
    Image8UC3(int r, int c) : rows(r), cols(c), data((size_t) r * c * 3, 0) {}
    bool empty() const { return data.empty(); }
    void set(int x, int y, uint8_t b, uint8_t g, uint8_t r) {
        if (x < 0 || y < 0 || x >= cols || y >= rows) return;
        uint8_t *p = &data[((size_t) y * cols + x) * 3];
        p[0] = b, p[1] = g, p[2] = r;
    }
    void circle(int cx, int cy, int radius, uint8_t b, uint8_t g, uint8_t r) {
        if (radius < 0) return;
        int x = radius, y = 0, err = 1 - radius;
        while (x >= y) {
            const int px[8] = {x, y, -y, -x, -x, -y, y, x}, py[8] = {y, x, x, y, -y, -x, -x, -y};
            for (int k = 0; k < 8; ++k) set(cx + px[k], cy + py[k], b, g, r);
            ++y;
            if (err < 0) err += 2 * y + 1;
            else {
                --x;
                err += 2 * (y - x) + 1;
            }
        }
    }
};

inline bool write_png(const std::string &path, const Image8UC3 &img) {
    if (img.empty()) return false;
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t n = 0; n < 256; ++n) {
            uint32_t c = n;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[n] = c;
        }
        init = true;
    }
    auto crc = [&](const std::vector<uint8_t> &v, size_t from) {
        uint32_t c = 0xFFFFFFFFu;
        for (size_t i = from; i < v.size(); ++i) c = table[(c ^ v[i]) & 0xFF] ^ (c >> 8);
        return c ^ 0xFFFFFFFFu;
    };
    auto be32 = [](std::vector<uint8_t> &v, uint32_t x) {
        for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t) (x >> s));
    };
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n'};
    auto chunk = [&](const char *type, const std::vector<uint8_t> &payload) {
        be32(out, (uint32_t) payload.size());
        const size_t from = out.size();
        out.insert(out.end(), type, type + 4);
        out.insert(out.end(), payload.begin(), payload.end());
        be32(out, crc(out, from));
    };
    std::vector<uint8_t> ihdr;
    be32(ihdr, (uint32_t) img.cols);
    be32(ihdr, (uint32_t) img.rows);
    ihdr.insert(ihdr.end(), {8, 2, 0, 0, 0});  // 8 bit, RGB, deflate, no filter, no interlace
    chunk("IHDR", ihdr);
    std::vector<uint8_t> raw;
    raw.reserve((size_t) img.rows * (img.cols * 3 + 1));
    for (int y = 0; y < img.rows; ++y) {
        raw.push_back(0);
        for (int x = 0; x < img.cols; ++x) {
            const uint8_t *p = &img.data[((size_t) y * img.cols + x) * 3];
            raw.push_back(p[2]), raw.push_back(p[1]), raw.push_back(p[0]);
        }
    }
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (size_t pos = 0; pos < raw.size() || pos == 0;) {
        const size_t len = std::min<size_t>(65535, raw.size() - pos);
        z.push_back(pos + len >= raw.size() ? 1 : 0);
        z.push_back((uint8_t) (len & 0xFF)), z.push_back((uint8_t) (len >> 8));
        z.push_back((uint8_t) (~len & 0xFF)), z.push_back((uint8_t) ((~len >> 8) & 0xFF));
        for (size_t i = 0; i < len; ++i) {
            a = (a + raw[pos + i]) % 65521u;
            b = (b + a) % 65521u;
        }
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + len);
        pos += len;
        if (len == 0) break;
    }
    be32(z, (b << 16) | a);
    chunk("IDAT", z);
    chunk("IEND", {});
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

}  // namespace ecb
#endif  // ECB_IMAGE_LITE_HPP
