// util.cpp — host helpers (see util.h).
#include "util.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>
#include <vector>

#include "../../../include/rtb.h"

namespace rtb {

static bool fileExists(const std::string& p)
{
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) return false;
    fclose(f);
    return true;
}

std::string resolvePath(const std::string& path, const std::string& assetDir)
{
    if (fileExists(path) || assetDir.empty() || (!path.empty() && path[0] == '/')) return path;
    const std::string alt = assetDir + "/" + path;
    return fileExists(alt) ? alt : path;
}

// The reference extracts numbers with `std::stringstream >> value` and only complains when the
// stream is neither at eof nor good (include/util.h:41-76): leading blanks are skipped, trailing
// text is ignored, a value that does not parse at all is an error.
template <typename T>
static T extract(const std::string& s, const char* what)
{
    T v{};
    std::stringstream ss(s);
    ss >> v;
    if (!ss.eof() && !ss.good()) throw Error(RTB_ERR_PARSE, std::string("cannot parse ") + what + " from '" + s + "'");
    return v;
}

float parseFloat(const std::string& s) { return extract<float>(s, "float"); }
int parseInt(const std::string& s) { return extract<int>(s, "int"); }
bool parseBool(const std::string& s) { return extract<bool>(s, "bool"); }

std::vector<std::string> splitString(const std::string& s, char delim)
{
    std::vector<std::string> cells;
    std::stringstream ss(s);
    std::string cell;
    while (std::getline(ss, cell, delim)) cells.push_back(cell);
    return cells;
}

Vec3f parseVec3(const std::string& s)
{
    const auto cells = splitString(s, ',');
    if (cells.size() != 3) throw Error(RTB_ERR_PARSE, "expected three comma separated numbers, got '" + s + "'");
    return { parseFloat(cells[0]), parseFloat(cells[1]), parseFloat(cells[2]) };
}

void loadBMP(const std::string& filename, std::vector<uint8_t>& rgb, int& width, int& height)
{
    FILE* f = fopen(filename.c_str(), "rb");
    if (!f) throw Error(RTB_ERR_IO, "Could not open .bmp file: " + filename);
    unsigned char header[54];
    if (fread(header, 1, sizeof header, f) != sizeof header) {
        fclose(f);
        throw Error(RTB_ERR_IO, "short .bmp header: " + filename);
    }
    int32_t w, h;
    memcpy(&w, header + 18, 4);
    memcpy(&h, header + 22, 4);
    if (w <= 0 || h <= 0) {
        fclose(f);
        throw Error(RTB_ERR_IO, "unsupported .bmp dimensions: " + filename);
    }
    width = w;
    height = h;
    rgb.assign((size_t)3 * w * h, 0);
    const size_t got = fread(rgb.data(), 1, rgb.size(), f);   // a short file leaves zeros, as the
    (void)got;                                                // reference's unchecked fread would
    fclose(f);
    // B,G,R -> R,G,B (util.cpp:100-106).  A 4096x4096 map is 16.7 M texels: the swap is split over a few threads (the three maps of
    // shotgun.scene were 135 of the 180 ms the scene takes to load)
    const size_t texels = rgb.size() / 3;
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t nThreads = texels >= (1u << 20) ? std::max<size_t>(1, std::min<size_t>(8, hw ? hw : 1)) : 1;
    auto swapRange = [&rgb](size_t a, size_t b) { for (size_t t = a; t < b; ++t) std::swap(rgb[3 * t], rgb[3 * t + 2]); };
    std::vector<std::thread> pool;
    for (size_t k = 1; k < nThreads; ++k) pool.emplace_back(swapRange, texels * k / nThreads, texels * (k + 1) / nThreads);
    swapRange(0, texels / nThreads);
    for (std::thread& t : pool) t.join();
}

namespace {
void writeBMP(const std::string& path, const unsigned char* body, int width, int height)
{
    const int pad = (4 - (width * 3) % 4) % 4;
    const uint32_t dataSize = (uint32_t)((width * 3 + pad) * height);
    unsigned char header[54] = { 0 };
    auto put32 = [&](int off, uint32_t v) { memcpy(header + off, &v, 4); };
    header[0] = 'B'; header[1] = 'M';
    put32(2, 54 + dataSize);
    put32(10, 54);
    put32(14, 40);
    put32(18, (uint32_t)width);
    put32(22, (uint32_t)height);
    header[26] = 1;
    header[28] = 24;
    put32(34, dataSize);
    put32(38, 2835);
    put32(42, 2835);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw Error(RTB_ERR_IO, "Could not open output file " + path);
    fwrite(header, 1, sizeof header, f);
    fwrite(body, 1, dataSize, f);
    fclose(f);
}
} // namespace

void saveBMP(const std::string& path, const float* fb, int width, int height)
{
    const int pad = (4 - (width * 3) % 4) % 4;
    std::vector<unsigned char> rows((size_t)(width * 3 + pad) * height, 0);
    quantiseBGR(fb, width, height, rows.data());
    writeBMP(path, rows.data(), width, height);
}

void quantiseBGR(const float* fb, int width, int height, unsigned char* rows)
{
    const int pad = (4 - (width * 3) % 4) % 4;
    size_t o = 0;
    for (int r = height - 1; r >= 0; --r) {
        const float* src = fb + (size_t)r * width * 3;
        for (int c = 0; c < width; ++c)
            for (int k = 2; k >= 0; --k) {
                const float v = std::max(0.0f, std::min(1.0f, src[c * 3 + k]));
                rows[o++] = (unsigned char)(v * 255);
            }
        o += pad;
    }
}

void saveBMPBytes(const std::string& path, const unsigned char* bgr, int width, int height) { writeBMP(path, bgr, width, height); }

} // namespace rtb
