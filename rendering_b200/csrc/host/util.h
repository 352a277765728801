// util.h — host helpers: error type, path resolution, BMP read/write, number parsing.
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "geometry.h"

namespace rtb {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// cwd-relative first (reference behaviour), then relative to `assetDir`
std::string resolvePath(const std::string& path, const std::string& assetDir);

// stream-extraction semantics of the reference's strTo* helpers (include/util.h:41-76)
float parseFloat(const std::string& s);
int parseInt(const std::string& s);
bool parseBool(const std::string& s);
Vec3f parseVec3(const std::string& s);
std::vector<std::string> splitString(const std::string& s, char delim);

// 24-bit BMP as the reference reads it: 54-byte header, width/height at 18/22, 3*w*h bytes,
// B<->R swapped (util.cpp:78-113).  Throws Error(RTB_ERR_IO).
void loadBMP(const std::string& filename, std::vector<uint8_t>& rgb, int& width, int& height);

// saveImage contract (util.cpp:15-76): bottom-up rows, BGR, byte = (uint8)(clamp(v,0,1)*255).
void saveBMP(const std::string& path, const float* fb, int width, int height);
// the conversion alone: out holds height rows of (3*width padded to 4) bytes, bottom-up, BGR (pad bytes untouched)
void quantiseBGR(const float* fb, int width, int height, unsigned char* out);
// same file from already-converted pixel bytes (rtb_render_bgr8: bottom-up rows, BGR, padded to 4)
void saveBMPBytes(const std::string& path, const unsigned char* bgr, int width, int height);

} // namespace rtb
