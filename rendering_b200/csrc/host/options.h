// options.h — per-scene Options and the process-global switches, same names and defaults as the
// reference's include/options.h:9-37 (the .scene [options] keys write into these).
#pragma once

#include <cstddef>
#include <string>

#include "geometry.h"

class Options {
public:
    size_t width = 800, height = 600;
    float bias = 0.0001f;
    int maxRayDepth = 10;
    int nWorkers = 32;            // kept for file-format compatibility; the GPU path ignores it
    Vec3f backgroundColor{ 0.0f, 0.0f, 0.0f };
    int acPenalty = 1;
    char skyboxNames[6][64] = { { 0 } };
    std::string imageName = "out";
};

namespace options {
inline bool outputProgress = true;
inline bool useBackfaceCulling = true;
inline bool collectStatistics = false;
inline bool enableOutput = true;
inline bool imageOutput = true;
inline bool useAC = true;
inline bool showAC = false;
inline bool useSkybox = false;
inline bool useTextures = true;
inline bool showNormals = false;
inline bool enableSSAA = true;
// restore the defaults above (the reference never needs this: one scene per process)
void resetDefaults();
}
