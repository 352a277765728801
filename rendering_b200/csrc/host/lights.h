// lights.h — light descriptions (reference include/lights.h:19-73).  Data only: the reference's
// virtual illuminate() is evaluated on the device (rendering_b200/csrc/cuda), never on the host.
#pragma once

#include <memory>
#include <vector>

#include "geometry.h"

enum class LightType { BaseLight, DistantLight, PointLight, AreaLight };

class Light {
public:
    virtual ~Light() = default;
    Vec3f color{ 1.0f };
    float intensity = 1.0f;
    LightType type = LightType::BaseLight;
};

class DistantLight : public Light {
public:
    // the default direction is normalised by the constructor (lights.cpp:11-16); a direction read
    // from the scene file is stored as written (scene.cpp:222)
    DistantLight() { type = LightType::DistantLight; dir.normalize(); }
    Vec3f dir{ 0, 0, -1 };
};

class PointLight : public Light {
public:
    PointLight() { type = LightType::PointLight; }
    Vec3f pos{ 0, 0, 0 };
};

class AreaLight : public Light {
public:
    AreaLight() { type = LightType::AreaLight; }
    // samples x samples grid over the parallelogram pos +- i/2 +- j/2 (lights.cpp:46-63)
    std::vector<Vec3f> samplePoints() const;
    Vec3f pos, i, j;
    int samples = 1;
};

using LightsVector = std::vector<std::unique_ptr<Light>>;
