// scene.h — Scene / Camera (reference include/scene.h:51-100).  Scene(const std::string&) loads a
// .scene file; render() is the drop-in boundary: it flattens the scene and calls the CUDA backend
// through the C ABI of include/rtb.h.  No CPU renderer exists in this library.
#pragma once

#include <string>

#include "geometry.h"
#include "lights.h"
#include "objects.h"
#include "options.h"

class Camera {
public:
    Vec3f pos{ 0, 0, 0 };
    Vec3f rot{ 0, 0, 0 };
    float fov = 60.0f;
    float zNear = 0.1f, zFar = 100.0f;
    // rotation matrix of Camera::getRay (scene.cpp:24-48); computed eagerly (the reference's lazy
    // initialisation is a data race)
    Matrix44f rotationMatrix() const { return Matrix44f::rotationDeg(rot); }
};

class Scene {
public:
    bool sceneLoadSuccess = true;
    ObjectVector objects;
    LightsVector lights;
    Options options;
    Camera camera;

    int skyboxWidth = 0, skyboxHeight = 0;
    TextureRGB8 skyboxes[6];

    Scene() = default;
    explicit Scene(const std::string& sceneName);          // prints + exit(-1) on error like LOG_ERROR()
    bool loadScene(const std::string& sceneName);          // throws rtb::Error
    bool loadSceneText(const std::string& text, const std::string& assetDir);
    void loadSkybox();

    // Renders through librtb_cuda.so (dlopen'ed on first use) and writes <image_name>.bmp.
    void render();

    std::string assetDir;   // fallback root for relative asset paths (directory of the scene file)
};
