// scene.cpp — .scene parser of the host library (reference Scene::loadScene, src/scene.cpp:62-334,
// and loadSkybox, :336-360).  Same file format, same key set, same defaults and the same ordering
// rule that `name=` loads the mesh with whatever pos/size/rot were given before it.
#include "scene.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "../../../include/rtb.h"
#include "util.h"

void options::resetDefaults()
{
    outputProgress = true;
    useBackfaceCulling = true;
    collectStatistics = false;
    enableOutput = true;
    imageOutput = true;
    useAC = true;
    showAC = false;
    useSkybox = false;
    useTextures = true;
    showNormals = false;
    enableSSAA = true;
}

std::vector<Vec3f> AreaLight::samplePoints() const
{
    std::vector<Vec3f> pts;
    const Vec3f corner = pos - (i / 2.0f) - (j / 2.0f);
    if (samples > 1) {
        for (int a = 0; a < samples; ++a)
            for (int b = 0; b < samples; ++b)
                pts.push_back(corner + (i * (((float)a) / (samples - 1))) + (j * (((float)b) / (samples - 1))));
    } else {
        pts.push_back(pos);
    }
    return pts;
}

namespace {

enum class Block { None, Options, Light, Object };

bool has(const std::string& s, const char* needle) { return s.find(needle) != std::string::npos; }

[[noreturn]] void parseError(const std::string& what, const std::string& line)
{
    throw rtb::Error(RTB_ERR_PARSE, what + ": '" + line + "'");
}

} // namespace

Scene::Scene(const std::string& sceneName)
{
    try {
        sceneLoadSuccess = loadScene(sceneName);
    } catch (const rtb::Error& e) {
        // LOG_ERROR() semantics of the reference (include/util.h:13-19)
        printf("Error: %s\n", e.what());
        std::exit(-1);
    }
}

bool Scene::loadScene(const std::string& scenePath)
{
    if (options::enableOutput) printf("Loading scene %s\n", scenePath.c_str());
    std::ifstream in(scenePath, std::ifstream::in);
    if (!in.good()) throw rtb::Error(RTB_ERR_IO, "Could not open scene file: " + scenePath);
    std::stringstream text;
    text << in.rdbuf();
    const size_t slash = scenePath.find_last_of('/');
    return loadSceneText(text.str(), slash == std::string::npos ? std::string(".") : scenePath.substr(0, slash));
}

bool Scene::loadSceneText(const std::string& text, const std::string& a_assetDir)
{
    assetDir = a_assetDir;
    std::istringstream in(text);
    Block block = Block::None;
    std::unique_ptr<Light> light;
    std::unique_ptr<Object> object;
    std::string line;

    while (in.good()) {
        std::getline(in, line);
        if (line.empty()) continue;

        // any line with '[' closes the block in progress (scene.cpp:96-107)
        if (has(line, "[")) {
            if (block == Block::Light) {
                if (!light) parseError("light block without type", line);
                lights.push_back(std::move(light));
            } else if (block == Block::Object) {
                if (!object) parseError("object block without type", line);
                objects.push_back(std::move(object));
            }
        }
        // "#[" comments out a whole block (scene.cpp:110-116)
        if (has(line, "#[")) {
            do {
                std::getline(in, line);
            } while (in.good() && (!has(line, "[") || has(line, "#[")));
            if (!in.good() && (!has(line, "[") || has(line, "#["))) break;
        }
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        if (line.empty()) continue;

        if (line[0] == '[') {
            if (line == "[options]") block = Block::Options;
            else if (line == "[light]") block = Block::Light;
            else if (line == "[object]") block = Block::Object;
            else if (line == "[end]") break;
            else parseError("unknown block", line);
            continue;
        }
        if (block == Block::None) continue;

        const size_t eq = line.find('=');
        if (eq == std::string::npos) parseError("expected key=value", line);
        std::string key = line.substr(0, eq);
        const std::string value = line.substr(eq + 1);

        if (block == Block::Options) {
            key.erase(std::remove(key.begin(), key.end(), ' '), key.end());
            key.erase(std::remove(key.begin(), key.end(), '\t'), key.end());
            if (key == "outputProgress") options::outputProgress = rtb::parseBool(value);
            else if (key == "useBackfaceCulling") options::useBackfaceCulling = rtb::parseBool(value);
            else if (key == "collectStatistics") options::collectStatistics = rtb::parseBool(value);
            else if (key == "enableOutput") options::enableOutput = rtb::parseBool(value);
            else if (key == "imageOutput") options::imageOutput = rtb::parseBool(value);
            else if (key == "useAC") options::useAC = rtb::parseBool(value);
            else if (key == "showAC") options::showAC = rtb::parseBool(value);
            else if (key == "useSkybox") options::useSkybox = rtb::parseBool(value);
            else if (key == "useTextures") options::useTextures = rtb::parseBool(value);
            else if (key == "showNormals") options::showNormals = rtb::parseBool(value);
            else if (key == "width") options.width = rtb::parseInt(value);
            else if (key == "height") options.height = rtb::parseInt(value);
            else if (key == "fov") camera.fov = rtb::parseFloat(value);
            else if (key == "image_name") options.imageName = value;
            else if (key == "n_workers") options.nWorkers = rtb::parseInt(value);
            else if (key == "max_ray_depth") options.maxRayDepth = rtb::parseInt(value);
            else if (key == "ac_penalty") options.acPenalty = rtb::parseInt(value);
            else if (key == "background_color") options.backgroundColor = rtb::parseVec3(value);
            else if (key == "position") camera.pos = rtb::parseVec3(value);
            else if (key == "rotation") camera.rot = rtb::parseVec3(value);
            else if (key == "skyboxes") {
                const auto names = rtb::splitString(value, ',');
                if (names.size() < 6) parseError("skyboxes needs six paths", line);
                for (int k = 0; k < 6; ++k) {
                    strncpy(options.skyboxNames[k], names[k].c_str(), 63);
                    options.skyboxNames[k][63] = 0;
                }
                options::useSkybox = true;
            } else {
                printf("Scene, unknown key: %s\n", key.c_str());
            }
        } else if (block == Block::Light) {
            if (key == "type") {
                if (value == "distant") light = std::make_unique<DistantLight>();
                else if (value == "point") light = std::make_unique<PointLight>();
                else if (value == "area") light = std::make_unique<AreaLight>();
                continue;
            }
            if (!light) { printf("Error, light type missing\n"); continue; }
            auto need = [&](LightType t) { if (light->type != t) parseError("key does not fit the light type", line); };
            if (key == "color") light->color = rtb::parseVec3(value);
            else if (key == "intensity") light->intensity = rtb::parseFloat(value);
            else if (key == "direction") { need(LightType::DistantLight); static_cast<DistantLight&>(*light).dir = rtb::parseVec3(value); }
            else if (key == "position") { need(LightType::PointLight); static_cast<PointLight&>(*light).pos = rtb::parseVec3(value); }
            else if (key == "pos") { need(LightType::AreaLight); static_cast<AreaLight&>(*light).pos = rtb::parseVec3(value); }
            else if (key == "i") { need(LightType::AreaLight); static_cast<AreaLight&>(*light).i = rtb::parseVec3(value); }
            else if (key == "j") { need(LightType::AreaLight); static_cast<AreaLight&>(*light).j = rtb::parseVec3(value); }
            else if (key == "samples") { need(LightType::AreaLight); static_cast<AreaLight&>(*light).samples = rtb::parseInt(value); }
        } else {
            if (key == "type") {
                if (value == "plane") object = std::make_unique<Plane>();
                else if (value == "sphere") object = std::make_unique<Sphere>();
                else if (value == "mesh") object = std::make_unique<Mesh>();
                continue;
            }
            if (!object) { printf("Error, object type missing\n"); continue; }
            if (key == "color") object->color = rtb::parseVec3(value);
            else if (key == "pos") object->pos = rtb::parseVec3(value);
            else if (key == "material") {
                const auto m = rtb::splitString(value, ',');
                if (m.empty()) parseError("empty material", line);
                if (m[0] == "transparent") {
                    if (m.size() < 2) parseError("transparent needs an index of refraction", line);
                    object->materialType = MaterialType::Transparent;
                    object->indexOfRefraction = rtb::parseFloat(m[1]);
                } else if (m[0] == "reflective") {
                    object->materialType = MaterialType::Reflective;
                } else if (m[0] == "phong") {
                    if (m.size() < 5) parseError("phong needs ambient,diffuse,specular,nSpecular", line);
                    object->materialType = MaterialType::Phong;
                    object->ambient = rtb::parseFloat(m[1]);
                    object->diffuse = rtb::parseFloat(m[2]);
                    object->specular = rtb::parseFloat(m[3]);
                    object->nSpecular = rtb::parseFloat(m[4]);
                }
            } else if (object->objectType == ObjectType::Sphere) {
                auto& sphere = static_cast<Sphere&>(*object);
                if (key == "radius") {
                    sphere.r = rtb::parseFloat(value);
                    sphere.r2 = powf(sphere.r, 2);
                }
            } else if (object->objectType == ObjectType::Plane) {
                if (key == "normal") static_cast<Plane&>(*object).normal = rtb::parseVec3(value);   // stays un-normalised
            } else if (object->objectType == ObjectType::Mesh) {
                auto& mesh = static_cast<Mesh&>(*object);
                if (key == "size") mesh.size = rtb::parseVec3(value);
                else if (key == "rot") mesh.rot = rtb::parseVec3(value);
                else if (key == "name") mesh.loadOBJ(rtb::resolvePath(value, assetDir), options);
                else if (key == "diffuse_map") mesh.diffuseMapLoaded = mesh.loadDiffuseMap(rtb::resolvePath(value, assetDir));
                else if (key == "normal_map") mesh.normalMapLoaded = mesh.loadNormalMap(rtb::resolvePath(value, assetDir));
                else if (key == "specular_map") mesh.specularMapLoaded = mesh.loadSpecularMap(rtb::resolvePath(value, assetDir));
            }
        }
    }

    if (options::useSkybox) loadSkybox();
    return true;
}

void Scene::loadSkybox()
{
    if (!options::useSkybox) return;
    for (int k = 0; k < 6; ++k) {
        int w = 0, h = 0;
        rtb::loadBMP(rtb::resolvePath(options.skyboxNames[k], assetDir), skyboxes[k].rgb, w, h);
        skyboxes[k].width = w;
        skyboxes[k].height = h;
        skyboxWidth = w;     // the reference keeps the dimensions of the LAST face (scene.cpp:343-347)
        skyboxHeight = h;
    }
}
