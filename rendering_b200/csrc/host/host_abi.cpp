// host_abi.cpp — extern "C" entry points of librtb_host.so (include/rtb.h, "host side").
#include <cstring>
#include <memory>
#include <mutex>
#include <string>

#include "../../../include/rtb.h"
#include "flatten.h"
#include "scene.h"
#include "util.h"

struct RtbHostScene {
    Scene scene;
    rtb::FlatScene flat;
};

namespace {
thread_local std::string g_lastError;
// the reference's option switches are process globals (include/options.h:26-36); loading is
// serialised so two loads cannot interleave their writes
std::mutex g_loadMutex;

template <typename F>
int guarded(F&& f)
{
    try {
        return f();
    } catch (const rtb::Error& e) {
        g_lastError = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return RTB_ERR_PARSE;
    }
}
} // namespace

extern "C" {

int rtb_scene_load(const char* scenePath, RtbHostScene** out)
{
    if (!scenePath || !out) { g_lastError = "null argument"; return RTB_ERR_ARG; }
    *out = nullptr;
    return guarded([&]() {
        std::lock_guard<std::mutex> lock(g_loadMutex);
        options::resetDefaults();
        options::enableOutput = false;
        auto hs = std::make_unique<RtbHostScene>();
        hs->scene.loadScene(scenePath);
        rtb::flatten(hs->scene, hs->flat);
        *out = hs.release();
        return RTB_OK;
    });
}

int rtb_scene_parse(const char* sceneText, const char* assetDir, RtbHostScene** out)
{
    if (!sceneText || !out) { g_lastError = "null argument"; return RTB_ERR_ARG; }
    *out = nullptr;
    return guarded([&]() {
        std::lock_guard<std::mutex> lock(g_loadMutex);
        options::resetDefaults();
        options::enableOutput = false;
        auto hs = std::make_unique<RtbHostScene>();
        hs->scene.loadSceneText(sceneText, assetDir ? assetDir : "");
        rtb::flatten(hs->scene, hs->flat);
        *out = hs.release();
        return RTB_OK;
    });
}

const RtbScene* rtb_scene_view(const RtbHostScene* hs) { return hs ? &hs->flat.view : nullptr; }

const char* rtb_scene_image_name(const RtbHostScene* hs) { return hs ? hs->flat.imageName.c_str() : nullptr; }

void rtb_scene_free(RtbHostScene* hs) { delete hs; }

int rtb_scene_tree_stats(const RtbHostScene* hs, int mesh, int64_t out[6])
{
    if (!hs || !out) { g_lastError = "null argument"; return RTB_ERR_ARG; }
    int k = 0;
    for (const auto& o : hs->scene.objects) {
        if (o->objectType != ObjectType::Mesh) continue;
        if (k++ == mesh) {
            const rtb::TreeStats st = rtb::treeStats(static_cast<const Mesh&>(*o));
            out[0] = st.nodes; out[1] = st.leaves; out[2] = st.refs; out[3] = st.maxLeaf; out[4] = st.maxDepth; out[5] = st.trisOutsideRoot;
            return RTB_OK;
        }
    }
    g_lastError = "mesh index out of range";
    return RTB_ERR_ARG;
}

int rtb_save_bmp(const char* path, const float* fb, int width, int height)
{
    if (!path || !fb || width <= 0 || height <= 0) { g_lastError = "bad argument"; return RTB_ERR_ARG; }
    return guarded([&]() { rtb::saveBMP(path, fb, width, height); return RTB_OK; });
}

int rtb_camera_from_angles(const float pos[3], const float rotDeg[3], float fovDeg, int width, int height, RtbCamera* out)
{
    if (!pos || !rotDeg || !out || width <= 0 || height <= 0) { g_lastError = "bad argument"; return RTB_ERR_ARG; }
    Camera cam;
    cam.pos = Vec3f{ pos[0], pos[1], pos[2] };
    cam.rot = Vec3f{ rotDeg[0], rotDeg[1], rotDeg[2] };
    cam.fov = fovDeg;
    *out = rtb::flattenCamera(cam, (size_t)width, (size_t)height);
    return RTB_OK;
}

int rtb_save_bmp_bgr8(const char* path, const uint8_t* bgr, int width, int height)
{
    if (!path || !bgr || width <= 0 || height <= 0) { g_lastError = "bad argument"; return RTB_ERR_ARG; }
    return guarded([&]() { rtb::saveBMPBytes(path, bgr, width, height); return RTB_OK; });
}

const char* rtb_host_last_error(void) { return g_lastError.c_str(); }

} // extern "C"
