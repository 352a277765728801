// render.cpp — Scene::render(): the drop-in boundary (reference src/scene.cpp:595-657).
//
// The reference allocates a Vec3f framebuffer, runs launchWorkers + launchSSAA on std::threads and
// saves a BMP.  Here the frame is produced by the CUDA backend behind the C ABI (include/rtb.h);
// librtb_cuda.so is dlopen'ed next to this library on first use so that the host library itself
// stays free of any CUDA dependency.  There is no CPU fallback: a missing backend or device is a
// hard error, like every LOG_ERROR() of the reference.
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../../include/rtb.h"
#include "flatten.h"
#include "scene.h"
#include "util.h"

namespace {

struct Backend {
    decltype(&rtb_create) create = nullptr;
    decltype(&rtb_render_bgr8) renderBgr8 = nullptr;
    decltype(&rtb_render_ac) renderAc = nullptr;
    decltype(&rtb_destroy) destroy = nullptr;
    decltype(&rtb_last_error) lastError = nullptr;
};

std::string ownDirectory()
{
    Dl_info info{};
    if (dladdr((void*)&ownDirectory, &info) && info.dli_fname) {
        const std::string p = info.dli_fname;
        const size_t slash = p.find_last_of('/');
        if (slash != std::string::npos) return p.substr(0, slash);
    }
    return ".";
}

Backend loadBackend()
{
    const char* env = getenv("RTB_CUDA_LIB");
    const std::string path = env ? env : ownDirectory() + "/librtb_cuda.so";
    void* so = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!so) throw rtb::Error(RTB_ERR_CUDA, std::string("cannot load CUDA backend: ") + dlerror());
    Backend b;
    b.create = (decltype(b.create))dlsym(so, "rtb_create");
    b.renderBgr8 = (decltype(b.renderBgr8))dlsym(so, "rtb_render_bgr8");
    b.renderAc = (decltype(b.renderAc))dlsym(so, "rtb_render_ac");
    b.destroy = (decltype(b.destroy))dlsym(so, "rtb_destroy");
    b.lastError = (decltype(b.lastError))dlsym(so, "rtb_last_error");
    if (!b.create || !b.renderBgr8 || !b.renderAc || !b.destroy || !b.lastError)
        throw rtb::Error(RTB_ERR_CUDA, "CUDA backend " + path + " lacks rtb_* symbols");
    return b;
}

double nowMs()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

void report(const char* name, double ms)
{
    if (options::enableOutput) printf("%-18s%lld ms\n", name, (long long)ms);   // Timer's format (include/timer.h:34)
}

} // namespace

void Scene::render()
{
    if (!sceneLoadSuccess) return;
    const double t0 = nowMs();
    try {
        static Backend backend = loadBackend();
        rtb::FlatScene flat;
        rtb::flatten(*this, flat);

        const char* devEnv = getenv("RTB_DEVICE");
        RtbHandle* handle = nullptr;
        if (backend.create(&flat.view, devEnv ? atoi(devEnv) : 0, RTB_CREATE_DEFAULT, &handle) != RTB_OK)
            throw rtb::Error(RTB_ERR_CUDA, backend.lastError());

        // the frame comes back as the BMP's pixel bytes: clamp / quantise / BGR / row flip of saveImage
        // (util.cpp:46-56) run on the device, a quarter of the float framebuffer's bytes cross PCIe
        std::vector<unsigned char> pixelBytes((size_t)((options.width * 3 + 3) & ~(size_t)3) * options.height, 0);
        RtbStats stats{};
        int rc;
        if (options::showAC) {
            // debug view of the tree (scene.cpp:607-635): float frame from the backend, converted here
            std::vector<float> frame((size_t)options.width * options.height * 3, 0.0f);
            rc = backend.renderAc(handle, frame.data(), nullptr, 0, nullptr, &stats);
            if (rc == RTB_OK) rtb::quantiseBGR(frame.data(), (int)options.width, (int)options.height, pixelBytes.data());
        } else {
            rc = backend.renderBgr8(handle, 0, (int)options.height, pixelBytes.data(), 0, nullptr, &stats);
        }
        const std::string err = rc == RTB_OK ? "" : backend.lastError();
        backend.destroy(handle);
        if (rc != RTB_OK) throw rtb::Error(rc, err);

        report("Render scene", stats.msPass1);
        report("Sobel filter", stats.msSobel);
        report("MSAA", stats.msSobel + stats.msSSAA);
        if (options::imageOutput) {
            const std::string path = options.imageName + ".bmp";
            rtb::saveBMPBytes(path, pixelBytes.data(), (int)options.width, (int)options.height);
            printf("Successfully wrote to output file %s\n", path.c_str());
        }
        if (options::collectStatistics) {
            printf("Statistics:\nRays casted:                        %10llu\n", (unsigned long long)stats.rays);
        }
    } catch (const rtb::Error& e) {
        printf("Error: %s\n", e.what());
        std::exit(-1);
    }
    report("Total time", nowMs() - t0);
    if (options::enableOutput) printf("\n");
}
