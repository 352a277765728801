// rtb_api.cu — extern "C" entry points of librtb_cuda.so (include/rtb.h, "device side") and the
// host-side frame scheduler that strings the wavefront kernels of rtb_kernels.cuh together.
//
// One RtbHandle owns: the scene resident in HBM (reference-tree nodes, triangle slots in leaf
// order, shading attributes, RGBA8 textures as cudaTextureObjects), the per-level queues, the
// colour-slot array whose first w*h entries are the framebuffer, and one CUDA stream.
// There is no CPU rendering path anywhere in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/rtb.h"
#include "rtb_kernels.cuh"
#include "scene_pack.h"

namespace {

thread_local std::string g_err;

struct CudaError { cudaError_t e; const char* what; };
#define CK(call)                                                         \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) throw CudaError{ e__, #call };           \
    } while (0)

// growable device buffer
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    void reserve(size_t need, cudaStream_t st, bool preserve)
    {
        if (need <= bytes) return;
        const size_t nb = std::max(need, bytes + bytes / 2);
        void* np = nullptr;
        CK(cudaMalloc(&np, nb));
        if (p) {
            if (preserve) CK(cudaMemcpyAsync(np, p, bytes, cudaMemcpyDeviceToDevice, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaFree(p));
        }
        p = np;
        bytes = nb;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct TextureRes {
    cudaArray_t array = nullptr;
    cudaTextureObject_t tex = 0;
    void* linear = nullptr;
};

struct QueueBufs {
    DevBuf o, d, dest;
    rtk::RayQueue view() const { return { o.as<float4>(), d.as<float4>(), dest.as<int>() }; }
    void reserve(size_t n, cudaStream_t st, bool preserve)
    {
        o.reserve(n * sizeof(float4), st, preserve);
        d.reserve(n * sizeof(float4), st, preserve);
        dest.reserve(n * sizeof(int), st, preserve);
    }
    void release() { o.release(); d.release(); dest.release(); }
};

} // namespace

struct RtbHandle {
    int device = 0;
    uint32_t createFlags = 0;
    cudaStream_t ownStream = nullptr;
    int smCount = 148;

    rt::Scene scene{};                 // header with DEVICE pointers
    std::vector<void*> allocations;    // scene-lifetime device allocations
    std::vector<TextureRes> textures;

    QueueBufs rays[2];
    DevBuf hitTuv, hitObj, surfP, surfN, surfC, vis, interiors, slots, flagged, rowsA, rowsB, userRays, outStage;
    rtk::Counters* dCtr = nullptr;
    rtk::Counters* hCtr = nullptr;     // pinned
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };

    // per-kernel timing: (kind, start, stop) spans recorded on the render stream, resolved at the end of a call
    struct Span { int kind; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> eventPool;
    size_t eventsUsed = 0;

    // per-call bookkeeping
    int slotCount = 0;
    int interiorCount = 0;
    RtbStats stats{};
};

namespace {

template <typename T>
T* upload(RtbHandle* h, const T* src, size_t n)
{
    if (n == 0) return nullptr;
    void* d = nullptr;
    CK(cudaMalloc(&d, n * sizeof(T)));
    h->allocations.push_back(d);
    CK(cudaMemcpy(d, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return static_cast<T*>(d);
}

rt::Image uploadImage(RtbHandle* h, const RtbImage& im)
{
    rt::Image out{};
    const std::vector<unsigned char> rgba = rtpack::packRGBA(im);
    if (rgba.empty()) return out;
    TextureRes res;
    const cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    CK(cudaMallocArray(&res.array, &fmt, im.width, im.height));
    CK(cudaMemcpy2DToArray(res.array, 0, 0, rgba.data(), (size_t)im.width * 4, (size_t)im.width * 4, im.height, cudaMemcpyHostToDevice));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = res.array;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;          // the reference samples nearest texel, no filtering
    td.readMode = cudaReadModeElementType;        // raw bytes; /256 happens in fetchTexel
    td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&res.tex, &rd, &td, nullptr));
    h->textures.push_back(res);
    out.tex = (unsigned long long)res.tex;
    out.rgba = nullptr;
    out.w = im.width;
    out.h = im.height;
    return out;
}

// Fast-path data of one mesh: search BVH over its unique triangles (bvh_build.h) and the tables that
// let the kernel evaluate the reference tree's eligibility rule for a single triangle.
void buildFastPath(RtbHandle* h, const RtbMesh& m, rt::Mesh& d)
{
    rtpack::FastPath fp;
    rtpack::packFastPath(m, fp);
    if (fp.maxDepth > rtk::kStackDepth) throw std::runtime_error("search BVH deeper than the traversal stack (64)");
    d.bvhNodes = reinterpret_cast<const float4*>(upload(h, fp.nodes.data(), fp.nodes.size()));
    d.bvhTris = upload(h, fp.tris.data(), fp.tris.size());
    d.triRefOff = upload(h, fp.triRefOff.data(), fp.triRefOff.size());
    d.triRefs = upload(h, fp.triRefs.data(), fp.triRefs.size());
    d.parent = upload(h, fp.parent.data(), fp.parent.size());
}

int gridFor(const RtbHandle* h, long long n, int block = rtk::kBlock, int perSm = 16)
{
    const long long blocks = (n + block - 1) / block;
    const long long cap = (long long)h->smCount * perSm;
    return (int)std::max(1LL, std::min(blocks, cap));
}

// persistent kernels: one resident wave (6 CTAs of 128 threads per SM with the 32 KiB stack), never more
// CTAs than there is work for
int persistentGrid(const RtbHandle* h, long long n)
{
    const long long blocks = (n + rtk::kBlock - 1) / rtk::kBlock;
    return (int)std::max(1LL, std::min(blocks, (long long)h->smCount * 6));
}

void launchCheck() { CK(cudaGetLastError()); }

cudaEvent_t takeEvent(RtbHandle* h)
{
    if (h->eventsUsed == h->eventPool.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        h->eventPool.push_back(e);
    }
    return h->eventPool[h->eventsUsed++];
}

// Brackets one kernel launch with CUDA events on the launching stream (RtbStats.msKernel).
struct KernelSpan {
    RtbHandle* h; cudaStream_t st; int kind; cudaEvent_t a, b;
    KernelSpan(RtbHandle* h_, cudaStream_t st_, int kind_) : h(h_), st(st_), kind(kind_)
    {
        a = takeEvent(h); b = takeEvent(h);
        CK(cudaEventRecord(a, st));
    }
    void done()
    {
        launchCheck();
        CK(cudaEventRecord(b, st));
        h->spans.push_back({ kind, a, b });
        h->stats.kernelLaunches++;
        h->stats.launchesKernel[kind]++;
    }
};

void resolveSpans(RtbHandle* h)
{
    for (const auto& sp : h->spans) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) h->stats.msKernel[sp.kind] += ms;
    }
    h->spans.clear();
    h->eventsUsed = 0;
}

// Makes room for a level of n rays: hit / surface / visibility records for n rays, up to 2n rays in
// the other queue, up to n more interior records and 2n more colour slots.
void reserveLevel(RtbHandle* h, cudaStream_t st, int cur, long long n)
{
    const long long S = std::max(1, h->scene.shadowRaysPerHit);
    if (n > (1LL << 28)) throw CudaError{ cudaErrorMemoryAllocation, "ray queue larger than 2^28 rays" };
    h->rays[cur].reserve((size_t)n, st, true);
    h->rays[cur ^ 1].reserve((size_t)(2 * n), st, false);
    h->hitTuv.reserve((size_t)n * sizeof(float4), st, false);
    h->hitObj.reserve((size_t)n * sizeof(int), st, false);
    h->surfP.reserve((size_t)n * sizeof(float4), st, false);
    h->surfN.reserve((size_t)n * sizeof(float4), st, false);
    h->surfC.reserve((size_t)n * sizeof(float4), st, false);
    h->vis.reserve((size_t)(n * S), st, false);
    h->interiors.reserve((size_t)(h->interiorCount + n) * sizeof(rtk::Interior), st, true);
    h->slots.reserve((size_t)(h->slotCount + 2 * n) * 3 * sizeof(float), st, true);
}

// Runs castRay for the n0 rays sitting in queue 0, level by level, then folds the interior records.
void runLevels(RtbHandle* h, cudaStream_t st, long long n0, uint64_t& tracedFirstLevel)
{
    const bool count = h->createFlags & RTB_CREATE_COUNTERS;
    const bool exact = h->createFlags & RTB_CREATE_EXACT_WALK;
    const rt::Scene& sc = h->scene;
    std::vector<std::pair<int, int>> levelRanges;
    int cur = 0;
    long long n = n0;
    tracedFirstLevel += (uint64_t)n0;
    const int interiorStart = h->interiorCount;
    (void)interiorStart;
    for (int depth = 0; depth <= sc.maxRayDepth && n > 0; ++depth) {
        reserveLevel(h, st, cur, n);
        if (depth > 0) h->stats.secondaryRays += (uint64_t)n;
        // reset the per-level counters, keep the running ones
        h->hCtr->nextRays = 0;
        h->hCtr->surfaces = 0;
        h->hCtr->interiors = h->interiorCount;
        h->hCtr->slots = h->slotCount;
        CK(cudaMemcpyAsync(h->dCtr, h->hCtr, offsetof(rtk::Counters, ssaaPixels), cudaMemcpyHostToDevice, st));
        h->stats.h2dBytes += offsetof(rtk::Counters, ssaaPixels);

        const rtk::RayQueue q = h->rays[cur].view(), next = h->rays[cur ^ 1].view();
        const rtk::HitQueue hits{ h->hitTuv.as<float4>(), h->hitObj.as<int>() };
        const rtk::SurfQueue surf{ h->surfP.as<float4>(), h->surfN.as<float4>(), h->surfC.as<float4>() };
        const int nextCap = (int)std::min<size_t>(h->rays[cur ^ 1].dest.bytes / sizeof(int), 1u << 30);
        const int interiorCap = (int)std::min<size_t>(h->interiors.bytes / sizeof(rtk::Interior), 1u << 30);
        const int slotCap = (int)std::min<size_t>(h->slots.bytes / (3 * sizeof(float)), 1u << 30);

        {
            KernelSpan ks(h, st, RTB_K_TRACE);
            if (count) rtk::k_trace<rtk::MODE_COUNT><<<gridFor(h, n), rtk::kBlock, 0, st>>>(sc, q, (int)n, hits, h->dCtr);
            else if (exact) rtk::k_trace<rtk::MODE_EXACT><<<gridFor(h, n), rtk::kBlock, 0, st>>>(sc, q, (int)n, hits, h->dCtr);
            else {
                CK(cudaMemsetAsync(&h->dCtr->walkCursor[0], 0, 2 * sizeof(unsigned long long), st));
                rtk::k_walk<false><<<persistentGrid(h, n), rtk::kBlock, 0, st>>>(sc, q, (int)n, hits, surf, h->vis.as<unsigned char>(), h->dCtr, &h->dCtr->walkCursor[0]);
            }
            ks.done();
        }
        {
            KernelSpan ks(h, st, RTB_K_SURFACE);
            rtk::k_surface<<<gridFor(h, n), rtk::kBlock, 0, st>>>(sc, q, (int)n, hits, surf, h->slots.as<float>(), h->dCtr);
            ks.done();
        }
        if (!(sc.flags & rt::FLAG_SHOW_NORMALS)) {
            if (sc.shadowRaysPerHit > 0) {
                const long long maxShadow = n * sc.shadowRaysPerHit;
                KernelSpan ks(h, st, RTB_K_SHADOW);
                if (count) rtk::k_shadow<rtk::MODE_COUNT><<<gridFor(h, maxShadow), rtk::kBlock, 0, st>>>(sc, q, surf, h->vis.as<unsigned char>(), h->dCtr);
                else if (exact) rtk::k_shadow<rtk::MODE_EXACT><<<gridFor(h, maxShadow), rtk::kBlock, 0, st>>>(sc, q, surf, h->vis.as<unsigned char>(), h->dCtr);
                else rtk::k_walk<true><<<persistentGrid(h, maxShadow), rtk::kBlock, 0, st>>>(sc, q, 0, hits, surf, h->vis.as<unsigned char>(), h->dCtr, &h->dCtr->walkCursor[1]);
                ks.done();
            }
            KernelSpan ks(h, st, RTB_K_SHADE);
            rtk::k_shade<<<gridFor(h, n), rtk::kBlock, 0, st>>>(sc, q, surf, h->vis.as<unsigned char>(), depth, next, nextCap,
                h->interiors.as<rtk::Interior>(), interiorCap, h->slots.as<float>(), slotCap, h->dCtr);
            ks.done();
        }
        CK(cudaMemcpyAsync(h->hCtr, h->dCtr, sizeof(rtk::Counters), cudaMemcpyDeviceToHost, st));
        h->stats.d2hBytes += sizeof(rtk::Counters);
        CK(cudaStreamSynchronize(st));
        if (h->hCtr->overflow) throw CudaError{ cudaErrorMemoryAllocation, "wavefront queue overflow" };
        h->stats.shadowRays += (uint64_t)h->hCtr->surfaces * (uint64_t)sc.shadowRaysPerHit;
        h->stats.shadowRaysSkipped = h->hCtr->shadowSkipped;
        h->stats.levels = std::max<uint32_t>(h->stats.levels, (uint32_t)depth + 1);
        levelRanges.push_back({ h->interiorCount, h->hCtr->interiors });
        h->interiorCount = h->hCtr->interiors;
        h->slotCount = h->hCtr->slots;
        n = h->hCtr->nextRays;
        cur ^= 1;
    }
    for (int l = (int)levelRanges.size() - 1; l >= 0; --l) {
        const int first = levelRanges[l].first, last = levelRanges[l].second;
        if (last > first) {
            KernelSpan ks(h, st, RTB_K_COMBINE);
            rtk::k_combine<<<gridFor(h, last - first), rtk::kBlock, 0, st>>>(h->interiors.as<rtk::Interior>(), first, last, h->slots.as<float>());
            ks.done();
        }
    }
    // the queue holding level 0 must be queue 0 again for the next caller
    if (h->rays[0].o.p == nullptr) h->rays[0].reserve(1, st, false);
}

void beginCall(RtbHandle* h)
{
    CK(cudaSetDevice(h->device));
    h->stats = RtbStats{};
    h->spans.clear();
    h->eventsUsed = 0;
    h->slotCount = 0;
    h->interiorCount = 0;
    std::memset(h->hCtr, 0, sizeof(rtk::Counters));
}

float elapsed(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

// Renders the rows in `owned` (ascending) into `out` (compact, owned rows in order).
int renderRows(RtbHandle* h, const std::vector<int>& owned, float* fb, float* pass1, int fbOnDevice, void* stream, RtbStats* statsOut)
{
    cudaStream_t st = stream ? (cudaStream_t)stream : h->ownStream;
    beginCall(h);
    const rt::Scene& sc = h->scene;
    const int w = sc.width, ht = sc.height;
    const size_t framePixels = (size_t)w * ht;

    // pass-1 rows: owned rows plus a one-row halo for the Sobel window, minus the never-rendered last row
    std::vector<int> p1rows;
    {
        std::vector<char> need(ht, 0);
        const bool ssaa = sc.flags & rt::FLAG_SSAA;
        for (int y : owned)
            for (int dy = ssaa ? -1 : 0; dy <= (ssaa ? 1 : 0); ++dy)
                if (y + dy >= 0 && y + dy < ht - 1) need[y + dy] = 1;
        for (int y = 0; y < ht; ++y) if (need[y]) p1rows.push_back(y);
    }
    h->rowsA.reserve(std::max<size_t>(1, p1rows.size()) * sizeof(int), st, false);
    h->rowsB.reserve(std::max<size_t>(1, owned.size()) * sizeof(int), st, false);
    CK(cudaMemcpyAsync(h->rowsA.p, p1rows.data(), p1rows.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->rowsB.p, owned.data(), owned.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    h->stats.h2dBytes += (p1rows.size() + owned.size()) * sizeof(int);

    CK(cudaEventRecord(h->ev[0], st));
    h->slotCount = (int)framePixels;
    h->slots.reserve(framePixels * 3 * sizeof(float), st, false);
    CK(cudaMemsetAsync(h->slots.p, 0, framePixels * 3 * sizeof(float), st));   // Vec3f() zero-init (scene.cpp:599)
    CK(cudaMemsetAsync(h->dCtr, 0, sizeof(rtk::Counters), st));

    const long long nPixels = (long long)p1rows.size() * (w - 1);
    if (nPixels > 0) {
        const long long n0 = rtk::raygenPaddedCount(w, (int)p1rows.size());   // whole 8x4 tiles, padding lanes idle
        reserveLevel(h, st, 0, n0);
        KernelSpan ks(h, st, RTB_K_RAYGEN);
        rtk::k_raygen<<<gridFor(h, n0), rtk::kBlock, 0, st>>>(sc, h->rowsA.as<int>(), (int)p1rows.size(), h->rays[0].view());
        ks.done();
        uint64_t padded = 0;
        runLevels(h, st, n0, padded);
        h->stats.primaryRays += (uint64_t)nPixels;
    }
    CK(cudaEventRecord(h->ev[1], st));

    const size_t outFloats = owned.size() * (size_t)w * 3;
    auto emit = [&](float* dst) {
        if (!dst || owned.empty()) return;
        float* target = dst;
        if (!fbOnDevice) {
            h->outStage.reserve(outFloats * sizeof(float), st, false);
            target = h->outStage.as<float>();
        }
        KernelSpan ks(h, st, RTB_K_OUTPUT);
        rtk::k_gather_rows<<<gridFor(h, (long long)outFloats), rtk::kBlock, 0, st>>>(h->slots.as<float>(), w, h->rowsB.as<int>(), (int)owned.size(), target);
        ks.done();
        if (!fbOnDevice) {
            CK(cudaMemcpyAsync(dst, target, outFloats * sizeof(float), cudaMemcpyDeviceToHost, st));
            h->stats.d2hBytes += outFloats * sizeof(float);
        }
    };
    if (pass1) { emit(pass1); if (!fbOnDevice) CK(cudaStreamSynchronize(st)); }

    int nFlagged = 0;
    if ((sc.flags & rt::FLAG_SSAA) && !owned.empty()) {
        h->flagged.reserve(owned.size() * (size_t)w * sizeof(int), st, false);
        {
            KernelSpan ks(h, st, RTB_K_SOBEL);
            rtk::k_sobel<<<gridFor(h, (long long)owned.size() * w), rtk::kBlock, 0, st>>>(w, ht, h->slots.as<float>(), h->rowsB.as<int>(),
                (int)owned.size(), h->flagged.as<int>(), h->dCtr);
            ks.done();
        }
        CK(cudaMemcpyAsync(h->hCtr, h->dCtr, sizeof(rtk::Counters), cudaMemcpyDeviceToHost, st));
        h->stats.d2hBytes += sizeof(rtk::Counters);
        CK(cudaEventRecord(h->ev[2], st));
        CK(cudaStreamSynchronize(st));
        nFlagged = h->hCtr->ssaaPixels;
        h->stats.ssaaPixels = (uint64_t)nFlagged;
        if (nFlagged > 0) {
            const long long n1 = 4LL * nFlagged;
            const int slotBase = h->slotCount;
            h->slotCount += (int)n1;
            reserveLevel(h, st, 0, n1);
            {
                KernelSpan ks(h, st, RTB_K_RAYGEN);
                rtk::k_ssaa_gen<<<gridFor(h, n1), rtk::kBlock, 0, st>>>(sc, h->flagged.as<int>(), nFlagged, slotBase, h->rays[0].view());
                ks.done();
            }
            runLevels(h, st, n1, h->stats.primaryRays);
            KernelSpan ks(h, st, RTB_K_OUTPUT);
            rtk::k_ssaa_resolve<<<gridFor(h, nFlagged), rtk::kBlock, 0, st>>>(h->flagged.as<int>(), nFlagged, slotBase, h->slots.as<float>());
            ks.done();
        }
    } else {
        CK(cudaEventRecord(h->ev[2], st));
    }
    emit(fb);
    CK(cudaEventRecord(h->ev[3], st));
    CK(cudaStreamSynchronize(st));

    if (h->createFlags & RTB_CREATE_COUNTERS) {
        CK(cudaMemcpy(h->hCtr, h->dCtr, sizeof(rtk::Counters), cudaMemcpyDeviceToHost));
        h->stats.boxTestsShadow = h->hCtr->boxTestsShadow;
        h->stats.triTestsShadow = h->hCtr->triTestsShadow;
        h->stats.boxTests = h->hCtr->boxTests + h->hCtr->boxTestsShadow;
        h->stats.triTests = h->hCtr->triTests + h->hCtr->triTestsShadow;
    }
    resolveSpans(h);
    h->stats.rays = h->stats.primaryRays + h->stats.secondaryRays + h->stats.shadowRays;
    h->stats.msPass1 = elapsed(h->ev[0], h->ev[1]);
    h->stats.msSobel = elapsed(h->ev[1], h->ev[2]);
    h->stats.msSSAA = elapsed(h->ev[2], h->ev[3]);
    h->stats.msTotal = elapsed(h->ev[0], h->ev[3]);
    if (statsOut) *statsOut = h->stats;
    return RTB_OK;
}

template <typename F>
int guarded(F&& f)
{
    try {
        return f();
    } catch (const CudaError& e) {
        g_err = std::string(e.what) + ": " + cudaGetErrorString(e.e);
        cudaGetLastError();
        return e.e == cudaErrorMemoryAllocation ? RTB_ERR_NOMEM : RTB_ERR_CUDA;
    } catch (const std::exception& e) {
        g_err = e.what();
        return RTB_ERR_CUDA;
    }
}

std::vector<int> stripRows(int height, int stripRowsN, int rank, int world)
{
    std::vector<int> rows;
    for (int y = 0; y < height; ++y)
        if ((y / stripRowsN) % world == rank) rows.push_back(y);
    return rows;
}

void destroyHandle(RtbHandle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (TextureRes& t : h->textures) {
        if (t.tex) cudaDestroyTextureObject(t.tex);
        if (t.array) cudaFreeArray(t.array);
    }
    for (void* p : h->allocations) cudaFree(p);
    h->rays[0].release(); h->rays[1].release();
    for (DevBuf* b : { &h->hitTuv, &h->hitObj, &h->surfP, &h->surfN, &h->surfC, &h->vis, &h->interiors, &h->slots, &h->flagged,
             &h->rowsA, &h->rowsB, &h->userRays, &h->outStage })
        b->release();
    if (h->dCtr) cudaFree(h->dCtr);
    if (h->hCtr) cudaFreeHost(h->hCtr);
    for (cudaEvent_t e : h->ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->eventPool) cudaEventDestroy(e);
    if (h->ownStream) cudaStreamDestroy(h->ownStream);
    delete h;
}

} // namespace

extern "C" {

int rtb_abi_version(void) { return RTB_ABI_VERSION; }

const char* rtb_last_error(void) { return g_err.c_str(); }

int rtb_create(const RtbScene* s, int device, uint32_t createFlags, RtbHandle** out)
{
    if (!s || !out) { g_err = "null argument"; return RTB_ERR_ARG; }
    *out = nullptr;
    if (s->abiVersion != RTB_ABI_VERSION) { g_err = "RtbScene.abiVersion mismatch"; return RTB_ERR_ARG; }
    if (s->width < 2 || s->height < 2) { g_err = "image must be at least 2x2"; return RTB_ERR_ARG; }
    RtbHandle* h = new RtbHandle();
    const int rc = guarded([&]() {
        int nDev = 0;
        CK(cudaGetDeviceCount(&nDev));
        if (device < 0 || device >= nDev) throw CudaError{ cudaErrorInvalidDevice, "rtb_create: no such CUDA device" };
        CK(cudaSetDevice(device));
        h->device = device;
        h->createFlags = createFlags;
        cudaDeviceProp prop{};
        CK(cudaGetDeviceProperties(&prop, device));
        h->smCount = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&h->ownStream, cudaStreamNonBlocking));
        for (cudaEvent_t& e : h->ev) CK(cudaEventCreate(&e));
        CK(cudaMalloc((void**)&h->dCtr, sizeof(rtk::Counters)));
        CK(cudaMallocHost((void**)&h->hCtr, sizeof(rtk::Counters)));

        rtpack::packHeader(*s, h->scene);
        std::vector<rt::Object> objects;
        for (int i = 0; i < s->nObjects; ++i) objects.push_back(rtpack::packObject(s->objects[i]));
        std::vector<rt::Light> lights;
        for (int i = 0; i < s->nLights; ++i) lights.push_back(rtpack::packLight(s->lights[i]));
        std::vector<rt::Mesh> meshes;
        for (int i = 0; i < s->nMeshes; ++i) {
            const RtbMesh& m = s->meshes[i];
            rtpack::PackedMesh pm;
            rtpack::packMesh(m, pm);
            if (pm.maxDepth > rtk::kStackDepth) throw std::runtime_error("mesh tree deeper than the traversal stack (64)");
            rt::Mesh d{};
            d.nodes = upload(h, pm.nodes.data(), pm.nodes.size());
            d.slots = upload(h, pm.slots.data(), pm.slots.size());
            d.nrm = upload(h, m.nrm, (size_t)m.nTris * 9);
            d.uv = upload(h, m.uv, (size_t)m.nTris * 6);
            d.tan = upload(h, m.tan, (size_t)m.nTris * 6);
            d.diffuse = uploadImage(h, m.diffuseMap);
            d.normal = uploadImage(h, m.normalMap);
            d.specular = uploadImage(h, m.specularMap);
            d.nNodes = m.nNodes; d.nSlots = m.nRefs; d.nTris = m.nTris; d.maxDepth = pm.maxDepth;
            if (m.nNodes > 0) buildFastPath(h, m, d);
            meshes.push_back(d);
        }
        h->scene.objects = upload(h, objects.data(), objects.size());
        h->scene.lights = upload(h, lights.data(), lights.size());
        h->scene.meshes = upload(h, meshes.data(), meshes.size());
        h->scene.areaPoints = upload(h, s->areaPoints, (size_t)s->nAreaPoints * 3);
        if (s->flags & RTB_FLAG_USE_SKYBOX)
            for (int k = 0; k < 6; ++k) h->scene.sky[k] = uploadImage(h, s->skybox[k]);
        return RTB_OK;
    });
    if (rc != RTB_OK) { destroyHandle(h); return rc; }
    *out = h;
    return RTB_OK;
}

int rtb_render(RtbHandle* h, int y0, int y1, float* fb, float* pass1, int fbOnDevice, void* stream, RtbStats* stats)
{
    if (!h || !fb) { g_err = "null argument"; return RTB_ERR_ARG; }
    if (y0 < 0 || y1 > h->scene.height || y0 > y1) { g_err = "row range outside the image"; return RTB_ERR_ARG; }
    return guarded([&]() {
        std::vector<int> rows;
        for (int y = y0; y < y1; ++y) rows.push_back(y);
        return renderRows(h, rows, fb, pass1, fbOnDevice, stream, stats);
    });
}

int rtb_strip_rows_owned(int height, int stripRowsN, int rank, int worldSize)
{
    if (height <= 0 || stripRowsN <= 0 || worldSize <= 0 || rank < 0 || rank >= worldSize) return RTB_ERR_ARG;
    return (int)stripRows(height, stripRowsN, rank, worldSize).size();
}

int rtb_render_strips(RtbHandle* h, int stripRowsN, int rank, int worldSize, float* fb, int fbOnDevice, void* stream,
    int* nRowsOut, RtbStats* stats)
{
    if (!h || !fb) { g_err = "null argument"; return RTB_ERR_ARG; }
    if (stripRowsN <= 0 || worldSize <= 0 || rank < 0 || rank >= worldSize) { g_err = "bad strip partition"; return RTB_ERR_ARG; }
    return guarded([&]() {
        const std::vector<int> rows = stripRows(h->scene.height, stripRowsN, rank, worldSize);
        if (nRowsOut) *nRowsOut = (int)rows.size();
        return renderRows(h, rows, fb, nullptr, fbOnDevice, stream, stats);
    });
}

int rtb_trace(RtbHandle* h, const float* rays, int nRays, float* tuv, int32_t* objTri)
{
    if (!h || !rays || !tuv || !objTri || nRays < 0) { g_err = "bad argument"; return RTB_ERR_ARG; }
    if (nRays == 0) return RTB_OK;
    return guarded([&]() {
        cudaStream_t st = h->ownStream;
        beginCall(h);
        reserveLevel(h, st, 0, nRays);
        h->userRays.reserve((size_t)nRays * 6 * sizeof(float), st, false);
        CK(cudaMemcpyAsync(h->userRays.p, rays, (size_t)nRays * 6 * sizeof(float), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(h->dCtr, 0, sizeof(rtk::Counters), st));
        rtk::k_rays_from_user<<<gridFor(h, nRays), rtk::kBlock, 0, st>>>(h->userRays.as<float>(), nRays, 0, h->rays[0].view());
        launchCheck();
        const rtk::HitQueue hits{ h->hitTuv.as<float4>(), h->hitObj.as<int>() };
        if (h->createFlags & (RTB_CREATE_EXACT_WALK | RTB_CREATE_COUNTERS))
            rtk::k_trace<rtk::MODE_EXACT><<<gridFor(h, nRays), rtk::kBlock, 0, st>>>(h->scene, h->rays[0].view(), nRays, hits, h->dCtr);
        else
            rtk::k_walk<false><<<persistentGrid(h, nRays), rtk::kBlock, 0, st>>>(h->scene, h->rays[0].view(), nRays, hits, rtk::SurfQueue{}, nullptr, h->dCtr, &h->dCtr->walkCursor[0]);
        launchCheck();
        std::vector<float4> t4(nRays);
        std::vector<int> ob(nRays);
        CK(cudaMemcpyAsync(t4.data(), hits.tuv, (size_t)nRays * sizeof(float4), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ob.data(), hits.obj, (size_t)nRays * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int i = 0; i < nRays; ++i) {
            tuv[3 * i] = t4[i].x; tuv[3 * i + 1] = t4[i].y; tuv[3 * i + 2] = t4[i].z;
            int tri;
            std::memcpy(&tri, &t4[i].w, 4);
            objTri[2 * i] = ob[i];
            objTri[2 * i + 1] = ob[i] < 0 ? -1 : tri;
        }
        return RTB_OK;
    });
}

int rtb_cast(RtbHandle* h, const float* rays, int nRays, float* rgb)
{
    if (!h || !rays || !rgb || nRays < 0) { g_err = "bad argument"; return RTB_ERR_ARG; }
    if (nRays == 0) return RTB_OK;
    return guarded([&]() {
        cudaStream_t st = h->ownStream;
        beginCall(h);
        h->slotCount = nRays;
        reserveLevel(h, st, 0, nRays);
        h->userRays.reserve((size_t)nRays * 6 * sizeof(float), st, false);
        CK(cudaMemcpyAsync(h->userRays.p, rays, (size_t)nRays * 6 * sizeof(float), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(h->dCtr, 0, sizeof(rtk::Counters), st));
        rtk::k_rays_from_user<<<gridFor(h, nRays), rtk::kBlock, 0, st>>>(h->userRays.as<float>(), nRays, 0, h->rays[0].view());
        launchCheck();
        runLevels(h, st, nRays, h->stats.primaryRays);
        CK(cudaMemcpyAsync(rgb, h->slots.p, (size_t)nRays * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return RTB_OK;
    });
}

int rtb_device_of(const RtbHandle* h) { return h ? h->device : RTB_ERR_ARG; }

void rtb_destroy(RtbHandle* h) { destroyHandle(h); }

} // extern "C"
