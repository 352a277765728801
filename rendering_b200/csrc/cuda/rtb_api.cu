// rtb_api.cu — extern "C" entry points of librtb_cuda.so (include/rtb.h, "device side") and the
// host-side frame scheduler that strings the wavefront kernels of rtb_kernels.cuh together.
//
// One RtbHandle owns: the scene resident in HBM (reference-tree nodes, triangle slots in leaf
// order, shading attributes, RGBA8 textures as cudaTextureObjects), the per-level queues, the
// colour-slot array whose first w*h entries are the framebuffer, and one CUDA stream.
// There is no CPU rendering path anywhere in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cfloat>
#include <cmath>
#include <chrono>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/rtb.h"
#include "rtb_kernels.cuh"
#include "rtb_tile.cuh"
#include "scene_pack.h"
#include "lbvh_build.cuh"

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

constexpr long long kEarlyOutRatio = 16;   // early output pays while the re-traced pixels are fewer than 1/16 of the frame's
enum OutputKind { OUT_FLOAT = 0, OUT_BGR8 = 1, OUT_SCATTER = 2 };   // OUT_SCATTER: rows go to their image position of a full-frame device buffer

struct RtbHandle {
    // the frame between rtb_render*_begin and rtb_render_end (beginRows / endRows)
    struct FramePlan {
        bool active = false;
        cudaStream_t st = nullptr;
        std::vector<int> owned;
        void* fb = nullptr; float* pass1 = nullptr; int fbOnDevice = 0; OutputKind kind = OUT_FLOAT;
        bool pipelinedCopy = false;      // OUT_BGR8 to a host buffer through copyStream (rtb_render_bgr8_begin)
        // early output: the pass-1 bytes leave for the host right after pass 1 (copy engine, beside Sobel and the SSAA pass), then
        // only the re-traced pixels are rewritten in place through fbDev, the device-side address of the (pinned) host buffer
        bool earlyOut = false;
        void* fbDev = nullptr;
        int earlyStage = 0;
        bool ssaa = false, literalWalk = false, culled = false;
        int genX0 = 0, genCols = 0, nGenRows = 0, nInitRows = 0;
        bool cover = false;                      // rays only for the 8x4 tiles of the resident kept list (k_tile_lists), the rest is filled
        bool rebuildLists = false;               // ... and this frame (re)builds the lists: camera or rows changed
        bool coverRead = false;                  // ... and reads their counters back (endRows)
        int tilesTotal = 0;                      // tiles of the generation grid
        long long livePixels = 0;                // pixels that get a generated primary ray
        long long nPixels = 0, n0 = 0, interiorPixels = 0, flaggedCap = 0;
        size_t outBytes = 0;
    } plan;
    int device = 0;
    uint32_t createFlags = 0;
    cudaStream_t ownStream = nullptr;
    int smCount = 148;

    rt::Scene scene{};                 // header with DEVICE pointers
    rt::Scene* sceneDev = nullptr;     // the same header resident in HBM
    rt::Scene* scenePinned = nullptr;  // pinned staging copy: rtb_set_camera writes it, the next render call uploads it on ITS stream
    bool sceneDirty = false;
    std::vector<void*> allocations;    // scene-lifetime device allocations
    std::vector<TextureRes> textures;
    bool kernelTiming = false;         // bracket every launch with CUDA events (RTB_CREATE_KERNEL_TIMING -> RtbStats.msKernel)
    std::vector<int> refTreeDepth;     // depth of every mesh's reference tree (showAC walk)
    // world-space boxes (lo.xyz, hi.xyz) around everything a primary ray can hit, for the screen-space bounds of the
    // geometry; `unbounded` when a plane is present or misses need their direction (skybox)
    std::vector<std::array<float, 6>> geomBounds;
    // finer boxes around the same geometry (search-BVH boxes a few levels down), resident on the device: every frame whose
    // camera or rows changed projects them into a coverage bitmap and turns that into the kept / skipped tile lists
    // (k_cover_mark, k_tile_lists) without the host touching a pixel; empty = no finer bound than primRect
    std::vector<std::array<float, 6>> coverBounds;
    const float* coverBoxesDev = nullptr;
    int nCoverBoxes = 0;
    DevBuf coverBits;
    rtk::CoverCtr* coverCtrDev = nullptr;
    rtk::CoverCtr* hCover = nullptr;   // pinned mirror, read back by the frames that rebuild the lists
    uint64_t cameraVersion = 1;        // bumped whenever the camera (hence primRect / the coverage) changes
    uint64_t tileListCamera = 0;       // cameraVersion the resident tile lists (tilesKept / tilesSkipped) were built for
    std::vector<int> tileListRows;     // ... and the generation rows
    long long tileListLive = -1;       // live pixels / length of the kept list as last read back (-1: not known yet)
    int tileListKept = -1;
    bool unbounded = false;
    int primRect[4] = { 0, 0, 0, 0 };  // pixel columns [x0,x1) and rows [y0,y1) primary rays are generated for
    uint64_t pendingH2D = 0;           // bytes uploaded by rtb_set_camera since the last render call (reported in its stats)
    float msBuildSearchBvh = 0.0f;     // time spent building the search BVHs at create
    int levels = 1;                    // recursion levels a ray tree can have: maxRayDepth+1 if any object spawns children
    bool spawns = false;               // some object is Reflective / Transparent (their shade branches exist even when maxRayDepth == 0)
    int stackEntries = 1;              // per-thread traversal stack entries the kernels need for this scene
    int walkBlocksPerSm[4] = { 1, 1, 1, 1 };   // resident CTAs per SM of k_walk<false, GEN 0..2> / k_walk<true> with that stack

    // tile pipeline (rtb_tile.cuh): per-group scratch slabs and their capacities (grown after an overflow, never shrunk)
    bool tilePipeline = true;
    DevBuf tileSlab;
    int tileCapRays = 0, tileCapInterior = 0;
    int tileBlocksPerSm[3] = { 0, 0, 0 };   // resident CTAs per SM of k_tile<GEN 0..2> for this scene (0 = not queried yet)
    // staged form (one 1024-thread CTA per SM, four groups, the top of one mesh's search BVH in shared memory)
    bool tileStaged = false;
    int stagedMeshIndex = -1;
    const float4* stagedNodesSrc = nullptr;
    const float4* stagedTrisSrc = nullptr;
    int stagedNodesAll = 0, stagedTrisAll = 0;   // that mesh's node / triangle counts
    int stagedNodes = 0, stagedTris = 0;         // what fits next to the stacks
    size_t smemOptin = 0;
    bool stagedAttrSet[3] = { false, false, false };

    QueueBufs rays[2];
    DevBuf hitTuv, hitObj, surfP, surfN, surfC, vis, interiors, slots, flagged, rowsA, rowsB, rowsC, userRays, outStage, tilesKept, tilesSkipped;
    DevBuf ctrBuf;                     // FrameCtr followed by 2 passes x (levels + 1) LevelCtr
    void* hCtr = nullptr;              // pinned mirror of ctrBuf
    void* hCtrDev = nullptr;           // ... and its device-side address: the counters come back by plain stores (k_store_words)
    size_t ctrBytes = 0;
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    // pipelined output (rtb_render_bgr8_begin): the frame's bytes leave on a second stream from one of two staging buffers, so
    // the device-to-host copy of frame i overlaps the kernels of frame i+1
    cudaStream_t copyStream = nullptr;
    DevBuf pipeStage[2];
    cudaEvent_t pipeReady[2] = { nullptr, nullptr }, pipeCopied[2] = { nullptr, nullptr };
    int pipeParity = 0;

    // capacities (in items) the buffers above currently provide; grown on demand, never shrunk
    long long capLevel0 = 0;           // rays of a level-0 queue (ray generation)
    long long capDeep = 0;             // rays of a deeper level's queue
    long long capInterior = 0;         // interior records of one frame
    long long capFlagged = 0;          // SSAA pixels
    long long capSlots = 0;
    // SSAA capacity hint carried from frame to frame: flagged pixels seen last time
    long long flaggedSeen = 0;
    std::vector<int> rowsAHost, rowsBHost, rowsCHost;   // lists currently resident in the DevBufs of the same name

    // per-kernel timing: (kind, start, stop) spans recorded on the render stream, resolved at the end of a call
    struct Span { int kind; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> eventPool;
    size_t eventsUsed = 0;

    RtbStats stats{};

    rtk::FrameCtr* dFrame() const { return ctrBuf.as<rtk::FrameCtr>(); }
    rtk::LevelCtr* dLevel(int pass, int level) const
    {
        return reinterpret_cast<rtk::LevelCtr*>(ctrBuf.as<char>() + sizeof(rtk::FrameCtr)) + (size_t)pass * (levels + 1) + level;
    }
    const rtk::FrameCtr* hFrame() const { return static_cast<const rtk::FrameCtr*>(hCtr); }
    const rtk::LevelCtr* hLevel(int pass, int level) const
    {
        return reinterpret_cast<const rtk::LevelCtr*>(static_cast<const char*>(hCtr) + sizeof(rtk::FrameCtr)) + (size_t)pass * (levels + 1) + level;
    }
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
    if (!im.rgb || im.width <= 0 || im.height <= 0) return out;
    const size_t nTexels = (size_t)im.width * im.height;
    // 3 B / texel over PCIe, widened to RGBA8 on the device, then laid out as a CUDA array (2D-local texture fetches).
    // Staging buffers are released on every path (a failing call unwinds through here).
    struct Staging {
        unsigned char* rgb = nullptr; uchar4* rgba = nullptr;
        ~Staging() { if (rgb) cudaFree(rgb); if (rgba) cudaFree(rgba); }
    } stg;
    cudaStream_t st = h->ownStream;
    CK(cudaMalloc((void**)&stg.rgb, nTexels * 3));
    CK(cudaMalloc((void**)&stg.rgba, nTexels * sizeof(uchar4)));
    CK(cudaMemcpyAsync(stg.rgb, im.rgb, nTexels * 3, cudaMemcpyHostToDevice, st));
    const int blocks = (int)std::min<size_t>((nTexels + 255) / 256, (size_t)h->smCount * 16);
    rtk::k_rgb_to_rgba<<<blocks, 256, 0, st>>>(stg.rgb, nTexels, stg.rgba);
    CK(cudaGetLastError());
    TextureRes res;
    const cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    CK(cudaMallocArray(&res.array, &fmt, im.width, im.height));
    h->textures.push_back(res);       // owned by the handle from here on (destroyHandle frees it even if a later step throws)
    CK(cudaMemcpy2DToArrayAsync(res.array, 0, 0, stg.rgba, (size_t)im.width * 4, (size_t)im.width * 4, im.height, cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = res.array;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;          // the reference samples nearest texel, no filtering
    td.readMode = cudaReadModeElementType;        // raw bytes; /256 happens in fetchTexel
    td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&res.tex, &rd, &td, nullptr));
    h->textures.back().tex = res.tex;
    out.tex = (unsigned long long)res.tex;
    out.rgba = nullptr;
    out.w = im.width;
    out.h = im.height;
    return out;
}

// Fast-path data of one mesh: search BVH over its unique triangles (bvh_build.h) and the tables that
// let the kernel evaluate the reference tree's eligibility rule for a single triangle.
int buildFastPath(RtbHandle* h, const RtbMesh& m, rt::Mesh& d, int meshIndex)
{
    const bool onDevice = h->createFlags & RTB_CREATE_DEVICE_BVH;
    rtpack::FastPath fp;
    int maxDepth = 0, nNodes = 0, nTris = 0;
    if (onDevice) {
        // eligibility tables from the host (they restate the reference tree); the search BVH itself on the device
        rtpack::packFastPath(m, fp, false);
        float* dPos = nullptr;
        CK(cudaMalloc((void**)&dPos, (size_t)m.nTris * 9 * sizeof(float)));
        lbvh::DeviceBvh bvh;
        cudaError_t e = cudaMemcpyAsync(dPos, m.pos, (size_t)m.nTris * 9 * sizeof(float), cudaMemcpyHostToDevice, h->ownStream);
        if (e == cudaSuccess) e = lbvh::buildOnDevice(dPos, m.nTris, fp.lo, fp.hi, fp.pad, h->ownStream, bvh);
        cudaFree(dPos);
        if (e != cudaSuccess) throw CudaError{ e, "device search-BVH build" };
        h->allocations.push_back(bvh.nodes);
        h->allocations.push_back(bvh.tris);
        d.bvhNodes = reinterpret_cast<const float4*>(bvh.nodes);
        d.bvhTris = bvh.tris;
        maxDepth = bvh.maxDepth; nNodes = bvh.nNodes; nTris = m.nTris;
        h->msBuildSearchBvh += bvh.buildMs;
        std::array<float, 6> b;
        for (int a = 0; a < 3; ++a) { b[a] = fp.lo[a] - fp.pad; b[3 + a] = fp.hi[a] + fp.pad; }
        // a NaN / inf vertex makes its triangle's box unbounded on the device too: no screen-space bound then
        bool finite = true;
        for (size_t i = 0; i < (size_t)m.nTris * 9; ++i) finite &= (m.pos[i] >= -FLT_MAX && m.pos[i] <= FLT_MAX);
        if (finite) { h->geomBounds.push_back(b); h->coverBounds.push_back(b); } else h->unbounded = true;
    } else {
        const auto t0 = std::chrono::steady_clock::now();
        rtpack::packFastPath(m, fp);
        h->msBuildSearchBvh += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        d.bvhNodes = reinterpret_cast<const float4*>(upload(h, fp.nodes.data(), fp.nodes.size()));
        d.bvhTris = upload(h, fp.tris.data(), fp.tris.size());
        maxDepth = fp.maxDepth; nNodes = (int)fp.nodes.size(); nTris = (int)(fp.tris.size() / 3);
        std::array<float, 6> b;
        if (rtpack::meshBounds(fp, b)) h->geomBounds.push_back(b);
        rtpack::meshCoverBoxes(fp, 10, h->coverBounds);
    }
    if (maxDepth > rtk::kStackDepth) throw std::runtime_error("search BVH deeper than the traversal stack (64)");
    if (nNodes > h->stagedNodesAll) {   // the largest search BVH is the one worth keeping in shared memory / prefetching
        h->stagedMeshIndex = meshIndex;
        h->stagedNodesAll = nNodes;
        h->stagedTrisAll = nTris;
    }
    d.triRefOff = upload(h, fp.triRefOff.data(), fp.triRefOff.size());
    d.triRefs = upload(h, fp.triRefs.data(), fp.triRefs.size());
    d.parent = upload(h, fp.parent.data(), fp.parent.size());
    return maxDepth;
}

void computePrimaryRect(RtbHandle* h)
{
    // the rectangle comes from the few root boxes (host, microseconds); the fine coverage inside it is the device's job
    rtpack::primaryRect(h->scene, h->geomBounds, h->unbounded, h->primRect);
    h->cameraVersion++;
}

size_t stackBytes(const RtbHandle* h) { return (size_t)h->stackEntries * rtk::kBlock * sizeof(int); }

// grid-stride kernels: enough CTAs to fill the machine, never more than the work needs
int gridFor(const RtbHandle* h, long long n, int block = rtk::kBlock, int perSm = 16)
{
    const long long blocks = (n + block - 1) / block;
    const long long cap = (long long)h->smCount * perSm;
    return (int)std::max(1LL, std::min(blocks, cap));
}

// persistent kernels: exactly one resident wave (occupancy queried at create time), never more CTAs than
// the queue's capacity could feed
int persistentGrid(const RtbHandle* h, int variant /* GEN_* or 3 = shadow */, long long cap)
{
    const long long blocks = (cap + rtk::kBlock - 1) / rtk::kBlock;
    return (int)std::max(1LL, std::min(blocks, (long long)h->smCount * h->walkBlocksPerSm[variant]));
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
        if (!h->kernelTiming) return;
        a = takeEvent(h); b = takeEvent(h);
        CK(cudaEventRecord(a, st));
    }
    void done(int launches = 1)     // launches: kernels enqueued inside the span (every one is counted)
    {
        launchCheck();
        h->stats.kernelLaunches += launches;
        h->stats.launchesKernel[kind] += launches;
        if (!h->kernelTiming) return;
        CK(cudaEventRecord(b, st));
        h->spans.push_back({ kind, a, b });
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

// Sizes every per-frame buffer for: level-0 queues of up to n0 rays, deeper levels of up to nDeep rays,
// nInterior interior records, nFlagged SSAA pixels on top of `framePixels` framebuffer slots.
// Slot layout: [0, framePixels) framebuffer | [framePixels, +4*capFlagged) SSAA samples | 2 per interior record.
void ensureCapacity(RtbHandle* h, cudaStream_t st, long long framePixels, long long n0, long long nDeep, long long nInterior, long long nFlagged,
    bool frameWideQueues = true)
{
    if (!frameWideQueues) {
        // tile pipeline: queues live in the groups' scratch slabs (enqueueTile); only the framebuffer and the SSAA list are frame-sized
        h->capFlagged = std::max(h->capFlagged, nFlagged);
        h->flagged.reserve((size_t)std::max(1LL, h->capFlagged) * sizeof(int), st, false);
        h->capSlots = std::max(h->capSlots, framePixels);
        h->slots.reserve((size_t)h->capSlots * 3 * sizeof(float), st, false);
        return;
    }
    if (std::max(n0, nDeep) > (1LL << 28)) throw CudaError{ cudaErrorMemoryAllocation, "ray queue larger than 2^28 rays" };
    h->capLevel0 = std::max(h->capLevel0, n0);
    h->capDeep = std::max(h->capDeep, h->levels > 1 ? nDeep : 0LL);
    h->capInterior = std::max(h->capInterior, h->levels > 1 ? nInterior : 0LL);
    h->capFlagged = std::max(h->capFlagged, nFlagged);
    const long long S = std::max(1, h->scene.shadowRaysPerHit);
    const long long nMax = std::max(h->capLevel0, h->capDeep);
    // level d lives in rays[d & 1]: queue 0 holds level 0 and the even deeper levels, queue 1 the odd ones
    h->rays[0].reserve((size_t)std::max(1LL, nMax), st, false);
    h->rays[1].reserve((size_t)std::max(1LL, h->capDeep), st, false);
    h->hitTuv.reserve((size_t)nMax * sizeof(float4), st, false);
    h->hitObj.reserve((size_t)nMax * sizeof(int), st, false);
    h->surfP.reserve((size_t)nMax * sizeof(float4), st, false);
    h->surfN.reserve((size_t)nMax * sizeof(float4), st, false);
    h->surfC.reserve((size_t)nMax * sizeof(float4), st, false);
    h->vis.reserve((size_t)(nMax * S), st, false);
    h->interiors.reserve((size_t)std::max(1LL, h->capInterior) * sizeof(rtk::Interior), st, false);
    h->flagged.reserve((size_t)std::max(1LL, h->capFlagged) * sizeof(int), st, false);
    h->capSlots = std::max(h->capSlots, framePixels + 4 * h->capFlagged + 2 * h->capInterior);
    h->slots.reserve((size_t)h->capSlots * 3 * sizeof(float), st, false);
}

// ---- tile pipeline ----------------------------------------------------------------------------------------------------
constexpr int kStagedGroups = 4;   // groups of the staged kernel's single CTA per SM
template <int GEN, bool DEEP, bool STATS>
void launchTile(int grid, size_t smem, cudaStream_t st, const rt::Scene& sc, const rtk::TileArgs& a)
{
    rtk::k_tile<GEN, DEEP, STATS, 1><<<grid, rtk::kTileThreads, smem, st>>>(sc, a);
}
template <int GEN>
void launchTileStaged(bool deep, bool& attrSet, size_t smemOptin, int grid, size_t smem, cudaStream_t st, const rt::Scene& sc, const rtk::TileArgs& a)
{
    if (!attrSet) {   // the attribute is per function, not per handle: always ask for the device's opt-in maximum
        if (deep) CK(cudaFuncSetAttribute(rtk::k_tile<GEN, true, false, kStagedGroups, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemOptin));
        else CK(cudaFuncSetAttribute(rtk::k_tile<GEN, false, false, kStagedGroups, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemOptin));
        attrSet = true;
    }
    if (deep) rtk::k_tile<GEN, true, false, kStagedGroups, true><<<grid, rtk::kTileThreads * kStagedGroups, smem, st>>>(sc, a);
    else rtk::k_tile<GEN, false, false, kStagedGroups, true><<<grid, rtk::kTileThreads * kStagedGroups, smem, st>>>(sc, a);
}
template <int GEN>
void launchTileGen(bool deep, bool stats, int grid, size_t smem, cudaStream_t st, const rt::Scene& sc, const rtk::TileArgs& a)
{
    if (deep) { if (stats) launchTile<GEN, true, true>(grid, smem, st, sc, a); else launchTile<GEN, true, false>(grid, smem, st, sc, a); }
    else { if (stats) launchTile<GEN, false, true>(grid, smem, st, sc, a); else launchTile<GEN, false, false>(grid, smem, st, sc, a); }
}
template <int GEN>
int tileOccupancy(bool deep, size_t smem)
{
    int b = 0;
    if (deep) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, rtk::k_tile<GEN, true, false, 1>, rtk::kTileThreads, smem));
    else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, rtk::k_tile<GEN, false, false, 1>, rtk::kTileThreads, smem));
    return std::max(1, b);
}

size_t tileSmemBytes(const RtbHandle* h)
{
    return ((sizeof(rtk::TileShared) + 15) & ~(size_t)15) + (size_t)h->stackEntries * rtk::kTileThreads * sizeof(int);
}

// castRay for `total` level-0 rays (an upper bound when only the device knows the count), tile by tile, in one launch.
void enqueueTile(RtbHandle* h, cudaStream_t st, int pass, int genKind, const rtk::GenArgs& gen, long long total, rtk::RayQueue userQ = rtk::RayQueue{},
    int nUser = 0)
{
    const bool deep = h->spawns;
    const bool stats = h->createFlags & RTB_CREATE_WALK_STATS;
    const bool staged = h->tileStaged && !stats && h->stagedNodes > 0 && !(h->createFlags & RTB_CREATE_DEVICE_BVH);
    const size_t smem = staged ? rtk::tileSmemStagedOffset(kStagedGroups, h->stackEntries) + (size_t)h->stagedNodes * 64 + (size_t)h->stagedTris * 48
                               : tileSmemBytes(h);
    if (!staged && h->tileBlocksPerSm[genKind] == 0)
        h->tileBlocksPerSm[genKind] = genKind == rtk::GEN_PRIMARY ? tileOccupancy<rtk::GEN_PRIMARY>(deep, smem)
            : genKind == rtk::GEN_SSAA ? tileOccupancy<rtk::GEN_SSAA>(deep, smem) : tileOccupancy<rtk::GEN_QUEUE>(deep, smem);
    const long long groups = staged ? (long long)h->smCount * kStagedGroups : (long long)h->smCount * h->tileBlocksPerSm[genKind];
    // tile size: ONE 32-ray batch per warp of the group.  Measured on B200 (tools/gpu_sweep.sh, frame ms cfg1 / cfg3 / cfg4 /
    // dragon): 256 rays 0.49 / 0.92 / 0.46 / 1.39, 512 rays 0.55 / 0.99 / 0.51 / 1.44, 1024 rays 0.67 / 1.15 / 0.62 / 1.64 —
    // small tiles keep the groups of an SM out of phase and the dynamic tile cursor balances the SMs.
    // Growing the tile when a rank of a multi-GPU frame has only 1-2 tiles per group (to save a "round") was measured too:
    // 1/4 of cfg4's rows 0.277 -> 0.317 ms, 1/8 0.267 -> 0.285 ms (tools/gpu_strips.py): kept at one batch per warp.
    // A pass with fewer 256-ray tiles than HALF the resident groups (small images, short SSAA lists) gets 64-ray tiles:
    // four times the groups share the work, and above all the few heavy tiles of a deep scene (the pixels of a glass sphere
    // spawn two children per level) are cut in four.  Measured (256 / 128 / 64 / 32 rays, pass 1 + SSAA ms): cfg1 256x256
    // depth 10 0.295+0.261 / 0.205+0.180 / 0.180+0.169 / 0.179+0.179; cfg2's SSAA list (142 tiles) 0.097 / 0.093 / 0.091 /
    // 0.120.  With enough tiles to fill the machine small tiles only idle lanes: cfg4 0.408 / 0.473 / 0.644 / 1.002 ms,
    // cfg3's SSAA list (344 tiles) 0.136 / 0.137 / 0.189 / 0.259.
    static const int forced = getenv("RTB_TILE_RAYS") ? atoi(getenv("RTB_TILE_RAYS")) : 0;
    const long long tiles256 = (total + rtk::kTileThreads - 1) / rtk::kTileThreads;
    // Where the tiles are uniform (no secondary rays) the cut only pays for really short passes: a rank's eighth of cfg4 (262
    // tiles) is 10 % slower with 64-ray tiles on 8 GPUs (0.297 vs 0.267 ms per frame), hence the second condition.
    const bool small = deep ? 2 * tiles256 < groups : 4 * tiles256 < groups;
    long long R = forced > 0 ? forced : (small ? 64 : rtk::kTileThreads);
    R = std::max<long long>(32, std::min(1024LL, R)) & ~31LL;
    const int S = h->scene.shadowRaysPerHit;
    h->tileCapRays = std::max<long long>(h->tileCapRays, deep ? 2 * R : R);
    h->tileCapInterior = deep ? std::max<long long>(h->tileCapInterior, 4 * R) : 0;
    const unsigned long long slab = rtk::tileSlabBytes(h->tileCapRays, h->tileCapInterior, 1024, S, h->levels);
    h->tileSlab.reserve((size_t)(groups * slab), st, false);
    const long long tiles = (total + R - 1) / R;
    const int grid = staged ? (int)std::max(1LL, std::min<long long>(h->smCount, (tiles + kStagedGroups - 1) / kStagedGroups))
                            : (int)std::max(1LL, std::min(groups, tiles));
    rtk::TileArgs a{};
    a.slab = h->tileSlab.as<char>();
    a.slabBytes = slab;
    a.capRays = h->tileCapRays;
    a.capInterior = h->tileCapInterior;
    a.levels = h->levels;
    a.tileRays = (int)R;
    a.stackEntries = h->stackEntries;
    a.tileCursor = &h->dFrame()->tileCursor[pass];
    a.ctr = h->dFrame();
    a.lv = h->dLevel(pass, 0);
    a.gen = gen;
    a.userQ = userQ;
    a.nUser = nUser;
    a.stagedMesh = staged ? h->scene.meshes + h->stagedMeshIndex : nullptr;
    a.stagedNodesSrc = h->stagedNodesSrc;
    a.stagedTrisSrc = h->stagedTrisSrc;
    a.stagedNodes = h->stagedNodes;
    a.stagedTris = h->stagedTris;
    static const int sortSurfaces = getenv("RTB_TILE_SORT") ? atoi(getenv("RTB_TILE_SORT")) : 0;
    a.sortSurfaces = deep ? sortSurfaces : 0;
    KernelSpan ks(h, st, pass == 0 ? RTB_K_TILE : RTB_K_TILE_SSAA);
    if (staged) {
        if (genKind == rtk::GEN_PRIMARY) launchTileStaged<rtk::GEN_PRIMARY>(deep, h->stagedAttrSet[genKind], h->smemOptin, grid, smem, st, h->scene, a);
        else if (genKind == rtk::GEN_SSAA) launchTileStaged<rtk::GEN_SSAA>(deep, h->stagedAttrSet[genKind], h->smemOptin, grid, smem, st, h->scene, a);
        else launchTileStaged<rtk::GEN_QUEUE>(deep, h->stagedAttrSet[genKind], h->smemOptin, grid, smem, st, h->scene, a);
    }
    else if (genKind == rtk::GEN_PRIMARY) launchTileGen<rtk::GEN_PRIMARY>(deep, stats, grid, smem, st, h->scene, a);
    else if (genKind == rtk::GEN_SSAA) launchTileGen<rtk::GEN_SSAA>(deep, stats, grid, smem, st, h->scene, a);
    else launchTileGen<rtk::GEN_QUEUE>(deep, stats, grid, smem, st, h->scene, a);
    ks.done();
}

// Enqueues castRay for the rays sitting in queue 0 of `pass`, level by level, then folds the interior
// records deepest level first.  Nothing here waits for the device: every kernel reads its work size
// from the LevelCtr its predecessor filled.
void enqueueLevels(RtbHandle* h, cudaStream_t st, int pass, long long framePixels, int genKind = rtk::GEN_QUEUE, rtk::GenArgs gen = rtk::GenArgs{})
{
    const bool count = h->createFlags & RTB_CREATE_COUNTERS;
    const bool exact = h->createFlags & RTB_CREATE_EXACT_WALK;
    const bool walkStats = h->createFlags & RTB_CREATE_WALK_STATS;
    const rt::Scene& sc = h->scene;
    const size_t smem = stackBytes(h);
    const rtk::HitQueue hits{ h->hitTuv.as<float4>(), h->hitObj.as<int>() };
    const rtk::SurfQueue surf{ h->surfP.as<float4>(), h->surfN.as<float4>(), h->surfC.as<float4>() };
    unsigned char* vis = h->vis.as<unsigned char>();
    const int slotBase = (int)(framePixels + 4 * h->capFlagged);
    for (int depth = 0; depth < h->levels; ++depth) {
        const long long cap = depth == 0 ? h->capLevel0 : h->capDeep;
        const long long capNext = h->capDeep;
        const rtk::RayQueue q = h->rays[depth & 1].view(), next = h->rays[(depth + 1) & 1].view();
        rtk::LevelCtr* lv = h->dLevel(pass, depth);
        {
            KernelSpan ks(h, st, RTB_K_TRACE);
            if (count) rtk::k_trace<rtk::MODE_COUNT><<<gridFor(h, cap), rtk::kBlock, smem, st>>>(sc, q, (int)cap, hits, h->dFrame(), lv);
            else if (exact) rtk::k_trace<rtk::MODE_EXACT><<<gridFor(h, cap), rtk::kBlock, smem, st>>>(sc, q, (int)cap, hits, h->dFrame(), lv);
            else if (walkStats) {   // same traversal with its own work counters (RTB_CREATE_WALK_STATS)
                if (depth == 0 && genKind == rtk::GEN_PRIMARY)
                    rtk::k_walk<false, rtk::GEN_PRIMARY, true><<<persistentGrid(h, 1, cap), rtk::kBlock, smem, st>>>(sc, q, (int)cap, hits, surf, vis, h->dFrame(), lv, gen);
                else if (depth == 0 && genKind == rtk::GEN_SSAA)
                    rtk::k_walk<false, rtk::GEN_SSAA, true><<<persistentGrid(h, 2, cap), rtk::kBlock, smem, st>>>(sc, q, (int)cap, hits, surf, vis, h->dFrame(), lv, gen);
                else
                    rtk::k_walk<false, rtk::GEN_QUEUE, true><<<persistentGrid(h, 0, cap), rtk::kBlock, smem, st>>>(sc, q, (int)cap, hits, surf, vis, h->dFrame(), lv, gen);
            }
            else if (depth == 0 && genKind == rtk::GEN_PRIMARY)
                rtk::k_walk<false, rtk::GEN_PRIMARY><<<persistentGrid(h, 1, cap), rtk::kBlock, smem, st>>>(sc, q, (int)cap, hits, surf, vis, h->dFrame(), lv, gen);
            else if (depth == 0 && genKind == rtk::GEN_SSAA)
                rtk::k_walk<false, rtk::GEN_SSAA><<<persistentGrid(h, 2, cap), rtk::kBlock, smem, st>>>(sc, q, (int)cap, hits, surf, vis, h->dFrame(), lv, gen);
            else
                rtk::k_walk<false, rtk::GEN_QUEUE><<<persistentGrid(h, 0, cap), rtk::kBlock, smem, st>>>(sc, q, (int)cap, hits, surf, vis, h->dFrame(), lv, gen);
            ks.done();
        }
        {
            KernelSpan ks(h, st, RTB_K_SURFACE);
            const int handleMisses = count || exact || depth > 0 || genKind == rtk::GEN_QUEUE;
            rtk::k_surface<<<gridFor(h, cap), rtk::kBlock, 0, st>>>(sc, q, (int)cap, hits, surf, h->slots.as<float>(), lv, handleMisses);
            ks.done();
        }
        if (!(sc.flags & rt::FLAG_SHOW_NORMALS)) {
            if (sc.shadowRaysPerHit > 0) {
                const long long maxShadow = cap * sc.shadowRaysPerHit;
                KernelSpan ks(h, st, RTB_K_SHADOW);
                if (count) rtk::k_shadow<rtk::MODE_COUNT><<<gridFor(h, maxShadow), rtk::kBlock, smem, st>>>(sc, q, surf, vis, h->dFrame(), lv);
                else if (exact) rtk::k_shadow<rtk::MODE_EXACT><<<gridFor(h, maxShadow), rtk::kBlock, smem, st>>>(sc, q, surf, vis, h->dFrame(), lv);
                else if (walkStats) rtk::k_walk<true, rtk::GEN_QUEUE, true><<<persistentGrid(h, 3, maxShadow), rtk::kBlock, smem, st>>>(sc, q, (int)cap, hits, surf, vis, h->dFrame(), lv, gen);
                else rtk::k_walk<true, rtk::GEN_QUEUE><<<persistentGrid(h, 3, maxShadow), rtk::kBlock, smem, st>>>(sc, q, (int)cap, hits, surf, vis, h->dFrame(), lv, gen);
                ks.done();
            }
            KernelSpan ks(h, st, RTB_K_SHADE);
            rtk::k_shade<<<gridFor(h, cap), rtk::kBlock, 0, st>>>(sc, q, surf, vis, depth, next, (int)capNext,
                h->interiors.as<rtk::Interior>(), (int)h->capInterior, h->slots.as<float>(), slotBase, (int)h->capSlots, h->dFrame(), lv);
            ks.done();
        }
    }
    for (int l = h->levels - 2; l >= 0; --l) {   // the deepest level never has children
        KernelSpan ks(h, st, RTB_K_COMBINE);
        // the LevelCtr entries of both passes are one flat array: records of earlier passes / shallower levels come first
        rtk::k_combine<<<gridFor(h, h->capInterior), rtk::kBlock, 0, st>>>(h->interiors.as<rtk::Interior>(), h->dLevel(0, 0),
            pass * (h->levels + 1) + l, (int)h->capInterior, h->slots.as<float>());
        ks.done();
    }
}

// A camera set since the last call reaches the device copy of the header here, stream-ordered before the kernels of
// this call and without a synchronisation (the call itself ends with one, so the staging copy is free again on return).
void uploadSceneHeader(RtbHandle* h, cudaStream_t st)
{
    if (!h->sceneDirty) return;
    CK(cudaMemcpyAsync(h->sceneDev, h->scenePinned, sizeof(rt::Scene), cudaMemcpyHostToDevice, st));
    h->sceneDirty = false;
}

void beginCall(RtbHandle* h)
{
    CK(cudaSetDevice(h->device));
    h->stats = RtbStats{};
    h->stats.msBuildSearchBvh = h->msBuildSearchBvh;
    h->stats.h2dBytes = h->pendingH2D;
    h->pendingH2D = 0;
    h->spans.clear();
    h->eventsUsed = 0;
}

float elapsed(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

// Reads the frame's counters back (the one synchronisation of a frame) and folds them into the stats.
// Returns the overflow bits.
int finishFrame(RtbHandle* h, cudaStream_t st, int passes, bool enqueueReadback = false)
{
    if (enqueueReadback) {
        CK(cudaMemcpyAsync(h->hCtr, h->ctrBuf.p, h->ctrBytes, cudaMemcpyDeviceToHost, st));
        h->stats.d2hBytes += h->ctrBytes;
    }
    CK(cudaStreamSynchronize(st));
    const rtk::FrameCtr* fc = h->hFrame();
    const uint64_t S = (uint64_t)h->scene.shadowRaysPerHit;
    for (int p = 0; p < passes; ++p)
        for (int l = 0; l < h->levels; ++l) {
            const rtk::LevelCtr* lv = h->hLevel(p, l);
            if (l > 0) h->stats.secondaryRays += (uint64_t)lv->nRays;
            if (!(h->scene.flags & rt::FLAG_SHOW_NORMALS)) h->stats.shadowRays += (uint64_t)lv->nSurf * S;
            if (lv->nRays > 0) h->stats.levels = std::max<uint32_t>(h->stats.levels, (uint32_t)l + 1);
        }
    h->stats.shadowRaysSkipped = fc->shadowSkipped;
    for (int k = 0; k < 2; ++k) {
        h->stats.walkNodes[k] = fc->walkNodes[k];
        h->stats.walkTris[k] = fc->walkTris[k];
        h->stats.walkEligibility[k] = fc->walkEligibility[k];
    }
    h->stats.ssaaPixels = (uint64_t)fc->ssaaPixels;
    if (h->createFlags & RTB_CREATE_COUNTERS) {
        h->stats.boxTestsShadow = fc->boxTestsShadow;
        h->stats.triTestsShadow = fc->triTestsShadow;
        h->stats.boxTests = fc->boxTests + fc->boxTestsShadow;
        h->stats.triTests = fc->triTests + fc->triTestsShadow;
    }
    return fc->overflow;
}

// After an overflow: grow what ran out (the counters say how much was wanted) so the re-run fits.
void growAfterOverflow(RtbHandle* h, int bits, int passes)
{
    if (bits & rtk::OVF_FLAGGED) h->capFlagged = std::max(2 * h->capFlagged, (long long)h->hFrame()->ssaaPixels);
    if (h->tilePipeline) {   // per-tile capacities: the counters do not say how much one tile wanted, so double
        if (bits & rtk::OVF_RAYS) h->tileCapRays *= 2;
        if (bits & rtk::OVF_INTERIORS) h->tileCapInterior *= 2;
        return;
    }
    if (bits & rtk::OVF_RAYS) {
        long long want = 2 * h->capDeep;
        for (int p = 0; p < passes; ++p)
            for (int l = 1; l <= h->levels; ++l) want = std::max(want, (long long)h->hLevel(p, l)->nRays);
        h->capDeep = want;
    }
    if (bits & rtk::OVF_INTERIORS) h->capInterior = std::max(2 * h->capInterior, (long long)h->hFrame()->interiors);
}

void uploadRows(RtbHandle* h, cudaStream_t st, DevBuf& buf, std::vector<int>& resident, const std::vector<int>& rows)
{
    if (rows == resident && buf.p) return;
    buf.reserve(std::max<size_t>(1, rows.size()) * sizeof(int), st, false);
    // the copy is stream-ordered, but `rows` may die before it runs: keep the source alive in `resident`
    CK(cudaStreamSynchronize(st));
    resident = rows;
    CK(cudaMemcpyAsync(buf.p, resident.data(), resident.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    h->stats.h2dBytes += resident.size() * sizeof(int);
}

// ---- one frame (or a rank's rows of it) -----------------------------------------------------------------------------
// A call is split in two so that a frame loop can keep the device busy: beginRows plans the frame and enqueues everything
// (all kernels, the output copy, the counter read-back) on the stream WITHOUT waiting; endRows waits, reads the counters,
// re-runs the frame with larger queues in the rare overflow case and fills the statistics.  The synchronous entry points
// are begin + end; rtb_render*_begin / rtb_render_end expose the halves (a caller may enqueue its own work — e.g. the
// multi-GPU exchange barrier — behind the frame before it waits).
void enqueueAttempt(RtbHandle* h)
{
    RtbHandle::FramePlan& f = h->plan;
    cudaStream_t st = f.st;
    const rt::Scene& sc = h->scene;
    const int w = sc.width, ht = sc.height;
    const long long framePixels = (long long)w * ht;
    const int* rect = h->primRect;
    const std::vector<int>& owned = f.owned;

    const long long level0 = std::max(f.n0, 4 * std::max(f.flaggedCap, h->capFlagged));
    const bool tile = h->tilePipeline;
    ensureCapacity(h, st, framePixels, level0, 2 * level0, 2 * level0, f.flaggedCap, !tile);
    const int sampleBase = (int)framePixels;

    CK(cudaEventRecord(h->ev[0], st));
    if (f.nInitRows > 0) {
        // culled: background colour where a primary ray would miss by construction; else Vec3f() zero-init (scene.cpp:599)
        KernelSpan ks(h, st, RTB_K_RAYGEN);
        // pixels that get a generated primary ray are written by pass 1 itself (hit or miss): skip them.  A listed row inside
        // the rectangle's row range is always a pass-1 row when SSAA is on (initRows = pass-1 rows + the last image row).
        const bool skip = f.ssaa && f.nGenRows > 0 && f.genCols > 0;
        const int skipY0 = f.culled ? rect[2] : 0, skipY1 = f.culled ? rect[3] : ht - 1;
        rtk::k_fill_background<<<gridFor(h, (long long)f.nInitRows * w), rtk::kBlock, 0, st>>>(h->slots.as<float>(), w, ht, h->rowsC.as<int>(),
            f.nInitRows, f.culled ? sc.background : rt::mk(0.0f, 0.0f, 0.0f), skip ? f.genX0 : 0, skip ? f.genX0 + f.genCols : 0, skipY0, skip ? skipY1 : skipY0);
        ks.done();
    }
    if (f.cover) {
        KernelSpan ks(h, st, RTB_K_RAYGEN);
        const int coverLaunches = f.rebuildLists ? 3 : 1;
        if (f.rebuildLists) {
            const int cellsX = (w + 7) / 8;
            CK(cudaMemsetAsync(h->coverBits.p, 0, (size_t)ht * ((cellsX + 31) / 32) * sizeof(unsigned), st));
            CK(cudaMemsetAsync(h->coverCtrDev, 0, sizeof(rtk::CoverCtr), st));
            rtk::k_cover_mark<<<gridFor(h, 32LL * h->nCoverBoxes), rtk::kBlock, 0, st>>>(h->sceneDev, h->coverBoxesDev, h->nCoverBoxes,
                h->coverBits.as<unsigned>(), cellsX, h->coverCtrDev);
            rtk::k_tile_lists<<<gridFor(h, f.tilesTotal), rtk::kBlock, 0, st>>>(h->coverBits.as<unsigned>(), cellsX, h->rowsA.as<int>(), f.nGenRows,
                f.genX0, f.genCols, h->tilesKept.as<int>(), h->tilesSkipped.as<int>(), h->coverCtrDev);
            CK(cudaMemcpyAsync(h->hCover, h->coverCtrDev, sizeof(rtk::CoverCtr), cudaMemcpyDeviceToHost, st));
            h->stats.d2hBytes += sizeof(rtk::CoverCtr);
            f.rebuildLists = false;   // a re-run after a queue overflow finds the lists resident
            f.coverRead = true;
        }
        rtk::k_fill_tiles<<<gridFor(h, 32LL * f.tilesTotal), rtk::kBlock, 0, st>>>(h->slots.as<float>(), w, h->tilesSkipped.as<int>(), h->coverCtrDev,
            h->rowsA.as<int>(), f.nGenRows, f.genX0, f.genCols, sc.background);
        ks.done(coverLaunches);
    }
    CK(cudaMemsetAsync(h->ctrBuf.p, 0, h->ctrBytes, st));
    if (f.n0 > 0) {
        if (f.literalWalk) {   // the literal reference walk reads a materialised queue
            KernelSpan ks(h, st, RTB_K_RAYGEN);
            rtk::k_raygen<<<gridFor(h, f.n0), rtk::kBlock, 0, st>>>(sc, h->rowsA.as<int>(), f.nGenRows, h->rays[0].view(), h->dLevel(0, 0));
            ks.done();
            enqueueLevels(h, st, 0, framePixels);
        } else if (tile) {
            enqueueTile(h, st, 0, rtk::GEN_PRIMARY, rtk::GenArgs{ h->rowsA.as<int>(), f.nGenRows, -1, f.genX0, f.genCols, h->slots.as<float>(), h->sceneDev,
                f.cover ? h->tilesKept.as<int>() : nullptr, f.cover ? &h->coverCtrDev->nKept : nullptr }, f.n0);
        } else {
            enqueueLevels(h, st, 0, framePixels, rtk::GEN_PRIMARY, rtk::GenArgs{ h->rowsA.as<int>(), f.nGenRows, 0, f.genX0, f.genCols, h->slots.as<float>(), h->sceneDev,
                f.cover ? h->tilesKept.as<int>() : nullptr, f.cover ? &h->coverCtrDev->nKept : nullptr });
        }
    }
    CK(cudaEventRecord(h->ev[1], st));

    const bool contiguous = !owned.empty() && owned.back() - owned.front() + 1 == (int)owned.size();
    auto emit = [&](void* dst, OutputKind k, size_t bytes) {
        if (!dst || owned.empty()) return;
        const bool direct = k == OUT_FLOAT && contiguous;   // the rows already lie contiguously in the slot array
        void* target = dst;
        if (!f.fbOnDevice && !direct) {
            h->outStage.reserve(bytes, st, false);
            target = h->outStage.p;
        }
        if (direct) {
            const float* src = h->slots.as<float>() + (size_t)owned.front() * w * 3;
            CK(cudaMemcpyAsync(dst, src, bytes, f.fbOnDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        } else {
            KernelSpan ks(h, st, RTB_K_OUTPUT);
            if (k == OUT_SCATTER)
                rtk::k_scatter_rows<<<gridFor(h, (long long)(bytes / 16)), rtk::kBlock, 0, st>>>(h->slots.as<float>(), w, h->rowsB.as<int>(),
                    (int)owned.size(), static_cast<float*>(target));
            else if (k == OUT_BGR8)
                rtk::k_quantize_bgr8<<<gridFor(h, (long long)(bytes / 4)), rtk::kBlock, 0, st>>>(h->slots.as<float>(), w, h->rowsB.as<int>(),
                    (int)owned.size(), static_cast<unsigned int*>(target));
            else
                rtk::k_gather_rows<<<gridFor(h, (long long)(bytes / 16)), rtk::kBlock, 0, st>>>(h->slots.as<float>(), w, h->rowsB.as<int>(),
                    (int)owned.size(), static_cast<float*>(target));
            ks.done();
            if (!f.fbOnDevice) CK(cudaMemcpyAsync(dst, target, bytes, cudaMemcpyDeviceToHost, st));
        }
        if (!f.fbOnDevice) h->stats.d2hBytes += bytes;
    };
    if (f.pass1) emit(f.pass1, OUT_FLOAT, owned.size() * (size_t)w * 3 * sizeof(float));
    if (f.earlyOut) {
        const int k = h->pipeParity;
        h->pipeParity ^= 1;
        f.earlyStage = k;
        const void* src = h->slots.as<float>() + (size_t)owned.front() * w * 3;     // float rows: straight out of the frame
        if (f.kind == OUT_BGR8) {
            h->pipeStage[k].reserve(f.outBytes, st, false);
            CK(cudaStreamWaitEvent(st, h->pipeCopied[k], 0));
            KernelSpan ks(h, st, RTB_K_OUTPUT);
            rtk::k_quantize_bgr8<<<gridFor(h, (long long)(f.outBytes / 4)), rtk::kBlock, 0, st>>>(h->slots.as<float>(), w, h->rowsB.as<int>(),
                (int)owned.size(), h->pipeStage[k].as<unsigned int>());
            ks.done();
            src = h->pipeStage[k].p;
        }
        // (float rows are read by the copy engine while the SSAA pass rewrites some of them: whatever it catches of those pixels
        // is rewritten below)
        CK(cudaEventRecord(h->pipeReady[k], st));
        CK(cudaStreamWaitEvent(h->copyStream, h->pipeReady[k], 0));
        CK(cudaMemcpyAsync(f.fb, src, f.outBytes, cudaMemcpyDeviceToHost, h->copyStream));
        CK(cudaEventRecord(h->pipeCopied[k], h->copyStream));
        h->stats.d2hBytes += f.outBytes;
    }

    if (f.ssaa) {
        {
            KernelSpan ks(h, st, RTB_K_SOBEL);
            rtk::k_sobel<<<gridFor(h, f.interiorPixels), rtk::kBlock, 0, st>>>(w, ht, h->slots.as<float>(), h->rowsB.as<int>(),
                (int)owned.size(), h->flagged.as<int>(), (int)h->capFlagged, h->dFrame());
            ks.done();
        }
        CK(cudaEventRecord(h->ev[2], st));
        if (f.literalWalk) {
            KernelSpan ks(h, st, RTB_K_RAYGEN);
            rtk::k_ssaa_gen<<<gridFor(h, 4 * h->capFlagged), rtk::kBlock, 0, st>>>(sc, h->flagged.as<int>(), (int)h->capFlagged, sampleBase,
                h->rays[0].view(), h->dFrame(), h->dLevel(1, 0));
            ks.done();
            enqueueLevels(h, st, 1, framePixels);
        } else if (tile) {
            // the tile kernel also takes the mean of each pixel's 4 samples: no separate resolve
            const long long hint = 4 * std::max(1024LL, h->flaggedSeen > 0 ? h->flaggedSeen : h->capFlagged / 4);
            enqueueTile(h, st, 1, rtk::GEN_SSAA, rtk::GenArgs{ h->flagged.as<int>(), (int)h->capFlagged, -1, 0, 0, h->slots.as<float>(), h->sceneDev },
                std::min(hint, 4 * h->capFlagged));
        } else {
            enqueueLevels(h, st, 1, framePixels, rtk::GEN_SSAA, rtk::GenArgs{ h->flagged.as<int>(), (int)h->capFlagged, sampleBase, 0, 0, h->slots.as<float>(), h->sceneDev });
        }
        if (!tile) {
            KernelSpan ks(h, st, RTB_K_OUTPUT);
            rtk::k_ssaa_resolve<<<gridFor(h, h->capFlagged), rtk::kBlock, 0, st>>>(h->flagged.as<int>(), (int)h->capFlagged, sampleBase,
                h->slots.as<float>(), h->dFrame());
            ks.done();
        }
    } else {
        CK(cudaEventRecord(h->ev[2], st));
    }
    if (f.earlyOut) {
        // the pass-1 bytes have arrived (or are about to); rewrite what the SSAA pass changed
        CK(cudaStreamWaitEvent(st, h->pipeCopied[f.earlyStage], 0));
        KernelSpan ks(h, st, RTB_K_OUTPUT);
        const int grid = gridFor(h, std::max(1LL, h->flaggedSeen > 0 ? h->flaggedSeen : h->capFlagged));
        if (f.kind == OUT_BGR8)
            rtk::k_patch_rows<false><<<grid, rtk::kBlock, 0, st>>>(h->slots.as<float>(), w, owned.front(), owned.back() + 1, h->flagged.as<int>(),
                (int)h->capFlagged, h->dFrame(), static_cast<unsigned int*>(f.fbDev), (long long)(f.outBytes / 4));
        else
            rtk::k_patch_rows<true><<<grid, rtk::kBlock, 0, st>>>(h->slots.as<float>(), w, owned.front(), owned.back() + 1, h->flagged.as<int>(),
                (int)h->capFlagged, h->dFrame(), static_cast<unsigned int*>(f.fbDev), (long long)(f.outBytes / 4));
        ks.done();
    } else if (f.pipelinedCopy && f.fb && !owned.empty()) {
        // bytes -> staging buffer k on the render stream; the copy to the host runs on copyStream behind an event, and the render
        // stream only waits for it again when buffer k is reused two frames later
        const int k = h->pipeParity;
        h->pipeParity ^= 1;
        h->pipeStage[k].reserve(f.outBytes, st, false);
        CK(cudaStreamWaitEvent(st, h->pipeCopied[k], 0));
        KernelSpan ks(h, st, RTB_K_OUTPUT);
        rtk::k_quantize_bgr8<<<gridFor(h, (long long)(f.outBytes / 4)), rtk::kBlock, 0, st>>>(h->slots.as<float>(), w, h->rowsB.as<int>(),
            (int)owned.size(), h->pipeStage[k].as<unsigned int>());
        ks.done();
        CK(cudaEventRecord(h->pipeReady[k], st));
        CK(cudaStreamWaitEvent(h->copyStream, h->pipeReady[k], 0));
        CK(cudaMemcpyAsync(f.fb, h->pipeStage[k].p, f.outBytes, cudaMemcpyDeviceToHost, h->copyStream));
        CK(cudaEventRecord(h->pipeCopied[k], h->copyStream));
        h->stats.d2hBytes += f.outBytes;
    } else {
        emit(f.fb, f.kind, f.outBytes);
    }
    CK(cudaEventRecord(h->ev[3], st));
    // the frame's counters, read back behind everything else (endRows waits for them)
    if (h->hCtrDev) {
        KernelSpan ks(h, st, RTB_K_OUTPUT);
        rtk::k_store_words<<<1, 256, 0, st>>>(h->ctrBuf.as<unsigned int>(), static_cast<unsigned int*>(h->hCtrDev), (int)(h->ctrBytes / 4));
        ks.done();
    } else {
        CK(cudaMemcpyAsync(h->hCtr, h->ctrBuf.p, h->ctrBytes, cudaMemcpyDeviceToHost, st));
    }
    h->stats.d2hBytes += h->ctrBytes;
}

// Plans the rows in `owned` (ascending) and enqueues the frame; output is compact (owned rows in order) unless kind == OUT_SCATTER.
void beginRows(RtbHandle* h, const std::vector<int>& owned, void* fb, float* pass1, int fbOnDevice, void* stream, OutputKind kind, bool pipelinedCopy = false)
{
    if (h->plan.active) throw std::runtime_error("a frame is already in flight on this handle: call rtb_render_end first");
    cudaStream_t st = stream ? (cudaStream_t)stream : h->ownStream;
    beginCall(h);
    uploadSceneHeader(h, st);
    const rt::Scene& sc = h->scene;
    const int w = sc.width, ht = sc.height;
    RtbHandle::FramePlan& f = h->plan;
    f = RtbHandle::FramePlan{};
    f.st = st; f.owned = owned; f.fb = fb; f.pass1 = pass1; f.fbOnDevice = fbOnDevice; f.kind = kind;
    f.pipelinedCopy = pipelinedCopy;
    f.ssaa = (sc.flags & rt::FLAG_SSAA) && !owned.empty();

    // pass-1 rows: owned rows plus a one-row halo for the Sobel window, minus the never-rendered last row
    std::vector<int> p1rows;
    {
        std::vector<char> need(ht, 0);
        for (int y : owned)
            for (int dy = f.ssaa ? -1 : 0; dy <= (f.ssaa ? 1 : 0); ++dy)
                if (y + dy >= 0 && y + dy < ht - 1) need[y + dy] = 1;
        for (int y = 0; y < ht; ++y) if (need[y]) p1rows.push_back(y);
    }
    // primary rays are generated only inside the screen-space bounds of the geometry (computePrimaryRect); the pixels
    // outside are misses by construction and receive the background colour from k_fill_background
    f.literalWalk = h->createFlags & (RTB_CREATE_COUNTERS | RTB_CREATE_EXACT_WALK);
    const int* rect = h->primRect;
    f.culled = !f.literalWalk && (rect[0] > 0 || rect[1] < w - 1 || rect[2] > 0 || rect[3] < ht - 1);
    std::vector<int> genRows;
    if (f.culled) { for (int y : p1rows) if (y >= rect[2] && y < rect[3]) genRows.push_back(y); }
    else genRows = p1rows;
    f.genX0 = f.culled ? rect[0] : 0;
    f.genCols = f.culled ? rect[1] - rect[0] : w - 1;
    f.nGenRows = (int)genRows.size();
    uploadRows(h, st, h->rowsA, h->rowsAHost, genRows);
    uploadRows(h, st, h->rowsB, h->rowsBHost, owned);
    // rows this call initialises: what it renders plus what its Sobel windows read (incl. the never-rendered last row)
    std::vector<int> initRows;
    {
        std::vector<char> need(ht, 0);
        for (int y : p1rows) need[y] = 1;
        for (int y : owned)
            for (int dy = -1; dy <= 1; ++dy)
                if (y + dy >= 0 && y + dy < ht) need[y + dy] = 1;
        for (int y = 0; y < ht; ++y) if (need[y]) initRows.push_back(y);
    }
    f.nInitRows = (int)initRows.size();
    uploadRows(h, st, h->rowsC, h->rowsCHost, initRows);

    f.nPixels = (long long)p1rows.size() * (w - 1);
    f.n0 = (!genRows.empty() && f.genCols > 0) ? rtk::raygenPaddedCount(f.genCols + 1, (int)genRows.size()) : 0;   // whole 8x4 tiles, padding lanes idle
    f.livePixels = f.culled ? (long long)genRows.size() * f.genCols : f.nPixels;
    static const bool noCover = getenv("RTB_NO_COVER") != nullptr;
    if (f.culled && f.n0 > 0 && h->nCoverBoxes > 0 && !noCover) {
        // Tiles of the generation rectangle that lie outside the projected coverage of the geometry get no rays at all.  The
        // two lists are built ON THE DEVICE by this frame's first kernels whenever the camera or the rows changed (a camera
        // sweep costs the host nothing) and stay resident otherwise; the host sizes grids with the grid's tile count and
        // learns the kept count / live pixels from the counter read-back.
        const int tilesX = (f.genCols + 7) / 8, tilesY = ((int)genRows.size() + 3) / 4, cellsX = (w + 7) / 8;
        f.cover = true;
        f.tilesTotal = tilesX * tilesY;
        if (h->tileListCamera != h->cameraVersion || h->tileListRows != genRows) {
            f.rebuildLists = true;
            h->coverBits.reserve((size_t)ht * ((cellsX + 31) / 32) * sizeof(unsigned), st, false);
            h->tilesKept.reserve((size_t)f.tilesTotal * sizeof(int), st, false);
            h->tilesSkipped.reserve((size_t)f.tilesTotal * sizeof(int), st, false);
            h->tileListCamera = h->cameraVersion;
            h->tileListRows = genRows;
            h->tileListLive = -1;
            h->tileListKept = -1;
        }
        f.n0 = 32LL * (h->tileListKept >= 0 ? h->tileListKept : f.tilesTotal);
    }
    f.interiorPixels = f.ssaa ? (long long)owned.size() * w : 0;
    // SSAA capacity: what the last frame flagged plus head-room, at least 1/16 of the owned pixels; a frame that
    // flags more sets OVF_FLAGGED and is re-run with the exact count
    f.flaggedCap = f.ssaa ? std::min(f.interiorPixels, std::max(f.interiorPixels / 16, h->flaggedSeen + h->flaggedSeen / 4 + 1024)) : 0;
    const size_t outRowBytes = kind == OUT_BGR8 ? (size_t)((w * 3 + 3) & ~3) : (size_t)w * 3 * sizeof(float);
    f.outBytes = owned.size() * outRowBytes;
    // early output needs: bytes for a host buffer the device can address (pinned), an SSAA pass to hide the copy behind, one
    // contiguous range of rows
    // ... and few enough re-traced pixels: rewriting one costs about as much PCIe time as copying kEarlyOutRatio pixels' bytes
    // (measured), so the last frame's count decides (RTB_EARLY_OUTPUT=0 / 1: never / always)
    const char* earlyEnv = getenv("RTB_EARLY_OUTPUT");
    const int earlyMode = earlyEnv ? atoi(earlyEnv) : -1;
    const bool fewFlagged = h->flaggedSeen > 0 && h->flaggedSeen * kEarlyOutRatio < (long long)owned.size() * w;
    // (a frame loop on the begin / end halves hides the copy behind the NEXT frame instead: pipelinedCopy)
    if ((earlyMode == 1 || (earlyMode < 0 && fewFlagged && !pipelinedCopy)) && (kind == OUT_BGR8 || kind == OUT_FLOAT) && !fbOnDevice && fb && f.ssaa
        && !f.literalWalk && owned.back() - owned.front() + 1 == (int)owned.size()) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, fb) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer) {
            f.earlyOut = true;
            f.fbDev = attr.devicePointer;
        } else {
            cudaGetLastError();
        }
    }
    f.active = true;
    enqueueAttempt(h);
}

int endRows(RtbHandle* h, RtbStats* statsOut)
{
    RtbHandle::FramePlan& f = h->plan;
    if (!f.active) throw std::runtime_error("rtb_render_end without a frame in flight");
    struct Done { RtbHandle::FramePlan& f; ~Done() { f.active = false; } } done{ f };
    const int passes = f.ssaa ? 2 : 1;
    for (int attempt = 0;; ++attempt) {
        const int overflow = finishFrame(h, f.st, passes);
        if (!overflow) break;
        if (attempt >= 8 + h->levels) throw CudaError{ cudaErrorMemoryAllocation, "wavefront queues still overflow after repeated growth" };
        // discard this attempt's statistics and run the frame again with larger queues
        growAfterOverflow(h, overflow, passes);
        f.flaggedCap = std::max(f.flaggedCap, h->capFlagged);
        const uint64_t h2d = h->stats.h2dBytes;
        resolveSpans(h);
        h->stats = RtbStats{};
        h->stats.msBuildSearchBvh = h->msBuildSearchBvh;
        h->stats.h2dBytes = h2d;
        enqueueAttempt(h);
    }
    resolveSpans(h);
    h->flaggedSeen = (long long)h->stats.ssaaPixels;
    h->stats.primaryRays = (uint64_t)f.nPixels + 4 * h->stats.ssaaPixels;
    if (f.cover) {
        if (f.coverRead) { h->tileListLive = h->hCover->livePixels; h->tileListKept = h->hCover->nKept; }
        f.livePixels = h->tileListLive;
    }
    h->stats.backgroundPixels = (uint64_t)(f.nPixels - f.livePixels);
    h->stats.rays = h->stats.primaryRays + h->stats.secondaryRays + h->stats.shadowRays;
    h->stats.msPass1 = elapsed(h->ev[0], h->ev[1]);
    h->stats.msSobel = elapsed(h->ev[1], h->ev[2]);
    h->stats.msSSAA = elapsed(h->ev[2], h->ev[3]);
    h->stats.msTotal = elapsed(h->ev[0], h->ev[3]);
    if (statsOut) *statsOut = h->stats;
    return RTB_OK;
}

int renderRows(RtbHandle* h, const std::vector<int>& owned, void* fb, float* pass1, int fbOnDevice, void* stream, RtbStats* statsOut,
    OutputKind kind = OUT_FLOAT)
{
    beginRows(h, owned, fb, pass1, fbOnDevice, stream, kind);
    return endRows(h, statsOut);
}

// castRay / trace on caller-supplied rays: one pass, level-0 queue = the rays themselves
void enqueueUserRays(RtbHandle* h, cudaStream_t st, const float* rays, int nRays, bool shade)
{
    beginCall(h);
    uploadSceneHeader(h, st);
    for (int attempt = 0;; ++attempt) {
        {
            const bool tileCast = shade && h->tilePipeline;
            ensureCapacity(h, st, nRays, nRays, tileCast ? 0 : 2LL * nRays, tileCast ? 0 : 2LL * nRays, 0);
        }
        h->userRays.reserve((size_t)nRays * 6 * sizeof(float), st, false);
        CK(cudaMemcpyAsync(h->userRays.p, rays, (size_t)nRays * 6 * sizeof(float), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(h->ctrBuf.p, 0, h->ctrBytes, st));
        rtk::k_rays_from_user<<<gridFor(h, nRays), rtk::kBlock, 0, st>>>(h->userRays.as<float>(), nRays, 0, h->rays[0].view(), h->dLevel(0, 0));
        launchCheck();
        if (shade && h->tilePipeline) {
            enqueueTile(h, st, 0, rtk::GEN_QUEUE, rtk::GenArgs{ nullptr, 0, -1, 0, 0, h->slots.as<float>(), h->sceneDev }, nRays, h->rays[0].view(), nRays);
        } else if (shade) {
            enqueueLevels(h, st, 0, nRays);
        } else {
            const rtk::HitQueue hits{ h->hitTuv.as<float4>(), h->hitObj.as<int>() };
            if (h->createFlags & (RTB_CREATE_EXACT_WALK | RTB_CREATE_COUNTERS))
                rtk::k_trace<rtk::MODE_EXACT><<<gridFor(h, nRays), rtk::kBlock, stackBytes(h), st>>>(h->scene, h->rays[0].view(), nRays, hits, h->dFrame(), h->dLevel(0, 0));
            else
                rtk::k_walk<false, rtk::GEN_QUEUE><<<persistentGrid(h, 0, nRays), rtk::kBlock, stackBytes(h), st>>>(h->scene, h->rays[0].view(), nRays, hits,
                    rtk::SurfQueue{}, nullptr, h->dFrame(), h->dLevel(0, 0), rtk::GenArgs{});
            launchCheck();
        }
        const int overflow = finishFrame(h, st, 1, true);
        if (!overflow) break;
        if (attempt >= 8 + h->levels) throw CudaError{ cudaErrorMemoryAllocation, "wavefront queues still overflow after repeated growth" };
        growAfterOverflow(h, overflow, 1);
        resolveSpans(h);
        h->stats = RtbStats{};
    }
    resolveSpans(h);
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
    } catch (const std::invalid_argument& e) {
        g_err = e.what();
        return RTB_ERR_ARG;
    } catch (const std::exception& e) {
        g_err = e.what();
        return RTB_ERR_CUDA;
    }
}

// Cyclic strips counted from row `origin` (the first row that can contain geometry): strip s = floor((y - origin) / stripRowsN)
// belongs to rank s mod world, so the rows that cost something are dealt evenly whatever lies above them.
std::vector<int> stripRows(int height, int stripRowsN, int rank, int world, int origin)
{
    std::vector<int> rows;
    for (int y = 0; y < height; ++y) {
        const int d = y - origin;
        const int s = d >= 0 ? d / stripRowsN : -((-d + stripRowsN - 1) / stripRowsN);
        if (((s % world) + world) % world == rank) rows.push_back(y);
    }
    return rows;
}

int stripOrigin(const RtbHandle* h)
{
    const bool literalWalk = h->createFlags & (RTB_CREATE_COUNTERS | RTB_CREATE_EXACT_WALK);
    return literalWalk ? 0 : std::max(0, std::min(h->primRect[2], h->scene.height - 1));
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
             &h->rowsA, &h->rowsB, &h->rowsC, &h->userRays, &h->outStage, &h->tileSlab, &h->tilesKept, &h->tilesSkipped, &h->coverBits })
        b->release();
    h->ctrBuf.release();
    if (h->hCtr) cudaFreeHost(h->hCtr);
    if (h->hCover) cudaFreeHost(h->hCover);
    if (h->scenePinned) cudaFreeHost(h->scenePinned);
    for (cudaEvent_t e : h->ev) if (e) cudaEventDestroy(e);
    if (h->copyStream) { cudaStreamSynchronize(h->copyStream); cudaStreamDestroy(h->copyStream); }
    for (int k = 0; k < 2; ++k) {
        if (h->pipeReady[k]) cudaEventDestroy(h->pipeReady[k]);
        if (h->pipeCopied[k]) cudaEventDestroy(h->pipeCopied[k]);
        h->pipeStage[k].release();
    }
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
        h->kernelTiming = createFlags & RTB_CREATE_KERNEL_TIMING;
        // the literal reference walk (parity of work counters) only exists in the frame-wide form
        h->tilePipeline = !(createFlags & (RTB_CREATE_WAVEFRONT | RTB_CREATE_COUNTERS | RTB_CREATE_EXACT_WALK));
        cudaDeviceProp prop{};
        CK(cudaGetDeviceProperties(&prop, device));
        h->smCount = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&h->ownStream, cudaStreamNonBlocking));
        for (cudaEvent_t& e : h->ev) CK(cudaEventCreate(&e));
        CK(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            CK(cudaEventCreateWithFlags(&h->pipeReady[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->pipeCopied[k], cudaEventDisableTiming));
        }

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
            h->refTreeDepth.push_back(pm.maxDepth);
            if (createFlags & (RTB_CREATE_COUNTERS | RTB_CREATE_EXACT_WALK)) h->stackEntries = std::max(h->stackEntries, pm.maxDepth + 1);
            // a mesh whose .obj yields no usable face still has a one-node tree (objects.cpp:389): nothing to search, the
            // kernels skip it (bvhTris stays null)
            if (m.nNodes > 0 && m.nTris > 0) h->stackEntries = std::max(h->stackEntries, buildFastPath(h, m, d, i) + 1);
            meshes.push_back(d);
        }
        for (int i = 0; i < s->nObjects; ++i) {
            rtpack::objectBounds(s->objects[i], h->geomBounds, h->unbounded);
            bool dummy = false;
            rtpack::objectBounds(s->objects[i], h->coverBounds, dummy);
        }
        if (!h->coverBounds.empty() && !h->unbounded) {
            h->coverBoxesDev = upload(h, &h->coverBounds[0][0], h->coverBounds.size() * 6);
            h->nCoverBoxes = (int)h->coverBounds.size();
            void* c = nullptr;
            CK(cudaMalloc(&c, sizeof(rtk::CoverCtr)));
            h->allocations.push_back(c);
            h->coverCtrDev = static_cast<rtk::CoverCtr*>(c);
            CK(cudaMallocHost(&h->hCover, sizeof(rtk::CoverCtr)));
        }
        computePrimaryRect(h);
        // recursion levels: only Reflective / Transparent hits spawn children (scene.cpp:854-941)
        bool spawns = false;
        for (int i = 0; i < s->nObjects; ++i)
            spawns |= s->objects[i].material == RTB_MAT_REFLECTIVE || s->objects[i].material == RTB_MAT_TRANSPARENT;
        h->levels = spawns ? std::max(0, s->maxRayDepth) + 1 : 1;
        h->spawns = spawns;
        if (h->levels > rtk::kMaxTileLevels) h->tilePipeline = false;   // the per-tile level table is fixed-size; deeper trees run frame-wide
        h->ctrBytes = sizeof(rtk::FrameCtr) + 2 * (size_t)(h->levels + 1) * sizeof(rtk::LevelCtr);
        h->ctrBuf.reserve(h->ctrBytes, h->ownStream, false);
        CK(cudaMallocHost(&h->hCtr, h->ctrBytes));
        std::memset(h->hCtr, 0, h->ctrBytes);
        if (cudaHostGetDevicePointer(&h->hCtrDev, h->hCtr, 0) != cudaSuccess) { h->hCtrDev = nullptr; cudaGetLastError(); }
        // one resident wave of the persistent traversal kernels with this scene's stack size
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->walkBlocksPerSm[0], rtk::k_walk<false, rtk::GEN_QUEUE>, rtk::kBlock, stackBytes(h)));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->walkBlocksPerSm[1], rtk::k_walk<false, rtk::GEN_PRIMARY>, rtk::kBlock, stackBytes(h)));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->walkBlocksPerSm[2], rtk::k_walk<false, rtk::GEN_SSAA>, rtk::kBlock, stackBytes(h)));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->walkBlocksPerSm[3], rtk::k_walk<true, rtk::GEN_QUEUE>, rtk::kBlock, stackBytes(h)));
        for (int& b : h->walkBlocksPerSm) b = std::max(1, b);
        h->scene.objects = upload(h, objects.data(), objects.size());
        h->scene.lights = upload(h, lights.data(), lights.size());
        h->scene.meshes = upload(h, meshes.data(), meshes.size());
        // staged tile kernel: how much of the largest search BVH fits into shared memory next to four groups' stacks
        h->smemOptin = prop.sharedMemPerBlockOptin;
        if (h->stagedMeshIndex >= 0) {
            const rt::Mesh& sm = meshes[h->stagedMeshIndex];
            h->stagedNodesSrc = sm.bvhNodes;
            h->stagedTrisSrc = sm.bvhTris;
            const long long avail = (long long)h->smemOptin - 1024 - (long long)rtk::tileSmemStagedOffset(kStagedGroups, h->stackEntries);
            h->stagedNodes = (int)std::max(0LL, std::min<long long>(h->stagedNodesAll, avail / 64));
            const long long left = avail - (long long)h->stagedNodes * 64;
            h->stagedTris = (h->stagedNodes == h->stagedNodesAll && left >= (long long)h->stagedTrisAll * 48) ? h->stagedTrisAll : 0;
        }
        if (const char* e = getenv("RTB_TILE_STAGED")) h->tileStaged = atoi(e) != 0;
        h->scene.areaPoints = upload(h, s->areaPoints, (size_t)s->nAreaPoints * 3);
        if (s->flags & RTB_FLAG_USE_SKYBOX) {
            // getSkybox indexes every face with one width / height (scene.cpp:381-442): six faces, all present, one size
            for (int k = 0; k < 6; ++k)
                if (!s->skybox[k].rgb || s->skybox[k].width <= 0 || s->skybox[k].height <= 0 || s->skybox[k].width != s->skybox[0].width
                    || s->skybox[k].height != s->skybox[0].height)
                    throw std::invalid_argument("RTB_FLAG_USE_SKYBOX needs six skybox faces of one size");
            for (int k = 0; k < 6; ++k) h->scene.sky[k] = uploadImage(h, s->skybox[k]);
        }
        // the header's resident copy (out-of-line device helpers read it): only now are all of its pointers final
        h->sceneDev = upload(h, &h->scene, 1);
        CK(cudaMallocHost(&h->scenePinned, sizeof(rt::Scene)));
        return RTB_OK;
    });
    if (rc != RTB_OK) { destroyHandle(h); return rc; }
    *out = h;
    return RTB_OK;
}

int rtb_set_camera(RtbHandle* h, const RtbCamera* camera)
{
    if (!h || !camera) { g_err = "null argument"; return RTB_ERR_ARG; }
    return guarded([&]() {
        const rt::V3 pos = rtpack::v3of(camera->pos);
        bool same = std::memcmp(&pos, &h->scene.camPos, sizeof pos) == 0 && std::memcmp(&camera->scale, &h->scene.camScale, sizeof(float)) == 0
            && std::memcmp(&camera->aspect, &h->scene.camAspect, sizeof(float)) == 0;
        for (int i = 0; i < 16; ++i) same = same && std::memcmp(&camera->rMatrix[i], &h->scene.camM[i], sizeof(float)) == 0;
        h->scene.camPos = pos;
        for (int i = 0; i < 16; ++i) h->scene.camM[i] = camera->rMatrix[i];
        h->scene.camScale = camera->scale;
        h->scene.camAspect = camera->aspect;
        *h->scenePinned = h->scene;          // uploaded by the next render call on its own stream (uploadSceneHeader)
        h->sceneDirty = true;
        h->pendingH2D += sizeof(rt::Scene);
        if (!same) computePrimaryRect(h);    // a camera that did not move keeps the rectangle and the resident tile lists
        return RTB_OK;
    });
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

int rtb_render_bgr8(RtbHandle* h, int y0, int y1, uint8_t* bgr, int onDevice, void* stream, RtbStats* stats)
{
    if (!h || !bgr) { g_err = "null argument"; return RTB_ERR_ARG; }
    if (y0 < 0 || y1 > h->scene.height || y0 > y1) { g_err = "row range outside the image"; return RTB_ERR_ARG; }
    return guarded([&]() {
        std::vector<int> rows;
        for (int y = y0; y < y1; ++y) rows.push_back(y);
        return renderRows(h, rows, bgr, nullptr, onDevice, stream, stats, OUT_BGR8);
    });
}

int rtb_render_ac(RtbHandle* h, float* fb, int32_t* counts, int onDevice, void* stream, RtbStats* stats)
{
    if (!h || !fb) { g_err = "null argument"; return RTB_ERR_ARG; }
    return guarded([&]() {
        cudaStream_t st = stream ? (cudaStream_t)stream : h->ownStream;
        beginCall(h);
        uploadSceneHeader(h, st);
        const int total = h->scene.width * h->scene.height;
        // the walk follows the REFERENCE tree, whose depth bounds the stack
        int depth = 1;
        for (int d : h->refTreeDepth) depth = std::max(depth, d + 1);
        const size_t smem = (size_t)depth * rtk::kBlock * sizeof(int);
        h->flagged.reserve((size_t)total * sizeof(int), st, false);          // per-pixel counts
        h->slots.reserve((size_t)total * 3 * sizeof(float), st, false);
        CK(cudaEventRecord(h->ev[0], st));
        CK(cudaMemsetAsync(h->ctrBuf.p, 0, h->ctrBytes, st));
        {
            KernelSpan ks(h, st, RTB_K_TRACE);
            rtk::k_count_ac<<<gridFor(h, total), rtk::kBlock, smem, st>>>(h->scene, h->flagged.as<int>(), h->dFrame());
            ks.done();
        }
        {
            KernelSpan ks(h, st, RTB_K_OUTPUT);
            rtk::k_ac_resolve<<<gridFor(h, total), rtk::kBlock, 0, st>>>(h->flagged.as<int>(), total, h->dFrame(), h->slots.as<float>());
            ks.done();
        }
        CK(cudaEventRecord(h->ev[3], st));
        const cudaMemcpyKind kind = onDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        CK(cudaMemcpyAsync(fb, h->slots.p, (size_t)total * 3 * sizeof(float), kind, st));
        if (counts) CK(cudaMemcpyAsync(counts, h->flagged.p, (size_t)total * sizeof(int), kind, st));
        if (!onDevice) h->stats.d2hBytes += (size_t)total * (3 * sizeof(float) + (counts ? sizeof(int) : 0));
        CK(cudaStreamSynchronize(st));
        resolveSpans(h);
        h->stats.primaryRays = (uint64_t)total;
        h->stats.msTotal = elapsed(h->ev[0], h->ev[3]);
        if (stats) *stats = h->stats;
        return RTB_OK;
    });
}

int rtb_strip_rows_owned(int height, int stripRowsN, int rank, int worldSize)
{
    if (height <= 0 || stripRowsN <= 0 || worldSize <= 0 || rank < 0 || rank >= worldSize) return RTB_ERR_ARG;
    return (int)stripRows(height, stripRowsN, rank, worldSize, 0).size();
}

int rtb_render_strips(RtbHandle* h, int stripRowsN, int rank, int worldSize, float* fb, int fbOnDevice, void* stream,
    int* nRowsOut, RtbStats* stats)
{
    if (!h || !fb) { g_err = "null argument"; return RTB_ERR_ARG; }
    if (stripRowsN <= 0 || worldSize <= 0 || rank < 0 || rank >= worldSize) { g_err = "bad strip partition"; return RTB_ERR_ARG; }
    return guarded([&]() {
        const std::vector<int> rows = stripRows(h->scene.height, stripRowsN, rank, worldSize, stripOrigin(h));
        if (nRowsOut) *nRowsOut = (int)rows.size();
        return renderRows(h, rows, fb, nullptr, fbOnDevice, stream, stats);
    });
}

int rtb_render_strips_to_frame(RtbHandle* h, int stripRowsN, int rank, int worldSize, float* frame, void* stream, RtbStats* stats)
{
    if (!h || !frame) { g_err = "null argument"; return RTB_ERR_ARG; }
    if (stripRowsN <= 0 || worldSize <= 0 || rank < 0 || rank >= worldSize) { g_err = "bad strip partition"; return RTB_ERR_ARG; }
    return guarded([&]() {
        const std::vector<int> rows = stripRows(h->scene.height, stripRowsN, rank, worldSize, stripOrigin(h));
        return renderRows(h, rows, frame, nullptr, 1, stream, stats, OUT_SCATTER);
    });
}

int rtb_render_begin(RtbHandle* h, int y0, int y1, float* fb, int fbOnDevice, void* stream)
{
    if (!h || !fb) { g_err = "null argument"; return RTB_ERR_ARG; }
    if (y0 < 0 || y1 > h->scene.height || y0 > y1) { g_err = "row range outside the image"; return RTB_ERR_ARG; }
    return guarded([&]() {
        std::vector<int> rows;
        for (int y = y0; y < y1; ++y) rows.push_back(y);
        beginRows(h, rows, fb, nullptr, fbOnDevice, stream, OUT_FLOAT);
        return RTB_OK;
    });
}

int rtb_render_strips_to_frame_begin(RtbHandle* h, int stripRowsN, int rank, int worldSize, float* frame, void* stream)
{
    if (!h || !frame) { g_err = "null argument"; return RTB_ERR_ARG; }
    if (stripRowsN <= 0 || worldSize <= 0 || rank < 0 || rank >= worldSize) { g_err = "bad strip partition"; return RTB_ERR_ARG; }
    return guarded([&]() {
        beginRows(h, stripRows(h->scene.height, stripRowsN, rank, worldSize, stripOrigin(h)), frame, nullptr, 1, stream, OUT_SCATTER);
        return RTB_OK;
    });
}

int rtb_render_bgr8_begin(RtbHandle* h, int y0, int y1, uint8_t* bgrHost)
{
    if (!h || !bgrHost) { g_err = "null argument"; return RTB_ERR_ARG; }
    if (y0 < 0 || y1 > h->scene.height || y0 > y1) { g_err = "row range outside the image"; return RTB_ERR_ARG; }
    return guarded([&]() {
        std::vector<int> rows;
        for (int y = y0; y < y1; ++y) rows.push_back(y);
        beginRows(h, rows, bgrHost, nullptr, 0, nullptr, OUT_BGR8, true);
        return RTB_OK;
    });
}

int rtb_output_sync(RtbHandle* h)
{
    if (!h) { g_err = "null argument"; return RTB_ERR_ARG; }
    return guarded([&]() {
        CK(cudaSetDevice(h->device));
        CK(cudaStreamSynchronize(h->copyStream));
        return RTB_OK;
    });
}

int rtb_render_end(RtbHandle* h, RtbStats* stats)
{
    if (!h) { g_err = "null argument"; return RTB_ERR_ARG; }
    return guarded([&]() { return endRows(h, stats); });
}

int rtb_frame_to_bgr8(RtbHandle* h, const float* frame, uint8_t* bgr, int onDevice, void* stream)
{
    if (!h || !frame || !bgr) { g_err = "null argument"; return RTB_ERR_ARG; }
    return guarded([&]() {
        CK(cudaSetDevice(h->device));
        cudaStream_t st = stream ? (cudaStream_t)stream : h->ownStream;
        const int w = h->scene.width, ht = h->scene.height;
        std::vector<int> rows(ht);
        for (int y = 0; y < ht; ++y) rows[y] = y;
        RtbStats keep = h->stats;
        uploadRows(h, st, h->rowsB, h->rowsBHost, rows);
        h->stats = keep;
        const size_t bytes = (size_t)ht * ((w * 3 + 3) & ~3);
        void* target = bgr;
        if (!onDevice) { h->outStage.reserve(bytes, st, false); target = h->outStage.p; }
        rtk::k_quantize_bgr8<<<gridFor(h, (long long)(bytes / 4)), rtk::kBlock, 0, st>>>(frame, w, h->rowsB.as<int>(), ht, static_cast<unsigned int*>(target));
        launchCheck();
        if (!onDevice) CK(cudaMemcpyAsync(bgr, target, bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return RTB_OK;
    });
}

int rtb_trace(RtbHandle* h, const float* rays, int nRays, float* tuv, int32_t* objTri)
{
    if (!h || !rays || !tuv || !objTri || nRays < 0) { g_err = "bad argument"; return RTB_ERR_ARG; }
    if (nRays == 0) return RTB_OK;
    return guarded([&]() {
        cudaStream_t st = h->ownStream;
        enqueueUserRays(h, st, rays, nRays, false);
        std::vector<float4> t4(nRays);
        std::vector<int> ob(nRays);
        CK(cudaMemcpyAsync(t4.data(), h->hitTuv.p, (size_t)nRays * sizeof(float4), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ob.data(), h->hitObj.p, (size_t)nRays * sizeof(int), cudaMemcpyDeviceToHost, st));
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
        enqueueUserRays(h, st, rays, nRays, true);
        CK(cudaMemcpyAsync(rgb, h->slots.p, (size_t)nRays * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return RTB_OK;
    });
}

int rtb_strip_origin(const RtbHandle* h) { return h ? stripOrigin(h) : RTB_ERR_ARG; }

int rtb_strip_rows(int height, int stripRowsN, int origin, int rank, int worldSize, int32_t* rowsOut)
{
    if (height <= 0 || stripRowsN <= 0 || worldSize <= 0 || rank < 0 || rank >= worldSize) return RTB_ERR_ARG;
    const std::vector<int> rows = stripRows(height, stripRowsN, rank, worldSize, origin);
    if (rowsOut) std::copy(rows.begin(), rows.end(), rowsOut);
    return (int)rows.size();
}

int rtb_device_of(const RtbHandle* h) { return h ? h->device : RTB_ERR_ARG; }

void rtb_destroy(RtbHandle* h) { destroyHandle(h); }

} // extern "C"
