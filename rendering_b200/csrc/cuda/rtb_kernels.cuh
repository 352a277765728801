// rtb_kernels.cuh — wavefront ray-tracing kernels (sm_100a).
//
// The reference renders one pixel at a time with a recursive castRay (scene.cpp:758-946).  Here a
// frame is a sequence of breadth-first LEVELS (level = recursion depth); each level runs as
// separate STAGES over compacted queues.  The stage bodies below (walkRays, surfaceStage, shadeStage, combineStage) are
// used twice: by the default tile pipeline (rtb_tile.cuh: k_tile runs all stages of a 256-ray tile inside one persistent
// kernel per pass) and by the frame-wide kernels of this file, one launch per stage and level (RTB_CREATE_WAVEFRONT, the
// literal-walk / counting handles):
//
//   k_walk<false, GEN>      closest hit over all objects (Render::trace + the search BVH); generates the primary
//                           and SSAA rays itself (renderWorker, SSAAworker) and resolves their misses
//   k_surface               hit -> surface record (compacts hits); misses of queued rays -> skybox   (castRay :762-775, :945)
//   k_walk<true>            any-hit shadow rays, one per (surface, light sample)   (castRay :785-787 ...)
//   k_shade                 light accumulation per material; resolves Diffuse/Phong, spawns the
//                           Reflective / Transparent children into the next level's queue
//   k_combine               folds child colours into their parents, deepest level first, with the
//                           reference's exact expression order (:858-890, :896-940)
//   k_sobel                 edge mask + compaction of flagged pixels   (launchSSAA :554-568)
//   k_ssaa_resolve          mean of the 4 re-traced samples            (SSAAworker :525-536)
//   k_fill_background, k_cover_mark / k_tile_lists / k_fill_tiles
//                           frame pre-fill; projected coverage of the geometry -> the 8x4 tiles primary rays are generated for
//   k_gather_rows / k_scatter_rows / k_quantize_bgr8, k_patch_rows, k_store_words
//                           output stages (incl. saveImage's conversion), early output's rewrite of the re-traced pixels in
//                           a pinned host buffer, the frame's counters into their host mirror
//   k_count_ac / k_ac_resolve, k_rgb_to_rgba
//                           showAC view, texture upload
//   k_raygen, k_ssaa_gen, k_trace, k_shadow
//                           the LITERAL reference walk (objects.cpp:587-631) over materialised queues: parity of the
//                           reference's work counters (RTB_CREATE_COUNTERS / RTB_CREATE_EXACT_WALK), not the timed path
//
// Colours travel through a "slot" array (3 floats per slot): slots [0, w*h) are the framebuffer,
// the rest is bump-allocated for SSAA samples and for the children of Reflective / Transparent hits.
// Every ray carries the slot its colour must land in, so queue order never affects the image.
#pragma once
#ifndef RTB_HOIST_NODES
#define RTB_HOIST_NODES 1
#endif

#include <cuda_runtime.h>

#include "rt_device.cuh"

namespace rtk {

using namespace rt;

constexpr int kBlock = 128;          // threads per CTA for the ray kernels
constexpr int kStackDepth = 64;      // per-thread traversal stack entries (shared memory)

// Device-side counters.  The host zeroes both structures once at the start of a frame and reads them
// back once at its end: nothing in between needs the host, so a whole frame (both passes, every
// recursion level) is enqueued without a single synchronisation.
struct FrameCtr {
    int interiors;       // Reflective / Transparent records handed out (persist until k_combine)
    int slots;           // child colour slots handed out (offset by the frame's slot base)
    int ssaaPixels;      // pixels flagged by k_sobel
    int overflow;        // OVF_* bits: a queue would have overflowed its capacity -> the host grows it and re-runs
    unsigned int shadowSkipped;   // shadow rays not traced because their result cannot affect the pixel
    int acMax;           // showAC debug view: largest per-pixel box count
    unsigned long long tileCursor[2];   // tile pipeline: next tile of pass 1 / of the SSAA pass
    unsigned long long boxTests, triTests;             // closest-hit rays (counting build)
    unsigned long long boxTestsShadow, triTestsShadow; // shadow rays (counting build)
    // fast path's own work (RTB_CREATE_WALK_STATS): search-BVH nodes fetched, triangles tested, eligibility evaluations
    unsigned long long walkNodes[2], walkTris[2], walkEligibility[2];   // [0] closest-hit rays, [1] shadow rays
};
enum { OVF_RAYS = 1, OVF_INTERIORS = 2, OVF_FLAGGED = 4 };

// One per (pass, recursion level).
struct LevelCtr {
    int nRays;           // rays in this level's queue (written by ray generation / the previous level's k_shade)
    int nSurf;           // compacted hits of this level (k_surface)
    int interiorEnd;     // one past the last interior record of this level (atomicMax by k_shade)
    int pad0;
    unsigned long long cursor[2];   // k_walk work cursors: [0] closest-hit rays, [1] shadow rays
};

// Structure-of-arrays ray queue (one level).
struct RayQueue {
    float4* o;       // origin.xyz
    float4* d;       // direction.xyz
    int* dest;       // colour slot
};

struct HitQueue {
    float4* tuv;     // t, u, v, triangle index (int bits)
    int* obj;        // object index, -1 = miss
};

struct SurfQueue {
    float4* pS;      // P.xyz, specular coefficient
    float4* nO;      // N.xyz, object index (int bits)
    float4* cR;      // colour.xyz, ray index (int bits)
};

// Reflective / Transparent hit waiting for its children (castRay's stack frame).
struct Interior {
    int dest;        // slot of this ray's own colour
    int child;       // first child slot: [child] refraction (or the single reflection), [child+1] reflection
    int kind;        // 1 = Reflective, 2 = Transparent with refraction, 3 = Transparent, total internal reflection
    float kr;
    float sx, sy, sz;   // specular light sum
};

__device__ __forceinline__ void storeSlot(float* slots, int s, V3 c)
{
    float* p = slots + (size_t)s * 3;
    p[0] = c.x; p[1] = c.y; p[2] = c.z;
}
__device__ __forceinline__ V3 loadSlot(const float* slots, int s)
{
    const float* p = slots + (size_t)s * 3;
    return mk(p[0], p[1], p[2]);
}

// Where colours go.  `dest` >= 0 names a slot of the frame-wide array `g` (slots [0, w*h) are the framebuffer);
// `dest` <= -2 names slot (-2 - dest) of the tile-local array `l` (SSAA samples and children of Reflective /
// Transparent hits of the tile pipeline, rtb_tile.cuh).  -1 marks a padding queue entry.
struct Slots {
    float* g;
    float* l;
};
__device__ __forceinline__ int localDest(int slot) { return -2 - slot; }
__device__ __forceinline__ float* slotAddr(const Slots& s, int dest)
{
    return dest >= 0 ? s.g + (size_t)dest * 3 : s.l + (size_t)(-2 - dest) * 3;
}
__device__ __forceinline__ void storeSlot(const Slots& s, int dest, V3 c)
{
    float* p = slotAddr(s, dest);
    p[0] = c.x; p[1] = c.y; p[2] = c.z;
}
__device__ __forceinline__ V3 loadSlot(const Slots& s, int dest)
{
    const float* p = slotAddr(s, dest);
    return mk(p[0], p[1], p[2]);
}

// warp-aggregated bump allocation: one atomicAdd per warp for `n` items per participating lane
__device__ __forceinline__ int warpAlloc(int* counter, bool want, int n)
{
    const unsigned mask = __activemask();
    const unsigned votes = __ballot_sync(mask, want);
    if (votes == 0) return -1;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(votes) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(votes) * n);
    base = __shfl_sync(mask, base, leader);
    return want ? base + __popc(votes & ((1u << lane) - 1)) * n : -1;
}

// ------------------------------------------------------------------------------------------------
// ray generation
// ------------------------------------------------------------------------------------------------
// rows[] lists the image rows to render (each < height-1); every row has width-1 pixels because the
// reference never renders the last column / row (getTiles, scene.cpp:369-372).  Queue order is 8x4
// pixel tiles (one warp = one tile) so neighbouring lanes traverse the same BVH nodes; lanes that
// fall outside the image are padding and carry dest = -1 (skipped by every later stage).
__host__ __device__ inline long long raygenPaddedCount(int width, int nRows)
{
    const long long tilesX = (width - 1 + 7) / 8, tilesY = (nRows + 3) / 4;
    return tilesX * tilesY * 32;
}
__global__ void k_raygen(Scene sc, const int* __restrict__ rows, int nRows, RayQueue q, LevelCtr* lv)
{
    const int wm1 = sc.width - 1;
    const int tilesX = (wm1 + 7) / 8;
    const long long total = raygenPaddedCount(sc.width, nRows);
    if (blockIdx.x == 0 && threadIdx.x == 0) lv->nRays = (int)total;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long tile = i >> 5;
        const int lane = (int)(i & 31);
        const int x = (int)(tile % tilesX) * 8 + (lane & 7);
        const int r = (int)(tile / tilesX) * 4 + (lane >> 3);
        if (x < wm1 && r < nRows) {
            const int y = rows[r];
            const V3 d = cameraDir(sc, (float)x + 0.5f, (float)y + 0.5f);
            q.o[i] = make_float4(sc.camPos.x, sc.camPos.y, sc.camPos.z, 0.0f);
            q.d[i] = make_float4(d.x, d.y, d.z, 0.0f);
            q.dest[i] = y * sc.width + x;
        } else {
            q.dest[i] = -1;
        }
    }
}

// caller-supplied rays (rtb_trace / rtb_cast): 6 floats per ray
__global__ void k_rays_from_user(const float* __restrict__ rays, int n, int destBase, RayQueue q, LevelCtr* lv)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) lv->nRays = n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        q.o[i] = make_float4(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2], 0.0f);
        q.d[i] = make_float4(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5], 0.0f);
        q.dest[i] = destBase + i;
    }
}

// ------------------------------------------------------------------------------------------------
// traversal
// ------------------------------------------------------------------------------------------------
// Walks one mesh's reference tree exactly like intersectAccelStruct (objects.cpp:587-631): every
// node whose box the ray LINE overlaps is visited, left before right, every triangle of every such
// leaf is tested, the first strictly smaller t wins.  ANY: stop at the first t < tLimit (shadow).
template <bool ANY, bool COUNT>
__device__ __forceinline__ bool walkMesh(const Scene& sc, const Mesh& me, const RayCtx& r, int* stack, float tLimit,
    float& tBest, float& uBest, float& vBest, int& triBest, unsigned long long& nBox, unsigned long long& nTri)
{
    if (me.nNodes == 0) return false;
    const bool useAC = sc.flags & FLAG_USE_AC;
    const bool cull = sc.flags & FLAG_CULL;
    bool found = false;
    int sp = 0;
    int node = 0;
    for (;;) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(me.nodes) + 2 * node);
        const float4 b = __ldg(reinterpret_cast<const float4*>(me.nodes) + 2 * node + 1);
        if (COUNT && useAC) nBox++;
        const bool inside = !useAC || lineHitsBox(r, a.x, a.y, a.z, a.w, b.x, b.y);
        if (inside) {
            const int link = __float_as_int(b.z), count = __float_as_int(b.w);
            if (count < 0) {
                stack[sp * kBlock] = link;
                sp++;
                node = node + 1;
                continue;
            }
            const float4* slot = reinterpret_cast<const float4*>(me.slots) + (size_t)link * 3;
            for (int s = 0; s < count; ++s, slot += 3) {
                const float4 p0 = __ldg(slot), p1 = __ldg(slot + 1), p2 = __ldg(slot + 2);
                float t, u, v;
                if (COUNT) nTri++;
                if (hitTriangle(r, mk(p0.x, p0.y, p0.z), mk(p1.x, p1.y, p1.z), mk(p2.x, p2.y, p2.z), cull, t, u, v)) {
                    if (ANY) {
                        if (t < tLimit) { if (!COUNT) return true; found = true; }
                    } else if (t < tBest) {
                        tBest = t; uBest = u; vBest = v; triBest = __float_as_int(p0.w); found = true;
                    }
                }
            }
        }
        if (sp == 0) break;
        sp--;
        node = stack[sp * kBlock];
    }
    return found;
}

enum { MODE_FAST = 0, MODE_EXACT = 1, MODE_COUNT = 2 };

// Render::trace for primary / secondary rays (scene.cpp:724-756)
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_trace(Scene sc, RayQueue q, int cap, HitQueue hits, FrameCtr* ctr, const LevelCtr* lv)
{
    extern __shared__ int stackMem[];
    int* stack = stackMem + threadIdx.x;
    unsigned long long nBox = 0, nTri = 0;
    const int n = min(lv->nRays, cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (q.dest[i] < 0) { hits.obj[i] = -1; continue; }   // padding lane of a ray-generation tile
        const float4 o4 = q.o[i], d4 = q.d[i];
        const RayCtx r = makeRay(mk(o4.x, o4.y, o4.z), mk(d4.x, d4.y, d4.z));
        float tNear = FLT_MAX, uN = -1.0f, vN = -1.0f;
        int objN = -1, triN = -1;
        for (int k = 0; k < sc.nObjects; ++k) {
            const Object& ob = sc.objects[k];
            float t = FLT_MAX, u = 0.0f, v = 0.0f;
            int tri = -1;
            bool ok;
            if (ob.type == OBJ_MESH) {
                if (MODE == MODE_FAST) ok = walkMeshFast<false>(sc, sc.meshes[ob.mesh], r, stack, kBlock, 0.0f, t, u, v, tri);
                else ok = walkMesh<false, MODE == MODE_COUNT>(sc, sc.meshes[ob.mesh], r, stack, 0.0f, t, u, v, tri, nBox, nTri);
            } else if (ob.type == OBJ_SPHERE) ok = hitSphere(r, ob.pos, ob.r2, t);
            else ok = hitPlane(r, ob.pos, ob.normal, t);
            if (ok && t < tNear) { tNear = t; uN = u; vN = v; objN = k; triN = tri; }
        }
        hits.tuv[i] = make_float4(tNear, uN, vN, __int_as_float(triN));
        hits.obj[i] = objN;
    }
    if (MODE == MODE_COUNT) {
        atomicAdd(&ctr->boxTests, nBox);
        atomicAdd(&ctr->triTests, nTri);
    }
}

// ------------------------------------------------------------------------------------------------
// surface stage: castRay between trace() and the light loops (scene.cpp:762-775, 945)
// ------------------------------------------------------------------------------------------------
// Threads `first`, `first + step`, ... of a group of `step` threads (a multiple of 32) share the work; `nSurf` is the
// group's compaction counter (global memory for the frame-wide pipeline, shared memory for a tile).
// MISSES = false compiles the miss branch (skybox lookup) out: rays generated inside the walk resolve their own misses.
// Sort key of a hit for the tile pipeline's optional surface sort (north star: "secondary rays sorted by material / direction"):
// material of the object hit (2 bits) and octant of the incoming ray's direction (3 bits).  Surfaces with one key sit next to
// each other in the compacted queue, so a warp of the shade stage runs one material branch and the children it spawns — the
// next level's queue — stay grouped by material and direction as well.
__device__ __forceinline__ int surfaceSortKey(const Scene& sc, int obj, float4 d)
{
    return (sc.objects[obj].material & 3) * 8 + ((d.x < 0.f) ? 1 : 0) + ((d.y < 0.f) ? 2 : 0) + ((d.z < 0.f) ? 4 : 0);
}

// MISSES = false compiles the miss branch (skybox lookup) out: rays generated inside the walk resolve their own misses.
// bins (tile pipeline, optional): 32 running slot counters, one per surfaceSortKey, pre-loaded with the bins' start offsets:
// a surface takes the next slot of its bin instead of the next slot of the queue (counting sort).
template <bool MISSES = true>
__device__ __forceinline__ void surfaceStage(const Scene& sc, RayQueue q, HitQueue hits, SurfQueue surf, const Slots& slots, int n, int first, int step,
    int* nSurf, int handleMisses, int* bins = nullptr)
{
    const int nPadded = (n + 31) & ~31;
    for (int i = first; i < nPadded; i += step) {
        bool wantSurface = false;
        Surface s;
        int obj = -1;
        // rays generated inside the walk resolve their own misses (handleMisses == 0): only obj is valid for those
        if (i < n) obj = hits.obj[i];
        if (i < n && (obj >= 0 || (MISSES && handleMisses && q.dest[i] != -1))) {
            const float4 d4 = q.d[i];
            const V3 d = mk(d4.x, d4.y, d4.z);
            if (MISSES && obj < 0) {
                storeSlot(slots, q.dest[i], skybox(sc, d));
            } else {
                const float4 o4 = q.o[i], h = hits.tuv[i];
                s = surfaceAt(sc, sc.objects[obj], mk(o4.x, o4.y, o4.z), d, h.x, h.y, h.z, __float_as_int(h.w));
                if (sc.flags & FLAG_SHOW_NORMALS) storeSlot(slots, q.dest[i], s.N / 2.0f + mk(0.5f, 0.5f, 0.5f));
                else wantSurface = true;
            }
        }
        int si;
        if (bins) si = wantSurface ? atomicAdd(&bins[surfaceSortKey(sc, obj, q.d[i])], 1) : -1;
        else si = warpAlloc(nSurf, wantSurface, 1);
        if (wantSurface) {
            surf.pS[si] = make_float4(s.P.x, s.P.y, s.P.z, s.specCoef);
            surf.nO[si] = make_float4(s.N.x, s.N.y, s.N.z, __int_as_float(obj));
            surf.cR[si] = make_float4(s.color.x, s.color.y, s.color.z, __int_as_float(i));
        }
    }
}

__global__ void __launch_bounds__(kBlock) k_surface(Scene sc, RayQueue q, int cap, HitQueue hits, SurfQueue surf,
    float* __restrict__ slots, LevelCtr* lv, int handleMisses)
{
    surfaceStage(sc, q, hits, surf, Slots{ slots, nullptr }, min(lv->nRays, cap), blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x,
        &lv->nSurf, handleMisses);
}

// light sample k of a surface: direction FROM the light (normalised for point / area) and distance
__device__ __forceinline__ bool shadowSample(const Scene& sc, int k, V3 P, V3& L, float& dist)
{
    for (int i = 0; i < sc.nLights; ++i) {
        const Light& li = sc.lights[i];
        const int cnt = (li.type == LIGHT_AREA) ? li.pointCount : 1;
        if (k < cnt) {
            if (li.type == LIGHT_AREA) {
                const float* ap = sc.areaPoints + (size_t)(li.pointOffset + k) * 3;
                L = P - mk(ap[0], ap[1], ap[2]);
                dist = length(L);            // intrShadInfo.tNear = lightDir.length()   (scene.cpp:800)
                L = normalize(L);
            } else {
                V3 I;
                illuminate(li, P, L, I, dist);
            }
            return li.type == LIGHT_AREA;
        }
        k -= cnt;
    }
    return false;
}

// Shadow trace (scene.cpp:787 etc.): Transparent objects cast no shadow (:733); an occluder counts
// only when it is closer than the light (`tNear < intrInfo.tNear`, tNear preloaded by illuminate).
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_shadow(Scene sc, RayQueue q, SurfQueue surf, unsigned char* __restrict__ vis, FrameCtr* ctr,
    const LevelCtr* lv)
{
    extern __shared__ int stackMem[];
    int* stack = stackMem + threadIdx.x;
    unsigned long long nBox = 0, nTri = 0;
    unsigned int nSkipped = 0;
    const int S = sc.shadowRaysPerHit;
    const int nSurf = lv->nSurf;
    const long long total = (long long)nSurf * S;
    for (long long qi = blockIdx.x * (long long)blockDim.x + threadIdx.x; qi < total; qi += (long long)gridDim.x * blockDim.x) {
        // light-major order: a warp works on ONE light sample for 32 neighbouring surfaces
        const int k = (int)(qi / nSurf), si = (int)(qi % nSurf);
        const float4 p4 = surf.pS[si], n4 = surf.nO[si];
        const V3 P = mk(p4.x, p4.y, p4.z), N = mk(n4.x, n4.y, n4.z);
        V3 L; float dist;
        const bool areaSample = shadowSample(sc, k, P, L, dist);
        if (MODE != MODE_COUNT) {
            // Dead-ray elision.  The visibility bit only ever multiplies max(0, N.-L) (Diffuse / Phong,
            // scene.cpp:788,820) and pow(max(0, R.-D), nSpecular) (every material but Diffuse, :824,:867,
            // :917); when those factors are exactly zero the reference's trace() cannot change the pixel,
            // so the ray is not traced.  Bit-identical image; counted in FrameCtr::shadowSkipped.
            const int mat = sc.objects[__float_as_int(n4.w)].material;
            bool needed = false;
            if (mat == MAT_DIFFUSE || mat == MAT_PHONG) needed = maxf_(0.f, dot(N, -L)) > 0.f;
            if (!needed && mat != MAT_DIFFUSE) {
                const float4 d4 = q.d[__float_as_int(surf.cR[si].w)];
                const float base = maxf_(0.f, dot(reflect(L, N), -mk(d4.x, d4.y, d4.z)));
                needed = base > 0.f || (!areaSample && !(sc.objects[__float_as_int(n4.w)].nSpecular > 0.f));
            }
            if (!needed) { vis[(size_t)si * S + k] = 0; nSkipped++; continue; }
        }
        const RayCtx r = makeRay(P + N * sc.bias, -L);
        float tNear = dist;
        bool blocked = false;
        for (int j = 0; j < sc.nObjects; ++j) {
            const Object& ob = sc.objects[j];
            if (ob.material == MAT_TRANSPARENT) continue;
            float t = FLT_MAX, u, v; int tri;
            bool ok;
            if (ob.type == OBJ_MESH) {
                if (MODE == MODE_COUNT) {
                    // full closest-hit walk so the work counters equal the reference's
                    ok = walkMesh<false, true>(sc, sc.meshes[ob.mesh], r, stack, 0.0f, t, u, v, tri, nBox, nTri);
                } else {
                    if (MODE == MODE_FAST) ok = walkMeshFast<true>(sc, sc.meshes[ob.mesh], r, stack, kBlock, tNear, t, u, v, tri);
                    else ok = walkMesh<true, false>(sc, sc.meshes[ob.mesh], r, stack, tNear, t, u, v, tri, nBox, nTri);
                    if (ok) { blocked = true; break; }
                    continue;
                }
            } else if (ob.type == OBJ_SPHERE) ok = hitSphere(r, ob.pos, ob.r2, t);
            else ok = hitPlane(r, ob.pos, ob.normal, t);
            if (ok && t < tNear) { tNear = t; blocked = true; if (MODE != MODE_COUNT) break; }
        }
        vis[(size_t)si * S + k] = blocked ? 0 : 1;
    }
    if (nSkipped) atomicAdd(&ctr->shadowSkipped, nSkipped);
    if (MODE == MODE_COUNT) {
        atomicAdd(&ctr->boxTestsShadow, nBox);
        atomicAdd(&ctr->triTestsShadow, nTri);
    }
}


// ------------------------------------------------------------------------------------------------
// persistent-threads traversal (fast path)
// ------------------------------------------------------------------------------------------------
// One kernel serves both ray kinds: ANY = false is Render::trace for primary / secondary rays
// (closest hit over all objects), ANY = true is the shadow trace (any occluder closer than the
// light, Transparent objects skipped).  The grid is sized to the machine, not to the queue: every
// warp pulls batches of consecutive rays from a global cursor (ballot + one atomicAdd per warp) whenever fewer
// than kRefillBelow of its lanes still hold a live ray (1 = finish the batch first, measured best; see the loop).
// Traversal of the search BVH uses a per-thread stack in shared memory and the decoupled stepping described at the loop.
//
// GEN selects where closest-hit rays come from:
//   GEN_QUEUE    the level's ray queue (secondary rays, caller-supplied rays)
//   GEN_PRIMARY  generated in place from the pixel grid, 8x4 tiles (renderWorker + Camera::getRay); no queue read
//   GEN_SSAA     generated in place from the flagged-pixel list, 4 samples per pixel (SSAAworker)
// Generated rays that miss everything are finished right here (skybox / background colour into the ray's slot);
// only hits are written to the queue for the surface and shade stages.
constexpr int kDone = (int)0x80000000;   // cursor value: no mesh traversal in progress
#ifndef RTB_REFILL_BELOW
#define RTB_REFILL_BELOW 1
#endif
constexpr int kRefillBelow = RTB_REFILL_BELOW;
#ifndef RTB_LEAF_BATCH
#define RTB_LEAF_BATCH 4
#endif
#ifndef RTB_INNER_STEPS
#define RTB_INNER_STEPS 8
#endif
constexpr int kLeafBatch = RTB_LEAF_BATCH, kInnerSteps = RTB_INNER_STEPS;

enum { GEN_QUEUE = 0, GEN_PRIMARY = 1, GEN_SSAA = 2 };

struct GenArgs {
    const int* list;     // GEN_PRIMARY: image rows to render; GEN_SSAA: flagged pixels
    int count;           // GEN_PRIMARY: number of rows; GEN_SSAA: capacity of the flagged list
    int slotBase;        // GEN_SSAA: first sample slot of the frame-wide slot array; -1 = samples go to tile-local slots
    int x0, cols;        // GEN_PRIMARY: pixel columns [x0, x0 + cols) to generate rays for (the screen-space bounds of the
                         // geometry; pixels outside were pre-filled with the background colour by k_fill_background)
    float* slots;        // frame-wide colour slots (misses of generated rays are resolved here)
    const Scene* scene;  // device-resident copy of the scene header for the out-of-line helpers below
    const int* tiles;    // GEN_PRIMARY: the 8x4 pixel tiles (index into the cols x rows tile grid) rays are generated for, or nullptr
    const int* nTiles;   //              = all of them; the list and its length live on the device (k_tile_lists): tiles outside the
                         //              geometry's projected coverage get the background colour from k_fill_tiles instead of rays
};
__device__ __forceinline__ long long primaryRayCount(const GenArgs& g)
{
    return g.tiles ? 32LL * *g.nTiles : raygenPaddedCount(g.cols + 1, g.count);
}

// Kept out of line so their FP64 normalisation / cube-map code does not raise the register count of the
// traversal loop (they run once per ray, the loop runs ~100 times).
__device__ __noinline__ V3 genCameraRay(const Scene* sc, float px, float py) { return cameraDir(*sc, px, py); }
__device__ __noinline__ V3 missColour(const Scene* sc, V3 d) { return skybox(*sc, d); }

#ifndef RTB_WALK_MIN_BLOCKS
#define RTB_WALK_MIN_BLOCKS 9   // 56 registers, 9 CTAs per SM: measured 1-2 % faster than 64 / 78 registers at 8 / 6 CTAs
#endif

// Part of one mesh's search BVH resident in shared memory (rtb_tile.cuh stages it with TMA bulk copies): nodes
// [0, nNodes) — the tree's top levels, or all of it — and, when the whole mesh fits, every leaf triangle.
struct StagedBvh {
    const Mesh* mesh;        // the mesh the copy belongs to (nullptr = nothing staged)
    unsigned nodesAddr;      // shared-memory byte addresses
    unsigned trisAddr;
    int nNodes;
    int nTris;               // 0 = triangles stay in global memory
};
__device__ __forceinline__ float4 ldsF4(unsigned addr)
{
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// per-thread tallies of a walk, flushed by the caller (once per kernel)
struct WalkAcc {
    unsigned nSkipped = 0;                                 // shadow rays elided
    unsigned long long nNodes[2] = { 0, 0 }, nTris[2] = { 0, 0 }, nElig[2] = { 0, 0 };   // STATS only; [0] closest-hit rays, [1] shadow rays
};

// The traversal loop.  All 32 lanes of a warp must call it together; warps are otherwise independent and pull rays
// [0, total) from `cursor` (global memory for a frame-wide launch, shared memory for one tile's stage).  Indices are
// relative to the queues handed in; generated rays map index `my` to pixel / sample number genOffset + my.
// STRIDE = distance in ints between a thread's consecutive stack entries (the thread count of the stack's owner).
// ANY and GEN are RUN-TIME, warp-uniform arguments on purpose: a fused kernel walks closest-hit rays, shadow rays and
// (deep scenes) queued secondary rays, and calls this function from ONE site in a small loop, so one copy of the loop
// (~25 KB of SASS) serves all of them.  As template parameters every kind had its own inlined copy and the kernels grew
// past the instruction cache (ncu: no-instruction stalls 1.1 per issue in the 120 KB k_tile; profiles/r2_experiments).
template <bool STATS, int STRIDE, typename CursorT, bool STAGED = false>
__device__ __forceinline__ void walkRays(const bool ANY, const int GEN, const Scene& sc, RayQueue q, HitQueue hits, SurfQueue surf, unsigned char* vis,
    int nSurfRaw, const GenArgs& gen, long long genOffset, const Slots& slots, CursorT* cursor, long long total, int* stack, WalkAcc& acc,
    const StagedBvh& sb = StagedBvh{})
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lanesBelow = (1u << lane) - 1;
    const bool cull = sc.flags & FLAG_CULL;
    const int S = sc.shadowRaysPerHit;
    const int nSurf = nSurfRaw > 0 ? nSurfRaw : 1;
    const int tilesX = (gen.cols + 7) / 8;

    bool have = false, exhausted = false;
    long long out = 0;                 // where the result goes: ray index (closest) or visibility index (ANY)
    int dest = -1;                     // colour slot of a generated ray
    RayCtx r = makeRay(mk(0.f, 0.f, 0.f), mk(0.f, 0.f, -1.f));
    float tNear = FLT_MAX, uN = -1.f, vN = -1.f;   // best over all objects / light distance
    int objN = -1, triN = -1;
    int obj = 0;                       // object loop position
    int cur = kDone, sp = 0;           // search-BVH cursor of the mesh being walked
    // The node array of the mesh being walked stays in registers: reading it through the Mesh record put a dependent load in
    // front of every node fetch.  The record itself is only needed at leaves and is looked up again there (RTB_HOIST_NODES=0:
    // the old form, kept for A/B).
#if RTB_HOIST_NODES
    const float4* nodesP = nullptr;
#else
    const Mesh* me = nullptr;
#endif
    bool found = false;
    int slotBest = 0x7fffffff, triM = -1;
    float tM = FLT_MAX, uM = 0.f, vM = 0.f;

    for (;;) {
        // ---- refill idle lanes from the cursor ----
        // One atomicAdd per refill hands every idle lane the next consecutive ray.  kRefillBelow = 1: a warp takes
        // 32 consecutive rays and finishes them all before fetching again.  Measured on B200 (cfg4 / dragon frame ms):
        // refilling when fewer than 22 / 16 / 10 / 6 / 1 lanes are live gives 0.677 / 0.661 / 0.655 / 0.565 / 0.565 and
        // 2.35 / 2.27 / 2.23 / 2.08 / 2.07; keeping every lane busy from a per-warp ring of prepared rays is slower still
        // (0.67 / 2.74).  Rays that start together stay in phase (same tree levels, same leaves at the same time);
        // a lane refilled mid-flight starts at the root while its neighbours are at leaves, and the warp pays for both.
        const unsigned idle = __ballot_sync(FULL, !have);
        if (idle && !exhausted) {
            const int nIdle = __popc(idle), leader = __ffs(idle) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = (unsigned long long)atomicAdd(cursor, (CursorT)nIdle);
            base = __shfl_sync(FULL, base, leader);
            if ((long long)base + nIdle >= total) exhausted = true;
            const long long my = (long long)base + __popc(idle & lanesBelow);
            const bool take = !have && my < total;
            if (take) {
                if (!ANY) {
                    bool live = true;
                    V3 o, d;
                    if (GEN == GEN_PRIMARY) {
                        const long long gmy = genOffset + my;
                        const long long tile = gen.tiles ? (long long)__ldg(gen.tiles + (gmy >> 5)) : (gmy >> 5);
                        const int xr = (int)(tile % tilesX) * 8 + ((int)gmy & 7);
                        const int x = gen.x0 + xr;
                        const int rr = (int)(tile / tilesX) * 4 + (((int)gmy >> 3) & 3);
                        live = xr < gen.cols && rr < gen.count;
                        if (live) {
                            const int y = __ldg(gen.list + rr);
                            o = sc.camPos;
                            d = genCameraRay(gen.scene, (float)x + 0.5f, (float)y + 0.5f);
                            dest = y * sc.width + x;
                        }
                    } else if (GEN == GEN_SSAA) {
                        const long long gmy = genOffset + my;
                        const int pix = gen.list[gmy >> 2], k = (int)gmy & 3;
                        const int y = pix / sc.width, x = pix - y * sc.width;
                        // sample order (.25,.25) (.25,.75) (.75,.25) (.75,.75)  (scene.cpp:527-534)
                        const float ox = (k & 2) ? 0.75f : 0.25f, oy = (k & 1) ? 0.75f : 0.25f;
                        o = sc.camPos;
                        d = genCameraRay(gen.scene, (float)x + ox, (float)y + oy);
                        dest = gen.slotBase < 0 ? localDest((int)my) : gen.slotBase + (int)gmy;
                    } else {
                        live = q.dest[my] != -1;     // padding entry
                        if (live) {
                            const float4 o4 = q.o[my], d4 = q.d[my];
                            o = mk(o4.x, o4.y, o4.z); d = mk(d4.x, d4.y, d4.z);
                        }
                    }
                    if (!live) {
                        hits.obj[my] = -1;
                    } else {
                        r = makeRay(o, d);
                        tNear = FLT_MAX; uN = -1.f; vN = -1.f; objN = -1; triN = -1;
                        out = my; obj = 0; cur = kDone; have = true;
                    }
                } else {
                    // light-major order: neighbouring lanes = neighbouring surfaces, same light sample
                    const int k = (int)(my / nSurf), si = (int)(my % nSurf);
                    const float4 p4 = surf.pS[si], n4 = surf.nO[si];
                    const V3 P = mk(p4.x, p4.y, p4.z), N = mk(n4.x, n4.y, n4.z);
                    V3 L; float dist;
                    const bool areaSample = shadowSample(sc, k, P, L, dist);
                    // dead-ray elision (see k_shadow): the bit cannot influence the pixel
                    const Object& ob = sc.objects[__float_as_int(n4.w)];
                    const int mat = ob.material;
                    bool needed = false;
                    if (mat == MAT_DIFFUSE || mat == MAT_PHONG) needed = maxf_(0.f, dot(N, -L)) > 0.f;
                    if (!needed && mat != MAT_DIFFUSE) {
                        const float4 d4 = q.d[__float_as_int(surf.cR[si].w)];
                        const float base2 = maxf_(0.f, dot(reflect(L, N), -mk(d4.x, d4.y, d4.z)));
                        needed = base2 > 0.f || (!areaSample && !(ob.nSpecular > 0.f));
                    }
                    out = (long long)si * S + k;
                    if (!needed) { vis[out] = 0; acc.nSkipped++; }
                    else {
                        r = makeRay(P + N * sc.bias, -L);
                        tNear = dist; obj = 0; cur = kDone; have = true;
                    }
                }
            }
        }
        if (__ballot_sync(FULL, have) == 0) {
            if (exhausted) break;
            continue;
        }
        // ---- traverse until the warp runs low on live rays ----
        for (;;) {
            if (have) {
                if (cur == kDone) {
                    // object loop of Render::trace (scene.cpp:731-754)
                    if (obj >= sc.nObjects) {
                        if (ANY) vis[out] = 1;
                        else if (GEN == GEN_QUEUE) {
                            hits.tuv[out] = make_float4(tNear, uN, vN, __int_as_float(triN)); hits.obj[out] = objN;
                        } else {
                            hits.obj[out] = objN;
                            if (objN < 0) {
                                storeSlot(slots, dest, missColour(gen.scene, r.d));        // castRay's miss branch (scene.cpp:945)
                            } else {
                                hits.tuv[out] = make_float4(tNear, uN, vN, __int_as_float(triN));
                                q.o[out] = make_float4(r.o.x, r.o.y, r.o.z, 0.0f);
                                q.d[out] = make_float4(r.d.x, r.d.y, r.d.z, 0.0f);
                                q.dest[out] = dest;
                            }
                        }
                        have = false;
                    } else {
                        const Object& ob = sc.objects[obj];
                        if (ANY && ob.material == MAT_TRANSPARENT) {
                            obj++;
                        } else if (ob.type == OBJ_MESH) {
#if RTB_HOIST_NODES
                            const Mesh* me = &sc.meshes[ob.mesh];
#else
                            me = &sc.meshes[ob.mesh];
#endif
                            if (me->nNodes == 0 || me->nTris == 0) obj++;   // missing .obj / no usable face: nothing to hit
                            else {
                                cur = 0; sp = 0; found = false; slotBest = 0x7fffffff; tM = tNear;
#if RTB_HOIST_NODES
                                nodesP = me->bvhNodes;
#endif
                            }
                        } else {
                            float t = FLT_MAX;
                            const bool ok = (ob.type == OBJ_SPHERE) ? hitSphere(r, ob.pos, ob.r2, t) : hitPlane(r, ob.pos, ob.normal, t);
                            if (ok && t < tNear) {
                                if (ANY) { vis[out] = 0; have = false; }
                                else { tNear = t; uN = 0.f; vN = 0.f; objN = obj; triN = -1; }
                            }
                            obj++;
                        }
                    }
                }
            }
            // Decoupled stepping.  A classic while-while loop lets every lane descend until ALL lanes of the warp have
            // reached a leaf, so one long descent stalls 31 parked lanes, and then runs the triangle code for whoever
            // has a leaf.  Here a round advances the lanes at inner nodes by at most kInnerSteps nodes, and lanes that
            // reached a leaf park until kLeafBatch of them wait (or no lane can descend): nobody waits for a whole
            // descent, and the triangle code runs for a group of lanes.  Measured on B200 against the while-while loop:
            // dragon frame 2.07 -> 1.43 ms, cfg4 0.565 -> 0.52 ms (kInnerSteps 8, kLeafBatch 1..4 equivalent; 1 inner
            // step per round or batches of 16+ leaves are 10-15 % slower).
            {
                const bool atInner = have && cur >= 0;
                const bool atLeaf = have && cur < 0 && cur != kDone;
                const unsigned innerM = __ballot_sync(FULL, atInner), leafM = __ballot_sync(FULL, atLeaf);
                const bool runInner = innerM != 0 && __popc(leafM) < kLeafBatch;
                if (runInner && atInner) {
#if RTB_HOIST_NODES
                    const Mesh* me = STAGED ? &sc.meshes[sc.objects[obj].mesh] : nullptr;
#endif
                    for (int step = 0; step < kInnerSteps && cur >= 0; ++step) {
                        float4 a, b, c, d;
                        if (STAGED && me == sb.mesh && cur < sb.nNodes) {
                            const unsigned na = sb.nodesAddr + (unsigned)cur * 64u;
                            a = ldsF4(na); b = ldsF4(na + 16); c = ldsF4(na + 32); d = ldsF4(na + 48);
                        } else {
#if RTB_HOIST_NODES
                            const float4* nd = nodesP + (size_t)cur * 4;
#else
                            const float4* nd = me->bvhNodes + (size_t)cur * 4;
#endif
                            a = __ldg(nd); b = __ldg(nd + 1); c = __ldg(nd + 2); d = __ldg(nd + 3);
                        }
                        if (STATS) acc.nNodes[ANY]++;
                        const float tFar = ANY ? tNear : tM;
                        bool h0, h1;
                        const float e0 = slabEntry(r, a.x, a.y, a.z, a.w, b.x, b.y, tFar, h0);
                        const float e1 = slabEntry(r, b.z, b.w, c.x, c.y, c.z, c.w, tFar, h1);
                        const int c0 = __float_as_int(d.x), c1 = __float_as_int(d.y);
                        if (h0 && h1) {
                            const bool swap = e1 < e0;
                            stack[sp * STRIDE] = swap ? c0 : c1;
                            sp++;
                            cur = swap ? c1 : c0;
                        } else if (h0) cur = c0;
                        else if (h1) cur = c1;
                        else if (sp > 0) { sp--; cur = stack[sp * STRIDE]; }
                        else { cur = kDone; break; }
                    }
                    if (cur == kDone) {   // mesh finished
                        if (!ANY && found) { tNear = tM; uN = uM; vN = vM; objN = obj; triN = triM; }
                        obj++;
                    }
                } else if (!runInner && atLeaf) {
                    const int code = ~cur;
                    const int first = code >> 3, count = (code & 7) + 1;
#if RTB_HOIST_NODES
                    const Mesh* me = &sc.meshes[sc.objects[obj].mesh];
#endif
                    const float4* tp = me->bvhTris + (size_t)first * 3;
                    const bool trisStaged = STAGED && me == sb.mesh && sb.nTris > 0;
                    unsigned ta = sb.trisAddr + (unsigned)first * 48u;
                    bool blocked = false;
                    for (int k = 0; k < count; ++k, tp += 3, ta += 48u) {
                        float4 p0, p1, p2;
                        if (trisStaged) { p0 = ldsF4(ta); p1 = ldsF4(ta + 16); p2 = ldsF4(ta + 32); }
                        else { p0 = __ldg(tp); p1 = __ldg(tp + 1); p2 = __ldg(tp + 2); }
                        float t, u, v;
                        if (STATS) acc.nTris[ANY]++;
                        if (!hitTriangle(r, mk(p0.x, p0.y, p0.z), mk(p1.x, p1.y, p1.z), mk(p2.x, p2.y, p2.z), cull, t, u, v)) continue;
                        const int tri = __float_as_int(p0.w);
                        if (ANY) {
                            if (STATS && t < tNear) acc.nElig[ANY]++;
                            if (t < tNear && eligibleSlot(sc, *me, r, tri) >= 0) { blocked = true; break; }
                        } else if (t < tM || (found && t == tM)) {
                            if (STATS) acc.nElig[ANY]++;
                            const int slot = eligibleSlot(sc, *me, r, tri);
                            if (slot >= 0 && (t < tM || slot < slotBest)) { tM = t; uM = u; vM = v; triM = tri; slotBest = slot; found = true; }
                        }
                    }
                    if (ANY && blocked) { vis[out] = 0; have = false; cur = kDone; }
                    else if (sp > 0) { sp--; cur = stack[sp * STRIDE]; }
                    else {
                        cur = kDone;
                        if (!ANY && found) { tNear = tM; uN = uM; vN = vM; objN = obj; triN = triM; }
                        obj++;
                    }
                }
            }
            const unsigned act = __ballot_sync(FULL, have);
            if (act == 0 || (!exhausted && __popc(act) < kRefillBelow)) break;
        }
    }
}

__device__ __forceinline__ void flushWalkAcc(FrameCtr* ctr, const WalkAcc& acc, bool stats)
{
    // one atomic per warp, not per thread
    unsigned sk = acc.nSkipped;
    for (int o = 16; o > 0; o >>= 1) sk += __shfl_xor_sync(0xffffffffu, sk, o);
    if ((threadIdx.x & 31) == 0 && sk) atomicAdd(&ctr->shadowSkipped, sk);
    (void)stats;
}

// The frame-wide form: one launch walks one level's whole queue (the grid is sized to the machine, every warp pulls
// batches from the level's global cursor).  The default pipeline runs the same loop per tile inside k_tile (rtb_tile.cuh).
template <bool ANY, int GEN, bool STATS = false>
__global__ void __launch_bounds__(kBlock, RTB_WALK_MIN_BLOCKS) k_walk(Scene sc, RayQueue q, int cap, HitQueue hits, SurfQueue surf,
    unsigned char* __restrict__ vis, FrameCtr* ctr, LevelCtr* lv, GenArgs gen)
{
    extern __shared__ int stackMem[];
    int* stack = stackMem + threadIdx.x;
    const int S = sc.shadowRaysPerHit;
    unsigned long long* cursor = &lv->cursor[ANY ? 1 : 0];
    const int nSurfRaw = ANY ? lv->nSurf : 0;
    long long total;
    if (ANY) total = (long long)nSurfRaw * S;
    else if (GEN == GEN_PRIMARY) total = primaryRayCount(gen);
    else if (GEN == GEN_SSAA) total = 4LL * min(ctr->ssaaPixels, gen.count);
    else total = min(lv->nRays, cap);
    if (!ANY && GEN != GEN_QUEUE && blockIdx.x == 0 && threadIdx.x == 0) lv->nRays = (int)total;
    WalkAcc acc;
    walkRays<STATS, kBlock>(ANY, GEN, sc, q, hits, surf, vis, nSurfRaw, gen, 0LL, Slots{ gen.slots, nullptr }, cursor, total, stack, acc);
    if (ANY) flushWalkAcc(ctr, acc, STATS);
    if (STATS) {
        atomicAdd(&ctr->walkNodes[ANY ? 1 : 0], acc.nNodes[ANY]);
        atomicAdd(&ctr->walkTris[ANY ? 1 : 0], acc.nTris[ANY]);
        atomicAdd(&ctr->walkEligibility[ANY ? 1 : 0], acc.nElig[ANY]);
    }
}

// ------------------------------------------------------------------------------------------------
// shade stage: the four material branches of castRay (scene.cpp:780-941)
// ------------------------------------------------------------------------------------------------
#ifndef RTB_SHADE_MIN_BLOCKS
#define RTB_SHADE_MIN_BLOCKS 8
#endif

// Where a shade stage puts what it spawns.  Frame-wide pipeline: global counters, child colours in the frame-wide slot
// array from slotBase on (slotCount hands them out).  Tile pipeline: the group's shared-memory counters, child colours
// in the tile-local slot array (record i owns local slots slotBase + 2 i, slotCount == nullptr).
struct ShadeOut {
    RayQueue next; int nextCap; int* nextCount;
    Interior* interiors; int interiorCap; int* interiorCount;
    int* interiorEnd;        // frame-wide only: atomicMax of (record index + 1) per level
    int* slotCount; int slotBase, slotCap;
    bool localSlots;
    int* overflow;
};

// DEEP = false compiles the Reflective / Transparent branches and the child bookkeeping out (no such object in the scene).
template <bool DEEP = true>
__device__ __forceinline__ void shadeStage(const Scene& sc, RayQueue q, SurfQueue surf, const unsigned char* vis, int depth, int n, int first, int step,
    const Slots& slots, const ShadeOut& o)
{
    const int S = sc.shadowRaysPerHit;
    const int nPadded = (n + 31) & ~31;
    for (int si = first; si < nPadded; si += step) {
        int nChildren = 0;
        bool wantInterior = false;
        Interior rec;
        V3 childO[2], childD[2];
        if (si < n) {
            const float4 p4 = surf.pS[si], n4 = surf.nO[si], c4 = surf.cR[si];
            const V3 P = mk(p4.x, p4.y, p4.z), N = mk(n4.x, n4.y, n4.z), color = mk(c4.x, c4.y, c4.z);
            const Object& ob = sc.objects[__float_as_int(n4.w)];
            const int ri = __float_as_int(c4.w);
            const float4 d4 = q.d[ri];
            const V3 dir = mk(d4.x, d4.y, d4.z);
            const int dest = q.dest[ri];
            const int mat = ob.material;

            V3 diff = mk(0.0f, 0.0f, 0.0f), spec = mk(0.0f, 0.0f, 0.0f);
            int k = 0;
            for (int i = 0; i < sc.nLights; ++i) {
                const Light& li = sc.lights[i];
                if (li.type != LIGHT_AREA) {
                    V3 L, I; float dist;
                    illuminate(li, P, L, I, dist);
                    const float v = vis[(size_t)si * S + k] ? 1.0f : 0.0f;
                    k++;
                    if (mat == MAT_DIFFUSE) {
                        diff = diff + I * (v * maxf_(0.f, dot(N, -L)));                       // :788
                    } else {
                        if (mat == MAT_PHONG) diff = diff + (I * v) * maxf_(0.f, dot(N, -L)); // :820
                        const V3 R = reflect(L, N);
                        spec = spec + (I * v) * powExact(maxf_(0.f, dot(R, -dir)), ob.nSpecular);   // :824, :867, :917
                    }
                } else {
                    const V3 I = areaIntensity(li, P);
                    float dsum = 0.0f, ssum = 0.0f;
                    for (int p = 0; p < li.pointCount; ++p, ++k) {
                        const float* ap = sc.areaPoints + (size_t)(li.pointOffset + p) * 3;
                        const V3 L = normalize(P - mk(ap[0], ap[1], ap[2]));
                        const float v = vis[(size_t)si * S + k] ? 1.0f : 0.0f;
                        dsum += v * maxf_(0.f, dot(N, -L));
                        const V3 R = reflect(L, N);
                        ssum += v * maxf_(0.f, dot(R, -dir));
                    }
                    const float cnt = (float)li.pointCount;
                    if (mat == MAT_DIFFUSE || mat == MAT_PHONG) diff = diff + I * (dsum / cnt);      // :806, :844
                    if (mat != MAT_DIFFUSE) spec = spec + I * powExact(ssum / cnt, ob.nSpecular);     // :846, :887, :937
                }
            }

            if (mat == MAT_DIFFUSE) {
                storeSlot(slots, dest, color * diff);                                                 // :809
            } else if (mat == MAT_PHONG || !DEEP) {
                storeSlot(slots, dest, color * ob.ambient + diff * ob.diffuse + spec * p4.w);         // :852
            } else if (mat == MAT_REFLECTIVE) {
                const V3 ro = P + N * sc.bias;
                const V3 rd = dir - N * (2 * dot(dir, N));                                           // :856
                if (depth + 1 > sc.maxRayDepth) {
                    storeSlot(slots, dest, skybox(sc, rd) * 0.8f + spec);                             // :760, :858, :890
                } else {
                    wantInterior = true; nChildren = 1;
                    rec.dest = dest; rec.kind = 1; rec.kr = 0.0f; rec.sx = spec.x; rec.sy = spec.y; rec.sz = spec.z;
                    childO[0] = ro; childD[0] = rd;
                }
            } else {
                const float kr = fresnel(dir, N, ob.ior);                                             // :893
                const bool outside = dot(dir, N) < 0;
                const V3 biasVec = N * sc.bias;
                const bool refr = kr < 1;
                V3 rdir = mk(0.0f, 0.0f, 0.0f), rorig = P;
                if (refr) {
                    rdir = normalize(refract(dir, N, ob.ior));                                        // :899
                    rorig = outside ? P - biasVec : P + biasVec;
                }
                const V3 fdir = normalize(reflect(dir, N));                                           // :905
                const V3 forig = outside ? P + biasVec : P - biasVec;
                if (depth + 1 > sc.maxRayDepth) {
                    V3 c = mk(0.0f, 0.0f, 0.0f);
                    if (refr) c = c + skybox(sc, rdir) * (1 - kr);
                    c = c + skybox(sc, fdir) * kr;
                    c = c + spec * kr;
                    storeSlot(slots, dest, c);
                } else {
                    wantInterior = true;
                    rec.dest = dest; rec.kr = kr; rec.sx = spec.x; rec.sy = spec.y; rec.sz = spec.z;
                    if (refr) { rec.kind = 2; nChildren = 2; childO[0] = rorig; childD[0] = rdir; childO[1] = forig; childD[1] = fdir; }
                    else { rec.kind = 3; nChildren = 1; childO[0] = forig; childD[0] = fdir; }
                }
            }
        }
        if (!DEEP) continue;
        // bump-allocate: interior record, two colour slots, nChildren queue entries
        const int ii = warpAlloc(o.interiorCount, wantInterior, 1);
        const int cs = o.slotBase + (o.slotCount ? warpAlloc(o.slotCount, wantInterior, 2) : 2 * ii);
        const int q1 = warpAlloc(o.nextCount, nChildren >= 1, 1);
        const int q2 = warpAlloc(o.nextCount, nChildren >= 2, 1);
        if (wantInterior) {
            const bool fits1 = q1 < o.nextCap, fits2 = nChildren < 2 || q2 < o.nextCap;
            if (ii >= o.interiorCap || cs + 2 > o.slotCap) {
                atomicOr(o.overflow, OVF_INTERIORS);
                // queue entries already claimed must not stay uninitialised: mark them as padding
                if (fits1) o.next.dest[q1] = -1;
                if (nChildren == 2 && fits2) o.next.dest[q2] = -1;
            } else {
                if (!fits1 || !fits2) atomicOr(o.overflow, OVF_RAYS);   // the host re-runs the frame with larger queues
                rec.child = cs;
                o.interiors[ii] = rec;
                if (o.interiorEnd) atomicMax(o.interiorEnd, ii + 1);
                // kind 3 keeps its single (reflection) child in slot child+1 so the combine stage reads one layout
                const int firstSlot = (rec.kind == 3) ? cs + 1 : cs;
                const int d1 = o.localSlots ? localDest(firstSlot) : firstSlot, d2 = o.localSlots ? localDest(cs + 1) : cs + 1;
                if (fits1) {
                    o.next.o[q1] = make_float4(childO[0].x, childO[0].y, childO[0].z, 0.0f);
                    o.next.d[q1] = make_float4(childD[0].x, childD[0].y, childD[0].z, 0.0f);
                    o.next.dest[q1] = d1;
                } else storeSlot(slots, d1, mk(0.0f, 0.0f, 0.0f));
                if (nChildren == 2) {
                    if (fits2) {
                        o.next.o[q2] = make_float4(childO[1].x, childO[1].y, childO[1].z, 0.0f);
                        o.next.d[q2] = make_float4(childD[1].x, childD[1].y, childD[1].z, 0.0f);
                        o.next.dest[q2] = d2;
                    } else storeSlot(slots, d2, mk(0.0f, 0.0f, 0.0f));
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kBlock, RTB_SHADE_MIN_BLOCKS) k_shade(Scene sc, RayQueue q, SurfQueue surf, const unsigned char* __restrict__ vis,
    int depth, RayQueue next, int nextCap, Interior* __restrict__ interiors, int interiorCap,
    float* __restrict__ slots, int slotBase, int slotCap, FrameCtr* ctr, LevelCtr* lv)
{
    const ShadeOut o{ next, nextCap, &(lv + 1)->nRays, interiors, interiorCap, &ctr->interiors, &lv->interiorEnd, &ctr->slots, slotBase, slotCap,
        false, &ctr->overflow };
    shadeStage(sc, q, surf, vis, depth, lv->nSurf, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, Slots{ slots, nullptr }, o);
}

// Fold children into parents for interior records [first, last): castRay's return path.  rec.child is a slot INDEX of
// the frame-wide array, or of the tile-local array when localChildren.
__device__ __forceinline__ void combineStage(const Interior* interiors, int first, int last, int tid, int step, const Slots& slots, bool localChildren)
{
    for (int i = first + tid; i < last; i += step) {
        const Interior rec = interiors[i];
        const V3 spec = mk(rec.sx, rec.sy, rec.sz);
        const int c0 = localChildren ? localDest(rec.child) : rec.child, c1 = localChildren ? localDest(rec.child + 1) : rec.child + 1;
        V3 c;
        if (rec.kind == 1) {
            c = loadSlot(slots, c0) * 0.8f;                        // hitColor = 0.8f * castRay(...)   (:858)
            c = c + spec;                                          // hitColor += specularComponent    (:890)
        } else {
            c = mk(0.0f, 0.0f, 0.0f);                              // hitColor = { 0 }                 (:896)
            if (rec.kind == 2) c = c + loadSlot(slots, c0) * (1 - rec.kr);   // :902
            c = c + loadSlot(slots, c1) * rec.kr;                  // :908
            c = c + spec * rec.kr;                                 // :940
        }
        storeSlot(slots, rec.dest, c);
    }
}

// Level `level`'s records are [max(interiorEnd of the shallower levels), interiorEnd[level]): levels run one
// after the other and records are bump-allocated, so each level owns one contiguous range.
__global__ void k_combine(const Interior* __restrict__ interiors, const LevelCtr* __restrict__ lv, int level, int interiorCap,
    float* __restrict__ slots)
{
    int first = 0;
    for (int j = 0; j < level; ++j) first = max(first, lv[j].interiorEnd);
    const int last = min(max(first, lv[level].interiorEnd), interiorCap);
    combineStage(interiors, first, last, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, Slots{ slots, nullptr }, false);
}

// ------------------------------------------------------------------------------------------------
// SSAA: Sobel mask on the unclamped float frame, then 4 re-traced samples per flagged pixel
// ------------------------------------------------------------------------------------------------
// rows[]: image rows owned by this call; only interior pixels get a flag (scene.cpp:554-555).  Pixels are
// visited in the 8x4 tiles of ray generation, so the compacted list keeps neighbouring pixels — and with them the 4
// samples of each — next to each other in the SSAA ray queue.
// One warp handles a block of kSobelTiles horizontally adjacent 8x4 tiles per iteration: the block's (8 kSobelTiles + 2) x 6
// pixel neighbourhood is staged in shared memory with all its coalesced row loads in flight at once, every lane then
// evaluates one pixel of each tile, and ONE atomicAdd reserves the list entries of the whole block (the kernel was bound by
// the latency of one small load group and one returning atomic per 32 pixels: 33 us for the 1080p frame).
constexpr int kSobelTiles = 4;
constexpr int kSobelRowFloats = (8 * kSobelTiles + 2) * 3;      // 102
__global__ void __launch_bounds__(kBlock) k_sobel(int width, int height, const float* __restrict__ fb, const int* __restrict__ rows, int nRows,
    int* __restrict__ flagged, int flaggedCap, FrameCtr* ctr)
{
    __shared__ float tileMem[kBlock / 32][6][kSobelRowFloats + 1];
    float (*tile)[kSobelRowFloats + 1] = tileMem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int blocksX = (width + 8 * kSobelTiles - 1) / (8 * kSobelTiles);
    const long long total = (long long)blocksX * ((nRows + 3) / 4);
    const long long warpId = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nWarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float op[3][3] = { { -1, 0, 1 }, { -2, 0, 2 }, { -1, 0, 1 } };
    for (long long blk = warpId; blk < total; blk += nWarps) {
        const int x0 = (int)(blk % blocksX) * 8 * kSobelTiles, r0 = (int)(blk / blocksX) * 4;
        const int r = r0 + (lane >> 3);
        const int yFirst = rows[r0], rLast = min(r0 + 3, nRows - 1);
        // the block's rows are consecutive image rows (always, unless a strip partition cuts through it): stage them;
        // otherwise every lane reads its 3x3 windows directly
        const bool staged = rows[rLast] - yFirst == rLast - r0 && yFirst >= 1 && yFirst + (rLast - r0) < height - 1;
        const int y = r < nRows ? rows[r] : 0;
        if (staged) {
            const int cBase = (x0 - 1) * 3;
            __syncwarp();
#pragma unroll
            for (int rr = 0; rr < 6; ++rr) {
                const size_t rowOff = (size_t)(yFirst - 1 + rr) * width * 3;
#pragma unroll
                for (int k = 0; k < (kSobelRowFloats + 31) / 32; ++k) {
                    const int j = k * 32 + lane, c = cBase + j;
                    if (j < kSobelRowFloats) tile[rr][j] = (yFirst - 1 + rr < height && c >= 0 && c < width * 3) ? fb[rowOff + c] : 0.0f;
                }
            }
            __syncwarp();
        }
        unsigned votes[kSobelTiles];
        int pixOf[kSobelTiles];
#pragma unroll
        for (int t = 0; t < kSobelTiles; ++t) {
            const int x = x0 + t * 8 + (lane & 7);
            const bool interior = x < width && r < nRows && y >= 1 && y < height - 1 && x >= 1 && x < width - 1;
            bool flag = false;
            if (interior) {
                V3 gx = mk(0.0f, 0.0f, 0.0f), gy = mk(0.0f, 0.0f, 0.0f);
                if (staged) {
                    const int ly = y - yFirst, lx = (t * 8 + (lane & 7)) * 3;   // window rows ly..ly+2, float columns lx..lx+8
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int b = 0; b < 3; ++b) {
                            const V3 c = mk(tile[ly + a][lx + 3 * b], tile[ly + a][lx + 3 * b + 1], tile[ly + a][lx + 3 * b + 2]);
                            gx = gx + c * op[a][b];
                            gy = gy + c * op[b][a];
                        }
                } else {
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int b = 0; b < 3; ++b) {
                            const V3 c = loadSlot(fb, (y - 1 + a) * width + x - 1 + b);
                            gx = gx + c * op[a][b];
                            gy = gy + c * op[b][a];
                        }
                }
                // val = sqrtf(powf(|Gx|,2) + powf(|Gy|,2)) > 0.5f with |G| = (float)sqrt((double)G.G)  (scene.cpp:565-566,
                // geometry.h:99).  Away from the threshold the float sum of squares decides with a wide margin and the
                // double-precision square roots are skipped.
                const float s2 = dot(gx, gx) + dot(gy, gy);
                if (s2 > 0.2501f) flag = true;
                else if (s2 < 0.2499f) flag = false;
                else {
                    // powf(v, 2) is compiled to v * v: GCC expands pow with the exponents -1, 0, 1, 2 inline at every
                    // optimisation level (no libm call in the reference binary's launchSSAA), unlike castRay's powf(x, nSpecular)
                    const float lx = length(gx), ly = length(gy);
                    flag = sqrtf(lx * lx + ly * ly) > 0.5f;
                }
            }
            votes[t] = __ballot_sync(0xffffffffu, flag);
            pixOf[t] = y * width + x;
        }
        int count = 0;
#pragma unroll
        for (int t = 0; t < kSobelTiles; ++t) count += __popc(votes[t]);
        if (count == 0) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(&ctr->ssaaPixels, count);
        base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
        for (int t = 0; t < kSobelTiles; ++t) {      // tile by tile: the list keeps the 8x4 locality
            if (votes[t] & (1u << lane)) {
                const int f = base + __popc(votes[t] & ((1u << lane) - 1));
                if (f < flaggedCap) flagged[f] = pixOf[t];
                else atomicOr(&ctr->overflow, OVF_FLAGGED);
            }
            base += __popc(votes[t]);
        }
    }
}

__global__ void k_ssaa_gen(Scene sc, const int* __restrict__ flagged, int flaggedCap, int slotBase, RayQueue q, const FrameCtr* ctr,
    LevelCtr* lv)
{
    const int total = min(ctr->ssaaPixels, flaggedCap) * 4;
    if (blockIdx.x == 0 && threadIdx.x == 0) lv->nRays = total;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int pix = flagged[i >> 2], k = i & 3;
        const int y = pix / sc.width, x = pix % sc.width;
        // sample order (.25,.25) (.25,.75) (.75,.25) (.75,.75)  (scene.cpp:527-534)
        const float ox = (k & 2) ? 0.75f : 0.25f, oy = (k & 1) ? 0.75f : 0.25f;
        const V3 d = cameraDir(sc, (float)x + ox, (float)y + oy);
        q.o[i] = make_float4(sc.camPos.x, sc.camPos.y, sc.camPos.z, 0.0f);
        q.d[i] = make_float4(d.x, d.y, d.z, 0.0f);
        q.dest[i] = slotBase + i;
    }
}

__global__ void k_ssaa_resolve(const int* __restrict__ flagged, int flaggedCap, int slotBase, float* __restrict__ slots, const FrameCtr* ctr)
{
    const int nFlagged = min(ctr->ssaaPixels, flaggedCap);
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nFlagged; f += gridDim.x * blockDim.x) {
        V3 c = mk(0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int k = 0; k < 4; ++k) c = c + loadSlot(slots, slotBase + f * 4 + k);
        storeSlot(slots, flagged[f], c / 4.0f);
    }
}

// ------------------------------------------------------------------------------------------------
// background fill
// ------------------------------------------------------------------------------------------------
// When the geometry's screen-space bounds cover only part of the frame, primary rays are generated for that part
// only; every other rendered pixel is what castRay returns for a miss without a skybox: the background colour
// (scene.cpp:383,945).  The last row and column stay black (never rendered, scene.cpp:369-372).
// Only the listed rows are written (a rank of a multi-GPU frame touches its strips and their halo, not the frame), and of
// those only the pixels outside columns [sx0, sx1) x rows [sy0, sy1): every pixel inside is a generated primary ray, whose
// colour the pass-1 kernel stores whether it hits or misses.
__global__ void k_fill_background(float* __restrict__ fb, int width, int height, const int* __restrict__ rows, int nRows, V3 bg,
    int sx0, int sx1, int sy0, int sy1)
{
    // one CTA per listed row at a time: no per-pixel division, rows inside the skip range only touch their two margins
    for (int r = blockIdx.x; r < nRows; r += gridDim.x) {
        const int y = rows[r];
        const bool skipRow = y >= sy0 && y < sy1 && sx1 > sx0;
        float* row = fb + (size_t)y * width * 3;
        for (int x = threadIdx.x; x < width; x += blockDim.x) {
            if (skipRow && x >= sx0 && x < sx1) {
                if (sx1 - sx0 >= (int)blockDim.x) { x = ((sx1 - (int)threadIdx.x + (int)blockDim.x - 1) / (int)blockDim.x) * (int)blockDim.x + (int)threadIdx.x - (int)blockDim.x; }
                continue;
            }
            const bool rendered = x < width - 1 && y < height - 1;
            row[3 * x] = rendered ? bg.x : 0.0f; row[3 * x + 1] = rendered ? bg.y : 0.0f; row[3 * x + 2] = rendered ? bg.z : 0.0f;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// screen coverage of the geometry -> the tiles primary rays are generated for (all on the device)
// ------------------------------------------------------------------------------------------------
// primaryRect (scene_pack.h, host) bounds the generation RECTANGLE with the root boxes.  Inside it most 8x4 tiles of a thin or
// diagonal object are still empty (cfg4: 1.0 M of 1.6 M generated primaries missed everything).  With finer boxes — a mesh's
// search-BVH child boxes 10 levels down, spheres' boxes — the same projection gives a per-row, per-8-pixel-cell COVERAGE bitmap;
// tiles outside it are misses by construction and get the background colour instead of rays.  Everything runs on the device
// (a camera sweep changes it every frame): k_cover_mark projects the boxes with the device copy of the camera, k_tile_lists
// turns the bitmap into the list of tiles to generate rays for and the list of tiles to fill.  The lists stay resident while
// camera and rows do not change.  tests/test_fast_path_cpu.py checks the same projection (scene_pack.h) against the oracle's
// hit pixels; the frames stay bit-identical (goldens, camera sweeps, rotated cameras).
struct CoverCtr {
    int nKept, nSkipped;
    int invalid;           // a box reaches behind the camera plane / is not finite: no finer bound, every tile is kept
    int pad;
    long long livePixels;  // pixels of the kept tiles (statistics)
};

// one warp per box: project its 8 corners (rt_device.cuh pixelBoundsOfBox), mark rows x cells of the expanded pixel rectangle
__global__ void k_cover_mark(const Scene* __restrict__ scp, const float* __restrict__ boxes, int nBoxes, unsigned* __restrict__ bits, int cellsX, CoverCtr* ctr)
{
    const int lane = threadIdx.x & 31;
    const int width = scp->width, height = scp->height;
    const int wordsPerRow = (cellsX + 31) / 32;
    for (long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; w < nBoxes; w += ((long long)gridDim.x * blockDim.x) >> 5) {
        double minX, maxX, minY, maxY;
        const bool ok = pixelBoundsOfBox(*scp, boxes + (size_t)w * 6, minX, maxX, minY, maxY);
        if (!ok) { if (lane == 0) atomicExch(&ctr->invalid, 1); continue; }
        const int wm1 = width - 1, hm1 = height - 1;
        const int px0 = (int)fmax(0.0, fmin((double)wm1, floor(minX) - 2)), px1 = (int)fmax(0.0, fmin((double)wm1, ceil(maxX) + 3));
        const int py0 = (int)fmax(0.0, fmin((double)hm1, floor(minY) - 2)), py1 = (int)fmax(0.0, fmin((double)hm1, ceil(maxY) + 3));
        if (px1 <= px0 || py1 <= py0) continue;
        const int c0 = px0 / 8, c1 = (px1 - 1) / 8;                       // cells [c0, c1]
        const int w0 = c0 / 32, w1 = c1 / 32;
        const int nWords = w1 - w0 + 1;
        for (int i = lane; i < (py1 - py0) * nWords; i += 32) {
            const int y = py0 + i / nWords, wd = w0 + i % nWords;
            const int lo = max(c0, wd * 32) - wd * 32, hi = min(c1, wd * 32 + 31) - wd * 32;
            const unsigned mask = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1)) & ~((1u << lo) - 1);
            atomicOr(&bits[(size_t)y * wordsPerRow + wd], mask);
        }
    }
}

// one thread per 8x4 tile of the generation grid: covered (or no valid bitmap) -> kept list, else -> skipped list
__global__ void k_tile_lists(const unsigned* __restrict__ bits, int cellsX, const int* __restrict__ rows, int nRows, int x0, int cols,
    int* __restrict__ kept, int* __restrict__ skipped, CoverCtr* ctr)
{
    const int tilesX = (cols + 7) / 8, tilesY = (nRows + 3) / 4;
    const int wordsPerRow = (cellsX + 31) / 32;
    const bool all = ctr->invalid != 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ((tilesX * tilesY + 31) & ~31); t += gridDim.x * blockDim.x) {
        bool valid = t < tilesX * tilesY, covered = false;
        int live = 0;
        if (valid) {
            const int tx = t % tilesX, ty = t / tilesX;
            const int xa = x0 + tx * 8, xb = min(xa + 7, x0 + cols - 1);
            const int r0 = ty * 4, r1 = min(r0 + 4, nRows);
            live = (r1 - r0) * (xb - xa + 1);
            covered = all;
            for (int rr = r0; rr < r1 && !covered; ++rr) {
                const unsigned* rowBits = bits + (size_t)rows[rr] * wordsPerRow;
                const int ca = xa / 8, cb = xb / 8;
                covered = ((rowBits[ca / 32] >> (ca & 31)) & 1u) || ((rowBits[cb / 32] >> (cb & 31)) & 1u);
            }
        }
        // warp-aggregated appends keep the lists in (nearly) raster order
        const int k = warpAlloc(&ctr->nKept, valid && covered, 1);
        const int s = warpAlloc(&ctr->nSkipped, valid && !covered, 1);
        if (valid && covered) { kept[k] = t; atomicAdd((unsigned long long*)&ctr->livePixels, (unsigned long long)live); }
        else if (valid) skipped[s] = t;
    }
}

// The tiles of the skipped list: their pixels are misses by construction -> background colour.  One warp per tile.
__global__ void k_fill_tiles(float* __restrict__ fb, int width, const int* __restrict__ tiles, const CoverCtr* __restrict__ ctr, const int* __restrict__ rows,
    int nRows, int x0, int cols, V3 bg)
{
    const int tilesX = (cols + 7) / 8;
    const int lane = threadIdx.x & 31;
    const int nTiles = ctr->nSkipped;
    for (long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; w < nTiles; w += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int tile = tiles[w];
        const int xr = (tile % tilesX) * 8 + (lane & 7), rr = (tile / tilesX) * 4 + (lane >> 3);
        if (xr < cols && rr < nRows) storeSlot(fb, rows[rr] * width + x0 + xr, bg);
    }
}

// ------------------------------------------------------------------------------------------------
// texture upload: loadBMP's 3-byte texels (util.cpp:78-113) -> RGBA8 on the device
// ------------------------------------------------------------------------------------------------
// The reference expands every map to float on the host (12 B / texel, objects.cpp:396-458); here the file's bytes go
// over PCIe as they are (3 B / texel) and are widened to the 4 B / texel a texture object needs by this kernel.
__global__ void k_rgb_to_rgba(const unsigned char* __restrict__ rgb, size_t nTexels, uchar4* __restrict__ rgba)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nTexels; i += (size_t)gridDim.x * blockDim.x)
        rgba[i] = make_uchar4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 255);
}

// ------------------------------------------------------------------------------------------------
// showAC debug view (scene.cpp:607-635): boxes of the reference tree a primary ray's line passes
// ------------------------------------------------------------------------------------------------
// Scene::countAC / recCountAC (scene.cpp:659-669, objects.cpp:572-585) for every pixel of the frame (this view
// renders the last row / column too, and its pixel centre is x+0.5).  counts: w*h ints.
__global__ void __launch_bounds__(kBlock) k_count_ac(Scene sc, int* __restrict__ counts, FrameCtr* ctr)
{
    extern __shared__ int stackMem[];
    int* stack = stackMem + threadIdx.x;
    const bool useAC = sc.flags & FLAG_USE_AC;
    const int total = sc.width * sc.height;
    int localMax = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / sc.width, x = i - y * sc.width;
        const RayCtx r = makeRay(sc.camPos, cameraDir(sc, (float)x, (float)y));
        int sum = 0;
        for (int k = 0; k < sc.nObjects; ++k) {
            const Object& ob = sc.objects[k];
            if (ob.type != OBJ_MESH) continue;
            const Mesh& me = sc.meshes[ob.mesh];
            if (me.nNodes == 0) continue;
            int sp = 0, node = 0;
            for (;;) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(me.nodes) + 2 * node);
                const float4 b = __ldg(reinterpret_cast<const float4*>(me.nodes) + 2 * node + 1);
                if (!useAC || lineHitsBox(r, a.x, a.y, a.z, a.w, b.x, b.y)) {
                    sum++;
                    if (__float_as_int(b.w) < 0) {          // inner node: left child next, right child later
                        stack[sp * kBlock] = __float_as_int(b.z);
                        sp++;
                        node = node + 1;
                        continue;
                    }
                }
                if (sp == 0) break;
                sp--;
                node = stack[sp * kBlock];
            }
        }
        counts[i] = sum;
        localMax = max(localMax, sum);
    }
    for (int o = 16; o > 0; o >>= 1) localMax = max(localMax, __shfl_xor_sync(0xffffffffu, localMax, o));
    if ((threadIdx.x & 31) == 0 && localMax > 0) atomicMax(&ctr->acMax, localMax);
}

// frameBuffer = Vec3f{ (float)count / acMax }  (scene.cpp:627-632; 0/0 = NaN when no pixel sees a box, as in the reference)
__global__ void k_ac_resolve(const int* __restrict__ counts, int total, const FrameCtr* ctr, float* __restrict__ fb)
{
    const int acMax = ctr->acMax;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float val = (float)counts[i];
        val /= acMax;
        storeSlot(fb, i, mk(val, val, val));
    }
}

// ------------------------------------------------------------------------------------------------
// output
// ------------------------------------------------------------------------------------------------
// copy owned rows of the full-frame slot region into a compact output, 16 bytes per thread
// (rowFloats = 3*width is a multiple of 4 whenever width is; the scalar tail covers the rest)
__global__ void k_gather_rows(const float* __restrict__ fb, int width, const int* __restrict__ rows, int nRows, float* __restrict__ out)
{
    const int rowFloats = width * 3;
    if ((rowFloats & 3) == 0 && (((size_t)fb | (size_t)out) & 15) == 0) {
        const int rowVecs = rowFloats >> 2;
        const long long total = (long long)nRows * rowVecs;
        const float4* src = reinterpret_cast<const float4*>(fb);
        float4* dst = reinterpret_cast<float4*>(out);
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const int r = (int)(i / rowVecs);
            dst[i] = src[(long long)rows[r] * rowVecs + (i - (long long)r * rowVecs)];
        }
        return;
    }
    const long long total = (long long)nRows * rowFloats;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / rowFloats);
        out[i] = fb[(long long)rows[r] * rowFloats + (i % rowFloats)];
    }
}

// Same rows, but written at their IMAGE position of a full-frame buffer: out may be a peer-mapped pointer to
// another GPU's framebuffer (NVLink stores), which makes this kernel the gather of the multi-GPU frame.
__global__ void k_scatter_rows(const float* __restrict__ fb, int width, const int* __restrict__ rows, int nRows, float* __restrict__ out)
{
    const int rowFloats = width * 3;
    if ((rowFloats & 3) == 0 && (((size_t)fb | (size_t)out) & 15) == 0) {
        const int rowVecs = rowFloats >> 2;
        const long long total = (long long)nRows * rowVecs;
        const float4* src = reinterpret_cast<const float4*>(fb);
        float4* dst = reinterpret_cast<float4*>(out);
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const int r = (int)(i / rowVecs);
            const long long at = (long long)rows[r] * rowVecs + (i - (long long)r * rowVecs);
            dst[at] = src[at];
        }
        return;
    }
    const long long total = (long long)nRows * rowFloats;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / rowFloats);
        const long long at = (long long)rows[r] * rowFloats + (i % rowFloats);
        out[at] = fb[at];
    }
}

// saveImage's pixel conversion on the device (util.cpp:46-56): rows bottom-up, B,G,R byte order, each channel
// (uint8)(clamp(0,1,v)*255), rows padded to a multiple of 4 bytes.  Output row j holds image row rows[nRows-1-j].
// One thread converts 4 bytes (= 4 channels) and stores them as one 32-bit word.
__device__ __forceinline__ unsigned int bgrWord(const float* __restrict__ srcRow, int width, int wordInRow)
{
    unsigned int word = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int byteInRow = wordInRow * 4 + b;
        unsigned int v = 0;
        if (byteInRow < width * 3) {
            const int px = byteInRow / 3, ch = 2 - (byteInRow - px * 3);   // B,G,R
            const float c = clampf_(0.0f, 1.0f, srcRow[px * 3 + ch]);
            v = (unsigned int)(unsigned char)(c * 255);
        }
        word |= v << (8 * b);
    }
    return word;
}

__global__ void k_quantize_bgr8(const float* __restrict__ fb, int width, const int* __restrict__ rows, int nRows, unsigned int* __restrict__ out)
{
    const int rowBytes = (width * 3 + 3) & ~3;
    const int rowWords = rowBytes >> 2;
    const long long total = (long long)nRows * rowWords;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i / rowWords);
        const int wordInRow = (int)(i - (long long)j * rowWords);
        out[i] = bgrWord(fb + (long long)rows[nRows - 1 - j] * width * 3, width, wordInRow);
    }
}

// Early output (rtb_api.cu enqueueAttempt): the pass-1 frame's bytes are already on their way to the host while Sobel and the
// SSAA pass run; afterwards only the pixels SSAA re-traced differ.  They are rewritten straight into the host buffer (`out` is
// the device-side address of pinned host memory: the stores cross PCIe as posted writes, whose cost is per transaction, not per
// byte).  So the unit is the 32-byte sector: a warp takes 32 flagged pixels (neighbours in the list are neighbours in the image,
// k_sobel), finds the distinct sectors their bytes fall in, and rewrites each sector whole — 8 lanes, one 32-bit word each,
// recomputed from the float frame — as ONE coalesced store.  A sector shared with the next warp is written twice with the same
// bytes.  Rows [y0, y1) of the image are the buffer's rows, bottom-up; outWords = words of the buffer.
// FLOATS = true: the same for the float32 frame (rtb_render into pinned host memory) — rows top-down, 12 bytes per pixel, and a
// sector's words are the frame's own.
template <bool FLOATS>
__global__ void k_patch_rows(const float* __restrict__ fb, int width, int y0, int y1, const int* __restrict__ flagged, int cap, const FrameCtr* __restrict__ ctr,
    unsigned int* __restrict__ out, long long outWords)
{
    const unsigned FULL = 0xffffffffu;
    const int n = min(ctr->ssaaPixels, cap);
    const int lane = threadIdx.x & 31;
    const int rowWords = ((width * 3 + 3) & ~3) >> 2;
    const int nChunks = (n + 31) / 32;
    for (int chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < nChunks; chunk += (gridDim.x * blockDim.x) >> 5) {
        const int i = chunk * 32 + lane;
        long long s0 = -1, s1 = -1;                    // sectors (32 B = 8 words) holding the pixel's first / last byte
        if (i < n) {
            const int pix = flagged[i];
            const int y = pix / width, x = pix - y * width;
            if (y >= y0 && y < y1) {
                const long long byte0 = FLOATS ? ((long long)(y - y0) * width + x) * 12 : (long long)(y1 - 1 - y) * rowWords * 4 + (long long)x * 3;
                s0 = byte0 >> 5;
                s1 = (byte0 + (FLOATS ? 11 : 2)) >> 5;
                if (s1 == s0) s1 = -1;
            }
        }
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            const long long s = pass == 0 ? s0 : s1;
            // one lane per distinct sector: the lowest lane of every group of equal values
            const unsigned same = __match_any_sync(FULL, s);
            const bool leader = s >= 0 && (__ffs(same) - 1) == lane;
            unsigned todo = __ballot_sync(FULL, leader);
            while (todo) {
                // the next (up to) four sectors, eight lanes each
                long long mine = -1;
                unsigned rest = todo;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int src = rest ? __ffs(rest) - 1 : 0;
                    const long long sg = __shfl_sync(FULL, s, src);
                    if (rest && (lane >> 3) == g) mine = sg;
                    rest &= rest - 1;
                }
                todo = rest;
                if (mine >= 0) {
                    const long long word = mine * 8 + (lane & 7);
                    if (word < outWords) {
                        if (FLOATS) {
                            out[word] = __float_as_uint(fb[(long long)y0 * width * 3 + word]);
                        } else {
                            const int j = (int)(word / rowWords), wordInRow = (int)(word - (long long)j * rowWords);
                            out[word] = bgrWord(fb + (long long)(y1 - 1 - j) * width * 3, width, wordInRow);
                        }
                    }
                }
            }
        }
    }
}

// the frame's counters into their pinned host mirror by plain stores: no copy-engine operation at the end of a frame, which
// would queue behind a frame-sized device-to-host copy in flight
__global__ void k_store_words(const unsigned int* __restrict__ src, unsigned int* __restrict__ dst, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

} // namespace rtk
