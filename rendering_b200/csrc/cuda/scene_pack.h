// scene_pack.h — host-side packing of the ABI's RtbScene into the device layouts of rt_device.cuh.
// Header-only, plain C++: rtb_api.cu copies the packed arrays to HBM; tests/shim uses the same
// packing to exercise rt_device.cuh on the CPU.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../../include/rtb.h"
#include "bvh_build.h"
#include "rt_device.cuh"

namespace rtpack {

struct PackedMesh {
    std::vector<rt::Node> nodes;
    std::vector<rt::TriSlot> slots;
    int maxDepth = 0;
};

inline void packMesh(const RtbMesh& m, PackedMesh& out)
{
    out.nodes.resize(m.nNodes);
    out.maxDepth = 0;
    for (int k = 0; k < m.nNodes; ++k) {
        const RtbNode& s = m.nodes[k];
        rt::Node& d = out.nodes[k];
        d.lox = s.lo[0]; d.loy = s.lo[1]; d.loz = s.lo[2];
        d.hix = s.hi[0]; d.hiy = s.hi[1]; d.hiz = s.hi[2];
        if (s.right >= 0) { d.link = s.right; d.count = -1; }
        else { d.link = s.firstRef; d.count = s.refCount; }
        out.maxDepth = std::max(out.maxDepth, s.depth);
    }
    out.slots.resize(m.nRefs);
    for (int r = 0; r < m.nRefs; ++r) {
        const int tri = m.refs[r];
        const float* p = m.pos + (size_t)tri * 9;
        rt::TriSlot& d = out.slots[r];
        std::memset(&d, 0, sizeof d);
        d.v0x = p[0]; d.v0y = p[1]; d.v0z = p[2];
        d.tri = tri;
        d.e1x = p[3] - p[0]; d.e1y = p[4] - p[1]; d.e1z = p[5] - p[2];   // v1 - v0 (objects.cpp:70)
        d.e2x = p[6] - p[0]; d.e2y = p[7] - p[1]; d.e2z = p[8] - p[2];   // v2 - v0 (objects.cpp:71)
    }
}

// Fast-path tables of one mesh (see rt_device.cuh walkMeshFast / eligibleSlot).
struct FastPath {
    std::vector<rtbvh::Node> nodes;
    std::vector<float4> tris;        // 3 per triangle, BVH leaf order
    std::vector<int> triRefOff;      // nTris + 1
    std::vector<int2> triRefs;       // (reference leaf node, slot), ascending slot per triangle
    std::vector<int> parent;         // reference-tree parents
    int maxDepth = 0;
    float pad = 0;
};

inline void packFastPath(const RtbMesh& m, FastPath& out)
{
    // padding: a hit accepted by the float Moller-Trumbore test lies within rounding distance of the
    // triangle; 1e-4 of the mesh diagonal is orders of magnitude above that
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (size_t i = 0; i < (size_t)m.nTris * 3; ++i)
        for (int a = 0; a < 3; ++a) {
            const float v = m.pos[i * 3 + a];
            if (v >= -FLT_MAX && v <= FLT_MAX) { lo[a] = std::min(lo[a], v); hi[a] = std::max(hi[a], v); }
        }
    double diag2 = 0;
    for (int a = 0; a < 3; ++a) diag2 += hi[a] > lo[a] ? (double)(hi[a] - lo[a]) * (hi[a] - lo[a]) : 0.0;
    out.pad = (float)(1e-4 * std::sqrt(diag2)) + 1e-7f;

    rtbvh::Builder builder;
    int maxLeaf = 4;
    if (const char* e = std::getenv("RTB_BVH_LEAF")) maxLeaf = std::atoi(e);   // tuning knob
    rtbvh::Result bvh = builder.build(m.pos, m.nTris, out.pad, maxLeaf);
    out.maxDepth = bvh.maxDepth;
    out.nodes.swap(bvh.nodes);
    out.tris.resize(bvh.triOrder.size() * 3);
    for (size_t k = 0; k < bvh.triOrder.size(); ++k) {
        const int tri = bvh.triOrder[k];
        const float* p = m.pos + (size_t)tri * 9;
        float triAsFloat;
        std::memcpy(&triAsFloat, &tri, 4);
        out.tris[k * 3 + 0] = float4{ p[0], p[1], p[2], triAsFloat };
        out.tris[k * 3 + 1] = float4{ p[3] - p[0], p[4] - p[1], p[5] - p[2], 0.0f };   // v1 - v0 (objects.cpp:70)
        out.tris[k * 3 + 2] = float4{ p[6] - p[0], p[7] - p[1], p[8] - p[2], 0.0f };   // v2 - v0 (objects.cpp:71)
    }

    out.parent.assign(m.nNodes, -1);
    std::vector<int>& off = out.triRefOff;
    off.assign(m.nTris + 1, 0);
    for (int k = 0; k < m.nNodes; ++k) {
        const RtbNode& n = m.nodes[k];
        if (n.right >= 0) { out.parent[k + 1] = k; out.parent[n.right] = k; }
        else for (int s = n.firstRef; s < n.firstRef + n.refCount; ++s) off[m.refs[s] + 1]++;
    }
    for (int t = 0; t < m.nTris; ++t) off[t + 1] += off[t];
    out.triRefs.resize(m.nRefs);
    std::vector<int> cursor(off.begin(), off.end() - 1);
    for (int k = 0; k < m.nNodes; ++k) {   // pre-order: leaves, hence slots, in ascending DFS order
        const RtbNode& n = m.nodes[k];
        if (n.right >= 0) continue;
        for (int s = n.firstRef; s < n.firstRef + n.refCount; ++s) out.triRefs[cursor[m.refs[s]]++] = int2{ k, s };
    }
}

inline std::vector<unsigned char> packRGBA(const RtbImage& im)
{
    std::vector<unsigned char> out;
    if (!im.rgb || im.width <= 0 || im.height <= 0) return out;
    const size_t n = (size_t)im.width * im.height;
    out.resize(n * 4);
    for (size_t i = 0; i < n; ++i) {
        out[i * 4 + 0] = im.rgb[i * 3 + 0];
        out[i * 4 + 1] = im.rgb[i * 3 + 1];
        out[i * 4 + 2] = im.rgb[i * 3 + 2];
        out[i * 4 + 3] = 255;
    }
    return out;
}

inline rt::V3 v3of(const float* p) { return rt::mk(p[0], p[1], p[2]); }

inline rt::Object packObject(const RtbObject& o)
{
    rt::Object d;
    d.type = o.type; d.material = o.material;
    d.color = v3of(o.color);
    d.ior = o.ior; d.ambient = o.ambient; d.diffuse = o.diffuse; d.specular = o.specular; d.nSpecular = o.nSpecular;
    d.pos = v3of(o.pos); d.r2 = o.r2; d.normal = v3of(o.normal); d.mesh = o.mesh;
    return d;
}

inline rt::Light packLight(const RtbLight& l)
{
    rt::Light d;
    d.type = l.type; d.color = v3of(l.color); d.intensity = l.intensity; d.v = v3of(l.v);
    d.pointOffset = l.pointOffset; d.pointCount = l.pointCount;
    return d;
}

// everything of rt::Scene that is not a pointer
inline void packHeader(const RtbScene& s, rt::Scene& d)
{
    std::memset(&d, 0, sizeof d);
    d.width = s.width; d.height = s.height; d.bias = s.bias; d.maxRayDepth = s.maxRayDepth;
    d.background = v3of(s.backgroundColor);
    d.flags = s.flags;
    d.camPos = v3of(s.camera.pos);
    for (int i = 0; i < 16; ++i) d.camM[i] = s.camera.rMatrix[i];
    d.camScale = s.camera.scale; d.camAspect = s.camera.aspect;
    d.nObjects = s.nObjects; d.nLights = s.nLights; d.nMeshes = s.nMeshes;
    int spp = 0;
    for (int i = 0; i < s.nLights; ++i) spp += (s.lights[i].type == RTB_LIGHT_AREA) ? s.lights[i].pointCount : 1;
    d.shadowRaysPerHit = spp;
}

} // namespace rtpack
