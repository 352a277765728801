// scene_pack.h — host-side packing of the ABI's RtbScene into the device layouts of rt_device.cuh.
// Header-only, plain C++: rtb_api.cu copies the packed arrays to HBM; tests/shim uses the same
// packing to exercise rt_device.cuh on the CPU.
#pragma once

#include <algorithm>
#include <cstring>
#include <vector>

#include "../../../include/rtb.h"
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
