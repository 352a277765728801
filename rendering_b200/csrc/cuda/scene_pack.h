// scene_pack.h — host-side packing of the ABI's RtbScene into the device layouts of rt_device.cuh.
// Header-only, plain C++: rtb_api.cu copies the packed arrays to HBM; tests/shim uses the same
// packing to exercise rt_device.cuh on the CPU.
#pragma once

#include <algorithm>
#include <array>
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
    int topNodes = 0;                // nodes [0, topNodes) are the tree's top levels in breadth-first order
    float lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 };   // finite bounds of the mesh's vertices
    float pad = 0;
};

// Renumbers the search BVH so that its top levels come first in breadth-first order (at most `topBudget` nodes: whole
// levels only), followed by the subtrees below them in the builder's depth-first order.  A prefix of the array is then
// "the top of the tree" — what the tile kernel stages into shared memory — while deep subtrees keep their locality.
// Only indices change: child order inside a node, boxes and leaf codes stay, so traversal visits the same nodes.
inline int reorderTopLevelsFirst(std::vector<rtbvh::Node>& nodes, int topBudget)
{
    const int n = (int)nodes.size();
    if (n <= 1) return n;
    std::vector<int> order;          // order[newIndex] = oldIndex
    order.reserve(n);
    std::vector<int> level{ 0 }, nextLevel;
    // whole levels while they fit the budget
    while (!level.empty() && (int)(order.size() + level.size()) <= topBudget) {
        nextLevel.clear();
        for (int k : level) {
            order.push_back(k);
            if (nodes[k].child0 >= 0) nextLevel.push_back(nodes[k].child0);
            if (nodes[k].child1 >= 0) nextLevel.push_back(nodes[k].child1);
        }
        level.swap(nextLevel);
    }
    const int nTop = (int)order.size();
    // the rest: depth-first below every frontier node, frontier in breadth-first order
    std::vector<int> stack;
    for (int root : level) {
        stack.push_back(root);
        while (!stack.empty()) {
            const int k = stack.back();
            stack.pop_back();
            order.push_back(k);
            if (nodes[k].child1 >= 0) stack.push_back(nodes[k].child1);
            if (nodes[k].child0 >= 0) stack.push_back(nodes[k].child0);
        }
    }
    std::vector<int> newIndex(n, -1);
    for (int i = 0; i < (int)order.size(); ++i) newIndex[order[i]] = i;
    std::vector<rtbvh::Node> out(order.size());
    for (int i = 0; i < (int)order.size(); ++i) {
        rtbvh::Node nd = nodes[order[i]];
        if (nd.child0 >= 0) nd.child0 = newIndex[nd.child0];
        if (nd.child1 >= 0) nd.child1 = newIndex[nd.child1];
        out[i] = nd;
    }
    nodes.swap(out);
    return nTop;
}

constexpr int kTopLevelBudget = 2048;   // nodes of the breadth-first prefix (128 KB): what shared memory can hold next to the stacks

// buildSearch = false: only the eligibility tables, the padding and the mesh bounds (the search BVH is then built on the
// device, lbvh_build.cuh)
inline void packFastPath(const RtbMesh& m, FastPath& out, bool buildSearch = true)
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
    for (int a = 0; a < 3; ++a) { out.lo[a] = lo[a]; out.hi[a] = hi[a]; }

    if (buildSearch) {
    rtbvh::Builder builder;
    int maxLeaf = 4;
    if (const char* e = std::getenv("RTB_BVH_LEAF")) maxLeaf = std::atoi(e);   // tuning knob
    rtbvh::Result bvh = builder.build(m.pos, m.nTris, out.pad, maxLeaf);
    out.maxDepth = bvh.maxDepth;
    out.topNodes = reorderTopLevelsFirst(bvh.nodes, kTopLevelBudget);
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

// Pixel rectangle {x0, x1, y0, y1} (columns [x0,x1), rows [y0,y1)) that contains the projection of every bounded
// object, expanded by 2 pixels (the projection uses the camera constants in double; rounding is ~1e-4 pixel).
// The whole rendered frame when anything is unbounded, behind / around the camera, or when missing rays need their
// direction (skybox).  bounds: world-space boxes lo.xyz, hi.xyz around everything a primary ray can hit.
// cover (optional): per image row and 8-pixel column cell (cover[y * cellsX + x / 8], cellsX = (width + 7) / 8), 1 where the
// projection of at least one box falls — with fine boxes (a mesh's search-BVH boxes a few levels down, meshCoverBoxes) the
// primary rays of every 8x4 tile outside it are misses by construction and are not generated at all.
inline bool pixelBoundsOfBox(const rt::Scene& sc, const std::array<float, 6>& b, double& minX, double& maxX, double& minY, double& maxY)
{
    return rt::pixelBoundsOfBox(sc, b.data(), minX, maxX, minY, maxY);      // rt_device.cuh: shared with the device's k_cover_mark
}

inline void primaryRect(const rt::Scene& sc, const std::vector<std::array<float, 6>>& bounds, bool unbounded, int r[4],
    std::vector<unsigned char>* cover = nullptr)
{
    const int wm1 = sc.width - 1, hm1 = sc.height - 1;
    r[0] = 0; r[1] = wm1; r[2] = 0; r[3] = hm1;
    if (cover) cover->clear();
    if (unbounded || (sc.flags & rt::FLAG_SKYBOX)) return;
    // The projection below inverts rMatrix by transposing its 3x3 block, which is only right for what Camera::getRay builds
    // (a pure rotation, scene.cpp:24-48).  rtb_set_camera accepts any matrix through the C ABI: anything else — scale, shear,
    // a translation row or a w column (mulRowVecMatrix applies both) — renders the whole frame, like the handles that never cull.
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            double d = 0;
            for (int k = 0; k < 3; ++k) d += (double)sc.camM[i * 4 + k] * sc.camM[j * 4 + k];
            if (!(std::fabs(d - (i == j ? 1.0 : 0.0)) <= 1e-4)) return;
        }
        if (sc.camM[i * 4 + 3] != 0.0f || sc.camM[3 * 4 + i] != 0.0f) return;
    }
    if (sc.camM[15] != 1.0f) return;
    auto clampi = [](double v, int lo, int hi) { return (int)std::max<double>(lo, std::min<double>(hi, v)); };
    const int cellsX = (sc.width + 7) / 8;
    std::vector<unsigned char> cells;
    if (cover) cells.assign((size_t)cellsX * sc.height, 0);
    double minX = 1e300, maxX = -1e300, minY = 1e300, maxY = -1e300;
    for (const auto& b : bounds) {
        double x0, x1, y0, y1;
        if (!pixelBoundsOfBox(sc, b, x0, x1, y0, y1)) return;
        minX = std::min(minX, x0); maxX = std::max(maxX, x1);
        minY = std::min(minY, y0); maxY = std::max(maxY, y1);
        if (cover) {
            const int px0 = clampi(std::floor(x0) - 2, 0, wm1), px1 = clampi(std::ceil(x1) + 3, 0, wm1);
            const int py0 = clampi(std::floor(y0) - 2, 0, hm1), py1 = clampi(std::ceil(y1) + 3, 0, hm1);
            if (px1 > px0 && py1 > py0)
                for (int y = py0; y < py1; ++y)
                    std::memset(&cells[(size_t)y * cellsX + px0 / 8], 1, (size_t)((px1 - 1) / 8 - px0 / 8 + 1));
        }
    }
    if (bounds.empty()) { r[0] = r[1] = r[2] = r[3] = 0; return; }
    r[0] = clampi(std::floor(minX) - 2, 0, wm1); r[1] = clampi(std::ceil(maxX) + 3, 0, wm1);
    r[2] = clampi(std::floor(minY) - 2, 0, hm1); r[3] = clampi(std::ceil(maxY) + 3, 0, hm1);
    if (r[1] <= r[0] || r[3] <= r[2]) r[0] = r[1] = r[2] = r[3] = 0;
    if (cover) cover->swap(cells);
}

// Boxes that together contain a mesh: its search BVH's child boxes `depth` levels down (leaves stop earlier).  Finer than
// the root box, so their projections hug the silhouette (primaryRect's `cover`).
inline void meshCoverBoxes(const FastPath& fp, int depth, std::vector<std::array<float, 6>>& out)
{
    if (fp.nodes.empty() || fp.tris.empty()) return;
    struct Item { int node, level; };
    std::vector<Item> stack{ { 0, 0 } };
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        const rtbvh::Node& n = fp.nodes[it.node];
        const float* lo[2] = { n.c0lo, n.c1lo };
        const float* hi[2] = { n.c0hi, n.c1hi };
        const int child[2] = { n.child0, n.child1 };
        for (int c = 0; c < 2; ++c) {
            if (lo[c][0] > hi[c][0]) continue;                               // empty child
            if (child[c] >= 0 && it.level + 1 < depth) stack.push_back({ child[c], it.level + 1 });
            else out.push_back({ lo[c][0], lo[c][1], lo[c][2], hi[c][0], hi[c][1], hi[c][2] });
        }
    }
}

// The bounds primaryRect needs: the padded root box of a mesh's search BVH ...
inline bool meshBounds(const FastPath& fp, std::array<float, 6>& b)
{
    if (fp.nodes.empty() || fp.tris.empty()) return false;
    const rtbvh::Node& root = fp.nodes[0];
    for (int a = 0; a < 3; ++a) {
        b[a] = std::min(root.c0lo[a], root.c1lo[a]);          // an empty child has lo = +FLT_MAX, hi = -FLT_MAX
        b[3 + a] = std::max(root.c0hi[a], root.c1hi[a]);
    }
    return true;
}
// ... a box around a sphere; a plane makes the scene unbounded
inline void objectBounds(const RtbObject& o, std::vector<std::array<float, 6>>& bounds, bool& unbounded)
{
    if (o.type == RTB_OBJ_PLANE) unbounded = true;
    else if (o.type == RTB_OBJ_SPHERE) {
        const float rad = std::sqrt(std::max(0.0f, o.r2)) * 1.001f + 1e-6f;
        if (!(rad < FLT_MAX)) unbounded = true;
        bounds.push_back({ o.pos[0] - rad, o.pos[1] - rad, o.pos[2] - rad, o.pos[0] + rad, o.pos[1] + rad, o.pos[2] + rad });
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
