// tests/shim/shim_render.cpp — TEST-ONLY host harness for rendering_b200/csrc/cuda/rt_device.cuh.
//
// Compiles the backend's __host__ __device__ arithmetic with g++ and drives it with a plain
// recursive renderer, so the per-ray math (camera, box/triangle/sphere/plane tests, surface data,
// lights, Fresnel, skybox) can be checked bit-for-bit against the reference on the CPU box, where
// no GPU exists.  It is NOT part of the product and NOT the oracle: nothing under rendering_b200/
// links it, and the wavefront plumbing of the kernels is exercised only by the -m gpu tests.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../rendering_b200/csrc/cuda/scene_pack.h"

using namespace rt;

namespace {

struct HostScene {
    Scene sc;
    std::vector<Object> objects;
    std::vector<Light> lights;
    std::vector<Mesh> meshes;
    std::vector<rtpack::PackedMesh> packed;
    std::vector<std::vector<unsigned char>> images;
    std::vector<rtpack::FastPath> fast;
    bool useFast = false;
    unsigned long long rays = 0, boxTests = 0, triTests = 0;
};

Image hostImage(HostScene& hs, const RtbImage& im)
{
    Image d{};
    hs.images.push_back(rtpack::packRGBA(im));
    if (!hs.images.back().empty()) { d.rgba = hs.images.back().data(); d.w = im.width; d.h = im.height; }
    return d;
}

struct Hit { float t, u, v; int obj, tri; };

bool traceRay(HostScene& hs, V3 o, V3 d, bool shadow, float tMax, Hit& hit)
{
    const Scene& sc = hs.sc;
    hs.rays++;
    const RayCtx r = makeRay(o, d);
    hit.obj = -1; hit.t = tMax; hit.tri = -1; hit.u = hit.v = 0;
    const bool cull = sc.flags & FLAG_CULL;
    for (int k = 0; k < sc.nObjects; ++k) {
        const Object& ob = sc.objects[k];
        if (shadow && ob.material == MAT_TRANSPARENT) continue;
        float t = FLT_MAX, u = 0, v = 0; int tri = -1; bool ok = false;
        if (ob.type == OBJ_MESH) {
            const Mesh& me = sc.meshes[ob.mesh];
            if (me.nNodes == 0) continue;
            if (hs.useFast) {
                int fstack[128];
                if (shadow) ok = walkMeshFast<true>(sc, me, r, fstack, 1, tMax, t, u, v, tri);
                else ok = walkMeshFast<false>(sc, me, r, fstack, 1, 0.0f, t, u, v, tri);
                if (ok && shadow) { hit.obj = k; return true; }
                if (ok && t < hit.t) { hit.t = t; hit.u = u; hit.v = v; hit.obj = k; hit.tri = tri; }
                continue;
            }
            int stack[128]; int sp = 0; int node = 0;
            for (;;) {
                const Node& n = me.nodes[node];
                hs.boxTests += (sc.flags & FLAG_USE_AC) ? 1 : 0;
                bool descend = !(sc.flags & FLAG_USE_AC) || lineHitsBox(r, n.lox, n.loy, n.loz, n.hix, n.hiy, n.hiz);
                if (descend) {
                    if (n.count < 0) { stack[sp++] = n.link; node = node + 1; continue; }
                    for (int s = n.link; s < n.link + n.count; ++s) {
                        const TriSlot& ts = me.slots[s];
                        float tt, uu, vv;
                        hs.triTests++;
                        if (hitTriangle(r, mk(ts.v0x, ts.v0y, ts.v0z), mk(ts.e1x, ts.e1y, ts.e1z), mk(ts.e2x, ts.e2y, ts.e2z), cull, tt, uu, vv) && tt < t) {
                            t = tt; u = uu; v = vv; tri = ts.tri; ok = true;
                        }
                    }
                }
                if (sp == 0) break;
                node = stack[--sp];
            }
        } else if (ob.type == OBJ_SPHERE) {
            ok = hitSphere(r, ob.pos, ob.r2, t);
        } else {
            ok = hitPlane(r, ob.pos, ob.normal, t);
        }
        if (ok && t < hit.t) { hit.t = t; hit.u = u; hit.v = v; hit.obj = k; hit.tri = tri; }
    }
    return hit.obj >= 0;
}

V3 castRay(HostScene& hs, V3 o, V3 d, int depth)
{
    const Scene& sc = hs.sc;
    if (depth > sc.maxRayDepth) return skybox(sc, d);
    Hit h;
    if (!traceRay(hs, o, d, false, FLT_MAX, h)) return skybox(sc, d);
    const Object& ob = sc.objects[h.obj];
    const Surface s = surfaceAt(sc, ob, o, d, h.t, h.u, h.v, h.tri);
    if (sc.flags & FLAG_SHOW_NORMALS) return s.N / 2.0f + mk(0.5f, 0.5f, 0.5f);
    V3 diff = mk(0, 0, 0), spec = mk(0, 0, 0);
    const V3 shadowOrig = s.P + s.N * sc.bias;
    for (int i = 0; i < sc.nLights; ++i) {
        const Light& li = sc.lights[i];
        Hit sh;
        if (li.type != LIGHT_AREA) {
            V3 L, I; float dist;
            illuminate(li, s.P, L, I, dist);
            const float vis = traceRay(hs, shadowOrig, -L, true, dist, sh) ? 0.0f : 1.0f;
            if (ob.material == MAT_DIFFUSE) {
                diff = diff + I * (vis * maxf_(0.f, dot(s.N, -L)));
            } else {
                if (ob.material == MAT_PHONG) diff = diff + (I * vis) * maxf_(0.f, dot(s.N, -L));
                const V3 R = reflect(L, s.N);
                spec = spec + (I * vis) * powExact(maxf_(0.f, dot(R, -d)), ob.nSpecular);
            }
        } else {
            const V3 I = areaIntensity(li, s.P);
            float dsum = 0, ssum = 0;
            for (int p = 0; p < li.pointCount; ++p) {
                const float* ap = sc.areaPoints + (size_t)(li.pointOffset + p) * 3;
                V3 L = s.P - mk(ap[0], ap[1], ap[2]);
                const float dist = length(L);
                L = normalize(L);
                const float vis = traceRay(hs, shadowOrig, -L, true, dist, sh) ? 0.0f : 1.0f;
                dsum += vis * maxf_(0.f, dot(s.N, -L));
                const V3 R = reflect(L, s.N);
                ssum += vis * maxf_(0.f, dot(R, -d));
            }
            const float n = (float)li.pointCount;
            if (ob.material == MAT_DIFFUSE || ob.material == MAT_PHONG) diff = diff + I * (dsum / n);
            if (ob.material != MAT_DIFFUSE) spec = spec + I * powExact(ssum / n, ob.nSpecular);
        }
    }
    if (ob.material == MAT_DIFFUSE) return s.color * diff;
    if (ob.material == MAT_PHONG) return s.color * ob.ambient + diff * ob.diffuse + spec * s.specCoef;
    if (ob.material == MAT_REFLECTIVE) {
        const V3 c = castRay(hs, s.P + s.N * sc.bias, d - s.N * (2 * dot(d, s.N)), depth + 1);
        return c * 0.8f + spec;
    }
    const float kr = fresnel(d, s.N, ob.ior);
    const bool outside = dot(d, s.N) < 0;
    const V3 biasVec = s.N * sc.bias;
    V3 color = mk(0, 0, 0);
    if (kr < 1) {
        const V3 rd = normalize(refract(d, s.N, ob.ior));
        const V3 ro = outside ? s.P - biasVec : s.P + biasVec;
        color = color + castRay(hs, ro, rd, depth + 1) * (1 - kr);
    }
    const V3 fd = normalize(reflect(d, s.N));
    const V3 fo = outside ? s.P + biasVec : s.P - biasVec;
    color = color + castRay(hs, fo, fd, depth + 1) * kr;
    color = color + spec * kr;
    return color;
}

} // namespace

extern "C" {

// Renders the whole frame on the CPU with the backend's device arithmetic.  pass1/final: h*w*3.
// counters: {rays, boxTests, triTests, ssaaPixels}
static int render_impl(const RtbScene* s, float* pass1, float* final, unsigned long long counters[4], bool fast);

int shim_render(const RtbScene* s, float* pass1, float* final, unsigned long long counters[4]) { return render_impl(s, pass1, final, counters, false); }
// same, but meshes are searched through the fast path (search BVH + eligibility tables)
int shim_render_fast(const RtbScene* s, float* pass1, float* final, unsigned long long counters[4]) { return render_impl(s, pass1, final, counters, true); }

// FNV-1a digest of a mesh's search BVH (nodes + leaf-ordered triangles); serial != 0 forces the single-threaded build
unsigned long long shim_bvh_digest(const RtbScene* s, int mesh, int serial)
{
    if (serial) setenv("RTB_BVH_SERIAL", "1", 1); else unsetenv("RTB_BVH_SERIAL");
    rtpack::FastPath fp;
    rtpack::packFastPath(s->meshes[mesh], fp);
    unsetenv("RTB_BVH_SERIAL");
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) { const unsigned char* c = static_cast<const unsigned char*>(p); for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; } };
    mix(fp.nodes.data(), fp.nodes.size() * sizeof(rtbvh::Node));
    mix(fp.tris.data(), fp.tris.size() * sizeof(float4));
    mix(&fp.maxDepth, sizeof fp.maxDepth);
    return h;
}

// rt_device.cuh's restatement of glibc powf over n (x, y) pairs
void shim_powf(const float* x, const float* y, int n, float* out)
{
    for (int i = 0; i < n; ++i) out[i] = rt::powfGlibc(x[i], y[i]);
}

// the C library's own powf over the same pairs (what the reference's std::pow(float, float) resolves to)
void shim_libm_powf(const float* x, const float* y, int n, float* out)
{
    for (int i = 0; i < n; ++i) { volatile float a = x[i], b = y[i]; out[i] = powf(a, b); }
}

// the pixel rectangle primary rays are limited to (scene_pack.h primaryRect), computed exactly like rtb_create does
int shim_primary_rect(const RtbScene* s, int rect[4])
{
    rt::Scene sc;
    rtpack::packHeader(*s, sc);
    std::vector<std::array<float, 6>> bounds;
    bool unbounded = false;
    for (int i = 0; i < s->nMeshes; ++i) {
        if (s->meshes[i].nNodes == 0) continue;
        rtpack::FastPath fp;
        rtpack::packFastPath(s->meshes[i], fp);
        std::array<float, 6> b;
        if (rtpack::meshBounds(fp, b)) bounds.push_back(b);
    }
    for (int i = 0; i < s->nObjects; ++i) rtpack::objectBounds(s->objects[i], bounds, unbounded);
    rtpack::primaryRect(sc, bounds, unbounded, rect);
    return 0;
}

// the finer bound rtb_create / rtb_set_camera compute the same way: cover[y * cellsX + x / 8] (cellsX = (width + 7) / 8) from the
// search-BVH boxes 10 levels down; returns 1 when a coverage mask exists (cover filled), 0 when only the rectangle applies
int shim_primary_cover(const RtbScene* s, int rect[4], unsigned char* cover)
{
    rt::Scene sc;
    rtpack::packHeader(*s, sc);
    std::vector<std::array<float, 6>> bounds;
    bool unbounded = false;
    for (int i = 0; i < s->nMeshes; ++i) {
        if (s->meshes[i].nNodes == 0 || s->meshes[i].nTris == 0) continue;
        rtpack::FastPath fp;
        rtpack::packFastPath(s->meshes[i], fp);
        rtpack::meshCoverBoxes(fp, 10, bounds);
    }
    for (int i = 0; i < s->nObjects; ++i) rtpack::objectBounds(s->objects[i], bounds, unbounded);
    std::vector<unsigned char> cells;
    rtpack::primaryRect(sc, bounds, unbounded, rect, &cells);
    if (cells.empty()) return 0;
    std::copy(cells.begin(), cells.end(), cover);
    return 1;
}

} // extern "C"

static int render_impl(const RtbScene* s, float* pass1, float* final, unsigned long long counters[4], bool fast)
{
    HostScene hs;
    hs.useFast = fast;
    hs.fast.resize(s->nMeshes);
    rtpack::packHeader(*s, hs.sc);
    hs.images.reserve(64);
    for (int i = 0; i < s->nObjects; ++i) hs.objects.push_back(rtpack::packObject(s->objects[i]));
    for (int i = 0; i < s->nLights; ++i) hs.lights.push_back(rtpack::packLight(s->lights[i]));
    hs.packed.resize(s->nMeshes);
    for (int i = 0; i < s->nMeshes; ++i) {
        rtpack::packMesh(s->meshes[i], hs.packed[i]);
        Mesh m{};
        m.nodes = hs.packed[i].nodes.data(); m.slots = hs.packed[i].slots.data();
        m.nrm = s->meshes[i].nrm; m.uv = s->meshes[i].uv; m.tan = s->meshes[i].tan;
        m.diffuse = hostImage(hs, s->meshes[i].diffuseMap);
        m.normal = hostImage(hs, s->meshes[i].normalMap);
        m.specular = hostImage(hs, s->meshes[i].specularMap);
        m.nNodes = s->meshes[i].nNodes; m.nSlots = s->meshes[i].nRefs; m.nTris = s->meshes[i].nTris;
        m.maxDepth = hs.packed[i].maxDepth;
        if (fast && m.nNodes > 0) {
            rtpack::packFastPath(s->meshes[i], hs.fast[i]);
            m.bvhNodes = reinterpret_cast<const float4*>(hs.fast[i].nodes.data());
            m.bvhTris = hs.fast[i].tris.data();
            m.triRefOff = hs.fast[i].triRefOff.data();
            m.triRefs = hs.fast[i].triRefs.data();
            m.parent = hs.fast[i].parent.data();
        }
        hs.meshes.push_back(m);
    }
    for (int k = 0; k < 6; ++k) hs.sc.sky[k] = hostImage(hs, s->skybox[k]);
    hs.sc.objects = hs.objects.data(); hs.sc.lights = hs.lights.data(); hs.sc.meshes = hs.meshes.data();
    hs.sc.areaPoints = s->areaPoints;

    const int w = s->width, h = s->height;
    std::vector<float> fb((size_t)w * h * 3, 0.0f);
    for (int y = 0; y + 1 < h; ++y)
        for (int x = 0; x + 1 < w; ++x) {
            const V3 c = castRay(hs, hs.sc.camPos, cameraDir(hs.sc, (float)x + 0.5f, (float)y + 0.5f), 0);
            float* p = &fb[((size_t)y * w + x) * 3];
            p[0] = c.x; p[1] = c.y; p[2] = c.z;
        }
    if (pass1) std::copy(fb.begin(), fb.end(), pass1);
    unsigned long long flagged = 0;
    if (hs.sc.flags & FLAG_SSAA) {
        std::vector<unsigned char> flag((size_t)w * h, 0);
        const float op[3][3] = { { -1, 0, 1 }, { -2, 0, 2 }, { -1, 0, 1 } };
        for (int i = 1; i < h - 1; ++i)
            for (int j = 1; j < w - 1; ++j) {
                V3 gx = mk(0, 0, 0), gy = mk(0, 0, 0);
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) {
                        const float* p = &fb[((size_t)(i - 1 + a) * w + j - 1 + b) * 3];
                        const V3 c = mk(p[0], p[1], p[2]);
                        gx = gx + c * op[a][b];
                        gy = gy + c * op[b][a];
                    }
                const float lx = length(gx), ly = length(gy);
                flag[(size_t)i * w + j] = sqrtf(lx * lx + ly * ly) > 0.5f;
            }
        std::vector<float> out = fb;
        for (int y = 0; y + 1 < h; ++y)
            for (int x = 0; x + 1 < w; ++x) {
                if (!flag[(size_t)y * w + x]) continue;
                flagged++;
                const float offs[4][2] = { { 0.25f, 0.25f }, { 0.25f, 0.75f }, { 0.75f, 0.25f }, { 0.75f, 0.75f } };
                V3 c = mk(0, 0, 0);
                for (int k = 0; k < 4; ++k)
                    c = c + castRay(hs, hs.sc.camPos, cameraDir(hs.sc, (float)x + offs[k][0], (float)y + offs[k][1]), 0);
                c = c / 4.0f;
                float* p = &out[((size_t)y * w + x) * 3];
                p[0] = c.x; p[1] = c.y; p[2] = c.z;
            }
        fb.swap(out);
    }
    if (final) std::copy(fb.begin(), fb.end(), final);
    if (counters) { counters[0] = hs.rays; counters[1] = hs.boxTests; counters[2] = hs.triTests; counters[3] = flagged; }
    return 0;
}


