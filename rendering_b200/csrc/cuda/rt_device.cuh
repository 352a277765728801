// rt_device.cuh — per-ray arithmetic of the CUDA backend.
//
// Every function here is the device-side statement of one reference routine (cited per function).
// The float pipeline is reproduced operation by operation: sums associate left to right, nothing
// is contracted to FMA (the library is built with -fmad=false), divisions and square roots are the
// IEEE-rounded ones, `normalize` goes through double exactly like the reference's unqualified
// sqrt() does (include/geometry.h:99-112), comparisons keep their NaN behaviour.
//
// The functions are __host__ __device__ so tests/ can also compile this header with g++ and check
// the arithmetic on the CPU box (tests/shim); the product only ever calls them from kernels.
#pragma once

#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#define RT_HD_COLD __host__ __device__ __noinline__     // once-per-ray routines with long bodies: one copy, called (see normalize)
#else
#define RT_HD inline
#define RT_HD_COLD inline
// plain-C++ stand-ins for the CUDA vector types (tests/shim compiles this header with g++)
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
#endif

namespace rt {

struct V3 { float x, y, z; };
struct V2 { float x, y; };

RT_HD V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_HD V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_HD V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
RT_HD V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_HD V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
RT_HD V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
RT_HD float dot(V3 a, V3 b) { float s = a.x * b.x; s = s + a.y * b.y; s = s + a.z * b.z; return s; }
RT_HD V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// std::max / std::min as the reference calls them (second operand wins only on strict compare)
RT_HD float maxf_(float a, float b) { return (a < b) ? b : a; }
RT_HD float minf_(float a, float b) { return (b < a) ? b : a; }
RT_HD float clampf_(float lo, float hi, float v) { return maxf_(lo, minf_(hi, v)); }   // include/util.h:26-29

// Vec3::length / normalize (include/geometry.h:99-112): double sqrt, double reciprocal.
// On the device the double-precision square root / division sequences (~100 instructions each) are kept OUT OF LINE:
// they run a few times per ray, and inlining them at every call site made the fused kernels' code (120 KB) thrash the
// instruction cache (ncu: no-instruction stalls, profiles/r2_experiments).
#if defined(__CUDA_ARCH__) && !defined(RTB_INLINE_FP64)
__device__ __noinline__ float rt_sqrtViaDouble(float l2) { return (float)sqrt((double)l2); }
__device__ __noinline__ float rt_rsqrtViaDouble(float l2) { return (float)(1.0 / sqrt((double)l2)); }
// PointLight::illuminate / area lights: min(1, intensity / (4 pi len2 / 1000)) evaluated in double (lights.cpp:34)
__device__ __noinline__ float rt_attenuation(float intensity, float len2)
{
    const double att = (double)intensity / (4 * 3.14159265358979323846 * (double)len2 / 1000);
    const float a = (float)att;
    return (a < 1.0f) ? a : 1.0f;      // std::min(1.0f, a): a wins only on strict compare
}
#define RT_NOINLINE_DEVICE __device__ __noinline__
#else
RT_HD float rt_sqrtViaDouble(float l2) { return (float)sqrt((double)l2); }
RT_HD float rt_rsqrtViaDouble(float l2) { return (float)(1.0 / sqrt((double)l2)); }
RT_HD float rt_attenuation(float intensity, float len2)
{
    const double att = (double)intensity / (4 * 3.14159265358979323846 * (double)len2 / 1000);
    const float a = (float)att;
    return (a < 1.0f) ? a : 1.0f;
}
#define RT_NOINLINE_DEVICE inline
#endif
RT_HD float length(V3 v) { return rt_sqrtViaDouble(dot(v, v)); }
RT_HD V3 normalize(V3 v)
{
    const float l2 = dot(v, v);
    if (l2 > 0) {
        const float k = rt_rsqrtViaDouble(l2);
        v.x *= k; v.y *= k; v.z *= k;
    }
    return v;
}

RT_HD float bitsToFloat(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { uint32_t u; float f; } c; c.u = u; return c.f;
#endif
}
// `float < 1e-8` with a DOUBLE literal (objects.cpp:76,79,810): true exactly for floats below the
// first float above 1e-8, which is 0x322BCC78
#define RT_EPS_1E8 (rt::bitsToFloat(0x322BCC78u))

// ------------------------------------------------------------------------------------------------
// device-resident scene description (built by rtb_create from the ABI's RtbScene)
// ------------------------------------------------------------------------------------------------
struct Image {
    unsigned long long tex;      // cudaTextureObject_t: RGBA8, point sampled, unnormalised coords
    const unsigned char* rgba;   // same texels, linear RGBA8 (host shim / linear fallback for huge maps)
    int w, h;
};

struct Object {
    int type, material;
    V3 color;
    float ior, ambient, diffuse, specular, nSpecular;
    V3 pos;
    float r2;
    V3 normal;
    int mesh;
};

struct Light {
    int type;
    V3 color;
    float intensity;
    V3 v;
    int pointOffset, pointCount;
};

// Reference-tree node, 32 B: a = (lo.x, lo.y, lo.z, hi.x), b = (hi.y, hi.z, link, count) where link
// is the right child's index for inner nodes (count < 0) or the first reference slot of a leaf.
struct Node { float lox, loy, loz, hix, hiy, hiz; int link; int count; };

// One triangle reference slot, 48 B, stored in the reference's leaf order: v0, e1 = v1-v0,
// e2 = v2-v0 (the same subtractions rayTriangleIntersect performs per test, objects.cpp:70-71).
struct TriSlot { float v0x, v0y, v0z; int tri; float e1x, e1y, e1z; int pad0; float e2x, e2y, e2z; int pad1; };

struct Mesh {
    const Node* nodes;
    const TriSlot* slots;
    const float* nrm;   // 9 per triangle
    const float* uv;    // 6 per triangle
    const float* tan;   // 6 per triangle
    Image diffuse, normal, specular;
    int nNodes, nSlots, nTris, maxDepth;
    // fast path (bvh_build.h): search BVH over unique triangles + eligibility tables
    const float4* bvhNodes;    // 4 x float4 per node
    const float4* bvhTris;     // 3 x float4 per triangle, BVH leaf order
    const int* triRefOff;      // nTris+1: range of a triangle's references in triRefs
    const int2* triRefs;       // (reference-tree leaf node, reference slot), ascending slot
    const int* parent;         // reference-tree parent of every node, -1 at the root
};

struct Scene {
    int width, height;
    float bias;
    int maxRayDepth;
    V3 background;
    unsigned flags;
    V3 camPos;
    float camM[16];
    float camScale, camAspect;
    int nObjects, nLights, nMeshes;
    int shadowRaysPerHit;          // sum over lights of (area ? pointCount : 1)
    const Object* objects;
    const Light* lights;
    const Mesh* meshes;
    const float* areaPoints;
    Image sky[6];
};

enum { FLAG_CULL = 1u, FLAG_USE_AC = 2u, FLAG_SKYBOX = 4u, FLAG_SHOW_NORMALS = 8u, FLAG_SSAA = 16u };
enum { OBJ_SPHERE = 1, OBJ_PLANE = 2, OBJ_MESH = 3 };
enum { MAT_DIFFUSE = 0, MAT_REFLECTIVE = 1, MAT_TRANSPARENT = 2, MAT_PHONG = 3 };
enum { LIGHT_DISTANT = 1, LIGHT_POINT = 2, LIGHT_AREA = 3 };

// ------------------------------------------------------------------------------------------------
// camera: renderWorker's getPixels lambda + Camera::getRay (scene.cpp:453-457, 19-54)
// ------------------------------------------------------------------------------------------------
RT_HD V3 mulRowVecMatrix(const float* m, V3 s, bool wZeroColumn)
{
    // Matrix44::multVecMatrix (include/geometry.h:289-307)
    float o[4];
    for (int j = 0; j < 4; ++j) {
        float a = s.x * m[0 * 4 + j];
        a = a + s.y * m[1 * 4 + j];
        a = a + s.z * m[2 * 4 + j];
        a = a + m[3 * 4 + j];
        o[j] = a;
    }
    V3 d = mk(o[0], o[1], o[2]);
    const float w = o[3];
    (void)wZeroColumn;
    if (w != 0.0f && w != 1.0f) {
        const float wi = 1.0f / w;
        d.x *= wi; d.y *= wi; d.z *= wi;
    }
    return d;
}

// px, py are the values handed to getPixels: (float)x + 0.5f for pass 1, +0.25f / +0.75f for SSAA
RT_HD V3 cameraDir(const Scene& sc, float px, float py)
{
    const float W = (float)sc.width, H = (float)sc.height;
    const float xPix = (2 * (px + 0.5f) / W - 1) * sc.camScale * sc.camAspect;
    const float yPix = -(2 * (py + 0.5f) / H - 1) * sc.camScale;
    return mulRowVecMatrix(sc.camM, normalize(mk(xPix, yPix, -1.0f)), false);
}

// ------------------------------------------------------------------------------------------------
// primitives
// ------------------------------------------------------------------------------------------------
struct RayCtx {   // ray plus the per-ray invariants intersectBox recomputes per call (objects.cpp:543-544)
    V3 o, d, inv;
    int sx, sy, sz;
};
RT_HD RayCtx makeRay(V3 o, V3 d)
{
    RayCtx r;
    r.o = o; r.d = d;
    r.inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    r.sx = r.inv.x < 0; r.sy = r.inv.y < 0; r.sz = r.inv.z < 0;
    return r;
}

// AccelerationStructure::intersectBox (objects.cpp:534-570): ray LINE against the slab box, no t
// range, comparisons written so NaNs fall through exactly as in the reference.
RT_HD bool lineHitsBox(const RayCtx& r, float lox, float loy, float loz, float hix, float hiy, float hiz)
{
    float tmin = ((r.sx ? hix : lox) - r.o.x) * r.inv.x;
    float tmax = ((r.sx ? lox : hix) - r.o.x) * r.inv.x;
    const float tymin = ((r.sy ? hiy : loy) - r.o.y) * r.inv.y;
    const float tymax = ((r.sy ? loy : hiy) - r.o.y) * r.inv.y;
    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    const float tzmin = ((r.sz ? hiz : loz) - r.o.z) * r.inv.z;
    const float tzmax = ((r.sz ? loz : hiz) - r.o.z) * r.inv.z;
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    return true;
}

// Same test, additionally reporting whether any slab product was NaN (0 * inf: a zero direction
// component with the origin exactly on a box plane).  NaNs make the reference's comparisons
// vacuous, which is the only way a child box can pass while an ancestor fails.
RT_HD bool lineHitsBoxNaN(const RayCtx& r, float lox, float loy, float loz, float hix, float hiy, float hiz, bool& sawNaN)
{
    float tmin = ((r.sx ? hix : lox) - r.o.x) * r.inv.x;
    float tmax = ((r.sx ? lox : hix) - r.o.x) * r.inv.x;
    const float tymin = ((r.sy ? hiy : loy) - r.o.y) * r.inv.y;
    const float tymax = ((r.sy ? loy : hiy) - r.o.y) * r.inv.y;
    const float tzmin = ((r.sz ? hiz : loz) - r.o.z) * r.inv.z;
    const float tzmax = ((r.sz ? loz : hiz) - r.o.z) * r.inv.z;
    sawNaN = (tmin != tmin) | (tmax != tmax) | (tymin != tymin) | (tymax != tymax) | (tzmin != tzmin) | (tzmax != tzmax);
    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    return true;
}

// Triangle::rayTriangleIntersect (objects.cpp:59-95), Moller-Trumbore with the reference's reject order
RT_HD bool hitTriangle(const RayCtx& r, V3 v0, V3 e1, V3 e2, bool cull, float& t, float& u, float& v)
{
    const V3 pvec = cross(r.d, e2);
    const float det = dot(e1, pvec);
    if (cull && det < RT_EPS_1E8) return false;
    if (fabsf(det) < RT_EPS_1E8) return false;
    const float invDet = 1 / det;
    const V3 tvec = r.o - v0;
    const float uu = dot(tvec, pvec) * invDet;
    if (uu < 0 || uu > 1) return false;
    const V3 qvec = cross(tvec, e1);
    const float vv = dot(r.d, qvec) * invDet;
    if (vv < 0 || uu + vv > 1) return false;
    const float tt = dot(e2, qvec) * invDet;
    if (tt < 0) return false;
    t = tt; u = uu; v = vv;
    return true;
}

// Sphere::intersectObject (objects.cpp:774-786)
RT_HD bool hitSphere(const RayCtx& r, V3 c, float r2, float& t)
{
    const V3 L = c - r.o;
    const float tca = dot(L, r.d);
    const float d2 = dot(L, L) - tca * tca;
    if (d2 > r2) return false;
    const float thc = sqrtf(r2 - d2);
    float t0 = tca - thc;
    const float t1 = tca + thc;
    if (t0 < 0) t0 = t1;
    if (t0 < 0) return false;
    t = t0;
    return true;
}

// Plane::intersectObject (objects.cpp:807-814)
RT_HD bool hitPlane(const RayCtx& r, V3 p, V3 n, float& t)
{
    const float denom = dot(r.d, n);
    if (fabsf(denom) < RT_EPS_1E8) return false;
    const float t0 = dot(p - r.o, n) / denom;
    t = t0;
    return (t0 >= 0);
}

// ------------------------------------------------------------------------------------------------
// textures, skybox
// ------------------------------------------------------------------------------------------------
RT_HD void fetchTexel(const Image& im, int x, int y, float& r, float& g, float& b)
{
#if defined(__CUDA_ARCH__)
    if (im.tex) {
        const uchar4 t = tex2D<uchar4>((cudaTextureObject_t)im.tex, (float)x + 0.5f, (float)y + 0.5f);
        r = (float)t.x; g = (float)t.y; b = (float)t.z;
    } else {
        const uchar4 t = reinterpret_cast<const uchar4*>(im.rgba)[(size_t)y * im.w + x];
        r = (float)t.x; g = (float)t.y; b = (float)t.z;
    }
#else
    const unsigned char* t = im.rgba + ((size_t)y * im.w + x) * 4;
    r = (float)t[0]; g = (float)t[1]; b = (float)t[2];
#endif
    r /= 256; g /= 256; b /= 256;   // the reference divides by 256, not 255 (objects.cpp:409, scene.cpp:354)
}

// texel addressing of getDiffuseColor / getSpecularValue / getSurfaceData (objects.cpp:144-147,
// 156-159): truncation, upper clamp only.  Negative / NaN coordinates index out of bounds in the
// reference (undefined behaviour); they are clamped to 0 here.
RT_HD int texIndex(int size, float coord)
{
    const float f = size * coord;
    int i = (f >= 2147483648.0f) ? size : ((f > -2147483648.0f) ? (int)f : 0);
    if (i >= size) i = size - 1;
    if (i < 0) i = 0;
    return i;
}

// Scene::getSkybox (scene.cpp:381-442)
RT_HD int skyPixel(float v, int size)
{
    const float f = (v + 1.0f) / 2.0f * size;
    int i = (f >= 2147483648.0f) ? size : ((f > -2147483648.0f) ? (int)f : 0);
    if (i >= size) i = size - 1;
    if (i < 0) i = 0;   // reference: out-of-bounds read for directions with a NaN component
    return i;
}
RT_HD V3 skybox(const Scene& sc, V3 dir)
{
    if (!(sc.flags & FLAG_SKYBOX)) return sc.background;
    const float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
    const float m = maxf_(ax, maxf_(ay, az));
    int face, i, j;
    const int W = sc.sky[0].w, H = sc.sky[0].h;
    if (m == az) {
        if (dir.z < 0) { const V3 a = dir * (1 / -dir.z); i = skyPixel(a.y, H); j = skyPixel(a.x, W); face = 1; }
        else           { const V3 a = dir * (1 / dir.z);  i = skyPixel(a.y, H); j = skyPixel(-a.x, W); face = 3; }
    } else if (m == ax) {
        if (dir.x < 0) { const V3 a = dir * (1 / -dir.x); i = skyPixel(a.y, H); j = skyPixel(-a.z, W); face = 0; }
        else           { const V3 a = dir * (1 / dir.x);  i = skyPixel(a.y, H); j = skyPixel(a.z, W); face = 2; }
    } else {
        if (dir.y < 0) { const V3 a = dir * (1 / -dir.y); i = skyPixel(a.z, H); j = skyPixel(a.x, W); face = 5; }
        else           { const V3 a = dir * (1 / dir.y);  i = skyPixel(a.z, H); j = skyPixel(a.x, W); face = 4; }
    }
    V3 c;
    fetchTexel(sc.sky[face], j, i, c.x, c.y, c.z);
    return c;
}

// ------------------------------------------------------------------------------------------------
// surface data
// ------------------------------------------------------------------------------------------------
struct Surface {
    V3 P, N, color;
    float specCoef;
};

// Object::getSurfaceData + getDiffuseColor + getSpecularValue for all three object kinds
// (objects.cpp:121-175, 788-796, 816-824; scene.cpp:768-775, 849-851)
RT_HD Surface surfaceAt(const Scene& sc, const Object& ob, V3 orig, V3 dir, float t, float u, float v, int tri)
{
    Surface s;
    s.P = orig + dir * t;
    s.color = ob.color;
    s.specCoef = ob.specular;
    if (ob.type == OBJ_SPHERE) {
        s.N = normalize(s.P - ob.pos);
    } else if (ob.type == OBJ_PLANE) {
        s.N = ob.normal;
    } else {
        const Mesh& me = sc.meshes[ob.mesh];
        const float* n = me.nrm + (size_t)tri * 9;
        const float* tc = me.uv + (size_t)tri * 6;
        const float w = 1 - u - v;
        V2 tex;
        tex.x = tc[2] * u + tc[4] * v + tc[0] * w;
        tex.y = tc[3] * u + tc[5] * v + tc[1] * w;
        const V3 nb = mk(n[3], n[4], n[5]), nc = mk(n[6], n[7], n[8]), na = mk(n[0], n[1], n[2]);
        // issue the (up to three) independent texel fetches before the dependent FP64 normalisations
        const bool shade = !(sc.flags & FLAG_SHOW_NORMALS);
        const bool wantN = me.normal.w > 0, wantD = shade && me.diffuse.w > 0, wantS = shade && ob.material == MAT_PHONG && me.specular.w > 0;
        float nr = 0, ng = 0, nbl = 0, dr = 0, dg = 0, db = 0, sr = 0, sg = 0, sb = 0;
        if (wantN) fetchTexel(me.normal, texIndex(me.normal.w, tex.x), texIndex(me.normal.h, tex.y), nr, ng, nbl);
        if (wantD) fetchTexel(me.diffuse, texIndex(me.diffuse.w, tex.x), texIndex(me.diffuse.h, tex.y), dr, dg, db);
        if (wantS) fetchTexel(me.specular, texIndex(me.specular.w, tex.x), texIndex(me.specular.h, tex.y), sr, sg, sb);
        s.N = normalize((nb * u + nc * v + na * w) / 3.0f);
        if (wantN) {
            const float* tb = me.tan + (size_t)tri * 6;
            // load-time conversion (objects.cpp:431-433) then the lookup's own normalize (:148)
            V3 tn = normalize(mk(nr * 2 - 1, -(ng * 2 - 1), nbl));
            tn = normalize(tn);
            // rows tangent, bitangent, N; fourth row and column zero (objects.cpp:133-139)
            V3 d;
            d.x = tn.x * tb[0]; d.x = d.x + tn.y * tb[3]; d.x = d.x + tn.z * s.N.x; d.x = d.x + 0.0f;
            d.y = tn.x * tb[1]; d.y = d.y + tn.y * tb[4]; d.y = d.y + tn.z * s.N.y; d.y = d.y + 0.0f;
            d.z = tn.x * tb[2]; d.z = d.z + tn.y * tb[5]; d.z = d.z + tn.z * s.N.z; d.z = d.z + 0.0f;
            float wq = tn.x * 0.0f; wq = wq + tn.y * 0.0f; wq = wq + tn.z * 0.0f; wq = wq + 0.0f;
            if (wq != 0.0f && wq != 1.0f) { const float wi = 1.0f / wq; d.x *= wi; d.y *= wi; d.z *= wi; }
            s.N = normalize(d);
        }
        if (wantD) s.color = mk(dr, dg, db);
        if (wantS) s.specCoef = (sr + sg + sb) / 3.0f;   // objects.cpp:455
    }
    return s;
}

// ------------------------------------------------------------------------------------------------
// lights and Fresnel optics
// ------------------------------------------------------------------------------------------------
// PointLight::illuminate / DistantLight::illuminate (lights.cpp:18-38).  L is the direction FROM the
// light; the shadow ray travels along -L up to `dist`.
RT_HD void illuminate(const Light& li, V3 P, V3& L, V3& I, float& dist)
{
    if (li.type == LIGHT_DISTANT) {
        L = li.v;
        I = li.color * li.intensity;
        dist = FLT_MAX;
    } else {
        L = P - li.v;
        I = li.color * rt_attenuation(li.intensity, dot(L, L));
        L = normalize(L);
        dist = length(P - li.v);
    }
}
// intensity of an area light at P (scene.cpp:795, 831, 874, 924)
RT_HD V3 areaIntensity(const Light& li, V3 P)
{
    const V3 d = P - li.v;
    return li.color * rt_attenuation(li.intensity, dot(d, d));
}

RT_HD V3 reflect(V3 dir, V3 n) { return dir - n * (2 * dot(dir, n)); }   // scene.cpp:672-675

RT_HD V3 refract(V3 dir, V3 n, float ior)   // scene.cpp:677-696
{
    float n1 = 1, n2 = ior;
    float cosi = clampf_(-1, 1, dot(dir, n));
    V3 mn = n;
    if (cosi < 0) cosi = -cosi;
    else { const float tmp = n1; n1 = n2; n2 = tmp; mn = -n; }
    const float rri = n1 / n2;
    const float k = 1 - rri * rri * (1 - cosi * cosi);
    if (k < 0) return mk(0.0f, 0.0f, 0.0f);
    return dir * rri + mn * (rri * cosi - sqrtf(k));
}

RT_HD float fresnel(V3 dir, V3 n, float ior)   // scene.cpp:698-722
{
    float n1 = 1, n2 = ior;
    float cosi = clampf_(-1, 1, dot(dir, n));
    if (cosi > 0) { const float tmp = n1; n1 = n2; n2 = tmp; }
    const float sint = n1 / n2 * sqrtf(maxf_(0.f, 1 - cosi * cosi));
    if (sint >= 1) return 1;
    const float cost = sqrtf(maxf_(0.f, 1 - sint * sint));
    cosi = fabsf(cosi);
    const float rs = ((n2 * cosi) - (n1 * cost)) / ((n2 * cosi) + (n1 * cost));
    const float rp = ((n1 * cosi) - (n2 * cost)) / ((n1 * cosi) + (n2 * cost));
    return (rs * rs + rp * rp) / 2;
}

// Pixel bounds of the projection of a world-space box (lo.xyz, hi.xyz): the inverse of cameraDir for the box's 8 corners, in
// double.  false when a corner is not finite or at / behind the camera plane (no bound).  One definition for the host's
// rectangle (scene_pack.h primaryRect), the device's coverage bitmap (k_cover_mark) and the CPU tests of both.
RT_HD bool pixelBoundsOfBox(const Scene& sc, const float* b, double& minX, double& maxX, double& minY, double& maxY)
{
    minX = 1e300; maxX = -1e300; minY = 1e300; maxY = -1e300;
    for (int c = 0; c < 8; ++c) {
        const double v0 = (double)b[(c & 1) ? 3 : 0] - sc.camPos.x, v1 = (double)b[(c & 2) ? 4 : 1] - sc.camPos.y, v2 = (double)b[(c & 4) ? 5 : 2] - sc.camPos.z;
        if (!(isfinite(v0) && isfinite(v1) && isfinite(v2))) return false;
        // world direction = camera direction (row vector) x rMatrix  =>  camera = world x rMatrix^T
        double cam[3];
        for (int i = 0; i < 3; ++i) cam[i] = v0 * sc.camM[i * 4 + 0] + v1 * sc.camM[i * 4 + 1] + v2 * sc.camM[i * 4 + 2];
        const double len = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
        if (!(cam[2] < -1e-4 * len)) return false;                   // at or behind the camera plane: no bound
        const double xPix = cam[0] / -cam[2], yPix = cam[1] / -cam[2];
        // renderWorker (scene.cpp:453-461): xPix = (2 (x + 1.0) / W - 1) scale aspect, yPix = -(2 (y + 1.0) / H - 1) scale
        const double px = (xPix / ((double)sc.camScale * sc.camAspect) + 1.0) * sc.width / 2.0 - 1.0;
        const double py = (-yPix / (double)sc.camScale + 1.0) * sc.height / 2.0 - 1.0;
        if (!(isfinite(px) && isfinite(py))) return false;
        minX = fmin(minX, px); maxX = fmax(maxX, px);
        minY = fmin(minY, py); maxY = fmax(maxY, py);
    }
    return true;
}

#if defined(__CUDA_ARCH__)
#define RT_LDG(p) __ldg(p)
#else
#define RT_LDG(p) (*(p))
#endif
// ------------------------------------------------------------------------------------------------
// powf, bit for bit
// ------------------------------------------------------------------------------------------------
// The reference's std::pow(float, float) (scene.cpp:824,846,867,887,917,937; :565) is glibc's powf, a THIRD-PARTY
// routine absent from /root/reference: glibc 2.39 (Ubuntu 24.04, the image both boxes run), sysdeps/ieee754/flt-32/
// e_powf.c + e_powf_log2_data.c + math/e_exp2f_data.c (Szabolcs Nagy's algorithm from ARM optimized-routines).
// It is NOT correctly rounded (0.82 ULP: log2(x) from a 16-entry table and a degree-5 polynomial, 2^(y log2 x) from a
// 32-entry table and a cubic, all in double), so a correctly rounded pow differs from it by one ulp on ~0.1 % of
// inputs.  This is a restatement of the published algorithm with the table constants of that glibc build and the
// operation grouping of the variant x86-64 selects on CPUs with FMA (__powf_fma: every a*b+c below is one fused
// operation — read off the library's disassembly; tests/test_powf.py pins it against the host's powf on millions of inputs).
#define RT_POWF_LOG2_TAB { \
    0x3ff661ec79f8f3beULL, 0xbfdefec65b963019ULL, 0x3ff571ed4aaf883dULL, 0xbfdb0b6832d4fca4ULL, \
    0x3ff49539f0f010b0ULL, 0xbfd7418b0a1fb77bULL, 0x3ff3c995b0b80385ULL, 0xbfd39de91a6dcf7bULL, \
    0x3ff30d190c8864a5ULL, 0xbfd01d9bf3f2b631ULL, 0x3ff25e227b0b8ea0ULL, 0xbfc97c1d1b3b7af0ULL, \
    0x3ff1bb4a4a1a343fULL, 0xbfc2f9e393af3c9fULL, 0x3ff12358f08ae5baULL, 0xbfb960cbbf788d5cULL, \
    0x3ff0953f419900a7ULL, 0xbfaa6f9db6475fceULL, 0x3ff0000000000000ULL, 0x0000000000000000ULL, \
    0x3fee608cfd9a47acULL, 0x3fb338ca9f24f53dULL, 0x3feca4b31f026aa0ULL, 0x3fc476a9543891baULL, \
    0x3feb2036576afce6ULL, 0x3fce840b4ac4e4d2ULL, 0x3fe9c2d163a1aa2dULL, 0x3fd40645f0c6651cULL, \
    0x3fe886e6037841edULL, 0x3fd88e9c2c1b9ff8ULL, 0x3fe767dcf5534862ULL, 0x3fdce0a44eb17bccULL }
#define RT_EXP2F_TAB { \
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, \
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, \
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL, \
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, \
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, \
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, \
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, \
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL }
static const unsigned long long kPowfLog2TabHost[32] = RT_POWF_LOG2_TAB;   // (invc, logc) pairs
static const unsigned long long kExp2fTabHost[32] = RT_EXP2F_TAB;
#if defined(__CUDACC__)
static __device__ const unsigned long long kPowfLog2TabDev[32] = RT_POWF_LOG2_TAB;
static __device__ const unsigned long long kExp2fTabDev[32] = RT_EXP2F_TAB;
#endif

RT_HD double bitsToDouble(unsigned long long u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    union { unsigned long long u; double d; } c; c.u = u; return c.d;
#endif
}
RT_HD unsigned long long doubleBits(double d)
{
#if defined(__CUDA_ARCH__)
    return (unsigned long long)__double_as_longlong(d);
#else
    union { double d; unsigned long long u; } c; c.d = d; return c.u;
#endif
}
RT_HD uint32_t floatBitsU(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
// single IEEE operations, immune to contraction / re-association by either compiler
#if defined(__CUDA_ARCH__)
RT_HD double dFma(double a, double b, double c) { return __fma_rn(a, b, c); }
RT_HD double dMul(double a, double b) { return __dmul_rn(a, b); }
RT_HD double dAdd(double a, double b) { return __dadd_rn(a, b); }
#else
RT_HD double dFma(double a, double b, double c) { return fma(a, b, c); }
RT_HD double dMul(double a, double b) { return a * b; }
RT_HD double dAdd(double a, double b) { return a + b; }
#endif

// e_powf.c checkint(): 0 = y is not an integer, 1 = odd integer, 2 = even integer
RT_HD int powfCheckInt(uint32_t iy)
{
    const int e = (int)((iy >> 23) & 0xff);
    if (e < 0x7f) return 0;
    if (e > 0x7f + 23) return 2;
    if (iy & ((1u << (0x7f + 23 - e)) - 1)) return 0;
    if (iy & (1u << (0x7f + 23 - e))) return 1;
    return 2;
}

RT_HD_COLD float powfGlibc(float x, float y)
{
#if defined(__CUDA_ARCH__)
    const unsigned long long* logTab = kPowfLog2TabDev;
    const unsigned long long* expTab = kExp2fTabDev;
#else
    const unsigned long long* logTab = kPowfLog2TabHost;
    const unsigned long long* expTab = kExp2fTabHost;
#endif
    unsigned long long signBias = 0;
    uint32_t ix = floatBitsU(x);
    const uint32_t iy = floatBitsU(y);
    const bool yZeroInfNan = 2 * iy - 1 >= 2u * 0x7f800000u - 1;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || yZeroInfNan) {
        // x is subnormal, zero, negative, inf or nan — or y is zero, inf or nan
        if (yZeroInfNan) {
            if (2 * iy == 0) return (2 * (ix ^ 0x00400000u) > 2u * 0x7fc00000u) ? x + y : 1.0f;        // signalling NaN ^ 0
            if (ix == 0x3f800000u) return (2 * (iy ^ 0x00400000u) > 2u * 0x7fc00000u) ? x + y : 1.0f;
            if (2 * ix > 2u * 0x7f800000u || 2 * iy > 2u * 0x7f800000u) return x + y;
            if (2 * ix == 2 * 0x3f800000u) return 1.0f;
            if ((2 * ix < 2 * 0x3f800000u) == !(iy & 0x80000000u)) return 0.0f;   // |x| < 1 && y == inf, |x| > 1 && y == -inf
            return y * y;
        }
        if (2 * ix - 1 >= 2u * 0x7f800000u - 1) {   // x is zero, inf or nan
            float x2 = x * x;
            if ((ix & 0x80000000u) && powfCheckInt(iy) == 1) x2 = -x2;
            return (iy & 0x80000000u) ? 1 / x2 : x2;
        }
        if (ix & 0x80000000u) {   // finite x < 0
            const int yint = powfCheckInt(iy);
            if (yint == 0) return (x - x) / (x - x);
            if (yint == 1) signBias = 1ull << (5 + 11);   // SIGN_BIAS = 1 << (EXP2F_TABLE_BITS + 11)
            ix &= 0x7fffffffu;
        }
        if (ix < 0x00800000u) {   // subnormal x: normalise so the exponent becomes negative
            ix = floatBitsU(bitsToFloat(ix) * 8388608.0f);
            ix &= 0x7fffffffu;
            ix -= 23u << 23;
        }
    }
    // log2_inline: x = 2^k z, z in [OFF, 2 OFF); log2(x) = k + log2(c) + log2(z / c) with c near the centre of z's sub-interval
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> (23 - 4)) % 16);
    const uint32_t top = tmp & 0xff800000u;
    const uint32_t iz = ix - top;
    const int k = (int)top >> 23;   // arithmetic shift
    const double invc = bitsToDouble(RT_LDG(logTab + 2 * i)), logc = bitsToDouble(RT_LDG(logTab + 2 * i + 1));
    const double z = (double)bitsToFloat(iz);
    const double A0 = bitsToDouble(0x3fd27616c9496e0bULL), A1 = bitsToDouble(0xbfd71969a075c67aULL), A2 = bitsToDouble(0x3fdec70a6ca7baddULL),
                 A3 = bitsToDouble(0xbfe7154748bef6c8ULL), A4 = bitsToDouble(0x3ff71547652ab82bULL);
    const double r = dFma(z, invc, -1.0);
    const double y0 = dAdd(logc, (double)k);
    const double r2 = dMul(r, r);
    const double yy = dFma(A0, r, A1);
    const double p = dFma(A2, r, A3);
    const double r4 = dMul(r2, r2);
    double q = dFma(A4, r, y0);
    q = dFma(p, r2, q);
    const double logx = dFma(yy, r4, q);
    const double ylogx = dMul((double)y, logx);   // cannot overflow: y is single precision
    if (((doubleBits(ylogx) >> 47) & 0xffff) >= (0x405f800000000000ULL >> 47)) {   // |y log2 x| >= 126
        const float sgn = signBias ? -1.0f : 1.0f;
        if (ylogx > bitsToDouble(0x405fffffffd1d571ULL)) return sgn * bitsToFloat(0x7f800000u);   // > 0x1.fffffffd1d571p+6: overflow
        if (ylogx <= -150.0) return sgn * 0.0f;                                                              // underflow
        if (ylogx < -149.0) return sgn * bitsToFloat(1u);   // __math_may_uflowf: 0x1.4p-75f squared, rounded to nearest
    }
    // exp2_inline: 2^x = 2^(k/32) 2^r with r in [-1/64, 1/64]
    const double shift = bitsToDouble(0x42e8000000000000ULL);   // 0x1.8p+52 / 32
    double kd = dAdd(ylogx, shift);
    const unsigned long long ki = doubleBits(kd);
    kd = dAdd(kd, -shift);
    const double rr = dAdd(ylogx, -kd);
    unsigned long long t = RT_LDG(expTab + (ki % 32));
    t += (ki + signBias) << (52 - 5);
    const double s = bitsToDouble(t);
    const double C0 = bitsToDouble(0x3fac6af84b912394ULL), C1 = bitsToDouble(0x3fcebfce50fac4f3ULL), C2 = bitsToDouble(0x3fe62e42ff0c52d6ULL);
    const double zz = dFma(C0, rr, C1);
    const double rr2 = dMul(rr, rr);
    double e = dFma(C2, rr, 1.0);
    e = dFma(zz, rr2, e);
    e = dMul(e, s);
    return (float)e;
}
RT_HD float powExact(float x, float y) { return powfGlibc(x, y); }

RT_HD int floatBits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    union { float f; int i; } c; c.f = f; return c.i;
#endif
}

// ------------------------------------------------------------------------------------------------
// fast path: search BVH + exact eligibility
// ------------------------------------------------------------------------------------------------
// What the reference tree computes for a ray is: among all (leaf, triangle) references whose leaf is
// ELIGIBLE — the ray line passes lineHitsBox for the leaf and every ancestor — the hit with the
// smallest t, ties going to the smallest reference slot (DFS order, strict `<`).  Boxes are nested
// (children are the parent cut at a plane) and the slab arithmetic is monotone, so a leaf that
// passes WITHOUT producing a NaN implies all its ancestors pass; only NaN cases walk the parents.
// Returns the smallest eligible reference slot of triangle `tri`, or -1.
RT_HD int eligibleSlot(const Scene& sc, const Mesh& me, const RayCtx& r, int tri)
{
    const int b = RT_LDG(me.triRefOff + tri), e = RT_LDG(me.triRefOff + tri + 1);
    if (!(sc.flags & FLAG_USE_AC)) return b < e ? RT_LDG(me.triRefs + b).y : -1;
    const float4* nodes = reinterpret_cast<const float4*>(me.nodes);
    for (int i = b; i < e; ++i) {
        const int2 ref = RT_LDG(me.triRefs + i);
        const float4 n0 = RT_LDG(nodes + 2 * ref.x), n1 = RT_LDG(nodes + 2 * ref.x + 1);
        bool nan;
        if (!lineHitsBoxNaN(r, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, nan)) continue;
        bool ok = true;
        if (nan) {
            for (int p = RT_LDG(me.parent + ref.x); p >= 0; p = RT_LDG(me.parent + p)) {
                const float4 a0 = RT_LDG(nodes + 2 * p), a1 = RT_LDG(nodes + 2 * p + 1);
                if (!lineHitsBox(r, a0.x, a0.y, a0.z, a0.w, a1.x, a1.y)) { ok = false; break; }
            }
        }
        if (ok) return ref.y;
    }
    return -1;
}

// conservative ray-segment / padded-box test for CULLING only (never decides a hit): NaN-ignoring
// min/max keep a NaN axis from rejecting
RT_HD float slabEntry(const RayCtx& r, float lox, float loy, float loz, float hix, float hiy, float hiz, float tFar, bool& hit)
{
    const float t0x = (lox - r.o.x) * r.inv.x, t1x = (hix - r.o.x) * r.inv.x;
    const float t0y = (loy - r.o.y) * r.inv.y, t1y = (hiy - r.o.y) * r.inv.y;
    const float t0z = (loz - r.o.z) * r.inv.z, t1z = (hiz - r.o.z) * r.inv.z;
    const float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
    const float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tFar));
    hit = tn <= tf;
    return tn;
}

// Closest hit (ANY = false) or occlusion test against tLimit (ANY = true) of one mesh.
template <bool ANY>
RT_HD bool walkMeshFast(const Scene& sc, const Mesh& me, const RayCtx& r, int* stack, int stackStride, float tLimit,
    float& tBest, float& uBest, float& vBest, int& triBest)
{
    if (me.nNodes == 0 || me.nTris == 0) return false;
    const bool cull = sc.flags & FLAG_CULL;
    bool found = false;
    int slotBest = 0x7fffffff;
    int sp = 0;
    int cur = 0;   // >= 0 inner node, < 0 leaf
    for (;;) {
        if (cur >= 0) {
            const float4* nd = me.bvhNodes + (size_t)cur * 4;
            const float4 a = RT_LDG(nd), b = RT_LDG(nd + 1), c = RT_LDG(nd + 2), d = RT_LDG(nd + 3);
            const float tFar = ANY ? tLimit : tBest;
            bool h0, h1;
            const float e0 = slabEntry(r, a.x, a.y, a.z, a.w, b.x, b.y, tFar, h0);
            const float e1 = slabEntry(r, b.z, b.w, c.x, c.y, c.z, c.w, tFar, h1);
            const int c0 = floatBits(d.x), c1 = floatBits(d.y);
            if (h0 && h1) {
                const bool swap = e1 < e0;
                stack[sp * stackStride] = swap ? c0 : c1;
                sp++;
                cur = swap ? c1 : c0;
                continue;
            }
            if (h0) { cur = c0; continue; }
            if (h1) { cur = c1; continue; }
        } else {
            const int code = ~cur;
            const int first = code >> 3, count = (code & 7) + 1;
            const float4* tp = me.bvhTris + (size_t)first * 3;
            for (int k = 0; k < count; ++k, tp += 3) {
                const float4 p0 = RT_LDG(tp), p1 = RT_LDG(tp + 1), p2 = RT_LDG(tp + 2);
                float t, u, v;
                if (!hitTriangle(r, mk(p0.x, p0.y, p0.z), mk(p1.x, p1.y, p1.z), mk(p2.x, p2.y, p2.z), cull, t, u, v)) continue;
                const int tri = floatBits(p0.w);
                if (ANY) {
                    if (t < tLimit && eligibleSlot(sc, me, r, tri) >= 0) return true;
                } else if (t < tBest || (found && t == tBest)) {
                    const int slot = eligibleSlot(sc, me, r, tri);
                    if (slot >= 0 && (t < tBest || slot < slotBest)) {
                        tBest = t; uBest = u; vBest = v; triBest = tri; slotBest = slot; found = true;
                    }
                }
            }
        }
        if (sp == 0) break;
        sp--;
        cur = stack[sp * stackStride];
    }
    return found;
}

} // namespace rt
