// flatten.cpp — turns the loaded Scene into the structure-of-arrays RtbScene the CUDA backend
// uploads (SURVEY.md 8b "Inputs read").  Host-only values that need libm (tanf, sinf, cosf) are
// evaluated here, exactly where the reference evaluates them on the CPU (scene.cpp:24-48,447-448).
#include "flatten.h"

#include <algorithm>
#include <cmath>

namespace rtb {

static RtbImage imageOf(const TextureRGB8& t, bool loaded)
{
    RtbImage im{};
    if (loaded && t.loaded()) {
        im.rgb = t.rgb.data();
        im.width = t.width;
        im.height = t.height;
    }
    return im;
}

static void put3(float* d, const Vec3f& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

RtbCamera flattenCamera(const Camera& camera, size_t width, size_t height)
{
    RtbCamera c{};
    put3(c.pos, camera.pos);
    const Matrix44f r = camera.rotationMatrix();
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) c.rMatrix[i * 4 + j] = r[i][j];
    c.scale = tanf(camera.fov * 0.5f / 180.0f * (float)(3.14159265358979323846));   // scene.cpp:447
    c.aspect = (width) / (float)height;                                              // scene.cpp:448
    return c;
}

void flatten(const Scene& scene, FlatScene& out)
{
    out = FlatScene{};
    RtbScene& s = out.view;
    s.abiVersion = RTB_ABI_VERSION;
    s.width = (int32_t)scene.options.width;
    s.height = (int32_t)scene.options.height;
    s.bias = scene.options.bias;
    s.maxRayDepth = scene.options.maxRayDepth;
    put3(s.backgroundColor, scene.options.backgroundColor);
    s.flags = (options::useBackfaceCulling ? RTB_FLAG_BACKFACE_CULLING : 0u) | (options::useAC ? RTB_FLAG_USE_AC : 0u)
        | (options::useSkybox ? RTB_FLAG_USE_SKYBOX : 0u) | (options::showNormals ? RTB_FLAG_SHOW_NORMALS : 0u)
        | (options::enableSSAA ? RTB_FLAG_ENABLE_SSAA : 0u);
    out.imageName = scene.options.imageName;

    s.camera = flattenCamera(scene.camera, scene.options.width, scene.options.height);

    size_t nMeshes = 0;
    for (const auto& o : scene.objects) nMeshes += (o->objectType == ObjectType::Mesh);
    out.meshStorage.resize(nMeshes);
    out.meshes.resize(nMeshes);

    int meshIndex = 0;
    for (const auto& o : scene.objects) {
        RtbObject f{};
        switch (o->materialType) {
        case MaterialType::Diffuse: f.material = RTB_MAT_DIFFUSE; break;
        case MaterialType::Reflective: f.material = RTB_MAT_REFLECTIVE; break;
        case MaterialType::Transparent: f.material = RTB_MAT_TRANSPARENT; break;
        case MaterialType::Phong: f.material = RTB_MAT_PHONG; break;
        }
        put3(f.color, o->color);
        f.ior = o->indexOfRefraction;
        f.ambient = o->ambient;
        f.diffuse = o->diffuse;
        f.specular = o->specular;
        f.nSpecular = o->nSpecular;
        put3(f.pos, o->pos);
        f.mesh = -1;
        if (o->objectType == ObjectType::Sphere) {
            f.type = RTB_OBJ_SPHERE;
            f.r2 = static_cast<const Sphere&>(*o).r2;
        } else if (o->objectType == ObjectType::Plane) {
            f.type = RTB_OBJ_PLANE;
            put3(f.normal, static_cast<const Plane&>(*o).normal);
        } else if (o->objectType == ObjectType::Mesh) {
            const Mesh& m = static_cast<const Mesh&>(*o);
            f.type = RTB_OBJ_MESH;
            f.mesh = meshIndex;
            FlatMeshStorage& st = out.meshStorage[meshIndex];
            const size_t n = m.allTris.size();
            st.pos.resize(n * 9); st.nrm.resize(n * 9); st.uv.resize(n * 6); st.tan.resize(n * 6);
            for (size_t i = 0; i < n; ++i) {
                const Triangle& t = m.allTris[i];
                put3(&st.pos[i * 9], t.a); put3(&st.pos[i * 9 + 3], t.b); put3(&st.pos[i * 9 + 6], t.c);
                put3(&st.nrm[i * 9], t.n_a); put3(&st.nrm[i * 9 + 3], t.n_b); put3(&st.nrm[i * 9 + 6], t.n_c);
                st.uv[i * 6 + 0] = t.t_a.x; st.uv[i * 6 + 1] = t.t_a.y;
                st.uv[i * 6 + 2] = t.t_b.x; st.uv[i * 6 + 3] = t.t_b.y;
                st.uv[i * 6 + 4] = t.t_c.x; st.uv[i * 6 + 5] = t.t_c.y;
                put3(&st.tan[i * 6], t.tangent); put3(&st.tan[i * 6 + 3], t.bitangent);
            }
            if (m.ac) {
                st.nodes.resize(m.ac->nodes.size());
                for (size_t k = 0; k < st.nodes.size(); ++k) {
                    const auto& src = m.ac->nodes[k];
                    RtbNode& d = st.nodes[k];
                    put3(d.lo, src.bounds[0]); put3(d.hi, src.bounds[1]);
                    d.right = src.right; d.firstRef = src.firstRef; d.refCount = src.refCount; d.depth = src.depth;
                }
                st.refs.assign(m.ac->refs.begin(), m.ac->refs.end());
            }
            RtbMesh& fm = out.meshes[meshIndex];
            fm.nTris = (int32_t)n;
            fm.nNodes = (int32_t)st.nodes.size();
            fm.nRefs = (int32_t)st.refs.size();
            fm.pos = st.pos.data(); fm.nrm = st.nrm.data(); fm.uv = st.uv.data(); fm.tan = st.tan.data();
            fm.nodes = st.nodes.data(); fm.refs = st.refs.data();
            fm.diffuseMap = imageOf(m.diffuseMap, m.diffuseMapLoaded);
            fm.normalMap = imageOf(m.normalMap, m.normalMapLoaded);
            fm.specularMap = imageOf(m.specularMap, m.specularMapLoaded);
            ++meshIndex;
        }
        out.objects.push_back(f);
    }

    for (const auto& l : scene.lights) {
        RtbLight f{};
        put3(f.color, l->color);
        f.intensity = l->intensity;
        if (l->type == LightType::DistantLight) {
            f.type = RTB_LIGHT_DISTANT;
            put3(f.v, static_cast<const DistantLight&>(*l).dir);
        } else if (l->type == LightType::PointLight) {
            f.type = RTB_LIGHT_POINT;
            put3(f.v, static_cast<const PointLight&>(*l).pos);
        } else {
            const auto& a = static_cast<const AreaLight&>(*l);
            f.type = RTB_LIGHT_AREA;
            put3(f.v, a.pos);
            const auto pts = a.samplePoints();
            f.pointOffset = (int32_t)(out.areaPoints.size() / 3);
            f.pointCount = (int32_t)pts.size();
            for (const Vec3f& p : pts) { out.areaPoints.push_back(p.x); out.areaPoints.push_back(p.y); out.areaPoints.push_back(p.z); }
        }
        out.lights.push_back(f);
    }

    s.nObjects = (int32_t)out.objects.size();
    s.nLights = (int32_t)out.lights.size();
    s.nMeshes = (int32_t)out.meshes.size();
    s.nAreaPoints = (int32_t)(out.areaPoints.size() / 3);
    s.objects = out.objects.data();
    s.lights = out.lights.data();
    s.meshes = out.meshes.data();
    s.areaPoints = out.areaPoints.data();
    if (options::useSkybox)
        for (int k = 0; k < 6; ++k) {
            s.skybox[k] = imageOf(scene.skyboxes[k], true);
            // the reference indexes every face with the last face's dimensions
            s.skybox[k].width = scene.skyboxWidth;
            s.skybox[k].height = scene.skyboxHeight;
        }
}

TreeStats treeStats(const Mesh& mesh)
{
    TreeStats st{};
    if (!mesh.ac) return st;
    st.nodes = (int64_t)mesh.ac->nodes.size();
    for (const auto& n : mesh.ac->nodes) {
        st.maxDepth = std::max<int64_t>(st.maxDepth, n.depth);
        if (n.right < 0) {
            st.leaves++;
            st.refs += n.refCount;
            st.maxLeaf = std::max<int64_t>(st.maxLeaf, n.refCount);
        }
    }
    const Vec3f lo = mesh.ac->rootBounds[0], hi = mesh.ac->rootBounds[1];
    auto outside = [&](const Vec3f& p) { return p.x < lo.x || p.y < lo.y || p.z < lo.z || p.x > hi.x || p.y > hi.y || p.z > hi.z; };
    for (const Triangle& t : mesh.allTris) st.trisOutsideRoot += (outside(t.a) || outside(t.b) || outside(t.c));
    return st;
}

} // namespace rtb
