// objects.h — scene objects (reference include/objects.h).  Same class and field names for the
// scene graph; meshes are stored structure-of-arrays and the per-mesh split tree is a flat
// pre-order node pool instead of a pointer tree, because both are uploaded to the GPU as they are.
// Intersection / shading methods of the reference (intersectObject, getSurfaceData, ...) have no
// host counterpart here: that work exists only as CUDA kernels.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "geometry.h"
#include "options.h"

enum class ObjectType { Object, Sphere, Plane, Mesh };
enum class MaterialType { Diffuse, Reflective, Transparent, Phong };

class Object {
public:
    virtual ~Object() = default;
    ObjectType objectType = ObjectType::Object;
    Vec3f pos{ 1.0f };
    Vec3f color{ 1.0f };
    MaterialType materialType = MaterialType::Diffuse;
    float indexOfRefraction = 1.4f;
    float ambient = 0.1f;
    float diffuse = 0.1f;
    float specular = 1.0f;
    float nSpecular = 5.0f;
};

class Sphere : public Object {
public:
    Sphere() { objectType = ObjectType::Sphere; pos = Vec3f(0.0f); }
    float r = 1.0f;
    float r2 = 1.0f;
};

class Plane : public Object {
public:
    Plane() { objectType = ObjectType::Plane; normal.normalize(); }
    Vec3f normal{ 0, 1, 0 };
};

// One triangle as the reference's Triangle class holds it (include/objects.h:49-67).
struct Triangle {
    Vec3f a, b, c;
    Vec3f n_a, n_b, n_c;
    Vec2f t_a, t_b, t_c;
    Vec3f tangent, bitangent;
};

// The per-mesh spatial split tree of the reference (AccelerationStructure, objects.cpp:470-526,
// 633-763), built bit-for-bit but stored flat: nodes in DFS pre-order (left child = index+1).
class AccelerationStructure {
public:
    struct Node {
        Vec3f bounds[2];
        int right = -1;       // index of right child; -1 => leaf
        int firstRef = 0;     // leaves: range in `refs`
        int refCount = 0;
        int depth = 1;
    };
    void setBounds(const Vec3f& lo, const Vec3f& hi) { rootBounds[0] = lo; rootBounds[1] = hi; }
    // Build over triangles [0, a.size()) given as position arrays; a_depth starts at 1.
    void setup(const std::vector<Triangle>& tris, const Options& options);

    static float calculateSAH(int orientation, const std::vector<Triangle>& tris, const std::vector<int>& ids,
        const Vec3f bounds[2], float boundary);
    static float binarySearchSAH(int orientation, const std::vector<Triangle>& tris, const std::vector<int>& ids,
        const Vec3f bounds[2], float left, float right);

    Vec3f rootBounds[2];
    std::vector<Node> nodes;
    std::vector<int> refs;

private:
    // a subtree in its own coordinates: node links and reference ranges are relative to its first node / first ref
    struct Subtree { std::vector<Node> nodes; std::vector<int> refs; };
    static void build(Subtree& out, const std::vector<Triangle>& tris, std::vector<int>& ids, const Vec3f bounds[2], int depth,
        const Options& options);
    static void append(Subtree& out, const Subtree& child);
};

struct TextureRGB8 {
    std::vector<uint8_t> rgb;   // R,G,B per texel, file row order (loadBMP, util.cpp:78-113)
    int width = 0, height = 0;
    bool loaded() const { return !rgb.empty(); }
};

class Mesh : public Object {
public:
    Mesh() { objectType = ObjectType::Mesh; }

    bool loadOBJ(const std::string& filename, const Options& options);
    bool loadDiffuseMap(const std::string& filename);
    bool loadNormalMap(const std::string& filename);
    bool loadSpecularMap(const std::string& filename);

    Vec3f size{ 0.0f };
    Vec3f rot{ 0.0f };
    std::vector<Triangle> allTris;
    std::unique_ptr<AccelerationStructure> ac;

    bool diffuseMapLoaded = false, normalMapLoaded = false, specularMapLoaded = false;
    TextureRGB8 diffuseMap, normalMap, specularMap;
};

using ObjectVector = std::vector<std::unique_ptr<Object>>;
