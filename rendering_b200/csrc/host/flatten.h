// flatten.h — Scene graph -> POD RtbScene (include/rtb.h) with owned storage.
#pragma once

#include <string>
#include <vector>

#include "../../../include/rtb.h"
#include "scene.h"

namespace rtb {

struct FlatMeshStorage {
    std::vector<float> pos, nrm, uv, tan;
    std::vector<RtbNode> nodes;
    std::vector<int32_t> refs;
};

// Owns every array an RtbScene points into.  Texture / skybox bytes are borrowed from the Scene,
// which must outlive the FlatScene.
struct FlatScene {
    RtbScene view{};
    std::vector<RtbObject> objects;
    std::vector<RtbLight> lights;
    std::vector<RtbMesh> meshes;
    std::vector<FlatMeshStorage> meshStorage;
    std::vector<float> areaPoints;
    std::string imageName;
};

void flatten(const Scene& scene, FlatScene& out);
// Camera (scene.h:51-65) -> the host-computed constants of Camera::getRay / renderWorker (scene.cpp:24-48, 447-448)
RtbCamera flattenCamera(const Camera& camera, size_t width, size_t height);

struct TreeStats { int64_t nodes, leaves, refs, maxLeaf, maxDepth, trisOutsideRoot; };
TreeStats treeStats(const Mesh& mesh);

} // namespace rtb
