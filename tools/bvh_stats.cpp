// tools/bvh_stats.cpp — development aid: traversal-step statistics of the search BVH on the CPU.
//   g++ -O2 -std=c++17 -ffp-contract=off -Iinclude tools/bvh_stats.cpp -Lrendering_b200 -lrtb_host -Wl,-rpath,$PWD/rendering_b200 -o /tmp/bvh_stats
//   cd scenes && /tmp/bvh_stats cfg4_shotgun_1080.scene
// Traces every primary ray of the scene through the same structure the kernels use and prints how many inner nodes,
// leaves and triangles a ray touches (mean / percentiles / max) — the quantities that bound k_walk.
#include <algorithm>
#include <cstdio>
#include <vector>

#include "../rendering_b200/csrc/cuda/scene_pack.h"

using namespace rt;

int main(int argc, char** argv)
{
    RtbHostScene* hs = nullptr;
    if (argc < 2 || rtb_scene_load(argv[1], &hs) != RTB_OK) { fprintf(stderr, "cannot load scene: %s\n", rtb_host_last_error()); return 1; }
    const RtbScene* s = rtb_scene_view(hs);
    Scene sc; rtpack::packHeader(*s, sc);
    for (int mi = 0; mi < s->nMeshes; ++mi) {
        const RtbMesh& m = s->meshes[mi];
        rtpack::FastPath fp; rtpack::packFastPath(m, fp);
        printf("mesh %d: %d tris, %zu bvh nodes, %zu leaf triangle slots, depth %d\n", mi, m.nTris, fp.nodes.size(), fp.tris.size() / 3, fp.maxDepth);
        std::vector<int> nodesV, leavesV, trisV;
        const bool cull = sc.flags & FLAG_CULL;
        for (int y = 0; y + 1 < sc.height; y += 2)
            for (int x = 0; x + 1 < sc.width; x += 2) {
                const V3 d = cameraDir(sc, (float)x + 0.5f, (float)y + 0.5f);
                const RayCtx r = makeRay(sc.camPos, d);
                int nNodes = 0, nLeaves = 0, nTris = 0;
                float tBest = FLT_MAX;
                int stack[128], sp = 0, cur = 0;
                for (;;) {
                    if (cur >= 0) {
                        const rtbvh::Node& nd = fp.nodes[cur];
                        nNodes++;
                        bool h0, h1;
                        const float e0 = slabEntry(r, nd.c0lo[0], nd.c0lo[1], nd.c0lo[2], nd.c0hi[0], nd.c0hi[1], nd.c0hi[2], tBest, h0);
                        const float e1 = slabEntry(r, nd.c1lo[0], nd.c1lo[1], nd.c1lo[2], nd.c1hi[0], nd.c1hi[1], nd.c1hi[2], tBest, h1);
                        if (h0 && h1) { const bool sw = e1 < e0; stack[sp++] = sw ? nd.child0 : nd.child1; cur = sw ? nd.child1 : nd.child0; continue; }
                        if (h0) { cur = nd.child0; continue; }
                        if (h1) { cur = nd.child1; continue; }
                    } else {
                        const int code = ~cur, first = code >> 3, count = (code & 7) + 1;
                        nLeaves++;
                        for (int k = 0; k < count; ++k) {
                            const float4* tp = &fp.tris[(size_t)(first + k) * 3];
                            float t, u, v;
                            nTris++;
                            if (hitTriangle(r, mk(tp[0].x, tp[0].y, tp[0].z), mk(tp[1].x, tp[1].y, tp[1].z), mk(tp[2].x, tp[2].y, tp[2].z), cull, t, u, v) && t < tBest) tBest = t;
                        }
                    }
                    if (sp == 0) break;
                    cur = stack[--sp];
                }
                if (nNodes > 1) { nodesV.push_back(nNodes); leavesV.push_back(nLeaves); trisV.push_back(nTris); }
            }
        auto report = [](const char* name, std::vector<int>& v) {
            if (v.empty()) return;
            std::sort(v.begin(), v.end());
            double sum = 0; for (int a : v) sum += a;
            printf("  %-14s rays %zu  mean %.1f  p50 %d  p90 %d  p99 %d  max %d\n", name, v.size(), sum / v.size(), v[v.size() / 2], v[v.size() * 9 / 10], v[v.size() * 99 / 100], v.back());
        };
        report("inner nodes", nodesV); report("leaves", leavesV); report("triangle tests", trisV);
    }
    rtb_scene_free(hs);
    return 0;
}
