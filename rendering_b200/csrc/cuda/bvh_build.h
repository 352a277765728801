// bvh_build.h — host-side builder of the traversal BVH used by the fast path.
//
// The reference's per-mesh tree (AccelerationStructure, objects.cpp:470-526) decides WHICH triangles
// a ray may hit (its "eligibility" semantics are reproduced exactly, see rtb_kernels.cuh) but is a
// poor search structure: loose boxes, no ordering, leaves of up to 15 756 triangles on the dragon.
// For speed the kernels search a second structure built here: a binned-SAH BVH2 over the mesh's
// UNIQUE triangles with tight (slightly padded) boxes and at most `maxLeaf` (default 4) triangles per leaf.
//
// Layout (what the GPU reads):
//   node k = 4 x float4: {c0.lo.xyz, c0.hi.x} {c0.hi.yz, c1.lo.xy} {c1.lo.z, c1.hi.xyz} {child0, child1, -, -}
//   child >= 0: inner node index; child < 0: leaf, ~child = (firstTriangle << 3) | (count - 1)
//   triangles in leaf order: 3 x float4 {v0.xyz, triangle id} {e1.xyz, -} {e2.xyz, -}
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <future>
#include <vector>

namespace rtbvh {

struct Box {
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX };
    float hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    void grow(const float* p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void grow(const Box& b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float area() const
    {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return (dx < 0 || dy < 0 || dz < 0) ? 0.0f : 2.0f * (dx * dy + dy * dz + dz * dx);
    }
};

struct Node {   // 64 bytes, read as 4 x float4
    float c0lo[3], c0hi[3], c1lo[3], c1hi[3];
    int32_t child0, child1, pad0, pad1;
};
static_assert(sizeof(Node) == 64, "BVH node must be 64 bytes");

struct Result {
    std::vector<Node> nodes;
    std::vector<int> triOrder;   // triangle ids in leaf order
    int maxDepth = 0;
};

constexpr int kMaxLeafLimit = 8;   // leaf codes keep (count - 1) in 3 bits
constexpr int kBins = 16;

class Builder {
public:
    // pos: 9 floats per triangle.  pad: absolute padding added to every leaf-level triangle box so a
    // hit reported by the float Moller-Trumbore test is never culled by the box test.
    Result build(const float* pos, int nTris, float pad, int maxLeaf = 4)
    {
        pos_ = pos;
        pad_ = pad;
        maxLeaf_ = std::min(kMaxLeafLimit, std::max(1, maxLeaf));
        res_ = Result{};
        boxes_.resize(nTris);
        cent_.resize((size_t)nTris * 3);
        ids_.resize(nTris);
        for (int i = 0; i < nTris; ++i) {
            Box b;
            for (int v = 0; v < 3; ++v) b.grow(pos + (size_t)i * 9 + v * 3);
            for (int a = 0; a < 3; ++a) {
                // NaN / inf vertices (degenerate inputs) would poison the boxes: clamp them out
                if (!(b.lo[a] >= -FLT_MAX && b.hi[a] <= FLT_MAX)) { b.lo[a] = -FLT_MAX; b.hi[a] = FLT_MAX; }
                cent_[(size_t)i * 3 + a] = 0.5f * b.lo[a] + 0.5f * b.hi[a];
                b.lo[a] -= pad;
                b.hi[a] += pad;
            }
            boxes_[i] = b;
            ids_[i] = i;
        }
        // the root is always an inner node so the kernel never has to special-case a leaf root
        res_.nodes.emplace_back();
        if (nTris == 0) {
            setEmpty(res_.nodes[0].c0lo, res_.nodes[0].c0hi);
            setEmpty(res_.nodes[0].c1lo, res_.nodes[0].c1hi);
            res_.nodes[0].child0 = res_.nodes[0].child1 = ~0;
            return res_;
        }
        parallel_ = std::getenv("RTB_BVH_SERIAL") == nullptr;
        splitInto(res_, 0, 0, nTris, 1);
        return res_;
    }

private:
    const float* pos_ = nullptr;
    float pad_ = 0;
    int maxLeaf_ = 4;
    bool parallel_ = true;
    Result res_;
    std::vector<Box> boxes_;
    std::vector<float> cent_;
    std::vector<int> ids_;

    static void setEmpty(float* lo, float* hi) { for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; } }

    Box rangeBox(int first, int last) const
    {
        Box b;
        for (int i = first; i < last; ++i) b.grow(boxes_[ids_[i]]);
        return b;
    }

    int makeLeaf(Result& res, int first, int last) const
    {
        const int start = (int)res.triOrder.size();
        for (int i = first; i < last; ++i) res.triOrder.push_back(ids_[i]);
        return ~((start << 3) | (last - first - 1));
    }

    // appends a subtree built on its own (another thread) and returns the index its root got; inner links and
    // leaf codes are shifted into `res`'s numbering, which reproduces the sequential (DFS) layout exactly
    static int append(Result& res, const Result& sub)
    {
        const int nodeBase = (int)res.nodes.size(), triBase = (int)res.triOrder.size();
        auto shift = [&](int child) {
            if (child >= 0) return child + nodeBase;
            const int code = ~child;
            return ~((((code >> 3) + triBase) << 3) | (code & 7));
        };
        for (Node n : sub.nodes) {
            n.child0 = shift(n.child0);
            n.child1 = shift(n.child1);
            res.nodes.push_back(n);
        }
        res.triOrder.insert(res.triOrder.end(), sub.triOrder.begin(), sub.triOrder.end());
        res.maxDepth = std::max(res.maxDepth, sub.maxDepth);
        return nodeBase;
    }

    // chooses a partition of ids_[first,last) and returns the split position
    int partition(int first, int last)
    {
        const int n = last - first;
        Box cb;
        for (int i = first; i < last; ++i) cb.grow(&cent_[(size_t)ids_[i] * 3]);
        int bestAxis = -1, bestBin = -1;
        float bestCost = FLT_MAX;
        for (int axis = 0; axis < 3; ++axis) {
            const float ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 0)) continue;
            Box binBox[kBins];
            int binCount[kBins] = { 0 };
            const float scale = kBins / ext;
            for (int i = first; i < last; ++i) {
                const int id = ids_[i];
                int b = (int)((cent_[(size_t)id * 3 + axis] - cb.lo[axis]) * scale);
                b = std::min(kBins - 1, std::max(0, b));
                binCount[b]++;
                binBox[b].grow(boxes_[id]);
            }
            float rightArea[kBins];
            int rightCount[kBins];
            Box acc;
            int cnt = 0;
            for (int b = kBins - 1; b > 0; --b) {
                acc.grow(binBox[b]);
                cnt += binCount[b];
                rightArea[b] = acc.area();
                rightCount[b] = cnt;
            }
            acc = Box();
            cnt = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                acc.grow(binBox[b]);
                cnt += binCount[b];
                if (cnt == 0 || rightCount[b + 1] == 0) continue;
                const float cost = acc.area() * cnt + rightArea[b + 1] * rightCount[b + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestBin = b; }
            }
        }
        if (bestAxis < 0) return first + n / 2;   // all centroids coincide: split by count
        const float ext = cb.hi[bestAxis] - cb.lo[bestAxis];
        const float scale = kBins / ext;
        const float lo = cb.lo[bestAxis];
        auto mid = std::partition(ids_.begin() + first, ids_.begin() + last, [&](int id) {
            int b = (int)((cent_[(size_t)id * 3 + bestAxis] - lo) * scale);
            b = std::min(kBins - 1, std::max(0, b));
            return b <= bestBin;
        });
        int m = (int)(mid - ids_.begin());
        if (m == first || m == last) m = first + n / 2;
        return m;
    }

    // ids_[first,last) are disjoint per subtree and boxes_ / cent_ are read-only, so the two halves of a large node are
    // built concurrently near the root, each into its own Result
    void splitInto(Result& res, int nodeIndex, int first, int last, int depth)
    {
        res.maxDepth = std::max(res.maxDepth, depth);
        const int n = last - first;
        int m;
        if (n <= 1) m = last;            // single triangle: second child stays empty
        else m = partition(first, last);
        const int ranges[2][2] = { { first, m }, { m, last } };
        int children[2];
        Box cbox[2];
        const bool fork = parallel_ && depth <= 3 && n > 30000 && m - first > maxLeaf_ && last - m > maxLeaf_;
        if (fork) {
            Result sub[2];
            for (int c = 0; c < 2; ++c) { cbox[c] = rangeBox(ranges[c][0], ranges[c][1]); sub[c].nodes.emplace_back(); }
            auto job = std::async(std::launch::async, [&]() { splitInto(sub[0], 0, ranges[0][0], ranges[0][1], depth + 1); });
            splitInto(sub[1], 0, ranges[1][0], ranges[1][1], depth + 1);
            job.get();
            children[0] = append(res, sub[0]);
            children[1] = append(res, sub[1]);
        } else {
            for (int c = 0; c < 2; ++c) {
                const int f = ranges[c][0], l = ranges[c][1];
                if (l - f <= 0) { children[c] = ~0; cbox[c] = Box(); continue; }
                cbox[c] = rangeBox(f, l);
                if (l - f <= maxLeaf_) {
                    children[c] = makeLeaf(res, f, l);
                } else {
                    children[c] = (int)res.nodes.size();
                    res.nodes.emplace_back();
                    splitInto(res, children[c], f, l, depth + 1);
                }
            }
        }
        Node& nd = res.nodes[nodeIndex];
        std::memcpy(nd.c0lo, cbox[0].lo, 12); std::memcpy(nd.c0hi, cbox[0].hi, 12);
        std::memcpy(nd.c1lo, cbox[1].lo, 12); std::memcpy(nd.c1hi, cbox[1].hi, 12);
        nd.child0 = children[0];
        nd.child1 = children[1];
        nd.pad0 = nd.pad1 = 0;
    }
};

} // namespace rtbvh
