// lbvh_build.cuh — search-BVH construction ON THE DEVICE (reference: the host-side tree build of
// AccelerationStructure::setup, objects.cpp:470-526,633-763, which this replaces for the SEARCH structure only).
//
// The reference's own split tree still comes from the host (it defines which triangles a ray may hit — eligibility — and
// must be reproduced bit for bit).  The structure the kernels actually search can be any valid BVH over the mesh's unique
// triangles: a different tree changes the nodes visited, never the hit (rt_device.cuh eligibleSlot decides), so frames are
// bit-identical to the host-built binned-SAH tree (tests/test_gpu_parity.py compares them on every golden).
//
// Linear BVH (Lauterbach et al. 2009 / Karras 2012 "Maximizing parallelism in the construction of BVHs"):
//   k_lbvh_keys    63-bit key per triangle: 30-bit Morton code of the centroid (mesh bounds) << 32 | triangle index (keys are unique)
//   cub::DeviceRadixSort::SortKeys                 (the one library call: a 64-bit key sort; everything else is below)
//   k_lbvh_leaves  triangles in sorted order in the kernels' 3 x float4 layout + their padded boxes
//   k_lbvh_tree    Karras' radix tree: every inner node finds its key range and split in parallel (count-leading-zeros of key XORs)
//   k_lbvh_fit     boxes bottom-up: the second child to arrive at a node (atomic counter) merges and climbs on
//   k_lbvh_emit    nodes in the traversal layout (rtbvh::Node: both children's boxes + links); a subtree of <= 4 triangles is a
//                  contiguous run of the sorted order, so it collapses into one leaf code for free; tree depth for the stack size
// Quality is below the host's binned SAH (more nodes visited per ray), construction takes a fraction of a millisecond instead of
// ~0.1 s for 250k triangles: the choice for geometry that changes per frame (RTB_CREATE_DEVICE_BVH).
#pragma once

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <cfloat>
#include <cstdint>

#include "bvh_build.h"

namespace lbvh {

constexpr int kLeafMax = 4;       // triangles per leaf, like the host builder's default
constexpr int kThreads = 256;

__device__ __forceinline__ unsigned expandBits10(unsigned v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ float finiteOr(float v, float d) { return (v >= -FLT_MAX && v <= FLT_MAX) ? v : d; }

// pos: 9 floats per triangle.  lo / inv: mesh bounds and 1 / extent per axis (0 for a flat axis)
__global__ void k_lbvh_keys(const float* __restrict__ pos, int n, float3 lo, float3 inv, unsigned long long* __restrict__ keys)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = pos + (size_t)i * 9;
    float c[3];
    for (int a = 0; a < 3; ++a) {
        const float v0 = finiteOr(p[a], 0.f), v1 = finiteOr(p[3 + a], 0.f), v2 = finiteOr(p[6 + a], 0.f);
        c[a] = 0.5f * fminf(v0, fminf(v1, v2)) + 0.5f * fmaxf(v0, fmaxf(v1, v2));
    }
    const float qx = fminf(fmaxf((c[0] - lo.x) * inv.x * 1024.f, 0.f), 1023.f);
    const float qy = fminf(fmaxf((c[1] - lo.y) * inv.y * 1024.f, 0.f), 1023.f);
    const float qz = fminf(fmaxf((c[2] - lo.z) * inv.z * 1024.f, 0.f), 1023.f);
    const unsigned morton = (expandBits10((unsigned)qx) << 2) | (expandBits10((unsigned)qy) << 1) | expandBits10((unsigned)qz);
    keys[i] = ((unsigned long long)morton << 32) | (unsigned)i;
}

__global__ void k_lbvh_leaves(const unsigned long long* __restrict__ keys, const float* __restrict__ pos, int n, float pad,
    float4* __restrict__ tris, float4* __restrict__ leafLo, float4* __restrict__ leafHi)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int tri = (int)(keys[k] & 0xffffffffu);
    const float* p = pos + (size_t)tri * 9;
    tris[k * 3 + 0] = make_float4(p[0], p[1], p[2], __int_as_float(tri));
    tris[k * 3 + 1] = make_float4(p[3] - p[0], p[4] - p[1], p[5] - p[2], 0.f);   // v1 - v0 (objects.cpp:70)
    tris[k * 3 + 2] = make_float4(p[6] - p[0], p[7] - p[1], p[8] - p[2], 0.f);   // v2 - v0 (objects.cpp:71)
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        lo[a] = fminf(p[a], fminf(p[3 + a], p[6 + a]));
        hi[a] = fmaxf(p[a], fmaxf(p[3 + a], p[6 + a]));
        // NaN / inf vertices (degenerate inputs) would poison the boxes: such a triangle's box covers everything, like on the host
        if (!(lo[a] >= -FLT_MAX && hi[a] <= FLT_MAX) || !(p[a] == p[a]) || !(p[3 + a] == p[3 + a]) || !(p[6 + a] == p[6 + a])) { lo[a] = -FLT_MAX; hi[a] = FLT_MAX; }
        else { lo[a] -= pad; hi[a] += pad; }
    }
    leafLo[k] = make_float4(lo[0], lo[1], lo[2], 0.f);
    leafHi[k] = make_float4(hi[0], hi[1], hi[2], 0.f);
}

__device__ __forceinline__ int lbvhDelta(const unsigned long long* keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    return __clzll((long long)(keys[i] ^ keys[j]));      // keys are unique: never 64
}

// Karras 2012, fig. 4.  Inner node i (0 .. n-2); children >= 0 are inner nodes, < 0 are leaves encoded as ~sortedIndex.
__global__ void k_lbvh_tree(const unsigned long long* __restrict__ keys, int n, int* __restrict__ childL, int* __restrict__ childR,
    int* __restrict__ first, int* __restrict__ last, int* __restrict__ parentInner, int* __restrict__ parentLeaf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (lbvhDelta(keys, n, i, i + 1) - lbvhDelta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dMin = lbvhDelta(keys, n, i, i - d);
    int lMax = 2;
    while (lbvhDelta(keys, n, i, i + lMax * d) > dMin) lMax *= 2;
    int l = 0;
    for (int t = lMax / 2; t >= 1; t /= 2)
        if (lbvhDelta(keys, n, i, i + (l + t) * d) > dMin) l += t;
    const int j = i + l * d;
    const int dNode = lbvhDelta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {      // binary search for the split, step sizes ceil(l / 2^k)
        if (lbvhDelta(keys, n, i, i + (s + t) * d) > dNode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int cl = (lo == gamma) ? ~gamma : gamma;
    const int cr = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    childL[i] = cl; childR[i] = cr;
    first[i] = lo; last[i] = hi;
    if (cl >= 0) parentInner[cl] = i; else parentLeaf[~cl] = i;
    if (cr >= 0) parentInner[cr] = i; else parentLeaf[~cr] = i;
    if (i == 0) parentInner[0] = -1;
}

__global__ void k_lbvh_fit(int n, const int* __restrict__ childL, const int* __restrict__ childR, const int* __restrict__ parentInner,
    const int* __restrict__ parentLeaf, const float4* leafLo, const float4* leafHi, float4* nodeLo, float4* nodeHi, int* arrived)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int node = parentLeaf[k];
    while (node >= 0) {
        __threadfence();                                   // this thread's child box before the arrival count
        if (atomicAdd(&arrived[node], 1) == 0) return;     // first child to arrive: the sibling finishes the node
        __threadfence();
        const int cl = childL[node], cr = childR[node];
        const volatile float4* lLo = cl >= 0 ? nodeLo + cl : leafLo + ~cl;
        const volatile float4* lHi = cl >= 0 ? nodeHi + cl : leafHi + ~cl;
        const volatile float4* rLo = cr >= 0 ? nodeLo + cr : leafLo + ~cr;
        const volatile float4* rHi = cr >= 0 ? nodeHi + cr : leafHi + ~cr;
        nodeLo[node] = make_float4(fminf(lLo->x, rLo->x), fminf(lLo->y, rLo->y), fminf(lLo->z, rLo->z), 0.f);
        nodeHi[node] = make_float4(fmaxf(lHi->x, rHi->x), fmaxf(lHi->y, rHi->y), fmaxf(lHi->z, rHi->z), 0.f);
        node = parentInner[node];
    }
}

// Traversal layout.  A child whose key range holds <= kLeafMax triangles becomes a leaf code ~((first << 3) | (count - 1)): the
// range is a contiguous run of the sorted triangle array.  Inner nodes inside such runs are simply never referenced.
__global__ void k_lbvh_emit(int n, const int* __restrict__ childL, const int* __restrict__ childR, const int* __restrict__ first,
    const int* __restrict__ last, const int* __restrict__ parentInner, const float4* __restrict__ leafLo, const float4* __restrict__ leafHi,
    const float4* __restrict__ nodeLo, const float4* __restrict__ nodeHi, rtbvh::Node* __restrict__ out, int* __restrict__ maxDepth)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    if (i != 0 && last[i] - first[i] + 1 <= kLeafMax) return;      // collapsed into its parent's leaf code
    rtbvh::Node nd;
    const int ch[2] = { childL[i], childR[i] };
    int link[2];
    for (int c = 0; c < 2; ++c) {
        const int k = ch[c];
        const float4 lo = k >= 0 ? nodeLo[k] : leafLo[~k], hi = k >= 0 ? nodeHi[k] : leafHi[~k];
        float* dlo = c == 0 ? nd.c0lo : nd.c1lo;
        float* dhi = c == 0 ? nd.c0hi : nd.c1hi;
        dlo[0] = lo.x; dlo[1] = lo.y; dlo[2] = lo.z;
        dhi[0] = hi.x; dhi[1] = hi.y; dhi[2] = hi.z;
        if (k < 0) link[c] = ~((~k << 3) | 0);
        else if (last[k] - first[k] + 1 <= kLeafMax) link[c] = ~((first[k] << 3) | (last[k] - first[k]));
        else link[c] = k;
    }
    nd.child0 = link[0]; nd.child1 = link[1]; nd.pad0 = nd.pad1 = 0;
    out[i] = nd;
    int depth = 2;                                         // this node and the level of its children
    for (int p = parentInner[i]; p >= 0; p = parentInner[p]) ++depth;
    atomicMax(maxDepth, depth);
}

// single triangle: the root is an inner node whose second child is empty (like the host builder)
__global__ void k_lbvh_single(const float4* __restrict__ leafLo, const float4* __restrict__ leafHi, rtbvh::Node* __restrict__ out, int* __restrict__ maxDepth)
{
    rtbvh::Node nd;
    const float4 lo = leafLo[0], hi = leafHi[0];
    nd.c0lo[0] = lo.x; nd.c0lo[1] = lo.y; nd.c0lo[2] = lo.z;
    nd.c0hi[0] = hi.x; nd.c0hi[1] = hi.y; nd.c0hi[2] = hi.z;
    for (int a = 0; a < 3; ++a) { nd.c1lo[a] = FLT_MAX; nd.c1hi[a] = -FLT_MAX; }
    nd.child0 = ~0; nd.child1 = ~0; nd.pad0 = nd.pad1 = 0;
    out[0] = nd;
    *maxDepth = 2;
}

struct DeviceBvh {
    rtbvh::Node* nodes = nullptr;   // max(1, n - 1) entries, root at 0 (entries inside collapsed subtrees are unused)
    float4* tris = nullptr;         // 3 per triangle, sorted order
    int nNodes = 0;
    int maxDepth = 0;
    float buildMs = 0.f;            // CUDA-event time of the device work (keys .. emit)
};

#define LBVH_CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return e__; } while (0)

// dPos: device copy of the mesh's 9 floats per triangle.  lo / hi: finite mesh bounds (host).  The caller owns out.nodes / out.tris.
inline cudaError_t buildOnDevice(const float* dPos, int n, const float lo[3], const float hi[3], float pad, cudaStream_t st, DeviceBvh& out)
{
    out = DeviceBvh{};
    if (n <= 0) return cudaSuccess;
    const int grid = (n + kThreads - 1) / kThreads;
    unsigned long long *keysA = nullptr, *keysB = nullptr;
    float4 *leafLo = nullptr, *leafHi = nullptr, *nodeLo = nullptr, *nodeHi = nullptr;
    int *childL = nullptr, *childR = nullptr, *first = nullptr, *last = nullptr, *parentInner = nullptr, *parentLeaf = nullptr, *arrived = nullptr, *dDepth = nullptr;
    void* temp = nullptr;
    size_t tempBytes = 0;
    const int inner = n > 1 ? n - 1 : 1;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t rc = cudaSuccess;
    auto body = [&]() -> cudaError_t {
        LBVH_CK(cudaMalloc(&keysA, (size_t)n * 8)); LBVH_CK(cudaMalloc(&keysB, (size_t)n * 8));
        LBVH_CK(cudaMalloc(&leafLo, (size_t)n * 16)); LBVH_CK(cudaMalloc(&leafHi, (size_t)n * 16));
        LBVH_CK(cudaMalloc(&nodeLo, (size_t)inner * 16)); LBVH_CK(cudaMalloc(&nodeHi, (size_t)inner * 16));
        LBVH_CK(cudaMalloc(&childL, (size_t)inner * 4)); LBVH_CK(cudaMalloc(&childR, (size_t)inner * 4));
        LBVH_CK(cudaMalloc(&first, (size_t)inner * 4)); LBVH_CK(cudaMalloc(&last, (size_t)inner * 4));
        LBVH_CK(cudaMalloc(&parentInner, (size_t)inner * 4)); LBVH_CK(cudaMalloc(&parentLeaf, (size_t)n * 4));
        LBVH_CK(cudaMalloc(&arrived, (size_t)inner * 4)); LBVH_CK(cudaMalloc(&dDepth, 4));
        LBVH_CK(cudaMalloc(&out.nodes, (size_t)inner * sizeof(rtbvh::Node)));
        LBVH_CK(cudaMalloc(&out.tris, (size_t)n * 3 * sizeof(float4)));
        LBVH_CK(cub::DeviceRadixSort::SortKeys(nullptr, tempBytes, keysA, keysB, n, 0, 62, st));
        LBVH_CK(cudaMalloc(&temp, tempBytes));
        LBVH_CK(cudaEventCreate(&e0)); LBVH_CK(cudaEventCreate(&e1));
        LBVH_CK(cudaEventRecord(e0, st));
        float3 l3 = make_float3(lo[0], lo[1], lo[2]), inv;
        inv.x = hi[0] > lo[0] ? 1.0f / (hi[0] - lo[0]) : 0.f;
        inv.y = hi[1] > lo[1] ? 1.0f / (hi[1] - lo[1]) : 0.f;
        inv.z = hi[2] > lo[2] ? 1.0f / (hi[2] - lo[2]) : 0.f;
        k_lbvh_keys<<<grid, kThreads, 0, st>>>(dPos, n, l3, inv, keysA);
        LBVH_CK(cub::DeviceRadixSort::SortKeys(temp, tempBytes, keysA, keysB, n, 0, 62, st));
        k_lbvh_leaves<<<grid, kThreads, 0, st>>>(keysB, dPos, n, pad, out.tris, leafLo, leafHi);
        LBVH_CK(cudaMemsetAsync(arrived, 0, (size_t)inner * 4, st));
        LBVH_CK(cudaMemsetAsync(dDepth, 0, 4, st));
        if (n == 1) {
            k_lbvh_single<<<1, 1, 0, st>>>(leafLo, leafHi, out.nodes, dDepth);
        } else {
            k_lbvh_tree<<<grid, kThreads, 0, st>>>(keysB, n, childL, childR, first, last, parentInner, parentLeaf);
            k_lbvh_fit<<<grid, kThreads, 0, st>>>(n, childL, childR, parentInner, parentLeaf, leafLo, leafHi, nodeLo, nodeHi, arrived);
            k_lbvh_emit<<<grid, kThreads, 0, st>>>(n, childL, childR, first, last, parentInner, leafLo, leafHi, nodeLo, nodeHi, out.nodes, dDepth);
        }
        LBVH_CK(cudaGetLastError());
        LBVH_CK(cudaEventRecord(e1, st));
        LBVH_CK(cudaMemcpyAsync(&out.maxDepth, dDepth, 4, cudaMemcpyDeviceToHost, st));
        LBVH_CK(cudaStreamSynchronize(st));
        LBVH_CK(cudaEventElapsedTime(&out.buildMs, e0, e1));
        out.nNodes = inner;
        return cudaSuccess;
    };
    rc = body();
    for (void* p : { (void*)keysA, (void*)keysB, (void*)leafLo, (void*)leafHi, (void*)nodeLo, (void*)nodeHi, (void*)childL, (void*)childR, (void*)first,
             (void*)last, (void*)parentInner, (void*)parentLeaf, (void*)arrived, (void*)dDepth, temp })
        if (p) cudaFree(p);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (rc != cudaSuccess) {
        if (out.nodes) cudaFree(out.nodes);
        if (out.tris) cudaFree(out.tris);
        out = DeviceBvh{};
    }
    return rc;
}

} // namespace lbvh
