// rtb_tile.cuh — the tile pipeline: castRay's whole recursion for a tile of rays inside ONE persistent kernel.
//
// The frame-wide pipeline (rtb_kernels.cuh: k_walk -> k_surface -> k_walk<ANY> -> k_shade per recursion level, then
// k_combine per level) synchronises the whole GPU between stages: every stage waits for the slowest ray of the
// frame, pays a launch and — on small or deep ray trees — runs at its latency floor (cfg1: 110 launches per frame).
// Here a GROUP of kTileThreads threads takes a tile of up to `tileRays` rays from a global tile cursor and runs the
// same stages on it back to back with group-wide barriers only:
//
//   walk (closest hit; level 0 generates its rays in place) -> surface -> walk (shadow) -> shade -> next level ...
//   -> fold interior records deepest level first -> (SSAA) mean of the 4 samples into the pixel
//
// over compacted queues that live in the group's own scratch slab (L1 / L2 resident, never frame-sized).  Groups are
// independent: while one waits for its longest ray, the other groups on the SM traverse, shade or fetch texels, so
// the stages of different tiles overlap on every SM and no stage ever waits for the whole frame.  A frame is
//   [background fill] -> k_tile<GEN_PRIMARY> -> k_sobel -> k_tile<GEN_SSAA> -> output
// whatever the recursion depth.  The stage bodies are the ones the frame-wide kernels use (walkRays, surfaceStage,
// shadeStage, combineStage), so both pipelines produce bit-identical frames (tests compare them).
//
// Colours of SSAA samples and of Reflective / Transparent children never leave the tile: they live in the group's
// tile-local slot array (Slots::l); only finished pixels are written to the framebuffer.
#pragma once

#include "rtb_kernels.cuh"

namespace rtk {

#ifndef RTB_TILE_THREADS
#define RTB_TILE_THREADS 256
#endif
constexpr int kTileThreads = RTB_TILE_THREADS;     // threads per group
constexpr int kMaxTileLevels = 64;    // recursion levels the per-group level table holds (deeper scenes use the frame-wide pipeline)
#ifndef RTB_TILE_MIN_BLOCKS
#define RTB_TILE_MIN_BLOCKS 4
#endif

// counters of the level with parity p; the other parity's set is cleared while this one is in use
struct TileLevelCtr {
    unsigned cursorClosest, cursorShadow;
    int nSurf, nNext;
};
struct TileShared {
    unsigned long long tile;
    int interiors;
    int pad;
    TileLevelCtr lv[2];
    int levelEnd[kMaxTileLevels + 1];   // interior records handed out up to and including level l
    int bins[32];                       // surface sort (optional): per-key counts, then running slot offsets
};

struct TileArgs {
    char* slab;                    // scratch: group g owns [g * slabBytes, (g + 1) * slabBytes)
    unsigned long long slabBytes;
    int capRays;                   // rays a level's queue of one tile can hold
    int capInterior;               // interior records of one tile (all levels)
    int levels;                    // recursion levels (1 when nothing spawns children)
    int tileRays;                  // rays of a level-0 tile, a multiple of 32
    int stackEntries;              // traversal stack entries per thread
    unsigned long long* tileCursor;
    FrameCtr* ctr;
    LevelCtr* lv;                  // this pass's level array: statistics only (rays / surfaces per level)
    GenArgs gen;
    RayQueue userQ;                // GEN_QUEUE: level-0 rays (caller-supplied), `nUser` of them
    int nUser;
    // STAGED kernels: the part of one mesh's search BVH every CTA copies into its shared memory at start
    const Mesh* stagedMesh;        // device address of that mesh's header (compared with the mesh being walked)
    const float4* stagedNodesSrc;  // nodes [0, stagedNodes): 4 x float4 each
    const float4* stagedTrisSrc;   // all leaf triangles (3 x float4 each) when stagedTris > 0
    int stagedNodes, stagedTris;
    int sortSurfaces;              // deep scenes: counting-sort each level's surfaces by (material, direction octant)
};

// ---- TMA bulk copy global -> shared, completion on an mbarrier (cp.async.bulk, SASS UBLKCP) ----
__device__ __forceinline__ unsigned smemAddrOf(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned bar, int arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkCopyToShared(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned bar, unsigned phase)
{
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    } while (!done);
}

// bytes of one group's slab for the given capacities (host and device agree through this one function)
__host__ __device__ inline unsigned long long tileSlabBytes(int capRays, int capInterior, int tileRays, int S, int levels)
{
    auto al = [](unsigned long long v) { return (v + 15ull) & ~15ull; };
    const unsigned long long n = (unsigned long long)capRays;
    unsigned long long b = 0;
    b += 2 * (al(n * 16) + al(n * 16) + al(n * 4));            // two ray queues: o, d, dest
    b += al(n * 16) + al(n * 4);                               // hits: tuv, obj
    b += 3 * al(n * 16);                                       // surface records
    b += al(n * (unsigned long long)(S > 0 ? S : 1));          // visibility bytes
    b += al((unsigned long long)capInterior * sizeof(Interior));
    b += al(((unsigned long long)tileRays + 2ull * capInterior) * 12);   // local slots: samples, then 2 per interior record
    (void)levels;
    return al(b);
}

struct TileScratch {
    RayQueue q[2];
    HitQueue hits;
    SurfQueue surf;
    unsigned char* vis;
    Interior* interiors;
    float* localSlots;
};

__device__ __forceinline__ TileScratch carveTileScratch(char* base, int capRays, int capInterior, int tileRays, int S)
{
    auto al = [](unsigned long long v) { return (v + 15ull) & ~15ull; };
    const unsigned long long n = (unsigned long long)capRays;
    TileScratch t;
    char* p = base;
    for (int k = 0; k < 2; ++k) {
        t.q[k].o = reinterpret_cast<float4*>(p); p += al(n * 16);
        t.q[k].d = reinterpret_cast<float4*>(p); p += al(n * 16);
        t.q[k].dest = reinterpret_cast<int*>(p); p += al(n * 4);
    }
    t.hits.tuv = reinterpret_cast<float4*>(p); p += al(n * 16);
    t.hits.obj = reinterpret_cast<int*>(p); p += al(n * 4);
    t.surf.pS = reinterpret_cast<float4*>(p); p += al(n * 16);
    t.surf.nO = reinterpret_cast<float4*>(p); p += al(n * 16);
    t.surf.cR = reinterpret_cast<float4*>(p); p += al(n * 16);
    t.vis = reinterpret_cast<unsigned char*>(p); p += al(n * (unsigned long long)(S > 0 ? S : 1));
    t.interiors = reinterpret_cast<Interior*>(p); p += al((unsigned long long)capInterior * sizeof(Interior));
    t.localSlots = reinterpret_cast<float*>(p);
    (void)tileRays;
    return t;
}

// barrier over one group: the whole CTA when it holds one group, else a named barrier (ids 1..NG) over kTileThreads threads
template <int NG>
__device__ __forceinline__ void groupSync(int group)
{
    if (NG == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kTileThreads) : "memory");
}

// shared-memory layout of a k_tile CTA: [TileShared x NG] [stacks: NG x stackEntries x kTileThreads ints] [mbarrier] [staged nodes] [staged triangles]
__host__ __device__ inline size_t tileSmemStateBytes(int ng) { return (sizeof(TileShared) * ng + 15) & ~(size_t)15; }
__host__ __device__ inline size_t tileSmemStagedOffset(int ng, int stackEntries)
{
    return (tileSmemStateBytes(ng) + (size_t)ng * stackEntries * kTileThreads * sizeof(int) + 16 + 127) & ~(size_t)127;
}

// GEN:    where level-0 rays come from (GEN_PRIMARY / GEN_SSAA generated in place, GEN_QUEUE = args.userQ).
// DEEP:   the scene can spawn secondary rays (levels > 1); false compiles the recursion out.
// NG:     groups per CTA.  NG = 1: several small CTAs per SM.  NG = 4: ONE 1024-thread CTA per SM whose four groups run
//         their tiles independently (named barriers) but share one copy of the scene in shared memory:
// STAGED: the CTA first copies the top of the search BVH (all of it, and the leaf triangles, when they fit) into shared
//         memory with TMA bulk copies and the walks read those nodes with ld.shared instead of through L1 / L2.
template <int GEN, bool DEEP, bool STATS, int NG, bool STAGED = false>
__global__ void __launch_bounds__(kTileThreads * NG, NG == 1 ? RTB_TILE_MIN_BLOCKS : 1) k_tile(Scene sc, TileArgs a)
{
    extern __shared__ __align__(128) unsigned char tileSmem[];
    const int group = threadIdx.x / kTileThreads, gtid = threadIdx.x % kTileThreads;
    TileShared* gsAll = reinterpret_cast<TileShared*>(tileSmem);
    TileShared& gs = gsAll[group];
    int* stackBase = reinterpret_cast<int*>(tileSmem + tileSmemStateBytes(NG));
    int* stack = stackBase + (size_t)group * a.stackEntries * kTileThreads + gtid;

    StagedBvh sb{};
    if (STAGED) {
        const size_t off = tileSmemStagedOffset(NG, a.stackEntries);
        const unsigned bar = smemAddrOf(tileSmem + off - 16);
        const unsigned nodeBytes = (unsigned)a.stagedNodes * 64u, triBytes = (unsigned)a.stagedTris * 48u;
        sb.mesh = a.stagedMesh;
        sb.nodesAddr = smemAddrOf(tileSmem + off);
        sb.trisAddr = sb.nodesAddr + nodeBytes;
        sb.nNodes = a.stagedNodes;
        sb.nTris = a.stagedTris;
        if (threadIdx.x == 0) mbarInit(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbarExpectTx(bar, nodeBytes + triBytes);
            const unsigned chunk = 32768u;
            for (unsigned o = 0; o < nodeBytes; o += chunk)
                bulkCopyToShared(sb.nodesAddr + o, reinterpret_cast<const char*>(a.stagedNodesSrc) + o, min(chunk, nodeBytes - o), bar);
            for (unsigned o = 0; o < triBytes; o += chunk)
                bulkCopyToShared(sb.trisAddr + o, reinterpret_cast<const char*>(a.stagedTrisSrc) + o, min(chunk, triBytes - o), bar);
        }
        mbarWait(bar, 0);
    }

    const int S = sc.shadowRaysPerHit;
    const int R = a.tileRays;
    const bool showNormals = sc.flags & FLAG_SHOW_NORMALS;
    const TileScratch ts = carveTileScratch(a.slab + (size_t)(blockIdx.x * NG + group) * a.slabBytes, a.capRays, a.capInterior, R, S);
    const Slots slots{ a.gen.slots, ts.localSlots };

    long long total;
    if (GEN == GEN_PRIMARY) total = primaryRayCount(a.gen);
    else if (GEN == GEN_SSAA) total = 4LL * min(a.ctr->ssaaPixels, a.gen.count);
    else total = a.nUser;
    if (GEN != GEN_QUEUE && blockIdx.x == 0 && threadIdx.x == 0) a.lv[0].nRays = (int)min(total, 0x7fffffffLL);

    WalkAcc acc;

    for (;;) {
        if (gtid == 0) {
            gs.tile = atomicAdd(a.tileCursor, 1ull);
            gs.interiors = 0;
            gs.lv[0] = TileLevelCtr{ 0u, 0u, 0, 0 };
            gs.lv[1] = TileLevelCtr{ 0u, 0u, 0, 0 };
        }
        groupSync<NG>(group);
        const long long base = (long long)gs.tile * R;
        if (base >= total) break;
        int n = (int)min((long long)R, total - base);
        int lastLevel = 0;

        for (int depth = 0; depth < (DEEP ? a.levels : 1); ++depth) {
            TileLevelCtr& lc = gs.lv[depth & 1];
            RayQueue q = ts.q[depth & 1];
            if (GEN == GEN_QUEUE && depth == 0) q = RayQueue{ a.userQ.o + base, a.userQ.d + base, a.userQ.dest + base };
            const RayQueue next = ts.q[(depth + 1) & 1];
            lastLevel = depth;
            int nSurf = 0;

            // Two walks per level — closest hit (Render::trace), then the shadow rays of its hits, light-major — through ONE
            // call site (see walkRays: one copy of the traversal loop in the kernel).  The surface stage sits between them.
#pragma unroll 1
            for (int phase = 0; phase < 2; ++phase) {
                const bool any = phase != 0;
                unsigned* cursor = &lc.cursorClosest;
                long long count = n;
                int genKind = (depth == 0) ? GEN : GEN_QUEUE;
                if (any) {
                    // the other parity's counters are idle now: every thread has read the previous level's nNext before its walk
                    if (gtid == 0) gs.lv[(depth + 1) & 1] = TileLevelCtr{ 0u, 0u, 0, 0 };
                    // ---- surface records, hits compacted (castRay :762-775); misses of queued rays -> skybox ----
                    int* bins = nullptr;
                    if (DEEP && a.sortSurfaces && !showNormals) {
                        // counting sort by (material, incoming direction octant): count, exclusive scan, then the surface stage
                        // takes slots from its bin's running offset
                        if (gtid < 32) gs.bins[gtid] = 0;
                        groupSync<NG>(group);
                        for (int i = gtid; i < n; i += kTileThreads) {
                            const int obj = ts.hits.obj[i];
                            if (obj >= 0) atomicAdd(&gs.bins[surfaceSortKey(sc, obj, q.d[i])], 1);
                        }
                        groupSync<NG>(group);
                        if (gtid < 32) {
                            const int own = gs.bins[gtid];
                            int incl = own;
                            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (gtid >= o) incl += v; }
                            gs.bins[gtid] = incl - own;
                            if (gtid == 31) lc.nSurf = incl;
                        }
                        groupSync<NG>(group);
                        bins = gs.bins;
                    }
                    surfaceStage<(DEEP || GEN == GEN_QUEUE)>(sc, q, ts.hits, ts.surf, slots, n, gtid, kTileThreads, &lc.nSurf, (depth > 0 || GEN == GEN_QUEUE) ? 1 : 0, bins);
                    groupSync<NG>(group);
                    nSurf = lc.nSurf;
                    if (showNormals || S <= 0 || nSurf <= 0) break;
                    cursor = &lc.cursorShadow;
                    count = (long long)nSurf * S;
                    genKind = GEN_QUEUE;
                }
                walkRays<STATS, kTileThreads, unsigned, STAGED>(any, genKind, sc, q, ts.hits, ts.surf, ts.vis, nSurf, a.gen, base, slots, cursor, count, stack, acc, sb);
                groupSync<NG>(group);
            }
            if (!showNormals) {
                // ---- shade; Reflective / Transparent hits leave a record and queue their children ----
                const ShadeOut so{ next, a.capRays, &lc.nNext, ts.interiors, a.capInterior, &gs.interiors, nullptr, nullptr, R, R + 2 * a.capInterior,
                    true, &a.ctr->overflow };
                shadeStage<DEEP>(sc, q, ts.surf, ts.vis, depth, nSurf, gtid, kTileThreads, slots, so);
                groupSync<NG>(group);
            }
            if (gtid == 0) {
                if (depth > 0) atomicAdd(&a.lv[depth].nRays, n);
                if (nSurf > 0) atomicAdd(&a.lv[depth].nSurf, nSurf);
                gs.levelEnd[depth] = min(gs.interiors, a.capInterior);
            }
            if (!DEEP) break;
            n = min(lc.nNext, a.capRays);
            if (n == 0) break;
        }

        if (DEEP) {
            // ---- castRay's return path: fold records, deepest level first ----
            groupSync<NG>(group);     // levelEnd of the last level is visible
            for (int l = lastLevel; l >= 0; --l) {
                const int first = l > 0 ? gs.levelEnd[l - 1] : 0, last = gs.levelEnd[l];
                if (last > first) combineStage(ts.interiors, first, last, gtid, kTileThreads, slots, true);
                groupSync<NG>(group);
            }
        }
        if (GEN == GEN_SSAA) {
            // ---- mean of the 4 re-traced samples (SSAAworker :525-536) ----
            const int nPix = (int)min((long long)R, total - base) >> 2;
            for (int p = gtid; p < nPix; p += kTileThreads) {
                V3 c = mk(0.0f, 0.0f, 0.0f);
#pragma unroll
                for (int k = 0; k < 4; ++k) c = c + loadSlot(slots, localDest(4 * p + k));
                storeSlot(a.gen.slots, a.gen.list[(base >> 2) + p], c / 4.0f);
            }
        }
        groupSync<NG>(group);     // everybody is done with this tile's shared state
    }

    flushWalkAcc(a.ctr, acc, STATS);
    if (STATS) {
        atomicAdd(&a.ctr->walkNodes[0], acc.nNodes[0]); atomicAdd(&a.ctr->walkTris[0], acc.nTris[0]); atomicAdd(&a.ctr->walkEligibility[0], acc.nElig[0]);
        atomicAdd(&a.ctr->walkNodes[1], acc.nNodes[1]); atomicAdd(&a.ctr->walkTris[1], acc.nTris[1]); atomicAdd(&a.ctr->walkEligibility[1], acc.nElig[1]);
    }
}

} // namespace rtk
