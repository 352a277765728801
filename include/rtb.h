/* rtb.h — C ABI of the B200 ray-tracing backend ("rtb") that sits behind Scene::render().
 *
 * The reference (holoskii/Rendering) has no FFI: its boundary is the C++ member function
 * `void Scene::render()` (reference include/scene.h:91, src/scene.cpp:595-657), which runs
 * launchWorkers (scene.cpp:470-506) and launchSSAA (scene.cpp:542-593) over a loaded Scene and
 * hands a `Vec3f frameBuffer[h*w]` to saveImage (src/util.cpp:15-76).  This header is the thin
 * `extern "C"` surface a maintainer binds instead: plain pointers and sizes, no C++ or torch types.
 *
 *   host side  (librtb_host.so, dependency-free C++17):  rtb_scene_load / rtb_scene_view / ...
 *       replaces Scene::loadScene (scene.cpp:62-334), Mesh::loadOBJ + AccelerationStructure::setup
 *       (objects.cpp:177-394, 470-526) and flattens the result into the POD `RtbScene` below.
 *   device side (librtb_cuda.so, CUDA sm_100a):            rtb_create / rtb_render / rtb_destroy
 *       replaces launchWorkers + launchSSAA, i.e. everything between "Scene is loaded" and
 *       "frameBuffer is full".  There is NO CPU fallback: every entry point fails with
 *       RTB_ERR_CUDA when no device is usable.
 *
 * All arrays are host pointers owned by the caller for the duration of rtb_create(); the library
 * copies what it needs to device memory.  Floats are IEEE-754 binary32.
 */
#ifndef RTB_H
#define RTB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB_ABI_VERSION 2

/* ---- status codes (reference: LOG_ERROR() prints and exit(-1)s, include/util.h:13-19) ---- */
#define RTB_OK             0
#define RTB_ERR_ARG       -1   /* null / inconsistent argument                                   */
#define RTB_ERR_IO        -2   /* file missing / unreadable (scene, obj, bmp)                    */
#define RTB_ERR_PARSE     -3   /* malformed .scene / .obj (reference would LOG_ERROR)            */
#define RTB_ERR_CUDA      -4   /* CUDA runtime error, or no device: no CPU fallback exists       */
#define RTB_ERR_NOMEM     -5
#define RTB_ERR_UNSUPPORTED -6

/* ---- enums mirror the reference's (include/objects.h:17-18, include/lights.h:11) ---- */
enum { RTB_OBJ_SPHERE = 1, RTB_OBJ_PLANE = 2, RTB_OBJ_MESH = 3 };
enum { RTB_MAT_DIFFUSE = 0, RTB_MAT_REFLECTIVE = 1, RTB_MAT_TRANSPARENT = 2, RTB_MAT_PHONG = 3 };
enum { RTB_LIGHT_DISTANT = 1, RTB_LIGHT_POINT = 2, RTB_LIGHT_AREA = 3 };

/* process-global switches of the reference's `namespace options` (include/options.h:26-36) */
#define RTB_FLAG_BACKFACE_CULLING  (1u << 0)   /* options::useBackfaceCulling (objects.cpp:75)   */
#define RTB_FLAG_USE_AC            (1u << 1)   /* options::useAC (objects.cpp:536)               */
#define RTB_FLAG_USE_SKYBOX        (1u << 2)   /* options::useSkybox (scene.cpp:383)             */
#define RTB_FLAG_SHOW_NORMALS      (1u << 3)   /* options::showNormals (scene.cpp:771)           */
#define RTB_FLAG_ENABLE_SSAA       (1u << 4)   /* options::enableSSAA (scene.cpp:604)            */

/* Camera (include/scene.h:51-65).  rMatrix = mz*my*mx built on the host with sinf/cosf exactly as
 * Camera::getRay does (scene.cpp:24-48), row-major 4x4, applied as row-vector x matrix.
 * scale = tanf(fov*0.5f/180.0f*(float)M_PI), aspect = width/(float)height (scene.cpp:447-448).   */
typedef struct RtbCamera {
    float pos[3];
    float rMatrix[16];
    float scale;
    float aspect;
} RtbCamera;

/* One scene object (include/objects.h:24-46, 165-191).  `mesh` indexes RtbScene.meshes for
 * RTB_OBJ_MESH, else -1.  Sphere: pos + r2 (= powf(r,2), scene.cpp:294).  Plane: pos + normal
 * (NOT normalised when it comes from the scene file, scene.cpp:299-301).                          */
typedef struct RtbObject {
    int32_t type;
    int32_t material;
    float   color[3];
    float   ior;
    float   ambient, diffuse, specular, nSpecular;
    float   pos[3];
    float   r2;
    float   normal[3];
    int32_t mesh;
} RtbObject;

/* One light (include/lights.h:19-73).  v = dir (distant, un-normalised from file) or pos (point /
 * area centre).  Area lights: sample points are precomputed on the host exactly as
 * AreaLight::setPoints (lights.cpp:46-63) and stored in RtbScene.areaPoints[pointOffset ...].     */
typedef struct RtbLight {
    int32_t type;
    float   color[3];
    float   intensity;
    float   v[3];
    int32_t pointOffset;
    int32_t pointCount;
} RtbLight;

/* Node of the reference's per-mesh split tree (include/objects.h:128-163), DFS pre-order:
 * the left child of inner node k is k+1, `right` is the right child's index; `right < 0` marks a
 * leaf whose triangle references are refs[firstRef .. firstRef+refCount).  Leaves are numbered in
 * the reference's traversal order (left before right, objects.cpp:601-619) so a reference slot
 * index orders exact-t ties the way the recursive strict `<` does.                                 */
typedef struct RtbNode {
    float   lo[3];
    float   hi[3];
    int32_t right;
    int32_t firstRef;
    int32_t refCount;
    int32_t depth;
} RtbNode;

/* 8-bit RGB image as loadBMP leaves it (util.cpp:78-113): 3 bytes/texel, R,G,B order (after the
 * reference's B<->R swap), row 0 = first row in the file (BMP bottom row), no padding.            */
typedef struct RtbImage {
    const uint8_t* rgb;
    int32_t width, height;
} RtbImage;

/* One triangle mesh (include/objects.h:69-121): SoA over triangles, in .obj face order.           */
typedef struct RtbMesh {
    int32_t nTris, nNodes, nRefs;
    const float*   pos;       /* nTris*9  : a.xyz b.xyz c.xyz (world space, objects.cpp:306-320)   */
    const float*   nrm;       /* nTris*9  : n_a n_b n_c                                            */
    const float*   uv;        /* nTris*6  : t_a t_b t_c                                            */
    const float*   tan;       /* nTris*6  : tangent.xyz bitangent.xyz (objects.cpp:43-55)          */
    const RtbNode* nodes;     /* nNodes                                                            */
    const int32_t* refs;      /* nRefs triangle indices, leaf by leaf                              */
    RtbImage diffuseMap, normalMap, specularMap;   /* rgb == NULL when not loaded                  */
} RtbMesh;

typedef struct RtbScene {
    int32_t   abiVersion;          /* RTB_ABI_VERSION                                              */
    int32_t   width, height;       /* Options::width/height (include/options.h:12)                 */
    float     bias;                /* Options::bias                                                */
    int32_t   maxRayDepth;         /* Options::maxRayDepth                                         */
    float     backgroundColor[3];
    uint32_t  flags;               /* RTB_FLAG_*                                                   */
    RtbCamera camera;
    int32_t   nObjects, nLights, nMeshes, nAreaPoints;
    const RtbObject* objects;
    const RtbLight*  lights;
    const RtbMesh*   meshes;
    const float*     areaPoints;   /* nAreaPoints*3                                                */
    RtbImage  skybox[6];           /* left,front,right,back,top,bottom (scene.cpp:336-360)         */
} RtbScene;

/* kernel kinds for RtbStats.msKernel / launchesKernel */
enum { RTB_K_RAYGEN = 0, RTB_K_TRACE = 1, RTB_K_SURFACE = 2, RTB_K_SHADOW = 3, RTB_K_SHADE = 4, RTB_K_COMBINE = 5,
       RTB_K_SOBEL = 6, RTB_K_OUTPUT = 7,
       RTB_K_TILE = 8,        /* the tile pipeline's fused kernel, pass 1 (all recursion levels of every tile)   */
       RTB_K_TILE_SSAA = 9,   /* the same kernel over the 4 samples of the flagged pixels, incl. their mean      */
       RTB_NKINDS = 10 };

/* Work counters of one rtb_render call (64-bit: the reference's are int and wrap, stats.h:11-16). */
typedef struct RtbStats {
    uint64_t rays;          /* Render::trace invocations: primary + secondary + shadow + SSAA     */
    uint64_t primaryRays, secondaryRays, shadowRays, ssaaPixels;
    uint64_t boxTests;      /* reference-walk box / triangle tests of ALL rays; only filled when   */
    uint64_t triTests;      /* the handle was created with RTB_CREATE_COUNTERS                     */
    uint64_t boxTestsShadow, triTestsShadow;      /* the shadow rays' share of the two above       */
    uint64_t h2dBytes, d2hBytes;                  /* host<->device bytes copied inside the call    */
    uint64_t shadowRaysSkipped;   /* shadow rays (counted in `rays`) whose visibility cannot affect the pixel
                                     and that the fast path therefore does not trace                      */
    uint64_t backgroundPixels;    /* primary rays (counted in `rays`) that lie outside the screen-space bounds of the
                                     geometry and are resolved by the background pre-fill instead of a traversal     */
    float    msBuildSearchBvh;    /* time rtb_create spent building the search BVHs (host: wall clock; device: CUDA events);
                                     a property of the handle, repeated in every call's stats                              */
    uint32_t kernelLaunches;
    uint32_t levels;
    float    msPass1, msSobel, msSSAA, msTotal;   /* CUDA-event times on the render stream         */
    float    msKernel[RTB_NKINDS];                /* device time per kernel kind (RTB_CREATE_KERNEL_TIMING) */
    uint32_t launchesKernel[RTB_NKINDS];
    /* the fast path's OWN work (RTB_CREATE_WALK_STATS): search-BVH nodes fetched (64 B each), triangles tested (48 B
     * each), eligibility evaluations; index 0 = closest-hit rays, 1 = shadow rays                                       */
    uint64_t walkNodes[2], walkTris[2], walkEligibility[2];
} RtbStats;

typedef struct RtbHandle RtbHandle;
typedef struct RtbHostScene RtbHostScene;

/* ===================== host side: librtb_host.so ===================== */

/* Parse a .scene file the way Scene::loadScene does (scene.cpp:62-334), load meshes/textures,
 * build each mesh's tree bit-for-bit like AccelerationStructure::setup (objects.cpp:470-526) and
 * flatten.  Relative asset paths are tried against the cwd first (reference behaviour), then
 * against the directory of the scene file.                                                         */
int  rtb_scene_load(const char* scenePath, RtbHostScene** out);
/* Same, from scene text in memory (assetDir resolves relative paths; may be NULL).                 */
int  rtb_scene_parse(const char* sceneText, const char* assetDir, RtbHostScene** out);
const RtbScene* rtb_scene_view(const RtbHostScene* hs);
const char*     rtb_scene_image_name(const RtbHostScene* hs);
void rtb_scene_free(RtbHostScene* hs);
/* Per-mesh tree statistics {nodes, leaves, refs, maxLeaf, maxDepth, trisOutsideRoot}.              */
int  rtb_scene_tree_stats(const RtbHostScene* hs, int mesh, int64_t out[6]);
/* saveImage contract (util.cpp:15-76) with the well-defined quantisation (uint8)(clamp(v)*255).    */
int  rtb_save_bmp(const char* path, const float* fb, int width, int height);
/* Camera constants for a position / Euler rotation (degrees) / field of view, computed exactly like the loader does
 * for the [options] keys position, rotation, fov (scene.cpp:170-171,182-185 -> Camera::getRay :24-48, renderWorker :447-448).
 * Feed the result to rtb_set_camera to move the camera of a resident scene between frames.                          */
int  rtb_camera_from_angles(const float pos[3], const float rotDeg[3], float fovDeg, int width, int height, RtbCamera* out);
/* Writes header + the pixel bytes produced by rtb_render_bgr8 for the full frame.                   */
int  rtb_save_bmp_bgr8(const char* path, const uint8_t* bgr, int width, int height);
const char* rtb_host_last_error(void);

/* ===================== device side: librtb_cuda.so ===================== */

#define RTB_CREATE_DEFAULT   0u
#define RTB_CREATE_COUNTERS  (1u << 0)   /* count box / triangle tests (slower; parity of work)   */
#define RTB_CREATE_EXACT_WALK (1u << 1)  /* traverse exactly like objects.cpp:587-631 (no culling) */
#define RTB_CREATE_WALK_STATS (1u << 3)  /* fill RtbStats.walkNodes / walkTris / walkEligibility (slower kernels)    */
#define RTB_CREATE_WAVEFRONT (1u << 4)   /* frame-wide level pipeline (one launch per stage and recursion level) instead of
                                            the default tile pipeline (whole recursion per tile inside one kernel); same bits */
#define RTB_CREATE_DEVICE_BVH (1u << 5)  /* build the search BVH on the device (linear BVH: Morton keys, radix sort, Karras tree) instead of
                                            the host's binned SAH: ~100x faster to build, somewhat slower to traverse, same frames   */
#define RTB_CREATE_KERNEL_TIMING (1u << 2) /* fill RtbStats.msKernel: CUDA events around every launch (costs ~6 us
                                              of stream time per launch, so it is off by default)               */

/* Upload a flattened scene to `device` and build the device acceleration data.                     */
int  rtb_create(const RtbScene* scene, int device, uint32_t createFlags, RtbHandle** out);

/* Replace the camera of a resident scene (the reference is single-shot: one Scene, one render(); a persistent handle
 * renders a camera sweep without re-uploading geometry or textures).  Takes effect from the next rtb_render* call.   */
int  rtb_set_camera(RtbHandle* h, const RtbCamera* camera);

/* Render rows [y0,y1) of the frame: pass 1 (launchWorkers) + Sobel + SSAA (launchSSAA), with the
 * reference's quirks (last row/column black, pixel centre x+1.0, Sobel over unclamped floats).
 * `fb` receives (y1-y0)*width*3 floats, row y0 first; it is a HOST pointer (copied back inside the
 * call) when fbOnDevice == 0, else a device pointer on the handle's device.  `pass1` (optional,
 * same shape and placement) receives the frame before SSAA.  `stream` is a cudaStream_t (NULL =
 * the handle's own stream).  For multi-GPU strips use rtb_render_strips.                           */
int  rtb_render(RtbHandle* h, int y0, int y1, float* fb, float* pass1, int fbOnDevice,
                void* stream, RtbStats* stats);

/* Same frame, delivered as the pixel bytes saveImage writes after the 54-byte BMP header (util.cpp:46-56):
 * rows bottom-up (image row y1-1 first), B,G,R per pixel, channel = (uint8)(clamp(0,1,v)*255), each row padded
 * to a multiple of 4 bytes — (y1-y0) * ((3*width+3)&~3) bytes.  The conversion runs on the device, so a host
 * buffer (onDevice == 0) receives a quarter of the bytes rtb_render copies back.                            */
int  rtb_render_bgr8(RtbHandle* h, int y0, int y1, uint8_t* bgr, int onDevice, void* stream, RtbStats* stats);

/* The showAC debug view (options::showAC, scene.cpp:607-635): fb = per-pixel count of reference-tree boxes passed by
 * the primary ray's line (Scene::countAC :659-669) / the largest count, on all three channels; every pixel of the
 * frame, pixel centre x+0.5.  `counts` (optional, width*height int32) receives the raw counts.                      */
int  rtb_render_ac(RtbHandle* h, float* fb, int32_t* counts, int onDevice, void* stream, RtbStats* stats);

/* Render the rows owned by `rank` under a cyclic strip partition: strip s (stripRows rows) belongs
 * to rank s % worldSize.  Output is compact: owned rows in ascending order; returns their count in
 * *nRowsOut.  One halo row either side of every strip is rendered locally for the Sobel window.    */
int  rtb_render_strips(RtbHandle* h, int stripRows, int rank, int worldSize, float* fb,
                       int fbOnDevice, void* stream, int* nRowsOut, RtbStats* stats);
/* Same partition, but every owned row y is written at frame + y*width*3 of a FULL-FRAME device buffer.  `frame`
 * may be a peer-mapped pointer to the root rank's framebuffer (CUDA IPC / symmetric memory over NVLink): the strips
 * then land in place through the output kernel's own stores and no separate gather or un-permute runs.  The caller
 * synchronises the ranks afterwards (a barrier).                                                                    */
int  rtb_render_strips_to_frame(RtbHandle* h, int stripRows, int rank, int worldSize, float* frame, void* stream, RtbStats* stats);
/* The same two calls split in halves for frame loops: *_begin plans the frame and ENQUEUES everything on the stream
 * (kernels, output, counter read-back) without waiting; the caller may enqueue its own work behind it (the multi-GPU
 * exchange barrier, the next frame's camera) and then calls rtb_render_end, which waits, re-runs the frame with larger
 * queues in the rare overflow case, and returns the statistics.  One frame in flight per handle.  `fb` / `frame` must
 * stay valid until rtb_render_end returns.                                                                          */
int  rtb_render_begin(RtbHandle* h, int y0, int y1, float* fb, int fbOnDevice, void* stream);
int  rtb_render_strips_to_frame_begin(RtbHandle* h, int stripRows, int rank, int worldSize, float* frame, void* stream);
int  rtb_render_end(RtbHandle* h, RtbStats* stats);
/* Frame loop with the output pipelined (camera sweeps): like rtb_render_bgr8 into a HOST buffer (pinned for a truly
 * asynchronous copy), but the bytes leave on a second stream from one of two staging buffers, so the device-to-host copy
 * of frame i overlaps the kernels of frame i+1.  rtb_render_end returns when the frame's kernels are done (statistics);
 * the bytes of that frame are complete after rtb_output_sync (or after the rtb_render_end of the frame after next).
 * Use at least two host buffers in turn.  Runs on the handle's own stream.                                          */
int  rtb_render_bgr8_begin(RtbHandle* h, int y0, int y1, uint8_t* bgrHost);
int  rtb_output_sync(RtbHandle* h);
/* saveImage's conversion (see rtb_render_bgr8) of an assembled full float frame resident on the handle's device.    */
int  rtb_frame_to_bgr8(RtbHandle* h, const float* frame, uint8_t* bgr, int onDevice, void* stream);
/* Number of rows rank owns under that partition with the strips counted from row 0 (for sizing buffers: an upper
 * bound is ceil(height / (stripRows*worldSize)) * stripRows + stripRows for any origin).                            */
int  rtb_strip_rows_owned(int height, int stripRows, int rank, int worldSize);
/* The strips of rtb_render_strips* are counted from the first image row that can contain geometry (so the rows that
 * cost something are dealt evenly): strip s = floor((y - origin) / stripRows) belongs to rank s mod worldSize.
 * rtb_strip_origin returns that row for the handle's current camera; rtb_strip_rows lists (rowsOut, may be NULL) and
 * counts the rows of `rank` for a given origin.                                                                     */
int  rtb_strip_origin(const RtbHandle* h);
int  rtb_strip_rows(int height, int stripRows, int origin, int rank, int worldSize, int32_t* rowsOut);

/* Closest-hit query on caller-supplied rays (orig.xyz dir.xyz per ray): the device equivalent of
 * Render::trace (scene.cpp:724-756).  out: per ray {t,u,v} floats and {object,tri} ints (-1 miss). */
int  rtb_trace(RtbHandle* h, const float* rays, int nRays, float* tuv, int32_t* objTri);
/* Render::castRay (scene.cpp:758-946) on caller-supplied rays; out rgb per ray.                    */
int  rtb_cast(RtbHandle* h, const float* rays, int nRays, float* rgb);

int  rtb_device_of(const RtbHandle* h);
void rtb_destroy(RtbHandle* h);
const char* rtb_last_error(void);
int  rtb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RTB_H */
