// objects.cpp — .obj / texture loading and the split-tree builder of the host library.
//
// Value-exact restatement of Mesh::loadOBJ (reference src/objects.cpp:177-394) and
// AccelerationStructure::setup + SAH helpers (objects.cpp:470-526, 633-763): their outputs —
// world-space vertices, normals, tangents, root bounds, tree shape, leaf order — are the inputs of
// the GPU hot path, so every float operation keeps the reference's order.
#include "objects.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <limits>

#include "../../../include/rtb.h"
#include "util.h"

namespace {

// face-index reader with the reference's semantics (objects.cpp:207-215): skip blanks, skip one
// '/', accumulate until blank, '/' or end; an empty field yields 0
size_t nextIndex(const char*& p)
{
    size_t v = 0;
    while (*p == ' ') ++p;
    if (*p == '/') ++p;
    for (; *p && *p != ' ' && *p != '/'; ++p) v = v * 10 + (size_t)(*p - '0');
    return v;
}

// `sscanf("%f ...")` equivalent: n whitespace-separated floats
bool readFloats(const char* s, float* out, int n)
{
    for (int i = 0; i < n; ++i) {
        char* end = nullptr;
        out[i] = strtof(s, &end);
        if (end == s) return false;
        s = end;
    }
    return true;
}

Triangle makeTriangle(const Vec3f& a, const Vec3f& b, const Vec3f& c)
{
    Triangle t;
    t.a = a; t.b = b; t.c = c;
    t.n_a = t.n_b = t.n_c = (b - a).crossProduct(c - a);   // unnormalised face normal (objects.cpp:17-21)
    return t;
}

Triangle makeTriangle(const Vec3f& a, const Vec3f& b, const Vec3f& c, const Vec3f& na, const Vec3f& nb, const Vec3f& nc)
{
    Triangle t = makeTriangle(a, b, c);
    t.n_a = na; t.n_b = nb; t.n_c = nc;
    return t;
}

Triangle makeTriangle(const Vec3f& a, const Vec3f& b, const Vec3f& c, const Vec3f& na, const Vec3f& nb, const Vec3f& nc,
    const Vec2f& ta, const Vec2f& tb, const Vec2f& tc)
{
    Triangle t = makeTriangle(a, b, c, na, nb, nc);
    t.t_a = ta; t.t_b = tb; t.t_c = tc;
    // per-triangle tangent frame for normal mapping (objects.cpp:43-55); unnormalised, may be
    // inf/NaN for degenerate uv triangles exactly like the reference
    const Vec3f e1 = b - a, e2 = c - a;
    const Vec2f d1 = tb - ta, d2 = tc - ta;
    const float f = 1.0f / (d1.x * d2.y - d2.x * d1.y);
    t.tangent = { f * (d2.y * e1.x - d1.y * e2.x), f * (d2.y * e1.y - d1.y * e2.y), f * (d2.y * e1.z - d1.y * e2.z) };
    t.bitangent = { f * (-d2.x * e1.x + d1.x * e2.x), f * (-d2.x * e1.y + d1.x * e2.y), f * (-d2.x * e1.z + d1.x * e2.z) };
    return t;
}

} // namespace

bool Mesh::loadOBJ(const std::string& filename, const Options& opts)
{
    const auto tStart = std::chrono::steady_clock::now();
    const Matrix44f rMatrix = Matrix44f::rotationDeg(rot);

    std::ifstream in(filename, std::ios::in);
    if (!in.good()) {
        // the reference prints and carries on with an empty mesh (objects.cpp:219-222)
        printf("Error, failed to load obj, filename: %s\n", filename.c_str());
        return false;
    }
    ac = std::make_unique<AccelerationStructure>();
    std::vector<Vec3f> positions, normals;
    std::vector<Vec2f> uvs;
    Vec3f lo(std::numeric_limits<float>::max());
    Vec3f hi(std::numeric_limits<float>::min());   // smallest POSITIVE float, as in objects.cpp:231
    bool placed = false;
    allTris.clear();

    auto vtx = [&](size_t i) -> const Vec3f& { return positions.at(i - 1); };
    auto nrm = [&](size_t i) -> const Vec3f& { return normals.at(i - 1); };
    auto tex = [&](size_t i) -> const Vec2f& { return uvs.at(i - 1); };

    std::string line;
    std::vector<size_t> vi, ti, ni;   // per-face index lists, reused
    do {
        std::getline(in, line);
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        if (line.empty()) continue;

        // first whitespace-delimited token, at most 31 characters (the reference reads it with sscanf("%s"), objects.cpp:241-251)
        char tag[32] = { 0 };
        {
            const char* c = line.c_str();
            while (*c == ' ' || *c == '\t' || *c == '\r' || *c == '\n' || *c == '\v' || *c == '\f') ++c;
            size_t n = 0;
            while (*c && !(*c == ' ' || *c == '\t' || *c == '\r' || *c == '\n' || *c == '\v' || *c == '\f') && n < 31) tag[n++] = *c++;
        }
        const size_t skip = strlen(tag) + 1;
        const char* body = line.c_str() + std::min(skip, line.size());

        if (strcmp(tag, "v") == 0) {
            float p[3];
            if (!readFloats(body, p, 3)) throw rtb::Error(RTB_ERR_PARSE, "bad 'v' line in " + filename);
            lo.x = std::min(p[0], lo.x); lo.y = std::min(p[1], lo.y); lo.z = std::min(p[2], lo.z);
            hi.x = std::max(p[0], hi.x); hi.y = std::max(p[1], hi.y); hi.z = std::max(p[2], hi.z);
            positions.emplace_back(p[0], p[1], p[2]);
        } else if (strcmp(tag, "vn") == 0) {
            float p[3];
            if (!readFloats(body, p, 3)) throw rtb::Error(RTB_ERR_PARSE, "bad 'vn' line in " + filename);
            normals.push_back(Vec3f(p[0], p[1], p[2]).normalize());
        } else if (strcmp(tag, "vt") == 0) {
            float p[2];
            if (!readFloats(body, p, 2)) throw rtb::Error(RTB_ERR_PARSE, "bad 'vt' line in " + filename);
            uvs.emplace_back(p[0], p[1]);
        } else if (strcmp(tag, "f") == 0) {
            if (!placed) {
                // First face: fit every vertex read so far into `size` keeping proportions, rotate,
                // translate (objects.cpp:282-331).
                placed = true;
                const Vec3f range = hi - lo;
                Vec3f fit = size;
                const bool flat = range.x < opts.bias || range.y < opts.bias || range.z < opts.bias;
                if (!flat) {
                    const Vec3f stretch = size / range;
                    const float least = std::min(stretch.x, std::min(stretch.y, stretch.z));
                    if (least == stretch.x) {
                        fit.y = fit.x / (range.x / range.y);
                        fit.z = fit.x / (range.x / range.z);
                    } else if (least == stretch.y) {
                        fit.x = fit.y / (range.y / range.x);
                        fit.z = fit.y / (range.y / range.z);
                    } else {
                        fit.x = fit.z / (range.z / range.x);
                        fit.y = fit.z / (range.z / range.y);
                    }
                }
                for (Vec3f& v : positions) {
                    v.x = fit.x * ((v.x - lo.x) / range.x - 0.5f);
                    v.y = fit.y * ((v.y - lo.y) / range.y - 0.5f);
                    v.z = fit.z * ((v.z - lo.z) / range.z - 0.5f);
                    v = rMatrix.multVecMatrix(v);
                    v.x += pos.x; v.y += pos.y; v.z += pos.z;
                    if (range.x < opts.bias) v.x = pos.x;
                    if (range.y < opts.bias) v.y = pos.y;
                    if (range.z < opts.bias) v.z = pos.z;
                }
                for (Vec3f& n : normals) n = rMatrix.multVecMatrix(n);
                // Root bounds: the reference rotates the SIZE VECTOR, not the box (objects.cpp:328-330),
                // so rotated meshes poke outside their root box; reproduced on purpose.
                Vec3f ext = rMatrix.multVecMatrix(fit);
                ext = Vec3f(std::fabs(ext.x), std::fabs(ext.y), std::fabs(ext.z));
                ac->setBounds(pos - ext / 2.0f, pos + ext / 2.0f);
            }

            int slashes = 0;
            for (const char* p = body; *p; ++p) slashes += (*p == '/');
            const char* p = body;
            vi.clear(); ti.clear(); ni.clear();
            if (slashes == 0) {
                for (size_t v; (v = nextIndex(p)) > 0;) vi.push_back(v);
                for (size_t i = 1; i + 1 < vi.size(); ++i)
                    allTris.push_back(makeTriangle(vtx(vi[0]), vtx(vi[i]), vtx(vi[i + 1])));
            } else if (slashes % 2 == 0) {
                for (size_t v; (v = nextIndex(p)) > 0;) {
                    const size_t t = nextIndex(p);
                    const size_t n = nextIndex(p);
                    vi.push_back(v);
                    if (t > 0) ti.push_back(t);
                    if (n > 0) ni.push_back(n);
                }
                for (size_t i = 1; i + 1 < vi.size(); ++i) {
                    if (ni.empty())
                        allTris.push_back(makeTriangle(vtx(vi[0]), vtx(vi[i]), vtx(vi[i + 1])));
                    else if (ti.empty())
                        allTris.push_back(makeTriangle(vtx(vi[0]), vtx(vi[i]), vtx(vi[i + 1]),
                            nrm(ni.at(0)), nrm(ni.at(i)), nrm(ni.at(i + 1))));
                    else
                        allTris.push_back(makeTriangle(vtx(vi[0]), vtx(vi[i]), vtx(vi[i + 1]),
                            nrm(ni.at(0)), nrm(ni.at(i)), nrm(ni.at(i + 1)),
                            tex(ti.at(0)), tex(ti.at(i)), tex(ti.at(i + 1))));
                }
            } else {
                printf("Unhandled slash count: %d\n", slashes);
            }
        }
    } while (in.good());

    const auto tParsed = std::chrono::steady_clock::now();
    ac->setup(allTris, opts);
    if (options::enableOutput) {   // the reference times this phase with a Timer("OBJ loading") (objects.cpp:217)
        const auto tDone = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return (long long)std::chrono::duration_cast<std::chrono::milliseconds>(b - a).count(); };
        printf("%-18s%lld ms\n", "OBJ loading", ms(tStart, tDone));
        if (getenv("RTB_HOST_TIMING")) printf("  parse %lld ms, tree %lld ms, %zu triangles\n", ms(tStart, tParsed), ms(tParsed, tDone), allTris.size());
    }
    return true;
}

static bool loadMap(const std::string& filename, TextureRGB8& t)
{
    if (!options::useTextures) return false;
    rtb::loadBMP(filename, t.rgb, t.width, t.height);
    return true;
}

// The reference expands maps to float at load (objects.cpp:396-458): byte/256, normal map to
// (2x-1, -(2y-1), z) normalised, specular to (r+g+b)/3.  Here the bytes stay bytes — the same
// arithmetic runs per lookup on the device — which is 4x less memory and exact.
bool Mesh::loadDiffuseMap(const std::string& filename) { return loadMap(filename, diffuseMap); }
bool Mesh::loadNormalMap(const std::string& filename) { return loadMap(filename, normalMap); }
bool Mesh::loadSpecularMap(const std::string& filename) { return loadMap(filename, specularMap); }

// ------------------------------------------------------------------------------------------------
// Split tree
// ------------------------------------------------------------------------------------------------

static inline bool touchesLow(const Triangle& t, int axis, float s) { return t.a[axis] <= s || t.b[axis] <= s || t.c[axis] <= s; }
static inline bool touchesHigh(const Triangle& t, int axis, float s) { return t.a[axis] >= s || t.b[axis] >= s || t.c[axis] >= s; }

// cost = nLeft*(s - lo) + nRight*(hi - s); straddlers count on both sides (objects.cpp:633-674)
float AccelerationStructure::calculateSAH(int axis, const std::vector<Triangle>& tris, const std::vector<int>& ids,
    const Vec3f bounds[2], float s)
{
    int nLow = 0, nHigh = 0;
    for (int id : ids) {
        nLow += touchesLow(tris[id], axis, s);
        nHigh += touchesHigh(tris[id], axis, s);
    }
    return nLow * (s - bounds[0][axis]) + nHigh * (bounds[1][axis] - s);
}

// bisection with +-0.05 probes, stops when the interval is below 0.1 ABSOLUTE units (objects.cpp:676-689)
float AccelerationStructure::binarySearchSAH(int axis, const std::vector<Triangle>& tris, const std::vector<int>& ids,
    const Vec3f bounds[2], float left, float right)
{
    for (;;) {
        const float mid = right - (right - left) / 2;
        if (right - left < 0.1f) return mid;
        if (calculateSAH(axis, tris, ids, bounds, mid - 0.05f) < calculateSAH(axis, tris, ids, bounds, mid + 0.05f))
            right = mid;
        else
            left = mid;
    }
}

void AccelerationStructure::setup(const std::vector<Triangle>& tris, const Options& opts)
{
    std::vector<int> ids(tris.size());
    for (size_t i = 0; i < ids.size(); ++i) ids[i] = (int)i;
    Subtree tree;
    build(tree, tris, ids, rootBounds, 1, opts);
    nodes.swap(tree.nodes);
    refs.swap(tree.refs);
}

// appends a subtree built on its own (another thread), shifting its links into `out`'s coordinates
void AccelerationStructure::append(Subtree& out, const Subtree& child)
{
    const int nodeBase = (int)out.nodes.size(), refBase = (int)out.refs.size();
    for (Node n : child.nodes) {
        if (n.right >= 0) n.right += nodeBase;
        else n.firstRef += refBase;
        out.nodes.push_back(n);
    }
    out.refs.insert(out.refs.end(), child.refs.begin(), child.refs.end());
}

// The two halves of a node are independent, so near the root they are built concurrently, each into its own Subtree,
// and concatenated in the reference's order (self, low subtree, high subtree): the result is the sequential one bit for bit.
void AccelerationStructure::build(Subtree& out, const std::vector<Triangle>& tris, std::vector<int>& ids, const Vec3f bounds[2], int depth,
    const Options& opts)
{
    std::vector<Node>& nodes = out.nodes;
    std::vector<int>& refs = out.refs;
    const int self = (int)nodes.size();
    nodes.emplace_back();
    nodes[self].bounds[0] = bounds[0];
    nodes[self].bounds[1] = bounds[1];
    nodes[self].depth = depth;

    auto makeLeaf = [&]() {
        nodes[self].right = -1;
        nodes[self].firstRef = (int)refs.size();
        nodes[self].refCount = (int)ids.size();
        refs.insert(refs.end(), ids.begin(), ids.end());
    };

    // deep enough for this many triangles (objects.cpp:477)
    if (ids.size() <= (size_t)depth * (size_t)opts.acPenalty) { makeLeaf(); return; }

    const Vec3f dim = bounds[1] - bounds[0];
    int axis;
    if (dim.x > dim.y && dim.x > dim.z) axis = 0;
    else if (dim.y > dim.z) axis = 1;
    else axis = 2;

    const float split = binarySearchSAH(axis, tris, ids, bounds, bounds[0][axis], bounds[1][axis]);
    std::vector<int> low, high;
    for (int id : ids) {
        if (touchesLow(tris[id], axis, split)) low.push_back(id);
        if (touchesHigh(tris[id], axis, split)) high.push_back(id);
    }
    // give up when a side is empty or duplication reaches 1.5x (objects.cpp:498)
    if (low.empty() || high.empty() || (double)(low.size() + high.size()) >= ids.size() * 1.5) { makeLeaf(); return; }

    // children are the parent's box cut at the plane — NOT tight boxes (objects.cpp:510-521)
    Vec3f lowBox[2] = { bounds[0], bounds[1] };
    Vec3f highBox[2] = { bounds[0], bounds[1] };
    lowBox[1][axis] = split;
    highBox[0][axis] = split;

    std::vector<int>().swap(ids);   // release before recursing: the dragon's tree is 25 deep
    const bool fork = depth <= 3 && low.size() + high.size() > 50000;
    if (!fork) {
        build(out, tris, low, lowBox, depth + 1, opts);
        nodes[self].right = (int)nodes.size();
        build(out, tris, high, highBox, depth + 1, opts);
        return;
    }
    Subtree lowTree, highTree;
    auto lowJob = std::async(std::launch::async, [&]() { build(lowTree, tris, low, lowBox, depth + 1, opts); });
    build(highTree, tris, high, highBox, depth + 1, opts);
    lowJob.get();
    append(out, lowTree);
    nodes[self].right = (int)nodes.size();
    append(out, highTree);
}
