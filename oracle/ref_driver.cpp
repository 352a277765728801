// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, not product code.
//
// Library-mode driver for the UNMODIFIED reference renderer.  It is compiled together with the
// reference's own sources where they lie (/root/reference/src/{scene,objects,lights,util}.cpp,
// main.cpp omitted) by oracle/Makefile into oracle/_ref/ref_driver (git-ignored, ships to the GPU
// box).  Nothing under rendering_b200/ may link or execute it; only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs do.
//
// All reference members are public (include/scene.h:68-100), so the driver can run the two render
// passes separately and dump float32 framebuffers, bypassing saveImage()'s char-cast UB
// (src/util.cpp:52).
//
//   ref_driver render <scene> <out_prefix> [workers]   pass-1 + final float32 framebuffers, timings
//   ref_driver bench  <scene> <repeat> [workers]       time launchWorkers + launchSSAA, best/all
//   ref_driver hits   <scene> <out.bin>                per-pixel primary Render::trace record
//   ref_driver tree   <scene> <out.bin>                per-mesh tree + triangle dump (DFS pre-order)
//   ref_driver cast   <scene> <rays.f32> <out.f32>     Render::castRay on caller-supplied rays
//   ref_driver stats  <scene> [workers]                reference's own counters (collectStatistics)
//   ref_driver ac     <scene> <out.i32>                per-pixel Scene::countAC of the showAC debug view (scene.cpp:607-635)
#include "scene.h"
#include "stats.h"
#include "util.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

double nowMs()
{
	using namespace std::chrono;
	return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

void dumpFloats(const std::string& path, const Vec3f* fb, size_t n)
{
	FILE* f = fopen(path.c_str(), "wb");
	if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
	fwrite(fb, sizeof(Vec3f), n, f);
	fclose(f);
}

struct TreeStats { long nodes = 0, leaves = 0, refs = 0, maxLeaf = 0, maxDepth = 0; };

void walk(const AccelerationStructure* n, int depth, TreeStats& st)
{
	st.nodes++;
	if (depth > st.maxDepth) st.maxDepth = depth;
	if (n->left) {
		walk(n->left.get(), depth + 1, st);
		walk(n->right.get(), depth + 1, st);
	} else {
		st.leaves++;
		st.refs += (long)n->tris.size();
		if ((long)n->tris.size() > st.maxLeaf) st.maxLeaf = (long)n->tris.size();
	}
}

void setWorkers(Scene& s, int workers)
{
	if (workers > 0) s.options.nWorkers = workers;
}

} // namespace

int main(int argc, char** argv)
{
	if (argc < 3) {
		fprintf(stderr, "usage: ref_driver render|bench|hits|tree|cast|stats <scene> ...\n");
		return 2;
	}
	const std::string mode = argv[1];
	const std::string scenePath = argv[2];
	options::outputProgress = false;

	if (mode == "stats") options::collectStatistics = true;
	double t0 = nowMs();
	Scene s(scenePath);
	double loadMs = nowMs() - t0;
	options::outputProgress = false;
	options::enableOutput = false;
	const size_t w = s.options.width, h = s.options.height;

	if (mode == "render") {
		if (argc < 4) return 2;
		const std::string prefix = argv[3];
		setWorkers(s, argc > 4 ? atoi(argv[4]) : 1);
		Vec3f* fb = new Vec3f[w * h];
		double a = nowMs();
		s.launchWorkers(fb);
		double renderMs = nowMs() - a;
		dumpFloats(prefix + ".pass1.f32", fb, w * h);
		a = nowMs();
		s.launchSSAA(fb);
		double ssaaMs = nowMs() - a;
		dumpFloats(prefix + ".final.f32", fb, w * h);
		double sum = 0;
		for (size_t i = 0; i < w * h; i++) sum += (double)fb[i].x + fb[i].y + fb[i].z;
		printf("{\"width\": %zu, \"height\": %zu, \"workers\": %d, \"load_ms\": %.3f, \"render_ms\": %.3f, \"ssaa_ms\": %.3f, \"fbsum\": %.6f}\n",
			w, h, s.options.nWorkers, loadMs, renderMs, ssaaMs, sum);
		delete[] fb;
	}
	else if (mode == "bench") {
		if (argc < 4) return 2;
		const int repeat = atoi(argv[3]);
		setWorkers(s, argc > 4 ? atoi(argv[4]) : 0);
		Vec3f* fb = new Vec3f[w * h];
		printf("{\"width\": %zu, \"height\": %zu, \"workers\": %d, \"frame_ms\": [", w, h, s.options.nWorkers);
		for (int r = 0; r < repeat; r++) {
			for (size_t i = 0; i < w * h; i++) fb[i] = Vec3f();
			double a = nowMs();
			s.launchWorkers(fb);
			s.launchSSAA(fb);
			double ms = nowMs() - a;
			printf("%s%.3f", r ? ", " : "", ms);
			fflush(stdout);
		}
		printf("]}\n");
		delete[] fb;
	}
	else if (mode == "hits") {
		if (argc < 4) return 2;
		FILE* f = fopen(argv[3], "wb");
		if (!f) return 2;
		const float scale = tanf(s.camera.fov * 0.5f / 180.0f * (float)(M_PI));
		const float aspect = (s.options.width) / (float)s.options.height;
		for (size_t y = 0; y + 1 < h; y++) {
			for (size_t x = 0; x + 1 < w; x++) {
				float X = (float)x + 0.5f, Y = (float)y + 0.5f;
				float xPix = (2 * (X + 0.5f) / (float)w - 1) * scale * aspect;
				float yPix = -(2 * (Y + 0.5f) / (float)h - 1) * scale;
				Ray ray = s.camera.getRay(xPix, yPix);
				IntersectInfo info;
				bool hit = Render::trace(ray, s.objects, info);
				int obj = -1, tri = -1;
				if (hit) {
					for (size_t k = 0; k < s.objects.size(); k++) if (s.objects[k].get() == info.hitObject) obj = (int)k;
					if (info.hitObject->objectType == ObjectType::Mesh) {
						const Mesh* m = static_cast<const Mesh*>(info.hitObject);
						for (size_t k = 0; k < m->allTris.size(); k++) if (m->allTris[k] == info.triPtr) { tri = (int)k; break; }
					}
				}
				float rec[9] = { ray.dir.x, ray.dir.y, ray.dir.z, hit ? info.tNear : -1.0f,
					hit ? info.uv.x : 0.0f, hit ? info.uv.y : 0.0f, 0, 0, 0 };
				memcpy(&rec[6], &obj, 4);
				memcpy(&rec[7], &tri, 4);
				fwrite(rec, 4, 8, f);
			}
		}
		fclose(f);
	}
	else if (mode == "tree") {
		if (argc < 4) return 2;
		FILE* f = fopen(argv[3], "wb");
		if (!f) return 2;
		for (size_t k = 0; k < s.objects.size(); k++) {
			if (s.objects[k]->objectType != ObjectType::Mesh) continue;
			const Mesh* m = static_cast<const Mesh*>(s.objects[k].get());
			TreeStats st;
			walk(m->ac.get(), 1, st);
			printf("{\"object\": %zu, \"tris\": %zu, \"nodes\": %ld, \"leaves\": %ld, \"refs\": %ld, \"maxLeaf\": %ld, \"maxDepth\": %ld}\n",
				k, m->allTris.size(), st.nodes, st.leaves, st.refs, st.maxLeaf, st.maxDepth);
			int header[2] = { (int)m->allTris.size(), (int)st.nodes };
			fwrite(header, 4, 2, f);
			// triangles: 9 pos, 9 normal, 6 uv, 3 tangent, 3 bitangent = 30 floats
			for (const Triangle* t : m->allTris) {
				float r[30] = { t->a.x, t->a.y, t->a.z, t->b.x, t->b.y, t->b.z, t->c.x, t->c.y, t->c.z,
					t->n_a.x, t->n_a.y, t->n_a.z, t->n_b.x, t->n_b.y, t->n_b.z, t->n_c.x, t->n_c.y, t->n_c.z,
					t->t_a.x, t->t_a.y, t->t_b.x, t->t_b.y, t->t_c.x, t->t_c.y,
					t->tangent.x, t->tangent.y, t->tangent.z, t->bitangent.x, t->bitangent.y, t->bitangent.z };
				fwrite(r, 4, 30, f);
			}
			// pointer → index map via sorted lookup would need <algorithm>; meshes here are small
			// enough for the linear scan only when leaves are small, so build an index map first.
			std::vector<std::pair<const Triangle*, int>> sorted;
			sorted.reserve(m->allTris.size());
			for (size_t i = 0; i < m->allTris.size(); i++) sorted.push_back({ m->allTris[i], (int)i });
			std::sort(sorted.begin(), sorted.end());
			struct Rec { static void go(FILE* f, const AccelerationStructure* n,
				const std::vector<std::pair<const Triangle*, int>>& sorted) {
				float b[6] = { n->bounds[0].x, n->bounds[0].y, n->bounds[0].z, n->bounds[1].x, n->bounds[1].y, n->bounds[1].z };
				fwrite(b, 4, 6, f);
				int isLeaf = n->left ? 0 : 1;
				int count = isLeaf ? (int)n->tris.size() : 0;
				fwrite(&isLeaf, 4, 1, f);
				fwrite(&count, 4, 1, f);
				if (isLeaf) {
					for (const Triangle* t : n->tris) {
						auto it = std::lower_bound(sorted.begin(), sorted.end(), std::make_pair(t, -1));
						int idx = it->second;
						fwrite(&idx, 4, 1, f);
					}
				} else {
					go(f, n->left.get(), sorted);
					go(f, n->right.get(), sorted);
				}
			} };
			Rec::go(f, m->ac.get(), sorted);
		}
		fclose(f);
	}
	else if (mode == "cast") {
		if (argc < 5) return 2;
		FILE* f = fopen(argv[3], "rb");
		if (!f) return 2;
		fseek(f, 0, SEEK_END);
		long bytes = ftell(f);
		fseek(f, 0, SEEK_SET);
		size_t n = bytes / (6 * sizeof(float));
		std::vector<float> rays(n * 6);
		if (fread(rays.data(), 4, n * 6, f) != n * 6) return 2;
		fclose(f);
		std::vector<Vec3f> out(n);
		for (size_t i = 0; i < n; i++) {
			Ray r{ Vec3f(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), Vec3f(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]) };
			out[i] = Render::castRay(r, s, 0);
		}
		dumpFloats(argv[4], out.data(), n);
	}
	else if (mode == "stats") {
		setWorkers(s, argc > 3 ? atoi(argv[3]) : 0);
		Vec3f* fb = new Vec3f[w * h];
		s.launchWorkers(fb);
		s.launchSSAA(fb);
		// the reference's counters are 32-bit (include/stats.h:11-16) and wrap on large scenes
		printf("{\"rays\": %d, \"box_tests\": %d, \"tri_tests\": %d, \"tri_copies\": %zu, \"tris\": %zu}\n",
			(int)stats::raysCasted, (int)stats::accelStructTests, (int)stats::rayTriTests,
			(size_t)stats::triCopiesCount, (size_t)stats::meshCount);
		delete[] fb;
	}
	else if (mode == "ac") {
		if (argc < 4) return 2;
		// the showAC branch of Scene::render() keeps its counts private and only writes the 8-bit BMP; drive the
		// same public calls (Camera::getRay, Scene::countAC) over the same pixel grid and dump the integers
		const float scale = tanf(s.camera.fov * 0.5f / 180.0f * (float)(M_PI));
		const float aspect = (s.options.width) / (float)s.options.height;
		std::vector<int> counts(w * h);
		int acMax = 0;
		for (size_t y = 0; y < h; y++)
			for (size_t x = 0; x < w; x++) {
				float xPix = (2 * (x + 0.5f) / (float)w - 1) * scale * aspect;
				float yPix = -(2 * (y + 0.5f) / (float)h - 1) * scale;
				Ray ray = s.camera.getRay(xPix, yPix);
				int val = s.countAC(ray);
				if (val > acMax) acMax = val;
				counts[x + y * w] = val;
			}
		FILE* f = fopen(argv[3], "wb");
		if (!f) return 2;
		fwrite(counts.data(), 4, counts.size(), f);
		fclose(f);
		printf("{\"width\": %zu, \"height\": %zu, \"ac_max\": %d}\n", w, h, acMax);
	}
	else {
		fprintf(stderr, "unknown mode %s\n", mode.c_str());
		return 2;
	}
	return 0;
}
