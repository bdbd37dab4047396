// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// extern "C" harness around the UNMODIFIED reference path tracer.  The reference headers and
// .cpp files are compiled from /root/reference where they lie (oracle/Makefile); this file
// only #includes them.  It restates the one piece of the path that cannot be compiled
// headless -- the per-pixel loop body of MyApp::Tick (myapp.cpp:163-175) and the
// Accumulator store (myapp.h:17-19, 57-59), which live in files that need GLFW/OpenGL --
// and it reads protected members of reference objects (-fno-access-control) to export the
// built BVH, camera and material constants for comparison with the B200 host mirror.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load the library built from this file.
#include "precomp.h"
#include "disney.h"
#include "integrator.h"
#include "bvhtrimesh.h"
#include "texture.h"
#include "scene.h"

#define TINYOBJLOADER_IMPLEMENTATION
#include "tiny_obj_loader.h"

#include <atomic>
#include <thread>

#include "scenes/config_scenes.h"   // ag-pathtracer_b200/host/scenes (shared, API-only source)

extern "C" void agpt_ref_seed_path(unsigned pixel_index, unsigned sample);
extern "C" void agpt_ref_set_state(unsigned s);
extern "C" unsigned agpt_ref_get_state();

namespace {

struct RefScene {
	Scene scene;
	std::unique_ptr<Camera> camera;
};

struct Hit {          // same layout as agpt_hit in include/agpt.h
	uint32_t found;
	int32_t prim;
	int32_t tri;
	float t;
};

struct WalkStats {
	uint64_t interior = 0, boxes = 0, tris = 0;
};

// Instrumented mirror of BVHTriMesh::RecursiveHit (bvhtrimesh.h:332-384): same box tests
// (the reference's Bounds::Intersect), same near/far rule, same leaf loop through the
// reference's TriangleIntersect, plus the triangle id and visit counters the unmodified
// code does not expose.  Cross-checked against Scene::Intersect on every ray below.
bool MirrorHit(const BVHTriMesh& m, const BVHNode& node, const Ray& ray, SurfaceInteraction& hit, int& tri, WalkStats& st) {
	bool any = false;
	if (node.count > 0) {
		for (int i = 0; i < node.count; i++) {
			int idx = m.primitives[node.first + i].index;
			st.tris++;
			if (m.TriangleIntersect(ray, idx, hit)) { any = true; tri = idx / 3; }
		}
		return any;
	}
	st.interior++;
	st.boxes += 2;
	BVHNode a = m.nodes[node.first], b = m.nodes[node.first + 1];
	float da, db;
	bool ha = a.bounds.Intersect(ray, da), hb = b.bounds.Intersect(ray, db);
	bool swapKids;
	if (ha && hb) swapKids = db < da;
	else if (ha || hb) swapKids = !ha;
	else return false;
	if (swapKids) { BVHNode t = a; a = b; b = t; }
	if (MirrorHit(m, a, ray, hit, tri, st)) any = true;
	if (ha && hb && MirrorHit(m, b, ray, hit, tri, st)) any = true;
	return any;
}

// Scene::Intersect (scene.h:5-13) with ids: primitives in list order, each shrinking ray.t.
bool MirrorSceneIntersect(const Scene& scene, const Ray& ray, SurfaceInteraction& hit, int& prim, int& tri, WalkStats& st) {
	bool found = false;
	for (size_t p = 0; p < scene.primitives.size(); p++) {
		const Intersectable* shape = scene.primitives[p].get();
		if (auto* bvh = dynamic_cast<const BVHTriMesh*>(shape)) {
			float dist;
			st.boxes++;
			if (!bvh->nodes[0].bounds.Intersect(ray, dist)) continue;
			int t = -1;
			if (MirrorHit(*bvh, bvh->nodes[0], ray, hit, t, st)) { found = true; prim = (int)p; tri = t; }
		}
		else if (auto* mesh = dynamic_cast<const TriangleMesh*>(shape)) {
			for (size_t i = 0; i < mesh->indices.size(); i += 3) {
				st.tris++;
				if (mesh->TriangleIntersect(ray, (int)i, hit)) { found = true; prim = (int)p; tri = (int)(i / 3); }
			}
		}
		else if (shape->Intersect(ray, hit)) { found = true; prim = (int)p; tri = -1; }
	}
	return found;
}

// Ray counting without touching the reference: a do-nothing Intersectable put at the FRONT of
// Scene::primitives is called exactly once per Scene::Intersect (scene.h:7-11) and once per
// Scene::IntersectP (scene.h:16-18, before any early out); it never reports a hit, so results
// are unchanged (primitive indices shift by one, so it is only used for timing runs).
thread_local unsigned long long t_raysClosest = 0, t_raysAny = 0;
std::atomic<unsigned long long> g_raysClosest(0), g_raysAny(0);
class RayCounter : public Intersectable {
public:
	RayCounter() : Intersectable(nullptr) {}
	bool Intersect(const Ray&, SurfaceInteraction&) const override { t_raysClosest++; return false; }
	bool IntersectP(const Ray&) const override { t_raysAny++; return false; }
};
void FlushRayCounts() {
	g_raysClosest += t_raysClosest; g_raysAny += t_raysAny;
	t_raysClosest = t_raysAny = 0;
}

template <typename F>
void ParallelRows(int y0, int y1, int threads, F&& body) {
	if (threads <= 1) { for (int y = y0; y < y1; y++) body(y); FlushRayCounts(); return; }
	std::atomic<int> next(y0);
	std::vector<std::thread> pool;
	for (int t = 0; t < threads; t++)
		pool.emplace_back([&] { for (int y; (y = next.fetch_add(1)) < y1;) body(y); FlushRayCounts(); });
	for (auto& t : pool) t.join();
}

inline uint32_t Bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

} // namespace

extern "C" {

void* agpt_ref_scene_create(int config, int level) {
	auto* rs = new RefScene();
	std::streambuf* old = std::cerr.rdbuf(nullptr);   // silence the BVH build chatter (bvhtrimesh.h:166,177)
	bool ok = agpt_scenes::BuildConfig(&rs->scene, config, level);
	std::cerr.rdbuf(old);
	if (!ok) { delete rs; return nullptr; }
	rs->camera.reset(new Camera(rs->scene.camera));
	return rs;
}

void agpt_ref_scene_destroy(void* h) { delete (RefScene*)h; }

// Timing scenes only: prepend the ray counter (shifts primitive indices by one).
void agpt_ref_scene_count_rays(void* h) {
	auto* rs = (RefScene*)h;
	rs->scene.primitives.insert(rs->scene.primitives.begin(), std::make_shared<RayCounter>());
}
// out2 = {Scene::Intersect calls, Scene::IntersectP calls} since the last reset.
void agpt_ref_ray_counts(unsigned long long* out2, int reset) {
	out2[0] = g_raysClosest; out2[1] = g_raysAny;
	if (reset) { g_raysClosest = 0; g_raysAny = 0; }
}

int agpt_ref_scene_counts(void* h, int* n_prims, int* n_lights) {
	auto* rs = (RefScene*)h;
	*n_prims = (int)rs->scene.primitives.size();
	*n_lights = (int)rs->scene.lights.size();
	return 0;
}

// Render samples [s0, s0+ns) of pixels [x0,x1) x [y0,y1) of a W x H film into
// out_rgba[(H-1-y)*W + x] (float4 per pixel, .w untouched), i.e. the Accumulator layout.
// The buffer is accumulated into, not cleared.  Returns the number of paths traced.
long long agpt_ref_render(void* h, int W, int H, int x0, int y0, int x1, int y1, int s0, int ns,
		int max_depth, int depth_arg, int threads, float* out_rgba) {
	auto* rs = (RefScene*)h;
	const Scene& scene = rs->scene;
	const Camera& camera = *rs->camera;
	PathTracer integrator(max_depth);
	const int width = W, height = H;
	ParallelRows(y0, y1, threads, [&](int y) {
		for (int s = s0; s < s0 + ns; s++)
			for (int x = x0; x < x1; x++) {
				agpt_ref_seed_path((unsigned)(y * W + x), (unsigned)s);
				float2 p(x + RandomFloat(), y + RandomFloat());      // myapp.cpp:165
				float2 uv(p.x / width, p.y / height);               // myapp.h:57-59
				Ray ray = camera.GetRay(uv.x, uv.y);                // myapp.cpp:167
				float3 clr = integrator.Li(ray, scene, depth_arg);  // myapp.cpp:168
				if (HasNans(clr) || std::isinf(Luminance(clr))) clr = float3(0.f);   // myapp.cpp:169-172
				float* px = out_rgba + 4 * ((size_t)(height - 1 - y) * width + x);  // myapp.h:17-19
				px[0] += clr.x; px[1] += clr.y; px[2] += clr.z;
			}
	});
	return (long long)(x1 - x0) * (y1 - y0) * ns;
}

// Primary-ray hit table for one sample index: out[y*W + x] = {found, prim, tri, t}.
// rays_out (optional, 8 floats per pixel): O.xyz, D.xyz, 2 jitter values -- lets a test
// compare the generated camera rays bit for bit.  stats_out (optional): interior visits,
// box tests, triangle tests summed over all rays.  Returns the number of rays on which the
// instrumented walk disagreed with the unmodified Scene::Intersect (must be 0).
long long agpt_ref_primary_hits(void* h, int W, int H, int sample, int threads, void* out_hits, float* rays_out, unsigned long long* stats_out) {
	auto* rs = (RefScene*)h;
	const Scene& scene = rs->scene;
	const Camera& camera = *rs->camera;
	Hit* out = (Hit*)out_hits;
	std::atomic<long long> mismatches(0);
	std::atomic<unsigned long long> nInt(0), nBox(0), nTri(0);
	const int width = W, height = H;
	ParallelRows(0, H, threads, [&](int y) {
		WalkStats st;
		for (int x = 0; x < W; x++) {
			agpt_ref_seed_path((unsigned)(y * W + x), (unsigned)sample);
			float2 p(x + RandomFloat(), y + RandomFloat());
			float2 uv(p.x / width, p.y / height);
			Ray ray = camera.GetRay(uv.x, uv.y);
			if (rays_out) {
				float* r = rays_out + 8 * ((size_t)y * W + x);
				r[0] = ray.O.x; r[1] = ray.O.y; r[2] = ray.O.z;
				r[3] = ray.D.x; r[4] = ray.D.y; r[5] = ray.D.z;
				r[6] = p.x; r[7] = p.y;
			}
			Ray probe = ray;
			SurfaceInteraction a, b;
			bool fa = scene.Intersect(ray, a);
			int prim = -1, tri = -1;
			bool fb = MirrorSceneIntersect(scene, probe, b, prim, tri, st);
			if (fa != fb || (fa && (a.shape != b.shape || Bits(ray.t) != Bits(probe.t)))) mismatches++;
			Hit& o = out[(size_t)y * W + x];
			o.found = fa ? 1u : 0u;
			o.prim = fa ? prim : -1;
			o.tri = fa ? tri : -1;
			o.t = fa ? ray.t : 0.f;
		}
		nInt += st.interior; nBox += st.boxes; nTri += st.tris;
	});
	if (stats_out) { stats_out[0] = nInt; stats_out[1] = nBox; stats_out[2] = nTri; }
	return mismatches;
}

// Trace caller-supplied rays (O.xyz, D.xyz [normalised by the Ray ctor], tmax) through
// Scene::Intersect / IntersectP.  any_hit != 0 -> out[i].found only.
long long agpt_ref_trace_rays(void* h, long long n, const float* rays7, int any_hit, int threads, void* out_hits, unsigned long long* stats_out) {
	auto* rs = (RefScene*)h;
	const Scene& scene = rs->scene;
	Hit* out = (Hit*)out_hits;
	std::atomic<long long> mismatches(0);
	std::atomic<unsigned long long> nInt(0), nBox(0), nTri(0);
	const int chunk = 1024;
	int nchunks = (int)((n + chunk - 1) / chunk);
	ParallelRows(0, nchunks, threads, [&](int c) {
		WalkStats st;
		for (long long i = (long long)c * chunk; i < std::min(n, (long long)(c + 1) * chunk); i++) {
			const float* r = rays7 + 7 * i;
			Ray ray(float3(r[0], r[1], r[2]), float3(r[3], r[4], r[5]), r[6]);
			Hit& o = out[i];
			if (any_hit) {
				o.found = scene.IntersectP(ray) ? 1u : 0u; o.prim = -1; o.tri = -1; o.t = 0.f;
				continue;
			}
			Ray probe = ray;
			SurfaceInteraction a, b;
			bool fa = scene.Intersect(ray, a);
			int prim = -1, tri = -1;
			bool fb = MirrorSceneIntersect(scene, probe, b, prim, tri, st);
			if (fa != fb || (fa && (a.shape != b.shape || Bits(ray.t) != Bits(probe.t)))) mismatches++;
			o.found = fa ? 1u : 0u; o.prim = fa ? prim : -1; o.tri = fa ? tri : -1; o.t = fa ? ray.t : 0.f;
		}
		nInt += st.interior; nBox += st.boxes; nTri += st.tris;
	});
	if (stats_out) { stats_out[0] = nInt; stats_out[1] = nBox; stats_out[2] = nTri; }
	return mismatches;
}

// Radiance of single camera paths, no accumulation: out[3*i..] = Li for pixel (xs[i], ys[i]),
// sample ss[i].  draws_out (optional) = number of RandomFloat() calls the path consumed.
void agpt_ref_li_pixels(void* h, int W, int H, int n, const int* xs, const int* ys, const int* ss, int max_depth, int depth_arg, float* out, int* draws_out) {
	auto* rs = (RefScene*)h;
	PathTracer integrator(max_depth);
	const int width = W, height = H;
	for (int i = 0; i < n; i++) {
		int x = xs[i], y = ys[i];
		agpt_ref_seed_path((unsigned)(y * W + x), (unsigned)ss[i]);
		unsigned s0 = agpt_ref_get_state();
		float2 p(x + RandomFloat(), y + RandomFloat());
		float2 uv(p.x / width, p.y / height);
		Ray ray = rs->camera->GetRay(uv.x, uv.y);
		float3 clr = integrator.Li(ray, rs->scene, depth_arg);
		out[3 * i] = clr.x; out[3 * i + 1] = clr.y; out[3 * i + 2] = clr.z;
		if (draws_out) {
			// count draws by replaying the stream until it reaches the final state
			unsigned target = agpt_ref_get_state(), s = s0;
			int k = 0;
			while (s != target && k < 100000) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; k++; }
			draws_out[i] = k;
		}
	}
}

// Accumulator::CopyToSurface (myapp.h:34-41) on an accumulator-layout float4 buffer: the loop body
// restated (Surface needs GLFW), its arithmetic -- float3 / float, lin2rgb, rgb2uint -- is the
// reference's own (template/precomp.h:642, common.h:41-51), compiled from where it lies.
void agpt_ref_resolve(const float* rgba, long long n, int samples, unsigned* out_rgb8) {
	for (long long i = 0; i < n; i++) {
		float3 px(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2]);
		auto rgb = lin2rgb(px / (float)samples);
		out_rgb8[i] = rgb2uint(rgb);
	}
}

// ---- exports of built reference objects (for comparison with the host mirror) ----------

// Camera: origin, lower_left_corner, horizontal, vertical, u, v (3 floats each) + lens_radius.
void agpt_ref_camera_export(void* h, float* out19) {
	const Camera& c = *((RefScene*)h)->camera;
	const float3* v[6] = { &c.origin, &c.lower_left_corner, &c.horizontal, &c.vertical, &c.u, &c.v };
	for (int i = 0; i < 6; i++) { out19[3 * i] = v[i]->x; out19[3 * i + 1] = v[i]->y; out19[3 * i + 2] = v[i]->z; }
	out19[18] = c.lens_radius;
}

static int CountNodes(const BVHTriMesh& m) {
	// nodes[] has no stored length (bvhtrimesh.h:172); recover it as 1 + the largest index reachable.
	int maxIdx = 0;
	std::vector<int> stack{ 0 };
	while (!stack.empty()) {
		int i = stack.back(); stack.pop_back();
		maxIdx = std::max(maxIdx, i);
		if (m.nodes[i].count == 0) { stack.push_back(m.nodes[i].first); stack.push_back(m.nodes[i].first + 1); }
	}
	return maxIdx + 1;
}

// kind: 0 sphere, 1 plane, 2 BVHTriMesh, 3 TriangleMesh.  counts: nodes, triangles, vertices,
// normals, texcoords.
int agpt_ref_prim_info(void* h, int prim, int* kind, int* counts5, int* has_material, int* is_light) {
	const Scene& scene = ((RefScene*)h)->scene;
	const Intersectable* s = scene.primitives[prim].get();
	for (int i = 0; i < 5; i++) counts5[i] = 0;
	*has_material = s->GetMaterial() != nullptr;
	*is_light = s->GetAreaLight() != nullptr;
	if (dynamic_cast<const Sphere*>(s)) *kind = 0;
	else if (dynamic_cast<const Plane*>(s)) *kind = 1;
	else if (auto* tm = dynamic_cast<const TriangleMesh*>(s)) {
		auto* bvh = dynamic_cast<const BVHTriMesh*>(s);
		*kind = bvh ? 2 : 3;
		counts5[0] = bvh ? CountNodes(*bvh) : 0;
		counts5[1] = (int)tm->indices.size() / 3;
		counts5[2] = (int)tm->vertices.size();
		counts5[3] = (int)tm->normals.size();
		counts5[4] = (int)tm->texcoords.size();
	}
	else return -1;
	return 0;
}

// nodes_out: 8 x 4 bytes per node, verbatim BVHNode.  leaf_tri_out: for every leaf slot i of
// primitives[], the original triangle number primitives[i].index / 3.
int agpt_ref_bvh_export(void* h, int prim, void* nodes_out, int* leaf_tri_out) {
	const Scene& scene = ((RefScene*)h)->scene;
	auto* bvh = dynamic_cast<const BVHTriMesh*>(scene.primitives[prim].get());
	if (!bvh) return -1;
	static_assert(sizeof(BVHNode) == 32, "BVHNode layout");
	int n = CountNodes(*bvh);
	memcpy(nodes_out, bvh->nodes, (size_t)n * sizeof(BVHNode));
	memset((char*)nodes_out + sizeof(BVHNode), 0, sizeof(BVHNode));   // slot 1 is never written upstream
	for (size_t i = 0; i < bvh->primitives.size(); i++) leaf_tri_out[i] = bvh->primitives[i].index / 3;
	return n;
}

// Mesh geometry in original (index) order: tri_verts_out = 9 floats per triangle.
int agpt_ref_mesh_export(void* h, int prim, float* tri_verts_out) {
	const Scene& scene = ((RefScene*)h)->scene;
	auto* tm = dynamic_cast<const TriangleMesh*>(scene.primitives[prim].get());
	if (!tm) return -1;
	for (size_t i = 0; i < tm->indices.size(); i++) {
		const float3& v = tm->vertices[tm->indices[i].vertex_index];
		tri_verts_out[3 * i] = v.x; tri_verts_out[3 * i + 1] = v.y; tri_verts_out[3 * i + 2] = v.z;
	}
	return (int)tm->indices.size() / 3;
}

// Material constants as the reference constructors derived them (material.h:14-49):
// out = { type(0 none,1 disney,2 mirror), diffuse R xyz, retro R xyz, retro roughness,
//         alphax, alphay, fresnel R0 xyz, metallic, eta, mirror R xyz, has_diffuse, has_retro }
int agpt_ref_material_export(void* h, int prim, float* out20) {
	const Scene& scene = ((RefScene*)h)->scene;
	const Material* m = scene.primitives[prim]->GetMaterial();
	for (int i = 0; i < 20; i++) out20[i] = 0;
	if (!m) return 0;
	if (auto* d = dynamic_cast<const DisneyMaterial*>(m)) {
		out20[0] = 1;
		if (d->diffuse) { out20[1] = d->diffuse->R.x; out20[2] = d->diffuse->R.y; out20[3] = d->diffuse->R.z; out20[18] = 1; }
		if (d->retro) { out20[4] = d->retro->R.x; out20[5] = d->retro->R.y; out20[6] = d->retro->R.z; out20[7] = d->retro->roughness; out20[19] = 1; }
		auto* dist = static_cast<const TrowbridgeReitzDistribution*>(d->microfacet->distribution);
		out20[8] = dist->alphax; out20[9] = dist->alphay;
		auto* fr = static_cast<const DisneyFresnel*>(d->microfacet->fresnel);
		out20[10] = fr->R0.x; out20[11] = fr->R0.y; out20[12] = fr->R0.z; out20[13] = fr->metallic; out20[14] = fr->eta;
		return 1;
	}
	if (auto* mm = dynamic_cast<const MirrorMaterial*>(m)) {
		out20[0] = 2;
		out20[15] = mm->reflection->R.x; out20[16] = mm->reflection->R.y; out20[17] = mm->reflection->R.z;
		return 2;
	}
	return -1;
}

// ---- per-function probes (differential tests against single reference functions) -------

// Bounds::Intersect (bvhtrimesh.h:18-36).  boxes: 6 floats (bmin, bmax); rays: O, D (used as
// given, NOT normalised), tmax = 7 floats.  out_hit[i] in {0,1}, out_t[i] = entry distance.
void agpt_ref_probe_bounds(int n, const float* boxes6, const float* rays7, int* out_hit, float* out_t) {
	for (int i = 0; i < n; i++) {
		Bounds b;
		for (int a = 0; a < 3; a++) { b.bmin3[a] = boxes6[6 * i + a]; b.bmax3[a] = boxes6[6 * i + 3 + a]; }
		Ray ray;
		const float* r = rays7 + 7 * i;
		ray.O = float3(r[0], r[1], r[2]); ray.D = float3(r[3], r[4], r[5]); ray.t = r[6];
		float t = 0;
		out_hit[i] = b.Intersect(ray, t) ? 1 : 0;
		out_t[i] = out_hit[i] ? t : 0.f;
	}
}

// The reflection half of the rough-dielectric extension, built from the reference's OWN classes: MicrofacetReflection
// (reflection.h:38-78) over the plain TrowbridgeReitzDistribution (microfacet.h:112-153) and FresnelDielectric(1, eta)
// (microfacet.h:220-228).  The reference never combines them (its only Fresnel in use is DisneyFresnel), but it can:
// this pins AGPT_LOBE_GLASS_REFLECT to reference code.  alpha = max(.001, roughness^2) as host/material.h GlassMaterial.
class RefGlassReflection : public Material {
public:
	RefGlassReflection(const float3& Kr, float roughness, float eta) {
		float a = std::max(.001f, roughness * roughness);
		lobe = std::make_shared<MicrofacetReflection>(Kr, new TrowbridgeReitzDistribution(a, a), new FresnelDielectric(1.f, eta));
	}
	void SetupBSDF(BSDF* bsdf) const override { bsdf->AddBxDF(lobe.get()); }
private:
	std::shared_ptr<MicrofacetReflection> lobe;
};

// BSDF probe on a flat shading frame.  For each i: builds a SurfaceInteraction from
// (p=0, dpdu, dpdv), optionally SetShadingGeometry(ss, ts), sets up the material's BSDF and
// evaluates f(wo,wi,skipSpecular), Pdf, and Sample_f(wo,u).
// mat: {type, color xyz, roughness, metallic} (type 1 disney, 2 mirror).
// in: dpdu3 dpdv3 wo3 wi3 u2 = 14 floats.  out: f3 pdf1 | wi3 f3 pdf1 specular1 = 12 floats.
void agpt_ref_probe_bsdf(int n, const float* mat6, const float* in14, int skip_specular, float* out12) {
	std::shared_ptr<Material> m;
	if ((int)mat6[0] == 1) m = DisneyMaterial::Make(float3(mat6[1], mat6[2], mat6[3]), mat6[4], mat6[5]);
	else if ((int)mat6[0] == 3) m = std::make_shared<RefGlassReflection>(float3(mat6[1], mat6[2], mat6[3]), mat6[4], mat6[5]);   // {3, Kr xyz, roughness, eta}
	else m = MirrorMaterial::Make(float3(mat6[1], mat6[2], mat6[3]));
	Sphere shape(float3(0.f), 1.f, m);
	for (int i = 0; i < n; i++) {
		const float* a = in14 + 14 * i;
		float3 dpdu(a[0], a[1], a[2]), dpdv(a[3], a[4], a[5]), wo(a[6], a[7], a[8]), wi(a[9], a[10], a[11]);
		float2 u(a[12], a[13]);
		SurfaceInteraction si(float3(0.f), float2(0, 0), wo, dpdu, dpdv, &shape);
		si.EvalMaterial();
		float* o = out12 + 12 * i;
		float3 f = si.bsdf.f(wo, wi, skip_specular != 0);
		o[0] = f.x; o[1] = f.y; o[2] = f.z;
		o[3] = si.bsdf.Pdf(wo, wi, skip_specular != 0);
		float3 wis(0.f);
		float pdf = 0;
		bool spec = false;
		float3 fs = si.bsdf.Sample_f(wo, &wis, u, &pdf, skip_specular != 0, &spec);
		o[4] = wis.x; o[5] = wis.y; o[6] = wis.z; o[7] = fs.x; o[8] = fs.y; o[9] = fs.z; o[10] = pdf; o[11] = spec ? 1.f : 0.f;
	}
}

// Sphere::Sample(ref,u) / Sphere::Pdf (intersectable.h:239-317).  in: center3 r1 refp3 u2 = 9;
// out: p3 n3 pdf1 pdfOnly1 = 8.
void agpt_ref_probe_sphere_sample(int n, const float* in9, float* out8) {
	for (int i = 0; i < n; i++) {
		const float* a = in9 + 9 * i;
		Sphere s(float3(a[0], a[1], a[2]), a[3], nullptr);
		SurfaceInteraction ref;
		ref.p = float3(a[4], a[5], a[6]);
		float pdf = 0;
		Interaction it = s.Sample(ref, float2(a[7], a[8]), &pdf);
		float* o = out8 + 8 * i;
		o[0] = it.p.x; o[1] = it.p.y; o[2] = it.p.z; o[3] = it.n.x; o[4] = it.n.y; o[5] = it.n.z; o[6] = pdf;
		o[7] = s.Pdf(ref, float3(0, 0, 1));
	}
}

// First k floats of the stream of (pixel_index, sample) -- pins the stream definition.
void agpt_ref_probe_stream(unsigned pixel_index, unsigned sample, int k, float* out) {
	agpt_ref_seed_path(pixel_index, sample);
	for (int i = 0; i < k; i++) out[i] = RandomFloat();
}

// Order in which g++ evaluates the two RandomFloat() calls of `float2 u(RandomFloat(),
// RandomFloat())` (integrator.h:102-103,171; myapp.cpp:165): out = {u.x, u.y, first draw, second draw}.
void agpt_ref_probe_draw_order(unsigned state, float* out4) {
	agpt_ref_set_state(state);
	float2 u(RandomFloat(), RandomFloat());
	agpt_ref_set_state(state);
	float a = RandomFloat();
	float b = RandomFloat();
	out4[0] = u.x; out4[1] = u.y; out4[2] = a; out4[3] = b;
}

int agpt_ref_sizes(int* out) {
	out[0] = sizeof(float3); out[1] = sizeof(BVHNode); out[2] = sizeof(Ray); out[3] = sizeof(Primitive);
	out[4] = sizeof(index_type); out[5] = sizeof(BSDF); out[6] = sizeof(SurfaceInteraction);
	return 7;
}

} // extern "C"
