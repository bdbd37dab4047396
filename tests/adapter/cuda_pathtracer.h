// tests/adapter/cuda_pathtracer.h -- the drop-in as a maintainer of ag-pathtracer would add it: ONE new header
// next to integrator.h, written against the REFERENCE's own classes (Scene, Intersectable, Sphere, Plane,
// TriangleMesh, BVHTriMesh, DisneyMaterial, MirrorMaterial, AreaLight, UniformInfiniteLight, InfiniteAreaLight,
// Camera -- /root/reference/*.h, unmodified), talking to the GPU only through the C ABI of include/agpt.h.
//
// This is INTEGRATION.md section A as code that compiles: oracle/Makefile builds it against the headers where
// they lie under /root/reference (target `adapter`), and tests/test_gpu_adapter.py renders reference Scene objects
// through it on the GPU and compares with the reference's own PathTracer::Li on the CPU.
//
// The reference keeps its scene data in protected / private members (bvhtrimesh.h:200-210, trianglemesh.h:43-50,
// camera.h:92-106, material.h:64-69, lights.h:85-86, intersectable.h:58-60): the translation unit that includes this
// header is compiled with g++ -fno-access-control (a maintainer would add `friend class CudaPathTracer;` instead).
// Nothing else of the reference changes.
#pragma once

#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "integrator.h"     // the reference's: Integrator, Scene, lights, materials (all #pragma once)
#include "bvhtrimesh.h"     // (holds non-inline definitions: exactly one translation unit may include it, bvhtrimesh.h:213-413)
#include "texture.h"
#include "sampling.h"
#include "agpt.h"

class CudaPathTracer : public Integrator {        // Integrator: integrator.h:26-31
public:
	explicit CudaPathTracer(int maxDepth = 5, int device = 0) : MaxDepth(maxDepth) {
		if (agpt_create(device, &ctx) != AGPT_OK) throw std::runtime_error(std::string("CudaPathTracer: ") + agpt_last_error());
	}
	~CudaPathTracer() { agpt_destroy(ctx); }
	CudaPathTracer(const CudaPathTracer&) = delete;
	CudaPathTracer& operator=(const CudaPathTracer&) = delete;

	// Once per scene (a Scene is immutable after MyApp::Init, myapp.cpp:121-135): the object graph -> the tables of agpt.h.
	void Upload(const Scene& scene) {
		std::vector<agpt_prim> prims;
		std::vector<agpt_sphere> spheres;
		std::vector<agpt_plane> planes;
		std::vector<agpt_mesh_desc> meshes;
		std::vector<agpt_material> materials;
		std::vector<agpt_light> lights;
		std::vector<std::vector<float>> verts, normals, uvs;       // per mesh, leaf order (kept alive until the uploads return)
		std::vector<std::vector<int32_t>> ids;
		std::map<const Material*, int> materialIndex;
		std::map<const Light*, int> lightIndex;
		std::map<const Intersectable*, int> primIndex;
		for (size_t i = 0; i < scene.lights.size(); i++) lightIndex[scene.lights[i].get()] = (int)i;
		verts.reserve(scene.primitives.size()); normals.reserve(scene.primitives.size()); uvs.reserve(scene.primitives.size()); ids.reserve(scene.primitives.size());

		for (size_t p = 0; p < scene.primitives.size(); p++) {      // list order is kept: it decides exact-t ties (scene.h:5-13)
			const Intersectable* shape = scene.primitives[p].get();
			primIndex[shape] = (int)p;
			agpt_prim row;
			row.material = -1;
			if (const Material* m = shape->GetMaterial()) {
				auto it = materialIndex.find(m);
				if (it == materialIndex.end()) {
					it = materialIndex.emplace(m, (int)materials.size()).first;
					materials.push_back(ExportMaterial(m));
				}
				row.material = it->second;
			}
			row.area_light = -1;
			if (const AreaLight* al = shape->GetAreaLight()) {
				auto it = lightIndex.find(al);
				if (it != lightIndex.end()) row.area_light = it->second;
			}
			if (auto* s = dynamic_cast<const Sphere*>(shape)) {
				row.type = AGPT_PRIM_SPHERE; row.payload = (int)spheres.size();
				agpt_sphere r;
				memset(&r, 0, sizeof(r));
				r.center[0] = s->Center.x; r.center[1] = s->Center.y; r.center[2] = s->Center.z; r.r = s->r; r.r2 = s->r2;
				spheres.push_back(r);
			}
			else if (auto* pl = dynamic_cast<const Plane*>(shape)) {
				row.type = AGPT_PRIM_PLANE; row.payload = (int)planes.size();
				agpt_plane r;
				memset(&r, 0, sizeof(r));
				r.o[0] = pl->O.x; r.o[1] = pl->O.y; r.o[2] = pl->O.z; r.half_x = pl->HalfSize.x; r.half_z = pl->HalfSize.y;
				planes.push_back(r);
			}
			else if (auto* mesh = dynamic_cast<const TriangleMesh*>(shape)) {
				auto* bvh = dynamic_cast<const BVHTriMesh*>(shape);
				row.type = bvh ? AGPT_PRIM_BVH_MESH : AGPT_PRIM_MESH; row.payload = (int)meshes.size();
				// leaf order: slot j holds the triangle primitives[j] names (bvhtrimesh.h:339); a plain mesh keeps index order
				std::vector<int> first;         // index of the triangle's first vertex in mesh->indices
				if (bvh) for (auto& pr : bvh->primitives) first.push_back(pr.index);
				else for (size_t i = 0; i < mesh->indices.size(); i += 3) first.push_back((int)i);
				verts.emplace_back(); normals.emplace_back(); uvs.emplace_back(); ids.emplace_back();
				auto& v = verts.back(); auto& n = normals.back(); auto& t = uvs.back(); auto& id = ids.back();
				for (int f : first) {
					id.push_back(f / 3);
					for (int k = 0; k < 3; k++) {
						const index_type& ix = mesh->indices[f + k];
						const float3& a = mesh->vertices[ix.vertex_index];
						v.insert(v.end(), { a.x, a.y, a.z, 0.f });
						if (!mesh->normals.empty()) { const float3& b = mesh->normals[ix.normal_index]; n.insert(n.end(), { b.x, b.y, b.z, 0.f }); }
						if (!mesh->texcoords.empty()) { const float2& c = mesh->texcoords[ix.texcoord_index]; t.insert(t.end(), { c.x, c.y }); }
					}
				}
				agpt_mesh_desc d;
				memset(&d, 0, sizeof(d));
				static_assert(sizeof(BVHNode) == sizeof(agpt_bvh_node), "BVHNode is uploaded verbatim");
				if (bvh) { d.nodes = reinterpret_cast<const agpt_bvh_node*>(bvh->nodes); d.n_nodes = CountNodes(*bvh); }
				d.n_tris = (int)id.size();
				d.tri_verts = v.data(); d.tri_ids = id.data();
				d.tri_normals = n.empty() ? nullptr : n.data();
				d.tri_uvs = t.empty() ? nullptr : t.data();
				meshes.push_back(d);
			}
			else throw std::runtime_error("CudaPathTracer: unknown Intersectable");
			prims.push_back(row);
		}
		agpt_envmap env = { 0, 0, nullptr, nullptr, nullptr, 0.f };
		std::vector<float> envRgb;
		for (auto& l : scene.lights) {
			agpt_light rec;
			memset(&rec, 0, sizeof(rec));
			rec.prim = -1;
			if (auto* al = dynamic_cast<const AreaLight*>(l.get())) {
				rec.type = AGPT_LIGHT_AREA;
				auto it = primIndex.find(al->Shape.get());
				if (it != primIndex.end()) rec.prim = it->second;
				rec.lemit[0] = al->Lemit.x; rec.lemit[1] = al->Lemit.y; rec.lemit[2] = al->Lemit.z;
			}
			else if (auto* ul = dynamic_cast<const UniformInfiniteLight*>(l.get())) {
				rec.type = AGPT_LIGHT_UNIFORM_INFINITE;
				rec.lemit[0] = ul->Lemit.x; rec.lemit[1] = ul->Lemit.y; rec.lemit[2] = ul->Lemit.z;
			}
			else if (auto* il = dynamic_cast<const InfiniteAreaLight*>(l.get())) {
				rec.type = AGPT_LIGHT_INFINITE_AREA;
				// texels and the distribution the constructor built (lights.cpp:31-48, sampling.h:20-36)
				env.width = il->Lmap->Width(); env.height = il->Lmap->Height();
				for (int i = 0; i < env.width * env.height; i++) { const float3& px = il->Lmap->pixels[i]; envRgb.insert(envRgb.end(), { px.x, px.y, px.z }); }
				env.rgb = envRgb.data(); env.func = il->distrib->func.data(); env.cdf = il->distrib->cdf.data(); env.func_int = il->distrib->funcInt;
			}
			else throw std::runtime_error("CudaPathTracer: unknown Light");
			lights.push_back(rec);
		}
		Check(agpt_upload_meshes(ctx, meshes.data(), (int)meshes.size()));
		Check(agpt_upload_spheres(ctx, spheres.data(), (int)spheres.size()));
		Check(agpt_upload_planes(ctx, planes.data(), (int)planes.size()));
		Check(agpt_upload_materials(ctx, materials.data(), (int)materials.size()));
		Check(agpt_upload_lights(ctx, lights.data(), (int)lights.size()));
		Check(agpt_upload_envmap(ctx, env.width > 0 ? &env : nullptr));
		Check(agpt_upload_primitives(ctx, prims.data(), (int)prims.size()));
		uploaded = &scene;
	}

	// The batched Tick body: replaces the for-y / for-x loop of MyApp::Tick (myapp.cpp:163-175) for samples
	// firstSample .. firstSample+numSamples-1.  `pixels` is Accumulator::pixels (myapp.h:8-13: float3 = 16 bytes,
	// row height-1-y), added to like successive Ticks add to it; the caller bumps Accumulator::samples.
	void Render(const Scene& scene, const Camera& camera, float3* pixels, int width, int height, int firstSample, int numSamples, int depth = 0) {
		if (uploaded != &scene) Upload(scene);
		agpt_camera c;
		const float3* src[6] = { &camera.origin, &camera.lower_left_corner, &camera.horizontal, &camera.vertical, &camera.u, &camera.v };
		float* dst[6] = { c.origin, c.lower_left_corner, c.horizontal, c.vertical, c.u, c.v };
		for (int i = 0; i < 6; i++) { dst[i][0] = src[i]->x; dst[i][1] = src[i]->y; dst[i][2] = src[i]->z; }
		c.lens_radius = camera.lens_radius;
		Check(agpt_set_camera(ctx, &c));
		Check(agpt_set_film(ctx, width, height));
		static_assert(sizeof(float3) == 16, "Accumulator::pixels is a float4-strided buffer");
		Check(agpt_write_accum(ctx, &pixels->x));
		Check(agpt_render(ctx, firstSample, numSamples, 1, MaxDepth, depth, 0));
		Check(agpt_read_accum(ctx, &pixels->x));
	}

	// Accumulator::CopyToSurface (myapp.h:34-41) of the film resident on the device: out[y*width + x] = 0x00RRGGBB
	void CopyToSurface(int samples, uint* out) { Check(agpt_resolve(ctx, samples, out)); }

	// Per-ray entry point (debug click, myapp.cpp:196-198).  The reference draws from its global generator
	// (RandomUInt, template.cpp:667-676); the path starts from one draw of it.
	float3 Li(const Ray& ray, const Scene& scene, int depth = 0) const override {
		if (uploaded != &scene) const_cast<CudaPathTracer*>(this)->Upload(scene);
		float r7[7] = { ray.O.x, ray.O.y, ray.O.z, ray.D.x, ray.D.y, ray.D.z, ray.t };
		uint32_t state = RandomUInt();
		float out[3] = { 0, 0, 0 };
		Check(agpt_li_rays(ctx, 1, r7, &state, MaxDepth, depth, AGPT_FLAG_RAYS_FINAL, out));
		return float3(out[0], out[1], out[2]);
	}

	agpt_ctx* Context() const { return ctx; }

private:
	static void Check(int status) { if (status != AGPT_OK) throw std::runtime_error(std::string("agpt: ") + agpt_last_error()); }

	// nodes[] has no stored length (bvhtrimesh.h:172): 1 + the largest index reachable from the root
	static int CountNodes(const BVHTriMesh& m) {
		int maxIdx = 0;
		std::vector<int> stack{ 0 };
		while (!stack.empty()) {
			int i = stack.back(); stack.pop_back();
			if (i > maxIdx) maxIdx = i;
			if (m.nodes[i].count == 0) { stack.push_back(m.nodes[i].first); stack.push_back(m.nodes[i].first + 1); }
		}
		return maxIdx + 1;
	}

	// the constants the material constructors derived (material.h:14-49,74-77), read back from the BxDF objects
	static agpt_material ExportMaterial(const Material* m) {
		agpt_material r;
		memset(&r, 0, sizeof(r));
		if (auto* d = dynamic_cast<const DisneyMaterial*>(m)) {
			r.type = AGPT_MAT_DISNEY;
			r.eta = d->eta;
			if (d->diffuse) { r.lobes |= AGPT_LOBE_DIFFUSE; r.diffuse_r[0] = d->diffuse->R.x; r.diffuse_r[1] = d->diffuse->R.y; r.diffuse_r[2] = d->diffuse->R.z; }
			if (d->retro) { r.lobes |= AGPT_LOBE_RETRO; r.roughness = d->retro->roughness; r.diffuse_r[0] = d->retro->R.x; r.diffuse_r[1] = d->retro->R.y; r.diffuse_r[2] = d->retro->R.z; }
			if (d->microfacet) {
				r.lobes |= AGPT_LOBE_MICROFACET;
				auto* dist = static_cast<const TrowbridgeReitzDistribution*>(d->microfacet->distribution);
				auto* fr = static_cast<const DisneyFresnel*>(d->microfacet->fresnel);
				r.alpha_x = dist->alphax; r.alpha_y = dist->alphay;
				r.spec_r0[0] = fr->R0.x; r.spec_r0[1] = fr->R0.y; r.spec_r0[2] = fr->R0.z; r.metallic = fr->metallic; r.eta = fr->eta;
			}
		}
		else if (auto* mm = dynamic_cast<const MirrorMaterial*>(m)) {
			r.type = AGPT_MAT_MIRROR; r.lobes = AGPT_LOBE_SPECULAR;
			r.mirror_r[0] = mm->reflection->R.x; r.mirror_r[1] = mm->reflection->R.y; r.mirror_r[2] = mm->reflection->R.z;
		}
		else throw std::runtime_error("CudaPathTracer: unknown Material");
		return r;
	}

	int MaxDepth;
	agpt_ctx* ctx = nullptr;
	const Scene* uploaded = nullptr;
};
