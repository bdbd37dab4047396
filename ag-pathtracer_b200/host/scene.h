// host/scene.h -- Scene with the reference's surface (/root/reference/scene.h:3-30): a flat
// list of primitives in insertion order, a list of lights, a CameraDesc -- plus Flatten(),
// which turns the object graph into the typed tables of include/agpt.h.
//
// Scene::Intersect / IntersectP (scene.h:5-19) are the device's closest-hit and any-hit
// kernels; list ORDER is preserved in the primitive table because it decides exact-t ties
// (a sphere accepts root == ray.t, triangles and planes reject t >= ray.t).
#pragma once

#include <atomic>
#include <map>

#include "precomp.h"
#include "camera.h"
#include "intersectable.h"
#include "trianglemesh.h"
#include "bvhtrimesh.h"
#include "lights.h"

// Everything agpt_upload_* needs, owned in one place so the pointers in `meshes` stay valid.
struct FlatScene {
	std::vector<agpt_prim> prims;
	std::vector<agpt_sphere> spheres;
	std::vector<agpt_plane> planes;
	std::vector<agpt_material> materials;
	std::vector<agpt_light> lights;
	std::vector<agpt_mesh_desc> meshes;
	std::vector<FlatTriangles> meshTris;
	agpt_envmap envmap = { 0, 0, nullptr, nullptr, nullptr, 0.f };   // borrowed from the scene's InfiniteAreaLight
	uint64_t Bytes() const {
		uint64_t b = prims.size() * sizeof(agpt_prim) + spheres.size() * sizeof(agpt_sphere) + planes.size() * sizeof(agpt_plane)
			+ materials.size() * sizeof(agpt_material) + lights.size() * sizeof(agpt_light);
		for (auto& m : meshes) b += (uint64_t)m.n_nodes * 32 + (uint64_t)m.n_tris * (48 + 4 + (m.tri_normals ? 48 : 0) + (m.tri_uvs ? 24 : 0));
		return b;
	}
};

class Scene {
public:
	void addAreaLight(std::shared_ptr<Intersectable> shape, const float3& L) {
		primitives.push_back(shape);
		lights.push_back(make_shared<AreaLight>(shape, L));
	}

	// Object graph -> device tables.  Materials are de-duplicated by pointer.
	std::shared_ptr<FlatScene> Flatten() const {
		auto flat = std::make_shared<FlatScene>();
		std::map<const Material*, int> matIndex;
		std::map<const Light*, int> lightIndex;
		std::map<const Intersectable*, int> primIndex;
		for (size_t i = 0; i < lights.size(); i++) lightIndex[lights[i].get()] = (int)i;
		flat->meshTris.reserve(primitives.size());
		for (size_t i = 0; i < primitives.size(); i++) {
			const Intersectable* shape = primitives[i].get();
			primIndex[shape] = (int)i;
			agpt_prim row;
			row.type = shape->Kind();
			row.material = -1;
			if (const Material* m = shape->GetMaterial()) {
				auto it = matIndex.find(m);
				if (it == matIndex.end()) {
					it = matIndex.emplace(m, (int)flat->materials.size()).first;
					flat->materials.push_back(m->Export());
				}
				row.material = it->second;
			}
			row.area_light = -1;
			if (const AreaLight* al = shape->GetAreaLight()) {
				auto it = lightIndex.find(al);
				if (it != lightIndex.end()) row.area_light = it->second;
			}
			switch (row.type) {
			case AGPT_PRIM_SPHERE:
				row.payload = (int)flat->spheres.size();
				flat->spheres.push_back(static_cast<const Sphere*>(shape)->Export());
				break;
			case AGPT_PRIM_PLANE:
				row.payload = (int)flat->planes.size();
				flat->planes.push_back(static_cast<const Plane*>(shape)->Export());
				break;
			default: {
				const TriangleMesh* mesh = static_cast<const TriangleMesh*>(shape);
				agpt_mesh_desc d;
				memset(&d, 0, sizeof(d));
				std::vector<int32_t> order;
				if (row.type == AGPT_PRIM_BVH_MESH) {
					const BVHTriMesh* bvh = static_cast<const BVHTriMesh*>(shape);
					d.nodes = reinterpret_cast<const agpt_bvh_node*>(bvh->Nodes().data());   // owned by the mesh
					d.n_nodes = (int)bvh->Nodes().size();
					order = bvh->LeafOrder();
				}
				else {
					order.resize(mesh->NumTriangles());
					for (size_t t = 0; t < order.size(); t++) order[t] = (int32_t)t;
				}
				flat->meshTris.push_back(mesh->ExportTriangles(order));
				const FlatTriangles& ft = flat->meshTris.back();
				d.n_tris = (int)ft.ids.size();
				d.tri_verts = ft.verts.data();
				d.tri_ids = ft.ids.data();
				d.tri_normals = ft.normals.empty() ? nullptr : ft.normals.data();
				d.tri_uvs = ft.uvs.empty() ? nullptr : ft.uvs.data();
				row.payload = (int)flat->meshes.size();
				flat->meshes.push_back(d);
			} break;
			}
			flat->prims.push_back(row);
		}
		for (auto& l : lights) {
			agpt_light rec;
			memset(&rec, 0, sizeof(rec));
			rec.type = l->Kind();
			rec.prim = -1;
			if (rec.type == AGPT_LIGHT_AREA) {
				auto it = primIndex.find(static_cast<const AreaLight*>(l.get())->Shape.get());
				if (it != primIndex.end()) rec.prim = it->second;
			}
			if (rec.type == AGPT_LIGHT_INFINITE_AREA) flat->envmap = static_cast<const InfiniteAreaLight*>(l.get())->Export();
			float3 e = l->Emission();
			rec.lemit[0] = e.x; rec.lemit[1] = e.y; rec.lemit[2] = e.z;
			flat->lights.push_back(rec);
		}
		return flat;
	}

	// What an uploaded copy of this scene is keyed on (CudaPathTracer): the object's serial number --
	// a different Scene later constructed at the same address gets a new one -- mixed with the
	// identities of its primitives and lights, so push_back / addAreaLight after an upload are seen.
	// Shapes, materials and lights are immutable after construction upstream, so identity is content.
	uint64_t Fingerprint() const {
		uint64_t h = 0xcbf29ce484222325ull ^ serial;
		auto mix = [&h](uint64_t v) { h = (h ^ v) * 0x100000001b3ull; };
		mix(primitives.size()); mix(lights.size());
		for (auto& p : primitives) mix((uint64_t)(uintptr_t)p.get());
		for (auto& l : lights) mix((uint64_t)(uintptr_t)l.get());
		return h;
	}

	vector<shared_ptr<Intersectable>> primitives;
	vector<shared_ptr<Light>> lights;
	CameraDesc camera;

private:
	static uint64_t NextSerial() { static std::atomic<uint64_t> next(1); return next.fetch_add(1); }
	uint64_t serial = NextSerial();
};
