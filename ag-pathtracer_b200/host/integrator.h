// host/integrator.h -- Integrator (reference surface, /root/reference/integrator.h:28-31)
// and CudaPathTracer, the B200 implementation of PathTracer (integrator.h:120-196).
//
// CudaPathTracer::Render is the batched form of MyApp::Tick's pixel loop
// (myapp.cpp:163-175): numSamples passes over every pixel of the accumulator, executed as
// wavefront kernels on the device through the C ABI (include/agpt.h).  Li() keeps the
// per-ray entry point (used upstream by the debug click, myapp.cpp:196-198) by tracing a
// one-path batch on the device.  No CPU integrator exists here: if libagpt cannot create a
// CUDA context the constructor throws.
#pragma once

#include <stdexcept>

#include "precomp.h"
#include "scene.h"
#include "accumulator.h"

class Integrator {
public:
	virtual ~Integrator() {}
	virtual float3 Li(const Ray& ray, const Scene& scene, int depth = 0) const = 0;
};

class CudaPathTracer : public Integrator {
public:
	explicit CudaPathTracer(int maxDepth = 5, int device = 0) : MaxDepth(maxDepth) {
		if (agpt_create(device, &ctx) != AGPT_OK) throw std::runtime_error(std::string("CudaPathTracer: ") + agpt_last_error());
	}
	~CudaPathTracer() override { if (ctx) agpt_destroy(ctx); }
	CudaPathTracer(const CudaPathTracer&) = delete;
	CudaPathTracer& operator=(const CudaPathTracer&) = delete;

	// Upload (or re-upload) the flattened scene.  Called implicitly by Render/Li whenever the
	// scene's fingerprint differs from the resident one (another Scene object, or primitives /
	// lights added since).
	void Upload(const Scene& scene) const {
		auto flat = scene.Flatten();
		Check(agpt_upload_meshes(ctx, flat->meshes.data(), (int)flat->meshes.size()));
		Check(agpt_upload_spheres(ctx, flat->spheres.data(), (int)flat->spheres.size()));
		Check(agpt_upload_planes(ctx, flat->planes.data(), (int)flat->planes.size()));
		Check(agpt_upload_materials(ctx, flat->materials.data(), (int)flat->materials.size()));
		Check(agpt_upload_lights(ctx, flat->lights.data(), (int)flat->lights.size()));
		Check(agpt_upload_envmap(ctx, flat->envmap.width > 0 ? &flat->envmap : nullptr));
		Check(agpt_upload_primitives(ctx, flat->prims.data(), (int)flat->prims.size()));
		uploaded = scene.Fingerprint();
	}

	// numSamples Tick bodies: samples firstSample .. firstSample+numSamples-1 of every pixel,
	// added to `acc` (which keeps its earlier contents, like successive Ticks do).
	void Render(const Scene& scene, const Camera& camera, Accumulator& acc, int firstSample, int numSamples,
			int depth = 0, uint32_t flags = 0) const {
		if (uploaded != scene.Fingerprint()) Upload(scene);
		agpt_camera cam = camera.Export();
		Check(agpt_set_camera(ctx, &cam));
		Check(agpt_set_film(ctx, acc.width, acc.height));
		Check(agpt_write_accum(ctx, &acc.Pixels()->x));
		Check(agpt_render(ctx, firstSample, numSamples, 1, MaxDepth, depth, flags));
		Check(agpt_read_accum(ctx, &acc.Pixels()->x));
		acc.SetNumSamples(acc.NumSamples() + numSamples);
	}

	// Single-ray entry point.  The reference's Li draws from the global generator; here the
	// ray gets the RNG stream of (pixel 0, sample liCalls++), film/camera are not involved.
	float3 Li(const Ray& ray, const Scene& scene, int depth = 0) const override;

	agpt_ctx* Context() const { return ctx; }
	int GetMaxDepth() const { return MaxDepth; }

protected:
	static void Check(int status) {
		if (status != AGPT_OK) throw std::runtime_error(std::string("agpt: ") + agpt_last_error());
	}
	int MaxDepth;
	agpt_ctx* ctx = nullptr;
	mutable uint64_t uploaded = 0;     // Scene::Fingerprint() of the resident tables (0: none)
	mutable unsigned liCalls = 0;
};

// (defined here so that the mirror stays header-only; libagpt.so is the only link dependency)
inline float3 CudaPathTracer::Li(const Ray& ray, const Scene& scene, int depth) const {
	if (uploaded != scene.Fingerprint()) Upload(scene);
	float r7[7] = { ray.O.x, ray.O.y, ray.O.z, ray.D.x, ray.D.y, ray.D.z, ray.t };
	uint32_t seed = 0x12345678u + 0x9e3779b9u * liCalls++;   // upstream's global seed, advanced per call
	float out[3] = { 0, 0, 0 };
	Check(agpt_li_rays(ctx, 1, r7, &seed, MaxDepth, depth, AGPT_FLAG_RAYS_FINAL, out));      // ray.D is final: the Ray ctor normalised it
	return float3(out[0], out[1], out[2]);
}

using PathTracer = CudaPathTracer;   // scene code written against the reference keeps compiling
