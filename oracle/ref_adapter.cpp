// oracle/ref_adapter.cpp -- TEST INFRASTRUCTURE, not product code.
//
// The reference-side adapter (tests/adapter/cuda_pathtracer.h: CudaPathTracer : Integrator over the REFERENCE's
// own Scene / BVHTriMesh / material / light / Camera classes) compiled against the unmodified headers under
// /root/reference and linked with libagpt.so.  Built as a library of its own (oracle/_ref/libagpt_ref_adapter.so)
// so that the plain reference library -- what bench.py's reference arm loads -- never pulls the product in.
// It reuses the harness (scene construction, the restated Tick body) by including it.
#include "ref_harness.cpp"

#include "cuda_pathtracer.h"        // tests/adapter

static thread_local std::string g_adapterError;

extern "C" {

const char* agpt_ref_adapter_error() { return g_adapterError.c_str(); }

// Samples [s0, s0+ns) of the whole W x H film: reference Scene + Camera objects -> adapter -> GPU.  Same output
// contract as agpt_ref_render (accumulated into out_rgba, accumulator layout).  rgb8_out (optional): the
// adapter's CopyToSurface of the film after these samples, assuming the film held samples_before samples.
int agpt_ref_adapter_render(void* h, int W, int H, int s0, int ns, int samples_before, int max_depth, int depth_arg, float* out_rgba, unsigned* rgb8_out) {
	auto* rs = (RefScene*)h;
	try {
		CudaPathTracer integrator(max_depth);
		static_assert(sizeof(float3) == 16, "float3 stride");
		integrator.Render(rs->scene, *rs->camera, reinterpret_cast<float3*>(out_rgba), W, H, s0, ns, depth_arg);
		if (rgb8_out) integrator.CopyToSurface(samples_before + ns, rgb8_out);
	}
	catch (const std::exception& e) { g_adapterError = e.what(); return -1; }
	return 0;
}

// Integrator::Li through the adapter against PathTracer::Li on the CPU for the camera rays of the given film
// positions: both start from generator state `state` (the adapter takes one RandomUInt() draw as the path's
// stream, so the CPU side discards one draw too).  out_gpu / out_cpu: 3 floats per ray.
int agpt_ref_adapter_li(void* h, int n, const float* uv2, unsigned state, int max_depth, int depth_arg, float* out_gpu, float* out_cpu) {
	auto* rs = (RefScene*)h;
	try {
		CudaPathTracer gpu(max_depth);
		PathTracer cpu(max_depth);
		const Integrator* both[2] = { &gpu, &cpu };
		for (int i = 0; i < n; i++) {
			Ray ray = rs->camera->GetRay(uv2[2 * i], uv2[2 * i + 1]);
			for (int k = 0; k < 2; k++) {
				agpt_ref_set_state(state + 7919u * (unsigned)i);
				if (k == 1) RandomUInt();
				Ray r = ray;
				float3 c = both[k]->Li(r, rs->scene, depth_arg);
				float* o = (k == 0 ? out_gpu : out_cpu) + 3 * i;
				o[0] = c.x; o[1] = c.y; o[2] = c.z;
			}
		}
	}
	catch (const std::exception& e) { g_adapterError = e.what(); return -1; }
	return 0;
}

} // extern "C"
