// examples/render_ppm.cpp -- scene code written in the reference's style (compare
// /root/reference/myapp.cpp:13-114,121-135) compiled against the B200 host mirror, rendering
// headless and writing the resolved frame (Accumulator::CopyToSurface semantics, done on the
// device by agpt_resolve) as a PPM.
//
//   make -C ag-pathtracer_b200 && g++ -std=c++17 -O2 -Iinclude -Iag-pathtracer_b200/host examples/render_ppm.cpp \
//       -Lag-pathtracer_b200 -lagpt -Wl,-rpath,$PWD/ag-pathtracer_b200 -o build/render_ppm && build/render_ppm out.ppm 7 64
#include "precomp.h"
#include "scene.h"
#include "integrator.h"
#include "scenes/config_scenes.h"

int main(int argc, char** argv) {
	const char* out = argc > 1 ? argv[1] : "out.ppm";
	int config = argc > 2 ? atoi(argv[2]) : 7;
	int spp = argc > 3 ? atoi(argv[3]) : 64;
	auto defaults = agpt_scenes::Defaults(config);
	int W = defaults.width ? defaults.width : 400, H = defaults.height ? defaults.height : 400;
	int maxDepth = defaults.max_depth ? defaults.max_depth : 5;

	auto scene = std::make_shared<Scene>();                 // MyApp::Init (myapp.cpp:121-135)
	if (!agpt_scenes::BuildConfig(scene.get(), config, 0)) { fprintf(stderr, "unknown config\n"); return 1; }
	Camera camera(scene->camera);
	auto integrator = std::make_shared<PathTracer>(maxDepth);    // PathTracer == CudaPathTracer here
	Accumulator accumulator(W, H);

	integrator->Render(*scene, camera, accumulator, 0, spp, defaults.depth_arg);   // spp x MyApp::Tick

	std::vector<uint32_t> rgb((size_t)W * H);
	if (agpt_resolve(integrator->Context(), accumulator.NumSamples(), rgb.data()) != AGPT_OK) { fprintf(stderr, "%s\n", agpt_last_error()); return 1; }
	FILE* f = fopen(out, "wb");
	if (!f) return 1;
	fprintf(f, "P6\n%d %d\n255\n", W, H);
	for (uint32_t p : rgb) { unsigned char c[3] = { (unsigned char)(p >> 16), (unsigned char)(p >> 8), (unsigned char)p }; fwrite(c, 1, 3, f); }
	fclose(f);
	agpt_stats st;
	agpt_get_stats(integrator->Context(), &st);
	printf("%s: %dx%d, %d spp, %.1f ms on the device, %.1f Mrays/s\n", out, W, H, spp, st.ms_render,
		(st.rays_closest + st.rays_shadow + st.rays_mis) / st.ms_render / 1e3);
	return 0;
}
