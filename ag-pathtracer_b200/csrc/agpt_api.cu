// agpt_api.cu -- libagpt.so: the extern "C" boundary of include/agpt.h over the sm_100a
// wavefront kernels.  Host code here only moves tables into HBM, sizes launches and drives
// the wave loop; every ray is traced and shaded on the device.  There is no CPU path: if no
// CUDA device can be opened agpt_create() fails and the caller gets the CUDA error text.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "agpt_kernels.cuh"
#include "agpt_multigpu.cuh"

// ---- error plumbing ----------------------------------------------------------------------
static thread_local std::string g_error;
static int Fail(int code, const std::string& msg) { g_error = msg; return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return Fail(AGPT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
#define NEED(cond, code, msg) do { if (!(cond)) return Fail(code, msg); } while (0)

template <typename T>
struct DevBuf {
	T* p = nullptr;
	size_t n = 0;
	cudaError_t Alloc(size_t count) {
		Free();
		if (count == 0) return cudaSuccess;
		cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
		if (e == cudaSuccess) n = count;
		return e;
	}
	cudaError_t Upload(const T* host, size_t count, cudaStream_t s) {
		cudaError_t e = Alloc(count);
		if (e != cudaSuccess || count == 0) return e;
		return cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s);
	}
	void Free() { if (p) cudaFree(p); p = nullptr; n = 0; }
	size_t Bytes() const { return n * sizeof(T); }
};

struct MeshStore {
	DevBuf<float4> nodes, tris, normals;
	DevBuf<float2> uvs;
	DevBuf<int> ids;
	void Free() { nodes.Free(); tris.Free(); normals.Free(); uvs.Free(); ids.Free(); }
};

#ifndef AGPT_BATCH_LOG2
#define AGPT_BATCH_LOG2 27
#endif
// Paths in flight per batch, at most.  2^27 = 134 M slots ~ 39 GB of wavefront state + queues (sized for 180 GB of
// HBM3e; the state is allocated for the batches a caller actually asks for: 16 spp at 1080p take 33 M slots).
// Bigger batches mean fuller waves and more rays per bucket, i.e. more coherent warps: cfg 3 runs a 16-spp
// share in 106 / 98 / 94 ms with batches of 2^23 / 2^24 / 2^25 slots (round 1's kernels) and in 57.9 / 55.8 /
// 54.4 ms with 2^25 / 2^26 / 2^27 (16 / 32 / 64 spp per batch at 1080p, round 2's).
static const size_t kMaxPathsPerBatch = (size_t)1 << AGPT_BATCH_LOG2;

struct agpt_ctx {
	int device = 0;
	cudaStream_t ownStream = nullptr, stream = nullptr;
	cudaStream_t sideStream = nullptr;   // any-hit trace of a wave runs here, beside the closest-hit trace
	cudaStream_t copyStream = nullptr;   // agpt_write_accum_begin: the film upload runs here, beside the render that follows
	cudaEvent_t evUpload = nullptr;
	bool uploadPending = false;          // a film upload is in flight: k_accumulate waits for it, every other film access joins it first
	cudaEvent_t evFork = nullptr, evJoin = nullptr;
	cudaEvent_t evA = nullptr, evB = nullptr, evC = nullptr, evD = nullptr;
	cudaEvent_t evRender0 = nullptr, evRender1 = nullptr;      // bracket of agpt_render (ms_render)
	int smCount = 0;

	// scene tables
	std::vector<MeshStore> meshStore;
	DevBuf<DMesh> meshes;
	DevBuf<agpt_sphere> spheres;
	DevBuf<agpt_plane> planes;
	DevBuf<agpt_prim> prims;
	DevBuf<agpt_instance> instances;     // extension: placed meshes
	std::vector<agpt_instance> hostInstances;
	int nInstances = 0;
	bool hasGlass = false;        // some material is a rough dielectric (extension): shade runs the GLASS instantiation
	DevBuf<int> sphereRun;
	DevBuf<float4> sphereRunBox;
	DevBuf<float4> keyBoxes;         // root boxes of the meshes that do not cover the scene (ray-bucket key)
	std::vector<agpt_bvh_node> hostRoots;   // node 0 of every mesh (count < 0: the mesh has no BVH)
	std::vector<agpt_sphere> hostSpheres;
	bool runsDirty = true;
	int maxSphereRun = 1 << 20;   // AGPT_MAX_SPHERE_RUN
	DevBuf<agpt_material> mats;
	DevBuf<agpt_light> lights;
	DevBuf<float> envRgb, envFunc, envCdf;
	float envFuncInt = 0;
	int envW = 0, envH = 0;
	std::vector<agpt_prim> hostPrims;
	std::vector<agpt_light> hostLights;
	int nMeshes = 0, nSpheres = 0, nPlanes = 0, nMats = 0;
	int maxBvhDepth = 0;          // deepest uploaded tree (levels below the root), validated against the traversal stack
	agpt_camera cam;
	bool haveCam = false;
	float boundsLo[3] = { -1, -1, -1 }, boundsHi[3] = { 1, 1, 1 };   // bounded geometry (mesh roots), for ray bucketing only
	float gridLo[3] = { -1, -1, -1 }, gridHi[3] = { 1, 1, 1 };       // ... refined from where the camera's rays actually land (CalibrateBucketGrid)
	bool gridCalibrated = false;
	bool calibrateGrid = true;    // AGPT_BUCKET_CALIBRATE=0: keep the grid on the mesh bounds
	int width = 0, height = 0;

	// film
	DevBuf<float4> accumOwn;
	DevBuf<uint32_t> resolved;    // agpt_resolve's packed output, kept between calls
	float4* accum = nullptr;      // accumOwn.p or caller-owned

	// wavefront state
	size_t capacity = 0;          // path slots allocated
	DevBuf<float4> f4[13];
	DevBuf<int> i32[3];
	DevBuf<uint32_t> u32[2];
	DevBuf<int> queues[6];        // closest A/B (2*cap), shadow A/B, active A/B
	DevBuf<int> sortedClosest;    // closest queue in bucket order (2*cap)
	DevBuf<unsigned short> keys[2], shadowKeys[2];
	DevBuf<int> sortedShadow;
	DevBuf<int> hist;             // AGPT_BUCKETS x { histogram, offsets, running }
	DevBuf<int> counts;           // 2 x 3
	DevBuf<int> survivors, survivorCount;   // paths k_shade_b works on this wave
	DevBuf<unsigned long long> traceCounters;   // [2][8] totals (closest-hit, any-hit), then [2][AGPT_MAX_WAVE_ROWS][8] per wave (FlushCounters)
	DevBuf<RayCounters> rayCounters;
	int* hostCounts = nullptr;    // pinned ring of kRing x 3 ints
	cudaEvent_t ringEvents[8] = {};

	// multi-process sharding: the other ranks' accumulators, mapped through CUDA IPC (agpt_open_peer_accums)
	float4* peerAccum[AGPT_MAX_PEERS] = {};
	int peerRank = -1, peerWorld = 0;

	agpt_stats stats;
	bool asyncWaves = false;      // AGPT_ASYNC_WAVES=1: run one wave ahead of the landed queue counts instead of syncing
	                              // every wave.  Measured slower (N=1: 133 vs 130.5 ms/step, N=8: 142.8 vs 139.3): the loose
	                              // launch bounds and the extra empty wave cost more than the ~30 us sync gaps they remove.
	size_t maxPathsPerBatch = kMaxPathsPerBatch;   // AGPT_BATCH_LOG2 (environment) overrides the compiled-in size
	bool overlapAny = true;       // AGPT_OVERLAP_ANY=0: any-hit trace on the main stream after the closest-hit trace
	bool bucketRays = true;       // bucket pass on the ray queues (AGPT_BUCKET_RAYS=0 turns it off)
};

// Runs of sphere primitives whose payloads are consecutive: the trace kernels walk such a run
// straight through the sphere table, and skip it when no ray of the warp can reach its box.
static int BuildSphereRuns(agpt_ctx* c) {
	const int n = (int)c->hostPrims.size();
	std::vector<int> run((size_t)n, 0);
	std::vector<float4> box(3 * (size_t)n, make_float4(0.f, 0.f, 0.f, 0.f));
	const std::vector<agpt_prim>& rows = c->hostPrims;
	for (int p = n - 1; p >= 0; p--) {
		if (rows[p].type != AGPT_PRIM_SPHERE) continue;
		bool chained = p + 1 < n && rows[p + 1].type == AGPT_PRIM_SPHERE && rows[p + 1].payload == rows[p].payload + 1;
		run[p] = chained && run[p + 1] < c->maxSphereRun ? run[p + 1] + 1 : 1;     // capped: shorter runs have tighter boxes
	}
	for (int p = 0; p < n; p++) {
		if (run[p] == 0) continue;
		float lo[3] = { 1e30f, 1e30f, 1e30f }, hi[3] = { -1e30f, -1e30f, -1e30f }, rmin = 1e30f;
		bool ok = true;
		for (int j = 0; j < run[p]; j++) {
			int k = rows[p + j].payload;
			if (k < 0 || k >= (int)c->hostSpheres.size()) { ok = false; break; }
			const agpt_sphere& sp = c->hostSpheres[k];
			if (!(sp.r > 0.f)) { ok = false; break; }
			for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], sp.center[a] - sp.r); hi[a] = fmaxf(hi[a], sp.center[a] + sp.r); }
			rmin = fminf(rmin, sp.r);
		}
		float4* b = &box[3 * (size_t)p];
		if (!ok) { b[2] = make_float4(0.f, 0.f, 0.f, -1.f); continue; }      // never cull
		const float grow = 0.05f * rmin;
		float ctr[3], diag2 = 0.f;
		for (int a = 0; a < 3; a++) { lo[a] -= grow; hi[a] += grow; ctr[a] = 0.5f * (lo[a] + hi[a]); float h = 0.5f * (hi[a] - lo[a]); diag2 += h * h; }
		// Rounding of Sphere::Intersect's discriminant is at most ~22 * 2^-24 |oc|^2 = 1.3e-6 |oc|^2 (every
		// product, sum and input rounding counted against us); a ray that misses the grown box has
		// d^2 - r^2 > 2 r grow = 0.1 rmin^2.  With |oc|^2 <= 2 (|O - ctr|^2 + diag2) kept below 2e4 rmin^2 the
		// rounding stays under 0.026 rmin^2: a factor 4 of slack on a worst-case bound.
		b[0] = make_float4(lo[0], lo[1], lo[2], hi[0]);
		b[1] = make_float4(hi[1], hi[2], 0.f, 0.f);
		b[2] = make_float4(ctr[0], ctr[1], ctr[2], 1.0e4f * rmin * rmin - diag2);
	}
	// ray-bucket key: the root boxes of BVH mesh primitives that are small against the union of all of them
	{
		std::vector<float4> kb;
		float lo[3] = { 1e30f, 1e30f, 1e30f }, hi[3] = { -1e30f, -1e30f, -1e30f };
		std::vector<int> meshPrims;
		for (int p = 0; p < n; p++)
			if (rows[p].type == AGPT_PRIM_BVH_MESH && rows[p].payload >= 0 && rows[p].payload < (int)c->hostRoots.size() && c->hostRoots[rows[p].payload].count >= 0) {
				meshPrims.push_back(rows[p].payload);
				for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], c->hostRoots[rows[p].payload].bmin[a]); hi[a] = fmaxf(hi[a], c->hostRoots[rows[p].payload].bmax[a]); }
			}
		double all = 1;
		for (int a = 0; a < 3; a++) all *= fmax((double)hi[a] - lo[a], 1e-30);
		for (int m : meshPrims) {
			const agpt_bvh_node& r = c->hostRoots[m];
			double v = 1;
			for (int a = 0; a < 3; a++) v *= fmax((double)r.bmax[a] - r.bmin[a], 1e-30);
			if (v <= 0.25 * all && kb.size() < 2 * 16) { kb.push_back(make_float4(r.bmin[0], r.bmin[1], r.bmin[2], r.bmax[0])); kb.push_back(make_float4(r.bmax[1], r.bmax[2], 0.f, 0.f)); }
		}
		CU(c->keyBoxes.Upload(kb.data(), kb.size(), c->stream));
	}
	CU(c->sphereRun.Upload(run.data(), (size_t)n, c->stream));
	CU(c->sphereRunBox.Upload(box.data(), box.size(), c->stream));
	CU(cudaStreamSynchronize(c->stream));
	c->runsDirty = false;
	return AGPT_OK;
}

static void SetBucketGrid(DScene& s, const float* lo, const float* hi) {
	for (int a = 0; a < 3; a++) { s.cellLo[a] = lo[a]; float e = hi[a] - lo[a]; s.cellScale[a] = e > 0 ? (float)(1 << AGPT_CELL_BITS) / e : 0.f; }
}

static DScene MakeScene(const agpt_ctx* c) {
	DScene s;
	s.keyBoxes = c->keyBoxes.p; s.n_keyBoxes = (int)(c->keyBoxes.n / 2);
	s.prims = c->prims.p; s.sphereRun = c->sphereRun.p; s.sphereRunBox = c->sphereRunBox.p; s.spheres = c->spheres.p; s.planes = c->planes.p; s.meshes = c->meshes.p; s.instances = c->instances.p;
	s.mats = c->mats.p; s.lights = c->lights.p;
	s.n_prims = (int)c->prims.n; s.n_lights = (int)c->lights.n;
	s.envRgb = c->envRgb.p; s.envFunc = c->envFunc.p; s.envCdf = c->envCdf.p; s.envFuncInt = c->envFuncInt; s.envW = c->envW; s.envH = c->envH;
	s.width = c->width; s.height = c->height;
	SetBucketGrid(s, c->gridCalibrated ? c->gridLo : c->boundsLo, c->gridCalibrated ? c->gridHi : c->boundsHi);
	s.cam = c->cam;
	return s;
}

static PathState MakePathState(agpt_ctx* c) {
	PathState p;
	p.rayO = c->f4[0].p; p.rayD = c->f4[1].p; p.hitA = c->f4[2].p; p.beta = c->f4[3].p; p.L = c->f4[4].p;
	p.neeLight = c->f4[5].p; p.neeMis = c->f4[6].p; p.neeBeta = c->f4[7].p;
	p.shO = c->f4[8].p; p.shD = c->f4[9].p; p.misO = c->f4[10].p; p.misD = c->f4[11].p; p.Lout = c->f4[12].p;
	p.hitSlot = c->i32[0].p; p.shadowOccluded = c->i32[1].p; p.misPrim = c->i32[2].p;
	p.rng = c->u32[0].p; p.flags = c->u32[1].p;
	p.slots = (int)c->capacity;
	return p;
}

static int EnsureCapacity(agpt_ctx* c, size_t paths) {
	if (paths <= c->capacity) return AGPT_OK;
	c->capacity = 0;                 // published again only when every buffer below is there (a failed allocation must not leave a stale size)
	for (auto& b : c->f4) CU(b.Alloc(paths));
	for (auto& b : c->i32) CU(b.Alloc(paths));
	for (auto& b : c->u32) CU(b.Alloc(paths));
	const size_t slack = 1024;       // kernels read queue[i] before they know the queue length (i < grid * block)
	CU(c->queues[0].Alloc(2 * paths + slack)); CU(c->queues[1].Alloc(2 * paths + slack));
	CU(c->sortedClosest.Alloc(2 * paths + slack)); CU(c->keys[0].Alloc(2 * paths + slack)); CU(c->keys[1].Alloc(2 * paths + slack));
	CU(c->shadowKeys[0].Alloc(paths + slack)); CU(c->shadowKeys[1].Alloc(paths + slack)); CU(c->sortedShadow.Alloc(paths + slack));
	for (int k = 2; k < 6; k++) CU(c->queues[k].Alloc(paths + slack));
	CU(c->survivors.Alloc(paths + slack));
	c->capacity = paths;
	return AGPT_OK;
}

static int BuildSphereRuns(agpt_ctx* c);

static int CheckReady(agpt_ctx* c, bool needFilm) {
	NEED(c != nullptr, AGPT_ERR_INVALID, "null context");
	NEED(c->prims.n > 0, AGPT_ERR_STATE, "no primitives uploaded (agpt_upload_primitives)");
	if (needFilm) {
		NEED(c->width > 0 && c->height > 0, AGPT_ERR_STATE, "film not set (agpt_set_film)");
		NEED(c->haveCam, AGPT_ERR_STATE, "camera not set (agpt_set_camera)");
	}
	// table consistency: every row must point inside its table
	for (size_t i = 0; i < c->hostPrims.size(); i++) {
		const agpt_prim& p = c->hostPrims[i];
		int limit = p.type == AGPT_PRIM_SPHERE ? c->nSpheres : p.type == AGPT_PRIM_PLANE ? c->nPlanes : p.type == AGPT_PRIM_INSTANCE ? c->nInstances : c->nMeshes;
		NEED(p.type >= 0 && p.type <= 4 && p.payload >= 0 && p.payload < limit, AGPT_ERR_INVALID, "primitive row points outside its shape table");
		NEED(p.material < c->nMats, AGPT_ERR_INVALID, "primitive row points outside the material table");
		NEED(p.area_light < (int)c->hostLights.size(), AGPT_ERR_INVALID, "primitive row points outside the light table");
	}
	for (auto& in : c->hostInstances)
		NEED(in.mesh >= 0 && in.mesh < c->nMeshes, AGPT_ERR_INVALID, "instance row points outside the mesh table");
	for (auto& l : c->hostLights)
		NEED(l.type != AGPT_LIGHT_INFINITE_AREA || c->envW > 0, AGPT_ERR_STATE, "InfiniteAreaLight without an environment map (agpt_upload_envmap)");
	for (auto& l : c->hostLights)
		NEED(l.type != AGPT_LIGHT_AREA || (l.prim >= 0 && l.prim < (int)c->hostPrims.size()), AGPT_ERR_INVALID, "area light without a primitive");
	// An AreaLight wraps exactly one shape upstream (scene.h addAreaLight; lights.h:70-86), and the exact
	// MIS-ray cull relies on it: the primitive a light names must be the only one that names the light.
	for (size_t i = 0; i < c->hostPrims.size(); i++) {
		int al = c->hostPrims[i].area_light;
		NEED(al < 0 || (c->hostLights[al].type == AGPT_LIGHT_AREA && c->hostLights[al].prim == (int)i), AGPT_ERR_INVALID,
			"primitive and area light do not name each other (one shape per AreaLight)");
	}
	if (c->mats.p == nullptr) {
		// a scene of null-material (emissive) shapes only is legal upstream; shade still needs a record to point idle lanes at
		CU(cudaSetDevice(c->device));
		agpt_material none;
		memset(&none, 0, sizeof(none));
		CU(c->mats.Upload(&none, 1, c->stream));
		CU(cudaStreamSynchronize(c->stream));
	}
	if (c->runsDirty) {
		CU(cudaSetDevice(c->device));
		int rcode = BuildSphereRuns(c);
		if (rcode != AGPT_OK) return rcode;
	}
	return AGPT_OK;
}

static inline int Blocks(size_t n, int threads) { return (int)((n + threads - 1) / threads); }

// A film upload started by agpt_write_accum_begin may still be in flight: whatever touches the film next waits for it.
static int JoinUpload(agpt_ctx* c) {
	if (c != nullptr && c->uploadPending) {
		CU(cudaSetDevice(c->device));
		CU(cudaStreamSynchronize(c->copyStream));
		c->uploadPending = false;
	}
	return AGPT_OK;
}
#define JOIN(c) do { int j_ = JoinUpload(c); if (j_ != AGPT_OK) return j_; } while (0)

// ---- launch helpers: template flags from run-time flags --------------------------------------
// (COUNT, FAST = !strictBoxes, INST = the scene has instances) -> one of eight instantiations
#define AGPT_DISPATCH3(count, fast, inst, CALL) do { \
	if (count) { if (fast) { if (inst) { CALL(true, true, true); } else { CALL(true, true, false); } } else { if (inst) { CALL(true, false, true); } else { CALL(true, false, false); } } } \
	else { if (fast) { if (inst) { CALL(false, true, true); } else { CALL(false, true, false); } } else { if (inst) { CALL(false, false, true); } else { CALL(false, false, false); } } } } while (0)

static void LaunchClosest(bool count, bool strictBoxes, bool inst, int blocks, cudaStream_t st, const DScene& sc, const PathState& ps, const int* queue, const int* n, unsigned long long* cnt, unsigned long long* row) {
#define CALL(C, F, I) k_trace_closest<C, F, I><<<blocks, AGPT_TRACE_THREADS, 0, st>>>(sc, ps, queue, n, cnt, row)
	AGPT_DISPATCH3(count, !strictBoxes, inst, CALL);
#undef CALL
}
static void LaunchAny(bool count, bool strictBoxes, bool inst, int blocks, cudaStream_t st, const DScene& sc, const PathState& ps, const int* queue, const int* n, unsigned long long* cnt, unsigned long long* row) {
#define CALL(C, F, I) k_trace_any<C, F, I><<<blocks, AGPT_TRACE_THREADS, 0, st>>>(sc, ps, queue, n, cnt, row)
	AGPT_DISPATCH3(count, !strictBoxes, inst, CALL);
#undef CALL
}
template <bool ANY>
static void LaunchTable(bool count, bool strictBoxes, bool inst, int blocks, cudaStream_t st, const DScene& sc, const float4* o, const float4* d, int n, agpt_hit* out, unsigned long long* cnt) {
#define CALL(C, F, I) k_trace_table<ANY, C, F, I><<<blocks, AGPT_TRACE_THREADS, 0, st>>>(sc, o, d, n, out, cnt)
	AGPT_DISPATCH3(count, !strictBoxes, inst, CALL);
#undef CALL
}


extern "C" {

const char* agpt_last_error(void) { return g_error.c_str(); }

int agpt_device_count(int* out) {
	NEED(out != nullptr, AGPT_ERR_INVALID, "null out");
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) { *out = 0; return Fail(AGPT_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
	*out = n;
	return AGPT_OK;
}

int agpt_create(int device, agpt_ctx** out) {
	NEED(out != nullptr, AGPT_ERR_INVALID, "null out");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0)
		return Fail(AGPT_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
			" (libagpt has no CPU path)");
	NEED(device >= 0 && device < n, AGPT_ERR_INVALID, "device index out of range");
	CU(cudaSetDevice(device));
	agpt_ctx* c = new agpt_ctx();
	c->device = device;
	memset(&c->stats, 0, sizeof(c->stats));
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, device));
	c->smCount = prop.multiProcessorCount;
	CU(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking));
	c->stream = c->ownStream;
	CU(cudaStreamCreateWithFlags(&c->sideStream, cudaStreamNonBlocking));
	CU(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
	CU(cudaEventCreateWithFlags(&c->evUpload, cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
	CU(cudaEventCreate(&c->evA)); CU(cudaEventCreate(&c->evB)); CU(cudaEventCreate(&c->evC)); CU(cudaEventCreate(&c->evD));
	CU(cudaEventCreate(&c->evRender0)); CU(cudaEventCreate(&c->evRender1));
	CU(c->counts.Alloc(6));
	CU(c->survivorCount.Alloc(1));
	CU(c->hist.Alloc(3 * AGPT_BUCKETS));
	CU(c->traceCounters.Alloc(2 * AGPT_WAVE_COUNTERS * (1 + AGPT_MAX_WAVE_ROWS)));
	CU(c->rayCounters.Alloc(1));
	CU(cudaMemset(c->traceCounters.p, 0, c->traceCounters.Bytes()));
	CU(cudaMemset(c->rayCounters.p, 0, c->rayCounters.Bytes()));
	CU(cudaMallocHost((void**)&c->hostCounts, 8 * 3 * sizeof(int)));
	for (auto& e : c->ringEvents) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	if (const char* e = getenv("AGPT_ASYNC_WAVES")) c->asyncWaves = atoi(e) != 0;
	if (const char* e = getenv("AGPT_BATCH_LOG2")) { int b = atoi(e); if (b >= 10 && b <= 28) c->maxPathsPerBatch = (size_t)1 << b; }
	if (const char* e = getenv("AGPT_BUCKET_CALIBRATE")) c->calibrateGrid = atoi(e) != 0;
	if (const char* e = getenv("AGPT_MAX_SPHERE_RUN")) { int v = atoi(e); if (v >= 1) c->maxSphereRun = v; }
	if (const char* e = getenv("AGPT_OVERLAP_ANY")) c->overlapAny = atoi(e) != 0;
	if (const char* e = getenv("AGPT_BUCKET_RAYS")) c->bucketRays = atoi(e) != 0;
	*out = c;
	return AGPT_OK;
}

int agpt_destroy(agpt_ctx* c) {
	JoinUpload(c);          // (an upload still in flight must not outlive its buffers; a failure here changes nothing about what follows)
	if (!c) return AGPT_OK;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	for (int r = 0; r < c->peerWorld; r++) if (r != c->peerRank && c->peerAccum[r]) cudaIpcCloseMemHandle(c->peerAccum[r]);
	for (auto& m : c->meshStore) m.Free();
	c->meshes.Free(); c->instances.Free(); c->spheres.Free(); c->planes.Free(); c->prims.Free(); c->sphereRun.Free(); c->sphereRunBox.Free(); c->keyBoxes.Free(); c->mats.Free(); c->lights.Free();
	c->accumOwn.Free(); c->resolved.Free();
	c->envRgb.Free(); c->envFunc.Free(); c->envCdf.Free();
	for (auto& b : c->f4) b.Free();
	for (auto& b : c->i32) b.Free();
	for (auto& b : c->u32) b.Free();
	for (auto& b : c->queues) b.Free();
	c->sortedClosest.Free(); c->keys[0].Free(); c->keys[1].Free(); c->hist.Free();
	c->shadowKeys[0].Free(); c->shadowKeys[1].Free(); c->sortedShadow.Free();
	c->survivors.Free(); c->survivorCount.Free();
	c->counts.Free(); c->traceCounters.Free(); c->rayCounters.Free();
	if (c->hostCounts) cudaFreeHost(c->hostCounts);
	for (auto& e : c->ringEvents) if (e) cudaEventDestroy(e);
	cudaEventDestroy(c->evA); cudaEventDestroy(c->evB); cudaEventDestroy(c->evC); cudaEventDestroy(c->evD);
	cudaEventDestroy(c->evRender0); cudaEventDestroy(c->evRender1);
	cudaEventDestroy(c->evFork); cudaEventDestroy(c->evJoin);
	cudaStreamDestroy(c->sideStream);
	cudaStreamDestroy(c->copyStream);
	cudaEventDestroy(c->evUpload);
	cudaStreamDestroy(c->ownStream);
	delete c;
	return AGPT_OK;
}

int agpt_set_stream(agpt_ctx* c, void* s) {
	NEED(c != nullptr, AGPT_ERR_INVALID, "null context");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	c->stream = s ? (cudaStream_t)s : c->ownStream;
	return AGPT_OK;
}

// ---- scene upload ------------------------------------------------------------------------
// One pass over a caller-supplied node table before anything is uploaded: every interior link and
// leaf range must stay inside its table, the links must form a TREE (a child reached twice -- a
// shared child or a link back to an ancestor -- would make the walk loop or double-count), and the
// tree must not be deeper than the traversal stack.  The reference recurses without a limit
// (bvhtrimesh.h:332-413); the device stack holds AGPT_STACK_SMEM + AGPT_STACK_LOCAL far children,
// far more than any SAH tree over 2^31 triangles needs, and deeper tables are refused here
// instead of corrupting memory there.
static int ValidateBvh(const agpt_mesh_desc& d, int meshIndex, int* depthOut) {
	*depthOut = 0;
	if (d.n_nodes == 0) return AGPT_OK;
	const std::string where = "mesh " + std::to_string(meshIndex) + ": ";
	NEED(d.n_nodes == 1 || d.n_nodes >= 3, AGPT_ERR_INVALID, where + "node table of 2 entries (root at 0, slot 1 unused, children from 2)");
	std::vector<unsigned char> seen((size_t)d.n_nodes, 0);
	std::vector<std::pair<int, int>> todo;      // node, depth
	todo.emplace_back(0, 0);
	seen[0] = 1;
	int maxDepth = 0;
	while (!todo.empty()) {
		auto [k, depth] = todo.back();
		todo.pop_back();
		const agpt_bvh_node& nd = d.nodes[k];
		if (depth > maxDepth) maxDepth = depth;
		if (nd.count > 0) {
			NEED(nd.first >= 0 && (long long)nd.first + nd.count <= d.n_tris, AGPT_ERR_INVALID, where + "BVH leaf range outside the triangle table");
			continue;
		}
		NEED(nd.count == 0, AGPT_ERR_INVALID, where + "BVH node with a negative count");
		NEED(nd.first >= 2 && nd.first + 1 < d.n_nodes, AGPT_ERR_INVALID, where + "BVH child link outside the node table");
		NEED(!seen[nd.first] && !seen[nd.first + 1], AGPT_ERR_INVALID, where + "BVH node table is not a tree (a child is linked twice, or links back to an ancestor)");
		seen[nd.first] = seen[nd.first + 1] = 1;
		NEED(depth + 1 <= AGPT_STACK_SMEM + AGPT_STACK_LOCAL, AGPT_ERR_INVALID,
			where + "BVH deeper than " + std::to_string(AGPT_STACK_SMEM + AGPT_STACK_LOCAL) + " levels (the device traversal stack)");
		todo.emplace_back(nd.first, depth + 1);
		todo.emplace_back(nd.first + 1, depth + 1);
	}
	*depthOut = maxDepth;
	return AGPT_OK;
}

int agpt_upload_meshes(agpt_ctx* c, const agpt_mesh_desc* meshes, int n) {
	NEED(c != nullptr && n >= 0 && (n == 0 || meshes != nullptr), AGPT_ERR_INVALID, "bad mesh table");
	// validate every descriptor before touching what is resident: a refused upload leaves the old scene intact
	int maxDepth = 0;
	for (int i = 0; i < n; i++) {
		const agpt_mesh_desc& d = meshes[i];
		NEED(d.n_tris >= 0 && d.n_nodes >= 0 && (d.n_tris == 0 || (d.tri_verts && d.tri_ids)), AGPT_ERR_INVALID, "mesh without triangle data");
		NEED(d.n_nodes == 0 || d.nodes != nullptr, AGPT_ERR_INVALID, "mesh node table missing");
		int depth = 0;
		int rcode = ValidateBvh(d, i, &depth);
		if (rcode != AGPT_OK) return rcode;
		if (depth > maxDepth) maxDepth = depth;
	}
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	c->gridCalibrated = false;
	// from here on a failure must not leave the device table pointing at freed buffers
	for (auto& m : c->meshStore) m.Free();
	c->meshStore.assign(n, MeshStore());
	c->meshes.Free();
	c->nMeshes = 0;
	c->maxBvhDepth = 0;
	std::vector<DMesh> table(n);
	for (int i = 0; i < n; i++) {
		const agpt_mesh_desc& d = meshes[i];
		MeshStore& st = c->meshStore[i];
		CU(st.nodes.Upload((const float4*)d.nodes, 2 * (size_t)d.n_nodes, c->stream));
		CU(st.tris.Upload((const float4*)d.tri_verts, 3 * (size_t)d.n_tris, c->stream));
		CU(st.ids.Upload(d.tri_ids, (size_t)d.n_tris, c->stream));
		if (d.tri_normals) CU(st.normals.Upload((const float4*)d.tri_normals, 3 * (size_t)d.n_tris, c->stream));
		if (d.tri_uvs) CU(st.uvs.Upload((const float2*)d.tri_uvs, 3 * (size_t)d.n_tris, c->stream));
		if (d.n_tris > 0) {
			k_flag_degenerate<<<Blocks(d.n_tris, 256), 256, 0, c->stream>>>(st.tris.p, st.uvs.p, d.n_tris);
			CU(cudaGetLastError());
			c->stats.kernel_launches++;
		}
		DMesh& m = table[i];
		m.nodes = st.nodes.p; m.tris = st.tris.p; m.ids = st.ids.p; m.normals = st.normals.p; m.uvs = st.uvs.p;
		m.n_nodes = d.n_nodes; m.n_tris = d.n_tris;
	}
	CU(c->meshes.Upload(table.data(), (size_t)n, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	c->nMeshes = n;
	c->maxBvhDepth = maxDepth;
	c->hostRoots.assign((size_t)n, agpt_bvh_node());
	for (int i = 0; i < n; i++) { if (meshes[i].n_nodes > 0) c->hostRoots[i] = meshes[i].nodes[0]; else c->hostRoots[i].count = -1; }
	c->runsDirty = true;
	bool any = false;
	for (int i = 0; i < n; i++) {
		if (meshes[i].n_nodes == 0) continue;
		const agpt_bvh_node& r = meshes[i].nodes[0];
		for (int a = 0; a < 3; a++) {
			c->boundsLo[a] = any ? (r.bmin[a] < c->boundsLo[a] ? r.bmin[a] : c->boundsLo[a]) : r.bmin[a];
			c->boundsHi[a] = any ? (r.bmax[a] > c->boundsHi[a] ? r.bmax[a] : c->boundsHi[a]) : r.bmax[a];
		}
		any = true;
	}
	return AGPT_OK;
}

#define SIMPLE_UPLOAD(fn, T, member, counter) \
	int fn(agpt_ctx* c, const T* rows, int n) { \
		NEED(c != nullptr && n >= 0 && (n == 0 || rows != nullptr), AGPT_ERR_INVALID, "bad table"); \
		CU(cudaSetDevice(c->device)); \
		CU(cudaStreamSynchronize(c->stream)); \
		CU(c->member.Upload(rows, (size_t)n, c->stream)); \
		CU(cudaStreamSynchronize(c->stream)); \
		counter; \
		return AGPT_OK; \
	}
SIMPLE_UPLOAD(agpt_upload_spheres, agpt_sphere, spheres, c->nSpheres = n; c->hostSpheres.assign(rows, rows + n); c->runsDirty = true)
SIMPLE_UPLOAD(agpt_upload_planes, agpt_plane, planes, c->nPlanes = n)
SIMPLE_UPLOAD(agpt_upload_materials, agpt_material, mats, c->nMats = n; c->hasGlass = false; for (int i = 0; i < n; i++) if (rows[i].lobes & (AGPT_LOBE_GLASS_REFLECT | AGPT_LOBE_GLASS_TRANSMIT)) c->hasGlass = true)
SIMPLE_UPLOAD(agpt_upload_lights, agpt_light, lights, c->hostLights.assign(rows, rows + n))
int agpt_upload_primitives(agpt_ctx* c, const agpt_prim* rows, int n) {
	NEED(c != nullptr && n >= 0 && (n == 0 || rows != nullptr), AGPT_ERR_INVALID, "bad table");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	CU(c->prims.Upload(rows, (size_t)n, c->stream));
	c->hostPrims.assign(rows, rows + n);
	c->runsDirty = true;
	CU(cudaStreamSynchronize(c->stream));
	return AGPT_OK;
}

SIMPLE_UPLOAD(agpt_upload_instances, agpt_instance, instances, c->nInstances = n; c->hostInstances.assign(rows, rows + n))

int agpt_upload_envmap(agpt_ctx* c, const agpt_envmap* env) {
	NEED(c != nullptr, AGPT_ERR_INVALID, "null context");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	c->envRgb.Free(); c->envFunc.Free(); c->envCdf.Free();
	c->envW = c->envH = 0; c->envFuncInt = 0;
	if (!env || env->width == 0) return AGPT_OK;
	NEED(env->width > 0 && env->height > 0 && env->rgb && env->func && env->cdf, AGPT_ERR_INVALID, "bad environment map");
	size_t n = (size_t)env->width * env->height;
	CU(c->envRgb.Upload(env->rgb, 3 * n, c->stream));
	CU(c->envFunc.Upload(env->func, n, c->stream));
	CU(c->envCdf.Upload(env->cdf, n + 1, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	c->envW = env->width; c->envH = env->height; c->envFuncInt = env->func_int;
	return AGPT_OK;
}

int agpt_set_camera(agpt_ctx* c, const agpt_camera* cam) {
	NEED(c != nullptr && cam != nullptr, AGPT_ERR_INVALID, "null camera");
	c->cam = *cam;
	c->haveCam = true;
	c->gridCalibrated = false;
	return AGPT_OK;
}

int agpt_set_film(agpt_ctx* c, int width, int height) {
	JOIN(c);
	NEED(c != nullptr && width > 0 && height > 0 && (long long)width * height < (1ll << 30), AGPT_ERR_INVALID, "bad film size");
	CU(cudaSetDevice(c->device));
	if (width == c->width && height == c->height) return AGPT_OK;
	CU(cudaStreamSynchronize(c->stream));
	bool own = c->accum == c->accumOwn.p;
	c->width = width; c->height = height;
	if (own) {
		CU(c->accumOwn.Alloc((size_t)width * height));
		c->accum = c->accumOwn.p;
		CU(cudaMemsetAsync(c->accum, 0, c->accumOwn.Bytes(), c->stream));
	}
	return AGPT_OK;
}

int agpt_scene_bytes(agpt_ctx* c, uint64_t* out) {
	NEED(c != nullptr && out != nullptr, AGPT_ERR_INVALID, "null argument");
	uint64_t b = c->envRgb.Bytes() + c->envFunc.Bytes() + c->envCdf.Bytes() + c->meshes.Bytes() + c->spheres.Bytes() + c->planes.Bytes() + c->prims.Bytes() + c->instances.Bytes() + c->mats.Bytes() + c->lights.Bytes();
	for (auto& m : c->meshStore) b += m.nodes.Bytes() + m.tris.Bytes() + m.normals.Bytes() + m.uvs.Bytes() + m.ids.Bytes();
	*out = b;
	return AGPT_OK;
}

// ---- accumulator -------------------------------------------------------------------------
int agpt_clear(agpt_ctx* c) {
	JOIN(c);
	NEED(c != nullptr && c->accum != nullptr, AGPT_ERR_STATE, "film not set");
	CU(cudaSetDevice(c->device));
	CU(cudaMemsetAsync(c->accum, 0, (size_t)c->width * c->height * sizeof(float4), c->stream));
	return AGPT_OK;
}
int agpt_accum_ptr_dev(agpt_ctx* c, void** p) {
	JOIN(c);
	NEED(c != nullptr && p != nullptr && c->accum != nullptr, AGPT_ERR_STATE, "film not set");
	*p = c->accum;
	return AGPT_OK;
}
int agpt_set_accum_dev(agpt_ctx* c, void* p) {
	JOIN(c);
	NEED(c != nullptr, AGPT_ERR_INVALID, "null context");
	NEED(c->width > 0, AGPT_ERR_STATE, "film not set");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	if (p) c->accum = (float4*)p;
	else {
		if (!c->accumOwn.p) { CU(c->accumOwn.Alloc((size_t)c->width * c->height)); CU(cudaMemset(c->accumOwn.p, 0, c->accumOwn.Bytes())); }
		c->accum = c->accumOwn.p;
	}
	return AGPT_OK;
}
int agpt_read_accum(agpt_ctx* c, float* host) {
	JOIN(c);
	NEED(c != nullptr && host != nullptr && c->accum != nullptr, AGPT_ERR_STATE, "film not set");
	CU(cudaSetDevice(c->device));
	CU(cudaMemcpyAsync(host, c->accum, (size_t)c->width * c->height * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return AGPT_OK;
}
int agpt_write_accum(agpt_ctx* c, const float* host) {
	JOIN(c);
	NEED(c != nullptr && host != nullptr && c->accum != nullptr, AGPT_ERR_STATE, "film not set");
	CU(cudaSetDevice(c->device));
	CU(cudaMemcpyAsync(c->accum, host, (size_t)c->width * c->height * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return AGPT_OK;
}
int agpt_write_accum_begin(agpt_ctx* c, const float* host) {
	JOIN(c);
	NEED(c != nullptr && host != nullptr && c->accum != nullptr, AGPT_ERR_STATE, "film not set");
	CU(cudaSetDevice(c->device));
	// everything queued on the render stream so far (a clear, an earlier render's accumulate) comes first
	CU(cudaEventRecord(c->evUpload, c->stream));
	CU(cudaStreamWaitEvent(c->copyStream, c->evUpload, 0));
	CU(cudaMemcpyAsync(c->accum, host, (size_t)c->width * c->height * sizeof(float4), cudaMemcpyHostToDevice, c->copyStream));
	CU(cudaEventRecord(c->evUpload, c->copyStream));
	c->uploadPending = true;
	return AGPT_OK;
}
int agpt_resolve(agpt_ctx* c, int samples, uint32_t* host) {
	JOIN(c);
	NEED(c != nullptr && host != nullptr && c->accum != nullptr && samples > 0, AGPT_ERR_STATE, "film not set or samples <= 0");
	CU(cudaSetDevice(c->device));
	int n = c->width * c->height;
	if (c->resolved.n != (size_t)n) CU(c->resolved.Alloc(n));
	k_resolve<<<Blocks(n, 256), 256, 0, c->stream>>>(c->accum, c->resolved.p, n, (float)samples);
	CU(cudaGetLastError());
	c->stats.kernel_launches++;
	CU(cudaMemcpyAsync(host, c->resolved.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return AGPT_OK;
}

// ---- the wave loop -------------------------------------------------------------------------
// Runs `n` already-generated paths (slots 0..n-1, queues A and counts A filled by k_generate) to
// completion.  Queue lengths stay on the device (kernels read them); the host sizes grids from
// upper bounds.  Default: after every wave the host waits for that wave's three counts (one
// 12-byte copy to pinned memory), so the bounds are exact and the loop ends with the wave that
// empties the active list.  Optional (asyncWaves): the host runs up to kMaxAhead waves ahead of
// the newest landed counts, tightening bounds as copies land (the active list only shrinks: a
// later wave has at most `active` paths, 2*active closest-hit rays and `active` shadow rays);
// empty waves launched past the end are no-ops.  AGPT_FLAG_TIMING brackets the kernel classes
// of every wave with events.
static const int kMaxAhead = 2, kRing = 8;
static const int kSmallWave = 16384;
// The bucket grid should resolve the part of the scene the rays are in, not the scene's bounding
// box (cfg 3: a 40-unit backdrop around 15 units of objects left 12 of the 512 cells in use).
// Once per scene and camera, after the first camera rays have been traced, a strided sample of
// their hit points is read back and the grid is laid over mean +- 2.5 sigma of it.  Ray order
// only: results do not depend on the grid.
static int CalibrateBucketGrid(agpt_ctx* c, const PathState& ps, int n) {
	const int samples = n < 65536 ? n : 65536;
	const size_t pitch = (size_t)(n / samples) * sizeof(float4);
	std::vector<float4> o(samples), d(samples), h(samples);
	CU(cudaMemcpy2DAsync(o.data(), sizeof(float4), (const float4*)ps.rayO, pitch, sizeof(float4), samples, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaMemcpy2DAsync(d.data(), sizeof(float4), (const float4*)ps.rayD, pitch, sizeof(float4), samples, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaMemcpy2DAsync(h.data(), sizeof(float4), (const float4*)ps.hitA, pitch, sizeof(float4), samples, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	double sum[3] = { 0, 0, 0 }, sq[3] = { 0, 0, 0 };
	long hits = 0;
	for (int i = 0; i < samples; i++) {
		int prim; memcpy(&prim, &h[i].w, sizeof(prim));
		float t = h[i].x;
		if (prim < 0 || !(t > 0.f) || !(t < 1e30f)) continue;
		double p[3] = { o[i].x + (double)t * d[i].x, o[i].y + (double)t * d[i].y, o[i].z + (double)t * d[i].z };
		for (int a = 0; a < 3; a++) { sum[a] += p[a]; sq[a] += p[a] * p[a]; }
		hits++;
	}
	c->gridCalibrated = true;
	for (int a = 0; a < 3; a++) { c->gridLo[a] = c->boundsLo[a]; c->gridHi[a] = c->boundsHi[a]; }
	if (hits < 256) return AGPT_OK;                       // too little to go by: keep the mesh bounds
	for (int a = 0; a < 3; a++) {
		double mean = sum[a] / hits, var = sq[a] / hits - mean * mean;
		double sigma = var > 0 ? sqrt(var) : 0;
		if (sigma <= 0) continue;
		c->gridLo[a] = (float)(mean - 2.5 * sigma);
		c->gridHi[a] = (float)(mean + 2.5 * sigma);
	}
	return AGPT_OK;
}

static int RunWaves(agpt_ctx* c, const DScene& scIn, PathState& ps, int n, int max_depth, int rr_depth_arg, uint32_t flags) {
	const bool count = flags & AGPT_FLAG_COUNTERS, timing = flags & AGPT_FLAG_TIMING, strictBoxes = flags & AGPT_FLAG_STRICT_BOXES;
	DScene sc = scIn;
	WaveQueues q[2];
	for (int k = 0; k < 2; k++) {
		q[k].closest = c->queues[0 + k].p; q[k].shadow = c->queues[2 + k].p; q[k].active = c->queues[4 + k].p;
		q[k].counts = c->counts.p + 3 * k;
		q[k].keys = c->keys[k].p; q[k].shadowKeys = c->shadowKeys[k].p;
	}
	int* bucketHist = c->hist.p;
	int* bucketOffsets = c->hist.p + AGPT_BUCKETS;
	int* bucketRunning = c->hist.p + 2 * AGPT_BUCKETS;
	const bool bucketing = c->bucketRays;
	const bool overlap = c->overlapAny && !timing;
	float msClosest = 0, msAny = 0, msShade = 0;
	unsigned long long* cntClosest = c->traceCounters.p;
	unsigned long long* cntAny = c->traceCounters.p + AGPT_WAVE_COUNTERS;
	unsigned long long* waveRows = c->traceCounters.p + 2 * AGPT_WAVE_COUNTERS;      // [kind][wave][8]

	int ubClosest = n, ubShadow = 0, ubActive = n;   // upper bounds of the current wave's queue lengths
	int cur = 0, wave = 0;
	int ringHead = 0, ringTail = 0;                  // copies in flight: [ringTail, ringHead)
	bool done = false;
	while (!done) {
		// ---- one wave on q[cur] -> q[cur ^ 1] ----
		const int* closestQueue = q[cur].closest;
		const int* shadowQueue = q[cur].shadow;
		WaveQueues qin = q[cur];
		if (wave > 0 && bucketing) {
			// bucket pass: rays that start in the same cell going the same way end up adjacent
			// (not worth its six launches for the last few thousand rays of a batch)
			const int perBlock = 256 * AGPT_BUCKET_ITEMS;
			if (ubClosest >= kSmallWave) {
				CU(cudaMemsetAsync(bucketHist, 0, AGPT_BUCKETS * sizeof(int), c->stream));
				k_bucket_hist<<<Blocks(ubClosest, perBlock), 256, 0, c->stream>>>(q[cur].keys, q[cur].counts + 0, bucketHist);
				k_bucket_scan<<<1, 1024, 0, c->stream>>>(bucketHist, bucketOffsets, bucketRunning);
				k_bucket_scatter<<<Blocks(ubClosest, perBlock), 256, 0, c->stream>>>(q[cur].closest, q[cur].keys, q[cur].counts + 0, bucketOffsets, bucketRunning, c->sortedClosest.p);
				c->stats.kernel_launches += 3;
				closestQueue = c->sortedClosest.p;
			}
			if (ubShadow >= kSmallWave) {
				CU(cudaMemsetAsync(bucketHist, 0, AGPT_BUCKETS * sizeof(int), c->stream));
				k_bucket_hist<<<Blocks(ubShadow, perBlock), 256, 0, c->stream>>>(q[cur].shadowKeys, q[cur].counts + 1, bucketHist);
				k_bucket_scan<<<1, 1024, 0, c->stream>>>(bucketHist, bucketOffsets, bucketRunning);
				k_bucket_scatter<<<Blocks(ubShadow, perBlock), 256, 0, c->stream>>>(q[cur].shadow, q[cur].shadowKeys, q[cur].counts + 1, bucketOffsets, bucketRunning, c->sortedShadow.p);
				c->stats.kernel_launches += 3;
				shadowQueue = c->sortedShadow.p;
			}
		}
		// The two traces of a wave are independent (different queues, different result arrays):
		// unless per-kernel timing is on, the any-hit trace runs on the side stream so that each
		// fills the SMs the other's tail leaves idle.
		const bool fork = overlap && ubClosest > 0 && ubShadow > 0;
		if (timing) CU(cudaEventRecord(c->evA, c->stream));
		if (fork) CU(cudaEventRecord(c->evFork, c->stream));
		if (ubClosest > 0) {
			LaunchClosest(count, strictBoxes, c->nInstances > 0, Blocks(ubClosest, AGPT_TRACE_THREADS), c->stream, sc, ps, closestQueue, q[cur].counts + 0, cntClosest,
				waveRows + (size_t)AGPT_WAVE_COUNTERS * (wave < AGPT_MAX_WAVE_ROWS ? wave : AGPT_MAX_WAVE_ROWS - 1));
			c->stats.kernel_launches++; c->stats.launches_closest++;
		}
		if (wave == 0 && bucketing && !c->gridCalibrated && c->calibrateGrid && n >= 4096) {
			// first camera rays of this scene and camera: lay the bucket grid where they land
			int rcode = CalibrateBucketGrid(c, ps, n);
			if (rcode != AGPT_OK) return rcode;
			SetBucketGrid(sc, c->gridLo, c->gridHi);
		}
		if (timing) CU(cudaEventRecord(c->evB, c->stream));
		if (fork) {
			CU(cudaStreamWaitEvent(c->sideStream, c->evFork, 0));
			LaunchAny(count, strictBoxes, c->nInstances > 0, Blocks(ubShadow, AGPT_TRACE_THREADS), c->sideStream, sc, ps, shadowQueue, q[cur].counts + 1, cntAny,
				waveRows + (size_t)AGPT_WAVE_COUNTERS * (AGPT_MAX_WAVE_ROWS + (wave < AGPT_MAX_WAVE_ROWS ? wave : AGPT_MAX_WAVE_ROWS - 1)));
			c->stats.kernel_launches++; c->stats.launches_any++;
			CU(cudaEventRecord(c->evJoin, c->sideStream));
			CU(cudaStreamWaitEvent(c->stream, c->evJoin, 0));
		} else if (ubShadow > 0) {
			LaunchAny(count, strictBoxes, c->nInstances > 0, Blocks(ubShadow, AGPT_TRACE_THREADS), c->stream, sc, ps, shadowQueue, q[cur].counts + 1, cntAny,
				waveRows + (size_t)AGPT_WAVE_COUNTERS * (AGPT_MAX_WAVE_ROWS + (wave < AGPT_MAX_WAVE_ROWS ? wave : AGPT_MAX_WAVE_ROWS - 1)));
			c->stats.kernel_launches++; c->stats.launches_any++;
		}
		CU(cudaMemsetAsync(q[cur ^ 1].counts, 0, 3 * sizeof(int), c->stream));
		if (timing) CU(cudaEventRecord(c->evC, c->stream));
		ShadeParams sp;
		sp.count = q[cur].counts + 2; sp.max_depth = max_depth; sp.rr_depth_arg = rr_depth_arg; sp.rr_by_bounce = (flags & AGPT_FLAG_RR_BY_BOUNCE) ? 1 : 0; sp.exact_counts = count ? 1 : 0;
		// shade: k_shade_a (per active entry: NEE fold, emission, termination -> survivor list),
		// then k_shade_b (per survivor: the BSDF work), blocks striding over the list
		CU(cudaMemsetAsync(c->survivorCount.p, 0, sizeof(int), c->stream));
		sp.count = c->survivorCount.p;
		const int shadeBlocks = Blocks(ubActive, AGPT_SHADE_THREADS);      // upper bound: blocks past the survivor count return at once
		if (c->envW > 0) k_shade_a<true><<<Blocks(ubActive, 256), 256, 0, c->stream>>>(sc, ps, qin.active, q[cur].counts + 2, c->survivors.p, c->survivorCount.p, max_depth);
		else k_shade_a<false><<<Blocks(ubActive, 256), 256, 0, c->stream>>>(sc, ps, qin.active, q[cur].counts + 2, c->survivors.p, c->survivorCount.p, max_depth);
		// (ENV: the scene has an InfiniteAreaLight, GLASS: a rough-dielectric material -- each instantiation carries only what it needs)
#define SHADE_B(E, G) k_shade_b<E, G><<<shadeBlocks, AGPT_SHADE_THREADS, 0, c->stream>>>(sc, ps, c->survivors.p, q[cur ^ 1], sp, c->rayCounters.p)
		if (c->envW > 0) { if (c->hasGlass) SHADE_B(true, true); else SHADE_B(true, false); }
		else { if (c->hasGlass) SHADE_B(false, true); else SHADE_B(false, false); }
#undef SHADE_B
		c->stats.kernel_launches++;
		c->stats.kernel_launches++; c->stats.launches_shade++;
		CU(cudaGetLastError());
		if (timing) CU(cudaEventRecord(c->evD, c->stream));
		// counts of the NEXT wave -> pinned ring slot, marked by an event
		int slot = ringHead % kRing;
		CU(cudaMemcpyAsync(c->hostCounts + 3 * slot, q[cur ^ 1].counts, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
		CU(cudaEventRecord(c->ringEvents[slot], c->stream));
		ringHead++;
		cur ^= 1; wave++;
		c->stats.waves++;
		// the next wave's bounds follow from this wave's: the active list only shrinks
		ubClosest = 2 * ubActive < (int)(2 * c->capacity) ? 2 * ubActive : (int)(2 * c->capacity);
		ubShadow = ubActive;

		// ---- consume the copies that have landed (block only when too far ahead, or when timing) ----
		while (ringTail < ringHead) {
			int s = ringTail % kRing;
			// small waves (the stragglers at the end of a batch) are launched from the last known
			// bounds without waiting for their counts: nothing to lose on a few thousand entries
			bool runAhead = c->asyncWaves || ubActive < kSmallWave;
			bool mustWait = timing || !runAhead || (ringHead - ringTail) >= kMaxAhead;
			cudaError_t e = mustWait ? cudaEventSynchronize(c->ringEvents[s]) : cudaEventQuery(c->ringEvents[s]);
			if (e == cudaErrorNotReady) break;
			if (e != cudaSuccess) return Fail(AGPT_ERR_CUDA, std::string("wave loop: ") + cudaGetErrorString(e));
			const int* hc = c->hostCounts + 3 * s;        // (closest, shadow, active) of wave ringTail + 1
			int landedWave = ringTail + 1;
			ringTail++;
			if (hc[2] == 0) { done = true; break; }
			if (landedWave == wave) { ubClosest = hc[0]; ubShadow = hc[1]; ubActive = hc[2]; }
			else {
				// older than the wave about to be launched: still bounds it
				if (hc[2] < ubActive) ubActive = hc[2];
				if (2 * hc[2] < ubClosest) ubClosest = 2 * hc[2];
				if (hc[2] < ubShadow) ubShadow = hc[2];
			}
		}
		if (timing) {
			float a = 0, b = 0, d = 0;
			cudaEventElapsedTime(&a, c->evA, c->evB);
			cudaEventElapsedTime(&b, c->evB, c->evC);
			cudaEventElapsedTime(&d, c->evC, c->evD);
			msClosest += a; msAny += b; msShade += d;
		}
	}
	c->stats.ms_trace_closest += msClosest; c->stats.ms_trace_any += msAny; c->stats.ms_shade += msShade;
	return AGPT_OK;
}

int agpt_render(agpt_ctx* c, int first_sample, int num_samples, int sample_stride, int max_depth, int rr_depth_arg, uint32_t flags) {
	int rcode = CheckReady(c, true);
	if (rcode != AGPT_OK) return rcode;
	NEED(num_samples >= 0 && sample_stride >= 1 && max_depth >= 0, AGPT_ERR_INVALID, "bad sample range or depth");
	NEED(c->accum != nullptr, AGPT_ERR_STATE, "no accumulator");
	CU(cudaSetDevice(c->device));
	const size_t wh = (size_t)c->width * c->height;
	int perBatch = (int)(c->maxPathsPerBatch / wh);
	if (perBatch < 1) perBatch = 1;
	if (perBatch > num_samples) perBatch = num_samples;
	if (num_samples == 0) return AGPT_OK;
	rcode = EnsureCapacity(c, wh * perBatch);
	if (rcode != AGPT_OK) return rcode;
	DScene sc = MakeScene(c);
	PathState ps = MakePathState(c);
	WaveQueues q0;
	q0.closest = c->queues[0].p; q0.shadow = c->queues[2].p; q0.active = c->queues[4].p; q0.counts = c->counts.p; q0.keys = nullptr; q0.shadowKeys = nullptr;
	// evA/evC/evD are reused inside RunWaves when timing; the render bracket has its own pair
	cudaEvent_t r0 = c->evRender0, r1 = c->evRender1;
	CU(cudaEventRecord(r0, c->stream));
	for (int done = 0; done < num_samples; done += perBatch) {
		int ns = num_samples - done < perBatch ? num_samples - done : perBatch;
		int n = (int)(wh * ns);
		GenParams g;
		memset(&g, 0, sizeof(g));
		g.n = n; g.first_sample = first_sample + done * sample_stride; g.sample_stride = sample_stride; g.tiled = 1; g.samples = ns;
		k_generate<<<Blocks(n, 256), 256, 0, c->stream>>>(sc, ps, q0, g);
		CU(cudaGetLastError());
		c->stats.kernel_launches++;
		c->stats.paths += (uint64_t)n;
		c->stats.rays_closest += (uint64_t)n;     // camera rays; the rest is counted on the device
		rcode = RunWaves(c, sc, ps, n, max_depth, rr_depth_arg, flags);
		if (rcode != AGPT_OK) return rcode;
		if (c->uploadPending) CU(cudaStreamWaitEvent(c->stream, c->evUpload, 0));      // the film this batch adds to may still be on its way (agpt_write_accum_begin)
		k_accumulate<<<Blocks(wh, 256), 256, 0, c->stream>>>(ps.Lout, c->accum, c->width, c->height, ns);
		CU(cudaGetLastError());
		c->stats.kernel_launches++;
	}
	CU(cudaEventRecord(r1, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	c->uploadPending = false;        // k_accumulate waited for it
	float ms = 0;
	cudaEventElapsedTime(&ms, r0, r1);
	c->stats.ms_render += ms;
	return AGPT_OK;
}

static int TraceTable(agpt_ctx* c, const DScene& sc, const float4* rayO, const float4* rayD, int n, int any_hit, uint32_t flags, agpt_hit* out_host) {
	DevBuf<agpt_hit> out;
	CU(out.Alloc(n));
	const bool count = flags & AGPT_FLAG_COUNTERS, strictBoxes = flags & AGPT_FLAG_STRICT_BOXES;
	int blocks = Blocks(n, AGPT_TRACE_THREADS);
	if (any_hit) LaunchTable<true>(count, strictBoxes, c->nInstances > 0, blocks, c->stream, sc, rayO, rayD, n, out.p, c->traceCounters.p + AGPT_WAVE_COUNTERS);
	else LaunchTable<false>(count, strictBoxes, c->nInstances > 0, blocks, c->stream, sc, rayO, rayD, n, out.p, c->traceCounters.p);
	CU(cudaGetLastError());
	c->stats.kernel_launches++;
	CU(cudaMemcpyAsync(out_host, out.p, (size_t)n * sizeof(agpt_hit), cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	out.Free();
	return AGPT_OK;
}

int agpt_trace_primary(agpt_ctx* c, int sample, uint32_t flags, agpt_hit* out_host) {
	int rcode = CheckReady(c, true);
	if (rcode != AGPT_OK) return rcode;
	NEED(out_host != nullptr, AGPT_ERR_INVALID, "null output");
	CU(cudaSetDevice(c->device));
	const size_t wh = (size_t)c->width * c->height;
	rcode = EnsureCapacity(c, wh);
	if (rcode != AGPT_OK) return rcode;
	DScene sc = MakeScene(c);
	PathState ps = MakePathState(c);
	WaveQueues q0;
	q0.closest = c->queues[0].p; q0.shadow = c->queues[2].p; q0.active = c->queues[4].p; q0.counts = c->counts.p; q0.keys = nullptr; q0.shadowKeys = nullptr;
	GenParams g;
	memset(&g, 0, sizeof(g));
	g.n = (int)wh; g.first_sample = sample; g.sample_stride = 1;
	k_generate<<<Blocks(wh, 256), 256, 0, c->stream>>>(sc, ps, q0, g);
	CU(cudaGetLastError());
	c->stats.kernel_launches++;
	return TraceTable(c, sc, ps.rayO, ps.rayD, (int)wh, 0, flags, out_host);
}

int agpt_trace_rays(agpt_ctx* c, int64_t n, const float* rays7, int any_hit, uint32_t flags, agpt_hit* out_host) {
	int rcode = CheckReady(c, false);
	if (rcode != AGPT_OK) return rcode;
	NEED(n >= 0 && n < (1ll << 30) && (n == 0 || (rays7 && out_host)), AGPT_ERR_INVALID, "bad ray table");
	if (n == 0) return AGPT_OK;
	CU(cudaSetDevice(c->device));
	// same normalisation as the Ray constructor, done on the device by k_generate
	rcode = EnsureCapacity(c, (size_t)n);
	if (rcode != AGPT_OK) return rcode;
	DScene sc = MakeScene(c);
	PathState ps = MakePathState(c);
	WaveQueues q0;
	q0.closest = c->queues[0].p; q0.shadow = c->queues[2].p; q0.active = c->queues[4].p; q0.counts = c->counts.p; q0.keys = nullptr; q0.shadowKeys = nullptr;
	DevBuf<float> rays;
	DevBuf<uint32_t> seeds;
	CU(rays.Upload(rays7, 7 * (size_t)n, c->stream));
	CU(seeds.Alloc((size_t)n));
	CU(cudaMemsetAsync(seeds.p, 0, seeds.Bytes(), c->stream));
	GenParams g;
	memset(&g, 0, sizeof(g));
	g.n = (int)n; g.rays7 = rays.p; g.seeds = seeds.p; g.raysFinal = (flags & AGPT_FLAG_RAYS_FINAL) ? 1 : 0;
	k_generate<<<Blocks(n, 256), 256, 0, c->stream>>>(sc, ps, q0, g);
	CU(cudaGetLastError());
	c->stats.kernel_launches++;
	rcode = TraceTable(c, sc, ps.rayO, ps.rayD, (int)n, any_hit, flags, out_host);
	rays.Free(); seeds.Free();
	return rcode;
}

static int LiGeneric(agpt_ctx* c, GenParams g, int max_depth, int rr_depth_arg, uint32_t flags, float* out_rgb) {
	int n = g.n;
	int rcode = EnsureCapacity(c, (size_t)n);
	if (rcode != AGPT_OK) return rcode;
	DScene sc = MakeScene(c);
	PathState ps = MakePathState(c);
	WaveQueues q0;
	q0.closest = c->queues[0].p; q0.shadow = c->queues[2].p; q0.active = c->queues[4].p; q0.counts = c->counts.p; q0.keys = nullptr; q0.shadowKeys = nullptr;
	k_generate<<<Blocks(n, 256), 256, 0, c->stream>>>(sc, ps, q0, g);
	CU(cudaGetLastError());
	c->stats.kernel_launches++;
	c->stats.paths += (uint64_t)n;
	c->stats.rays_closest += (uint64_t)n;
	rcode = RunWaves(c, sc, ps, n, max_depth, rr_depth_arg, flags & AGPT_FLAG_RR_BY_BOUNCE);
	if (rcode != AGPT_OK) return rcode;
	std::vector<float4> host(n);
	CU(cudaMemcpyAsync(host.data(), ps.Lout, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	for (int i = 0; i < n; i++) { out_rgb[3 * i] = host[i].x; out_rgb[3 * i + 1] = host[i].y; out_rgb[3 * i + 2] = host[i].z; }
	return AGPT_OK;
}

int agpt_li_pixels(agpt_ctx* c, int n, const int* xs, const int* ys, const int* ss, int max_depth, int rr_depth_arg, uint32_t flags, float* out_rgb) {
	int rcode = CheckReady(c, true);
	if (rcode != AGPT_OK) return rcode;
	NEED(n >= 0 && (n == 0 || (xs && ys && ss && out_rgb)), AGPT_ERR_INVALID, "bad pixel list");
	if (n == 0) return AGPT_OK;
	CU(cudaSetDevice(c->device));
	DevBuf<int> dx, dy, ds;
	CU(dx.Upload(xs, n, c->stream)); CU(dy.Upload(ys, n, c->stream)); CU(ds.Upload(ss, n, c->stream));
	GenParams g;
	memset(&g, 0, sizeof(g));
	g.n = n; g.xs = dx.p; g.ys = dy.p; g.ss = ds.p;
	rcode = LiGeneric(c, g, max_depth, rr_depth_arg, flags, out_rgb);
	dx.Free(); dy.Free(); ds.Free();
	return rcode;
}

int agpt_li_rays(agpt_ctx* c, int n, const float* rays7, const uint32_t* rng_states, int max_depth, int rr_depth_arg, uint32_t flags, float* out_rgb) {
	int rcode = CheckReady(c, false);
	if (rcode != AGPT_OK) return rcode;
	NEED(n >= 0 && (n == 0 || (rays7 && rng_states && out_rgb)), AGPT_ERR_INVALID, "bad ray list");
	if (n == 0) return AGPT_OK;
	CU(cudaSetDevice(c->device));
	DevBuf<float> rays;
	DevBuf<uint32_t> seeds;
	CU(rays.Upload(rays7, 7 * (size_t)n, c->stream));
	CU(seeds.Upload(rng_states, (size_t)n, c->stream));
	GenParams g;
	memset(&g, 0, sizeof(g));
	g.n = n; g.rays7 = rays.p; g.seeds = seeds.p; g.raysFinal = (flags & AGPT_FLAG_RAYS_FINAL) ? 1 : 0;
	rcode = LiGeneric(c, g, max_depth, rr_depth_arg, flags, out_rgb);
	rays.Free(); seeds.Free();
	return rcode;
}

// ---- multi-GPU: sample-index sharding ---------------------------------------------
// One context per GPU, every context holds the whole scene; GPU g renders s = first + g, first + g + G, ...
static int CheckGroup(agpt_ctx** ctxs, int n) {
	for (int g = 0; ctxs != nullptr && g < n; g++) JOIN(ctxs[g]);
	NEED(ctxs != nullptr && n >= 1 && n <= AGPT_MAX_PEERS, AGPT_ERR_INVALID, "bad context list (1.." + std::to_string(AGPT_MAX_PEERS) + " contexts)");
	for (int g = 0; g < n; g++) {
		NEED(ctxs[g] != nullptr && ctxs[g]->accum != nullptr, AGPT_ERR_STATE, "context " + std::to_string(g) + ": film not set");
		NEED(ctxs[g]->width == ctxs[0]->width && ctxs[g]->height == ctxs[0]->height, AGPT_ERR_INVALID, "contexts have different film sizes");
		for (int h = 0; h < g; h++) NEED(ctxs[h] != ctxs[g] && ctxs[h]->device != ctxs[g]->device, AGPT_ERR_INVALID, "two contexts on one GPU");
	}
	return AGPT_OK;
}

// Peer access between every pair of the group's GPUs; *all = false if some pair has none.
static int EnablePeerAccess(agpt_ctx** ctxs, int n, bool* all) {
	*all = true;
	for (int a = 0; a < n; a++)
		for (int b = 0; b < n; b++) {
			if (a == b) continue;
			int can = 0;
			CU(cudaDeviceCanAccessPeer(&can, ctxs[a]->device, ctxs[b]->device));
			if (!can) { *all = false; continue; }
			CU(cudaSetDevice(ctxs[a]->device));
			cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[b]->device, 0);
			if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
			else if (e != cudaSuccess) return Fail(AGPT_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
		}
	return AGPT_OK;
}

static int SyncGroup(agpt_ctx** ctxs, int n) {
	for (int g = 0; g < n; g++) { CU(cudaSetDevice(ctxs[g]->device)); CU(cudaStreamSynchronize(ctxs[g]->stream)); }
	return AGPT_OK;
}

static NcclApi g_nccl;
static std::vector<int> g_ncclDevices;
static std::vector<agpt_ncclComm_t> g_ncclComms;

static int NcclAllReduce(agpt_ctx** ctxs, int n) {
	NEED(g_nccl.Load(), AGPT_ERR_STATE, g_nccl.error);
	std::vector<int> devs(n);
	for (int g = 0; g < n; g++) devs[g] = ctxs[g]->device;
	if (devs != g_ncclDevices) {
		for (auto cm : g_ncclComms) g_nccl.CommDestroy(cm);
		g_ncclComms.assign(n, nullptr);
		int r = g_nccl.CommInitAll(g_ncclComms.data(), n, devs.data());
		if (r != 0) { g_ncclComms.clear(); g_ncclDevices.clear(); return Fail(AGPT_ERR_CUDA, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r)); }
		g_ncclDevices = devs;
	}
	const size_t count = 4 * (size_t)ctxs[0]->width * ctxs[0]->height;
	int r = g_nccl.GroupStart();
	for (int g = 0; g < n && r == 0; g++) {
		CU(cudaSetDevice(ctxs[g]->device));
		r = g_nccl.AllReduce(ctxs[g]->accum, ctxs[g]->accum, count, /*ncclFloat*/ 7, /*ncclSum*/ 0, g_ncclComms[g], ctxs[g]->stream);
	}
	int r2 = g_nccl.GroupEnd();
	if (r != 0 || r2 != 0) return Fail(AGPT_ERR_CUDA, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r != 0 ? r : r2));
	return AGPT_OK;
}

static bool WantNccl() { const char* e = getenv("AGPT_REDUCE"); return e && std::string(e) == "nccl"; }

static void SliceOf(int g, int n, size_t wh, int* first, int* count) {
	size_t a = wh * (size_t)g / n, b = wh * (size_t)(g + 1) / n;
	*first = (int)a; *count = (int)(b - a);
}

int agpt_reduce_accum(agpt_ctx** ctxs, int n, int root) {
	int rcode = CheckGroup(ctxs, n);
	if (rcode != AGPT_OK) return rcode;
	NEED(root >= -1 && root < n, AGPT_ERR_INVALID, "bad root");
	if (n == 1) return AGPT_OK;
	bool peers = false;
	rcode = EnablePeerAccess(ctxs, n, &peers);
	if (rcode != AGPT_OK) return rcode;
	rcode = SyncGroup(ctxs, n);                    // every render has landed before any accumulator is read
	if (rcode != AGPT_OK) return rcode;
	const size_t wh = (size_t)ctxs[0]->width * ctxs[0]->height;
	for (int g = 0; g < n; g++) { CU(cudaSetDevice(ctxs[g]->device)); CU(cudaEventRecord(ctxs[g]->evRender0, ctxs[g]->stream)); }
	int path = 1;
	if (!peers || WantNccl()) {
		path = 2;
		rcode = NcclAllReduce(ctxs, n);
		if (rcode != AGPT_OK) return rcode;
	}
	else {
		PeerAccums pa;
		pa.n = n;
		for (int g = 0; g < n; g++) pa.p[g] = ctxs[g]->accum;
		for (int g = 0; g < n; g++) {
			if (root >= 0 && g != root) continue;
			int first = 0, count = (int)wh;
			if (root < 0) SliceOf(g, n, wh, &first, &count);
			CU(cudaSetDevice(ctxs[g]->device));
			if (root < 0) k_allreduce_slice<<<Blocks(count, 256), 256, 0, ctxs[g]->stream>>>(pa, first, count);
			else k_reduce_resolve<<<Blocks(count, 256), 256, 0, ctxs[g]->stream>>>(pa, first, count, 1.f, ctxs[g]->accum, nullptr);
			CU(cudaGetLastError());
			ctxs[g]->stats.kernel_launches++;
		}
	}
	for (int g = 0; g < n; g++) { CU(cudaSetDevice(ctxs[g]->device)); CU(cudaEventRecord(ctxs[g]->evRender1, ctxs[g]->stream)); }
	rcode = SyncGroup(ctxs, n);
	if (rcode != AGPT_OK) return rcode;
	for (int g = 0; g < n; g++) {
		float ms = 0;
		cudaEventElapsedTime(&ms, ctxs[g]->evRender0, ctxs[g]->evRender1);
		ctxs[g]->stats.ms_reduce += ms; ctxs[g]->stats.reduce_path = (uint32_t)path;
	}
	return AGPT_OK;
}

int agpt_allreduce_accum(agpt_ctx** ctxs, int n) { return agpt_reduce_accum(ctxs, n, -1); }

// Fused reduce + CopyToSurface: GPU g sums ITS slice of the film over all accumulators (rank order, peer
// loads) and packs it; the G slices go to the host over G PCIe links at once.  keep_sum != 0 also leaves
// the summed film in ctxs[0]'s accumulator (slices written there by peer stores).
int agpt_reduce_resolve(agpt_ctx** ctxs, int n, int samples, int keep_sum, uint32_t* host_rgb8) {
	int rcode = CheckGroup(ctxs, n);
	if (rcode != AGPT_OK) return rcode;
	NEED(samples > 0 && host_rgb8 != nullptr, AGPT_ERR_INVALID, "samples <= 0 or null output");
	bool peers = n == 1;
	if (n > 1) { rcode = EnablePeerAccess(ctxs, n, &peers); if (rcode != AGPT_OK) return rcode; }
	if (!peers || (n > 1 && WantNccl())) {
		// no peer memory between some pair: NCCL all-reduce, then the single-GPU resolve
		rcode = agpt_reduce_accum(ctxs, n, -1);
		if (rcode != AGPT_OK) return rcode;
		return agpt_resolve(ctxs[0], samples, host_rgb8);
	}
	rcode = SyncGroup(ctxs, n);
	if (rcode != AGPT_OK) return rcode;
	const size_t wh = (size_t)ctxs[0]->width * ctxs[0]->height;
	PeerAccums pa;
	pa.n = n;
	for (int g = 0; g < n; g++) pa.p[g] = ctxs[g]->accum;
	for (int g = 0; g < n; g++) {
		agpt_ctx* c = ctxs[g];
		int first, count;
		SliceOf(g, n, wh, &first, &count);
		CU(cudaSetDevice(c->device));
		if (c->resolved.n != wh) CU(c->resolved.Alloc(wh));
		CU(cudaEventRecord(c->evRender0, c->stream));
		// (the sum may be written over ctxs[0]'s own addend: pixel i is read and written by this one thread only)
		k_reduce_resolve<<<Blocks(count, 256), 256, 0, c->stream>>>(pa, first, count, (float)samples, keep_sum ? ctxs[0]->accum : nullptr, c->resolved.p);
		CU(cudaGetLastError());
		c->stats.kernel_launches++;
		CU(cudaEventRecord(c->evRender1, c->stream));
		CU(cudaMemcpyAsync(host_rgb8 + first, c->resolved.p + first, (size_t)count * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	}
	rcode = SyncGroup(ctxs, n);
	if (rcode != AGPT_OK) return rcode;
	for (int g = 0; g < n; g++) {
		float ms = 0;
		cudaEventElapsedTime(&ms, ctxs[g]->evRender0, ctxs[g]->evRender1);
		ctxs[g]->stats.ms_reduce += ms; ctxs[g]->stats.reduce_path = 1;
	}
	return AGPT_OK;
}

// num_samples Tick bodies split over the group: context g renders first_sample + g, + g + n, ... (one host
// thread per GPU: agpt_render paces its waves from the host).
int agpt_render_multi(agpt_ctx** ctxs, int n, int first_sample, int num_samples, int max_depth, int rr_depth_arg, uint32_t flags) {
	int rcode = CheckGroup(ctxs, n);
	if (rcode != AGPT_OK) return rcode;
	NEED(num_samples >= 0, AGPT_ERR_INVALID, "bad sample range");
	std::vector<int> rc((size_t)n, AGPT_OK);
	std::vector<std::string> msg((size_t)n);
	std::vector<std::thread> pool;
	for (int g = 0; g < n; g++) {
		int count = num_samples > g ? (num_samples - g + n - 1) / n : 0;
		pool.emplace_back([=, &rc, &msg] {
			if (count > 0) rc[g] = agpt_render(ctxs[g], first_sample + g, count, n, max_depth, rr_depth_arg, flags);
			if (rc[g] != AGPT_OK) msg[g] = g_error;         // (thread-local message of the worker)
		});
	}
	for (auto& t : pool) t.join();
	for (int g = 0; g < n; g++) if (rc[g] != AGPT_OK) return Fail(rc[g], "GPU " + std::to_string(ctxs[g]->device) + ": " + msg[g]);
	return AGPT_OK;
}

// ---- multi-process sharding (one rank per GPU: torchrun, MPI): peers' accumulators through CUDA IPC ----
int agpt_accum_ipc_handle(agpt_ctx* c, void* handle64) {
	JOIN(c);
	NEED(c != nullptr && handle64 != nullptr, AGPT_ERR_INVALID, "null argument");
	NEED(c->accum != nullptr && c->accum == c->accumOwn.p, AGPT_ERR_STATE, "the context must own its accumulator (agpt_set_film, no agpt_set_accum_dev)");
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
	CU(cudaSetDevice(c->device));
	cudaIpcMemHandle_t h;
	CU(cudaIpcGetMemHandle(&h, c->accumOwn.p));
	memcpy(handle64, &h, sizeof(h));
	return AGPT_OK;
}

int agpt_close_peer_accums(agpt_ctx* c) {
	NEED(c != nullptr, AGPT_ERR_INVALID, "null context");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	for (int r = 0; r < c->peerWorld; r++) if (r != c->peerRank && c->peerAccum[r]) cudaIpcCloseMemHandle(c->peerAccum[r]);
	memset(c->peerAccum, 0, sizeof(c->peerAccum));
	c->peerWorld = 0; c->peerRank = -1;
	return AGPT_OK;
}

int agpt_open_peer_accums(agpt_ctx* c, int rank, int world, const void* handles64) {
	JOIN(c);
	NEED(c != nullptr && handles64 != nullptr && world >= 1 && world <= AGPT_MAX_PEERS && rank >= 0 && rank < world, AGPT_ERR_INVALID, "bad rank / world / handles");
	NEED(c->accum != nullptr, AGPT_ERR_STATE, "film not set");
	int rcode = agpt_close_peer_accums(c);
	if (rcode != AGPT_OK) return rcode;
	for (int r = 0; r < world; r++) {
		if (r == rank) { c->peerAccum[r] = c->accum; continue; }
		cudaIpcMemHandle_t h;
		memcpy(&h, (const char*)handles64 + 64 * (size_t)r, sizeof(h));
		void* p = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) { c->peerWorld = r; c->peerRank = rank; agpt_close_peer_accums(c); return Fail(AGPT_ERR_CUDA, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e)); }
		c->peerAccum[r] = (float4*)p;
	}
	c->peerRank = rank; c->peerWorld = world;
	return AGPT_OK;
}

static int PeerTable(agpt_ctx* c, PeerAccums* pa) {
	NEED(c != nullptr && c->peerWorld >= 1, AGPT_ERR_STATE, "peer accumulators not open (agpt_open_peer_accums)");
	pa->n = c->peerWorld;
	for (int r = 0; r < c->peerWorld; r++) pa->p[r] = c->peerAccum[r];
	pa->p[c->peerRank] = c->accum;
	return AGPT_OK;
}

// This rank's slice of the all-reduce.  The CALLER provides the two barriers: every rank has finished
// rendering before any rank calls, and no rank reads its accumulator before every rank has returned.
int agpt_allreduce_accum_peers(agpt_ctx* c) {
	JOIN(c);
	PeerAccums pa;
	int rcode = PeerTable(c, &pa);
	if (rcode != AGPT_OK) return rcode;
	CU(cudaSetDevice(c->device));
	int first, count;
	SliceOf(c->peerRank, c->peerWorld, (size_t)c->width * c->height, &first, &count);
	CU(cudaEventRecord(c->evRender0, c->stream));
	k_allreduce_slice<<<Blocks(count, 256), 256, 0, c->stream>>>(pa, first, count);
	CU(cudaGetLastError());
	c->stats.kernel_launches++;
	CU(cudaEventRecord(c->evRender1, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	float ms = 0;
	cudaEventElapsedTime(&ms, c->evRender0, c->evRender1);
	c->stats.ms_reduce += ms; c->stats.reduce_path = 1;
	return AGPT_OK;
}

// The whole film on this rank (the root): sum of all ranks' accumulators in rank order -> packed pixels on the
// host; keep_sum != 0 also stores the sum in this rank's accumulator.  Reads the peers, writes none of them:
// only the barrier BEFORE the call is needed.
int agpt_reduce_resolve_peers(agpt_ctx* c, int samples, int keep_sum, uint32_t* host_rgb8) {
	JOIN(c);
	PeerAccums pa;
	int rcode = PeerTable(c, &pa);
	if (rcode != AGPT_OK) return rcode;
	NEED(samples > 0 && host_rgb8 != nullptr, AGPT_ERR_INVALID, "samples <= 0 or null output");
	CU(cudaSetDevice(c->device));
	const size_t wh = (size_t)c->width * c->height;
	if (c->resolved.n != wh) CU(c->resolved.Alloc(wh));
	CU(cudaEventRecord(c->evRender0, c->stream));
	k_reduce_resolve<<<Blocks(wh, 256), 256, 0, c->stream>>>(pa, 0, (int)wh, (float)samples, keep_sum ? c->accum : nullptr, c->resolved.p);
	CU(cudaGetLastError());
	c->stats.kernel_launches++;
	CU(cudaEventRecord(c->evRender1, c->stream));
	CU(cudaMemcpyAsync(host_rgb8, c->resolved.p, wh * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	float ms = 0;
	cudaEventElapsedTime(&ms, c->evRender0, c->evRender1);
	c->stats.ms_reduce += ms; c->stats.reduce_path = 1;
	return AGPT_OK;
}

// ---- pinned host memory for accumulators (fast H2D/D2H of the float4 film) --------------------
int agpt_host_alloc(size_t bytes, void** out) {
	NEED(out != nullptr && bytes > 0, AGPT_ERR_INVALID, "bad allocation request");
	*out = nullptr;
	CU(cudaMallocHost(out, bytes));
	return AGPT_OK;
}
int agpt_host_free(void* p) {
	if (p) CU(cudaFreeHost(p));
	return AGPT_OK;
}

// ---- observability -----------------------------------------------------------------------
int agpt_get_stats(agpt_ctx* c, agpt_stats* out) {
	NEED(c != nullptr && out != nullptr, AGPT_ERR_INVALID, "null argument");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	unsigned long long tc[2 * AGPT_WAVE_COUNTERS];
	RayCounters rc;
	CU(cudaMemcpy(tc, c->traceCounters.p, sizeof(tc), cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(&rc, c->rayCounters.p, sizeof(rc), cudaMemcpyDeviceToHost));
	*out = c->stats;
	for (int k = 0; k < 2; k++) {
		const unsigned long long* t = tc + AGPT_WAVE_COUNTERS * k;
		out->node_visits[k] = t[1]; out->box_tests[k] = t[2]; out->tri_tests[k] = t[3]; out->analytic_tests[k] = t[4];
		out->warp_steps[k] = t[5]; out->lane_steps[k] = t[6];
	}
	out->rays_closest += rc.rays_closest; out->rays_shadow = rc.rays_shadow; out->rays_mis = rc.rays_mis; out->rays_skip = rc.rays_skip; out->rays_mis_culled = rc.rays_mis_culled; out->rays_tail_culled = rc.rays_tail_culled;
	out->ms_other = out->ms_render - out->ms_trace_closest - out->ms_trace_any - out->ms_shade;
	return AGPT_OK;
}
int agpt_get_wave_stats(agpt_ctx* c, int kind, agpt_wave_stats* out, int max_waves, int* n_waves) {
	NEED(c != nullptr && out != nullptr && n_waves != nullptr && (kind == 0 || kind == 1) && max_waves >= 1, AGPT_ERR_INVALID, "bad argument");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	std::vector<unsigned long long> rows((size_t)AGPT_MAX_WAVE_ROWS * AGPT_WAVE_COUNTERS);
	CU(cudaMemcpy(rows.data(), c->traceCounters.p + 2 * AGPT_WAVE_COUNTERS + (size_t)kind * AGPT_MAX_WAVE_ROWS * AGPT_WAVE_COUNTERS, rows.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	int n = 0;
	for (int w = 0; w < AGPT_MAX_WAVE_ROWS && w < max_waves; w++) {
		const unsigned long long* r = &rows[(size_t)w * AGPT_WAVE_COUNTERS];
		out[w].rays = r[0]; out[w].node_visits = r[1]; out[w].box_tests = r[2]; out[w].tri_tests = r[3]; out[w].analytic_tests = r[4];
		out[w].warp_steps = r[5]; out[w].lane_steps = r[6]; out[w].reserved = 0;
		if (r[0]) n = w + 1;
	}
	*n_waves = n;
	return AGPT_OK;
}

int agpt_probe_bandwidth(agpt_ctx* c, size_t bytes, int iters, float* gbs_out) {
	NEED(c != nullptr && gbs_out != nullptr && bytes >= 4096 && iters >= 1, AGPT_ERR_INVALID, "bad probe arguments");
	CU(cudaSetDevice(c->device));
	DevBuf<uint4> buf;
	DevBuf<unsigned> sink;
	const size_t n16 = bytes / 16;
	CU(buf.Alloc(n16)); CU(sink.Alloc(1));
	CU(cudaMemsetAsync(buf.p, 0, n16 * 16, c->stream));
	const int blocks = c->smCount * 4;
	k_probe_bandwidth<<<blocks, 512, 0, c->stream>>>(buf.p, n16, 1, sink.p);          // warm-up: page in, fill L2
	CU(cudaEventRecord(c->evA, c->stream));
	k_probe_bandwidth<<<blocks, 512, 0, c->stream>>>(buf.p, n16, iters, sink.p);
	CU(cudaEventRecord(c->evB, c->stream));
	CU(cudaGetLastError());
	c->stats.kernel_launches += 2;
	CU(cudaStreamSynchronize(c->stream));
	float ms = 0;
	cudaEventElapsedTime(&ms, c->evA, c->evB);
	*gbs_out = ms > 0 ? (float)((double)n16 * 16 * iters / (ms * 1e-3) / 1e9) : 0.f;
	buf.Free(); sink.Free();
	return AGPT_OK;
}

int agpt_debug_status(agpt_ctx* c, uint64_t* out4) {
	NEED(c != nullptr && out4 != nullptr, AGPT_ERR_INVALID, "null argument");
	out4[0] = out4[1] = out4[2] = out4[3] = 0;
#ifdef AGPT_DEBUG
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	unsigned long long v[4];
	CU(cudaMemcpyFromSymbol(v, g_agptDebug, sizeof(v)));
	for (int k = 0; k < 4; k++) out4[k] = v[k];
#endif
	return AGPT_OK;
}
int agpt_reset_stats(agpt_ctx* c) {
	NEED(c != nullptr, AGPT_ERR_INVALID, "null context");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	memset(&c->stats, 0, sizeof(c->stats));
	CU(cudaMemset(c->traceCounters.p, 0, c->traceCounters.Bytes()));
	CU(cudaMemset(c->rayCounters.p, 0, c->rayCounters.Bytes()));
	return AGPT_OK;
}

} // extern "C"

// ---- per-function probes -------------------------------------------------------------------
__global__ void k_probe_bounds(int n, const float* boxes6, const float* rays7, int* out_hit, float* out_t) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float* b = boxes6 + 6 * i;
	const float* r = rays7 + 7 * i;
	float t = 0;
	bool h = BoundsIntersect(f3(b[0], b[1], b[2]), f3(b[3], b[4], b[5]), f3(r[0], r[1], r[2]), f3(r[3], r[4], r[5]), r[6], t);
	out_hit[i] = h ? 1 : 0;
	out_t[i] = h ? t : 0.f;
}

// BSDF::f / Pdf / Sample_f through the SAME functions k_shade_b runs (VertexBsdfInit, EvalLobes,
// FinishEval, SampleLobeDir, FinishSample), against golden vectors of the reference's own methods.
__global__ void k_probe_bsdf(int n, agpt_material mat, const float* in14, int skipSpecular, float* out12) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float* a = in14 + 14 * i;
	float3 dpdu = f3(a[0], a[1], a[2]), dpdv = f3(a[3], a[4], a[5]), wo = f3(a[6], a[7], a[8]), wi = f3(a[9], a[10], a[11]);
	float2 u = make_float2(a[12], a[13]);
	DSurface si;
	SurfaceInit(si, f3(0.f), dpdu, dpdv);
	VertexBsdf vb;
	VertexBsdfInit<true>(vb, si, &mat, wo);
	float* o = out12 + 12 * i;
	// f(wo, wi) and Pdf(wo, wi): zero when wo.z == 0 (reflection.h:117,178)
	float3 f = f3(0.f);
	float pdf = 0.f;
	LobeEval ev;
	if (vb.woOk) {
		EvalLobes<true>(vb, WorldToLocal(vb.b, wi), ev);
		f = FinishEval<true>(vb, ev, wi, &pdf);
	}
	o[0] = f.x; o[1] = f.y; o[2] = f.z; o[3] = pdf;
	// Sample_f(wo, &wi, u, &pdf, skipSpecular, &sampledSpecular)
	DirSample smp;
	SampleLobeDir<true>(vb, u, skipSpecular != 0, smp);
	float3 wis = f3(0.f), fs = f3(0.f);
	float pdfs = 0.f;
	if (smp.ok) {
		ev.f = f3(0.f); ev.pdfCos = 0.f; ev.pdfMicro = 0.f; ev.fT = f3(0.f); ev.pdfT = 0.f;
		if (smp.lobe != AGPT_LOBE_SPECULAR) EvalLobes<true>(vb, smp.wi, ev);
		wis = LocalToWorld(vb.b, smp.wi);
		fs = FinishSample<true>(vb, smp, ev, wis, &pdfs);
	}
	bool spec = vb.woOk && smp.matching > 0 && smp.lobe == AGPT_LOBE_SPECULAR;
	o[4] = wis.x; o[5] = wis.y; o[6] = wis.z; o[7] = fs.x; o[8] = fs.y; o[9] = fs.z; o[10] = pdfs; o[11] = spec ? 1.f : 0.f;
}

__global__ void k_probe_sphere_sample(int n, const float* in9, float* out8) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float* a = in9 + 9 * i;
	agpt_sphere s;
	s.center[0] = a[0]; s.center[1] = a[1]; s.center[2] = a[2]; s.r = a[3]; s.r2 = a[3] * a[3];
	float3 p, nn;
	float pdf = 0;
	SphereSampleFrom(s, f3(a[4], a[5], a[6]), make_float2(a[7], a[8]), &p, &nn, &pdf);
	float* o = out8 + 8 * i;
	o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = nn.x; o[4] = nn.y; o[5] = nn.z; o[6] = pdf;
	o[7] = SpherePdfFrom(s, f3(a[4], a[5], a[6]));
}

__global__ void k_probe_stream(uint32_t pixel, uint32_t sample, int k, float* out) {
	uint32_t s = StreamSeed(pixel, sample);
	for (int i = 0; i < k; i++) out[i] = RandomFloat(s);
}

template <typename TIn, typename TOut, typename Launch>
static int ProbeRun(agpt_ctx* c, const TIn* in, size_t nIn, TOut* out, size_t nOut, Launch launch) {
	DevBuf<TIn> din;
	DevBuf<TOut> dout;
	CU(din.Upload(in, nIn, c->stream));
	CU(dout.Alloc(nOut));
	launch(din.p, dout.p);
	CU(cudaGetLastError());
	c->stats.kernel_launches++;
	CU(cudaMemcpyAsync(out, dout.p, nOut * sizeof(TOut), cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	din.Free(); dout.Free();
	return AGPT_OK;
}

extern "C" {

int agpt_probe_bounds(agpt_ctx* c, int n, const float* boxes6, const float* rays7, int* out_hit, float* out_t) {
	NEED(c != nullptr && n > 0 && boxes6 && rays7 && out_hit && out_t, AGPT_ERR_INVALID, "bad probe arguments");
	CU(cudaSetDevice(c->device));
	DevBuf<float> db, dr, dt;
	DevBuf<int> dh;
	CU(db.Upload(boxes6, 6 * (size_t)n, c->stream)); CU(dr.Upload(rays7, 7 * (size_t)n, c->stream));
	CU(dh.Alloc(n)); CU(dt.Alloc(n));
	k_probe_bounds<<<Blocks(n, 128), 128, 0, c->stream>>>(n, db.p, dr.p, dh.p, dt.p);
	CU(cudaGetLastError());
	c->stats.kernel_launches++;
	CU(cudaMemcpyAsync(out_hit, dh.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CU(cudaMemcpyAsync(out_t, dt.p, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	db.Free(); dr.Free(); dt.Free(); dh.Free();
	return AGPT_OK;
}

int agpt_probe_bsdf(agpt_ctx* c, int n, const agpt_material* mat, const float* in14, int skip_specular, float* out12) {
	NEED(c != nullptr && n > 0 && mat && in14 && out12, AGPT_ERR_INVALID, "bad probe arguments");
	CU(cudaSetDevice(c->device));
	agpt_material m = *mat;
	return ProbeRun(c, in14, 14 * (size_t)n, out12, 12 * (size_t)n, [&](const float* din, float* dout) {
		k_probe_bsdf<<<Blocks(n, 128), 128, 0, c->stream>>>(n, m, din, skip_specular, dout);
	});
}

int agpt_probe_sphere_sample(agpt_ctx* c, int n, const float* in9, float* out8) {
	NEED(c != nullptr && n > 0 && in9 && out8, AGPT_ERR_INVALID, "bad probe arguments");
	CU(cudaSetDevice(c->device));
	return ProbeRun(c, in9, 9 * (size_t)n, out8, 8 * (size_t)n, [&](const float* din, float* dout) {
		k_probe_sphere_sample<<<Blocks(n, 128), 128, 0, c->stream>>>(n, din, dout);
	});
}

int agpt_probe_stream(agpt_ctx* c, uint32_t pixel_index, uint32_t sample, int k, float* out) {
	NEED(c != nullptr && k > 0 && out, AGPT_ERR_INVALID, "bad probe arguments");
	CU(cudaSetDevice(c->device));
	DevBuf<float> d;
	CU(d.Alloc(k));
	k_probe_stream<<<1, 1, 0, c->stream>>>(pixel_index, sample, k, d.p);
	CU(cudaGetLastError());
	c->stats.kernel_launches++;
	CU(cudaMemcpyAsync(out, d.p, k * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	d.Free();
	return AGPT_OK;
}

} // extern "C"
