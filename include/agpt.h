/* agpt.h -- C ABI of the B200-native path-tracing hot path (libagpt.so).
 *
 * The reference (voxel-tracer/ag-pathtracer) has no FFI or plugin layer: its hot path is
 * entered per ray through `virtual float3 Integrator::Li(const Ray&, const Scene&, int depth)`
 * (integrator.h:28-31) and per frame through the pixel loop of `MyApp::Tick`
 * (myapp.cpp:163-175) storing through `Accumulator::AddSample` (myapp.h:17-19).  This header
 * is the boundary the drop-in inserts there: host C++ (the reference's own classes, or the
 * mirror in ag-pathtracer_b200/host/) flattens a Scene once and hands (scene, camera,
 * sample range) to the device; the device hands an accumulator back.
 *
 * Conventions: plain C types only; every function returns 0 on success or a negative
 * agpt_status, with a message in agpt_last_error(); no exceptions cross the boundary; one
 * context per GPU; functions are thread-compatible per context (not thread-safe); all
 * pointers are HOST pointers to caller-owned memory that may be freed when the call
 * returns, unless the name says `_dev`.  There is no CPU fallback: without a usable CUDA
 * device agpt_create() fails and nothing else can be called.
 */
#ifndef AGPT_H
#define AGPT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct agpt_ctx agpt_ctx;

typedef enum agpt_status {
	AGPT_OK = 0,
	AGPT_ERR_INVALID = -1,   /* bad argument or inconsistent scene tables */
	AGPT_ERR_CUDA = -2,      /* CUDA runtime error (message carries cudaGetErrorString) */
	AGPT_ERR_STATE = -3,     /* call order: film / camera / scene not set yet */
	AGPT_ERR_NOMEM = -4
} agpt_status;

/* ---- scene tables ------------------------------------------------------------------ */

/* Verbatim BVHNode of the reference (bvhtrimesh.h:126-130): 32 bytes = 2 x 128-bit loads.
 * count > 0: leaf, `first` = offset of its first triangle in the mesh's leaf-ordered
 * triangle array; count == 0: interior, children at nodes[first] and nodes[first+1]
 * (one 64-byte line, bvhtrimesh.h:175).  Root at 0, slot 1 unused. */
typedef struct agpt_bvh_node {
	float bmin[3];
	float bmax[3];
	int32_t first;
	int32_t count;
} agpt_bvh_node;

/* One TriangleMesh / BVHTriMesh.  Triangles are LEAF-ORDERED: slot j holds the triangle
 * that `primitives[j]` refers to in the reference (bvhtrimesh.h:339), so a leaf reads
 * tri_verts[first .. first+count) directly instead of chasing
 * primitives[].index -> indices[].vertex_index -> vertices[].
 * n_nodes == 0: plain TriangleMesh, tested brute force in slot order with no bounds test
 * (trianglemesh.h:25-35). */
typedef struct agpt_mesh_desc {
	const agpt_bvh_node* nodes;  /* n_nodes entries, or NULL */
	int32_t n_nodes;
	int32_t n_tris;
	const float* tri_verts;      /* n_tris x 3 x float4: v0 v1 v2; .w unused */
	const int32_t* tri_ids;      /* n_tris: original triangle number (indices[3*id..]) reported in agpt_hit.tri */
	const float* tri_normals;    /* n_tris x 3 x float4 (n0 n1 n2) or NULL: mesh has no normals (trianglemesh.cpp:86) */
	const float* tri_uvs;        /* n_tris x 3 x float2 or NULL: default uvs (0,0),(1,0),(1,1) (trianglemesh.cpp:52-56) */
} agpt_mesh_desc;

typedef struct agpt_sphere {     /* intersectable.h:159-162,319-321 */
	float center[3];
	float r;
	float r2;                    /* radius*radius as the Sphere ctor rounds it */
	float pad[3];
} agpt_sphere;

typedef struct agpt_plane {      /* intersectable.h:119-121,154-156: XZ plane through o, normal +y */
	float o[3];
	float half_x;
	float half_z;
	float pad[3];
} agpt_plane;

enum { AGPT_PRIM_SPHERE = 0, AGPT_PRIM_PLANE = 1, AGPT_PRIM_BVH_MESH = 2, AGPT_PRIM_MESH = 3,
       AGPT_PRIM_INSTANCE = 4 /* EXTENSION: a placed mesh, payload = index into the instance table */ };

/* EXTENSION (BASELINE config 4 says "instanced"): one placement of a mesh.  The reference has no
 * transforms at trace time (scene.h:5-28: a flat list, geometry baked); what an instance means is therefore
 * defined here and stated on the CPU in oracle/agpt_oracle.cpp (parity is against that, not against the reference):
 *   - the ray is taken to object space, O' = W2O * (O,1), D' = W2O * (D,0), D' NOT re-normalised, so the ray
 *     parameter t is the same in both spaces and ray.t is shared with the rest of the scene;
 *   - BVHTriMesh::Intersect / IntersectP run unchanged on (O', D', t) over the shared mesh;
 *   - the hit point is O + t*D (world ray); the triangle's vertices go to world space through O2W before the
 *     partial derivatives are formed, the interpolated shading normal goes through transpose(W2O).
 * Matrices are 3x4 row-major affine, rows of [R | t]; products are summed left to right ((m0*x + m1*y) + m2*z) + m3. */
typedef struct agpt_instance {
	int32_t mesh;                  /* index into the mesh table (shared by any number of instances) */
	int32_t pad[3];
	float object_to_world[12];
	float world_to_object[12];
} agpt_instance;

/* Scene::primitives in list order (scene.h:5-13,27): order decides exact-t ties. */
typedef struct agpt_prim {
	int32_t type;        /* AGPT_PRIM_* */
	int32_t payload;     /* index into the spheres / planes / meshes table of that type */
	int32_t material;    /* index into materials, -1 = nullptr material (emissive shape, integrator.h:152-161) */
	int32_t area_light;  /* index into lights of the AreaLight wrapping this shape, -1 = none; that light's
	                        `prim` must be this row (one shape per AreaLight, as upstream), else AGPT_ERR_INVALID */
} agpt_prim;

enum { AGPT_MAT_DISNEY = 1, AGPT_MAT_MIRROR = 2, AGPT_MAT_GLASS = 3 /* EXTENSION: rough dielectric */ };
enum { AGPT_LOBE_DIFFUSE = 1, AGPT_LOBE_RETRO = 2, AGPT_LOBE_MICROFACET = 4, AGPT_LOBE_SPECULAR = 8,
       /* EXTENSION (BASELINE config 5 says "rough-glass"): a rough dielectric interface as PBRT-v3's GlassMaterial
        * builds it.  GLASS_REFLECT is the reference's own MicrofacetReflection (reflection.h:38-78) over its plain
        * TrowbridgeReitzDistribution (G = 1/(1 + Lambda(wo) + Lambda(wi)), microfacet.h:103-105) and its FresnelDielectric(1, eta)
        * (microfacet.h:220-228).  GLASS_TRANSMIT -- MicrofacetTransmission (PBRT-v3 reflection.cpp, radiance transport) -- and the
        * rule that a transmission lobe contributes to BSDF::f when wi and wo lie on opposite sides of the geometric normal do not
        * exist upstream (its BSDF::f only sums when both lie on the same side, reflection.h:114-123): parity for those is against
        * oracle/agpt_oracle.cpp only.  bxdfs[] order: GLASS_REFLECT, GLASS_TRANSMIT. */
       AGPT_LOBE_GLASS_REFLECT = 16, AGPT_LOBE_GLASS_TRANSMIT = 32 };

/* Constants the reference's material constructors derive once (material.h:14-49,74-77). */
typedef struct agpt_material {
	int32_t type;         /* AGPT_MAT_* */
	uint32_t lobes;       /* AGPT_LOBE_* in BSDF::bxdfs[] order: diffuse, retro, microfacet | specular */
	float roughness;      /* DisneyRetro::roughness */
	float metallic;       /* DisneyFresnel::metallic */
	float diffuse_r[3];   /* (1-metallic)*color: R of DisneyDiffuse and DisneyRetro; GLASS: T of the transmission lobe */
	float eta;            /* 1.5 (material.h:65) */
	float spec_r0[3];     /* DisneyFresnel::R0 = Lerp(metallic, SchlickR0FromEta(eta), color) */
	float alpha_x;        /* max(.001, roughness^2) */
	float mirror_r[3];    /* SpecularReflection::R; GLASS: R of the reflection lobe */
	float alpha_y;
} agpt_material;

enum { AGPT_LIGHT_AREA = 0, AGPT_LIGHT_UNIFORM_INFINITE = 1, AGPT_LIGHT_INFINITE_AREA = 2 };

/* Scene::lights in list order (lights.h:37-51,72-87). */
typedef struct agpt_light {
	int32_t type;        /* AGPT_LIGHT_* */
	int32_t prim;        /* AREA: scene primitive index of the emitting shape; else -1 */
	float pad[2];
	float lemit[3];
	float pad2;
} agpt_light;

/* InfiniteAreaLight (lights.h:52-70, lights.cpp:31-112): lat-long HDR map (HDRTexture,
 * texture.h:41-84) plus the piecewise-constant distribution over its texels (Distribution1D,
 * sampling.h:20-69) that the light's constructor builds (lights.cpp:33-47).  One per scene. */
typedef struct agpt_envmap {
	int32_t width, height;
	const float* rgb;        /* width*height x 3 floats, row-major, as stbi_loadf decodes the .hdr */
	const float* func;       /* width*height: max(r,g,b) * sin(theta_row) */
	const float* cdf;        /* width*height + 1 */
	float func_int;          /* Distribution1D::funcInt */
} agpt_envmap;

/* Derived camera vectors exactly as Camera::updateCoords leaves them (camera.h:77-90);
 * computed on the host (tan, double->float), the device only adds and multiplies. */
typedef struct agpt_camera {
	float origin[3];
	float lower_left_corner[3];
	float horizontal[3];
	float vertical[3];
	float u[3];
	float v[3];
	float lens_radius;
} agpt_camera;

typedef struct agpt_hit {
	uint32_t found;
	int32_t prim;        /* index in Scene::primitives */
	int32_t tri;         /* original triangle number inside that mesh, -1 for sphere / plane */
	float t;
} agpt_hit;

typedef struct agpt_stats {
	uint64_t paths;            /* camera paths started */
	uint64_t rays_closest;     /* path rays through Scene::Intersect (integrator.h:136), skip-through rays included */
	uint64_t rays_shadow;      /* any-hit visibility rays (integrator.h:50) */
	uint64_t rays_mis;         /* closest-hit MIS rays (integrator.h:79) */
	uint64_t rays_skip;        /* of rays_closest: continuation through null-material shapes (integrator.h:158) */
	uint64_t rays_mis_culled;  /* MIS rays the reference traces but that provably cannot reach their light: not traced, not counted above.
	                            * Exact in renders made with AGPT_FLAG_COUNTERS; otherwise an upper bound (such samples are dropped before their BSDF
	                            * value is known, and the reference does not trace a sample whose value is black) */
	uint64_t rays_tail_culled; /* path rays the reference traces at bounces == MaxDepth and then discards (integrator.h:139,150): not traced, not counted above */
	/* traversal work, only counted with AGPT_FLAG_COUNTERS; [0] closest-hit kernel, [1] any-hit kernel */
	uint64_t node_visits[2];   /* interior sibling pairs fetched (64 B each) */
	uint64_t box_tests[2];
	uint64_t tri_tests[2];     /* triangles fetched and tested (48 B each) */
	uint64_t analytic_tests[2];/* sphere / plane records tested (32 B each) */
	uint64_t warp_steps[2];    /* trips of the lockstep BVH walk, counted once per warp */
	uint64_t lane_steps[2];    /* lanes that had work in those trips: lane_steps / warp_steps = rays alive per step (of 32) */
	uint64_t kernel_launches;  /* launches of this library's kernels since agpt_reset_stats */
	uint64_t launches_closest; /* of which closest-hit trace, any-hit trace, shade */
	uint64_t launches_any;
	uint64_t launches_shade;
	uint64_t waves;            /* wavefront iterations */
	float ms_render;           /* device time of agpt_render calls (CUDA events on the context stream) */
	float ms_trace_closest;    /* per kernel class, only with AGPT_FLAG_TIMING */
	float ms_trace_any;
	float ms_shade;
	float ms_other;            /* ms_render minus the three above (generate, accumulate, queue bookkeeping) */
	float ms_reduce;           /* device time of this context's share of agpt_reduce_* / agpt_allreduce_* calls */
	uint32_t reduce_path;      /* how the last one ran: 1 = this library's peer-memory kernels, 2 = ncclAllReduce */
} agpt_stats;

/* agpt_render / agpt_trace_* flags */
enum {
	AGPT_FLAG_COUNTERS = 1u,      /* count node visits / box / triangle tests (slower kernels) */
	AGPT_FLAG_TIMING = 2u,        /* bracket every kernel class with CUDA events (serialises) */
	AGPT_FLAG_STRICT_BOXES = 4u,  /* every slab test with the reference's six IEEE divisions; the default is the
	                                 exact-filtered test (same decisions, divisions only inside a guard band) */
	AGPT_FLAG_RAYS_FINAL = 8u,    /* agpt_trace_rays / agpt_li_rays: directions are used as given -- the rays come from host
	                                 Ray objects, whose constructor has normalised them already (camera.h:7) */
	AGPT_FLAG_RR_BY_BOUNCE = 16u  /* EXTENSION (not reference behaviour): Russian roulette keyed on the path's own
	                                 bounce index -- live once bounces > 3, the rule BASELINE config 5 names ("16 bounces with Russian
	                                 roulette") -- instead of on Li's constant depth argument (integrator.h:180, dead for depth = 0).
	                                 rr_depth_arg is ignored.  Checked against oracle/agpt_oracle.cpp, the only statement of it. */
};

/* ---- lifecycle --------------------------------------------------------------------- */

int agpt_create(int device, agpt_ctx** out);                 /* fails loudly without CUDA */
int agpt_destroy(agpt_ctx* ctx);
const char* agpt_last_error(void);
int agpt_device_count(int* out);
/* Launch on the caller's CUDA stream (cudaStream_t as void*); NULL = the context's own. */
int agpt_set_stream(agpt_ctx* ctx, void* cuda_stream);

/* ---- scene upload (replaces the Scene the reference keeps in host memory, scene.h:27-29) */

int agpt_upload_meshes(agpt_ctx* ctx, const agpt_mesh_desc* meshes, int n);
int agpt_upload_spheres(agpt_ctx* ctx, const agpt_sphere* spheres, int n);
int agpt_upload_planes(agpt_ctx* ctx, const agpt_plane* planes, int n);
int agpt_upload_primitives(agpt_ctx* ctx, const agpt_prim* prims, int n);
int agpt_upload_instances(agpt_ctx* ctx, const agpt_instance* instances, int n);  /* EXTENSION; rows of AGPT_PRIM_INSTANCE point here */
int agpt_upload_materials(agpt_ctx* ctx, const agpt_material* materials, int n);
int agpt_upload_lights(agpt_ctx* ctx, const agpt_light* lights, int n);
int agpt_upload_envmap(agpt_ctx* ctx, const agpt_envmap* env);          /* NULL clears it */
int agpt_set_camera(agpt_ctx* ctx, const agpt_camera* camera);          /* Camera, camera.h:38-56 */
int agpt_set_film(agpt_ctx* ctx, int width, int height);                /* Accumulator(w,h), myapp.h:10-13 */
int agpt_scene_bytes(agpt_ctx* ctx, uint64_t* out);                     /* resident scene bytes in HBM */

/* ---- the hot path -------------------------------------------------------------------- */

/* num_samples iterations of the MyApp::Tick pixel loop (myapp.cpp:163-175) for sample
 * indices [first_sample, first_sample+num_samples) stepping by sample_stride (1 on one GPU,
 * G when G GPUs split the samples s = g mod G), PathTracer(max_depth)::Li(ray, scene,
 * rr_depth_arg) per pixel, accumulated into the device float4[W*H] accumulator. */
int agpt_render(agpt_ctx* ctx, int first_sample, int num_samples, int sample_stride,
		int max_depth, int rr_depth_arg, uint32_t flags);

/* Hit table of the camera rays of one sample index: out_host[y*W + x] (parity gate). */
int agpt_trace_primary(agpt_ctx* ctx, int sample, uint32_t flags, agpt_hit* out_host);

/* Scene::Intersect (any_hit = 0) / Scene::IntersectP (any_hit != 0) for caller rays:
 * rays7 = O.xyz, D.xyz (normalised on the device like the Ray ctor, camera.h:7, unless AGPT_FLAG_RAYS_FINAL), tmax. */
int agpt_trace_rays(agpt_ctx* ctx, int64_t n, const float* rays7, int any_hit, uint32_t flags, agpt_hit* out_host);

/* Radiance of single camera paths without accumulation (Integrator::Li for the debug
 * click, myapp.cpp:196-198): pixel (xs[i], ys[i]), sample ss[i] -> out_rgb[3*i..]. */
int agpt_li_pixels(agpt_ctx* ctx, int n, const int* xs, const int* ys, const int* ss,
		int max_depth, int rr_depth_arg, uint32_t flags, float* out_rgb);

/* Integrator::Li(ray, scene, depth) for caller-supplied rays (integrator.h:28-31): rays7 as
 * in agpt_trace_rays, rng_states[i] = xorshift32 state the path starts from (the reference
 * draws from its one global generator instead). */
int agpt_li_rays(agpt_ctx* ctx, int n, const float* rays7, const uint32_t* rng_states,
		int max_depth, int rr_depth_arg, uint32_t flags, float* out_rgb);

/* ---- accumulator (Accumulator, myapp.h:8-68) ----------------------------------------- */

int agpt_clear(agpt_ctx* ctx);                                  /* Accumulator::Clear */
int agpt_accum_ptr_dev(agpt_ctx* ctx, void** dev_ptr);          /* float4[W*H] in HBM, row (H-1-y): for NCCL */
int agpt_set_accum_dev(agpt_ctx* ctx, void* dev_ptr);           /* accumulate into caller-owned device memory */
int agpt_read_accum(agpt_ctx* ctx, float* host_rgba);           /* D2H of float4[W*H] */
int agpt_write_accum(agpt_ctx* ctx, const float* host_rgba);    /* H2D (resume) */
/* The same upload, started and not waited for: it runs beside the agpt_render that follows (the film is only needed when a
 * batch is added to it, at the end) -- the batched Tick of CudaPathTracer::Render hides its H2D copy this way.  The host buffer
 * must stay untouched until the next call on this context that reads or writes the film returns (agpt_render, agpt_read_accum,
 * agpt_resolve, the reduce calls ...: each waits for the upload first).  Page-locked memory (agpt_host_alloc) for a true overlap. */
int agpt_write_accum_begin(agpt_ctx* ctx, const float* host_rgba);
/* Accumulator::CopyToSurface (myapp.h:34-41): /samples, pow(1/2.2), 8-bit pack 0x00RRGGBB. */
int agpt_resolve(agpt_ctx* ctx, int samples, uint32_t* host_rgb8);

/* ---- multi-GPU: sample-index sharding (north star "NCCL reduces the accumulators over NVLink") ----
 * One context per GPU of ONE process, each holding the whole scene and a film of the same size.  GPU g of G
 * renders the samples s = first + g (mod G) into its own accumulator (agpt_render with sample_stride = G, or
 * agpt_render_multi, which runs the G renders on G host threads); the frame is the sum of the accumulators.
 *
 * agpt_reduce_accum sums them with this library's own kernel over NVLink peer memory -- each GPU owns a slice
 * of the film, loads it from all accumulators, adds in RANK ORDER and stores the sum back to all (root = -1) or
 * the root's GPU does it for the whole film (root >= 0: only ctxs[root] receives the sum).  Rank-order summation
 * makes the result reproducible and equal to ((a0 + a1) + a2) + ... .  Where some pair of GPUs has no peer access,
 * or with AGPT_REDUCE=nccl in the environment, it runs ncclAllReduce instead (libnccl is dlopen'ed on first use).
 * agpt_allreduce_accum(ctxs, n) == agpt_reduce_accum(ctxs, n, -1).
 *
 * agpt_reduce_resolve is the fused end of a render: sum + Accumulator::CopyToSurface (myapp.h:34-41) in ONE
 * kernel per GPU (its slice of the film), packed pixels straight to the host over all GPUs' PCIe links;
 * keep_sum != 0 also leaves the summed film in ctxs[0]'s accumulator. */
int agpt_render_multi(agpt_ctx** ctxs, int n, int first_sample, int num_samples, int max_depth, int rr_depth_arg, uint32_t flags);
int agpt_reduce_accum(agpt_ctx** ctxs, int n, int root);
int agpt_allreduce_accum(agpt_ctx** ctxs, int n);
int agpt_reduce_resolve(agpt_ctx** ctxs, int n, int samples, int keep_sum, uint32_t* host_rgb8);

/* The same exchange between PROCESSES (one rank per GPU: torchrun, MPI).  Each rank exports its context-owned
 * accumulator as a 64-byte CUDA IPC handle; after the ranks have exchanged the handles (any transport), each
 * opens the others' accumulators and the same peer-memory kernels run on them.  The caller provides the
 * barriers: before either call every rank must have finished rendering; after agpt_allreduce_accum_peers no
 * rank may touch its accumulator until every rank has returned.  agpt_reduce_resolve_peers runs on the root
 * rank only and writes nothing to the peers. */
int agpt_accum_ipc_handle(agpt_ctx* ctx, void* handle64);
int agpt_open_peer_accums(agpt_ctx* ctx, int rank, int world, const void* handles64 /* world x 64 bytes, rank order */);
int agpt_close_peer_accums(agpt_ctx* ctx);
int agpt_allreduce_accum_peers(agpt_ctx* ctx);
int agpt_reduce_resolve_peers(agpt_ctx* ctx, int samples, int keep_sum, uint32_t* host_rgb8);

/* Page-locked host memory for accumulator buffers (what MALLOC64 is upstream, myapp.h:11):
 * agpt_read_accum / agpt_write_accum on such a buffer run at PCIe speed. */
int agpt_host_alloc(size_t bytes, void** out);
int agpt_host_free(void* p);

/* ---- observability ------------------------------------------------------------------- */

/* Traversal work per WAVE of the renders made with AGPT_FLAG_COUNTERS since agpt_reset_stats.  Wave w of a batch
 * traces the path rays of bounce w (and the shadow / MIS rays of the vertex before it), so this is the per-bounce
 * N_int / N_tri / divergence report of SURVEY 8d (RecursiveHit, bvhtrimesh.h:332-384).  kind 0 = closest-hit kernel
 * (path + MIS rays), 1 = any-hit kernel (shadow rays).  Fills up to max_waves rows; *n_waves = rows in use. */
typedef struct agpt_wave_stats {
	uint64_t rays, node_visits, box_tests, tri_tests, analytic_tests, warp_steps, lane_steps, reserved;
} agpt_wave_stats;
int agpt_get_wave_stats(agpt_ctx* ctx, int kind, agpt_wave_stats* out, int max_waves, int* n_waves);
/* Read bandwidth (GB/s) a plain streaming kernel of this library reaches on this GPU over `bytes` of device memory
 * read `iters` times: bytes well inside the L2 size measures L2, bytes far beyond it HBM (roofline denominators). */
int agpt_probe_bandwidth(agpt_ctx* ctx, size_t bytes, int iters, float* gbs_out);
int agpt_get_stats(agpt_ctx* ctx, agpt_stats* out);
int agpt_reset_stats(agpt_ctx* ctx);
/* Builds with -DAGPT_DEBUG (libagpt_debug.so) check stack depth, node / triangle / primitive indices and queue
 * slots inside the kernels.  out4 = { failed checks, code of the first failure, its value, checks executed }
 * since the library was loaded; a release build reports 0 checks executed. */
int agpt_debug_status(agpt_ctx* ctx, uint64_t* out4);

/* ---- per-function probes (differential tests against single reference functions) ----- */

/* Bounds::Intersect (bvhtrimesh.h:18-36): boxes6 = bmin,bmax; rays7 = O,D (as given),tmax. */
int agpt_probe_bounds(agpt_ctx* ctx, int n, const float* boxes6, const float* rays7, int* out_hit, float* out_t);
/* BSDF::f / Pdf / Sample_f on a frame built from (dpdu, dpdv) (reflection.h:114-188):
 * in14 = dpdu3 dpdv3 wo3 wi3 u2; out12 = f3 pdf1 | sampled wi3 f3 pdf1 specular1. */
int agpt_probe_bsdf(agpt_ctx* ctx, int n, const agpt_material* mat, const float* in14, int skip_specular, float* out12);
/* Sphere::Sample(ref,u) / Sphere::Pdf (intersectable.h:239-317): in9 = center3 r refp3 u2;
 * out8 = p3 n3 pdf pdf_only. */
int agpt_probe_sphere_sample(agpt_ctx* ctx, int n, const float* in9, float* out8);
/* First k floats of the RNG stream of (pixel_index, sample). */
int agpt_probe_stream(agpt_ctx* ctx, uint32_t pixel_index, uint32_t sample, int k, float* out);

#ifdef __cplusplus
}
#endif
#endif /* AGPT_H */
