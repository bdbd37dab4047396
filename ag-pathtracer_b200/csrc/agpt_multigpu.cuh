// agpt_multigpu.cuh -- the exchange step of sample-index sharding as this library's own
// kernels over NVLink peer memory, plus the NCCL route.
//
// GPU g of G renders the samples s = g (mod G) of every pixel into its own float4[W*H]
// accumulator; the frame is the sum of the G accumulators, then Accumulator::CopyToSurface
// (myapp.h:34-41).  Two peer-memory kernels do that here:
//
//   k_allreduce_slice   GPU g owns the g-th slice of the film: it loads that slice from every
//                       GPU's accumulator (its own from HBM, the others as NVLink P2P loads), adds
//                       them IN RANK ORDER and stores the sum back into all G accumulators (P2P
//                       stores).  Nobody else touches those pixels, so the G kernels need no
//                       synchronisation with one another -- only "all renders done" before and
//                       "all kernels done" after, which the host provides.  Each GPU moves
//                       2 (G-1)/G of a film over NVLink, like a ring all-reduce, in one launch.
//   k_reduce_resolve    the fused end of a render: sum in rank order + /samples + gamma + 8-bit pack
//                       in one pass, without materialising the summed film: reduce (NCCL's job in the
//                       north star) and CopyToSurface (the reference's display step) in one kernel.
//
// Summation in rank order makes the result reproducible run to run and equal to
// ((a0 + a1) + a2) + ... computed anywhere -- NCCL's ring / tree / NVLS orders are not.
// The NCCL route (ncclAllReduce through a dlopen'ed libnccl, the library torch already loaded if
// there is one) is kept for pairs of GPUs without peer access and for A/B measurements.
#pragma once

#include <dlfcn.h>

#include "agpt_kernels.cuh"

#define AGPT_MAX_PEERS 16

struct PeerAccums {
	float4* p[AGPT_MAX_PEERS];     // accumulators in rank order, as addresses valid on the launching GPU
	int n;
};

// pixels [first, first + count): sum over ranks in order, write the sum to every rank's accumulator
__global__ void __launch_bounds__(256) k_allreduce_slice(PeerAccums acc, int first, int count) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	i += first;
	float4 s = acc.p[0][i];
	for (int r = 1; r < acc.n; r++) { float4 v = acc.p[r][i]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
	for (int r = 0; r < acc.n; r++) acc.p[r][i] = s;
}

// pixels [first, first + count): sum over ranks in order -> optional sumOut (this GPU's memory) -> packed pixel
__global__ void __launch_bounds__(256) k_reduce_resolve(PeerAccums acc, int first, int count, float samples, float4* sumOut, uint32_t* out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	i += first;
	float4 s = acc.p[0][i];
	for (int r = 1; r < acc.n; r++) { float4 v = acc.p[r][i]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
	if (sumOut) sumOut[i] = s;
	if (out) out[i] = ResolvePixel(s, samples);
}

// ---- NCCL through dlopen: no link-time dependency, and one libnccl per process -------------------
typedef struct ncclComm* agpt_ncclComm_t;
struct NcclApi {
	void* lib = nullptr;
	int (*CommInitAll)(agpt_ncclComm_t*, int, const int*) = nullptr;
	int (*CommDestroy)(agpt_ncclComm_t) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, agpt_ncclComm_t, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
	std::string error;
	bool Load() {
		if (lib) return true;
		// a libnccl already in the process (torch's) first; then the system's
		const char* names[] = { "libnccl.so.2", "libnccl.so" };
		for (const char* nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD); if (lib) break; }
		for (const char* nm : names) { if (lib) break; lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); }
		if (!lib) { error = std::string("libnccl not found: ") + dlerror(); return false; }
		CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
		CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
		AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
		GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
		GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
		GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
		if (!CommInitAll || !CommDestroy || !AllReduce || !GroupStart || !GroupEnd || !GetErrorString) { error = "libnccl lacks an expected symbol"; lib = nullptr; return false; }
		return true;
	}
};
