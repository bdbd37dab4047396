// host/bvhtrimesh.h -- BVHTriMesh: binned-SAH BVH over a triangle mesh, built on the host
// into the flat array the device traverses.
//
// Counterpart of /root/reference/bvhtrimesh.h:152-330.  The north star keeps the build on
// the host "with the same topology": this builder makes the decisions of the reference's
// BuildRecursive (:213-310) -- 12 centroid buckets on the longest centroid axis, cost
// 1 + (n0*A0 + n1*A1)/A, leaf iff n <= maxPrimsInNode and minCost >= n, median split by
// nth_element for n <= 2, leaf when all centroids coincide -- with the same float
// arithmetic and the same libstdc++ partition / nth_element, so node for node and slot for
// slot it equals what upstream builds (tests compare the arrays byte for byte with the
// oracle's export).  Unlike upstream it does not go through a shared_ptr build tree: nodes
// are written straight into the flattened layout of FlattenBVHTree (:312-330): root at 0,
// slot 1 unused, each sibling pair at an even index (one 64-byte line), depth-first with
// the left subtree first.  Interior bounds are the union of the children's, which for
// min/max is exactly the bounds of the range.
//
// Traversal (RecursiveHit / RecursiveHitP, :332-413) is device code; see
// csrc/agpt_trace.cuh.
#pragma once

#include "precomp.h"
#include "trianglemesh.h"

#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <unistd.h>

struct alignas(32) BVHNode {      // == agpt_bvh_node == upstream BVHNode (bvhtrimesh.h:126-130)
	float bmin3[3];
	float bmax3[3];
	int first;
	int count;
};
static_assert(sizeof(BVHNode) == sizeof(agpt_bvh_node), "BVHNode must stay 32 bytes");

class BVHTriMesh : public TriangleMesh {
public:
	BVHTriMesh(shared_ptr<TriangleMesh> trimesh, std::shared_ptr<Material> mat, int maxPrimsInNode = 1)
		: TriangleMesh(trimesh, mat) {
		Build(maxPrimsInNode);
	}

	// Process-wide build settings.  Defaults come from the environment: AGPT_BUILD_THREADS
	// (else the hardware concurrency, at most 16) and AGPT_BVH_CACHE_DIR (else no cache).
	struct BuildOptions {
		int threads = 1;
		std::string cacheDir;
	};
	static BuildOptions& Options() {
		static BuildOptions opt = [] {
			BuildOptions o;
			unsigned hc = std::thread::hardware_concurrency();
			o.threads = hc == 0 ? 1 : (hc > 16 ? 16 : (int)hc);
			if (const char* e = getenv("AGPT_BUILD_THREADS")) { int t = atoi(e); if (t >= 1) o.threads = t; }
			if (const char* e = getenv("AGPT_BVH_CACHE_DIR")) o.cacheDir = e;
			return o;
		}();
		return opt;
	}

	int Kind() const override { return AGPT_PRIM_BVH_MESH; }
	const std::vector<BVHNode>& Nodes() const { return nodes; }
	// leafOrder[j] = original triangle number held by leaf slot j (upstream primitives[j].index / 3)
	const std::vector<int32_t>& LeafOrder() const { return leafOrder; }

private:
	struct Box {
		float lo[3] = { 1e34f, 1e34f, 1e34f };
		float hi[3] = { -1e34f, -1e34f, -1e34f };
		void Grow(const float* p) {
			for (int a = 0; a < 3; a++) { lo[a] = lo[a] < p[a] ? lo[a] : p[a]; hi[a] = hi[a] > p[a] ? hi[a] : p[a]; }
		}
		void Grow(const Box& b) {
			for (int a = 0; a < 3; a++) { lo[a] = lo[a] < b.lo[a] ? lo[a] : b.lo[a]; hi[a] = hi[a] > b.hi[a] ? hi[a] : b.hi[a]; }
		}
		float Extent(int a) const { return hi[a] - lo[a]; }
		int LongestAxis() const {
			int a = 0;
			if (Extent(1) > Extent(0)) a = 1;
			if (Extent(2) > Extent(a)) a = 2;
			return a;
		}
		float SurfaceArea() const {
			float dx = Extent(0), dy = Extent(1), dz = Extent(2);
			return 2 * (dx * dy + dx * dz + dy * dz);
		}
		// position of p[axis] inside the box as a fraction (Bounds::Offset, bvhtrimesh.h:76-82)
		float Offset(const float* p, int axis) const {
			float o = p[axis] - lo[axis];
			if (hi[axis] > lo[axis]) o /= hi[axis] - lo[axis];
			return o;
		}
	};
	struct BuildPrim {
		int tri;            // original triangle number
		Box box;
		float centroid[3];
	};

	static constexpr int kBuckets = 12;

	static int BucketOf(const Box& centroidBox, const BuildPrim& p, int axis) {
		int b = (int)(kBuckets * centroidBox.Offset(p.centroid, axis));
		return b == kBuckets ? kBuckets - 1 : b;
	}

	// ---- build ---------------------------------------------------------------------------
	// SURVEY 8f row 3 (the step before the path): upstream's BuildRecursive + FlattenBVHTree cost
	// ~15 s and ~3.5 GB of shared_ptr nodes at 10 M triangles.  Here the same decisions are
	// taken (a) without a pointer tree, (b) by several threads -- the two halves of a split are
	// independent, so the top levels fork; every subtree is built into its own array with
	// local numbering and the arrays are spliced in depth-first order, which reproduces the
	// serial numbering exactly -- and (c) optionally not at all: with a cache directory set, the
	// flattened arrays are stored under a hash of the mesh and simply read back next time.
	void Build(int maxPrims) {
		const int nTris = NumTriangles();
		const BuildOptions& opt = Options();
		uint64_t key = 0;
		if (!opt.cacheDir.empty()) {
			key = MeshKey(maxPrims);
			if (LoadCache(CachePath(opt.cacheDir, key), key, nTris)) return;
		}
		prims.resize(nTris);
		for (int t = 0; t < nTris; t++) {
			BuildPrim& p = prims[t];
			p.tri = t;
			for (int k = 0; k < 3; k++) p.box.Grow(&vertices[indices[3 * t + k].vertex_index].x);
			for (int a = 0; a < 3; a++) p.centroid[a] = (p.box.lo[a] + p.box.hi[a]) * 0.5f;
		}
		nodes.assign(2, BVHNode{});   // root + the unused slot 1
		if (nTris > 0) {
			int forkDepth = 0;
			for (int t = 1; t < opt.threads; t *= 2) forkDepth++;
			if (forkDepth > 0) forkDepth++;             // twice as many subtrees as threads evens out their sizes
			Top top;
			BuildTop(top, 0, nTris, maxPrims, forkDepth);
			Emit(top, 0);
		}
		leafOrder.resize(nTris);
		for (int j = 0; j < nTris; j++) leafOrder[j] = prims[j].tri;
		prims.clear();
		prims.shrink_to_fit();
		if (!opt.cacheDir.empty()) SaveCache(CachePath(opt.cacheDir, key), key);
	}

	static void StoreNode(BVHNode& n, const Box& b, int first, int count) {
		for (int a = 0; a < 3; a++) { n.bmin3[a] = b.lo[a]; n.bmax3[a] = b.hi[a]; }
		n.first = first;
		n.count = count;
	}

	// The decision of one BuildRecursive call for prims[start,end): returns false for a leaf,
	// else reorders the range and returns the split position.
	bool DecideSplit(int start, int end, int maxPrims, Box& bounds, int& mid) {
		bounds = Box();
		for (int i = start; i < end; i++) bounds.Grow(prims[i].box);
		const int n = end - start;
		if (n == 1) return false;

		Box centroidBox;
		for (int i = start; i < end; i++) centroidBox.Grow(prims[i].centroid);
		const int axis = centroidBox.LongestAxis();
		if (centroidBox.lo[axis] == centroidBox.hi[axis]) return false;

		mid = (start + end) / 2;
		if (n <= 2) {
			std::nth_element(prims.begin() + start, prims.begin() + mid, prims.begin() + end,
				[axis](const BuildPrim& a, const BuildPrim& b) { return a.centroid[axis] < b.centroid[axis]; });
			return true;
		}
		int counts[kBuckets] = {};
		Box boxes[kBuckets];
		for (int i = start; i < end; i++) {
			int b = BucketOf(centroidBox, prims[i], axis);
			counts[b]++;
			boxes[b].Grow(prims[i].box);
		}
		// SAH cost of splitting after bucket i; first strictly smaller cost wins
		float minCost = 0;
		int minBucket = 0;
		for (int i = 0; i < kBuckets - 1; i++) {
			Box b0, b1;
			int c0 = 0, c1 = 0;
			for (int j = 0; j <= i; j++) { b0.Grow(boxes[j]); c0 += counts[j]; }
			for (int j = i + 1; j < kBuckets; j++) { b1.Grow(boxes[j]); c1 += counts[j]; }
			float cost = 1 + (c0 * b0.SurfaceArea() + c1 * b1.SurfaceArea()) / bounds.SurfaceArea();
			if (i == 0 || cost < minCost) { minCost = cost; minBucket = i; }
		}
		float leafCost = (float)n;
		if (!(n > maxPrims || minCost < leafCost)) return false;
		BuildPrim* pmid = std::partition(&prims[start], &prims[end - 1] + 1,
			[&](const BuildPrim& p) { return BucketOf(centroidBox, p, axis) <= minBucket; });
		mid = (int)(pmid - &prims[0]);
		return true;
	}

	// Subtree of prims[start,end) into `out`, numbered as if it were a whole tree: its root in
	// out[slot], children pairs appended at even indices, left subtree first.
	void BuildLocal(std::vector<BVHNode>& out, int slot, int start, int end, int maxPrims) {
		Box bounds;
		int mid = 0;
		if (!DecideSplit(start, end, maxPrims, bounds, mid)) { StoreNode(out[slot], bounds, start, end - start); return; }
		int pair = (int)out.size();            // children go to the next free even index
		out.resize(out.size() + 2);
		StoreNode(out[slot], bounds, pair, 0);
		BuildLocal(out, pair, start, mid, maxPrims);
		BuildLocal(out, pair + 1, mid, end, maxPrims);
	}

	// Top of the tree while forking: a split whose halves are built concurrently, or a task
	// (a whole subtree in `local`, layout [root, unused, pairs...]).
	struct Top {
		Box bounds;
		std::unique_ptr<Top> left, right;
		std::vector<BVHNode> local;
	};
	static constexpr int kForkGrain = 4096;    // ranges below this are not worth a thread

	void BuildTop(Top& t, int start, int end, int maxPrims, int forkDepth) {
		int mid = 0;
		if (forkDepth == 0 || end - start <= kForkGrain || !DecideSplit(start, end, maxPrims, t.bounds, mid)) {
			// (when DecideSplit says "leaf" it has not touched the range: BuildLocal decides the same again)
			t.local.assign(2, BVHNode{});
			BuildLocal(t.local, 0, start, end, maxPrims);
			return;
		}
		t.left.reset(new Top());
		t.right.reset(new Top());
		std::thread other([&] { BuildTop(*t.left, start, mid, maxPrims, forkDepth - 1); });
		BuildTop(*t.right, mid, end, maxPrims, forkDepth - 1);
		other.join();
	}

	// Splice in depth-first order: exactly the indices a serial build hands out.
	void Emit(const Top& t, int slot) {
		if (!t.left) {
			const int shift = (int)nodes.size() - 2;       // local pair index 2 lands at nodes.size()
			nodes[slot] = t.local[0];
			if (nodes[slot].count == 0) nodes[slot].first += shift;
			const size_t base = nodes.size();
			nodes.insert(nodes.end(), t.local.begin() + 2, t.local.end());
			for (size_t i = base; i < nodes.size(); i++) if (nodes[i].count == 0) nodes[i].first += shift;
			return;
		}
		int pair = (int)nodes.size();
		nodes.resize(nodes.size() + 2);
		StoreNode(nodes[slot], t.bounds, pair, 0);
		Emit(*t.left, pair);
		Emit(*t.right, pair + 1);
	}

	// ---- cache of the flattened arrays ---------------------------------------------------------
	struct CacheHeader {
		char magic[8];          // "AGPTBVH1"
		uint64_t key;           // hash of vertex positions, vertex indices and maxPrimsInNode
		int32_t nTris, nNodes;
	};
	uint64_t MeshKey(int maxPrims) const {
		uint64_t h = 1469598103934665603ull;
		auto mix = [&h](const void* p, size_t n) {
			// FNV-1a over 8-byte words (+ tail bytes)
			const unsigned char* b = (const unsigned char*)p;
			size_t i = 0;
			for (; i + 8 <= n; i += 8) { uint64_t w; memcpy(&w, b + i, 8); h = (h ^ w) * 1099511628211ull; }
			for (; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
		};
		for (const auto& v : vertices) { float xyz[3] = { v.x, v.y, v.z }; mix(xyz, sizeof(xyz)); }      // (not the float3's padding lane)
		for (const auto& ix : indices) { int v = ix.vertex_index; mix(&v, sizeof(v)); }
		mix(&maxPrims, sizeof(maxPrims));
		return h;
	}
	static std::string CachePath(const std::string& dir, uint64_t key) {
		char name[40];
		snprintf(name, sizeof(name), "/%016llx.agbvh", (unsigned long long)key);
		return dir + name;
	}
	bool LoadCache(const std::string& path, uint64_t key, int nTris) {
		FILE* f = fopen(path.c_str(), "rb");
		if (!f) return false;
		CacheHeader h;
		bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, "AGPTBVH1", 8) == 0 && h.key == key && h.nTris == nTris && h.nNodes >= 2;
		if (ok) {
			nodes.resize(h.nNodes);
			leafOrder.resize(nTris);
			ok = fread(nodes.data(), sizeof(BVHNode), nodes.size(), f) == nodes.size() &&
				fread(leafOrder.data(), sizeof(int32_t), leafOrder.size(), f) == leafOrder.size() && fgetc(f) == EOF;
		}
		fclose(f);
		if (!ok) { nodes.clear(); leafOrder.clear(); }      // stale or damaged: rebuild (and overwrite)
		return ok;
	}
	void SaveCache(const std::string& path, uint64_t key) const {
		std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
		FILE* f = fopen(tmp.c_str(), "wb");
		if (!f) return;                                      // the cache is best effort
		CacheHeader h;
		memcpy(h.magic, "AGPTBVH1", 8);
		h.key = key; h.nTris = (int32_t)leafOrder.size(); h.nNodes = (int32_t)nodes.size();
		bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(nodes.data(), sizeof(BVHNode), nodes.size(), f) == nodes.size() &&
			fwrite(leafOrder.data(), sizeof(int32_t), leafOrder.size(), f) == leafOrder.size();
		ok = fclose(f) == 0 && ok;
		if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
	}

	std::vector<BuildPrim> prims;
	std::vector<BVHNode> nodes;
	std::vector<int32_t> leafOrder;
};
